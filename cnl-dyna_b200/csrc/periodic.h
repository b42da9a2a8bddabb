// Translation classes of the mesh components (symmetric-path assembly).
//
// CMUT arrays are copies of one membrane mesh on a lattice (cnld/abstract.py matrix_array /
// linear_array; cnld/mesh.py Mesh.from_abstract translates one membrane mesh to every membrane
// position), and the single-layer kernel depends on x - y only.  The block of Z that couples
// component a with component b is therefore a function of the offset between them: blocks with the
// same offset are equal, blocks with opposite offsets are transposes of each other.  The plan below
// lists which blocks are integrated (one per offset class) and which are copies.  Detection is
// purely geometric -- connected components with identical connectivity whose vertices differ by a
// translation to within a few ulps -- and falls back to "integrate everything" otherwise.
#pragma once
#include <vector>

namespace cnld {

struct WorkItem { unsigned t_begin, t_end, s_begin, s_end; };  // row triangles x column-triangle chunk
struct CopyItem { unsigned dst_r0, dst_c0, src_r0, src_c0, transposed; };

struct TranslationPlan {
  bool periodic = false;
  unsigned ncomp = 1;     // components (membranes)
  unsigned nvc = 0;       // vertices per component (periodic only)
  unsigned ntc = 0;       // triangles per component
  unsigned nclass = 0;    // integrated off-diagonal blocks
  std::vector<WorkItem> work;
  std::vector<CopyItem> copy;
};

// row_threads x col_chunk is the CTA tile of the regular-pair kernel.
void build_translation_plan(unsigned nv, unsigned nt, const double *x, const unsigned *tri,
                            unsigned row_threads, unsigned col_chunk, bool allow_periodic,
                            TranslationPlan &plan);

}  // namespace cnld
