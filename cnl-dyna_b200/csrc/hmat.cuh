// Device H-matrix (see hmatrix.cu).
#pragma once
#include "bem.cuh"
#include "htree.h"

namespace cnld {

struct HLeafDev {
  uint r0, m, c0, n;  // row / column range in cluster order
  uint adm;           // 1: low rank (U V^T), 0: dense
  uint k, kcap;       // rank found by ACA / capacity
  uint rtl, ctl;      // triangle-list ids of the row / column cluster (dense leaves)
  uint pad;
  size_t off;         // offset into the pool (complex doubles)
  size_t foff;        // offset into the pivot-flag scratch (low-rank leaves)
};

struct HMat {
  const BemDev *bem = nullptr;
  uint n = 0, kmax = 0, nnear = 0, nfar = 0;
  ClusterTree tree;
  std::vector<BlockLeaf> blocks;
  std::vector<HLeafDev> leaf;
  size_t pool_size = 0, near_entries = 0;
  bool assembled = false;
  HLeafDev *d_leaf = nullptr;
  uint *d_near = nullptr, *d_far = nullptr, *d_perm = nullptr, *d_pos = nullptr, *d_cpos = nullptr;
  uint *d_ctri_ptr = nullptr, *d_ctri = nullptr, *d_v2t_ptr = nullptr, *d_v2t = nullptr;
  cplx *d_pool = nullptr;
  unsigned char *d_flags = nullptr;
  uint *d_err = nullptr;
  cplx *d_xp = nullptr, *d_yp = nullptr;  // permuted work vectors
  uint vec_cap = 0;
};

struct LeafRange { uint r0, m, c0, n; bool admissible; };
int hmat_create_from_ranges(HMat &H, const BemDev &bem, const uint *htri, const std::vector<uint> &perm,
                            const std::vector<LeafRange> &ranges, uint kmax);
int hmat_from_host(HMat &H, int device, uint n, const std::vector<uint> &perm, const std::vector<HLeafDev> &leaf,
                   const std::vector<cplx> &pool);
int hmat_create(HMat &H, const BemDev &bem, const double *hx, const uint *htri, uint clf, double eta, int admis,
                bool strict, uint kmax);
void hmat_free(HMat &H);
int hmat_assemble(cudaStream_t st, HMat &H, double kr, double ki, double accur);
int hmat_matvec(cudaStream_t st, HMat &H, cplx alpha, const cplx *d_x, cplx *d_y, uint nrhs);
int hmat_expand(cudaStream_t st, HMat &H, cplx *d_Z, size_t ld);
int hmat_fetch_leaves(HMat &H);
int nearfield_block(cudaStream_t st, const BemDev &bem, const uint *htri, const uint *ridx, uint rows,
                    const uint *cidx, uint cols, double kr, double ki, cplx *d_out);

}  // namespace cnld
