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
  uint capped_leaves = 0;   // low-rank leaves of the last assembly whose ACA stopped at the rank capacity, eps_aca not met
  HLeafDev *d_leaf = nullptr;
  uint *d_near = nullptr, *d_far = nullptr, *d_perm = nullptr, *d_pos = nullptr, *d_cpos = nullptr;
  uint *d_ctri_ptr = nullptr, *d_ctri = nullptr, *d_v2t_ptr = nullptr, *d_v2t = nullptr;
  cplx *d_pool = nullptr;
  unsigned char *d_flags = nullptr;
  uint *d_err = nullptr;
  cplx *d_xp = nullptr, *d_yp = nullptr;  // permuted work vectors
  uint vec_cap = 0;
  // batched matvec (nrhs >= 4): bands of 16 rows in cluster order, each with the list of leaves that
  // intersect it; low-rank leaves first contract t = V^T x into a scratch (k x nrhs per leaf)
  uint nband = 0, nmvfar = 0;
  uint *d_band_ptr = nullptr, *d_mvfar = nullptr;   // band -> dense rounds; low-rank leaf ids
  // dense round = up to 16 columns of a dense leaf that intersects the band
  struct DRound { unsigned long long a_off; uint r0, m, x_off, cnt; uint pad[2]; };
  DRound *d_dround = nullptr;
  // band -> packed (low-rank leaf, rank index) items, rebuilt after every assembly (ranks change)
  struct FarItem { unsigned long long u_off, t_idx; uint r0, m; };   // U column at pool[u_off + row - r0]; t row index
  uint *d_bfar_ptr = nullptr;
  FarItem *d_bfar = nullptr;
  size_t bfar_cap = 0;
  unsigned long long *d_toff = nullptr;   // per leaf: offset (in rank indices) into the scratch
  unsigned long long ksum = 0;            // rows of the scratch (rank indices, padded per column cluster to 16)
  // stacked t-pass: the low-rank leaves of one column cluster contract the same x|cols, so their rank
  // indices are laid out consecutively in the scratch (padded to a multiple of 16 per column cluster) and
  // a CTA owns a band of 16 scratch rows: t|band = [V_1 .. V_q]^T|band x|cols -- full 16-row DMMA tiles
  // instead of one CTA per leaf with k (3..10) useful rows
  struct TBand { uint c0, n, rows, pad; };
  TBand *d_tband = nullptr;
  unsigned long long *d_vrow = nullptr;   // per scratch row: pool offset of its V column (~0: padding row)
  size_t tband_cap = 0, vrow_cap = 0;
  uint ntband = 0;
  std::vector<uint> far_by_col;           // low-rank leaf ids sorted by column cluster (structure, built once)
  cplx *d_T = nullptr;
  size_t T_cap = 0;
  bool mv_ready = false;
  // translation classes: leaf `dup[2i]` shares the pool data (and, after assembly, the rank) of leaf `dup[2i+1]`
  uint ndup = 0;
  uint *d_dup = nullptr;
  size_t unique_entries = 0;   // pool entries actually stored / assembled
};

struct LeafRange { uint r0, m, c0, n; bool admissible; };
// hx (optional): host vertex coordinates; with them, and a mesh that is translated copies of one
// component (bem.periodic), leaves that are translates of an earlier leaf share its pool data
int hmat_create_from_ranges(HMat &H, const BemDev &bem, const uint *htri, const std::vector<uint> &perm,
                            const std::vector<LeafRange> &ranges, uint kmax, const double *hx = nullptr);
int hmat_from_host(HMat &H, int device, uint n, const std::vector<uint> &perm, const std::vector<HLeafDev> &leaf,
                   const std::vector<cplx> &pool);
int hmat_create(HMat &H, const BemDev &bem, const double *hx, const uint *htri, uint clf, double eta, int admis,
                bool strict, uint kmax);
void hmat_free(HMat &H);
int hmat_assemble(cudaStream_t st, HMat &H, double kr, double ki, double accur);
int hmat_matvec(cudaStream_t st, HMat &H, cplx alpha, const cplx *d_x, cplx *d_y, uint nrhs);
int hmat_expand(cudaStream_t st, HMat &H, cplx *d_Z, size_t ld);
int hmat_compress_dense(cudaStream_t st, HMat &H, const cplx *d_Z, size_t ld, double accur);
int hmat_fetch_leaves(HMat &H);
// (re)build the band lists and rank offsets of the batched matvec from H.leaf (after the ranks changed)
int hmat_prepare_batched(HMat &H, bool fetch);
int nearfield_block(cudaStream_t st, const BemDev &bem, const uint *htri, const uint *ridx, uint rows,
                    const uint *cidx, uint cols, double kr, double ki, cplx *d_out);

}  // namespace cnld
