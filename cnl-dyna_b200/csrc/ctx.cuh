// Private definition of cnld_ctx (shared by capi.cu and hapi.cu).
#pragma once
#include "../../include/cnld_b200.h"
#include "bem.cuh"

namespace cnld {
namespace sweep {
constexpr int MAX_SLOTS = 8;
constexpr double TWO_PI = 6.283185307179586476925286766559;

struct Slot {
  cudaStream_t st = nullptr;      // high priority: latency-bound kernels
  cudaStream_t st_lo = nullptr;   // medium priority: trailing updates (DMMA tensor pipe)
  cudaStream_t st_asm = nullptr;  // lowest priority: assembly.  (Measured on B200: DMMA and DFMA share the
                                  // FP64 units -- a co-resident assembly CTA slows the ZGEMM CTAs one for one --
                                  // so the assembly is NOT hidden under another frequency's factorisation.)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cplx *G = nullptr;      // nv x nv
  cplx *X = nullptr;      // nv x P
  cplx *hX = nullptr;     // pinned staging for x_node
  // symmetric path: sparse Sauter-Schwab slots, their asymmetry, refinement vectors (nv x P each)
  cplx *sval = nullptr, *dval = nullptr, *E = nullptr, *E2 = nullptr;
  double *norms = nullptr;  // [sum |e_last|^2, sum |x|^2] of the last refinement step
  bool was_sym = false;
  LuWork lu;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  long pending = -1;      // frequency index whose x_node sits in hX
};
}  // namespace sweep
}  // namespace cnld

using namespace cnld;
using namespace cnld::sweep;

struct cnld_ctx {
  BemDev bem;
  int nslots = 2;
  Slot slots[MAX_SLOTS];
  bool ws_ready = false;
  // FEM operators, one CSR pattern
  size_t nnz = 0;
  int64_t *d_indptr = nullptr;
  uint *d_indices = nullptr;
  double *d_K = nullptr, *d_M = nullptr, *d_B = nullptr;
  // host copies (block detection for the GMRES preconditioner, hapi.cu)
  std::vector<int64_t> h_indptr;
  std::vector<uint> h_indices;
  std::vector<double> h_K, h_M, h_B;
  std::vector<double> h_x;   // mesh kept for the cluster tree
  std::vector<uint> h_t;
  // loads
  uint P = 0;
  int64_t *d_f_indptr = nullptr, *d_a_indptr = nullptr;
  uint *d_f_idx = nullptr, *d_a_idx = nullptr;
  double *d_f_val = nullptr, *d_a_val = nullptr;
  uint8_t *d_ob = nullptr;
  std::vector<uint> f_first_row;   // per source patch: first loaded row (empty: columns not ordered by row)
  // results of the last sweep
  cplx *d_xpatch = nullptr;
  size_t xpatch_cap = 0;
  cplx *d_xnode = nullptr;  // only for run_device with keep_node
  double stage[5] = {0, 0, 0, 0, 0};
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  bool profile = false;
  // symmetric path (cnld_b200_set_symmetric): factor the symmetric part, refine with the sparse asymmetry
  int symmetric = 0, refine_steps = 1;
  double sym_tol = 1e-11;       // estimated remaining relative error above which a frequency is redone (general path)
  unsigned sym_redone = 0;      // frequencies redone in the last sweep
  double sym_max_ratio = 0.0;   // largest |e_last| / |x| of the last sweep
  // strictly upper asymmetry of the FEM operators, CSR by row: (K, M, B)[r,c] - (K, M, B)[c,r]
  uint m_nnz = 0;
  uint *d_m_rowptr = nullptr, *d_m_col = nullptr;
  double *d_m_dK = nullptr, *d_m_dM = nullptr, *d_m_dB = nullptr;
  double kstat[3] = {0, 0, 0};  // trailing ZGEMM: flops, seconds, launches (last sweep)
  // caller's x_node buffer page-locked in place (direct D2H, no staging copy); re-used across calls
  void *reg_ptr = nullptr;
  size_t reg_bytes = 0;
  // pressure scratch
  double *d_sp = nullptr;
  PresPlan pplan;            // triangle chunks of the many-source pressure path (pfield.cu)
  bool pplan_ready = false;
};

