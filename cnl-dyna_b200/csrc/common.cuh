// Shared declarations for libcnld_b200 (sm_100a).  Complex numbers are double2
// (re, im), layout-identical to C99 double _Complex / H2Lib `field`.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

typedef double2 cplx;
typedef unsigned int uint;

namespace cnld {

void set_error(const char *fmt, ...);
extern std::atomic<unsigned long long> g_launches;
inline void count_launch(unsigned n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
// executed real flops of every ZGEMM launched so far (4M: 8 m n k, 3M: 6 m n k, skipped tiles excluded)
extern std::atomic<unsigned long long> g_dmma_flops;

#define CNLD_CK(call)                                                                     \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      cnld::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return -1;                                                                          \
    }                                                                                     \
  } while (0)

#define CNLD_CK_LAUNCH()                                                                  \
  do {                                                                                    \
    cnld::count_launch();                                                                 \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) {                                                             \
      cnld::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return -1;                                                                          \
    }                                                                                     \
  } while (0)

__host__ __device__ inline cplx cmul(cplx a, cplx b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ inline cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ inline cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ inline cplx cinv(cplx a) {
  double d = 1.0 / (a.x * a.x + a.y * a.y);
  return make_double2(a.x * d, -a.y * d);
}
// c += a*b
__device__ __forceinline__ void cfma(cplx &c, cplx a, cplx b) {
  c.x = fma(a.x, b.x, c.x);
  c.x = fma(-a.y, b.y, c.x);
  c.y = fma(a.x, b.y, c.y);
  c.y = fma(a.y, b.x, c.y);
}
// c -= a*b
__device__ __forceinline__ void cfms(cplx &c, cplx a, cplx b) {
  c.x = fma(-a.x, b.x, c.x);
  c.x = fma(a.y, b.y, c.x);
  c.y = fma(-a.x, b.y, c.y);
  c.y = fma(-a.y, b.x, c.y);
}

extern int g_zgemm_variant;
extern int g_lu_outer;
extern int g_lu_reserve_sms;
extern int g_lu_outer_inverse;  // lu_solve through inverted outer diagonal blocks
// ---- dense kernels (zgemm.cu / lu.cu) --------------------------------------
// C(MxN) = beta*C + alpha*A(MxK)*B(KxN); alpha in {+1,-1}, beta in {0,1};
// column-major, interleaved complex.  In-place (C aliasing A or B) is safe only
// when one CTA tile spans the full extent that is re-read (see lu.cu).
int zgemm(cudaStream_t st, uint M, uint N, uint K, double alpha, const cplx *A, size_t lda,
          const cplx *B, size_t ldb, double beta, cplx *C, size_t ldc);

// same contract, 3M complex product (C must not alias A or B)
// lower_only: skip the tiles that lie entirely above the diagonal of C (M == N); *entries (optional)
// returns the number of output entries actually computed.
int zgemm3m(cudaStream_t st, uint M, uint N, uint K, double alpha, const cplx *A, size_t lda,
            const cplx *B, size_t ldb, double beta, cplx *C, size_t ldc, int reserve_sms = 0,
            bool lower_only = false, double *entries = nullptr);
extern int g_zgemm_3m;

// Two-level batch of a ZGEMM: item q = o * per + j (blockIdx.z) shifts the operand pointers by
// o * s?o + j * s?j elements.
struct ZBatch {
  uint per = 1;
  long long sAo = 0, sAj = 0, sBo = 0, sBj = 0, sCo = 0, sCj = 0;
};
int zgemm_batched(cudaStream_t st, uint M, uint N, uint K, double alpha, const cplx *A, size_t lda,
                  const cplx *B, size_t ldb, double beta, cplx *C, size_t ldc, uint nbatch, const ZBatch &bt);

int zgemm3m_batched(cudaStream_t st, uint M, uint N, uint K, double alpha, const cplx *A, size_t lda,
                    const cplx *B, size_t ldb, double beta, cplx *C, size_t ldc, uint nbatch, const ZBatch &bt);

struct LuWork {
  uint n = 0, nb = 0, nblk = 0, nouter = 0, nbo = 0;
  // inverses of the full nbo x nbo diagonal blocks of L and R (lu_invert_outer), used by lu_solve
  cplx *linvO = nullptr, *uinvO = nullptr, *tmpO = nullptr;
  uint inv_nbo = 0, inv_nfull = 0;   // geometry the buffers were allocated / filled for
  bool inv_ready = false;
  mutable cplx *ytmp = nullptr;      // nbo x nrhs scratch of lu_solve
  mutable size_t ytmp_elems = 0;
  cplx *linv = nullptr;  // [nblk][nb*nb]
  cplx *uinv = nullptr;  // [nblk][nb*nb]
  uint *info = nullptr;  // device flag: 0 ok else first failing pivot + 1
  cplx *panel = nullptr; // scratch for out-of-place panel updates
  cudaStream_t lo = nullptr;  // optional low-priority stream for the big trailing updates
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool prof = false;     // record events around every trailing-update ZGEMM
  std::vector<cudaEvent_t> pe;
  std::vector<double> pflops;  // executed flops of the launch bracketed by pe[2i], pe[2i+1]
};
int lu_work_alloc(LuWork &w, uint n);
void lu_work_free(LuWork &w);
// form inv(L_kk), inv(R_kk) of an already factored matrix (stand-alone lrsolve entry)
int lu_invert_diag(cudaStream_t st, LuWork &w, cplx *a, size_t lda);
// inverses of the full outer diagonal blocks (recursive doubling from the nb-block inverses)
int lu_invert_outer(cudaStream_t st, LuWork &w, const cplx *a, size_t lda);
void lu_profile(LuWork &w, bool on);
int lu_trailing_stats(const LuWork &w, double *seconds, double *flops, double *launches);
// unpivoted LU in place on device matrix a (n x n, lda); keeps block inverses in w
int lu_factor(cudaStream_t st, LuWork &w, cplx *a, size_t lda);
// same result for a complex SYMMETRIC matrix given by its lower triangle, half the flops
int lu_factor_sym(cudaStream_t st, LuWork &w, cplx *a, size_t lda);
// X (n x nrhs, ldx) <- (LU)^-1 X
int lu_solve(cudaStream_t st, const LuWork &w, const cplx *a, size_t lda, cplx *x, uint nrhs,
             size_t ldx, const uint *col_first_row = nullptr);

}  // namespace cnld
