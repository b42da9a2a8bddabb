// Blocked UNPIVOTED complex-FP64 LU and multi-RHS triangular solves.
// Replaces lrdecomp_amatrix / lrsolve_amatrix_avector
// (include/factorizations.h:173-182, :233; call sites
// cnld/compressed_formats.py:247-252, :259-262).  Like the reference there is
// no pivoting: A = L R with unit-lower L, in place; a zero / non-finite pivot is
// reported through `info` (index + 1) exactly as lrdecomp_amatrix's return value.
//
// Right-looking, block size 64.  The diagonal block is factored by one CTA in
// shared memory, which also forms inv(L_kk) and inv(R_kk); both panel updates
// and all triangular solves then become DMMA ZGEMMs (zgemm.cu), as does the
// rank-64 trailing update that carries ~99% of the flops.
#include "common.cuh"

namespace cnld {
namespace {
constexpr int NB = 64;
constexpr int LD = NB + 1;  // odd leading dimension (in 16-byte elements): conflict free both ways
constexpr int DIAG_THREADS = 512;
constexpr int DIAG_SMEM = 2 * NB * LD * (int)sizeof(cplx);

// FACTOR=true : factor diagonal block `b0` (grid = 1).
// FACTOR=false: the matrix already holds L\R; only form the block inverses, one CTA per
//               diagonal block (grid = nblk) -- used by the stand-alone lrsolve entry.
// SYM=true: only the lower triangle of the block is valid on entry (symmetric path); the upper
//           triangle is mirrored while loading.
template <bool FACTOR, bool SYM = false>
__global__ void __launch_bounds__(DIAG_THREADS, 1)
k_diag(cplx *a0, size_t lda, uint n, uint b0, cplx *linv0, cplx *uinv0, uint *info) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ cplx P[NB];   // reciprocal pivots
  cplx *S = reinterpret_cast<cplx *>(smem_raw);
  cplx *X = S + NB * LD;
  const int tid = threadIdx.x;
  const uint r = tid & (NB - 1), g = tid / NB;   // thread = (row r, column group g of NG)
  constexpr uint NG = DIAG_THREADS / NB;
  const uint b = b0 + blockIdx.x, base = b * NB;
  const uint nbk = (n - base < (uint)NB) ? n - base : NB;
  cplx *a = a0 + (size_t)base * lda + base;
  cplx *linv = linv0 + (size_t)b * NB * NB, *uinv = uinv0 + (size_t)b * NB * NB;
  for (uint j = g; j < nbk; j += NG)
    if (r < nbk) S[j * LD + r] = (SYM && r < j) ? a[(size_t)r * lda + j] : a[(size_t)j * lda + r];
  __syncthreads();
  // One barrier per elimination step: column k of L stays unscaled in S until the end (every thread
  // multiplies by the reciprocal pivot itself -- the same two roundings as scaling first), and the
  // thread that produces the next pivot also publishes its reciprocal.
  auto publish = [&](uint k, cplx piv) {
    double mag = piv.x * piv.x + piv.y * piv.y;
    if (FACTOR && !(mag > 0.0 && mag < 1.79e308)) atomicCAS(info, 0u, base + k + 1);
    P[k] = cinv(piv);
  };
  if (FACTOR) {
    if (tid == 0) publish(0, S[0]);
    __syncthreads();
    for (uint k = 0; k + 1 < nbk; k++) {
      if (r > k && r < nbk) {
        const cplx l = cmul(S[k * LD + r], P[k]);
        for (uint j = k + 1 + g; j < nbk; j += NG) {
          cplx v = S[j * LD + r];
          cfms(v, l, S[j * LD + k]);
          S[j * LD + r] = v;
          if (r == k + 1 && j == k + 1) publish(k + 1, v);
        }
      }
      __syncthreads();
    }
    for (uint j = g; j < nbk; j += NG)   // scale L's columns, write the block back
      if (r < nbk) {
        cplx v = S[j * LD + r];
        if (r > j) { v = cmul(v, P[j]); S[j * LD + r] = v; }
        a[(size_t)j * lda + r] = v;
      }
  } else {
    if (tid < (int)nbk) publish(tid, S[tid * LD + tid]);
  }
  // X = inv(L), L unit lower: forward elimination on the identity
  for (uint j = g; j < NB; j += NG) X[j * LD + r] = make_double2(r == j ? 1.0 : 0.0, 0.0);
  __syncthreads();
  for (uint k = 0; k + 1 < nbk; k++) {
    if (r > k && r < nbk) {
      const cplx l = S[k * LD + r];
      for (uint c = g; c <= k; c += NG) cfms(X[c * LD + r], l, X[c * LD + k]);
    }
    __syncthreads();
  }
  for (uint j = g; j < NB; j += NG) linv[j * NB + r] = X[j * LD + r];
  if (SYM) {
    // symmetric block: R = D L^T, so inv(R) = inv(L)^T D^-1 -- a scaled transpose instead of a second elimination
    for (uint j = g; j < NB; j += NG) {
      cplx v = make_double2(r == j ? 1.0 : 0.0, 0.0);
      if (r < nbk && j < nbk) v = cmul(X[r * LD + j], P[j]);
      uinv[j * NB + r] = v;
    }
    return;
  }
  __syncthreads();
  // X = inv(R), R upper: backward elimination on the identity; row kk is scaled by the reciprocal
  // pivot on the fly and in place only at the end (one barrier per step again)
  for (uint j = g; j < NB; j += NG) X[j * LD + r] = make_double2(r == j ? 1.0 : 0.0, 0.0);
  __syncthreads();
  for (uint kk = nbk; kk-- > 1;) {
    if (r < kk) {
      const cplx u = S[kk * LD + r], pk = P[kk];
      for (uint c = kk + g; c < nbk; c += NG) cfms(X[c * LD + r], u, cmul(X[c * LD + kk], pk));
    }
    __syncthreads();
  }
  for (uint j = g; j < NB; j += NG) {
    cplx v = X[j * LD + r];
    if (r < nbk && j < nbk) v = cmul(v, P[r]);
    uinv[j * NB + r] = v;
  }
}
}  // namespace

int lu_work_alloc(LuWork &w, uint n) {
  w.n = n;
  w.nb = NB;
  w.nblk = (n + NB - 1) / NB;
  CNLD_CK(cudaMalloc(&w.linv, sizeof(cplx) * (size_t)w.nblk * NB * NB));
  CNLD_CK(cudaMalloc(&w.uinv, sizeof(cplx) * (size_t)w.nblk * NB * NB));
  CNLD_CK(cudaMalloc(&w.info, sizeof(uint)));
  w.pflops.assign(w.nblk + 1, 0.0);
  return 0;
}

void lu_work_free(LuWork &w) {
  cudaFree(w.linvO); cudaFree(w.uinvO); cudaFree(w.tmpO); cudaFree(w.ytmp);
  cudaFree(w.linv);
  cudaFree(w.uinv);
  cudaFree(w.info);
  cudaFree(w.panel);
  for (auto &e : w.pe) cudaEventDestroy(e);
  if (w.ev_fork) cudaEventDestroy(w.ev_fork);
  if (w.ev_join) cudaEventDestroy(w.ev_join);
  w = LuWork();
}

static int diag_attr() {
  static bool attr_set[64] = {false};
  int dev = 0;
  CNLD_CK(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    CNLD_CK(cudaFuncSetAttribute(k_diag<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM));
    CNLD_CK((cudaFuncSetAttribute(k_diag<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM)));
    CNLD_CK(cudaFuncSetAttribute(k_diag<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM));
    attr_set[dev & 63] = true;
  }
  return 0;
}

int lu_invert_diag(cudaStream_t st, LuWork &w, cplx *a, size_t lda) {
  if (diag_attr()) return -1;
  k_diag<false><<<w.nblk, DIAG_THREADS, DIAG_SMEM, st>>>(a, lda, w.n, 0, w.linv, w.uinv, w.info);
  CNLD_CK_LAUNCH();
  return 0;
}

void lu_profile(LuWork &w, bool on) {
  if (on && w.pe.empty()) {
    w.pe.resize(2 * (size_t)w.nblk);
    for (auto &e : w.pe) cudaEventCreate(&e);
  }
  w.prof = on;
}

// seconds and flops spent in the trailing-update ZGEMMs of the last lu_factor on this work
int lu_trailing_stats(const LuWork &w, double *seconds, double *flops, double *launches) {
  *seconds = *flops = *launches = 0.0;
  if (!w.prof) return 0;
  for (uint ob = 0; ob < w.nouter && ob < w.pflops.size(); ob++) {
    if (w.pflops[ob] <= 0.0) continue;
    float ms = 0.f;
    CNLD_CK(cudaEventElapsedTime(&ms, w.pe[2 * ob], w.pe[2 * ob + 1]));
    *seconds += ms * 1e-3;
    *flops += w.pflops[ob] * (g_zgemm_3m ? 1.0 : 4.0 / 3.0);  // executed real flops (3M: 6 m n k)
    *launches += 1.0;
  }
  return 0;
}

// Two-level right-looking factorisation.  Outer blocks of NBO columns: inside an outer block
// the 64-wide steps update only the block column (all rows below) and the block row (all columns
// to the right); the rest of the matrix receives ONE rank-NBO update per outer block, so ~95% of
// the flops run as K = NBO ZGEMMs (DMMA efficiency grows with K: 26.6 / 32.4 / 33.7 TFLOP/s at
// K = 64 / 256 / 512 on B200).
//
// Look-ahead (when a companion stream w.lo exists): the rank-NBO update of outer block O is split
// into T1 = the strip that the NEXT outer block's panel work touches (its block column and block
// row) and T2 = everything beyond.  T2(O) runs on the companion stream with a few SMs left free
// while the latency-bound panel loop of block O+1 runs on the main stream:
//     main : inner(O) -> [wait T2(O-1)] T1(O) -> inner(O+1) -> [wait T2(O)] T1(O+1) -> ...
//     lo   :                      [wait T1(O)] T2(O)   ->            [wait T1(O+1)] T2(O+1) ...
int g_lu_outer = 512;
int g_lu_outer_inverse = 1;
int g_lu_reserve_sms = 12;

static int inner_loop(cudaStream_t st, LuWork &w, cplx *a, size_t lda, uint O, uint OE) {
  const uint n = w.n;
  for (uint o = O; o < OE; o += NB) {
    const uint b = o / NB;
    const uint nbk = (OE - o < (uint)NB) ? OE - o : NB;
    const uint below = n - o - nbk;  // rows below / columns right of the diagonal block
    cplx *akk = a + (size_t)o * lda + o;
    cplx *li = w.linv + (size_t)b * NB * NB, *ui = w.uinv + (size_t)b * NB * NB;
    k_diag<true><<<1, DIAG_THREADS, DIAG_SMEM, st>>>(a, lda, n, b, w.linv, w.uinv, w.info);
    CNLD_CK_LAUNCH();
    if (!below) continue;
    cplx *a21 = akk + nbk, *a12 = akk + (size_t)nbk * lda, *a22 = a12 + nbk;
    // in place: one CTA tile (64x64) spans the whole 64-wide panel it re-reads
    if (zgemm(st, below, nbk, nbk, 1.0, a21, lda, ui, NB, 0.0, a21, lda)) return -1;
    if (zgemm(st, nbk, below, nbk, 1.0, li, NB, a12, lda, 0.0, a12, lda)) return -1;
    const uint in_cols = OE - o - nbk;  // columns (and rows) left inside the outer block
    if (in_cols) {
      // block column: all rows below x remaining columns of the outer block
      if (zgemm3m(st, below, in_cols, nbk, -1.0, a21, lda, a12, lda, 1.0, a22, lda)) return -1;
      // block row: remaining rows of the outer block x columns right of the outer block
      const uint right = n - OE;
      if (right && zgemm3m(st, in_cols, right, nbk, -1.0, a21, lda, a12 + (size_t)in_cols * lda, lda, 1.0,
                           a22 + (size_t)in_cols * lda, lda)) return -1;
    }
  }
  return 0;
}

int lu_factor(cudaStream_t st, LuWork &w, cplx *a, size_t lda) {
  if (diag_attr()) return -1;
  const uint n = w.n;
  const uint NBO = (uint)((g_lu_outer / NB) * NB > 0 ? (g_lu_outer / NB) * NB : NB);
  CNLD_CK(cudaMemsetAsync(w.info, 0, sizeof(uint), st));
  w.inv_ready = false;
  const bool ahead = w.lo != nullptr;
  bool t2_pending = false;
  uint ob = 0;
  for (uint O = 0; O < n; O += NBO, ob++) {
    const uint W = (n - O < NBO) ? n - O : NBO;
    const uint OE = O + W;  // end of the outer block
    if (inner_loop(st, w, a, lda, O, OE)) return -1;
    const uint rem = n - OE;
    if (!rem) break;
    cplx *l = a + (size_t)O * lda + OE, *u = a + (size_t)OE * lda + O, *t = a + (size_t)OE * lda + OE;
    if (!ahead) {
      if (w.prof) CNLD_CK(cudaEventRecord(w.pe[2 * ob], st));
      if (zgemm3m(st, rem, rem, W, -1.0, l, lda, u, lda, 1.0, t, lda)) return -1;
      if (w.prof) CNLD_CK(cudaEventRecord(w.pe[2 * ob + 1], st));
      w.pflops[ob] = 6.0 * rem * rem * W;
      continue;
    }
    const uint W2 = (rem < NBO) ? rem : NBO;  // width of the next outer block
    // T1: needs the previous T2 (it covered this strip)
    if (t2_pending) CNLD_CK(cudaStreamWaitEvent(st, w.ev_join, 0));
    if (zgemm3m(st, rem, W2, W, -1.0, l, lda, u, lda, 1.0, t, lda)) return -1;
    const uint rest = rem - W2;
    if (rest) {
      if (zgemm3m(st, W2, rest, W, -1.0, l, lda, u + (size_t)W2 * lda, lda, 1.0, t + (size_t)W2 * lda, lda)) return -1;
      // T2 on the companion stream, leaving a few SMs to the panel loop of the next block
      CNLD_CK(cudaEventRecord(w.ev_fork, st));
      CNLD_CK(cudaStreamWaitEvent(w.lo, w.ev_fork, 0));
      if (w.prof) CNLD_CK(cudaEventRecord(w.pe[2 * ob], w.lo));
      if (zgemm3m(w.lo, rest, rest, W, -1.0, l + W2, lda, u + (size_t)W2 * lda, lda, 1.0,
                  t + (size_t)W2 * lda + W2, lda, g_lu_reserve_sms)) return -1;
      if (w.prof) CNLD_CK(cudaEventRecord(w.pe[2 * ob + 1], w.lo));
      CNLD_CK(cudaEventRecord(w.ev_join, w.lo));
      t2_pending = true;
      w.pflops[ob] = 6.0 * rest * rest * W;
    } else {
      t2_pending = false;
      w.pflops[ob] = 0.0;
    }
  }
  if (ahead && t2_pending) CNLD_CK(cudaStreamWaitEvent(st, w.ev_join, 0));
  w.nouter = ob;
  w.nbo = NBO;
  return g_lu_outer_inverse ? lu_invert_outer(st, w, a, lda) : 0;
}

// dst(r, c) = d(r) * src(c, r):  U block = D * L block^T  (G = L D L^T  =>  R = D L^T), 32x32 tiles
__global__ void __launch_bounds__(256)
k_ut(uint R, uint Cn, const cplx *__restrict__ src, const cplx *__restrict__ diag, cplx *dst, size_t lda) {
  __shared__ cplx tile[32][33];
  const uint r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const uint tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (uint k = ty; k < 32; k += 8) {
    uint c = c0 + tx, r = r0 + k;  // src(c, r): c contiguous
    if (c < Cn && r < R) tile[k][tx] = src[(size_t)r * lda + c];
  }
  __syncthreads();
  for (uint k = ty; k < 32; k += 8) {
    uint r = r0 + tx, c = c0 + k;  // dst(r, c): r contiguous
    if (r < R && c < Cn) dst[(size_t)c * lda + r] = cmul(diag[(size_t)r * (lda + 1)], tile[tx][k]);
  }
}

static int ut(cudaStream_t st, uint R, uint Cn, const cplx *src, const cplx *diag, cplx *dst, size_t lda) {
  if (!R || !Cn) return 0;
  dim3 grid((R + 31) / 32, (Cn + 31) / 32);
  k_ut<<<grid, 256, 0, st>>>(R, Cn, src, diag, dst, lda);
  CNLD_CK_LAUNCH();
  return 0;
}

// Symmetric variant: a holds the LOWER triangle of a complex symmetric matrix S.  The unpivoted
// elimination of S is S = L R with R = D L^T, so only L is computed by ZGEMMs; every block of R is
// obtained from L by a scaled transposition, the block-row products disappear and the trailing
// updates touch only the tiles on or below the diagonal: (4/3) n^3 conventional flops instead of
// (8/3) n^3.  On exit the buffer holds L (unit lower) and R (upper) exactly like lu_factor's, so
// lu_solve applies unchanged.  Look-ahead as in lu_factor.
int lu_factor_sym(cudaStream_t st, LuWork &w, cplx *a, size_t lda) {
  if (diag_attr()) return -1;
  const uint n = w.n;
  const uint NBO = (uint)((g_lu_outer / NB) * NB > 0 ? (g_lu_outer / NB) * NB : NB);
  CNLD_CK(cudaMemsetAsync(w.info, 0, sizeof(uint), st));
  w.inv_ready = false;
  const bool ahead = w.lo != nullptr;
  bool t2_pending = false;
  uint ob = 0;
  for (uint O = 0; O < n; O += NBO, ob++) {
    const uint W = (n - O < NBO) ? n - O : NBO;
    const uint OE = O + W;
    for (uint o = O; o < OE; o += NB) {
      const uint b = o / NB;
      const uint nbk = (OE - o < (uint)NB) ? OE - o : NB;
      const uint below = n - o - nbk;
      cplx *akk = a + (size_t)o * lda + o;
      k_diag<true, true><<<1, DIAG_THREADS, DIAG_SMEM, st>>>(a, lda, n, b, w.linv, w.uinv, w.info);
      CNLD_CK_LAUNCH();
      if (!below) continue;
      cplx *a21 = akk + nbk, *a12 = akk + (size_t)nbk * lda, *a22 = a12 + nbk;
      if (zgemm(st, below, nbk, nbk, 1.0, a21, lda, w.uinv + (size_t)b * NB * NB, NB, 0.0, a21, lda)) return -1;
      const uint in_cols = OE - o - nbk;
      if (in_cols) {
        if (ut(st, nbk, in_cols, a21, akk, a12, lda)) return -1;
        if (zgemm3m(st, below, in_cols, nbk, -1.0, a21, lda, a12, lda, 1.0, a22, lda)) return -1;
      }
    }
    const uint rem = n - OE;
    if (!rem) break;
    cplx *l = a + (size_t)O * lda + OE, *u = a + (size_t)OE * lda + O, *t = a + (size_t)OE * lda + OE;
    if (ut(st, W, rem, l, a + (size_t)O * lda + O, u, lda)) return -1;
    const uint W2 = (rem < NBO) ? rem : NBO;
    if (t2_pending) CNLD_CK(cudaStreamWaitEvent(st, w.ev_join, 0));
    // T1: the block column of the next outer block (all rows below)
    if (zgemm3m(st, rem, W2, W, -1.0, l, lda, u, lda, 1.0, t, lda)) return -1;
    const uint rest = rem - W2;
    t2_pending = false;
    w.pflops[ob] = 0.0;
    if (rest) {
      cudaStream_t ts = st;
      if (ahead) {
        CNLD_CK(cudaEventRecord(w.ev_fork, st));
        CNLD_CK(cudaStreamWaitEvent(w.lo, w.ev_fork, 0));
        ts = w.lo;
      }
      if (w.prof) CNLD_CK(cudaEventRecord(w.pe[2 * ob], ts));
      double tiles = 0.0;
      if (zgemm3m(ts, rest, rest, W, -1.0, l + W2, lda, u + (size_t)W2 * lda, lda, 1.0, t + (size_t)W2 * lda + W2, lda,
                  ahead ? g_lu_reserve_sms : 0, true, &tiles)) return -1;
      if (w.prof) CNLD_CK(cudaEventRecord(w.pe[2 * ob + 1], ts));
      if (ahead) {
        CNLD_CK(cudaEventRecord(w.ev_join, w.lo));
        t2_pending = true;
      }
      w.pflops[ob] = 6.0 * tiles * W;  // tiles = computed output entries
    }
  }
  if (t2_pending) CNLD_CK(cudaStreamWaitEvent(st, w.ev_join, 0));
  w.nouter = ob;
  w.nbo = NBO;
  return g_lu_outer_inverse ? lu_invert_outer(st, w, a, lda) : 0;
}

// ---- inverses of the outer diagonal blocks -----------------------------------------------------
// zero the nbo x nbo inverse blocks and put the nb x nb block inverses (k_diag) on their diagonals
__global__ void k_inv_init(uint nfull, uint nbo, const cplx *__restrict__ linv, const cplx *__restrict__ uinv,
                           cplx *linvO, cplx *uinvO) {
  const size_t tot = (size_t)nfull * nbo * nbo;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < tot; idx += (size_t)gridDim.x * blockDim.x) {
    const uint O = (uint)(idx / ((size_t)nbo * nbo));
    const uint rem = (uint)(idx % ((size_t)nbo * nbo));
    const uint i = rem % nbo, j = rem / nbo;
    cplx l = make_double2(0.0, 0.0), u = l;
    if (i / NB == j / NB) {
      const size_t src = ((size_t)O * (nbo / NB) + i / NB) * NB * NB + (size_t)(j % NB) * NB + (i % NB);
      l = linv[src];
      u = uinv[src];
    }
    linvO[idx] = l;
    uinvO[idx] = u;
  }
}

// inv([A 0; C B]) = [A^-1 0; -B^-1 C A^-1  B^-1] (L) and inv([A C; 0 B]) = [A^-1  -A^-1 C B^-1; 0 B^-1] (R),
// doubling the block size from nb to nbo, all outer blocks and all pairs of a level in one launch.
int lu_invert_outer(cudaStream_t st, LuWork &w, const cplx *a, size_t lda) {
  const uint nbo = w.nbo, nfull = w.n / (nbo ? nbo : 1);
  w.inv_ready = false;
  if (!nbo || nbo == (uint)NB || !nfull || (nbo & (nbo - 1)) || nbo % NB) return 0;  // needs nbo = nb * 2^k
  if (w.inv_nbo != nbo || w.inv_nfull != nfull) {
    cudaFree(w.linvO); cudaFree(w.uinvO); cudaFree(w.tmpO);
    w.linvO = w.uinvO = w.tmpO = nullptr;
    const size_t blk = (size_t)nbo * nbo;
    CNLD_CK(cudaMalloc(&w.linvO, sizeof(cplx) * nfull * blk));
    CNLD_CK(cudaMalloc(&w.uinvO, sizeof(cplx) * nfull * blk));
    CNLD_CK(cudaMalloc(&w.tmpO, sizeof(cplx) * nfull * blk / 2));
    w.inv_nbo = nbo; w.inv_nfull = nfull;
  }
  k_inv_init<<<1184, 256, 0, st>>>(nfull, nbo, w.linv, w.uinv, w.linvO, w.uinvO);
  CNLD_CK_LAUNCH();
  const long long blk = (long long)nbo * nbo;
  for (uint s = NB; 2 * s <= nbo; s *= 2) {
    const uint per = nbo / (2 * s), nbatch = nfull * per;
    const long long dA = (long long)(lda + 1), dI = (long long)(nbo + 1);
    ZBatch bt;
    bt.per = per;
    // L:  T = C * Ainv  (C = a[r0 + s, r0])
    bt.sAo = nbo * dA; bt.sAj = 2 * s * dA;
    bt.sBo = blk;      bt.sBj = 2 * s * dI;
    bt.sCo = blk / 2;  bt.sCj = (long long)s * s;
    if (zgemm_batched(st, s, s, s, 1.0, a + s, lda, w.linvO, nbo, 0.0, w.tmpO, s, nbatch, bt)) return -1;
    //     new block = -Binv * T
    bt.sAo = blk;      bt.sAj = 2 * s * dI;
    bt.sBo = blk / 2;  bt.sBj = (long long)s * s;
    bt.sCo = blk;      bt.sCj = 2 * s * dI;
    if (zgemm_batched(st, s, s, s, -1.0, w.linvO + s * dI, nbo, w.tmpO, s, 0.0, w.linvO + s, nbo, nbatch, bt)) return -1;
    // R:  T = C * Binv  (C = a[r0, r0 + s])
    bt.sAo = nbo * dA; bt.sAj = 2 * s * dA;
    bt.sBo = blk;      bt.sBj = 2 * s * dI;
    bt.sCo = blk / 2;  bt.sCj = (long long)s * s;
    if (zgemm_batched(st, s, s, s, 1.0, a + (size_t)s * lda, lda, w.uinvO + s * dI, nbo, 0.0, w.tmpO, s, nbatch, bt)) return -1;
    //     new block = -Ainv * T
    bt.sAo = blk;      bt.sAj = 2 * s * dI;
    bt.sBo = blk / 2;  bt.sBj = (long long)s * s;
    bt.sCo = blk;      bt.sCj = 2 * s * dI;
    if (zgemm_batched(st, s, s, s, -1.0, w.uinvO, nbo, w.tmpO, s, 0.0, w.uinvO + (size_t)s * nbo, nbo, nbatch, bt)) return -1;
  }
  w.inv_ready = true;
  return 0;
}

// Forward / backward substitution for nrhs right-hand sides (n x nrhs, ldx), in place.  Outer blocks
// whose inverse is available (lu_invert_outer) take one product with the inverse and one rank-nbo
// update of everything beyond; the others (a ragged last block, stand-alone lrsolve) go through
// block steps with the inverted nb x nb diagonal blocks.
// col_first_row (optional, host, nrhs entries, non-decreasing): row of the first non-zero of every
// right-hand side.  The forward substitution then skips, per outer block, the columns that are still
// zero there (patch loads: column s is non-zero only on its own membrane, generate_db.py:84) -- at the
// 4x4 array half of the (block, column) pairs.
int lu_solve(cudaStream_t st, const LuWork &w, const cplx *a, size_t lda, cplx *x, uint nrhs,
             size_t ldx, const uint *col_first_row) {
  const uint n = w.n;
  auto active = [&](uint row_end) -> uint {   // columns whose first non-zero row is below row_end
    if (!col_first_row) return nrhs;
    uint c = 0;
    while (c < nrhs && col_first_row[c] < row_end) c++;
    return c;
  };
  const bool inv = w.inv_ready && g_lu_outer_inverse;
  const uint NBO = inv ? w.inv_nbo : (uint)((g_lu_outer / NB) * NB > 0 ? (g_lu_outer / NB) * NB : NB);
  if (inv && w.ytmp_elems < (size_t)NBO * nrhs) {
    cudaFree(w.ytmp);
    w.ytmp = nullptr;
    CNLD_CK(cudaMalloc(&w.ytmp, sizeof(cplx) * (size_t)NBO * nrhs));
    w.ytmp_elems = (size_t)NBO * nrhs;
  }
  const uint nouter = (n + NBO - 1) / NBO;
  for (uint ob = 0; ob < nouter; ob++) {  // L y = b
    const uint O = ob * NBO, OE = (n - O < NBO) ? n : O + NBO;
    const uint na = active(OE);     // the other columns are zero in rows [0, OE) and stay zero
    if (!na) continue;
    if (inv && ob < w.inv_nfull) {
      if (zgemm3m(st, NBO, na, NBO, 1.0, w.linvO + (size_t)ob * NBO * NBO, NBO, x + O, ldx, 0.0, w.ytmp, NBO)) return -1;
      CNLD_CK(cudaMemcpy2DAsync(x + O, sizeof(cplx) * ldx, w.ytmp, sizeof(cplx) * NBO, sizeof(cplx) * NBO, na,
                                cudaMemcpyDeviceToDevice, st));
      if (n > OE && zgemm3m(st, n - OE, na, NBO, -1.0, a + (size_t)O * lda + OE, lda, w.ytmp, NBO, 1.0, x + OE, ldx))
        return -1;
      continue;
    }
    for (uint o = O; o < OE; o += NB) {
      const uint b = o / NB, nbk = (OE - o < (uint)NB) ? OE - o : NB, in_rows = OE - o - nbk;
      cplx *xb = x + o;
      if (zgemm(st, nbk, na, nbk, 1.0, w.linv + (size_t)b * NB * NB, NB, xb, ldx, 0.0, xb, ldx)) return -1;
      if (in_rows && zgemm3m(st, in_rows, na, nbk, -1.0, a + (size_t)o * lda + o + nbk, lda, xb, ldx, 1.0,
                             xb + nbk, ldx)) return -1;
    }
    if (n > OE && zgemm3m(st, n - OE, na, OE - O, -1.0, a + (size_t)O * lda + OE, lda, x + O, ldx, 1.0, x + OE, ldx))
      return -1;
  }
  for (uint ob = nouter; ob-- > 0;) {  // R x = y
    const uint O = ob * NBO, OE = (n - O < NBO) ? n : O + NBO;
    if (inv && ob < w.inv_nfull) {
      if (zgemm3m(st, NBO, nrhs, NBO, 1.0, w.uinvO + (size_t)ob * NBO * NBO, NBO, x + O, ldx, 0.0, w.ytmp, NBO)) return -1;
      CNLD_CK(cudaMemcpy2DAsync(x + O, sizeof(cplx) * ldx, w.ytmp, sizeof(cplx) * NBO, sizeof(cplx) * NBO, nrhs,
                                cudaMemcpyDeviceToDevice, st));
      if (O && zgemm3m(st, O, nrhs, NBO, -1.0, a + (size_t)O * lda, lda, w.ytmp, NBO, 1.0, x, ldx)) return -1;
      continue;
    }
    const uint nin = (OE - O + NB - 1) / NB;
    for (uint ib = nin; ib-- > 0;) {
      const uint o = O + ib * NB, b = o / NB, nbk = (OE - o < (uint)NB) ? OE - o : NB;
      cplx *xb = x + o;
      if (zgemm(st, nbk, nrhs, nbk, 1.0, w.uinv + (size_t)b * NB * NB, NB, xb, ldx, 0.0, xb, ldx)) return -1;
      if (o > O && zgemm3m(st, o - O, nrhs, nbk, -1.0, a + (size_t)o * lda + O, lda, xb, ldx, 1.0, x + O, ldx)) return -1;
    }
    if (O && zgemm3m(st, O, nrhs, OE - O, -1.0, a + (size_t)O * lda, lda, x + O, ldx, 1.0, x, ldx)) return -1;
  }
  return 0;
}

}  // namespace cnld
