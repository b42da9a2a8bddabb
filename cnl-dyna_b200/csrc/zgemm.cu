// Complex-FP64 GEMM on the FP64 tensor pipe (DMMA.8x8x4 via mma.sync.m8n8k4.f64;
// tcgen05 has no f64 kind, and every larger f64 mma shape lowers to DMMA.8x8x4 on
// sm_100a -- checked with cuobjdump).  This is the one dense contraction of the
// hot path: the trailing update of the blocked unpivoted LU
// (lrdecomp_amatrix, include/factorizations.h:173-182), plus the triangular
// panel/solve products (lrsolve_amatrix_avector, :233).
//
// C(MxN) = beta*C + alpha*A(MxK)*B(KxN), column-major interleaved complex.
// Complex product as 4 real DMMAs per 8x8x4 tile:
//   Cre += Are*Bre + (-Aim)*Bim ;  Cim += Are*Bim + Aim*Bre
// CTA tile 128x64, 8 warps (4x2), warp tile 32x32 (128 accumulator registers),
// K chunks of 8 through a 4-stage cp.async ring.  Shared-memory leading dimensions
// are chosen so the 16-byte fragment loads are bank-conflict free:
//   A tile [k][i], ld = 130 (== 2 mod 8);  B tile [j][k], ld = 12 (== 4 mod 8).
#include <algorithm>

#include "common.cuh"

namespace cnld {

namespace {
constexpr int BK = 8;
constexpr int BS_LD = BK + 4;
template <int WM, int WN, int STAGES_>
struct Cfg {
  static constexpr int BM = 32 * WM, BN = 32 * WN, STAGES = STAGES_, NT = 32 * WM * WN;
  static constexpr int AS_LD = BM + 2;
  static constexpr int A_STAGE = BK * AS_LD;  // cplx elements
  static constexpr int B_STAGE = BN * BS_LD;
  static constexpr int SMEM_BYTES = STAGES * (A_STAGE + B_STAGE) * (int)sizeof(cplx);
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <class CF, int MINB>
__global__ void __launch_bounds__(CF::NT, MINB)
zgemm_kernel(uint M, uint N, uint K, double alpha, const cplx *__restrict__ A0, size_t lda,
             const cplx *__restrict__ B0, size_t ldb, double beta, cplx *C0, size_t ldc, const ZBatch bt) {
  const long long bo = blockIdx.z / bt.per, bj = blockIdx.z % bt.per;
  const cplx *__restrict__ A = A0 + bo * bt.sAo + bj * bt.sAj;
  const cplx *__restrict__ B = B0 + bo * bt.sBo + bj * bt.sBj;
  cplx *C = C0 + bo * bt.sCo + bj * bt.sCj;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx *As = reinterpret_cast<cplx *>(smem_raw);
  constexpr int BM = CF::BM, BN = CF::BN, STAGES = CF::STAGES, NT = CF::NT, AS_LD = CF::AS_LD,
                A_STAGE = CF::A_STAGE, B_STAGE = CF::B_STAGE;
  constexpr int WMC = BM / 32;
  cplx *Bs = As + STAGES * A_STAGE;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WMC, wn = warp / WMC;
  const uint m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int nk = (K + BK - 1) / BK;

  auto load_stage = [&](int kc, int stage) {
    const uint k0 = kc * BK;
    cplx *as = As + stage * A_STAGE;
    cplx *bs = Bs + stage * B_STAGE;
    const int i = tid % BM;
#pragma unroll
    for (int r = 0; r < BM * BK / NT; r++) {
      int kk = tid / BM + (NT / BM) * r;
      bool ok = (m0 + i < M) && (k0 + kk < K);
      const cplx *src = ok ? (A + (size_t)(k0 + kk) * lda + (m0 + i)) : A;
      cp_async16(as + kk * AS_LD + i, src, ok);
    }
    const int kb = tid & 7;
#pragma unroll
    for (int r = 0; r < BN * BK / NT; r++) {
      int j = (tid >> 3) + (NT / 8) * r;
      bool ok = (n0 + j < N) && (k0 + kb < K);
      const cplx *src = ok ? (B + (size_t)(n0 + j) * ldb + (k0 + kb)) : B;
      cp_async16(bs + j * BS_LD + kb, src, ok);
    }
  };

  double cre[4][4][2], cim[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) cre[a][b][0] = cre[a][b][1] = cim[a][b][0] = cim[a][b][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }

  const int fr = lane >> 2, fc = lane & 3;
  for (int kc = 0; kc < nk; kc++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int nx = kc + STAGES - 1;
      if (nx < nk) load_stage(nx, nx % STAGES);
      cp_async_commit();
    }
    const cplx *as = As + (kc % STAGES) * A_STAGE + wm * 32 + fr;
    const cplx *bs = Bs + (kc % STAGES) * B_STAGE + (wn * 32 + fr) * BS_LD + fc;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ks++) {
      double are[4], aim[4], nai[4], bre[4], bim[4];
#pragma unroll
      for (int mi = 0; mi < 4; mi++) {
        cplx v = as[(ks * 4 + fc) * AS_LD + mi * 8];
        are[mi] = v.x; aim[mi] = v.y; nai[mi] = -v.y;
      }
#pragma unroll
      for (int ni = 0; ni < 4; ni++) {
        cplx v = bs[ni * 8 * BS_LD + ks * 4];
        bre[ni] = v.x; bim[ni] = v.y;
      }
#pragma unroll
      for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
          dmma(cre[mi][ni][0], cre[mi][ni][1], are[mi], bre[ni]);
          dmma(cim[mi][ni][0], cim[mi][ni][1], are[mi], bim[ni]);
          dmma(cre[mi][ni][0], cre[mi][ni][1], nai[mi], bim[ni]);
          dmma(cim[mi][ni][0], cim[mi][ni][1], aim[mi], bre[ni]);
        }
    }
  }
  cp_async_wait<0>();

  // epilogue: lane holds C[fr][2*fc + e] of each 8x8 block
  // (all loads of a row-group are issued before its stores: C is not restrict-qualified, so a
  //  naive load/add/store chain would serialise 32 HBM round trips per thread)
#pragma unroll
  for (int mi = 0; mi < 4; mi++) {
    const uint row = m0 + wm * 32 + mi * 8 + fr;
    const bool rok = row < M;
    cplx old[4][2];
#pragma unroll
    for (int ni = 0; ni < 4; ni++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        uint col = n0 + wn * 32 + ni * 8 + fc * 2 + e;
        old[ni][e] = make_double2(0.0, 0.0);
        if (beta != 0.0 && rok && col < N) old[ni][e] = __ldcg(C + (size_t)col * ldc + row);
      }
#pragma unroll
    for (int ni = 0; ni < 4; ni++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        uint col = n0 + wn * 32 + ni * 8 + fc * 2 + e;
        if (rok && col < N)
          C[(size_t)col * ldc + row] = make_double2(fma(alpha, cre[mi][ni][e], old[ni][e].x),
                                                    fma(alpha, cim[mi][ni][e], old[ni][e].y));
      }
  }
}

// ---- 3M variant -------------------------------------------------------------------------------
// Complex product with three real products instead of four:
//   X = Are*Bre, Y = Aim*Bim, W = (Are+Aim)*(Bre+Bim);  Cre = X - Y,  Cim = W - X - Y
// 25 % fewer DMMAs on the FP64 tensor pipe, which is what bounds the trailing updates.  Three
// accumulator sets per warp tile, so the warp tile is 32x24 (144 accumulator registers) and the
// CTA tile 64x48 (4 warps), still 2 CTAs/SM.  Used for C = C - A*B with C not aliasing A or B.
constexpr int T3_BM = 64, T3_BN = 48, T3_NT = 128, T3_STAGES = 4, T3_GM = 16;
constexpr int T3_AS_LD = T3_BM + 2;
constexpr int T3_A_STAGE = BK * T3_AS_LD, T3_B_STAGE = T3_BN * BS_LD;
constexpr int T3_SMEM = T3_STAGES * (T3_A_STAGE + T3_B_STAGE) * (int)sizeof(cplx);

__global__ void __launch_bounds__(T3_NT, 2)
zgemm3m_kernel(uint M, uint N, uint K, double alpha, const cplx *__restrict__ A0, size_t lda,
               const cplx *__restrict__ B0, size_t ldb, double beta, cplx *C0, size_t ldc, int lower_only,
               uint ntiles1, uint nbatch, const ZBatch bt) {
  const uint ntiles = ntiles1 * nbatch;   // batch item q owns tiles [q * ntiles1, (q + 1) * ntiles1)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx *As = reinterpret_cast<cplx *>(smem_raw);
  cplx *Bs = As + T3_STAGES * T3_A_STAGE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 1, wn = warp >> 1;
  // Persistent CTAs (grid = 2 per SM): every CTA walks the tile list, so the kernel never has
  // pending CTAs and the block scheduler can place a co-resident assembly CTA of another frequency
  // (FP64 ALU pipe) next to the two GEMM CTAs (DMMA tensor pipe) of an SM.
  // Grouped rasterisation: consecutive tiles walk T3_GM row tiles x all column tiles, so the CTAs
  // in flight share a few MB of A/B panels in L2 instead of streaming a whole panel.
  const uint mt = (M + T3_BM - 1) / T3_BM, nt = (N + T3_BN - 1) / T3_BN;
  const uint gsz_full = T3_GM * nt;
  const int fr = lane >> 2, fc = lane & 3;
  const int nk = (K + BK - 1) / BK;

  // tile index -> (m0, n0); false for a tile that lies entirely above the diagonal (lower_only)
  auto coords = [&](uint tile, uint &m0, uint &n0) -> bool {
    tile %= ntiles1;
    uint first, gm, within;
    if (!lower_only) {
      const uint grp = tile / gsz_full;
      first = grp * T3_GM;
      gm = min((uint)T3_GM, mt - first);
      within = tile % gsz_full;
    } else {
      // compacted list: row group g keeps only the column tiles that reach its last row's diagonal,
      // so every CTA of the persistent grid gets the same number of computed tiles
      uint t = tile;
      first = 0;
      for (;;) {
        gm = min((uint)T3_GM, mt - first);
        const uint cols = min(nt, ((first + gm) * T3_BM + T3_BN - 1) / T3_BN);
        if (t < gm * cols) break;
        t -= gm * cols;
        first += T3_GM;
      }
      within = t;
    }
    m0 = (first + within % gm) * T3_BM;
    n0 = (within / gm) * T3_BN;
    return !(lower_only && m0 + T3_BM <= n0);
  };

  auto batch_off = [&](uint tile, long long sa, long long sb) -> long long {
    const uint q = tile / ntiles1;
    return (long long)(q / bt.per) * sa + (long long)(q % bt.per) * sb;
  };
  auto load_stage = [&](const cplx *__restrict__ A, const cplx *__restrict__ B, uint m0, uint n0, int kc, int stage) {
    const uint k0 = kc * BK;
    cplx *as = As + stage * T3_A_STAGE;
    cplx *bs = Bs + stage * T3_B_STAGE;
    const int i = tid % T3_BM;
#pragma unroll
    for (int r = 0; r < T3_BM * BK / T3_NT; r++) {
      int kk = tid / T3_BM + (T3_NT / T3_BM) * r;
      bool ok = (m0 + i < M) && (k0 + kk < K);
      const cplx *src = ok ? (A + (size_t)(k0 + kk) * lda + (m0 + i)) : A;
      cp_async16(as + kk * T3_AS_LD + i, src, ok);
    }
    const int kb = tid & 7;
#pragma unroll
    for (int r = 0; r < T3_BN * BK / T3_NT; r++) {
      int j = (tid >> 3) + (T3_NT / 8) * r;
      bool ok = (n0 + j < N) && (k0 + kb < K);
      const cplx *src = ok ? (B + (size_t)(n0 + j) * ldb + (k0 + kb)) : B;
      cp_async16(bs + j * BS_LD + kb, src, ok);
    }
  };
  auto prologue = [&](const cplx *A, const cplx *B, uint m0, uint n0) {
#pragma unroll
    for (int s = 0; s < T3_STAGES - 1; s++) {
      if (s < nk) load_stage(A, B, m0, n0, s, s);
      cp_async_commit();
    }
  };

  uint tile = blockIdx.x, m0 = 0, n0 = 0;
  while (tile < ntiles && !coords(tile, m0, n0)) tile += gridDim.x;
  if (tile >= ntiles) return;
  const cplx *A = A0 + batch_off(tile, bt.sAo, bt.sAj), *B = B0 + batch_off(tile, bt.sBo, bt.sBj);
  prologue(A, B, m0, n0);
  for (;;) {
  cplx *C = C0 + batch_off(tile, bt.sCo, bt.sCj);
  // the C tile is read ~70 us from now: pull its 384 lines into L2 while the main loop runs
  if (beta != 0.0) {
#pragma unroll
    for (int r = 0; r < (T3_BN * T3_BM / 8) / T3_NT; r++) {
      const uint l = tid + r * T3_NT, col = n0 + l / (T3_BM / 8), row = m0 + (l % (T3_BM / 8)) * 8;
      if (col < N && row < M) asm volatile("prefetch.global.L2 [%0];" ::"l"(C + (size_t)col * ldc + row));
    }
  }
  double X[4][3][2], Y[4][3][2], W[4][3][2];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 3; b++)
#pragma unroll
      for (int e = 0; e < 2; e++) X[a][b][e] = Y[a][b][e] = W[a][b][e] = 0.0;

  for (int kc = 0; kc < nk; kc++) {
    cp_async_wait<T3_STAGES - 2>();
    __syncthreads();
    {
      int nx = kc + T3_STAGES - 1;
      if (nx < nk) load_stage(A, B, m0, n0, nx, nx % T3_STAGES);
      cp_async_commit();
    }
    const cplx *as = As + (kc % T3_STAGES) * T3_A_STAGE + wm * 32 + fr;
    const cplx *bs = Bs + (kc % T3_STAGES) * T3_B_STAGE + (wn * 24 + fr) * BS_LD + fc;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ks++) {
      double are[4], aim[4], asu[4], bre[3], bim[3], bsu[3];
#pragma unroll
      for (int mi = 0; mi < 4; mi++) {
        cplx v = as[(ks * 4 + fc) * T3_AS_LD + mi * 8];
        are[mi] = v.x; aim[mi] = v.y; asu[mi] = v.x + v.y;
      }
#pragma unroll
      for (int ni = 0; ni < 3; ni++) {
        cplx v = bs[ni * 8 * BS_LD + ks * 4];
        bre[ni] = v.x; bim[ni] = v.y; bsu[ni] = v.x + v.y;
      }
#pragma unroll
      for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 3; ni++) {
          dmma(X[mi][ni][0], X[mi][ni][1], are[mi], bre[ni]);
          dmma(Y[mi][ni][0], Y[mi][ni][1], aim[mi], bim[ni]);
          dmma(W[mi][ni][0], W[mi][ni][1], asu[mi], bsu[ni]);
        }
    }
  }
  cp_async_wait<0>();
  __syncthreads();  // every warp is done with the shared-memory ring
  // start the next tile's first stages now, so they fly while this tile's epilogue runs.
  // (Static tile list on purpose: a ticket counter balances one launch better -- 29.6 vs 27.5 TFLOP/s
  // for the first trailing update -- but then the kernel holds every SM until its last tile, and the
  // latency-bound chains of the other frequencies in flight starve: 6.8 vs 7.2 frequencies/s.)
  uint ntile = tile + gridDim.x, nm0 = 0, nn0 = 0;
  while (ntile < ntiles && !coords(ntile, nm0, nn0)) ntile += gridDim.x;
  const bool more = ntile < ntiles;
  const cplx *An = A, *Bn = B;
  if (more) {
    An = A0 + batch_off(ntile, bt.sAo, bt.sAj);
    Bn = B0 + batch_off(ntile, bt.sBo, bt.sBj);
    prologue(An, Bn, nm0, nn0);
  }
#pragma unroll
  for (int mh = 0; mh < 2; mh++) {
    cplx old[2][3][2];
#pragma unroll
    for (int mi = 0; mi < 2; mi++) {
      const uint row = m0 + wm * 32 + (mh * 2 + mi) * 8 + fr;
#pragma unroll
      for (int ni = 0; ni < 3; ni++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          uint col = n0 + wn * 24 + ni * 8 + fc * 2 + e;
          old[mi][ni][e] = make_double2(0.0, 0.0);
          if (beta != 0.0 && row < M && col < N) old[mi][ni][e] = __ldcg(C + (size_t)col * ldc + row);
        }
    }
#pragma unroll
    for (int mi = 0; mi < 2; mi++) {
      const uint row = m0 + wm * 32 + (mh * 2 + mi) * 8 + fr;
#pragma unroll
      for (int ni = 0; ni < 3; ni++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          uint col = n0 + wn * 24 + ni * 8 + fc * 2 + e;
          if (row < M && col < N) {
            double cr = X[mh * 2 + mi][ni][e] - Y[mh * 2 + mi][ni][e];
            double ci = W[mh * 2 + mi][ni][e] - X[mh * 2 + mi][ni][e] - Y[mh * 2 + mi][ni][e];
            C[(size_t)col * ldc + row] = make_double2(fma(alpha, cr, old[mi][ni][e].x), fma(alpha, ci, old[mi][ni][e].y));
          }
        }
    }
  }
  if (!more) break;
  tile = ntile; m0 = nm0; n0 = nn0; A = An; B = Bn;
  }  // tile loop
}
}  // namespace

int g_zgemm_3m = 1;  // use the 3M kernel for non-aliased accumulating products

static int zgemm3m_impl(cudaStream_t st, uint M, uint N, uint K, double alpha, const cplx *A, size_t lda,
                        const cplx *B, size_t ldb, double beta, cplx *C, size_t ldc, int reserve_sms, bool lower_only,
                        double *entries, uint nbatch, const ZBatch &bt) {
  double cnt = (double)M * N;
  if (lower_only) {
    cnt = 0.0;
    for (uint n0 = 0; n0 < N; n0 += T3_BN) {
      const uint nn = std::min((uint)T3_BN, N - n0);
      // row tiles with m0 + BM > n0
      const uint mfirst = (n0 + 1 > (uint)T3_BM) ? ((n0 + 1 - T3_BM + T3_BM - 1) / T3_BM) * T3_BM : 0;
      if (mfirst < M) cnt += (double)(M - mfirst) * nn;
    }
  }
  if (entries) *entries = cnt;
  if (M == 0 || N == 0) return 0;
  if (!g_zgemm_3m) return zgemm_batched(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, nbatch, bt);
  g_dmma_flops.fetch_add((unsigned long long)(6.0 * cnt * K * nbatch), std::memory_order_relaxed);
  static bool attr_set[64] = {false};
  int dev = 0;
  CNLD_CK(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    CNLD_CK(cudaFuncSetAttribute(zgemm3m_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SMEM));
    attr_set[dev & 63] = true;
  }
  static int nsm[64] = {0};
  if (!nsm[dev & 63]) CNLD_CK(cudaDeviceGetAttribute(&nsm[dev & 63], cudaDevAttrMultiProcessorCount, dev));
  const uint mt = (M + T3_BM - 1) / T3_BM, ntn = (N + T3_BN - 1) / T3_BN;
  uint tiles = mt * ntn;
  if (lower_only) {  // same enumeration as the kernel's compacted list
    tiles = 0;
    for (uint first = 0; first < mt; first += T3_GM) {
      const uint gm = std::min((uint)T3_GM, mt - first);
      tiles += gm * std::min(ntn, ((first + gm) * T3_BM + T3_BN - 1) / T3_BN);
    }
  }
  // reserve_sms: leave that many SMs without a CTA of this (persistent) kernel, for concurrent
  // latency-bound kernels on another stream
  const int use_sms = std::max(1, nsm[dev & 63] - std::max(0, reserve_sms));
  dim3 grid(std::min(tiles * nbatch, (uint)(2 * use_sms)));
  zgemm3m_kernel<<<grid, T3_NT, T3_SMEM, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, lower_only ? 1 : 0, tiles,
                                               nbatch, bt);
  CNLD_CK_LAUNCH();
  return 0;
}

int zgemm3m(cudaStream_t st, uint M, uint N, uint K, double alpha, const cplx *A, size_t lda,
            const cplx *B, size_t ldb, double beta, cplx *C, size_t ldc, int reserve_sms, bool lower_only,
            double *entries) {
  return zgemm3m_impl(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, reserve_sms, lower_only, entries, 1, ZBatch());
}

// batched 3M product (C must not alias A or B)
int zgemm3m_batched(cudaStream_t st, uint M, uint N, uint K, double alpha, const cplx *A, size_t lda,
                    const cplx *B, size_t ldb, double beta, cplx *C, size_t ldc, uint nbatch, const ZBatch &bt) {
  if (!nbatch) return 0;
  return zgemm3m_impl(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, 0, false, nullptr, nbatch, bt);
}

namespace {
}  // namespace

int g_zgemm_variant = 1;  // 0: 128x64 tile, 1 CTA/SM;  1: 64x64 tile, 2 CTAs/SM

template <class CF, int MINB>
static int launch(cudaStream_t st, uint M, uint N, uint K, double alpha, const cplx *A, size_t lda,
                  const cplx *B, size_t ldb, double beta, cplx *C, size_t ldc, uint nbatch = 1,
                  const ZBatch &bt = ZBatch()) {
  static bool attr_set[64] = {false};
  int dev = 0;
  CNLD_CK(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    CNLD_CK(cudaFuncSetAttribute(zgemm_kernel<CF, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 CF::SMEM_BYTES));
    attr_set[dev & 63] = true;
  }
  g_dmma_flops.fetch_add((unsigned long long)(8.0 * M * N * K * nbatch), std::memory_order_relaxed);
  dim3 grid((M + CF::BM - 1) / CF::BM, (N + CF::BN - 1) / CF::BN, nbatch);
  zgemm_kernel<CF, MINB><<<grid, CF::NT, CF::SMEM_BYTES, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bt);
  CNLD_CK_LAUNCH();
  return 0;
}

int zgemm(cudaStream_t st, uint M, uint N, uint K, double alpha, const cplx *A, size_t lda,
          const cplx *B, size_t ldb, double beta, cplx *C, size_t ldc) {
  if (M == 0 || N == 0) return 0;
  if (g_zgemm_variant == 0) return launch<Cfg<4, 2, 4>, 1>(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  return launch<Cfg<2, 2, 4>, 2>(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
}

int zgemm_batched(cudaStream_t st, uint M, uint N, uint K, double alpha, const cplx *A, size_t lda,
                  const cplx *B, size_t ldb, double beta, cplx *C, size_t ldc, uint nbatch, const ZBatch &bt) {
  if (M == 0 || N == 0 || nbatch == 0) return 0;
  return launch<Cfg<2, 2, 4>, 2>(st, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, nbatch, bt);
}

}  // namespace cnld
