// Time-domain FIR convolution of the contact simulation (C ABI: cnld_b200_fir_*).
//
// Restates fir_conv_cy (cnld/simulation.pyx:51-77),
//     x[j] = dt * sum_i sum_{k < min(nsample, nfir - offset)} fir[i][j][k + offset] * p[nsample - 1 - k][i],
// the convolution FixedStepSolver runs several times per time step (:400-410: `_fir_conv_base` over the
// whole history with offset 1, once per step, and `_fir_conv_add` for the current sample with offset 0 in
// every fixed-point iteration, :459-503).  The patch-to-patch impulse-response table fir[P][P][nfir]
// (663 MB at the 4x4 array) stays on the device; one convolution reads it once -- an HBM-bound kernel
// (2 flops per 8 bytes).  Thread block = (source i, slice of the destinations): the source's history window
// sits in shared memory and every warp streams whole contiguous table rows against it; the per-(j, i)
// partial sums are reduced over i in a fixed order, so results do not depend on scheduling.
// The stepping entry points keep the pressure history on the device too: a step uploads one row of P values.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

#include "../../include/cnld_b200.h"
#include "common.cuh"

using namespace cnld;

struct cnld_fir {
  int device = 0;
  uint nsrc = 0, ndest = 0, nfir = 0, ngroup = 1, per = 0;
  double *d_fir = nullptr;     // [nsrc][ndest][nfir]
  double *d_hist = nullptr;    // source-major history: hist[i][t], t < cap (resident stepping)
  double *d_p = nullptr;       // staging of a caller-provided history, source-major
  double *d_part = nullptr, *d_x = nullptr, *d_row = nullptr;
  size_t hist_cap = 0, p_cap = 0, nhist = 0;
  cudaStream_t st = nullptr;
};

namespace {
constexpr int FC_THREADS = 256, FC_KTILE = 4096;

// CTA = (source i, slice of the destinations); the source's history window is staged once in shared memory
// (reversed, so lag k pairs with table entry k) and every warp streams whole table rows fir[i][j][offset ..
// offset + K) -- 32 KB of contiguous doubles per row at 4000 taps -- against it.  partial[j][i] = row dot window.
__global__ void __launch_bounds__(FC_THREADS)
k_fir_conv(uint nsrc, uint ndest, uint nfir, uint per, uint K, uint offset, const double *__restrict__ fir,
           const double *__restrict__ h, size_t hs, size_t last, double *partial, uint pj, uint pi) {
  __shared__ double win[FC_KTILE];
  // the solver's step kernel is launched as a programmatic dependent: it may become resident (and stage its
  // frequency-independent operands) while this grid drains; it still waits for this grid's results
  asm volatile("griddepcontrol.launch_dependents;");
  const uint i = blockIdx.x, g = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint j0 = g * per, j1 = min(ndest, j0 + per);
  const double *hp = h + (size_t)i * hs + last;
  for (uint k0 = 0; k0 < K; k0 += FC_KTILE) {
    const uint kt = min((uint)FC_KTILE, K - k0);
    __syncthreads();
    for (uint k = threadIdx.x; k < kt; k += FC_THREADS) win[k] = hp[-(long long)(k0 + k)];
    __syncthreads();
    for (uint j = j0 + warp; j < j1; j += FC_THREADS / 32) {
      const double *f = fir + ((size_t)i * ndest + j) * nfir + offset + k0;
      // 16-byte streaming loads, four per lane in flight (2 KB per warp): a scalar head element when the row
      // segment starts on an odd double, pairs (head + 2q, head + 2q + 1) after it, a scalar tail element
      const uint head = (uint)((reinterpret_cast<uintptr_t>(f) >> 3) & 1u) & (kt > 0u);
      const uint np = (kt - head) >> 1;
      const double2 *f2 = reinterpret_cast<const double2 *>(f + head);
      const double *w = win + head;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      uint q = lane;
      for (; q + 96 < np; q += 128) {
        const double2 v0 = __ldcs(f2 + q), v1 = __ldcs(f2 + q + 32), v2 = __ldcs(f2 + q + 64), v3 = __ldcs(f2 + q + 96);
        a0 = fma(v0.y, w[2 * q + 1], fma(v0.x, w[2 * q], a0));
        a1 = fma(v1.y, w[2 * q + 65], fma(v1.x, w[2 * q + 64], a1));
        a2 = fma(v2.y, w[2 * q + 129], fma(v2.x, w[2 * q + 128], a2));
        a3 = fma(v3.y, w[2 * q + 193], fma(v3.x, w[2 * q + 192], a3));
      }
      for (; q < np; q += 32) {
        const double2 v0 = __ldcs(f2 + q);
        a0 = fma(v0.y, w[2 * q + 1], fma(v0.x, w[2 * q], a0));
      }
      if (lane == 0 && head) a1 = fma(f[0], win[0], a1);
      if (lane == 1 && head + 2 * np < kt) a1 = fma(f[kt - 1], win[kt - 1], a1);
      double acc = (a0 + a1) + (a2 + a3);
      for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (lane == 0) {
        double *dst = partial + (size_t)j * pj + (size_t)i * pi;
        *dst = k0 ? *dst + acc : acc;
      }
    }
  }
}
__global__ void k_fir_finish(uint ndest, uint nsrc, const double *__restrict__ partial, double dt, double *x) {
  const uint j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ndest) return;
  double s = 0.0;
  for (uint i = 0; i < nsrc; i++) s += partial[(size_t)j * nsrc + i];
  x[j] = s * dt;
}
// p[t][i] (time-major, as the caller holds it) -> h[i][t]
__global__ void k_fir_transpose(uint nsample, uint nsrc, const double *__restrict__ p, double *h) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)nsample * nsrc) return;
  const uint i = e % nsrc;
  const size_t t = e / nsrc;
  h[(size_t)i * nsample + t] = p[e];
}
__global__ void k_fir_push(uint nsrc, size_t cap, size_t pos, const double *__restrict__ row, double *hist) {
  const uint i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nsrc) hist[(size_t)i * cap + pos] = row[i];
}

int run_conv(cnld_fir *f, const double *d_h, size_t hs, size_t last, uint K, uint offset, double dt, double *x_host) {
  if (K) {
    k_fir_conv<<<dim3(f->nsrc, f->ngroup), FC_THREADS, 0, f->st>>>(f->nsrc, f->ndest, f->nfir, f->per, K, offset, f->d_fir, d_h,
                                                                  hs, last, f->d_part, f->nsrc, 1);
    CNLD_CK_LAUNCH();
  } else {
    CNLD_CK(cudaMemsetAsync(f->d_part, 0, sizeof(double) * (size_t)f->ndest * f->nsrc, f->st));
  }
  k_fir_finish<<<(f->ndest + 127) / 128, 128, 0, f->st>>>(f->ndest, f->nsrc, f->d_part, dt, f->d_x);
  CNLD_CK_LAUNCH();
  CNLD_CK(cudaMemcpyAsync(x_host, f->d_x, sizeof(double) * f->ndest, cudaMemcpyDeviceToHost, f->st));
  CNLD_CK(cudaStreamSynchronize(f->st));
  return 0;
}
}  // namespace

extern "C" {

cnld_fir *cnld_b200_fir_create(int device, cnld_uint nsrc, cnld_uint ndest, cnld_uint nfir, const double *fir) {
  if (cnld_b200_device_count() <= device) { set_error("no CUDA device %d (libcnld_b200 has no CPU fallback)", device); return nullptr; }
  if (!nsrc || !ndest || !nfir || !fir) { set_error("fir_create: empty table"); return nullptr; }
  cnld_fir *f = new cnld_fir();
  f->device = device; f->nsrc = nsrc; f->ndest = ndest; f->nfir = nfir;
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
  // CTAs = sources x destination slices: about four per SM, at least 8 destinations (one per warp) per slice
  f->ngroup = std::max(1u, std::min((ndest + 7) / 8, (uint)((4 * nsm + nsrc - 1) / nsrc)));
  f->per = (ndest + f->ngroup - 1) / f->ngroup;
  f->ngroup = (ndest + f->per - 1) / f->per;
  const size_t tab = (size_t)nsrc * ndest * nfir;
  bool ok = cudaSetDevice(device) == cudaSuccess && cudaMalloc(&f->d_fir, sizeof(double) * tab) == cudaSuccess &&
            cudaMalloc(&f->d_part, sizeof(double) * (size_t)ndest * nsrc) == cudaSuccess &&
            cudaMalloc(&f->d_x, sizeof(double) * ndest) == cudaSuccess && cudaMalloc(&f->d_row, sizeof(double) * nsrc) == cudaSuccess &&
            cudaMemcpy(f->d_fir, fir, sizeof(double) * tab, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaStreamCreateWithFlags(&f->st, cudaStreamNonBlocking) == cudaSuccess;
  if (!ok) {
    set_error("fir_create: %s", cudaGetErrorString(cudaGetLastError()));
    cnld_b200_fir_destroy(f);
    return nullptr;
  }
  return f;
}

void cnld_b200_fir_destroy(cnld_fir *f) {
  if (!f) return;
  cudaSetDevice(f->device);
  cudaFree(f->d_fir); cudaFree(f->d_hist); cudaFree(f->d_p); cudaFree(f->d_part); cudaFree(f->d_x); cudaFree(f->d_row);
  if (f->st) cudaStreamDestroy(f->st);
  delete f;
}

/* fir_conv_cy(fir, p, dt, offset) (simulation.pyx:51-77): p[nsample][nsrc] HOST, x[ndest] HOST */
int cnld_b200_fir_conv(cnld_fir *f, const double *p, cnld_uint nsample, double dt, int offset, double *x) {
  if (!f || !x || (nsample && !p) || offset < 0) { set_error("fir_conv: bad arguments"); return -1; }
  CNLD_CK(cudaSetDevice(f->device));
  const uint K = (uint)offset >= f->nfir ? 0u : std::min(nsample, f->nfir - (uint)offset);
  if (nsample) {
    const size_t need = 2 * (size_t)nsample * f->nsrc;
    if (f->p_cap < need) {
      cudaFree(f->d_p);
      f->d_p = nullptr;
      CNLD_CK(cudaMalloc(&f->d_p, sizeof(double) * need));
      f->p_cap = need;
    }
    double *raw = f->d_p + (size_t)nsample * f->nsrc;
    CNLD_CK(cudaMemcpyAsync(raw, p, sizeof(double) * (size_t)nsample * f->nsrc, cudaMemcpyHostToDevice, f->st));
    k_fir_transpose<<<(unsigned)(((size_t)nsample * f->nsrc + 255) / 256), 256, 0, f->st>>>(nsample, f->nsrc, raw, f->d_p);
    CNLD_CK_LAUNCH();
  }
  return run_conv(f, f->d_p, nsample, nsample ? nsample - 1 : 0, K, (uint)offset, dt, x);
}

/* Resident stepping (simulation.pyx:400-410, :459-503).  push: the total pressure of a finished time step
 * (p_row[nsrc]) joins the device history.  base: `_fir_conv_base` = convolution of the whole history with
 * offset 1.  add: `_fir_conv_add` = the current sample alone with offset 0. */
int cnld_b200_fir_reset(cnld_fir *f) {
  if (!f) { set_error("null fir"); return -1; }
  f->nhist = 0;
  return 0;
}
int cnld_b200_fir_push(cnld_fir *f, const double *p_row) {
  if (!f || !p_row) { set_error("fir_push: null argument"); return -1; }
  CNLD_CK(cudaSetDevice(f->device));
  if (f->nhist == f->hist_cap) {     // grow geometrically; the history is kept whole (nfir lags are read)
    const size_t cap = std::max<size_t>(2 * f->hist_cap, 1024);
    double *nh = nullptr;
    CNLD_CK(cudaMalloc(&nh, sizeof(double) * cap * f->nsrc));
    if (f->nhist)
      CNLD_CK(cudaMemcpy2DAsync(nh, sizeof(double) * cap, f->d_hist, sizeof(double) * f->hist_cap, sizeof(double) * f->nhist, f->nsrc,
                                cudaMemcpyDeviceToDevice, f->st));
    CNLD_CK(cudaStreamSynchronize(f->st));
    cudaFree(f->d_hist);
    f->d_hist = nh;
    f->hist_cap = cap;
  }
  CNLD_CK(cudaMemcpyAsync(f->d_row, p_row, sizeof(double) * f->nsrc, cudaMemcpyHostToDevice, f->st));
  k_fir_push<<<(f->nsrc + 127) / 128, 128, 0, f->st>>>(f->nsrc, f->hist_cap, f->nhist, f->d_row, f->d_hist);
  CNLD_CK_LAUNCH();
  f->nhist++;
  return 0;
}
int cnld_b200_fir_base(cnld_fir *f, double dt, double *x) {
  if (!f || !x) { set_error("fir_base: null argument"); return -1; }
  CNLD_CK(cudaSetDevice(f->device));
  const uint K = f->nfir > 1 ? (uint)std::min<size_t>(f->nhist, f->nfir - 1) : 0u;
  return run_conv(f, f->d_hist, f->hist_cap, f->nhist ? f->nhist - 1 : 0, K, 1, dt, x);
}
int cnld_b200_fir_add(cnld_fir *f, const double *p_row, double dt, double *x) {
  if (!f || !x || !p_row) { set_error("fir_add: null argument"); return -1; }
  CNLD_CK(cudaSetDevice(f->device));
  CNLD_CK(cudaMemcpyAsync(f->d_row, p_row, sizeof(double) * f->nsrc, cudaMemcpyHostToDevice, f->st));
  return run_conv(f, f->d_row, 1, 0, 1, 0, dt, x);      // history of one sample per source: stride 1
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// Fixed-step time-domain solver on the device (C ABI: cnld_b200_td_*).
//
// Restates FixedStepSolver (cnld/simulation.pyx:297-555).  One time step (`step`, :459-503) is
//   blind estimate  x = conv(p_tot[:s], offset 1)                                   (:447-457)
//   fixed point     x <- base + dt * fir[:, :, 0]^T p_tot(x),  p_tot = p_cont_spr(x) + p_cont_dmp + p_es(v, x)
//                   until max|dx| <= atol or maxiter, first with the contact damper of the previous estimate
//                   (zero), then with the damper frozen at the converged estimate   (:412-445, :466-498)
// The reference recomputes `_fir_conv_base` in the first `_update_x`; it is the blind estimate's convolution of
// the same rows, so one pass over the table per step is all the HBM traffic a step needs.  Here nothing leaves
// the device between steps: the total-pressure history IS the convolution's input (source-major rows) and the
// state arrays are downloaded when the caller asks for them.  How a run of samples is scheduled:
//   * one sample (nsteps = 1):   k_fir_conv over the whole history, then the step kernel;
//   * a block of up to TD_B = 16 samples: the history BEFORE the block is convolved once for all of them
//     (k_fir_far_mma: every table entry loaded once, 16 lags each, on DMMA), the step kernel of sample s adds the
//     lags that reach into the block's own samples (compact copy of the first lags, L2);
//   * consecutive full blocks: the pass for block k + 1 runs ahead on a second stream under the samples of block k
//     over the rows older than block k, k_fir_mid adds block k's samples at the boundary;
//   * the step kernel (k_td_step_cl, a cluster of 4 CTAs, or k_td_step on one CTA) is a programmatic dependent of
//     the kernel before it and does everything that does not depend on the previous sample before its wait.
// np.vectorize quirk kept on purpose (results identical to the reference's): the contact closures return the
// Python int 0 outside contact and np.vectorize derives its output dtype from the first element, so whenever
// patch 0 is out of contact the contact pressures of all patches are truncated toward zero (int64).
struct cnld_td {
  int device = 0;
  uint P = 0, nfir = 0, nt = 0, ngroup = 1, per = 0, cur = 1;
  uint ngroup_far = 1, per_far = 0;      // k_fir_far: destination slices sized to one wave of its own occupancy
  uint ntap = 0;                 // taps that matter: 1 + the last tap that is nonzero anywhere (<= nfir)
  double dt = 0, e0 = 0, k = 0, n = 0, x0 = 0, lmbd = 0, atol = 0;
  int maxiter = 0;
  bool pdl = true;               // step kernel launched as a programmatic dependent of the convolution (CNLD_B200_TD_PDL=0: plain)
  bool cluster = true;           // step kernel on a cluster of TD_CL CTAs (CNLD_B200_TD_CLUSTER=0: one CTA)
  uint block = 16;               // samples per pass over the table (CNLD_B200_TD_BLOCK=1: one pass per sample)
  int mma = 1;                   // history pass on the FP64 tensor pipe over an interleaved second copy of the table;
                                 // CNLD_B200_TD_MMA=0: the DFMA kernel on the row layout, blocks of 8, no second copy
  size_t ld = 0;                 // row stride of the device table: rows padded so that tap 1 sits on a 32-byte boundary
  double *tab = nullptr;         // tap 0 of row 0 (= d_fir + 3)
  double *d_firn = nullptr;      // fir[:, :, 1 .. TD_NL] as [lag - 1][source][destination]
  double *d_tab8 = nullptr;      // mma == 1: second copy of the table, interleaved by tiles of 8 destinations (k_td_interleave)
  size_t ng4 = 0;                // groups of 4 taps per row in that copy
  double *d_fir = nullptr, *d_fir0 = nullptr, *d_v = nullptr, *d_geff = nullptr, *d_hist = nullptr, *d_state = nullptr,
         *d_part = nullptr, *d_err = nullptr;
  int *d_iters = nullptr;
  unsigned long long *d_trace = nullptr;
  cudaStream_t st = nullptr, st2 = nullptr;       // st2: run-ahead history passes (lower priority)
  cudaEvent_t ev_hist = nullptr, ev_far = nullptr;
  bool overlap = true;           // history pass of block k + 1 under the samples of block k (CNLD_B200_TD_OVERLAP=0: in line)
  uint pbuf = 0;                 // plane buffer (of two) the current block reads
  uint ngroup_narrow = 1, per_narrow = 0;
};

namespace {
constexpr int TD_THREADS = 1024;
constexpr int TD_B = 16;       // blocked stepping: at most this many samples are served by one pass over the table
constexpr int TD_NL = 2 * TD_B - 1;   // lags of the compact copy: the step kernels add the block's own samples (lags < TD_B),
                                      // k_fir_mid the previous block's when the history pass ran ahead (lags <= TD_NL)
enum { TD_X = 0, TD_U, TD_PES, TD_SPR, TD_DMP, TD_PTOT, TD_NSTATE };

struct TdArgs {
  uint P, nt, s;
  double dt, e0h, k, n, x0, lmbd, atol;   // e0h = -e_0 / 2
  int maxiter, stage;                     // stage: fir[:, :, 0] fits in shared memory next to the work vectors
  uint near;                              // blocked stepping: lags 1 .. near are added from the block's own samples
  uint early;                             // the history pass and the samples up to s - 2 are complete when this kernel starts
  const double *firn;                     // fir[:, :, 1 .. TD_NL] as [lag - 1][source][destination]
  const double *fir0, *v, *geff, *part;
  double *hist, *state, *err;
  int *iters;
  unsigned long long *trace;              // optional [nt][8] %globaltimer marks of the kernel's phases (CNLD_B200_TD_TRACE=1)
};

__device__ __forceinline__ void td_mark(const TdArgs &a, int k) {
  if (a.trace && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.trace[(size_t)a.s * 8 + k] = t;
  }
}

__global__ void k_td_gather_fir0(uint P, size_t ld, const double *__restrict__ tab, double *fir0) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < (size_t)P * P) fir0[e] = tab[e * ld];
}
// 1 + the largest tap index that is nonzero in any (source, destination) pair: a Kramers-Kronig (causal) table is
// exactly zero in its second half (impulse_response.py:115-125), and taps that are zero everywhere add nothing
__global__ void k_td_last_tap(size_t rows, uint nfir, size_t ld, const double *__restrict__ tab, uint *last) {
  const size_t r = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const double *f = tab + r * ld;
  uint best = 0;
  for (uint k = threadIdx.x & 31; k < nfir; k += 32)
    if (f[k] != 0.0) best = k + 1;
  for (int off = 16; off > 0; off >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, off));
  if ((threadIdx.x & 31) == 0 && best) atomicMax(last, best);
}

// firn[lag - 1][i][j] = fir[i][j][lag], lag = 1 .. TD_NL (zero beyond the table)
__global__ void k_td_gather_near(uint P, uint nfir, size_t ld, const double *__restrict__ tab, double *firn) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x, PP = (size_t)P * P;
  if (e >= PP * TD_NL) return;
  const uint lag = (uint)(e / PP) + 1;
  firn[e] = lag < nfir ? tab[(e % PP) * ld + lag] : 0.0;
}

// Blocked stepping, the part of the next TD_B samples' convolutions that the finished history already fixes:
//   part[b][i][j] = sum_{tau > b} fir[i][j][tau] * h_i[s0 + b - tau],   b = 0 .. TD_B - 1,
// (h_i = total pressure of source i, rows < s0 known).  Every table entry is loaded once and feeds TD_B lags: the
// pass costs what one convolution costs and serves TD_B samples.  Same CTA shape as k_fir_conv.  The solver stores
// its table with 32-byte aligned rows starting at tap 1, so a lane takes four consecutive taps (one sector, two
// 16-byte loads) against twelve window values.  Window: logical[q] = W(k0 + q - B), W(m) = h_i[s0 - 1 - m] (zero
// outside 0 <= m < s0: lags that reach into the block or before the first sample drop out by themselves).  A lane
// reads its twelve window values as six 16-byte chunks at a lane stride of 32 bytes; adjacent chunks are swapped in
// every other 128-byte row so that the eight lanes of a quarter warp hit all 32 banks whatever the chunk offset.
__device__ __forceinline__ uint fw_phys(uint m) {
  const uint c = m >> 1;
  return ((c ^ ((c >> 3) & 1u)) << 1) | (m & 1u);
}
__device__ __forceinline__ double2 fw_lds128(const double *p) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
  return v;
}

template <int B>
__global__ void __launch_bounds__(FC_THREADS, 3)
k_fir_far(uint P, size_t ld, uint per, uint K, uint s0, const double *__restrict__ tab, const double *__restrict__ h, size_t hs,
          double *part) {
  constexpr uint WLEN = FC_KTILE + B + 8;
  __shared__ __align__(16) double wbuf[(WLEN + 15) & ~15u];
  asm volatile("griddepcontrol.launch_dependents;");
  const uint i = blockIdx.x, g = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint j0 = g * per, j1 = min(P, j0 + per);
  const double *hp = h + (size_t)i * hs;
  const size_t PP = (size_t)P * P;
  for (uint k0 = 0; k0 < K; k0 += FC_KTILE) {           // taps tau = 1 + k0 + t, t < kt
    const uint kt = min((uint)FC_KTILE, K - k0), nq = kt >> 2;
    __syncthreads();
    for (uint q = threadIdx.x; q < kt + B + 4; q += FC_THREADS) {
      const long long m = (long long)k0 + (long long)q - B;
      wbuf[fw_phys(q)] = (m >= 0 && m < (long long)s0) ? hp[s0 - 1 - m] : 0.0;
    }
    __syncthreads();
    for (uint j = j0 + warp; j < j1; j += FC_THREADS / 32) {
      const double *f = tab + ((size_t)i * P + j) * ld + 1 + k0;      // 32-byte aligned (ld % 4 == 0, tap 1 on a boundary)
      double acc[B];
#pragma unroll
      for (int b = 0; b < B; b++) acc[b] = 0.0;
      const double2 *f2 = reinterpret_cast<const double2 *>(f);
      // three groups of four taps per lane in flight: a buffer is refilled right after it has been used and is
      // needed again two groups later (rotation by unrolling, no register copies that would wait for the load)
      constexpr int NBUF = 3;
      double2 bu[NBUF][2];
#pragma unroll
      for (int k = 0; k < NBUF; k++)
        if (lane + 32 * k < nq) { bu[k][0] = __ldcs(f2 + 2 * (lane + 32 * k)); bu[k][1] = __ldcs(f2 + 2 * (lane + 32 * k) + 1); }
      for (uint qb = lane; qb < nq; qb += 32 * NBUF) {
#pragma unroll
        for (int k = 0; k < NBUF; k++) {
          const uint q = qb + 32 * k;
          if (q < nq) {
            double wv[12];                              // logical[4 q + x]
#pragma unroll
            for (int x = 0; x < 12; x += 2) {
              const double2 w2 = fw_lds128(wbuf + fw_phys(4 * q + x));
              wv[x] = w2.x;
              wv[x + 1] = w2.y;
            }
            const double2 u0 = bu[k][0], u1 = bu[k][1];
#pragma unroll
            for (int b = 0; b < B; b++) {               // tap 4 q + u, lag b -> logical[4 q + u - b + B]
              acc[b] = fma(u0.x, wv[B - b], acc[b]);
              acc[b] = fma(u0.y, wv[B + 1 - b], acc[b]);
              acc[b] = fma(u1.x, wv[B + 2 - b], acc[b]);
              acc[b] = fma(u1.y, wv[B + 3 - b], acc[b]);
            }
            if (q + 32 * NBUF < nq) {
              bu[k][0] = __ldcs(f2 + 2 * (q + 32 * NBUF));
              bu[k][1] = __ldcs(f2 + 2 * (q + 32 * NBUF) + 1);
            }
          }
        }
      }
      const uint t = 4 * nq + lane;                     // up to three taps after the last full group
      if (t < kt) {
        const double fv = f[t];
#pragma unroll
        for (int b = 0; b < B; b++) acc[b] = fma(fv, wbuf[fw_phys(t - b + B)], acc[b]);
      }
#pragma unroll
      for (int b = 0; b < B; b++) {
        double a = acc[b];
        for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
        if (lane == 0) {
          double *dst = part + (size_t)b * PP + (size_t)i * P + j;
          *dst = k0 ? *dst + a : a;
        }
      }
    }
  }
}

// out[j] = sum_i M[i * P + j] * (W ? w[i] : 1): the sources are dealt to G thread groups, the group sums are added
// in group order (fixed association, independent of timing)
// The same history pass on the FP64 tensor pipe.  For one source i the pass is a product of the table rows with a
// Toeplitz matrix of the window: part[n][i][j] = sum_t fir[i][j][1 + t] * W(t - n), an (8 destinations) x (taps) x
// (8 lags) GEMM per row tile -- DMMA.8x8x4 with A = table entries straight from global memory (lane = (row, four taps):
// two 16-byte loads of one sector give the A operands of four DMMAs), B = window values from shared memory (one
// wavefront: the 32 lanes read 15 consecutive doubles), C = the 8 x 8 block of `part` (no shuffles, no reduction
// over lanes).  Per KB of table: 2 LDG.128 + 4 NB LDS.64 + 4 NB DMMA instead of 2 + 12 + 64 DFMA per 128 taps -- the
// L1TEX unit no longer limits the pass, and 16 lags (NB = 2) cost the same table traffic as 8.  The eight warps of a
// CTA split the taps of a row tile; their C blocks are added in warp order.
// Walking eight table rows at once, 128 bytes at a time in each, thrashes the DRAM pages (3.2 TB/s measured on the row
// layout), so the pass reads a second copy of the table interleaved by row tiles (k_td_interleave): a warp request is
// 1 KB contiguous.  Measured on the way: taps in the M dimension instead (one row per warp, n' = lag - m as columns)
// needs no copy but NB + 1 column blocks, 184 us at 16 lags against 152 us here.
__device__ __forceinline__ void fw_dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// tab8[((i * tiles + jt) * ng4 + g4) * 32 + r * 4 + u] = fir[i][8 jt + r][1 + 4 g4 + u]  (zero outside the table)
__global__ void k_td_interleave(uint P, uint nfir, size_t ld, size_t ng4, const double *__restrict__ tab, double *tab8) {
  const size_t tiles = (P + 7) >> 3, total = (size_t)P * tiles * ng4 * 32;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const uint u = e & 3, r = (e >> 2) & 7;
    const size_t g4 = (e >> 5) % ng4, it = (e >> 5) / ng4, jt = it % tiles, i = it / tiles;
    const size_t j = 8 * jt + r, tau = 1 + 4 * g4 + u;
    tab8[e] = (j < P && tau < nfir) ? tab[(i * P + j) * ld + tau] : 0.0;
  }
}

template <int NB>
__global__ void __launch_bounds__(FC_THREADS)
k_fir_far_mma(uint P, size_t ng4, uint per, uint K, uint s0, uint lag0, const double *__restrict__ tab8, const double *__restrict__ h,
              size_t hs, double *part) {
  // lag0 > 0: the pass runs ahead -- s0 is the first sample of the block BEFORE the one it serves, the lags are
  // lag0 + n (rows >= s0 are not known yet; the step kernels add them)
  constexpr uint B = 8 * NB, NW = FC_THREADS / 32;
  __shared__ __align__(16) double wbuf[FC_KTILE + B + 32];    // logical[q] = W(k0 + q - B - lag0)
  __shared__ double red[NW][NB][64];
  asm volatile("griddepcontrol.launch_dependents;");
  const uint i = blockIdx.x, g = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint j0 = g * per, j1 = min(P, j0 + per), r = lane >> 2, kq = lane & 3;
  const double *hp = h + (size_t)i * hs;
  const size_t PP = (size_t)P * P;
  for (uint k0 = 0; k0 < K; k0 += FC_KTILE) {           // taps tau = 1 + k0 + t, t < kt
    // a chunk = 16 taps: lane (row r, kq) takes the four taps 16 ch + 4 kq .. + 3 of its row (one 32-byte sector, the
    // warp's request covers a 128-byte line in each of its eight rows) = the A operands of four DMMAs
    const uint kt = min((uint)FC_KTILE, K - k0), nchunk = (kt + 15) >> 4, cpw = (nchunk + NW - 1) / NW;
    __syncthreads();
    for (uint q = threadIdx.x; q < 16 * nchunk + B + 1; q += FC_THREADS) {
      const long long m = (long long)k0 + (long long)q - B - lag0;
      wbuf[q] = (m >= 0 && m < (long long)s0) ? hp[s0 - 1 - m] : 0.0;
    }
    __syncthreads();
    const uint c0 = min(nchunk, warp * cpw), c1 = min(nchunk, c0 + cpw);
    for (uint jt = j0; jt < j1; jt += 8) {
      // interleaved copy: [source][row tile][group of 4 taps][row 0..7][4 taps] -- the warp's request is 1 KB contiguous
      const double *f = tab8 + ((((size_t)i * ((P + 7) >> 3) + (jt >> 3)) * ng4 + (k0 >> 2) + kq) * 8 + r) * 4;
      double acc[NB][2];
#pragma unroll
      for (int nb = 0; nb < NB; nb++) acc[nb][0] = acc[nb][1] = 0.0;
      auto load = [&](uint ch, double (&v)[4]) {        // zero past the end of the tile
        const uint t = 16 * ch + 4 * kq;
        if (t + 3 < kt) {
          const double2 lo = __ldcs(reinterpret_cast<const double2 *>(f + 128 * ch)), hi = __ldcs(reinterpret_cast<const double2 *>(f + 128 * ch) + 1);
          v[0] = lo.x; v[1] = lo.y; v[2] = hi.x; v[3] = hi.y;
        } else {
#pragma unroll
          for (int u = 0; u < 4; u++) v[u] = t + u < kt ? f[128 * ch + u] : 0.0;
        }
      };
      auto mma = [&](uint ch, const double (&v)[4]) {
        const double *w = wbuf + 16 * ch + 4 * kq + B - r;         // lag n = 8 nb + r: logical[t - n + B]
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
          for (int nb = 0; nb < NB; nb++) fw_dmma(acc[nb][0], acc[nb][1], v[u], w[u - 8 * nb]);
      };
      // two pairs of chunks in rotation: a pair is reloaded right after its DMMAs and used again one pair later, so four
      // to eight 16-byte loads per lane are in flight under the DMMAs of the other pair (no register copies)
      double va[4], vb[4], vc[4], vd[4];
      if (c0 < c1) load(c0, va);
      if (c0 + 1 < c1) load(c0 + 1, vb);
      for (uint ch = c0; ch < c1; ch += 4) {
        if (ch + 2 < c1) load(ch + 2, vc);
        if (ch + 3 < c1) load(ch + 3, vd);
        mma(ch, va);
        if (ch + 1 < c1) mma(ch + 1, vb);
        if (ch + 4 < c1) load(ch + 4, va);
        if (ch + 5 < c1) load(ch + 5, vb);
        if (ch + 2 < c1) mma(ch + 2, vc);
        if (ch + 3 < c1) mma(ch + 3, vd);
      }
#pragma unroll
      for (int nb = 0; nb < NB; nb++) {
        red[warp][nb][2 * lane] = acc[nb][0];
        red[warp][nb][2 * lane + 1] = acc[nb][1];
      }
      __syncthreads();
      // C fragment element e of lane l: row l / 4, lag 8 nb + 2 (l % 4) + e; the warps' blocks are added in warp order
      for (uint e = threadIdx.x; e < NB * 64; e += FC_THREADS) {
        const uint nb = e >> 6, x = e & 63, l = x >> 1, row = jt + (l >> 2), n = 8 * nb + 2 * (l & 3) + (x & 1);
        double sum = 0.0;
#pragma unroll
        for (uint w = 0; w < NW; w++) sum += red[w][nb][x];
        if (row < j1) {
          double *dst = part + (size_t)n * PP + (size_t)i * P + row;
          *dst = k0 ? *dst + sum : sum;
        }
      }
      __syncthreads();
    }
  }
}

// CG: M is global memory another kernel wrote -- read it through L2 (ld.global.cg); otherwise M is in shared memory.
// w (W only) is always in shared memory.  Pointer increments and 32-bit shared addresses: this loop is most of what
// the single CTA of the step kernel executes.
__device__ __forceinline__ double td_lds(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
template <bool W, bool CG = false>
__device__ __forceinline__ void td_colsum(uint nsrc, uint P, const double *M, const double *w, double *gsum, double *out) {
  const uint T = TD_THREADS, Pp = min((P + 31u) & ~31u, T), G = T / Pp, g = threadIdx.x / Pp, jj = threadIdx.x - g * Pp;
  const uint chunk = (nsrc + G - 1) / G;
  if (g < G) {
    const uint i0 = min(nsrc, g * chunk), i1 = min(nsrc, i0 + chunk), n = i1 - i0;
    const unsigned wb = W ? (unsigned)__cvta_generic_to_shared(w) + i0 * 8u : 0u;
    for (uint j = jj; j < P; j += Pp) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      const double *pg = M + (size_t)i0 * P + j;      // CG
      unsigned ps = CG ? 0u : (unsigned)__cvta_generic_to_shared(M) + (i0 * P + j) * 8u, pw = wb;
      const unsigned st = P * 8u;
      uint i = 0;
      for (; i + 3 < n; i += 4) {
        double m0, m1, m2, m3;
        if (CG) {
          m0 = __ldcg(pg); m1 = __ldcg(pg + P); m2 = __ldcg(pg + 2 * (size_t)P); m3 = __ldcg(pg + 3 * (size_t)P);
          pg += 4 * (size_t)P;
        } else {
          m0 = td_lds(ps); m1 = td_lds(ps + st); m2 = td_lds(ps + 2 * st); m3 = td_lds(ps + 3 * st);
          ps += 4 * st;
        }
        if (W) {
          a0 = fma(m0, td_lds(pw), a0);
          a1 = fma(m1, td_lds(pw + 8), a1);
          a2 = fma(m2, td_lds(pw + 16), a2);
          a3 = fma(m3, td_lds(pw + 24), a3);
          pw += 32;
        } else {
          a0 += m0; a1 += m1; a2 += m2; a3 += m3;
        }
      }
      for (; i < n; i++) {
        double m0;
        if (CG) { m0 = __ldcg(pg); pg += P; } else { m0 = td_lds(ps); ps += st; }
        if (W) { a0 = fma(m0, td_lds(pw), a0); pw += 8; } else a0 += m0;
      }
      gsum[(size_t)g * P + j] = (a0 + a1) + (a2 + a3);
    }
  }
  __syncthreads();
  for (uint j = threadIdx.x; j < P; j += T) {
    double s = 0.0;
    for (uint q = 0; q < G; q++) s += gsum[q * P + j];
    out[j] = s;
  }
  __syncthreads();
}

// The step kernel is ONE CTA executing a long, mostly serial program; every helper exists once in the code
// (noinline functions, loops that are not unrolled): with the reference's control flow written out call by call the
// kernel had 26 000 instructions, each executed once or twice, and ran at the speed of its instruction fetches.
struct TdSm {
  double *base, *x, *xp, *u, *pes, *spr, *dmp, *pt, *v, *ge, *tmp, *wn, *red, *gsum;
  const double *fir0;
  bool fir0_shared;
};

__device__ __forceinline__ void td_far_sum(const TdArgs &a, const TdSm &m) {       // sum over the sources of the history pass
  td_colsum<false, true>(a.P, a.P, a.part, nullptr, m.gsum, m.tmp);
}
// tmp += sum over lags lo .. hi of fir[:, :, lag]^T p_tot[s - lag] (lags and sources flattened into one column sum)
__device__ __forceinline__ void td_near_sum(const TdArgs &a, const TdSm &m, uint lo, uint hi) {
  const uint P = a.P, T = TD_THREADS, tid = threadIdx.x, n = (hi - lo + 1) * P;
  const double *pt_rows = a.state + (size_t)TD_PTOT * a.nt * P;
  for (uint r = 0; r <= hi - lo; r++)
    for (uint j = tid; j < P; j += T) m.wn[r * P + j] = __ldcg(pt_rows + (size_t)(a.s - lo - r) * P + j);
  __syncthreads();
  td_colsum<true, true>(n, P, a.firn + (size_t)(lo - 1) * P * P, m.wn, m.gsum, m.base);
  for (uint j = tid; j < P; j += T) m.tmp[j] += m.base[j];
  __syncthreads();
}
__device__ __noinline__ double td_pow_general(double a, double n) { return pow(a, n); }
__device__ __forceinline__ double td_pow(double a, double n) { return n == 1.0 ? a : n == 2.0 ? a * a : td_pow_general(a, n); }

// the velocity of _update_all (:415) enters nothing inside the fixed point: it is evaluated where it is read, by the
// damper update and for the stored row
__device__ __forceinline__ void td_update_u(const TdArgs &a, const TdSm &m) {
  for (uint j = threadIdx.x; j < a.P; j += TD_THREADS) m.u[j] = a.s ? (m.x[j] - m.xp[j]) / a.dt : 0.0;
  __syncthreads();
}
__device__ __forceinline__ void td_update_all(const TdArgs &a, const TdSm &m) {    // _update_all, simulation.pyx:412-421 (without u)
  const uint P = a.P, T = TD_THREADS, tid = threadIdx.x;
  const bool first_free = m.x[0] >= a.x0;
  for (uint j = tid; j < P; j += T) {
    const double xj = m.x[j], d = xj + m.ge[j];
    m.pes[j] = (a.e0h * (m.v[j] * m.v[j])) / (d * d);
    double c = xj >= a.x0 ? 0.0 : a.k * td_pow(a.x0 - xj, a.n);
    if (first_free) c = trunc(c);
    m.spr[j] = c;
    m.pt[j] = (c + m.dmp[j]) + m.pes[j];
  }
  __syncthreads();
}
__device__ __forceinline__ void td_update_dmp(const TdArgs &a, const TdSm &m) {    // _update_cont_dmp, :423-428
  const uint P = a.P, T = TD_THREADS, tid = threadIdx.x;
  const bool first_free = m.x[0] >= a.x0;
  for (uint j = tid; j < P; j += T) {
    double c = m.x[j] >= a.x0 ? 0.0 : -(a.lmbd * td_pow(a.x0 - m.x[j], a.n)) * m.u[j];
    if (first_free) c = trunc(c);
    m.dmp[j] = c;
  }
  __syncthreads();
}
// second half of _update_x (:436-443): x <- base + dt * tmp, returns max|dx|
__device__ __forceinline__ double td_finish_x(const TdArgs &a, const TdSm &m) {
  const uint P = a.P, T = TD_THREADS, tid = threadIdx.x;
  double e = 0.0;
  for (uint j = tid; j < P; j += T) {
    const double xn = __dadd_rn(m.base[j], __dmul_rn(m.tmp[j], a.dt));
    const double dj = fabs(m.x[j] - xn);
    e = (dj > e || dj != dj) ? dj : e;                // NaN sticks, like np.max
    m.x[j] = xn;
  }
  for (int off = 16; off > 0; off >>= 1) {
    const double o = __shfl_xor_sync(0xffffffffu, e, off);
    e = (o > e || o != o) ? o : e;
  }
  if ((tid & 31) == 0) m.red[tid >> 5] = e;
  __syncthreads();
  e = (tid & 31) < T / 32 ? m.red[tid & 31] : 0.0;    // every warp reduces the 32 warp maxima again
  for (int off = 16; off > 0; off >>= 1) {
    const double o = __shfl_xor_sync(0xffffffffu, e, off);
    e = (o > e || o != o) ? o : e;
  }
  __syncthreads();
  return e;
}
__device__ __forceinline__ double td_update_x(const TdArgs &a, const TdSm &m) {    // _update_x, :430-445
  if (m.fir0_shared) td_colsum<true, false>(a.P, a.P, m.fir0, m.pt, m.gsum, m.tmp);
  else td_colsum<true, true>(a.P, a.P, m.fir0, m.pt, m.gsum, m.tmp);
  return td_finish_x(a, m);
}

__global__ void __launch_bounds__(TD_THREADS) k_td_step(TdArgs a) {
  extern __shared__ double sm[];
  const uint P = a.P, T = TD_THREADS, tid = threadIdx.x, s = a.s;
  TdSm m;
  m.base = sm; m.x = m.base + P; m.xp = m.x + P; m.u = m.xp + P; m.pes = m.u + P; m.spr = m.pes + P; m.dmp = m.spr + P;
  m.pt = m.dmp + P; m.v = m.pt + P; m.ge = m.v + P; m.tmp = m.ge + P; m.wn = m.tmp + P; m.red = m.wn + (size_t)TD_NL * P;
  m.gsum = m.red + 32;
  m.fir0 = a.fir0;
  m.fir0_shared = a.stage != 0;
  double *S = a.state;
  const size_t plane = (size_t)a.nt * P, PP = (size_t)P * P;
  td_mark(a, 0);
  // Prologue, overlapped with the previous kernel of the stream (programmatic dependent launch): the drive, the
  // gaps and the zero-lag slice of the table are constants of the run.  In a block of samples a kernel releases
  // its dependent only after its own dependency is met, so when sample s starts, everything up to sample s - 2 and
  // the history pass of the block are complete: their part of the convolution (all of it except lag 1) is summed
  // before the wait as well, through L2.
  if (a.stage) {
    const uint Pp = min((P + 31u) & ~31u, T), G = T / Pp;
    double *f0s = sm + (((size_t)(11 + TD_NL) * P + 32 + (size_t)G * P + 1) & ~(size_t)1);   // 16-byte aligned
    const size_t n2 = PP / 2;
    const double2 *src = reinterpret_cast<const double2 *>(a.fir0);
    double2 *dst = reinterpret_cast<double2 *>(f0s);
    for (size_t e = tid; e < n2; e += T) dst[e] = src[e];
    if (tid == 0 && (P & 1)) f0s[PP - 1] = a.fir0[PP - 1];
    m.fir0 = f0s;
  }
  for (uint j = tid; j < P; j += T) {
    m.v[j] = a.v[(size_t)s * P + j];
    m.ge[j] = a.geff[j];
    m.dmp[j] = 0.0;
    m.x[j] = 0.0;
    m.tmp[j] = 0.0;
  }
  __syncthreads();
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {              // 0: before the dependency is met, 1: after
    if (pass) {
      td_mark(a, 1);
      asm volatile("griddepcontrol.wait;" ::: "memory");   // the previous kernel of the stream is complete
      asm volatile("griddepcontrol.launch_dependents;");   // the next sample's kernel may run its prologue now
      td_mark(a, 2);
      for (uint j = tid; j < P; j += T) m.xp[j] = s ? S[TD_X * plane + (size_t)(s - 1) * P + j] : 0.0;
    }
    if (s && (pass == 0) == (a.early != 0)) td_far_sum(a, m);
    // lags into the block's own samples: 2 .. near early, lag 1 (or all of them) once the previous sample is final
    const uint lo = pass ? 1u : 2u, hi = pass ? (a.early ? min(a.near, 1u) : a.near) : (a.early ? a.near : 0u);
    if (s && hi >= lo) td_near_sum(a, m, lo, hi);
  }
  if (s == 0) {                                       // initial state (:335)
    td_update_all(a, m);
  } else {
    for (uint j = tid; j < P; j += T) m.x[j] = m.base[j] = m.tmp[j] * a.dt;      // _blind_x (:447-457)
    __syncthreads();
    td_mark(a, 3);
    // step (:459-503) as one loop: [damper update] -> _update_all -> decide -> _update_x.  Phase 0 iterates without
    // the contact damper, phase 1 with the damper frozen at phase 0's estimate; the first _update_x of a phase is
    // unconditional, up to maxiter - 1 more follow while max|dx| > atol; `it` counts like the reference's `i`
    int phase = 0, q = 0, it = 2;
    bool need_dmp = false;
    double err = 0.0;
#pragma unroll 1
    for (;;) {
      if (need_dmp) {
        td_update_u(a, m);
        td_update_dmp(a, m);
        need_dmp = false;
      }
      td_update_all(a, m);
      if (q == 0 || (q < a.maxiter && !(err <= a.atol))) {
        err = td_update_x(a, m);
        if (q > 0) it++;
        q++;
        continue;
      }
      if (phase == 0) {
        td_mark(a, 4);
        phase = 1;
        q = 0;
        need_dmp = true;
        continue;
      }
      break;
    }
    if (tid == 0) {
      a.err[s] = err;
      a.iters[s] = it;
    }
    td_mark(a, 5);
  }
  td_update_u(a, m);
  for (uint j = tid; j < P; j += T) {
    const size_t r = (size_t)s * P + j;
    S[TD_X * plane + r] = m.x[j];
    S[TD_U * plane + r] = m.u[j];
    S[TD_PES * plane + r] = m.pes[j];
    S[TD_SPR * plane + r] = m.spr[j];
    S[TD_DMP * plane + r] = m.dmp[j];
    S[TD_PTOT * plane + r] = m.pt[j];
    a.hist[(size_t)j * a.nt + s] = m.pt[j];           // the next steps' convolution input
  }
  td_mark(a, 6);
}

// ---- the step kernel on a cluster of TD_CL CTAs ----------------------------------------------------------------
// Same program, but the products over the sources are dealt to the CTAs of a thread-block cluster: CTA c owns a slice
// of the sources, keeps its rows of fir[:, :, 0] in REGISTERS for the whole kernel (6 doubles per thread at P = 144:
// the single-CTA kernel re-reads 166 KB of shared memory per fixed-point iteration and is bound by that), sums its
// slice, writes the partial sums into every CTA's shared memory (distributed shared memory) and, after one cluster
// barrier, every CTA adds the TD_CL partial vectors in rank order.  Everything else -- electrostatic and contact
// pressures, the error norm, the control flow -- is computed redundantly by every CTA from identical inputs with
// identical instructions, so all CTAs take the same branches and meet at the same barriers.  Rank 0 writes the rows.
#ifndef CNLD_TD_CL
#define CNLD_TD_CL 4                      // step kernel to step kernel at P = 144: 2 CTAs 24.7 us, 4 CTAs 19.7 us, 8 CTAs 20.3 us
#endif
constexpr int TD_CL = CNLD_TD_CL;         // cluster size
constexpr int TD_MR = 32 / CNLD_TD_CL;    // fir[:, :, 0] rows per thread kept in registers (at most)
namespace cg = cooperative_groups;

struct TdCl {
  double *gsum, *gpart;                   // [G][P] group sums of this CTA; [2][TD_CL][P] partial vectors of all CTAs
  uint rank, buf;
};

// out[j] = sum over ALL sources i < nsrc of M[i * P + j] * (W ? w[i] : 1); REG: this thread's entries of M are in mreg
template <bool W, bool REG>
__device__ __forceinline__ void td_cl_colsum(cg::cluster_group &cl, TdCl &c, uint nsrc, uint P, const double *M, const double *w,
                                             const double (&mreg)[TD_MR], double *out) {
  const uint T = TD_THREADS, Pp = min((P + 31u) & ~31u, T), G = T / Pp, g = threadIdx.x / Pp, jj = threadIdx.x - g * Pp;
  const uint per = (nsrc + TD_CL - 1) / TD_CL, lo = min(nsrc, c.rank * per), nc = min(nsrc, lo + per) - lo;
  const uint chunk = (nc + G - 1) / G;
  if (g < G) {
    const uint o0 = min(nc, g * chunk), n = min(nc, o0 + chunk) - o0, i0 = lo + o0;
    const unsigned wb = W ? (unsigned)__cvta_generic_to_shared(w) + i0 * 8u : 0u;
    for (uint j = jj; j < P; j += Pp) {
      double a0 = 0.0, a1 = 0.0;
      if (REG) {
#pragma unroll
        for (int q = 0; q < TD_MR; q += 2) {
          if (q < (int)n) a0 = fma(mreg[q], td_lds(wb + 8 * q), a0);
          if (q + 1 < (int)n) a1 = fma(mreg[q + 1], td_lds(wb + 8 * q + 8), a1);
        }
      } else {
        const double *pg = M + (size_t)i0 * P + j;
        unsigned pw = wb;
        uint i = 0;
        for (; i + 1 < n; i += 2) {
          const double m0 = __ldcg(pg), m1 = __ldcg(pg + P);
          pg += 2 * (size_t)P;
          if (W) {
            a0 = fma(m0, td_lds(pw), a0);
            a1 = fma(m1, td_lds(pw + 8), a1);
            pw += 16;
          } else {
            a0 += m0;
            a1 += m1;
          }
        }
        if (i < n) a0 = W ? fma(__ldcg(pg), td_lds(pw), a0) : a0 + __ldcg(pg);
      }
      c.gsum[(size_t)g * P + j] = a0 + a1;
    }
  }
  __syncthreads();
  double *mine = c.gpart + ((size_t)c.buf * TD_CL + c.rank) * P;
  for (uint j = threadIdx.x; j < P; j += T) {
    double sum = 0.0;
    for (uint q = 0; q < G; q++) sum += c.gsum[q * P + j];
#pragma unroll
    for (uint r = 0; r < TD_CL; r++) cl.map_shared_rank(mine, r)[j] = sum;
  }
  cl.sync();
  const double *all = c.gpart + (size_t)c.buf * TD_CL * P;
  for (uint j = threadIdx.x; j < P; j += T) {
    double sum = 0.0;
#pragma unroll
    for (uint r = 0; r < TD_CL; r++) sum += all[r * P + j];
    out[j] = sum;
  }
  c.buf ^= 1u;          // the next product writes the other buffer: a CTA one barrier ahead cannot overwrite what a slower one still reads
  __syncthreads();
}

__global__ void __cluster_dims__(TD_CL, 1, 1) __launch_bounds__(TD_THREADS) k_td_step_cl(TdArgs a) {
  extern __shared__ double sm[];
  cg::cluster_group cl = cg::this_cluster();
  const uint P = a.P, T = TD_THREADS, tid = threadIdx.x, s = a.s;
  const uint Pp = min((P + 31u) & ~31u, T), G = T / Pp;
  TdSm m;
  m.base = sm; m.x = m.base + P; m.xp = m.x + P; m.u = m.xp + P; m.pes = m.u + P; m.spr = m.pes + P; m.dmp = m.spr + P;
  m.pt = m.dmp + P; m.v = m.pt + P; m.ge = m.v + P; m.tmp = m.ge + P; m.wn = m.tmp + P; m.red = m.wn + (size_t)TD_NL * P;
  m.gsum = m.red + 32;
  m.fir0 = a.fir0;
  m.fir0_shared = false;
  TdCl c;
  c.gsum = m.gsum;
  c.gpart = m.gsum + (size_t)G * P;
  c.rank = cl.block_rank();
  c.buf = 0;
  double *S = a.state;
  const size_t plane = (size_t)a.nt * P, PP = (size_t)P * P;
  const bool lead = c.rank == 0;
  if (lead) td_mark(a, 0);
  // this thread's entries of fir[:, :, 0]: sources of this CTA's slice, group g, destination jj (constants of the run)
  double mreg[TD_MR];
  {
    const uint g = tid / Pp, jj = tid - g * Pp, per = (P + TD_CL - 1) / TD_CL, lo = min(P, c.rank * per), nc = min(P, lo + per) - lo;
    const uint chunk = (nc + G - 1) / G, o0 = min(nc, g * chunk), n = g < G ? min(nc, o0 + chunk) - o0 : 0u;
#pragma unroll
    for (int q = 0; q < TD_MR; q++) mreg[q] = (q < (int)n && jj < P) ? a.fir0[(size_t)(lo + o0 + q) * P + jj] : 0.0;
  }
  const bool reg = a.stage != 0;          // host: P <= TD_THREADS and at most TD_MR rows per thread
  for (uint j = tid; j < P; j += T) {
    m.v[j] = a.v[(size_t)s * P + j];
    m.ge[j] = a.geff[j];
    m.dmp[j] = 0.0;
    m.x[j] = 0.0;
    m.tmp[j] = 0.0;
  }
  __syncthreads();
  cl.sync();                               // every CTA of the cluster is resident before the first remote store
  const double *pt_rows = S + (size_t)TD_PTOT * plane;
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {              // 0: before the dependency is met, 1: after (see k_td_step)
    if (pass) {
      if (lead) td_mark(a, 1);
      asm volatile("griddepcontrol.wait;" ::: "memory");
      asm volatile("griddepcontrol.launch_dependents;");
      if (lead) td_mark(a, 2);
      for (uint j = tid; j < P; j += T) m.xp[j] = s ? S[TD_X * plane + (size_t)(s - 1) * P + j] : 0.0;
    }
    if (s && (pass == 0) == (a.early != 0)) td_cl_colsum<false, false>(cl, c, P, P, a.part, nullptr, mreg, m.tmp);
    const uint lo = pass ? 1u : 2u, hi = pass ? (a.early ? min(a.near, 1u) : a.near) : (a.early ? a.near : 0u);
    if (s && hi >= lo) {
      for (uint r = 0; r <= hi - lo; r++)
        for (uint j = tid; j < P; j += T) m.wn[r * P + j] = __ldcg(pt_rows + (size_t)(s - lo - r) * P + j);
      __syncthreads();
      td_cl_colsum<true, false>(cl, c, (hi - lo + 1) * P, P, a.firn + (size_t)(lo - 1) * PP, m.wn, mreg, m.base);
      for (uint j = tid; j < P; j += T) m.tmp[j] += m.base[j];
      __syncthreads();
    }
  }
  if (s == 0) {
    td_update_all(a, m);
  } else {
    for (uint j = tid; j < P; j += T) m.x[j] = m.base[j] = m.tmp[j] * a.dt;
    __syncthreads();
    if (lead) td_mark(a, 3);
    int phase = 0, q = 0, it = 2;
    bool need_dmp = false;
    double err = 0.0;
#pragma unroll 1
    for (;;) {
      if (need_dmp) {
        td_update_u(a, m);
        td_update_dmp(a, m);
        need_dmp = false;
      }
      td_update_all(a, m);
      if (q == 0 || (q < a.maxiter && !(err <= a.atol))) {
        if (reg) td_cl_colsum<true, true>(cl, c, P, P, a.fir0, m.pt, mreg, m.tmp);
        else td_cl_colsum<true, false>(cl, c, P, P, a.fir0, m.pt, mreg, m.tmp);
        err = td_finish_x(a, m);
        if (q > 0) it++;
        q++;
        continue;
      }
      if (phase == 0) {
        if (lead) td_mark(a, 4);
        phase = 1;
        q = 0;
        need_dmp = true;
        continue;
      }
      break;
    }
    if (lead && tid == 0) {
      a.err[s] = err;
      a.iters[s] = it;
    }
    if (lead) td_mark(a, 5);
  }
  td_update_u(a, m);
  if (lead) {
    for (uint j = tid; j < P; j += T) {
      const size_t r = (size_t)s * P + j;
      S[TD_X * plane + r] = m.x[j];
      S[TD_U * plane + r] = m.u[j];
      S[TD_PES * plane + r] = m.pes[j];
      S[TD_SPR * plane + r] = m.spr[j];
      S[TD_DMP * plane + r] = m.dmp[j];
      S[TD_PTOT * plane + r] = m.pt[j];
      a.hist[(size_t)j * a.nt + s] = m.pt[j];
    }
    td_mark(a, 6);
  }
  cl.sync();                               // no CTA leaves while another may still address its shared memory
}

size_t td_smem_cl(uint P) {
  const uint Pp = std::min((P + 31u) & ~31u, (uint)TD_THREADS), G = TD_THREADS / Pp;
  return sizeof(double) * ((size_t)(11 + TD_NL) * P + 32 + (size_t)G * P + (size_t)2 * TD_CL * P);
}
// the register path of the cluster kernel: every destination has its own thread and a thread holds at most TD_MR rows
bool td_cl_reg(uint P) {
  if (P > (uint)TD_THREADS) return false;
  const uint Pp = std::min((P + 31u) & ~31u, (uint)TD_THREADS), G = TD_THREADS / Pp, nc = (P + TD_CL - 1) / TD_CL;
  return (nc + G - 1) / G <= (uint)TD_MR;
}

size_t td_smem_base(uint P) {
  const uint Pp = std::min((P + 31u) & ~31u, (uint)TD_THREADS), G = TD_THREADS / Pp;
  return sizeof(double) * (((size_t)(11 + TD_NL) * P + 32 + (size_t)G * P + 1) & ~(size_t)1);
}
bool td_stage(uint P) { return td_smem_base(P) + sizeof(double) * P * P <= 224 * 1024; }
size_t td_smem(uint P) { return td_smem_base(P) + (td_stage(P) ? sizeof(double) * P * P : 0); }

int td_launch_kernel(cnld_td *t, uint s, uint near, uint early, const double *part) {
  const int stage = t->cluster ? (td_cl_reg(t->P) ? 1 : 0) : (td_stage(t->P) ? 1 : 0);
  TdArgs a{t->P, t->nt, s, t->dt, -t->e0 / 2, t->k, t->n, t->x0, t->lmbd, t->atol, t->maxiter, stage, near,
           t->pdl ? early : 0u, t->d_firn, t->d_fir0, t->d_v, t->d_geff, part, t->d_hist, t->d_state, t->d_err, t->d_iters,
           t->d_trace};
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(t->cluster ? TD_CL : 1);
  cfg.blockDim = dim3(TD_THREADS);
  cfg.dynamicSmemBytes = t->cluster ? td_smem_cl(t->P) : td_smem(t->P);
  cfg.stream = t->st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = t->pdl ? 1 : 0;
  if (t->cluster) CNLD_CK(cudaLaunchKernelEx(&cfg, k_td_step_cl, a));
  else CNLD_CK(cudaLaunchKernelEx(&cfg, k_td_step, a));
  CNLD_CK_LAUNCH();
  return 0;
}

// one sample: one pass over the table (k_fir_conv, lags 1 .. min(s, taps - 1)), then the step kernel
int td_launch_step(cnld_td *t, uint s) {
  if (s) {
    const uint K = t->ntap > 1 ? std::min(s, t->ntap - 1) : 0u;
    if (K) {
      k_fir_conv<<<dim3(t->P, t->ngroup), FC_THREADS, 0, t->st>>>(t->P, t->P, (uint)t->ld, t->per, K, 1, t->tab, t->d_hist, t->nt,
                                                                   s - 1, t->d_part, 1, t->P);
      CNLD_CK_LAUNCH();
    } else {
      CNLD_CK(cudaMemsetAsync(t->d_part, 0, sizeof(double) * (size_t)t->P * t->P, t->st));
    }
  }
  return td_launch_kernel(t, s, 0, 0, t->d_part);
}

// What a run-ahead history pass could not know: the TD_B samples of the block before the one it serves.
//   part[n][i][j] += sum_{r < TD_B} fir[i][j][n + TD_B - r] * p_tot[s0 - TD_B + r][i],   n = 0 .. TD_B - 1
// from the compact copy of the first TD_NL lags (5 MB at C3, L2): CTA = source i, thread = destination j holds its
// 31 lags in registers.  Runs once per block between the previous block's last sample and this block's first.
__global__ void __launch_bounds__(256) k_fir_mid(uint P, uint s0, const double *__restrict__ firn, const double *__restrict__ pt_rows,
                                                 double *part) {
  __shared__ double pv[TD_B];
  asm volatile("griddepcontrol.launch_dependents;");
  const uint i = blockIdx.x;
  const size_t PP = (size_t)P * P;
  if (threadIdx.x < TD_B) pv[threadIdx.x] = pt_rows[(size_t)(s0 - TD_B + threadIdx.x) * P + i];
  __syncthreads();
  for (uint j = threadIdx.x; j < P; j += blockDim.x) {
    double f[TD_NL];
#pragma unroll
    for (int l = 0; l < TD_NL; l++) f[l] = __ldcg(firn + (size_t)l * PP + (size_t)i * P + j);     // lag l + 1
#pragma unroll
    for (int n = 0; n < TD_B; n++) {
      double acc = 0.0;
#pragma unroll
      for (int r = 0; r < TD_B; r++) acc = fma(f[n + TD_B - r - 1], pv[r], acc);
      part[(size_t)n * PP + (size_t)i * P + j] += acc;
    }
  }
}

// history pass for the nb <= TD_B samples of a block: into plane buffer `buf`, on stream st.  lag0 = 0: the history before
// s0 for the block starting at s0.  lag0 = TD_B (run-ahead): the history before s0 for the block starting at s0 + TD_B.
int td_launch_far(cnld_td *t, cudaStream_t st, uint s0, uint nb, uint lag0, uint buf, bool narrow) {
  const size_t PP = (size_t)t->P * t->P;
  double *part = t->d_part + (size_t)buf * TD_B * PP;
  const uint K = t->ntap > 1 ? std::min(s0 + lag0 + nb - 1, t->ntap - 1) : 0u;
  if (!K) {
    CNLD_CK(cudaMemsetAsync(part, 0, sizeof(double) * PP * TD_B, st));
    return 0;
  }
  // narrow: about two CTAs per SM, so that the step kernels of the running block find room next to the pass
  const uint ng = narrow ? t->ngroup_narrow : t->ngroup_far, per = narrow ? t->per_narrow : t->per_far;
  const dim3 grid(t->P, ng);
  if (!t->mma) k_fir_far<8><<<grid, FC_THREADS, 0, st>>>(t->P, t->ld, per, K, s0, t->tab, t->d_hist, t->nt, part);
  else if (nb <= 8) k_fir_far_mma<1><<<grid, FC_THREADS, 0, st>>>(t->P, t->ng4, per, K, s0, lag0, t->d_tab8, t->d_hist, t->nt, part);
  else k_fir_far_mma<2><<<grid, FC_THREADS, 0, st>>>(t->P, t->ng4, per, K, s0, lag0, t->d_tab8, t->d_hist, t->nt, part);
  CNLD_CK_LAUNCH();
  return 0;
}

// nb <= TD_B samples starting at s0 >= 1.  The history before the block is convolved once for all of them
// (td_launch_far), the step kernels add the lags that reach into the block itself from fir[:, :, 1 .. nb - 1].
// ahead: the pass for this block already ran under the previous block (stream st2, rows before the PREVIOUS block
// only, plane buffer t->pbuf); k_fir_mid adds the previous block's samples.
// prefetch: start the run-ahead pass for the next block now.
int td_launch_block(cnld_td *t, uint s0, uint nb, bool ahead, bool prefetch) {
  const size_t PP = (size_t)t->P * t->P;
  if (ahead) {
    CNLD_CK(cudaStreamWaitEvent(t->st, t->ev_far, 0));
    k_fir_mid<<<t->P, 256, 0, t->st>>>(t->P, s0, t->d_firn, t->d_state + (size_t)TD_PTOT * t->nt * t->P,
                                      t->d_part + (size_t)t->pbuf * TD_B * PP);
    CNLD_CK_LAUNCH();
  } else if (td_launch_far(t, t->st, s0, nb, 0, t->pbuf, false)) return -1;
  if (prefetch) {                       // everything before s0 is final once the work queued on st so far is done
    CNLD_CK(cudaEventRecord(t->ev_hist, t->st));
    CNLD_CK(cudaStreamWaitEvent(t->st2, t->ev_hist, 0));
    if (td_launch_far(t, t->st2, s0, TD_B, TD_B, t->pbuf ^ 1u, true)) return -1;
    CNLD_CK(cudaEventRecord(t->ev_far, t->st2));
  }
  const double *part = t->d_part + (size_t)t->pbuf * TD_B * PP;
  for (uint b = 0; b < nb; b++) {
    if (td_launch_kernel(t, s0 + b, t->ntap > 1 ? std::min(b, t->ntap - 1) : 0u, b >= 1, part + (size_t)b * PP)) return -1;
  }
  if (prefetch) t->pbuf ^= 1u;
  return 0;
}
}  // namespace

extern "C" {

cnld_td *cnld_b200_td_create(int device, cnld_uint npatch, cnld_uint nfir, const double *fir, cnld_uint nt, const double *v,
                             const double *gap_eff, double dt, double e0, double k, double n, double x0, double lmbd,
                             double atol, int maxiter) {
  if (cnld_b200_device_count() <= device) { set_error("no CUDA device %d (libcnld_b200 has no CPU fallback)", device); return nullptr; }
  if (!npatch || !nfir || nt < 2 || !fir || !v || !gap_eff || !(dt > 0) || maxiter < 1) { set_error("td_create: bad arguments"); return nullptr; }
  const size_t smem = td_smem(npatch);
  if (smem > 227 * 1024) { set_error("td_create: %u patches need %zu bytes of shared memory in the step kernel", npatch, smem); return nullptr; }
  cnld_td *t = new cnld_td();
  t->device = device; t->P = npatch; t->nfir = nfir; t->nt = nt;
  t->dt = dt; t->e0 = e0; t->k = k; t->n = n; t->x0 = x0; t->lmbd = lmbd; t->atol = atol; t->maxiter = maxiter;
  if (const char *e = getenv("CNLD_B200_TD_PDL")) t->pdl = atoi(e) != 0;
  if (const char *e = getenv("CNLD_B200_TD_CLUSTER")) t->cluster = atoi(e) != 0;
  if (const char *e = getenv("CNLD_B200_TD_OVERLAP")) t->overlap = atoi(e) != 0;
  if (td_smem_cl(npatch) > 200 * 1024) t->cluster = false;
  if (const char *e = getenv("CNLD_B200_TD_MMA")) t->mma = atoi(e) != 0;
  if (const char *e = getenv("CNLD_B200_TD_BLOCK")) t->block = (uint)std::min(std::max(atoi(e), 1), TD_B);
  if (t->mma) {                         // the DMMA pass reads a second, interleaved copy of the table: only if both fit
    size_t fr = 0, tot = 0;
    const size_t one = sizeof(double) * (size_t)npatch * (((size_t)npatch + 7) & ~(size_t)7) * ((size_t)nfir + 16);
    if (cudaSetDevice(device) == cudaSuccess && cudaMemGetInfo(&fr, &tot) == cudaSuccess && 2 * one + ((size_t)1 << 30) > fr) t->mma = 0;
  }
  if (!t->mma) t->block = std::min(t->block, 8u);
  t->ld = ((size_t)nfir + 3 + 3) & ~(size_t)3;
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
  t->ngroup = std::max(1u, std::min((npatch + 7) / 8, (uint)((4 * nsm + npatch - 1) / npatch)));
  t->per = (npatch + t->ngroup - 1) / t->ngroup;
  t->ngroup = (npatch + t->per - 1) / t->per;
  int occ_far = 4;
  if (t->mma) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_far, k_fir_far_mma<2>, FC_THREADS, 0);
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_far, k_fir_far<8>, FC_THREADS, 0);
  t->ngroup_far = std::max(1u, std::min((npatch + 7) / 8, (uint)(std::max(occ_far, 1) * nsm) / npatch));
  t->per_far = (npatch + t->ngroup_far - 1) / t->ngroup_far;
  if (t->mma == 1) t->per_far = (t->per_far + 7) & ~7u;          // whole tiles of 8 destinations per CTA
  t->ngroup_far = (npatch + t->per_far - 1) / t->per_far;
  t->ngroup_narrow = std::max(1u, std::min((npatch + 7) / 8, (uint)(2 * nsm) / npatch));
  t->per_narrow = ((npatch + t->ngroup_narrow - 1) / t->ngroup_narrow + 7) & ~7u;
  t->ngroup_narrow = (npatch + t->per_narrow - 1) / t->per_narrow;
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  const size_t PP = (size_t)npatch * npatch, tab = PP * t->ld + 4, rows = (size_t)nt * npatch;
  bool ok = cudaSetDevice(device) == cudaSuccess && cudaMalloc(&t->d_fir, sizeof(double) * tab) == cudaSuccess &&
            cudaMemset(t->d_fir, 0, sizeof(double) * tab) == cudaSuccess &&
            cudaMalloc(&t->d_fir0, sizeof(double) * PP) == cudaSuccess && cudaMalloc(&t->d_part, sizeof(double) * PP * TD_B * 2) == cudaSuccess &&
            cudaMalloc(&t->d_firn, sizeof(double) * PP * TD_NL) == cudaSuccess &&
            cudaMalloc(&t->d_v, sizeof(double) * rows) == cudaSuccess && cudaMalloc(&t->d_geff, sizeof(double) * npatch) == cudaSuccess &&
            cudaMalloc(&t->d_hist, sizeof(double) * rows) == cudaSuccess &&
            cudaMalloc(&t->d_state, sizeof(double) * rows * TD_NSTATE) == cudaSuccess &&
            cudaMalloc(&t->d_err, sizeof(double) * nt) == cudaSuccess && cudaMalloc(&t->d_iters, sizeof(int) * nt) == cudaSuccess &&
            (!getenv("CNLD_B200_TD_TRACE") || (cudaMalloc(&t->d_trace, sizeof(unsigned long long) * nt * 8) == cudaSuccess &&
                                                cudaMemset(t->d_trace, 0, sizeof(unsigned long long) * nt * 8) == cudaSuccess)) &&
            cudaMemcpy2D(t->d_fir + 3, sizeof(double) * t->ld, fir, sizeof(double) * nfir, sizeof(double) * nfir, PP,
                         cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemcpy(t->d_v, v, sizeof(double) * rows, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemcpy(t->d_geff, gap_eff, sizeof(double) * npatch, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaStreamCreateWithPriority(&t->st, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
            cudaStreamCreateWithPriority(&t->st2, cudaStreamNonBlocking, prio_lo) == cudaSuccess &&
            cudaEventCreateWithFlags(&t->ev_hist, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&t->ev_far, cudaEventDisableTiming) == cudaSuccess &&
            cudaFuncSetAttribute(k_td_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
            (!t->cluster ||
             cudaFuncSetAttribute(k_td_step_cl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)td_smem_cl(npatch)) == cudaSuccess);
  if (ok) {
    t->tab = t->d_fir + 3;
    k_td_gather_fir0<<<(unsigned)((PP + 255) / 256), 256, 0, t->st>>>(npatch, t->ld, t->tab, t->d_fir0);
    k_td_gather_near<<<(unsigned)((PP * TD_NL + 255) / 256), 256, 0, t->st>>>(npatch, nfir, t->ld, t->tab, t->d_firn);
    if (t->mma == 1) {
      t->ng4 = (((size_t)nfir + 3) / 4 + 3) & ~(size_t)3;       // whole chunks of 16 taps
      const size_t n8 = (size_t)npatch * ((npatch + 7) / 8) * t->ng4 * 32;
      ok = ok && cudaMalloc(&t->d_tab8, sizeof(double) * n8) == cudaSuccess;
      if (ok) k_td_interleave<<<(unsigned)std::min<size_t>((n8 + 255) / 256, 1u << 20), 256, 0, t->st>>>(npatch, nfir, t->ld, t->ng4, t->tab, t->d_tab8);
      count_launch();
    }
    uint *d_last = reinterpret_cast<uint *>(t->d_iters);        // scratch until reset() clears it
    ok = ok && cudaMemsetAsync(d_last, 0, sizeof(uint), t->st) == cudaSuccess;
    k_td_last_tap<<<(unsigned)((PP + 7) / 8), 256, 0, t->st>>>(PP, nfir, t->ld, t->tab, d_last);
    count_launch(3);
    ok = ok && cudaMemcpyAsync(&t->ntap, d_last, sizeof(uint), cudaMemcpyDeviceToHost, t->st) == cudaSuccess &&
         cudaStreamSynchronize(t->st) == cudaSuccess && cudaGetLastError() == cudaSuccess && cnld_b200_td_reset(t) == 0;
  }
  if (!ok) {
    set_error("td_create: %s", cudaGetErrorString(cudaGetLastError()));
    cnld_b200_td_destroy(t);
    return nullptr;
  }
  return t;
}

void cnld_b200_td_destroy(cnld_td *t) {
  if (!t) return;
  cudaSetDevice(t->device);
  cudaFree(t->d_fir); cudaFree(t->d_fir0); cudaFree(t->d_firn); cudaFree(t->d_tab8); cudaFree(t->d_part); cudaFree(t->d_v); cudaFree(t->d_geff); cudaFree(t->d_hist);
  cudaFree(t->d_state); cudaFree(t->d_err); cudaFree(t->d_iters); cudaFree(t->d_trace);
  if (t->st) cudaStreamDestroy(t->st);
  if (t->st2) cudaStreamDestroy(t->st2);
  if (t->ev_hist) cudaEventDestroy(t->ev_hist);
  if (t->ev_far) cudaEventDestroy(t->ev_far);
  delete t;
}

/* FixedStepSolver.reset (simulation.pyx:516-527): clear the state, initial `_update_all` at sample 0, current_step = 1 */
int cnld_b200_td_reset(cnld_td *t) {
  if (!t) { set_error("null td solver"); return -1; }
  CNLD_CK(cudaSetDevice(t->device));
  const size_t rows = (size_t)t->nt * t->P;
  CNLD_CK(cudaMemsetAsync(t->d_state, 0, sizeof(double) * rows * TD_NSTATE, t->st));
  CNLD_CK(cudaMemsetAsync(t->d_hist, 0, sizeof(double) * rows, t->st));
  CNLD_CK(cudaMemsetAsync(t->d_err, 0, sizeof(double) * t->nt, t->st));
  CNLD_CK(cudaMemsetAsync(t->d_iters, 0, sizeof(int) * t->nt, t->st));
  if (td_launch_step(t, 0)) return -1;
  CNLD_CK(cudaStreamSynchronize(t->st));
  t->cur = 1;
  t->pbuf = 0;
  return 0;
}

/* nsteps x FixedStepSolver.step (simulation.pyx:459-503), no host round trip between the steps */
int cnld_b200_td_step(cnld_td *t, cnld_uint nsteps) {
  if (!t) { set_error("null td solver"); return -1; }
  if ((size_t)t->cur + nsteps > t->nt) { set_error("td_step: %u steps from sample %u run past the %u samples of the time grid", nsteps, t->cur, t->nt); return -1; }
  CNLD_CK(cudaSetDevice(t->device));
  bool ahead = false;                   // the history pass of the next block to launch has run ahead
  for (uint left = nsteps; left;) {
    const uint nb = std::min(left, t->block);
    if (nb > 1) {
      // run ahead for the next block if this one is a full block of TD_B samples and a next block exists
      const bool prefetch = t->overlap && t->mma && t->block == (uint)TD_B && nb == (uint)TD_B && left - nb > 1;
      if (td_launch_block(t, t->cur, nb, ahead, prefetch)) return -1;
      ahead = prefetch;
    } else {
      if (td_launch_step(t, t->cur)) return -1;
      ahead = false;
    }
    t->cur += nb;
    left -= nb;
  }
  CNLD_CK(cudaStreamSynchronize(t->st2));
  CNLD_CK(cudaStreamSynchronize(t->st));
  return 0;
}

/* diagnostic (CNLD_B200_TD_TRACE=1 at create): %globaltimer marks [ns] of the step kernels of samples [s0, s1), 8 per
 * sample: kernel start, prologue done, dependency met, base ready, first fixed point done, second done, rows written */
int cnld_b200_td_trace(cnld_td *t, cnld_uint s0, cnld_uint s1, unsigned long long *marks) {
  if (!t || s0 > s1 || s1 > t->nt || !marks) { set_error("td_trace: bad arguments"); return -1; }
  if (!t->d_trace) { set_error("td_trace: create the solver with CNLD_B200_TD_TRACE=1"); return -1; }
  CNLD_CK(cudaSetDevice(t->device));
  CNLD_CK(cudaMemcpy(marks, t->d_trace + (size_t)s0 * 8, sizeof(unsigned long long) * 8 * (s1 - s0), cudaMemcpyDeviceToHost));
  return 0;
}

cnld_uint cnld_b200_td_current_step(const cnld_td *t) { return t ? t->cur : 0; }
/* taps the convolution reads: 1 + the last tap that is nonzero in any patch pair (trailing all-zero taps are skipped) */
cnld_uint cnld_b200_td_taps(const cnld_td *t) { return t ? t->ntap : 0; }

/* state rows [i0, i1) -> HOST arrays [i1 - i0][npatch] (any of them may be NULL) */
int cnld_b200_td_state(cnld_td *t, cnld_uint i0, cnld_uint i1, double *x, double *u, double *p_es, double *p_cont_spr,
                       double *p_cont_dmp, double *p_tot) {
  if (!t || i0 > i1 || i1 > t->nt) { set_error("td_state: bad range"); return -1; }
  CNLD_CK(cudaSetDevice(t->device));
  double *dst[TD_NSTATE] = {x, u, p_es, p_cont_spr, p_cont_dmp, p_tot};
  const size_t plane = (size_t)t->nt * t->P, off = (size_t)i0 * t->P, cnt = (size_t)(i1 - i0) * t->P;
  for (int a = 0; a < TD_NSTATE; a++)
    if (dst[a] && cnt) CNLD_CK(cudaMemcpyAsync(dst[a], t->d_state + a * plane + off, sizeof(double) * cnt, cudaMemcpyDeviceToHost, t->st));
  CNLD_CK(cudaStreamSynchronize(t->st));
  return 0;
}

/* final fixed-point error and convolution count of the steps that produced samples [s0, s1), s0 >= 1 */
int cnld_b200_td_info(cnld_td *t, cnld_uint s0, cnld_uint s1, double *error, int *iters) {
  if (!t || s0 > s1 || s1 > t->nt) { set_error("td_info: bad range"); return -1; }
  CNLD_CK(cudaSetDevice(t->device));
  if (error && s1 > s0) CNLD_CK(cudaMemcpyAsync(error, t->d_err + s0, sizeof(double) * (s1 - s0), cudaMemcpyDeviceToHost, t->st));
  if (iters && s1 > s0) CNLD_CK(cudaMemcpyAsync(iters, t->d_iters + s0, sizeof(int) * (s1 - s0), cudaMemcpyDeviceToHost, t->st));
  CNLD_CK(cudaStreamSynchronize(t->st));
  return 0;
}

}  // extern "C"
