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
// the device between steps: the total-pressure history IS the convolution's input (source-major rows), the
// step's scalar logic runs in one CTA right behind the convolution on the same stream, and the state arrays are
// downloaded when the caller asks for them.
// np.vectorize quirk kept on purpose (results identical to the reference's): the contact closures return the
// Python int 0 outside contact and np.vectorize derives its output dtype from the first element, so whenever
// patch 0 is out of contact the contact pressures of all patches are truncated toward zero (int64).
struct cnld_td {
  int device = 0;
  uint P = 0, nfir = 0, nt = 0, ngroup = 1, per = 0, cur = 1;
  uint ntap = 0;                 // taps that matter: 1 + the last tap that is nonzero anywhere (<= nfir)
  double dt = 0, e0 = 0, k = 0, n = 0, x0 = 0, lmbd = 0, atol = 0;
  int maxiter = 0;
  bool pdl = true;               // step kernel launched as a programmatic dependent of the convolution (CNLD_B200_TD_PDL=0: plain)
  double *d_fir = nullptr, *d_fir0 = nullptr, *d_v = nullptr, *d_geff = nullptr, *d_hist = nullptr, *d_state = nullptr,
         *d_part = nullptr, *d_err = nullptr;
  int *d_iters = nullptr;
  cudaStream_t st = nullptr;
};

namespace {
constexpr int TD_THREADS = 1024;
enum { TD_X = 0, TD_U, TD_PES, TD_SPR, TD_DMP, TD_PTOT, TD_NSTATE };

struct TdArgs {
  uint P, nt, s;
  double dt, e0h, k, n, x0, lmbd, atol;   // e0h = -e_0 / 2
  int maxiter, stage;                     // stage: fir[:, :, 0] fits in shared memory next to the work vectors
  const double *fir0, *v, *geff, *part;
  double *hist, *state, *err;
  int *iters;
};

__global__ void k_td_gather_fir0(uint P, uint nfir, const double *__restrict__ fir, double *fir0) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < (size_t)P * P) fir0[e] = fir[e * nfir];
}
// 1 + the largest tap index that is nonzero in any (source, destination) pair: a Kramers-Kronig (causal) table is
// exactly zero in its second half (impulse_response.py:115-125), and taps that are zero everywhere add nothing
__global__ void k_td_last_tap(size_t rows, uint nfir, const double *__restrict__ fir, uint *last) {
  const size_t r = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const double *f = fir + r * nfir;
  uint best = 0;
  for (uint k = threadIdx.x & 31; k < nfir; k += 32)
    if (f[k] != 0.0) best = k + 1;
  for (int off = 16; off > 0; off >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, off));
  if ((threadIdx.x & 31) == 0 && best) atomicMax(last, best);
}

__device__ __forceinline__ double td_pow(double a, double n) {
  return n == 1.0 ? a : n == 2.0 ? a * a : pow(a, n);
}

// out[j] = sum_i M[i * P + j] * (W ? w[i] : 1): the sources are dealt to G thread groups, the group sums are added
// in group order (fixed association, independent of timing)
template <bool W>
__device__ void td_colsum(uint P, const double *__restrict__ M, const double *w, double *gsum, double *out) {
  const uint T = TD_THREADS, Pp = min((P + 31u) & ~31u, T), G = T / Pp, g = threadIdx.x / Pp, jj = threadIdx.x % Pp;
  const uint chunk = (P + G - 1) / G;
  if (g < G) {
    const uint i0 = g * chunk, i1 = min(P, i0 + chunk);
    for (uint j = jj; j < P; j += Pp) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      uint i = i0;
      for (; i + 3 < i1; i += 4) {
        const double m0 = M[(size_t)i * P + j], m1 = M[(size_t)(i + 1) * P + j], m2 = M[(size_t)(i + 2) * P + j],
                     m3 = M[(size_t)(i + 3) * P + j];
        a0 = W ? fma(m0, w[i], a0) : a0 + m0;
        a1 = W ? fma(m1, w[i + 1], a1) : a1 + m1;
        a2 = W ? fma(m2, w[i + 2], a2) : a2 + m2;
        a3 = W ? fma(m3, w[i + 3], a3) : a3 + m3;
      }
      for (; i < i1; i++) a0 = W ? fma(M[(size_t)i * P + j], w[i], a0) : a0 + M[(size_t)i * P + j];
      gsum[(size_t)g * P + j] = (a0 + a1) + (a2 + a3);
    }
  }
  __syncthreads();
  for (uint j = threadIdx.x; j < P; j += T) {
    double s = 0.0;
    for (uint q = 0; q < G; q++) s += gsum[(size_t)q * P + j];
    out[j] = s;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(TD_THREADS) k_td_step(TdArgs a) {
  extern __shared__ double sm[];
  const uint P = a.P, T = TD_THREADS, tid = threadIdx.x, s = a.s;
  double *base = sm, *x = base + P, *xp = x + P, *u = xp + P, *pes = u + P, *spr = pes + P, *dmp = spr + P, *pt = dmp + P,
         *v = pt + P, *ge = v + P, *tmp = ge + P, *red = tmp + P, *gsum = red + 32;
  double *S = a.state;
  const size_t plane = (size_t)a.nt * P;
  // prologue on data no earlier kernel of the step writes: drive, gaps and the zero-lag slice of the table
  const double *fir0 = a.fir0;
  if (a.stage) {
    const uint Pp = min((P + 31u) & ~31u, T), G = T / Pp;
    double *f0s = sm + (((size_t)11 * P + 32 + (size_t)G * P + 1) & ~(size_t)1);   // 16-byte aligned
    const size_t n2 = (size_t)P * P / 2;
    const double2 *src = reinterpret_cast<const double2 *>(a.fir0);
    double2 *dst = reinterpret_cast<double2 *>(f0s);
    for (size_t e = tid; e < n2; e += T) dst[e] = src[e];
    if (tid == 0 && (P & 1)) f0s[(size_t)P * P - 1] = a.fir0[(size_t)P * P - 1];
    fir0 = f0s;
  }
  for (uint j = tid; j < P; j += T) {
    v[j] = a.v[(size_t)s * P + j];
    ge[j] = a.geff[j];
    dmp[j] = 0.0;
    x[j] = 0.0;
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");   // the convolution (and with it every earlier step) is complete
  for (uint j = tid; j < P; j += T) xp[j] = s ? S[TD_X * plane + (size_t)(s - 1) * P + j] : 0.0;
  __syncthreads();

  auto update_all = [&]() {                           // _update_all, simulation.pyx:412-421
    const bool first_free = x[0] >= a.x0;
    for (uint j = tid; j < P; j += T) {
      const double xj = x[j], d = xj + ge[j];
      u[j] = s ? (xj - xp[j]) / a.dt : 0.0;
      pes[j] = (a.e0h * (v[j] * v[j])) / (d * d);
      double c = xj >= a.x0 ? 0.0 : a.k * td_pow(a.x0 - xj, a.n);
      if (first_free) c = trunc(c);
      spr[j] = c;
      pt[j] = (c + dmp[j]) + pes[j];
    }
    __syncthreads();
  };
  auto update_x = [&]() -> double {                   // _update_x, :430-445
    td_colsum<true>(P, fir0, pt, gsum, tmp);
    double e = 0.0;
    for (uint j = tid; j < P; j += T) {
      const double xn = __dadd_rn(base[j], __dmul_rn(tmp[j], a.dt));
      const double dj = fabs(x[j] - xn);
      e = (dj > e || dj != dj) ? dj : e;              // NaN sticks, like np.max
      x[j] = xn;
    }
    for (int off = 16; off > 0; off >>= 1) {
      const double o = __shfl_xor_sync(0xffffffffu, e, off);
      e = (o > e || o != o) ? o : e;
    }
    if ((tid & 31) == 0) red[tid >> 5] = e;
    __syncthreads();
    e = red[0];
    for (uint w = 1; w < T / 32; w++) e = (red[w] > e || red[w] != red[w]) ? red[w] : e;
    __syncthreads();
    return e;
  };

  if (s == 0) {                                       // initial state (:335)
    update_all();
  } else {
    td_colsum<false>(P, a.part, nullptr, gsum, tmp);  // finish the convolution: sum over the sources, times dt
    for (uint j = tid; j < P; j += T) x[j] = base[j] = tmp[j] * a.dt;      // _blind_x
    __syncthreads();
    update_all();
    double err = update_x();
    update_all();
    int it = 2;
    for (int q = 0; q < a.maxiter - 1; q++) {
      if (err <= a.atol) break;
      err = update_x();
      update_all();
      it++;
    }
    {                                                 // _update_cont_dmp (:423-428)
      const bool first_free = x[0] >= a.x0;
      for (uint j = tid; j < P; j += T) {
        double c = x[j] >= a.x0 ? 0.0 : -(a.lmbd * td_pow(a.x0 - x[j], a.n)) * u[j];
        if (first_free) c = trunc(c);
        tmp[j] = c;
      }
      __syncthreads();
      for (uint j = tid; j < P; j += T) dmp[j] = tmp[j];
      __syncthreads();
    }
    update_all();
    err = update_x();
    update_all();
    for (int q = 0; q < a.maxiter - 1; q++) {
      if (err <= a.atol) break;
      err = update_x();
      update_all();
      it++;
    }
    if (tid == 0) {
      a.err[s] = err;
      a.iters[s] = it;
    }
  }
  for (uint j = tid; j < P; j += T) {
    const size_t r = (size_t)s * P + j;
    S[TD_X * plane + r] = x[j];
    S[TD_U * plane + r] = u[j];
    S[TD_PES * plane + r] = pes[j];
    S[TD_SPR * plane + r] = spr[j];
    S[TD_DMP * plane + r] = dmp[j];
    S[TD_PTOT * plane + r] = pt[j];
    a.hist[(size_t)j * a.nt + s] = pt[j];             // the next steps' convolution input
  }
}

size_t td_smem_base(uint P) {
  const uint Pp = std::min((P + 31u) & ~31u, (uint)TD_THREADS), G = TD_THREADS / Pp;
  return sizeof(double) * (((size_t)11 * P + 32 + (size_t)G * P + 1) & ~(size_t)1);
}
bool td_stage(uint P) { return td_smem_base(P) + sizeof(double) * P * P <= 224 * 1024; }
size_t td_smem(uint P) { return td_smem_base(P) + (td_stage(P) ? sizeof(double) * P * P : 0); }

int td_launch_step(cnld_td *t, uint s) {
  if (s) {
    const uint K = t->ntap > 1 ? std::min(s, t->ntap - 1) : 0u;
    if (K) {
      k_fir_conv<<<dim3(t->P, t->ngroup), FC_THREADS, 0, t->st>>>(t->P, t->P, t->nfir, t->per, K, 1, t->d_fir, t->d_hist, t->nt,
                                                                   s - 1, t->d_part, 1, t->P);
      CNLD_CK_LAUNCH();
    } else {
      CNLD_CK(cudaMemsetAsync(t->d_part, 0, sizeof(double) * (size_t)t->P * t->P, t->st));
    }
  }
  TdArgs a{t->P, t->nt, s, t->dt, -t->e0 / 2, t->k, t->n, t->x0, t->lmbd, t->atol, t->maxiter, td_stage(t->P) ? 1 : 0, t->d_fir0,
           t->d_v, t->d_geff, t->d_part, t->d_hist, t->d_state, t->d_err, t->d_iters};
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(1);
  cfg.blockDim = dim3(TD_THREADS);
  cfg.dynamicSmemBytes = td_smem(t->P);
  cfg.stream = t->st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = t->pdl ? 1 : 0;
  CNLD_CK(cudaLaunchKernelEx(&cfg, k_td_step, a));
  CNLD_CK_LAUNCH();
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
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
  t->ngroup = std::max(1u, std::min((npatch + 7) / 8, (uint)((4 * nsm + npatch - 1) / npatch)));
  t->per = (npatch + t->ngroup - 1) / t->ngroup;
  t->ngroup = (npatch + t->per - 1) / t->per;
  const size_t PP = (size_t)npatch * npatch, tab = PP * nfir, rows = (size_t)nt * npatch;
  bool ok = cudaSetDevice(device) == cudaSuccess && cudaMalloc(&t->d_fir, sizeof(double) * tab) == cudaSuccess &&
            cudaMalloc(&t->d_fir0, sizeof(double) * PP) == cudaSuccess && cudaMalloc(&t->d_part, sizeof(double) * PP) == cudaSuccess &&
            cudaMalloc(&t->d_v, sizeof(double) * rows) == cudaSuccess && cudaMalloc(&t->d_geff, sizeof(double) * npatch) == cudaSuccess &&
            cudaMalloc(&t->d_hist, sizeof(double) * rows) == cudaSuccess &&
            cudaMalloc(&t->d_state, sizeof(double) * rows * TD_NSTATE) == cudaSuccess &&
            cudaMalloc(&t->d_err, sizeof(double) * nt) == cudaSuccess && cudaMalloc(&t->d_iters, sizeof(int) * nt) == cudaSuccess &&
            cudaMemcpy(t->d_fir, fir, sizeof(double) * tab, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemcpy(t->d_v, v, sizeof(double) * rows, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemcpy(t->d_geff, gap_eff, sizeof(double) * npatch, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaStreamCreateWithFlags(&t->st, cudaStreamNonBlocking) == cudaSuccess &&
            cudaFuncSetAttribute(k_td_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess;
  if (ok) {
    k_td_gather_fir0<<<(unsigned)((PP + 255) / 256), 256, 0, t->st>>>(npatch, nfir, t->d_fir, t->d_fir0);
    uint *d_last = reinterpret_cast<uint *>(t->d_iters);        // scratch until reset() clears it
    ok = cudaMemsetAsync(d_last, 0, sizeof(uint), t->st) == cudaSuccess;
    k_td_last_tap<<<(unsigned)((PP + 7) / 8), 256, 0, t->st>>>(PP, nfir, t->d_fir, d_last);
    count_launch(2);
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
  cudaFree(t->d_fir); cudaFree(t->d_fir0); cudaFree(t->d_part); cudaFree(t->d_v); cudaFree(t->d_geff); cudaFree(t->d_hist);
  cudaFree(t->d_state); cudaFree(t->d_err); cudaFree(t->d_iters);
  if (t->st) cudaStreamDestroy(t->st);
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
  return 0;
}

/* nsteps x FixedStepSolver.step (simulation.pyx:459-503), no host round trip between the steps */
int cnld_b200_td_step(cnld_td *t, cnld_uint nsteps) {
  if (!t) { set_error("null td solver"); return -1; }
  if ((size_t)t->cur + nsteps > t->nt) { set_error("td_step: %u steps from sample %u run past the %u samples of the time grid", nsteps, t->cur, t->nt); return -1; }
  CNLD_CK(cudaSetDevice(t->device));
  for (uint q = 0; q < nsteps; q++) {
    if (td_launch_step(t, t->cur)) return -1;
    t->cur++;
  }
  CNLD_CK(cudaStreamSynchronize(t->st));
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
