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
           const double *__restrict__ h, size_t hs, size_t last, double *partial, uint ngroup) {
  __shared__ double win[FC_KTILE];
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
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      uint k = lane;
      for (; k + 96 < kt; k += 128) {
        a0 = fma(__ldcs(f + k), win[k], a0);
        a1 = fma(__ldcs(f + k + 32), win[k + 32], a1);
        a2 = fma(__ldcs(f + k + 64), win[k + 64], a2);
        a3 = fma(__ldcs(f + k + 96), win[k + 96], a3);
      }
      for (; k < kt; k += 32) a0 = fma(__ldcs(f + k), win[k], a0);
      double acc = (a0 + a1) + (a2 + a3);
      for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (lane == 0) {
        double *dst = partial + (size_t)j * nsrc + i;
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
                                                                  hs, last, f->d_part, f->ngroup);
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
