// Micro-benchmarks that provide the FP64 roofline denominators the driver's
// MEASURED_PEAKS.json does not carry (it has HBM GB/s and bf16 TFLOP/s only):
//   * the DMMA ZGEMM of zgemm.cu on device-resident data,
//   * the plain FP64 FMA pipe (independent DFMA chains).
#include "../../include/cnld_b200.h"
#include "common.cuh"

namespace cnld {
namespace {
__global__ void k_fill(cplx *p, size_t n, unsigned seed) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned h = (unsigned)i * 2654435761u + seed;
  h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
  unsigned h2 = h * 3266489917u + 0x9e3779b9u;
  p[i] = make_double2((h & 0xffff) / 65536.0 - 0.5, (h2 & 0xffff) / 65536.0 - 0.5);
}

__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b) {
  double r[16];
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = fma(r[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += r[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace
}  // namespace cnld

using namespace cnld;

extern "C" int cnld_b200_bench_zgemm(int device, cnld_uint m, cnld_uint n, cnld_uint k, int reps,
                                     double *tflops) {
  if (cnld_b200_device_count() <= device) { set_error("no CUDA device %d", device); return -1; }
  CNLD_CK(cudaSetDevice(device));
  cplx *A = nullptr, *B = nullptr, *C = nullptr;
  CNLD_CK(cudaMalloc(&A, sizeof(cplx) * (size_t)m * k));
  CNLD_CK(cudaMalloc(&B, sizeof(cplx) * (size_t)k * n));
  CNLD_CK(cudaMalloc(&C, sizeof(cplx) * (size_t)m * n));
  k_fill<<<(unsigned)(((size_t)m * k + 255) / 256), 256>>>(A, (size_t)m * k, 1u);
  k_fill<<<(unsigned)(((size_t)k * n + 255) / 256), 256>>>(B, (size_t)k * n, 2u);
  k_fill<<<(unsigned)(((size_t)m * n + 255) / 256), 256>>>(C, (size_t)m * n, 3u);
  count_launch(3);
  cudaEvent_t e0, e1;
  CNLD_CK(cudaEventCreate(&e0));
  CNLD_CK(cudaEventCreate(&e1));
  int rc = 0;
  for (int w = 0; w < 2 && !rc; w++) rc = zgemm3m(0, m, n, k, -1.0, A, m, B, k, 1.0, C, m);
  if (!rc) {
    CNLD_CK(cudaEventRecord(e0, 0));
    for (int r = 0; r < reps && !rc; r++) rc = zgemm3m(0, m, n, k, -1.0, A, m, B, k, 1.0, C, m);
    CNLD_CK(cudaEventRecord(e1, 0));
    CNLD_CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CNLD_CK(cudaEventElapsedTime(&ms, e0, e1));
    *tflops = 8.0 * m * n * k * reps / (ms * 1e-3) / 1e12;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(A); cudaFree(B); cudaFree(C);
  return rc;
}

extern "C" int cnld_b200_bench_dfma(int device, int reps, double *tflops) {
  if (cnld_b200_device_count() <= device) { set_error("no CUDA device %d", device); return -1; }
  CNLD_CK(cudaSetDevice(device));
  const int blocks = 148 * 8, threads = 256, iters = 4096;
  double *out = nullptr;
  CNLD_CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  CNLD_CK(cudaEventCreate(&e0));
  CNLD_CK(cudaEventCreate(&e1));
  k_dfma<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
  CNLD_CK(cudaEventRecord(e0, 0));
  for (int r = 0; r < reps; r++) k_dfma<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
  count_launch(reps + 1);
  CNLD_CK(cudaEventRecord(e1, 0));
  CNLD_CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CNLD_CK(cudaEventElapsedTime(&ms, e0, e1));
  *tflops = 2.0 * 16.0 * iters * (double)blocks * threads * reps / (ms * 1e-3) / 1e12;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  return 0;
}

extern "C" int cnld_b200_set_zgemm_variant(int v) { cnld::g_zgemm_variant = v; return 0; }
extern "C" int cnld_b200_set_lu_outer(int v) { cnld::g_lu_outer = v; return 0; }
extern "C" int cnld_b200_set_zgemm_3m(int v) { cnld::g_zgemm_3m = v; return 0; }

extern "C" int cnld_b200_set_lu_outer_inverse(int v) { cnld::g_lu_outer_inverse = v; return 0; }
namespace cnld_hapi { extern int g_hsweep_fluid_precond; extern int g_hsweep_coarse; }
extern "C" int cnld_b200_set_hsweep_coarse(int v) { cnld_hapi::g_hsweep_coarse = v; return 0; }
extern "C" int cnld_b200_set_hsweep_fluid_precond(int v) { cnld_hapi::g_hsweep_fluid_precond = v; return 0; }
extern "C" int cnld_b200_set_lu_reserve(int v) { cnld::g_lu_reserve_sms = v; return 0; }
namespace cnld { extern int g_hmv_tstack; extern int g_hmv_ywarps; }
extern "C" int cnld_b200_set_hmv_ywarps(int v) { cnld::g_hmv_ywarps = v; return 0; }
extern "C" int cnld_b200_set_hmv_tstack(int v) { cnld::g_hmv_tstack = v; return 0; }
