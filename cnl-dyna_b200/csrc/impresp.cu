// Frequency response -> impulse response on the device (C ABI: cnld_b200_fft_to_fir).
//
// Restates impulse_response.fft_to_fir / fft_to_sir (cnld/impulse_response.py:191-224) for a batch of
// one-sided spectra s[b][0..nf):
//   interp_fft      (:161-175) magnitude and unwrapped phase interpolated linearly on an n-times finer grid
//   one_to_two(odd) (:11-46)   conjugate mirror to nfft = 2 nfi - 1 bins
//   kramers_kronig2 (:115-125) optional: imaginary part rebuilt from the real part by a Hilbert transform
//   real(ifft) * fs (:200-206)
// The mirrored spectrum is conjugate symmetric, so the inverse transform of nfft bins is a real sum over
// the nfi one-sided bins,
//   h[n] = fs/nfft * Re( sum_k w_k s_k e^{+2 pi i k n / nfft} ),   w_0 = 1, w_k = 2,
// and the Kramers-Kronig step (Hilbert transform of the even real part along the frequency axis, then
// the same inverse transform) reduces to the same sum over Re(s_k) alone with the second half of the
// output zeroed and the first half doubled (derivation in DESIGN.md): a causal response.  The sum is one
// complex GEMM  [nfft x nfi] . [nfi x batch]  with an exactly reduced twiddle matrix (k n mod nfft in
// integers), run on the FP64 tensor-core ZGEMM -- no FFT library, and exact to rounding where an FFT
// accumulates O(log n) errors.  nfft is odd by construction (2 nfi - 1), so there is no radix-2 shortcut
// to lose; at P x P = 20 736 spectra of 4 000 bins it is 5e12 flop, 0.2 s.
#include <algorithm>
#include <cmath>

#include "../../include/cnld_b200.h"
#include "common.cuh"

using namespace cnld;

namespace {
constexpr double IR_TWO_PI = 6.283185307179586476925286766559;

// interp_fft for one spectrum per thread: |s| and the unwrapped phase, linear in between.  fi = linspace(f0, f_last,
// (nf - 1) n + 1); the segment of fi[j] is found like searchsorted(f, fi, 'right') - 1 clipped to [0, nf - 2].
__global__ void k_ir_interp(uint nf, uint nfi, uint interp, uint nb, const double *__restrict__ f, const cplx *__restrict__ s,
                            int real_only, cplx *out) {
  const uint b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const cplx *sp = s + (size_t)b * nf;
  cplx *op = out + (size_t)b * nfi;
  if (interp == 1) {
    for (uint k = 0; k < nf; k++) op[k] = real_only ? make_double2(sp[k].x, 0.0) : sp[k];
    return;
  }
  const double f0 = f[0], fl = f[nf - 1], step = (fl - f0) / (double)(nfi - 1);   // numpy.linspace
  double m0 = hypot(sp[0].x, sp[0].y), p0 = atan2(sp[0].y, sp[0].x);   // unwrapped phase so far
  uint seg = 0;
  double m1 = hypot(sp[1].x, sp[1].y), a1 = atan2(sp[1].y, sp[1].x);
  auto unwrap = [](double prev, double raw) {       // numpy.unwrap: jumps larger than pi are taken out
    double d = raw - prev;
    d -= IR_TWO_PI * rint(d / IR_TWO_PI);
    // numpy maps a jump of exactly +-pi to the sign of the original difference
    if (d == -M_PI && raw - prev > 0) d = M_PI;
    return prev + d;
  };
  double p1 = unwrap(p0, a1);
  for (uint j = 0; j < nfi; j++) {
    const double fj = (j + 1 == nfi) ? fl : f0 + (double)j * step;
    while (seg + 2 < nf && fj >= f[seg + 1]) {      // advance to the segment [f[seg], f[seg+1]) that holds fj
      seg++;
      m0 = m1; p0 = p1;
      m1 = hypot(sp[seg + 1].x, sp[seg + 1].y);
      p1 = unwrap(p0, atan2(sp[seg + 1].y, sp[seg + 1].x));
    }
    const double df = f[seg + 1] - f[seg], t = fj - f[seg];
    const double mi = (m1 - m0) / df * t + m0, pi = (p1 - p0) / df * t + p0;
    double sn, cs;
    sincos(pi, &sn, &cs);
    op[j] = make_double2(mi * cs, real_only ? 0.0 : mi * sn);
  }
}

// E[n + k nfft] = w_k exp(+2 pi i (k n mod nfft) / nfft)
__global__ void k_ir_twiddle(uint nfft, uint nfi, cplx *E) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)nfft * nfi) return;
  const uint n = e % nfft, k = e / nfft;
  const unsigned long long r = ((unsigned long long)n * k) % nfft;
  double sn, cs;
  sincospi(2.0 * (double)r / (double)nfft, &sn, &cs);
  const double w = k ? 2.0 : 1.0;
  E[e] = make_double2(w * cs, w * sn);
}

// h[b][n] = scale * c_n * Re(C[n + b nfft]); Kramers-Kronig: c = 1 (n = 0), 2 (n <= (nfft-1)/2), 0 above
__global__ void k_ir_finish(uint nfft, uint nb, const cplx *__restrict__ C, double scale, int kkr, double *h) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)nfft * nb) return;
  const uint n = e % nfft;
  double c = 1.0;
  if (kkr) c = n == 0 ? 1.0 : (n <= (nfft - 1) / 2 ? 2.0 : 0.0);
  h[e] = scale * c * C[e].x;
}
}  // namespace

extern "C" {

int cnld_b200_fft_to_fir(int device, cnld_uint nf, const double *freqs, cnld_uint nbatch, const cnld_field *s,
                         cnld_uint interp, int use_kkr, double *t, double *h) {
  if (cnld_b200_device_count() <= device) { set_error("no CUDA device %d (libcnld_b200 has no CPU fallback)", device); return -1; }
  if (nf < 2 || !freqs || !s || !h || interp < 1) { set_error("fft_to_fir: need at least two frequencies, spectra and an output"); return -1; }
  for (uint k = 1; k < nf; k++)
    if (!(freqs[k] > freqs[k - 1])) { set_error("fft_to_fir: frequencies must increase"); return -1; }
  CNLD_CK(cudaSetDevice(device));
  const uint nfi = (nf - 1) * interp + 1, nfft = 2 * nfi - 1;
  const double dfi = (freqs[nf - 1] - freqs[0]) / (double)(nfi - 1);   // fi[1] - fi[0] of the linspace grid
  const double fs = dfi * nfft;
  if (t) for (uint n = 0; n < nfft; n++) t[n] = (double)n / fs;
  if (!nbatch) return 0;
  // batch chunks sized so that spectra + product + result stay within ~2 GB
  const size_t per = sizeof(cplx) * ((size_t)nf + nfi + nfft) + sizeof(double) * nfft;
  const uint chunk = (uint)std::max<size_t>(1, std::min<size_t>(nbatch, (2ull << 30) / per));
  double *d_f = nullptr, *d_h = nullptr;
  cplx *d_s = nullptr, *d_si = nullptr, *d_E = nullptr, *d_C = nullptr;
  int rc = -1;
  do {
    if (cudaMalloc(&d_f, sizeof(double) * nf) != cudaSuccess || cudaMalloc(&d_s, sizeof(cplx) * (size_t)chunk * nf) != cudaSuccess ||
        cudaMalloc(&d_si, sizeof(cplx) * (size_t)chunk * nfi) != cudaSuccess || cudaMalloc(&d_E, sizeof(cplx) * (size_t)nfft * nfi) != cudaSuccess ||
        cudaMalloc(&d_C, sizeof(cplx) * (size_t)chunk * nfft) != cudaSuccess || cudaMalloc(&d_h, sizeof(double) * (size_t)chunk * nfft) != cudaSuccess) break;
    if (cudaMemcpy(d_f, freqs, sizeof(double) * nf, cudaMemcpyHostToDevice) != cudaSuccess) break;
    k_ir_twiddle<<<(unsigned)(((size_t)nfft * nfi + 255) / 256), 256>>>(nfft, nfi, d_E);
    count_launch();
    bool ok = true;
    for (uint b0 = 0; b0 < nbatch && ok; b0 += chunk) {
      const uint nb = std::min(chunk, nbatch - b0);
      ok = cudaMemcpy(d_s, reinterpret_cast<const cplx *>(s) + (size_t)b0 * nf, sizeof(cplx) * (size_t)nb * nf, cudaMemcpyHostToDevice) == cudaSuccess;
      if (!ok) break;
      k_ir_interp<<<(nb + 63) / 64, 64>>>(nf, nfi, interp, nb, d_f, d_s, use_kkr ? 1 : 0, d_si);
      count_launch();
      if (zgemm3m(0, nfft, nb, nfi, 1.0, d_E, nfft, d_si, nfi, 0.0, d_C, nfft)) { rc = -2; ok = false; break; }
      k_ir_finish<<<(unsigned)(((size_t)nfft * nb + 255) / 256), 256>>>(nfft, nb, d_C, fs / (double)nfft, use_kkr ? 1 : 0, d_h);
      count_launch();
      ok = cudaMemcpy(h + (size_t)b0 * nfft, d_h, sizeof(double) * (size_t)nb * nfft, cudaMemcpyDeviceToHost) == cudaSuccess;
    }
    if (ok) rc = 0;
  } while (0);
  if (rc == -1) set_error("fft_to_fir: %s", cudaGetErrorString(cudaGetLastError()));
  cudaFree(d_f); cudaFree(d_s); cudaFree(d_si); cudaFree(d_E); cudaFree(d_C); cudaFree(d_h);
  return rc ? -1 : 0;
}

}  // extern "C"
