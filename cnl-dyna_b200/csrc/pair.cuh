// Device-side building blocks shared by the dense assembly (assembly.cu) and the H-matrix
// builders (hmatrix.cu): the Helmholtz kernel and the regular / singular panel-pair rules.
#pragma once
#include "bem.cuh"

namespace cnld {

constexpr double INV4PI = 0.0795774715459476678844418816863;

// sin / cos in FP64 for |x| < ~1e6 (k r is at most a few hundred on this path): two-constant
// Cody-Waite reduction by pi/2 with FMAs, fdlibm minimax kernels on [-pi/4, pi/4] (error < 1 ulp),
// quadrant from the low bits of the round-to-nearest magic-number add.  Branch-free, no local
// memory, ~22 FP64 instructions -- the library sincos() carries a Payne-Hanek slow path, a stack
// frame and about twice the instruction count, and this call dominates the assembly kernels.
// (coefficients live in constant memory so DFMA takes them as c[bank][offset] operands instead of
//  materialising 64-bit immediates through UMOV pairs)
__constant__ double c_sincos[16] = {
    6755399441055744.0,          // [0] 1.5 * 2^52
    0.63661977236758134308,      // [1] 2/pi
    1.57079632679489655800e+00,  // [2] pi/2 hi
    6.12323399573676603587e-17,  // [3] pi/2 lo
    1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,   // [4..9] sin
    -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01,
    -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,  // [10..15] cos
    2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02};

__device__ __forceinline__ void sincos_fast(double x, double &sn, double &cs) {
  double t = fma(x, c_sincos[1], c_sincos[0]);
  int q = __double2loint(t);
  double fn = t - c_sincos[0];
  double r = fma(-fn, c_sincos[2], x);
  r = fma(-fn, c_sincos[3], r);
  double r2 = r * r;
  double ps = c_sincos[4];
  ps = fma(ps, r2, c_sincos[5]);
  ps = fma(ps, r2, c_sincos[6]);
  ps = fma(ps, r2, c_sincos[7]);
  ps = fma(ps, r2, c_sincos[8]);
  ps = fma(ps, r2, c_sincos[9]);
  double s = fma(ps * r2, r, r);
  double pc = c_sincos[10];
  pc = fma(pc, r2, c_sincos[11]);
  pc = fma(pc, r2, c_sincos[12]);
  pc = fma(pc, r2, c_sincos[13]);
  pc = fma(pc, r2, c_sincos[14]);
  pc = fma(pc, r2, c_sincos[15]);
  double c = fma(pc * r2, r2, fma(-0.5, r2, 1.0));
  // quadrant: q & 1 swaps, bit 1 of q flips sin, bit 1 of (q + 1) flips cos
  double a = (q & 1) ? c : s, b = (q & 1) ? s : c;
  sn = (q & 2) ? -a : a;
  cs = ((q + 1) & 2) ? -b : b;
}

// 1/sqrt(x) for positive, normal x (squared distances between distinct quadrature points: never
// zero, denormal or infinite on this path).  MUFU.RSQ64H seed + one third-order step
// y = y0 + y0 e (1/2 + 3/8 e), e = 1 - x y0^2 -- the same arithmetic as the library rsqrt() minus its
// range check, whose branch (BSSY / BRA / BSYNC per call) keeps the compiler from interleaving
// independent kernel evaluations (ncu: 'wait' stalls on every dependent DFMA of a serialised chain).
__device__ __forceinline__ double rsqrt_pos(double x) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  const double e = fma(x, -(y0 * y0), 1.0);
  const double p = fma(e, 0.375, 0.5);
  return fma(p, y0 * e, y0);
}

// exp(i kr r) * exp(-ki r) / r  for squared distance r2
template <bool CPLXK>
__device__ __forceinline__ cplx helmholtz(double r2, double kr, double ki) {
  double rinv = rsqrt_pos(r2);
  double r = r2 * rinv;
  if (CPLXK) rinv *= exp(-ki * r);
  double s, c;
  sincos_fast(kr * r, s, c);
  return make_double2(c * rinv, s * rinv);
}

// 3x3 block of a DISJOINT triangle pair with the tensor Duffy-Gauss rule (NQ points per triangle):
// loc[i][j] = sum_p sum_q w_p phi_i(p) w_q phi_j(q) G(x_p, y_q)   (no g_t g_s / 4 pi scaling)
template <int NQ, bool CPLXK, typename YP>
__device__ __forceinline__ void regular_pair(const double (&xp)[NQ][3], YP yq, const RegRule &R, double kr,
                                             double ki, cplx (&loc)[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) loc[i][j] = make_double2(0.0, 0.0);
#pragma unroll
  for (int p = 0; p < NQ; p++) {
    cplx a0 = make_double2(0.0, 0.0), a1 = a0, a2 = a0;
#pragma unroll
    for (int q = 0; q < NQ; q++) {
      double dx = xp[p][0] - yq[3 * q], dy = xp[p][1] - yq[3 * q + 1], dz = xp[p][2] - yq[3 * q + 2];
      cplx kv = helmholtz<CPLXK>(dx * dx + dy * dy + dz * dz, kr, ki);
      double w0 = R.wphi[q][0], w1 = R.wphi[q][1], w2 = R.wphi[q][2];
      a0.x = fma(w0, kv.x, a0.x); a0.y = fma(w0, kv.y, a0.y);
      a1.x = fma(w1, kv.x, a1.x); a1.y = fma(w1, kv.y, a1.y);
      a2.x = fma(w2, kv.x, a2.x); a2.y = fma(w2, kv.y, a2.y);
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      double wi = R.wphi[p][i];
      loc[i][0].x = fma(wi, a0.x, loc[i][0].x); loc[i][0].y = fma(wi, a0.y, loc[i][0].y);
      loc[i][1].x = fma(wi, a1.x, loc[i][1].x); loc[i][1].y = fma(wi, a1.y, loc[i][1].y);
      loc[i][2].x = fma(wi, a2.x, loc[i][2].x); loc[i][2].y = fma(wi, a2.y, loc[i][2].y);
    }
  }
}

// One touching pair, evaluated cooperatively by a warp: V[0..2] / V[3..5] are the (permuted)
// vertices of the row / column triangle, shared vertices first.  On return every lane holds
// acc[2*(3i+j)], acc[2*(3i+j)+1] = re, im of sum_q w_q a_i b_j G  (no g_t g_s / 4 pi scaling).
template <bool CPLXK>
__device__ __forceinline__ void singular_pair_warp(const double (&V)[6][3], const SingTables &tab, uint family,
                                                   double kr, double ki, int lane, double (&acc)[18]) {
  const uint n = tab.n[family];
  const double *xt = tab.p[family], *xs = xt + n, *yt = xs + n, *ys = yt + n, *w = ys + n;
#pragma unroll
  for (int i = 0; i < 18; i++) acc[i] = 0.0;
  for (uint q = lane; q < n; q += 32) {
    double a[3] = {1.0 - xt[q], xt[q] - xs[q], xs[q]};
    double b[3] = {1.0 - yt[q], yt[q] - ys[q], ys[q]};
    double dx = V[0][0] * a[0] + V[1][0] * a[1] + V[2][0] * a[2] - (V[3][0] * b[0] + V[4][0] * b[1] + V[5][0] * b[2]);
    double dy = V[0][1] * a[0] + V[1][1] * a[1] + V[2][1] * a[2] - (V[3][1] * b[0] + V[4][1] * b[1] + V[5][1] * b[2]);
    double dz = V[0][2] * a[0] + V[1][2] * a[1] + V[2][2] * a[2] - (V[3][2] * b[0] + V[4][2] * b[1] + V[5][2] * b[2]);
    cplx kv = helmholtz<CPLXK>(dx * dx + dy * dy + dz * dz, kr, ki);
    double wq = w[q];
    kv.x *= wq; kv.y *= wq;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        double ab = a[i] * b[j];
        acc[2 * (3 * i + j)] = fma(ab, kv.x, acc[2 * (3 * i + j)]);
        acc[2 * (3 * i + j) + 1] = fma(ab, kv.y, acc[2 * (3 * i + j) + 1]);
      }
  }
#pragma unroll
  for (int i = 0; i < 18; i++)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
}

}  // namespace cnld
