// Device-side building blocks shared by the dense assembly (assembly.cu) and the H-matrix
// builders (hmatrix.cu): the Helmholtz kernel and the regular / singular panel-pair rules.
#pragma once
#include "bem.cuh"

namespace cnld {

constexpr double INV4PI = 0.0795774715459476678844418816863;

// exp(i kr r) * exp(-ki r) / r  for squared distance r2
template <bool CPLXK>
__device__ __forceinline__ cplx helmholtz(double r2, double kr, double ki) {
  double rinv = rsqrt(r2);
  double r = r2 * rinv;
  if (CPLXK) rinv *= exp(-ki * r);
  double s, c;
  sincos(kr * r, &s, &c);
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
      double w0 = R.w[q] * R.phi[q][0], w1 = R.w[q] * R.phi[q][1], w2 = R.w[q] * R.phi[q][2];
      a0.x = fma(w0, kv.x, a0.x); a0.y = fma(w0, kv.y, a0.y);
      a1.x = fma(w1, kv.x, a1.x); a1.y = fma(w1, kv.y, a1.y);
      a2.x = fma(w2, kv.x, a2.x); a2.y = fma(w2, kv.y, a2.y);
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      double wi = R.w[p] * R.phi[p][i];
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
