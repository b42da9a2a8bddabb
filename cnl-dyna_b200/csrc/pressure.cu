// Field pressure radiated by the array surface.
// Restates bem_pressure.mesh_pres_vector (cnld/bem_pressure.pyx:54-88):
//   p_vec[n] = -(k c)^2 * 2 rho * sum_{tri} sum_{g=1..3} (1/3) N_n(xi_g) area e^{-j k r_g}/(4 pi r_g)
// with the 3-point rule (1/6,1/6),(2/3,1/6),(1/6,2/3) (:25-27) and the source z forced to 0
// (:81), and the consumer loop of array_patch_pres_imp_resp (:103-112), p = p_vec . disp.
//
// k_pres_vector: reference-shaped output pvec[nobs][nv] (few observation points).
// k_pressure   : 10^5..10^6 observation points.  The node sum is reordered so the
//   displacement is interpolated to the quadrature points once
//   (D_g = scale * area/3/(4 pi) * sum_l N_l(xi_g) disp[v_l]) and every thread owns one
//   observation point, streaming the 3*nt source points through shared memory:
//   p(obs) = sum_g D_g * e^{-j k r}/r.  FP64 sincos/rsqrt bound.
#include "pair.cuh"

namespace cnld {
namespace {
__constant__ double c_xi[3] = {1.0 / 6, 2.0 / 3, 1.0 / 6};
__constant__ double c_eta[3] = {1.0 / 6, 1.0 / 6, 2.0 / 3};

__global__ void k_pres_vector(uint nt, const double *__restrict__ x, const uint *__restrict__ tri,
                              const double *__restrict__ g, const double *__restrict__ obs, double k,
                              double scale, uint nv, cplx *pvec) {
  const uint o = blockIdx.x;
  const double ox = obs[3 * o], oy = obs[3 * o + 1], oz = obs[3 * o + 2];
  double *p = reinterpret_cast<double *>(pvec + (size_t)o * nv);
  for (uint t = threadIdx.x; t < nt; t += blockDim.x) {
    uint v0 = tri[3 * t], v1 = tri[3 * t + 1], v2 = tri[3 * t + 2];
    double x1 = x[3 * v0], y1 = x[3 * v0 + 1], x2 = x[3 * v1], y2 = x[3 * v1 + 1], x3 = x[3 * v2],
           y3 = x[3 * v2 + 1];
    double da = 0.5 * g[t];
    for (int q = 0; q < 3; q++) {
      double xi = c_xi[q], eta = c_eta[q], n0 = 1.0 - xi - eta;
      double xs = x1 * n0 + x2 * xi + x3 * eta, ys = y1 * n0 + y2 * xi + y3 * eta;
      double r2 = (ox - xs) * (ox - xs) + (oy - ys) * (oy - ys) + oz * oz;
      double rinv = rsqrt(r2), r = r2 * rinv, s, c;
      sincos_fast(k * r, s, c);
      double f = scale * (1.0 / 3) * INV4PI * rinv * da;
      double cr = f * c, ci = -f * s;
      atomicAdd(p + 2 * (size_t)v0, n0 * cr);  atomicAdd(p + 2 * (size_t)v0 + 1, n0 * ci);
      atomicAdd(p + 2 * (size_t)v1, xi * cr);  atomicAdd(p + 2 * (size_t)v1 + 1, xi * ci);
      atomicAdd(p + 2 * (size_t)v2, eta * cr); atomicAdd(p + 2 * (size_t)v2 + 1, eta * ci);
    }
  }
}

// source points (x, y) of the 3-point rule, [3*nt][2]
__global__ void k_pres_points(uint nt, const double *__restrict__ x, const uint *__restrict__ tri,
                              double *sp) {
  uint t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  uint v0 = tri[3 * t], v1 = tri[3 * t + 1], v2 = tri[3 * t + 2];
  for (int q = 0; q < 3; q++) {
    double xi = c_xi[q], eta = c_eta[q], n0 = 1.0 - xi - eta;
    sp[2 * (3 * (size_t)t + q)] = x[3 * v0] * n0 + x[3 * v1] * xi + x[3 * v2] * eta;
    sp[2 * (3 * (size_t)t + q) + 1] = x[3 * v0 + 1] * n0 + x[3 * v1 + 1] * xi + x[3 * v2 + 1] * eta;
  }
}

// D[src][3*t+q] = scale * area/3/(4 pi) * sum_l N_l disp[src][v_l]
__global__ void k_pres_contract(uint nt, uint nv, uint nsrc, const uint *__restrict__ tri,
                                const double *__restrict__ g, const cplx *__restrict__ disp,
                                double scale, cplx *D) {
  uint t = blockIdx.x * blockDim.x + threadIdx.x;
  uint src = blockIdx.y;
  if (t >= nt || src >= nsrc) return;
  const cplx *d = disp + (size_t)src * nv;
  cplx d0 = d[tri[3 * t]], d1 = d[tri[3 * t + 1]], d2 = d[tri[3 * t + 2]];
  double f = scale * 0.5 * g[t] * (1.0 / 3) * INV4PI;
  for (int q = 0; q < 3; q++) {
    double xi = c_xi[q], eta = c_eta[q], n0 = 1.0 - xi - eta;
    D[(size_t)src * 3 * nt + 3 * (size_t)t + q] =
        make_double2(f * (n0 * d0.x + xi * d1.x + eta * d2.x), f * (n0 * d0.y + xi * d1.y + eta * d2.y));
  }
}

constexpr int PR_THREADS = 128;
constexpr int PR_CHUNK = 512;

template <int NS>
__global__ void __launch_bounds__(PR_THREADS)
k_pressure(uint nobs, const double *__restrict__ obs, uint npts, const double *__restrict__ sp,
           const cplx *__restrict__ D, size_t dstride, double k, cplx *out, size_t ostride) {
  __shared__ double s_xy[PR_CHUNK][2];
  __shared__ cplx s_d[NS][PR_CHUNK];
  const uint o = blockIdx.x * PR_THREADS + threadIdx.x;
  const bool live = o < nobs;
  double ox = 0, oy = 0, oz2 = 1;
  if (live) { ox = obs[3 * (size_t)o]; oy = obs[3 * (size_t)o + 1]; oz2 = obs[3 * (size_t)o + 2]; oz2 *= oz2; }
  cplx acc[NS];
#pragma unroll
  for (int s = 0; s < NS; s++) acc[s] = make_double2(0.0, 0.0);
  for (uint p0 = 0; p0 < npts; p0 += PR_CHUNK) {
    uint np = min((uint)PR_CHUNK, npts - p0);
    __syncthreads();
    for (uint i = threadIdx.x; i < np; i += PR_THREADS) {
      s_xy[i][0] = sp[2 * (size_t)(p0 + i)];
      s_xy[i][1] = sp[2 * (size_t)(p0 + i) + 1];
#pragma unroll
      for (int s = 0; s < NS; s++) s_d[s][i] = D[(size_t)s * dstride + p0 + i];
    }
    __syncthreads();
#pragma unroll 4
    for (uint i = 0; i < np; i++) {
      double dx = ox - s_xy[i][0], dy = oy - s_xy[i][1];
      double r2 = fma(dx, dx, fma(dy, dy, oz2));
      double rinv = rsqrt(r2), r = r2 * rinv, sn, cs;
      sincos_fast(k * r, sn, cs);
      double er = cs * rinv, ei = -sn * rinv;
#pragma unroll
      for (int s = 0; s < NS; s++) {
        cplx d = s_d[s][i];
        acc[s].x = fma(er, d.x, fma(-ei, d.y, acc[s].x));
        acc[s].y = fma(er, d.y, fma(ei, d.x, acc[s].y));
      }
    }
  }
  if (live) {
#pragma unroll
    for (int s = 0; s < NS; s++) out[(size_t)s * ostride + o] = acc[s];
  }
}
}  // namespace

int pres_vector(cudaStream_t st, const BemDev &b, const double *d_obs, uint nobs, double k, double c,
                double rho, cplx *d_pvec) {
  CNLD_CK(cudaMemsetAsync(d_pvec, 0, sizeof(cplx) * (size_t)nobs * b.nv, st));
  double scale = -(k * c) * (k * c) * rho * 2.0;
  k_pres_vector<<<nobs, 256, 0, st>>>(b.nt, b.x, b.t, b.g, d_obs, k, scale, b.nv, d_pvec);
  CNLD_CK_LAUNCH();
  return 0;
}

int pres_points(cudaStream_t st, const BemDev &b, double *d_sp) {
  k_pres_points<<<(b.nt + 127) / 128, 128, 0, st>>>(b.nt, b.x, b.t, d_sp);
  CNLD_CK_LAUNCH();
  return 0;
}

// out[src][obs] for one wavenumber; d_disp [nsrc][nv]; d_D scratch [nsrc][3 nt]
int pressure_field(cudaStream_t st, const BemDev &b, const double *d_obs, uint nobs, const double *d_sp,
                   double k, double c, double rho, const cplx *d_disp, uint nsrc, cplx *d_D, cplx *d_out) {
  double scale = -(k * c) * (k * c) * rho * 2.0;
  dim3 gc((b.nt + 127) / 128, nsrc);
  k_pres_contract<<<gc, 128, 0, st>>>(b.nt, b.nv, nsrc, b.t, b.g, d_disp, scale, d_D);
  CNLD_CK_LAUNCH();
  const uint npts = 3 * b.nt;
  const uint grid = (nobs + PR_THREADS - 1) / PR_THREADS;
  uint s = 0;
  while (s < nsrc) {
    uint left = nsrc - s;
    const cplx *D = d_D + (size_t)s * npts;
    cplx *out = d_out + (size_t)s * nobs;
    if (left >= 4) { k_pressure<4><<<grid, PR_THREADS, 0, st>>>(nobs, d_obs, npts, d_sp, D, npts, k, out, nobs); s += 4; }
    else if (left >= 2) { k_pressure<2><<<grid, PR_THREADS, 0, st>>>(nobs, d_obs, npts, d_sp, D, npts, k, out, nobs); s += 2; }
    else { k_pressure<1><<<grid, PR_THREADS, 0, st>>>(nobs, d_obs, npts, d_sp, D, npts, k, out, nobs); s += 1; }
    CNLD_CK_LAUNCH();
  }
  return 0;
}

}  // namespace cnld
