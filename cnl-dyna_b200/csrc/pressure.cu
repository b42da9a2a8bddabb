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
//   p(obs) = sum_g D_g * e^{-j k r}/r.  FP64 sincos/rsqrt bound.  Right for 1..4 sources.
// k_pvec_tiles + zgemm3m: many sources (nsrc = P = 144 at C3).  e^{-j k r}/r is evaluated ONCE per
//   (observation point, source point) and folded to node space on chip
//   (PV[obs][node] = mesh_pres_vector of the reference, for a whole tile of observation points),
//   then p[src][obs] = PV . disp[src] is one DMMA ZGEMM (K = nv instead of 3 nt source points).
#include "pair.cuh"

#include <algorithm>
#include <numeric>

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
      double rinv = rsqrt_pos(r2), r = r2 * rinv, s, c;
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
      double rinv = rsqrt_pos(r2), r = r2 * rinv, sn, cs;
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

// ---- node-space pressure vectors for a tile of observation points -----------------------------
// The triangles are cut into spatially compact chunks (<= PV_TC triangles touching <= PV_NC nodes,
// host-built once per mesh, PresPlan).  A CTA owns PV_OT observation points (one per thread) and
// walks all chunks: per triangle three kernel evaluations, combined with the hat values of the
// 3-point rule ((2/3,1/6,1/6) and its rotations: node l gets w (S/6 + E_l/2), S = E_0+E_1+E_2)
// and accumulated into the chunk's [node][obs] tile in shared memory (a thread only touches its
// own column: no atomics, no conflicts).  After the chunk the tile is written to PV (column-major
// obs x node: coalesced along obs); a node that an earlier chunk already wrote is accumulated --
// by the same thread that wrote it, so plain loads/stores are ordered.
template <int OT, int NC, int TC, int TB>
struct PvCfg {
  static constexpr int SMEM = NC * OT * (int)sizeof(cplx);
};

// TB triangles (3 TB independent kernel evaluations) are evaluated before their shared-memory updates;
// the staged triangle records live in a separate static array, so the compiler knows the tile updates
// do not alias them and can overlap the next batch's evaluations with this batch's updates.
template <int OT, int NC, int TC, int TB, int MINB>
__global__ void __launch_bounds__(OT, MINB)
k_pvec_tiles(uint nobs, const double *__restrict__ obs, uint nchunk, const uint *__restrict__ tptr,
             const double *__restrict__ tri, const uint *__restrict__ nptr, const uint *__restrict__ node,
             double k, double scale, cplx *PV, size_t lda) {
  extern __shared__ __align__(16) unsigned char pv_smem[];
  __shared__ __align__(16) double stage[2][(TC + TB) * 8];
  cplx *tile = reinterpret_cast<cplx *>(pv_smem);
  const int tid = threadIdx.x;
  const uint o = blockIdx.x * OT + tid;
  const bool live = o < nobs;
  double ox = 0, oy = 0, oz2 = 1;
  if (live) { ox = obs[3 * (size_t)o]; oy = obs[3 * (size_t)o + 1]; oz2 = obs[3 * (size_t)o + 2]; oz2 *= oz2; }
  cplx *col = tile + tid;
  const double w6 = scale * (1.0 / 6), w2 = scale * 0.5;
  for (uint c = 0; c < nchunk; c++) {
    const uint t0 = tptr[c], ntc = tptr[c + 1] - t0, n0 = nptr[c], nn = nptr[c + 1] - n0;
    double *sg = stage[c & 1];
    {
      const double2 *src = reinterpret_cast<const double2 *>(tri + (size_t)t0 * 8);
      double2 *dst = reinterpret_cast<double2 *>(sg);
      for (uint i = tid; i < ntc * 4; i += OT) dst[i] = src[i];
      // pad to a multiple of TB with copies of the first record at weight 0
      const uint npad = (ntc + TB - 1) / TB * TB;
      for (uint i = ntc * 4 + tid; i < npad * 4; i += OT) {
        double2 v = src[i & 3];
        if ((i & 3) == 3) v.x = 0.0;
        dst[i] = v;
      }
    }
    for (uint ln = 0; ln < nn; ln++) col[ln * OT] = make_double2(0.0, 0.0);
    __syncthreads();
    for (uint i = 0; i < ntc; i += TB) {
      double er[TB][3], ei[TB][3], a6[TB], a2[TB];
      unsigned long long sl[TB];
#pragma unroll
      for (int b = 0; b < TB; b++) {
        const double2 *e = reinterpret_cast<const double2 *>(sg + 8 * (i + b));
        const double2 p0 = e[0], p1 = e[1], p2 = e[2], ws = e[3];
        const double px[3] = {p0.x, p1.x, p2.x}, py[3] = {p0.y, p1.y, p2.y};
#pragma unroll
        for (int q = 0; q < 3; q++) {
          double dx = ox - px[q], dy = oy - py[q];
          double r2 = fma(dx, dx, fma(dy, dy, oz2));
          double rinv = rsqrt_pos(r2), r = r2 * rinv, sn, cs;
          sincos_fast(k * r, sn, cs);
          er[b][q] = cs * rinv; ei[b][q] = -sn * rinv;
        }
        a6[b] = ws.x * w6; a2[b] = ws.x * w2;
        sl[b] = (unsigned long long)__double_as_longlong(ws.y);
      }
#pragma unroll
      for (int b = 0; b < TB; b++) {
        const double sr = a6[b] * (er[b][0] + er[b][1] + er[b][2]), si = a6[b] * (ei[b][0] + ei[b][1] + ei[b][2]);
#pragma unroll
        for (int l = 0; l < 3; l++) {
          const uint s = (uint)(sl[b] >> (8 * l)) & 0xffu;
          cplx v = col[s * OT];
          v.x += fma(a2[b], er[b][l], sr);
          v.y += fma(a2[b], ei[b][l], si);
          col[s * OT] = v;
        }
      }
    }
    if (live) {
#pragma unroll 4
      for (uint ln = 0; ln < nn; ln++) {
        const uint nd = node[n0 + ln];
        cplx *dst = PV + (size_t)(nd & 0x7fffffffu) * lda + o;
        cplx v = col[ln * OT];
        if (nd >> 31) { cplx old = *dst; v.x += old.x; v.y += old.y; }
        *dst = v;
      }
    }
  }
}

// variants (cnld_b200_set_pvec_variant): {observation points per CTA, nodes per chunk, triangles per
// chunk, triangles per batch, CTAs per SM}
struct PvVariant { int ot, nc, tc, tb, ctas; };
constexpr PvVariant PV_VARIANTS[] = {{128, 48, 64, 1, 2}, {128, 48, 64, 2, 2}, {128, 24, 32, 1, 4}, {128, 24, 32, 2, 4},
                                     {64, 48, 64, 2, 3}, {128, 32, 40, 2, 3}, {128, 16, 20, 1, 6}, {128, 24, 32, 3, 4}};
constexpr int PV_NVARIANTS = (int)(sizeof(PV_VARIANTS) / sizeof(PV_VARIANTS[0]));
}  // namespace

int pres_vector(cudaStream_t st, const BemDev &b, const double *d_obs, uint nobs, double k, double c,
                double rho, cplx *d_pvec) {
  CNLD_CK(cudaMemsetAsync(d_pvec, 0, sizeof(cplx) * (size_t)nobs * b.nv, st));
  double scale = -(k * c) * (k * c) * rho * 2.0;
  k_pres_vector<<<nobs, 256, 0, st>>>(b.nt, b.x, b.t, b.g, d_obs, k, scale, b.nv, d_pvec);
  CNLD_CK_LAUNCH();
  return 0;
}


int g_pvec_variant = 3;

// ---- PresPlan: triangle chunks of k_pvec_tiles (host, once per mesh) ----------------------------
namespace {
struct PlanBuilder {
  const double *x; const uint *t; uint nv;
  std::vector<double> cen;           // centroids [nt][3]
  std::vector<uint> order;           // triangle indices, permuted in place by the bisection
  std::vector<uint> tptr, nptr, node;
  std::vector<double> tri;
  std::vector<uint8_t> touched;
  std::vector<uint> scratch;
  const double *g;
  uint max_tc = 64, max_nc = 48;
  uint distinct_nodes(uint lo, uint hi) {
    scratch.clear();
    for (uint i = lo; i < hi; i++) for (int l = 0; l < 3; l++) scratch.push_back(t[3 * order[i] + l]);
    std::sort(scratch.begin(), scratch.end());
    scratch.erase(std::unique(scratch.begin(), scratch.end()), scratch.end());
    return (uint)scratch.size();
  }
  void emit(uint lo, uint hi) {
    distinct_nodes(lo, hi);                      // scratch = sorted unique node ids of the chunk
    for (uint i = lo; i < hi; i++) {
      const uint tt = order[i];
      const uint v[3] = {t[3 * tt], t[3 * tt + 1], t[3 * tt + 2]};
      static const double xi[3] = {1.0 / 6, 2.0 / 3, 1.0 / 6}, eta[3] = {1.0 / 6, 1.0 / 6, 2.0 / 3};
      for (int q = 0; q < 3; q++) {              // bem_pressure.pyx:25-27, :75-80 (source z forced to 0)
        const double n0 = 1.0 - xi[q] - eta[q];
        tri.push_back(x[3 * v[0]] * n0 + x[3 * v[1]] * xi[q] + x[3 * v[2]] * eta[q]);
        tri.push_back(x[3 * v[0] + 1] * n0 + x[3 * v[1] + 1] * xi[q] + x[3 * v[2] + 1] * eta[q]);
      }
      tri.push_back(0.5 * g[tt] * (1.0 / 3) * 0.0795774715459476678844418816863);
      unsigned long long sl = 0;
      for (int l = 0; l < 3; l++) {
        const uint s = (uint)(std::lower_bound(scratch.begin(), scratch.end(), v[l]) - scratch.begin());
        sl |= (unsigned long long)s << (8 * l);
      }
      double sd;
      std::memcpy(&sd, &sl, sizeof(sd));
      tri.push_back(sd);
    }
    for (uint nd : scratch) {
      node.push_back(nd | (touched[nd] ? 0x80000000u : 0u));
      touched[nd] = 1;
    }
    tptr.push_back((uint)(tri.size() / 8));
    nptr.push_back((uint)node.size());
  }
  void split(uint lo, uint hi) {
    if (hi - lo <= max_tc && distinct_nodes(lo, hi) <= max_nc) { emit(lo, hi); return; }
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (uint i = lo; i < hi; i++)
      for (int d = 0; d < 3; d++) {
        mn[d] = std::min(mn[d], cen[3 * order[i] + d]);
        mx[d] = std::max(mx[d], cen[3 * order[i] + d]);
      }
    int ax = 0;
    for (int d = 1; d < 3; d++) if (mx[d] - mn[d] > mx[ax] - mn[ax]) ax = d;
    const uint mid = lo + (hi - lo) / 2;
    std::nth_element(order.begin() + lo, order.begin() + mid, order.begin() + hi, [&](uint a, uint b) {
      const double ca = cen[3 * a + ax], cb = cen[3 * b + ax];
      return ca < cb || (ca == cb && a < b);
    });
    split(lo, mid);
    split(mid, hi);
  }
};
}  // namespace

void pres_plan_free(PresPlan &p) {
  cudaFree(p.tptr); cudaFree(p.tri); cudaFree(p.nptr); cudaFree(p.node);
  p = PresPlan();
}

// host part: chunks of the mesh (x, t, g = 2 * area per triangle)
static void pres_plan_host(uint nv, uint nt, const double *hx, const uint *ht, const double *hg, PlanBuilder &pb) {
  pb.max_tc = PV_VARIANTS[g_pvec_variant].tc; pb.max_nc = PV_VARIANTS[g_pvec_variant].nc;
  pb.x = hx; pb.t = ht; pb.nv = nv; pb.g = hg;
  pb.cen.resize(3 * (size_t)nt);
  for (uint tt = 0; tt < nt; tt++)
    for (int d = 0; d < 3; d++)
      pb.cen[3 * (size_t)tt + d] = (hx[3 * ht[3 * tt] + d] + hx[3 * ht[3 * tt + 1] + d] + hx[3 * ht[3 * tt + 2] + d]) / 3.0;
  pb.order.resize(nt);
  std::iota(pb.order.begin(), pb.order.end(), 0u);
  pb.touched.assign(nv, 0);
  pb.tptr.push_back(0); pb.nptr.push_back(0);
  if (nt) pb.split(0, nt);
}

// Host-only view of the plan (tests): counts = {chunks, node entries, max triangles per chunk, max
// nodes per chunk, accumulate-flagged entries, vertices without a triangle}; order[nt] (nullable) =
// the triangles in chunk order.
int pres_plan_stats(uint nv, uint nt, const double *hx, const uint *ht, uint *counts, uint *order) {
  std::vector<double> hg(nt, 1.0);
  PlanBuilder pb;
  pres_plan_host(nv, nt, hx, ht, hg.data(), pb);
  uint mt = 0, mn = 0, acc = 0, unt = 0;
  for (size_t c = 0; c + 1 < pb.tptr.size(); c++) {
    mt = std::max(mt, pb.tptr[c + 1] - pb.tptr[c]);
    mn = std::max(mn, pb.nptr[c + 1] - pb.nptr[c]);
  }
  for (uint nd : pb.node) acc += nd >> 31;
  for (uint v = 0; v < nv; v++) unt += !pb.touched[v];
  counts[0] = (uint)pb.tptr.size() - 1; counts[1] = (uint)pb.node.size(); counts[2] = mt; counts[3] = mn;
  counts[4] = acc; counts[5] = unt;
  if (order) std::copy(pb.order.begin(), pb.order.end(), order);
  return 0;
}

int pres_plan_build(const BemDev &b, const double *hx, const uint *ht, PresPlan &p) {
  PlanBuilder pb;
  std::vector<double> hg(b.nt);
  CNLD_CK(cudaMemcpy(hg.data(), b.g, sizeof(double) * b.nt, cudaMemcpyDeviceToHost));
  pres_plan_host(b.nv, b.nt, hx, ht, hg.data(), pb);
  p.variant = g_pvec_variant;
  p.nchunk = (uint)pb.tptr.size() - 1;
  p.nentries = (uint)pb.node.size();
  p.untouched.clear();
  for (uint v = 0; v < b.nv; v++) if (!pb.touched[v]) p.untouched.push_back(v);
  CNLD_CK(cudaMalloc(&p.tptr, sizeof(uint) * pb.tptr.size()));
  CNLD_CK(cudaMalloc(&p.nptr, sizeof(uint) * pb.nptr.size()));
  CNLD_CK(cudaMalloc(&p.tri, sizeof(double) * std::max<size_t>(pb.tri.size(), 8)));
  CNLD_CK(cudaMalloc(&p.node, sizeof(uint) * std::max<size_t>(pb.node.size(), 1)));
  CNLD_CK(cudaMemcpy(p.tptr, pb.tptr.data(), sizeof(uint) * pb.tptr.size(), cudaMemcpyHostToDevice));
  CNLD_CK(cudaMemcpy(p.nptr, pb.nptr.data(), sizeof(uint) * pb.nptr.size(), cudaMemcpyHostToDevice));
  CNLD_CK(cudaMemcpy(p.tri, pb.tri.data(), sizeof(double) * pb.tri.size(), cudaMemcpyHostToDevice));
  CNLD_CK(cudaMemcpy(p.node, pb.node.data(), sizeof(uint) * pb.node.size(), cudaMemcpyHostToDevice));
  return 0;
}

uint pres_tile_obs() { return PV_VARIANTS[g_pvec_variant].ot; }
uint pres_tile_ctas() { return PV_VARIANTS[g_pvec_variant].ctas; }
int pres_set_variant(int v) {
  if (v < 0 || v >= PV_NVARIANTS) { set_error("pvec variant %d outside 0..%d", v, PV_NVARIANTS - 1); return -1; }
  g_pvec_variant = v;
  return 0;
}

namespace {
template <int V>
int launch_pvec(cudaStream_t st, const PresPlan &p, const double *d_obs, uint mc, double k, double scale, cplx *d_PV,
                size_t lda) {
  constexpr PvVariant C = PV_VARIANTS[V];
  auto kern = k_pvec_tiles<C.ot, C.nc, C.tc, C.tb, C.ctas>;
  constexpr int SMEM = C.nc * C.ot * (int)sizeof(cplx);
  static bool attr_set[64] = {false};
  int dev = 0;
  CNLD_CK(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    CNLD_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set[dev & 63] = true;
  }
  kern<<<(mc + C.ot - 1) / C.ot, C.ot, SMEM, st>>>(mc, d_obs, p.nchunk, p.tptr, p.tri, p.nptr, p.node, k, scale, d_PV, lda);
  CNLD_CK_LAUNCH();
  return 0;
}
}  // namespace

// PV (column-major mc x nv, leading dimension lda >= mc) = pressure vectors of mc observation points
int pres_pvec_tiles(cudaStream_t st, const PresPlan &p, const BemDev &b, const double *d_obs, uint mc, double k,
                    double c, double rho, cplx *d_PV, size_t lda) {
  // a vertex no triangle references never gets written by the kernel: its column must read as zero
  for (uint v : p.untouched) CNLD_CK(cudaMemsetAsync(d_PV + (size_t)v * lda, 0, sizeof(cplx) * mc, st));
  const double scale = -(k * c) * (k * c) * rho * 2.0;
  switch (p.variant) {
    case 0: return launch_pvec<0>(st, p, d_obs, mc, k, scale, d_PV, lda);
    case 1: return launch_pvec<1>(st, p, d_obs, mc, k, scale, d_PV, lda);
    case 2: return launch_pvec<2>(st, p, d_obs, mc, k, scale, d_PV, lda);
    case 3: return launch_pvec<3>(st, p, d_obs, mc, k, scale, d_PV, lda);
    case 4: return launch_pvec<4>(st, p, d_obs, mc, k, scale, d_PV, lda);
    case 5: return launch_pvec<5>(st, p, d_obs, mc, k, scale, d_PV, lda);
    case 6: return launch_pvec<6>(st, p, d_obs, mc, k, scale, d_PV, lda);
    case 7: return launch_pvec<7>(st, p, d_obs, mc, k, scale, d_PV, lda);
  }
  set_error("pvec variant %d", p.variant);
  return -1;
}

// out[src][obs0 .. obs0 + mc) (row stride ldo) = PV (mc x nv) . disp[src][:]   -- one DMMA ZGEMM
int pres_contract(cudaStream_t st, const BemDev &b, const cplx *d_PV, size_t lda, uint mc, const cplx *d_disp,
                  uint nsrc, cplx *d_out, size_t ldo) {
  return zgemm3m(st, mc, nsrc, b.nv, 1.0, d_PV, lda, d_disp, b.nv, 0.0, d_out, ldo);
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
