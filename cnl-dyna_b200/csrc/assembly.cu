// Galerkin assembly of the Helmholtz single-layer operator with linear hat
// functions, kernel exp(i k r)/(4 pi r):  Z_ij = int int phi_i(x) g(x,y) phi_j(y).
// Replaces assemble_bem3d_amatrix -> bem->nearfield = assemble_ll_near_bem3d
// (include/bem3d.h:3028, :1728-1735; kernel form include/helmholtzbem3d.h:32-35 and
// scratch/demo_zmatrix_hmatrix.c:97-121; call site cnld/compressed_formats.py:629-634).
//
// Two kernels accumulate alpha * Z into a pre-initialised matrix (the sweep
// initialises it with K - w^2 M - j w B and passes alpha = -2 rho w^2, so G is formed
// without an extra pass over HBM -- cnld/scripts/generate_db.py:54-59):
//   k_regular : every disjoint triangle pair, q^4 tensor Duffy-Gauss points.  One
//               thread owns a row triangle (lanes run along rows = the contiguous
//               direction of the column-major matrix), a CTA streams a chunk of
//               column triangles through shared memory.  FP64 sincos + rsqrt.
//   k_singular: one warp per touching pair (identical / common edge / common vertex),
//               lanes stride the Sauter-Schwab points, warp-shuffle reduction.
#include "pair.cuh"

namespace cnld {

namespace {
constexpr int MAXNQ = 16;


// per-triangle setup: Gram determinant g = |(B-A)x(C-A)| = 2*area, unit normal, and the
// regular rule's points in global coordinates (H2Lib USE_TRIQUADPOINTS, singquad2d.h:120-135)
__global__ void k_tri_setup(uint nt, uint nq, const __grid_constant__ RegRule R,
                            const double *__restrict__ x, const uint *__restrict__ t, double *g,
                            double *nrm, double *qp) {
  uint i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nt) return;
  const double *A = x + 3 * (size_t)t[3 * i], *B = x + 3 * (size_t)t[3 * i + 1],
               *C = x + 3 * (size_t)t[3 * i + 2];
  double u0 = B[0] - A[0], u1 = B[1] - A[1], u2 = B[2] - A[2];
  double v0 = C[0] - A[0], v1 = C[1] - A[1], v2 = C[2] - A[2];
  double c0 = u1 * v2 - u2 * v1, c1 = u2 * v0 - u0 * v2, c2 = u0 * v1 - u1 * v0;
  double nn = sqrt(c0 * c0 + c1 * c1 + c2 * c2);
  g[i] = nn;
  nrm[3 * i] = c0 / nn; nrm[3 * i + 1] = c1 / nn; nrm[3 * i + 2] = c2 / nn;
  for (uint p = 0; p < nq; p++)
    for (int d = 0; d < 3; d++)
      qp[((size_t)i * nq + p) * 3 + d] = A[d] * R.phi[p][0] + B[d] * R.phi[p][1] + C[d] * R.phi[p][2];
}

constexpr int REG_THREADS = 128;
constexpr int REG_CHUNK = 64;

template <int NQ, bool CPLXK>
__global__ void __launch_bounds__(REG_THREADS)
k_regular(uint nt, const __grid_constant__ RegRule R, const uint *__restrict__ tri,
          const double *__restrict__ qp, const double *__restrict__ g, double kr, double ki,
          cplx alpha, cplx *Z, size_t ld) {
  __shared__ double s_qp[REG_CHUNK][NQ][3];
  __shared__ uint s_tv[REG_CHUNK][3];
  __shared__ double s_g[REG_CHUNK];
  const uint t = blockIdx.x * REG_THREADS + threadIdx.x;
  const uint s0 = blockIdx.y * REG_CHUNK;
  const uint ns = min((uint)REG_CHUNK, nt - s0);
  for (uint i = threadIdx.x; i < ns * NQ * 3; i += REG_THREADS)
    (&s_qp[0][0][0])[i] = qp[(size_t)s0 * NQ * 3 + i];
  for (uint i = threadIdx.x; i < ns * 3; i += REG_THREADS) (&s_tv[0][0])[i] = tri[(size_t)s0 * 3 + i];
  for (uint i = threadIdx.x; i < ns; i += REG_THREADS) s_g[i] = g[s0 + i];
  __syncthreads();
  if (t >= nt) return;
  uint tv[3] = {tri[3 * (size_t)t], tri[3 * (size_t)t + 1], tri[3 * (size_t)t + 2]};
  double xp[NQ][3];
#pragma unroll
  for (int p = 0; p < NQ; p++)
#pragma unroll
    for (int d = 0; d < 3; d++) xp[p][d] = qp[((size_t)t * NQ + p) * 3 + d];
  const double gt = g[t] * INV4PI;

  for (uint sl = 0; sl < ns; sl++) {
    const uint sv0 = s_tv[sl][0], sv1 = s_tv[sl][1], sv2 = s_tv[sl][2];
    bool touch = (tv[0] == sv0) | (tv[0] == sv1) | (tv[0] == sv2) | (tv[1] == sv0) | (tv[1] == sv1) |
                 (tv[1] == sv2) | (tv[2] == sv0) | (tv[2] == sv1) | (tv[2] == sv2);
    if (touch) continue;  // handled by k_singular
    cplx loc[3][3];
    regular_pair<NQ, CPLXK>(xp, &s_qp[sl][0][0], R, kr, ki, loc);
    const double sc = gt * s_g[sl];
    const cplx f = make_double2(alpha.x * sc, alpha.y * sc);
    const uint sv[3] = {sv0, sv1, sv2};
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int i = 0; i < 3; i++) {
        cplx v = cmul(loc[i][j], f);
        double *dst = reinterpret_cast<double *>(Z + (size_t)sv[j] * ld + tv[i]);
        atomicAdd(dst, v.x);
        atomicAdd(dst + 1, v.y);
      }
  }
}

constexpr int SING_WARPS = 8;

template <bool CPLXK>
__global__ void __launch_bounds__(SING_WARPS * 32)
k_singular(size_t npairs, const TouchPair *__restrict__ pairs, const double *__restrict__ x,
           const double *__restrict__ g, SingTables tab, double kr, double ki, cplx alpha, cplx *Z,
           size_t ld) {
  const size_t pid = (size_t)blockIdx.x * SING_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pid >= npairs) return;
  const TouchPair P = pairs[pid];
  double V[6][3];
#pragma unroll
  for (int v = 0; v < 3; v++)
#pragma unroll
    for (int d = 0; d < 3; d++) {
      V[v][d] = x[3 * (size_t)P.rv[v] + d];
      V[3 + v][d] = x[3 * (size_t)P.cv[v] + d];
    }
  double acc[18];
  singular_pair_warp<CPLXK>(V, tab, P.family, kr, ki, lane, acc);
  if (lane == 0) {
    double sc = g[P.t] * g[P.s] * INV4PI;
    cplx f = make_double2(alpha.x * sc, alpha.y * sc);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        cplx v = cmul(make_double2(acc[2 * (3 * i + j)], acc[2 * (3 * i + j) + 1]), f);
        double *dst = reinterpret_cast<double *>(Z + (size_t)P.cv[j] * ld + P.rv[i]);
        atomicAdd(dst, v.x);
        atomicAdd(dst + 1, v.y);
      }
  }
}


// ---- symmetric path ---------------------------------------------------------------------------
// Only pairs with s < t are evaluated; the 3x3 block goes to the LOWER triangle: entry
// (max(row, col), min(row, col)).  For a disjoint pair the transposed pair (s, t) has exactly the
// transposed block (the tensor rule and the kernel are symmetric), so this fills the lower
// triangle completely with half the kernel evaluations.
template <int NQ, bool CPLXK>
__global__ void __launch_bounds__(REG_THREADS)
k_regular_sym(const WorkItem *__restrict__ work, const __grid_constant__ RegRule R, const uint *__restrict__ tri,
              const double *__restrict__ qp, const double *__restrict__ g, double kr, double ki,
              cplx alpha, cplx *Z, size_t ld) {
  __shared__ double s_qp[REG_CHUNK][NQ][3];
  __shared__ uint s_tv[REG_CHUNK][3];
  __shared__ double s_g[REG_CHUNK];
  const WorkItem wi = work[blockIdx.x];   // row triangles [t_begin, t_end) x column chunk [s_begin, s_end)
  const uint t = wi.t_begin + threadIdx.x;
  const uint nt = wi.t_end;
  const uint s0 = wi.s_begin;
  const uint ns = wi.s_end - wi.s_begin;
  for (uint i = threadIdx.x; i < ns * NQ * 3; i += REG_THREADS)
    (&s_qp[0][0][0])[i] = qp[(size_t)s0 * NQ * 3 + i];
  for (uint i = threadIdx.x; i < ns * 3; i += REG_THREADS) (&s_tv[0][0])[i] = tri[(size_t)s0 * 3 + i];
  for (uint i = threadIdx.x; i < ns; i += REG_THREADS) s_g[i] = g[s0 + i];
  __syncthreads();
  if (t >= nt) return;
  uint tv[3] = {tri[3 * (size_t)t], tri[3 * (size_t)t + 1], tri[3 * (size_t)t + 2]};
  double xp[NQ][3];
#pragma unroll
  for (int p = 0; p < NQ; p++)
#pragma unroll
    for (int d = 0; d < 3; d++) xp[p][d] = qp[((size_t)t * NQ + p) * 3 + d];
  const double gt = g[t] * INV4PI;
  const uint smax = min(ns, t > s0 ? t - s0 : 0u);  // s < t
  for (uint sl = 0; sl < smax; sl++) {
    const uint sv0 = s_tv[sl][0], sv1 = s_tv[sl][1], sv2 = s_tv[sl][2];
    bool touch = (tv[0] == sv0) | (tv[0] == sv1) | (tv[0] == sv2) | (tv[1] == sv0) | (tv[1] == sv1) |
                 (tv[1] == sv2) | (tv[2] == sv0) | (tv[2] == sv1) | (tv[2] == sv2);
    if (touch) continue;
    cplx loc[3][3];
    regular_pair<NQ, CPLXK>(xp, &s_qp[sl][0][0], R, kr, ki, loc);
    const double sc = gt * s_g[sl];
    const cplx f = make_double2(alpha.x * sc, alpha.y * sc);
    const uint sv[3] = {sv0, sv1, sv2};
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int i = 0; i < 3; i++) {
        cplx v = cmul(loc[i][j], f);
        const uint r = max(tv[i], sv[j]), c = min(tv[i], sv[j]);
        double *dst = reinterpret_cast<double *>(Z + (size_t)c * ld + r);
        atomicAdd(dst, v.x);
        atomicAdd(dst + 1, v.y);
      }
  }
}

// touching pairs -> sparse slots (no alpha: it is applied when the slots are folded in)
template <bool CPLXK>
__global__ void __launch_bounds__(SING_WARPS * 32)
k_singular_slots(size_t npairs, const TouchPair *__restrict__ pairs, const uint *__restrict__ pair_slot,
                 const double *__restrict__ x, const double *__restrict__ g, SingTables tab, double kr,
                 double ki, cplx *sval) {
  const size_t pid = (size_t)blockIdx.x * SING_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pid >= npairs) return;
  const TouchPair P = pairs[pid];
  double V[6][3];
#pragma unroll
  for (int v = 0; v < 3; v++)
#pragma unroll
    for (int d = 0; d < 3; d++) {
      V[v][d] = x[3 * (size_t)P.rv[v] + d];
      V[3 + v][d] = x[3 * (size_t)P.cv[v] + d];
    }
  double acc[18];
  singular_pair_warp<CPLXK>(V, tab, P.family, kr, ki, lane, acc);
  if (lane < 9) {
    const double sc = g[P.t] * g[P.s] * INV4PI;
    double re = 0.0, im = 0.0;
#pragma unroll
    for (int e = 0; e < 9; e++)
      if (lane == e) { re = acc[2 * e]; im = acc[2 * e + 1]; }
    double *dst = reinterpret_cast<double *>(sval + pair_slot[9 * pid + lane]);
    atomicAdd(dst, re * sc);
    atomicAdd(dst + 1, im * sc);
  }
}

// replicate the integrated blocks: dst block = src block (or its transpose), blk x blk each
__global__ void __launch_bounds__(256)
k_block_copy(const CopyItem *__restrict__ items, uint blk, cplx *G, size_t ld) {
  __shared__ cplx tile[32][33];
  const CopyItem it = items[blockIdx.z];
  const uint i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;   // tile of the destination block
  const uint tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (!it.transposed) {
    for (uint k = ty; k < 32; k += 8) {
      const uint i = i0 + tx, j = j0 + k;
      if (i < blk && j < blk) G[(size_t)(it.dst_c0 + j) * ld + it.dst_r0 + i] = G[(size_t)(it.src_c0 + j) * ld + it.src_r0 + i];
    }
    return;
  }
  for (uint k = ty; k < 32; k += 8) {   // src(j, i): rows j contiguous
    const uint j = j0 + tx, i = i0 + k;
    if (i < blk && j < blk) tile[k][tx] = G[(size_t)(it.src_c0 + i) * ld + it.src_r0 + j];
  }
  __syncthreads();
  for (uint k = ty; k < 32; k += 8) {
    const uint i = i0 + tx, j = j0 + k;
    if (i < blk && j < blk) G[(size_t)(it.dst_c0 + j) * ld + it.dst_r0 + i] = tile[tx][k];
  }
}

// nslot0: slots of the integrated component(s); slot s takes its value from slot s % nslot0
__global__ void k_slots_fold(uint nslot, uint nslot0, const uint *__restrict__ row, const uint *__restrict__ col,
                             const uint *__restrict__ tr, const cplx *__restrict__ sval, cplx alpha, cplx *G,
                             size_t ld, cplx *dval) {
  uint s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslot) return;
  const uint r = row[s], c = col[s];
  cplx v = cmul(alpha, sval[s % nslot0]);
  if (r >= c) {
    cplx *d = G + (size_t)c * ld + r;  // every lower slot is unique; the regular kernel has finished
    d->x += v.x;
    d->y += v.y;
    dval[s] = make_double2(0.0, 0.0);
  } else {
    cplx w = cmul(alpha, sval[tr[s] % nslot0]);
    dval[s] = make_double2(v.x - w.x, v.y - w.y);
  }
}

template <int NQ>
int launch_regular_sym(cudaStream_t st, const BemDev &b, double kr, double ki, cplx alpha, cplx *Z, size_t ld) {
  const BemDev::SymPlan &pl = b.plan[(b.periodic && b.use_periodic) ? 1 : 0];
  if (!pl.nwork) return 0;
  if (ki != 0.0)
    k_regular_sym<NQ, true><<<pl.nwork, REG_THREADS, 0, st>>>(pl.work, b.rule, b.t, b.qp, b.g, kr, ki, alpha, Z, ld);
  else
    k_regular_sym<NQ, false><<<pl.nwork, REG_THREADS, 0, st>>>(pl.work, b.rule, b.t, b.qp, b.g, kr, ki, alpha, Z, ld);
  CNLD_CK_LAUNCH();
  return 0;
}


template <int NQ>
int launch_regular(cudaStream_t st, const BemDev &b, double kr, double ki, cplx alpha, cplx *Z, size_t ld) {
  dim3 grid((b.nt + REG_THREADS - 1) / REG_THREADS, (b.nt + REG_CHUNK - 1) / REG_CHUNK);
  if (ki != 0.0)
    k_regular<NQ, true><<<grid, REG_THREADS, 0, st>>>(b.nt, b.rule, b.t, b.qp, b.g, kr, ki, alpha, Z, ld);
  else
    k_regular<NQ, false><<<grid, REG_THREADS, 0, st>>>(b.nt, b.rule, b.t, b.qp, b.g, kr, ki, alpha, Z, ld);
  CNLD_CK_LAUNCH();
  return 0;
}
}  // namespace

int bem_create(BemDev &b, int device, uint nv, uint nt, const double *x, const uint *t, uint q, uint q2) {
  if (q < 1 || q > 4) { set_error("q_reg must be 1..4 (got %u)", q); return -1; }
  if (q2 < 1 || q2 > 12) { set_error("q_sing must be 1..12 (got %u)", q2); return -1; }
  b = BemDev();
  b.device = device; b.nv = nv; b.nt = nt; b.q = q; b.q2 = q2; b.nq = q * q;
  CNLD_CK(cudaSetDevice(device));
  CNLD_CK(cudaMalloc(&b.x, sizeof(double) * 3 * (size_t)nv));
  CNLD_CK(cudaMalloc(&b.t, sizeof(uint) * 3 * (size_t)nt));
  CNLD_CK(cudaMalloc(&b.g, sizeof(double) * nt));
  CNLD_CK(cudaMalloc(&b.n, sizeof(double) * 3 * (size_t)nt));
  CNLD_CK(cudaMalloc(&b.qp, sizeof(double) * 3 * (size_t)nt * b.nq));
  CNLD_CK(cudaMemcpy(b.x, x, sizeof(double) * 3 * (size_t)nv, cudaMemcpyHostToDevice));
  CNLD_CK(cudaMemcpy(b.t, t, sizeof(uint) * 3 * (size_t)nt, cudaMemcpyHostToDevice));
  // regular rule (kernel parameter, lands in the constant bank)
  std::vector<double> rt, rs, rw;
  build_triangle_rule(q, rt, rs, rw);
  memset(&b.rule, 0, sizeof(b.rule));
  for (uint p = 0; p < b.nq && p < MAXNQ; p++) {
    b.rule.w[p] = rw[p];
    b.rule.phi[p][0] = 1.0 - rt[p]; b.rule.phi[p][1] = rt[p] - rs[p]; b.rule.phi[p][2] = rs[p];
    for (int d = 0; d < 3; d++) b.rule.wphi[p][d] = b.rule.w[p] * b.rule.phi[p][d];
  }
  // singular families
  for (int f = FAM_VERT; f <= FAM_ID; f++) {
    PairRule R;
    build_pair_rule(f, q, q2, R);
    size_t n = R.size();
    std::vector<double> flat;
    flat.reserve(5 * n);
    flat.insert(flat.end(), R.xt.begin(), R.xt.end());
    flat.insert(flat.end(), R.xs.begin(), R.xs.end());
    flat.insert(flat.end(), R.yt.begin(), R.yt.end());
    flat.insert(flat.end(), R.ys.begin(), R.ys.end());
    flat.insert(flat.end(), R.w.begin(), R.w.end());
    double *d = nullptr;
    CNLD_CK(cudaMalloc(&d, sizeof(double) * 5 * n));
    CNLD_CK(cudaMemcpy(d, flat.data(), sizeof(double) * 5 * n, cudaMemcpyHostToDevice));
    b.sing.p[f] = d;
    b.sing.n[f] = (uint)n;
  }
  std::vector<TouchPair> pairs;
  build_touching_pairs(nv, nt, t, pairs);
  b.npairs = pairs.size();
  CNLD_CK(cudaMalloc(&b.pairs, sizeof(TouchPair) * pairs.size()));
  CNLD_CK(cudaMemcpy(b.pairs, pairs.data(), sizeof(TouchPair) * pairs.size(), cudaMemcpyHostToDevice));
  {
    TouchSlots TS;
    build_touch_slots(pairs, TS);
    b.nslot = (uint)TS.row.size();
    CNLD_CK(cudaMalloc(&b.pair_slot, sizeof(uint) * TS.pair_slot.size()));
    CNLD_CK(cudaMalloc(&b.slot_row, sizeof(uint) * b.nslot));
    CNLD_CK(cudaMalloc(&b.slot_col, sizeof(uint) * b.nslot));
    CNLD_CK(cudaMalloc(&b.slot_tr, sizeof(uint) * b.nslot));
    CNLD_CK(cudaMemcpy(b.pair_slot, TS.pair_slot.data(), sizeof(uint) * TS.pair_slot.size(), cudaMemcpyHostToDevice));
    CNLD_CK(cudaMemcpy(b.slot_row, TS.row.data(), sizeof(uint) * b.nslot, cudaMemcpyHostToDevice));
    CNLD_CK(cudaMemcpy(b.slot_col, TS.col.data(), sizeof(uint) * b.nslot, cudaMemcpyHostToDevice));
    CNLD_CK(cudaMemcpy(b.slot_tr, TS.tr.data(), sizeof(uint) * b.nslot, cudaMemcpyHostToDevice));
    std::vector<uint> rowptr(nv + 1, 0);
    for (uint r : TS.row) rowptr[r + 1]++;
    for (uint v = 0; v < nv; v++) rowptr[v + 1] += rowptr[v];
    CNLD_CK(cudaMalloc(&b.slot_rowptr, sizeof(uint) * (nv + 1)));
    CNLD_CK(cudaMemcpy(b.slot_rowptr, rowptr.data(), sizeof(uint) * (nv + 1), cudaMemcpyHostToDevice));
  }
  for (int per = 0; per < 2; per++) {
    TranslationPlan P;
    build_translation_plan(nv, nt, x, t, REG_THREADS, REG_CHUNK, per == 1, P);
    if (per == 1) {
      if (!P.periodic) break;
      // the sparse slots / touching pairs must repeat with the components too
      if (b.nslot % P.ncomp || b.npairs % P.ncomp) break;
      bool ok = true;
      const uint ns0 = b.nslot / P.ncomp;
      for (size_t i = 0; i < b.npairs / P.ncomp && ok; i++) ok = pairs[i].t < P.ntc && pairs[i].s < P.ntc;
      if (!ok) break;
      b.periodic = true; b.ncomp = P.ncomp; b.nclass = P.nclass;
      b.plan[1].nslot0 = ns0; b.plan[1].npairs0 = b.npairs / P.ncomp; b.plan[1].blk = P.nvc;
    } else {
      b.plan[0].nslot0 = b.nslot; b.plan[0].npairs0 = b.npairs; b.plan[0].blk = nv;
    }
    BemDev::SymPlan &pl = b.plan[per];
    pl.nwork = (uint)P.work.size(); pl.ncopy = (uint)P.copy.size();
    CNLD_CK(cudaMalloc(&pl.work, sizeof(WorkItem) * (P.work.size() + 1)));
    CNLD_CK(cudaMemcpy(pl.work, P.work.data(), sizeof(WorkItem) * P.work.size(), cudaMemcpyHostToDevice));
    CNLD_CK(cudaMalloc(&pl.copy, sizeof(CopyItem) * (P.copy.size() + 1)));
    CNLD_CK(cudaMemcpy(pl.copy, P.copy.data(), sizeof(CopyItem) * P.copy.size(), cudaMemcpyHostToDevice));
  }
  k_tri_setup<<<(nt + 127) / 128, 128>>>(nt, b.nq, b.rule, b.x, b.t, b.g, b.n, b.qp);
  CNLD_CK_LAUNCH();
  CNLD_CK(cudaDeviceSynchronize());
  return 0;
}

void bem_free(BemDev &b) {
  cudaFree(b.x); cudaFree(b.t); cudaFree(b.g); cudaFree(b.n); cudaFree(b.qp); cudaFree(b.pairs);
  cudaFree(b.pair_slot); cudaFree(b.slot_row); cudaFree(b.slot_col); cudaFree(b.slot_tr); cudaFree(b.slot_rowptr);
  for (int f = 0; f < 4; f++) cudaFree(const_cast<double *>(b.sing.p[f]));
  for (int per = 0; per < 2; per++) { cudaFree(b.plan[per].work); cudaFree(b.plan[per].copy); }
  b = BemDev();
}

int bem_assemble(cudaStream_t st, const BemDev &b, double kr, double ki, cplx alpha, cplx *Z, size_t ld) {
  int rc;
  switch (b.nq) {
  case 1: rc = launch_regular<1>(st, b, kr, ki, alpha, Z, ld); break;
  case 4: rc = launch_regular<4>(st, b, kr, ki, alpha, Z, ld); break;
  case 9: rc = launch_regular<9>(st, b, kr, ki, alpha, Z, ld); break;
  default: rc = launch_regular<16>(st, b, kr, ki, alpha, Z, ld); break;
  }
  if (rc) return rc;
  unsigned grid = (unsigned)((b.npairs + SING_WARPS - 1) / SING_WARPS);
  if (ki != 0.0)
    k_singular<true><<<grid, SING_WARPS * 32, 0, st>>>(b.npairs, b.pairs, b.x, b.g, b.sing, kr, ki, alpha, Z, ld);
  else
    k_singular<false><<<grid, SING_WARPS * 32, 0, st>>>(b.npairs, b.pairs, b.x, b.g, b.sing, kr, ki, alpha, Z, ld);
  CNLD_CK_LAUNCH();
  return 0;
}

int bem_assemble_sym(cudaStream_t st, const BemDev &b, double kr, double ki, cplx alpha, cplx *G, size_t ld,
                     cplx *sval, cplx *dval) {
  const BemDev::SymPlan &pl = b.plan[(b.periodic && b.use_periodic) ? 1 : 0];
  CNLD_CK(cudaMemsetAsync(sval, 0, sizeof(cplx) * pl.nslot0, st));
  int rc;
  switch (b.nq) {
  case 1: rc = launch_regular_sym<1>(st, b, kr, ki, alpha, G, ld); break;
  case 4: rc = launch_regular_sym<4>(st, b, kr, ki, alpha, G, ld); break;
  case 9: rc = launch_regular_sym<9>(st, b, kr, ki, alpha, G, ld); break;
  default: rc = launch_regular_sym<16>(st, b, kr, ki, alpha, G, ld); break;
  }
  if (rc) return rc;
  if (pl.ncopy) {  // blocks with the same offset between their components are equal (periodic.h)
    dim3 grid((pl.blk + 31) / 32, (pl.blk + 31) / 32, pl.ncopy);
    k_block_copy<<<grid, 256, 0, st>>>(pl.copy, pl.blk, G, ld);
    CNLD_CK_LAUNCH();
  }
  if (pl.npairs0) {
    unsigned grid = (unsigned)((pl.npairs0 + SING_WARPS - 1) / SING_WARPS);
    if (ki != 0.0)
      k_singular_slots<true><<<grid, SING_WARPS * 32, 0, st>>>(pl.npairs0, b.pairs, b.pair_slot, b.x, b.g, b.sing, kr, ki, sval);
    else
      k_singular_slots<false><<<grid, SING_WARPS * 32, 0, st>>>(pl.npairs0, b.pairs, b.pair_slot, b.x, b.g, b.sing, kr, ki, sval);
    CNLD_CK_LAUNCH();
  }
  if (b.nslot) {
    k_slots_fold<<<(b.nslot + 255) / 256, 256, 0, st>>>(b.nslot, pl.nslot0, b.slot_row, b.slot_col, b.slot_tr, sval, alpha, G, ld, dval);
    CNLD_CK_LAUNCH();
  }
  return 0;
}

}  // namespace cnld
