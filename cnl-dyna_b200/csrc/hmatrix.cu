// H-matrix path for arrays too large to factor densely: host-built cluster / block tree
// (cluster.cpp), device-side leaf builders and the batched H-matrix matvec.
//   inadmissible leaf -> dense block          = bem->nearfield        (include/bem3d.h:370-385)
//   admissible leaf   -> partial-pivot ACA    = decomp_partialaca_rkmatrix with entry =
//                                               bem->nearfield_far    (include/aca.h:79-117, bem3d.h:405)
//   y += alpha Z x                            = addeval_hmatrix_avector (include/hmatrix.h:392-419)
// Replaces setup_hmatrix_aprx_paca_bem3d + assemble_bem3d_hmatrix (bem3d.h:2627, :3056; call
// sites cnld/compressed_formats.py:675-694).
//
// Layout in HBM: all leaves live in one pool of complex doubles; DOFs are permuted into cluster
// order so every leaf is a contiguous (row range) x (column range).  A dense leaf is an m x n
// column-major tile; a low-rank leaf stores U (m x kcap) then V (n x kcap), block ~ U V^T.
//   k_near  : one CTA per inadmissible leaf, one thread per disjoint triangle pair, touching pairs
//             queued in shared memory and evaluated one warp per pair (Sauter-Schwab rules).
//   k_aca   : one CTA per admissible leaf; residual rows / columns generated on demand, entry by
//             entry, from the vertex -> triangle fans with the regular rule.
//   k_hmv_* : leaf-parallel matvec, HBM-bound streaming of the pool.
#include "hmat.cuh"
#include <unordered_map>
#include <cmath>

#include <algorithm>
#include <map>

#include "pair.cuh"

namespace cnld {

namespace {

__device__ __forceinline__ void red_add(cplx *p, cplx v) {
  atomicAdd(reinterpret_cast<double *>(p), v.x);
  atomicAdd(reinterpret_cast<double *>(p) + 1, v.y);
}

constexpr int NEAR_THREADS = 128;
constexpr int NEAR_QCAP = 3072;

template <int NQ, bool CPLXK>
__global__ void __launch_bounds__(NEAR_THREADS)
k_near(const HLeafDev *__restrict__ leaves, const uint *__restrict__ near_ids, const __grid_constant__ RegRule R,
       const SingTables tab, const double *__restrict__ x, const uint *__restrict__ tri,
       const double *__restrict__ qp, const double *__restrict__ g, const uint *__restrict__ pos,
       const uint *__restrict__ cposv, const uint *__restrict__ ctri_ptr, const uint *__restrict__ ctri, double kr,
       double ki, cplx *pool, uint *err) {
  __shared__ uint2 queue[NEAR_QCAP];
  __shared__ uint qn;
  const HLeafDev L = leaves[near_ids[blockIdx.x]];
  const uint *rt = ctri + ctri_ptr[L.rtl], *ct = ctri + ctri_ptr[L.ctl];
  const uint nr = ctri_ptr[L.rtl + 1] - ctri_ptr[L.rtl], nc = ctri_ptr[L.ctl + 1] - ctri_ptr[L.ctl];
  cplx *D = pool + L.off;
  if (threadIdx.x == 0) qn = 0;
  __syncthreads();
  for (uint p = threadIdx.x; p < nr * nc; p += NEAR_THREADS) {
    const uint t = rt[p / nc], s = ct[p % nc];
    const uint tv[3] = {tri[3 * (size_t)t], tri[3 * (size_t)t + 1], tri[3 * (size_t)t + 2]};
    const uint sv[3] = {tri[3 * (size_t)s], tri[3 * (size_t)s + 1], tri[3 * (size_t)s + 2]};
    bool touch = false;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) touch |= (tv[i] == sv[j]);
    if (touch) {
      uint slot = atomicAdd(&qn, 1u);
      if (slot < NEAR_QCAP) queue[slot] = make_uint2(t, s);
      else atomicExch(err, 1u);
      continue;
    }
    double xp[NQ][3];
#pragma unroll
    for (int a = 0; a < NQ; a++)
#pragma unroll
      for (int d = 0; d < 3; d++) xp[a][d] = qp[((size_t)t * NQ + a) * 3 + d];
    cplx loc[3][3];
    regular_pair<NQ, CPLXK>(xp, qp + (size_t)s * NQ * 3, R, kr, ki, loc);
    const double sc = g[t] * g[s] * INV4PI;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      uint pr = pos[tv[i]] - L.r0;
      if (pr >= L.m) continue;
#pragma unroll
      for (int j = 0; j < 3; j++) {
        uint pc = cposv[sv[j]] - L.c0;
        if (pc < L.n) red_add(D + (size_t)pc * L.m + pr, make_double2(loc[i][j].x * sc, loc[i][j].y * sc));
      }
    }
  }
  __syncthreads();
  const uint nq = min(qn, (uint)NEAR_QCAP);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint e = warp; e < nq; e += NEAR_THREADS / 32) {
    const uint t = queue[e].x, s = queue[e].y;
    uint tv[3] = {tri[3 * (size_t)t], tri[3 * (size_t)t + 1], tri[3 * (size_t)t + 2]};
    uint sv[3] = {tri[3 * (size_t)s], tri[3 * (size_t)s + 1], tri[3 * (size_t)s + 2]};
    // shared vertices to the front, same slot in both (select_quadrature_singquad2d)
    uint tp[3], sp[3], c = 0;
    bool tu[3] = {false, false, false}, su[3] = {false, false, false};
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        if (tv[i] == sv[j]) { tp[c] = i; sp[c] = j; tu[i] = su[j] = true; c++; break; }
    uint a = c, b = c;
    for (int i = 0; i < 3; i++) { if (!tu[i]) tp[a++] = i; if (!su[i]) sp[b++] = i; }
    uint rv[3] = {tv[tp[0]], tv[tp[1]], tv[tp[2]]}, cv[3] = {sv[sp[0]], sv[sp[1]], sv[sp[2]]};
    double V[6][3];
#pragma unroll
    for (int v = 0; v < 3; v++)
#pragma unroll
      for (int d = 0; d < 3; d++) { V[v][d] = x[3 * (size_t)rv[v] + d]; V[3 + v][d] = x[3 * (size_t)cv[v] + d]; }
    double acc[18];
    singular_pair_warp<CPLXK>(V, tab, c, kr, ki, lane, acc);
    if (lane == 0) {
      const double sc = g[t] * g[s] * INV4PI;
      for (int i = 0; i < 3; i++) {
        uint pr = pos[rv[i]] - L.r0;
        if (pr >= L.m) continue;
        for (int j = 0; j < 3; j++) {
          uint pc = cposv[cv[j]] - L.c0;
          if (pc < L.n) red_add(D + (size_t)pc * L.m + pr, make_double2(acc[2 * (3 * i + j)] * sc, acc[2 * (3 * i + j) + 1] * sc));
        }
      }
    }
  }
}

// Galerkin entry Z[vi, vj] for well separated vertices (every triangle pair disjoint):
// sum over the two triangle fans with the regular rule (bem->nearfield_far)
template <int NQ, bool CPLXK>
__device__ cplx far_entry(uint vi, uint vj, const RegRule &R, const double *__restrict__ qp,
                          const double *__restrict__ g, const uint *__restrict__ v2t_ptr,
                          const uint *__restrict__ v2t, double kr, double ki) {
  cplx sum = make_double2(0.0, 0.0);
  for (uint a = v2t_ptr[vi]; a < v2t_ptr[vi + 1]; a++) {
    const uint t = v2t[a] >> 2, li = v2t[a] & 3;
    double xp[NQ][3];
#pragma unroll
    for (int p = 0; p < NQ; p++)
#pragma unroll
      for (int d = 0; d < 3; d++) xp[p][d] = qp[((size_t)t * NQ + p) * 3 + d];
    for (uint b = v2t_ptr[vj]; b < v2t_ptr[vj + 1]; b++) {
      const uint s = v2t[b] >> 2, lj = v2t[b] & 3;
      const double *yq = qp + (size_t)s * NQ * 3;
      cplx acc = make_double2(0.0, 0.0);
#pragma unroll
      for (int p = 0; p < NQ; p++) {
        cplx ap = make_double2(0.0, 0.0);
#pragma unroll
        for (int q = 0; q < NQ; q++) {
          double dx = xp[p][0] - yq[3 * q], dy = xp[p][1] - yq[3 * q + 1], dz = xp[p][2] - yq[3 * q + 2];
          cplx kv = helmholtz<CPLXK>(dx * dx + dy * dy + dz * dz, kr, ki);
          double wq = R.wphi[q][lj];
          ap.x = fma(wq, kv.x, ap.x);
          ap.y = fma(wq, kv.y, ap.y);
        }
        double wp = R.wphi[p][li];
        acc.x = fma(wp, ap.x, acc.x);
        acc.y = fma(wp, ap.y, acc.y);
      }
      double sc = g[t] * g[s] * INV4PI;
      sum.x = fma(acc.x, sc, sum.x);
      sum.y = fma(acc.y, sc, sum.y);
    }
  }
  return sum;
}

constexpr int ACA_THREADS = 256;

// block-wide argmax of (value, index); result broadcast to all threads
__device__ void block_argmax(double &val, uint &idx, double *sv, uint *si) {
  for (int off = 16; off > 0; off >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, val, off);
    uint oi = __shfl_xor_sync(0xffffffffu, idx, off);
    if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) { sv[warp] = val; si[warp] = idx; }
  __syncthreads();
  val = sv[0]; idx = si[0];
  for (int w = 1; w < ACA_THREADS / 32; w++)
    if (sv[w] > val || (sv[w] == val && si[w] < idx)) { val = sv[w]; idx = si[w]; }
}
__device__ double block_sum(double v, double *sv) {
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sv[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < ACA_THREADS / 32; w++) s += sv[w];
  return s;
}

// Partial-pivot ACA of one admissible leaf by a CTA (aca.h:114-117 decomp_partialaca_rkmatrix); the matrix
// entries come from `entry(row vertex, column vertex)`: the far-field quadrature (k_aca) or an explicit
// dense matrix (k_aca_dense: re-compression after dense-expansion H-arithmetic, h2compat_h.cu).
template <class Entry>
__device__ __forceinline__ void aca_leaf(HLeafDev *leaves, const uint lid, const Entry &entry,
                                         const uint *__restrict__ perm, double accur, cplx *pool, unsigned char *flags,
                                         uint *capped) {
  __shared__ double s_val[ACA_THREADS / 32];
  __shared__ uint s_idx[ACA_THREADS / 32];
  const HLeafDev L = leaves[lid];
  const uint m = L.m, n = L.n, kcap = L.kcap;
  cplx *U = pool + L.off, *V = U + (size_t)m * kcap;
  unsigned char *rused = flags + L.foff, *cused = rused + m;
  const uint *ridx = perm + L.r0, *cidx = perm + L.c0;
  for (uint i = threadIdx.x; i < m + n; i += ACA_THREADS) rused[i] = 0;
  __syncthreads();
  uint k = 0, ik = 0;
  double start = 0.0;
  bool met = false;   // the loop ended because the accuracy was reached or the block was exhausted
  while (k < kcap) {
    // residual row ik -> V[:, k]
    const uint vi = ridx[ik];
    double best = -1.0;
    uint bj = 0xffffffffu;
    for (uint j = threadIdx.x; j < n; j += ACA_THREADS) {
      cplx r = entry(vi, cidx[j]);
      for (uint l = 0; l < k; l++) cfms(r, U[(size_t)l * m + ik], V[(size_t)l * n + j]);
      V[(size_t)k * n + j] = r;
      double a = r.x * r.x + r.y * r.y;
      if (!cused[j] && a > best) { best = a; bj = j; }
    }
    if (threadIdx.x == 0) rused[ik] = 1;
    block_argmax(best, bj, s_val, s_idx);
    if (!(best > 0.0)) {
      // zero residual row: continue with the first unused row, if any
      double nb = -1.0;
      uint nx = 0xffffffffu;
      for (uint i = threadIdx.x; i < m; i += ACA_THREADS)
        if (!rused[i] && i < nx) { nx = i; nb = 1.0; }
      // argmax breaks ties towards the smaller index -> picks the smallest unused row
      block_argmax(nb, nx, s_val, s_idx);
      if (!(nb > 0.0)) { met = true; break; }
      ik = nx;
      continue;
    }
    const uint jk = bj;
    const cplx piv = cinv(V[(size_t)k * n + jk]);
    __syncthreads();
    double nv2 = 0.0;
    for (uint j = threadIdx.x; j < n; j += ACA_THREADS) {
      cplx v = cmul(V[(size_t)k * n + j], piv);
      V[(size_t)k * n + j] = v;
      nv2 += v.x * v.x + v.y * v.y;
    }
    if (threadIdx.x == 0) cused[jk] = 1;
    // residual column jk -> U[:, k]
    const uint vj = cidx[jk];
    double nu2 = 0.0;
    best = -1.0;
    uint bi = 0xffffffffu;
    __syncthreads();
    for (uint i = threadIdx.x; i < m; i += ACA_THREADS) {
      cplx c = entry(ridx[i], vj);
      for (uint l = 0; l < k; l++) cfms(c, U[(size_t)l * m + i], V[(size_t)l * n + jk]);
      U[(size_t)k * m + i] = c;
      double a = c.x * c.x + c.y * c.y;
      nu2 += a;
      if (!rused[i] && a > best) { best = a; bi = i; }
    }
    nu2 = block_sum(nu2, s_val);
    nv2 = block_sum(nv2, s_val);
    block_argmax(best, bi, s_val, s_idx);
    const double e = sqrt(nu2) * sqrt(nv2);
    if (k == 0) start = e;
    k++;
    if (e <= accur * start) { met = true; break; }
    if (bi == 0xffffffffu) { met = true; break; }
    ik = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    leaves[lid].k = k;
    // rank capacity reached before eps_aca was: H2Lib's ACA would have kept growing the rank.  A full-rank
    // factorisation (k == min(m, n)) is exact, not capped.
    if (!met && k == kcap && kcap < min(m, n)) atomicAdd(capped, 1u);
  }
}

template <int NQ, bool CPLXK>
struct QuadEntry {
  const RegRule &R;
  const double *__restrict__ qp, *__restrict__ g;
  const uint *__restrict__ v2t_ptr, *__restrict__ v2t;
  double kr, ki;
  __device__ __forceinline__ cplx operator()(uint vi, uint vj) const {
    return far_entry<NQ, CPLXK>(vi, vj, R, qp, g, v2t_ptr, v2t, kr, ki);
  }
};
template <int NQ, bool CPLXK>
__global__ void __launch_bounds__(ACA_THREADS)
k_aca(HLeafDev *leaves, const uint *__restrict__ far_ids, const __grid_constant__ RegRule R,
      const double *__restrict__ qp, const double *__restrict__ g, const uint *__restrict__ perm,
      const uint *__restrict__ v2t_ptr, const uint *__restrict__ v2t, double kr, double ki, double accur,
      cplx *pool, unsigned char *flags, uint *capped) {
  const QuadEntry<NQ, CPLXK> entry{R, qp, g, v2t_ptr, v2t, kr, ki};
  aca_leaf(leaves, far_ids[blockIdx.x], entry, perm, accur, pool, flags, capped);
}

struct DenseEntry {
  const cplx *__restrict__ Z;
  size_t ld;
  __device__ __forceinline__ cplx operator()(uint vi, uint vj) const { return Z[(size_t)vj * ld + vi]; }
};
__global__ void __launch_bounds__(ACA_THREADS)
k_aca_dense(HLeafDev *leaves, const uint *__restrict__ far_ids, const cplx *__restrict__ Z, size_t ld,
            const uint *__restrict__ perm, double accur, cplx *pool, unsigned char *flags, uint *capped) {
  const DenseEntry entry{Z, ld};
  aca_leaf(leaves, far_ids[blockIdx.x], entry, perm, accur, pool, flags, capped);
}
// inadmissible leaves of the re-compression: the sub-block itself
__global__ void k_gather_dense(const HLeafDev *__restrict__ leaves, const uint *__restrict__ near_ids,
                               const cplx *__restrict__ Z, size_t ld, const uint *__restrict__ perm, cplx *pool) {
  const HLeafDev L = leaves[near_ids[blockIdx.x]];
  for (size_t e = threadIdx.x; e < (size_t)L.m * L.n; e += blockDim.x)
    pool[L.off + e] = Z[(size_t)perm[L.c0 + e / L.m] * ld + perm[L.r0 + e % L.m]];
}

// ---- matvec: y_perm += alpha * Z * x_perm over the leaves --------------------------------

__global__ void k_permute_in(uint n, uint nrhs, const uint *__restrict__ perm, const cplx *__restrict__ x,
                             cplx *xp) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * nrhs) return;
  uint r = i / n, p = i % n;
  xp[i] = x[(size_t)r * n + perm[p]];
}
__global__ void k_permute_out(uint n, uint nrhs, const uint *__restrict__ perm, cplx alpha, const cplx *__restrict__ yp,
                              cplx *y) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * nrhs) return;
  uint r = i / n, p = i % n;
  cplx v = cmul(alpha, yp[i]);
  cplx *d = y + (size_t)r * n + perm[p];
  d->x += v.x;
  d->y += v.y;
}

// ---- batched matvec (several right-hand sides) ----------------------------------------------------
// With nrhs vectors a leaf is a small GEMM.  Two passes, no atomics:
//   1. every low-rank leaf:  t = V^T x|cols   (k x nrhs, into a scratch)
//   2. every band of 16 rows (cluster order): y|band = sum over the leaves that intersect the band of
//      A|band x|cols (dense) or U|band t (low rank); the band's CTA owns its rows of y.
// A band's work is a list of rounds of up to 16 contraction indices: first the dense leaves that
// intersect it (16 columns per round, static), then its (low-rank leaf, rank index) items 16 at a
// time (rebuilt after every assembly because the ranks change).  Per round the operands are gathered
// into shared memory with cp.async, software-pipelined (round c is contracted while round c+1 is in
// flight and the descriptors of round c+2 are read), and contracted on the FP64 tensor pipe: the
// band is a 16 x K x (16 TC) complex GEMM, 3M product (X = ArBr, Y = AiBi, W = (Ar+Ai)(Br+Bi)) on
// DMMA.8x8x4 -- a tenth of the instructions of the DFMA formulation, which was issue/latency bound
// (ncu: FP64 pipe 41 % active, long-scoreboard stalls).
constexpr int BAND = 16, AS_LDB = BAND + 2;

// The batched kernels keep their work vectors DOF-major, [dof][rhs]: the 16 x (16 TC) operand tile of a
// round is then 16 contiguous runs of 16 TC right-hand sides instead of 16 TC runs of 16 DOFs.
__global__ void k_permute_in_t(uint n, uint nrhs, const uint *__restrict__ perm, const cplx *__restrict__ x, cplx *xp) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // i = pos * nrhs + r
  if (i >= (size_t)n * nrhs) return;
  const uint r = i % nrhs, p = i / nrhs;
  xp[i] = x[(size_t)r * n + perm[p]];
}
__global__ void k_permute_out_t(uint n, uint nrhs, const uint *__restrict__ perm, cplx alpha, const cplx *__restrict__ yp,
                                cplx *y) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * nrhs) return;
  const uint r = i % nrhs, p = i / nrhs;
  const cplx v = cmul(alpha, yp[i]);
  cplx *d = y + (size_t)r * n + perm[p];
  d->x += v.x;
  d->y += v.y;
}

__device__ __forceinline__ void hm_cp_async16(void *smem, const void *gmem, bool valid) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  int sz = valid ? 16 : 0;   // src-size 0: zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void hm_dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// CTA = 2 warps; warp w owns the right-hand sides [8 TC w, 8 TC (w+1)) of the pass (TC n-tiles of 8).
// TMODE = false: CTA = band of 16 rows of y.   TMODE = true: CTA = low-rank leaf, rows = rank indices,
// t = V^T x|cols written to the scratch ([rank index][rhs]); the rounds are 16 columns of the leaf.
template <int TC, bool TMODE, int NW = 2>
__global__ void __launch_bounds__(32 * NW)
k_hmv_gemm(const uint *__restrict__ band_ptr, const HMat::DRound *__restrict__ dround, const uint *__restrict__ bfar_ptr,
           const HMat::FarItem *__restrict__ bfar, const uint *__restrict__ far_ids, const HLeafDev *__restrict__ leaves,
           const unsigned long long *__restrict__ toff, const cplx *__restrict__ pool, cplx *T,
           const cplx *__restrict__ xp, cplx *yp, uint ntot, uint nrhs) {
  constexpr int BT_THREADS = 32 * NW;        // NW warps, each TC n-tiles of 8 right-hand sides
  constexpr int RP = 8 * TC * NW, XS_LD = RP + 2;   // leading dimensions == 2 (mod 8): conflict-free fragment loads
  extern __shared__ __align__(16) unsigned char hg_smem[];
  cplx (*As)[16][AS_LDB] = reinterpret_cast<cplx (*)[16][AS_LDB]>(hg_smem);                                  // [buf][contraction index][row]
  cplx (*Xs)[16][XS_LD] = reinterpret_cast<cplx (*)[16][XS_LD]>(hg_smem + sizeof(cplx) * 2 * 16 * AS_LDB);   // [buf][contraction index][rhs]
  __shared__ HMat::FarItem Is[3][16];     // per contraction index: U/A column, row range, x or t index
  __shared__ uint s_cnt[3], s_far[3];
  const uint tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint fr = lane >> 2, fc = lane & 3;
  uint b0 = 0, db = 0, nd = 0, fb = 0, fe = 0, R = 0, lk = 0, ln = 0, lc0 = 0;
  size_t voff = 0;
  unsigned long long t0 = 0;
  if (!TMODE) {
    b0 = blockIdx.x * BAND;
    db = band_ptr[blockIdx.x]; nd = band_ptr[blockIdx.x + 1] - db;
    fb = bfar_ptr[blockIdx.x]; fe = bfar_ptr[blockIdx.x + 1];
    R = nd + (fe - fb + 15) / 16;
  } else {
    const uint id = far_ids[blockIdx.x];
    const HLeafDev L = leaves[id];
    lk = L.k; ln = L.n; lc0 = L.c0;
    voff = L.off + (size_t)L.m * L.kcap;
    t0 = toff[id];
    R = (ln + 15) / 16;
    if (!lk) return;
  }

  for (uint l0 = 0; l0 < (TMODE ? lk : 1u); l0 += 16)    // rank indices 16 at a time (TMODE), once otherwise
  for (uint r0 = 0; r0 < nrhs; r0 += RP) {
    const uint kk = TMODE ? min(16u, lk - l0) : 16u;
    // descriptors of round c -> registers of threads 0..15 (one contraction index each)
    auto fetch_items = [&](uint c, HMat::FarItem &it, uint &cnt) {
      if (TMODE) {
        cnt = min(16u, ln - c * 16);
        it.u_off = voff + (size_t)l0 * ln + c * 16 + tid;   // V[j, l0]: rows (rank indices) at stride n
        it.t_idx = lc0 + c * 16 + tid;
        it.r0 = 0; it.m = kk;
      } else if (c < nd) {
        const HMat::DRound D = dround[db + c];
        cnt = D.cnt;
        it.u_off = D.a_off + (size_t)tid * D.m;
        it.t_idx = D.x_off + tid;
        it.r0 = D.r0; it.m = D.m;
      } else {
        const uint f0 = fb + (c - nd) * 16;
        cnt = min(16u, fe - f0);
        if (tid < cnt) it = bfar[f0 + tid];
      }
    };
    auto store_items = [&](uint c, const HMat::FarItem &it, uint cnt) {
      if (tid < 16) Is[c % 3][tid] = it;
      if (tid == 0) { s_cnt[c % 3] = cnt; s_far[c % 3] = (!TMODE && c >= nd) ? 1u : 0u; }
    };
    // operands of round c -> shared-memory buffers c & 1 (indices beyond cnt are zero-filled)
    auto issue = [&](uint c) {
      const uint ib = c % 3, buf = c & 1, cnt = s_cnt[ib];
      const bool far = s_far[ib] != 0;
      if (TMODE) {
        for (uint e = tid; e < 16 * BAND; e += BT_THREADS) {
          const uint j = e % 16, i = e / 16;   // consecutive threads walk a row of V^T (contiguous in j)
          const bool ok = j < cnt && i < kk;
          hm_cp_async16(&As[buf][j][i], ok ? pool + Is[ib][j].u_off + (size_t)i * ln : pool, ok);
        }
      } else {
        for (uint e = tid; e < 16 * BAND; e += BT_THREADS) {
          const uint i = e % BAND, j = e / BAND;
          const HMat::FarItem it = Is[ib][j];
          const long long row = (long long)b0 + i - (long long)it.r0;
          const bool ok = j < cnt && row >= 0 && row < (long long)it.m;
          hm_cp_async16(&As[buf][j][i], ok ? pool + it.u_off + row : pool, ok);
        }
      }
      if (!far) {
        for (uint e = tid; e < 16 * RP; e += BT_THREADS) {
          const uint r = e % RP, j = e / RP;
          const bool ok = j < cnt && r0 + r < nrhs;
          hm_cp_async16(&Xs[buf][j][r], ok ? xp + (size_t)Is[ib][j].t_idx * nrhs + r0 + r : xp, ok);
        }
      } else {
        for (uint e = tid; e < 16 * RP; e += BT_THREADS) {
          const uint r = e % RP, j = e / RP;
          const bool ok = j < cnt && r0 + r < nrhs;
          hm_cp_async16(&Xs[buf][j][r], ok ? T + (size_t)Is[ib][j].t_idx * nrhs + r0 + r : T, ok);
        }
      }
    };
    double X[2][TC][2], Y[2][TC][2], W[2][TC][2];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
      for (int q = 0; q < TC; q++)
#pragma unroll
        for (int e = 0; e < 2; e++) X[a][q][e] = Y[a][q][e] = W[a][q][e] = 0.0;
    HMat::FarItem it;
    uint cnt = 0;
    __syncthreads();   // previous pass is done with every buffer
    if (R > 0) { fetch_items(0, it, cnt); store_items(0, it, cnt); }
    if (R > 1) { fetch_items(1, it, cnt); store_items(1, it, cnt); }
    __syncthreads();
    if (R > 0) issue(0);
    asm volatile("cp.async.commit_group;\n" ::);
    for (uint c = 0; c < R; c++) {
      __syncthreads();   // items of round c+1 visible; round c-1 contracted: its buffers are free
      if (c + 1 < R) issue(c + 1);
      asm volatile("cp.async.commit_group;\n" ::);
      if (c + 2 < R) fetch_items(c + 2, it, cnt);
      asm volatile("cp.async.wait_group 1;\n" ::);
      __syncthreads();   // operands of round c have landed
      const uint buf = c & 1, jn = (s_cnt[c % 3] + 3) / 4;   // k4 steps (zero-filled tail)
      const int nmi = (TMODE && kk <= 8) ? 1 : 2;             // ranks <= 8: the second row tile is all zero
      for (uint ks = 0; ks < jn; ks++) {
        double are[2], aim[2], asu[2], bre[TC], bim[TC], bsu[TC];
#pragma unroll
        for (int mi = 0; mi < 2; mi++) {
          const cplx v = As[buf][ks * 4 + fc][mi * 8 + fr];
          are[mi] = v.x; aim[mi] = v.y; asu[mi] = v.x + v.y;
        }
#pragma unroll
        for (int ni = 0; ni < TC; ni++) {
          const cplx v = Xs[buf][ks * 4 + fc][(warp * TC + ni) * 8 + fr];
          bre[ni] = v.x; bim[ni] = v.y; bsu[ni] = v.x + v.y;
        }
#pragma unroll
        for (int mi = 0; mi < 2; mi++)
          if (mi < nmi)
#pragma unroll
            for (int ni = 0; ni < TC; ni++) {
              hm_dmma(X[mi][ni][0], X[mi][ni][1], are[mi], bre[ni]);
              hm_dmma(Y[mi][ni][0], Y[mi][ni][1], aim[mi], bim[ni]);
              hm_dmma(W[mi][ni][0], W[mi][ni][1], asu[mi], bsu[ni]);
            }
      }
      if (c + 2 < R) store_items(c + 2, it, cnt);   // slot (c+2) % 3 = (c-1) % 3: round c-1 is finished
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
#pragma unroll
    for (int mi = 0; mi < 2; mi++)
#pragma unroll
      for (int ni = 0; ni < TC; ni++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const uint i = mi * 8 + fr, r = r0 + (warp * TC + ni) * 8 + fc * 2 + e;
          const cplx v = make_double2(X[mi][ni][e] - Y[mi][ni][e], W[mi][ni][e] - X[mi][ni][e] - Y[mi][ni][e]);
          if (TMODE) {
            if (i < kk && r < nrhs) T[(size_t)(t0 + l0 + i) * nrhs + r] = v;
          } else {
            if (b0 + i < ntot && r < nrhs) yp[(size_t)(b0 + i) * nrhs + r] = v;
          }
        }
  }
}


// ---- stacked t-pass -----------------------------------------------------------------------------
// CTA = band of 16 rows of the scratch (consecutive rank indices of the low-rank leaves of ONE column
// cluster), NW warps x TC n-tiles of 8 right-hand sides per pass.  Round = 16 columns of the cluster.
// All (pass, round) pairs form one software pipeline (the next pair's operands are in flight while the
// current one is contracted), so column clusters of 16..32 columns -- most of them -- do not drain it
// after every pass.  No descriptors are read inside the loop: a row's operand address is its V column
// plus the column index.
constexpr int TS_LDA = 20;   // As[row][col]: conflict-free for the cp.async writes and the fragment loads
template <int NW, int TC>
__global__ void __launch_bounds__(32 * NW)
k_hmv_tstack(const HMat::TBand *__restrict__ tband, const unsigned long long *__restrict__ vrow,
             const cplx *__restrict__ pool, cplx *T, const cplx *__restrict__ xp, uint nrhs) {
  constexpr int NT = 32 * NW, RP = 8 * TC * NW, XS_LD = RP + 2;
  extern __shared__ __align__(16) unsigned char ts_smem[];
  cplx (*As)[16][TS_LDA] = reinterpret_cast<cplx (*)[16][TS_LDA]>(ts_smem);                                  // [buf][row][col]
  cplx (*Xs)[16][XS_LD] = reinterpret_cast<cplx (*)[16][XS_LD]>(ts_smem + sizeof(cplx) * 2 * 16 * TS_LDA);   // [buf][col][rhs]
  __shared__ unsigned long long s_row[16];
  const uint tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint fr = lane >> 2, fc = lane & 3;
  const HMat::TBand B = tband[blockIdx.x];
  if (tid < 16) s_row[tid] = vrow[(size_t)blockIdx.x * 16 + tid];
  __syncthreads();
  const uint R = (B.n + 15) / 16, npass = (nrhs + RP - 1) / RP, total = R * npass;
  const int nmi = B.rows <= 8 ? 1 : 2;
  auto issue = [&](uint s) {
    const uint c = s % R, r0 = (s / R) * RP, buf = s & 1, cnt = min(16u, B.n - c * 16);
    for (uint e = tid; e < 256; e += NT) {
      const uint j = e & 15, i = e >> 4;
      const unsigned long long base = s_row[i];
      const bool ok = j < cnt && base != ~0ull;
      hm_cp_async16(&As[buf][i][j], ok ? pool + base + c * 16 + j : pool, ok);
    }
    for (uint e = tid; e < 16 * RP; e += NT) {
      const uint r = e % RP, j = e / RP;
      const bool ok = j < cnt && r0 + r < nrhs;
      hm_cp_async16(&Xs[buf][j][r], ok ? xp + (size_t)(B.c0 + c * 16 + j) * nrhs + r0 + r : xp, ok);
    }
  };
  double X[2][TC][2], Y[2][TC][2], W[2][TC][2];
#pragma unroll
  for (int a = 0; a < 2; a++)
#pragma unroll
    for (int q = 0; q < TC; q++)
#pragma unroll
      for (int e = 0; e < 2; e++) X[a][q][e] = Y[a][q][e] = W[a][q][e] = 0.0;
  issue(0);
  asm volatile("cp.async.commit_group;\n" ::);
  for (uint s = 0; s < total; s++) {
    __syncthreads();   // pair s-1 is contracted: its buffers are free
    if (s + 1 < total) issue(s + 1);
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 1;\n" ::);
    __syncthreads();   // operands of pair s have landed
    const uint c = s % R, buf = s & 1, jn = (min(16u, B.n - c * 16) + 3) / 4;
    for (uint ks = 0; ks < jn; ks++) {
      double are[2], aim[2], asu[2], bre[TC], bim[TC], bsu[TC];
#pragma unroll
      for (int mi = 0; mi < 2; mi++) {
        const cplx v = As[buf][mi * 8 + fr][ks * 4 + fc];
        are[mi] = v.x; aim[mi] = v.y; asu[mi] = v.x + v.y;
      }
#pragma unroll
      for (int ni = 0; ni < TC; ni++) {
        const cplx v = Xs[buf][ks * 4 + fc][(warp * TC + ni) * 8 + fr];
        bre[ni] = v.x; bim[ni] = v.y; bsu[ni] = v.x + v.y;
      }
#pragma unroll
      for (int mi = 0; mi < 2; mi++)
        if (mi < nmi)
#pragma unroll
          for (int ni = 0; ni < TC; ni++) {
            hm_dmma(X[mi][ni][0], X[mi][ni][1], are[mi], bre[ni]);
            hm_dmma(Y[mi][ni][0], Y[mi][ni][1], aim[mi], bim[ni]);
            hm_dmma(W[mi][ni][0], W[mi][ni][1], asu[mi], bsu[ni]);
          }
    }
    if (c + 1 == R) {   // last round of this pass: write the band's t rows, restart the accumulators
      const uint r0 = (s / R) * RP;
#pragma unroll
      for (int mi = 0; mi < 2; mi++)
#pragma unroll
        for (int ni = 0; ni < TC; ni++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const uint i = mi * 8 + fr, r = r0 + (warp * TC + ni) * 8 + fc * 2 + e;
            if (i < B.rows && r < nrhs)
              T[((size_t)blockIdx.x * 16 + i) * nrhs + r] =
                  make_double2(X[mi][ni][e] - Y[mi][ni][e], W[mi][ni][e] - X[mi][ni][e] - Y[mi][ni][e]);
            X[mi][ni][e] = Y[mi][ni][e] = W[mi][ni][e] = 0.0;
          }
    }
  }
  asm volatile("cp.async.wait_group 0;\n" ::);
}

// ---- matvec with 1-3 vectors: HBM-bound, so the kernels only have to read every leaf once, coalesced, with
// enough independent loads in flight per warp to cover the DRAM latency ----
// t-pass: one warp per low-rank leaf, lanes along the columns (V[:, l] is contiguous).  LC rank indices are
// contracted together: LC independent 16-byte loads per lane before the first use (a leaf is only k x n =
// 5..10 x 16..64 entries, so one load per lane and rank index would leave the warp waiting for DRAM k times),
// then LC interleaved shuffle reductions.
constexpr int V1_THREADS = 128;
// Compiler fence for values: everything that produced v[] (the loads) is scheduled before this point and every
// use after it (one asm statement: separate ones would let a use slip between them).  Without it ptxas sinks
// each load next to its first use, which leaves two loads in flight per lane.
__device__ __forceinline__ void loads_done(cplx (&v)[4]) {
  asm volatile("" : "+d"(v[0].x), "+d"(v[0].y), "+d"(v[1].x), "+d"(v[1].y), "+d"(v[2].x), "+d"(v[2].y), "+d"(v[3].x), "+d"(v[3].y));
}
__device__ __forceinline__ void loads_done(cplx (&v)[8]) {
  asm volatile("" : "+d"(v[0].x), "+d"(v[0].y), "+d"(v[1].x), "+d"(v[1].y), "+d"(v[2].x), "+d"(v[2].y), "+d"(v[3].x), "+d"(v[3].y),
                    "+d"(v[4].x), "+d"(v[4].y), "+d"(v[5].x), "+d"(v[5].y), "+d"(v[6].x), "+d"(v[6].y), "+d"(v[7].x), "+d"(v[7].y));
}
template <int NR>
__global__ void __launch_bounds__(V1_THREADS, 4)
k_hmv1_t(uint nfar, const uint *__restrict__ far_ids, const HLeafDev *__restrict__ leaves,
         const unsigned long long *__restrict__ toff, const cplx *__restrict__ pool, const cplx *__restrict__ xp,
         cplx *T, uint ntot) {
  constexpr int LC = NR == 1 ? 8 : 4;
  const uint w = (blockIdx.x * V1_THREADS + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= nfar) return;
  const uint id = far_ids[w];
  const HLeafDev L = leaves[id];
  const cplx *V = pool + L.off + (size_t)L.m * L.kcap;
  const cplx *x = xp + L.c0;
  cplx *Tl = T + toff[id] * NR;
  for (uint l0 = 0; l0 < L.k; l0 += LC) {
    cplx s[LC][NR];
#pragma unroll
    for (int c = 0; c < LC; c++)
#pragma unroll
      for (int r = 0; r < NR; r++) s[c][r] = make_double2(0.0, 0.0);
    for (uint j = lane; j < L.n; j += 32) {
      cplx a[LC], xv[NR];
#pragma unroll
      for (int c = 0; c < LC; c++)       // clamped row: rows beyond the rank repeat the last one and are not stored
        a[c] = __ldg(V + (size_t)min(l0 + c, L.k - 1) * L.n + j);
#pragma unroll
      for (int r = 0; r < NR; r++) xv[r] = x[(size_t)r * ntot + j];
      loads_done(a);
#pragma unroll
      for (int c = 0; c < LC; c++)
#pragma unroll
        for (int r = 0; r < NR; r++) cfma(s[c][r], a[c], xv[r]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
      for (int c = 0; c < LC; c++)
#pragma unroll
        for (int r = 0; r < NR; r++) {
          s[c][r].x += __shfl_xor_sync(0xffffffffu, s[c][r].x, off);
          s[c][r].y += __shfl_xor_sync(0xffffffffu, s[c][r].y, off);
        }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < LC; c++)
        if (l0 + c < L.k) {
#pragma unroll
          for (int r = 0; r < NR; r++) Tl[(size_t)(l0 + c) * NR + r] = s[c][r];
        }
    }
  }
}

// t-pass over the stacked layout (the batched t-pass's own descriptors): the scratch rows of one column
// cluster are consecutive, 16 per band, and vrow[] holds the pool offset of every row's V column -- so a warp
// needs no leaf record at all: one band descriptor + 16 row offsets, then the columns as independent 16-byte
// loads (half-warp = one row: a row's n entries are contiguous; 8 rows x 2 column steps = 16 loads per lane
// in flight), 4-step reductions, 16 consecutive results.  Warps walk the bands with a grid stride and fetch
// the next band's descriptors before they contract the current one.
__device__ __forceinline__ void loads_done(cplx (&p)[8], cplx (&q)[8]) {
  // the real parts only: a 16-byte load delivers both parts, so ordering x orders the load
  asm volatile("" : "+d"(p[0].x), "+d"(p[1].x), "+d"(p[2].x), "+d"(p[3].x), "+d"(p[4].x), "+d"(p[5].x), "+d"(p[6].x), "+d"(p[7].x),
                    "+d"(q[0].x), "+d"(q[1].x), "+d"(q[2].x), "+d"(q[3].x), "+d"(q[4].x), "+d"(q[5].x), "+d"(q[6].x), "+d"(q[7].x));
}
__device__ __forceinline__ void loads_done(cplx (&p)[4], cplx (&q)[4]) {
  asm volatile("" : "+d"(p[0].x), "+d"(p[1].x), "+d"(p[2].x), "+d"(p[3].x), "+d"(q[0].x), "+d"(q[1].x), "+d"(q[2].x), "+d"(q[3].x));
}
template <int NR>
__global__ void __launch_bounds__(V1_THREADS, 3)
k_hmv1_tb(uint ntband, const HMat::TBand *__restrict__ tband, const unsigned long long *__restrict__ vrow,
          const cplx *__restrict__ pool, const cplx *__restrict__ xp, cplx *T, uint ntot) {
  constexpr int RU = NR == 1 ? 8 : 4;        // rows per lane and pass (register budget: RU x NR accumulators)
  const uint W = gridDim.x * (V1_THREADS / 32), lane = threadIdx.x & 31;
  uint w = (blockIdx.x * V1_THREADS + threadIdx.x) >> 5;
  if (w >= ntband) return;
  const uint i = lane & 15, h = lane >> 4;
  HMat::TBand B = tband[w];
  unsigned long long off[8];
#pragma unroll
  for (int u = 0; u < 8; u++) off[u] = __ldg(vrow + (size_t)w * 16 + h + 2 * u);
  while (true) {
    const uint wn = w + W, wc = wn < ntband ? wn : w;
    const HMat::TBand Bn = tband[wc];
    unsigned long long offn[8];
#pragma unroll
    for (int u = 0; u < 8; u++) offn[u] = __ldg(vrow + (size_t)wc * 16 + h + 2 * u);
    // padding rows read row 0 of the band (always present, a valid address) and are masked out at the end
    const unsigned long long off0 = __shfl_sync(0xffffffffu, off[0], 0);
    double msk[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      msk[u] = off[u] != ~0ull ? 1.0 : 0.0;
      if (off[u] == ~0ull) off[u] = off0;
    }
    const cplx *x = xp + B.c0;
#pragma unroll
    for (int pass = 0; pass < 8 / RU; pass++) {
      if (pass * 2 * RU >= (int)B.rows) break;
      cplx s[RU][NR];
#pragma unroll
      for (int u = 0; u < RU; u++)
#pragma unroll
        for (int r = 0; r < NR; r++) s[u][r] = make_double2(0.0, 0.0);
      for (uint j = i; j < B.n; j += 32) {
        const bool two = j + 16 < B.n;
        const uint j2 = two ? j + 16 : j;
        const double m2 = two ? 1.0 : 0.0;
        cplx a[RU], a2[RU], xv[NR], xv2[NR];
#pragma unroll
        for (int u = 0; u < RU; u++) {
          a[u] = __ldg(pool + off[pass * RU + u] + j);
          a2[u] = __ldg(pool + off[pass * RU + u] + j2);
        }
#pragma unroll
        for (int r = 0; r < NR; r++) {
          xv[r] = x[(size_t)r * ntot + j];
          const cplx t2 = x[(size_t)r * ntot + j2];
          xv2[r] = make_double2(t2.x * m2, t2.y * m2);
        }
        loads_done(a, a2);
#pragma unroll
        for (int u = 0; u < RU; u++)
#pragma unroll
          for (int r = 0; r < NR; r++) {
            cfma(s[u][r], a[u], xv[r]);
            cfma(s[u][r], a2[u], xv2[r]);
          }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1)
#pragma unroll
        for (int u = 0; u < RU; u++)
#pragma unroll
          for (int r = 0; r < NR; r++) {
            s[u][r].x += __shfl_xor_sync(0xffffffffu, s[u][r].x, o);
            s[u][r].y += __shfl_xor_sync(0xffffffffu, s[u][r].y, o);
          }
      if (i == 0) {
#pragma unroll
        for (int u = 0; u < RU; u++) {
          const uint row = h + 2 * (pass * RU + u);
          if (row < B.rows) {
#pragma unroll
            for (int r = 0; r < NR; r++)
              T[((size_t)w * 16 + row) * NR + r] =
                  make_double2(s[u][r].x * msk[pass * RU + u], s[u][r].y * msk[pass * RU + u]);
          }
        }
      }
    }
    if (wn >= ntband) break;
    w = wn;
    B = Bn;
#pragma unroll
    for (int u = 0; u < 8; u++) off[u] = offn[u];
  }
}

// y-pass: one CTA per band of 16 rows, its dense rounds and low-rank items dealt round-robin to the four
// warps (12 176 bands at the 16x16 array are not enough warps to fill the chip one warp per band); lane =
// (row, parity of the contraction index): a column of a leaf's 16 band rows is 256 contiguous bytes, two
// columns per warp load.  All loads of a round (8 per lane) and of four low-rank items are issued before
// the first use, and the next round's descriptor is fetched while the current one is contracted.
template <int NR>
__global__ void __launch_bounds__(V1_THREADS, 6)
k_hmv1_band(uint nband, const uint *__restrict__ band_ptr, const HMat::DRound *__restrict__ dround,
            const uint *__restrict__ bfar_ptr, const HMat::FarItem *__restrict__ bfar, const cplx *__restrict__ pool,
            const cplx *__restrict__ T, const cplx *__restrict__ xp, cplx *yp, uint ntot) {
  constexpr int NW = V1_THREADS / 32;
  __shared__ cplx red[NW][BAND][NR];
  const uint b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint i = lane & 15, h = lane >> 4;
  const long long row = (long long)b * BAND + i;
  const cplx zero = make_double2(0.0, 0.0);
  cplx acc[NR];
#pragma unroll
  for (int r = 0; r < NR; r++) acc[r] = zero;
  {
    const uint d1 = band_ptr[b + 1];
    uint d = band_ptr[b] + warp;
    HMat::DRound D{};
    if (d < d1) D = dround[d];
    while (d < d1) {
      const uint dn = d + NW;
      HMat::DRound Dn{};
      if (dn < d1) Dn = dround[dn];
      const long long rl = row - (long long)D.r0;
      const bool ok = rl >= 0 && rl < (long long)D.m;
      const cplx *col = pool + D.a_off + (ok ? rl : 0);
      const cplx *x = xp + D.x_off;
      // unconditional loads from clamped (always valid) addresses, masked afterwards: predicated loads are
      // scheduled next to their uses, two at a time
      cplx a[8], xv[NR][8];
      const uint jmax = D.cnt - 1;
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const uint j = min(h + 2 * u, jmax);
        a[u] = __ldg(col + (size_t)j * D.m);
#pragma unroll
        for (int r = 0; r < NR; r++) xv[r][u] = __ldg(x + (size_t)r * ntot + j);
      }
      loads_done(a);
#pragma unroll
      for (int r = 0; r < NR; r++) loads_done(xv[r]);
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const double msk = ok && h + 2 * u <= jmax ? 1.0 : 0.0;
        const cplx am = make_double2(a[u].x * msk, a[u].y * msk);
#pragma unroll
        for (int r = 0; r < NR; r++) cfma(acc[r], am, xv[r][u]);
      }
      D = Dn;
      d = dn;
    }
  }
  {
    // four items (stride 2 NW) per trip; the next trip's item records are fetched before this trip's data
    const uint f1 = bfar_ptr[b + 1];
    uint f = bfar_ptr[b] + warp * 2 + h;
    HMat::FarItem it[4], itn[4];
    if (f < f1) {
#pragma unroll
      for (int u = 0; u < 4; u++) it[u] = bfar[min(f + 2 * NW * u, f1 - 1)];
    }
    while (f < f1) {
      const uint fn = f + 8 * NW;
      if (fn < f1) {
#pragma unroll
        for (int u = 0; u < 4; u++) itn[u] = bfar[min(fn + 2 * NW * u, f1 - 1)];
      }
      cplx a[4], t[NR][4];
      double msk[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const long long rl = row - (long long)it[u].r0;
        const bool ok = f + 2 * NW * u < f1 && rl >= 0 && rl < (long long)it[u].m;
        msk[u] = ok ? 1.0 : 0.0;
        a[u] = __ldg(pool + it[u].u_off + (ok ? rl : 0));
#pragma unroll
        for (int r = 0; r < NR; r++) t[r][u] = __ldg(T + (size_t)it[u].t_idx * NR + r);
      }
      loads_done(a);
#pragma unroll
      for (int r = 0; r < NR; r++) loads_done(t[r]);
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const cplx am = make_double2(a[u].x * msk[u], a[u].y * msk[u]);
#pragma unroll
        for (int r = 0; r < NR; r++) cfma(acc[r], am, t[r][u]);
      }
      f = fn;
#pragma unroll
      for (int u = 0; u < 4; u++) it[u] = itn[u];
    }
  }
#pragma unroll
  for (int r = 0; r < NR; r++) {
    acc[r].x += __shfl_xor_sync(0xffffffffu, acc[r].x, 16);
    acc[r].y += __shfl_xor_sync(0xffffffffu, acc[r].y, 16);
    if (h == 0) red[warp][i][r] = acc[r];
  }
  __syncthreads();
  if (warp == 0 && h == 0 && row < (long long)ntot) {
#pragma unroll
    for (int r = 0; r < NR; r++) {
      cplx sum = red[0][i][r];
#pragma unroll
      for (int q = 1; q < NW; q++) { sum.x += red[q][i][r].x; sum.y += red[q][i][r].y; }
      yp[(size_t)r * ntot + row] = sum;
    }
  }
}

__global__ void k_expand(const HLeafDev *__restrict__ leaves, const cplx *__restrict__ pool,
                         const uint *__restrict__ perm, cplx *Z, size_t ld) {
  const HLeafDev L = leaves[blockIdx.x];
  const cplx *base = pool + L.off;
  for (size_t e = threadIdx.x; e < (size_t)L.m * L.n; e += blockDim.x) {
    uint i = e % L.m, j = e / L.m;
    cplx v;
    if (!L.adm) v = base[e];
    else {
      v = make_double2(0.0, 0.0);
      const cplx *U = base, *V = base + (size_t)L.m * L.kcap;
      for (uint l = 0; l < L.k; l++) cfma(v, U[(size_t)l * L.m + i], V[(size_t)l * L.n + j]);
    }
    Z[(size_t)perm[L.c0 + j] * ld + perm[L.r0 + i]] = v;
  }
}

__global__ void k_untile(const HLeafDev *__restrict__ leaves, const cplx *__restrict__ pool, cplx *out, size_t ld) {
  const HLeafDev L = leaves[blockIdx.x];
  for (size_t e = threadIdx.x; e < (size_t)L.m * L.n; e += blockDim.x)
    out[(size_t)(L.c0 + e / L.m) * ld + L.r0 + e % L.m] = pool[L.off + e];
}

template <typename T>
int to_device(T *&d, const std::vector<T> &h) {
  CNLD_CK(cudaMalloc(&d, sizeof(T) * (h.size() ? h.size() : 1)));
  if (!h.empty()) CNLD_CK(cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
  return 0;
}
}  // namespace

int hmat_create(HMat &H, const BemDev &bem, const double *hx, const uint *htri, uint clf, double eta, int admis,
                bool strict, uint kmax) {
  build_vertex_cluster_tree(bem.nv, bem.nt, hx, htri, clf, H.tree);
  build_block_leaves(H.tree, eta, (Admis)admis, strict, H.blocks);
  std::vector<LeafRange> ranges;
  for (const BlockLeaf &B : H.blocks) {
    const ClusterNode &Rn = H.tree.node[B.rc], &Cn = H.tree.node[B.cc];
    ranges.push_back({Rn.start, Rn.size, Cn.start, Cn.size, B.admissible});
  }
  return hmat_create_from_ranges(H, bem, htri, H.tree.perm, ranges, kmax, hx);
}

// leaves given as (row range, column range) of the permutation `perm` (cluster order -> DOF)
// Translation classes of the leaves.  The mesh is ncomp translated copies of one component (periodic.h) and
// the kernel depends on x - y only, so a leaf whose row and column vertices are the translates -- by one
// common lattice vector, in the same cluster order -- of another leaf's holds the same block.  Key of a
// leaf: for every row and column vertex (local vertex id, quantised offset of its component relative to
// the component of the first row vertex).  Leaves are matched by a 64-bit hash of the key and then
// compared element by element; rep[b] = first leaf with the same key.
static void leaf_translation_classes(const BemDev &bem, const double *hx, const std::vector<uint> &perm,
                                     const std::vector<LeafRange> &ranges, std::vector<uint> &rep) {
  rep.resize(ranges.size());
  for (size_t b = 0; b < ranges.size(); b++) rep[b] = (uint)b;
  if (!hx || !bem.periodic || !bem.use_periodic || bem.ncomp < 2) return;
  const uint nvc = bem.plan[1].blk, nc = bem.ncomp;
  double maxabs = 0.0;
  for (size_t i = 0; i < 3 * (size_t)bem.nv; i++) maxabs = std::max(maxabs, std::fabs(hx[i]));
  const double quantum = 1e-7 * (maxabs > 0.0 ? maxabs : 1.0);   // lattice offsets differ by far more, rounding by far less
  std::vector<long long> q(3 * (size_t)nc);
  for (uint c = 0; c < nc; c++)
    for (int d = 0; d < 3; d++) q[3 * c + d] = std::llround((hx[3 * (size_t)(c * nvc) + d] - hx[d]) / quantum);
  auto key_elem = [&](uint v, uint a0, int d) -> long long {   // d = 0..2: offset components, 3: local id
    const uint c = v / nvc;
    return d == 3 ? (long long)(v % nvc) : q[3 * c + d] - q[3 * a0 + d];
  };
  auto hash_leaf = [&](const LeafRange &B) -> unsigned long long {
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&](long long x) { h = (h ^ (unsigned long long)x) * 1099511628211ull; };
    mix(B.m); mix(B.n); mix(B.admissible ? 1 : 0);
    const uint a0 = perm[B.r0] / nvc;
    for (uint i = 0; i < B.m; i++) for (int d = 0; d < 4; d++) mix(key_elem(perm[B.r0 + i], a0, d));
    for (uint i = 0; i < B.n; i++) for (int d = 0; d < 4; d++) mix(key_elem(perm[B.c0 + i], a0, d));
    return h;
  };
  auto same = [&](const LeafRange &A, const LeafRange &B) -> bool {
    if (A.m != B.m || A.n != B.n || A.admissible != B.admissible) return false;
    const uint a0 = perm[A.r0] / nvc, b0 = perm[B.r0] / nvc;
    for (uint i = 0; i < A.m; i++) for (int d = 0; d < 4; d++)
      if (key_elem(perm[A.r0 + i], a0, d) != key_elem(perm[B.r0 + i], b0, d)) return false;
    for (uint i = 0; i < A.n; i++) for (int d = 0; d < 4; d++)
      if (key_elem(perm[A.c0 + i], a0, d) != key_elem(perm[B.c0 + i], b0, d)) return false;
    return true;
  };
  std::unordered_map<unsigned long long, std::vector<uint>> seen;
  seen.reserve(ranges.size() / 4 + 16);
  for (size_t b = 0; b < ranges.size(); b++) {
    if (!ranges[b].m || !ranges[b].n) continue;
    std::vector<uint> &cand = seen[hash_leaf(ranges[b])];
    bool found = false;
    for (uint r : cand)
      if (same(ranges[r], ranges[b])) { rep[b] = r; found = true; break; }
    if (!found) cand.push_back((uint)b);
  }
}

__global__ void k_copy_ranks(uint ndup, const uint *__restrict__ dup, HLeafDev *leaves) {
  const uint i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ndup) leaves[dup[2 * i]].k = leaves[dup[2 * i + 1]].k;
}

int hmat_create_from_ranges(HMat &H, const BemDev &bem, const uint *htri, const std::vector<uint> &perm,
                            const std::vector<LeafRange> &ranges, uint kmax, const double *hx) {
  H.bem = &bem;
  H.n = bem.nv;
  H.kmax = kmax;
  if (&perm != &H.tree.perm) H.tree.perm = perm;
  H.tree.pos.resize(bem.nv);
  for (uint i = 0; i < bem.nv; i++) H.tree.pos[H.tree.perm[i]] = i;
  // vertex -> triangles, and per cluster range the triangles touching it (dense leaves only)
  std::vector<uint> vstart(bem.nv + 1, 0), vadj(3 * (size_t)bem.nt), v2t(3 * (size_t)bem.nt);
  for (size_t i = 0; i < 3 * (size_t)bem.nt; i++) vstart[htri[i] + 1]++;
  for (uint v = 0; v < bem.nv; v++) vstart[v + 1] += vstart[v];
  {
    std::vector<uint> fill(vstart.begin(), vstart.end() - 1);
    for (uint t = 0; t < bem.nt; t++)
      for (uint d = 0; d < 3; d++) { uint e = fill[htri[3 * (size_t)t + d]]++; vadj[e] = t; v2t[e] = t * 4 + d; }
  }
  std::map<std::pair<uint, uint>, uint> list_of;  // (start, size) -> triangle list id
  std::vector<uint> ctri_ptr(1, 0), ctri;
  auto tri_list = [&](uint start, uint size) -> uint {
    auto it = list_of.find({start, size});
    if (it != list_of.end()) return it->second;
    std::vector<uint> ts;
    for (uint i = 0; i < size; i++) {
      uint v = H.tree.perm[start + i];
      for (uint e = vstart[v]; e < vstart[v + 1]; e++) ts.push_back(vadj[e]);
    }
    std::sort(ts.begin(), ts.end());
    ts.erase(std::unique(ts.begin(), ts.end()), ts.end());
    ctri.insert(ctri.end(), ts.begin(), ts.end());
    ctri_ptr.push_back((uint)ctri.size());
    uint id = (uint)ctri_ptr.size() - 2;
    list_of[{start, size}] = id;
    return id;
  };
  H.leaf.clear();
  std::vector<uint> near_ids, far_ids, rep, dup;
  leaf_translation_classes(bem, hx, H.tree.perm, ranges, rep);
  size_t off = 0, foff = 0;
  for (size_t b = 0; b < ranges.size(); b++) {
    const LeafRange &B = ranges[b];
    HLeafDev L;
    memset(&L, 0, sizeof(L));
    L.r0 = B.r0; L.m = B.m; L.c0 = B.c0; L.n = B.n;
    L.adm = B.admissible ? 1 : 0;
    if (rep[b] != b) {   // translate of an earlier leaf: same data, nothing to assemble
      const HLeafDev &Rp = H.leaf[rep[b]];
      L.off = Rp.off; L.kcap = Rp.kcap; L.foff = Rp.foff; L.rtl = Rp.rtl; L.ctl = Rp.ctl;
      if (L.adm) { dup.push_back((uint)b); dup.push_back(rep[b]); }
      H.leaf.push_back(L);
      continue;
    }
    L.off = off;
    if (B.admissible) {
      L.kcap = std::min(std::min(L.m, L.n), kmax);
      L.foff = foff;
      foff += (size_t)L.m + L.n;
      off += (size_t)(L.m + L.n) * L.kcap;
      far_ids.push_back((uint)b);
    } else {
      L.rtl = tri_list(B.r0, B.m);
      L.ctl = tri_list(B.c0, B.n);
      off += (size_t)L.m * L.n;
      near_ids.push_back((uint)b);
    }
    H.leaf.push_back(L);
  }
  H.pool_size = off;
  H.nnear = (uint)near_ids.size();
  H.nfar = (uint)far_ids.size();
  H.near_entries = 0;   // logical entries (every leaf, shared or not)
  for (const HLeafDev &L : H.leaf) if (!L.adm) H.near_entries += (size_t)L.m * L.n;
  H.ndup = (uint)(dup.size() / 2);
  if (to_device(H.d_dup, dup)) return -1;
  CNLD_CK(cudaSetDevice(bem.device));
  if (to_device(H.d_leaf, H.leaf) || to_device(H.d_near, near_ids) || to_device(H.d_far, far_ids) ||
      to_device(H.d_perm, H.tree.perm) || to_device(H.d_pos, H.tree.pos) || to_device(H.d_ctri_ptr, ctri_ptr) ||
      to_device(H.d_ctri, ctri) || to_device(H.d_v2t_ptr, vstart) || to_device(H.d_v2t, v2t))
    return -1;
  CNLD_CK(cudaMalloc(&H.d_pool, sizeof(cplx) * (off ? off : 1)));
  CNLD_CK(cudaMalloc(&H.d_flags, foff ? foff : 1));
  CNLD_CK(cudaMalloc(&H.d_err, 2 * sizeof(uint)));
  return 0;
}

// matvec-only H-matrix from host leaf data: `leaf` (off/k/kcap filled) and a packed pool
int hmat_from_host(HMat &H, int device, uint n, const std::vector<uint> &perm, const std::vector<HLeafDev> &leaf,
                   const std::vector<cplx> &pool) {
  H.bem = nullptr;
  H.n = n;
  H.tree.perm = perm;
  H.leaf = leaf;
  H.pool_size = pool.size();
  CNLD_CK(cudaSetDevice(device));
  if (to_device(H.d_leaf, H.leaf) || to_device(H.d_perm, H.tree.perm) || to_device(H.d_pool, pool)) return -1;
  H.assembled = true;
  return 0;
}

// band lists (structure) and rank offsets (after every assembly) of the batched matvec
int hmat_prepare_batched(HMat &H, bool fetch) {
  if (fetch && hmat_fetch_leaves(H)) return -1;
  if (!H.d_band_ptr) {   // structure: bands x dense rounds, list of low-rank leaves
    H.nband = (H.n + BAND - 1) / BAND;
    std::vector<uint> ptr(H.nband + 1, 0), mvfar;
    for (size_t id = 0; id < H.leaf.size(); id++) {
      const HLeafDev &L = H.leaf[id];
      if (!L.m || !L.n) continue;
      if (L.adm) { mvfar.push_back((uint)id); continue; }
      for (uint b = L.r0 / BAND; b <= (L.r0 + L.m - 1) / BAND; b++) ptr[b + 1] += (L.n + 15) / 16;
    }
    for (uint b = 0; b < H.nband; b++) ptr[b + 1] += ptr[b];
    std::vector<HMat::DRound> rounds(ptr[H.nband]);
    std::vector<uint> fill(ptr.begin(), ptr.end() - 1);
    for (size_t id = 0; id < H.leaf.size(); id++) {
      const HLeafDev &L = H.leaf[id];
      if (!L.m || !L.n || L.adm) continue;
      for (uint b = L.r0 / BAND; b <= (L.r0 + L.m - 1) / BAND; b++)
        for (uint j0 = 0; j0 < L.n; j0 += 16)
          rounds[fill[b]++] = {L.off + (size_t)j0 * L.m, L.r0, L.m, L.c0 + j0, std::min(16u, L.n - j0), {0, 0}};
    }
    H.nmvfar = (uint)mvfar.size();
    H.far_by_col = mvfar;   // grouped by column cluster for the stacked t-pass
    std::stable_sort(H.far_by_col.begin(), H.far_by_col.end(), [&](uint a, uint b) {
      const HLeafDev &A = H.leaf[a], &B = H.leaf[b];
      return A.c0 != B.c0 ? A.c0 < B.c0 : A.n < B.n;
    });
    if (to_device(H.d_band_ptr, ptr) || to_device(H.d_dround, rounds) || to_device(H.d_mvfar, mvfar)) return -1;
    CNLD_CK(cudaMalloc(&H.d_toff, sizeof(unsigned long long) * (H.leaf.size() + 1)));
    CNLD_CK(cudaMalloc(&H.d_bfar_ptr, sizeof(uint) * (H.nband + 1)));
  }
  // ranks: scratch offsets and the packed (leaf, rank index) items of every band.  The scratch rows of
  // the low-rank leaves of one column cluster are consecutive and padded to a multiple of 16 (stacked
  // t-pass: one CTA per 16 scratch rows).
  std::vector<unsigned long long> toff(H.leaf.size() + 1, 0);
  std::vector<uint> fptr(H.nband + 1, 0);
  std::vector<HMat::TBand> tbands;
  std::vector<unsigned long long> vrow;
  unsigned long long sum = 0;
  for (size_t q = 0; q < H.far_by_col.size();) {
    const HLeafDev &L0 = H.leaf[H.far_by_col[q]];
    const unsigned long long base = sum;
    size_t e = q;
    for (; e < H.far_by_col.size(); e++) {
      const uint id = H.far_by_col[e];
      const HLeafDev &L = H.leaf[id];
      if (L.c0 != L0.c0 || L.n != L0.n) break;
      toff[id] = sum;
      const size_t voff = L.off + (size_t)L.m * L.kcap;
      for (uint l = 0; l < L.k; l++) vrow.push_back(voff + (size_t)l * L.n);
      sum += L.k;
    }
    const unsigned long long rows = sum - base;
    for (unsigned long long r = 0; r < rows; r += 16)
      tbands.push_back({L0.c0, L0.n, (uint)std::min<unsigned long long>(16, rows - r), 0});
    sum = base + (rows + 15) / 16 * 16;
    vrow.resize(sum, ~0ull);
    q = e;
  }
  for (size_t id = 0; id < H.leaf.size(); id++) {
    const HLeafDev &L = H.leaf[id];
    if (!L.adm || !L.m || !L.n) continue;
    for (uint b = L.r0 / BAND; b <= (L.r0 + L.m - 1) / BAND; b++) fptr[b + 1] += L.k;
  }
  H.ksum = sum;
  H.ntband = (uint)tbands.size();
  if (H.tband_cap < tbands.size()) {
    cudaFree(H.d_tband);
    H.d_tband = nullptr;
    H.tband_cap = tbands.size() + tbands.size() / 4 + 16;
    CNLD_CK(cudaMalloc(&H.d_tband, sizeof(HMat::TBand) * H.tband_cap));
  }
  if (H.vrow_cap < vrow.size()) {
    cudaFree(H.d_vrow);
    H.d_vrow = nullptr;
    H.vrow_cap = vrow.size() + vrow.size() / 4 + 16;
    CNLD_CK(cudaMalloc(&H.d_vrow, sizeof(unsigned long long) * H.vrow_cap));
  }
  CNLD_CK(cudaMemcpy(H.d_tband, tbands.data(), sizeof(HMat::TBand) * tbands.size(), cudaMemcpyHostToDevice));
  CNLD_CK(cudaMemcpy(H.d_vrow, vrow.data(), sizeof(unsigned long long) * vrow.size(), cudaMemcpyHostToDevice));
  for (uint b = 0; b < H.nband; b++) fptr[b + 1] += fptr[b];
  std::vector<HMat::FarItem> items(fptr[H.nband]);
  {
    std::vector<uint> fill(fptr.begin(), fptr.end() - 1);
    for (size_t id = 0; id < H.leaf.size(); id++) {
      const HLeafDev &L = H.leaf[id];
      if (!L.adm || !L.m || !L.n) continue;
      for (uint b = L.r0 / BAND; b <= (L.r0 + L.m - 1) / BAND; b++)
        for (uint l = 0; l < L.k; l++) items[fill[b]++] = {L.off + (size_t)l * L.m, toff[id] + l, L.r0, L.m};
    }
  }
  if (H.bfar_cap < items.size()) {
    cudaFree(H.d_bfar);
    H.d_bfar = nullptr;
    H.bfar_cap = items.size() + items.size() / 4 + 16;
    CNLD_CK(cudaMalloc(&H.d_bfar, sizeof(HMat::FarItem) * H.bfar_cap));
  }
  CNLD_CK(cudaMemcpy(H.d_bfar, items.data(), sizeof(HMat::FarItem) * items.size(), cudaMemcpyHostToDevice));
  CNLD_CK(cudaMemcpy(H.d_bfar_ptr, fptr.data(), sizeof(uint) * (H.nband + 1), cudaMemcpyHostToDevice));
  CNLD_CK(cudaMemcpy(H.d_toff, toff.data(), sizeof(unsigned long long) * H.leaf.size(), cudaMemcpyHostToDevice));
  H.mv_ready = true;
  return 0;
}

void hmat_free(HMat &H) {
  cudaFree(H.d_band_ptr); cudaFree(H.d_dround); cudaFree(H.d_mvfar); cudaFree(H.d_toff); cudaFree(H.d_T);
  cudaFree(H.d_bfar_ptr); cudaFree(H.d_bfar); cudaFree(H.d_dup); cudaFree(H.d_tband); cudaFree(H.d_vrow);
  cudaFree(H.d_leaf); cudaFree(H.d_near); cudaFree(H.d_far); cudaFree(H.d_perm); cudaFree(H.d_pos);
  cudaFree(H.d_ctri_ptr); cudaFree(H.d_ctri); cudaFree(H.d_v2t_ptr); cudaFree(H.d_v2t); cudaFree(H.d_pool);
  cudaFree(H.d_flags); cudaFree(H.d_err); cudaFree(H.d_xp); cudaFree(H.d_yp); cudaFree(H.d_cpos);
  H = HMat();
}

template <int NQ>
static int assemble_nq(cudaStream_t st, HMat &H, double kr, double ki, double accur) {
  const BemDev &b = *H.bem;
  if (H.nnear) {
    if (ki != 0.0)
      k_near<NQ, true><<<H.nnear, NEAR_THREADS, 0, st>>>(H.d_leaf, H.d_near, b.rule, b.sing, b.x, b.t, b.qp, b.g,
                                                          H.d_pos, H.d_cpos ? H.d_cpos : H.d_pos, H.d_ctri_ptr,
                                                          H.d_ctri, kr, ki, H.d_pool, H.d_err);
    else
      k_near<NQ, false><<<H.nnear, NEAR_THREADS, 0, st>>>(H.d_leaf, H.d_near, b.rule, b.sing, b.x, b.t, b.qp, b.g,
                                                           H.d_pos, H.d_cpos ? H.d_cpos : H.d_pos, H.d_ctri_ptr,
                                                           H.d_ctri, kr, ki, H.d_pool, H.d_err);
    CNLD_CK_LAUNCH();
  }
  if (H.nfar) {
    if (ki != 0.0)
      k_aca<NQ, true><<<H.nfar, ACA_THREADS, 0, st>>>(H.d_leaf, H.d_far, b.rule, b.qp, b.g, H.d_perm, H.d_v2t_ptr,
                                                       H.d_v2t, kr, ki, accur, H.d_pool, H.d_flags, H.d_err + 1);
    else
      k_aca<NQ, false><<<H.nfar, ACA_THREADS, 0, st>>>(H.d_leaf, H.d_far, b.rule, b.qp, b.g, H.d_perm, H.d_v2t_ptr,
                                                        H.d_v2t, kr, ki, accur, H.d_pool, H.d_flags, H.d_err + 1);
    CNLD_CK_LAUNCH();
  }
  return 0;
}

int hmat_assemble(cudaStream_t st, HMat &H, double kr, double ki, double accur) {
  CNLD_CK(cudaMemsetAsync(H.d_pool, 0, sizeof(cplx) * H.pool_size, st));
  CNLD_CK(cudaMemsetAsync(H.d_err, 0, 2 * sizeof(uint), st));
  int rc;
  switch (H.bem->nq) {
  case 1: rc = assemble_nq<1>(st, H, kr, ki, accur); break;
  case 4: rc = assemble_nq<4>(st, H, kr, ki, accur); break;
  case 9: rc = assemble_nq<9>(st, H, kr, ki, accur); break;
  default: set_error("H-matrix path supports q_reg <= 3"); return -1;
  }
  if (rc) return rc;
  if (H.ndup) {   // translates take the rank of the leaf whose data they share
    k_copy_ranks<<<(H.ndup + 255) / 256, 256, 0, st>>>(H.ndup, H.d_dup, H.d_leaf);
    CNLD_CK_LAUNCH();
  }
  uint errs[2] = {0, 0};
  CNLD_CK(cudaMemcpyAsync(errs, H.d_err, 2 * sizeof(uint), cudaMemcpyDeviceToHost, st));
  CNLD_CK(cudaStreamSynchronize(st));
  const uint err = errs[0];
  H.capped_leaves = errs[1];
  if (err) { set_error("H-matrix near-field: touching-pair queue overflow (leaf too large; raise NEAR_QCAP)"); return -1; }
  H.assembled = true;
  H.mv_ready = false;   // ranks changed: the batched matvec refreshes its offsets on first use
  return 0;
}

int g_hmv_tstack = 1;
int g_hmv_ywarps = 0;   // 0: pick by the batch width

// Fill every leaf of H from an explicit dense matrix on the device (original DOF order, column-major):
// inadmissible leaves take their sub-block, admissible leaves its partial-pivot ACA to `accur`.
int hmat_compress_dense(cudaStream_t st, HMat &H, const cplx *d_Z, size_t ld, double accur) {
  CNLD_CK(cudaMemsetAsync(H.d_pool, 0, sizeof(cplx) * H.pool_size, st));
  CNLD_CK(cudaMemsetAsync(H.d_err, 0, 2 * sizeof(uint), st));
  if (H.nnear) {
    k_gather_dense<<<H.nnear, 256, 0, st>>>(H.d_leaf, H.d_near, d_Z, ld, H.d_perm, H.d_pool);
    CNLD_CK_LAUNCH();
  }
  if (H.nfar) {
    k_aca_dense<<<H.nfar, ACA_THREADS, 0, st>>>(H.d_leaf, H.d_far, d_Z, ld, H.d_perm, accur, H.d_pool, H.d_flags, H.d_err + 1);
    CNLD_CK_LAUNCH();
  }
  uint errs[2] = {0, 0};
  CNLD_CK(cudaMemcpyAsync(errs, H.d_err, 2 * sizeof(uint), cudaMemcpyDeviceToHost, st));
  CNLD_CK(cudaStreamSynchronize(st));
  H.capped_leaves = errs[1];
  H.assembled = true;
  H.mv_ready = false;
  return 0;
}

// device vectors in ORIGINAL DOF order, nrhs of them with stride n:  y += alpha * Z x
int hmat_matvec(cudaStream_t st, HMat &H, cplx alpha, const cplx *d_x, cplx *d_y, uint nrhs) {
  if (!H.assembled) { set_error("H-matrix has not been assembled"); return -1; }
  if (H.vec_cap < nrhs) {
    cudaFree(H.d_xp); cudaFree(H.d_yp);
    H.d_xp = H.d_yp = nullptr;
    CNLD_CK(cudaMalloc(&H.d_xp, sizeof(cplx) * (size_t)H.n * nrhs));
    CNLD_CK(cudaMalloc(&H.d_yp, sizeof(cplx) * (size_t)H.n * nrhs));
    H.vec_cap = nrhs;
  }
  const size_t tot = (size_t)H.n * nrhs;
  const unsigned gb = (unsigned)((tot + 255) / 256);
  if (nrhs >= 4) {
    if (!H.mv_ready) {
      CNLD_CK(cudaStreamSynchronize(st));
      if (hmat_prepare_batched(H, H.bem != nullptr)) return -1;
    }
    const size_t need = (size_t)H.ksum * nrhs;
    if (H.T_cap < need) {
      CNLD_CK(cudaStreamSynchronize(st));
      cudaFree(H.d_T);
      H.d_T = nullptr;
      CNLD_CK(cudaMalloc(&H.d_T, sizeof(cplx) * (need ? need : 1)));
      H.T_cap = need;
    }
    k_permute_in_t<<<gb, 256, 0, st>>>(H.n, nrhs, H.d_perm, d_x, H.d_xp);
    CNLD_CK_LAUNCH();
    // t-pass: right-hand sides per pass = 8 TC NW; the width that wastes the fewest columns, ties to the wider
    if (!g_hmv_tstack && H.nmvfar) {   // debug: one CTA per low-rank leaf (round-1 kernel)
      k_hmv_gemm<4, true><<<H.nmvfar, 64, (int)sizeof(cplx) * 2 * 16 * (AS_LDB + 16 * 4 + 2), st>>>(H.d_band_ptr, H.d_dround, H.d_bfar_ptr, H.d_bfar, H.d_mvfar,
                                                           H.d_leaf, H.d_toff, H.d_pool, H.d_T, H.d_xp, H.d_yp, H.n, nrhs);
      CNLD_CK_LAUNCH();
    } else if (H.ntband) {
      int bw = 2, bt = 2;
      uint bestw = ~0u;
      for (int w : {2, 4})
        for (int c : {2, 3, 4}) {
          const uint rp = 8 * c * w, waste = (nrhs + rp - 1) / rp * rp - nrhs;
          if (waste < bestw || (waste == bestw && 8 * c * w > 8 * bt * bw)) { bestw = waste; bw = w; bt = c; }
        }
#define CNLD_TS(NWV, TCV)                                                                                          \
  do {                                                                                                             \
    constexpr int SM = (int)sizeof(cplx) * 2 * 16 * (TS_LDA + 8 * TCV * NWV + 2);                                  \
    static bool attr[64] = {false};                                                                                \
    int dev = 0;                                                                                                   \
    CNLD_CK(cudaGetDevice(&dev));                                                                                  \
    if (!attr[dev & 63]) {                                                                                         \
      CNLD_CK(cudaFuncSetAttribute(k_hmv_tstack<NWV, TCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));      \
      attr[dev & 63] = true;                                                                                       \
    }                                                                                                              \
    k_hmv_tstack<NWV, TCV><<<H.ntband, 32 * NWV, SM, st>>>(H.d_tband, H.d_vrow, H.d_pool, H.d_T, H.d_xp, nrhs);    \
  } while (0)
      if (bw == 2) { if (bt == 2) CNLD_TS(2, 2); else if (bt == 3) CNLD_TS(2, 3); else CNLD_TS(2, 4); }
      else { if (bt == 2) CNLD_TS(4, 2); else if (bt == 3) CNLD_TS(4, 3); else CNLD_TS(4, 4); }
#undef CNLD_TS
      CNLD_CK_LAUNCH();
    }
    // y-pass: right-hand sides per pass = 8 TC NW, chosen like the t-pass (fewest wasted columns, ties to the wider:
    // a wider pass re-reads the leaf data and the round descriptors fewer times)
    {
      int bw = 2, bt = 2;
      uint bestw = ~0u;
      for (int w : {2, 4})
        for (int c : {2, 3, 4}) {
          const uint rp = 8 * c * w, waste = (nrhs + rp - 1) / rp * rp - nrhs;
          if (waste < bestw || (waste == bestw && 8 * c * w > 8 * bt * bw)) { bestw = waste; bw = w; bt = c; }
        }
      if (g_hmv_ywarps == 2 || g_hmv_ywarps == 4) {   // debug override of the warp count
        bw = g_hmv_ywarps;
        bestw = ~0u;
        for (int c : {2, 3, 4}) {
          const uint rp = 8 * c * bw, waste = (nrhs + rp - 1) / rp * rp - nrhs;
          if (waste <= bestw) { bestw = waste; bt = c; }
        }
      }
#define CNLD_HMV(NWV, TCV)                                                                                         \
  do {                                                                                                             \
    constexpr int SM = (int)sizeof(cplx) * 2 * 16 * (AS_LDB + 8 * TCV * NWV + 2);                                  \
    static bool attr[64] = {false};                                                                                \
    int dev = 0;                                                                                                   \
    CNLD_CK(cudaGetDevice(&dev));                                                                                  \
    if (!attr[dev & 63]) {                                                                                         \
      CNLD_CK(cudaFuncSetAttribute(k_hmv_gemm<TCV, false, NWV>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM)); \
      attr[dev & 63] = true;                                                                                       \
    }                                                                                                              \
    k_hmv_gemm<TCV, false, NWV><<<H.nband, 32 * NWV, SM, st>>>(H.d_band_ptr, H.d_dround, H.d_bfar_ptr, H.d_bfar,     \
        H.d_mvfar, H.d_leaf, H.d_toff, H.d_pool, H.d_T, H.d_xp, H.d_yp, H.n, nrhs);                                \
  } while (0)
      if (bw == 2) { if (bt == 2) CNLD_HMV(2, 2); else if (bt == 3) CNLD_HMV(2, 3); else CNLD_HMV(2, 4); }
      else { if (bt == 2) CNLD_HMV(4, 2); else if (bt == 3) CNLD_HMV(4, 3); else CNLD_HMV(4, 4); }
#undef CNLD_HMV
      CNLD_CK_LAUNCH();
    }
    k_permute_out_t<<<gb, 256, 0, st>>>(H.n, nrhs, H.d_perm, alpha, H.d_yp, d_y);
    CNLD_CK_LAUNCH();
    return 0;
  }
  // 1-3 vectors: every leaf is read once, coalesced (HBM-bound)
  if (!H.mv_ready) {
    CNLD_CK(cudaStreamSynchronize(st));
    if (hmat_prepare_batched(H, H.bem != nullptr)) return -1;
  }
  {
    const size_t need = (size_t)H.ksum * nrhs;
    if (H.T_cap < need) {
      CNLD_CK(cudaStreamSynchronize(st));
      cudaFree(H.d_T);
      H.d_T = nullptr;
      CNLD_CK(cudaMalloc(&H.d_T, sizeof(cplx) * (need ? need : 1)));
      H.T_cap = need;
    }
  }
  k_permute_in<<<gb, 256, 0, st>>>(H.n, nrhs, H.d_perm, d_x, H.d_xp);
  CNLD_CK_LAUNCH();
  const unsigned gt = (unsigned)(((size_t)H.nmvfar * 32 + V1_THREADS - 1) / V1_THREADS);
  const unsigned gy = H.nband;          // one CTA per band
  // grid-stride t-pass: 3 CTAs per SM resident, 8 rounds of CTAs (tail balance), never more warps than bands
  const unsigned gtb = (unsigned)std::min<size_t>(((size_t)H.ntband * 32 + V1_THREADS - 1) / V1_THREADS, (size_t)148 * 3 * 8);
#define CNLD_HMV1(NRV)                                                                                             \
  do {                                                                                                             \
    if (H.nmvfar && !g_hmv_tstack) {                                                                                \
      k_hmv1_t<NRV><<<gt, V1_THREADS, 0, st>>>(H.nmvfar, H.d_mvfar, H.d_leaf, H.d_toff, H.d_pool, H.d_xp, H.d_T, H.n); \
      CNLD_CK_LAUNCH();                                                                                            \
    } else if (H.ntband) {                                                                                         \
      k_hmv1_tb<NRV><<<gtb, V1_THREADS, 0, st>>>(H.ntband, H.d_tband, H.d_vrow, H.d_pool, H.d_xp, H.d_T, H.n);    \
      CNLD_CK_LAUNCH();                                                                                            \
    }                                                                                                              \
    k_hmv1_band<NRV><<<gy, V1_THREADS, 0, st>>>(H.nband, H.d_band_ptr, H.d_dround, H.d_bfar_ptr, H.d_bfar, H.d_pool,  \
                                                H.d_T, H.d_xp, H.d_yp, H.n);                                       \
    CNLD_CK_LAUNCH();                                                                                              \
  } while (0)
  if (nrhs == 1) CNLD_HMV1(1);
  else if (nrhs == 2) CNLD_HMV1(2);
  else CNLD_HMV1(3);
#undef CNLD_HMV1
  k_permute_out<<<gb, 256, 0, st>>>(H.n, nrhs, H.d_perm, alpha, H.d_yp, d_y);
  CNLD_CK_LAUNCH();
  return 0;
}

int hmat_expand(cudaStream_t st, HMat &H, cplx *d_Z, size_t ld) {
  if (!H.assembled) { set_error("H-matrix has not been assembled"); return -1; }
  k_expand<<<(unsigned)H.leaf.size(), 256, 0, st>>>(H.d_leaf, H.d_pool, H.d_perm, d_Z, ld);
  CNLD_CK_LAUNCH();
  return 0;
}

// bem->nearfield(ridx, cidx): dense sub-block for arbitrary vertex index sets, tiled 32 x 256 so the
// touching-pair queue of k_near cannot overflow.  d_out is rows x cols column-major (ld = rows).
int nearfield_block(cudaStream_t st, const BemDev &bem, const uint *htri, const uint *ridx, uint rows,
                    const uint *cidx, uint cols, double kr, double ki, cplx *d_out) {
  const uint NONE = 0xffffffffu, TR = 32, TC = 256;
  HMat H;
  H.bem = &bem;
  H.n = bem.nv;
  std::vector<uint> rpos(bem.nv, NONE), cpos(bem.nv, NONE);
  for (uint i = 0; i < rows; i++) {
    if (ridx[i] >= bem.nv || rpos[ridx[i]] != NONE) { set_error("nearfield: bad or repeated row index"); return -1; }
    rpos[ridx[i]] = i;
  }
  for (uint i = 0; i < cols; i++) {
    if (cidx[i] >= bem.nv || cpos[cidx[i]] != NONE) { set_error("nearfield: bad or repeated column index"); return -1; }
    cpos[cidx[i]] = i;
  }
  std::vector<uint> vstart(bem.nv + 1, 0), vadj(3 * (size_t)bem.nt);
  for (size_t i = 0; i < 3 * (size_t)bem.nt; i++) vstart[htri[i] + 1]++;
  for (uint v = 0; v < bem.nv; v++) vstart[v + 1] += vstart[v];
  {
    std::vector<uint> fill(vstart.begin(), vstart.end() - 1);
    for (uint t = 0; t < bem.nt; t++)
      for (uint d = 0; d < 3; d++) vadj[fill[htri[3 * (size_t)t + d]]++] = t;
  }
  std::vector<uint> ctri_ptr(1, 0), ctri;
  auto tri_list = [&](const uint *idx, uint cnt) -> uint {
    std::vector<uint> ts;
    for (uint i = 0; i < cnt; i++)
      for (uint e = vstart[idx[i]]; e < vstart[idx[i] + 1]; e++) ts.push_back(vadj[e]);
    std::sort(ts.begin(), ts.end());
    ts.erase(std::unique(ts.begin(), ts.end()), ts.end());
    ctri.insert(ctri.end(), ts.begin(), ts.end());
    ctri_ptr.push_back((uint)ctri.size());
    return (uint)ctri_ptr.size() - 2;
  };
  std::vector<uint> rlist, clist;
  for (uint r0 = 0; r0 < rows; r0 += TR) rlist.push_back(tri_list(ridx + r0, std::min(TR, rows - r0)));
  for (uint c0 = 0; c0 < cols; c0 += TC) clist.push_back(tri_list(cidx + c0, std::min(TC, cols - c0)));
  // tiles write straight into the strided output: leaf offset = element (r0, c0), "m" = tile rows.
  // k_near addresses D[pc * L.m + pr]; with a strided target we need ld = rows, so every tile is
  // assembled into a compact scratch tile and copied out afterwards.
  std::vector<uint> near_ids;
  size_t off = 0;
  for (uint ci = 0, c0 = 0; c0 < cols; c0 += TC, ci++)
    for (uint ri = 0, r0 = 0; r0 < rows; r0 += TR, ri++) {
      HLeafDev L;
      memset(&L, 0, sizeof(L));
      L.r0 = r0; L.m = std::min(TR, rows - r0); L.c0 = c0; L.n = std::min(TC, cols - c0);
      L.rtl = rlist[ri]; L.ctl = clist[ci];
      L.off = off;
      off += (size_t)L.m * L.n;
      near_ids.push_back((uint)H.leaf.size());
      H.leaf.push_back(L);
    }
  H.pool_size = off;
  H.nnear = (uint)near_ids.size();
  H.nfar = 0;
  CNLD_CK(cudaSetDevice(bem.device));
  std::vector<uint> perm(bem.nv);
  for (uint i = 0; i < bem.nv; i++) perm[i] = i;
  int rc = 0;
  if (to_device(H.d_leaf, H.leaf) || to_device(H.d_near, near_ids) || to_device(H.d_pos, rpos) || to_device(H.d_cpos, cpos) ||
      to_device(H.d_ctri_ptr, ctri_ptr) || to_device(H.d_ctri, ctri) || to_device(H.d_perm, perm)) rc = -1;
  if (!rc && cudaMalloc(&H.d_pool, sizeof(cplx) * (off ? off : 1)) != cudaSuccess) { set_error("nearfield: out of memory"); rc = -1; }
  if (!rc && cudaMalloc(&H.d_err, 2 * sizeof(uint)) != cudaSuccess) { set_error("nearfield: out of memory"); rc = -1; }
  if (!rc) rc = hmat_assemble(st, H, kr, ki, 0.0);
  if (!rc) {
    k_untile<<<(unsigned)H.leaf.size(), 256, 0, st>>>(H.d_leaf, H.d_pool, d_out, rows);
    CNLD_CK_LAUNCH();
    CNLD_CK(cudaStreamSynchronize(st));
  }
  hmat_free(H);
  return rc;
}

int hmat_fetch_leaves(HMat &H) {
  CNLD_CK(cudaMemcpy(H.leaf.data(), H.d_leaf, sizeof(HLeafDev) * H.leaf.size(), cudaMemcpyDeviceToHost));
  return 0;
}

}  // namespace cnld
