// C ABI of the H-matrix path (include/cnld_b200.h, "cnld_b200_hmat_*" / "cnld_b200_hsweep_run")
// and the GPU GMRES that replaces the reference's H-LU for arrays too large to factor densely
// (BASELINE.json north_star; Krylov formulation of scratch/freq_resp_db_krylov.py:38-79 and
// solve_gmres_hmatrix_avector, include/krylovsolvers.h:283).
//
// Per frequency:  G x = F[:, s]  with  G = (K - w^2 M - j w B) + (-2 rho w^2) Z_H(k),
// left-preconditioned by the block-diagonal inverse of the FEM operator (one dense inverse per
// membrane type and frequency, the `Gfe_inv` of the reference's Krylov driver), restarted
// GMRES with classical Gram-Schmidt + re-orthogonalisation, all right-hand sides of a batch
// iterated in lock-step so every H-matrix leaf is streamed from HBM once per iteration for the
// whole batch.
#include <cmath>
#include <complex>

#include "ctx.cuh"
#include "hmat.cuh"

struct cnld_hmat {
  cnld_ctx *ctx = nullptr;
  HMat H;
  unsigned last_capped = 0;   // rank-capped leaves summed over the assemblies of the last hsweep_run
  // coarse space of the two-level GMRES preconditioner (cnld_b200_hmat_set_coarse): m functions per node
  // (membrane eigenmodes, zero on clamped nodes), W[n][m], and their in-vacuo resonance frequencies
  uint coarse_m = 0, last_coarse_m = 0;
  std::vector<double> coarse_W, coarse_f;
};

namespace cnld_hapi {
typedef std::complex<double> zc;
constexpr double H_TWO_PI = 6.283185307179586476925286766559;

// ---- batched vector kernels; vectors are [r][n] with stride n --------------------------------
// Y[r][i] = sum_j (K - w^2 M - j w B)[i,j] X[r][j]: the sparse FEM operator (about 20 entries per row)
// applied to all right-hand sides.  Thread = (row, chunk of RC right-hand sides): the row's pattern and
// values are read once per chunk; consecutive threads take consecutive rows, so the gathers of one
// right-hand side fall into neighbouring sectors.
template <int RC>
__global__ void __launch_bounds__(256)
k_spmm_mbk(uint n, uint nb, const int64_t *__restrict__ indptr, const uint *__restrict__ indices,
           const double *__restrict__ K, const double *__restrict__ M, const double *__restrict__ B,
           double omega, const cplx *__restrict__ X, cplx *Y) {
  const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint nch = (nb + RC - 1) / RC;
  if (id >= (size_t)n * nch) return;
  const uint i = id % n, r0 = (uint)(id / n) * RC;
  const double w2 = omega * omega;
  cplx acc[RC];
#pragma unroll
  for (int q = 0; q < RC; q++) acc[q] = make_double2(0.0, 0.0);
  for (int64_t p = indptr[i]; p < indptr[i + 1]; p++) {
    const cplx a = make_double2(K[p] - w2 * M[p], -omega * B[p]);
    const cplx *x = X + (size_t)r0 * n + indices[p];
#pragma unroll
    for (int q = 0; q < RC; q++)
      if (r0 + q < nb) cfma(acc[q], a, x[(size_t)q * n]);
  }
#pragma unroll
  for (int q = 0; q < RC; q++)
    if (r0 + q < nb) Y[(size_t)(r0 + q) * n + i] = acc[q];
}

__global__ void k_block_dense(uint nb_, uint r0, const int64_t *__restrict__ indptr, const uint *__restrict__ indices,
                              const double *__restrict__ K, const double *__restrict__ M, const double *__restrict__ B,
                              double omega, cplx *D) {
  // dense copy of one diagonal block of the FEM operator (column-major nb_ x nb_, zero-initialised)
  uint i = blockIdx.x;
  const double w2 = omega * omega;
  for (int64_t p = indptr[r0 + i] + threadIdx.x; p < indptr[r0 + i + 1]; p += blockDim.x) {
    uint j = indices[p] - r0;
    if (j < nb_) D[(size_t)j * nb_ + i] = make_double2(K[p] - w2 * M[p], -omega * B[p]);
  }
}
__global__ void k_axpy_scaled(size_t n, double a, const cplx *__restrict__ x, cplx *y) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { y[i].x += a * x[i].x; y[i].y += a * x[i].y; }
}
__global__ void k_identity(uint n, cplx *D) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)n * n) D[i] = make_double2((i % n) == (i / n) ? 1.0 : 0.0, 0.0);
}

// Hd[i][r] = <V_i[r], W[r]>, i = 0..nv-1 (conjugate-linear in V); grid (nb, nv)
__global__ void __launch_bounds__(256) k_dots(uint n, uint nb, const cplx *__restrict__ Vb, const cplx *__restrict__ W, cplx *Hd) {
  __shared__ double sr[8], si[8];
  const uint r = blockIdx.x, i = blockIdx.y;
  const cplx *v = Vb + ((size_t)i * nb + r) * n, *w = W + (size_t)r * n;
  double ar = 0.0, ai = 0.0;
  for (uint e = threadIdx.x; e < n; e += 256) {
    cplx a = v[e], b = w[e];
    ar += a.x * b.x + a.y * b.y;
    ai += a.x * b.y - a.y * b.x;
  }
  for (int off = 16; off > 0; off >>= 1) { ar += __shfl_xor_sync(0xffffffffu, ar, off); ai += __shfl_xor_sync(0xffffffffu, ai, off); }
  if ((threadIdx.x & 31) == 0) { sr[threadIdx.x >> 5] = ar; si[threadIdx.x >> 5] = ai; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; k++) { ar += sr[k]; ai += si[k]; }
    Hd[(size_t)i * nb + r] = make_double2(ar, ai);
  }
}
// W[r] -= sum_i Hd[i][r] V_i[r]
__global__ void k_project(uint n, uint nb, uint nv, const cplx *__restrict__ Vb, const cplx *__restrict__ Hd, cplx *W) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint r = blockIdx.y;
  if (e >= n) return;
  cplx w = W[(size_t)r * n + e];
  for (uint i = 0; i < nv; i++) cfms(w, Hd[(size_t)i * nb + r], Vb[((size_t)i * nb + r) * n + e]);
  W[(size_t)r * n + e] = w;
}
// out[r] = W[r] * s[r]   (s real)
__global__ void k_scale_to(uint n, const cplx *__restrict__ W, const double *__restrict__ s, cplx *out) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint r = blockIdx.y;
  if (e >= n) return;
  cplx w = W[(size_t)r * n + e];
  out[(size_t)r * n + e] = make_double2(w.x * s[r], w.y * s[r]);
}
// X[r] += sum_j y[j][r] V_j[r]
__global__ void k_combine(uint n, uint nb, uint nv, const cplx *__restrict__ Vb, const cplx *__restrict__ y, cplx *X) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint r = blockIdx.y;
  if (e >= n) return;
  cplx x = X[(size_t)r * n + e];
  for (uint j = 0; j < nv; j++) cfma(x, y[(size_t)j * nb + r], Vb[((size_t)j * nb + r) * n + e]);
  X[(size_t)r * n + e] = x;
}
__global__ void k_sub(size_t tot, const cplx *__restrict__ A, const cplx *__restrict__ Bv, cplx *C) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < tot) C[e] = make_double2(A[e].x - Bv[e].x, A[e].y - Bv[e].y);
}
__global__ void k_rhs_batch(uint n, uint s0, const int64_t *__restrict__ indptr, const uint *__restrict__ idx,
                            const double *__restrict__ val, cplx *X) {
  uint col = s0 + blockIdx.x;
  for (int64_t p = indptr[col] + threadIdx.x; p < indptr[col + 1]; p += blockDim.x)
    X[(size_t)blockIdx.x * n + idx[p]] = make_double2(val[p], 0.0);
}
__global__ void __launch_bounds__(256)
k_epilogue_batch(uint nv, uint P, uint s0, cplx *X, const uint8_t *__restrict__ ob, const int64_t *__restrict__ a_indptr,
                 const uint *__restrict__ a_idx, const double *__restrict__ a_val, cplx *xpatch) {
  // x = conj(x); x[ob] = 0; x_patch[src][dst] = AVG[:, dst] . x   (generate_db.py:89-99)
  const uint src = s0 + blockIdx.x;
  cplx *x = X + (size_t)blockIdx.x * nv;
  for (uint i = threadIdx.x; i < nv; i += blockDim.x) {
    cplx v = x[i];
    v.y = -v.y;
    if (ob[i]) v = make_double2(0.0, 0.0);
    x[i] = v;
  }
  __syncthreads();
  const uint warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (uint dst = warp; dst < P; dst += nw) {
    double sr = 0.0, si = 0.0;
    for (int64_t p = a_indptr[dst] + lane; p < a_indptr[dst + 1]; p += 32) {
      cplx v = x[a_idx[p]];
      sr = fma(a_val[p], v.x, sr);
      si = fma(a_val[p], v.y, si);
    }
    for (int off = 16; off > 0; off >>= 1) { sr += __shfl_xor_sync(0xffffffffu, sr, off); si += __shfl_xor_sync(0xffffffffu, si, off); }
    if (lane == 0) xpatch[(size_t)src * P + dst] = make_double2(sr, si);
  }
}

// device buffer released on every exit path
template <typename T>
struct DevBuf {
  T *p = nullptr;
  ~DevBuf() { cudaFree(p); }
  int alloc(size_t count) {
    if (cudaMalloc(&p, sizeof(T) * (count ? count : 1)) != cudaSuccess) {
      p = nullptr;
      set_error("out of device memory (%zu bytes)", sizeof(T) * count);
      cudaGetLastError();
      return -1;
    }
    return 0;
  }
  T *release() { T *q = p; p = nullptr; return q; }
};

// ---- coarse space kernels: coarse DOF (b, j) = mode j restricted to diagonal block b ------------------------
// C[(b, j)][r] = sum_{i in block b} W[i][j] X[r][i]   (W real); grid (nblocks, nr), block 128 threads
__global__ void __launch_bounds__(128)
k_coarse_restrict(uint n, uint m, uint mtot, const uint *__restrict__ r0, const uint *__restrict__ sz,
                  const double *__restrict__ W, const cplx *__restrict__ X, cplx *C, size_t ldc) {
  __shared__ double sr[4][16], si[4][16];
  const uint b = blockIdx.x, r = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const cplx *x = X + (size_t)r * n + r0[b];
  const double *w = W + (size_t)r0[b] * mtot;
  double ar[16], ai[16];
#pragma unroll
  for (int j = 0; j < 16; j++) ar[j] = ai[j] = 0.0;
  for (uint i = threadIdx.x; i < sz[b]; i += 128) {
    const cplx v = x[i];
#pragma unroll
    for (int j = 0; j < 16; j++)
      if (j < (int)m) { const double wj = w[(size_t)i * mtot + j]; ar[j] = fma(wj, v.x, ar[j]); ai[j] = fma(wj, v.y, ai[j]); }
  }
#pragma unroll
  for (int j = 0; j < 16; j++) {
    for (int off = 16; off > 0; off >>= 1) { ar[j] += __shfl_xor_sync(0xffffffffu, ar[j], off); ai[j] += __shfl_xor_sync(0xffffffffu, ai[j], off); }
    if (lane == 0) { sr[warp][j] = ar[j]; si[warp][j] = ai[j]; }
  }
  __syncthreads();
  if (threadIdx.x < m) {
    const uint j = threadIdx.x;
    C[(size_t)r * ldc + (size_t)b * m + j] = make_double2(sr[0][j] + sr[1][j] + sr[2][j] + sr[3][j], si[0][j] + si[1][j] + si[2][j] + si[3][j]);
  }
}
// Y[r][i] (+)= sum_j W[i][j] C[(block of i, j)][r]
__global__ void k_coarse_prolong(uint n, uint m, uint mtot, const uint *__restrict__ blk_of, const double *__restrict__ W,
                                 const cplx *__restrict__ C, size_t ldc, cplx *Y, int accumulate) {
  const uint i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (i >= n) return;
  const cplx *c = C + (size_t)r * ldc + (size_t)blk_of[i] * m;
  const double *w = W + (size_t)i * mtot;
  cplx s = accumulate ? Y[(size_t)r * n + i] : make_double2(0.0, 0.0);
  for (uint j = 0; j < m; j++) { s.x = fma(w[j], c[j].x, s.x); s.y = fma(w[j], c[j].y, s.y); }
  Y[(size_t)r * n + i] = s;
}

struct Blocks {  // diagonal blocks of the FEM operator and their types (identical membranes share one)
  std::vector<uint> r0, sz, type;
  std::vector<uint> rep;  // representative block of each type
  bool exact = true;      // every entry of a block's rows lies inside the block (FEM operator == its block diagonal)
  struct Run { uint first, count; };   // consecutive blocks of one type: one batched ZGEMM
  std::vector<Run> runs;
};

// block-diagonal structure from the CSR pattern; blocks with identical K, M, B get one type
void detect_blocks(const cnld_ctx *c, Blocks &B) {
  const uint n = c->bem.nv;
  uint start = 0, reach = 0;
  for (uint i = 0; i < n; i++) {
    for (int64_t p = c->h_indptr[i]; p < c->h_indptr[i + 1]; p++) reach = std::max(reach, c->h_indices[p]);
    // tiny blocks (clamped boundary nodes have identity rows) are merged into the block that follows
    if (reach <= i && (i + 1 - start >= 8 || i + 1 == n)) {
      B.r0.push_back(start); B.sz.push_back(i + 1 - start); start = i + 1; reach = i + 1;
    }
  }
  auto same = [&](uint a, uint b) {
    if (B.sz[a] != B.sz[b]) return false;
    int64_t pa = c->h_indptr[B.r0[a]], pb = c->h_indptr[B.r0[b]];
    int64_t na = c->h_indptr[B.r0[a] + B.sz[a]] - pa, nb = c->h_indptr[B.r0[b] + B.sz[b]] - pb;
    if (na != nb) return false;
    for (int64_t q = 0; q < na; q++) {
      if (c->h_indices[pa + q] - B.r0[a] != c->h_indices[pb + q] - B.r0[b]) return false;
      if (c->h_K[pa + q] != c->h_K[pb + q] || c->h_M[pa + q] != c->h_M[pb + q] || c->h_B[pa + q] != c->h_B[pb + q]) return false;
    }
    return true;
  };
  for (uint b = 0; b < B.r0.size(); b++) {
    uint t = (uint)B.rep.size();
    for (uint k = 0; k < B.rep.size(); k++) if (same(B.rep[k], b)) { t = k; break; }
    if (t == B.rep.size()) B.rep.push_back(b);
    B.type.push_back(t);
  }
  for (uint b = 0; b < B.r0.size() && B.exact; b++)
    for (int64_t p = c->h_indptr[B.r0[b]]; p < c->h_indptr[B.r0[b] + B.sz[b]]; p++)
      if (c->h_indices[p] < B.r0[b] || c->h_indices[p] >= B.r0[b] + B.sz[b]) { B.exact = false; break; }
  for (uint b = 0; b < B.r0.size();) {   // blocks partition [0, n) in order, so same-type neighbours are contiguous
    uint e = b + 1;
    while (e < B.r0.size() && B.type[e] == B.type[b]) e++;
    B.runs.push_back({b, e - b});
    b = e;
  }
}

int g_hsweep_fluid_precond = 1;
int g_hsweep_coarse = 1;   // 0: ignore the coarse space; n > 1: use exactly n functions per membrane

struct Gmres {
  cnld_ctx *c;
  HMat *H;
  cudaStream_t st;
  uint n, nb, restart;
  double omega, coef, kwave = 0.0;
  int fluid_precond = 1;   // add the membrane's own block of coef * Z to the block-Jacobi preconditioner
  Blocks blk;
  std::vector<cplx *> d_inv;  // per block type: dense inverse
  cplx *Vb = nullptr, *W = nullptr, *T = nullptr, *X = nullptr, *Bp = nullptr, *Hd = nullptr, *yd = nullptr;
  double *sd = nullptr;

  // y|block = D_type x|block for every diagonal block: one batched ZGEMM per run of same-type blocks
  int apply_blockdiag(const std::vector<cplx *> &mats, const cplx *x, cplx *y, uint nr) {
    for (const Blocks::Run &R : blk.runs) {
      const uint b = R.first, sz = blk.sz[b];
      ZBatch bt;
      bt.per = R.count;
      bt.sBj = sz; bt.sCj = sz;   // operand A (the block) is shared by the run
      if (zgemm3m_batched(st, sz, nr, sz, 1.0, mats[blk.type[b]], sz, x + blk.r0[b], n, 0.0, y + blk.r0[b], n, R.count, bt))
        return -1;
    }
    return 0;
  }
  int apply_G(const cplx *x, cplx *y, uint nr) {  // y = MBK x + coef Z x
    // the FEM operator as the sparse matrix it is (round 1 applied it as one dense 761 x 761 block per
    // membrane: 38 times its flops)
    constexpr int RC = 8;
    const size_t work = (size_t)n * ((nr + RC - 1) / RC);
    k_spmm_mbk<RC><<<(unsigned)((work + 255) / 256), 256, 0, st>>>(n, nr, c->d_indptr, c->d_indices, c->d_K, c->d_M, c->d_B,
                                                                  omega, x, y);
    CNLD_CK_LAUNCH();
    if (coef != 0.0 && hmat_matvec(st, *H, make_double2(coef, 0.0), x, y, nr)) return -1;
    return 0;
  }
  int apply_P(const cplx *x, cplx *y, uint nr) {  // y = M^-1 x
    if (!cm) return apply_blockdiag(d_inv, x, y, nr);   // one level: blockdiag(MBK + own fluid block)^-1
    // Two-level (A-DEF2): q = W E^-1 W^H x (coarse correction in the span of the membranes' first cm
    // eigenmodes, E = W^H G W), then the block-Jacobi smoother on what the coarse space did not explain,
    // y = q + P (x - G q), with G q = (G W) c from the stored G W: no extra H-matvec.  The acoustic
    // crosstalk between membranes acts mostly through their low modes; with those handled exactly the
    // iteration count no longer grows with the array size or near the immersed resonances.
    const uint nc = nblk * cm;
    dim3 gr(nblk, nr);
    k_coarse_restrict<<<gr, 128, 0, st>>>(n, cm, cmtot, d_r0, d_sz, d_Wc, x, Cc, nc);
    CNLD_CK_LAUNCH();
    if (zgemm3m(st, nc, nr, nc, 1.0, Einv, nc, Cc, nc, 0.0, Cc2, nc)) return -1;
    CNLD_CK(cudaMemcpyAsync(Tm, x, sizeof(cplx) * (size_t)n * nr, cudaMemcpyDeviceToDevice, st));
    if (zgemm3m(st, n, nr, nc, -1.0, GW, n, Cc2, nc, 1.0, Tm, n)) return -1;
    if (apply_blockdiag(d_inv, Tm, y, nr)) return -1;
    dim3 gp((n + 127) / 128, nr);
    k_coarse_prolong<<<gp, 128, 0, st>>>(n, cm, cmtot, d_blk_of, d_Wc, Cc2, nc, y, 1);
    CNLD_CK_LAUNCH();
    return 0;
  }
  // coarse space (optional): cm of the cmtot functions per node are in use at this frequency
  uint cm = 0, cmtot = 0, nblk = 0;
  double *d_Wc = nullptr;
  uint *d_r0 = nullptr, *d_sz = nullptr, *d_blk_of = nullptr;
  cplx *GW = nullptr, *Einv = nullptr, *Cc = nullptr, *Cc2 = nullptr, *Tm = nullptr;
  size_t gw_cap = 0;
  int alloc_coarse(const std::vector<double> &Wh, uint mtot) {
    cmtot = mtot;
    nblk = (uint)blk.r0.size();
    std::vector<uint> blk_of(n, 0);
    for (uint b = 0; b < nblk; b++) for (uint i = 0; i < blk.sz[b]; i++) blk_of[blk.r0[b] + i] = b;
    const size_t ncmax = (size_t)nblk * mtot;
    CNLD_CK(cudaMalloc(&d_Wc, sizeof(double) * Wh.size()));
    CNLD_CK(cudaMemcpy(d_Wc, Wh.data(), sizeof(double) * Wh.size(), cudaMemcpyHostToDevice));
    CNLD_CK(cudaMalloc(&d_r0, sizeof(uint) * nblk));
    CNLD_CK(cudaMalloc(&d_sz, sizeof(uint) * nblk));
    CNLD_CK(cudaMalloc(&d_blk_of, sizeof(uint) * n));
    CNLD_CK(cudaMemcpy(d_r0, blk.r0.data(), sizeof(uint) * nblk, cudaMemcpyHostToDevice));
    CNLD_CK(cudaMemcpy(d_sz, blk.sz.data(), sizeof(uint) * nblk, cudaMemcpyHostToDevice));
    CNLD_CK(cudaMemcpy(d_blk_of, blk_of.data(), sizeof(uint) * n, cudaMemcpyHostToDevice));
    CNLD_CK(cudaMalloc(&GW, sizeof(cplx) * (size_t)n * ncmax));
    CNLD_CK(cudaMalloc(&Einv, sizeof(cplx) * ncmax * ncmax));
    CNLD_CK(cudaMalloc(&Cc, sizeof(cplx) * ncmax * std::max<size_t>(nb, 1)));
    CNLD_CK(cudaMalloc(&Cc2, sizeof(cplx) * ncmax * std::max<size_t>(nb, 1)));
    CNLD_CK(cudaMalloc(&Tm, sizeof(cplx) * (size_t)n * nb));
    return 0;
  }
  // per frequency: G W (through the H-matvec, nb columns at a time), E = W^H G W and its inverse
  int setup_coarse(uint m_use) {
    cm = 0;
    if (!cmtot || !m_use) return 0;
    const uint m = std::min(m_use, cmtot), nc = nblk * m;
    for (uint c0 = 0; c0 < nc; c0 += nb) {
      const uint nr = std::min(nb, nc - c0);
      // columns c0 .. c0 + nr of W: unit coarse vectors prolonged (X as scratch)
      CNLD_CK(cudaMemsetAsync(Cc, 0, sizeof(cplx) * (size_t)nc * nr, st));
      std::vector<cplx> unit((size_t)nc * nr, make_double2(0.0, 0.0));
      for (uint r = 0; r < nr; r++) unit[(size_t)r * nc + c0 + r] = make_double2(1.0, 0.0);
      CNLD_CK(cudaMemcpyAsync(Cc, unit.data(), sizeof(cplx) * unit.size(), cudaMemcpyHostToDevice, st));
      CNLD_CK(cudaStreamSynchronize(st));
      dim3 gp((n + 127) / 128, nr);
      k_coarse_prolong<<<gp, 128, 0, st>>>(n, m, cmtot, d_blk_of, d_Wc, Cc, nc, Tm, 0);
      CNLD_CK_LAUNCH();
      if (apply_G(Tm, GW + (size_t)c0 * n, nr)) return -1;
    }
    // E (nc x nc, column major) = W^H (G W), then E^-1 by LU against the identity
    DevBuf<cplx> E;
    if (E.alloc((size_t)nc * nc)) return -1;
    for (uint c0 = 0; c0 < nc; c0 += 4096) {
      const uint nr = std::min(4096u, nc - c0);
      dim3 gr(nblk, nr);
      k_coarse_restrict<<<gr, 128, 0, st>>>(n, m, cmtot, d_r0, d_sz, d_Wc, GW + (size_t)c0 * n, E.p + (size_t)c0 * nc, nc);
      CNLD_CK_LAUNCH();
    }
    k_identity<<<(unsigned)(((size_t)nc * nc + 255) / 256), 256, 0, st>>>(nc, Einv);
    CNLD_CK_LAUNCH();
    LuWork w;
    uint info = 0;
    const int lrc = (lu_work_alloc(w, nc) || lu_factor(st, w, E.p, nc) || lu_solve(st, w, E.p, nc, Einv, nc, nc) ||
                     cudaMemcpyAsync(&info, w.info, sizeof(uint), cudaMemcpyDeviceToHost, st) != cudaSuccess) ? -1 : 0;
    cudaStreamSynchronize(st);
    lu_work_free(w);
    if (lrc) return -1;
    if (info) return 0;      // singular coarse operator (e.g. f = 0 with a degenerate space): stay one-level
    cm = m;
    return 0;
  }
  int setup_precond() {
    for (uint t = 0; t < blk.rep.size(); t++) {
      const uint b = blk.rep[t], sz = blk.sz[b];
      if (sz > 8192) { set_error("FEM diagonal block of %u DOFs is too large for the block preconditioner", sz); return -1; }
      DevBuf<cplx> D, I;
      if (D.alloc((size_t)sz * sz) || I.alloc((size_t)sz * sz)) return -1;
      CNLD_CK(cudaMemsetAsync(D.p, 0, sizeof(cplx) * (size_t)sz * sz, st));
      k_block_dense<<<sz, 128, 0, st>>>(sz, blk.r0[b], c->d_indptr, c->d_indices, c->d_K, c->d_M, c->d_B, omega, D.p);
      CNLD_CK_LAUNCH();
      k_identity<<<(unsigned)(((size_t)sz * sz + 255) / 256), 256, 0, st>>>(sz, I.p);
      CNLD_CK_LAUNCH();
      if (fluid_precond && coef != 0.0) {
        // self-radiation of the block: its own dense block of Z (exact near-field quadrature), so the
        // preconditioner is (MBK + coef Z)_block^-1 -- the fluid loading a membrane puts on itself is
        // the largest part of coef Z and costs one sz x sz assembly per block type and frequency
        DevBuf<cplx> Zb;
        if (Zb.alloc((size_t)sz * sz)) return -1;
        std::vector<uint> idx(sz);
        for (uint i = 0; i < sz; i++) idx[i] = blk.r0[b] + i;
        if (nearfield_block(st, c->bem, c->h_t.data(), idx.data(), sz, idx.data(), sz, kwave, 0.0, Zb.p)) return -1;
        k_axpy_scaled<<<(unsigned)(((size_t)sz * sz + 255) / 256), 256, 0, st>>>((size_t)sz * sz, coef, Zb.p, D.p);
        CNLD_CK_LAUNCH();
        CNLD_CK(cudaStreamSynchronize(st));
      }
      LuWork w;
      const int lrc = (lu_work_alloc(w, sz) || lu_factor(st, w, D.p, sz) || lu_solve(st, w, D.p, sz, I.p, sz, sz)) ? -1 : 0;
      cudaStreamSynchronize(st);
      lu_work_free(w);
      if (lrc) return -1;
      cplx *Iq = I.release();
      if (t < d_inv.size()) { cudaFree(d_inv[t]); d_inv[t] = Iq; } else d_inv.push_back(Iq);
    }
    return 0;
  }
  int alloc() {
    const size_t v = (size_t)n * nb;
    CNLD_CK(cudaMalloc(&Vb, sizeof(cplx) * v * (restart + 1)));
    CNLD_CK(cudaMalloc(&W, sizeof(cplx) * v));
    CNLD_CK(cudaMalloc(&T, sizeof(cplx) * v));
    CNLD_CK(cudaMalloc(&X, sizeof(cplx) * v));
    CNLD_CK(cudaMalloc(&Bp, sizeof(cplx) * v));
    CNLD_CK(cudaMalloc(&Hd, sizeof(cplx) * (size_t)(restart + 2) * nb));
    CNLD_CK(cudaMalloc(&yd, sizeof(cplx) * (size_t)(restart + 1) * nb));
    CNLD_CK(cudaMalloc(&sd, sizeof(double) * nb));
    return 0;
  }
  void release() {
    cudaFree(Vb); cudaFree(W); cudaFree(T); cudaFree(X); cudaFree(Bp); cudaFree(Hd); cudaFree(yd); cudaFree(sd);
    cudaFree(d_Wc); cudaFree(d_r0); cudaFree(d_sz); cudaFree(d_blk_of); cudaFree(GW); cudaFree(Einv); cudaFree(Cc); cudaFree(Cc2); cudaFree(Tm);
    for (auto p : d_inv) cudaFree(p);
    d_inv.clear();
  }

  // Solve P G X = P B for nr right-hand sides held in B (device, [r][n]); result in X.
  int solve(const cplx *Bdev, uint nr, double tol, uint maxiter, uint *iters) {
    const size_t v = (size_t)n * nr;
    const unsigned gv = (unsigned)((v + 255) / 256);
    dim3 ge((n + 255) / 256, nr);
    std::vector<double> bnorm(nr), resid(nr), scal(nr);
    std::vector<zc> hh((size_t)(restart + 2) * nb);
    std::vector<char> done(nr, 0);
    std::vector<uint> it(nr, 0);
    CNLD_CK(cudaMemsetAsync(X, 0, sizeof(cplx) * v, st));
    if (apply_P(Bdev, Bp, nr)) return -1;
    auto norms = [&](const cplx *vec, std::vector<double> &out) -> int {
      dim3 g(nr, 1);
      k_dots<<<g, 256, 0, st>>>(n, nr, vec, vec, Hd);   // Vb index 0 -> vec itself (stride nb*n unused for i = 0)
      CNLD_CK_LAUNCH();
      CNLD_CK(cudaMemcpyAsync(hh.data(), Hd, sizeof(cplx) * nr, cudaMemcpyDeviceToHost, st));
      CNLD_CK(cudaStreamSynchronize(st));
      for (uint r = 0; r < nr; r++) out[r] = std::sqrt(std::max(hh[r].real(), 0.0));
      return 0;
    };
    if (norms(Bp, bnorm)) return -1;
    for (uint r = 0; r < nr; r++) if (bnorm[r] == 0.0) done[r] = 1;
    uint total = 0;
    while (total < maxiter) {
      // residual of the preconditioned system
      if (total == 0) {
        CNLD_CK(cudaMemcpyAsync(W, Bp, sizeof(cplx) * v, cudaMemcpyDeviceToDevice, st));
      } else {
        if (apply_G(X, T, nr) || apply_P(T, W, nr)) return -1;
        k_sub<<<gv, 256, 0, st>>>(v, Bp, W, W);
        CNLD_CK_LAUNCH();
      }
      if (norms(W, resid)) return -1;
      bool all = true;
      for (uint r = 0; r < nr; r++) {
        if (!done[r] && resid[r] <= tol * bnorm[r]) done[r] = 1;
        all &= (bool)done[r];
        scal[r] = resid[r] > 0.0 ? 1.0 / resid[r] : 0.0;
      }
      if (all) break;
      CNLD_CK(cudaMemcpyAsync(sd, scal.data(), sizeof(double) * nr, cudaMemcpyHostToDevice, st));
      k_scale_to<<<ge, 256, 0, st>>>(n, W, sd, Vb);
      CNLD_CK_LAUNCH();
      // per-rhs Hessenberg least squares by Givens rotations (host, tiny)
      const uint m = restart;
      std::vector<zc> Hm((size_t)nr * (m + 1) * m, zc(0)), gvec((size_t)nr * (m + 1), zc(0)), cs((size_t)nr * m), sn((size_t)nr * m);
      for (uint r = 0; r < nr; r++) gvec[(size_t)r * (m + 1)] = resid[r];
      uint j = 0;
      for (; j < m && total < maxiter; j++, total++) {
        const cplx *vj = Vb + (size_t)j * nb * n;
        if (apply_G(vj, T, nr) || apply_P(T, W, nr)) return -1;
        std::vector<zc> hcol((size_t)(j + 2) * nr, zc(0));
        for (int pass = 0; pass < 2; pass++) {  // classical Gram-Schmidt, twice
          dim3 gd(nr, j + 1);
          k_dots<<<gd, 256, 0, st>>>(n, nb, Vb, W, Hd);
          CNLD_CK_LAUNCH();
          k_project<<<ge, 256, 0, st>>>(n, nb, j + 1, Vb, Hd, W);
          CNLD_CK_LAUNCH();
          CNLD_CK(cudaMemcpyAsync(hh.data(), Hd, sizeof(cplx) * (size_t)(j + 1) * nb, cudaMemcpyDeviceToHost, st));
          CNLD_CK(cudaStreamSynchronize(st));
          for (uint i = 0; i <= j; i++)
            for (uint r = 0; r < nr; r++) hcol[(size_t)i * nr + r] += hh[(size_t)i * nb + r];
        }
        std::vector<double> wn(nr);
        if (norms(W, wn)) return -1;
        bool alldone = true;
        for (uint r = 0; r < nr; r++) {
          scal[r] = wn[r] > 0.0 ? 1.0 / wn[r] : 0.0;
          if (done[r]) continue;
          zc *Hr = Hm.data() + (size_t)r * (m + 1) * m, *gr = gvec.data() + (size_t)r * (m + 1);
          zc *c_ = cs.data() + (size_t)r * m, *s_ = sn.data() + (size_t)r * m;
          std::vector<zc> col(j + 2);
          for (uint i = 0; i <= j; i++) col[i] = hcol[(size_t)i * nr + r];
          col[j + 1] = wn[r];
          for (uint i = 0; i < j; i++) {
            zc t0 = std::conj(c_[i]) * col[i] + std::conj(s_[i]) * col[i + 1];
            col[i + 1] = -s_[i] * col[i] + c_[i] * col[i + 1];
            col[i] = t0;
          }
          double d = std::sqrt(std::norm(col[j]) + std::norm(col[j + 1]));
          if (d == 0.0) { c_[j] = 1.0; s_[j] = 0.0; }
          else { c_[j] = col[j] / d; s_[j] = col[j + 1] / d; }
          // rotation G = [conj(c) conj(s); -s c] maps (col_j, col_{j+1}) -> (d, 0)
          col[j] = d;
          col[j + 1] = 0.0;
          zc g0 = std::conj(c_[j]) * gr[j] + std::conj(s_[j]) * gr[j + 1];
          gr[j + 1] = -s_[j] * gr[j] + c_[j] * gr[j + 1];
          gr[j] = g0;
          for (uint i = 0; i <= j + 1 && i <= m; i++) if (i <= j) Hr[(size_t)i * m + j] = col[i];
          it[r]++;
          if (std::abs(gr[j + 1]) <= tol * bnorm[r]) done[r] = 2;  // converged inside this cycle
        }
        for (uint r = 0; r < nr; r++) alldone &= (done[r] != 0);
        CNLD_CK(cudaMemcpyAsync(sd, scal.data(), sizeof(double) * nr, cudaMemcpyHostToDevice, st));
        k_scale_to<<<ge, 256, 0, st>>>(n, W, sd, Vb + (size_t)(j + 1) * nb * n);
        CNLD_CK_LAUNCH();
        CNLD_CK(cudaStreamSynchronize(st));
        if (alldone) { j++; total++; break; }
      }
      // back substitution per rhs with the number of columns that rhs actually used
      std::vector<zc> yh((size_t)m * nb, zc(0));
      for (uint r = 0; r < nr; r++) {
        if (done[r] == 1) continue;  // was already converged before this cycle: no update
        zc *Hr = Hm.data() + (size_t)r * (m + 1) * m, *gr = gvec.data() + (size_t)r * (m + 1);
        uint kk = j;
        std::vector<zc> y(kk, zc(0));
        for (uint ii = kk; ii-- > 0;) {
          zc s = gr[ii];
          for (uint l = ii + 1; l < kk; l++) s -= Hr[(size_t)ii * m + l] * y[l];
          y[ii] = (Hr[(size_t)ii * m + ii] != zc(0)) ? s / Hr[(size_t)ii * m + ii] : zc(0);
        }
        for (uint l = 0; l < kk; l++) yh[(size_t)l * nb + r] = y[l];
        if (done[r] == 2) done[r] = 1;
      }
      CNLD_CK(cudaMemcpyAsync(yd, yh.data(), sizeof(cplx) * (size_t)m * nb, cudaMemcpyHostToDevice, st));
      k_combine<<<ge, 256, 0, st>>>(n, nb, j, Vb, yd, X);
      CNLD_CK_LAUNCH();
      CNLD_CK(cudaStreamSynchronize(st));
      // Every right-hand side converged inside this cycle by the Arnoldi residual estimate, and no restart has
      // happened: the estimate is the residual of X up to rounding (CGS2 keeps the basis orthogonal to 1e-15),
      // so the explicit residual of the next cycle -- a full operator application, a fifth of the work at 4
      // iterations -- is skipped.  After a restart the explicit residual is always formed.
      bool finished = true;
      for (uint r = 0; r < nr; r++) finished &= (done[r] == 1);
      if (finished && total <= m) break;
    }
    if (iters) for (uint r = 0; r < nr; r++) iters[r] = it[r];
    // Right-hand sides that ran out of iterations: measure the true preconditioned residual of what is
    // returned, so the caller gets a status instead of a silently unconverged solution (the reference's
    // lu() raises on a failed factorisation, compressed_formats.py:247-252).
    bool pending = false;
    for (uint r = 0; r < nr; r++) pending |= !done[r];
    if (pending) {
      if (apply_G(X, T, nr) || apply_P(T, W, nr)) return -1;
      k_sub<<<gv, 256, 0, st>>>(v, Bp, W, W);
      CNLD_CK_LAUNCH();
      if (norms(W, resid)) return -1;
      for (uint r = 0; r < nr; r++) {
        const double rel = bnorm[r] > 0.0 ? resid[r] / bnorm[r] : 0.0;
        if (!done[r] && rel > tol) { unconverged++; worst = std::max(worst, rel); }
      }
    }
    return 0;
  }
  unsigned unconverged = 0;   // right-hand sides above the tolerance after maxiter steps (whole sweep)
  double worst = 0.0;         // their largest relative residual
};
}  // namespace cnld_hapi
using namespace cnld_hapi;

extern "C" {

cnld_hmat *cnld_b200_hmat_create(cnld_ctx *c, cnld_uint clf, double eta, int admis, int strict, cnld_uint kmax) {
  if (!c) { set_error("null context"); return nullptr; }
  if (clf < 1 || !(eta > 0.0) || admis < 0 || admis > 3) { set_error("bad cluster / admissibility parameters"); return nullptr; }
  cnld_hmat *h = new cnld_hmat();
  h->ctx = c;
  if (hmat_create(h->H, c->bem, c->h_x.data(), c->h_t.data(), clf, eta, admis, strict != 0, kmax ? kmax : 64)) {
    delete h;
    return nullptr;
  }
  return h;
}

void cnld_b200_hmat_destroy(cnld_hmat *h) {
  if (!h) return;
  cudaSetDevice(h->ctx->bem.device);
  hmat_free(h->H);
  delete h;
}

int cnld_b200_hmat_info(cnld_hmat *h, double *out) {
  if (!h) { set_error("null H-matrix"); return -1; }
  HMat &H = h->H;
  double used = (double)H.near_entries, maxrank = 0;
  if (H.assembled) {
    if (hmat_fetch_leaves(H)) return -1;
    for (auto &L : H.leaf)
      if (L.adm) { used += (double)L.k * (L.m + L.n); maxrank = std::max(maxrank, (double)L.k); }
  }
  double nadm = 0, ninadm = 0;
  for (auto &L : H.leaf) (L.adm ? nadm : ninadm) += 1.0;
  out[0] = (double)H.tree.node.size(); out[1] = (double)H.leaf.size(); out[2] = nadm; out[3] = ninadm;
  out[4] = (double)H.pool_size; out[5] = used; out[6] = maxrank; out[7] = H.n;
  return 0;
}

int cnld_b200_hmat_sharing(cnld_hmat *h, double *out) {
  if (!h) { set_error("null H-matrix"); return -1; }
  HMat &H = h->H;
  out[0] = H.nfar; out[1] = H.nnear; out[2] = (double)H.leaf.size() - H.nfar - H.nnear; out[3] = (double)H.pool_size;
  return 0;
}

int cnld_b200_hmat_tree(cnld_hmat *h, cnld_uint *perm, int *son0, int *son1, cnld_uint *start, cnld_uint *size,
                        double *bmin, double *bmax, cnld_uint *leaf_rc, cnld_uint *leaf_cc, uint8_t *leaf_adm,
                        cnld_uint *leaf_rank) {
  if (!h) { set_error("null H-matrix"); return -1; }
  HMat &H = h->H;
  if (perm) memcpy(perm, H.tree.perm.data(), sizeof(uint) * H.n);
  for (size_t i = 0; i < H.tree.node.size(); i++) {
    const ClusterNode &N = H.tree.node[i];
    if (son0) son0[i] = N.son[0];
    if (son1) son1[i] = N.son[1];
    if (start) start[i] = N.start;
    if (size) size[i] = N.size;
    for (int d = 0; d < 3; d++) { if (bmin) bmin[3 * i + d] = N.bmin[d]; if (bmax) bmax[3 * i + d] = N.bmax[d]; }
  }
  if (leaf_rank && H.assembled && hmat_fetch_leaves(H)) return -1;
  for (size_t b = 0; b < H.blocks.size(); b++) {
    if (leaf_rc) leaf_rc[b] = H.blocks[b].rc;
    if (leaf_cc) leaf_cc[b] = H.blocks[b].cc;
    if (leaf_adm) leaf_adm[b] = H.blocks[b].admissible;
    if (leaf_rank) leaf_rank[b] = H.leaf[b].adm ? H.leaf[b].k : 0;
  }
  return 0;
}

int cnld_b200_hmat_assemble(cnld_hmat *h, double k_re, double k_im, double eps_aca) {
  if (!h) { set_error("null H-matrix"); return -1; }
  CNLD_CK(cudaSetDevice(h->ctx->bem.device));
  return hmat_assemble(0, h->H, k_re, k_im, eps_aca);
}

int cnld_b200_hmat_matvec(cnld_hmat *h, cnld_field alpha, const cnld_field *x, cnld_field *y, cnld_uint nrhs) {
  if (!h) { set_error("null H-matrix"); return -1; }
  CNLD_CK(cudaSetDevice(h->ctx->bem.device));
  const size_t tot = (size_t)h->H.n * nrhs;
  cplx *dx = nullptr, *dy = nullptr;
  CNLD_CK(cudaMalloc(&dx, sizeof(cplx) * tot));
  CNLD_CK(cudaMalloc(&dy, sizeof(cplx) * tot));
  int rc = 0;
  do {
    if (cudaMemcpy(dx, x, sizeof(cplx) * tot, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(dy, y, sizeof(cplx) * tot, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("H2D failed"); rc = -1; break; }
    if (hmat_matvec(0, h->H, make_double2(alpha.re, alpha.im), dx, dy, nrhs)) { rc = -1; break; }
    cudaError_t e = cudaMemcpy(y, dy, sizeof(cplx) * tot, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("hmat_matvec: %s", cudaGetErrorString(e)); rc = -1; }
  } while (0);
  cudaFree(dx); cudaFree(dy);
  return rc;
}

// device-resident timing of the matvec (seconds per call, CUDA events on the launch stream, after one
// untimed call); the vectors hold a fixed pseudo-random pattern
int cnld_b200_hmat_bench_matvec(cnld_hmat *h, cnld_uint nrhs, int reps, double *seconds) {
  if (!h || !seconds || !nrhs || reps < 1) { set_error("bench_matvec: bad arguments"); return -1; }
  CNLD_CK(cudaSetDevice(h->ctx->bem.device));
  const size_t tot = (size_t)h->H.n * nrhs;
  std::vector<cplx> hx(tot);
  unsigned long long sd = 88172645463325252ull;
  for (size_t i = 0; i < tot; i++) {
    sd ^= sd << 13; sd ^= sd >> 7; sd ^= sd << 17;
    hx[i] = make_double2((double)(sd & 0xffff) / 65536.0 - 0.5, (double)((sd >> 16) & 0xffff) / 65536.0 - 0.5);
  }
  cplx *dx = nullptr, *dy = nullptr;
  CNLD_CK(cudaMalloc(&dx, sizeof(cplx) * tot));
  CNLD_CK(cudaMalloc(&dy, sizeof(cplx) * tot));
  cudaEvent_t e0, e1;
  CNLD_CK(cudaEventCreate(&e0));
  CNLD_CK(cudaEventCreate(&e1));
  int rc = -1;
  do {
    if (cudaMemcpy(dx, hx.data(), sizeof(cplx) * tot, cudaMemcpyHostToDevice) != cudaSuccess) break;
    if (cudaMemset(dy, 0, sizeof(cplx) * tot) != cudaSuccess) break;
    if (hmat_matvec(0, h->H, make_double2(1.0, 0.0), dx, dy, nrhs)) break;
    cudaEventRecord(e0, 0);
    bool ok = true;
    for (int r = 0; r < reps && ok; r++) ok = hmat_matvec(0, h->H, make_double2(1.0, 0.0), dx, dy, nrhs) == 0;
    cudaEventRecord(e1, 0);
    if (!ok || cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *seconds = ms * 1e-3 / reps;
    rc = 0;
  } while (0);
  if (rc) set_error("bench_matvec: %s", cudaGetErrorString(cudaGetLastError()));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(dx); cudaFree(dy);
  return rc;
}

int cnld_b200_hmat_todense(cnld_hmat *h, cnld_field *Z, cnld_uint ld) {
  if (!h) { set_error("null H-matrix"); return -1; }
  CNLD_CK(cudaSetDevice(h->ctx->bem.device));
  const size_t n = h->H.n;
  cplx *dZ = nullptr;
  CNLD_CK(cudaMalloc(&dZ, sizeof(cplx) * n * n));
  int rc = 0;
  do {
    if (hmat_expand(0, h->H, dZ, n)) { rc = -1; break; }
    cudaError_t e = cudaMemcpy2D(Z, sizeof(cplx) * ld, dZ, sizeof(cplx) * n, sizeof(cplx) * n, n, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("hmat_todense: %s", cudaGetErrorString(e)); rc = -1; }
  } while (0);
  cudaFree(dZ);
  return rc;
}

int cnld_b200_hsweep_run(cnld_ctx *c, cnld_hmat *h, const double *freqs, cnld_uint nf, double cs, double rho,
                         double eps_aca, double tol, cnld_uint restart, cnld_uint maxiter, cnld_uint batch,
                         cnld_field *x_patch, cnld_field *x_node, cnld_uint *iters, double *t_freq) {
  if (!c || !h || h->ctx != c) { set_error("hsweep: context / H-matrix mismatch"); return -1; }
  if (!c->d_indptr || !c->d_f_indptr) { set_error("cnld_b200_set_mbk / set_loads have not been called"); return -1; }
  CNLD_CK(cudaSetDevice(c->bem.device));
  const uint n = c->bem.nv, P = c->P;
  struct GmresGuard {
    Gmres G;
    ~GmresGuard() { G.release(); }
  } guard;
  Gmres &G = guard.G;
  G.c = c; G.H = &h->H; G.st = 0; G.n = n;
  G.nb = std::min(batch ? batch : 32u, P);
  G.restart = restart ? restart : 50;
  detect_blocks(c, G.blk);
  if (G.alloc()) return -1;
  if (h->coarse_m && g_hsweep_coarse) {
    if (h->coarse_W.size() != (size_t)n * h->coarse_m) { set_error("hsweep: coarse space does not match the mesh"); return -1; }
    if (h->coarse_m > 16) { set_error("hsweep: at most 16 coarse functions per node"); return -1; }
    if (G.alloc_coarse(h->coarse_W, h->coarse_m)) return -1;
  }
  DevBuf<cplx> dB, dxp;
  if (dB.alloc((size_t)n * G.nb) || dxp.alloc((size_t)P * P)) return -1;
  struct Events {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~Events() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
  } ev;
  CNLD_CK(cudaEventCreate(&ev.e0));
  CNLD_CK(cudaEventCreate(&ev.e1));
  int rc = 0;
  h->last_capped = 0;
  for (uint i = 0; i < nf && !rc; i++) {
    const double f = freqs[i], omega = H_TWO_PI * f;
    G.omega = omega;
    G.coef = -omega * omega * 2.0 * rho;
    G.kwave = omega / cs;
    G.fluid_precond = g_hsweep_fluid_precond;
    cudaEventRecord(ev.e0, 0);
    if (f != 0.0) {
      if (hmat_assemble(0, h->H, omega / cs, 0.0, eps_aca)) { rc = -1; break; }
      h->last_capped += h->H.capped_leaves;
    }
    if (G.setup_precond()) { rc = -1; break; }
    if (G.cmtot) {
      // modes whose in-vacuo resonance lies below ~2.2 f matter at f (immersion lowers the resonances); at least 4.
      // Measured at the 16x16 array (scripts/time_hsweep.py): 8.6 MHz m = 2 / 4 / 6 -> 2.73 / 2.56 / 2.76 s,
      // 28.6 MHz m = 4 / 6 / 10 -> 5.1 / 3.6 / 5.2 s, 42.3 MHz m = 6 / 10 / 16 -> 7.8 / 4.5 / 5.4 s: too few modes cost
      // iterations, too many cost the dense G W product of every application.
      uint m_use = 0;
      for (uint j = 0; j < h->coarse_m; j++) if (h->coarse_f[j] <= 2.2 * f) m_use = j + 1;
      m_use = std::max(m_use, std::min(4u, h->coarse_m));
      if (g_hsweep_coarse > 1) m_use = std::min<uint>(g_hsweep_coarse, h->coarse_m);   // debug: fixed count
      if (f == 0.0) m_use = 0;
      if (G.setup_coarse(m_use)) { rc = -1; break; }
      h->last_coarse_m = G.cm;
    }
    for (uint s0 = 0; s0 < P && !rc; s0 += G.nb) {
      const uint nr = std::min(G.nb, P - s0);
      cudaMemsetAsync(dB.p, 0, sizeof(cplx) * (size_t)n * nr, 0);
      k_rhs_batch<<<nr, 128>>>(n, s0, c->d_f_indptr, c->d_f_idx, c->d_f_val, dB.p);
      cnld::count_launch();
      if (G.solve(dB.p, nr, tol, maxiter ? maxiter : 500, iters ? iters + (size_t)i * P + s0 : nullptr)) { rc = -1; break; }
      k_epilogue_batch<<<nr, 256>>>(n, P, s0, G.X, c->d_ob, c->d_a_indptr, c->d_a_idx, c->d_a_val, dxp.p);
      cnld::count_launch();
      if (x_node && cudaMemcpy(reinterpret_cast<cplx *>(x_node) + ((size_t)i * P + s0) * n, G.X, sizeof(cplx) * (size_t)n * nr,
                               cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("D2H failed"); rc = -1; }
    }
    cudaEventRecord(ev.e1, 0);
    if (x_patch && cudaMemcpy(reinterpret_cast<cplx *>(x_patch) + (size_t)i * P * P, dxp.p, sizeof(cplx) * (size_t)P * P,
                              cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("D2H failed"); rc = -1; }
    if (cudaEventSynchronize(ev.e1) != cudaSuccess) { set_error("hsweep: %s", cudaGetErrorString(cudaGetLastError())); rc = -1; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev.e0, ev.e1);
    if (t_freq) t_freq[i] = ms * 1e-3;
  }
  if (rc) { cudaDeviceSynchronize(); return rc; }
  if (G.unconverged) {
    // results are in the output buffers; the caller decides (rc 2 = "not converged", not a CUDA failure)
    set_error("GMRES: %u right-hand side(s) did not reach tol %.1e within %u iterations (largest relative residual %.2e)",
              G.unconverged, tol, maxiter ? maxiter : 500u, G.worst);
    return 2;
  }
  return 0;
}

/* Coarse space of the two-level preconditioner: W[n][m] (node major, m <= 16) = the first m eigenmodes of
 * every membrane's in-vacuo plate problem K v = lambda M v on its own nodes (zero on clamped nodes and on
 * every other membrane), fmode[m] = their resonance frequencies in Hz.  m = 0 removes it. */
int cnld_b200_hmat_set_coarse(cnld_hmat *h, cnld_uint m, const double *W, const double *fmode) {
  if (!h || (m && (!W || !fmode)) || m > 16) { set_error("hmat_set_coarse: bad arguments (m <= 16)"); return -1; }
  h->coarse_m = m;
  h->coarse_W.assign(W, W + (size_t)h->H.n * m);
  h->coarse_f.assign(fmode, fmode + m);
  return 0;
}
int cnld_b200_hmat_coarse_used(cnld_hmat *h) { return h ? (int)h->last_coarse_m : -1; }

int cnld_b200_hmat_capped(cnld_hmat *h, unsigned *last_assembly, unsigned *last_sweep) {
  if (!h) { set_error("null H-matrix"); return -1; }
  if (last_assembly) *last_assembly = h->H.capped_leaves;
  if (last_sweep) *last_sweep = h->last_capped;
  return 0;
}

}  // extern "C"
