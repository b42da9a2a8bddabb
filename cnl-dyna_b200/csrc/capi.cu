// C ABI of libcnld_b200.so (see include/cnld_b200.h) and the per-frequency sweep
// pipeline: init G with K - w^2 M - j w B -> accumulate (-2 rho w^2) Z(k) -> unpivoted LU
// -> solve all source patches -> conj / boundary mask / patch average
// (cnld/scripts/generate_db.py:30-130).
#include "../../include/cnld_b200.h"

#include <cmath>
#include <mutex>

#include "hmat.cuh"

namespace cnld {
std::atomic<unsigned long long> g_launches{0};
std::atomic<unsigned long long> g_dmma_flops{0};
static thread_local std::string t_err;
static std::string g_err;
static std::mutex g_err_mu;
void set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  std::lock_guard<std::mutex> lk(g_err_mu);
  g_err = buf;
}

}  // namespace cnld

#include "ctx.cuh"
#include <algorithm>
#include <map>
using namespace cnld;
using namespace cnld::sweep;

namespace cnld_capi {

template <bool ADD>
__global__ void k_scatter_mbk(uint nv, const int64_t *__restrict__ indptr, const uint *__restrict__ indices,
                              const double *__restrict__ K, const double *__restrict__ M,
                              const double *__restrict__ B, double omega, cplx *G, size_t ld) {
  // one warp per row: G[i,j] = K - w^2 M - j w B   (cnld/fem.py:1305)
  uint row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint lane = threadIdx.x & 31;
  if (row >= nv) return;
  double w2 = omega * omega;
  for (int64_t p = indptr[row] + lane; p < indptr[row + 1]; p += 32) {
    cplx *d = G + (size_t)indices[p] * ld + row;
    const cplx v = make_double2(K[p] - w2 * M[p], -omega * B[p]);
    if (ADD) { d->x += v.x; d->y += v.y; }   // pattern entries are unique: no atomics needed
    else *d = v;
  }
}

__global__ void k_build_rhs(uint P, const int64_t *__restrict__ indptr, const uint *__restrict__ idx,
                            const double *__restrict__ val, cplx *X, size_t ldx) {
  uint col = blockIdx.x;
  if (col >= P) return;
  for (int64_t p = indptr[col] + threadIdx.x; p < indptr[col + 1]; p += blockDim.x)
    X[(size_t)col * ldx + idx[p]] = make_double2(val[p], 0.0);
}

// x = conj(x); x[ob] = 0; x_patch[src][dst] = sum_i AVG[i,dst] x[i]   (generate_db.py:89-99)
__global__ void __launch_bounds__(256)
k_epilogue(uint nv, uint P, cplx *X, size_t ldx, const uint8_t *__restrict__ ob,
           const int64_t *__restrict__ a_indptr, const uint *__restrict__ a_idx,
           const double *__restrict__ a_val, cplx *xpatch) {
  const uint src = blockIdx.x;
  cplx *x = X + (size_t)src * ldx;
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
    for (int off = 16; off > 0; off >>= 1) {
      sr += __shfl_xor_sync(0xffffffffu, sr, off);
      si += __shfl_xor_sync(0xffffffffu, si, off);
    }
    if (lane == 0) xpatch[(size_t)src * P + dst] = make_double2(sr, si);
  }
}

// device copy of a host array; the allocation is kept when the new contents fit (no cudaFree /
// cudaMalloc -- both synchronise the device -- when a sweep re-uploads operators of the same size)
std::mutex g_cap_mu;
std::map<const void *, size_t> g_cap;
template <typename T>
int upload(T *&d, const T *h, size_t n) {
  const size_t bytes = sizeof(T) * (n ? n : 1);
  {
    std::lock_guard<std::mutex> lk(g_cap_mu);
    auto it = d ? g_cap.find(d) : g_cap.end();
    if (d && (it == g_cap.end() || it->second < bytes)) {
      if (it != g_cap.end()) g_cap.erase(it);
      cudaFree(d);
      d = nullptr;
    }
    if (!d) {
      CNLD_CK(cudaMalloc(&d, bytes));
      g_cap[d] = bytes;
    }
  }
  if (n) CNLD_CK(cudaMemcpy(d, h, sizeof(T) * n, cudaMemcpyHostToDevice));
  return 0;
}
template <typename T>
void dev_free(T *&d) {
  if (!d) return;
  {
    std::lock_guard<std::mutex> lk(g_cap_mu);
    g_cap.erase(d);
  }
  cudaFree(d);
  d = nullptr;
}

int ensure_workspace(cnld_ctx *c, bool need_host_stage) {
  const size_t nv = c->bem.nv;
  CNLD_CK(cudaSetDevice(c->bem.device));
  for (int s = 0; s < c->nslots; s++) {
    Slot &S = c->slots[s];
    if (!S.st) {
      int lo_p = 0, hi_p = 0;
      CNLD_CK(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));   // (least, greatest); greatest is numerically lowest
      CNLD_CK(cudaStreamCreateWithPriority(&S.st, cudaStreamNonBlocking, hi_p));
      // trailing updates: a different priority per slot, so the slots cannot stay in lock-step
      // (one factorisation owns the tensor pipe while the other slot assembles on the FP64 pipe)
      int mid = hi_p + 1 + s;
      if (mid > lo_p - 1) mid = lo_p - 1;
      if (mid < hi_p) mid = hi_p;
      CNLD_CK(cudaStreamCreateWithPriority(&S.st_lo, cudaStreamNonBlocking, mid));
      CNLD_CK(cudaStreamCreateWithPriority(&S.st_asm, cudaStreamNonBlocking, lo_p));
      CNLD_CK(cudaEventCreateWithFlags(&S.ev_fork, cudaEventDisableTiming));
      CNLD_CK(cudaEventCreateWithFlags(&S.ev_join, cudaEventDisableTiming));
    }
    if (!S.G) CNLD_CK(cudaMalloc(&S.G, sizeof(cplx) * nv * nv));
    if (!S.X) CNLD_CK(cudaMalloc(&S.X, sizeof(cplx) * nv * (c->P ? c->P : 1)));
    if (c->symmetric && !S.E) {
      CNLD_CK(cudaMalloc(&S.E, sizeof(cplx) * nv * (c->P ? c->P : 1)));
      CNLD_CK(cudaMalloc(&S.E2, sizeof(cplx) * nv * (c->P ? c->P : 1)));
      CNLD_CK(cudaMalloc(&S.sval, sizeof(cplx) * (c->bem.nslot + 1)));
      CNLD_CK(cudaMalloc(&S.norms, 2 * sizeof(double)));
      CNLD_CK(cudaMalloc(&S.dval, sizeof(cplx) * (c->bem.nslot + 1)));
    }
    if (!S.lu.linv) {
      if (lu_work_alloc(S.lu, (uint)nv)) return -1;
      S.lu.lo = S.st_lo;   // companion stream for the look-ahead trailing updates
      CNLD_CK(cudaEventCreateWithFlags(&S.lu.ev_fork, cudaEventDisableTiming));
      CNLD_CK(cudaEventCreateWithFlags(&S.lu.ev_join, cudaEventDisableTiming));
    }
    lu_profile(S.lu, c->profile);
    for (int e = 0; e < 4; e++)
      if (!S.ev[e]) CNLD_CK(cudaEventCreate(&S.ev[e]));
    if (need_host_stage && !S.hX) CNLD_CK(cudaMallocHost(&S.hX, sizeof(cplx) * nv * c->P));
  }
  c->ws_ready = true;
  return 0;
}

void free_workspace(cnld_ctx *c) {
  for (int s = 0; s < MAX_SLOTS; s++) {
    Slot &S = c->slots[s];
    if (S.st) cudaStreamSynchronize(S.st);
    cudaFree(S.G); cudaFree(S.X); cudaFree(S.E); cudaFree(S.E2); cudaFree(S.norms); cudaFree(S.sval); cudaFree(S.dval);
    if (S.hX) cudaFreeHost(S.hX);
    lu_work_free(S.lu);
    for (int e = 0; e < 4; e++) if (S.ev[e]) cudaEventDestroy(S.ev[e]);
    if (S.st) cudaStreamDestroy(S.st);
    if (S.st_lo) cudaStreamDestroy(S.st_lo);
    if (S.st_asm) cudaStreamDestroy(S.st_asm);
    if (S.ev_fork) cudaEventDestroy(S.ev_fork);
    if (S.ev_join) cudaEventDestroy(S.ev_join);
    S = Slot();
  }
  c->ws_ready = false;
}

// R(:, p) = -Delta * E(:, p).  Delta is strictly upper and sparse: the Sauter-Schwab asymmetry sits in
// dval at the slots with row < col (slots sorted by (row, col)), the asymmetry of the FEM operators (if
// any) in a second CSR.  One CTA per row, one thread per right-hand side.
__global__ void k_delta_apply(uint nv, uint P, const uint *__restrict__ rowptr, const uint *__restrict__ col,
                              const cplx *__restrict__ dval, const uint *__restrict__ m_rowptr,
                              const uint *__restrict__ m_col, const double *__restrict__ m_dK,
                              const double *__restrict__ m_dM, const double *__restrict__ m_dB, double omega,
                              const cplx *__restrict__ E, cplx *R, size_t ld) {
  const uint r = blockIdx.x;
  const uint s0 = rowptr[r], s1 = rowptr[r + 1];
  const uint q0 = m_rowptr ? m_rowptr[r] : 0, q1 = m_rowptr ? m_rowptr[r + 1] : 0;
  const double w2 = omega * omega;
  for (uint p = threadIdx.x; p < P; p += blockDim.x) {
    cplx acc = make_double2(0.0, 0.0);
    for (uint s = s0; s < s1; s++) {
      const uint cidx = col[s];
      if (cidx > r) cfms(acc, dval[s], E[(size_t)p * ld + cidx]);
    }
    for (uint q = q0; q < q1; q++)
      cfms(acc, make_double2(m_dK[q] - w2 * m_dM[q], -omega * m_dB[q]), E[(size_t)p * ld + m_col[q]]);
    R[(size_t)p * ld + r] = acc;
  }
}
// x += e; norms[0] += |e|^2, norms[1] += |x|^2  (convergence estimate of the refinement)
__global__ void __launch_bounds__(256) k_axpy(size_t n, const cplx *__restrict__ e, cplx *x, double *norms) {
  double ne = 0.0, nx = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    cplx v = x[i], w = e[i];
    v = make_double2(v.x + w.x, v.y + w.y);
    x[i] = v;
    ne += w.x * w.x + w.y * w.y;
    nx += v.x * v.x + v.y * v.y;
  }
  for (int o = 16; o; o >>= 1) { ne += __shfl_xor_sync(0xffffffffu, ne, o); nx += __shfl_xor_sync(0xffffffffu, nx, o); }
  __shared__ double sh[2][8];
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = ne; sh[1][threadIdx.x >> 5] = nx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++) { ne += sh[0][w]; nx += sh[1][w]; }
    atomicAdd(norms, ne); atomicAdd(norms + 1, nx);
  }
}

// enqueue one frequency on a slot; results: xpatch slice (device), X holds x_node
int enqueue_frequency(cnld_ctx *c, Slot &S, double f, double cs, double rho, cplx *d_xpatch_f, bool allow_sym = true) {
  const uint nv = c->bem.nv, P = c->P;
  const double omega = TWO_PI * f, k = omega / cs;
  // symmetric path: G = S + Delta with S symmetric (lower triangle of G mirrored) and Delta the sparse,
  // strictly upper asymmetry of the Sauter-Schwab entries.  S is factored at half the cost and
  // x = G^-1 f is recovered by the stationary iteration e <- -S^-1 Delta e, x <- x + e.
  const bool sym = c->symmetric != 0 && allow_sym;
  S.was_sym = sym;
  CNLD_CK(cudaEventRecord(S.ev[0], S.st));
  CNLD_CK(cudaMemsetAsync(S.G, 0, sizeof(cplx) * (size_t)nv * nv, S.st));
  if (!sym) {
    k_scatter_mbk<false><<<(nv * 32 + 255) / 256, 256, 0, S.st>>>(nv, c->d_indptr, c->d_indices, c->d_K, c->d_M,
                                                               c->d_B, omega, S.G, nv);
    CNLD_CK_LAUNCH();
  }
  if (sym && f == 0.0) CNLD_CK(cudaMemsetAsync(S.dval, 0, sizeof(cplx) * (c->bem.nslot + 1), S.st));
  if (f != 0.0) {
    // Gbe = -omg^2 * 2 * rho * Z   (generate_db.py:55-56)
    cudaStream_t as = S.st;
    if (c->nslots > 1) {  // throughput kernel -> low-priority companion stream
      CNLD_CK(cudaEventRecord(S.ev_fork, S.st));
      CNLD_CK(cudaStreamWaitEvent(S.st_asm, S.ev_fork, 0));
      as = S.st_asm;
    }
    const cplx alpha = make_double2(-omega * omega * 2.0 * rho, 0.0);
    if (sym ? bem_assemble_sym(as, c->bem, k, 0.0, alpha, S.G, nv, S.sval, S.dval)
            : bem_assemble(as, c->bem, k, 0.0, alpha, S.G, nv)) return -1;
    if (c->nslots > 1) {
      CNLD_CK(cudaEventRecord(S.ev_join, S.st_asm));
      CNLD_CK(cudaStreamWaitEvent(S.st, S.ev_join, 0));
    }
  }
  if (sym) {  // after the assembly: the translation-class copies overwrite whole blocks of G
    k_scatter_mbk<true><<<(nv * 32 + 255) / 256, 256, 0, S.st>>>(nv, c->d_indptr, c->d_indices, c->d_K, c->d_M,
                                                              c->d_B, omega, S.G, nv);
    CNLD_CK_LAUNCH();
  }
  CNLD_CK(cudaEventRecord(S.ev[1], S.st));
  if (sym ? lu_factor_sym(S.st, S.lu, S.G, nv) : lu_factor(S.st, S.lu, S.G, nv)) return -1;
  CNLD_CK(cudaEventRecord(S.ev[2], S.st));
  CNLD_CK(cudaMemsetAsync(S.X, 0, sizeof(cplx) * (size_t)nv * P, S.st));
  k_build_rhs<<<P, 128, 0, S.st>>>(P, c->d_f_indptr, c->d_f_idx, c->d_f_val, S.X, nv);
  CNLD_CK_LAUNCH();
  if (lu_solve(S.st, S.lu, S.G, nv, S.X, P, nv, c->f_first_row.empty() ? nullptr : c->f_first_row.data())) return -1;
  {
    const cplx *src = S.X;
    cplx *dst = S.E, *other = S.E2;
    for (int it = 0; sym && it < c->refine_steps; it++) {
      k_delta_apply<<<nv, 160, 0, S.st>>>(nv, P, c->bem.slot_rowptr, c->bem.slot_col, S.dval, c->d_m_rowptr,
                                          c->d_m_col, c->d_m_dK, c->d_m_dM, c->d_m_dB, omega, src, dst, nv);
      CNLD_CK_LAUNCH();
      if (lu_solve(S.st, S.lu, S.G, nv, dst, P, nv)) return -1;
      const size_t tot = (size_t)nv * P;
      CNLD_CK(cudaMemsetAsync(S.norms, 0, 2 * sizeof(double), S.st));
      k_axpy<<<1184, 256, 0, S.st>>>(tot, dst, S.X, S.norms);
      CNLD_CK_LAUNCH();
      src = dst;
      cplx *t = dst; dst = other; other = t;
    }
  }
  k_epilogue<<<P, 256, 0, S.st>>>(nv, P, S.X, nv, c->d_ob, c->d_a_indptr, c->d_a_idx, c->d_a_val,
                                  d_xpatch_f);
  CNLD_CK_LAUNCH();
  CNLD_CK(cudaEventRecord(S.ev[3], S.st));
  return 0;
}

int check_ready(cnld_ctx *c) {
  if (!c) { set_error("null context"); return -1; }
  if (!c->d_indptr) { set_error("cnld_b200_set_mbk has not been called"); return -1; }
  if (!c->d_f_indptr) { set_error("cnld_b200_set_loads has not been called"); return -1; }
  return 0;
}

int sweep_inner(cnld_ctx *c, const double *freqs, uint nf, double cs, double rho, cnld_field *x_patch,
                cnld_field *x_node, double *t_freq, bool device_only);
// On any failure the other slots may still have kernels and asynchronous copies in flight that target
// the caller's buffers: wait for every stream before the error is returned (the caller is then free to
// drop its arrays), and forget the half-finished frequencies.
int sweep_impl(cnld_ctx *c, const double *freqs, uint nf, double cs, double rho, cnld_field *x_patch,
               cnld_field *x_node, double *t_freq, bool device_only) {
  const int rc = sweep_inner(c, freqs, nf, cs, rho, x_patch, x_node, t_freq, device_only);
  if (rc && c) {
    const std::string msg = cnld_b200_last_error();
    for (int s = 0; s < c->nslots; s++) {
      Slot &S = c->slots[s];
      if (S.st) cudaStreamSynchronize(S.st);
      if (S.st_lo) cudaStreamSynchronize(S.st_lo);
      if (S.st_asm) cudaStreamSynchronize(S.st_asm);
      S.pending = -1;
    }
    cudaGetLastError();
    set_error("%s", msg.c_str());
  }
  return rc;
}

int sweep_inner(cnld_ctx *c, const double *freqs, uint nf, double cs, double rho, cnld_field *x_patch,
                cnld_field *x_node, double *t_freq, bool device_only) {
  if (check_ready(c)) return -1;
  const size_t nv = c->bem.nv, P = c->P;
  // Node displacements go straight into the caller's buffer when it lies inside the region the caller
  // page-locked with cnld_b200_pin_host; otherwise through pinned staging + a host copy.
  bool direct = false;
  if (x_node && nf && c->reg_ptr) {
    const char *lo = reinterpret_cast<const char *>(x_node), *hi = lo + sizeof(cplx) * (size_t)nf * P * nv;
    const char *rlo = reinterpret_cast<const char *>(c->reg_ptr);
    direct = lo >= rlo && hi <= rlo + c->reg_bytes;
  }
  if (ensure_workspace(c, x_node != nullptr && !direct)) return -1;
  if (c->xpatch_cap < nf) {
    cudaFree(c->d_xpatch);
    c->d_xpatch = nullptr;
    CNLD_CK(cudaMalloc(&c->d_xpatch, sizeof(cplx) * (size_t)nf * P * P));
    c->xpatch_cap = nf;
  }
  for (int s = 0; s < c->nslots; s++) c->slots[s].pending = -1;
  // whole-sweep device time: begin event gates every slot stream, end event joins them
  if (!c->ev_begin) { CNLD_CK(cudaEventCreate(&c->ev_begin)); CNLD_CK(cudaEventCreate(&c->ev_end)); }
  CNLD_CK(cudaEventRecord(c->ev_begin, c->slots[0].st));
  for (int s = 1; s < c->nslots; s++) CNLD_CK(cudaStreamWaitEvent(c->slots[s].st, c->ev_begin, 0));
  c->kstat[0] = c->kstat[1] = c->kstat[2] = 0.0;
  double acc[3] = {0, 0, 0};
  unsigned long long l0 = g_launches.load();
  c->sym_redone = 0;
  c->sym_max_ratio = 0.0;
  bool end_recorded = false;
  auto drain = [&](Slot &S) -> int {
    if (S.pending < 0) return 0;
    CNLD_CK(cudaEventSynchronize(S.ev[3]));
    if (S.was_sym && c->refine_steps > 0) {
      // convergence of x <- x + e: |e_last| / |x| estimates rho^steps, the remaining error rho^(steps+1)
      double nr[2] = {0, 0};
      CNLD_CK(cudaMemcpy(nr, S.norms, sizeof(nr), cudaMemcpyDeviceToHost));
      double ratio = (nr[1] > 0.0) ? std::sqrt(nr[0] / nr[1]) : 0.0;
      if (!(ratio == ratio)) ratio = 1.0;  // NaN
      if (ratio > c->sym_max_ratio) c->sym_max_ratio = ratio;
      double remaining = std::pow(ratio, (c->refine_steps + 1.0) / c->refine_steps);
      if (remaining > c->sym_tol) {  // not converged: redo this frequency with the general elimination
        c->sym_redone++;
        if (enqueue_frequency(c, S, freqs[S.pending], cs, rho, c->d_xpatch + (size_t)S.pending * P * P, false)) return -1;
        if (x_node) CNLD_CK(cudaMemcpyAsync(direct ? (void *)(reinterpret_cast<cplx *>(x_node) + (size_t)S.pending * P * nv) : (void *)S.hX,
                                            S.X, sizeof(cplx) * P * nv, cudaMemcpyDeviceToHost, S.st));
        if (x_node) CNLD_CK(cudaEventRecord(S.ev[3], S.st));
        if (end_recorded) CNLD_CK(cudaEventRecord(c->ev_end, S.st));
        CNLD_CK(cudaEventSynchronize(S.ev[3]));
      }
    }
    float ms[3];
    for (int e = 0; e < 3; e++) { CNLD_CK(cudaEventElapsedTime(&ms[e], S.ev[e], S.ev[e + 1])); acc[e] += ms[e] * 1e-3; }
    if (t_freq) t_freq[S.pending] = (ms[0] + ms[1] + ms[2]) * 1e-3;
    if (x_node && !direct) memcpy(reinterpret_cast<cplx *>(x_node) + (size_t)S.pending * P * nv, S.hX, sizeof(cplx) * P * nv);
    if (c->profile) {
      double sec, fl, nl;
      if (lu_trailing_stats(S.lu, &sec, &fl, &nl)) return -1;
      c->kstat[0] += fl; c->kstat[1] += sec; c->kstat[2] += nl;
    }
    uint info = 0;
    CNLD_CK(cudaMemcpy(&info, S.lu.info, sizeof(uint), cudaMemcpyDeviceToHost));
    if (info) { set_error("failed to calculate LU decomposition (pivot %u, frequency index %ld)", info - 1, S.pending); return -2; }
    S.pending = -1;
    return 0;
  };
  for (uint i = 0; i < nf; i++) {
    Slot &S = c->slots[i % c->nslots];
    int rc = drain(S);
    if (rc) return rc;
    if (enqueue_frequency(c, S, freqs[i], cs, rho, c->d_xpatch + (size_t)i * P * P)) return -1;
    if (x_node) CNLD_CK(cudaMemcpyAsync(direct ? (void *)(reinterpret_cast<cplx *>(x_node) + (size_t)i * P * nv) : (void *)S.hX,
                                        S.X, sizeof(cplx) * P * nv, cudaMemcpyDeviceToHost, S.st));
    if (x_node) CNLD_CK(cudaEventRecord(S.ev[3], S.st));
    S.pending = i;
  }
  for (int s = 1; s < c->nslots; s++)
    if (c->slots[s].pending >= 0) CNLD_CK(cudaStreamWaitEvent(c->slots[0].st, c->slots[s].ev[3], 0));
  CNLD_CK(cudaEventRecord(c->ev_end, c->slots[0].st));
  end_recorded = true;
  for (int s = 0; s < c->nslots; s++) {
    int rc = drain(c->slots[(nf + s) % c->nslots]);
    if (rc) return rc;
  }
  CNLD_CK(cudaEventSynchronize(c->ev_end));
  {
    float ms = 0.f;
    CNLD_CK(cudaEventElapsedTime(&ms, c->ev_begin, c->ev_end));
    c->stage[4] = ms * 1e-3;
  }
  if (!device_only && x_patch)
    CNLD_CK(cudaMemcpy(x_patch, c->d_xpatch, sizeof(cplx) * (size_t)nf * P * P, cudaMemcpyDeviceToHost));
  c->stage[0] = acc[0]; c->stage[1] = acc[1]; c->stage[2] = acc[2];
  c->stage[3] = (double)(g_launches.load() - l0);
  return 0;
}
}  // namespace cnld_capi
using namespace cnld_capi;

extern "C" {

const char *cnld_b200_last_error(void) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  t_err = g_err;
  return t_err.c_str();
}
const char *cnld_b200_version(void) { return "cnld_b200 0.1 (sm_100a)"; }
int cnld_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
unsigned long long cnld_b200_launch_count(void) { return g_launches.load(); }
unsigned long long cnld_b200_dmma_flop_count(void) { return g_dmma_flops.load(); }

cnld_ctx *cnld_b200_create(int device, cnld_uint nv, cnld_uint nt, const double *x, const cnld_uint *t,
                           cnld_uint q_reg, cnld_uint q_sing) {
  if (cnld_b200_device_count() <= device) {
    set_error("no CUDA device %d (libcnld_b200 has no CPU fallback)", device);
    return nullptr;
  }
  for (size_t i = 0; i < 3 * (size_t)nt; i++)
    if (t[i] >= nv) { set_error("triangle vertex index out of range"); return nullptr; }
  cnld_ctx *c = new cnld_ctx();
  c->h_x.assign(x, x + 3 * (size_t)nv);
  c->h_t.assign(t, t + 3 * (size_t)nt);
  if (bem_create(c->bem, device, nv, nt, x, t, q_reg, q_sing)) { delete c; return nullptr; }
  return c;
}

void cnld_b200_destroy(cnld_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->bem.device);
  free_workspace(c);
  dev_free(c->d_indptr); dev_free(c->d_indices); dev_free(c->d_K); dev_free(c->d_M); dev_free(c->d_B);
  dev_free(c->d_f_indptr); dev_free(c->d_a_indptr); dev_free(c->d_f_idx); dev_free(c->d_a_idx);
  dev_free(c->d_f_val); dev_free(c->d_a_val); dev_free(c->d_ob); cudaFree(c->d_xpatch); cudaFree(c->d_xnode);
  cudaFree(c->d_sp);
  if (c->pplan_ready) pres_plan_free(c->pplan);
  dev_free(c->d_m_rowptr); dev_free(c->d_m_col); dev_free(c->d_m_dK); dev_free(c->d_m_dM); dev_free(c->d_m_dB);
  if (c->ev_begin) { cudaEventDestroy(c->ev_begin); cudaEventDestroy(c->ev_end); }
  if (c->reg_ptr) cudaHostUnregister(c->reg_ptr);
  bem_free(c->bem);
  delete c;
}

int cnld_b200_pin_host(cnld_ctx *c, void *ptr, size_t bytes) {
  if (!c) { set_error("null context"); return -1; }
  CNLD_CK(cudaSetDevice(c->bem.device));
  if (c->reg_ptr) { cudaHostUnregister(c->reg_ptr); c->reg_ptr = nullptr; c->reg_bytes = 0; }
  if (!ptr || !bytes) return 0;
  if (cudaHostRegister(ptr, bytes, cudaHostRegisterDefault) != cudaSuccess) {
    cudaGetLastError();
    set_error("cudaHostRegister failed for %zu bytes (the sweep falls back to pinned staging)", bytes);
    return -1;
  }
  c->reg_ptr = ptr;
  c->reg_bytes = bytes;
  return 0;
}

int cnld_b200_set_streams(cnld_ctx *c, int n) {
  if (!c || n < 1 || n > MAX_SLOTS) { set_error("nstreams must be 1..%d", MAX_SLOTS); return -1; }
  if (n != c->nslots) { cudaSetDevice(c->bem.device); free_workspace(c); c->nslots = n; }
  return 0;
}

int cnld_b200_get_geometry(cnld_ctx *c, double *g, double *n) {
  if (!c) { set_error("null context"); return -1; }
  CNLD_CK(cudaSetDevice(c->bem.device));
  if (g) CNLD_CK(cudaMemcpy(g, c->bem.g, sizeof(double) * c->bem.nt, cudaMemcpyDeviceToHost));
  if (n) CNLD_CK(cudaMemcpy(n, c->bem.n, sizeof(double) * 3 * c->bem.nt, cudaMemcpyDeviceToHost));
  return 0;
}

int cnld_b200_assemble_slp_ll(cnld_ctx *c, double k_re, double k_im, cnld_field *Z, cnld_uint ld) {
  if (!c) { set_error("null context"); return -1; }
  const size_t nv = c->bem.nv;
  if (ld < nv) { set_error("ld < rows"); return -1; }
  CNLD_CK(cudaSetDevice(c->bem.device));
  cplx *dZ = nullptr;
  CNLD_CK(cudaMalloc(&dZ, sizeof(cplx) * nv * nv));
  int rc = 0;
  do {
    if (cudaMemsetAsync(dZ, 0, sizeof(cplx) * nv * nv, 0) != cudaSuccess) { set_error("memset failed"); rc = -1; break; }
    if (bem_assemble(0, c->bem, k_re, k_im, make_double2(1.0, 0.0), dZ, nv)) { rc = -1; break; }
    cudaError_t e = cudaMemcpy2D(Z, sizeof(cplx) * ld, dZ, sizeof(cplx) * nv, sizeof(cplx) * nv, nv,
                                 cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("copy back: %s", cudaGetErrorString(e)); rc = -1; }
  } while (0);
  cudaFree(dZ);
  return rc;
}

int cnld_b200_nearfield_slp_ll(cnld_ctx *c, double k_re, double k_im, const cnld_uint *ridx, cnld_uint rows,
                               const cnld_uint *cidx, cnld_uint cols, cnld_field *N, cnld_uint ld) {
  if (!c) { set_error("null context"); return -1; }
  if (ld < rows) { set_error("ld < rows"); return -1; }
  if (rows == 0 || cols == 0) return 0;
  CNLD_CK(cudaSetDevice(c->bem.device));
  cplx *d = nullptr;
  CNLD_CK(cudaMalloc(&d, sizeof(cplx) * (size_t)rows * cols));
  int rc = nearfield_block(0, c->bem, c->h_t.data(), ridx, rows, cidx, cols, k_re, k_im, d);
  if (!rc) {
    cudaError_t e = cudaMemcpy2D(N, sizeof(cplx) * ld, d, sizeof(cplx) * rows, sizeof(cplx) * rows, cols, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("nearfield: %s", cudaGetErrorString(e)); rc = -1; }
  }
  cudaFree(d);
  return rc;
}

static int lrdecomp_impl(int device, cnld_uint n, cnld_field *a, cnld_uint ld, cnld_uint *info, bool sym) {
  if (cnld_b200_device_count() <= device) { set_error("no CUDA device %d (no CPU fallback)", device); return -1; }
  CNLD_CK(cudaSetDevice(device));
  cplx *dA = nullptr;
  CNLD_CK(cudaMalloc(&dA, sizeof(cplx) * (size_t)n * n));
  LuWork w;
  int rc = 0;
  do {
    if (cudaMemcpy2D(dA, sizeof(cplx) * n, a, sizeof(cplx) * ld, sizeof(cplx) * n, n, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("H2D failed"); rc = -1; break; }
    if (lu_work_alloc(w, n) || (sym ? lu_factor_sym(0, w, dA, n) : lu_factor(0, w, dA, n))) { rc = -1; break; }
    uint inf = 0;
    if (cudaMemcpy(&inf, w.info, sizeof(uint), cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("D2H failed"); rc = -1; break; }
    if (info) *info = inf;
    if (cudaMemcpy2D(a, sizeof(cplx) * ld, dA, sizeof(cplx) * n, sizeof(cplx) * n, n, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("D2H failed"); rc = -1; }
  } while (0);
  lu_work_free(w);
  cudaFree(dA);
  return rc;
}
int cnld_b200_lrdecomp(int device, cnld_uint n, cnld_field *a, cnld_uint ld, cnld_uint *info) {
  return lrdecomp_impl(device, n, a, ld, info, false);
}
int cnld_b200_lrdecomp_sym(int device, cnld_uint n, cnld_field *a, cnld_uint ld, cnld_uint *info) {
  return lrdecomp_impl(device, n, a, ld, info, true);
}

int cnld_b200_lrsolve(int device, cnld_uint n, const cnld_field *lu, cnld_uint ld, cnld_field *x,
                      cnld_uint nrhs, cnld_uint ldx) {
  if (cnld_b200_device_count() <= device) { set_error("no CUDA device %d (no CPU fallback)", device); return -1; }
  CNLD_CK(cudaSetDevice(device));
  cplx *dA = nullptr, *dX = nullptr;
  CNLD_CK(cudaMalloc(&dA, sizeof(cplx) * (size_t)n * n));
  CNLD_CK(cudaMalloc(&dX, sizeof(cplx) * (size_t)n * nrhs));
  LuWork w;
  int rc = 0;
  do {
    if (cudaMemcpy2D(dA, sizeof(cplx) * n, lu, sizeof(cplx) * ld, sizeof(cplx) * n, n, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy2D(dX, sizeof(cplx) * n, x, sizeof(cplx) * ldx, sizeof(cplx) * n, nrhs, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("H2D failed"); rc = -1; break; }
    if (lu_work_alloc(w, n) || lu_invert_diag(0, w, dA, n) || lu_solve(0, w, dA, n, dX, nrhs, n)) { rc = -1; break; }
    if (cudaMemcpy2D(x, sizeof(cplx) * ldx, dX, sizeof(cplx) * n, sizeof(cplx) * n, nrhs, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("D2H failed"); rc = -1; }
  } while (0);
  lu_work_free(w);
  cudaFree(dA);
  cudaFree(dX);
  return rc;
}

int cnld_b200_zgemm(int device, cnld_uint m, cnld_uint n, cnld_uint k, double alpha, const cnld_field *A,
                    cnld_uint lda, const cnld_field *B, cnld_uint ldb, double beta, cnld_field *C,
                    cnld_uint ldc) {
  if (cnld_b200_device_count() <= device) { set_error("no CUDA device %d (no CPU fallback)", device); return -1; }
  if (!((alpha == 1.0 || alpha == -1.0) && (beta == 0.0 || beta == 1.0))) { set_error("alpha must be +-1, beta 0 or 1"); return -1; }
  CNLD_CK(cudaSetDevice(device));
  cplx *dA = nullptr, *dB = nullptr, *dC = nullptr;
  CNLD_CK(cudaMalloc(&dA, sizeof(cplx) * (size_t)m * k + 16));
  CNLD_CK(cudaMalloc(&dB, sizeof(cplx) * (size_t)k * n + 16));
  CNLD_CK(cudaMalloc(&dC, sizeof(cplx) * (size_t)m * n + 16));
  int rc = 0;
  do {
    if (cudaMemcpy2D(dA, sizeof(cplx) * m, A, sizeof(cplx) * lda, sizeof(cplx) * m, k, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy2D(dB, sizeof(cplx) * k, B, sizeof(cplx) * ldb, sizeof(cplx) * k, n, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy2D(dC, sizeof(cplx) * m, C, sizeof(cplx) * ldc, sizeof(cplx) * m, n, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("H2D failed"); rc = -1; break; }
    if (zgemm(0, m, n, k, alpha, dA, m, dB, k, beta, dC, m)) { rc = -1; break; }
    cudaError_t e = cudaMemcpy2D(C, sizeof(cplx) * ldc, dC, sizeof(cplx) * m, sizeof(cplx) * m, n, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("zgemm: %s", cudaGetErrorString(e)); rc = -1; }
  } while (0);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  return rc;
}

int cnld_b200_set_mbk(cnld_ctx *c, const int64_t *indptr, const cnld_uint *indices, const double *K,
                      const double *M, const double *B) {
  if (!c) { set_error("null context"); return -1; }
  CNLD_CK(cudaSetDevice(c->bem.device));
  const size_t nv = c->bem.nv;
  size_t nnz = (size_t)indptr[nv];
  for (size_t i = 0; i < nnz; i++) if (indices[i] >= nv) { set_error("MBK column index out of range"); return -1; }
  c->nnz = nnz;
  c->h_indptr.assign(indptr, indptr + nv + 1);
  c->h_indices.assign(indices, indices + nnz);
  c->h_K.assign(K, K + nnz); c->h_M.assign(M, M + nnz); c->h_B.assign(B, B + nnz);
  if (upload(c->d_indptr, indptr, nv + 1) || upload(c->d_indices, indices, nnz) || upload(c->d_K, K, nnz) ||
      upload(c->d_M, M, nnz) || upload(c->d_B, B, nnz)) return -1;
  // strictly upper asymmetry (symmetric path): Delta[r, c] = A[r, c] - A[c, r] for r < c, where the
  // symmetric part S is defined by the LOWER triangle; an entry missing from the pattern counts as 0
  {
    struct Trip { uint r, c; double k, m, b; };
    std::vector<Trip> trips;
    bool sorted = true;   // scipy's canonical CSR: binary search; otherwise scan the row
    for (size_t r = 0; r < nv && sorted; r++)
      for (int64_t p = indptr[r] + 1; p < indptr[r + 1]; p++) if (indices[p - 1] >= indices[p]) { sorted = false; break; }
    auto find = [&](size_t r, uint col) -> int64_t {
      if (sorted) {
        const cnld_uint *b = indices + indptr[r], *e = indices + indptr[r + 1];
        const cnld_uint *it = std::lower_bound(b, e, col);
        return (it != e && *it == col) ? (int64_t)(it - indices) : -1;
      }
      for (int64_t p = indptr[r]; p < indptr[r + 1]; p++) if (indices[p] == col) return p;
      return -1;
    };
    for (size_t r = 0; r < nv; r++)
      for (int64_t p = indptr[r]; p < indptr[r + 1]; p++) {
        const uint col = indices[p];
        if (col > r) {  // upper entry: compare with its lower partner
          int64_t q = find(col, (uint)r);
          double a = K[p] - (q >= 0 ? K[q] : 0.0), b = M[p] - (q >= 0 ? M[q] : 0.0), d = B[p] - (q >= 0 ? B[q] : 0.0);
          if (a != 0.0 || b != 0.0 || d != 0.0) trips.push_back({(uint)r, col, a, b, d});
        } else if (col < r && find(col, (uint)r) < 0) {  // lower entry without an upper partner
          trips.push_back({col, (uint)r, -K[p], -M[p], -B[p]});
        }
      }
    std::stable_sort(trips.begin(), trips.end(), [](const Trip &x, const Trip &y) { return x.r < y.r; });
    std::vector<uint> rp(nv + 1, 0), mc;
    std::vector<double> dK, dM, dB;
    for (const Trip &t : trips) { rp[t.r + 1]++; mc.push_back(t.c); dK.push_back(t.k); dM.push_back(t.m); dB.push_back(t.b); }
    for (size_t r = 0; r < nv; r++) rp[r + 1] += rp[r];
    c->m_nnz = (uint)mc.size();
    if (!c->m_nnz) {
      dev_free(c->d_m_rowptr); dev_free(c->d_m_col); dev_free(c->d_m_dK); dev_free(c->d_m_dM); dev_free(c->d_m_dB);
    } else {
      if (upload(c->d_m_rowptr, rp.data(), nv + 1) || upload(c->d_m_col, mc.data(), mc.size()) ||
          upload(c->d_m_dK, dK.data(), dK.size()) || upload(c->d_m_dM, dM.data(), dM.size()) ||
          upload(c->d_m_dB, dB.data(), dB.size())) return -1;
    }
  }
  return 0;
}

int cnld_b200_set_symmetric(cnld_ctx *c, int mode, int refine_steps) {
  if (!c || mode < 0 || mode > 1 || refine_steps < 0 || refine_steps > 16) { set_error("set_symmetric: mode 0/1, refine_steps 0..16"); return -1; }
  if (mode != c->symmetric) { cudaSetDevice(c->bem.device); free_workspace(c); }
  c->symmetric = mode;
  c->refine_steps = refine_steps;
  return 0;
}
int cnld_b200_symmetric_stats(const cnld_ctx *c, unsigned *mbk_asym_nnz, unsigned *redone, double *max_ratio) {
  if (!c) return -1;
  if (mbk_asym_nnz) *mbk_asym_nnz = c->m_nnz;
  if (redone) *redone = c->sym_redone;
  if (max_ratio) *max_ratio = c->sym_max_ratio;
  return c->symmetric;
}
// Host-only: the work / copy lists the symmetric-path assembly would use for this mesh (no device needed).
int cnld_b200_translation_plan(cnld_uint nv, cnld_uint nt, const double *x, const cnld_uint *t, int allow,
                               cnld_uint *counts, cnld_uint *work, cnld_uint work_cap, cnld_uint *copy,
                               cnld_uint copy_cap) {
  for (size_t i = 0; i < 3 * (size_t)nt; i++)
    if (t[i] >= nv) { set_error("triangle vertex index out of range"); return -1; }
  TranslationPlan P;
  build_translation_plan(nv, nt, x, t, 128, 64, allow != 0, P);
  if (counts) {
    counts[0] = P.periodic ? 1 : 0; counts[1] = P.ncomp; counts[2] = P.nclass;
    counts[3] = (cnld_uint)P.work.size(); counts[4] = (cnld_uint)P.copy.size(); counts[5] = P.nvc;
  }
  if (work) for (size_t i = 0; i < P.work.size() && i < work_cap; i++) memcpy(work + 4 * i, &P.work[i], 16);
  if (copy) for (size_t i = 0; i < P.copy.size() && i < copy_cap; i++) memcpy(copy + 5 * i, &P.copy[i], 20);
  return 0;
}

int cnld_b200_set_translation_classes(cnld_ctx *c, int on) {
  if (!c) { set_error("null context"); return -1; }
  c->bem.use_periodic = on ? 1 : 0;
  return 0;
}
int cnld_b200_translation_info(const cnld_ctx *c, unsigned *ncomp, unsigned *nclass) {
  if (!c) return -1;
  if (ncomp) *ncomp = c->bem.ncomp;
  if (nclass) *nclass = c->bem.nclass;
  return (c->bem.periodic && c->bem.use_periodic) ? 1 : 0;
}
int cnld_b200_set_symmetric_tol(cnld_ctx *c, double tol) {
  if (!c || !(tol > 0.0)) { set_error("set_symmetric_tol: tol > 0"); return -1; }
  c->sym_tol = tol;
  return 0;
}

int cnld_b200_set_loads(cnld_ctx *c, cnld_uint P, const int64_t *f_indptr, const cnld_uint *f_idx,
                        const double *f_val, const int64_t *a_indptr, const cnld_uint *a_idx,
                        const double *a_val, const uint8_t *ob) {
  if (!c) { set_error("null context"); return -1; }
  CNLD_CK(cudaSetDevice(c->bem.device));
  const size_t nv = c->bem.nv;
  for (int64_t i = 0; i < f_indptr[P]; i++) if (f_idx[i] >= nv) { set_error("F row index out of range"); return -1; }
  for (int64_t i = 0; i < a_indptr[P]; i++) if (a_idx[i] >= nv) { set_error("AVG row index out of range"); return -1; }
  if (P != c->P) { free_workspace(c); c->xpatch_cap = 0; }
  c->P = P;
  // first loaded row of every source patch: lets the forward substitution skip columns that are still zero
  c->f_first_row.assign(P, (uint)nv);
  for (cnld_uint s = 0; s < P; s++)
    for (int64_t q = f_indptr[s]; q < f_indptr[s + 1]; q++) c->f_first_row[s] = std::min<uint>(c->f_first_row[s], f_idx[q]);
  // non-decreasing lower envelope (suffix minimum): column s and every later one are zero above f_first_row[s]
  for (cnld_uint s = P; s-- > 1;) c->f_first_row[s - 1] = std::min(c->f_first_row[s - 1], c->f_first_row[s]);
  if (upload(c->d_f_indptr, f_indptr, (size_t)P + 1) || upload(c->d_f_idx, f_idx, (size_t)f_indptr[P]) ||
      upload(c->d_f_val, f_val, (size_t)f_indptr[P]) || upload(c->d_a_indptr, a_indptr, (size_t)P + 1) ||
      upload(c->d_a_idx, a_idx, (size_t)a_indptr[P]) || upload(c->d_a_val, a_val, (size_t)a_indptr[P]) ||
      upload(c->d_ob, ob, nv)) return -1;
  return 0;
}

int cnld_b200_sweep_run(cnld_ctx *c, const double *freqs, cnld_uint nf, double cs, double rho,
                        cnld_field *x_patch, cnld_field *x_node, double *t_freq) {
  return sweep_impl(c, freqs, nf, cs, rho, x_patch, x_node, t_freq, false);
}

int cnld_b200_sweep_run_device(cnld_ctx *c, const double *freqs, cnld_uint nf, double cs, double rho) {
  return sweep_impl(c, freqs, nf, cs, rho, nullptr, nullptr, nullptr, true);
}

int cnld_b200_sweep_fetch(cnld_ctx *c, cnld_uint nf, cnld_field *x_patch, cnld_field *x_node) {
  if (!c || !c->d_xpatch || nf > c->xpatch_cap) { set_error("no sweep results to fetch"); return -1; }
  if (x_node) { set_error("x_node is not retained by sweep_run_device; use cnld_b200_sweep_run"); return -1; }
  CNLD_CK(cudaSetDevice(c->bem.device));
  CNLD_CK(cudaMemcpy(x_patch, c->d_xpatch, sizeof(cplx) * (size_t)nf * c->P * c->P, cudaMemcpyDeviceToHost));
  return 0;
}

int cnld_b200_sweep_stage_times(cnld_ctx *c, double *out) {
  if (!c) { set_error("null context"); return -1; }
  for (int i = 0; i < 5; i++) out[i] = c->stage[i];
  return 0;
}

int cnld_b200_set_profile(cnld_ctx *c, int on) {
  if (!c) { set_error("null context"); return -1; }
  c->profile = on != 0;
  for (int s = 0; s < c->nslots; s++)
    if (c->slots[s].lu.linv) lu_profile(c->slots[s].lu, c->profile);
  return 0;
}

int cnld_b200_sweep_kernel_stats(cnld_ctx *c, double *out) {
  if (!c) { set_error("null context"); return -1; }
  for (int i = 0; i < 3; i++) out[i] = c->kstat[i];
  return 0;
}

int cnld_b200_pres_vector(cnld_ctx *c, const double *obs, cnld_uint nobs, double k, double cs, double rho,
                          cnld_field *pvec) {
  if (!c) { set_error("null context"); return -1; }
  CNLD_CK(cudaSetDevice(c->bem.device));
  double *d_obs = nullptr;
  cplx *d_p = nullptr;
  if (cudaMalloc(&d_obs, sizeof(double) * 3 * (size_t)nobs) != cudaSuccess ||
      cudaMalloc(&d_p, sizeof(cplx) * (size_t)nobs * c->bem.nv) != cudaSuccess) {
    set_error("pres_vector: out of device memory (%u points)", nobs);
    cudaGetLastError();
    cudaFree(d_obs); cudaFree(d_p);
    return -1;
  }
  int rc = 0;
  do {
    if (cudaMemcpy(d_obs, obs, sizeof(double) * 3 * (size_t)nobs, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("H2D failed"); rc = -1; break; }
    if (pres_vector(0, c->bem, d_obs, nobs, k, cs, rho, d_p)) { rc = -1; break; }
    cudaError_t e = cudaMemcpy(pvec, d_p, sizeof(cplx) * (size_t)nobs * c->bem.nv, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("pres_vector: %s", cudaGetErrorString(e)); rc = -1; }
  } while (0);
  cudaFree(d_obs);
  cudaFree(d_p);
  return rc;
}

int cnld_b200_pressure(cnld_ctx *c, const double *obs, cnld_uint nobs, const double *ks, cnld_uint nk,
                       double cs, double rho, const cnld_field *disp, cnld_uint nsrc, cnld_field *p) {
  if (!c) { set_error("null context"); return -1; }
  CNLD_CK(cudaSetDevice(c->bem.device));
  const size_t nv = c->bem.nv, nt = c->bem.nt;
  if (nsrc > 4) {
    // many sources: evaluate the kernel once per point pair, contract in node space on the DMMA ZGEMM
    cnld_pfield *pf = cnld_b200_pfield_create(c, obs, nobs, nsrc);
    if (!pf) return -1;
    int rc = 0;
    for (uint ik = 0; ik < nk && !rc; ik++)
      rc = cnld_b200_pfield_run(pf, ks[ik], cs, rho, disp + (size_t)ik * nsrc * nv, nsrc, p + (size_t)ik * nsrc * nobs);
    cnld_b200_pfield_destroy(pf);
    return rc;
  }
  double *d_obs = nullptr;
  cplx *d_disp = nullptr, *d_D = nullptr, *d_out = nullptr;
  if (!c->d_sp) {
    CNLD_CK(cudaMalloc(&c->d_sp, sizeof(double) * 2 * 3 * nt));
    if (pres_points(0, c->bem, c->d_sp)) return -1;
  }
  if (cudaMalloc(&d_obs, sizeof(double) * 3 * (size_t)nobs) != cudaSuccess || cudaMalloc(&d_disp, sizeof(cplx) * nsrc * nv) != cudaSuccess ||
      cudaMalloc(&d_D, sizeof(cplx) * nsrc * 3 * nt) != cudaSuccess || cudaMalloc(&d_out, sizeof(cplx) * (size_t)nsrc * nobs) != cudaSuccess) {
    set_error("pressure: out of device memory (%u points, %u sources)", nobs, nsrc);
    cudaGetLastError();
    cudaFree(d_obs); cudaFree(d_disp); cudaFree(d_D); cudaFree(d_out);
    return -1;
  }
  int rc = 0;
  do {
    if (cudaMemcpy(d_obs, obs, sizeof(double) * 3 * (size_t)nobs, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("H2D failed"); rc = -1; break; }
    for (uint ik = 0; ik < nk && !rc; ik++) {
      const cplx *hd = reinterpret_cast<const cplx *>(disp) + (size_t)ik * nsrc * nv;
      if (cudaMemcpyAsync(d_disp, hd, sizeof(cplx) * nsrc * nv, cudaMemcpyHostToDevice, 0) != cudaSuccess) { set_error("H2D failed"); rc = -1; break; }
      if (pressure_field(0, c->bem, d_obs, nobs, c->d_sp, ks[ik], cs, rho, d_disp, nsrc, d_D, d_out)) { rc = -1; break; }
      cudaError_t e = cudaMemcpy(reinterpret_cast<cplx *>(p) + (size_t)ik * nsrc * nobs, d_out,
                                 sizeof(cplx) * (size_t)nsrc * nobs, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { set_error("pressure: %s", cudaGetErrorString(e)); rc = -1; }
    }
  } while (0);
  cudaFree(d_obs); cudaFree(d_disp); cudaFree(d_D); cudaFree(d_out);
  return rc;
}

}  // extern "C"
