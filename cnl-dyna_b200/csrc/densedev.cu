// Device-resident dense matrices behind the H2Lib-named (host pointer) entry points.
//
// H2Lib's API hands every call a HOST matrix (cnld/h2lib/amatrix.pyx aliases it with numpy), and the
// reference calls lrsolve_amatrix_avector once per source patch with the same LU
// (cnld/scripts/generate_db.py:82-91).  Re-uploading 3.5 GB per call would make the PCIe bus the
// solver, so the device copy made by the first call (lrdecomp keeps the factors it just computed) is
// kept and found again by host pointer; a fingerprint of sampled entries guards against the host
// array having been rewritten in between.  Also here: the dense kernels the compat layer needs beyond
// the sweep's (matrix-vector products with full / triangular parts and their adjoints, triangular
// solves, Cholesky), so that none of the H2Lib-named entry points does its arithmetic on the host.
#include <algorithm>
#include <cstdlib>
#include <list>
#include <mutex>

#include "densedev.cuh"

namespace cnld {
namespace dd {
namespace {
std::mutex g_mu;
std::list<Resident *> g_cache;   // most recently used first
constexpr size_t MAX_RESIDENT = 3;
constexpr uint NSAMPLE = 4096;

// Fingerprint = the whole diagonal + NSAMPLE entries spread evenly over the matrix.  It catches the host
// array being reused for another matrix or rescaled / refactored in place; a write that touches none of the
// sampled entries is NOT detected (cnl-dyna never writes into a matrix between its lrsolve calls;
// CNLD_B200_NO_RESIDENT=1 turns the cache off and every call uploads again).
size_t sample_index(uint s, uint rows, uint cols, size_t ldh) {
  const uint nd = std::min(rows, cols);
  if (s < nd) return (size_t)s * ldh + s;
  s -= nd;
  const unsigned long long tot = (unsigned long long)rows * cols;
  const unsigned long long e = (tot * (2ull * s + 1)) / (2ull * NSAMPLE) % (tot ? tot : 1);
  return (size_t)(e / rows) * ldh + (size_t)(e % rows);
}
uint sample_count(const Resident &R) { return NSAMPLE + std::min(R.rows, R.cols); }
void take_fingerprint(Resident &R, const cplx *host) {
  R.finger.resize(sample_count(R));
  for (uint s = 0; s < R.finger.size(); s++) R.finger[s] = host[sample_index(s, R.rows, R.cols, R.ldh)];
}
bool fingerprint_ok(const Resident &R, const cplx *host) {
  static const bool off = getenv("CNLD_B200_NO_RESIDENT") != nullptr;
  if (off) return false;
  for (uint s = 0; s < R.finger.size(); s++) {
    const cplx v = host[sample_index(s, R.rows, R.cols, R.ldh)];
    if (memcmp(&v, &R.finger[s], sizeof(cplx)) != 0) return false;
  }
  return true;
}
void destroy(Resident *R) {
  if (R->w_ready) lu_work_free(R->w);
  cudaFree(R->d);
  delete R;
}

// ---- kernels ------------------------------------------------------------------------------------
// y += alpha * T(A) x, thread per row (column-major: consecutive threads read consecutive rows)
__global__ void k_gemv_n(uint rows, uint cols, const cplx *__restrict__ A, size_t ld, int tri, cplx alpha,
                         const cplx *__restrict__ x, cplx *y) {
  const uint i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  uint j0 = 0, j1 = cols;
  if (tri == TRI_UPPER) j0 = i;
  else if (tri == TRI_UNIT_LOWER) j1 = min(i, cols);
  else if (tri == TRI_LOWER) j1 = min(i + 1, cols);
  cplx s = make_double2(0.0, 0.0);
  for (uint j = j0; j < j1; j++) cfma(s, A[(size_t)j * ld + i], x[j]);
  if (tri == TRI_UNIT_LOWER && i < cols) { s.x += x[i].x; s.y += x[i].y; }
  const cplx v = cmul(alpha, s);
  y[i].x += v.x; y[i].y += v.y;
}
// y += alpha * T(A)^H x, warp per column (a column is contiguous)
__global__ void k_gemv_c(uint rows, uint cols, const cplx *__restrict__ A, size_t ld, int tri, cplx alpha,
                         const cplx *__restrict__ x, cplx *y) {
  const uint j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (j >= cols) return;
  uint i0 = 0, i1 = rows;
  if (tri == TRI_UPPER) i1 = min(j + 1, rows);
  else if (tri == TRI_UNIT_LOWER) i0 = j + 1;
  else if (tri == TRI_LOWER) i0 = j;
  cplx s = make_double2(0.0, 0.0);
  const cplx *col = A + (size_t)j * ld;
  for (uint i = i0 + lane; i < i1; i += 32) {
    const cplx a = col[i], b = x[i];
    s.x += a.x * b.x + a.y * b.y;
    s.y += a.x * b.y - a.y * b.x;
  }
  for (int off = 16; off > 0; off >>= 1) { s.x += __shfl_xor_sync(0xffffffffu, s.x, off); s.y += __shfl_xor_sync(0xffffffffu, s.y, off); }
  if (lane == 0) {
    if (tri == TRI_UNIT_LOWER && j < rows) { s.x += x[j].x; s.y += x[j].y; }
    const cplx v = cmul(alpha, s);
    y[j].x += v.x; y[j].y += v.y;
  }
}

// In-place triangular solve with one CTA.  adjoint = false: column sweep (x_i -= a_ij x_j down / up a
// column); adjoint = true: op(A) = A^H, whose rows are A's columns, so x_j needs a dot product with
// column j.  One barrier per unknown; meant for the single-vector compat calls, not the sweep.
__global__ void __launch_bounds__(1024)
k_trsv(uint n, const cplx *__restrict__ A, size_t ld, int lower, int unit, int adjoint, cplx *x) {
  __shared__ double sr[32], si[32];
  __shared__ cplx piv;
  const uint tid = threadIdx.x, nt = blockDim.x;
  const bool forward = adjoint ? !lower : (bool)lower;   // op(A) is lower triangular
  for (uint step = 0; step < n; step++) {
    const uint j = forward ? step : n - 1 - step;
    const cplx *col = A + (size_t)j * ld;
    if (!adjoint) {
      if (tid == 0) {
        cplx v = x[j];
        if (!unit) v = cmul(v, cinv(col[j]));
        x[j] = v;
        piv = v;
      }
      __syncthreads();
      const cplx xj = piv;
      if (lower) { for (uint i = j + 1 + tid; i < n; i += nt) cfms(x[i], col[i], xj); }
      else { for (uint i = tid; i < j; i += nt) cfms(x[i], col[i], xj); }
      __syncthreads();
    } else {
      // x_j = (x_j - sum_i conj(a_ij) x_i) / conj(a_jj), i over the already solved unknowns
      double ar = 0.0, ai = 0.0;
      const uint i0 = lower ? j + 1 : 0, i1 = lower ? n : j;
      for (uint i = i0 + tid; i < i1; i += nt) {
        const cplx a = col[i], b = x[i];
        ar += a.x * b.x + a.y * b.y;
        ai += a.x * b.y - a.y * b.x;
      }
      for (int off = 16; off > 0; off >>= 1) { ar += __shfl_xor_sync(0xffffffffu, ar, off); ai += __shfl_xor_sync(0xffffffffu, ai, off); }
      if ((tid & 31) == 0) { sr[tid >> 5] = ar; si[tid >> 5] = ai; }
      __syncthreads();
      if (tid == 0) {
        for (uint w = 1; w < (nt + 31) / 32; w++) { ar += sr[w]; ai += si[w]; }
        cplx v = make_double2(x[j].x - ar, x[j].y - ai);
        if (!unit) { const cplx d = col[j]; v = cmul(v, cinv(make_double2(d.x, -d.y))); }
        x[j] = v;
      }
      __syncthreads();
    }
  }
}

// ---- Cholesky A = L L^H (lower), blocked right-looking, NB = 32 -------------------------------------
constexpr int CH_NB = 32;
__global__ void __launch_bounds__(CH_NB *CH_NB)
k_potf2(uint n, uint k0, cplx *A, size_t ld, uint *info) {
  __shared__ cplx s[CH_NB][CH_NB + 1];
  const uint nb = min((uint)CH_NB, n - k0), i = threadIdx.x % CH_NB, j = threadIdx.x / CH_NB;
  if (i < nb && j < nb) s[i][j] = A[(size_t)(k0 + j) * ld + k0 + i];
  __syncthreads();
  for (uint k = 0; k < nb; k++) {
    if (threadIdx.x == 0) {
      const double d = s[k][k].x;
      if (!(d > 0.0) || !isfinite(d)) { if (*info == 0) *info = k0 + k + 1; s[k][k] = make_double2(1.0, 0.0); }
      else s[k][k] = make_double2(sqrt(d), 0.0);
    }
    __syncthreads();
    if (j == 0 && i > k && i < nb) { const double r = 1.0 / s[k][k].x; s[i][k].x *= r; s[i][k].y *= r; }
    __syncthreads();
    if (i > k && j > k && i >= j && i < nb && j < nb) {   // a_ij -= l_ik conj(l_jk)
      const cplx a = s[i][k], b = s[j][k];
      s[i][j].x -= a.x * b.x + a.y * b.y;
      s[i][j].y -= a.y * b.x - a.x * b.y;
    }
    __syncthreads();
  }
  if (i < nb && j < nb) A[(size_t)(k0 + j) * ld + k0 + i] = i >= j ? s[i][j] : make_double2(0.0, 0.0);
}
// L21 = A21 L11^-H: thread per row, forward substitution against the diagonal block in shared memory
__global__ void __launch_bounds__(128)
k_chol_panel(uint n, uint k0, cplx *A, size_t ld) {
  __shared__ cplx L[CH_NB][CH_NB + 1];
  const uint nb = min((uint)CH_NB, n - k0);
  for (uint e = threadIdx.x; e < CH_NB * CH_NB; e += blockDim.x) {
    const uint i = e % CH_NB, j = e / CH_NB;
    L[i][j] = (i < nb && j < nb) ? A[(size_t)(k0 + j) * ld + k0 + i] : make_double2(0.0, 0.0);
  }
  __syncthreads();
  const uint r = k0 + nb + blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  cplx l[CH_NB];
  for (uint j = 0; j < nb; j++) {
    cplx v = A[(size_t)(k0 + j) * ld + r];
    for (uint p = 0; p < j; p++) {   // v -= l_p conj(L_jp)
      const cplx a = l[p], b = L[j][p];
      v.x -= a.x * b.x + a.y * b.y;
      v.y -= a.y * b.x - a.x * b.y;
    }
    const double d = 1.0 / L[j][j].x;
    l[j] = make_double2(v.x * d, v.y * d);
    A[(size_t)(k0 + j) * ld + r] = l[j];
  }
}
// A22 -= L21 L21^H on the tiles on or below the diagonal (32 x 32 per CTA)
__global__ void __launch_bounds__(256)
k_chol_update(uint n, uint k0, uint nb, cplx *A, size_t ld) {
  if (blockIdx.y > blockIdx.x) return;
  __shared__ cplx P[32][CH_NB + 1], Q[32][CH_NB + 1];
  const uint r0 = k0 + nb + blockIdx.x * 32, c0 = k0 + nb + blockIdx.y * 32;
  for (uint e = threadIdx.x; e < 32 * CH_NB; e += 256) {
    const uint i = e % 32, p = e / 32;
    P[i][p] = (r0 + i < n && p < nb) ? A[(size_t)(k0 + p) * ld + r0 + i] : make_double2(0.0, 0.0);
    Q[i][p] = (c0 + i < n && p < nb) ? A[(size_t)(k0 + p) * ld + c0 + i] : make_double2(0.0, 0.0);
  }
  __syncthreads();
  const uint i = threadIdx.x % 32;
  for (uint jj = threadIdx.x / 32; jj < 32; jj += 8) {
    const uint row = r0 + i, col = c0 + jj;
    if (row >= n || col >= n || row < col) continue;
    double sr = 0.0, si = 0.0;
#pragma unroll 8
    for (uint p = 0; p < CH_NB; p++) {
      const cplx a = P[i][p], b = Q[jj][p];
      sr += a.x * b.x + a.y * b.y;
      si += a.y * b.x - a.x * b.y;
    }
    cplx *d = A + (size_t)col * ld + row;
    d->x -= sr; d->y -= si;
  }
}
__global__ void k_dotc(size_t n, const cplx *__restrict__ x, const cplx *__restrict__ y, double *out) {   // <x, y> = sum conj(x) y
  double sr = 0.0, si = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    sr += x[i].x * y[i].x + x[i].y * y[i].y;
    si += x[i].x * y[i].y - x[i].y * y[i].x;
  }
  for (int off = 16; off > 0; off >>= 1) { sr += __shfl_xor_sync(0xffffffffu, sr, off); si += __shfl_xor_sync(0xffffffffu, si, off); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(out, sr); atomicAdd(out + 1, si); }
}
__global__ void k_scal(size_t n, cplx alpha, cplx *x) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = cmul(alpha, x[i]);
}
__global__ void k_axpy_c(size_t n, cplx alpha, const cplx *__restrict__ x, cplx *y) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cfma(y[i], alpha, x[i]);
}
}  // namespace

int device_id() {
  static int dev = -1;
  if (dev < 0) {
    const char *e = getenv("CNLD_B200_DEVICE");   // one process per GPU: the launcher exports the rank's device
    dev = e ? atoi(e) : 0;
    if (!e) { int cur = 0; if (cudaGetDevice(&cur) == cudaSuccess) dev = cur; }
  }
  return dev;
}

Resident *find(const void *host, uint rows, uint cols, size_t ldh) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto it = g_cache.begin(); it != g_cache.end(); ++it) {
    Resident *R = *it;
    if (R->host != host) continue;
    if (R->rows == rows && R->cols == cols && R->ldh == ldh && fingerprint_ok(*R, static_cast<const cplx *>(host))) {
      g_cache.erase(it);
      g_cache.push_front(R);
      return R;
    }
    g_cache.erase(it);      // same address, different content: stale
    destroy(R);
    return nullptr;
  }
  return nullptr;
}

void drop(const void *host) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto it = g_cache.begin(); it != g_cache.end();)
    if ((*it)->host == host) { destroy(*it); it = g_cache.erase(it); } else ++it;
}

void drop_all() {
  std::lock_guard<std::mutex> lk(g_mu);
  for (Resident *R : g_cache) destroy(R);
  g_cache.clear();
}

// register a device matrix (ownership passes to the cache) as the copy of `host`
Resident *adopt(const void *host, uint rows, uint cols, size_t ldh, cplx *d, int kind) {
  drop(host);
  Resident *R = new Resident();
  R->host = host; R->rows = rows; R->cols = cols; R->ldh = ldh; R->d = d; R->kind = kind;
  take_fingerprint(*R, static_cast<const cplx *>(host));
  std::lock_guard<std::mutex> lk(g_mu);
  g_cache.push_front(R);
  while (g_cache.size() > MAX_RESIDENT) { destroy(g_cache.back()); g_cache.pop_back(); }
  return R;
}

Resident *resident(const void *host, uint rows, uint cols, size_t ldh, int kind) {
  if (Resident *R = find(host, rows, cols, ldh)) return R;
  if (cudaSetDevice(device_id()) != cudaSuccess) { set_error("no CUDA device %d (libcnld_b200 has no CPU fallback)", device_id()); return nullptr; }
  cplx *d = nullptr;
  if (cudaMalloc(&d, sizeof(cplx) * std::max<size_t>((size_t)rows * cols, 1)) != cudaSuccess) {
    cudaGetLastError();
    drop_all();             // make room and retry once
    if (cudaMalloc(&d, sizeof(cplx) * std::max<size_t>((size_t)rows * cols, 1)) != cudaSuccess) {
      set_error("out of device memory for a %u x %u matrix", rows, cols);
      cudaGetLastError();
      return nullptr;
    }
  }
  if (rows && cols &&
      cudaMemcpy2D(d, sizeof(cplx) * rows, host, sizeof(cplx) * ldh, sizeof(cplx) * rows, cols, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("H2D failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
    return nullptr;
  }
  return adopt(host, rows, cols, ldh, d, kind);
}

int gemv(cudaStream_t st, const cplx *A, size_t ld, uint rows, uint cols, int tri, int adjoint, cplx alpha, const cplx *x,
         cplx *y) {
  if (!rows || !cols) return 0;
  if (!adjoint) k_gemv_n<<<(rows + 127) / 128, 128, 0, st>>>(rows, cols, A, ld, tri, alpha, x, y);
  else k_gemv_c<<<(unsigned)(((size_t)cols * 32 + 127) / 128), 128, 0, st>>>(rows, cols, A, ld, tri, alpha, x, y);
  CNLD_CK_LAUNCH();
  return 0;
}

int trsv(cudaStream_t st, const cplx *A, size_t ld, uint n, bool lower, bool unit, bool adjoint, cplx *x) {
  if (!n) return 0;
  k_trsv<<<1, 1024, 0, st>>>(n, A, ld, lower ? 1 : 0, unit ? 1 : 0, adjoint ? 1 : 0, x);
  CNLD_CK_LAUNCH();
  return 0;
}

int potrf(cudaStream_t st, cplx *A, size_t ld, uint n, uint *d_info) {
  for (uint k0 = 0; k0 < n; k0 += CH_NB) {
    const uint nb = std::min((uint)CH_NB, n - k0), rest = n - k0 - nb;
    k_potf2<<<1, CH_NB * CH_NB, 0, st>>>(n, k0, A, ld, d_info);
    CNLD_CK_LAUNCH();
    if (!rest) break;
    k_chol_panel<<<(rest + 127) / 128, 128, 0, st>>>(n, k0, A, ld);
    CNLD_CK_LAUNCH();
    const uint tiles = (rest + 31) / 32;
    k_chol_update<<<dim3(tiles, tiles), 256, 0, st>>>(n, k0, nb, A, ld);
    CNLD_CK_LAUNCH();
  }
  return 0;
}

int axpy(cudaStream_t st, size_t n, cplx alpha, const cplx *x, cplx *y) {
  if (!n) return 0;
  k_axpy_c<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, alpha, x, y);
  CNLD_CK_LAUNCH();
  return 0;
}

int scal(cudaStream_t st, size_t n, cplx alpha, cplx *x) {
  if (!n) return 0;
  k_scal<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, alpha, x);
  CNLD_CK_LAUNCH();
  return 0;
}

// <x, y> = sum conj(x_i) y_i of device vectors (synchronises the stream)
int dotc(cudaStream_t st, size_t n, const cplx *x, const cplx *y, cplx *result) {
  static double *d_acc = nullptr;
  if (!d_acc) CNLD_CK(cudaMalloc(&d_acc, 2 * sizeof(double)));
  CNLD_CK(cudaMemsetAsync(d_acc, 0, 2 * sizeof(double), st));
  if (n) {
    k_dotc<<<(unsigned)std::min<size_t>((n + 255) / 256, 1024), 256, 0, st>>>(n, x, y, d_acc);
    CNLD_CK_LAUNCH();
  }
  double h[2] = {0, 0};
  CNLD_CK(cudaMemcpyAsync(h, d_acc, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CNLD_CK(cudaStreamSynchronize(st));
  *result = make_double2(h[0], h[1]);
  return 0;
}

// host-vector wrappers ---------------------------------------------------------------------------------
int with_vectors(uint nx, const cplx *hx, uint ny, cplx *hy, bool y_in, const std::function<int(const cplx *, cplx *)> &f) {
  cplx *dx = nullptr, *dy = nullptr;
  int rc = -1;
  do {
    if (cudaMalloc(&dx, sizeof(cplx) * std::max(nx, 1u)) != cudaSuccess || cudaMalloc(&dy, sizeof(cplx) * std::max(ny, 1u)) != cudaSuccess) break;
    if (nx && cudaMemcpy(dx, hx, sizeof(cplx) * nx, cudaMemcpyHostToDevice) != cudaSuccess) break;
    if (y_in ? (ny && cudaMemcpy(dy, hy, sizeof(cplx) * ny, cudaMemcpyHostToDevice) != cudaSuccess)
             : (cudaMemset(dy, 0, sizeof(cplx) * std::max(ny, 1u)) != cudaSuccess)) break;
    if (f(dx, dy)) { rc = -2; break; }
    if (ny && cudaMemcpy(hy, dy, sizeof(cplx) * ny, cudaMemcpyDeviceToHost) != cudaSuccess) break;
    rc = 0;
  } while (0);
  if (rc == -1) set_error("device vector transfer failed: %s", cudaGetErrorString(cudaGetLastError()));
  cudaFree(dx); cudaFree(dy);
  return rc;
}

}  // namespace dd
}  // namespace cnld
