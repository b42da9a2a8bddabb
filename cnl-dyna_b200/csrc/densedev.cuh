// Device-resident dense matrices + dense helper kernels of the H2Lib-named entry points (densedev.cu).
#pragma once
#include <functional>
#include <vector>

#include "common.cuh"

namespace cnld {
namespace dd {

enum { TRI_FULL = 0, TRI_UPPER = 1, TRI_UNIT_LOWER = 2, TRI_LOWER = 3 };
enum { KIND_PLAIN = 0, KIND_LU = 1, KIND_CHOL = 2 };

struct Resident {
  const void *host = nullptr;   // the host array this is the copy of (identity + fingerprint)
  uint rows = 0, cols = 0;
  size_t ldh = 0;
  cplx *d = nullptr;            // rows x cols, leading dimension rows
  int kind = KIND_PLAIN;
  LuWork w;                     // KIND_LU: block inverses of the factors
  bool w_ready = false;
  std::vector<cplx> finger;
};

int device_id();   // CNLD_B200_DEVICE, else the calling thread's current CUDA device
Resident *find(const void *host, uint rows, uint cols, size_t ldh);
Resident *resident(const void *host, uint rows, uint cols, size_t ldh, int kind = KIND_PLAIN);   // find or upload
Resident *adopt(const void *host, uint rows, uint cols, size_t ldh, cplx *d, int kind);
void drop(const void *host);
void drop_all();

// y += alpha * op(T(A)) x on device vectors; T = full / upper / unit lower / lower part, op = identity or adjoint
int gemv(cudaStream_t st, const cplx *A, size_t ld, uint rows, uint cols, int tri, int adjoint, cplx alpha, const cplx *x,
         cplx *y);
int trsv(cudaStream_t st, const cplx *A, size_t ld, uint n, bool lower, bool unit, bool adjoint, cplx *x);
int potrf(cudaStream_t st, cplx *A, size_t ld, uint n, uint *d_info);   // A = L L^H in place (lower), *d_info = failing pivot + 1
int axpy(cudaStream_t st, size_t n, cplx alpha, const cplx *x, cplx *y);
int scal(cudaStream_t st, size_t n, cplx alpha, cplx *x);
int dotc(cudaStream_t st, size_t n, const cplx *x, const cplx *y, cplx *result);   // sum conj(x) y, synchronises
// f(dx, dy) with device copies of host vectors x (in) and y (in/out if y_in, else zero-initialised out)
int with_vectors(uint nx, const cplx *hx, uint ny, cplx *hy, bool y_in, const std::function<int(const cplx *, cplx *)> &f);

}  // namespace dd
}  // namespace cnld
