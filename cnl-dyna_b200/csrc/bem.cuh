// Device-side BEM context (mesh + quadrature tables), see assembly.cu.
#pragma once
#include "common.cuh"
#include "quadrature.h"
#include "periodic.h"

namespace cnld {

struct SingTables {
  const double *p[4] = {nullptr, nullptr, nullptr, nullptr};  // per family: xt|xs|yt|ys|w, n each
  uint n[4] = {0, 0, 0, 0};
};

struct RegRule {        // single-triangle Duffy-Gauss rule, passed by value to kernels
  double w[16];         // weights (sum 1/2)
  double phi[16][3];    // hat values (1-t, t-s, s) at the points
  double wphi[16][3];   // w * phi
};

struct BemDev {
  int device = 0;
  uint nv = 0, nt = 0, q = 0, q2 = 0, nq = 0;
  double *x = nullptr;   // [nv][3]
  uint *t = nullptr;     // [nt][3]
  double *g = nullptr;   // [nt]   Gram determinant = 2*area
  double *n = nullptr;   // [nt][3] unit normals
  double *qp = nullptr;  // [nt][nq][3] regular-rule points in global coordinates
  RegRule rule;
  SingTables sing;
  TouchPair *pairs = nullptr;
  size_t npairs = 0;
  // sparse pattern of the entries fed by touching pairs (symmetric path, see quadrature.h)
  uint nslot = 0;
  uint *pair_slot = nullptr, *slot_row = nullptr, *slot_col = nullptr, *slot_tr = nullptr;
  uint *slot_rowptr = nullptr;  // nv + 1: slots are sorted by (row, col)
  // symmetric-path work lists (periodic.h): [0] integrate every pair s < t, [1] one block per
  // translation class + copies (only when the mesh is translated copies of one component)
  struct SymPlan {
    uint nwork = 0, ncopy = 0, blk = 0;
    WorkItem *work = nullptr;
    CopyItem *copy = nullptr;
    uint nslot0 = 0;       // slots of the integrated component(s)
    size_t npairs0 = 0;    // touching pairs of the integrated component(s)
  } plan[2];
  bool periodic = false;   // plan[1] is valid
  int use_periodic = 1;    // use it when valid
  uint ncomp = 1, nclass = 0;
};

int bem_create(BemDev &b, int device, uint nv, uint nt, const double *x, const uint *t, uint q, uint q2);
void bem_free(BemDev &b);
// Z += alpha * Z_slp(k); Z must be initialised by the caller.
int bem_assemble(cudaStream_t st, const BemDev &b, double kr, double ki, cplx alpha, cplx *Z, size_t ld);

// Symmetric path: G_lower += alpha * Z_lower(k) where every disjoint pair is evaluated once (its
// transposed contribution is identical), touching pairs are accumulated per sparse slot in `sval`
// (caller-provided, nslot entries), folded into the lower triangle, and the asymmetry of the
// Sauter-Schwab entries is returned in dval[slot] = alpha * (Z[r,c] - Z[c,r]) for slots with r < c.
int bem_assemble_sym(cudaStream_t st, const BemDev &b, double kr, double ki, cplx alpha, cplx *G, size_t ld,
                     cplx *sval, cplx *dval);
}  // namespace cnld

namespace cnld {
// pressure.cu
int pres_vector(cudaStream_t st, const BemDev &b, const double *d_obs, uint nobs, double k, double c,
                double rho, cplx *d_pvec);
int pres_points(cudaStream_t st, const BemDev &b, double *d_sp);
// many-source path: triangle chunks (host-built once per mesh) -> node-space pressure vectors of a
// tile of observation points -> one ZGEMM against the displacements
struct PresPlan {
  int variant = 0;                     // kernel geometry the chunks were cut for (PV_VARIANTS)
  uint nchunk = 0, nentries = 0;       // nentries = sum of chunk-local nodes (PV writes per observation point)
  uint *tptr = nullptr, *nptr = nullptr, *node = nullptr;
  double *tri = nullptr;
  std::vector<uint> untouched;         // vertices no triangle references
};
int pres_plan_build(const BemDev &b, const double *hx, const uint *ht, PresPlan &p);
void pres_plan_free(PresPlan &p);
int pres_plan_stats(uint nv, uint nt, const double *hx, const uint *ht, uint *counts, uint *order);
uint pres_tile_obs();
uint pres_tile_ctas();
int pres_set_variant(int v);
extern int g_pvec_variant;
int pres_pvec_tiles(cudaStream_t st, const PresPlan &p, const BemDev &b, const double *d_obs, uint mc, double k,
                    double c, double rho, cplx *d_PV, size_t lda);
int pres_contract(cudaStream_t st, const BemDev &b, const cplx *d_PV, size_t lda, uint mc, const cplx *d_disp,
                  uint nsrc, cplx *d_out, size_t ldo);
int pressure_field(cudaStream_t st, const BemDev &b, const double *d_obs, uint nobs, const double *d_sp,
                   double k, double c, double rho, const cplx *d_disp, uint nsrc, cplx *d_D, cplx *d_out);
}  // namespace cnld
