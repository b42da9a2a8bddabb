// H2Lib-named entry points (include/cnld_b200_h2lib.h): the symbols cnl-dyna's Cython layer
// binds (cnld/h2lib/_*.pxd).  Containers live in host memory exactly as H2Lib's do (numpy
// views alias them); assembly / factorisation / solves run on the GPU through the same
// kernels as the batched sweep.  Mesh generation follows the H2Lib contracts in
// include/surface3d.h and include/macrosurface3d.h.
#include "../../include/cnld_b200_h2lib.h"

#include <cmath>
#include <algorithm>
#include <cstdlib>
#include <map>
#include <random>

#include "../../include/cnld_b200.h"
#include "common.cuh"
#include "densedev.cuh"

using cnld::set_error;

namespace {
inline field mk(double re, double im) { field f; f.re = re; f.im = im; return f; }
inline field fmul(field a, field b) { return mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
inline field fadd(field a, field b) { return mk(a.re + b.re, a.im + b.im); }
inline field fconj(field a) { return mk(a.re, -a.im); }
[[noreturn]] void die(const char *what) {
  // H2Lib asserts on contract violations; keep that behaviour (no exceptions over the C ABI)
  fprintf(stderr, "libcnld_b200: %s: %s\n", what, cnld_b200_last_error());
  abort();
}
}  // namespace

extern "C" {

/* ---------------------------------------------------------------- basic ---- */
void init_h2lib(int *, char ***) {}
void uninit_h2lib(void) {}
void freemem(void *ptr) { free(ptr); }

/* -------------------------------------------------------------- avector ---- */
pavector new_avector(uint dim) {
  pavector v = (pavector)malloc(sizeof(avector));
  v->v = (field *)malloc(sizeof(field) * (dim ? dim : 1));
  v->dim = dim;
  v->owner = nullptr;
  return v;
}
pavector new_zero_avector(uint dim) {
  pavector v = new_avector(dim);
  memset(v->v, 0, sizeof(field) * dim);
  return v;
}
void del_avector(pavector v) {
  if (!v) return;
  if (!v->owner) free(v->v);
  free(v);
}
void clear_avector(pavector v) { memset(v->v, 0, sizeof(field) * v->dim); }
void copy_avector(pcavector v, pavector w) {
  if (v->dim != w->dim) { set_error("copy_avector: dimension mismatch"); die("copy_avector"); }
  memcpy(w->v, v->v, sizeof(field) * v->dim);
}
void random_avector(pavector v) {
  static std::mt19937_64 rng(42);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  for (uint i = 0; i < v->dim; i++) v->v[i] = mk(U(rng), U(rng));
}

/* -------------------------------------------------------------- amatrix ---- */
pamatrix new_amatrix(uint rows, uint cols) {
  pamatrix a = (pamatrix)malloc(sizeof(amatrix));
  size_t n = (size_t)rows * cols;
  a->a = (field *)malloc(sizeof(field) * (n ? n : 1));
  a->ld = rows;
  a->rows = rows;
  a->cols = cols;
  a->owner = nullptr;
  return a;
}
pamatrix new_zero_amatrix(uint rows, uint cols) {
  pamatrix a = new_amatrix(rows, cols);
  memset(a->a, 0, sizeof(field) * (size_t)rows * cols);
  return a;
}
void del_amatrix(pamatrix a) {
  if (!a) return;
  cnld::dd::drop(a->a);
  if (!a->owner) free(a->a);
  free(a);
}
pamatrix clone_amatrix(pcamatrix src) {
  pamatrix a = new_amatrix(src->rows, src->cols);
  for (uint j = 0; j < src->cols; j++)
    memcpy(a->a + (size_t)j * a->ld, src->a + (size_t)j * src->ld, sizeof(field) * src->rows);
  return a;
}
void scale_amatrix(field alpha, pamatrix a) {
  cnld::dd::drop(a->a);
  for (uint j = 0; j < a->cols; j++) {
    field *c = a->a + (size_t)j * a->ld;
    for (uint i = 0; i < a->rows; i++) c[i] = fmul(alpha, c[i]);
  }
}
void conjugate_amatrix(pamatrix a) {
  cnld::dd::drop(a->a);
  for (uint j = 0; j < a->cols; j++) {
    field *c = a->a + (size_t)j * a->ld;
    for (uint i = 0; i < a->rows; i++) c[i].im = -c[i].im;
  }
}
void addeval_amatrix_avector(field alpha, pcamatrix a, pcavector src, pavector trg) {
  // trg += alpha * A src on the GPU; the device copy of A is kept for the next call (densedev.cu)
  if (src->dim != a->cols || trg->dim != a->rows) { set_error("addeval_amatrix_avector: dimension mismatch"); die("addeval"); }
  cnld::dd::Resident *R = cnld::dd::resident(a->a, a->rows, a->cols, a->ld);
  if (!R) die("addeval_amatrix_avector");
  const cplx al = make_double2(alpha.re, alpha.im);
  if (cnld::dd::with_vectors(src->dim, (const cplx *)src->v, trg->dim, (cplx *)trg->v, true, [&](const cplx *dx, cplx *dy) {
        return cnld::dd::gemv(0, R->d, R->rows, R->rows, R->cols, cnld::dd::TRI_FULL, 0, al, dx, dy);
      }))
    die("addeval_amatrix_avector");
}
void add_amatrix(field alpha, bool atrans, pcamatrix a, pamatrix b) {
  cnld::dd::drop(b->a);
  for (uint j = 0; j < b->cols; j++)
    for (uint i = 0; i < b->rows; i++) {
      field v = atrans ? fconj(a->a[(size_t)i * a->ld + j]) : a->a[(size_t)j * a->ld + i];
      field *d = b->a + (size_t)j * b->ld + i;
      *d = fadd(*d, fmul(alpha, v));
    }
}
size_t getsize_amatrix(pcamatrix a) { return sizeof(amatrix) + sizeof(field) * (size_t)a->rows * a->cols; }

void addmul_amatrix(field alpha, bool atrans, pcamatrix a, bool btrans, pcamatrix b, pamatrix c) {
  // C += alpha * op(A) op(B) on the DMMA ZGEMM; op = conjugate transpose as in H2Lib
  uint m = c->rows, n = c->cols, k = atrans ? a->rows : a->cols;
  std::vector<field> A((size_t)m * k), B((size_t)k * n), T((size_t)m * n, mk(0, 0));
  for (uint p = 0; p < k; p++)
    for (uint i = 0; i < m; i++)
      A[(size_t)p * m + i] = atrans ? fconj(a->a[(size_t)i * a->ld + p]) : a->a[(size_t)p * a->ld + i];
  for (uint j = 0; j < n; j++)
    for (uint p = 0; p < k; p++)
      B[(size_t)j * k + p] = btrans ? fconj(b->a[(size_t)p * b->ld + j]) : b->a[(size_t)j * b->ld + p];
  if (cnld_b200_zgemm(0, m, n, k, 1.0, (const cnld_field *)A.data(), m, (const cnld_field *)B.data(), k, 0.0,
                      (cnld_field *)T.data(), m))
    die("addmul_amatrix");
  for (uint j = 0; j < n; j++)
    for (uint i = 0; i < m; i++) {
      field *d = c->a + (size_t)j * c->ld + i;
      *d = fadd(*d, fmul(alpha, T[(size_t)j * m + i]));
    }
}

/* --------------------------------------------------------- sparsematrix ---- */
psparsematrix new_raw_sparsematrix(uint rows, uint cols, uint nz) {
  psparsematrix s = (psparsematrix)malloc(sizeof(sparsematrix));
  s->rows = rows; s->cols = cols; s->nz = nz;
  s->row = (uint *)calloc((size_t)rows + 1, sizeof(uint));
  s->col = (uint *)calloc(nz ? nz : 1, sizeof(uint));
  s->coeff = (field *)calloc(nz ? nz : 1, sizeof(field));
  return s;
}
psparsematrix new_identity_sparsematrix(uint rows, uint cols) {
  uint n = rows < cols ? rows : cols;
  psparsematrix s = new_raw_sparsematrix(rows, cols, n);
  for (uint i = 0; i <= rows; i++) s->row[i] = i < n ? i : n;
  for (uint i = 0; i < n; i++) { s->col[i] = i; s->coeff[i] = mk(1.0, 0.0); }
  return s;
}
void del_sparsematrix(psparsematrix a) {
  if (!a) return;
  free(a->row); free(a->col); free(a->coeff); free(a);
}
field addentry_sparsematrix(psparsematrix a, uint row, uint col, field x) {
  for (uint p = a->row[row]; p < a->row[row + 1]; p++)
    if (a->col[p] == col) { a->coeff[p] = fadd(a->coeff[p], x); return a->coeff[p]; }
  set_error("addentry_sparsematrix: (%u,%u) not in the pattern", row, col);
  die("addentry_sparsematrix");
}
void setentry_sparsematrix(psparsematrix a, uint row, uint col, field x) {
  for (uint p = a->row[row]; p < a->row[row + 1]; p++)
    if (a->col[p] == col) { a->coeff[p] = x; return; }
  set_error("setentry_sparsematrix: (%u,%u) not in the pattern", row, col);
  die("setentry_sparsematrix");
}
size_t getsize_sparsematrix(pcsparsematrix a) {
  return sizeof(sparsematrix) + sizeof(uint) * ((size_t)a->rows + 1 + a->nz) + sizeof(field) * a->nz;
}
void sort_sparsematrix(psparsematrix a) {
  // H2Lib: diagonal entry first in every row, the rest ascending by column
  for (uint i = 0; i < a->rows; i++) {
    uint lo = a->row[i], hi = a->row[i + 1];
    std::vector<std::pair<uint, field>> r;
    for (uint p = lo; p < hi; p++) r.push_back({a->col[p], a->coeff[p]});
    std::stable_sort(r.begin(), r.end(), [i](const std::pair<uint, field> &x, const std::pair<uint, field> &y) {
      if ((x.first == i) != (y.first == i)) return x.first == i;
      return x.first < y.first;
    });
    for (uint p = lo; p < hi; p++) { a->col[p] = r[p - lo].first; a->coeff[p] = r[p - lo].second; }
  }
}
void clear_sparsematrix(psparsematrix a) { memset(a->coeff, 0, sizeof(field) * a->nz); }
void add_sparsematrix_amatrix(field alpha, bool atrans, pcsparsematrix a, pamatrix b) {
  for (uint i = 0; i < a->rows; i++)
    for (uint p = a->row[i]; p < a->row[i + 1]; p++) {
      uint j = a->col[p];
      field v = a->coeff[p];
      field *d = atrans ? b->a + (size_t)i * b->ld + j : b->a + (size_t)j * b->ld + i;
      if (atrans) v = fconj(v);
      *d = fadd(*d, fmul(alpha, v));
    }
}
void addeval_sparsematrix_avector(field alpha, pcsparsematrix a, pcavector x, pavector y) {
  for (uint i = 0; i < a->rows; i++) {
    field s = mk(0, 0);
    for (uint p = a->row[i]; p < a->row[i + 1]; p++) s = fadd(s, fmul(a->coeff[p], x->v[a->col[p]]));
    y->v[i] = fadd(y->v[i], fmul(alpha, s));
  }
}

/* ------------------------------------------------------------ surface3d ---- */
psurface3d new_surface3d(uint vertices, uint edges, uint triangles) {
  psurface3d g = (psurface3d)calloc(1, sizeof(surface3d));
  g->vertices = vertices; g->edges = edges; g->triangles = triangles;
  g->x = (real(*)[3])calloc((size_t)vertices ? vertices : 1, sizeof(real[3]));
  g->e = (uint(*)[2])calloc((size_t)edges ? edges : 1, sizeof(uint[2]));
  g->t = (uint(*)[3])calloc((size_t)triangles ? triangles : 1, sizeof(uint[3]));
  g->s = (uint(*)[3])calloc((size_t)triangles ? triangles : 1, sizeof(uint[3]));
  g->n = (real(*)[3])calloc((size_t)triangles ? triangles : 1, sizeof(real[3]));
  g->g = (real *)calloc((size_t)triangles ? triangles : 1, sizeof(real));
  return g;
}
void del_surface3d(psurface3d g) {
  if (!g) return;
  free(g->x); free(g->e); free(g->t); free(g->s); free(g->n); free(g->g); free(g);
}
void prepare_surface3d(psurface3d gr) {
  // normals, Gram determinants (2*area), hmin/hmax over the triangle edges
  real hmin = 1e300, hmax = 0.0;
  for (uint i = 0; i < gr->triangles; i++) {
    const real *A = gr->x[gr->t[i][0]], *B = gr->x[gr->t[i][1]], *C = gr->x[gr->t[i][2]];
    real ab[3], ac[3], bc[3];
    for (int d = 0; d < 3; d++) { ab[d] = B[d] - A[d]; ac[d] = C[d] - A[d]; bc[d] = C[d] - B[d]; }
    real nx = ab[1] * ac[2] - ab[2] * ac[1], ny = ab[2] * ac[0] - ab[0] * ac[2], nz = ab[0] * ac[1] - ab[1] * ac[0];
    real len = std::sqrt(nx * nx + ny * ny + nz * nz);
    gr->g[i] = len;
    gr->n[i][0] = nx / len; gr->n[i][1] = ny / len; gr->n[i][2] = nz / len;
    for (const real *v : {ab, ac, bc}) {
      real h = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
      hmin = std::min(hmin, h);
      hmax = std::max(hmax, h);
    }
  }
  gr->hmin = hmin;
  gr->hmax = hmax;
}
psurface3d merge_surface3d(pcsurface3d a, pcsurface3d b) {
  psurface3d g = new_surface3d(a->vertices + b->vertices, a->edges + b->edges, a->triangles + b->triangles);
  memcpy(g->x, a->x, sizeof(real[3]) * a->vertices);
  memcpy(g->x + a->vertices, b->x, sizeof(real[3]) * b->vertices);
  memcpy(g->e, a->e, sizeof(uint[2]) * a->edges);
  memcpy(g->t, a->t, sizeof(uint[3]) * a->triangles);
  memcpy(g->s, a->s, sizeof(uint[3]) * a->triangles);
  for (uint i = 0; i < b->edges; i++)
    for (int d = 0; d < 2; d++) g->e[a->edges + i][d] = b->e[i][d] + a->vertices;
  for (uint i = 0; i < b->triangles; i++)
    for (int d = 0; d < 3; d++) {
      g->t[a->triangles + i][d] = b->t[i][d] + a->vertices;
      g->s[a->triangles + i][d] = b->s[i][d] + a->edges;
    }
  prepare_surface3d(g);
  return g;
}
psurface3d refine_red_surface3d(psurface3d gr) {
  // regular ("red") refinement: edge midpoints become vertices, each triangle -> 4
  const uint V = gr->vertices, E = gr->edges, T = gr->triangles;
  psurface3d r = new_surface3d(V + E, 2 * E + 3 * T, 4 * T);
  memcpy(r->x, gr->x, sizeof(real[3]) * V);
  for (uint i = 0; i < E; i++) {
    for (int d = 0; d < 3; d++) r->x[V + i][d] = 0.5 * (gr->x[gr->e[i][0]][d] + gr->x[gr->e[i][1]][d]);
    r->e[2 * i][0] = gr->e[i][0]; r->e[2 * i][1] = V + i;       // half edge at e[i][0]
    r->e[2 * i + 1][0] = V + i;   r->e[2 * i + 1][1] = gr->e[i][1];  // half edge at e[i][1]
  }
  auto half = [&](uint edge, uint vertex) { return gr->e[edge][0] == vertex ? 2 * edge : 2 * edge + 1; };
  for (uint i = 0; i < T; i++) {
    const uint *v = gr->t[i], *s = gr->s[i];
    uint m[3] = {V + s[0], V + s[1], V + s[2]};  // midpoint opposite vertex j
    uint ie[3];                                  // inner edge opposite corner j: m[j+1]-m[j+2]
    for (int j = 0; j < 3; j++) {
      ie[j] = 2 * E + 3 * i + j;
      r->e[ie[j]][0] = m[(j + 1) % 3];
      r->e[ie[j]][1] = m[(j + 2) % 3];
    }
    for (int j = 0; j < 3; j++) {  // corner triangle at v[j]: (v[j], m[j+2], m[j+1])
      uint *t = r->t[4 * i + j], *ss = r->s[4 * i + j];
      t[0] = v[j]; t[1] = m[(j + 2) % 3]; t[2] = m[(j + 1) % 3];
      ss[0] = ie[j];
      ss[1] = half(s[(j + 1) % 3], v[j]);
      ss[2] = half(s[(j + 2) % 3], v[j]);
    }
    uint *t = r->t[4 * i + 3], *ss = r->s[4 * i + 3];
    t[0] = m[0]; t[1] = m[1]; t[2] = m[2];
    ss[0] = ie[0]; ss[1] = ie[1]; ss[2] = ie[2];
  }
  prepare_surface3d(r);
  return r;
}
void translate_surface3d(psurface3d gr, real *t) {
  for (uint i = 0; i < gr->vertices; i++)
    for (int d = 0; d < 3; d++) gr->x[i][d] += t[d];
}
uint check_surface3d(pcsurface3d gr) {
  uint problems = 0;
  for (uint i = 0; i < gr->edges; i++)
    if (gr->e[i][0] >= gr->vertices || gr->e[i][1] >= gr->vertices || gr->e[i][0] == gr->e[i][1]) problems++;
  for (uint i = 0; i < gr->triangles; i++)
    for (int j = 0; j < 3; j++) {
      if (gr->t[i][j] >= gr->vertices || gr->s[i][j] >= gr->edges) { problems++; continue; }
      uint a = gr->t[i][(j + 1) % 3], b = gr->t[i][(j + 2) % 3];
      const uint *e = gr->e[gr->s[i][j]];
      if (!((e[0] == a && e[1] == b) || (e[0] == b && e[1] == a))) problems++;
    }
  return problems;
}
bool isclosed_surface3d(pcsurface3d gr) {
  std::vector<uint> cnt(gr->edges, 0);
  for (uint i = 0; i < gr->triangles; i++)
    for (int j = 0; j < 3; j++) cnt[gr->s[i][j]]++;
  for (uint c : cnt) if (c != 2) return false;
  return true;
}
bool isoriented_surface3d(pcsurface3d gr) {
  // every inner edge must be traversed in opposite directions by its two triangles
  std::vector<int> dir(gr->edges, 0);
  for (uint i = 0; i < gr->triangles; i++)
    for (int j = 0; j < 3; j++) {
      uint a = gr->t[i][(j + 1) % 3];
      uint ed = gr->s[i][j];
      int d = (gr->e[ed][0] == a) ? 1 : -1;
      if (dir[ed] == d) return false;
      dir[ed] += d;
    }
  return true;
}

/* ------------------------------------------------------- macrosurface3d ---- */
void cnld_b200_phi_affine(uint i, real a, real b, void *data, real xt[3]) {
  pcmacrosurface3d mg = (pcmacrosurface3d)data;
  const real *p0 = mg->x[mg->t[i][0]], *p1 = mg->x[mg->t[i][1]], *p2 = mg->x[mg->t[i][2]];
  for (int d = 0; d < 3; d++) xt[d] = p0[d] * (1.0 - a - b) + p1[d] * a + p2[d] * b;
}
void cnld_b200_phi_circle(uint i, real a, real b, void *data, real xt[3]) {
  cnld_b200_phi_affine(i, a, b, data, xt);  // square -> disc, cnld/h2lib/macrosurface3d.pyx:124-145
  real l1 = std::fabs(xt[0]) + std::fabs(xt[1]), l2 = std::sqrt(xt[0] * xt[0] + xt[1] * xt[1]);
  if (l2 > 0.0) { xt[0] = xt[0] * l1 / l2; xt[1] = xt[1] * l1 / l2; }
}
void cnld_b200_phi_sphere(uint i, real a, real b, void *data, real xt[3]) {
  cnld_b200_phi_affine(i, a, b, data, xt);
  real n = std::sqrt(xt[0] * xt[0] + xt[1] * xt[1] + xt[2] * xt[2]);
  for (int d = 0; d < 3; d++) xt[d] /= n;
}
pmacrosurface3d new_macrosurface3d(uint vertices, uint edges, uint triangles) {
  pmacrosurface3d m = (pmacrosurface3d)calloc(1, sizeof(macrosurface3d));
  m->vertices = vertices; m->edges = edges; m->triangles = triangles;
  m->x = (real(*)[3])calloc(vertices ? vertices : 1, sizeof(real[3]));
  m->e = (uint(*)[2])calloc(edges ? edges : 1, sizeof(uint[2]));
  m->t = (uint(*)[3])calloc(triangles ? triangles : 1, sizeof(uint[3]));
  m->s = (uint(*)[3])calloc(triangles ? triangles : 1, sizeof(uint[3]));
  m->phi = cnld_b200_phi_affine;
  m->phidata = m;
  return m;
}
void del_macrosurface3d(pmacrosurface3d m) {
  if (!m) return;
  free(m->x); free(m->e); free(m->t); free(m->s); free(m);
}
pmacrosurface3d new_sphere_macrosurface3d(void) {
  // octahedron, outward-oriented
  pmacrosurface3d m = new_macrosurface3d(6, 12, 8);
  const real X[6][3] = {{0, 0, 1}, {1, 0, 0}, {0, 1, 0}, {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}};
  memcpy(m->x, X, sizeof(X));
  const uint E[12][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {1, 2}, {2, 3}, {3, 4}, {4, 1}, {5, 1}, {5, 2}, {5, 3}, {5, 4}};
  memcpy(m->e, E, sizeof(E));
  const uint T[8][3] = {{0, 1, 2}, {0, 2, 3}, {0, 3, 4}, {0, 4, 1}, {5, 2, 1}, {5, 3, 2}, {5, 4, 3}, {5, 1, 4}};
  memcpy(m->t, T, sizeof(T));
  std::map<std::pair<uint, uint>, uint> eid;
  for (uint i = 0; i < 12; i++) eid[{std::min(E[i][0], E[i][1]), std::max(E[i][0], E[i][1])}] = i;
  for (uint i = 0; i < 8; i++)
    for (int j = 0; j < 3; j++) {
      uint a = T[i][(j + 1) % 3], b = T[i][(j + 2) % 3];
      m->s[i][j] = eid[{std::min(a, b), std::max(a, b)}];
    }
  m->phi = cnld_b200_phi_sphere;
  return m;
}

psurface3d build_from_macrosurface3d_surface3d(pcmacrosurface3d mg, uint split) {
  // Numbering: macro vertices; then split-1 inner vertices per macro edge, walking e[.][0] ->
  // e[.][1]; then the inner lattice points of each macro triangle, row by row.  Edges: `split`
  // pieces per macro edge, then the inner edges of each macro triangle.  Triangles keep the
  // macro triangle's orientation.
  const uint n = split, MV = mg->vertices, ME = mg->edges, MT = mg->triangles;
  if (n < 1) { set_error("build_from_macrosurface3d_surface3d: split must be >= 1"); die("macrosurface3d"); }
  const uint inner = (n >= 2) ? (n - 1) * (n - 2) / 2 : 0;
  psurface3d gr = new_surface3d(MV + ME * (n - 1) + MT * inner, ME * n + MT * (3 * n * (n - 1) / 2), MT * n * n);
  auto edge_vertex = [&](uint E, uint k) -> uint {  // k-th lattice point of macro edge E, k = 0..n
    if (k == 0) return mg->e[E][0];
    if (k == n) return mg->e[E][1];
    return MV + E * (n - 1) + (k - 1);
  };
  for (uint E = 0; E < ME; E++)
    for (uint k = 0; k < n; k++) { gr->e[E * n + k][0] = edge_vertex(E, k); gr->e[E * n + k][1] = edge_vertex(E, k + 1); }
  uint nextv = MV + ME * (n - 1), nexte = ME * n, nextt = 0;
  std::vector<char> placed(MV + ME * (n - 1), 0);
  const uint rowlen = n + 1;
  std::vector<uint> vid(rowlen * rowlen), eh(rowlen * rowlen), ev(rowlen * rowlen), ed(rowlen * rowlen);
  const uint NONE = ~0u;
  auto at = [&](uint i, uint j) { return j * rowlen + i; };
  for (uint T = 0; T < MT; T++) {
    const uint *tv = mg->t[T];
    std::fill(vid.begin(), vid.end(), NONE);
    std::fill(eh.begin(), eh.end(), NONE);
    std::fill(ev.begin(), ev.end(), NONE);
    std::fill(ed.begin(), ed.end(), NONE);
    // the three sides: side c lies opposite corner c and runs corner c+1 -> corner c+2
    for (uint c = 0; c < 3; c++) {
      const uint E = mg->s[T][c], from = tv[(c + 1) % 3], to = tv[(c + 2) % 3];
      bool fwd;
      if (mg->e[E][0] == from && mg->e[E][1] == to) fwd = true;
      else if (mg->e[E][0] == to && mg->e[E][1] == from) fwd = false;
      else { set_error("macrosurface3d: edge %u does not match triangle %u", E, T); die("macrosurface3d"); }
      for (uint k = 0; k <= n; k++) {
        uint i, j;
        switch (c) {
        case 0: i = n - k; j = k; break;   // (n,0) -> (0,n)
        case 1: i = 0; j = n - k; break;   // (0,n) -> (0,0)
        default: i = k; j = 0;             // (0,0) -> (n,0)
        }
        vid[at(i, j)] = edge_vertex(E, fwd ? k : n - k);
        if (k < n) {
          uint piece = E * n + (fwd ? k : n - 1 - k);
          if (c == 0) ed[at(n - k - 1, k)] = piece;
          else if (c == 1) ev[at(0, n - k - 1)] = piece;
          else eh[at(k, 0)] = piece;
        }
      }
    }
    for (uint j = 0; j <= n; j++)
      for (uint i = 0; i + j <= n; i++) {
        uint &v = vid[at(i, j)];
        if (v == NONE) v = nextv++;
        else if (placed[v]) continue;
        else placed[v] = 1;
        mg->phi(T, (real)i / n, (real)j / n, mg->phidata, gr->x[v]);
      }
    auto new_edge = [&](uint a, uint b) { gr->e[nexte][0] = a; gr->e[nexte][1] = b; return nexte++; };
    for (uint j = 0; j < n; j++)
      for (uint i = 0; i + j < n; i++) {
        if (eh[at(i, j)] == NONE) eh[at(i, j)] = new_edge(vid[at(i, j)], vid[at(i + 1, j)]);
        if (ev[at(i, j)] == NONE) ev[at(i, j)] = new_edge(vid[at(i, j)], vid[at(i, j + 1)]);
        if (ed[at(i, j)] == NONE) ed[at(i, j)] = new_edge(vid[at(i + 1, j)], vid[at(i, j + 1)]);
      }
    for (uint j = 0; j < n; j++)
      for (uint i = 0; i + j < n; i++) {
        uint *t = gr->t[nextt], *s = gr->s[nextt];
        t[0] = vid[at(i, j)]; t[1] = vid[at(i + 1, j)]; t[2] = vid[at(i, j + 1)];
        s[0] = ed[at(i, j)]; s[1] = ev[at(i, j)]; s[2] = eh[at(i, j)];
        nextt++;
        if (i + j + 1 < n) {
          t = gr->t[nextt]; s = gr->s[nextt];
          t[0] = vid[at(i + 1, j)]; t[1] = vid[at(i + 1, j + 1)]; t[2] = vid[at(i, j + 1)];
          s[0] = eh[at(i, j + 1)]; s[1] = ed[at(i, j)]; s[2] = ev[at(i + 1, j)];
          nextt++;
        }
      }
  }
  prepare_surface3d(gr);
  return gr;
}

/* ---------------------------------------------------------------- bem3d ---- */
pbem3d new_bem3d(pcsurface3d gr, basisfunctionbem3d row_basis, basisfunctionbem3d col_basis) {
  pbem3d b = (pbem3d)calloc(1, sizeof(bem3d));
  b->gr = gr;
  b->row_basis = row_basis;
  b->col_basis = col_basis;
  b->kernel_const = mk(0.0795774715459476678844418816863, 0.0);
  return b;
}
void del_bem3d(pbem3d bem) {
  if (!bem) return;
  if (bem->priv) cnld_b200_destroy((cnld_ctx *)bem->priv);
  free(bem);
}
pbem3d new_slp_helmholtz_bem3d(field k, pcsurface3d gr, uint q_regular, uint q_singular,
                               basisfunctionbem3d row_basis, basisfunctionbem3d col_basis) {
  if (row_basis != BASIS_LINEAR_BEM3D || col_basis != BASIS_LINEAR_BEM3D) {
    set_error("new_slp_helmholtz_bem3d: only linear/linear basis functions are on the cnl-dyna hot path");
    return nullptr;
  }
  pbem3d b = new_bem3d(gr, row_basis, col_basis);
  b->k = k;
  b->q_regular = q_regular;
  b->q_singular = q_singular;
  b->priv = cnld_b200_create(cnld::dd::device_id(), gr->vertices, gr->triangles, &gr->x[0][0], &gr->t[0][0], q_regular, q_singular);
  if (!b->priv) { free(b); return nullptr; }
  return b;
}
/* The double-layer operator is bound by cnld/h2lib/_helmholtzbem3d.pxd:13 but never constructed by
 * cnl-dyna (ZFullMatrix / ZHMatrix build the single-layer operator only, compressed_formats.py:629,675);
 * it is outside the hot path and has no kernel here.  The symbol exists so that the
 * reference's extension modules link; calling it reports that and returns NULL. */
pbem3d new_dlp_helmholtz_bem3d(field, pcsurface3d, uint, uint, basisfunctionbem3d, basisfunctionbem3d, field) {
  set_error("new_dlp_helmholtz_bem3d: the double-layer potential operator is not implemented in libcnld_b200 "
            "(cnl-dyna's frequency-response path only uses new_slp_helmholtz_bem3d)");
  fprintf(stderr, "libcnld_b200: %s\n", cnld_b200_last_error());
  return nullptr;
}
void assemble_bem3d_amatrix(pbem3d bem, pamatrix G) {
  cnld::dd::drop(G->a);
  if (!bem || !bem->priv) { set_error("assemble_bem3d_amatrix: bem object has no device state"); die("assemble_bem3d_amatrix"); }
  if (G->rows != bem->gr->vertices || G->cols != bem->gr->vertices) { set_error("assemble_bem3d_amatrix: matrix must be vertices x vertices"); die("assemble_bem3d_amatrix"); }
  if (cnld_b200_assemble_slp_ll((cnld_ctx *)bem->priv, bem->k.re, bem->k.im, (cnld_field *)G->a, G->ld))
    die("assemble_bem3d_amatrix");
}

/* ------------------------------------------------------- factorizations ---- */
uint lrdecomp_amatrix(pamatrix a) {
  // unpivoted LU on the GPU; the factors stay resident so that the per-source-patch lrsolve calls of
  // generate_db.process (cnld/scripts/generate_db.py:82-91) do not re-upload them
  if (a->rows != a->cols) { set_error("lrdecomp_amatrix: matrix must be square"); die("lrdecomp_amatrix"); }
  const uint n = a->rows;
  if (cnld_b200_device_count() <= cnld::dd::device_id()) { set_error("no CUDA device %d (libcnld_b200 has no CPU fallback)", cnld::dd::device_id()); die("lrdecomp_amatrix"); }
  cnld::dd::drop(a->a);
  cudaSetDevice(cnld::dd::device_id());
  cplx *dA = nullptr;
  cnld::LuWork w;
  uint info = 0;
  bool ok = cudaMalloc(&dA, sizeof(cplx) * std::max<size_t>((size_t)n * n, 1)) == cudaSuccess;
  if (!ok) { cnld::dd::drop_all(); ok = cudaMalloc(&dA, sizeof(cplx) * std::max<size_t>((size_t)n * n, 1)) == cudaSuccess; }
  ok = ok && (!n || cudaMemcpy2D(dA, sizeof(cplx) * n, a->a, sizeof(cplx) * a->ld, sizeof(cplx) * n, n, cudaMemcpyHostToDevice) == cudaSuccess);
  ok = ok && !cnld::lu_work_alloc(w, n) && !cnld::lu_factor(0, w, dA, n);
  ok = ok && cudaMemcpy(&info, w.info, sizeof(uint), cudaMemcpyDeviceToHost) == cudaSuccess;
  ok = ok && (!n || cudaMemcpy2D(a->a, sizeof(cplx) * a->ld, dA, sizeof(cplx) * n, sizeof(cplx) * n, n, cudaMemcpyDeviceToHost) == cudaSuccess);
  if (!ok) {
    if (cudaGetLastError() != cudaSuccess) set_error("lrdecomp_amatrix: CUDA failure");
    cnld::lu_work_free(w);
    cudaFree(dA);
    die("lrdecomp_amatrix");
  }
  if (info) { cnld::lu_work_free(w); cudaFree(dA); return info; }
  cnld::dd::Resident *R = cnld::dd::adopt(a->a, n, n, a->ld, dA, cnld::dd::KIND_LU);
  R->w = w;
  R->w_ready = true;
  return 0;
}
static cnld::dd::Resident *resident_lu(pcamatrix a, const char *who) {
  cnld::dd::Resident *R = cnld::dd::resident(a->a, a->rows, a->cols, a->ld, cnld::dd::KIND_LU);
  if (!R) die(who);
  if (!R->w_ready) {   // factors that were not computed in this process' cache: rebuild the block inverses once
    if (cnld::lu_work_alloc(R->w, a->rows) || cnld::lu_invert_diag(0, R->w, R->d, R->rows)) die(who);
    R->w_ready = true;
    R->kind = cnld::dd::KIND_LU;
  }
  return R;
}
void lrsolve_amatrix_avector(bool atrans, pcamatrix a, pavector x) {
  if (a->rows != a->cols || x->dim != a->rows) { set_error("lrsolve_amatrix_avector: dimension mismatch"); die("lrsolve_amatrix_avector"); }
  cnld::dd::Resident *R = resident_lu(a, "lrsolve_amatrix_avector");
  const uint n = a->rows;
  int rc;
  if (!atrans) {
    rc = cnld::dd::with_vectors(0, nullptr, n, (cplx *)x->v, true,
                                [&](const cplx *, cplx *dx) { return cnld::lu_solve(0, R->w, R->d, n, dx, 1, n); });
  } else {   // (L R)^H x = b:  R^H y = b, L^H x = y
    rc = cnld::dd::with_vectors(0, nullptr, n, (cplx *)x->v, true, [&](const cplx *, cplx *dx) {
      return cnld::dd::trsv(0, R->d, n, n, false, false, true, dx) || cnld::dd::trsv(0, R->d, n, n, true, true, true, dx);
    });
  }
  if (rc) die("lrsolve_amatrix_avector");
}
void triangularsolve_amatrix_avector(bool alower, bool aunit, bool atrans, pcamatrix a, pavector x) {
  if (a->rows != a->cols || x->dim != a->rows) { set_error("triangularsolve_amatrix_avector: dimension mismatch"); die("triangularsolve_amatrix_avector"); }
  cnld::dd::Resident *R = cnld::dd::resident(a->a, a->rows, a->cols, a->ld);
  if (!R) die("triangularsolve_amatrix_avector");
  const uint n = a->rows;
  if (cnld::dd::with_vectors(0, nullptr, n, (cplx *)x->v, true,
                             [&](const cplx *, cplx *dx) { return cnld::dd::trsv(0, R->d, n, n, alower, aunit, atrans, dx); }))
    die("triangularsolve_amatrix_avector");
}
/* Cholesky A = L L^H of a Hermitian positive definite matrix, lower triangle in place (upper part
 * zeroed); returns 0 or the index (+1) of the first non-positive pivot (include/factorizations.h:281). */
uint choldecomp_amatrix(pamatrix a) {
  if (a->rows != a->cols) { set_error("choldecomp_amatrix: matrix must be square"); die("choldecomp_amatrix"); }
  const uint n = a->rows;
  cnld::dd::drop(a->a);
  cnld::dd::Resident *R = cnld::dd::resident(a->a, n, n, a->ld, cnld::dd::KIND_CHOL);
  if (!R) die("choldecomp_amatrix");
  uint *d_info = nullptr, info = 0;
  bool ok = cudaMalloc(&d_info, sizeof(uint)) == cudaSuccess && cudaMemset(d_info, 0, sizeof(uint)) == cudaSuccess;
  ok = ok && !cnld::dd::potrf(0, R->d, n, n, d_info);
  ok = ok && cudaMemcpy(&info, d_info, sizeof(uint), cudaMemcpyDeviceToHost) == cudaSuccess;
  ok = ok && (!n || cudaMemcpy2D(a->a, sizeof(cplx) * a->ld, R->d, sizeof(cplx) * n, sizeof(cplx) * n, n, cudaMemcpyDeviceToHost) == cudaSuccess);
  cudaFree(d_info);
  if (!ok) { set_error("choldecomp_amatrix: CUDA failure: %s", cudaGetErrorString(cudaGetLastError())); die("choldecomp_amatrix"); }
  cplx *keep = R->d;
  R->d = nullptr;                       // re-register under the new host content (the factor)
  cnld::dd::drop(a->a);
  if (info) { cudaFree(keep); return info; }
  cnld::dd::adopt(a->a, n, n, a->ld, keep, cnld::dd::KIND_CHOL);
  return 0;
}
void cholsolve_amatrix_avector(pcamatrix a, pavector x) {
  if (a->rows != a->cols || x->dim != a->rows) { set_error("cholsolve_amatrix_avector: dimension mismatch"); die("cholsolve_amatrix_avector"); }
  cnld::dd::Resident *R = cnld::dd::resident(a->a, a->rows, a->cols, a->ld, cnld::dd::KIND_CHOL);
  if (!R) die("cholsolve_amatrix_avector");
  const uint n = a->rows;
  if (cnld::dd::with_vectors(0, nullptr, n, (cplx *)x->v, true, [&](const cplx *, cplx *dx) {
        return cnld::dd::trsv(0, R->d, n, n, true, false, false, dx) || cnld::dd::trsv(0, R->d, n, n, true, false, true, dx);
      }))
    die("cholsolve_amatrix_avector");
}

}  // extern "C"
