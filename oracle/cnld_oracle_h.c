/* ============================================================================
 * cnld_oracle_h.c -- CPU ORACLE, hierarchical-matrix part (test infrastructure).
 *
 * Restates, from the header contracts (H2Lib sources are not in /root/reference):
 *   build_bem3d_cluster -> build_adaptive_cluster   include/bem3d.h:1493, cluster.h:133-152,
 *                                                   clustergeometry.h:45-75
 *   build_strict_block / build_nonstrict_block      include/block.h:216-248
 *   admissible_2_cluster                            include/block.h:78-95
 *   decomp_partialaca_rkmatrix                      include/aca.h:79-117
 *   addeval_hmatrix_avector                         include/hmatrix.h:392-419
 * Call sites: cnld/compressed_formats.py:675-694.   "parity unpinned" like the rest of the
 * H2Lib-backed oracle (see cnld_oracle.c).  Flat arrays instead of pointer trees:
 *   clusters: son0/son1 (-1 for leaves), start/size into the permutation `perm`
 *             (perm[start..start+size) = original DOF indices), bmin/bmax[3]
 *   leaves  : (row cluster, col cluster, admissible)
 * ========================================================================== */
#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef double real;
typedef double _Complex field;
typedef unsigned int uint;
#define ORC_API __attribute__((visibility("default")))

void orc_nearfield_slp_ll(uint nv, uint nt, const real *x_, const uint *t_, const real *g, real kr,
                          real ki, uint q, uint q2, const uint *ridx, uint rows, const uint *cidx,
                          uint cols, field *N, uint ld);

/* ---- cluster tree ---------------------------------------------------------------------- */
typedef struct {
  int *son0, *son1;
  uint *start, *size;
  real *bmin, *bmax; /* [ncl][3] */
  uint ncl, cap;
} ctree;

static uint ct_new(ctree *T)
{
  if (T->ncl == T->cap) {
    T->cap = T->cap ? 2 * T->cap : 1024;
    T->son0 = realloc(T->son0, sizeof(int) * T->cap);
    T->son1 = realloc(T->son1, sizeof(int) * T->cap);
    T->start = realloc(T->start, sizeof(uint) * T->cap);
    T->size = realloc(T->size, sizeof(uint) * T->cap);
    T->bmin = realloc(T->bmin, sizeof(real) * 3 * T->cap);
    T->bmax = realloc(T->bmax, sizeof(real) * 3 * T->cap);
  }
  return T->ncl++;
}

/* build_adaptive_cluster: split the bounding box of the characteristic points along its
 * longest side at the midpoint while size > clf; leaves get the union of the support boxes,
 * inner nodes the union of their sons' boxes. */
static uint ct_build(ctree *T, const real *px, const real *smin, const real *smax, uint *perm,
                     uint start, uint size, uint clf)
{
  uint id = ct_new(T);
  uint *idx = perm + start;
  T->start[id] = start;
  T->size[id] = size;
  T->son0[id] = T->son1[id] = -1;
  int split = 0;
  if (size > clf) {
    real hmin[3], hmax[3];
    for (int d = 0; d < 3; d++) { hmin[d] = 1e300; hmax[d] = -1e300; }
    for (uint i = 0; i < size; i++)
      for (int d = 0; d < 3; d++) {
        real v = px[3 * idx[i] + d];
        if (v < hmin[d]) hmin[d] = v;
        if (v > hmax[d]) hmax[d] = v;
      }
    int dir = 0;
    real a = hmax[0] - hmin[0];
    for (int d = 1; d < 3; d++) {
      real m = hmax[d] - hmin[d];
      if (a < m) { a = m; dir = d; }
    }
    if (a > 0.0) {
      real m = (hmax[dir] + hmin[dir]) / 2.0;
      uint size0 = 0;
      for (uint i = 0; i < size; i++)
        if (px[3 * idx[i] + dir] < m) {
          uint j = idx[i]; idx[i] = idx[size0]; idx[size0] = j;
          size0++;
        }
      uint s0 = ct_build(T, px, smin, smax, perm, start, size0, clf);
      uint s1 = ct_build(T, px, smin, smax, perm, start + size0, size - size0, clf);
      T->son0[id] = (int)s0;
      T->son1[id] = (int)s1;
      for (int d = 0; d < 3; d++) {
        T->bmin[3 * id + d] = fmin(T->bmin[3 * s0 + d], T->bmin[3 * s1 + d]);
        T->bmax[3 * id + d] = fmax(T->bmax[3 * s0 + d], T->bmax[3 * s1 + d]);
      }
      split = 1;
    }
  }
  if (!split) {
    for (int d = 0; d < 3; d++) { T->bmin[3 * id + d] = 1e300; T->bmax[3 * id + d] = -1e300; }
    for (uint i = 0; i < size; i++)
      for (int d = 0; d < 3; d++) {
        if (smin[3 * idx[i] + d] < T->bmin[3 * id + d]) T->bmin[3 * id + d] = smin[3 * idx[i] + d];
        if (smax[3 * idx[i] + d] > T->bmax[3 * id + d]) T->bmax[3 * id + d] = smax[3 * idx[i] + d];
      }
  }
  return id;
}

/* build_bem3d_cluster for linear basis functions: DOF = vertex, characteristic point = vertex,
 * support box = bounding box of all triangles at the vertex.
 * Outputs are malloc'ed; free with orc_free. */
ORC_API uint orc_build_cluster(uint nv, uint nt, const real *x, const uint *t, uint clf, uint *perm,
                               int **son0, int **son1, uint **start, uint **size, real **bmin,
                               real **bmax)
{
  real *smin = malloc(sizeof(real) * 3 * nv), *smax = malloc(sizeof(real) * 3 * nv);
  for (uint i = 0; i < nv; i++)
    for (int d = 0; d < 3; d++) { smin[3 * i + d] = x[3 * i + d]; smax[3 * i + d] = x[3 * i + d]; }
  for (uint k = 0; k < nt; k++)
    for (int a = 0; a < 3; a++) {
      uint v = t[3 * k + a];
      for (int b = 0; b < 3; b++) {
        const real *p = x + 3 * t[3 * k + b];
        for (int d = 0; d < 3; d++) {
          if (p[d] < smin[3 * v + d]) smin[3 * v + d] = p[d];
          if (p[d] > smax[3 * v + d]) smax[3 * v + d] = p[d];
        }
      }
    }
  for (uint i = 0; i < nv; i++) perm[i] = i;
  ctree T;
  memset(&T, 0, sizeof(T));
  ct_build(&T, x, smin, smax, perm, 0, nv, clf);
  free(smin);
  free(smax);
  *son0 = T.son0; *son1 = T.son1; *start = T.start; *size = T.size; *bmin = T.bmin; *bmax = T.bmax;
  return T.ncl;
}

ORC_API void orc_free(void *p) { free(p); }

/* ---- block tree ------------------------------------------------------------------------ */
/* admissible_2_cluster: max(diam_2(rc), diam_2(cc)) <= eta * dist_2(rc, cc) on bounding boxes */
static int adm2(const real *bmin, const real *bmax, uint r, uint c, real eta)
{
  real dr = 0, dc = 0, dist = 0;
  for (int d = 0; d < 3; d++) {
    real a = bmax[3 * r + d] - bmin[3 * r + d], b = bmax[3 * c + d] - bmin[3 * c + d];
    dr += a * a;
    dc += b * b;
    real g = fmax(0.0, fmax(bmin[3 * r + d] - bmax[3 * c + d], bmin[3 * c + d] - bmax[3 * r + d]));
    dist += g * g;
  }
  real diam = sqrt(fmax(dr, dc));
  return diam <= eta * sqrt(dist);
}

typedef struct { uint *rc, *cc; int *adm; uint n, cap; } leaflist;
static void ll_push(leaflist *L, uint r, uint c, int a)
{
  if (L->n == L->cap) {
    L->cap = L->cap ? 2 * L->cap : 4096;
    L->rc = realloc(L->rc, sizeof(uint) * L->cap);
    L->cc = realloc(L->cc, sizeof(uint) * L->cap);
    L->adm = realloc(L->adm, sizeof(int) * L->cap);
  }
  L->rc[L->n] = r; L->cc[L->n] = c; L->adm[L->n] = a; L->n++;
}

static void bt_build(leaflist *L, const int *son0, const int *son1, const real *bmin, const real *bmax,
                     uint r, uint c, real eta, int strict)
{
  if (adm2(bmin, bmax, r, c, eta)) { ll_push(L, r, c, 1); return; }
  int rs = son0[r] >= 0, cs = son0[c] >= 0;
  if (rs && cs) {
    /* sons ordered like H2Lib: son[i + j*rsons] -> column-major over (row son, col son) */
    uint rr[2] = { (uint)son0[r], (uint)son1[r] }, ccs[2] = { (uint)son0[c], (uint)son1[c] };
    for (int j = 0; j < 2; j++)
      for (int i = 0; i < 2; i++) bt_build(L, son0, son1, bmin, bmax, rr[i], ccs[j], eta, strict);
  } else if (!strict && rs) {
    bt_build(L, son0, son1, bmin, bmax, (uint)son0[r], c, eta, strict);
    bt_build(L, son0, son1, bmin, bmax, (uint)son1[r], c, eta, strict);
  } else if (!strict && cs) {
    bt_build(L, son0, son1, bmin, bmax, r, (uint)son0[c], eta, strict);
    bt_build(L, son0, son1, bmin, bmax, r, (uint)son1[c], eta, strict);
  } else
    ll_push(L, r, c, 0);
}

ORC_API uint orc_build_block(const int *son0, const int *son1, const real *bmin, const real *bmax, real eta,
                             int strict, uint **rc, uint **cc, int **adm)
{
  leaflist L;
  memset(&L, 0, sizeof(L));
  bt_build(&L, son0, son1, bmin, bmax, 0, 0, eta, strict);
  *rc = L.rc; *cc = L.cc; *adm = L.adm;
  return L.n;
}

/* ---- partial-pivot ACA on an admissible block --------------------------------------------
 * A ~ sum_l u_l v_l^T with rows/columns of the residual generated on demand through the
 * entry callback (here the Galerkin entries of the SLP operator).  Start row 0; column pivot =
 * largest |residual row entry|; next row pivot = largest |residual column entry| among unused
 * rows; stop when |u_k||v_k| <= accur * |u_0||v_0| or the rank reaches min(rows, cols).
 * U is rows x kmax, V is cols x kmax (column-major); returns the rank. */
ORC_API uint orc_aca_block(uint nv, uint nt, const real *x, const uint *t, const real *g, real kr, real ki,
                           uint q, uint q2, const uint *ridx, uint rows, const uint *cidx, uint cols,
                           real accur, uint kmax, field *U, field *V)
{
  if (kmax > rows) kmax = rows;
  if (kmax > cols) kmax = cols;
  char *rused = calloc(rows, 1), *cused = calloc(cols, 1);
  field *row = malloc(sizeof(field) * cols), *col = malloc(sizeof(field) * rows);
  uint k = 0, ik = 0;
  real start = 0.0;
  while (k < kmax) {
    /* residual row ik */
    orc_nearfield_slp_ll(nv, nt, x, t, g, kr, ki, q, q2, ridx + ik, 1, cidx, cols, row, 1);
    for (uint l = 0; l < k; l++)
      for (uint j = 0; j < cols; j++) row[j] -= U[(size_t)l * rows + ik] * V[(size_t)l * cols + j];
    rused[ik] = 1;
    uint jk = 0;
    real best = -1.0;
    for (uint j = 0; j < cols; j++)
      if (!cused[j] && cabs(row[j]) > best) { best = cabs(row[j]); jk = j; }
    if (!(best > 0.0)) {
      /* zero residual row: try the next unused row */
      uint nx = rows;
      for (uint i = 0; i < rows; i++) if (!rused[i]) { nx = i; break; }
      if (nx == rows) break;
      ik = nx;
      continue;
    }
    cused[jk] = 1;
    field piv = 1.0 / row[jk];
    for (uint j = 0; j < cols; j++) V[(size_t)k * cols + j] = row[j] * piv;
    /* residual column jk */
    orc_nearfield_slp_ll(nv, nt, x, t, g, kr, ki, q, q2, ridx, rows, cidx + jk, 1, col, rows);
    for (uint l = 0; l < k; l++)
      for (uint i = 0; i < rows; i++) col[i] -= U[(size_t)l * rows + i] * V[(size_t)l * cols + jk];
    real nu = 0.0, nvv = 0.0;
    for (uint i = 0; i < rows; i++) { U[(size_t)k * rows + i] = col[i]; nu += creal(col[i]) * creal(col[i]) + cimag(col[i]) * cimag(col[i]); }
    for (uint j = 0; j < cols; j++) { field v = V[(size_t)k * cols + j]; nvv += creal(v) * creal(v) + cimag(v) * cimag(v); }
    real err = sqrt(nu) * sqrt(nvv);
    if (k == 0) start = err;
    k++;
    if (err <= accur * start) break;
    uint nx = rows;
    best = -1.0;
    for (uint i = 0; i < rows; i++)
      if (!rused[i] && cabs(col[i]) > best) { best = cabs(col[i]); nx = i; }
    if (nx == rows) break;
    ik = nx;
  }
  free(rused); free(cused); free(row); free(col);
  return k;
}
