/* ============================================================================
 * cnld_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * Plain-C restatement of the cnl-dyna frequency-response hot path as it runs
 * through H2Lib on the CPU.  H2Lib's .c sources are NOT in /root/reference
 * (only include/ headers; the library is cloned from github.com/H2Lib/H2Lib branch
 * "community" -- no commit pin -- at container build time,
 * containers/Singularity:38-44), so every function below restates the
 * *published contract* in the reference's headers plus the published
 * algorithm (Sauter/Schwab, "Boundary Element Methods", Sect. 5.2 for the
 * singular panel-pair rules) and is anchored on the reference's own call sites.
 *
 * PARITY STATUS: "parity unpinned" for the H2Lib part -- the reference holds
 * no golden vectors / known-answer tests for this path and
 * libh2 cannot be built here.  The oracle therefore carries its own
 * known-answer tests (tests/test_oracle_kat.py).  The Python-side pieces of
 * the path (FEM inputs, field pressure, impulse response) ARE pinned against
 * the reference's own code, see oracle/pyoracle.py and tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 * ========================================================================== */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double real;
typedef double _Complex field;
typedef unsigned int uint;

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------
 * Gauss-Legendre points/weights on [-1,1].
 * Contract: include/gaussquad.h:19-38 (assemble_gauss: zeros of the m-th
 * Legendre polynomial; exact for degree 2m-1).  H2Lib gets them from the
 * Jacobi-matrix eigenproblem; we use Newton on P_m, which converges to the
 * same zeros to machine precision.
 * ---------------------------------------------------------------------- */
ORC_API void orc_gauss(uint m, real *x, real *w)
{
  for (uint i = 0; i < m; i++) {
    real z = cos(M_PI * (i + 0.75) / (m + 0.5));
    real pp = 1.0;
    for (int it = 0; it < 100; it++) {
      real p1 = 1.0, p2 = 0.0;
      for (uint j = 1; j <= m; j++) {
        real p3 = p2;
        p2 = p1;
        p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
      }
      pp = m * (z * p1 - p2) / (z * z - 1.0);
      real dz = p1 / pp;
      z -= dz;
      if (fabs(dz) < 1e-16) break;
    }
    /* ascending order like an eigenvalue solver returns them */
    x[m - 1 - i] = z;
    w[m - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
  }
}

/* ------------------------------------------------------------------------
 * Macro-surface refinement.
 * Contract: include/macrosurface3d.h:48-101 (struct, phi callback on the
 * reference triangle {s,t>=0, s+t<=1}), :187-199
 * (build_from_macrosurface3d_surface3d: every macro edge split into `split`
 * parts, split*split plane triangles per macro triangle).
 * Parametrizations restate cnld/h2lib/macrosurface3d.pyx:83-145
 * (kind 0 = 'square'/'cube' = affine, kind 1 = 'circle').
 * Node ordering: macro vertices, then the split-1 inner vertices of each
 * macro edge (from e[.][0] to e[.][1]), then the inner vertices of each macro
 * triangle row by row.  (H2Lib's exact inner ordering is not visible from
 * the headers; any consistent ordering is a permutation of the DOFs.)
 * ---------------------------------------------------------------------- */
static void orc_phi(int kind, const real (*mx)[3], const uint *mt, real a, real b,
                    real xt[3])
{
  const real *p0 = mx[mt[0]], *p1 = mx[mt[1]], *p2 = mx[mt[2]];
  for (int d = 0; d < 3; d++)
    xt[d] = p0[d] * (1.0 - a - b) + p1[d] * a + p2[d] * b;
  if (kind == 1) { /* circle_parametrization, macrosurface3d.pyx:124-145 */
    real n1 = fabs(xt[0]) + fabs(xt[1]);
    real n2 = sqrt(xt[0] * xt[0] + xt[1] * xt[1]);
    if (n2 > 0.0) {
      xt[0] = xt[0] * n1 / n2;
      xt[1] = xt[1] * n1 / n2;
    }
  }
}

ORC_API void orc_refined_counts(uint mv, uint me, uint mt, uint split, uint *v,
                                uint *e, uint *t)
{
  *v = mv + me * (split - 1) + mt * ((split - 1) * (split - 2) / 2);
  *e = me * split + mt * (3 * split * (split - 1) / 2);
  *t = mt * split * split;
}

/* local grid index (i,j), i+j<=n, row-major by j */
static inline uint gidx(uint n, uint i, uint j) { return j * (n + 1) - j * (j - 1) / 2 + i; }

ORC_API void orc_refine_macrosurface(int kind, uint mv, uint me, uint mtn,
                                     const real *mx_, const uint *me_, const uint *mt_,
                                     const uint *ms_, uint split, real *x_, uint *e_,
                                     uint *t_, uint *s_)
{
  const real(*mx)[3] = (const real(*)[3])mx_;
  const uint(*med)[2] = (const uint(*)[2])me_;
  const uint(*mtr)[3] = (const uint(*)[3])mt_;
  const uint(*mse)[3] = (const uint(*)[3])ms_;
  real(*x)[3] = (real(*)[3])x_;
  uint(*e)[2] = (uint(*)[2])e_;
  uint(*t)[3] = (uint(*)[3])t_;
  uint(*s)[3] = (uint(*)[3])s_;
  const uint n = split;
  uint nv = mv + me * (n - 1);
  uint ne = me * n;
  uint nt = 0;
  uint ngrid = (n + 1) * (n + 2) / 2;
  uint *vloc = (uint *)malloc(sizeof(uint) * ngrid);
  /* local edge ids: family 0 horizontal (i,j)-(i+1,j); 1 vertical (i,j)-(i,j+1);
     2 diagonal (i+1,j)-(i,j+1); each indexed by gidx of (i,j) */
  uint *eloc = (uint *)malloc(sizeof(uint) * ngrid * 3);
  char *vdone = (char *)calloc(mv + me * (n - 1), 1);

  /* sub-edges on macro edges: vertex chain a=0, inner 1..n-1, b=n */
  for (uint E = 0; E < me; E++) {
    for (uint k = 0; k < n; k++) {
      uint va = (k == 0) ? med[E][0] : mv + E * (n - 1) + (k - 1);
      uint vb = (k == n - 1) ? med[E][1] : mv + E * (n - 1) + k;
      e[E * n + k][0] = va;
      e[E * n + k][1] = vb;
    }
  }

  for (uint T = 0; T < mtn; T++) {
    const uint *tv = mtr[T];
    /* boundary vertices / edges of the local grid from the macro topology.
       local corner 0=(0,0), 1=(n,0), 2=(0,n); macro edge s[T][c] lies opposite
       corner c. */
    for (uint j = 0; j <= n; j++)
      for (uint i = 0; i + j <= n; i++) vloc[gidx(n, i, j)] = (uint)-1;
    vloc[gidx(n, 0, 0)] = tv[0];
    vloc[gidx(n, n, 0)] = tv[1];
    vloc[gidx(n, 0, n)] = tv[2];
    for (uint q = 0; q < ngrid * 3; q++) eloc[q] = (uint)-1;
    for (uint c = 0; c < 3; c++) {
      uint E = mse[T][c];
      uint from = tv[(c + 1) % 3], to = tv[(c + 2) % 3];
      int fwd;
      if (med[E][0] == from && med[E][1] == to) fwd = 1;
      else if (med[E][0] == to && med[E][1] == from) fwd = 0;
      else { fprintf(stderr, "orc_refine: edge/triangle mismatch\n"); abort(); }
      for (uint k = 0; k <= n; k++) {
        /* k-th point walking from `from` to `to` */
        uint i, j;
        if (c == 0) { i = n - k; j = k; }       /* corner1 -> corner2 */
        else if (c == 1) { i = 0; j = n - k; }  /* corner2 -> corner0 */
        else { i = k; j = 0; }                  /* corner0 -> corner1 */
        uint kk = fwd ? k : n - k;              /* position along macro edge */
        if (k > 0 && k < n) vloc[gidx(n, i, j)] = mv + E * (n - 1) + (kk - 1);
        if (k < n) {
          uint ke = fwd ? k : n - 1 - k; /* sub-edge between point k and k+1 */
          uint id = E * n + ke;
          if (c == 0) eloc[gidx(n, n - k - 1, k) * 3 + 2] = id; /* diag (i+1,j)-(i,j+1) */
          else if (c == 1) eloc[gidx(n, 0, n - k - 1) * 3 + 1] = id;
          else eloc[gidx(n, k, 0) * 3 + 0] = id;
        }
      }
    }
    /* vertex coordinates: macro + edge vertices evaluated through phi of the
       first triangle that touches them, inner vertices are new */
    for (uint j = 0; j <= n; j++)
      for (uint i = 0; i + j <= n; i++) {
        uint g = gidx(n, i, j);
        if (vloc[g] == (uint)-1) vloc[g] = nv++;
        else if (vdone[vloc[g]]) continue;
        else vdone[vloc[g]] = 1;
        orc_phi(kind, mx, tv, (real)i / n, (real)j / n, x[vloc[g]]);
      }
    /* inner edges */
    for (uint j = 0; j < n; j++)
      for (uint i = 0; i + j < n; i++) {
        uint g = gidx(n, i, j);
        if (eloc[g * 3 + 0] == (uint)-1) {
          e[ne][0] = vloc[g]; e[ne][1] = vloc[gidx(n, i + 1, j)];
          eloc[g * 3 + 0] = ne++;
        }
        if (eloc[g * 3 + 1] == (uint)-1) {
          e[ne][0] = vloc[g]; e[ne][1] = vloc[gidx(n, i, j + 1)];
          eloc[g * 3 + 1] = ne++;
        }
        if (eloc[g * 3 + 2] == (uint)-1) {
          e[ne][0] = vloc[gidx(n, i + 1, j)]; e[ne][1] = vloc[gidx(n, i, j + 1)];
          eloc[g * 3 + 2] = ne++;
        }
      }
    /* triangles, counter-clockwise like the macro triangle */
    for (uint j = 0; j < n; j++)
      for (uint i = 0; i + j < n; i++) {
        uint g = gidx(n, i, j);
        /* upright: (i,j),(i+1,j),(i,j+1) */
        t[nt][0] = vloc[g];
        t[nt][1] = vloc[gidx(n, i + 1, j)];
        t[nt][2] = vloc[gidx(n, i, j + 1)];
        s[nt][0] = eloc[g * 3 + 2];
        s[nt][1] = eloc[g * 3 + 1];
        s[nt][2] = eloc[g * 3 + 0];
        nt++;
        if (i + j + 1 < n) {
          /* inverted: (i+1,j),(i+1,j+1),(i,j+1) */
          uint g1 = gidx(n, i + 1, j), g2 = gidx(n, i, j + 1);
          t[nt][0] = vloc[g1];
          t[nt][1] = vloc[gidx(n, i + 1, j + 1)];
          t[nt][2] = vloc[g2];
          s[nt][0] = eloc[g2 * 3 + 0]; /* (i,j+1)-(i+1,j+1), opposite (i+1,j) */
          s[nt][1] = eloc[g * 3 + 2];  /* (i+1,j)-(i,j+1), opposite (i+1,j+1) */
          s[nt][2] = eloc[g1 * 3 + 1]; /* (i+1,j)-(i+1,j+1), opposite (i,j+1) */
          nt++;
        }
      }
  }
  free(vloc);
  free(eloc);
  free(vdone);
}

/* ------------------------------------------------------------------------
 * prepare_surface3d: normals, Gram determinants g = 2*area, hmin, hmax.
 * Contract: include/surface3d.h:45-71,101.
 * ---------------------------------------------------------------------- */
ORC_API void orc_prepare_surface(uint nv, uint nt, const real *x_, const uint *t_,
                                 real *n_, real *g, real *hminmax)
{
  const real(*x)[3] = (const real(*)[3])x_;
  const uint(*t)[3] = (const uint(*)[3])t_;
  real(*nn)[3] = (real(*)[3])n_;
  real hmin = 1e300, hmax = 0.0;
  (void)nv;
  for (uint i = 0; i < nt; i++) {
    const real *A = x[t[i][0]], *B = x[t[i][1]], *C = x[t[i][2]];
    real u[3], v[3], w[3];
    for (int d = 0; d < 3; d++) { u[d] = B[d] - A[d]; v[d] = C[d] - A[d]; w[d] = C[d] - B[d]; }
    real c0 = u[1] * v[2] - u[2] * v[1];
    real c1 = u[2] * v[0] - u[0] * v[2];
    real c2 = u[0] * v[1] - u[1] * v[0];
    real nrm = sqrt(c0 * c0 + c1 * c1 + c2 * c2);
    g[i] = nrm;
    nn[i][0] = c0 / nrm; nn[i][1] = c1 / nrm; nn[i][2] = c2 / nrm;
    real l[3] = { sqrt(u[0]*u[0]+u[1]*u[1]+u[2]*u[2]), sqrt(v[0]*v[0]+v[1]*v[1]+v[2]*v[2]),
                  sqrt(w[0]*w[0]+w[1]*w[1]+w[2]*w[2]) };
    for (int d = 0; d < 3; d++) { if (l[d] < hmin) hmin = l[d]; if (l[d] > hmax) hmax = l[d]; }
  }
  hminmax[0] = hmin;
  hminmax[1] = hmax;
}

/* ------------------------------------------------------------------------
 * Panel-pair quadrature rules (singquad2d).
 * Contract: include/singquad2d.h:38-144 (four families id/edge/vert/dist with
 * n points each; x/y hold the two reference coordinates of each point,
 * stored as x[q], x[q+n]), :179 (build_singquad2d(gr,q,q2): q = Gauss order
 * for distant pairs, q2 for the singular families).
 * Reference triangle {0<=s<=t<=1}; a point (t,s) maps to
 *   A(1-t) + B(t-s) + C s               (linear hats 1-t, t-s, s).
 * Rules: Sauter/Schwab Sect. 5.2 relative coordinates + Duffy:
 *   identical   6 regions, weight xi^3 eta1^2 eta2
 *   common edge 5 regions, weight xi^3 eta1^2 (first) / xi^3 eta1^2 eta2
 *   common vert 2 regions, weight xi^3 eta2
 *   distant     tensor Duffy-Gauss, weight t_x * t_y
 * => n_id = 6 q2^4, n_edge = 5 q2^4, n_vert = 2 q2^4, n_dist = q^4.
 * The weights integrate over the product of two reference triangles (area
 * 1/2 each); entries are later scaled by g_t g_s = (2 area_t)(2 area_s).
 * ---------------------------------------------------------------------- */
ORC_API uint orc_singquad_count(int family, uint q, uint q2)
{
  uint p = q2 * q2 * q2 * q2;
  switch (family) {
  case 0: return q * q * q * q;
  case 1: return 2 * p;
  case 2: return 5 * p;
  default: return 6 * p;
  }
}

/* family: 0 dist, 1 vert, 2 edge, 3 id.  Outputs xt,xs (first triangle),
 * yt,ys (second triangle), w; each of length orc_singquad_count(). */
ORC_API void orc_singquad_rule(int family, uint q, uint q2, real *xt, real *xs, real *yt,
                               real *ys, real *w)
{
  uint m = (family == 0) ? q : q2;
  real *gx = (real *)malloc(sizeof(real) * m), *gw = (real *)malloc(sizeof(real) * m);
  orc_gauss(m, gx, gw);
  for (uint i = 0; i < m; i++) { gx[i] = 0.5 * (gx[i] + 1.0); gw[i] *= 0.5; }
  uint n = 0;
  for (uint i = 0; i < m; i++)
    for (uint j = 0; j < m; j++)
      for (uint k = 0; k < m; k++)
        for (uint l = 0; l < m; l++) {
          real wt = gw[i] * gw[j] * gw[k] * gw[l];
          if (family == 0) {
            /* Duffy on each triangle: t = a, s = a*b */
            xt[n] = gx[i]; xs[n] = gx[i] * gx[j];
            yt[n] = gx[k]; ys[n] = gx[k] * gx[l];
            w[n] = wt * gx[i] * gx[k];
            n++;
            continue;
          }
          real xi = gx[i], e1 = gx[j], e2 = gx[k], e3 = gx[l];
          if (family == 1) {
            real ww = wt * xi * xi * xi * e2;
            xt[n] = xi; xs[n] = xi * e1; yt[n] = xi * e2; ys[n] = xi * e2 * e3; w[n++] = ww;
            xt[n] = xi * e2; xs[n] = xi * e2 * e3; yt[n] = xi; ys[n] = xi * e1; w[n++] = ww;
          } else if (family == 2) {
            real w1 = wt * xi * xi * xi * e1 * e1;
            real w2 = w1 * e2;
            xt[n] = xi; xs[n] = xi * e1 * e3;
            yt[n] = xi * (1.0 - e1 * e2); ys[n] = xi * e1 * (1.0 - e2); w[n++] = w1;
            xt[n] = xi; xs[n] = xi * e1;
            yt[n] = xi * (1.0 - e1 * e2 * e3); ys[n] = xi * e1 * e2 * (1.0 - e3); w[n++] = w2;
            xt[n] = xi * (1.0 - e1 * e2); xs[n] = xi * e1 * (1.0 - e2);
            yt[n] = xi; ys[n] = xi * e1 * e2 * e3; w[n++] = w2;
            xt[n] = xi * (1.0 - e1 * e2 * e3); xs[n] = xi * e1 * e2 * (1.0 - e3);
            yt[n] = xi; ys[n] = xi * e1; w[n++] = w2;
            xt[n] = xi * (1.0 - e1 * e2 * e3); xs[n] = xi * e1 * (1.0 - e2 * e3);
            yt[n] = xi; ys[n] = xi * e1 * e2; w[n++] = w2;
          } else {
            real ww = wt * xi * xi * xi * e1 * e1 * e2;
            real a1 = xi, a2 = xi * (1.0 - e1 + e1 * e2);
            real b1 = xi * (1.0 - e1 * e2 * e3), b2 = xi * (1.0 - e1);
            xt[n] = a1; xs[n] = a2; yt[n] = b1; ys[n] = b2; w[n++] = ww;
            xt[n] = b1; xs[n] = b2; yt[n] = a1; ys[n] = a2; w[n++] = ww;
            a1 = xi; a2 = xi * e1 * (1.0 - e2 + e2 * e3);
            b1 = xi * (1.0 - e1 * e2); b2 = xi * e1 * (1.0 - e2);
            xt[n] = a1; xs[n] = a2; yt[n] = b1; ys[n] = b2; w[n++] = ww;
            xt[n] = b1; xs[n] = b2; yt[n] = a1; ys[n] = a2; w[n++] = ww;
            a1 = xi * (1.0 - e1 * e2 * e3); a2 = xi * e1 * (1.0 - e2 * e3);
            b1 = xi; b2 = xi * e1 * (1.0 - e2);
            xt[n] = a1; xs[n] = a2; yt[n] = b1; ys[n] = b2; w[n++] = ww;
            xt[n] = b1; xs[n] = b2; yt[n] = a1; ys[n] = a2; w[n++] = ww;
          }
        }
  free(gx);
  free(gw);
}

/* ------------------------------------------------------------------------
 * select_quadrature_singquad2d: count common vertices of (t,s) and return
 * vertex permutations tp,sp that move the shared vertices to the front
 * (same shared vertex at the same position in both).
 * Contract: include/singquad2d.h:283-303.
 * ---------------------------------------------------------------------- */
static uint orc_select(const uint *tv, const uint *sv, uint *tp, uint *sp)
{
  uint p = 0;
  for (uint i = 0; i < 3; i++)
    for (uint j = 0; j < 3; j++)
      if (tv[i] == sv[j]) { tp[p] = i; sp[p] = j; p++; break; }
  uint q = p;
  for (uint i = 0; i < 3; i++) {
    uint j;
    for (j = 0; j < q && tp[j] != i; j++) ;
    if (j == q) tp[q++] = i;
  }
  q = p;
  for (uint i = 0; i < 3; i++) {
    uint j;
    for (j = 0; j < q && sp[j] != i; j++) ;
    if (j == q) sp[q++] = i;
  }
  return p;
}

/* Helmholtz single-layer kernel e^{ikr}/r for complex k = kr + i*ki:
 * cos(kr r)/r + i sin(kr r)/r, times exp(-ki r).  The 1/(4 pi) is applied
 * once per entry (KERNEL_CONST_HELMHOLTZBEM3D, include/helmholtzbem3d.h:35).
 * Form per scratch/demo_zmatrix_hmatrix.c:97-121. */
static inline field orc_kernel(real dx, real dy, real dz, real kr, real ki)
{
  real r2 = dx * dx + dy * dy + dz * dz;
  real rinv = 1.0 / sqrt(r2);
  real r = r2 * rinv;
  if (ki != 0.0) rinv *= exp(-ki * r);
  real ph = kr * r;
  return rinv * cos(ph) + I * (rinv * sin(ph));
}

typedef struct {
  uint n[4];
  real *xt[4], *xs[4], *yt[4], *ys[4], *w[4];
} orc_rules;

static void orc_rules_init(orc_rules *R, uint q, uint q2)
{
  for (int f = 0; f < 4; f++) {
    uint n = orc_singquad_count(f, q, q2);
    R->n[f] = n;
    R->xt[f] = (real *)malloc(sizeof(real) * n * 5);
    R->xs[f] = R->xt[f] + n; R->yt[f] = R->xs[f] + n; R->ys[f] = R->yt[f] + n;
    R->w[f] = R->ys[f] + n;
    orc_singquad_rule(f, q, q2, R->xt[f], R->xs[f], R->yt[f], R->ys[f], R->w[f]);
  }
}
static void orc_rules_free(orc_rules *R) { for (int f = 0; f < 4; f++) free(R->xt[f]); }

/* 3x3 local Galerkin block of the pair (t,s), linear/linear:
 *   loc[i][j] = g_t g_s /(4 pi) * sum_q w_q phi_i(x_q) phi_j(y_q) G(x_q,y_q)
 * indexed by the ORIGINAL local vertex numbers of t and s.
 * Contract: include/bem3d.h:359-406 (N_ij = int int phi_i g psi_j), :1728-1735
 * (assemble_ll_near_bem3d: rule picked by number of shared vertices). */
static void orc_pair_ll(const orc_rules *R, const real (*x)[3], const uint *tv,
                        const uint *sv, real gt, real gs, real kr, real ki, field loc[3][3])
{
  uint tp[3], sp[3];
  uint c = orc_select(tv, sv, tp, sp);
  const real *A = x[tv[tp[0]]], *B = x[tv[tp[1]]], *C = x[tv[tp[2]]];
  const real *D = x[sv[sp[0]]], *E = x[sv[sp[1]]], *F = x[sv[sp[2]]];
  field acc[3][3];
  memset(acc, 0, sizeof(acc));
  uint n = R->n[c];
  const real *xt = R->xt[c], *xs = R->xs[c], *yt = R->yt[c], *ys = R->ys[c], *w = R->w[c];
  for (uint q = 0; q < n; q++) {
    real a[3] = { 1.0 - xt[q], xt[q] - xs[q], xs[q] };
    real b[3] = { 1.0 - yt[q], yt[q] - ys[q], ys[q] };
    real dx = A[0] * a[0] + B[0] * a[1] + C[0] * a[2] - (D[0] * b[0] + E[0] * b[1] + F[0] * b[2]);
    real dy = A[1] * a[0] + B[1] * a[1] + C[1] * a[2] - (D[1] * b[0] + E[1] * b[1] + F[1] * b[2]);
    real dz = A[2] * a[0] + B[2] * a[1] + C[2] * a[2] - (D[2] * b[0] + E[2] * b[1] + F[2] * b[2]);
    field kv = w[q] * orc_kernel(dx, dy, dz, kr, ki);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) acc[i][j] += (a[i] * b[j]) * kv;
  }
  real sc = gt * gs * 0.0795774715459476678844418816863; /* 1/(4 pi) */
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) loc[tp[i]][sp[j]] = acc[i][j] * sc;
}

/* ------------------------------------------------------------------------
 * assemble_bem3d_amatrix for the SLP Helmholtz operator, linear/linear.
 * Contract: include/bem3d.h:3020-3029 (dense N x N Galerkin matrix, N = number
 * of vertices), column-major with leading dimension ld (include/amatrix.h:42-48).
 * Call site: cnld/compressed_formats.py:629-634.
 * ---------------------------------------------------------------------- */
ORC_API void orc_assemble_slp_ll(uint nv, uint nt, const real *x_, const uint *t_,
                                 const real *g, real kr, real ki, uint q, uint q2,
                                 field *Z, uint ld)
{
  const real(*x)[3] = (const real(*)[3])x_;
  const uint(*t)[3] = (const uint(*)[3])t_;
  orc_rules R;
  orc_rules_init(&R, q, q2);
  for (size_t c = 0; c < nv; c++) memset(Z + c * ld, 0, sizeof(field) * nv);
  /* vertex -> triangle adjacency (bem->v2t) so each thread owns whole rows */
  uint *cnt = (uint *)calloc(nv + 1, sizeof(uint));
  for (uint i = 0; i < nt; i++) for (int d = 0; d < 3; d++) cnt[t[i][d] + 1]++;
  for (uint i = 0; i < nv; i++) cnt[i + 1] += cnt[i];
  uint *adj = (uint *)malloc(sizeof(uint) * 3 * nt);
  uint *pos = (uint *)malloc(sizeof(uint) * nv);
  memcpy(pos, cnt, sizeof(uint) * nv);
  for (uint i = 0; i < nt; i++) for (int d = 0; d < 3; d++) adj[pos[t[i][d]]++] = i * 3 + d;
  free(pos);
  /* pair-parallel over row triangles; rows of Z touched by different row
     triangles overlap, so accumulate per row triangle into a private strip
     and reduce by vertex afterwards: strip[tt][li][col] */
  (void)adj; (void)cnt;
#pragma omp parallel
  {
    field *strip = (field *)malloc(sizeof(field) * 3 * (size_t)nv);
#pragma omp for schedule(dynamic, 4)
    for (uint tt = 0; tt < nt; tt++) {
      memset(strip, 0, sizeof(field) * 3 * (size_t)nv);
      for (uint ss = 0; ss < nt; ss++) {
        field loc[3][3];
        orc_pair_ll(&R, x, t[tt], t[ss], g[tt], g[ss], kr, ki, loc);
        for (int i = 0; i < 3; i++)
          for (int j = 0; j < 3; j++) strip[(size_t)i * nv + t[ss][j]] += loc[i][j];
      }
      for (int i = 0; i < 3; i++) {
        uint row = t[tt][i];
        for (uint col = 0; col < nv; col++) {
          field v = strip[(size_t)i * nv + col];
          real *dst = (real *)(Z + (size_t)col * ld + row);
#pragma omp atomic
          dst[0] += creal(v);
#pragma omp atomic
          dst[1] += cimag(v);
        }
      }
    }
    free(strip);
  }
  free(adj);
  free(cnt);
  orc_rules_free(&R);
}

/* Sub-block variant = bem->nearfield(ridx, cidx, ...) (include/bem3d.h:370-385):
 * fills N[rows x cols] for the given vertex index sets.  Used by the ACA and
 * H-matrix parts and as an independent cross-check of the dense assembly. */
ORC_API void orc_nearfield_slp_ll(uint nv, uint nt, const real *x_, const uint *t_,
                                  const real *g, real kr, real ki, uint q, uint q2,
                                  const uint *ridx, uint rows, const uint *cidx, uint cols,
                                  field *N, uint ld)
{
  const real(*x)[3] = (const real(*)[3])x_;
  const uint(*t)[3] = (const uint(*)[3])t_;
  orc_rules R;
  orc_rules_init(&R, q, q2);
  int *rmap = (int *)malloc(sizeof(int) * nv), *cmap = (int *)malloc(sizeof(int) * nv);
  for (uint i = 0; i < nv; i++) rmap[i] = cmap[i] = -1;
  for (uint i = 0; i < rows; i++) rmap[ridx[i]] = (int)i;
  for (uint i = 0; i < cols; i++) cmap[cidx[i]] = (int)i;
  for (uint c = 0; c < cols; c++) memset(N + (size_t)c * ld, 0, sizeof(field) * rows);
  for (uint tt = 0; tt < nt; tt++) {
    if (rmap[t[tt][0]] < 0 && rmap[t[tt][1]] < 0 && rmap[t[tt][2]] < 0) continue;
    for (uint ss = 0; ss < nt; ss++) {
      if (cmap[t[ss][0]] < 0 && cmap[t[ss][1]] < 0 && cmap[t[ss][2]] < 0) continue;
      field loc[3][3];
      orc_pair_ll(&R, x, t[tt], t[ss], g[tt], g[ss], kr, ki, loc);
      for (int i = 0; i < 3; i++) {
        int r = rmap[t[tt][i]];
        if (r < 0) continue;
        for (int j = 0; j < 3; j++) {
          int c = cmap[t[ss][j]];
          if (c >= 0) N[(size_t)c * ld + r] += loc[i][j];
        }
      }
    }
  }
  free(rmap);
  free(cmap);
  orc_rules_free(&R);
}

/* ------------------------------------------------------------------------
 * lrdecomp_amatrix / lrsolve_amatrix_avector: UNPIVOTED LU, L unit-lower,
 * in place, column-major.  Returns 0 on success, k+1 if pivot k is zero.
 * Contract: include/factorizations.h:173-182, :233.
 * Call sites: cnld/compressed_formats.py:247-252, :259-262.
 * ---------------------------------------------------------------------- */
ORC_API uint orc_lrdecomp(uint n, field *a, uint ld)
{
  for (uint k = 0; k < n; k++) {
    field piv = a[(size_t)k * ld + k];
    if (piv == 0.0) return k + 1;
    field alpha = 1.0 / piv;
    for (uint i = k + 1; i < n; i++) a[(size_t)k * ld + i] *= alpha;
#pragma omp parallel for schedule(static) if (n - k > 256)
    for (uint j = k + 1; j < n; j++) {
      field u = a[(size_t)j * ld + k];
      const field *l = a + (size_t)k * ld;
      field *c = a + (size_t)j * ld;
      for (uint i = k + 1; i < n; i++) c[i] -= l[i] * u;
    }
  }
  return 0;
}

ORC_API void orc_lrsolve(uint n, const field *a, uint ld, field *x)
{
  for (uint j = 0; j < n; j++) { /* L y = b, unit diagonal */
    field xj = x[j];
    for (uint i = j + 1; i < n; i++) x[i] -= a[(size_t)j * ld + i] * xj;
  }
  for (uint jj = n; jj-- > 0;) { /* R x = y */
    x[jj] /= a[(size_t)jj * ld + jj];
    field xj = x[jj];
    for (uint i = 0; i < jj; i++) x[i] -= a[(size_t)jj * ld + i] * xj;
  }
}

/* ------------------------------------------------------------------------
 * Field pressure: restates cnld/bem_pressure.pyx:54-88 (mesh_pres_vector)
 * with gauss_quadrature(2) (:25-27: points (1/6,1/6),(2/3,1/6),(1/6,2/3),
 * weights 1/3), kernel e^{-jkr}/(4 pi r) (:41-49), source z forced to 0 (:81),
 * scale -(k c)^2 * rho * 2 (:88).  One field point per call row; nobs points.
 * p_vec is [nobs][nv].
 * ---------------------------------------------------------------------- */
ORC_API void orc_pres_vector(uint nv, uint nt, const real *x_, const uint *t_,
                             const real *area, const real *obs, uint nobs, real k, real c,
                             real rho, field *pvec)
{
  const real(*x)[3] = (const real(*)[3])x_;
  const uint(*t)[3] = (const uint(*)[3])t_;
  static const real gxi[3] = { 1.0 / 6, 2.0 / 3, 1.0 / 6 }, get[3] = { 1.0 / 6, 1.0 / 6, 2.0 / 3 };
  const real gw = 1.0 / 3;
#pragma omp parallel for schedule(static)
  for (uint o = 0; o < nobs; o++) {
    field *p = pvec + (size_t)o * nv;
    memset(p, 0, sizeof(field) * nv);
    real ox = obs[3 * o], oy = obs[3 * o + 1], oz = obs[3 * o + 2];
    for (uint tt = 0; tt < nt; tt++) {
      const real *P1 = x[t[tt][0]], *P2 = x[t[tt][1]], *P3 = x[t[tt][2]];
      for (int q = 0; q < 3; q++) {
        real xi = gxi[q], eta = get[q];
        real xs = P1[0] * (1 - xi - eta) + P2[0] * xi + P3[0] * eta;
        real ys = P1[1] * (1 - xi - eta) + P2[1] * xi + P3[1] * eta;
        real r = sqrt((ox - xs) * (ox - xs) + (oy - ys) * (oy - ys) + oz * oz);
        field cf = gw * (cexp(-I * k * r) / (4 * M_PI * r)) * area[tt];
        p[t[tt][0]] += (1 - xi - eta) * cf;
        p[t[tt][1]] += xi * cf;
        p[t[tt][2]] += eta * cf;
      }
    }
    real sc = -(k * c) * (k * c) * rho * 2;
    for (uint i = 0; i < nv; i++) p[i] *= sc;
  }
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------
 * Timing helper for large arrays (bench.py --config c4, CPU baseline): assemble a LIST of H-matrix
 * leaves -- inadmissible: the dense Galerkin block, admissible: partial-pivot ACA exactly as
 * orc_aca_block -- and apply them to one vector, with vertex -> triangle adjacency built once
 * (orc_nearfield_slp_ll scans all T^2 triangle pairs per call, which is fine for the parity cases
 * and useless at T = 3.7e5), OpenMP over the leaves.  Same entries, same ACA, same stopping rule.
 * leaf l: rows ridx[rptr[l] .. rptr[l+1]), columns cidx[cptr[l] .. cptr[l+1]), adm[l].
 * out[0] = assembly seconds, out[1] = matvec seconds, out[2] = stored entries, out[3] = largest rank;
 * y (nv) += sum over the listed leaves of leaf * x.
 * ---------------------------------------------------------------------- */
#include <omp.h>
typedef struct { uint *tl; uint n; } orc_trilist;
static void orc_collect_tris(const uint *idx, uint cnt, const uint *vptr, const uint *vadj, int *stamp, int mark, orc_trilist *out)
{
  out->n = 0;
  for (uint i = 0; i < cnt; i++)
    for (uint e = vptr[idx[i]]; e < vptr[idx[i] + 1]; e++)
      if (stamp[vadj[e]] != mark) { stamp[vadj[e]] = mark; out->tl[out->n++] = vadj[e]; }
}
ORC_API void orc_hleaves_sample(uint nv, uint nt, const real *x_, const uint *t_, const real *g, real kr, real ki,
                                uint q, uint q2, uint nleaf, const uint *rptr, const uint *ridx, const uint *cptr,
                                const uint *cidx, const int *adm, real accur, uint kmax, const field *xin, field *y,
                                real *out)
{
  const real(*x)[3] = (const real(*)[3])x_;
  const uint(*t)[3] = (const uint(*)[3])t_;
  uint *vptr = calloc(nv + 1, sizeof(uint)), *vadj = malloc(sizeof(uint) * 3 * (size_t)nt), *fill = malloc(sizeof(uint) * nv);
  for (uint tt = 0; tt < nt; tt++) for (int d = 0; d < 3; d++) vptr[t[tt][d] + 1]++;
  for (uint v = 0; v < nv; v++) vptr[v + 1] += vptr[v];
  memcpy(fill, vptr, sizeof(uint) * nv);
  for (uint tt = 0; tt < nt; tt++) for (int d = 0; d < 3; d++) vadj[fill[t[tt][d]]++] = tt;
  free(fill);
  field **data = calloc(nleaf, sizeof(field *));
  uint *rank = calloc(nleaf, sizeof(uint));
  double stored = 0.0, maxrank = 0.0;
  double t0 = omp_get_wtime();
#pragma omp parallel
  {
    orc_rules R;
    orc_rules_init(&R, q, q2);
    int *rmap = malloc(sizeof(int) * nv), *cmap = malloc(sizeof(int) * nv), *stamp = malloc(sizeof(int) * nt);
    for (uint i = 0; i < nv; i++) rmap[i] = cmap[i] = -1;
    for (uint i = 0; i < nt; i++) stamp[i] = -1;
    orc_trilist rt = {malloc(sizeof(uint) * nt), 0}, ct = {malloc(sizeof(uint) * nt), 0};
    int mark = 0;
#pragma omp for schedule(dynamic, 4) reduction(+ : stored) reduction(max : maxrank)
    for (uint l = 0; l < nleaf; l++) {
      const uint *ri = ridx + rptr[l], *ci = cidx + cptr[l];
      const uint m = rptr[l + 1] - rptr[l], n = cptr[l + 1] - cptr[l];
      for (uint i = 0; i < m; i++) rmap[ri[i]] = (int)i;
      for (uint j = 0; j < n; j++) cmap[ci[j]] = (int)j;
      orc_collect_tris(ri, m, vptr, vadj, stamp, mark++, &rt);
      orc_collect_tris(ci, n, vptr, vadj, stamp, mark++, &ct);
      if (!adm[l]) {
        field *N = calloc((size_t)m * n, sizeof(field));
        for (uint a = 0; a < rt.n; a++)
          for (uint b = 0; b < ct.n; b++) {
            const uint tt = rt.tl[a], ss = ct.tl[b];
            field loc[3][3];
            orc_pair_ll(&R, x, t[tt], t[ss], g[tt], g[ss], kr, ki, loc);
            for (int i = 0; i < 3; i++) {
              const int r = rmap[t[tt][i]];
              if (r < 0) continue;
              for (int j = 0; j < 3; j++) { const int c = cmap[t[ss][j]]; if (c >= 0) N[(size_t)c * m + r] += loc[i][j]; }
            }
          }
        data[l] = N;
        stored += (double)m * n;
      } else {
        uint kc = kmax < m ? kmax : m;
        if (kc > n) kc = n;
        field *U = calloc((size_t)(m + n) * kc, sizeof(field)), *V = U + (size_t)m * kc;
        field *row = malloc(sizeof(field) * n), *col = malloc(sizeof(field) * m);
        char *rused = calloc(m, 1), *cused = calloc(n, 1);
        uint k = 0, ik = 0;
        real start = 0.0;
        while (k < kc) {
          memset(row, 0, sizeof(field) * n);            /* residual row ik: triangles of that vertex x column triangles */
          for (uint e = vptr[ri[ik]]; e < vptr[ri[ik] + 1]; e++) {
            const uint tt = vadj[e];
            int li = 0;
            for (int d = 0; d < 3; d++) if (t[tt][d] == ri[ik]) li = d;
            for (uint b = 0; b < ct.n; b++) {
              const uint ss = ct.tl[b];
              field loc[3][3];
              orc_pair_ll(&R, x, t[tt], t[ss], g[tt], g[ss], kr, ki, loc);
              for (int j = 0; j < 3; j++) { const int c = cmap[t[ss][j]]; if (c >= 0) row[c] += loc[li][j]; }
            }
          }
          for (uint p = 0; p < k; p++) for (uint j = 0; j < n; j++) row[j] -= U[(size_t)p * m + ik] * V[(size_t)p * n + j];
          rused[ik] = 1;
          uint jk = 0;
          real best = -1.0;
          for (uint j = 0; j < n; j++) if (!cused[j] && cabs(row[j]) > best) { best = cabs(row[j]); jk = j; }
          if (!(best > 0.0)) {
            uint nx = m;
            for (uint i = 0; i < m; i++) if (!rused[i]) { nx = i; break; }
            if (nx == m) break;
            ik = nx;
            continue;
          }
          cused[jk] = 1;
          const field piv = 1.0 / row[jk];
          for (uint j = 0; j < n; j++) V[(size_t)k * n + j] = row[j] * piv;
          memset(col, 0, sizeof(field) * m);             /* residual column jk */
          for (uint e = vptr[ci[jk]]; e < vptr[ci[jk] + 1]; e++) {
            const uint ss = vadj[e];
            int lj = 0;
            for (int d = 0; d < 3; d++) if (t[ss][d] == ci[jk]) lj = d;
            for (uint a = 0; a < rt.n; a++) {
              const uint tt = rt.tl[a];
              field loc[3][3];
              orc_pair_ll(&R, x, t[tt], t[ss], g[tt], g[ss], kr, ki, loc);
              for (int i = 0; i < 3; i++) { const int r = rmap[t[tt][i]]; if (r >= 0) col[r] += loc[i][lj]; }
            }
          }
          for (uint p = 0; p < k; p++) for (uint i = 0; i < m; i++) col[i] -= U[(size_t)p * m + i] * V[(size_t)p * n + jk];
          real nu = 0.0, nvv = 0.0;
          for (uint i = 0; i < m; i++) { U[(size_t)k * m + i] = col[i]; nu += creal(col[i]) * creal(col[i]) + cimag(col[i]) * cimag(col[i]); }
          for (uint j = 0; j < n; j++) { const field v = V[(size_t)k * n + j]; nvv += creal(v) * creal(v) + cimag(v) * cimag(v); }
          const real err = sqrt(nu) * sqrt(nvv);
          if (k == 0) start = err;
          k++;
          if (err <= accur * start) break;
          uint nx = m;
          best = -1.0;
          for (uint i = 0; i < m; i++) if (!rused[i] && cabs(col[i]) > best) { best = cabs(col[i]); nx = i; }
          if (nx == m) break;
          ik = nx;
        }
        free(row); free(col); free(rused); free(cused);
        data[l] = U;
        rank[l] = k;
        stored += (double)k * (m + n);
        if ((double)k > maxrank) maxrank = (double)k;
      }
      for (uint i = 0; i < m; i++) rmap[ri[i]] = -1;
      for (uint j = 0; j < n; j++) cmap[ci[j]] = -1;
    }
    free(rmap); free(cmap); free(stamp); free(rt.tl); free(ct.tl);
    orc_rules_free(&R);
  }
  double t1 = omp_get_wtime();
  /* matvec over the sampled leaves: disjoint row sets are not guaranteed, so accumulate per leaf under a lock-free
     per-thread buffer would need nv * threads; leaves are few here: serialise the scatter, parallelise the products */
  for (uint l = 0; l < nleaf; l++) {
    const uint *ri = ridx + rptr[l], *ci = cidx + cptr[l];
    const uint m = rptr[l + 1] - rptr[l], n = cptr[l + 1] - cptr[l];
    if (!adm[l]) {
      const field *N = data[l];
      for (uint j = 0; j < n; j++) { const field xj = xin[ci[j]]; for (uint i = 0; i < m; i++) y[ri[i]] += N[(size_t)j * m + i] * xj; }
    } else {
      uint kc = kmax < m ? kmax : m;
      if (kc > n) kc = n;
      const field *U = data[l], *V = U + (size_t)m * kc;
      for (uint p = 0; p < rank[l]; p++) {
        field s = 0.0;
        for (uint j = 0; j < n; j++) s += V[(size_t)p * n + j] * xin[ci[j]];
        for (uint i = 0; i < m; i++) y[ri[i]] += U[(size_t)p * m + i] * s;
      }
    }
  }
  double t2 = omp_get_wtime();
  for (uint l = 0; l < nleaf; l++) free(data[l]);
  free(data); free(rank); free(vptr); free(vadj);
  out[0] = t1 - t0; out[1] = t2 - t1; out[2] = stored; out[3] = maxrank;
}
