// Host-side, frequency-independent setup of the panel-pair quadrature:
// Gauss-Legendre nodes, the four rule families of H2Lib's singquad2d
// (include/singquad2d.h:38-144: distant / common vertex / common edge / identical;
// build_singquad2d(gr, q, q2), :179) and the list of touching triangle pairs with
// their vertex permutations (select_quadrature_singquad2d, :283-303).
//
// Reference triangle {0 <= s <= t <= 1}; (t, s) -> A(1-t) + B(t-s) + C s.
// Singular families follow Sauter & Schwab, Boundary Element Methods, Sect. 5.2
// (relative coordinates + Duffy); point counts 2/5/6 * q2^4 and q^4.
#include "quadrature.h"

#include <cmath>

namespace cnld {

void gauss_legendre(unsigned m, std::vector<double> &x, std::vector<double> &w) {
  // zeros of P_m by Newton iteration from the Chebyshev-like guess (include/gaussquad.h:19-38
  // defines the rule; H2Lib solves the Jacobi eigenproblem, same nodes)
  x.assign(m, 0.0);
  w.assign(m, 0.0);
  const double pi = 3.14159265358979323846;
  for (unsigned r = 0; r < (m + 1) / 2; r++) {
    double z = std::cos(pi * (r + 0.75) / (m + 0.5)), dp = 1.0;
    for (int it = 0; it < 64; it++) {
      double pa = 1.0, pb = 0.0;
      for (unsigned j = 1; j <= m; j++) {
        double pc = pb;
        pb = pa;
        pa = ((2.0 * j - 1.0) * z * pb - (j - 1.0) * pc) / j;
      }
      dp = m * (z * pa - pb) / (z * z - 1.0);
      double step = pa / dp;
      z -= step;
      if (std::fabs(step) < 1e-16) break;
    }
    double wt = 2.0 / ((1.0 - z * z) * dp * dp);
    x[m - 1 - r] = z;  w[m - 1 - r] = wt;
    x[r] = -z;         w[r] = wt;
  }
  if (m & 1) x[m / 2] = 0.0;
}

namespace {
// One Sauter-Schwab region: both reference points as functions of (xi, e1, e2, e3)
// and the Jacobian's extra factor beyond xi^3.
struct Pt4 { double xt, xs, yt, ys; };
typedef Pt4 (*RegionMap)(double, double, double, double);

// identical panels: three maps, each used with (x,y) and swapped -> 6 regions; jac xi^3 e1^2 e2
Pt4 id_a(double xi, double e1, double e2, double e3) {
  return {xi, xi * (1 - e1 + e1 * e2), xi * (1 - e1 * e2 * e3), xi * (1 - e1)};
}
Pt4 id_b(double xi, double e1, double e2, double e3) {
  return {xi, xi * e1 * (1 - e2 + e2 * e3), xi * (1 - e1 * e2), xi * e1 * (1 - e2)};
}
Pt4 id_c(double xi, double e1, double e2, double e3) {
  return {xi * (1 - e1 * e2 * e3), xi * e1 * (1 - e2 * e3), xi, xi * e1 * (1 - e2)};
}
// common edge: 5 regions; first has jac xi^3 e1^2, the others xi^3 e1^2 e2
Pt4 ed_0(double xi, double e1, double e2, double e3) {
  return {xi, xi * e1 * e3, xi * (1 - e1 * e2), xi * e1 * (1 - e2)};
}
Pt4 ed_1(double xi, double e1, double e2, double e3) {
  return {xi, xi * e1, xi * (1 - e1 * e2 * e3), xi * e1 * e2 * (1 - e3)};
}
Pt4 ed_2(double xi, double e1, double e2, double e3) {
  return {xi * (1 - e1 * e2), xi * e1 * (1 - e2), xi, xi * e1 * e2 * e3};
}
Pt4 ed_3(double xi, double e1, double e2, double e3) {
  return {xi * (1 - e1 * e2 * e3), xi * e1 * e2 * (1 - e3), xi, xi * e1};
}
Pt4 ed_4(double xi, double e1, double e2, double e3) {
  return {xi * (1 - e1 * e2 * e3), xi * e1 * (1 - e2 * e3), xi, xi * e1 * e2};
}
// common vertex: one map and its swap; jac xi^3 e2
Pt4 vt_a(double xi, double e1, double e2, double e3) { return {xi, xi * e1, xi * e2, xi * e2 * e3}; }
Pt4 swap(Pt4 p) { return {p.yt, p.ys, p.xt, p.xs}; }
}  // namespace

void build_pair_rule(int family, unsigned q, unsigned q2, PairRule &R) {
  const unsigned m = (family == FAM_DIST) ? q : q2;
  std::vector<double> gx, gw;
  gauss_legendre(m, gx, gw);
  for (unsigned i = 0; i < m; i++) { gx[i] = 0.5 * (gx[i] + 1.0); gw[i] *= 0.5; }
  R.xt.clear(); R.xs.clear(); R.yt.clear(); R.ys.clear(); R.w.clear();
  auto push = [&](Pt4 p, double wt) {
    R.xt.push_back(p.xt); R.xs.push_back(p.xs); R.yt.push_back(p.yt); R.ys.push_back(p.ys);
    R.w.push_back(wt);
  };
  for (unsigned i = 0; i < m; i++)
    for (unsigned j = 0; j < m; j++)
      for (unsigned k = 0; k < m; k++)
        for (unsigned l = 0; l < m; l++) {
          const double wt = gw[i] * gw[j] * gw[k] * gw[l];
          const double a = gx[i], b = gx[j], c = gx[k], d = gx[l];
          switch (family) {
          case FAM_DIST:  // Duffy-Gauss on each triangle
            push({a, a * b, c, c * d}, wt * a * c);
            break;
          case FAM_VERT: {
            double jw = wt * a * a * a * c;
            Pt4 p = vt_a(a, b, c, d);
            push(p, jw); push(swap(p), jw);
          } break;
          case FAM_EDGE: {
            double j0 = wt * a * a * a * b * b, j1 = j0 * c;
            push(ed_0(a, b, c, d), j0);
            push(ed_1(a, b, c, d), j1); push(ed_2(a, b, c, d), j1);
            push(ed_3(a, b, c, d), j1); push(ed_4(a, b, c, d), j1);
          } break;
          default: {
            double jw = wt * a * a * a * b * b * c;
            Pt4 p = id_a(a, b, c, d); push(p, jw); push(swap(p), jw);
            p = id_b(a, b, c, d);     push(p, jw); push(swap(p), jw);
            p = id_c(a, b, c, d);     push(p, jw); push(swap(p), jw);
          }
          }
        }
}

void build_triangle_rule(unsigned q, std::vector<double> &t, std::vector<double> &s,
                         std::vector<double> &w) {
  std::vector<double> gx, gw;
  gauss_legendre(q, gx, gw);
  t.clear(); s.clear(); w.clear();
  for (unsigned i = 0; i < q; i++)
    for (unsigned j = 0; j < q; j++) {
      double a = 0.5 * (gx[i] + 1.0), b = 0.5 * (gx[j] + 1.0);
      t.push_back(a); s.push_back(a * b); w.push_back(0.25 * gw[i] * gw[j] * a);
    }
}

// Vertex permutations that bring the shared vertices of (tv, sv) to the front, same shared
// vertex at the same slot in both; returns the number of shared vertices.
unsigned match_triangles(const unsigned *tv, const unsigned *sv, unsigned *tp, unsigned *sp) {
  unsigned c = 0;
  bool tused[3] = {false, false, false}, sused[3] = {false, false, false};
  for (unsigned i = 0; i < 3; i++)
    for (unsigned j = 0; j < 3; j++)
      if (tv[i] == sv[j]) { tp[c] = i; sp[c] = j; tused[i] = sused[j] = true; c++; break; }
  unsigned a = c, b = c;
  for (unsigned i = 0; i < 3; i++) {
    if (!tused[i]) tp[a++] = i;
    if (!sused[i]) sp[b++] = i;
  }
  return c;
}

void build_touching_pairs(unsigned nv, unsigned nt, const unsigned *tri, std::vector<TouchPair> &out) {
  // vertex -> incident triangles (CSR), then for every t the union of triangles at its vertices
  std::vector<unsigned> start(nv + 1, 0), adj(3 * (size_t)nt);
  for (size_t i = 0; i < 3 * (size_t)nt; i++) start[tri[i] + 1]++;
  for (unsigned v = 0; v < nv; v++) start[v + 1] += start[v];
  std::vector<unsigned> fill(start.begin(), start.end() - 1);
  for (unsigned t = 0; t < nt; t++)
    for (int d = 0; d < 3; d++) adj[fill[tri[3 * t + d]]++] = t;
  out.clear();
  std::vector<unsigned> cand;
  for (unsigned t = 0; t < nt; t++) {
    cand.clear();
    for (int d = 0; d < 3; d++) {
      unsigned v = tri[3 * t + d];
      for (unsigned p = start[v]; p < start[v + 1]; p++) cand.push_back(adj[p]);
    }
    std::sort(cand.begin(), cand.end());
    cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
    for (unsigned s : cand) {
      unsigned tp[3], sp[3];
      unsigned c = match_triangles(tri + 3 * t, tri + 3 * s, tp, sp);
      TouchPair P;
      for (int d = 0; d < 3; d++) { P.rv[d] = tri[3 * t + tp[d]]; P.cv[d] = tri[3 * s + sp[d]]; }
      P.t = t;
      P.s = s;
      P.family = c;
      out.push_back(P);
    }
  }
}

void build_touch_slots(const std::vector<TouchPair> &pairs, TouchSlots &S) {
  std::vector<unsigned long long> keys;
  keys.reserve(pairs.size() * 9);
  for (const TouchPair &P : pairs)
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) keys.push_back(((unsigned long long)P.rv[i] << 32) | P.cv[j]);
  std::vector<unsigned long long> uniq(keys);
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  auto find = [&](unsigned long long k) { return (unsigned)(std::lower_bound(uniq.begin(), uniq.end(), k) - uniq.begin()); };
  S.row.resize(uniq.size()); S.col.resize(uniq.size()); S.tr.resize(uniq.size());
  for (size_t s = 0; s < uniq.size(); s++) { S.row[s] = (unsigned)(uniq[s] >> 32); S.col[s] = (unsigned)(uniq[s] & 0xffffffffu); }
  for (size_t s = 0; s < uniq.size(); s++) S.tr[s] = find(((unsigned long long)S.col[s] << 32) | S.row[s]);
  S.pair_slot.resize(keys.size());
  for (size_t k = 0; k < keys.size(); k++) S.pair_slot[k] = find(keys[k]);
}

}  // namespace cnld
