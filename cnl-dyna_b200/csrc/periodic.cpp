#include "periodic.h"

#include <algorithm>
#include <cmath>
#include <numeric>

namespace cnld {
namespace {

void add_rectangle(std::vector<WorkItem> &work, unsigned T0, unsigned T1, unsigned S0, unsigned S1, bool self,
                   unsigned rt, unsigned cc) {
  for (unsigned tb = T0; tb < T1; tb += rt) {
    const unsigned te = std::min(T1, tb + rt);
    for (unsigned sb = S0; sb < S1; sb += cc) {
      if (self && sb + 1 >= te) break;  // pairs need s < t
      work.push_back({tb, te, sb, std::min(S1, sb + cc)});
    }
  }
}

unsigned find_root(std::vector<unsigned> &p, unsigned v) {
  while (p[v] != v) { p[v] = p[p[v]]; v = p[v]; }
  return v;
}

// true when the mesh is ncomp translated copies of its first component, stored one after another
bool detect(unsigned nv, unsigned nt, const double *x, const unsigned *tri, TranslationPlan &P) {
  if (nv == 0 || nt == 0) return false;
  std::vector<unsigned> parent(nv);
  std::iota(parent.begin(), parent.end(), 0u);
  for (unsigned t = 0; t < nt; t++) {
    unsigned a = find_root(parent, tri[3 * t]);
    for (int d = 1; d < 3; d++) {
      unsigned b = find_root(parent, tri[3 * t + d]);
      if (a != b) { if (b < a) std::swap(a, b); parent[b] = a; }   // root = smallest vertex
    }
  }
  // with "root = smallest vertex" a component stored contiguously has root == its first vertex
  std::vector<unsigned> roots;
  for (unsigned v = 0; v < nv; v++) if (find_root(parent, v) == v) roots.push_back(v);
  const unsigned nc = (unsigned)roots.size();
  if (nc < 2 || nv % nc || nt % nc) return false;
  const unsigned nvc = nv / nc, ntc = nt / nc;
  for (unsigned c = 0; c < nc; c++) if (roots[c] != c * nvc) return false;
  for (unsigned v = 0; v < nv; v++) if (find_root(parent, v) != (v / nvc) * nvc) return false;
  double maxabs = 0.0;
  for (size_t i = 0; i < 3 * (size_t)nv; i++) maxabs = std::max(maxabs, std::fabs(x[i]));
  const double tol = 32.0 * 2.220446049250313e-16 * maxabs;
  for (unsigned c = 1; c < nc; c++) {
    const unsigned v0 = c * nvc, t0 = c * ntc;
    for (unsigned i = 0; i < 3 * ntc; i++)
      if (tri[3 * (size_t)t0 + i] != tri[i] + v0) return false;   // same connectivity, same order
    const double off[3] = {x[3 * (size_t)v0] - x[0], x[3 * (size_t)v0 + 1] - x[1], x[3 * (size_t)v0 + 2] - x[2]};
    for (unsigned i = 0; i < nvc; i++)
      for (int d = 0; d < 3; d++)
        if (std::fabs(x[3 * (size_t)(v0 + i) + d] - x[3 * (size_t)i + d] - off[d]) > tol) return false;
  }
  P.ncomp = nc; P.nvc = nvc; P.ntc = ntc;
  return true;
}
}  // namespace

void build_translation_plan(unsigned nv, unsigned nt, const double *x, const unsigned *tri,
                            unsigned rt, unsigned cc, bool allow_periodic, TranslationPlan &P) {
  P = TranslationPlan();
  P.ntc = nt;
  if (!allow_periodic || !detect(nv, nt, x, tri, P)) {
    P.periodic = false; P.ncomp = 1; P.nvc = nv; P.ntc = nt;
    add_rectangle(P.work, 0, nt, 0, nt, true, rt, cc);
    return;
  }
  P.periodic = true;
  const unsigned nc = P.ncomp, nvc = P.nvc, ntc = P.ntc;
  double maxabs = 0.0;
  for (size_t i = 0; i < 3 * (size_t)nv; i++) maxabs = std::max(maxabs, std::fabs(x[i]));
  const double tol = 64.0 * 2.220446049250313e-16 * maxabs;
  // diagonal blocks: component 0 is integrated (pairs s < t), the others are copies
  add_rectangle(P.work, 0, ntc, 0, ntc, true, rt, cc);
  for (unsigned a = 1; a < nc; a++) P.copy.push_back({a * nvc, a * nvc, 0, 0, 0});
  // off-diagonal blocks (a > b): one integrated block per offset class
  struct Cls { double d[3]; unsigned a, b; };
  std::vector<Cls> classes;
  for (unsigned a = 1; a < nc; a++)
    for (unsigned b = 0; b < a; b++) {
      double d[3];
      for (int k = 0; k < 3; k++) d[k] = x[3 * (size_t)(a * nvc) + k] - x[3 * (size_t)(b * nvc) + k];
      int found = -1, transposed = 0;
      for (size_t c = 0; c < classes.size() && found < 0; c++) {
        const double *e = classes[c].d;
        if (std::fabs(d[0] - e[0]) <= tol && std::fabs(d[1] - e[1]) <= tol && std::fabs(d[2] - e[2]) <= tol) { found = (int)c; transposed = 0; }
        else if (std::fabs(d[0] + e[0]) <= tol && std::fabs(d[1] + e[1]) <= tol && std::fabs(d[2] + e[2]) <= tol) { found = (int)c; transposed = 1; }
      }
      if (found < 0) {
        classes.push_back({{d[0], d[1], d[2]}, a, b});
        add_rectangle(P.work, a * ntc, (a + 1) * ntc, b * ntc, (b + 1) * ntc, false, rt, cc);
      } else {
        P.copy.push_back({a * nvc, b * nvc, classes[found].a * nvc, classes[found].b * nvc, (unsigned)transposed});
      }
    }
  P.nclass = (unsigned)classes.size();
}

}  // namespace cnld
