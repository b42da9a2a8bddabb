// Host-side quadrature tables (see quadrature.cpp).
#pragma once
#include <algorithm>
#include <vector>

namespace cnld {

enum { FAM_DIST = 0, FAM_VERT = 1, FAM_EDGE = 2, FAM_ID = 3 };

struct PairRule {  // n points: (xt,xs) on the first triangle, (yt,ys) on the second
  std::vector<double> xt, xs, yt, ys, w;
  size_t size() const { return w.size(); }
};

struct TouchPair {    // triangle pair sharing >= 1 vertex (48 bytes, mirrored on device)
  unsigned rv[3];     // row-triangle vertex ids, shared vertices first
  unsigned cv[3];     // col-triangle vertex ids, same shared vertices at the same slots
  unsigned t, s;      // triangle ids
  unsigned family;    // number of shared vertices 1..3
  unsigned pad[3];
};

void gauss_legendre(unsigned m, std::vector<double> &x, std::vector<double> &w);
void build_pair_rule(int family, unsigned q, unsigned q2, PairRule &R);
// single-triangle Duffy-Gauss rule (q*q points) on {0<=s<=t<=1}; weights sum to 1/2
void build_triangle_rule(unsigned q, std::vector<double> &t, std::vector<double> &s,
                         std::vector<double> &w);
unsigned match_triangles(const unsigned *tv, const unsigned *sv, unsigned *tp, unsigned *sp);
void build_touching_pairs(unsigned nv, unsigned nt, const unsigned *tri, std::vector<TouchPair> &out);

// Sparse pattern of all matrix entries that receive contributions from touching pairs:
// slot s = (row[s], col[s]), tr[s] = slot of the transposed entry, pair_slot[9*p + 3*i + j] = slot of
// entry (rv[i], cv[j]) of pair p.  Used by the symmetric path: the Sauter-Schwab rules are not
// symmetric under swapping the two triangles, so these entries carry the (tiny) asymmetry of Z.
struct TouchSlots {
  std::vector<unsigned> row, col, tr, pair_slot;
};
void build_touch_slots(const std::vector<TouchPair> &pairs, TouchSlots &S);

}  // namespace cnld
