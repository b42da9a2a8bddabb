#include "htree.h"

#include <algorithm>
#include <cmath>
#include <limits>

namespace cnld {

namespace {
struct Geometry {  // clustergeometry: points + support boxes per DOF
  const double *pt;
  std::vector<double> smin, smax;
};

unsigned grow(ClusterTree &T, const Geometry &G, unsigned start, unsigned size, unsigned clf, unsigned level) {
  const unsigned id = (unsigned)T.node.size();
  T.node.emplace_back();
  T.node[id].start = start;
  T.node[id].size = size;
  T.node[id].level = level;
  unsigned *idx = T.perm.data() + start;
  bool is_leaf = true;
  if (size > clf) {
    // bounding box of the characteristic points, longest side, midpoint split
    double lo[3], hi[3];
    for (int d = 0; d < 3; d++) { lo[d] = std::numeric_limits<double>::max(); hi[d] = -lo[d]; }
    for (unsigned i = 0; i < size; i++)
      for (int d = 0; d < 3; d++) {
        double v = G.pt[3 * (size_t)idx[i] + d];
        lo[d] = std::min(lo[d], v);
        hi[d] = std::max(hi[d], v);
      }
    int dir = 0;
    double ext = hi[0] - lo[0];
    for (int d = 1; d < 3; d++)
      if (ext < hi[d] - lo[d]) { ext = hi[d] - lo[d]; dir = d; }
    if (ext > 0.0) {
      const double mid = (hi[dir] + lo[dir]) / 2.0;
      unsigned left = 0;  // in-place partition by swapping to the front (order as H2Lib leaves it)
      for (unsigned i = 0; i < size; i++)
        if (G.pt[3 * (size_t)idx[i] + dir] < mid) std::swap(idx[i], idx[left++]);
      unsigned a = grow(T, G, start, left, clf, level + 1);
      unsigned b = grow(T, G, start + left, size - left, clf, level + 1);
      ClusterNode &me = T.node[id];
      me.son[0] = (int)a;
      me.son[1] = (int)b;
      for (int d = 0; d < 3; d++) {
        me.bmin[d] = std::min(T.node[a].bmin[d], T.node[b].bmin[d]);
        me.bmax[d] = std::max(T.node[a].bmax[d], T.node[b].bmax[d]);
      }
      is_leaf = false;
    }
  }
  if (is_leaf) {
    ClusterNode &me = T.node[id];
    for (int d = 0; d < 3; d++) { me.bmin[d] = std::numeric_limits<double>::max(); me.bmax[d] = -me.bmin[d]; }
    for (unsigned i = 0; i < size; i++)
      for (int d = 0; d < 3; d++) {
        me.bmin[d] = std::min(me.bmin[d], G.smin[3 * (size_t)idx[i] + d]);
        me.bmax[d] = std::max(me.bmax[d], G.smax[3 * (size_t)idx[i] + d]);
      }
  }
  return id;
}
}  // namespace

void build_vertex_cluster_tree(unsigned nv, unsigned nt, const double *x, const unsigned *tri, unsigned clf,
                               ClusterTree &T) {
  Geometry G;
  G.pt = x;
  G.smin.assign(x, x + 3 * (size_t)nv);
  G.smax.assign(x, x + 3 * (size_t)nv);
  for (unsigned k = 0; k < nt; k++)
    for (int a = 0; a < 3; a++) {
      const size_t v = tri[3 * (size_t)k + a];
      for (int b = 0; b < 3; b++) {
        const double *p = x + 3 * (size_t)tri[3 * (size_t)k + b];
        for (int d = 0; d < 3; d++) {
          G.smin[3 * v + d] = std::min(G.smin[3 * v + d], p[d]);
          G.smax[3 * v + d] = std::max(G.smax[3 * v + d], p[d]);
        }
      }
    }
  T.node.clear();
  T.perm.resize(nv);
  for (unsigned i = 0; i < nv; i++) T.perm[i] = i;
  grow(T, G, 0, nv, clf, 0);
  T.pos.resize(nv);
  for (unsigned i = 0; i < nv; i++) T.pos[T.perm[i]] = i;
}

bool is_admissible(const ClusterNode &r, const ClusterNode &c, double eta, Admis kind) {
  double dr2 = 0, dc2 = 0, dist2 = 0, drm = 0, dcm = 0, distm = 0;
  for (int d = 0; d < 3; d++) {
    double a = r.bmax[d] - r.bmin[d], b = c.bmax[d] - c.bmin[d];
    double gap = std::max(0.0, std::max(r.bmin[d] - c.bmax[d], c.bmin[d] - r.bmax[d]));
    dr2 += a * a; dc2 += b * b; dist2 += gap * gap;
    drm = std::max(drm, a); dcm = std::max(dcm, b); distm = std::max(distm, gap);
  }
  switch (kind) {
  case ADMIS_2: return std::sqrt(std::max(dr2, dc2)) <= eta * std::sqrt(dist2);
  case ADMIS_2_MIN: return std::sqrt(std::min(dr2, dc2)) <= eta * std::sqrt(dist2);
  case ADMIS_MAX: return std::max(drm, dcm) <= eta * distm;
  default: {  // spheres around the box centres
    double cd2 = 0;
    for (int d = 0; d < 3; d++) {
      double m = 0.5 * (r.bmax[d] + r.bmin[d]) - 0.5 * (c.bmax[d] + c.bmin[d]);
      cd2 += m * m;
    }
    double rr = 0.5 * std::sqrt(dr2), rc = 0.5 * std::sqrt(dc2);
    double dist = std::sqrt(cd2) - rr - rc;
    return 2.0 * std::max(rr, rc) <= eta * std::max(dist, 0.0);
  }
  }
}

namespace {
void descend(const ClusterTree &T, unsigned r, unsigned c, double eta, Admis kind, bool strict,
             std::vector<BlockLeaf> &out) {
  const ClusterNode &R = T.node[r], &C = T.node[c];
  if (is_admissible(R, C, eta, kind)) { out.push_back({r, c, true}); return; }
  const bool rs = R.son[0] >= 0, cs = C.son[0] >= 0;
  if (rs && cs) {
    for (int j = 0; j < 2; j++)      // H2Lib son order: son[i + j*rsons]
      for (int i = 0; i < 2; i++) descend(T, (unsigned)R.son[i], (unsigned)C.son[j], eta, kind, strict, out);
  } else if (!strict && rs) {
    for (int i = 0; i < 2; i++) descend(T, (unsigned)R.son[i], c, eta, kind, strict, out);
  } else if (!strict && cs) {
    for (int j = 0; j < 2; j++) descend(T, r, (unsigned)C.son[j], eta, kind, strict, out);
  } else {
    out.push_back({r, c, false});
  }
}
}  // namespace

void build_block_leaves(const ClusterTree &T, double eta, Admis kind, bool strict, std::vector<BlockLeaf> &leaves) {
  leaves.clear();
  descend(T, 0, 0, eta, kind, strict, leaves);
}

}  // namespace cnld
