// Host-side cluster tree / block tree of the H-matrix path (frequency independent).
// Restates build_bem3d_cluster -> build_adaptive_cluster (include/bem3d.h:1493,
// cluster.h:133-152, clustergeometry.h:45-75), build_strict_block / build_nonstrict_block
// (block.h:216-248) and the admissibility predicates (block.h:78-140).
#pragma once
#include <cstdint>
#include <vector>

namespace cnld {

struct ClusterNode {
  int son[2] = {-1, -1};
  unsigned start = 0, size = 0;  // range in the permutation
  double bmin[3], bmax[3];
  unsigned level = 0;
};

struct ClusterTree {
  std::vector<ClusterNode> node;  // node[0] is the root
  std::vector<unsigned> perm;     // cluster order -> original DOF
  std::vector<unsigned> pos;      // original DOF -> cluster order
};

enum Admis { ADMIS_2 = 0, ADMIS_MAX = 1, ADMIS_SPHERE = 2, ADMIS_2_MIN = 3 };

struct BlockLeaf {
  unsigned rc, cc;  // cluster ids
  bool admissible;
};

// DOF = vertex (linear basis): characteristic point = vertex, support box = box of its triangle fan
void build_vertex_cluster_tree(unsigned nv, unsigned nt, const double *x, const unsigned *tri, unsigned clf,
                               ClusterTree &T);
bool is_admissible(const ClusterNode &r, const ClusterNode &c, double eta, Admis kind);
void build_block_leaves(const ClusterTree &T, double eta, Admis kind, bool strict, std::vector<BlockLeaf> &leaves);

}  // namespace cnld
