// ORACLE (test infrastructure).  Exact k-NN + pcl::Registration::getFitnessScore.
//
// Stands in for pcl::search::KdTree -> pcl::KdTreeFLANN -> flann::KDTreeSingleIndex<L2_Simple<float>>
// (PCL / FLANN 1.9, un-vendored; call sites FG:133, FG:254, and PCL registration.hpp getFitnessScore
// called at GBS:321).  FLANN's search is exact (eps = 0), so what must be restated is the metric:
// L2_Simple accumulates diff*diff left to right in f32.  Ties are arbitrary in FLANN; the oracle
// resolves them towards the smaller point index so that results are reproducible.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <numeric>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle.hpp"

namespace lgs_oracle {

static inline float dist2(const P4& a, const P4& b) {
  float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  return (dx * dx + dy * dy) + dz * dz;
}

static inline float box_dist2(const float* bmin, const float* bmax, const P4& q) {
  // per-axis gap computed with the same subtraction the point metric uses, so that for every point
  // p inside the box fl-gap <= |fl(q-p)| and the f32 sum below is <= dist2(q,p) (rounding is monotone)
  float g[3];
  const float c[3] = {q.x, q.y, q.z};
  for (int a = 0; a < 3; a++) {
    float lo = bmin[a] - c[a], hi = c[a] - bmax[a];
    g[a] = lo > 0.f ? lo : (hi > 0.f ? hi : 0.f);
  }
  return (g[0] * g[0] + g[1] * g[1]) + g[2] * g[2];
}

int KdTree::build_rec(int begin, int end) {
  Node nd;
  nd.left = nd.right = -1;
  nd.begin = begin;
  nd.end = end;
  nd.dim = 0;
  nd.split_lo = nd.split_hi = 0;
  for (int a = 0; a < 3; a++) {
    nd.bmin[a] = FLT_MAX;
    nd.bmax[a] = -FLT_MAX;
  }
  for (int i = begin; i < end; i++) {
    const float c[3] = {pts[i].x, pts[i].y, pts[i].z};
    for (int a = 0; a < 3; a++) {
      nd.bmin[a] = std::min(nd.bmin[a], c[a]);
      nd.bmax[a] = std::max(nd.bmax[a], c[a]);
    }
  }
  int id = static_cast<int>(nodes.size());
  nodes.push_back(nd);
  const int leaf_max = 12;
  if (end - begin > leaf_max) {
    int dim = 0;
    float ext = -1;
    for (int a = 0; a < 3; a++)
      if (nd.bmax[a] - nd.bmin[a] > ext) {
        ext = nd.bmax[a] - nd.bmin[a];
        dim = a;
      }
    if (ext > 0) {
      int mid = (begin + end) / 2;
      // order/pts are permuted together through an index sort on the sub-range
      std::vector<int> perm(end - begin);
      std::iota(perm.begin(), perm.end(), begin);
      auto key = [&](int i) { return dim == 0 ? pts[i].x : (dim == 1 ? pts[i].y : pts[i].z); };
      std::nth_element(perm.begin(), perm.begin() + (mid - begin), perm.end(), [&](int a, int b) {
        float ka = key(a), kb = key(b);
        return ka < kb || (ka == kb && order[a] < order[b]);
      });
      std::vector<P4> tp(end - begin);
      std::vector<int32_t> to(end - begin);
      for (int i = 0; i < end - begin; i++) {
        tp[i] = pts[perm[i]];
        to[i] = order[perm[i]];
      }
      std::copy(tp.begin(), tp.end(), pts.begin() + begin);
      std::copy(to.begin(), to.end(), order.begin() + begin);
      int l = build_rec(begin, mid);
      int r = build_rec(mid, end);
      nodes[id].left = l;
      nodes[id].right = r;
      nodes[id].dim = dim;
    }
  }
  return id;
}

void KdTree::build(const P4* p, size_t n) {
  nodes.clear();
  pts.assign(p, p + n);
  order.resize(n);
  std::iota(order.begin(), order.end(), 0);
  if (n) {
    nodes.reserve(2 * n / 6 + 16);
    build_rec(0, static_cast<int>(n));
  }
}

int KdTree::knn(const P4& q, int k, int32_t* idx, float* d2) const {
  int found = 0;
  if (nodes.empty() || k <= 0) return 0;
  auto better = [](float da, int ia, float db, int ib) { return da < db || (da == db && ia < ib); };
  int stack[128];
  float sdist[128];
  int sp = 0;
  stack[sp] = 0;
  sdist[sp++] = box_dist2(nodes[0].bmin, nodes[0].bmax, q);
  while (sp) {
    --sp;
    int ni = stack[sp];
    float bd = sdist[sp];
    if (found == k && bd > d2[k - 1]) continue;
    const Node& nd = nodes[ni];
    if (nd.left < 0) {
      for (int i = nd.begin; i < nd.end; i++) {
        float d = dist2(q, pts[i]);
        int oi = order[i];
        if (found < k || better(d, oi, d2[found - 1], idx[found - 1])) {
          int pos = found < k ? found : k - 1;
          if (found < k) found++;
          while (pos > 0 && better(d, oi, d2[pos - 1], idx[pos - 1])) {
            d2[pos] = d2[pos - 1];
            idx[pos] = idx[pos - 1];
            --pos;
          }
          d2[pos] = d;
          idx[pos] = oi;
        }
      }
    } else {
      float dl = box_dist2(nodes[nd.left].bmin, nodes[nd.left].bmax, q);
      float dr = box_dist2(nodes[nd.right].bmin, nodes[nd.right].bmax, q);
      // push the farther child first so the nearer one is visited next
      if (dl <= dr) {
        stack[sp] = nd.right; sdist[sp++] = dr;
        stack[sp] = nd.left;  sdist[sp++] = dl;
      } else {
        stack[sp] = nd.left;  sdist[sp++] = dl;
        stack[sp] = nd.right; sdist[sp++] = dr;
      }
    }
  }
  return found;
}

// pcl::Registration::getFitnessScore(max_range) (PCL registration.hpp): transform the source by the
// final transformation, exact 1-NN squared distance (f32) in the target, mean (f64) of those
// <= max_range, DBL_MAX if none.  The reference loop is serial; num_threads only speeds the oracle
// up (per-point distances are gathered, then summed serially in index order).
double fitness_score(const KdTree& tree, const P4* src, size_t n, const float* T, double max_range, int num_threads) {
  std::vector<float> d(n);
  (void)num_threads;
#pragma omp parallel for num_threads(num_threads) schedule(static)
  for (long i = 0; i < static_cast<long>(n); i++) {
    P4 q = transform_point(T, src[i]);
    int32_t id;
    float dd = FLT_MAX;
    tree.knn(q, 1, &id, &dd);
    d[i] = dd;
  }
  double score = 0.0;
  int nr = 0;
  for (size_t i = 0; i < n; i++) {
    if (d[i] <= max_range) {
      score += d[i];
      nr++;
    }
  }
  return nr > 0 ? score / nr : std::numeric_limits<double>::max();
}

}  // namespace lgs_oracle
