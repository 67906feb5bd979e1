// Exact nearest-neighbour search on the GPU: an implicit, complete binary BVH over the Morton-sorted
// target cloud.  Stands in for pcl::search::KdTree / FLANN KDTreeSingleIndex (exact, L2_Simple<float>)
// at FG:133 (1-NN per LM iteration), FG:254 (k-NN covariances) and in pcl::Registration::getFitnessScore
// (GBS:321).  Exactness: the f32 box distance uses the same subtraction / multiply / add sequence as the
// f32 point distance, and IEEE rounding is monotone, so box_dist2(q, B) <= dist2(q, p) for every p in B;
// pruning only on box_dist2 > current worst therefore never discards a candidate.  Ties are resolved
// towards the smaller original point index, which makes results independent of traversal order.
#pragma once
#include "common.cuh"

namespace lgs {

constexpr int kLeafSize = 8;
constexpr int kMaxK = 32;

struct NNView {
  const float4* spts;  // Morton-sorted points; .w carries the original index (int bits)
  const float4* bmin;  // per node (heap order): xyz = box min
  const float4* bmax;
  int n;               // points
  int n_leaves_p2;     // padded leaf count (power of two)
};

struct NNIndex {
  DevBuf spts, bmin, bmax, codes, codes_alt, perm, perm_alt, small;
  int64_t n = 0;
  int n_leaves_p2 = 0;
  int build(lgs_ctx* ctx, const float4* pts, int64_t n);
  NNView view() const { return NNView{spts.as<float4>(), bmin.as<float4>(), bmax.as<float4>(), static_cast<int>(n), n_leaves_p2}; }
  void release() {
    for (DevBuf* b : {&spts, &bmin, &bmax, &codes, &codes_alt, &perm, &perm_alt, &small}) b->release();
    n = 0;
  }
};

// mean squared 1-NN distance of T * src in the index (pcl::Registration::getFitnessScore)
int nn_fitness(lgs_ctx* ctx, const NNIndex& index, const float4* src, int64_t n_src, const float* T16, double max_range, double* fitness);

// k-NN of every indexed point among the indexed points themselves (self included): out_idx / out_d2 are
// n x k, rows addressed by ORIGINAL point index, ascending (d2, idx).
int nn_self_knn(lgs_ctx* ctx, const NNIndex& index, int k, int* out_idx_dev, float* out_d2_dev);

// k-NN of arbitrary queries (device float4), rows addressed by query index
int nn_knn(lgs_ctx* ctx, const NNIndex& index, const float4* queries, int64_t m, int k, int* out_idx_dev, float* out_d2_dev);

#ifdef __CUDACC__
__device__ __forceinline__ float nn_dist2(float qx, float qy, float qz, const float4& p) {
  const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
__device__ __forceinline__ float nn_box_dist2(float qx, float qy, float qz, const float4& lo, const float4& hi) {
  float gx = __fsub_rn(lo.x, qx), hx = __fsub_rn(qx, hi.x);
  float gy = __fsub_rn(lo.y, qy), hy = __fsub_rn(qy, hi.y);
  float gz = __fsub_rn(lo.z, qz), hz = __fsub_rn(qz, hi.z);
  gx = gx > 0.f ? gx : (hx > 0.f ? hx : 0.f);
  gy = gy > 0.f ? gy : (hy > 0.f ? hy : 0.f);
  gz = gz > 0.f ? gz : (hz > 0.f ? hz : 0.f);
  return __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
}

// 1-NN traversal.  Returns the squared distance and the original index of the nearest indexed point.
__device__ __forceinline__ void nn_search1(const NNView& v, float qx, float qy, float qz, float& best_d, int& best_i) {
  best_d = __int_as_float(0x7f800000);
  best_i = 0x7fffffff;
  if (v.n == 0) return;
  const int first_leaf = v.n_leaves_p2 - 1;
  int stack[48];
  float sdist[48];
  int sp = 0;
  stack[0] = 0;
  sdist[0] = nn_box_dist2(qx, qy, qz, __ldg(v.bmin), __ldg(v.bmax));
  sp = 1;
  while (sp > 0) {
    --sp;
    const int node = stack[sp];
    const float bd = sdist[sp];
    if (bd > best_d) continue;
    if (node >= first_leaf) {
      const int b = (node - first_leaf) * kLeafSize;
      const int e = min(b + kLeafSize, v.n);
      for (int j = b; j < e; j++) {
        const float4 p = __ldg(v.spts + j);
        const float d = nn_dist2(qx, qy, qz, p);
        const int oi = __float_as_int(p.w);
        if (d < best_d || (d == best_d && oi < best_i)) {
          best_d = d;
          best_i = oi;
        }
      }
    } else {
      const int l = 2 * node + 1, r = l + 1;
      const float dl = nn_box_dist2(qx, qy, qz, __ldg(v.bmin + l), __ldg(v.bmax + l));
      const float dr = nn_box_dist2(qx, qy, qz, __ldg(v.bmin + r), __ldg(v.bmax + r));
      if (dl <= dr) {
        if (dr <= best_d) { stack[sp] = r; sdist[sp++] = dr; }
        if (dl <= best_d) { stack[sp] = l; sdist[sp++] = dl; }
      } else {
        if (dl <= best_d) { stack[sp] = l; sdist[sp++] = dl; }
        if (dr <= best_d) { stack[sp] = r; sdist[sp++] = dr; }
      }
    }
  }
}
#endif

}  // namespace lgs
