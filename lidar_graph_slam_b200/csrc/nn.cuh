// Exact nearest-neighbour search on the GPU: an implicit 32-ary bounding-volume hierarchy over the Morton-sorted
// target cloud, searched by one WARP per query.  Stands in for pcl::search::KdTree / FLANN KDTreeSingleIndex (exact,
// L2_Simple<float>) at FG:133 (1-NN per LM iteration), FG:254 (k-NN covariances) and in
// pcl::Registration::getFitnessScore (GBS:321).
//
// Layout.  Level 0 = leaves of 32 consecutive sorted points; a node of level l + 1 owns 32 consecutive boxes of level
// l.  Boxes of all levels sit in two float4 arrays (min / max), level l at off[l], cnt[l] of them.  No pointers, no
// per-node child counts: the children of node j are boxes 32 j .. min(32 j + 32, cnt) - 1 of the level below.  The
// hierarchy is padded to 3 (clouds up to 2^20 points) or 6 levels (up to 2^31) so that only two traversal depths are
// instantiated; padding levels hold a single box.
//
// Search.  The 32 lanes of a warp test the 32 children of a node (or the 32 points of a leaf) at once: one coalesced
// 512-byte load, one distance each, and warp votes / redux to pick the nearest child still worth visiting.  Control
// flow is warp-uniform and the traversal state lives in registers (recursion is unrolled by templates over the
// level), so there is no per-thread stack in local memory.  The k best candidates are a sorted list held one entry
// per lane (k <= 32); a candidate is inserted with one vote, one shuffle-up and one select.
//
// Exactness.  The f32 box distance uses the same subtraction / multiply / add sequence as the f32 point distance, and
// IEEE rounding is monotone, so box_dist2(q, B) <= dist2(q, p) for every p in B; pruning only on box_dist2 > current
// worst therefore never discards a candidate.  Ties are resolved towards the smaller original point index, which
// makes results independent of the traversal order: any exact method returns the same rows.
#pragma once
#include "common.cuh"

namespace lgs {

constexpr int kLeaf = 32;       // points per leaf = children per node = lanes per warp
constexpr int kMaxK = 32;
constexpr int kMaxLevels = 6;

struct NNView {
  const float4* spts;  // Morton-sorted points; .w carries the original index (int bits)
  const float4* bmin;  // boxes, all levels concatenated: xyz = box min
  const float4* bmax;
  int n;               // points
  int n_levels;        // 3 or 6 (0 for an empty index)
  int off[kMaxLevels];
  int cnt[kMaxLevels];
};

struct NNIndex {
  DevBuf spts, bmin, bmax, codes, codes_alt, perm, perm_alt, small;
  int64_t n = 0;
  int n_levels = 0;
  int off[kMaxLevels] = {0, 0, 0, 0, 0, 0};
  int cnt[kMaxLevels] = {0, 0, 0, 0, 0, 0};
  int build(lgs_ctx* ctx, const float4* pts, int64_t n);
  NNView view() const {
    NNView v;
    v.spts = spts.as<float4>();
    v.bmin = bmin.as<float4>();
    v.bmax = bmax.as<float4>();
    v.n = static_cast<int>(n);
    v.n_levels = n_levels;
    for (int l = 0; l < kMaxLevels; l++) {
      v.off[l] = off[l];
      v.cnt[l] = cnt[l];
    }
    return v;
  }
  void release() {
    for (DevBuf* b : {&spts, &bmin, &bmax, &codes, &codes_alt, &perm, &perm_alt, &small}) b->release();
    n = 0;
    n_levels = 0;
  }
};

// mean squared 1-NN distance of T * src in the index (pcl::Registration::getFitnessScore)
int nn_fitness(lgs_ctx* ctx, const NNIndex& index, const float4* src, int64_t n_src, const float* T16, double max_range, double* fitness);

// k-NN of every indexed point among the indexed points themselves (self included): out_idx / out_d2 are
// n x k, rows addressed by ORIGINAL point index, ascending (d2, idx).
int nn_self_knn(lgs_ctx* ctx, const NNIndex& index, int k, int* out_idx_dev, float* out_d2_dev);

// k-NN of the indexed points list[0 .. *count_dev) (original point indices; the count lives in device memory, `capacity`
// bounds it): row j of out_idx_dev belongs to point list[j].  No host round trip: the launch covers the capacity.
int nn_knn_list(lgs_ctx* ctx, const NNIndex& index, const float4* pts, const int* list_dev, const int* count_dev, int64_t capacity, int k,
                int* out_idx_dev);

// k-NN of arbitrary queries (device float4), rows addressed by query index
int nn_knn(lgs_ctx* ctx, const NNIndex& index, const float4* queries, int64_t m, int k, int* out_idx_dev, float* out_d2_dev);

#ifdef __CUDACC__
constexpr unsigned kFullMask = 0xffffffffu;
constexpr unsigned kNoBox = 0xffffffffu;  // above the bit pattern of +inf: "no child here / already visited"

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

// Squared distances are non-negative, so their bit patterns order like unsigned integers: redux.sync.min.u32 gives
// the warp minimum in one instruction.
__device__ __forceinline__ unsigned nn_child_dists(const NNView& v, int level, int first, int count, float qx, float qy, float qz, int lane) {
  unsigned cd = kNoBox;
  if (lane < count) {
    const float4 lo = __ldg(v.bmin + v.off[level] + first + lane);
    const float4 hi = __ldg(v.bmax + v.off[level] + first + lane);
    cd = __float_as_uint(nn_box_dist2(qx, qy, qz, lo, hi));
  }
  return cd;
}

// ---- 1-NN -------------------------------------------------------------------------------------------------------
// best_d / best_i are warp-uniform.
__device__ __forceinline__ void nn1_leaf(const NNView& v, int leaf, float qx, float qy, float qz, int lane, float& best_d, int& best_i) {
  const int j = leaf * kLeaf + lane;
  unsigned db = kNoBox;
  int oi = 0x7fffffff;
  if (j < v.n) {
    const float4 p = __ldg(v.spts + j);
    db = __float_as_uint(nn_dist2(qx, qy, qz, p));
    oi = __float_as_int(p.w);
  }
  const unsigned m = __reduce_min_sync(kFullMask, db);
  const float md = __uint_as_float(m);
  if (m == kNoBox || md > best_d) return;
  const int mi = __reduce_min_sync(kFullMask, db == m ? oi : 0x7fffffff);
  if (md < best_d || mi < best_i) {
    best_d = md;
    best_i = mi;
  }
}

template <int LC>  // LC = level of the children
__device__ __forceinline__ void nn1_descend(const NNView& v, int first, int count, float qx, float qy, float qz, int lane, float& best_d, int& best_i) {
  unsigned cd = nn_child_dists(v, LC, first, count, qx, qy, qz, lane);
  while (true) {
    const unsigned m = __reduce_min_sync(kFullMask, cd);
    if (m == kNoBox || __uint_as_float(m) > best_d) break;
    const int pick = __ffs(__ballot_sync(kFullMask, cd == m)) - 1;
    if (lane == pick) cd = kNoBox;
    const int c = first + pick;
    if constexpr (LC == 0) {
      nn1_leaf(v, c, qx, qy, qz, lane, best_d, best_i);
    } else {
      nn1_descend<LC - 1>(v, c * kLeaf, min(kLeaf, v.cnt[LC - 1] - c * kLeaf), qx, qy, qz, lane, best_d, best_i);
    }
  }
}

// squared distance and original index of the nearest indexed point (inf / INT_MAX for an empty index); all 32
// lanes of the warp must call with the same query
__device__ __forceinline__ void nn_search1_warp(const NNView& v, float qx, float qy, float qz, int lane, float& best_d, int& best_i) {
  best_d = __int_as_float(0x7f800000);
  best_i = 0x7fffffff;
  if (v.n_levels == 3)
    nn1_descend<2>(v, 0, v.cnt[2], qx, qy, qz, lane, best_d, best_i);
  else if (v.n_levels == 6)
    nn1_descend<5>(v, 0, v.cnt[5], qx, qy, qz, lane, best_d, best_i);
}

// The same search started from a known indexed point (seed_i, at squared distance seed_d from the query; seed_i < 0: no
// seed).  Any indexed point is a valid upper bound and the walk still visits every box that could hold a closer point
// or an equally close one with a smaller index, so the result is the same exact nearest neighbour - found after fewer
// node visits.  Iterative registration uses the previous iteration's neighbour as the seed.
__device__ __forceinline__ void nn_search1_warp_seeded(const NNView& v, float qx, float qy, float qz, int lane, float seed_d, int seed_i, float& best_d,
                                                       int& best_i) {
  best_d = seed_i >= 0 ? seed_d : __int_as_float(0x7f800000);
  best_i = seed_i >= 0 ? seed_i : 0x7fffffff;
  if (v.n_levels == 3)
    nn1_descend<2>(v, 0, v.cnt[2], qx, qy, qz, lane, best_d, best_i);
  else if (v.n_levels == 6)
    nn1_descend<5>(v, 0, v.cnt[5], qx, qy, qz, lane, best_d, best_i);
}

// ---- k-NN -------------------------------------------------------------------------------------------------------
// The warp holds a list of 32 candidates sorted ascending by (d2, idx), entry j in lane j.  A candidate is ONE 64-bit key:
// the bit pattern of its squared distance (non-negative, so it orders like an unsigned integer) above its original index,
// which makes the lexicographic (d2, idx) comparison a single unsigned 64-bit compare - the compare-exchange network of
// the merge path is ~45 % of this kernel's instructions (ncu), and the two-float-one-int comparison was most of that.
// worst (warp-uniform) mirrors the distance of entry k - 1.  Empty entries are (+inf, INT_MAX).
typedef unsigned long long nnkey_t;
constexpr nnkey_t kNNEmpty = (static_cast<nnkey_t>(0x7f800000u) << 32) | 0x7fffffffu;
__device__ __forceinline__ nnkey_t nn_make_key(float d2, int idx) {
  return (static_cast<nnkey_t>(__float_as_uint(d2)) << 32) | static_cast<unsigned>(idx);
}
__device__ __forceinline__ float nn_key_dist(nnkey_t k) { return __uint_as_float(static_cast<unsigned>(k >> 32)); }
__device__ __forceinline__ int nn_key_index(nnkey_t k) { return static_cast<int>(static_cast<unsigned>(k)); }

struct NNKBest {
  nnkey_t key;      // this lane's list entry
  nnkey_t worst_k;  // entry k - 1 (warp-uniform)
  float worst;      // its distance
  bool fresh;       // nothing inserted yet (warp-uniform)
  // views used by the callers
  __device__ __forceinline__ float bd() const { return nn_key_dist(key); }
  __device__ __forceinline__ int bi() const { return nn_key_index(key); }
};

// candidates of one leaf below the current k-th distance from which one sort + merge (about 20 compare-exchange stages) is
// cheaper than an insertion each (about 30 instructions per insertion; ncu: 84 insertions per query on a 0.25 m sweep)
constexpr int kMergeThreshold = 6;

__device__ __forceinline__ void nnk_leaf(const NNView& v, int leaf, float qx, float qy, float qz, int lane, int k, NNKBest& B) {
  const int j = leaf * kLeaf + lane;
  nnkey_t c = kNNEmpty;
  if (j < v.n) {
    const float4 p = __ldg(v.spts + j);
    c = nn_make_key(nn_dist2(qx, qy, qz, p), __float_as_int(p.w));
  }
  unsigned q = 0;
  if (!B.fresh) q = __ballot_sync(kFullMask, c < B.worst_k);
  if (B.fresh || __popc(q) >= kMergeThreshold) {
    // sort the leaf's 32 candidates (bitonic network over the lanes) ...
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
      for (int s = size >> 1; s > 0; s >>= 1) {
        const nnkey_t o = __shfl_xor_sync(kFullMask, c, s);
        const bool want_min = ((lane & size) == 0) == ((lane & s) == 0);
        if (want_min == (o < c)) c = o;  // min keeper takes a smaller partner, max keeper takes a partner that is not smaller
      }
    }
    if (B.fresh) {
      // ... first leaf: the list is empty, the sorted candidates are the list
      B.fresh = false;
      B.key = c;
    } else {
      // ... a leaf with many candidates below the current bound: one merge instead of an insertion per candidate.  The 32
      // smallest of (list, candidates): lane j keeps the smaller of list[j] and candidates[31 - j] (that sequence is
      // bitonic), then a bitonic merge sorts it.  Keys are distinct, so the result is the list the insertions would build.
      const nnkey_t r = __shfl_sync(kFullMask, c, 31 - lane);
      if (r < B.key) B.key = r;
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        const nnkey_t o = __shfl_xor_sync(kFullMask, B.key, s);
        const bool want_min = (lane & s) == 0;
        if (want_min == (o < B.key)) B.key = o;
      }
    }
  } else {
    while (q) {
      const int src = __ffs(q) - 1;
      q &= q - 1;
      const nnkey_t cc = __shfl_sync(kFullMask, c, src);
      if (!(cc < B.worst_k)) continue;  // the list tightened since the vote
      const int pos = __popc(__ballot_sync(kFullMask, B.key < cc));  // entries ahead of the candidate: a prefix
      const nnkey_t up = __shfl_up_sync(kFullMask, B.key, 1);
      if (lane > pos)
        B.key = up;
      else if (lane == pos)
        B.key = cc;
      B.worst_k = __shfl_sync(kFullMask, B.key, k - 1);
    }
    B.worst = nn_key_dist(B.worst_k);
    return;
  }
  B.worst_k = __shfl_sync(kFullMask, B.key, k - 1);
  B.worst = nn_key_dist(B.worst_k);
}

template <int LC>
__device__ __forceinline__ void nnk_descend(const NNView& v, int first, int count, float qx, float qy, float qz, int lane, int k, NNKBest& B) {
  unsigned cd = nn_child_dists(v, LC, first, count, qx, qy, qz, lane);
  while (true) {
    const unsigned m = __reduce_min_sync(kFullMask, cd);
    if (m == kNoBox || __uint_as_float(m) > B.worst) break;
    const int pick = __ffs(__ballot_sync(kFullMask, cd == m)) - 1;
    if (lane == pick) cd = kNoBox;
    const int c = first + pick;
    if constexpr (LC == 0) {
      nnk_leaf(v, c, qx, qy, qz, lane, k, B);
    } else {
      nnk_descend<LC - 1>(v, c * kLeaf, min(kLeaf, v.cnt[LC - 1] - c * kLeaf), qx, qy, qz, lane, k, B);
    }
  }
}

// after the call lane j < k holds the j-th nearest indexed point in B.key (B.bd() / B.bi()); bi == INT_MAX where the index
// has fewer than k points
__device__ __forceinline__ void nn_searchk_warp(const NNView& v, float qx, float qy, float qz, int lane, int k, NNKBest& B) {
  B.key = B.worst_k = kNNEmpty;
  B.worst = __int_as_float(0x7f800000);
  B.fresh = true;
  if (v.n_levels == 3)
    nnk_descend<2>(v, 0, v.cnt[2], qx, qy, qz, lane, k, B);
  else if (v.n_levels == 6)
    nnk_descend<5>(v, 0, v.cnt[5], qx, qy, qz, lane, k, B);
}
#endif

}  // namespace lgs
