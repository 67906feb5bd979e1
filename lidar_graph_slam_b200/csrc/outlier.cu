// pcl::StatisticalOutlierRemoval<PointXYZI> on the GPU: the stage the prefilter node runs right after the voxel grid
// (points_prefiltering.cpp:79-80,132-140; mean_k = 30, stddev = 1.2 by default, launch/points_prefiltering.launch.xml:4-5).
//
//   NNIndex::build     Morton sort + implicit 32-ary BVH over the cloud (nn.cu)
//   sor_mean_dist      one warp per point: exact (mean_k + 1)-NN of the point in its own cloud, entry 0 (the point
//                      itself) skipped, the f32 square roots of the rest summed in f64 in ascending-neighbour order,
//                      distances[i] = float(sum / mean_k)   -- every operation as PCL's applyFilterIndices does it
//   sor_stats          one CTA, fixed-order f64 reduction of sum and sum of (f32) squares -> mean, sample standard
//                      deviation, threshold = mean + mul * stddev
//   scan_select        keep point i iff !(distances[i] > threshold) (or the complement), order preserved (sort.cuh)
//
// Membership is decided by f32 distances against an f64 threshold; the only difference from a serial evaluation is
// the order of the two global f64 sums (tree instead of index order, ~1e-16 relative on the threshold).
#include <cmath>

#include "nn.cuh"
#include "sort.cuh"

struct lgs_sor {
  lgs_ctx* ctx = nullptr;
  int mean_k = 1;          // pcl::StatisticalOutlierRemoval default
  double std_mul = 0.0;
  int negative = 0;
  lgs::NNIndex index;
  lgs::DevBuf cloud, dist, out, keep, small;
};

namespace lgs {

struct SorStats {
  double mean, stddev, threshold;
  int n_out;
  int pad;
};

constexpr int kSorBlock = 256;
constexpr int kSorWarps = kSorBlock / 32;

__global__ void __launch_bounds__(kSorBlock) sor_mean_dist_kernel(NNView v, int k, int mean_k, float* __restrict__ distances) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = blockIdx.x * kSorWarps + warp; t < v.n; t += gridDim.x * kSorWarps) {
    const float4 qp = __ldg(v.spts + t);  // sorted order: neighbouring warps walk the same part of the tree
    NNKBest B;
    nn_searchk_warp(v, qp.x, qp.y, qp.z, lane, k, B);
    const float r = (lane >= 1 && lane < k && B.bi() != 0x7fffffff) ? __fsqrt_rn(B.bd()) : 0.f;
    double s = 0.0;
    for (int j = 1; j < k; j++) s = __dadd_rn(s, static_cast<double>(__shfl_sync(kFullMask, r, j)));  // ascending-neighbour order
    if (lane == 0) distances[__float_as_int(qp.w)] = static_cast<float>(__ddiv_rn(s, static_cast<double>(mean_k)));
  }
}

__global__ void __launch_bounds__(1024) sor_stats_kernel(const float* __restrict__ distances, int n, double std_mul, SorStats* __restrict__ st) {
  __shared__ double s_sum[32], s_sq[32];
  double sum = 0, sq = 0;
  for (int i = threadIdx.x; i < n; i += 1024) {
    const float d = distances[i];
    sum = __dadd_rn(sum, static_cast<double>(d));
    sq = __dadd_rn(sq, static_cast<double>(__fmul_rn(d, d)));
  }
  for (int off = 16; off > 0; off >>= 1) {
    sum = __dadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, off));
    sq = __dadd_rn(sq, __shfl_xor_sync(0xffffffffu, sq, off));
  }
  if ((threadIdx.x & 31) == 0) {
    s_sum[threadIdx.x >> 5] = sum;
    s_sq[threadIdx.x >> 5] = sq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    sum = sq = 0;
    for (int w = 0; w < 32; w++) {
      sum = __dadd_rn(sum, s_sum[w]);
      sq = __dadd_rn(sq, s_sq[w]);
    }
    const double valid = static_cast<double>(n);
    const double mean = __ddiv_rn(sum, valid);
    const double variance = __ddiv_rn(__dsub_rn(sq, __ddiv_rn(__dmul_rn(sum, sum), valid)), __dsub_rn(valid, 1.0));
    const double stddev = sqrt(variance);
    st->mean = mean;
    st->stddev = stddev;
    st->threshold = __dadd_rn(mean, __dmul_rn(std_mul, stddev));
  }
}

struct SorKeep {
  const float4* pts;
  const float* distances;
  const SorStats* st;
  int negative;
  float4* out;
  unsigned char* keep;
  __device__ bool flag(int64_t i) const {
    const bool outlier = static_cast<double>(distances[i]) > st->threshold;
    return negative ? outlier : !outlier;
  }
  __device__ void emit(int64_t i, int64_t pos, bool f) const {
    if (keep) keep[i] = f ? 1 : 0;
    if (f && out) out[pos] = pts[i];
  }
};

static int sor_run(lgs_sor* s, const float4* pts, int64_t n, float4* out_dev, unsigned char* keep_dev, float* dist_dev, lgs_sor_info* info) {
  lgs_ctx* ctx = s->ctx;
  cudaStream_t st = ctx->stream;
  memset(info, 0, sizeof(*info));
  if (n == 0) return LGS_OK;
  LGS_REQUIRE(s->mean_k >= 1 && s->mean_k + 1 <= kMaxK, "mean_k must be in [1, 31]");
  LGS_TRY(s->index.build(ctx, pts, n));
  LGS_TRY(s->small.reserve(sizeof(SorStats)));
  if (!dist_dev) {
    LGS_TRY(s->dist.reserve(static_cast<size_t>(n) * 4));
    dist_dev = s->dist.as<float>();
  }
  SorStats* stats = s->small.as<SorStats>();
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + kSorWarps - 1) / kSorWarps, kNumSMs * 64)));
  sor_mean_dist_kernel<<<grid, kSorBlock, 0, st>>>(s->index.view(), s->mean_k + 1, s->mean_k, dist_dev);
  sor_stats_kernel<<<1, 1024, 0, st>>>(dist_dev, static_cast<int>(n), s->std_mul, stats);
  ctx->launches += 2;
  LGS_TRY(scan_select(ctx, SorKeep{pts, dist_dev, stats, s->negative, out_dev, keep_dev}, n, &stats->n_out));
  LGS_TRY(ctx->pin.reserve(256));
  SorStats* h = ctx->pin.as<SorStats>();
  LGS_CUDA(cudaMemcpyAsync(h, stats, sizeof(SorStats), cudaMemcpyDeviceToHost, st));
  LGS_CUDA(cudaStreamSynchronize(st));
  info->n_out = h->n_out;
  info->mean = h->mean;
  info->stddev = h->stddev;
  info->threshold = h->threshold;
  return LGS_OK;
}

}  // namespace lgs

extern "C" {

int lgs_sor_create(lgs_ctx* ctx, lgs_sor** out) {
  LGS_REQUIRE(ctx && out, "null argument");
  lgs_sor* s = new lgs_sor;
  s->ctx = ctx;
  *out = s;
  return LGS_OK;
}

void lgs_sor_destroy(lgs_sor* s) {
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  s->index.release();
  for (lgs::DevBuf* b : {&s->cloud, &s->dist, &s->out, &s->keep, &s->small}) b->release();
  delete s;
}

int lgs_sor_set_mean_k(lgs_sor* s, int32_t k) {
  LGS_REQUIRE(s, "null");
  LGS_REQUIRE(k >= 1 && k + 1 <= lgs::kMaxK, "mean_k must be in [1, 31]");
  s->mean_k = k;
  return LGS_OK;
}
int lgs_sor_set_stddev_mul_thresh(lgs_sor* s, double m) {
  LGS_REQUIRE(s, "null");
  s->std_mul = m;
  return LGS_OK;
}
int lgs_sor_set_negative(lgs_sor* s, int32_t negative) {
  LGS_REQUIRE(s, "null");
  s->negative = negative ? 1 : 0;
  return LGS_OK;
}

int lgs_sor_filter_dev(lgs_sor* s, const float* pts_dev, int64_t n, float* out_pts_dev, uint8_t* out_keep_dev, float* out_distances_dev,
                       lgs_sor_info* info) {
  LGS_NVTX("lgs_sor_filter_dev");
  LGS_REQUIRE(s && info, "null argument");
  LGS_REQUIRE(n >= 0 && n < (int64_t(1) << 31), "point count out of range");
  LGS_TRY(lgs::use_device(s->ctx));
  return lgs::sor_run(s, reinterpret_cast<const float4*>(pts_dev), n, reinterpret_cast<float4*>(out_pts_dev), out_keep_dev, out_distances_dev, info);
}

int lgs_sor_filter(lgs_sor* s, const void* pts, int64_t n, int32_t stride_bytes, float* out_pts, uint8_t* out_keep, float* out_distances,
                   lgs_sor_info* info) {
  LGS_NVTX("lgs_sor_filter");
  LGS_REQUIRE(s && info, "null argument");
  LGS_REQUIRE(n >= 0 && n < (int64_t(1) << 31), "point count out of range");
  lgs_ctx* ctx = s->ctx;
  LGS_TRY(lgs::use_device(ctx));
  LGS_TRY(lgs::upload_cloud(ctx, pts, n, stride_bytes, &s->cloud));
  const size_t nn = static_cast<size_t>(n > 0 ? n : 1);
  LGS_TRY(s->out.reserve(nn * 16));
  LGS_TRY(s->keep.reserve(nn));
  LGS_TRY(s->dist.reserve(nn * 4));
  LGS_TRY(lgs::sor_run(s, s->cloud.as<float4>(), n, s->out.as<float4>(), s->keep.as<unsigned char>(), s->dist.as<float>(), info));
  cudaStream_t st = ctx->stream;
  if (out_pts && info->n_out) LGS_CUDA(cudaMemcpyAsync(out_pts, s->out.p, static_cast<size_t>(info->n_out) * 16, cudaMemcpyDeviceToHost, st));
  if (out_keep && n) LGS_CUDA(cudaMemcpyAsync(out_keep, s->keep.p, static_cast<size_t>(n), cudaMemcpyDeviceToHost, st));
  if (out_distances && n) LGS_CUDA(cudaMemcpyAsync(out_distances, s->dist.p, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, st));
  LGS_CUDA(cudaStreamSynchronize(st));
  return LGS_OK;
}

}  // extern "C"
