// Batched loop-closure verification: GBS:297-322 (VoxelGrid 0.5 m on the submap, setInputTarget / setInputSource,
// align with identity guess, getFinalTransformation / getFitnessScore / hasConverged) for a list of candidate
// pairs.  Pairs are independent, so W host workers each drive their own stream and registration object and pull
// pairs from a shared counter; kernels of different pairs overlap on the device.  Every pair runs entirely on one
// stream with deterministic reductions, so a pair's record does not depend on W, on the pair order, or on which
// GPU of the box it was dealt to.
#include <atomic>
#include <limits>
#include <mutex>
#include <thread>
#include <vector>

#include "gicp.cuh"
#include "voxel_common.cuh"

using namespace lgs;

namespace {

struct BatchShared {
  const lgs_batch_params* bp;
  int device;
  int64_t n_pairs;
  const void* const* scans;
  const int64_t* n_scan;
  const void* const* submaps;
  const int64_t* n_submap;
  int32_t stride;
  const float* guesses;
  int32_t pair_id0;
  lgs_align_result* records;
  std::atomic<int64_t> next{0};
  std::atomic<int> failed{0};
  std::mutex err_mu;
  std::string err;
};

int run_pair(lgs_ctx* ctx, lgs_gicp* gicp, lgs_ndt* ndt, DevBuf* sub_raw, DevBuf* sub_ds, BatchShared* S, int64_t i) {
  const lgs_batch_params& bp = *S->bp;
  lgs_align_result& rec = S->records[i];
  memset(&rec, 0, sizeof(rec));
  // submap -> (optional) VoxelGrid -> device-resident target
  LGS_TRY(upload_cloud(ctx, S->submaps[i], S->n_submap[i], S->stride, sub_raw));
  const float* tgt_dev = sub_raw->as<float>();
  int64_t n_tgt = S->n_submap[i];
  if (bp.submap_leaf > 0 && n_tgt > 0) {
    LGS_TRY(sub_ds->reserve(static_cast<size_t>(n_tgt) * 16));
    const float leaf[3] = {bp.submap_leaf, bp.submap_leaf, bp.submap_leaf};
    lgs_voxelgrid_info info;
    LGS_TRY(voxelgrid_device(ctx, sub_raw->as<float4>(), n_tgt, leaf, 0, -1.0, nullptr, sub_ds->as<float4>(), nullptr, nullptr, &info));
    tgt_dev = sub_ds->as<float>();
    n_tgt = info.n_out;
  }
  const float* guess = S->guesses ? S->guesses + 16 * i : nullptr;
  const double max_range = bp.fitness_max_range > 0 ? bp.fitness_max_range : std::numeric_limits<double>::max();
  if (bp.method == LGS_METHOD_GICP) {
    LGS_TRY(lgs_gicp_set_target_dev(gicp, tgt_dev, n_tgt));
    LGS_TRY(lgs_gicp_set_source(gicp, S->scans[i], S->n_scan[i], S->stride));
    LGS_TRY(lgs_gicp_align(gicp, guess, &rec, nullptr));
    LGS_TRY(lgs_gicp_fitness(gicp, max_range, &rec.fitness));
  } else {
    LGS_TRY(lgs_ndt_set_target_dev(ndt, tgt_dev, n_tgt));
    LGS_TRY(lgs_ndt_set_source(ndt, S->scans[i], S->n_scan[i], S->stride));
    LGS_TRY(lgs_ndt_align(ndt, guess, &rec, nullptr));
    LGS_TRY(lgs_ndt_fitness(ndt, max_range, &rec.fitness));
  }
  rec.pair_id = S->pair_id0 + static_cast<int32_t>(i);
  return LGS_OK;
}

void worker(BatchShared* S) {
  lgs_ctx* ctx = nullptr;
  lgs_gicp* gicp = nullptr;
  lgs_ndt* ndt = nullptr;
  DevBuf sub_raw, sub_ds;
  int rc = lgs_ctx_create(S->device, nullptr, &ctx);
  const lgs_batch_params& bp = *S->bp;
  if (rc == LGS_OK) {
    if (bp.method == LGS_METHOD_GICP) {
      rc = lgs_gicp_create(ctx, &gicp);
      if (rc == LGS_OK) {
        if (bp.k_correspondences > 0) lgs_gicp_set_correspondence_randomness(gicp, bp.k_correspondences);
        if (bp.max_iterations > 0) lgs_gicp_set_maximum_iterations(gicp, bp.max_iterations);
        if (bp.transformation_epsilon > 0) lgs_gicp_set_transformation_epsilon(gicp, bp.transformation_epsilon);
        if (bp.max_correspondence_distance > 0) lgs_gicp_set_max_correspondence_distance(gicp, bp.max_correspondence_distance);
      }
    } else {
      rc = lgs_ndt_create(ctx, &ndt);
      if (rc == LGS_OK) {
        if (bp.ndt_resolution > 0) lgs_ndt_set_resolution(ndt, bp.ndt_resolution);
        if (bp.ndt_step_size > 0) lgs_ndt_set_step_size(ndt, bp.ndt_step_size);
        if (bp.max_iterations > 0) lgs_ndt_set_maximum_iterations(ndt, bp.max_iterations);
        if (bp.transformation_epsilon > 0) lgs_ndt_set_transformation_epsilon(ndt, bp.transformation_epsilon);
      }
    }
  }
  while (rc == LGS_OK && !S->failed.load()) {
    const int64_t i = S->next.fetch_add(1);
    if (i >= S->n_pairs) break;
    rc = run_pair(ctx, gicp, ndt, &sub_raw, &sub_ds, S, i);
  }
  if (rc != LGS_OK) {
    std::lock_guard<std::mutex> lk(S->err_mu);
    if (!S->failed.exchange(rc)) S->err = lgs_last_error();
  }
  if (ctx) {
    cudaSetDevice(S->device);
    cudaStreamSynchronize(ctx->stream);
  }
  sub_raw.release();
  sub_ds.release();
  if (gicp) lgs_gicp_destroy(gicp);
  if (ndt) lgs_ndt_destroy(ndt);
  if (ctx) lgs_ctx_destroy(ctx);
}

}  // namespace

extern "C" int lgs_batch_align(int device, void* cuda_stream, const lgs_batch_params* params, int64_t n_pairs, const void* const* scans,
                               const int64_t* n_scan, const void* const* submaps, const int64_t* n_submap, int32_t stride_bytes,
                               const float* guesses16, int32_t pair_id0, lgs_align_result* records, void* records_dev) {
  LGS_REQUIRE(params && records, "null argument");
  LGS_REQUIRE(n_pairs >= 0, "negative pair count");
  LGS_REQUIRE(params->method == LGS_METHOD_GICP || params->method == LGS_METHOD_NDT, "unknown method");
  LGS_REQUIRE(n_pairs == 0 || (scans && n_scan && submaps && n_submap), "null pair arrays");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    set_error("no CUDA device available; this library has no CPU fallback");
    return LGS_ERR_CUDA;
  }
  BatchShared S;
  S.bp = params;
  S.device = device;
  S.n_pairs = n_pairs;
  S.scans = scans;
  S.n_scan = n_scan;
  S.submaps = submaps;
  S.n_submap = n_submap;
  S.stride = stride_bytes;
  S.guesses = guesses16;
  S.pair_id0 = pair_id0;
  S.records = records;
  int W = params->n_workers > 0 ? params->n_workers : 4;
  if (W > n_pairs) W = static_cast<int>(std::max<int64_t>(n_pairs, 1));
  if (W > 32) W = 32;
  std::vector<std::thread> threads;
  for (int w = 0; w < W; w++) threads.emplace_back(worker, &S);
  for (auto& t : threads) t.join();
  if (S.failed.load()) {
    set_error("lgs_batch_align: %s", S.err.c_str());
    return S.failed.load();
  }
  if (records_dev && n_pairs) {
    LGS_CUDA(cudaSetDevice(device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    LGS_CUDA(cudaMemcpyAsync(records_dev, records, static_cast<size_t>(n_pairs) * sizeof(lgs_align_result), cudaMemcpyHostToDevice, st));
    LGS_CUDA(cudaStreamSynchronize(st));
  }
  return LGS_OK;
}
