// Ingest of the sensor_msgs/PointCloud2 byte layout: the raw message payload goes to the GPU as it is and is
// repacked there into packed float4 xyzi, replacing the host-side pcl::fromROSMsg of the nodes (PPF:65-70, LSM:122-130,
// GBS:287-295).  Field semantics follow pcl::fromROSMsg<pcl::PointXYZI>: a field is mapped
// only when its name AND datatype match the point type (x, y, z, intensity as FLOAT32); an intensity field of another
// datatype is not mapped and the point keeps its default intensity 0.
//
// Kernel: records are unaligned in general (Velodyne's point_step is 22), so a tile of 256 records is brought into
// shared memory with one 1-D bulk async copy (cp.async.bulk, completion on an mbarrier; 256 * point_step bytes is a
// multiple of 16 for every point_step) while the previous tile is being extracted: two stages per CTA, persistent grid.
// Extraction reads the four fields byte-wise from shared memory and stores one coalesced float4 per point.
// Algorithmic bytes: n * (point_step + 16); HBM-bound.
#include <algorithm>

#include "common.cuh"

namespace lgs {

constexpr int kTileRecs = 256;

struct Pc2Fields {
  int point_step;
  int off[4];  // x, y, z, intensity (-1: none)
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float load_f32_bytes(const unsigned char* p) {
  const unsigned u = static_cast<unsigned>(p[0]) | (static_cast<unsigned>(p[1]) << 8) | (static_cast<unsigned>(p[2]) << 16) | (static_cast<unsigned>(p[3]) << 24);
  return __uint_as_float(u);
}

// raw holds n records of point_step bytes and is readable up to the next multiple of 16 bytes past its end
__global__ void __launch_bounds__(kTileRecs) pc2_repack_kernel(const unsigned char* __restrict__ raw, int64_t n, Pc2Fields F, float4* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char pc2_smem[];
  __shared__ __align__(8) unsigned long long bar[2];
  const size_t tile_bytes = static_cast<size_t>(kTileRecs) * F.point_step;
  const int64_t ntiles = (n + kTileRecs - 1) / kTileRecs;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int64_t tile, int stage) {
    const int64_t recs = min(static_cast<int64_t>(kTileRecs), n - tile * kTileRecs);
    const unsigned bytes = static_cast<unsigned>((recs * F.point_step + 15) & ~static_cast<int64_t>(15));
    mbar_expect_tx(&bar[stage], bytes);
    bulk_g2s(pc2_smem + stage * tile_bytes, raw + tile * tile_bytes, bytes, &bar[stage]);
  };
  int64_t tile = blockIdx.x;
  if (threadIdx.x == 0 && tile < ntiles) issue(tile, 0);
  unsigned phase[2] = {0, 0};
  int stage = 0;
  for (; tile < ntiles; tile += gridDim.x, stage ^= 1) {
    const int64_t next = tile + gridDim.x;
    if (threadIdx.x == 0 && next < ntiles) issue(next, stage ^ 1);  // the other stage was drained before the last barrier
    mbar_wait(&bar[stage], phase[stage]);
    phase[stage] ^= 1;
    const int64_t i = tile * kTileRecs + threadIdx.x;
    if (i < n) {
      const unsigned char* rec = pc2_smem + stage * tile_bytes + static_cast<size_t>(threadIdx.x) * F.point_step;
      float4 v;
      v.x = load_f32_bytes(rec + F.off[0]);
      v.y = load_f32_bytes(rec + F.off[1]);
      v.z = load_f32_bytes(rec + F.off[2]);
      v.w = F.off[3] >= 0 ? load_f32_bytes(rec + F.off[3]) : 0.0f;
      out[i] = v;
    }
    __syncthreads();  // everyone is done with this stage before it is refilled
  }
}

}  // namespace lgs

using namespace lgs;

extern "C" {

int lgs_cloud_from_pointcloud2(lgs_ctx* ctx, const void* data, const lgs_pc2_layout* L, float* out_dev, int64_t* n_points) {
  LGS_NVTX("lgs_cloud_from_pointcloud2");
  LGS_REQUIRE(ctx && L && n_points, "null argument");
  LGS_REQUIRE(L->point_step >= 12 && L->point_step <= 256, "point_step must be in [12, 256]");
  LGS_REQUIRE(L->is_bigendian == 0, "big-endian PointCloud2 payloads are not supported");
  const int64_t w = L->width, h = L->height;
  const int64_t n = w * h;
  *n_points = n;
  LGS_REQUIRE(L->row_step == 0 || static_cast<int64_t>(L->row_step) == w * L->point_step, "row_step must equal width * point_step (no row padding)");
  const int32_t offs[3] = {L->offset_x, L->offset_y, L->offset_z};
  for (int a = 0; a < 3; a++) {
    LGS_REQUIRE(L->datatype_xyz == LGS_PC2_FLOAT32, "x, y, z must be FLOAT32 fields (pcl::fromROSMsg maps nothing else onto PointXYZI)");
    LGS_REQUIRE(offs[a] >= 0 && offs[a] + 4 <= static_cast<int32_t>(L->point_step), "x/y/z offset outside the record");
  }
  Pc2Fields F;
  F.point_step = static_cast<int>(L->point_step);
  F.off[0] = L->offset_x;
  F.off[1] = L->offset_y;
  F.off[2] = L->offset_z;
  // an intensity field whose datatype differs from the point type's is not mapped by fromROSMsg: intensity stays 0
  F.off[3] = (L->offset_intensity >= 0 && L->datatype_intensity == LGS_PC2_FLOAT32) ? L->offset_intensity : -1;
  if (F.off[3] >= 0) LGS_REQUIRE(F.off[3] + 4 <= F.point_step, "intensity offset outside the record");
  if (n == 0) return LGS_OK;
  LGS_REQUIRE(data && out_dev, "null cloud");
  LGS_TRY(use_device(ctx));
  const size_t bytes = static_cast<size_t>(n) * F.point_step;
  LGS_TRY(ctx->raw.reserve(bytes + 16));
  LGS_CUDA(cudaMemcpyAsync(ctx->raw.p, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
  const int smem = 2 * kTileRecs * F.point_step;
  static int smem_set_dev[64] = {};  // dynamic shared memory above 48 KB needs the opt-in attribute, per device
  int& smem_set = smem_set_dev[ctx->device & 63];
  if (smem > 48 * 1024 && smem > smem_set) {
    LGS_CUDA(cudaFuncSetAttribute(pc2_repack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kTileRecs * 256));
    smem_set = 2 * kTileRecs * 256;
  }
  const int64_t ntiles = (n + kTileRecs - 1) / kTileRecs;
  const int grid = static_cast<int>(std::min<int64_t>(ntiles, static_cast<int64_t>(kNumSMs) * 4));
  pc2_repack_kernel<<<grid, kTileRecs, smem, ctx->stream>>>(ctx->raw.as<unsigned char>(), n, F, reinterpret_cast<float4*>(out_dev));
  ctx->launches++;
  LGS_CUDA(cudaGetLastError());
  return LGS_OK;
}

}  // extern "C"
