// Build and query kernels of the exact nearest-neighbour index (see nn.cuh).
#include <cub/cub.cuh>

#include <cfloat>

#include "nn.cuh"

namespace lgs {

struct BoxAcc {
  unsigned mn[3];
  unsigned mx[3];
};

__global__ void nn_box_init_kernel(BoxAcc* acc) {
  if (threadIdx.x < 3) {
    acc->mn[threadIdx.x] = 0xFFFFFFFFu;
    acc->mx[threadIdx.x] = 0u;
  }
}

__global__ void __launch_bounds__(256) nn_bbox_kernel(const float4* __restrict__ pts, int64_t n, BoxAcc* __restrict__ acc) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 p = pts[i];
    mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
    mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
    mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
  }
  for (int off = 16; off > 0; off >>= 1)
    for (int a = 0; a < 3; a++) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], off));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], off));
    }
  if ((threadIdx.x & 31) == 0)
    for (int a = 0; a < 3; a++) {
      atomicMin(&acc->mn[a], enc_f(mn[a]));
      atomicMax(&acc->mx[a], enc_f(mx[a]));
    }
}

__device__ __forceinline__ unsigned expand10(unsigned v) {
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}

__global__ void __launch_bounds__(256) nn_morton_kernel(const float4* __restrict__ pts, int64_t n, const BoxAcc* __restrict__ acc,
                                                       unsigned* __restrict__ codes, unsigned* __restrict__ perm) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  float lo[3], hi[3];
  for (int a = 0; a < 3; a++) {
    lo[a] = dec_f_host(acc->mn[a]);
    hi[a] = dec_f_host(acc->mx[a]);
  }
  const float c[3] = {p.x, p.y, p.z};
  unsigned q[3];
  for (int a = 0; a < 3; a++) {
    float ext = hi[a] - lo[a];
    float t = ext > 0.f ? (c[a] - lo[a]) / ext : 0.f;
    t = fminf(fmaxf(t * 1024.0f, 0.0f), 1023.0f);
    q[a] = static_cast<unsigned>(t);
  }
  codes[i] = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
  perm[i] = static_cast<unsigned>(i);
}

__global__ void __launch_bounds__(256) nn_gather_kernel(const float4* __restrict__ pts, const unsigned* __restrict__ perm, int64_t n,
                                                       float4* __restrict__ spts) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  unsigned src = perm[i];
  float4 p = pts[src];
  p.w = __int_as_float(static_cast<int>(src));
  spts[i] = p;
}

__global__ void __launch_bounds__(256) nn_leaf_box_kernel(const float4* __restrict__ spts, int n, int n_leaves_p2, float4* __restrict__ bmin,
                                                         float4* __restrict__ bmax) {
  int leaf = blockIdx.x * blockDim.x + threadIdx.x;
  if (leaf >= n_leaves_p2) return;
  const float inf = __int_as_float(0x7f800000);
  float mn[3] = {inf, inf, inf}, mx[3] = {-inf, -inf, -inf};
  int b = leaf * kLeafSize, e = min(b + kLeafSize, n);
  for (int j = b; j < e; j++) {
    float4 p = spts[j];
    mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
    mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
    mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
  }
  int node = n_leaves_p2 - 1 + leaf;
  bmin[node] = make_float4(mn[0], mn[1], mn[2], 0.f);
  bmax[node] = make_float4(mx[0], mx[1], mx[2], 0.f);
}

// one level of the bottom-up merge: nodes [first, first + count)
__global__ void __launch_bounds__(256) nn_merge_level_kernel(int first, int count, float4* __restrict__ bmin, float4* __restrict__ bmax) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  int node = first + k;
  int l = 2 * node + 1, r = l + 1;
  float4 a = bmin[l], b = bmin[r];
  bmin[node] = make_float4(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), 0.f);
  a = bmax[l];
  b = bmax[r];
  bmax[node] = make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), 0.f);
}

int NNIndex::build(lgs_ctx* ctx, const float4* pts, int64_t n_in) {
  LGS_REQUIRE(n_in >= 0 && n_in < (int64_t(1) << 31), "point count out of range");
  n = n_in;
  n_leaves_p2 = 1;
  if (n == 0) return LGS_OK;
  cudaStream_t st = ctx->stream;
  const int n_leaves = static_cast<int>((n + kLeafSize - 1) / kLeafSize);
  while (n_leaves_p2 < n_leaves) n_leaves_p2 <<= 1;
  const size_t n_nodes = 2 * static_cast<size_t>(n_leaves_p2) - 1;
  LGS_TRY(spts.reserve(static_cast<size_t>(n) * 16));
  LGS_TRY(bmin.reserve(n_nodes * 16));
  LGS_TRY(bmax.reserve(n_nodes * 16));
  LGS_TRY(codes.reserve(static_cast<size_t>(n) * 4));
  LGS_TRY(codes_alt.reserve(static_cast<size_t>(n) * 4));
  LGS_TRY(perm.reserve(static_cast<size_t>(n) * 4));
  LGS_TRY(perm_alt.reserve(static_cast<size_t>(n) * 4));
  LGS_TRY(small.reserve(64));
  BoxAcc* acc = small.as<BoxAcc>();
  nn_box_init_kernel<<<1, 32, 0, st>>>(acc);
  nn_bbox_kernel<<<std::min(grid_for(n, 256), kNumSMs * 8), 256, 0, st>>>(pts, n, acc);
  cub::DoubleBuffer<unsigned> dk(codes.as<unsigned>(), codes_alt.as<unsigned>());
  cub::DoubleBuffer<unsigned> dv(perm.as<unsigned>(), perm_alt.as<unsigned>());
  nn_morton_kernel<<<grid_for(n, 256), 256, 0, st>>>(pts, n, acc, dk.Current(), dv.Current());
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, static_cast<int>(n), 0, 30, st);
  LGS_TRY(ctx->cub_tmp.reserve(tb));
  cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tb, dk, dv, static_cast<int>(n), 0, 30, st);
  nn_gather_kernel<<<grid_for(n, 256), 256, 0, st>>>(pts, dv.Current(), n, spts.as<float4>());
  nn_leaf_box_kernel<<<grid_for(n_leaves_p2, 256), 256, 0, st>>>(spts.as<float4>(), static_cast<int>(n), n_leaves_p2, bmin.as<float4>(),
                                                                bmax.as<float4>());
  ctx->launches += 5 + 6;
  for (int count = n_leaves_p2 / 2; count >= 1; count >>= 1) {
    nn_merge_level_kernel<<<grid_for(count, 256), 256, 0, st>>>(count - 1, count, bmin.as<float4>(), bmax.as<float4>());
    ctx->launches++;
  }
  LGS_CUDA(cudaGetLastError());
  return LGS_OK;
}

// ---------------------------------------------------------------------------------------------
// fitness: pcl::Registration::getFitnessScore

constexpr int kFitBlock = 128;

__global__ void __launch_bounds__(kFitBlock) nn_fitness_kernel(NNView v, const float4* __restrict__ src, int n, const float* __restrict__ Tdev,
                                                              double max_range, double* __restrict__ partials, double* __restrict__ result,
                                                              unsigned* __restrict__ counter) {
  __shared__ float T[16];
  if (threadIdx.x < 16) T[threadIdx.x] = Tdev[threadIdx.x];
  __syncthreads();
  double sum = 0.0, cnt = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = src[i];
    const float3 q = transform_pcl(T, p.x, p.y, p.z);
    float d;
    int id;
    nn_search1(v, q.x, q.y, q.z, d, id);
    if (static_cast<double>(d) <= max_range) {
      sum += static_cast<double>(d);
      cnt += 1.0;
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, off);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
  }
  __shared__ double ssum[kFitBlock / 32], scnt[kFitBlock / 32];
  __shared__ bool is_last;
  if ((threadIdx.x & 31) == 0) {
    ssum[threadIdx.x >> 5] = sum;
    scnt[threadIdx.x >> 5] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0, c = 0;
    for (int w = 0; w < kFitBlock / 32; w++) {
      s += ssum[w];
      c += scnt[w];
    }
    partials[2 * blockIdx.x] = s;
    partials[2 * blockIdx.x + 1] = c;
    __threadfence();
    is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    double s = 0, c = 0;
    for (unsigned b = 0; b < gridDim.x; b++) {
      s += partials[2 * b];
      c += partials[2 * b + 1];
    }
    result[0] = s;
    result[1] = c;
    *counter = 0;
  }
}

int nn_fitness(lgs_ctx* ctx, const NNIndex& index, const float4* src, int64_t n_src, const float* T16, double max_range, double* fitness) {
  *fitness = std::numeric_limits<double>::max();
  if (n_src == 0 || index.n == 0) return LGS_OK;
  cudaStream_t st = ctx->stream;
  const int grid = std::max(1, std::min(grid_for(n_src, kFitBlock), kNumSMs * 16));
  LGS_TRY(ctx->tmp[0].reserve(static_cast<size_t>(grid) * 16 + 256));
  double* partials = ctx->tmp[0].as<double>();
  double* result = partials + 2 * grid;
  unsigned* counter = reinterpret_cast<unsigned*>(result + 2);
  float* Tdev = reinterpret_cast<float*>(result + 4);
  LGS_TRY(ctx->pin_up.reserve(64));
  memcpy(ctx->pin_up.p, T16, 64);
  LGS_CUDA(cudaMemsetAsync(counter, 0, 4, st));
  LGS_CUDA(cudaMemcpyAsync(Tdev, ctx->pin_up.p, 64, cudaMemcpyHostToDevice, st));
  nn_fitness_kernel<<<grid, kFitBlock, 0, st>>>(index.view(), src, static_cast<int>(n_src), Tdev, max_range, partials, result, counter);
  ctx->launches++;
  LGS_CUDA(cudaGetLastError());
  LGS_TRY(ctx->pin.reserve(64));
  LGS_CUDA(cudaMemcpyAsync(ctx->pin.p, result, 16, cudaMemcpyDeviceToHost, st));
  LGS_CUDA(cudaStreamSynchronize(st));
  const double s = ctx->pin.as<double>()[0], c = ctx->pin.as<double>()[1];
  if (c > 0) *fitness = s / c;
  return LGS_OK;
}

// ---------------------------------------------------------------------------------------------
// k-NN

template <bool SELF>
__global__ void __launch_bounds__(128) nn_knn_kernel(NNView v, const float4* __restrict__ queries, int m, int k, int* __restrict__ out_idx,
                                                    float* __restrict__ out_d2) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  float4 qp = SELF ? __ldg(v.spts + t) : queries[t];
  const int row = SELF ? __float_as_int(qp.w) : t;
  const float qx = qp.x, qy = qp.y, qz = qp.z;
  float bd[kMaxK];
  int bi[kMaxK];
  int found = 0;
  const float inf = __int_as_float(0x7f800000);
  float worst = inf;
  int worst_i = 0x7fffffff;
  const int first_leaf = v.n_leaves_p2 - 1;
  int stack[48];
  float sdist[48];
  int sp = 1;
  stack[0] = 0;
  sdist[0] = nn_box_dist2(qx, qy, qz, __ldg(v.bmin), __ldg(v.bmax));
  while (sp > 0) {
    --sp;
    const int node = stack[sp];
    if (sdist[sp] > worst) continue;
    if (node >= first_leaf) {
      const int b = (node - first_leaf) * kLeafSize;
      const int e = min(b + kLeafSize, v.n);
      for (int j = b; j < e; j++) {
        const float4 p = __ldg(v.spts + j);
        const float d = nn_dist2(qx, qy, qz, p);
        const int oi = __float_as_int(p.w);
        if (found < k || d < worst || (d == worst && oi < worst_i)) {
          int pos = found < k ? found : k - 1;
          if (found < k) found++;
          while (pos > 0 && (d < bd[pos - 1] || (d == bd[pos - 1] && oi < bi[pos - 1]))) {
            bd[pos] = bd[pos - 1];
            bi[pos] = bi[pos - 1];
            --pos;
          }
          bd[pos] = d;
          bi[pos] = oi;
          if (found == k) {
            worst = bd[k - 1];
            worst_i = bi[k - 1];
          }
        }
      }
    } else {
      const int l = 2 * node + 1, r = l + 1;
      const float dl = nn_box_dist2(qx, qy, qz, __ldg(v.bmin + l), __ldg(v.bmax + l));
      const float dr = nn_box_dist2(qx, qy, qz, __ldg(v.bmin + r), __ldg(v.bmax + r));
      if (dl <= dr) {
        if (dr <= worst) { stack[sp] = r; sdist[sp++] = dr; }
        if (dl <= worst) { stack[sp] = l; sdist[sp++] = dl; }
      } else {
        if (dl <= worst) { stack[sp] = l; sdist[sp++] = dl; }
        if (dr <= worst) { stack[sp] = r; sdist[sp++] = dr; }
      }
    }
  }
  for (int j = 0; j < k; j++) {
    out_idx[static_cast<size_t>(row) * k + j] = j < found ? bi[j] : -1;
    if (out_d2) out_d2[static_cast<size_t>(row) * k + j] = j < found ? bd[j] : 0.f;
  }
}

int nn_self_knn(lgs_ctx* ctx, const NNIndex& index, int k, int* out_idx_dev, float* out_d2_dev) {
  LGS_REQUIRE(k >= 1 && k <= kMaxK, "k must be in [1, 32]");
  if (index.n == 0) return LGS_OK;
  nn_knn_kernel<true><<<grid_for(index.n, 128), 128, 0, ctx->stream>>>(index.view(), nullptr, static_cast<int>(index.n), k, out_idx_dev, out_d2_dev);
  ctx->launches++;
  LGS_CUDA(cudaGetLastError());
  return LGS_OK;
}

int nn_knn(lgs_ctx* ctx, const NNIndex& index, const float4* queries, int64_t m, int k, int* out_idx_dev, float* out_d2_dev) {
  LGS_REQUIRE(k >= 1 && k <= kMaxK, "k must be in [1, 32]");
  if (m == 0) return LGS_OK;
  if (index.n == 0) {
    LGS_CUDA(cudaMemsetAsync(out_idx_dev, 0xFF, static_cast<size_t>(m) * k * 4, ctx->stream));
    if (out_d2_dev) LGS_CUDA(cudaMemsetAsync(out_d2_dev, 0, static_cast<size_t>(m) * k * 4, ctx->stream));
    return LGS_OK;
  }
  nn_knn_kernel<false><<<grid_for(m, 128), 128, 0, ctx->stream>>>(index.view(), queries, static_cast<int>(m), k, out_idx_dev, out_d2_dev);
  ctx->launches++;
  LGS_CUDA(cudaGetLastError());
  return LGS_OK;
}

}  // namespace lgs

extern "C" int lgs_knn(lgs_ctx* ctx, const void* pts, int64_t n, int32_t stride_bytes, const void* queries, int64_t m, int32_t qstride_bytes, int32_t k,
                       int32_t* idx, float* d2) {
  using namespace lgs;
  LGS_REQUIRE(ctx && idx, "null argument");
  LGS_REQUIRE(k >= 1 && k <= kMaxK, "k must be in [1, 32]");
  LGS_TRY(use_device(ctx));
  DevBuf cloud, q, oi, od;
  NNIndex index;
  int rc = LGS_OK;
  do {
    if ((rc = upload_cloud(ctx, pts, n, stride_bytes, &cloud)) != LGS_OK) break;
    if ((rc = upload_cloud(ctx, queries, m, qstride_bytes, &q)) != LGS_OK) break;
    if ((rc = index.build(ctx, cloud.as<float4>(), n)) != LGS_OK) break;
    if ((rc = oi.reserve(static_cast<size_t>(m > 0 ? m : 1) * k * 4)) != LGS_OK) break;
    if ((rc = od.reserve(static_cast<size_t>(m > 0 ? m : 1) * k * 4)) != LGS_OK) break;
    if ((rc = nn_knn(ctx, index, q.as<float4>(), m, k, oi.as<int>(), od.as<float>())) != LGS_OK) break;
    if (m) {
      cudaMemcpyAsync(idx, oi.p, static_cast<size_t>(m) * k * 4, cudaMemcpyDeviceToHost, ctx->stream);
      if (d2) cudaMemcpyAsync(d2, od.p, static_cast<size_t>(m) * k * 4, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      set_error("lgs_knn: stream synchronize failed");
      rc = LGS_ERR_CUDA;
    }
  } while (false);
  cloud.release();
  q.release();
  oi.release();
  od.release();
  index.release();
  return rc;
}
