// Latency of the dependent chains that order-preserving f64 accumulation is made of, one warp, B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o chain chain.cu && ./chain
#include <cstdio>
#include <cuda_runtime.h>

template <int V>
__global__ void k(const float* __restrict__ in, int n, double* out, long long* cyc) {
  const int lane = threadIdx.x;
  double acc = 0.0;
  const long long t0 = clock64();
  if (V == 0) {  // pure DADD chain
    double v = in[lane];
    for (int i = 0; i < n; i++) acc = __dadd_rn(acc, v);
  } else if (V == 1) {  // DMUL + DADD chain (product independent)
    double v = in[lane];
    for (int i = 0; i < n; i++) acc = __dadd_rn(acc, __dmul_rn(v, v + i));
  } else if (V == 2) {  // shuffle + widen + DMUL + DADD per member, rolled loop over the 32 lanes' values
    for (int c = 0; c < n; c += 32) {
      const float x = in[(c + lane) & 1023];
      for (int t = 0; t < 32; t++) {
        const float a = __shfl_sync(0xffffffffu, x, t);
        acc = __dadd_rn(acc, __dmul_rn((double)a, (double)a));
      }
    }
  } else if (V == 3) {  // same, unrolled by 8: products first, additions in sequence
    for (int c = 0; c < n; c += 32) {
      const float x = in[(c + lane) & 1023];
#pragma unroll 1
      for (int t0_ = 0; t0_ < 32; t0_ += 8) {
        double p[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float a = __shfl_sync(0xffffffffu, x, t0_ + j);
          p[j] = __dmul_rn((double)a, (double)a);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) acc = __dadd_rn(acc, p[j]);
      }
    }
  } else if (V == 4) {  // FADD chain for comparison
    float f = 0.f, v = in[lane];
    for (int i = 0; i < n; i++) f = __fadd_rn(f, v);
    acc = f;
  } else if (V == 6 || V == 7) {  // the voxel_sums_kernel body: 3 shuffles, mask select, 2 widenings, DMUL, DADD; 7: widen BEFORE the shuffle
    const int ui = lane < 3 ? lane : (lane < 6 ? 0 : (lane < 8 ? 1 : 2));
    const int vi = lane < 3 ? 3 : (lane < 6 ? lane - 3 : (lane < 8 ? lane - 5 : 2));
    const unsigned ux = ui == 0 ? ~0u : 0u, uy = ui == 1 ? ~0u : 0u, uz = ui == 2 ? ~0u : 0u;
    const unsigned wx = vi == 0 ? ~0u : 0u, wy = vi == 1 ? ~0u : 0u, wz = vi == 2 ? ~0u : 0u, w1 = vi == 3 ? 0x3f800000u : 0u;
    for (int c = 0; c < n; c += 32) {
      const float cx = in[(c + lane) & 1023], cy = in[(c + lane + 7) & 1023], cz = in[(c + lane + 13) & 1023];
#pragma unroll 4
      for (int t = 0; t < 32; t++) {
        const unsigned x = __shfl_sync(0xffffffffu, __float_as_uint(cx), t), y = __shfl_sync(0xffffffffu, __float_as_uint(cy), t),
                       z = __shfl_sync(0xffffffffu, __float_as_uint(cz), t);
        const float uf = __uint_as_float((x & ux) | (y & uy) | (z & uz));
        const float wf = __uint_as_float((x & wx) | (y & wy) | (z & wz) | w1);
        acc = __dadd_rn(acc, __dmul_rn((double)uf, (double)wf));
      }
    }
  } else if (V == 5) {  // DFMA chain
    double v = in[lane];
    for (int i = 0; i < n; i++) acc = __fma_rn(acc, 1.0000001, v);
  }
  const long long t1 = clock64();
  out[lane] = acc;
  if (lane == 0) *cyc = t1 - t0;
}

int main() {
  float* in;
  double* out;
  long long* cyc;
  cudaMalloc(&in, 4096);
  cudaMalloc(&out, 4096);
  cudaMallocManaged(&cyc, 8);
  float h[1024];
  for (int i = 0; i < 1024; i++) h[i] = 1.0f + i * 1e-3f;
  cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
  const int n = 32768;
  const char* names[] = {"DADD chain", "DMUL+DADD chain", "shfl+cvt+DMUL+DADD rolled", "same, unrolled by 8", "FADD chain", "DFMA chain", "voxel_sums body (3 shfl, masks)", "-"};
#define RUN(V)                                                               \
  for (int r = 0; r < 2; r++) {                                              \
    k<V><<<1, 32>>>(in, n, out, cyc);                                        \
    cudaDeviceSynchronize();                                                 \
  }                                                                          \
  printf("%-32s %.1f cycles per element\n", names[V], double(*cyc) / n);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6)
  return 0;
}
