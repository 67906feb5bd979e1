// Pipe-throughput microbenchmark for the instructions the NDT derivative kernel is made of (sm_100a):
// scalar FMUL/FADD, packed FMUL2/FFMA2 (mul.rn.f32x2 / fma.rn.f32x2), F2F.F64.F32, DADD and a term-like mix.
// Prints warp-instructions per clock per SM for each.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false pipes.cu -o pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 c; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b)); return c; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 d) { u64 c; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(c) : "l"(a), "l"(b), "l"(d)); return c; }

constexpr int ITERS = 4096;
constexpr int CH = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed, u64 one, int iters) {
  float a[CH], b[CH];
  double d[CH];
  u64 p[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) { a[c] = seed + c + threadIdx.x; b[c] = 1.0f + 1e-7f * c; d[c] = c; p[c] = pk(a[c], b[c]); }
  const u64 q = pk(1.0000001f, 0.9999999f);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < CH; c++) {
      if (MODE == 0) { a[c] = __fmul_rn(a[c], b[c]); }                              // FMUL
      if (MODE == 1) { a[c] = __fadd_rn(__fmul_rn(a[c], b[c]), b[c]); }             // FMUL + FADD (unfused)
      if (MODE == 2) { p[c] = mul2(p[c], q); }                                      // FMUL2
      if (MODE == 3) { p[c] = fma2(p[c], one, q); }                                 // FFMA2 (packed add via *1.0)
      if (MODE == 4) { d[c] += static_cast<double>(a[c]); a[c] = __fmul_rn(a[c], b[c]); }  // F2F + DADD + FMUL
      if (MODE == 5) { d[c] += 1.25; }                                              // DADD
      if (MODE == 6) { d[c] += static_cast<double>(__int_as_float(__float_as_int(a[c]) + it)); }  // F2F + DADD + IADD
      if (MODE == 7) { p[c] = mul2(p[c], q); d[c] += 1.25; }                        // FMUL2 + DADD
      if (MODE == 8) { a[c] = fmaf(a[c], b[c], b[c]); }                             // FFMA
    }
  }
  float s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) { float x, y; upk(p[c], x, y); s += a[c] + static_cast<float>(d[c]) + x + y; }
  if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, int instr_per_chain, int sms, double mhz) {
  float* out;
  cudaMalloc(&out, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int wps : {4, 8, 16, 32}) {  // warps per SM
    int ctas = sms * (wps * 32 / 256 > 0 ? wps * 32 / 256 : 1);
    int thr = wps * 32 >= 256 ? 256 : wps * 32;
    k<MODE><<<ctas, thr>>>(out, 1.0f, 0x3f8000003f800000ull, ITERS);
    cudaEventRecord(e0);
    k<MODE><<<ctas, thr>>>(out, 1.0f, 0x3f8000003f800000ull, ITERS);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double winstr = double(ITERS) * CH * instr_per_chain * wps;  // per SM
    double clk = ms * 1e-3 * mhz * 1e6;
    printf("%-28s warps/SM %2d  %.3f ms  %.2f warp-instr/clk/SM\n", name, wps, ms, winstr / clk);
  }
  cudaFree(out);
}

int main() {
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double mhz = clk_khz / 1000.0;
  printf("%s, %d SMs, %.0f MHz (assumed for the per-clock figures)\n", pr.name, pr.multiProcessorCount, mhz);
  int sms = pr.multiProcessorCount;
  run<0>("FMUL", 1, sms, mhz);
  run<8>("FFMA", 1, sms, mhz);
  run<1>("FMUL+FADD", 2, sms, mhz);
  run<2>("FMUL2", 1, sms, mhz);
  run<3>("FFMA2(x*1+y)", 1, sms, mhz);
  run<4>("F2F.F64.F32+DADD+FMUL", 3, sms, mhz);
  run<5>("DADD", 1, sms, mhz);
  run<6>("F2F+DADD+IADD", 3, sms, mhz);
  run<7>("FMUL2+DADD", 2, sms, mhz);
  return 0;
}
