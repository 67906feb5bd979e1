// Launch-floor microbenchmark (sm_100a): how long is a kernel of the derivative kernel's shape (148 CTAs x 384 threads,
// 190 KB dynamic shared memory, ~1.7 KB of parameters) when it does NOTHING, and what do the pieces around the real
// work cost: the last-CTA handshake (store + __threadfence + atomic), a 128-bit store to mapped host memory (the
// result mailbox), and alternating shared-memory carve-outs between consecutive launches.
// Times with CUDA events around each launch (as lgs_ndt_profile does) and around a train of launches.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 launch_floor.cu -o launch_floor
#include <cstdio>
#include <cuda_runtime.h>

struct Params { float a[420]; };  // ~1.7 KB like EvalParams + CellTable

template <int MODE>
__global__ void __launch_bounds__(384, 1) k(Params P, double* partials, unsigned* counter, unsigned long long* host_box, unsigned long long token) {
  extern __shared__ unsigned char smem[];
  __shared__ bool is_last;
  if (MODE >= 1) {
    if (threadIdx.x < 32) partials[blockIdx.x * 32 + threadIdx.x] = P.a[threadIdx.x] + smem[threadIdx.x];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (is_last) {
      __threadfence();
      if (threadIdx.x < 32) {
        double v = 0;
        for (unsigned b = 0; b < gridDim.x; b++) v += __ldcg(partials + b * 32 + threadIdx.x);
        if (threadIdx.x == 0) *counter = 0;
        if (MODE == 2) asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(host_box + 2 * threadIdx.x), "l"((unsigned long long)__double_as_longlong(v)), "l"(token) : "memory");
        else partials[148 * 32 + threadIdx.x] = v;
      }
    }
  } else if (P.a[0] == 123.f) partials[threadIdx.x] = smem[threadIdx.x];
}

__global__ void small_kernel(double* p) { if (p[0] == 42.0) p[1] = 1; }

template <typename F>
void bench(const char* name, F launch, cudaStream_t st, int reps = 300) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 20; i++) launch(i);
  cudaStreamSynchronize(st);
  float per = 0;
  for (int i = 0; i < reps; i++) {   // one launch at a time, events around it, host sync after (the align loop's pattern)
    cudaEventRecord(e0, st); launch(i); cudaEventRecord(e1, st); cudaStreamSynchronize(st);
    float ms; cudaEventElapsedTime(&ms, e0, e1); per += ms;
  }
  cudaEventRecord(e0, st);
  for (int i = 0; i < reps; i++) launch(i);   // back-to-back train
  cudaEventRecord(e1, st); cudaStreamSynchronize(st);
  float train; cudaEventElapsedTime(&train, e0, e1);
  printf("%-58s single %.2f us   back-to-back %.2f us\n", name, 1e3 * per / reps, 1e3 * train / reps);
}

int main() {
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  double* partials; cudaMalloc(&partials, 149 * 32 * 8 + 1024); cudaMemset(partials, 0, 149 * 32 * 8);
  unsigned* counter; cudaMalloc(&counter, 64); cudaMemset(counter, 0, 64);
  unsigned long long* box; cudaHostAlloc(&box, 4096, cudaHostAllocMapped); unsigned long long* dbox; cudaHostGetDevicePointer(&dbox, box, 0);
  Params P = {};
  const int big = 190 * 1024;
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  bench("empty, 148x384, no dynamic smem", [&](int) { k<0><<<148, 384, 0, st>>>(P, partials, counter, dbox, 1); }, st);
  bench("empty, 148x384, 190 KB dynamic smem", [&](int) { k<0><<<148, 384, big, st>>>(P, partials, counter, dbox, 1); }, st);
  bench("last-CTA handshake, result to device memory, 190 KB", [&](int) { k<1><<<148, 384, big, st>>>(P, partials, counter, dbox, 1); }, st);
  bench("last-CTA handshake, result to host mailbox, 190 KB", [&](int i) { k<2><<<148, 384, big, st>>>(P, partials, counter, dbox, i); }, st);
  bench("handshake + mailbox, alternating with a 1-CTA kernel", [&](int i) { k<2><<<148, 384, big, st>>>(P, partials, counter, dbox, i); small_kernel<<<1, 32, 0, st>>>(partials); }, st);
  bench("empty, 148x384, 96 KB dynamic smem", [&](int) { k<0><<<148, 384, 96 * 1024, st>>>(P, partials, counter, dbox, 1); }, st);
  bench("empty, 296x192, 95 KB dynamic smem", [&](int) { k<0><<<296, 192, 95 * 1024, st>>>(P, partials, counter, dbox, 1); }, st);
  // host round trip: launch + spin on the mailbox token (no events, wall clock)
  {
    volatile unsigned long long* vb = box;
    for (int i = 0; i < 50; i++) { k<2><<<148, 384, big, st>>>(P, partials, counter, dbox, 1000 + i); while (vb[1] != 1000ull + i) {} }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaStreamSynchronize(st);
    cudaEventRecord(e0, st);
    for (int i = 0; i < 300; i++) { k<2><<<148, 384, big, st>>>(P, partials, counter, dbox, 5000 + i); while (vb[1] != 5000ull + i) {} }
    cudaEventRecord(e1, st); cudaStreamSynchronize(st);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-58s %.2f us per launch + mailbox round trip\n", "launch -> spin on mailbox -> next launch", 1e3 * ms / 300);
  }
  return 0;
}
