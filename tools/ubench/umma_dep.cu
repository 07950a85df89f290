// Micro-benchmark: do consecutive small-N tcgen05.mma (A in tensor memory) stall on each other only when their D column
// windows overlap?  One issuing thread, N = 32, precomputed descriptors; `nwin` disjoint windows used round-robin.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../basic_pitch_b200/csrc umma_dep.cu -o umma_dep
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace bp;

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.eq.b32 p, 0, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc)
      : "memory");
}

template <int N, int NWIN, int STRIDE, bool SS>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc(128, N);
    const uint64_t bd = make_desc(smem_u32(smem), N * 16, 128);
    const uint64_t ad = make_desc(smem_u32(smem) + 32768, 128 * 16, 128);
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t d = tmem + 256 + (j % NWIN) * STRIDE;
        if (SS)
          umma_bf16(d, ad, bd, idesc, 1u);
        else
          umma_ts(d, tmem + (j & 7) * 8, bd, idesc);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int N, int NWIN, int STRIDE, bool SS>
void run(const char* what) {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k<N, NWIN, STRIDE, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 4000;
  for (int rep = 0; rep < 2; ++rep) k<N, NWIN, STRIDE, SS><<<1, 128, 64 * 1024>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-44s N=%3d windows=%d stride=%3d  %.1f cycles/MMA  %s\n", what, N, NWIN, STRIDE, (double)h / iters, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<32, 1, 0, false>("TS same D");
  run<32, 2, 32, false>("TS two disjoint windows");
  run<32, 2, 10, false>("TS two overlapping windows (stride 10)");
  run<32, 4, 32, false>("TS four disjoint windows");
  run<32, 8, 32, false>("TS eight disjoint windows");
  run<16, 1, 0, false>("TS same D");
  run<16, 4, 16, false>("TS four disjoint windows");
  run<64, 1, 0, false>("TS same D");
  run<64, 2, 64, false>("TS two disjoint");
  run<128, 1, 0, false>("TS same D");
  run<128, 2, 128, false>("TS two disjoint");
  run<32, 1, 0, true>("SS same D");
  run<32, 4, 32, true>("SS four disjoint windows");
  run<128, 1, 0, true>("SS same D");
  return 0;
}
