// Micro-benchmark: issue rate / throughput of tcgen05.mma (kind::f16, bf16, M = 128, K = 16, operands in shared memory)
// as a function of N, to decide tile shapes of the conv / CQT kernels (is a small-N MMA bound by the A-operand read?).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../basic_pitch_b200/csrc umma_n.cu -o umma_n
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace bp;

template <int N>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters, int a_stride16 /* A start advance per MMA, >>4 */) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
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
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 170 * 1024;
    // descriptors precomputed: the loop body is the MMA itself (a single issuing thread needs ~60 cycles per MMA when it
    // also builds descriptors, which would hide the N dependence)
    uint64_t ad[4], bd[4];
    for (int j = 0; j < 4; ++j) {
      ad[j] = make_desc(a0 + (uint32_t)((j * a_stride16) << 4), 132 * 16, 128);
      bd[j] = make_desc(b0 + j * 4096, 2048, 128);
    }
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) umma_bf16(tmem + (j & 1) * 256, ad[j & 3], bd[j & 3], idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int N>
void run(int iters, int stride) {
  long long* d;
  cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  for (int rep = 0; rep < 2; ++rep) k<N><<<148, 128, 210 * 1024>>>(d, iters, stride);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < 148; ++i) s += h[i];
  printf("N=%3d a_stride=%4d  %.1f cycles/MMA  (floor 128*N/256 = %d)  %s\n", N, stride * 16, s / 148 / iters, N / 2,
         cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  const int iters = 4000;
  for (int stride : {0, 1, 33}) {
    run<32>(iters, stride);
    run<64>(iters, stride);
    run<80>(iters, stride);
    run<96>(iters, stride);
    run<128>(iters, stride);
    run<160>(iters, stride);
    run<192>(iters, stride);
    run<256>(iters, stride);
  }
  return 0;
}
