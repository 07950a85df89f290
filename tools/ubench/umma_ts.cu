// Micro-test: tcgen05.mma with the A operand in TMEM (kind::f16, bf16, M = 128, K = 16) — operand layout (row = lane,
// two bf16 per 32-bit column, K = 16 -> 8 columns), arbitrary D column offsets for small-N MMAs, throughput.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../basic_pitch_b200/csrc umma_ts.cu -o umma_ts
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "tc_ptx.cuh"
using namespace bp;

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return v;
}

// A: [128][32] floats (2 k-steps), B: [N][32] floats; D = A * B^T  [128][N]; D placed at column d_col, A at column a_col
__global__ void __launch_bounds__(128, 1) test_kernel(const float* A, const float* B, float* D, int N, int a_col, int d_col,
                                                      long long* cyc, int iters) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, row = threadIdx.x;
  // B canonical K-major no-swizzle: per k-step [kchunk 2][N][8] bf16 (LBO = N*16 bytes, SBO = 128)
  __nv_bfloat16* sb = reinterpret_cast<__nv_bfloat16*>(smem);
  for (int i = threadIdx.x; i < 2 * 2 * N * 8; i += 128) {
    const int ks = i / (2 * N * 8), r = i % (2 * N * 8), kc = r / (N * 8), n = (r / 8) % N, e = r % 8;
    sb[i] = __float2bfloat16_rn(B[n * 32 + ks * 16 + kc * 8 + e]);
  }
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
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  // zero D area, store A (two k-steps = 16 columns)
  for (int ks = 0; ks < 2; ++ks) {
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) {
      __nv_bfloat162 p = __floats2bfloat162_rn(A[row * 32 + ks * 16 + 2 * j], A[row * 32 + ks * 16 + 2 * j + 1]);
      v[j] = *reinterpret_cast<uint32_t*>(&p);
    }
    tmem_st8(lane_base + a_col + ks * 8, v);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(128, N);
    const uint32_t b0 = smem_u32(smem);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
      for (int ks = 0; ks < 2; ++ks)
        umma_ts(tmem + d_col, tmem + a_col + ks * 8, make_desc(b0 + ks * 2 * N * 16, N * 16, 128), idesc, (it | ks) ? 1u : 0u);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    cyc[0] = clock64() - t0;
  }
  __syncthreads();
  tc_fence_after();
  for (int n = 0; n < N; ++n) D[row * N + n] = __uint_as_float(tmem_ld1(lane_base + d_col + n));
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

static float bfr(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  u &= 0xffff0000u;
  memcpy(&x, &u, 4);
  return x;
}

int main() {
  std::vector<float> A(128 * 32), B(256 * 32);
  srand(1);
  for (auto& v : A) v = bfr((rand() % 2001 - 1000) / 500.f);
  for (auto& v : B) v = bfr((rand() % 2001 - 1000) / 500.f);
  float *dA, *dB, *dD;
  long long* dc;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, B.size() * 4);
  cudaMalloc(&dD, 128 * 256 * 4);
  cudaMalloc(&dc, 8);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int Ns[] = {16, 32, 48, 112, 128};
  const int dcols[] = {256, 258, 260, 264, 266, 272, 286, 330};
  for (int N : Ns)
    for (int dcol : dcols)
      for (int acol : {0, 8, 100}) {
        for (int iters : {1, 200}) {
          cudaMemset(dD, 0, 128 * 256 * 4);
          test_kernel<<<1, 128, 64 * 1024>>>(dA, dB, dD, N, acol, dcol, dc, iters);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("N=%d dcol=%d acol=%d: %s\n", N, dcol, acol, cudaGetErrorString(e));
            return 1;
          }
          std::vector<float> D(128 * N);
          long long cyc;
          cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
          cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
          double maxerr = 0;
          for (int r = 0; r < 128; ++r)
            for (int n = 0; n < N; ++n) {
              double ref = 0;
              for (int k = 0; k < 32; ++k) ref += (double)A[r * 32 + k] * B[n * 32 + k];
              maxerr = fmax(maxerr, fabs(ref * iters - D[r * N + n]) / iters);
            }
          if (iters == 1)
            printf("N=%3d dcol=%3d acol=%3d  max err %.3g ", N, dcol, acol, maxerr);
          else
            printf(" | x200: err %.3g, %.1f cycles/MMA\n", maxerr, (double)cyc / (2 * iters));
        }
      }
  return 0;
}
