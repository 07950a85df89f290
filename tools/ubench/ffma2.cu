// Micro-benchmark of the FP32 FMA forms used by the fused epilogues (sm_100a).  FMA/clk/SM for
//  A scalar FFMA  acc = o * UR + acc        (uniform-register weight, what nvcc emits for __constant__ weights)
//  B scalar FFMA  acc = o * R + acc         (weight in a vector register)
//  C FFMA2        acc2 = {o,o} * UR2 + acc2 (broadcast scalar x packed uniform weight pair)
//  D FFMA2        acc2 = o2 * {UR,UR} + acc2 (packed inputs x broadcast uniform weight)
//  E FFMA2        acc2 = o2 * R2 + acc2     (all packed vector registers)
#include <cstdio>
#include <cuda_runtime.h>
__constant__ float c_w[64];
constexpr int ITER = 2048, ACC = 16;
__device__ __forceinline__ unsigned long long pk(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float fma_s(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void fma_p(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
template <int MODE>
__global__ void k(float* out, const float* in) {
  float o[4];
  for (int j = 0; j < 4; ++j) o[j] = in[threadIdx.x + j * 32];
  float r[4];
  for (int j = 0; j < 4; ++j) r[j] = in[threadIdx.x + 128 + j * 32];
  float s = 0;
  if (MODE <= 1) {
    float acc[ACC];
    for (int j = 0; j < ACC; ++j) acc[j] = threadIdx.x + j;
    for (int i = 0; i < ITER; ++i) {
#pragma unroll
      for (int j = 0; j < ACC; ++j) acc[j] = fma_s(o[j & 3], MODE == 0 ? c_w[j] : r[(j >> 2) & 3], acc[j]);
    }
    for (int j = 0; j < ACC; ++j) s += acc[j];
  } else {
    unsigned long long acc[ACC / 2], o2[4], r2[4];
    for (int j = 0; j < 4; ++j) o2[j] = MODE == 2 ? pk(o[j], o[j]) : pk(o[j], o[(j + 1) & 3]);
    for (int j = 0; j < 4; ++j) r2[j] = pk(r[j], r[(j + 1) & 3]);
    for (int j = 0; j < ACC / 2; ++j) acc[j] = pk(threadIdx.x + j, threadIdx.x - j);
    for (int i = 0; i < ITER; ++i) {
#pragma unroll
      for (int j = 0; j < ACC / 2; ++j) {
        if (MODE == 2) fma_p(acc[j], o2[j & 3], pk(c_w[2 * j], c_w[2 * j + 1]));
        if (MODE == 3) fma_p(acc[j], o2[j & 3], pk(c_w[j], c_w[j]));
        if (MODE == 4) fma_p(acc[j], o2[j & 3], r2[(j >> 1) & 3]);
      }
    }
    for (int j = 0; j < ACC / 2; ++j) {
      float2 t = *reinterpret_cast<float2*>(&acc[j]);
      s += t.x + t.y;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
static void run(const char* name, float* out, const float* in, int blocks, int threads) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k<MODE><<<blocks, threads>>>(out, in);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int r = 0; r < 5; ++r) k<MODE><<<blocks, threads>>>(out, in);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  int clk, sms;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double fma = 5.0 * ITER * ACC * blocks * threads;
  printf("%-44s %8.3f ms %6.1f FMA/clk/SM\n", name, ms, fma / (ms * 1e-3) / (clk * 1e3) / sms);
}
int main() {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int threads = 256, blocks = sms * 8;
  float *out, *in;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaMalloc(&in, sizeof(float) * 1024);
  cudaMemset(in, 0, sizeof(float) * 1024);
  float hw[64];
  for (int i = 0; i < 64; ++i) hw[i] = 1.0f + 1e-7f * i;
  cudaMemcpyToSymbol(c_w, hw, sizeof(hw));
  run<0>("A FFMA  R = R * UR + R", out, in, blocks, threads);
  run<1>("B FFMA  R = R * R + R", out, in, blocks, threads);
  run<2>("C FFMA2 R2 = R.bcast * UR2 + R2", out, in, blocks, threads);
  run<3>("D FFMA2 R2 = R2 * UR.bcast + R2", out, in, blocks, threads);
  run<4>("E FFMA2 R2 = R2 * R2 + R2", out, in, blocks, threads);
  return 0;
}
