// Micro-benchmark: FP32 FMA issue rate per SM for (a) 3-register FFMA, (b) FFMA with a constant-bank operand,
// (c) packed fma.rn.f32x2.  Prints FMA/clk/SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 ffma.cu -o ffma
#include <cstdio>
#include <cuda_runtime.h>
__constant__ float c_w[64];
constexpr int ITER = 4096, ACC = 16;

__global__ void k_reg(float* out, float x, float w) {
  float acc[ACC];
  float ws[4] = {w, w * 1.1f, w * 1.2f, w * 1.3f};
  for (int j = 0; j < ACC; ++j) acc[j] = threadIdx.x + j;
  for (int i = 0; i < ITER; ++i) {
#pragma unroll
    for (int j = 0; j < ACC; ++j) acc[j] = fmaf(acc[j], ws[j & 3], x);
  }
  float s = 0;
  for (int j = 0; j < ACC; ++j) s += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_reg3(float* out, float x, float w) {  // acc = o * w + acc with all three distinct registers
  float acc[ACC];
  float o[4] = {x, x * 1.1f, x * 1.2f, x * 1.3f};
  float ws[4] = {w, w * 1.1f, w * 1.2f, w * 1.3f};
  for (int j = 0; j < ACC; ++j) acc[j] = threadIdx.x + j;
  for (int i = 0; i < ITER; ++i) {
#pragma unroll
    for (int j = 0; j < ACC; ++j) acc[j] = fmaf(o[j & 3], ws[(j >> 2) & 3], acc[j]);
    o[i & 3] += 1e-9f;
  }
  float s = 0;
  for (int j = 0; j < ACC; ++j) s += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_const(float* out, float x) {  // acc = o * c[..] + acc
  float acc[ACC];
  float o[4] = {x, x * 1.1f, x * 1.2f, x * 1.3f};
  for (int j = 0; j < ACC; ++j) acc[j] = threadIdx.x + j;
  for (int i = 0; i < ITER; ++i) {
#pragma unroll
    for (int j = 0; j < ACC; ++j) acc[j] = fmaf(o[j & 3], c_w[j], acc[j]);
    o[i & 3] += 1e-9f;
  }
  float s = 0;
  for (int j = 0; j < ACC; ++j) s += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ void fma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__global__ void k_pack(float* out, float x, float w) {  // ACC/2 packed accumulators
  unsigned long long acc[ACC / 2], o[4], ws[4];
  for (int j = 0; j < 4; ++j) {
    float2 t = make_float2(x * (1.f + 0.1f * j), x * (1.f + 0.1f * j));
    o[j] = *reinterpret_cast<unsigned long long*>(&t);
    float2 u = make_float2(w * (1.f + 0.1f * j), w * (1.05f + 0.1f * j));
    ws[j] = *reinterpret_cast<unsigned long long*>(&u);
  }
  for (int j = 0; j < ACC / 2; ++j) {
    float2 t = make_float2(threadIdx.x + j, threadIdx.x - j);
    acc[j] = *reinterpret_cast<unsigned long long*>(&t);
  }
  for (int i = 0; i < ITER; ++i) {
#pragma unroll
    for (int j = 0; j < ACC / 2; ++j) fma2(acc[j], o[j & 3], ws[(j >> 1) & 3]);
  }
  float s = 0;
  for (int j = 0; j < ACC / 2; ++j) {
    float2 t = *reinterpret_cast<float2*>(&acc[j]);
    s += t.x + t.y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F>
static void run(const char* name, F launch, double fma_per_thread, int blocks, int threads) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int r = 0; r < 5; ++r) launch();
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double fma = 5.0 * fma_per_thread * blocks * threads;
  printf("%-28s %8.3f ms  %7.2f TFMA/s  %6.1f FMA/clk/SM (at %d MHz nominal)\n", name, ms, fma / (ms * 1e-3) / 1e12,
         fma / (ms * 1e-3) / (clk * 1e3) / sms, clk / 1000);
}
int main() {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int threads = 256, blocks = sms * 8;
  float* out;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  float hw[64];
  for (int i = 0; i < 64; ++i) hw[i] = 1.0f + 1e-7f * i;
  cudaMemcpyToSymbol(c_w, hw, sizeof(hw));
  run("FFMA acc*w+x (reuse)", [&] { k_reg<<<blocks, threads>>>(out, 1e-3f, 0.999f); }, (double)ITER * ACC, blocks, threads);
  run("FFMA o*w+acc (3 regs)", [&] { k_reg3<<<blocks, threads>>>(out, 1e-3f, 0.999f); }, (double)ITER * ACC, blocks, threads);
  run("FFMA o*c[]+acc (const)", [&] { k_const<<<blocks, threads>>>(out, 1e-3f); }, (double)ITER * ACC, blocks, threads);
  run("FFMA2 packed f32x2", [&] { k_pack<<<blocks, threads>>>(out, 1e-3f, 0.999f); }, (double)ITER * ACC, blocks, threads);
  return 0;
}
