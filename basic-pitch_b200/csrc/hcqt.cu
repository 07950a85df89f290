// Harmonic constant-Q front end, FP32 path.
//
// Replaces the first 213 nodes of the deployed graph, i.e.
//   CQT2010v2.call        reference: basic_pitch/layers/nnaudio.py:623-661
//   get_cqt_complex       reference: basic_pitch/layers/nnaudio.py:216-256
//   downsampling_by_n     reference: basic_pitch/layers/nnaudio.py:259-284
//   NormalizedLog.call    reference: basic_pitch/layers/signal.py:171-185
//   BatchNormalization    reference: basic_pitch/models.py:188-189 (folded scalar affine)
// Kernels:
//   decimate_kernel  x_{o+1}[n] = sum_k LP[k] * x_o[2n + k - 127]           (8 launches per chunk)
//   cqt_kernel       frames(172) x taps(256) x 72 projection per octave, magnitude*sqrt(len),
//                    10*log10(p + 1e-10), per-window min/max (atomics)       (1 launch per chunk)
//   lognorm_kernel   (L - min) / (max - min) * bn_scale + bn_bias            (1 launch per chunk)
#include "kernels.cuh"

namespace bp {

__constant__ float c_lowpass[kTaps];

void upload_lowpass(const float* d_lp, cudaStream_t st) {
  cudaMemcpyToSymbolAsync(c_lowpass, d_lp, sizeof(float) * kTaps, 0, cudaMemcpyDeviceToDevice, st);
}

// ------------------------------------------------------------------------------------------------
// Half-band FIR + decimate by 2.  One CTA = 512 outputs of one window; thread t owns outputs
// 4t..4t+3.  The 1278 input samples of the tile are stored de-interleaved into 8 phases so that
// lanes read consecutive words (no bank conflicts) while each loaded sample feeds up to 4 FMAs;
// the taps come from the constant bank as immediate operands of the (fully unrolled) FMAs.
// ------------------------------------------------------------------------------------------------
constexpr int kDecTile = 512;
constexpr int kDecThreads = 128;
constexpr int kDecPhaseStride = 164;  // == 4 (mod 32): conflict-free scatter and gather

__global__ void __launch_bounds__(kDecThreads) decimate_kernel(const float* __restrict__ audio,
                                                               const WinDesc* __restrict__ desc,  // stage 0 only
                                                               const float* __restrict__ src,     // chain, stage >= 1
                                                               float* __restrict__ dst, int src_off, int dst_off,
                                                               int len_in, int len_out, int from_audio) {
  __shared__ float ph[8 * kDecPhaseStride];
  const int b = blockIdx.y;
  const int n0 = blockIdx.x * kDecTile;
  const int gbase = 2 * n0 - 127;  // input index of local index 0

  if (from_audio) {
    long long base;
    int lo, hi;
    if (desc) {
      WinDesc d = desc[b];
      base = d.base;
      lo = d.lo;
      hi = d.hi;
    } else {
      base = (long long)b * kWinSamples;
      lo = 0;
      hi = kWinSamples;
    }
    for (int li = threadIdx.x; li < 8 * 160; li += kDecThreads) {
      int gi = gbase + li;
      float v = (gi >= lo && gi < hi) ? __ldg(audio + base + gi) : 0.f;
      ph[(li & 7) * kDecPhaseStride + (li >> 3)] = v;
    }
  } else {
    const float* s = src + (size_t)b * kChainStride + src_off;
    for (int li = threadIdx.x; li < 8 * 160; li += kDecThreads) {
      int gi = gbase + li;
      float v = (gi >= 0 && gi < len_in) ? s[gi] : 0.f;
      ph[(li & 7) * kDecPhaseStride + (li >> 3)] = v;
    }
  }
  __syncthreads();

  const int t = threadIdx.x;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll
  for (int u = 0; u < kTaps + 6; ++u) {
    float v = ph[(u & 7) * kDecPhaseStride + t + (u >> 3)];
    if (u < kTaps) acc0 = fmaf(c_lowpass[u], v, acc0);
    if (u >= 2 && u - 2 < kTaps) acc1 = fmaf(c_lowpass[u - 2], v, acc1);
    if (u >= 4 && u - 4 < kTaps) acc2 = fmaf(c_lowpass[u - 4], v, acc2);
    if (u >= 6 && u - 6 < kTaps) acc3 = fmaf(c_lowpass[u - 6], v, acc3);
  }
  float* d = dst + (size_t)b * kChainStride + dst_off;
  int n = n0 + 4 * t;
  if (n + 3 < len_out) {
    *reinterpret_cast<float4*>(d + n) = make_float4(acc0, acc1, acc2, acc3);
  } else {
    if (n < len_out) d[n] = acc0;
    if (n + 1 < len_out) d[n + 1] = acc1;
    if (n + 2 < len_out) d[n + 2] = acc2;
  }
}

void launch_decimate(const float* audio, const WinDesc* desc, float* chain, int stage, int n_windows,
                     cudaStream_t st) {
  // stage s: x_s -> x_{s+1}
  const int len_in = octave_len(stage), len_out = octave_len(stage + 1);
  dim3 grid((len_out + kDecTile - 1) / kDecTile, n_windows);
  decimate_kernel<<<grid, kDecThreads, 0, st>>>(audio, desc, chain, chain, stage ? chain_off(stage) : 0,
                                                 chain_off(stage + 1), len_in, len_out, stage == 0);
}

// ------------------------------------------------------------------------------------------------
// CQT projection of one (window, octave): C[t][n] = sum_k xpad[t*hop + k] * W[k][n], t < 172,
// n < 72 (columns interleave real/imag of the 36 bins), reflect padding of 128 each side.
// 288 threads; thread (ty, tx) accumulates frames {ty, ty+43, ty+86, ty+129} x columns 12tx..12tx+11.
// K is consumed in chunks of 64 staged through shared memory (A gathered through the reflect map).
// Epilogue: magnitude * sqrt(len_k), power, 10*log10(p + 1e-10), block min/max -> atomics.
// ------------------------------------------------------------------------------------------------
constexpr int kCqtThreads = 288;
constexpr int kCqtKc = 64;
constexpr int kCqtApad = kCqtKc + 1;
constexpr int kCqtCols = 72;

__global__ void __launch_bounds__(kCqtThreads) cqt_kernel(const float* __restrict__ audio,
                                                          const WinDesc* __restrict__ desc,
                                                          const float* __restrict__ chain,
                                                          const float* __restrict__ wt,     // [256][72]
                                                          const float* __restrict__ scale,  // [309]
                                                          float* __restrict__ logmag,       // [B][172][309]
                                                          unsigned int* __restrict__ minmax /* [B][2] */) {
  extern __shared__ float smem[];
  float* As = smem;                       // [172][65]
  float* Ws = smem + kFrames * kCqtApad;  // [64][72]
  __shared__ float red_min[kCqtThreads / 32], red_max[kCqtThreads / 32];

  const int b = blockIdx.y;
  const int o = blockIdx.x;  // octave, 0 = top
  const int hop = 256 >> o;
  const int len = octave_len(o);

  long long base = 0;
  int lo = 0, hi = len;
  const float* src;
  if (o == 0) {
    if (desc) {
      WinDesc d = desc[b];
      base = d.base;
      lo = d.lo;
      hi = d.hi;
    } else {
      base = (long long)b * kWinSamples;
    }
    src = audio;
  } else {
    src = chain + (size_t)b * kChainStride + chain_off(o);
  }

  const int tid = threadIdx.x;
  const bool active = tid < 258;
  const int ty = active ? tid / 6 : 0;
  const int tx = active ? tid % 6 : 0;

  float acc[4][12];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 12; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < kTaps; k0 += kCqtKc) {
    __syncthreads();
    for (int e = tid; e < kFrames * kCqtKc; e += kCqtThreads) {
      int t = e >> 6, kk = e & 63;
      int idx = t * hop + k0 + kk - 128;
      if (idx < 0) idx = -idx;
      if (idx >= len) idx = 2 * (len - 1) - idx;
      float v = (idx >= lo && idx < hi) ? __ldg(src + base + idx) : 0.f;
      As[t * kCqtApad + kk] = v;
    }
    for (int e = tid; e < kCqtKc * kCqtCols / 4; e += kCqtThreads)
      reinterpret_cast<float4*>(Ws)[e] = __ldg(reinterpret_cast<const float4*>(wt + k0 * kCqtCols) + e);
    __syncthreads();
    if (active) {
#pragma unroll 4
      for (int kk = 0; kk < kCqtKc; ++kk) {
        float a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[(ty + 43 * i) * kCqtApad + kk];
        const float4* wp = reinterpret_cast<const float4*>(Ws + kk * kCqtCols + 12 * tx);
        float4 w0 = wp[0], w1 = wp[1], w2 = wp[2];
        float w[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 12; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
      }
    }
  }

  float vmin = INFINITY, vmax = -INFINITY;
  if (active) {
#pragma unroll
    for (int jb = 0; jb < 6; ++jb) {
      int g = (8 - o) * kBinsPerOctave + 6 * tx + jb - 15;
      if (g < 0) continue;
      float s = __ldg(scale + g);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float re = __fmul_rn(acc[i][2 * jb], s), im = __fmul_rn(acc[i][2 * jb + 1], s);
        float mag = sqrtf(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)));
        float p = __fadd_rn(__fmul_rn(mag, mag), 1e-10f);
        float L = __fmul_rn(__fmul_rn(logf(p), 0.4342944622039795f), 10.0f);
        logmag[((size_t)b * kFrames + ty + 43 * i) * kCqtBins + g] = L;
        vmin = fminf(vmin, L);
        vmax = fmaxf(vmax, L);
      }
    }
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) {
    vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, off));
    vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, off));
  }
  if ((tid & 31) == 0) {
    red_min[tid >> 5] = vmin;
    red_max[tid >> 5] = vmax;
  }
  __syncthreads();
  if (tid == 0) {
    for (int i = 1; i < kCqtThreads / 32; ++i) {
      vmin = fminf(vmin, red_min[i]);
      vmax = fmaxf(vmax, red_max[i]);
    }
    atomicMin(minmax + 2 * b, float_to_ordered(vmin));
    atomicMax(minmax + 2 * b + 1, float_to_ordered(vmax));
  }
}

__global__ void minmax_init_kernel(unsigned int* mm, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    mm[2 * i] = 0xffffffffu;
    mm[2 * i + 1] = 0u;
  }
}

void launch_cqt(const float* audio, const WinDesc* desc, const float* chain, const float* wt, const float* scale,
                float* logmag, unsigned int* minmax, int n_windows, cudaStream_t st) {
  static bool attr_set = false;
  const int smem = (kFrames * kCqtApad + kCqtKc * kCqtCols) * sizeof(float);
  if (!attr_set) {
    cudaFuncSetAttribute(cqt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr_set = true;
  }
  minmax_init_kernel<<<(n_windows + 255) / 256, 256, 0, st>>>(minmax, n_windows);
  cqt_kernel<<<dim3(kOctaves, n_windows), kCqtThreads, smem, st>>>(audio, desc, chain, wt, scale, logmag, minmax);
}

// ------------------------------------------------------------------------------------------------
// Per-window normalisation + folded BatchNorm, in place.
// ------------------------------------------------------------------------------------------------
__global__ void lognorm_kernel(float* __restrict__ y, const unsigned int* __restrict__ minmax,
                               const float* __restrict__ bn) {
  const int b = blockIdx.y;
  const float bn_scale = __ldg(bn), bn_bias = __ldg(bn + 1);
  const float mn = ordered_to_float(minmax[2 * b]);
  const float mx = __fsub_rn(ordered_to_float(minmax[2 * b + 1]), mn);
  float* p = y + (size_t)b * kFrames * kCqtBins;
  const int n = kFrames * kCqtBins;  // 53148 = 4 * 13287
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n / 4; i += gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<float4*>(p)[i];
    float* e = reinterpret_cast<float*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float q = (mx == 0.f) ? 0.f : __fdiv_rn(__fsub_rn(e[j], mn), mx);
      e[j] = __fadd_rn(__fmul_rn(q, bn_scale), bn_bias);
    }
    reinterpret_cast<float4*>(p)[i] = v;
  }
}

void launch_lognorm(float* y, const unsigned int* minmax, const float* bn, int n_windows, cudaStream_t st) {
  lognorm_kernel<<<dim3(13, n_windows), 256, 0, st>>>(y, minmax, bn);
}

}  // namespace bp
