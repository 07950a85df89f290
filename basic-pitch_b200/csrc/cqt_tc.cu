// Constant-Q projection on the tensor cores: per octave the 172 x 256 x 72 contraction
//   C[(b,t)][n] = sum_k xpad_o[b][t*hop_o + k] * W[k][n]          (n interleaves real / imaginary parts of 36 bins)
// as tcgen05.mma kind::tf32 with the "3xTF32" split (x = hi + lo, products hi*hi + hi*lo + lo*hi, FP32 accumulation in
// TMEM): a single TF32 or bf16 product is far from the 1e-3 bar for this stage, the
// split reproduces FP32-class results.
//
// Replaces (together with the unchanged decimation chain) reference: basic_pitch/layers/nnaudio.py:216-256
// (`get_cqt_complex`, reflect pad + two strided conv1d per octave), :642-661 (concat, sqrt(len) scaling, magnitude) and
// layers/signal.py:174-176 (power -> 10*log10(. + 1e-10)); the per-window min / max feed lognorm_kernel (hcqt.cu).
//
// The A operand is an overlapping strided view of the signal (row t starts at sample t*hop), which no UMMA/TMA
// descriptor can express for hop*4 B < 16 B or non-canonical pitches, so it is staged explicitly ("im2col" into the
// canonical K-major core-matrix layout) by four producer warps that also do the reflect padding and the hi/lo split.
//
// item = (M-tile of 128 frames, octave); per item 4 K-chunks of 64 taps, each chunk = 8 k-steps x 3 products:
//   warps 0-3   producers: gather 128 x 64 samples, split, st.shared into [plane][k/4][row][4] ; one lane bulk-copies
//               the matching 40 KB slice of the split kernel matrix W (UBLKCP) ; fence.proxy.async ; mbarrier arrive
//   warp 4      MMA issuer (one elected lane): 24 tcgen05.mma per chunk, tcgen05.commit frees the stage / publishes TMEM
//   warps 5-8   epilogue: tcgen05.ld 80 columns, magnitude * sqrt(len), log-power, store, per-window min/max (atomics)
// Shared memory: 2 stages x (64 KB A + 40 KB W); TMEM: 2 accumulators of 128 x 80 (256 columns allocated).
#include <vector>

#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace bp {

namespace cq {
constexpr int kMTile = 128;
constexpr int kKc = 64;                          // taps per chunk
constexpr int kN = 80;                           // 72 columns padded to a multiple of 16
constexpr int kAPlane = (kKc / 4) * kMTile * 16;  // 32768 B : [16 k-chunks of 4][128 rows][16 B]
constexpr int kWPlane = (kKc / 4) * kN * 16;      // 20480 B : [16][80][16 B]
constexpr int kStageBytes = 2 * kAPlane + 2 * kWPlane;  // 106496
constexpr int kStages = 2;
constexpr int kThreads = 288;
constexpr int kSmemBytes = kStages * kStageBytes + 256;
}  // namespace cq

// kernel matrix, split and laid out per chunk: wtc[chunk 4][plane 2][k/4 16][n 80][4]
void build_cqt_tc_weights(const float* cqt_real /* [36][256] */, const float* cqt_imag, std::vector<float>& out) {
  out.assign((size_t)4 * 2 * 16 * cq::kN * 4, 0.f);
  for (int k = 0; k < kTaps; ++k)
    for (int n = 0; n < 72; ++n) {
      const float w = (n & 1) ? cqt_imag[(n >> 1) * kTaps + k] : cqt_real[(n >> 1) * kTaps + k];
      uint32_t u;
      memcpy(&u, &w, 4);
      u &= 0xffffe000u;  // TF32 keeps the top 19 bits
      float hi;
      memcpy(&hi, &u, 4);
      const float lo = w - hi;
      const int c = k / cq::kKc, kk = k % cq::kKc;
      const size_t base = (size_t)c * 2 * 16 * cq::kN * 4;
      const size_t off = ((size_t)(kk >> 2) * cq::kN + n) * 4 + (kk & 3);
      out[base + off] = hi;
      out[base + (size_t)16 * cq::kN * 4 + off] = lo;
    }
}

// instruction descriptor, kind::tf32: D = f32 (bit 4), A = B = TF32 (format 2 at bits 7, 10), both K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

struct CqtTcArgs {
  const float* audio;
  const WinDesc* desc;   // may be null: window b = audio + b*43844
  const float* chain;    // decimated signals x_1..x_8
  const float* wtc;      // split kernel matrix, 4 chunks of 40 KB
  const float* scale;    // [309] sqrt(kernel length)
  float* logmag;         // [B][172][309]
  unsigned int* minmax;  // [B][2] ordered-uint min / max
  int n_windows, n_mtiles;
};

__global__ void __launch_bounds__(cq::kThreads, 1) cqt_tc_kernel(const CqtTcArgs a) {
  using namespace cq;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full = bars;              // [kStages]  128 producer arrivals + the bytes of the W slice
  uint64_t* empty = bars + kStages;   // [kStages]
  uint64_t* tmem_full = bars + 2 * kStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full + s, 128);
      mbar_init(empty + s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full + i, 1);
      mbar_init(tmem_empty + i, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_items = a.n_mtiles * kOctaves;
  const int total_frames = a.n_windows * kFrames;

  if (warp < 4) {
    // ------------------------------ producers: one row (frame) per thread ------------------------------
    const int r = threadIdx.x;  // 0..127
    uint32_t stage = 0, ph = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int mt = it / kOctaves, o = it % kOctaves;
      const int hop = 256 >> o;
      const int len = octave_len(o);
      const int m = mt * kMTile + r;
      const bool live = m < total_frames;
      const int b = live ? m / kFrames : 0;
      const int t = live ? m - b * kFrames : 0;
      const float* src;
      int lo = 0, hi = len;
      if (o == 0) {
        if (a.desc) {
          const WinDesc d = a.desc[b];
          src = a.audio + d.base;
          lo = d.lo;
          hi = d.hi;
        } else {
          src = a.audio + (long long)b * kWinSamples;
        }
      } else {
        src = a.chain + (size_t)b * kChainStride + chain_off(o);
      }
      const int i0 = t * hop - 128;  // signal index of tap 0 of this frame
      for (int c = 0; c < kTaps / kKc; ++c) {
        mbar_wait(empty + stage, ph ^ 1);
        unsigned char* sa = smem + stage * kStageBytes;
        if (threadIdx.x == 0) {
          mbar_expect_tx_only(full + stage, 2 * kWPlane);  // the bulk copy of the W slice completes on the same barrier
          bulk_g2s(sa + 2 * kAPlane, a.wtc + (size_t)c * (2 * kWPlane / 4), 2 * kWPlane, full + stage);
        }
        const int ib = i0 + c * kKc;
        const bool interior = live && ib >= lo && ib + kKc <= hi && ib >= 0 && ib + kKc <= len;
#pragma unroll 4
        for (int kc = 0; kc < kKc / 4; ++kc) {
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            int idx = ib + 4 * kc + j;
            float x = 0.f;
            if (interior) {
              x = __ldg(src + idx);
            } else if (live) {
              if (idx < 0) idx = -idx;
              if (idx >= len) idx = 2 * (len - 1) - idx;
              if (idx >= lo && idx < hi) x = __ldg(src + idx);
            }
            v[j] = x;
          }
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(v[0]) & 0xffffe000u);
          h.y = __uint_as_float(__float_as_uint(v[1]) & 0xffffe000u);
          h.z = __uint_as_float(__float_as_uint(v[2]) & 0xffffe000u);
          h.w = __uint_as_float(__float_as_uint(v[3]) & 0xffffe000u);
          l.x = v[0] - h.x;
          l.y = v[1] - h.y;
          l.z = v[2] - h.z;
          l.w = v[3] - h.w;
          *reinterpret_cast<float4*>(sa + (kc * kMTile + r) * 16) = h;
          *reinterpret_cast<float4*>(sa + kAPlane + (kc * kMTile + r) * 16) = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
        mbar_arrive(full + stage);
        if (++stage == kStages) {
          stage = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 4) {
    // ------------------------------ MMA issuer ------------------------------
    constexpr uint32_t idesc = make_idesc_tf32(128, kN);
    const uint32_t leader = elect_one() ? 1u : 0u;
    uint32_t stage = 0, ph = 0, icount = 0;
    uint32_t ph_t[2] = {0, 0};
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const uint32_t buf = icount & 1u;
      mbar_wait(tmem_empty + buf, ph_t[buf] ^ 1);
      ph_t[buf] ^= 1;
      tc_fence_after();
      const uint32_t d = tmem_base + buf * 128u;
      for (int c = 0; c < kTaps / kKc; ++c) {
        mbar_wait(full + stage, ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * kStageBytes);
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < kKc / 8; ++ks) {
            // one k-step = 8 taps = two 16-byte k-chunks, LBO apart
            const uint64_t a_hi = make_desc(sa + ks * 2 * (kMTile * 16), kMTile * 16, 128);
            const uint64_t a_lo = make_desc(sa + kAPlane + ks * 2 * (kMTile * 16), kMTile * 16, 128);
            const uint64_t b_hi = make_desc(sa + 2 * kAPlane + ks * 2 * (kN * 16), kN * 16, 128);
            const uint64_t b_lo = make_desc(sa + 2 * kAPlane + kWPlane + ks * 2 * (kN * 16), kN * 16, 128);
            umma_tf32(d, a_hi, b_hi, idesc, (c | ks) ? 1u : 0u);
            umma_tf32(d, a_hi, b_lo, idesc, 1u);
            umma_tf32(d, a_lo, b_hi, idesc, 1u);
          }
          umma_commit(empty + stage);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          ph ^= 1;
        }
      }
      if (leader) umma_commit(tmem_full + buf);
      __syncwarp();
      ++icount;
    }
  } else {
    // ------------------------------ epilogue (warps 5..8) ------------------------------
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    uint32_t ph_t[2] = {0, 0};
    uint32_t icount = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int mt = it / kOctaves, o = it % kOctaves;
      const int m = mt * kMTile + row;
      const bool live = m < total_frames;
      const int b = live ? m / kFrames : -1;
      const uint32_t buf = icount & 1u;
      mbar_wait(tmem_full + buf, ph_t[buf]);
      ph_t[buf] ^= 1;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 128u;
      float vmin = INFINITY, vmax = -INFINITY;
      float* out = a.logmag + (size_t)m * kCqtBins;  // m == b*172 + t
      const int g0 = (8 - o) * kBinsPerOctave - 15;  // global bin of this octave's bin 0 (may be negative for o = 8)
#pragma unroll 1
      for (int part = 0; part < 5; ++part) {  // 5 x 16 columns = 8 bins each (the last part holds bins 32..35 + padding)
        uint32_t v[16];
        tmem_ld16_nowait(taddr + part * 16, v);
        tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int jb = 0; jb < 8; ++jb) {
            const int j = part * 8 + jb;
            const int g = g0 + j;
            if (j < kBinsPerOctave && g >= 0) {
              const float s = __ldg(a.scale + g);
              const float re = __fmul_rn(__uint_as_float(v[2 * jb]), s), im = __fmul_rn(__uint_as_float(v[2 * jb + 1]), s);
              const float mag = sqrtf(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)));
              const float p = __fadd_rn(__fmul_rn(mag, mag), 1e-10f);
              const float L = __fmul_rn(__fmul_rn(logf(p), 0.4342944622039795f), 10.0f);
              out[g] = L;
              vmin = fminf(vmin, L);
              vmax = fmaxf(vmax, L);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + buf);
      // per-window min / max: one atomic pair per warp when the whole warp sits in one window
      const int b0 = __shfl_sync(0xffffffffu, b, 0);
      const bool uniform = __all_sync(0xffffffffu, b == b0);
      if (uniform) {
        if (b0 >= 0) {
#pragma unroll
          for (int off = 16; off; off >>= 1) {
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, off));
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, off));
          }
          if (lane == 0 && vmin <= vmax) {
            atomicMin(a.minmax + 2 * b0, float_to_ordered(vmin));
            atomicMax(a.minmax + 2 * b0 + 1, float_to_ordered(vmax));
          }
        }
      } else if (live && vmin <= vmax) {
        atomicMin(a.minmax + 2 * b, float_to_ordered(vmin));
        atomicMax(a.minmax + 2 * b + 1, float_to_ordered(vmax));
      }
      ++icount;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
  }
}

__global__ void minmax_init_kernel2(unsigned int* mm, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    mm[2 * i] = 0xffffffffu;
    mm[2 * i + 1] = 0u;
  }
}

void cqt_tc_setup() {
  cudaFuncSetAttribute(cqt_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cq::kSmemBytes);
}

void launch_cqt_tc(const float* audio, const WinDesc* desc, const float* chain, const float* wtc, const float* scale,
                   float* logmag, unsigned int* minmax, int n_windows, int n_sms, cudaStream_t st) {
  minmax_init_kernel2<<<(n_windows + 255) / 256, 256, 0, st>>>(minmax, n_windows);
  CqtTcArgs a;
  a.audio = audio;
  a.desc = desc;
  a.chain = chain;
  a.wtc = wtc;
  a.scale = scale;
  a.logmag = logmag;
  a.minmax = minmax;
  a.n_windows = n_windows;
  a.n_mtiles = (n_windows * kFrames + cq::kMTile - 1) / cq::kMTile;
  const int n_items = a.n_mtiles * kOctaves;
  const int grid = n_items < n_sms ? n_items : n_sms;
  cqt_tc_kernel<<<grid, cq::kThreads, cq::kSmemBytes, st>>>(a);
}

}  // namespace bp
