// Contour convolution (8 -> 8 channels, 3 x 39 taps, 65 % of the model's FLOPs) on the 5th-gen tensor
// cores: tcgen05.mma (kind::f16, bf16 operands, fp32 accumulators in TMEM), operands staged in shared
// memory by bulk async copies (UBLKCP) signalled through mbarriers, warp-specialised roles.
//
// Replaces node 231 (+ ReLU 232) of the deployed graph — reference: basic_pitch/models.py:241-250 — and
// the harmonic stacking in front of it (reference: basic_pitch/nn.py:69-88), which is folded into the
// weight operand and never materialised.
//
// Formulation ("Toeplitz along frequency on 16-bin aligned chunks")
//   rows  m = b*173 + t            time frames of all windows of the chunk, one zero row between windows
//   D[m][(fl,co)] (+)= A[m+dt][16q .. 16q+15] x T(ci,dt,rel)[16][(fl,co)]       rel = q - ft
//     A      the normalised CQT y itself (NOT the 8-channel stack), rows shifted by the time tap dt
//     T      128 x 16 "weight tile": T[k][(fl,co)] = W[co][ci][dt][df], df = 16*rel + k - shift_ci - fl + 19
//            (zero outside 0 <= df < 39); tiles that touch the edges of the stacked image carry the zero
//            gating of the reference's "same" padding (boundary variants)
//   every (ft, ci, dt, q) with a non-empty tile is one K=16 MMA step of shape 128 x 128 x 16
// Precision: both operands are split x = hi + lo (bf16 each) and three products are accumulated
// (hi*hi + hi*lo + lo*hi) in fp32, which keeps the posteriorgrams within ~1e-5 of the FP32 path
// (SURVEY.md Appendix C.4); a single bf16 product would miss the 1e-3 bar.
//
// Work decomposition: unit = (M-tile of 128 rows, group of 4 frequency tiles of 16 bins).  A CTA (1 per
// SM, persistent) walks units u = blockIdx.x, +gridDim.x, ...:
//   warp 0      producer: bulk-copies the 130 x 320 bf16 hi/lo data tile (k-chunk-major, 166 KB) and streams
//               the weight tiles of the group's program (8 KB each) through a 4-stage ring
//   warp 1      MMA issuer (one elected lane): 3 x tcgen05.mma per program use, tcgen05.commit to free the
//               weight stage / publish the accumulators; owns the TMEM allocation (512 columns)
//   warps 2-5   epilogue: tcgen05.ld the 4 x 128 accumulator columns, + bias, ReLU, store NHWC
#include <cuda_bf16.h>

#include <vector>

#include "kernels.cuh"

namespace bp {

namespace tc {
constexpr int kRowsPerWindow = kFrames + 1;  // 173: one zero separator row after every window
constexpr int kMTile = 128;
constexpr int kDataRows = kMTile + 2;                        // 130
constexpr int kChunks8 = 40;                                 // 320 bins / 8
constexpr int kDataLbo = kDataRows * 16;                     // 2080 B between 8-element k-chunks
constexpr int kDataPlaneBytes = kChunks8 * kDataLbo;         // 83200
constexpr int kDataBytes = 2 * kDataPlaneBytes;              // hi + lo = 166400
constexpr int kTileBytes = 8192;                             // weight tile: [plane 2][kchunk 2][128][8] bf16
constexpr int kStages = 4;
constexpr int kFTiles = 17;                                  // ceil(264 / 16)
constexpr int kGroups = 5;                                   // 4 + 4 + 4 + 4 + 1 frequency tiles
constexpr int kThreads = 192;
constexpr int kSmemBytes = kDataBytes + kStages * kTileBytes + 256;
constexpr int kShift[kHarmonics] = {-36, 0, 36, 57, 72, 84, 93, 101};
}  // namespace tc

// ------------------------------------------------------------------------------------------------
// Host: weight tiles + per-group programs
// ------------------------------------------------------------------------------------------------
static inline uint16_t f2bf(float x) {  // round-to-nearest-even float -> bf16 bits
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

void TcContourPlan::build(const float* W /* [8][8][3][39] */) {
  using namespace tc;
  tiles.clear();
  tile_seq.clear();
  use_words.clear();
  step_use_off.clear();
  group_step_off.assign(kGroups + 1, 0);

  struct Key {
    int ci, dt, rel, variant;  // variant 0 = interior, 1 = low boundary chunk, 2 = high boundary chunk
  };
  std::vector<Key> keys;
  auto find_or_add = [&](Key k) -> int {
    for (size_t i = 0; i < keys.size(); ++i)
      if (keys[i].ci == k.ci && keys[i].dt == k.dt && keys[i].rel == k.rel && keys[i].variant == k.variant) return (int)i;
    keys.push_back(k);
    // materialise the tile: [plane][kchunk][n][8]
    const int s = kShift[k.ci];
    const size_t base = tiles.size();
    tiles.resize(base + kTileBytes / 2, 0);
    for (int kk = 0; kk < 16; ++kk) {
      for (int n = 0; n < 128; ++n) {
        const int fl = n >> 3, co = n & 7;
        const int df = 16 * k.rel + kk - s - fl + 19;
        float w = 0.f;
        if (df >= 0 && df < 39) {
          bool keep = true;
          if (k.variant == 1) {  // chunk straddles g = 0: bins u < s are outside the stacked image
            const int qb = s / 16;
            keep = (16 * qb + kk - s) >= 0;
          } else if (k.variant == 2) {  // chunk straddles g = 264
            const int qh = (kContourBins + s) / 16;
            keep = (16 * qh + kk - s) < kContourBins;
          }
          if (keep) w = W[((co * 8 + k.ci) * 3 + k.dt) * 39 + df];
        }
        const uint16_t hi = f2bf(w);
        const uint16_t lo = f2bf(w - bf2f(hi));
        const size_t off = (size_t)(kk >> 3) * 128 * 8 + (size_t)n * 8 + (kk & 7);
        tiles[base + off] = hi;
        tiles[base + 2048 + off] = lo;
      }
    }
    return (int)keys.size() - 1;
  };

  for (int g = 0; g < kGroups; ++g) {
    const int ft0 = 4 * g, nft = (g == kGroups - 1) ? 1 : 4;
    // collect uses per tile id, keeping a deterministic order (ci, dt, rel, variant)
    struct Use {
      int tile, ftl, q, dt;
    };
    std::vector<Use> uses;
    for (int ci = 0; ci < 8; ++ci) {
      const int s = kShift[ci];
      for (int dt = 0; dt < 3; ++dt) {
        for (int rel = -5; rel <= 9; ++rel) {
          for (int variant = 0; variant < 3; ++variant) {
            for (int ftl = 0; ftl < nft; ++ftl) {
              const int ft = ft0 + ftl, q = ft + rel;
              if (q < 0 || q >= 20) continue;
              // does this (ft, q) pair contribute at all?  u in [16q, 16q+16), f in [16ft, 16ft+16)
              bool any = false;
              for (int kk = 0; kk < 16 && !any; ++kk)
                for (int fl = 0; fl < 16 && !any; ++fl) {
                  const int u = 16 * q + kk, f = 16 * ft + fl, gg = u - s;
                  const int df = gg - f + 19;
                  if (df >= 0 && df < 39 && gg >= 0 && gg < kContourBins && u < kCqtBins && f < kContourBins) any = true;
                }
              if (!any) continue;
              // which variant does this chunk need?
              int need = 0;
              if (s > 0 && s % 16 != 0 && q == s / 16) need = 1;
              if (kContourBins + s < 320 && (kContourBins + s) % 16 != 0 && q == (kContourBins + s) / 16) need = 2;
              if (need != variant) continue;
              const int id = find_or_add(Key{ci, dt, rel, variant});
              uses.push_back(Use{id, ftl, q, dt});
            }
          }
        }
      }
    }
    // steps: consecutive uses of the same tile share one staged copy
    std::vector<bool> seen(4, false);
    size_t i = 0;
    while (i < uses.size()) {
      size_t j = i;
      while (j < uses.size() && uses[j].tile == uses[i].tile) ++j;
      tile_seq.push_back(uses[i].tile);
      step_use_off.push_back((int)use_words.size());
      for (size_t u = i; u < j; ++u) {
        const bool first = !seen[uses[u].ftl];
        seen[uses[u].ftl] = true;
        use_words.push_back((uint32_t)uses[u].ftl | ((uint32_t)uses[u].q << 2) | ((uint32_t)uses[u].dt << 7) |
                            ((uint32_t)(first ? 1 : 0) << 9));
      }
      i = j;
    }
    group_step_off[g + 1] = (int)tile_seq.size();
  }
  step_use_off.push_back((int)use_words.size());
  n_tiles = (int)keys.size();
}

// ------------------------------------------------------------------------------------------------
// Device helpers (inline PTX)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// K-major, no-swizzle shared-memory matrix descriptor (SM100 "version 1"):
//   [0,14) start >> 4, [16,30) leading-dimension byte offset >> 4 (between the two 8-element k-chunks),
//   [32,46) stride byte offset >> 4 (between 8-row groups), [46,48) = 1, layout type [61,64) = 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor, kind::f16: D = f32 (bit 4), A = B = bf16 (bits 7, 10), both K-major, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// y (fp32, [B][172][309]) -> bf16 hi/lo planes in the k-chunk-major row layout the MMA reads:
//   yhl[plane][q8 (40)][row d (rows_total)][8],  d = 1 + b*173 + t, every other row zero.
// ------------------------------------------------------------------------------------------------
__global__ void y_split_kernel(const float* __restrict__ y, __nv_bfloat16* __restrict__ yhl, int n_windows,
                               int rows_total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one (row, q8) per thread
  const long long total = (long long)rows_total * tc::kChunks8;
  if (idx >= total) return;
  const int q8 = (int)(idx % tc::kChunks8);
  const int d = (int)(idx / tc::kChunks8);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  const int m = d - 1;
  if (m >= 0) {
    const int b = m / tc::kRowsPerWindow, t = m - b * tc::kRowsPerWindow;
    if (b < n_windows && t < kFrames) {
      const float* src = y + ((size_t)b * kFrames + t) * kCqtBins + q8 * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (q8 * 8 + j < kCqtBins) v[j] = __ldg(src + j);
    }
  }
  __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    hi[j] = __float2bfloat16_rn(v[j]);
    lo[j] = __float2bfloat16_rn(v[j] - __bfloat162float(hi[j]));
  }
  const size_t plane = (size_t)tc::kChunks8 * rows_total * 8;
  const size_t off = ((size_t)q8 * rows_total + d) * 8;
  *reinterpret_cast<uint4*>(yhl + off) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(yhl + plane + off) = *reinterpret_cast<const uint4*>(lo);
}

// ------------------------------------------------------------------------------------------------
// The tensor-core kernel
// ------------------------------------------------------------------------------------------------
struct TcArgs {
  const __nv_bfloat16* yhl;     // [2][40][rows_total][8]
  const uint16_t* tiles;        // [n_tiles][8192 B]
  const int* tile_seq;          // per step: tile id
  const int* step_use_off;      // [n_steps + 1]
  const uint32_t* use_words;    // packed uses
  const float* bias;            // [8]
  float* out;                   // [B][172][264][8]  (NHWC)
  int group_step_off[tc::kGroups + 1];
  int rows_total;
  int n_mtiles;
  int n_windows;
};

__global__ void __launch_bounds__(tc::kThreads, 1) contour_tc_kernel(const TcArgs a) {
  using namespace tc;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* s_data = smem;                          // [2 planes][40 chunks][130 rows][16 B]
  unsigned char* s_w = smem + kDataBytes;                // [kStages][8192]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDataBytes + kStages * kTileBytes);
  uint64_t* full_w = bars;             // [kStages]
  uint64_t* empty_w = bars + kStages;  // [kStages]
  uint64_t* data_full = bars + 2 * kStages;
  uint64_t* data_empty = data_full + 1;
  uint64_t* tmem_full = data_full + 2;
  uint64_t* tmem_empty = data_full + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(data_full + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_w + s, 1);
      mbar_init(empty_w + s, 1);
    }
    mbar_init(data_full, 1);
    mbar_init(data_empty, 1);
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_units = a.n_mtiles * kGroups;

  if (warp == 0) {
    // ------------------------------ producer ------------------------------
    if (lane == 0) {
      uint32_t stage = 0, ph_w = 0, ph_d = 0;
      const size_t plane_elems = (size_t)kChunks8 * a.rows_total * 8;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int mt = u / kGroups, g = u % kGroups;
        mbar_wait(data_empty, ph_d ^ 1);
        mbar_expect_tx(data_full, kDataBytes);
        for (int p = 0; p < 2; ++p)
          for (int c = 0; c < kChunks8; ++c)
            bulk_g2s(s_data + p * kDataPlaneBytes + c * kDataLbo,
                     a.yhl + p * plane_elems + ((size_t)c * a.rows_total + (size_t)mt * kMTile) * 8, kDataLbo, data_full);
        ph_d ^= 1;
        for (int s = a.group_step_off[g]; s < a.group_step_off[g + 1]; ++s) {
          mbar_wait(empty_w + stage, ph_w ^ 1);
          mbar_expect_tx(full_w + stage, kTileBytes);
          bulk_g2s(s_w + stage * kTileBytes, a.tiles + (size_t)__ldg(a.tile_seq + s) * (kTileBytes / 2), kTileBytes,
                   full_w + stage);
          if (++stage == kStages) {
            stage = 0;
            ph_w ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(128, 128);
      uint32_t stage = 0, ph_w = 0, ph_d = 0, ph_t = 0;
      const uint32_t data_addr = smem_u32(s_data);
      const uint32_t w_addr = smem_u32(s_w);
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int g = u % kGroups;
        mbar_wait(data_full, ph_d);
        mbar_wait(tmem_empty, ph_t ^ 1);
        tc_fence_after();
        for (int s = a.group_step_off[g]; s < a.group_step_off[g + 1]; ++s) {
          mbar_wait(full_w + stage, ph_w);
          tc_fence_after();
          const uint32_t wb = w_addr + stage * kTileBytes;
          const uint64_t b_hi = make_desc(wb, 2048, 128);
          const uint64_t b_lo = make_desc(wb + 4096, 2048, 128);
          const int u0 = __ldg(a.step_use_off + s), u1 = __ldg(a.step_use_off + s + 1);
          for (int k = u0; k < u1; ++k) {
            const uint32_t w = __ldg(a.use_words + k);
            const uint32_t ftl = w & 3u, q = (w >> 2) & 31u, dt = (w >> 7) & 3u, first = (w >> 9) & 1u;
            const uint32_t aoff = (2u * q) * kDataLbo + dt * 16u;
            const uint64_t a_hi = make_desc(data_addr + aoff, kDataLbo, 128);
            const uint64_t a_lo = make_desc(data_addr + kDataPlaneBytes + aoff, kDataLbo, 128);
            const uint32_t d = tmem_base + ftl * 128u;
            umma_bf16(d, a_hi, b_hi, idesc, first ? 0u : 1u);
            umma_bf16(d, a_hi, b_lo, idesc, 1u);
            umma_bf16(d, a_lo, b_hi, idesc, 1u);
          }
          umma_commit(empty_w + stage);
          if (++stage == kStages) {
            stage = 0;
            ph_w ^= 1;
          }
        }
        umma_commit(tmem_full);   // accumulators of this unit are complete
        umma_commit(data_empty);  // and the data tile may be overwritten
        ph_d ^= 1;
        ph_t ^= 1;
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2..5) ------------------------------
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    float bias[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bias[j] = __ldg(a.bias + j);
    uint32_t ph_t = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const int mt = u / kGroups, g = u % kGroups;
      const int nft = (g == kGroups - 1) ? 1 : 4;
      const int m = mt * kMTile + row;
      const int b = m / kRowsPerWindow, t = m - b * kRowsPerWindow;
      const bool live = (b < a.n_windows) && (t < kFrames);
      mbar_wait(tmem_full, ph_t);
      tc_fence_after();
      for (int ftl = 0; ftl < nft; ++ftl) {
        const int f0 = (4 * g + ftl) * 16;
#pragma unroll 1
        for (int c4 = 0; c4 < 4; ++c4) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ftl * 128 + c4 * 32), v);
          const int f = f0 + c4 * 4;  // 4 bins x 8 channels
          if (live && f < kContourBins) {
            float4* dst = reinterpret_cast<float4*>(a.out + (((size_t)b * kFrames + t) * kContourBins + f) * 8);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 o;
              o.x = fmaxf(__uint_as_float(v[4 * i + 0]) + bias[(4 * i + 0) & 7], 0.f);
              o.y = fmaxf(__uint_as_float(v[4 * i + 1]) + bias[(4 * i + 1) & 7], 0.f);
              o.z = fmaxf(__uint_as_float(v[4 * i + 2]) + bias[(4 * i + 2) & 7], 0.f);
              o.w = fmaxf(__uint_as_float(v[4 * i + 3]) + bias[(4 * i + 3) & 7], 0.f);
              if (f + (i >> 1) < kContourBins) dst[i] = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
      ph_t ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------
int tc_rows_total(int n_windows) {
  const int rows = n_windows * tc::kRowsPerWindow;
  const int n_mtiles = (rows + tc::kMTile - 1) / tc::kMTile;
  return n_mtiles * tc::kMTile + 2;
}

void tc_setup() {
  cudaFuncSetAttribute(contour_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
}

void launch_contour1_tc(const float* y, __nv_bfloat16* yhl, const TcContourDev& dev, const float* bias, float* c1_nhwc,
                        int n_windows, int n_sms, cudaStream_t st) {
  const int rows_total = tc_rows_total(n_windows);
  const long long cells = (long long)rows_total * tc::kChunks8;
  y_split_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(y, yhl, n_windows, rows_total);
  TcArgs a;
  a.yhl = yhl;
  a.tiles = dev.tiles;
  a.tile_seq = dev.tile_seq;
  a.step_use_off = dev.step_use_off;
  a.use_words = dev.use_words;
  a.bias = bias;
  a.out = c1_nhwc;
  for (int g = 0; g <= tc::kGroups; ++g) a.group_step_off[g] = dev.group_step_off[g];
  a.rows_total = rows_total;
  a.n_mtiles = (rows_total - 2) / tc::kMTile;
  a.n_windows = n_windows;
  const int n_units = a.n_mtiles * tc::kGroups;
  const int grid = n_units < n_sms ? n_units : n_sms;
  contour_tc_kernel<<<grid, tc::kThreads, tc::kSmemBytes, st>>>(a);
}

}  // namespace bp
