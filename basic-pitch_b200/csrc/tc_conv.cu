// The two convolutions that read the harmonic stack — contour (8 -> 8 channels, 3 x 39 taps, 65 % of the
// model's FLOPs) and onset (8 -> 32 channels, 5 x 5 taps, frequency stride 3, 18 %) — on the 5th-gen tensor
// cores: tcgen05.mma (kind::f16, bf16 operands, fp32 accumulators in TMEM), operands staged in shared
// memory by bulk async copies (UBLKCP) signalled through mbarriers, warp-specialised roles.
//
// Replaces nodes 231/232 (contour conv + ReLU, reference: basic_pitch/models.py:241-250) and 230/243 (onset
// conv + ReLU, reference: basic_pitch/models.py:295-304) of the deployed graph, and the harmonic stacking in
// front of them (reference: basic_pitch/nn.py:69-88), which is folded into the weight operand and never
// materialised.  One kernel, two specs (TcConvSpec).
//
// Formulation ("Toeplitz along frequency on 16-bin aligned chunks")
//   rows  m = b*174 + t            time frames of all windows of the chunk, two zero rows between windows
//   D[m][(fl,co)] (+)= A[m+dt][16q .. 16q+15] x T(ci,dt,off)[16][(fl,co)]
//     A      the normalised CQT y itself (NOT the 8-channel stack), rows shifted by the time tap dt
//     T      16 x 128 "weight tile": T[k][(fl,co)] = W[co][ci][dt][df] with
//            df = (16q + k) - shift_ci - SF*(ft*FLT + fl) + PL   (zero outside 0 <= df < KW and outside the
//            stacked image 0 <= g < 264: tiles that touch its edges are boundary variants)
//            N = FLT output bins x COUT channels = 128 (contour 16 x 8, onset 4 x 32); a tile depends on
//            (ci, dt, 16q - SF*FLT*ft), so frequency tiles SF*FLT*d = 16*j bins apart share tiles
//   every (ft, ci, dt, q) with a non-empty tile is one K=16 MMA step of shape 128 x 128 x 16
// Precision: both operands are split x = hi + lo (bf16 each) and three products are accumulated
// (hi*hi + hi*lo + lo*hi) in fp32, which keeps the posteriorgrams within ~1e-5 of the FP32 path
// (SURVEY.md Appendix C.4); a single bf16 product would miss the 1e-3 bar.
//
// Work decomposition: item = (M-tile of 128 rows, split s of S over the frequency groups); group = up to 2
// frequency tiles that share weight tiles (2 x 128 TMEM columns; the 512 columns hold two groups, so the
// epilogue of one group overlaps the MMAs of the next).  A CTA (1 per SM, persistent) walks items
// i = blockIdx.x, +gridDim.x, ...:
//   warp 0      producer: bulk-copies the (128+KH-1) x 320 bf16 hi/lo data tile (k-chunk-major) once per item and
//               streams the weight tiles of each group's program (8 KB each) through a 6-stage ring
//   warps 1, 6  MMA issuers, one per accumulator slot of the group (instruction issue, not the tensor pipe, limits
//               a single issuing warp at this MMA size): program words from constant memory, descriptors are
//               base + precomputed offset, 3 x tcgen05.mma per step by one elected lane, tcgen05.commit frees the
//               weight stage / publishes the accumulators
//   warps 2-5   epilogue: tcgen05.ld the accumulator columns, + bias, ReLU, float4 stores along frequency (planar)
#include <cuda_bf16.h>

#include <vector>

#include "kernels.cuh"

namespace bp {

namespace tc {
constexpr int kRowsPerWindow = kFrames + 2;  // 174: two zero separator rows after every window (time pad <= 2)
constexpr int kLeadRows = 2;                 // zero rows in front of the first window
constexpr int kMTile = 128;
constexpr int kMaxDataRows = kMTile + 4;                         // 132 (KH = 5)
constexpr int kChunks8 = 40;                                     // 320 bins / 8
constexpr int kMaxDataBytes = 2 * kChunks8 * kMaxDataRows * 16;  // hi + lo = 168960
constexpr int kTileBytes = 8192;                                 // weight tile: [plane 2][kchunk 2][128][8] bf16
constexpr int kStages = 6;
constexpr int kMaxSteps = 1024;                                  // program steps per layer (constant memory)
constexpr int kMaxGroups = 15;
constexpr int kThreads = 224;
constexpr int kSmemBytes = kMaxDataBytes + kStages * kTileBytes + 512;
constexpr int kShift[kHarmonics] = {-36, 0, 36, 57, 72, 84, 93, 101};
// step word of a slot: [0,14) A start-address offset >> 4, [15] first MMA into that accumulator; kNoUse = the
// slot's frequency tile does not use this step's weight tile
constexpr uint32_t kUseFirstAcc = 1u << 15, kNoUse = 0xffffffffu;
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

TcConvSpec tc_contour_spec() { return TcConvSpec{3, 39, 1, 1, 19, 8, 16, 264}; }
TcConvSpec tc_onset_spec() { return TcConvSpec{5, 5, 3, 2, 1, 32, 4, 88}; }

void TcConvPlan::build(const TcConvSpec& sp, const float* W /* [COUT][8][KH][KW] */) {
  using namespace tc;
  spec = sp;
  tiles.clear();
  tile_seq.clear();
  slot_words[0].clear();
  slot_words[1].clear();
  group_step_off.clear();
  group_ft.clear();
  const int n_ft = (sp.WOUT + sp.FLT - 1) / sp.FLT;
  const int data_rows = kMTile + sp.KH - 1;
  const int lbo16 = data_rows;  // (rows * 16 B) >> 4
  // frequency tiles d apart share weight tiles when SF*FLT*d is a multiple of 16 bins
  int stride = 1;
  while ((sp.SF * sp.FLT * stride) % 16 != 0) ++stride;

  struct Key {
    int ci, dt, off, variant;
  };
  std::vector<Key> keys;
  auto find_or_add = [&](Key k) -> int {
    for (size_t i = 0; i < keys.size(); ++i)
      if (keys[i].ci == k.ci && keys[i].dt == k.dt && keys[i].off == k.off && keys[i].variant == k.variant) return (int)i;
    keys.push_back(k);
    const int s = kShift[k.ci];
    const size_t base = tiles.size();
    tiles.resize(base + kTileBytes / 2, 0);
    for (int kk = 0; kk < 16; ++kk) {
      for (int n = 0; n < 128; ++n) {
        const int fl = n / sp.COUT, co = n % sp.COUT;
        const int df = k.off + kk - s - sp.SF * fl + sp.PL;
        float w = 0.f;
        if (df >= 0 && df < sp.KW) {
          bool keep = true;
          if (k.variant == 1) {  // chunk straddles g = 0: bins u < s lie outside the stacked image
            keep = (16 * (s / 16) + kk - s) >= 0;
          } else if (k.variant == 2) {  // chunk straddles g = 264
            keep = (16 * ((kContourBins + s) / 16) + kk - s) < kContourBins;
          }
          if (keep) w = W[((co * 8 + k.ci) * sp.KH + k.dt) * sp.KW + df];
        }
        const uint16_t hi = f2bf(w);
        const uint16_t lo = f2bf(w - bf2f(hi));
        const size_t o = (size_t)(kk >> 3) * 128 * 8 + (size_t)n * 8 + (kk & 7);
        tiles[base + o] = hi;
        tiles[base + 2048 + o] = lo;
      }
    }
    return (int)keys.size() - 1;
  };

  // groups: pairs {ft, ft + stride} (or singles)
  std::vector<bool> taken(n_ft, false);
  group_step_off.push_back(0);
  n_uses = 0;
  for (int ft_a = 0; ft_a < n_ft; ++ft_a) {
    if (taken[ft_a]) continue;
    taken[ft_a] = true;
    int ft_b = ft_a + stride;
    if (ft_b < n_ft && !taken[ft_b])
      taken[ft_b] = true;
    else
      ft_b = -1;
    const int fts[2] = {ft_a, ft_b};
    group_ft.push_back(ft_a);
    group_ft.push_back(ft_b);
    struct Use {
      int tile, slot, q, dt;
    };
    std::vector<Use> uses;
    for (int ci = 0; ci < 8; ++ci) {
      const int s = kShift[ci];
      for (int dt = 0; dt < sp.KH; ++dt) {
        for (int off = -160; off <= 480; ++off) {
          for (int variant = 0; variant < 3; ++variant) {
            for (int slot = 0; slot < 2; ++slot) {
              const int ft = fts[slot];
              if (ft < 0) continue;
              const int num = off + sp.SF * sp.FLT * ft;
              if (num < 0 || num % 16 != 0) continue;
              const int q = num / 16;
              if (q >= 20) continue;
              bool any = false;
              for (int kk = 0; kk < 16 && !any; ++kk)
                for (int fl = 0; fl < sp.FLT && !any; ++fl) {
                  const int u = 16 * q + kk, f = ft * sp.FLT + fl, gg = u - s;
                  const int df = gg - sp.SF * f + sp.PL;
                  if (df >= 0 && df < sp.KW && gg >= 0 && gg < kContourBins && u < kCqtBins && f < sp.WOUT) any = true;
                }
              if (!any) continue;
              int need = 0;
              if (s > 0 && s % 16 != 0 && q == s / 16) need = 1;
              if (kContourBins + s < 320 && (kContourBins + s) % 16 != 0 && q == (kContourBins + s) / 16) need = 2;
              if (need != variant) continue;
              uses.push_back(Use{find_or_add(Key{ci, dt, off, variant}), slot, q, dt});
            }
          }
        }
      }
    }
    bool seen[2] = {false, false};
    size_t i = 0;
    while (i < uses.size()) {
      size_t j = i;
      while (j < uses.size() && uses[j].tile == uses[i].tile) ++j;
      tile_seq.push_back(uses[i].tile);
      uint32_t w[2] = {kNoUse, kNoUse};
      for (size_t u = i; u < j; ++u) {
        const int sl = uses[u].slot;
        // a (tile, frequency tile) pair determines the chunk q, so a slot uses a tile at most once
        w[sl] = (uint32_t)(2 * uses[u].q * lbo16 + uses[u].dt);
        if (!seen[sl]) w[sl] |= kUseFirstAcc;
        seen[sl] = true;
        ++n_uses;
      }
      slot_words[0].push_back(w[0]);
      slot_words[1].push_back(w[1]);
      i = j;
    }
    group_step_off.push_back((int)tile_seq.size());
  }
  n_tiles = (int)keys.size();
  n_groups = (int)group_ft.size() / 2;
}

// The MMA programs live in constant memory: the issuing warp indexes them with warp-uniform values, so the words,
// the descriptors derived from them and the loop state stay in uniform registers (no per-use R2UR traffic).
// They depend only on the layer geometry (TcConvSpec), not on the weights.
__constant__ uint32_t c_prog[2][2][tc::kMaxSteps];  // [layer][slot][step]
__constant__ int c_tile_seq[2][tc::kMaxSteps];       // [layer][step] -> weight tile id
__constant__ int c_group_step_off[2][tc::kMaxGroups + 1];
__constant__ int c_group_ft[2][2 * tc::kMaxGroups];

int tc_upload_program(int layer, const TcConvPlan& pl, cudaStream_t st) {
  if (layer < 0 || layer > 1 || (int)pl.tile_seq.size() > tc::kMaxSteps - 1 || pl.n_groups > tc::kMaxGroups) return -1;
  for (int sl = 0; sl < 2; ++sl)
    cudaMemcpyToSymbolAsync(c_prog, pl.slot_words[sl].data(), pl.slot_words[sl].size() * 4,
                            ((size_t)layer * 2 + sl) * tc::kMaxSteps * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyToSymbolAsync(c_tile_seq, pl.tile_seq.data(), pl.tile_seq.size() * 4, (size_t)layer * tc::kMaxSteps * 4,
                          cudaMemcpyHostToDevice, st);
  cudaMemcpyToSymbolAsync(c_group_step_off, pl.group_step_off.data(), pl.group_step_off.size() * 4,
                          (size_t)layer * (tc::kMaxGroups + 1) * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyToSymbolAsync(c_group_ft, pl.group_ft.data(), pl.group_ft.size() * 4, (size_t)layer * 2 * tc::kMaxGroups * 4,
                          cudaMemcpyHostToDevice, st);
  return cudaStreamSynchronize(st) == cudaSuccess ? 0 : -1;
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
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
// The three split-precision products of one program use, issued by the elected lane only (PTX predication, no
// branch): D (+)= Ahi*Bhi ; D += Ahi*Blo ; D += Alo*Bhi.  Descriptors are passed as (low word, shared high word).
__device__ __forceinline__ void umma_bf16_x3(uint32_t tmem_d, uint32_t a_hi_lo32, uint32_t a_lo_lo32, uint32_t b_hi_lo32,
                                             uint32_t b_lo_lo32, uint32_t desc_hi32, uint32_t idesc, uint32_t accumulate,
                                             uint32_t leader) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q, t;\n\t"
      ".reg .b64 dah, dal, dbh, dbl;\n\t"
      "setp.ne.b32 p, %7, 0;\n\t"
      "setp.ne.b32 q, %8, 0;\n\t"
      "setp.eq.b32 t, 0, 0;\n\t"
      "mov.b64 dah, {%1, %5};\n\t"
      "mov.b64 dal, {%2, %5};\n\t"
      "mov.b64 dbh, {%3, %5};\n\t"
      "mov.b64 dbl, {%4, %5};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbh, %6, p;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbl, %6, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dal, dbh, %6, t;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(a_hi_lo32), "r"(a_lo_lo32), "r"(b_hi_lo32), "r"(b_lo_lo32), "r"(desc_hi32), "r"(idesc), "r"(accumulate),
      "r"(leader)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pred(uint64_t* bar, uint32_t leader) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(leader)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// y (fp32, [B][172][309]) -> bf16 hi/lo planes in the k-chunk-major row layout the MMA reads:
//   yhl[plane][q8 (40)][row d (rows_total)][8],  d = 2 + b*174 + t, every other row zero.
// ------------------------------------------------------------------------------------------------
__global__ void y_split_kernel(const float* __restrict__ y, __nv_bfloat16* __restrict__ yhl, int n_windows,
                               int rows_total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one (row, q8) per thread
  const long long total = (long long)rows_total * tc::kChunks8;
  if (idx >= total) return;
  const int d = (int)(idx % rows_total);  // rows fastest: 16-byte stores of a warp are contiguous
  const int q8 = (int)(idx / rows_total);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  const int m = d - tc::kLeadRows;
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
  const float* bias;            // [COUT]
  int layer;                    // which constant-memory program (0 contour, 1 onset)
  float* out;                   // [B][COUT][172][WOUT]  (planar, like the FP32 path)
  int rows_total, n_mtiles, n_windows;
  int n_groups, n_split;        // an item covers groups [s*n_groups/n_split, (s+1)*n_groups/n_split)
  int data_rows, row0;          // tile rows (128 + KH - 1); first data row of M-tile 0 (= 2 - PT)
  int cout, flt, wout;
};

__global__ void __launch_bounds__(tc::kThreads, 1) conv_tc_kernel(const TcArgs a) {
  using namespace tc;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* s_data = smem;                    // [2 planes][40 chunks][data_rows][16 B]
  unsigned char* s_w = smem + kMaxDataBytes;       // [kStages][8192]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kMaxDataBytes + kStages * kTileBytes);
  uint64_t* full_w = bars;             // [kStages]
  uint64_t* empty_w = bars + kStages;  // [kStages]
  uint64_t* data_full = bars + 2 * kStages;
  uint64_t* data_empty = data_full + 1;
  uint64_t* tmem_full = data_full + 2;   // [2]
  uint64_t* tmem_empty = data_full + 4;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(data_full + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t lbo = (uint32_t)a.data_rows * 16u;
  const uint32_t plane_bytes = kChunks8 * lbo;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_w + s, 1);
      mbar_init(empty_w + s, 2);  // one arrival per MMA warp
    }
    mbar_init(data_full, 1);
    mbar_init(data_empty, 2);
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full + i, 2);
      mbar_init(tmem_empty + i, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  const int n_items = a.n_mtiles * a.n_split;

  if (warp == 0) {
    // ------------------------------ producer ------------------------------
    if (lane == 0) {
      uint32_t stage = 0, ph_w = 0, ph_d = 0;
      const size_t plane_elems = (size_t)kChunks8 * a.rows_total * 8;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int mt = it / a.n_split, sp = it % a.n_split;
        const int g0 = sp * a.n_groups / a.n_split, g1 = (sp + 1) * a.n_groups / a.n_split;
        mbar_wait(data_empty, ph_d ^ 1);
        mbar_expect_tx(data_full, 2 * plane_bytes);
        const size_t row = (size_t)mt * kMTile + a.row0;
        for (int p = 0; p < 2; ++p)
          for (int c = 0; c < kChunks8; ++c)
            bulk_g2s(s_data + p * plane_bytes + c * lbo, a.yhl + p * plane_elems + ((size_t)c * a.rows_total + row) * 8,
                     lbo, data_full);
        ph_d ^= 1;
        const int s0 = c_group_step_off[a.layer][g0], s1 = c_group_step_off[a.layer][g1];
        for (int s = s0; s < s1; ++s) {
          mbar_wait(empty_w + stage, ph_w ^ 1);
          mbar_expect_tx(full_w + stage, kTileBytes);
          bulk_g2s(s_w + stage * kTileBytes, a.tiles + (size_t)c_tile_seq[a.layer][s] * (kTileBytes / 2), kTileBytes,
                   full_w + stage);
          if (++stage == kStages) {
            stage = 0;
            ph_w ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 || warp == 6) {
    // ------------------------------ MMA issuers: warp 1 -> accumulator slot 0, warp 6 -> slot 1 ---------------------
    constexpr uint32_t idesc = make_idesc(128, 128);
    const int slot = (warp == 1) ? 0 : 1;
    const uint32_t leader = elect_one() ? 1u : 0u;
    uint32_t stage = 0, ph_w = 0, ph_d = 0;
    uint32_t ph_t[2] = {0, 0};
    uint32_t gcount = 0;  // groups issued so far by this CTA -> TMEM buffer = gcount & 1
    // descriptor words: low = start >> 4 | (LBO >> 4) << 16 ; high = SBO >> 4 | version 1 << 14 (shared by all)
    const uint32_t desc_hi32 = (128u >> 4) | (1u << 14);
    const uint32_t a_hi_base = ((smem_u32(s_data) >> 4) & 0x3fffu) | ((uint32_t)a.data_rows << 16);
    const uint32_t a_lo_base = a_hi_base + (plane_bytes >> 4);
    const uint32_t b_base = ((smem_u32(s_w) >> 4) & 0x3fffu) | ((2048u >> 4) << 16);
    const uint32_t* prog = c_prog[a.layer][slot];
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int sp = it % a.n_split;
      const int g0 = sp * a.n_groups / a.n_split, g1 = (sp + 1) * a.n_groups / a.n_split;
      mbar_wait(data_full, ph_d);
      ph_d ^= 1;
      for (int g = g0; g < g1; ++g) {
        const uint32_t buf = gcount & 1u;
        mbar_wait(tmem_empty + buf, ph_t[buf] ^ 1);
        ph_t[buf] ^= 1;
        tc_fence_after();
        const int s0 = c_group_step_off[a.layer][g], s1 = c_group_step_off[a.layer][g + 1];
        const uint32_t d = tmem_base + buf * 256u + (uint32_t)slot * 128u;
        uint32_t w = prog[s0];
        for (int s = s0; s < s1; ++s) {
          const uint32_t w_next = prog[s + 1];  // (one word past the end is inside the array)
          mbar_wait(full_w + stage, ph_w);
          if (w != kNoUse) {
            tc_fence_after();
            const uint32_t off = w & 0x3fffu;
            const uint32_t bl = b_base + stage * (kTileBytes >> 4);
            umma_bf16_x3(d, a_hi_base + off, a_lo_base + off, bl, bl + 256u, desc_hi32, idesc,
                         (w & kUseFirstAcc) ? 0u : 1u, leader);
            umma_commit_pred(empty_w + stage, leader);
          } else if (leader) {
            mbar_arrive(empty_w + stage);
          }
          if (++stage == kStages) {
            stage = 0;
            ph_w ^= 1;
          }
          w = w_next;
        }
        umma_commit_pred(tmem_full + buf, leader);  // this slot's accumulator is complete
        ++gcount;
      }
      umma_commit_pred(data_empty, leader);  // the data tile may be overwritten
    }
  } else {
    // ------------------------------ epilogue (warps 2..5) ------------------------------
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    float* s_bias = reinterpret_cast<float*>(tmem_slot + 2);  // 32 floats behind the barriers
    if (warp == 2) s_bias[lane] = __ldg(a.bias + (lane % a.cout));
    asm volatile("bar.sync 1, 128;" ::: "memory");
    uint32_t ph_t[2] = {0, 0};
    uint32_t gcount = 0;
    const size_t chan_pitch = (size_t)kFrames * a.wout;  // planar output [B][COUT][172][WOUT]
    const int cshift = (a.cout == 8) ? 0 : 1;           // column of (fl, co): COUT = 8 -> fl*8 + co ; 32 -> fl*32 + co
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int mt = it / a.n_split, sp = it % a.n_split;
      const int g0 = sp * a.n_groups / a.n_split, g1 = (sp + 1) * a.n_groups / a.n_split;
      const int m = mt * kMTile + row;
      const int b = m / kRowsPerWindow, t = m - b * kRowsPerWindow;
      const bool live = (b < a.n_windows) && (t < kFrames);
      float* orow = a.out + (size_t)b * a.cout * chan_pitch + (size_t)t * a.wout;
      for (int g = g0; g < g1; ++g) {
        const uint32_t buf = gcount & 1u;
        mbar_wait(tmem_full + buf, ph_t[buf]);
        ph_t[buf] ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int slot = 0; slot < 2; ++slot) {
          const int ft = c_group_ft[a.layer][2 * g + slot];
          if (ft < 0) continue;
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 256u + (uint32_t)slot * 128u;
          // one pass = 4 consecutive bins x 8 channels: 4 passes cover the 128 columns of the tile
#pragma unroll 1
          for (int pass = 0; pass < 4; ++pass) {
            uint32_t v[32];  // v[fq * 8 + c] : bin quad index fq, channel c of the block
            int f, co0;
            if (cshift == 0) {  // COUT = 8: 32 consecutive columns are 4 bins x 8 channels
              tmem_ld32_nowait(taddr + pass * 32, v);
              f = ft * 16 + pass * 4;
              co0 = 0;
            } else {  // COUT = 32: the tile is 4 bins x 32 channels; take 8 channels of each bin
              uint32_t(&v0)[8] = *reinterpret_cast<uint32_t(*)[8]>(&v[0]);
              uint32_t(&v1)[8] = *reinterpret_cast<uint32_t(*)[8]>(&v[8]);
              uint32_t(&v2)[8] = *reinterpret_cast<uint32_t(*)[8]>(&v[16]);
              uint32_t(&v3)[8] = *reinterpret_cast<uint32_t(*)[8]>(&v[24]);
              tmem_ld8_nowait(taddr + 0 * 32 + pass * 8, v0);
              tmem_ld8_nowait(taddr + 1 * 32 + pass * 8, v1);
              tmem_ld8_nowait(taddr + 2 * 32 + pass * 8, v2);
              tmem_ld8_nowait(taddr + 3 * 32 + pass * 8, v3);
              f = ft * 4;
              co0 = pass * 8;
            }
            tmem_ld_wait();
            if (live && f < a.wout) {
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float bv = s_bias[(co0 + c) & 31];
                float4 o;
                o.x = fmaxf(__uint_as_float(v[0 * 8 + c]) + bv, 0.f);
                o.y = fmaxf(__uint_as_float(v[1 * 8 + c]) + bv, 0.f);
                o.z = fmaxf(__uint_as_float(v[2 * 8 + c]) + bv, 0.f);
                o.w = fmaxf(__uint_as_float(v[3 * 8 + c]) + bv, 0.f);
                *reinterpret_cast<float4*>(orow + (size_t)(co0 + c) * chan_pitch + f) = o;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty + buf);
        ++gcount;
      }
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
  return n_mtiles * tc::kMTile + 4;
}

void tc_setup() {
  cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
}

void launch_y_split(const float* y, __nv_bfloat16* yhl, int n_windows, cudaStream_t st) {
  const int rows_total = tc_rows_total(n_windows);
  const long long cells = (long long)rows_total * tc::kChunks8;
  y_split_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(y, yhl, n_windows, rows_total);
}

void launch_conv_tc(const __nv_bfloat16* yhl, const TcConvDev& dev, const float* bias, float* out_nhwc, int n_windows,
                    int n_sms, cudaStream_t st) {
  const TcConvSpec& sp = dev.spec;
  TcArgs a;
  a.yhl = yhl;
  a.tiles = dev.tiles;
  a.tile_seq = dev.tile_seq;
  a.bias = bias;
  a.layer = dev.layer;
  a.out = out_nhwc;
  a.rows_total = tc_rows_total(n_windows);
  a.n_mtiles = (a.rows_total - 4) / tc::kMTile;
  a.n_windows = n_windows;
  a.n_groups = dev.n_groups;
  // few M-tiles (small batches): split the frequency groups of an M-tile over several CTAs
  int split = 1;
  if (a.n_mtiles < n_sms) split = (n_sms + a.n_mtiles - 1) / a.n_mtiles;
  if (split > dev.n_groups) split = dev.n_groups;
  a.n_split = split;
  a.data_rows = tc::kMTile + sp.KH - 1;
  a.row0 = tc::kLeadRows - sp.PT;
  a.cout = sp.COUT;
  a.flt = sp.FLT;
  a.wout = sp.WOUT;
  const int n_items = a.n_mtiles * a.n_split;
  const int grid = n_items < n_sms ? n_items : n_sms;
  conv_tc_kernel<<<grid, tc::kThreads, tc::kSmemBytes, st>>>(a);
}

}  // namespace bp
