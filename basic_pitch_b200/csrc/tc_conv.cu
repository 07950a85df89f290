// The three wide convolutions — contour (8 -> 8 channels, 3 x 39 taps, 65 % of the model's FLOPs), onset
// (8 -> 32 channels, 5 x 5 taps, frequency stride 3, 18 %) and note (1 -> 32 channels, 7 x 7, stride 3, 4.5 %) —
// on the 5th-gen tensor cores: tcgen05.mma (kind::f16, bf16 operands, fp32 accumulators in TMEM), operands staged in shared
// memory by bulk async copies (UBLKCP) signalled through mbarriers, warp-specialised roles.
//
// Replaces nodes 231/232 (contour conv + ReLU, reference: basic_pitch/models.py:241-250) and 230/243 (onset
// conv + ReLU, reference: basic_pitch/models.py:295-304) of the deployed graph, and the harmonic stacking in
// front of them (reference: basic_pitch/nn.py:69-88), which is folded into the weight operand and never
// materialised; node 238/239 (note conv1 + ReLU, models.py:270-279) reads the contour posteriorgram instead.
// One kernel template, three specs (TcConvSpec).  The epilogue (EPI 1 / 2 / 3 = onset / note / contour) also reduces
// the FOLLOWING single-output convolution (onset conv2 models.py:305-313, note conv2 :282-290, contour conv2 :254-262)
// over its input channels and frequency taps inside the thread that owns the frame and stores time-tap planes with
// halo columns, so the 8- / 32-channel activations never reach HBM and the second convs degenerate to a sum of 5 / 3 /
// 7 time-tap planes (cnn.cu: halo_tapsum_kernel).  EPI 0 stores the contour activations channels-last (path 2, tests).
//
// Formulation ("Toeplitz along frequency on aligned chunks")
//   rows  m = b*174 + t            time frames of all windows of the chunk, two zero rows between windows
//   D[m][(fl,co)] (+)= A[m+dt][8c .. 8c+15] x T(dt,off)[16][(fl,co)]
//     A      the normalised CQT y itself (NOT the 8-channel stack), rows shifted by the time tap dt; K = 16 bins that
//            start on an 8-bin chunk of the k-chunk-major layout
//     T      16 x 128 "weight tile": T[k][(fl,co)] = sum_ci W[co][ci][dt][df_ci] with
//            df_ci = (8c + k) - shift_ci - SF*(ft*FLT + fl) + PL   (terms outside 0 <= df < KW or outside the stacked
//            image 0 <= g < 264 dropped).  The harmonic channels are shifted views of one image, so they are MERGED
//            in the weight operand: the stack conv is a single-channel conv of y whose taps are the union of the
//            shifted per-channel taps (contour: 8 x 39 = 312 taps on 176 distinct offsets -> 12 instead of 32
//            K-steps per time tap and frequency tile).
//            N = FLT output bins x COUT channels = 128 (contour 16 x 8, onset / note 4 x 32); a tile depends on
//            (dt, 8c - SF*FLT*ft), so frequency tiles SF*FLT*d = 8*j bins apart share tiles (de-duplicated by
//            content)
//   every (ft, dt, c) with a non-empty tile is one K=16 MMA step of shape 128 x 128 x 16
// Precision: both operands are split x = hi + lo (bf16 each) and three products are accumulated
// (hi*hi + hi*lo + lo*hi) in fp32, which keeps the posteriorgrams within ~1e-5 of the FP32 path
// (SURVEY.md Appendix C.4); a single bf16 product would miss the 1e-3 bar.
//
// Work decomposition: item = (M-tile of 128 rows, split s of S over the frequency groups); group = up to 2
// frequency tiles that share weight tiles (2 x 128 TMEM columns; the 512 columns hold two groups, so the
// epilogue of one group overlaps the MMAs of the next).  A CTA (1 per SM, persistent) walks items
// i = blockIdx.x, +gridDim.x, ...:
//   warp 8      producer: bulk-copies the (128+KH-1) x 320 bf16 hi/lo data tile (k-chunk-major) once per item and
//               streams the weight tiles of each group's program (8 KB each) through a 6-stage ring
//   warps 9, 10 MMA issuers, one per accumulator slot of the group (instruction issue, not the tensor pipe, limits
//               a single issuing warp at this MMA size): program words from constant memory, descriptors are
//               base + precomputed offset, 3 x tcgen05.mma per step by one elected lane, tcgen05.commit frees the
//               weight stage / publishes the accumulators
//   warps 0-3, 4-7  epilogue, one warpgroup-like set of 4 warps (= the 4 TMEM lane quadrants) per accumulator slot:
//               tcgen05.ld the accumulator columns, + bias, ReLU, then the fused reduction of the next conv on packed
//               FP32 FMAs and time-fastest stores of the tap planes
#include <cuda_bf16.h>

#include <algorithm>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace bp {

namespace tc {
constexpr int kMTile = 128;
constexpr int kMaxDataBytes = 2 * 40 * (kMTile + 4) * 16;        // hi + lo planes, 40 chunks x 132 rows = 168960
constexpr int kTileBytes = 8192;                                 // weight tile: [plane 2][kchunk 2][128][8] bf16
constexpr int kStages = 6;
constexpr int kMaxSteps = 1024;                                  // program steps per layer (constant memory)
constexpr int kMaxGroups = 15;
constexpr int kThreads = 352;  // 11 warps: 2 x 4 epilogue warps, producer, 2 MMA issuers
// The issue arbiter of an SM sub-partition prefers the highest warp id, so the latency-critical single-thread roles
// (producer, MMA issuers) get the highest ids and are never starved by the FFMA streams of the epilogue warps.
constexpr int kProducerWarp = 8, kMmaWarp0 = 9, kMmaWarp1 = 10;
constexpr int kSmemBytes = kMaxDataBytes + kStages * kTileBytes + 512;
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

//                                      KH KW SF PT PL COUT FLT WOUT n_ci  shifts                              bins ch8 rows/win lead epi taps
TcConvSpec tc_contour_spec() { return {3, 39, 1, 1, 19, 8, 16, 264, 8, {-36, 0, 36, 57, 72, 84, 93, 101}, 309, 40, 174, 2, 0, 0}; }
TcConvSpec tc_onset_spec() { return {5, 5, 3, 2, 1, 32, 4, 88, 8, {-36, 0, 36, 57, 72, 84, 93, 101}, 309, 40, 174, 2, 1, 9}; }
TcConvSpec tc_note_spec() { return {7, 7, 3, 3, 2, 32, 4, 88, 1, {0, 0, 0, 0, 0, 0, 0, 0}, 264, 34, 175, 3, 2, 21}; }

void TcConvPlan::build(const TcConvSpec& sp, const float* W /* [COUT][n_ci][KH][KW] */) {
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
  // K = 16 steps start on 8-bin chunk boundaries (the k-chunk-major layout makes any chunk index a legal
  // descriptor start), so frequency tiles d apart share weight tiles when SF*FLT*d is a multiple of 8 bins
  int stride = 1;
  while ((sp.SF * sp.FLT * stride) % 8 != 0) ++stride;

  // weight tiles are de-duplicated by content (boundary clipping makes otherwise equal keys differ and vice versa)
  std::unordered_map<uint64_t, std::vector<int>> by_hash;
  int n_keys = 0;
  std::vector<uint16_t> scratch(kTileBytes / 2);
  // The n_ci input channels are shifted views of ONE image (harmonic stacking, nn.py:69-88), so the stack conv is a
  // single-channel conv of y with the merged kernel  Wm[co][dt][u] = sum_ci W[co][ci][dt][u - shift_ci + PL]  wherever
  // the stacked pixel exists (0 <= g = u_abs - shift_ci < 264): the contour taps of the 8 harmonics (8 x 39 = 312)
  // cover only 176 distinct bin offsets, the onset clusters of the upper harmonics overlap too.  A tile therefore
  // holds the SUM over the channels: tile for time tap dt, K rows [clip_lo, 16) of the 16 bins that start at bin
  // 8*c, frequency tile ft (rows below clip_lo belong to the previous step when the last chunk pair is clamped).
  auto find_or_add = [&](int dt, int c, int ft, int clip_lo) -> int {
    bool any = false;
    std::fill(scratch.begin(), scratch.end(), (uint16_t)0);
    for (int kk = clip_lo; kk < 16; ++kk) {
      const int u = 8 * c + kk;  // bin of y
      if (u >= sp.data_bins) continue;
      for (int n = 0; n < 128; ++n) {
        const int fl = n / sp.COUT, co = n % sp.COUT;
        const int f = ft * sp.FLT + fl;  // (columns f >= WOUT are computed like the others and dropped by the epilogue)
        double acc = 0.0;
        bool hit = false;
        for (int ci = 0; ci < sp.n_ci; ++ci) {
          const int gg = u - sp.shifts[ci];  // bin of the stacked image
          const int df = gg - sp.SF * f + sp.PL;
          if (df < 0 || df >= sp.KW || gg < 0 || gg >= kContourBins) continue;
          acc += (double)W[((co * sp.n_ci + ci) * sp.KH + dt) * sp.KW + df];
          hit = true;
        }
        if (!hit) continue;
        const float w = (float)acc;
        const uint16_t hi = f2bf(w);
        const uint16_t lo = f2bf(w - bf2f(hi));
        const size_t o = (size_t)(kk >> 3) * 128 * 8 + (size_t)n * 8 + (kk & 7);
        scratch[o] = hi;
        scratch[2048 + o] = lo;
        any = true;
      }
    }
    if (!any) return -1;
    uint64_t h = 1469598103934665603ull;
    for (uint16_t v : scratch) h = (h ^ v) * 1099511628211ull;
    for (int id : by_hash[h])
      if (std::memcmp(tiles.data() + (size_t)id * (kTileBytes / 2), scratch.data(), kTileBytes) == 0) return id;
    tiles.insert(tiles.end(), scratch.begin(), scratch.end());
    by_hash[h].push_back(n_keys);
    return n_keys++;
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
      int tile, slot, c8, dt, off;
    };
    std::vector<Use> uses;
    for (int dt = 0; dt < sp.KH; ++dt) {
      std::vector<Use> cand;
      for (int slot = 0; slot < 2; ++slot) {
        const int ft = fts[slot];
        if (ft < 0) continue;
        // 8-bin blocks of y this frequency tile reads through any channel
        std::vector<bool> need(sp.chunks8, false);
        for (int ci = 0; ci < sp.n_ci; ++ci)
          for (int fl = 0; fl < sp.FLT; ++fl) {
            const int f = ft * sp.FLT + fl;
            if (f >= sp.WOUT) continue;
            for (int df = 0; df < sp.KW; ++df) {
              const int gg = sp.SF * f - sp.PL + df, u = gg + sp.shifts[ci];
              if (gg < 0 || gg >= kContourBins || u < 0 || u >= sp.data_bins) continue;
              need[u / 8] = true;
            }
          }
        // cover the needed blocks with K = 16 steps (two adjacent blocks), left to right
        for (int c8 = 0; c8 < sp.chunks8;) {
          if (!need[c8]) {
            ++c8;
            continue;
          }
          const int c = std::min(c8, sp.chunks8 - 2);  // both k-chunks of the step must exist in the data tile
          const int off = 8 * c - sp.SF * sp.FLT * ft;
          const int tile = find_or_add(dt, c, ft, 8 * (c8 - c));
          if (tile >= 0) cand.push_back(Use{tile, slot, c, dt, off});
          c8 += 2;
        }
      }
      std::stable_sort(cand.begin(), cand.end(), [](const Use& a, const Use& b) {
        return a.off != b.off ? a.off < b.off : (a.tile != b.tile ? a.tile < b.tile : a.slot < b.slot);
      });
      uses.insert(uses.end(), cand.begin(), cand.end());
    }
    bool seen[2] = {false, false};
    size_t i = 0;
    while (i < uses.size()) {
      size_t j = i;
      while (j < uses.size() && uses[j].tile == uses[i].tile && (j == i || uses[j].slot != uses[j - 1].slot)) ++j;
      tile_seq.push_back(uses[i].tile);
      uint32_t w[2] = {kNoUse, kNoUse};
      for (size_t u = i; u < j; ++u) {
        const int sl = uses[u].slot;
        w[sl] = (uint32_t)(uses[u].c8 * lbo16 + uses[u].dt);  // A start offset >> 4: chunk c8, row dt
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
  n_tiles = n_keys;
  n_groups = (int)group_ft.size() / 2;
}

// The MMA programs live in constant memory: the issuing warp indexes them with warp-uniform values, so the words,
// the descriptors derived from them and the loop state stay in uniform registers (no per-use R2UR traffic).
// They depend only on the layer geometry (TcConvSpec), not on the weights.
__constant__ uint32_t c_prog[3][2][tc::kMaxSteps];  // [layer][slot][step]
__constant__ int c_tile_seq[3][tc::kMaxSteps];       // [layer][step] -> weight tile id
__constant__ int c_group_step_off[3][tc::kMaxGroups + 1];
__constant__ int c_group_ft[3][2 * tc::kMaxGroups];
// epilogue constants: conv1 bias and the weights of the fused channel reduction (conv2), [channel][tap]
__constant__ float c_bias1[3][32];
// onset / note conv2 weights as pairs of time taps for the packed FMAs: [channel][df][pair p] = (w2[c][2p][df], w2[c][2p+1][df])
// (onset 3 time taps -> 2 pairs, note 7 -> 4 pairs, the odd last one padded with 0)
__constant__ float2 c_red_onset[32][3][2];
__constant__ float2 c_red_note[32][3][4];
// contour conv2 [dt][channel][6 pairs]: an input bin at even offset bl feeds the output pairs (bl,bl+1), (bl+2,bl+3),
// (bl+4,bl+5) with weights (w4,w3), (w2,w1), (w0,0); at odd bl the pairs (bl-1,bl), (bl+1,bl+2), (bl+3,bl+4) with
// (0,w4), (w3,w2), (w1,w0)   [output offset j = bl + 4 - df]
__constant__ float2 c_red_contour[5][8][6];

void tc_upload_epilogue(const float* contour1_b, const float* onset1_b, const float* note1_b, const float* onset2_w,
                        const float* note2_w, const float* contour2_w, cudaStream_t st) {
  float b[3][32] = {};
  for (int i = 0; i < 8; ++i) b[0][i] = contour1_b[i];
  for (int i = 0; i < 32; ++i) b[1][i] = onset1_b[i], b[2][i] = note1_b[i];
  float2 ro[32][3][2], rn[32][3][4];
  for (int c = 0; c < 32; ++c)
    for (int df = 0; df < 3; ++df) {
      // channel 0 of onset conv2 is the note input (models.py:305: concat[note, onset1]); weights [1][C][KH][3]
      for (int dt = 0; dt < 4; ++dt) (&ro[c][df][0].x)[dt] = dt < 3 ? onset2_w[(1 + c) * 9 + dt * 3 + df] : 0.f;
      for (int dt = 0; dt < 8; ++dt) (&rn[c][df][0].x)[dt] = dt < 7 ? note2_w[c * 21 + dt * 3 + df] : 0.f;
    }
  float2 rc[5][8][6];
  for (int c = 0; c < 8; ++c)
    for (int dt = 0; dt < 5; ++dt) {
      const float* w = contour2_w + (c * 5 + dt) * 5;  // [1][8][5][5], w[df]
      rc[dt][c][0] = make_float2(w[4], w[3]);
      rc[dt][c][1] = make_float2(w[2], w[1]);
      rc[dt][c][2] = make_float2(w[0], 0.f);
      rc[dt][c][3] = make_float2(0.f, w[4]);
      rc[dt][c][4] = make_float2(w[3], w[2]);
      rc[dt][c][5] = make_float2(w[1], w[0]);
    }
  cudaMemcpyToSymbolAsync(c_red_contour, rc, sizeof(rc), 0, cudaMemcpyHostToDevice, st);
  cudaMemcpyToSymbolAsync(c_bias1, b, sizeof(b), 0, cudaMemcpyHostToDevice, st);
  cudaMemcpyToSymbolAsync(c_red_onset, ro, sizeof(ro), 0, cudaMemcpyHostToDevice, st);
  cudaMemcpyToSymbolAsync(c_red_note, rn, sizeof(rn), 0, cudaMemcpyHostToDevice, st);
  cudaStreamSynchronize(st);
}

int tc_upload_program(int layer, const TcConvPlan& pl, cudaStream_t st) {
  if (layer < 0 || layer > 2 || (int)pl.tile_seq.size() > tc::kMaxSteps - 1 || pl.n_groups > tc::kMaxGroups) return -1;
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
// src (fp32, [B][172][bins]) -> bf16 hi/lo planes in the k-chunk-major row layout the MMA reads:
//   dst[plane][q8 (chunks8)][row d (rows_total)][8],  d = lead + b*rows_per_window + t, every other row zero.
// Used for y (309 bins -> 40 chunks) and for the contour posteriorgram (264 bins -> 34 chunks).
// ------------------------------------------------------------------------------------------------
__global__ void split_kernel(const float* __restrict__ src, int bins, __nv_bfloat16* __restrict__ dst, int n_windows,
                             int rows_used, int rows_total /* stride */, int chunks8, int rows_per_window, int lead) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one (row, q8) per thread
  const long long total = (long long)rows_used * chunks8;
  if (idx >= total) return;
  const int d = (int)(idx % rows_used);  // rows fastest: 16-byte stores of a warp are contiguous
  const int q8 = (int)(idx / rows_used);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  const int m = d - lead;
  if (m >= 0) {
    const int b = m / rows_per_window, t = m - b * rows_per_window;
    if (b < n_windows && t < kFrames) {
      const float* p = src + ((size_t)b * kFrames + t) * bins + q8 * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (q8 * 8 + j < bins) v[j] = __ldg(p + j);
    }
  }
  __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    hi[j] = __float2bfloat16_rn(v[j]);
    lo[j] = __float2bfloat16_rn(v[j] - __bfloat162float(hi[j]));
  }
  const size_t plane = (size_t)chunks8 * rows_total * 8;
  const size_t off = ((size_t)q8 * rows_total + d) * 8;
  *reinterpret_cast<uint4*>(dst + off) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(dst + plane + off) = *reinterpret_cast<const uint4*>(lo);
}

// ------------------------------------------------------------------------------------------------
// The tensor-core kernel
// ------------------------------------------------------------------------------------------------
struct TcArgs {
  const __nv_bfloat16* data;    // [2][chunks8][rows_total][8]
  const uint16_t* tiles;        // [n_tiles][8192 B]
  float* out;                   // EPI 0: [B][172][WOUT][COUT] channels-last ; EPI 1/2: [B][taps][WOUT][172] time-fastest
  int layer;                    // which constant-memory program (0 contour, 1 onset, 2 note)
  int rows_total, n_mtiles, n_windows;
  int n_groups, n_split;        // an item covers groups [s*n_groups/n_split, (s+1)*n_groups/n_split)
  int data_rows, row0;          // tile rows (128 + KH - 1); first data row of M-tile 0 (= lead - PT)
  int chunks8, rows_per_window;
  int cout, flt, wout;
};

// Fused second conv of the onset / note branch (32 -> 1 channels, KH x 3 taps, models.py:305-313 / 282-290) in the
// epilogue.  The tile is 4 bins x 32 channels of relu(conv1) for one frame per thread; channels and frequency taps are
// reduced in the thread:
//   Q[dt][j][t] = sum_{c, df} relu(conv1)[c][t][4 ft + j + df - 2] * w2[c][dt][df]     j = 0 .. 5  (bins 4 ft - 1 + j)
// and the time taps (and the two halo columns of the neighbouring tiles) are summed by halo_tapsum_kernel (cnn.cu).
// Q is time-fastest ([B][22 tiles][KH][6][172]): the 32 lanes of a warp hold 32 consecutive frames, so every store
// writes one contiguous 128-byte run; it is half the size of one plane per (dt, df) tap.  Packed FMAs over pairs of
// time taps; the weight pairs are uniform-register operands loaded from constant memory at static offsets (LDCU.128),
// two bins per loaded pair, so no weight lives in a vector register.
template <int LAYER, int KH>
__device__ __forceinline__ void reduce_store(uint32_t taddr, const float2 (&red)[32][3][(KH + 1) / 2],
                                             float* dst /* (b, ft, dt 0, j 0, t) */, bool live) {
  constexpr int NP = (KH + 1) / 2;
  // rolling window over the output offsets: half h (input bins 2h, 2h+1) touches j = 2h .. 2h+3 = accw[0..3]; after it
  // j = 2h and 2h+1 are complete.  The half loop is not unrolled (the weights stay LDCU operands instead of registers).
  float2 accw[4][NP];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int p = 0; p < NP; ++p) accw[k][p] = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    uint32_t v0[32], v1[32];
    tmem_ld32_nowait(taddr + h * 64, v0);
    tmem_ld32_nowait(taddr + h * 64 + 32, v1);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const float o0 = fmaxf(__uint_as_float(v0[c]) + c_bias1[LAYER][c], 0.f);
      const float o1 = fmaxf(__uint_as_float(v1[c]) + c_bias1[LAYER][c], 0.f);
#pragma unroll
      for (int df = 0; df < 3; ++df)
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const float2 w = red[c][df][p];
          if ((KH & 1) && p == NP - 1) {  // odd last time tap: a scalar FMA instead of half an empty pair
            accw[2 - df][p].x = fmaf(o0, w.x, accw[2 - df][p].x);
            accw[3 - df][p].x = fmaf(o1, w.x, accw[3 - df][p].x);
          } else {
            ffma2(accw[2 - df][p], o0, w);  // input bin fl feeds output offset j = fl - df + 2
            ffma2(accw[3 - df][p], o1, w);
          }
        }
    }
    float* d = dst + (size_t)(2 * h) * kFrames;
    if (live) {
#pragma unroll
      for (int dt = 0; dt < KH; ++dt)
#pragma unroll
        for (int k = 0; k < 2; ++k) d[(size_t)(dt * 6 + k) * kFrames] = (dt & 1) ? accw[k][dt >> 1].y : accw[k][dt >> 1].x;
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      accw[0][p] = accw[2][p];
      accw[1][p] = accw[3][p];
      accw[2][p] = accw[3][p] = make_float2(0.f, 0.f);
    }
  }
  if (live) {  // j = 4, 5: the halo columns of the next tile
#pragma unroll
    for (int dt = 0; dt < KH; ++dt)
#pragma unroll
      for (int k = 0; k < 2; ++k)
        dst[(size_t)(dt * 6 + 4 + k) * kFrames] = (dt & 1) ? accw[k][dt >> 1].y : accw[k][dt >> 1].x;
  }
}

// Fused contour conv2 (8 -> 1 channels, 5 x 5 taps, models.py:252-259) in the contour epilogue.  A tile holds 16 bins
// x 8 channels of relu(conv1) for one frame per thread; the channel and frequency taps are reduced in the thread,
//   Q[dt][j][t] = sum_{c, df} relu(conv1)[c][t][16 ft + j - df] * w2[c][dt][df]      j = 0 .. 19  (bins 16 ft - 2 + j),
// and the five time taps are summed by halo_tapsum_kernel (cnn.cu), which also adds the four halo columns of the
// neighbouring tiles.  Q is time-fastest ([B][17 tiles][5][20][172]), so every store is a contiguous 128-byte run,
// and 19 % smaller than the channels-last activations it replaces; the 8-channel image never reaches HBM.
// Pass 0 applies bias + ReLU (and zeroes the bins >= 264 of the last tile) in place in TMEM; the dt loop is not
// unrolled, so only the 40 weights of one time tap are live.
__device__ __forceinline__ void contour_quad(const uint32_t (&v)[32], int dt, float2 (&acc)[10], int c4) {
#pragma unroll
  for (int bl = 0; bl < 4; ++bl)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float o = __uint_as_float(v[bl * 8 + c]);
      const int j2 = (4 * c4 + bl) >> 1;  // pair index of output offsets (2 j2, 2 j2 + 1)
      if ((bl & 1) == 0) {
        ffma2(acc[j2], o, c_red_contour[dt][c][0]);
        ffma2(acc[j2 + 1], o, c_red_contour[dt][c][1]);
        ffma2(acc[j2 + 2], o, c_red_contour[dt][c][2]);
      } else {
        ffma2(acc[j2], o, c_red_contour[dt][c][3]);
        ffma2(acc[j2 + 1], o, c_red_contour[dt][c][4]);
        ffma2(acc[j2 + 2], o, c_red_contour[dt][c][5]);
      }
    }
}

__device__ __forceinline__ void contour_reduce_store(uint32_t taddr, int n_valid /* live columns of the tile */,
                                                     float* dst /* (b, ft, dt 0, j 0, t) */, bool live) {
#pragma unroll 1
  for (int c4 = 0; c4 < 4; ++c4) {
    uint32_t v[32];
    tmem_ld32_nowait(taddr + c4 * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float o = fmaxf(__uint_as_float(v[i]) + c_bias1[0][i & 7], 0.f);
      v[i] = (c4 * 32 + i < n_valid) ? __float_as_uint(o) : 0u;
    }
    tmem_st32(taddr + c4 * 32, v);
  }
  tmem_st_wait();
#pragma unroll 1
  for (int dt = 0; dt < 5; ++dt) {
    float2 acc[10];  // output offsets j = 0 .. 19 as pairs
#pragma unroll
    for (int j = 0; j < 10; ++j) acc[j] = make_float2(0.f, 0.f);
    uint32_t v0[32], v1[32];
    tmem_ld32_nowait(taddr, v0);
    tmem_ld_wait();
    tmem_ld32_nowait(taddr + 32, v1);
    contour_quad(v0, dt, acc, 0);
    tmem_ld_wait();
    tmem_ld32_nowait(taddr + 64, v0);
    contour_quad(v1, dt, acc, 1);
    tmem_ld_wait();
    tmem_ld32_nowait(taddr + 96, v1);
    contour_quad(v0, dt, acc, 2);
    tmem_ld_wait();
    contour_quad(v1, dt, acc, 3);
    if (live) {
#pragma unroll
      for (int j = 0; j < 20; ++j) dst[(size_t)(dt * 20 + j) * kFrames] = (j & 1) ? acc[j >> 1].y : acc[j >> 1].x;
    }
  }
}

template <int EPI>
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

  // Broadcasting the warp index lets the compiler keep the role branches and the producer / MMA loop state in uniform
  // registers (no R2UR before every UTCHMMA: ~40 instead of ~60 instructions per step, measured -5 % on the contour and
  // -3 % on the onset kernel).  The epilogue-bound note kernel measured 20 % slower that way, so it keeps per-thread
  // values.
  constexpr bool kUniformRoles = (EPI != 2);
  const int warp_t = threadIdx.x >> 5;
  const int warp = kUniformRoles ? __shfl_sync(0xffffffffu, warp_t, 0) : warp_t;
  const int lane = threadIdx.x & 31;
  const uint32_t lbo = (uint32_t)a.data_rows * 16u;
  const uint32_t plane_bytes = (uint32_t)a.chunks8 * lbo;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_w + s, 1);
      mbar_init(empty_w + s, 2);  // one arrival per MMA warp
    }
    mbar_init(data_full, 1);
    mbar_init(data_empty, 2);
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full + i, 2);
      mbar_init(tmem_empty + i, 8);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  const int n_items = a.n_mtiles * a.n_split;

  if (warp == kProducerWarp) {
    // ------------------------------ producer ------------------------------
    if (lane == 0) {
      uint32_t stage = 0, ph_w = 0, ph_d = 0;
      const size_t plane_elems = (size_t)a.chunks8 * a.rows_total * 8;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int mt = it / a.n_split, sp = it % a.n_split;
        const int g0 = sp * a.n_groups / a.n_split, g1 = (sp + 1) * a.n_groups / a.n_split;
        mbar_wait(data_empty, ph_d ^ 1);
        mbar_expect_tx(data_full, 2 * plane_bytes);
        const size_t row = (size_t)mt * kMTile + a.row0;
        for (int p = 0; p < 2; ++p)
          for (int c = 0; c < a.chunks8; ++c)
            bulk_g2s(s_data + p * plane_bytes + c * lbo, a.data + p * plane_elems + ((size_t)c * a.rows_total + row) * 8,
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
  } else if (warp == kMmaWarp0 || warp == kMmaWarp1) {
    // ------------------------------ MMA issuers: one warp per accumulator slot ---------------------
    constexpr uint32_t idesc = make_idesc(128, 128);
    const int slot = (warp == kMmaWarp0) ? 0 : 1;
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
    // ------------------------------ epilogue (warps 0..3 -> slot 0, warps 4..7 -> slot 1) ------------------------------
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int slot = warp >> 2;
    const int row = quad * 32 + lane;
    uint32_t ph_t[2] = {0, 0};
    uint32_t gcount = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int mt = it / a.n_split, sp = it % a.n_split;
      const int g0 = sp * a.n_groups / a.n_split, g1 = (sp + 1) * a.n_groups / a.n_split;
      const int m = mt * kMTile + row;
      const int b = m / a.rows_per_window, t = m - b * a.rows_per_window;
      const bool live = (b < a.n_windows) && (t < kFrames);
      for (int g = g0; g < g1; ++g) {
        const uint32_t buf = gcount & 1u;
        mbar_wait(tmem_full + buf, ph_t[buf]);
        ph_t[buf] ^= 1;
        tc_fence_after();
        const int ft = c_group_ft[a.layer][2 * g + slot];
        if (ft >= 0) {
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 256u + (uint32_t)slot * 128u;
          if constexpr (EPI == 0) {
            // contour: 16 bins x 8 channels, bias + ReLU, channels-last rows of 128 contiguous floats
            const int n_valid = min(a.flt, a.wout - ft * a.flt) * a.cout;
            float* dst = a.out + ((size_t)b * kFrames + t) * ((size_t)a.wout * a.cout) + (size_t)ft * 128;
#pragma unroll 1
            for (int c4 = 0; c4 < 4; ++c4) {
              uint32_t v[32];
              tmem_ld32_nowait(taddr + c4 * 32, v);
              tmem_ld_wait();
              if (live) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  float4 o;
                  o.x = fmaxf(__uint_as_float(v[4 * i + 0]) + c_bias1[0][(4 * i + 0) & 7], 0.f);
                  o.y = fmaxf(__uint_as_float(v[4 * i + 1]) + c_bias1[0][(4 * i + 1) & 7], 0.f);
                  o.z = fmaxf(__uint_as_float(v[4 * i + 2]) + c_bias1[0][(4 * i + 2) & 7], 0.f);
                  o.w = fmaxf(__uint_as_float(v[4 * i + 3]) + c_bias1[0][(4 * i + 3) & 7], 0.f);
                  if (c4 * 32 + 4 * i < n_valid) reinterpret_cast<float4*>(dst + c4 * 32)[i] = o;
                }
              }
            }
          } else if constexpr (EPI == 3) {
            const int n_valid = min(a.flt, a.wout - ft * a.flt) * a.cout;
            float* dst = a.out + ((size_t)b * 17 + ft) * (5 * 20 * kFrames) + t;
            contour_reduce_store(taddr, n_valid, dst, live);
          } else {
            // onset / note: the tile is 4 bins x 32 channels; reduce channels and frequency taps of the next conv
            constexpr int KH2 = (EPI == 1) ? 3 : 7;
            float* dst = a.out + ((size_t)b * 22 + ft) * (KH2 * 6 * kFrames) + t;
            if constexpr (EPI == 1) {
              reduce_store<1, 3>(taddr, c_red_onset, dst, live);
            } else {
              reduce_store<2, 7>(taddr, c_red_note, dst, live);
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
  if (warp == kMmaWarp0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------
int tc_rows_total(int n_windows, int rows_per_window) {
  const int rows = n_windows * rows_per_window;
  const int n_mtiles = (rows + tc::kMTile - 1) / tc::kMTile;
  return n_mtiles * tc::kMTile + 8;
}

void tc_setup() {
  cudaFuncSetAttribute(conv_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
  cudaFuncSetAttribute(conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
  cudaFuncSetAttribute(conv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
  cudaFuncSetAttribute(conv_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
}

void launch_split(const float* src, __nv_bfloat16* dst, const TcConvSpec& sp, int n_windows, int rows_stride,
                  cudaStream_t st) {
  const int rows_used = tc_rows_total(n_windows, sp.rows_per_window);  // <= rows_stride
  const long long cells = (long long)rows_used * sp.chunks8;
  split_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(src, sp.data_bins, dst, n_windows, rows_used, rows_stride,
                                                               sp.chunks8, sp.rows_per_window, sp.lead_rows);
}

void launch_conv_tc(const __nv_bfloat16* data, const TcConvDev& dev, float* out, int n_windows, int rows_stride,
                    int n_sms, cudaStream_t st, bool fuse_next) {
  const TcConvSpec& sp = dev.spec;
  TcArgs a;
  a.data = data;
  a.tiles = dev.tiles;
  a.out = out;
  a.layer = dev.layer;
  a.rows_total = rows_stride;  // row stride of the split layout (fixed per model, independent of the batch)
  a.n_mtiles = (tc_rows_total(n_windows, sp.rows_per_window) - 8) / tc::kMTile;
  a.n_windows = n_windows;
  a.n_groups = dev.n_groups;
  // An item is (M-tile, one of `split` runs of frequency groups).  Pick the split that minimises the number of waves
  // times the work per item, the data-tile load counted as half a group: full chunks run unsplit, partial chunks and
  // small batches spread over all SMs.
  int split = 1;
  double best = 1e30;
  for (int s = 1; s <= dev.n_groups; ++s) {
    const int waves = (a.n_mtiles * s + n_sms - 1) / n_sms;
    const double cost = waves * ((dev.n_groups + s - 1) / s + 0.5);
    if (cost < best - 1e-9) best = cost, split = s;
  }
  a.n_split = split;
  a.data_rows = tc::kMTile + sp.KH - 1;
  a.row0 = sp.lead_rows - sp.PT;
  a.chunks8 = sp.chunks8;
  a.rows_per_window = sp.rows_per_window;
  a.cout = sp.COUT;
  a.flt = sp.FLT;
  a.wout = sp.WOUT;
  const int n_items = a.n_mtiles * a.n_split;
  const int grid = n_items < n_sms ? n_items : n_sms;
  if (sp.epi == 0 && fuse_next)
    conv_tc_kernel<3><<<grid, tc::kThreads, tc::kSmemBytes, st>>>(a);
  else if (sp.epi == 0)
    conv_tc_kernel<0><<<grid, tc::kThreads, tc::kSmemBytes, st>>>(a);
  else if (sp.epi == 1)
    conv_tc_kernel<1><<<grid, tc::kThreads, tc::kSmemBytes, st>>>(a);
  else
    conv_tc_kernel<2><<<grid, tc::kThreads, tc::kSmemBytes, st>>>(a);
}

}  // namespace bp
