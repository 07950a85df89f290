// The three wide convolutions — contour (8 -> 8 channels, 3 x 39 taps, 65 % of the model's FLOPs), onset
// (8 -> 32 channels, 5 x 5 taps, frequency stride 3, 18 %) and note (1 -> 32 channels, 7 x 7, stride 3, 4.5 %) —
// on the 5th-gen tensor cores: tcgen05.mma (kind::f16, bf16 operands, fp32 accumulators in TMEM), operands staged in shared
// memory by bulk async copies (UBLKCP) signalled through mbarriers, warp-specialised roles.
//
// Replaces nodes 231/232 (contour conv + ReLU, reference: basic_pitch/models.py:241-250) and 230/243 (onset
// conv + ReLU, reference: basic_pitch/models.py:295-304) of the deployed graph, and the harmonic stacking in
// front of them (reference: basic_pitch/nn.py:69-88), which is folded into the weight operand and never
// materialised; node 238/239 (note conv1 + ReLU, models.py:270-279) reads the contour posteriorgram instead.
// One kernel template, three specs (TcConvSpec).  The epilogue (EPI 1 / 2 / 3 = onset / note / contour) also computes
// the FOLLOWING single-output convolution (onset conv2 models.py:305-313, note conv2 :282-290, contour conv2 :254-262)
// completely, so neither the 8- / 32-channel activations nor any partial sums of them reach HBM:
//   * channels and frequency taps are reduced inside the thread that owns the frame (packed FP32 FMAs),
//   * the time taps are summed across the lanes of the warp (the 32 lanes hold 32 consecutive frames: shuffles) and
//     across the four epilogue warps of an accumulator slot through a small shared-memory exchange; M-tiles overlap
//     by KH2 - 1 rows, so every frame is complete in exactly one tile.  The thread of tile row r finishes the output
//     frame r - H (H = KH2 / 2), so every tap comes from a lane at or below its own: each lane adds its taps in the
//     same order whatever its position in the tile, and a frame's value does not depend on the batch around it,
//   * the frequency halo between neighbouring tiles is a register carry: a slot walks its frequency tiles in ascending
//     order; only where two tile RANGES meet (slot 0 | slot 1, or the group splits of a small batch) the two partial
//     sums go to a small edge buffer and edge_fix_kernel finishes those 4 (contour) / 2 bins,
//   * bias, sigmoid (+ the note input channel of the onset conv2, + unwrap inference.py:247-279) and the store.
// EPI 0 stores the contour activations channels-last (path 2, activation-level tests).
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
// Work decomposition: item = (M-tile of 128 rows, split s of S over the frequency groups); group g = the two
// frequency tiles {g, g + G0} (they share weight tiles; 2 x 128 TMEM columns; the 512 columns hold two groups, so the
// epilogue of one group overlaps the MMAs of the next).  A CTA (1 per SM, persistent) walks items
// i = blockIdx.x, +gridDim.x, ...:
//   warp 8      producer: bulk-copies the (128+KH-1) x 320 bf16 hi/lo data tile (k-chunk-major) once per item and
//               streams the weight tiles of each group's program (8 KB each) through a 6-stage ring
//   warps 9, 10 MMA issuers, one per accumulator slot of the group (instruction issue, not the tensor pipe, limits
//               a single issuing warp at this MMA size): program words from constant memory, descriptors are
//               base + precomputed offset, 3 x tcgen05.mma per step by one elected lane, tcgen05.commit frees the
//               weight stage / publishes the accumulators
//   warps 0-3, 4-7  epilogue, one warpgroup-like set of 4 warps (= the 4 TMEM lane quadrants) per accumulator slot:
//               tcgen05.ld the accumulator columns, + bias, ReLU, then the whole next conv as described above
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
// time-halo exchange between the four epilogue warps of a slot: [slot 2][buffer 2][warp 4][kXchgFloats]
constexpr int kXchgFloats = 200;  // contour: 10 published lane values x 20 output offsets (note 21 x 6, onset 3 x 6)
constexpr int kXchgBytes = 2 * 2 * 4 * kXchgFloats * 4;
constexpr int kSmemBytes = kMaxDataBytes + kStages * kTileBytes + kXchgBytes + 512;
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

//                                      KH KW SF PT PL COUT FLT WOUT n_ci  shifts                              bins ch8 rows/win lead epi KH2 HALO G0
TcConvSpec tc_contour_spec() { return {3, 39, 1, 1, 19, 8, 16, 264, 8, {-36, 0, 36, 57, 72, 84, 93, 101}, 309, 40, 174, 3, 0, 5, 2, 9}; }
TcConvSpec tc_onset_spec() { return {5, 5, 3, 2, 1, 32, 4, 88, 8, {-36, 0, 36, 57, 72, 84, 93, 101}, 309, 40, 174, 3, 1, 3, 1, 12}; }
TcConvSpec tc_note_spec() { return {7, 7, 3, 3, 2, 32, 4, 88, 1, {0, 0, 0, 0, 0, 0, 0, 0}, 264, 34, 175, 6, 2, 7, 1, 12}; }

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
  // descriptor start), so frequency tiles d apart share weight tiles when SF*FLT*d is a multiple of 8 bins; the two
  // tiles of a group are G0 apart: slot 0 walks tiles 0 .. G0-1, slot 1 tiles G0 .. n_ft-1, both in ascending order
  // (the fused epilogue carries the frequency halo of the next conv from tile to tile in registers)
  const int stride = sp.G0;

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

  // groups: pairs {g, g + G0} (or singles)
  group_step_off.push_back(0);
  n_uses = 0;
  for (int ft_a = 0; ft_a < stride; ++ft_a) {
    const int ft_b = ft_a + stride < n_ft ? ft_a + stride : -1;
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
// epilogue constants: conv1 bias, conv2 bias, and the weights of the fused conv2, [channel][tap]
__constant__ float c_bias1[3][32];
__constant__ float c_bias2[3];
// onset / note conv2 weights as pairs of time taps for the packed FMAs: [channel][df][pair p] = (w2[c][2p][df], w2[c][2p+1][df])
// (onset 3 time taps -> 2 pairs, note 7 -> 4 pairs, the odd last one padded with 0)
__constant__ float2 c_red_onset[32][3][2];
__constant__ float2 c_red_note[32][3][4];
// channel 0 of the onset conv2 multiplies the note posteriorgram (models.py:305: concat[note, onset1]): [dt][df]
__constant__ float c_onset_note_w[9];
// contour conv2 [dt][channel][6 pairs]: an input bin at even offset bl feeds the output pairs (bl,bl+1), (bl+2,bl+3),
// (bl+4,bl+5) with weights (w4,w3), (w2,w1), (w0,0); at odd bl the pairs (bl-1,bl), (bl+1,bl+2), (bl+3,bl+4) with
// (0,w4), (w3,w2), (w1,w0)   [output offset j = bl + 4 - df]
__constant__ float2 c_red_contour[5][8][6];

void tc_upload_epilogue(const float* contour1_b, const float* onset1_b, const float* note1_b, const float* onset2_w,
                        const float* note2_w, const float* contour2_w, const float* contour2_b, const float* onset2_b,
                        const float* note2_b, cudaStream_t st) {
  float b[3][32] = {};
  for (int i = 0; i < 8; ++i) b[0][i] = contour1_b[i];
  for (int i = 0; i < 32; ++i) b[1][i] = onset1_b[i], b[2][i] = note1_b[i];
  const float b2[3] = {contour2_b[0], onset2_b[0], note2_b[0]};
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
  cudaMemcpyToSymbolAsync(c_bias2, b2, sizeof(b2), 0, cudaMemcpyHostToDevice, st);
  cudaMemcpyToSymbolAsync(c_red_onset, ro, sizeof(ro), 0, cudaMemcpyHostToDevice, st);
  cudaMemcpyToSymbolAsync(c_red_note, rn, sizeof(rn), 0, cudaMemcpyHostToDevice, st);
  cudaMemcpyToSymbolAsync(c_onset_note_w, onset2_w, 9 * sizeof(float), 0, cudaMemcpyHostToDevice, st);
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
// fp32 rows -> bf16 hi/lo planes in the k-chunk-major row layout the MMA reads:
//   dst[plane][q8 (chunks8)][row d (rows_total)][8],  d = lead + b*rows_per_window + t, every other row zero.
// lognorm_split_kernel: the log-magnitude of the CQT kernel -> NormalizedLog (reference: layers/signal.py:177-183:
//   (L - min) / (max - min), 0 when max == min) + folded BatchNorm (models.py:188-189) as the split operand of the
//   contour / onset convs (309 bins -> 40 chunks).  (The fp32 copy is only produced on request: launch_lognorm.)
// The contour posteriorgram reaches the note conv in the same layout (264 bins -> 34 chunks), written by the contour
// epilogue itself.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_split8(const float (&v)[8], __nv_bfloat16* dst, size_t off, size_t plane) {
  __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    hi[j] = __float2bfloat16_rn(v[j]);
    lo[j] = __float2bfloat16_rn(v[j] - __bfloat162float(hi[j]));
  }
  *reinterpret_cast<uint4*>(dst + off) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(dst + plane + off) = *reinterpret_cast<const uint4*>(lo);
}

__global__ void lognorm_split_kernel(const float* __restrict__ y, const unsigned int* __restrict__ minmax,
                                     const float* __restrict__ bn, __nv_bfloat16* __restrict__ dst, int n_windows,
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
      const float bn_scale = __ldg(bn), bn_bias = __ldg(bn + 1);
      const float mn = ordered_to_float(minmax[2 * b]);
      const float mx = __fsub_rn(ordered_to_float(minmax[2 * b + 1]), mn);
      const float* p = y + ((size_t)b * kFrames + t) * kCqtBins + q8 * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (q8 * 8 + j < kCqtBins) {
          const float q = (mx == 0.f) ? 0.f : __fdiv_rn(__fsub_rn(__ldg(p + j), mn), mx);
          v[j] = __fadd_rn(__fmul_rn(q, bn_scale), bn_bias);
        }
    }
  }
  store_split8(v, dst, ((size_t)q8 * rows_total + d) * 8, (size_t)chunks8 * rows_total * 8);
}

// ------------------------------------------------------------------------------------------------
// The tensor-core kernel
// ------------------------------------------------------------------------------------------------
struct TcArgs {
  const __nv_bfloat16* data;    // [2][chunks8][rows_total][8]
  const uint16_t* tiles;        // [n_tiles][8192 B]
  TcOut o;                      // where the results go (kernels.cuh)
  int edge_rows;                // row stride of o.edge: [edge slot][side 2][KE][edge_rows]
  int layer;                    // which constant-memory program (0 contour, 1 onset, 2 note)
  int rows_total, n_mtiles, n_windows;
  int n_groups, n_split;        // an item covers groups [s*n_groups/n_split, (s+1)*n_groups/n_split)
  int data_rows, row0;          // tile rows (128 + KH - 1); first data row of M-tile 0
  int ms, h2;                   // M-tile stride (128 - 2*h2) and time halo of the fused conv2
  int chunks8, rows_per_window;
  int cout, flt, wout, n_ft, g0;
};

__device__ __forceinline__ float sigmoidf_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

__device__ __forceinline__ void slot_barrier(int slot) {  // the four epilogue warps of one accumulator slot
  asm volatile("bar.sync %0, 128;" ::"r"(1 + slot) : "memory");
}

// Time taps of the fused conv2 across frames.  The thread of tile row r finishes output frame q = r - H of the tile's row
// space: out[q] = sum_dt P_dt[q + dt - H] = sum_a P_{2H-a}[row r - a], a = 0 .. 2H, i.e. every source is `a` lanes BELOW the
// thread.  Sources inside the warp come by shuffle; lanes < a of warps 1..3 take them from the values the previous warp
// published (lanes 32-a .. 31 -> entries a(a-1)/2 + lane - (32-a)) after the slot barrier (time_edges).  A lane adds
// a = 0, 1, .., 2H in this order in both places, so the rounding of a frame does not depend on its row in the tile.
template <int H, int NJ>
__device__ __forceinline__ void time_tap(float (&S)[NJ], const float (&P)[NJ], int a, int lane, float* pub) {
  if (a == 0) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) S[j] += P[j];
    return;
  }
  const bool ok = lane >= a;
  const bool publish = lane >= 32 - a;
  float* e = pub + (a * (a - 1) / 2 + lane - (32 - a)) * NJ;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const float v = __shfl_up_sync(0xffffffffu, P[j], a);
    if (ok) S[j] += v;
    if (publish) e[j] = P[j];
  }
}
template <int H, int NJ>
__device__ __forceinline__ void time_edges(float (&S)[NJ], int quad, int lane, const float* xb /* [4][kXchgFloats] */) {
  if (quad == 0) return;
#pragma unroll
  for (int a = 1; a <= 2 * H; ++a) {
    if (lane < a) {  // source `a` rows below: lane 32 - a + lane of the previous warp
      const float* e = xb + (quad - 1) * tc::kXchgFloats + (a * (a - 1) / 2 + lane) * NJ;
#pragma unroll
      for (int j = 0; j < NJ; ++j) S[j] += e[j];
    }
  }
}

// Edge-buffer slot of the tile range that slot `s` of split `q` walks (it starts at tile g0(q) + s * G0): s * n_split + q.
// The range that ENDS below it is slot s of split q - 1, or, for (s, q) = (1, 0), slot 0 of the last split.

// What the epilogue thread of one frame knows about where its results go.  All stores are coalesced: the 32 lanes of a
// warp hold 32 consecutive frames, and every layout below has the frame index fastest.
//   pitch layers (note / onset)  pitch-major planes  [pitch][frame]
//   contour                      chunk-major         [8-bin chunk][frame][8]  (fp32), same shape as the bf16 hi/lo
//                                split layout [plane][chunk][row][8] that the note conv reads
struct RowOut {
  float* raw;    // pitch layers: raw_pm + b*172 + t        ; contour: raw_cm + (b*172 + t)*8            (nullptr: not stored)
  float* unw;    // pitch layers: unw_pm + unwrapped frame  ; contour: unw_cm + unwrapped frame * 8      (nullptr: not stored)
  __nv_bfloat16* chl;  // contour: chl + data row * 8
  float* edge;   // edge buffer column of this frame: edge + R (nullptr: row not complete / not live)
  const float* note_col;  // EPI 1: note_raw_pm + b*172 + t
  int t;         // frame inside the window (of the OUTPUT frame this thread finishes)
  bool ok;       // live output frame whose time taps are complete in this M-tile
  int e_lo, e_hi;  // edge slots of this range's start and of the range above its end (-1: none)
};

// Time taps of one output column (onset / note layers): acc holds the KH2 per-frame partial sums of the column as pairs
// of time taps; returns the sum over the taps available inside the warp and publishes the top lanes (see time_tap).
template <int KH2>
__device__ __forceinline__ float time_taps_col(const float2 (&acc)[(KH2 + 1) / 2], int lane, float* pub_col /* pub + j */) {
  float s = 0.f;
#pragma unroll
  for (int a = 0; a < KH2; ++a) {
    const int dt = KH2 - 1 - a;
    const float p = (dt & 1) ? acc[dt >> 1].y : acc[dt >> 1].x;
    if (a == 0) {
      s += p;
    } else {
      const float v = __shfl_up_sync(0xffffffffu, p, a);
      if (lane >= a) s += v;
      if (lane >= 32 - a) pub_col[(a * (a - 1) / 2 + lane - (32 - a)) * 6] = p;
    }
  }
  return s;
}

// Fused second conv of the onset / note branch (32 -> 1 channels, KH x 3 taps, models.py:305-313 / 282-290).  The tile
// is 4 bins x 32 channels of relu(conv1) for one frame per thread; channels and frequency taps are reduced in the thread:
//   P[dt][j] = sum_{c, df} relu(conv1)[c][t][4 ft + j + df - 2] * w2[c][dt][df]     j = 0 .. 5  (bins 4 ft - 1 + j)
// and the time taps follow per finished column (time_taps_col): S[j] = sum_dt P[dt][j][frame + dt - H].
// Packed FMAs over pairs of time taps; the weight pairs are uniform-register operands loaded from constant memory at
// static offsets (LDCU.128).  Rolling window over the output offsets: half h (input bins 2h, 2h+1) touches j = 2h .. 2h+3 =
// accw[0..3]; after it j = 2h and 2h+1 are complete.  The half loop is NOT unrolled (unrolled, the compiler keeps the
// weights in vector registers and spills); finished columns are pushed through S like a shift register.
template <int LAYER, int KH>
__device__ __forceinline__ void pitch_tile_sums(uint32_t taddr, const float2 (&red)[32][3][(KH + 1) / 2], bool live, int lane,
                                                float* pub, float (&S)[6]) {
  constexpr int NP = (KH + 1) / 2;
  float2 accw[4][NP];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int p = 0; p < NP; ++p) accw[k][p] = make_float2(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 6; ++j) S[j] = 0.f;
#pragma unroll 1
  for (int h = 0; h < 3; ++h) {
    if (h < 2) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {  // input bin fl = 2h + q feeds output offsets j = fl - df + 2 = 2h + (q + 2 - df)
        uint32_t v[32];
        tmem_ld32_nowait(taddr + (2 * h + q) * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float o = live ? fmaxf(__uint_as_float(v[c]) + c_bias1[LAYER][c], 0.f) : 0.f;
#pragma unroll
          for (int df = 0; df < 3; ++df)
#pragma unroll
            for (int p = 0; p < NP; ++p) {
              const float2 w = red[c][df][p];
              if ((KH & 1) && p == NP - 1) {  // odd last time tap: a scalar FMA instead of half an empty pair
                accw[q + 2 - df][p].x = fmaf(o, w.x, accw[q + 2 - df][p].x);
              } else {
                ffma2(accw[q + 2 - df][p], o, w);
              }
            }
        }
      }
    }
    // columns j = 2h, 2h + 1 are complete (h == 2: the halo columns of the next tile)
    const float s0 = time_taps_col<KH>(accw[0], lane, pub + 2 * h);
    const float s1 = time_taps_col<KH>(accw[1], lane, pub + 2 * h + 1);
#pragma unroll
    for (int j = 0; j < 4; ++j) S[j] = S[j + 2];
    S[4] = s0;
    S[5] = s1;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      accw[0][p] = accw[2][p];
      accw[1][p] = accw[3][p];
      accw[2][p] = accw[3][p] = make_float2(0.f, 0.f);
    }
  }
}

// Frequency halo + finish for the onset / note layers (FLT = 4, halo 1): S[j] is the time-complete sum for bin 4 ft - 1 + j.
template <int EPI>
__device__ __forceinline__ void finish_pitch_tile(const TcArgs& a, const RowOut& ro, float (&S)[6], float (&carry)[2], int ft,
                                                  bool first, bool last) {
  constexpr int L = EPI == 1 ? 1 : 2;
  const bool lower = ft > 0;
  if (first) {
    if (lower && ro.edge) {
      float* e = ro.edge + (size_t)(ro.e_lo * 2 + 0) * 2 * a.edge_rows;
      e[0] = S[0];
      e[a.edge_rows] = S[1];
    }
  } else {
    S[0] += carry[0];
    S[1] += carry[1];
  }
  const int jlo = first ? (lower ? 2 : 1) : 0;
  const int jhi = (ft == a.n_ft - 1) ? 5 : 4;  // the last tile also finishes its top bin (no tile above)
  if (ro.ok) {
    float nv[3][6];
    if constexpr (EPI == 1) {  // note frames t-1 .. t+1, pitches 4 ft - 2 .. 4 ft + 3 (zero outside the image); lanes run along t
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const int f = 4 * ft - 2 + c;
        const bool fin = (unsigned)f < (unsigned)kPitches;
        const float* col = ro.note_col + (size_t)f * a.o.raw_rows;
#pragma unroll
        for (int r = 0; r < 3; ++r)
          nv[r][c] = (fin && (unsigned)(ro.t + r - 1) < (unsigned)kFrames) ? __ldg(col + r - 1) : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      if (j < jlo || j >= jhi) continue;
      const int f = 4 * ft - 1 + j;
      float x = S[j] + c_bias2[L];
      if constexpr (EPI == 1) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int df = 0; df < 3; ++df)
            if (j + df < 6) x = fmaf(nv[r][j + df], c_onset_note_w[r * 3 + df], x);  // pitch f + df - 1 = 4 ft - 2 + (j + df)
      }
      const float v = sigmoidf_fast(x);
      if (ro.raw) ro.raw[(size_t)f * a.o.raw_rows] = v;
      if (ro.unw) ro.unw[(size_t)f * a.o.frame_stride] = v;
    }
  }
  carry[0] = S[4];
  carry[1] = S[5];
  if (last && ft < a.n_ft - 1 && ro.edge && ro.e_hi >= 0) {
    float* e = ro.edge + (size_t)(ro.e_hi * 2 + 1) * 2 * a.edge_rows;
    e[0] = S[4];
    e[a.edge_rows] = S[5];
  }
}

// One finished 8-bin chunk of the contour posteriorgram for one frame: bf16 hi/lo into the operand layout of the note
// conv, fp32 into the chunk-major posteriorgram.
__device__ __forceinline__ void store_contour_chunk(const TcArgs& a, const RowOut& ro, int chunk, const float (&v)[8]) {
  store_split8(v, ro.chl, (size_t)chunk * a.o.chl_rows * 8, (size_t)a.o.chl_chunks * a.o.chl_rows * 8);
  const float4 x0 = make_float4(v[0], v[1], v[2], v[3]), x1 = make_float4(v[4], v[5], v[6], v[7]);
  if (ro.raw) {
    float4* d = reinterpret_cast<float4*>(ro.raw + (size_t)chunk * a.o.raw_rows * 8);
    d[0] = x0;
    d[1] = x1;
  }
  if (ro.unw) {
    float4* d = reinterpret_cast<float4*>(ro.unw + (size_t)chunk * a.o.frame_stride * 8);
    d[0] = x0;
    d[1] = x1;
  }
}

// Fused contour conv2 (8 -> 1 channels, 5 x 5 taps, models.py:252-259) in the contour epilogue.  A tile holds 16 bins
// x 8 channels of relu(conv1) for one frame per thread; the channel and frequency taps are reduced in the thread,
//   P[dt][j] = sum_{c, df} relu(conv1)[c][t][16 ft + j - df] * w2[c][dt][df]      j = 0 .. 19  (bins 16 ft - 2 + j).
// Pass 0 applies bias + ReLU (and zeroes the bins >= 264 of the last tile and frames that are not live) in place in TMEM;
// the dt loop is not unrolled, so only the 40 weights of one time tap are live.
__device__ __forceinline__ void contour_pair(const uint32_t (&v)[16], int dt, float2 (&acc)[10], int c2) {
#pragma unroll
  for (int bl = 0; bl < 2; ++bl)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float o = __uint_as_float(v[bl * 8 + c]);
      const int j2 = c2;  // pair index of output offsets (2 j2, 2 j2 + 1) of input bin 2 c2 + bl
      if (bl == 0) {
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

__device__ __forceinline__ void contour_tile(const TcArgs& a, const RowOut& ro, uint32_t taddr, int n_valid, bool live,
                                             int ft, bool first, bool last, int quad, int lane, int slot, float* xb,
                                             float (&carry)[4], float (&hold)[6]) {
#pragma unroll 1
  for (int c4 = 0; c4 < 4; ++c4) {
    uint32_t v[32];
    tmem_ld32_nowait(taddr + c4 * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float o = fmaxf(__uint_as_float(v[i]) + c_bias1[0][i & 7], 0.f);
      v[i] = (live && c4 * 32 + i < n_valid) ? __float_as_uint(o) : 0u;
    }
    tmem_st32(taddr + c4 * 32, v);
  }
  tmem_st_wait();
  float S[20];
#pragma unroll
  for (int j = 0; j < 20; ++j) S[j] = 0.f;
  float* pub = xb + quad * tc::kXchgFloats;
#pragma unroll 1
  for (int ta = 0; ta < 5; ++ta) {  // source row `ta` below the thread's: time tap dt = 4 - ta (see time_tap)
    const int dt = 4 - ta;
    float2 acc[10];  // output offsets j = 0 .. 19 as pairs
#pragma unroll
    for (int j = 0; j < 10; ++j) acc[j] = make_float2(0.f, 0.f);
    // 16 accumulator columns (2 bins x 8 channels) per load, the next load in flight while this one is consumed
    uint32_t v0[16], v1[16];
    tmem_ld16_nowait(taddr, v0);
#pragma unroll
    for (int c2 = 0; c2 < 8; c2 += 2) {
      tmem_ld_wait();
      tmem_ld16_nowait(taddr + (c2 + 1) * 16, v1);
      contour_pair(v0, dt, acc, c2);
      tmem_ld_wait();
      if (c2 + 2 < 8) tmem_ld16_nowait(taddr + (c2 + 2) * 16, v0);
      contour_pair(v1, dt, acc, c2 + 1);
    }
    float P[20];
#pragma unroll
    for (int j = 0; j < 10; ++j) P[2 * j] = acc[j].x, P[2 * j + 1] = acc[j].y;
    time_tap<2, 20>(S, P, ta, lane, pub);
  }
  __syncwarp();
  slot_barrier(slot);
  time_edges<2, 20>(S, quad, lane, xb);
  // frequency halo: S[j] <-> bin 16 ft - 2 + j; bins 16 ft - 2 .. 16 ft + 1 also get the top four sums of the tile below.
  // Finished bins leave in aligned 8-bin chunks: chunk 2 ft - 1 = the six bins held back from the previous tile + j = 0, 1;
  // chunk 2 ft = j = 2 .. 9; j = 10 .. 15 are held for the next tile.  Where a range starts / ends, the four partial sums
  // AND the six finished bins next to them go to the edge buffer (10 values per side); edge_fix_kernel assembles the two
  // chunks around the boundary.
  const bool lower = ft > 0;
  if (first) {
    if (lower && ro.edge) {
      float* e = ro.edge + (size_t)(ro.e_lo * 2 + 0) * 10 * a.edge_rows;
#pragma unroll
      for (int k = 0; k < 4; ++k) e[(size_t)k * a.edge_rows] = S[k];
#pragma unroll
      for (int k = 0; k < 6; ++k) e[(size_t)(4 + k) * a.edge_rows] = sigmoidf_fast(S[4 + k] + c_bias2[0]);  // bins 16 ft + 2 .. + 7
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) S[k] += carry[k];
  }
  float fin[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) fin[j] = sigmoidf_fast(S[j] + c_bias2[0]);
  if (ro.ok) {
    if (!first) {  // chunk 2 ft - 1: bins 16 ft - 8 .. 16 ft - 1
      const float v[8] = {hold[0], hold[1], hold[2], hold[3], hold[4], hold[5], fin[0], fin[1]};
      store_contour_chunk(a, ro, 2 * ft - 1, v);
    }
    if (!first || !lower) {  // chunk 2 ft: bins 16 ft .. 16 ft + 7 (at a range start above tile 0 the fix-up writes it)
      const float v[8] = {fin[2], fin[3], fin[4], fin[5], fin[6], fin[7], fin[8], fin[9]};
      store_contour_chunk(a, ro, 2 * ft, v);
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) hold[k] = fin[10 + k];
#pragma unroll
  for (int k = 0; k < 4; ++k) carry[k] = S[16 + k];
  if (last && ft < a.n_ft - 1 && ro.edge && ro.e_hi >= 0) {
    float* e = ro.edge + (size_t)(ro.e_hi * 2 + 1) * 10 * a.edge_rows;
#pragma unroll
    for (int k = 0; k < 4; ++k) e[(size_t)k * a.edge_rows] = S[16 + k];
#pragma unroll
    for (int k = 0; k < 6; ++k) e[(size_t)(4 + k) * a.edge_rows] = fin[10 + k];  // bins 16 ft + 8 .. + 13
  }
}

template <int EPI>
__global__ void __launch_bounds__(tc::kThreads, 1) conv_tc_kernel(const TcArgs a) {
  using namespace tc;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* s_data = smem;                    // [2 planes][40 chunks][data_rows][16 B]
  unsigned char* s_w = smem + kMaxDataBytes;       // [kStages][8192]
  float* s_x = reinterpret_cast<float*>(smem + kMaxDataBytes + kStages * kTileBytes);  // [slot][buf][warp][kXchgFloats]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kMaxDataBytes + kStages * kTileBytes + kXchgBytes);
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
        const size_t row = (size_t)mt * a.ms + a.row0;
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
    uint32_t xbuf = 0;  // exchange buffer of this slot, toggled per tile
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int mt = it / a.n_split, sp = it % a.n_split;
      const int g0 = sp * a.n_groups / a.n_split, g1 = (sp + 1) * a.n_groups / a.n_split;
      // conv1 row of this thread (it provides relu(conv1) of that frame to the fused conv2) ...
      const int m = mt * a.ms - a.h2 + row;  // row of the (window, frame) space: m = b * rows_per_window + t
      const int b1 = m >= 0 ? m / a.rows_per_window : 0, t1 = m - b1 * a.rows_per_window;
      const bool live = m >= 0 && (b1 < a.n_windows) && (t1 < kFrames);
      // ... and the output frame it finishes: h2 rows earlier (all time taps of the fused conv2 then lie at or below the
      // thread's own row, see time_tap); rows < 2 h2 of the tile are finished by the previous tile
      const int q = m - a.h2;
      const int b = q >= 0 ? q / a.rows_per_window : 0, t = q - b * a.rows_per_window;
      RowOut ro{};
      float carry[4] = {0.f, 0.f, 0.f, 0.f};
      float hold[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if constexpr (EPI != 0) {
        ro.e_lo = slot * a.n_split + sp;
        ro.e_hi = (sp + 1 < a.n_split) ? slot * a.n_split + sp + 1 : (slot == 0 ? a.n_split : -1);
        ro.ok = row >= 2 * a.h2 && q >= 0 && b < a.n_windows && t < kFrames;
        ro.t = t;
        if (ro.ok) {
          ro.edge = a.o.edge + q;
          int uf = -1;  // unwrapped frame (reference: inference.py:247-279), if this frame is kept
          if (a.o.ud) {
            const UnwrapDesc u = a.o.ud[b];
            const int tt = t - kOverlapHalf;
            if ((unsigned)tt < (unsigned)max(u.rows, 0)) uf = (int)(u.dst_base + tt);
          }
          if constexpr (EPI == 3) {
            ro.chl = a.o.chl + ((size_t)a.o.chl_lead + (size_t)b * a.o.chl_rpw + t) * 8;
            if (a.o.raw) ro.raw = a.o.raw + ((size_t)b * kFrames + t) * 8;
            if (uf >= 0) ro.unw = a.o.unwrapped + (size_t)uf * 8;
          } else {
            if (a.o.raw) ro.raw = a.o.raw + (size_t)b * kFrames + t;
            if (uf >= 0) ro.unw = a.o.unwrapped + uf;
            if constexpr (EPI == 1) ro.note_col = a.o.note_raw + (size_t)b * kFrames + t;
          }
        }
      }
      for (int g = g0; g < g1; ++g) {
        const uint32_t buf = gcount & 1u;
        mbar_wait(tmem_full + buf, ph_t[buf]);
        ph_t[buf] ^= 1;
        tc_fence_after();
        const int ft = c_group_ft[a.layer][2 * g + slot];
        if (ft >= 0) {
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 256u + (uint32_t)slot * 128u;
          // first / last tile of this slot's ascending range inside the item
          const bool first = (g == g0);
          const bool last = (g == g1 - 1) || (ft == a.n_ft - 1);
          if constexpr (EPI == 0) {
            // contour: 16 bins x 8 channels, bias + ReLU, channels-last rows of 128 contiguous floats
            const int n_valid = min(a.flt, a.wout - ft * a.flt) * a.cout;
            float* dst = a.o.act + ((size_t)b * kFrames + t) * ((size_t)a.wout * a.cout) + (size_t)ft * 128;
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
            float* xb = s_x + (slot * 2 + xbuf) * 4 * kXchgFloats;
            contour_tile(a, ro, taddr, n_valid, live, ft, first, last, quad, lane, slot, xb, carry, hold);
            xbuf ^= 1u;
          } else {
            // onset / note: the tile is 4 bins x 32 channels; the whole next conv (32 -> 1, KH2 x 3) follows
            constexpr int KH2 = (EPI == 1) ? 3 : 7, H = KH2 / 2;
            float* xb = s_x + (slot * 2 + xbuf) * 4 * kXchgFloats;
            float S[6];
            if constexpr (EPI == 1) {
              pitch_tile_sums<1, 3>(taddr, c_red_onset, live, lane, xb + quad * kXchgFloats, S);
            } else {
              pitch_tile_sums<2, 7>(taddr, c_red_note, live, lane, xb + quad * kXchgFloats, S);
            }
            __syncwarp();
            slot_barrier(slot);
            time_edges<H, 6>(S, quad, lane, xb);
            xbuf ^= 1u;
            float c2[2] = {carry[0], carry[1]};
            finish_pitch_tile<EPI>(a, ro, S, c2, ft, first, last);
            carry[0] = c2[0];
            carry[1] = c2[1];
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
// Where two tile ranges meet (frequency tile ft_b = first tile of a range, ft_b > 0) the 2 * HALO bins
// FLT * ft_b - HALO + k got one partial sum from each side: finish them here.  Pitch layers: 2 bins per frame.  Contour:
// 4 bins, and with the six finished bins each side left next to them the two 8-bin chunks around the boundary.
// ------------------------------------------------------------------------------------------------
struct EdgeFixArgs {
  TcOut o;
  int edge_rows, n_rows;        // rows of the (window, frame) space covered by the M-tiles
  int n_edges;                  // 2 * n_split slots: slot e = s * n_split + q starts at tile q * n_groups / n_split + s * G0
  int n_split, n_groups, g0, n_ft;
  int layer, wout, rows_per_window, n_windows;
};

__global__ void edge_fix_kernel(const EdgeFixArgs a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per_edge = a.layer == 0 ? 1 : 2;  // contour: one thread per (frame, edge); pitch layers: per (frame, edge, bin)
  const long long total = (long long)a.n_rows * a.n_edges * per_edge;
  if (idx >= total) return;
  const int R = (int)(idx % a.n_rows);  // rows fastest: coalesced reads of the edge buffer, coalesced stores
  const int ek = (int)(idx / a.n_rows);
  const int e = ek / per_edge, k = ek - e * per_edge;
  const int b = R / a.rows_per_window, t = R - b * a.rows_per_window;
  if (b >= a.n_windows || t >= kFrames) return;
  const int es = e / a.n_split, eq = e - es * a.n_split;
  const int ft_b = eq * a.n_groups / a.n_split + es * a.g0;
  if (ft_b <= 0 || ft_b >= a.n_ft) return;  // not a boundary between two ranges
  int uf = -1;
  if (a.o.ud) {
    const UnwrapDesc u = a.o.ud[b];
    const int tt = t - kOverlapHalf;
    if ((unsigned)tt < (unsigned)max(u.rows, 0)) uf = (int)(u.dst_base + tt);
  }
  if (a.layer == 0) {
    const float* lo = a.o.edge + (size_t)(e * 2 + 0) * 10 * a.edge_rows + R;  // from the range that starts at ft_b
    const float* hi = a.o.edge + (size_t)(e * 2 + 1) * 10 * a.edge_rows + R;  // from the range that ends at ft_b - 1
    float fx[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)  // (S + carry) + bias, the order of the in-kernel carry
      fx[q] = sigmoidf_fast((lo[(size_t)q * a.edge_rows] + hi[(size_t)q * a.edge_rows]) + c_bias2[0]);
    float va[8], vb[8];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      va[q] = hi[(size_t)(4 + q) * a.edge_rows];      // bins 16 ft_b - 8 .. - 3
      vb[2 + q] = lo[(size_t)(4 + q) * a.edge_rows];  // bins 16 ft_b + 2 .. + 7
    }
    va[6] = fx[0], va[7] = fx[1], vb[0] = fx[2], vb[1] = fx[3];
    __nv_bfloat16* chl = a.o.chl + ((size_t)a.o.chl_lead + (size_t)b * a.o.chl_rpw + t) * 8;
    const size_t plane = (size_t)a.o.chl_chunks * a.o.chl_rows * 8;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int chunk = 2 * ft_b - 1 + c;
      const float(&v)[8] = c ? vb : va;
      store_split8(v, chl, (size_t)chunk * a.o.chl_rows * 8, plane);
      const float4 x0 = make_float4(v[0], v[1], v[2], v[3]), x1 = make_float4(v[4], v[5], v[6], v[7]);
      if (a.o.raw) {
        float4* d = reinterpret_cast<float4*>(a.o.raw + ((size_t)chunk * a.o.raw_rows + (size_t)b * kFrames + t) * 8);
        d[0] = x0, d[1] = x1;
      }
      if (uf >= 0) {
        float4* d = reinterpret_cast<float4*>(a.o.unwrapped + ((size_t)chunk * a.o.frame_stride + uf) * 8);
        d[0] = x0, d[1] = x1;
      }
    }
    return;
  }
  const int f = 4 * ft_b - 1 + k;
  if (f < 0 || f >= a.wout) return;
  float x = (a.o.edge[((size_t)(e * 2 + 0) * 2 + k) * a.edge_rows + R] +
             a.o.edge[((size_t)(e * 2 + 1) * 2 + k) * a.edge_rows + R]) + c_bias2[a.layer];  // (S + carry) + bias
  if (a.o.note_raw) {
#pragma unroll
    for (int dt = 0; dt < 3; ++dt)
#pragma unroll
      for (int df = 0; df < 3; ++df) {
        const int tt = t + dt - 1, ff = f + df - 1;
        if ((unsigned)tt < (unsigned)kFrames && (unsigned)ff < (unsigned)kPitches)
          x = fmaf(__ldg(a.o.note_raw + (size_t)ff * a.o.raw_rows + (size_t)b * kFrames + tt), c_onset_note_w[dt * 3 + df], x);
      }
  }
  const float v = sigmoidf_fast(x);
  if (a.o.raw) a.o.raw[(size_t)f * a.o.raw_rows + (size_t)b * kFrames + t] = v;
  if (uf >= 0) a.o.unwrapped[(size_t)f * a.o.frame_stride + uf] = v;
}

// ------------------------------------------------------------------------------------------------
int tc_rows_total(int n_windows, int rows_per_window) {
  // lead rows + the rows of the windows + what the last (overlapping) M-tile and its time taps may touch
  return n_windows * rows_per_window + tc::kMTile + 16;
}

void tc_setup() {
  cudaFuncSetAttribute(conv_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
  cudaFuncSetAttribute(conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
  cudaFuncSetAttribute(conv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
  cudaFuncSetAttribute(conv_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
}

void launch_lognorm_split(const float* y, const unsigned int* minmax, const float* bn, __nv_bfloat16* dst,
                          const TcConvSpec& sp, int n_windows, int rows_stride, cudaStream_t st) {
  const int rows_used = tc_rows_total(n_windows, sp.rows_per_window);  // <= rows_stride
  const long long cells = (long long)rows_used * sp.chunks8;
  lognorm_split_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(y, minmax, bn, dst, n_windows, rows_used, rows_stride,
                                                                       sp.chunks8, sp.rows_per_window, sp.lead_rows);
}

size_t tc_edge_floats(const TcConvSpec& sp, int n_windows) {
  const int ms = tc::kMTile - (sp.KH2 - 1);
  const int n_mtiles = (n_windows * sp.rows_per_window + ms - 1) / ms;
  const int per_side = sp.epi == 0 ? 10 : 2;
  return (size_t)2 * sp.G0 * 2 * per_side * ((size_t)n_mtiles * ms);  // at most 2 * G0 range starts
}

void launch_conv_tc(const __nv_bfloat16* data, const TcConvDev& dev, const TcOut& o, int n_windows, int rows_stride,
                    int n_sms, cudaStream_t st, bool fuse_next) {
  const TcConvSpec& sp = dev.spec;
  const bool fused = sp.epi != 0 || fuse_next;
  TcArgs a{};
  a.data = data;
  a.tiles = dev.tiles;
  a.o = o;
  a.layer = dev.layer;
  a.rows_total = rows_stride;  // row stride of the split layout (fixed per model, independent of the batch)
  a.h2 = fused ? (sp.KH2 - 1) / 2 : 0;
  a.ms = tc::kMTile - 2 * a.h2;
  a.n_mtiles = (n_windows * sp.rows_per_window + a.ms - 1) / a.ms;
  a.edge_rows = a.n_mtiles * a.ms;
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
  a.row0 = sp.lead_rows - sp.PT - a.h2;
  a.chunks8 = sp.chunks8;
  a.rows_per_window = sp.rows_per_window;
  a.cout = sp.COUT;
  a.flt = sp.FLT;
  a.wout = sp.WOUT;
  a.n_ft = (sp.WOUT + sp.FLT - 1) / sp.FLT;
  a.g0 = sp.G0;
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
  if (fused) {  // slot s of split q covers tiles [g0(q) + s*G0, g1(q) + s*G0): edge slot s*split + q
    EdgeFixArgs ef{};
    ef.o = o;
    ef.edge_rows = a.edge_rows;
    ef.n_rows = n_windows * sp.rows_per_window;
    ef.n_edges = 2 * split;
    ef.n_split = split;
    ef.n_groups = dev.n_groups;
    ef.g0 = sp.G0;
    ef.n_ft = a.n_ft;
    ef.layer = sp.epi;  // c_bias2 index: 0 contour, 1 onset, 2 note
    ef.wout = sp.WOUT;
    ef.rows_per_window = sp.rows_per_window;
    ef.n_windows = n_windows;
    const long long total = (long long)ef.n_rows * ef.n_edges * (sp.epi == 0 ? 1 : 2);
    edge_fix_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ef);
  }
}

}  // namespace bp
