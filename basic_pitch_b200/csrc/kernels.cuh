// Launch wrappers of the hot-path kernels (host-callable; all asynchronous on `st`).
#pragma once
#include <cuda_bf16.h>

#include <cstring>
#include <vector>

#include "common.cuh"

namespace bp {

// ---- hcqt.cu ------------------------------------------------------------------------------------
void hcqt_setup();  // per device: shared-memory opt-in of the FP32 CQT kernel
void upload_lowpass(const float* h_lp, cudaStream_t st);  // host taps -> __constant__ tap pairs
void launch_decimate(const float* audio, const WinDesc* desc, float* chain, int stage, int n_windows, cudaStream_t st);
void launch_cqt(const float* audio, const WinDesc* desc, const float* chain, const float* wt, const float* scale,
                float* logmag, unsigned int* minmax, int n_windows, cudaStream_t st);
void launch_lognorm(float* y, const unsigned int* minmax, const float* bn /* device: scale, bias */, int n_windows,
                    cudaStream_t st);

// ---- cnn.cu (FP32 FFMA path) -------------------------------------------------------------------
// Transposed weights wT[(ci*KH+dt)*KW+df][COUT]; activations are planar [B][C][172][W].
struct CnnWeights {
  const float *contour1_wT, *contour1_b;
  const float *contour2_wT, *contour2_b;
  const float *note1_wT, *note1_b;
  const float *note2_wT, *note2_b;
  const float *onset1_wT, *onset1_b;
  const float *onset2_wT, *onset2_b;
};
void cnn_setup();
void launch_contour1(const float* y, const CnnWeights& w, float* c1, int n_windows, cudaStream_t st);
void launch_contour2(const float* c1, const CnnWeights& w, float* contour, int n_windows, cudaStream_t st);
void launch_note1(const float* contour, const CnnWeights& w, float* n1, int n_windows, cudaStream_t st);
void launch_note2(const float* n1, const CnnWeights& w, float* note, int n_windows, cudaStream_t st);
void launch_onset1(const float* y, const CnnWeights& w, float* o1, int n_windows, cudaStream_t st);
void launch_onset2(const float* note, const float* o1, const CnnWeights& w, float* onset, int n_windows,
                   cudaStream_t st);


// ---- tc_conv.cu (tcgen05 path of the three wide convolutions, each with its following conv fused) ------------
struct TcConvSpec {
  int KH, KW, SF, PT, PL, COUT, FLT, WOUT;  // taps, frequency stride, pads, channels, bins per 128-column tile, output bins
  int n_ci;                                 // input channels (harmonics of y, or 1 for the contour posteriorgram)
  int shifts[8];                            // frequency shift of every input channel (harmonic stacking)
  int data_bins, chunks8;                   // bins of the input rows and 8-bin chunks of the split layout (even)
  int rows_per_window, lead_rows;           // row layout of the split input: 172 frames + zero rows per window
  int epi;                                  // 0 contour, 1 onset, 2 note
  int KH2, HALO, G0;                        // fused next conv: time taps, frequency halo; tiles of a group are G0 apart
};
TcConvSpec tc_contour_spec();
TcConvSpec tc_onset_spec();
TcConvSpec tc_note_spec();
struct TcConvPlan {  // host side: weight tiles + the per-group MMA programs
  TcConvSpec spec{};
  std::vector<uint16_t> tiles;          // n_tiles x 4096 bf16 : [plane hi/lo][k-chunk 2][n 128][8]
  std::vector<int> tile_seq;            // per step: tile id
  std::vector<uint32_t> slot_words[2];  // per step and accumulator slot: A offset >> 4 | first << 15, or 0xffffffff
  std::vector<int> group_step_off;      // [n_groups + 1]
  std::vector<int> group_ft;            // [n_groups][2] frequency tiles of the group (-1 = none)
  int n_tiles = 0, n_groups = 0, n_uses = 0;
  void build(const TcConvSpec& spec, const float* w /* [COUT][n_ci][KH][KW] */);
};
struct TcConvDev {
  TcConvSpec spec;
  const uint16_t* tiles;
  const uint16_t* b2;  // conv2 weight tiles of the fused epilogue (tc_build_b2)
  int n_groups;
  int layer;  // index of the program in constant memory (0 contour, 1 onset, 2 note)
};
int tc_upload_program(int layer, const TcConvPlan& plan, cudaStream_t st);  // 0 on success
void tc_upload_epilogue(const float* contour1_b, const float* onset1_b, const float* note1_b, const float* onset2_w,
                        const float* contour2_b, const float* onset2_b, const float* note2_b, cudaStream_t st);
// bf16 hi/lo weight tiles of the fused second conv (epi: 0 contour, 1 onset, 2 note; w2 = that conv's weights), see TcB2
void tc_build_b2(int epi, const float* w2, std::vector<uint16_t>& out);
int tc_rows_total(int n_windows, int rows_per_window);
size_t tc_edge_floats(const TcConvSpec& spec, int n_windows);  // size of the edge buffer of a fused layer
int tc_setup();  // 0 on success
// Where window w's centre frames go in the unwrapped (per-file) posteriorgrams (reference: inference.py:247-279).
struct UnwrapDesc {
  long long dst_base;  // first output frame this window contributes to
  int rows;            // how many of its 142 centre frames are kept (may be <= 0)
  int pad;
};
// NormalizedLog + folded BatchNorm of the CQT log-magnitudes straight into the bf16 hi/lo split operand (k-chunk-major
// row layout of tc_contour_spec() / tc_onset_spec()); `y` itself is left untouched (launch_lognorm makes the fp32 copy).
// rows_stride: row stride of the split layout (>= tc_rows_total(n_windows, ...)); fixed per model so that rows the
// kernels never write (separators, pads) keep their zeros across batches of different size
void launch_lognorm_split(const float* y, const unsigned int* minmax, const float* bn, __nv_bfloat16* dst,
                          const TcConvSpec& spec, int n_windows, int rows_stride, cudaStream_t st);
// Outputs of a fused layer.  Device-internal posteriorgram layouts have the FRAME index fastest, so that the epilogue
// threads (one frame each, 32 consecutive frames per warp) store coalesced, and so that the decode kernels read coalesced:
//   note / onset   pitch-major   pm[pitch][frame]                 frame stride = number of frames the buffer holds
//   contour        chunk-major   cm[8-bin chunk (33)][frame][8]
// `raw` holds all 172 frames of every window of the chunk (frame index b*172 + t, stride `raw_rows`); with `ud` the centre
// frames of every window also / instead go to their unwrapped position (reference: inference.py:247-279) in `unwrapped`,
// whose frame stride is `frame_stride`.
struct TcOut {
  float* raw = nullptr;
  float* unwrapped = nullptr;
  long long frame_stride = 0;
  long long raw_rows = 0;           // frame stride of `raw` and `note_raw`
  const UnwrapDesc* ud = nullptr;
  const float* note_raw = nullptr;  // onset layer: the raw pitch-major note posteriorgram (input channel 0 of its conv2)
  __nv_bfloat16* chl = nullptr;     // contour layer: split operand of the note conv [2][chl_chunks][chl_rows][8]
  int chl_rows = 0, chl_chunks = 0, chl_rpw = 0, chl_lead = 0;
  float* edge = nullptr;            // >= tc_edge_floats(spec, n_windows) floats of scratch
  float* act = nullptr;             // EPI 0 (path 2): channels-last activations [B][172][WOUT][COUT]
};
// fuse_next (contour layer only; the other two always fuse): also compute the following conv in the epilogue instead
// of storing the channels-last activations
void launch_conv_tc(const __nv_bfloat16* data, const TcConvDev& dev, const TcOut& out, int n_windows, int rows_stride,
                    int n_sms, cudaStream_t st, bool fuse_next = false);
// contour conv2 on the channels-last output of the tensor-core contour conv; also emits the bf16 hi/lo split of the
// contour posteriorgram in the layout of tc_note_spec()
void launch_contour2_tc(const float* c1_nhwc, const CnnWeights& w, float* contour, __nv_bfloat16* chl, int rows_total,
                        int n_windows, cudaStream_t st);

// ---- cqt_tc.cu (tcgen05 path of the constant-Q projection, three-way bf16 split) ----------------------------
void build_cqt_tc_weights(const float* cqt_real, const float* cqt_imag, std::vector<uint16_t>& out);
void cqt_tc_setup();
void launch_cqt_tc(const float* audio, const WinDesc* desc, const float* chain, const uint16_t* wtc, const float* scale,
                   float* logmag, unsigned int* minmax, int n_windows, int n_sms, cudaStream_t st);

// ---- layout.cu: internal frame-fastest layouts (pm / cm, see TcOut) <-> row-major [frame][bins] of the C ABI ------
// rows -> internal: without `ud` frames [0, n_frames) go to dst_frame0 ..; with `ud` `rows` is a chunk of raw windows
// [n_windows][172][bins] and the kept centre frames of every window go to their unwrapped position
void launch_rows_to_pm(const float* rows, long long n_frames, int width, float* pm, long long stride, long long dst_frame0,
                       cudaStream_t st, const UnwrapDesc* ud = nullptr, int n_windows = 0);
void launch_rows_to_cm(const float* rows, long long n_frames, float* cm, long long stride, long long dst_frame0,
                       cudaStream_t st, const UnwrapDesc* ud = nullptr, int n_windows = 0);
void launch_pm_to_rows(const float* pm, long long stride, long long src_frame0, long long n_frames, int width, float* rows,
                       cudaStream_t st);
void launch_cm_to_rows(const float* cm, long long stride, long long src_frame0, long long n_frames, float* rows,
                       cudaStream_t st);

// ---- ingest.cu: PCM of any rate / channel count -> mono float32 at 22 050 Hz (format: 0 f32, 1 s16, 2 s32, 3 u8) ----
long long ingest_output_length(long long n_frames, int sample_rate);
std::vector<double> ingest_filter(int up, int down);  // host-only: the low-pass design (unit DC gain)
int launch_ingest(int device, const void* d_pcm, int format, long long n_frames, int channels, int sample_rate, float* d_out,
                  cudaStream_t st);  // 0 ok, -1 CUDA error, -2 unsupported argument

// ---- decode.cu ----------------------------------------------------------------------------------
struct DecodeParamsDev {
  double onset_thresh, frame_thresh;
  int min_note_len, energy_tol, infer_onsets, melodia, lo_col, hi_col;
};
struct DecodeBuffers {
  const long long* frame_off;     // [n_files+1] device; file i covers frames [frame_off[i], frame_off[i+1])
  float* energy;                  // [total_frames*88] "remaining energy", column-major per file
  unsigned int* candbits;         // [(total_frames*88+31)/32 + 1] onset-candidate bit per cell
  unsigned int* max_onset;        // [n_files]
  unsigned long long* max_fd;     // [n_files]
  const long long* slot_off;      // [n_files+1] device; note slots of file i
  int* note_count;                // [n_files]
  int* note_start;                // [slots]
  int* note_end;
  int* note_pitch;
  int* overflow;                  // [1] set when a file ran out of slots
  float* blk_max;                 // [88 * decode_block_slots(total_frames, n_files)] per-column maxima of 256-frame blocks
  int* blk_arg;                   //   (melodia loop of long files) and their frame indices
};
long long decode_block_slots(long long total_frames, int n_files);
void launch_decode_notes(const float* note, const float* onset, const DecodeBuffers& buf, int n_files,
                         long long total_frames, const DecodeParamsDev& p, cudaStream_t st);
void launch_infer_onsets(const float* note, const float* onset, const DecodeBuffers& buf, int n_files, long long total_frames,
                         double* out64 /* [total_frames][88] */, cudaStream_t st);
// amplitude (NumPy pairwise mean) + pitch bends for compacted notes
void launch_note_finish(const float* note, const float* contour, const long long* note_frame_base /*[n_notes]*/,
                        const int* start, const int* end, const int* pitch, float* amp, const int* bend_off,
                        int* bends, int n_notes, int with_bends, const double* gauss /*[51] device*/,
                        cudaStream_t st);

}  // namespace bp
