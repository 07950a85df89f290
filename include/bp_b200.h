/*
 * bp_b200.h — C ABI of the Blackwell-native basic-pitch hot path (libbp_b200.so, sm_100a).
 *
 * The reference has no FFI: its boundary is a Python API in front of a third-party ML runtime
 * (SURVEY.md §8b).  Each entry point below names the reference interface it stands in for
 * (paths relative to the reference repository root).  All pointers are plain host or device
 * pointers; no torch / numpy types cross this boundary.  Unless stated otherwise arrays are
 * C-contiguous float32.
 *
 * Error convention: every function returning `int` returns BP_OK (0) or a negative BP_E_* code;
 * `bp_last_error()` returns a thread-local, human-readable message for the last failure.
 * There is no CPU fallback: without a CUDA device every call fails with BP_E_CUDA.
 *
 * Threading: one `bp_model_t` is bound to one CUDA device and must not be used from two host
 * threads at the same time (the reference's `Model` is likewise single-threaded,
 * basic_pitch/inference.py:71-182).  Different models (e.g. one per GPU) are independent.
 */
#ifndef BP_B200_H
#define BP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BP_OK 0
#define BP_E_INVALID (-1)  /* bad argument / malformed weight blob                     */
#define BP_E_CUDA (-2)     /* CUDA runtime error (message carries cudaGetErrorString)  */
#define BP_E_CAPACITY (-3) /* caller-provided output capacity too small                */
#define BP_E_NOMEM (-4)

/* Geometry of the model (reference: basic_pitch/constants.py:25-47, inference.py:185-191,302-305). */
#define BP_SAMPLE_RATE 22050
#define BP_WINDOW_SAMPLES 43844 /* AUDIO_N_SAMPLES                                     */
#define BP_WINDOW_FRAMES 172    /* ANNOT_N_FRAMES                                      */
#define BP_N_PITCHES 88         /* note / onset bins                                   */
#define BP_N_CONTOUR_BINS 264   /* contour bins                                        */
#define BP_OVERLAP_FRAMES 30    /* DEFAULT_OVERLAPPING_FRAMES                          */
#define BP_HOP_SAMPLES 36164    /* AUDIO_N_SAMPLES - 30*256                            */
#define BP_HOP_FRAMES 142       /* frames kept per window after unwrapping             */
#define BP_LEAD_ZEROS 3840      /* zeros prepended to every file (overlap_len / 2)     */

typedef struct bp_model bp_model_t;

/* Parameters of the note decode.
 * reference: basic_pitch/note_creation.py:360-371 (output_to_notes_polyphonic arguments) and
 * :52-63 (model_output_to_notes).  The pitch-column range replaces min_freq/max_freq: the host
 * converts Hz to columns exactly as constrain_frequency does (note_creation.py:333-336). */
typedef struct bp_decode_params {
  double onset_thresh;       /* default 0.5  */
  double frame_thresh;       /* default 0.3 ; must be >= 0 when melodia_trick (reference loops forever otherwise) */
  int32_t min_note_len;      /* frames, default 11 */
  int32_t energy_tol;        /* default 11 */
  int32_t infer_onsets;      /* default 1  */
  int32_t melodia_trick;     /* default 1  */
  int32_t include_pitch_bends; /* default 1 */
  int32_t min_pitch_idx;     /* columns < this are zeroed (default 0)   */
  int32_t max_pitch_idx;     /* columns >= this are zeroed (default 88) */
  int32_t reserved;
} bp_decode_params_t;

/* Note events of a batch of files, structure-of-arrays, caller-allocated (host memory).
 * Notes of file i occupy [note_off[i], note_off[i+1]) in generation order (onset-loop notes latest
 * first, then melodia notes — the order of the reference's list, note_creation.py:410-509).
 * Pitch bends of note j occupy bends[bend_off[j] .. bend_off[j+1]) (1/3-semitone units,
 * note_creation.py:215-217); bend_off has note capacity + 1 entries. */
typedef struct bp_notes {
  int32_t note_capacity;  /* in: capacity of the per-note arrays                     */
  int32_t bend_capacity;  /* in: capacity of `bends`                                 */
  int32_t* note_off;      /* out: [n_files + 1]                                      */
  int32_t* start_frame;   /* out: [note_capacity]                                    */
  int32_t* end_frame;     /* out: [note_capacity]                                    */
  int32_t* pitch_midi;    /* out: [note_capacity]  (column + 21)                     */
  float* amplitude;       /* out: [note_capacity]                                    */
  int32_t* bend_off;      /* out: [note_capacity + 1]                                */
  int32_t* bends;         /* out: [bend_capacity]; may be NULL if !include_pitch_bends */
} bp_notes_t;

/* ---- library ---------------------------------------------------------------------------- */
int bp_version(void);
const char* bp_last_error(void);
void bp_default_decode_params(bp_decode_params_t* p);

/* Number of model windows / output frames for a file of n_samples 22 050 Hz samples.
 * reference: basic_pitch/inference.py:194-219 (window_audio_file), :247-279 (unwrap_output). */
int64_t bp_num_windows(int64_t n_samples);
int64_t bp_num_frames(int64_t n_samples);

/* ---- model lifetime ----------------------------------------------------------------------
 * reference: Model.__init__ (basic_pitch/inference.py:78-154) loading
 * the files under basic_pitch/saved_models/icassp_2022.  `blob` is the BPW1 tensor blob
 * (basic_pitch_b200/weights.py); it is copied, the caller keeps ownership. */
int bp_model_create(const void* blob, size_t nbytes, int device, bp_model_t** out);
void bp_model_destroy(bp_model_t* m);
int bp_model_device(const bp_model_t* m);
/* Device pointer + size of the packed parameter block (for the one NCCL broadcast at init). */
int bp_model_param_block(bp_model_t* m, void** d_ptr, size_t* nbytes);
/* Re-derive internal (split / transposed) weight layouts after the block was overwritten. */
int bp_model_refresh(bp_model_t* m);
/* Kernel launches issued by this model since creation (for bench.py's gpu_launches). */
int64_t bp_model_launch_count(const bp_model_t* m);

/* ---- stage 1+2: audio windows -> posteriorgrams -------------------------------------------
 * reference: Model.predict (basic_pitch/inference.py:156-182), batched: audio [n][43844] ->
 * note [n][172][88], onset [n][172][88], contour [n][172][264].
 * _device: pointers are device memory on the model's device; work is enqueued on `stream`
 * (a cudaStream_t) and NOT synchronised.  _host: host pointers, copies inside, synchronous. */
int bp_forward_device(bp_model_t* m, const float* d_audio, int64_t n_windows, float* d_note, float* d_onset,
                      float* d_contour, void* stream);
int bp_forward_host(bp_model_t* m, const float* h_audio, int64_t n_windows, float* h_note, float* h_onset,
                    float* h_contour);

/* ---- run_inference for whole files ---------------------------------------------------------
 * reference: run_inference (basic_pitch/inference.py:282-330) = get_audio_input windowing
 * (:222-244) + Model.predict per window + unwrap_output (:247-279), for a batch of files.
 * audio: the files' samples back to back; sample_off[n_files+1] gives each file's range.
 * Outputs: unwrapped posteriorgrams of all files back to back; file i has
 * bp_num_frames(len_i) frames starting at frame_off[i] (frame_off[n_files+1] is written by the
 * call; host array in both variants).  Output arrays must hold sum_i bp_num_frames(len_i) frames. */
int bp_run_inference_device(bp_model_t* m, const float* d_audio, const int64_t* h_sample_off, int32_t n_files,
                            float* d_note, float* d_onset, float* d_contour, int64_t* h_frame_off, void* stream);
int bp_run_inference_host(bp_model_t* m, const float* h_audio, const int64_t* h_sample_off, int32_t n_files,
                          float* h_note, float* h_onset, float* h_contour, int64_t* h_frame_off);

/* ---- stage 3: posteriorgrams -> note events -------------------------------------------------
 * reference: model_output_to_notes without the MIDI object (basic_pitch/note_creation.py:52-111):
 * output_to_notes_polyphonic (:360-511) + get_pitch_bends (:182-219), for a batch of files whose
 * (unwrapped) posteriorgrams lie back to back, file i covering frames [frame_off[i], frame_off[i+1]).
 * The posteriorgrams are not modified (the reference zeroes out-of-range pitch columns in place,
 * note_creation.py:338-341; the Python wrapper reproduces that on its own arrays).
 * `notes` arrays are host memory in both variants; the call synchronises `stream` before returning. */
int bp_decode_device(bp_model_t* m, const float* d_note, const float* d_onset, const float* d_contour,
                     const int64_t* h_frame_off, int32_t n_files, const bp_decode_params_t* params, bp_notes_t* notes,
                     void* stream);
int bp_decode_host(bp_model_t* m, const float* h_note, const float* h_onset, const float* h_contour,
                   const int64_t* h_frame_off, int32_t n_files, const bp_decode_params_t* params, bp_notes_t* notes);

/* ---- the whole path: predict() for a batch of files -------------------------------------------
 * reference: predict (basic_pitch/inference.py:431-506) minus file I/O and the MIDI object:
 * run_inference + model_output_to_notes.  Posteriorgram outputs are optional (pass NULL to keep
 * them on the device only).  `h_frame_off` [n_files+1] is always written. */
int bp_transcribe_host(bp_model_t* m, const float* h_audio, const int64_t* h_sample_off, int32_t n_files,
                       const bp_decode_params_t* params, float* h_note, float* h_onset, float* h_contour,
                       int64_t* h_frame_off, bp_notes_t* notes);
int bp_transcribe_device(bp_model_t* m, const float* d_audio, const int64_t* h_sample_off, int32_t n_files,
                         const bp_decode_params_t* params, int64_t* h_frame_off, bp_notes_t* notes, void* stream);

/* ---- the two sub-steps of the decode the reference also exposes as functions --------------------------------------
 * bp_infer_onsets_host — reference: basic_pitch/note_creation.py:289-311 get_infered_onsets (n_diff = 2):
 *   out[t][f] (float64) = max(onsets, max(onsets) * frame_diff / max(frame_diff)) of one file, h_onset / h_note (T,88) f32.
 *   All-NaN when max(frame_diff) == 0, like the reference.
 * bp_pitch_bends_host — reference: note_creation.py:182-219 get_pitch_bends (n_bins_tolerance = 25): for every given
 *   note (start_frame, end_frame, pitch_midi) the per-frame pitch-bend estimates (1/3-semitone units) from the contour
 *   posteriorgram (T,264); h_bend_off[n_notes+1] receives the offsets into h_bends (BP_E_CAPACITY: "need N"). */
int bp_infer_onsets_host(bp_model_t* m, const float* h_onset, const float* h_note, int64_t n_frames, double* h_out);
int bp_pitch_bends_host(bp_model_t* m, const float* h_contour, int64_t n_frames, int32_t n_notes, const int32_t* h_start,
                        const int32_t* h_end, const int32_t* h_pitch_midi, int32_t* h_bend_off, int32_t* h_bends,
                        int64_t bend_capacity);

/* ---- introspection used by tests / profiling ----------------------------------------------------
 * Copies an internal activation of the most recent bp_forward_* call for window 0..n-1 to host.
 * which: 0 = CQT log-magnitude after normalisation+BN (n,172,309); 1 = contour conv1 output
 * (n,8,172,264); 2 = note conv1 output (n,32,172,88); 3 = onset conv1 output (n,32,172,88) — 2 and 3 exist only on
 * the FP32 path (bp_model_set_path(m, 0)) and 1 on paths 0 and 2: the tensor-core kernels reduce these activations
 * against the following convolution in their epilogues and never store them.
 * Only valid when the batch fitted in one internal chunk (n <= bp_model_chunk_windows). */
int bp_debug_activation(bp_model_t* m, int which, float* h_out, int64_t n_windows);
int64_t bp_model_chunk_windows(const bp_model_t* m);
/* Selects the arithmetic path: 0 = FP32 FFMA kernels everywhere; 1 = tensor-core kernels (tcgen05, split-bf16
 * operands, FP32 accumulate) for the constant-Q projection and the three wide convolutions, each with the following
 * single-output convolution reduced in its epilogue (default); 2 = as 1 but the contour convolution stores its
 * 8-channel activations (bp_debug_activation which = 1). */
int bp_model_set_path(bp_model_t* m, int path);

/* `bp_transcribe_host` for audio that is NOT packed: file i is audio[i][0 .. n_samples[i]) in ordinary (pageable) host
 * memory — what a caller holding one array per file has (reference: basic_pitch/inference.py:509-604, `predict_and_save`
 * loops `predict` over `audio_path_list`).  The library gathers the files sub-batch by sub-batch into pinned staging with
 * a few host threads, so that the gather and upload of sub-batch k+1 overlap the kernels of k, and streams the
 * posteriorgrams of each finished sub-batch back (h_note / h_onset / h_contour, row-major [total_frames][88 | 88 | 264],
 * any may be NULL) while later sub-batches compute; buffers from bp_host_alloc make that copy asynchronous.
 * Everything else (h_frame_off, notes, errors) as bp_transcribe_host. */
int bp_transcribe_files_host(bp_model_t* m, const float* const* audio, const int64_t* n_samples, int32_t n_files,
                             const bp_decode_params_t* params, float* h_note, float* h_onset, float* h_contour,
                             int64_t* h_frame_off, bp_notes_t* notes);
/* Page-locked host memory for the outputs of the host entry points (cudaHostAlloc / cudaFreeHost); NULL on failure. */
void* bp_host_alloc(size_t bytes);
void bp_host_free(void* p);
/* After a BP_E_CAPACITY failure of a decode / transcribe call on this thread: the capacities that call needed
 * (0 = that one was sufficient), so that the caller can retry without parsing the error text. */
void bp_last_required(int64_t* note_capacity, int64_t* bend_capacity);

/* Audio ingest — `librosa.load(path, sr=22050, mono=True)` minus the container decode (reference:
 * basic_pitch/inference.py:239): interleaved PCM frames of `channels` channels at `sample_rate` Hz -> mono float32 at
 * 22 050 Hz.  sample_format: 0 float32, 1 int16 (/ 2^15), 2 int32 (/ 2^31; 24-bit WAV left-justified), 3 uint8
 * ((x - 128) / 128).  Channels are averaged in float32; other rates go through a Kaiser-windowed polyphase FIR (pass
 * band 0.913 x Nyquist of the lower rate, 125 dB stop band) applied like scipy.signal.resample_poly.  The output has
 * bp_resampled_length(n_frames, sample_rate) = ceil(n_frames * 22050 / sample_rate) samples.
 * _device: PCM and output in device memory, asynchronous on `stream` — the output can feed bp_transcribe_device directly;
 * _host: both in host memory (the PCM crosses PCIe as stored: 2 bytes per sample for 16-bit audio). */
int64_t bp_resampled_length(int64_t n_frames, int32_t sample_rate);
int bp_load_pcm_device(bp_model_t* m, const void* d_pcm, int32_t sample_format, int64_t n_frames, int32_t channels,
                       int32_t sample_rate, float* d_audio, void* stream);
int bp_load_pcm_host(bp_model_t* m, const void* h_pcm, int32_t sample_format, int64_t n_frames, int32_t channels,
                     int32_t sample_rate, float* h_audio);

/* Batched writers (host threads, no GPU): one Standard MIDI File and / or one note-event CSV per file of a batch, straight
 * from the concatenated note arrays (times already in seconds).  Replaces, for batches, note_events_to_midi (reference:
 * basic_pitch/note_creation.py:222-271, incl. drop_overlapping_pitch_bends :274-286 when multiple_pitch_bends = 0) +
 * PrettyMIDI.write (inference.py:574-584) and save_note_events (inference.py:409-428).  File i owns notes
 * [note_off[i], note_off[i+1]); note j owns bends [bend_off[j], bend_off[j+1]) (bend_off may be NULL: no bends).
 * midi_paths / csv_paths: arrays of n_files paths (the array or single entries may be NULL = skip).  n_threads <= 0: auto. */
int bp_write_note_files(int32_t n_files, const char* const* midi_paths, const char* const* csv_paths,
                        const int32_t* note_off, const double* start_s, const double* end_s, const int32_t* pitch_midi,
                        const float* amplitude, const int32_t* bend_off, const int32_t* bends,
                        int32_t multiple_pitch_bends, double midi_tempo, int32_t n_threads);

/* Host-only: the low-pass of the ingest resampler for the reduced ratio up / down (unit DC gain, before the gain `up`);
 * returns the number of taps, copies them if capacity allows.  Tests pin it against scipy.signal.firwin. */
int64_t bp_debug_resample_filter(int32_t up, int32_t down, double* taps, int64_t capacity);

/* Host-only (no GPU needed): builds the tensor-core plan (split-bf16 Toeplitz weight tiles and the per-group MMA
 * programs, csrc/tc_conv.cu) of the contour conv (which = 0, w = [8][8][3][39]), the onset conv (which = 1,
 * w = [32][8][5][5]) or the note conv (which = 2, w = [32][1][7][7]) so that tests can emulate the program on the CPU.  sizes[4] = {n_tiles, n_steps, n_uses,
 * n_groups}; pass NULL arrays to query sizes first.  tiles: n_tiles x 4096 bf16 ([plane hi/lo][k-chunk 2][n 128][8]);
 * slot_words: [2][n_steps]; group_step_off has n_groups + 1 entries, group_ft n_groups x 2. */
int bp_debug_tc_plan(int which, const float* w, int32_t* sizes, uint16_t* tiles, int32_t* tile_seq, uint32_t* slot_words,
                     int32_t* group_step_off, int32_t* group_ft);

/* Host-only: the bf16 hi/lo weight tiles of the fused SECOND convolution of a tensor-core layer (csrc/tc_conv.cu, TcB2):
 * which = 0 contour conv2 (w2 = [1][8][5][5], reference models.py:254-262), 1 onset conv2 ([1][33][3][3], models.py:305-313),
 * 2 note conv2 ([1][32][7][3], models.py:282-290).  sizes[5] = {n_tiles, N, time taps, accumulator columns, columns per output offset}; tiles
 * (may be NULL to query sizes): n_tiles x [plane hi/lo][k-chunk 2][n N][8] bf16. */
int bp_debug_tc_b2(int which, const float* w2, int32_t* sizes, uint16_t* tiles);

/* Per-kernel device timing for the roofline line of bench.py: records CUDA events on the launching
 * stream around every launch of one kernel family (0 = contour conv 3x39, 1 = onset conv 5x5,
 * 2 = CQT projection + log-normalise, 3 = decimation chain, 4 = the remaining small convs,
 * 5 = decode (prep / candidates / sequential loops), 6 = amplitudes + pitch bends;
 * -1 = off, the default) and resets the accumulators.  bp_model_profile_read synchronises the device
 * and returns the summed interval time, the number of intervals and the windows processed. */
int bp_model_profile(bp_model_t* m, int which);
int bp_model_profile_read(bp_model_t* m, double* total_ms, int64_t* n_intervals, int64_t* n_windows);

#ifdef __cplusplus
}
#endif
#endif /* BP_B200_H */
