// C ABI of the hot path (include/bp_b200.h): model lifetime, workspace, chunked launch sequences.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "bp_b200.h"
#include "kernels.cuh"

using namespace bp;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

}  // namespace
namespace bp {
int writer_fail(int code, const std::string& msg) { return fail(code, msg); }  // writers.cu reports through bp_last_error
}
namespace {

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      return fail(BP_E_CUDA, std::string(#call) + " failed: " + cudaGetErrorString(e_) + " (" __FILE__ ":" + \
                                 std::to_string(__LINE__) + ")");                                        \
  } while (0)

#define CKL()                                                                                            \
  do {                                                                                                   \
    cudaError_t e_ = cudaGetLastError();                                                                 \
    if (e_ != cudaSuccess)                                                                               \
      return fail(BP_E_CUDA, std::string("kernel launch failed: ") + cudaGetErrorString(e_) + " (" __FILE__ ":" + \
                                 std::to_string(__LINE__) + ")");                                        \
  } while (0)

// A device buffer that only ever grows.
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

// transposed / interleaved weight layouts derived from the parameter block
struct DerivedLayout {
  static constexpr int cqt_wt = 0;                           // [256][72] columns: re0,im0,re1,im1,...
  static constexpr int contour1_wT = cqt_wt + 256 * 72;      // [936][8]
  static constexpr int contour2_wT = contour1_wT + 936 * 8;  // [200][1]
  static constexpr int note1_wT = contour2_wT + 200;         // [49][32]
  static constexpr int note2_wT = note1_wT + 49 * 32;        // [672][1]
  static constexpr int onset1_wT = note2_wT + 672;           // [200][32]
  static constexpr int onset2_wT = onset1_wT + 200 * 32;     // [297][1] (+3)
  static constexpr int total = onset2_wT + 300;
};

__global__ void derive_kernel(const float* __restrict__ P, float* __restrict__ D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 256 * 72) {
    int k = i / 72, n = i % 72;
    int bin = n >> 1;
    D[DerivedLayout::cqt_wt + i] = (n & 1) ? P[ParamLayout::cqt_imag + bin * 256 + k] : P[ParamLayout::cqt_real + bin * 256 + k];
  }
  auto tr = [&](int src, int dst, int cout, int kk) {  // [cout][kk] -> [kk][cout]
    if (i < cout * kk) {
      int k = i / cout, c = i % cout;
      D[dst + i] = P[src + c * kk + k];
    }
  };
  tr(ParamLayout::contour1_w, DerivedLayout::contour1_wT, 8, 936);
  tr(ParamLayout::contour2_w, DerivedLayout::contour2_wT, 1, 200);
  tr(ParamLayout::note1_w, DerivedLayout::note1_wT, 32, 49);
  tr(ParamLayout::note2_w, DerivedLayout::note2_wT, 1, 672);
  tr(ParamLayout::onset1_w, DerivedLayout::onset1_wT, 32, 200);
  tr(ParamLayout::onset2_w, DerivedLayout::onset2_wT, 1, 297);
}

__global__ void desc_upload_kernel(const int4* __restrict__ a_src, int4* __restrict__ a_dst, const int4* __restrict__ b_src,
                                   int4* __restrict__ b_dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    a_dst[i] = a_src[i];
    b_dst[i] = b_src[i];
  }
}

__global__ void compact_notes_kernel(const long long* __restrict__ frame_off, const long long* __restrict__ slot_off,
                                     const int* __restrict__ note_off, const int* __restrict__ s_start,
                                     const int* __restrict__ s_end, const int* __restrict__ s_pitch,
                                     int* __restrict__ start, int* __restrict__ end, int* __restrict__ pitch,
                                     long long* __restrict__ note_base) {
  const int file = blockIdx.x;
  const int n = note_off[file + 1] - note_off[file];
  const long long s0 = slot_off[file];
  const int d0 = note_off[file];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    start[d0 + i] = s_start[s0 + i];
    end[d0 + i] = s_end[s0 + i];
    pitch[d0 + i] = s_pitch[s0 + i];
    note_base[d0 + i] = frame_off[file];
  }
}

}  // namespace

struct bp_model {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // host->device audio copies of bp_transcribe_host run ahead of the compute stream
  std::vector<cudaEvent_t> copy_ev;
  float* d_params = nullptr;
  float* d_derived = nullptr;
  double* d_gauss = nullptr;
  CnnWeights cw{};
  int chunk = 206;  // windows per launch sequence: the M-tiles of every tensor-core layer fill whole waves (bp_model_create)
  int path = 1;  // 0 = FP32 FFMA everywhere, 1 = tcgen05 with fused epilogues, 2 = tcgen05 keeping the contour activations
  int n_sms = 148;
  struct TcLayer {
    TcConvPlan plan;
    TcConvDev dev{};
    DevBuf<uint16_t> tiles, b2;
  } tc_contour, tc_onset, tc_note;
  DevBuf<__nv_bfloat16> yhl, chl;
  std::vector<float> h_params;  // host copy of the parameter block (weight-dependent __constant__ data is re-uploaded
                                // from it whenever another model used the device's constant bank in between)
  DevBuf<uint16_t> cqt_wtc;  // three-way bf16 split of the CQT kernel matrix (tensor-core path)
  size_t chl_zeroed = 0;  // elements of chl known to hold zeros in every row/bin the kernels never write
  int64_t launches = 0;
  // forward workspace (chunk windows)
  DevBuf<float> chain, y, c1, n1, o1;
  DevBuf<float> raw_note, raw_onset, raw_contour;  // row-major raw windows [nb][172][*] (FP32 path inside run_inference)
  DevBuf<float> i_note, i_onset, i_contour;        // internal raw: pm [88][chunk*172], cm [33][chunk*172][8] (tensor-core paths)
  DevBuf<float> u_note, u_onset, u_contour;        // internal unwrapped posteriorgrams of a call: pm [88][F], cm [33][F][8]
  bool y_is_log = false;  // tensor-core paths: `y` still holds the raw log-magnitudes (bp_debug_activation normalises)
  DevBuf<unsigned int> minmax;
  DevBuf<float> edge;  // partial sums where two frequency-tile ranges of a fused conv meet (tc_conv.cu)
  DevBuf<WinDesc> wdesc;
  DevBuf<UnwrapDesc> udesc;
  // staging for the host entry points
  DevBuf<float> st_audio, st_note, st_onset, st_contour;
  DevBuf<unsigned char> st_pcm;  // bp_load_pcm_host
  // bp_transcribe_files_host: pinned gather buffers (one sub-batch of audio each), the device->host stream of the
  // posteriorgrams and its events
  float* gather[3] = {nullptr, nullptr, nullptr};
  size_t gather_cap = 0;  // floats per buffer
  cudaStream_t d2h_stream = nullptr;
  std::vector<cudaEvent_t> conv_ev;
  // decode workspace
  DevBuf<long long> d_frame_off, d_slot_off, d_note_base;
  DevBuf<float> energy, d_amp;
  DevBuf<double> d_onset64;
  DevBuf<unsigned int> candbits, max_onset;
  DevBuf<float> blk_max;
  DevBuf<int> blk_arg;
  DevBuf<unsigned long long> max_fd;
  DevBuf<int> note_count, slot_start, slot_end, slot_pitch, overflow, d_note_off, d_start, d_end, d_pitch, d_bend_off,
      d_bends;
  int64_t last_forward_n = 0;
  int last_path = 0;
  // optional per-kernel timing (bench.py roofline): CUDA events around one kernel family
  WinDesc* h_wd = nullptr;        // pinned staging of the window / unwrap descriptors (bp_run_inference_device)
  UnwrapDesc* h_ud = nullptr;
  size_t h_desc_cap = 0;
  cudaEvent_t desc_ev = nullptr;
  bool desc_pending = false;
  int profile_which = -1;  // -1 off; 0 contour1, 1 onset1, 2 cqt, 3 decimate chain, 4 small convs, 5 decode, 6 note finish
  std::vector<cudaEvent_t> prof_ev;
  size_t prof_used = 0;
  int64_t prof_windows = 0;
};

namespace {

// Weight-dependent data in __constant__ memory (FIR taps, conv1 biases, fused conv2 weights) is shared by all models
// of a process on one device; remember whose values are resident.
// The bank is guarded per device: a launch sequence that depends on it (forward_chunk) holds the device's mutex while
// it enqueues, and a change of owner first drains the device so that kernels of the previous owner that are still in
// flight never see the new values.  Models with different weights may therefore be used from several host threads;
// they serialise at chunk granularity.
thread_local int64_t g_need_notes = 0, g_need_bends = 0;  // bp_last_required
const bp_model* g_const_owner[64] = {};
std::mutex g_const_mu[64];

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

int parse_blob(const void* blob, size_t nbytes, std::vector<float>& params) {
  const unsigned char* b = static_cast<const unsigned char*>(blob);
  if (nbytes < 8 || std::memcmp(b, "BPW1", 4) != 0) return fail(BP_E_INVALID, "weight blob: bad magic (expected BPW1)");
  uint32_t count;
  std::memcpy(&count, b + 4, 4);
  size_t pos = 8;
  params.assign(ParamLayout::total, 0.f);
  struct Slot {
    const char* name;
    int off;
    int n;
  };
  const Slot slots[] = {
      {"cqt_real", ParamLayout::cqt_real, 36 * 256},  {"cqt_imag", ParamLayout::cqt_imag, 36 * 256},
      {"lowpass", ParamLayout::lowpass, 256},         {"cqt_scale", ParamLayout::cqt_scale, 309},
      {"bn_scale", ParamLayout::bn, 1},               {"bn_bias", ParamLayout::bn + 1, 1},
      {"contour1_w", ParamLayout::contour1_w, 7488},  {"contour1_b", ParamLayout::contour1_b, 8},
      {"contour2_w", ParamLayout::contour2_w, 200},   {"contour2_b", ParamLayout::contour2_b, 1},
      {"note1_w", ParamLayout::note1_w, 1568},        {"note1_b", ParamLayout::note1_b, 32},
      {"note2_w", ParamLayout::note2_w, 672},         {"note2_b", ParamLayout::note2_b, 1},
      {"onset1_w", ParamLayout::onset1_w, 6400},      {"onset1_b", ParamLayout::onset1_b, 32},
      {"onset2_w", ParamLayout::onset2_w, 297},       {"onset2_b", ParamLayout::onset2_b, 1},
  };
  unsigned found = 0;
  for (uint32_t t = 0; t < count; ++t) {
    if (pos + 4 > nbytes) return fail(BP_E_INVALID, "weight blob: truncated");
    uint32_t nl;
    std::memcpy(&nl, b + pos, 4);
    pos += 4;
    if (nl > 64 || pos + nl > nbytes) return fail(BP_E_INVALID, "weight blob: bad tensor name");
    std::string name(reinterpret_cast<const char*>(b + pos), nl);
    pos += nl + ((4 - nl % 4) % 4);
    if (pos + 4 > nbytes) return fail(BP_E_INVALID, "weight blob: truncated");
    uint32_t nd;
    std::memcpy(&nd, b + pos, 4);
    pos += 4;
    if (nd > 8 || pos + 4 * nd > nbytes) return fail(BP_E_INVALID, "weight blob: bad rank");
    size_t n = 1;
    for (uint32_t d = 0; d < nd; ++d) {
      uint32_t v;
      std::memcpy(&v, b + pos, 4);
      pos += 4;
      n *= v;
    }
    if (pos + 4 * n > nbytes) return fail(BP_E_INVALID, "weight blob: truncated tensor " + name);
    for (size_t s = 0; s < sizeof(slots) / sizeof(slots[0]); ++s) {
      if (name == slots[s].name) {
        if ((int)n != slots[s].n) return fail(BP_E_INVALID, "weight blob: tensor " + name + " has wrong size");
        std::memcpy(params.data() + slots[s].off, b + pos, 4 * n);
        found |= 1u << s;
      }
    }
    pos += 4 * n;
  }
  if (found != (1u << (sizeof(slots) / sizeof(slots[0]))) - 1) return fail(BP_E_INVALID, "weight blob: missing tensors");
  return BP_OK;
}

int upload_constants(bp_model* m, cudaStream_t st) {
  const float* hp = m->h_params.data();
  if (m->device >= 0 && m->device < 64 && g_const_owner[m->device] && g_const_owner[m->device] != m)
    CK(cudaDeviceSynchronize());  // kernels of the previous owner may still be reading the bank
  upload_lowpass(hp + ParamLayout::lowpass, st);
  tc_upload_epilogue(hp + ParamLayout::contour1_b, hp + ParamLayout::onset1_b, hp + ParamLayout::note1_b,
                     hp + ParamLayout::onset2_w, hp + ParamLayout::contour2_b, hp + ParamLayout::onset2_b, hp + ParamLayout::note2_b, st);
  CKL();
  if (m->device >= 0 && m->device < 64) g_const_owner[m->device] = m;
  return BP_OK;
}

int derive(bp_model* m, cudaStream_t st) {
  derive_kernel<<<(256 * 72 + 255) / 256, 256, 0, st>>>(m->d_params, m->d_derived);
  CKL();
  m->launches += 1;
  // tensor-core plans: split-bf16 Toeplitz weight tiles + MMA programs (host-built from the parameter block)
  m->h_params.resize(ParamLayout::total);
  std::vector<float>& hp = m->h_params;
  CK(cudaMemcpyAsync(hp.data(), m->d_params, sizeof(float) * ParamLayout::total, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const TcConvSpec specs[3] = {tc_contour_spec(), tc_onset_spec(), tc_note_spec()};
  bp_model::TcLayer* layers[3] = {&m->tc_contour, &m->tc_onset, &m->tc_note};
  const float* wsrc[3] = {hp.data() + ParamLayout::contour1_w, hp.data() + ParamLayout::onset1_w,
                          hp.data() + ParamLayout::note1_w};
  const float* w2src[3] = {hp.data() + ParamLayout::contour2_w, hp.data() + ParamLayout::onset2_w,
                           hp.data() + ParamLayout::note2_w};
  for (int l = 0; l < 3; ++l) {
    bp_model::TcLayer& L = *layers[l];
    L.plan.build(specs[l], wsrc[l]);
    const TcConvPlan& pl = L.plan;
    if (tc_upload_program(l, pl, st) != 0)
      return fail(BP_E_INVALID, "tensor-core program does not fit its constant-memory area");
    CK(L.tiles.reserve(pl.tiles.size()));
    CK(cudaMemcpyAsync(L.tiles.p, pl.tiles.data(), pl.tiles.size() * 2, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    std::vector<uint16_t> b2;
    tc_build_b2(l, w2src[l], b2);
    CK(L.b2.reserve(b2.size()));
    CK(cudaMemcpyAsync(L.b2.p, b2.data(), b2.size() * 2, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    L.dev = TcConvDev{pl.spec, L.tiles.p, L.b2.p, pl.n_groups, l};
  }
  {
    std::vector<uint16_t> wtc;
    build_cqt_tc_weights(hp.data() + ParamLayout::cqt_real, hp.data() + ParamLayout::cqt_imag, wtc);
    CK(m->cqt_wtc.reserve(wtc.size()));
    CK(cudaMemcpyAsync(m->cqt_wtc.p, wtc.data(), wtc.size() * 2, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
  }
  std::unique_lock<std::mutex> const_lock;
  if (m->device >= 0 && m->device < 64) const_lock = std::unique_lock<std::mutex>(g_const_mu[m->device]);
  return upload_constants(m, st);
}

int ensure_forward_ws(bp_model* m, int nb) {
  CK(m->chain.reserve((size_t)nb * kChainStride));
  CK(m->y.reserve((size_t)nb * kFrames * kCqtBins));
  // the default path never materialises the 8- / 32-channel activations
  if (m->path != 1) CK(m->c1.reserve((size_t)nb * 8 * kFrames * kContourBins));
  if (m->path == 0) {
    CK(m->n1.reserve((size_t)nb * 32 * kFrames * kPitches));
    CK(m->o1.reserve((size_t)nb * 32 * kFrames * kPitches));
  }
  CK(m->minmax.reserve((size_t)nb * 2));
  if (m->path >= 1) {
    const size_t fr = (size_t)m->chunk * kFrames;
    CK(m->i_note.reserve(kPitches * fr));
    CK(m->i_onset.reserve(kPitches * fr));
    if (m->path == 1) CK(m->i_contour.reserve((size_t)kContourBins * fr));
  }
  CK(m->edge.reserve(std::max({tc_edge_floats(tc_contour_spec(), nb), tc_edge_floats(tc_onset_spec(), nb),
                               tc_edge_floats(tc_note_spec(), nb)})));
  // split layouts use the row stride of a full chunk whatever the batch size (see launch_conv_tc)
  CK(m->yhl.reserve((size_t)2 * 40 * 8 * tc_rows_total(m->chunk, tc_contour_spec().rows_per_window)));
  {
    const TcConvSpec ns = tc_note_spec();
    const size_t need = (size_t)2 * ns.chunks8 * 8 * tc_rows_total(m->chunk, ns.rows_per_window);
    const __nv_bfloat16* before = m->chl.p;
    CK(m->chl.reserve(need));
    if (m->chl.p != before || m->chl_zeroed < m->chl.cap) {  // separator rows / pad bins are never written: zero once
      CK(cudaMemset(m->chl.p, 0, m->chl.cap * sizeof(__nv_bfloat16)));
      m->chl_zeroed = m->chl.cap;
    }
  }
  return BP_OK;
}

struct ProfScope {
  bp_model* m;
  cudaStream_t st;
  bool on;
  ProfScope(bp_model* m_, int which, cudaStream_t st_) : m(m_), st(st_), on(m_->profile_which == which) {
    if (on) rec();
  }
  ~ProfScope() {
    if (on) rec();
  }
  void rec() {
    if (m->prof_used == m->prof_ev.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      m->prof_ev.push_back(e);
    }
    cudaEventRecord(m->prof_ev[m->prof_used++], st);
  }
};

// Internal (frame-fastest) posteriorgrams of a call, see TcOut in kernels.cuh.
struct PostI {
  float *note, *onset, *contour;  // pm [88][stride], pm [88][stride], cm [33][stride][8]
  long long stride;
};

// HCQT + CNN for `nb` windows (nb <= chunk).
// Without `ud` (bp_forward_*): raw row-major outputs note / onset / contour [nb][172][*].
// With `ud` (bp_run_inference_*): the centre frames of every window go to their unwrapped position in the internal
// posteriorgrams `u` — straight from the fused epilogues on the tensor-core path, through a layout conversion of the
// row-major raw windows on the FP32 path; note / onset / contour are then optional row-major scratch.
int forward_chunk(bp_model* m, const float* audio, const WinDesc* desc, int nb, float* note, float* onset,
                  float* contour, cudaStream_t st, const UnwrapDesc* ud = nullptr, const PostI* u = nullptr) {
  float* chain = m->chain.p;
  if (m->profile_which >= 0) m->prof_windows += nb;
  std::unique_lock<std::mutex> const_lock;
  if (m->device >= 0 && m->device < 64) const_lock = std::unique_lock<std::mutex>(g_const_mu[m->device]);
  if (m->device >= 0 && m->device < 64 && g_const_owner[m->device] != m) {
    int rc = upload_constants(m, st);
    if (rc) return rc;
  }
  {
    ProfScope ps(m, 3, st);
    for (int s = 0; s < 8; ++s) launch_decimate(audio, desc, chain, s, nb, st);
  }
  const TcConvSpec cs = tc_contour_spec(), ns = tc_note_spec();
  const int ystride = tc_rows_total(m->chunk, cs.rows_per_window), cstride = tc_rows_total(m->chunk, ns.rows_per_window);
  const long long fr = (long long)m->chunk * kFrames;  // frame stride of the internal raw buffers
  const long long nfr = (long long)nb * kFrames;
  {
    ProfScope ps(m, 2, st);
    if (m->path >= 1) {
      launch_cqt_tc(audio, desc, chain, m->cqt_wtc.p, m->d_params + ParamLayout::cqt_scale, m->y.p, m->minmax.p, nb,
                    m->n_sms, st);
      // NormalizedLog + BatchNorm straight into the bf16 hi/lo split the convs read
      launch_lognorm_split(m->y.p, m->minmax.p, m->d_params + ParamLayout::bn, m->yhl.p, cs, nb, ystride, st);
      m->y_is_log = true;
    } else {
      launch_cqt(audio, desc, chain, m->d_derived + DerivedLayout::cqt_wt, m->d_params + ParamLayout::cqt_scale, m->y.p,
                 m->minmax.p, nb, st);
      launch_lognorm(m->y.p, m->minmax.p, m->d_params + ParamLayout::bn, nb, st);
      m->y_is_log = false;
    }
  }
  int extra = 0;  // launches beyond the fixed sequence
  if (m->path >= 1) {
    {
      ProfScope ps(m, 0, st);
      TcOut o;
      o.edge = m->edge.p;
      if (m->path == 1) {  // fused contour conv2: finished posteriorgram (internal layouts) + the note conv's operand
        o.raw = ud ? nullptr : m->i_contour.p;
        o.raw_rows = fr;
        o.ud = ud;
        o.unwrapped = ud ? u->contour : nullptr;
        o.frame_stride = ud ? u->stride : 0;
        o.chl = m->chl.p;
        o.chl_rows = cstride;
        o.chl_chunks = ns.chunks8;
        o.chl_rpw = ns.rows_per_window;
        o.chl_lead = ns.lead_rows;
      } else {
        o.act = m->c1.p;  // channels-last activations
      }
      launch_conv_tc(m->yhl.p, m->tc_contour.dev, o, nb, ystride, m->n_sms, st, /*fuse_next=*/m->path == 1);
    }
    {
      ProfScope ps(m, 4, st);
      if (m->path == 2) {
        launch_contour2_tc(m->c1.p, m->cw, contour, m->chl.p, cstride, nb, st);
        if (ud) launch_rows_to_cm(contour, 0, u->contour, u->stride, 0, st, ud, nb), ++extra;
      }
      TcOut o;
      o.raw = m->i_note.p;  // the onset conv2 reads the raw note rows of the whole window
      o.raw_rows = fr;
      o.ud = ud;
      o.unwrapped = ud ? u->note : nullptr;
      o.frame_stride = ud ? u->stride : 0;
      o.edge = m->edge.p;
      launch_conv_tc(m->chl.p, m->tc_note.dev, o, nb, cstride, m->n_sms, st);
    }
    {
      ProfScope ps(m, 1, st);
      TcOut o;
      o.raw = ud ? nullptr : m->i_onset.p;
      o.raw_rows = fr;
      o.ud = ud;
      o.unwrapped = ud ? u->onset : nullptr;
      o.frame_stride = ud ? u->stride : 0;
      o.note_raw = m->i_note.p;
      o.edge = m->edge.p;
      launch_conv_tc(m->yhl.p, m->tc_onset.dev, o, nb, ystride, m->n_sms, st);
    }
    if (!ud) {  // raw windows for the caller: internal -> row-major
      launch_pm_to_rows(m->i_note.p, fr, 0, nfr, kPitches, note, st);
      launch_pm_to_rows(m->i_onset.p, fr, 0, nfr, kPitches, onset, st);
      if (m->path == 1) launch_cm_to_rows(m->i_contour.p, fr, 0, nfr, contour, st), ++extra;
      extra += 2;
    }
  } else {
    {
      ProfScope ps(m, 0, st);
      launch_contour1(m->y.p, m->cw, m->c1.p, nb, st);
    }
    {
      ProfScope ps(m, 4, st);
      launch_contour2(m->c1.p, m->cw, contour, nb, st);
      launch_note1(contour, m->cw, m->n1.p, nb, st);
      launch_note2(m->n1.p, m->cw, note, nb, st);
    }
    {
      ProfScope ps(m, 1, st);
      launch_onset1(m->y.p, m->cw, m->o1.p, nb, st);
    }
    {
      ProfScope ps(m, 4, st);
      launch_onset2(note, m->o1.p, m->cw, onset, nb, st);
    }
    if (ud) {
      launch_rows_to_pm(note, 0, kPitches, u->note, u->stride, 0, st, ud, nb);
      launch_rows_to_pm(onset, 0, kPitches, u->onset, u->stride, 0, st, ud, nb);
      launch_rows_to_cm(contour, 0, u->contour, u->stride, 0, st, ud, nb);
      extra += 3;
    }
  }
  CKL();
  // decimation (4 + tail), min/max init + CQT, log-normalise (+ split); tensor-core convs: 3 x (conv + edge fix)
  // (path 2: contour conv2 instead of one edge fix); FP32 path: 6 convs
  m->launches += 5 + 2 + 1 + 6 + extra;
  m->last_path = m->path;
  return BP_OK;
}

int validate_params(const bp_decode_params_t* p) {
  if (!p) return fail(BP_E_INVALID, "decode params: null");
  if (!(p->frame_thresh == p->frame_thresh) || !(p->onset_thresh == p->onset_thresh))
    return fail(BP_E_INVALID, "decode params: NaN threshold");
  if (p->melodia_trick && p->frame_thresh < 0)
    return fail(BP_E_INVALID, "decode params: frame_thresh < 0 with melodia_trick never terminates (the reference loops forever)");
  if (p->energy_tol < 1) return fail(BP_E_INVALID, "decode params: energy_tol must be >= 1");
  if (p->min_note_len < 0) return fail(BP_E_INVALID, "decode params: min_note_len must be >= 0");
  return BP_OK;
}

}  // namespace

extern "C" {

int bp_version(void) { return 100; }
const char* bp_last_error(void) { return g_err.c_str(); }

void bp_default_decode_params(bp_decode_params_t* p) {
  if (!p) return;
  p->onset_thresh = 0.5;
  p->frame_thresh = 0.3;
  p->min_note_len = 11;
  p->energy_tol = 11;
  p->infer_onsets = 1;
  p->melodia_trick = 1;
  p->include_pitch_bends = 1;
  p->min_pitch_idx = 0;
  p->max_pitch_idx = BP_N_PITCHES;
  p->reserved = 0;
}

int64_t bp_num_windows(int64_t n_samples) {
  if (n_samples < 0) return 0;
  return (n_samples + kLeadZeros + kHopSamples - 1) / kHopSamples;
}
int64_t bp_num_frames(int64_t n_samples) {
  if (n_samples <= 0) return 0;
  return (int64_t)((double)n_samples / (double)kHopSamples * (double)kHopFrames);
}

static int model_init(bp_model* m, const std::vector<float>& params, const cudaDeviceProp& prop);

int bp_model_create(const void* blob, size_t nbytes, int device, bp_model_t** out) {
  if (!blob || !out) return fail(BP_E_INVALID, "bp_model_create: null argument");
  *out = nullptr;
  std::vector<float> params;
  int rc = parse_blob(blob, nbytes, params);
  if (rc) return rc;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(BP_E_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                               "); this library has no CPU path");
  if (device < 0 || device >= ndev) return fail(BP_E_INVALID, "bp_model_create: bad device index");
  DeviceGuard g(device);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(BP_E_CUDA, std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major) +
                               std::to_string(prop.minor) + "; this library is built for sm_100a only");
  bp_model* m = new bp_model();
  m->device = device;
  rc = model_init(m, params, prop);
  if (rc) {  // every failure path releases the streams and device buffers created so far
    const std::string msg = g_err;
    bp_model_destroy(m);
    g_err = msg;
    return rc;
  }
  *out = m;
  return BP_OK;
}

static int model_init(bp_model* m, const std::vector<float>& params, const cudaDeviceProp& prop) {
  int rc;
  CK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
  CK(cudaMalloc(&m->d_params, sizeof(float) * ParamLayout::total));
  CK(cudaMalloc(&m->d_derived, sizeof(float) * DerivedLayout::total));
  CK(cudaMalloc(&m->d_gauss, sizeof(double) * 51));
  // stream-ordered with derive_kernel below (m->stream is non-blocking: it does not wait for the legacy stream)
  CK(cudaMemcpyAsync(m->d_params, params.data(), sizeof(float) * ParamLayout::total, cudaMemcpyHostToDevice, m->stream));
  double gauss[51];
  for (int i = 0; i < 51; ++i) {  // scipy.signal.windows.gaussian(51, std=5): exp(-n^2 / (2*std^2))
    double n = (double)i - 25.0;
    gauss[i] = std::exp(-(n * n) / 50.0);
  }
  CK(cudaMemcpyAsync(m->d_gauss, gauss, sizeof(gauss), cudaMemcpyHostToDevice, m->stream));
  CK(cudaStreamSynchronize(m->stream));  // `gauss` is a stack array
  const float* P = m->d_params;
  const float* D = m->d_derived;
  m->cw = CnnWeights{D + DerivedLayout::contour1_wT, P + ParamLayout::contour1_b, D + DerivedLayout::contour2_wT,
                     P + ParamLayout::contour2_b,    D + DerivedLayout::note1_wT, P + ParamLayout::note1_b,
                     D + DerivedLayout::note2_wT,    P + ParamLayout::note2_b,    D + DerivedLayout::onset1_wT,
                     P + ParamLayout::onset1_b,      D + DerivedLayout::onset2_wT, P + ParamLayout::onset2_b};
  hcqt_setup();
  cnn_setup();
  if (tc_setup() != 0) {
    cudaGetLastError();
    return fail(BP_E_CUDA, "tensor-core conv kernels: shared-memory opt-in failed");
  }
  cqt_tc_setup();
  m->n_sms = prop.multiProcessorCount;
  // largest chunk whose M-tiles make at most two per SM in every layer.  M-tiles advance by 128 - (KH2 - 1) rows (they
  // overlap by the time taps of the fused conv2): 122 rows of 175 per window in the note layer, 124 / 126 of 174 in the
  // contour / onset layers -> 206 windows on 148 SMs
  {
    const TcConvSpec ns = tc_note_spec();
    m->chunk = std::max(1, 2 * m->n_sms * (128 - (ns.KH2 - 1)) / ns.rows_per_window);
  }
  rc = derive(m, m->stream);
  if (rc) return rc;
  CK(cudaStreamSynchronize(m->stream));
  return BP_OK;
}

void bp_model_destroy(bp_model_t* m) {
  if (!m) return;
  DeviceGuard g(m->device);
  cudaDeviceSynchronize();
  if (m->device >= 0 && m->device < 64) {
    std::lock_guard<std::mutex> lk(g_const_mu[m->device]);
    if (g_const_owner[m->device] == m) g_const_owner[m->device] = nullptr;
  }
  m->chain.release(); m->y.release(); m->c1.release(); m->n1.release(); m->o1.release();
  m->raw_note.release(); m->raw_onset.release(); m->raw_contour.release(); m->minmax.release(); m->edge.release();
  m->i_note.release(); m->i_onset.release(); m->i_contour.release(); m->u_note.release(); m->u_onset.release();
  m->u_contour.release();
  m->wdesc.release(); m->udesc.release(); m->st_audio.release(); m->st_note.release(); m->st_onset.release();
  m->st_contour.release(); m->st_pcm.release(); m->d_frame_off.release(); m->d_slot_off.release(); m->d_note_base.release();
  m->energy.release(); m->d_amp.release(); m->candbits.release(); m->blk_max.release(); m->blk_arg.release(); m->max_onset.release(); m->max_fd.release();
  m->note_count.release(); m->slot_start.release(); m->slot_end.release(); m->slot_pitch.release();
  m->overflow.release(); m->d_note_off.release(); m->d_start.release(); m->d_end.release(); m->d_pitch.release();
  m->d_bend_off.release(); m->d_bends.release(); m->d_onset64.release();
  m->yhl.release();
  m->chl.release();
  m->cqt_wtc.release();
  for (bp_model::TcLayer* L : {&m->tc_contour, &m->tc_onset, &m->tc_note}) L->tiles.release(), L->b2.release();
  if (m->d_params) cudaFree(m->d_params);
  if (m->d_derived) cudaFree(m->d_derived);
  if (m->d_gauss) cudaFree(m->d_gauss);
  for (cudaEvent_t e : m->prof_ev) cudaEventDestroy(e);
  if (m->desc_ev) cudaEventDestroy(m->desc_ev);
  if (m->h_wd) cudaFreeHost(m->h_wd);
  if (m->h_ud) cudaFreeHost(m->h_ud);
  for (cudaEvent_t e : m->copy_ev) cudaEventDestroy(e);
  if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
  if (m->d2h_stream) cudaStreamDestroy(m->d2h_stream);
  for (float* b : m->gather)
    if (b) cudaFreeHost(b);
  for (cudaEvent_t e : m->conv_ev) cudaEventDestroy(e);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

int bp_model_device(const bp_model_t* m) { return m ? m->device : -1; }
int64_t bp_model_launch_count(const bp_model_t* m) { return m ? m->launches : 0; }
int64_t bp_model_chunk_windows(const bp_model_t* m) { return m ? m->chunk : 0; }

int bp_model_param_block(bp_model_t* m, void** d_ptr, size_t* nbytes) {
  if (!m || !d_ptr || !nbytes) return fail(BP_E_INVALID, "bp_model_param_block: null argument");
  *d_ptr = m->d_params;
  *nbytes = sizeof(float) * ParamLayout::total;
  return BP_OK;
}

int bp_model_refresh(bp_model_t* m) {
  if (!m) return fail(BP_E_INVALID, "bp_model_refresh: null model");
  DeviceGuard g(m->device);
  CK(cudaDeviceSynchronize());
  int rc = derive(m, m->stream);
  if (rc) return rc;
  CK(cudaStreamSynchronize(m->stream));
  return BP_OK;
}

int bp_model_set_path(bp_model_t* m, int path) {
  if (!m) return fail(BP_E_INVALID, "bp_model_set_path: null model");
  if (path < 0 || path > 2)
    return fail(BP_E_INVALID, "bp_model_set_path: path must be 0 (FP32 FFMA), 1 (tcgen05, fused epilogues) or 2 (tcgen05, "
                              "contour activations kept)");
  m->path = path;
  return BP_OK;
}

int bp_forward_device(bp_model_t* m, const float* d_audio, int64_t n_windows, float* d_note, float* d_onset,
                      float* d_contour, void* stream) {
  if (!m || !d_audio || !d_note || !d_onset || !d_contour) return fail(BP_E_INVALID, "bp_forward_device: null argument");
  if (n_windows < 0) return fail(BP_E_INVALID, "bp_forward_device: negative window count");
  DeviceGuard g(m->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = ensure_forward_ws(m, (int)std::min<int64_t>(n_windows, m->chunk));
  if (rc) return rc;
  for (int64_t c0 = 0; c0 < n_windows; c0 += m->chunk) {
    int nb = (int)std::min<int64_t>(m->chunk, n_windows - c0);
    rc = forward_chunk(m, d_audio + c0 * kWinSamples, nullptr, nb, d_note + c0 * kFrames * kPitches,
                       d_onset + c0 * kFrames * kPitches, d_contour + c0 * kFrames * kContourBins, st);
    if (rc) return rc;
  }
  m->last_forward_n = n_windows;
  return BP_OK;
}

int bp_forward_host(bp_model_t* m, const float* h_audio, int64_t n_windows, float* h_note, float* h_onset,
                    float* h_contour) {
  if (!m || !h_audio || !h_note || !h_onset || !h_contour) return fail(BP_E_INVALID, "bp_forward_host: null argument");
  if (n_windows < 0) return fail(BP_E_INVALID, "bp_forward_host: negative window count");
  DeviceGuard g(m->device);
  CK(m->st_audio.reserve((size_t)n_windows * kWinSamples));
  CK(m->st_note.reserve((size_t)n_windows * kFrames * kPitches));
  CK(m->st_onset.reserve((size_t)n_windows * kFrames * kPitches));
  CK(m->st_contour.reserve((size_t)n_windows * kFrames * kContourBins));
  cudaStream_t st = m->stream;
  CK(cudaMemcpyAsync(m->st_audio.p, h_audio, sizeof(float) * n_windows * kWinSamples, cudaMemcpyHostToDevice, st));
  int rc = bp_forward_device(m, m->st_audio.p, n_windows, m->st_note.p, m->st_onset.p, m->st_contour.p, st);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h_note, m->st_note.p, sizeof(float) * n_windows * kFrames * kPitches, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_onset, m->st_onset.p, sizeof(float) * n_windows * kFrames * kPitches, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_contour, m->st_contour.p, sizeof(float) * n_windows * kFrames * kContourBins,
                     cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return BP_OK;
}

// HCQT + CNN + unwrap for a batch of files into the internal posteriorgrams `u` (frames of file i at
// frame_base + h_frame_off[i] ..); h_frame_off is relative to the batch (h_frame_off[0] = 0).
static int run_inference_internal(bp_model* m, const float* d_audio, const int64_t* h_sample_off, int32_t n_files,
                                  const PostI& u, int64_t frame_base, int64_t* h_frame_off, cudaStream_t st) {
  std::vector<WinDesc> wd;
  std::vector<UnwrapDesc> ud;
  h_frame_off[0] = 0;
  for (int i = 0; i < n_files; ++i) {
    const int64_t n = h_sample_off[i + 1] - h_sample_off[i];
    if (n < 0) return fail(BP_E_INVALID, "run_inference: sample offsets must be non-decreasing");
    const int64_t nw = bp_num_windows(n), nf = bp_num_frames(n);
    for (int64_t w = 0; w < nw; ++w) {
      const int64_t s = w * kHopSamples - kLeadZeros;  // window start relative to the file
      WinDesc d;
      d.base = h_sample_off[i] + s;
      d.lo = (int)std::max<int64_t>(0, -s);
      d.hi = (int)std::max<int64_t>(d.lo, std::min<int64_t>(kWinSamples, n - s));
      wd.push_back(d);
      UnwrapDesc x;
      x.dst_base = frame_base + h_frame_off[i] + w * kHopFrames;
      x.rows = (int)std::max<int64_t>(0, std::min<int64_t>(kHopFrames, nf - w * kHopFrames));
      x.pad = 0;
      ud.push_back(x);
    }
    h_frame_off[i + 1] = h_frame_off[i] + nf;
  }
  const int64_t nwin = (int64_t)wd.size();
  if (nwin == 0) return BP_OK;
  if (!d_audio) return fail(BP_E_INVALID, "run_inference: null audio");
  const int chunk = m->chunk;
  int rc = ensure_forward_ws(m, (int)std::min<int64_t>(nwin, chunk));
  if (rc) return rc;
  const int nbmax = (int)std::min<int64_t>(nwin, chunk);
  if (m->path == 0) {  // row-major raw windows of the FP32 kernels, converted + unwrapped per chunk
    CK(m->raw_note.reserve((size_t)nbmax * kFrames * kPitches));
    CK(m->raw_onset.reserve((size_t)nbmax * kFrames * kPitches));
  }
  if (m->path != 1) CK(m->raw_contour.reserve((size_t)nbmax * kFrames * kContourBins));
  CK(m->wdesc.reserve(nwin));
  CK(m->udesc.reserve(nwin));
  // descriptors go through pinned staging owned by the model, guarded by an event that completes with the copy (not with
  // the kernels queued behind it): no stream synchronisation here, so back-to-back calls (the sub-batches of
  // bp_transcribe_host) keep the GPU queue full
  if (m->desc_pending) {
    CK(cudaEventSynchronize(m->desc_ev));
    m->desc_pending = false;
  }
  if ((size_t)nwin > m->h_desc_cap) {
    if (m->h_wd) cudaFreeHost(m->h_wd);
    if (m->h_ud) cudaFreeHost(m->h_ud);
    m->h_wd = nullptr;
    m->h_ud = nullptr;
    m->h_desc_cap = 0;
    const size_t cap = std::max<size_t>((size_t)nwin, 4096);
    CK(cudaMallocHost(&m->h_wd, sizeof(WinDesc) * cap));
    CK(cudaMallocHost(&m->h_ud, sizeof(UnwrapDesc) * cap));
    m->h_desc_cap = cap;
  }
  if (!m->desc_ev) CK(cudaEventCreateWithFlags(&m->desc_ev, cudaEventDisableTiming));
  std::memcpy(m->h_wd, wd.data(), sizeof(WinDesc) * nwin);
  std::memcpy(m->h_ud, ud.data(), sizeof(UnwrapDesc) * nwin);
  // a kernel reads the pinned (device-mapped) staging directly: a copy-engine transfer would queue behind the audio
  // uploads that bp_transcribe_host has in flight on the copy stream
  static_assert(sizeof(WinDesc) == 16 && sizeof(UnwrapDesc) == 16, "descriptor upload moves 16-byte records");
  desc_upload_kernel<<<(unsigned)((nwin + 255) / 256), 256, 0, st>>>(
      reinterpret_cast<const int4*>(m->h_wd), reinterpret_cast<int4*>(m->wdesc.p), reinterpret_cast<const int4*>(m->h_ud),
      reinterpret_cast<int4*>(m->udesc.p), (int)nwin);
  CKL();
  m->launches += 1;
  CK(cudaEventRecord(m->desc_ev, st));
  m->desc_pending = true;
  for (int64_t c0 = 0; c0 < nwin; c0 += chunk) {
    const int nb = (int)std::min<int64_t>(chunk, nwin - c0);
    rc = forward_chunk(m, d_audio, m->wdesc.p + c0, nb, m->raw_note.p, m->raw_onset.p, m->raw_contour.p, st,
                       m->udesc.p + c0, &u);
    if (rc) return rc;
  }
  m->last_forward_n = std::min<int64_t>(nwin, chunk) == nwin ? nwin : 0;
  return BP_OK;
}

// internal posteriorgram buffers of the model for `total_frames` frames
static int reserve_internal(bp_model* m, int64_t total_frames, PostI* u) {
  const size_t F = (size_t)((total_frames + 31) / 32 * 32 + 32);
  CK(m->u_note.reserve(kPitches * F));
  CK(m->u_onset.reserve(kPitches * F));
  CK(m->u_contour.reserve((size_t)kContourBins * F));
  *u = PostI{m->u_note.p, m->u_onset.p, m->u_contour.p, (long long)F};
  return BP_OK;
}

// internal -> the row-major arrays of the C ABI (any of the destinations may be null)
static void internal_to_rows(bp_model* m, const PostI& u, int64_t total_frames, float* d_note, float* d_onset,
                             float* d_contour, cudaStream_t st) {
  if (d_note) launch_pm_to_rows(u.note, u.stride, 0, total_frames, kPitches, d_note, st), m->launches += 1;
  if (d_onset) launch_pm_to_rows(u.onset, u.stride, 0, total_frames, kPitches, d_onset, st), m->launches += 1;
  if (d_contour) launch_cm_to_rows(u.contour, u.stride, 0, total_frames, d_contour, st), m->launches += 1;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int bp_run_inference_device(bp_model_t* m, const float* d_audio, const int64_t* h_sample_off, int32_t n_files,
                            float* d_note, float* d_onset, float* d_contour, int64_t* h_frame_off, void* stream) {
  if (!m || !h_sample_off || !h_frame_off || n_files < 0)
    return fail(BP_E_INVALID, "bp_run_inference_device: bad argument");
  DeviceGuard g(m->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t total_frames = 0;
  for (int i = 0; i < n_files; ++i) total_frames += bp_num_frames(h_sample_off[i + 1] - h_sample_off[i]);
  if (total_frames > 0 && (!d_audio || !d_note || !d_onset || !d_contour))
    return fail(BP_E_INVALID, "bp_run_inference_device: null buffer");
  if (!aligned16(d_contour)) return fail(BP_E_INVALID, "bp_run_inference_device: d_contour must be 16-byte aligned");
  PostI u;
  int rc = reserve_internal(m, total_frames, &u);
  if (rc) return rc;
  rc = run_inference_internal(m, d_audio, h_sample_off, n_files, u, 0, h_frame_off, st);
  if (rc) return rc;
  internal_to_rows(m, u, total_frames, d_note, d_onset, d_contour, st);
  CKL();
  return BP_OK;
}

int bp_run_inference_host(bp_model_t* m, const float* h_audio, const int64_t* h_sample_off, int32_t n_files,
                          float* h_note, float* h_onset, float* h_contour, int64_t* h_frame_off) {
  if (!m || !h_sample_off || !h_frame_off || n_files < 0) return fail(BP_E_INVALID, "bp_run_inference_host: bad argument");
  DeviceGuard g(m->device);
  const int64_t n_samples = h_sample_off[n_files] - h_sample_off[0];
  int64_t total_frames = 0;
  for (int i = 0; i < n_files; ++i) total_frames += bp_num_frames(h_sample_off[i + 1] - h_sample_off[i]);
  cudaStream_t st = m->stream;
  CK(m->st_audio.reserve((size_t)std::max<int64_t>(n_samples, 1)));
  CK(m->st_note.reserve((size_t)total_frames * kPitches + 1));
  CK(m->st_onset.reserve((size_t)total_frames * kPitches + 1));
  CK(m->st_contour.reserve((size_t)total_frames * kContourBins + 1));
  if (n_samples > 0)
    CK(cudaMemcpyAsync(m->st_audio.p, h_audio + h_sample_off[0], sizeof(float) * n_samples, cudaMemcpyHostToDevice, st));
  std::vector<int64_t> rel(n_files + 1);
  for (int i = 0; i <= n_files; ++i) rel[i] = h_sample_off[i] - h_sample_off[0];
  int rc = bp_run_inference_device(m, m->st_audio.p, rel.data(), n_files, m->st_note.p, m->st_onset.p, m->st_contour.p,
                                   h_frame_off, st);
  if (rc) return rc;
  if (total_frames > 0) {
    if (!h_note || !h_onset || !h_contour) return fail(BP_E_INVALID, "bp_run_inference_host: null output");
    CK(cudaMemcpyAsync(h_note, m->st_note.p, sizeof(float) * total_frames * kPitches, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_onset, m->st_onset.p, sizeof(float) * total_frames * kPitches, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_contour, m->st_contour.p, sizeof(float) * total_frames * kContourBins, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  return BP_OK;
}

int bp_decode_device(bp_model_t* m, const float* d_note, const float* d_onset, const float* d_contour,
                     const int64_t* h_frame_off, int32_t n_files, const bp_decode_params_t* params, bp_notes_t* notes,
                     void* stream) {
  if (!m || !h_frame_off || !notes || n_files < 0) return fail(BP_E_INVALID, "bp_decode_device: bad argument");
  int rc = validate_params(params);
  if (rc) return rc;
  if (!notes->note_off || !notes->bend_off) return fail(BP_E_INVALID, "bp_decode_device: notes arrays missing");
  DeviceGuard g(m->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long total_frames = h_frame_off[n_files] - h_frame_off[0];
  notes->note_off[0] = 0;
  notes->bend_off[0] = 0;
  if (n_files == 0) return BP_OK;
  if (h_frame_off[0] != 0) return fail(BP_E_INVALID, "bp_decode_device: frame_off[0] must be 0");
  if (total_frames > 0 && (!d_note || !d_onset || (params->include_pitch_bends && !d_contour)))
    return fail(BP_E_INVALID, "bp_decode_device: null posteriorgram");

  DecodeParamsDev dp;
  dp.onset_thresh = params->onset_thresh;
  dp.frame_thresh = params->frame_thresh;
  dp.min_note_len = params->min_note_len;
  dp.energy_tol = params->energy_tol;
  dp.infer_onsets = params->infer_onsets;
  dp.melodia = params->melodia_trick;
  dp.lo_col = std::max(0, std::min<int>(params->min_pitch_idx, kPitches));
  dp.hi_col = std::max(0, std::min<int>(params->max_pitch_idx, kPitches));

  std::vector<long long> foff(n_files + 1), soff(n_files + 1);
  for (int i = 0; i <= n_files; ++i) {
    foff[i] = h_frame_off[i];
    if (i && foff[i] < foff[i - 1]) return fail(BP_E_INVALID, "bp_decode_device: frame offsets must be non-decreasing");
  }
  std::vector<int> counts(n_files);
  for (int attempt = 0; attempt < 2; ++attempt) {
    soff[0] = 0;
    for (int i = 0; i < n_files; ++i) {
      long long T = foff[i + 1] - foff[i];
      soff[i + 1] = soff[i] + (attempt == 0 ? std::min<long long>(T * kPitches, 8 * T + 64) : T * kPitches);
    }
    const size_t cells = (size_t)total_frames * kPitches;
    CK(m->d_frame_off.reserve(n_files + 1));
    CK(m->d_slot_off.reserve(n_files + 1));
    CK(m->energy.reserve(cells + 1));
    CK(m->candbits.reserve(cells / 32 + 2));
    CK(m->blk_max.reserve((size_t)kPitches * decode_block_slots(total_frames, n_files)));
    CK(m->blk_arg.reserve((size_t)kPitches * decode_block_slots(total_frames, n_files)));
    CK(m->max_onset.reserve(n_files));
    CK(m->max_fd.reserve(n_files));
    CK(m->note_count.reserve(n_files));
    CK(m->overflow.reserve(1));
    CK(m->slot_start.reserve((size_t)soff[n_files] + 1));
    CK(m->slot_end.reserve((size_t)soff[n_files] + 1));
    CK(m->slot_pitch.reserve((size_t)soff[n_files] + 1));
    CK(cudaMemcpyAsync(m->d_frame_off.p, foff.data(), sizeof(long long) * (n_files + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(m->d_slot_off.p, soff.data(), sizeof(long long) * (n_files + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(m->overflow.p, 0, sizeof(int), st));
    DecodeBuffers b;
    b.frame_off = m->d_frame_off.p;
    b.energy = m->energy.p;
    b.candbits = m->candbits.p;
    b.max_onset = m->max_onset.p;
    b.max_fd = m->max_fd.p;
    b.slot_off = m->d_slot_off.p;
    b.note_count = m->note_count.p;
    b.note_start = m->slot_start.p;
    b.note_end = m->slot_end.p;
    b.note_pitch = m->slot_pitch.p;
    b.overflow = m->overflow.p;
    b.blk_max = m->blk_max.p;
    b.blk_arg = m->blk_arg.p;
    {
      ProfScope ps(m, 5, st);
      launch_decode_notes(d_note, d_onset, b, n_files, total_frames, dp, st);
    }
    CKL();
    m->launches += total_frames > 0 ? 3 : 1;
    int overflow = 0;
    CK(cudaMemcpyAsync(counts.data(), m->note_count.p, sizeof(int) * n_files, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&overflow, m->overflow.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (!overflow) break;
    if (attempt == 1) return fail(BP_E_CUDA, "bp_decode_device: note slots overflowed at full capacity (internal error)");
  }
  long long n_notes = 0;
  for (int i = 0; i < n_files; ++i) {
    n_notes += counts[i];
    if (n_notes > notes->note_capacity) {
      long long need = 0;
      for (int j = 0; j < n_files; ++j) need += counts[j];
      return g_need_notes = need, fail(BP_E_CAPACITY, "bp_decode_device: note_capacity too small, need " + std::to_string(need));
    }
    notes->note_off[i + 1] = (int32_t)n_notes;
  }
  if (n_notes == 0) return BP_OK;
  if (!notes->start_frame || !notes->end_frame || !notes->pitch_midi || !notes->amplitude)
    return fail(BP_E_INVALID, "bp_decode_device: notes arrays missing");

  CK(m->d_note_off.reserve(n_files + 1));
  CK(m->d_start.reserve(n_notes));
  CK(m->d_end.reserve(n_notes));
  CK(m->d_pitch.reserve(n_notes));
  CK(m->d_amp.reserve(n_notes));
  CK(m->d_note_base.reserve(n_notes));
  CK(m->d_bend_off.reserve(n_notes + 1));
  CK(cudaMemcpyAsync(m->d_note_off.p, notes->note_off, sizeof(int) * (n_files + 1), cudaMemcpyHostToDevice, st));
  compact_notes_kernel<<<n_files, 128, 0, st>>>(m->d_frame_off.p, m->d_slot_off.p, m->d_note_off.p, m->slot_start.p,
                                                m->slot_end.p, m->slot_pitch.p, m->d_start.p, m->d_end.p, m->d_pitch.p,
                                                m->d_note_base.p);
  CKL();
  m->launches += 1;
  CK(cudaMemcpyAsync(notes->start_frame, m->d_start.p, sizeof(int) * n_notes, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(notes->end_frame, m->d_end.p, sizeof(int) * n_notes, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(notes->pitch_midi, m->d_pitch.p, sizeof(int) * n_notes, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  long long n_bends = 0;
  const int with_bends = params->include_pitch_bends ? 1 : 0;
  for (long long j = 0; j < n_notes; ++j) {
    if (with_bends) n_bends += notes->end_frame[j] - notes->start_frame[j];
    if (n_bends > 0x7fffffffLL) return fail(BP_E_CAPACITY, "bp_decode_device: more than 2^31 pitch-bend values");
    notes->bend_off[j + 1] = (int32_t)n_bends;
  }
  if (with_bends && n_bends > notes->bend_capacity)
    return g_need_bends = n_bends, fail(BP_E_CAPACITY, "bp_decode_device: bend_capacity too small, need " + std::to_string(n_bends));
  if (with_bends && n_bends > 0 && !notes->bends) return fail(BP_E_INVALID, "bp_decode_device: bends array missing");
  CK(m->d_bends.reserve((size_t)n_bends + 1));
  CK(cudaMemcpyAsync(m->d_bend_off.p, notes->bend_off, sizeof(int) * (n_notes + 1), cudaMemcpyHostToDevice, st));
  {
    ProfScope ps(m, 6, st);
    launch_note_finish(d_note, d_contour, m->d_note_base.p, m->d_start.p, m->d_end.p, m->d_pitch.p, m->d_amp.p,
                       m->d_bend_off.p, m->d_bends.p, (int)n_notes, with_bends, m->d_gauss, st);
  }
  CKL();
  m->launches += 1;
  CK(cudaMemcpyAsync(notes->amplitude, m->d_amp.p, sizeof(float) * n_notes, cudaMemcpyDeviceToHost, st));
  if (with_bends && n_bends > 0)
    CK(cudaMemcpyAsync(notes->bends, m->d_bends.p, sizeof(int) * n_bends, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return BP_OK;
}

int bp_infer_onsets_host(bp_model_t* m, const float* h_onset, const float* h_note, int64_t n_frames, double* h_out) {
  if (!m || n_frames < 0) return fail(BP_E_INVALID, "bp_infer_onsets_host: bad argument");
  if (n_frames == 0) return BP_OK;
  if (!h_onset || !h_note || !h_out) return fail(BP_E_INVALID, "bp_infer_onsets_host: null array");
  DeviceGuard g(m->device);
  cudaStream_t st = m->stream;
  const size_t cells = (size_t)n_frames * kPitches;
  CK(m->st_note.reserve(cells + 1));
  CK(m->st_onset.reserve(cells + 1));
  CK(m->energy.reserve(cells + 1));
  CK(m->candbits.reserve(cells / 32 + 2));
  CK(m->max_onset.reserve(1));
  CK(m->max_fd.reserve(1));
  CK(m->d_frame_off.reserve(2));
  CK(m->d_onset64.reserve(cells));
  const long long foff[2] = {0, (long long)n_frames};
  CK(cudaMemcpyAsync(m->d_frame_off.p, foff, sizeof(foff), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(m->st_note.p, h_note, sizeof(float) * cells, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(m->st_onset.p, h_onset, sizeof(float) * cells, cudaMemcpyHostToDevice, st));
  DecodeBuffers b{};
  b.frame_off = m->d_frame_off.p;
  b.energy = m->energy.p;
  b.candbits = m->candbits.p;
  b.max_onset = m->max_onset.p;
  b.max_fd = m->max_fd.p;
  launch_infer_onsets(m->st_note.p, m->st_onset.p, b, 1, (long long)n_frames, m->d_onset64.p, st);
  CKL();
  m->launches += 2;
  CK(cudaMemcpyAsync(h_out, m->d_onset64.p, sizeof(double) * cells, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return BP_OK;
}

int bp_pitch_bends_host(bp_model_t* m, const float* h_contour, int64_t n_frames, int32_t n_notes, const int32_t* h_start,
                        const int32_t* h_end, const int32_t* h_pitch_midi, int32_t* h_bend_off, int32_t* h_bends,
                        int64_t bend_capacity) {
  if (!m || n_frames < 0 || n_notes < 0 || !h_bend_off) return fail(BP_E_INVALID, "bp_pitch_bends_host: bad argument");
  h_bend_off[0] = 0;
  if (n_notes == 0) return BP_OK;
  if (!h_contour || !h_start || !h_end || !h_pitch_midi) return fail(BP_E_INVALID, "bp_pitch_bends_host: null array");
  long long total = 0;
  for (int j = 0; j < n_notes; ++j) {
    if (h_start[j] < 0 || h_end[j] <= h_start[j] || h_end[j] > n_frames)
      return fail(BP_E_INVALID, "bp_pitch_bends_host: note frames must satisfy 0 <= start < end <= n_frames");
    if (h_pitch_midi[j] < 21 || h_pitch_midi[j] >= 21 + kPitches)
      return fail(BP_E_INVALID, "bp_pitch_bends_host: pitch outside 21..108");
    total += h_end[j] - h_start[j];
    if (total > 0x7fffffffLL) return fail(BP_E_CAPACITY, "bp_pitch_bends_host: more than 2^31 pitch-bend values");
    h_bend_off[j + 1] = (int32_t)total;
  }
  if (total > bend_capacity) return fail(BP_E_CAPACITY, "bp_pitch_bends_host: bend_capacity too small, need " + std::to_string(total));
  if (!h_bends) return fail(BP_E_INVALID, "bp_pitch_bends_host: bends array missing");
  DeviceGuard g(m->device);
  cudaStream_t st = m->stream;
  CK(m->st_contour.reserve((size_t)n_frames * kContourBins + 1));
  CK(m->st_note.reserve((size_t)n_frames * kPitches + 1));  // the kernel also averages the note posteriorgram: zeros here
  CK(m->d_start.reserve(n_notes));
  CK(m->d_end.reserve(n_notes));
  CK(m->d_pitch.reserve(n_notes));
  CK(m->d_amp.reserve(n_notes));
  CK(m->d_note_base.reserve(n_notes));
  CK(m->d_bend_off.reserve(n_notes + 1));
  CK(m->d_bends.reserve((size_t)total + 1));
  CK(cudaMemsetAsync(m->st_note.p, 0, sizeof(float) * (size_t)n_frames * kPitches, st));
  CK(cudaMemsetAsync(m->d_note_base.p, 0, sizeof(long long) * n_notes, st));
  CK(cudaMemcpyAsync(m->st_contour.p, h_contour, sizeof(float) * (size_t)n_frames * kContourBins, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(m->d_start.p, h_start, sizeof(int) * n_notes, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(m->d_end.p, h_end, sizeof(int) * n_notes, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(m->d_pitch.p, h_pitch_midi, sizeof(int) * n_notes, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(m->d_bend_off.p, h_bend_off, sizeof(int) * (n_notes + 1), cudaMemcpyHostToDevice, st));
  launch_note_finish(m->st_note.p, m->st_contour.p, m->d_note_base.p, m->d_start.p, m->d_end.p, m->d_pitch.p, m->d_amp.p,
                     m->d_bend_off.p, m->d_bends.p, n_notes, 1, m->d_gauss, st);
  CKL();
  m->launches += 1;
  CK(cudaMemcpyAsync(h_bends, m->d_bends.p, sizeof(int) * (size_t)total, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return BP_OK;
}

int bp_decode_host(bp_model_t* m, const float* h_note, const float* h_onset, const float* h_contour,
                   const int64_t* h_frame_off, int32_t n_files, const bp_decode_params_t* params, bp_notes_t* notes) {
  if (!m || !h_frame_off || n_files < 0) return fail(BP_E_INVALID, "bp_decode_host: bad argument");
  DeviceGuard g(m->device);
  const int64_t total = h_frame_off[n_files];
  cudaStream_t st = m->stream;
  CK(m->st_note.reserve((size_t)total * kPitches + 1));
  CK(m->st_onset.reserve((size_t)total * kPitches + 1));
  CK(m->st_contour.reserve((size_t)total * kContourBins + 1));
  if (total > 0) {
    if (!h_note || !h_onset || !h_contour) return fail(BP_E_INVALID, "bp_decode_host: null posteriorgram");
    CK(cudaMemcpyAsync(m->st_note.p, h_note, sizeof(float) * total * kPitches, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(m->st_onset.p, h_onset, sizeof(float) * total * kPitches, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(m->st_contour.p, h_contour, sizeof(float) * total * kContourBins, cudaMemcpyHostToDevice, st));
  }
  return bp_decode_device(m, m->st_note.p, m->st_onset.p, m->st_contour.p, h_frame_off, n_files, params, notes, st);
}

int bp_transcribe_device(bp_model_t* m, const float* d_audio, const int64_t* h_sample_off, int32_t n_files,
                         const bp_decode_params_t* params, int64_t* h_frame_off, bp_notes_t* notes, void* stream) {
  if (!m || !h_sample_off || !h_frame_off || n_files < 0) return fail(BP_E_INVALID, "bp_transcribe_device: bad argument");
  int rc = validate_params(params);
  if (rc) return rc;
  DeviceGuard g(m->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t total_frames = 0;
  for (int i = 0; i < n_files; ++i) total_frames += bp_num_frames(h_sample_off[i + 1] - h_sample_off[i]);
  CK(m->st_note.reserve((size_t)total_frames * kPitches + 4));
  CK(m->st_onset.reserve((size_t)total_frames * kPitches + 4));
  CK(m->st_contour.reserve((size_t)total_frames * kContourBins + 4));
  PostI u;
  rc = reserve_internal(m, total_frames, &u);
  if (rc) return rc;
  rc = run_inference_internal(m, d_audio, h_sample_off, n_files, u, 0, h_frame_off, st);
  if (rc) return rc;
  internal_to_rows(m, u, total_frames, m->st_note.p, m->st_onset.p, m->st_contour.p, st);
  CKL();
  return bp_decode_device(m, m->st_note.p, m->st_onset.p, m->st_contour.p, h_frame_off, n_files, params, notes, stream);
}

int bp_transcribe_host(bp_model_t* m, const float* h_audio, const int64_t* h_sample_off, int32_t n_files,
                       const bp_decode_params_t* params, float* h_note, float* h_onset, float* h_contour,
                       int64_t* h_frame_off, bp_notes_t* notes) {
  if (!m || !h_sample_off || !h_frame_off || n_files < 0) return fail(BP_E_INVALID, "bp_transcribe_host: bad argument");
  int rc = validate_params(params);
  if (rc) return rc;
  DeviceGuard g(m->device);
  const int64_t n_samples = h_sample_off[n_files] - h_sample_off[0];
  cudaStream_t st = m->stream;
  CK(m->st_audio.reserve((size_t)std::max<int64_t>(n_samples, 1)));
  if (n_samples > 0 && !h_audio) return fail(BP_E_INVALID, "bp_transcribe_host: null audio");
  std::vector<int64_t> rel(n_files + 1);
  int64_t total_frames = 0;
  for (int i = 0; i <= n_files; ++i) rel[i] = h_sample_off[i] - h_sample_off[0];
  for (int i = 0; i < n_files; ++i) {
    if (rel[i + 1] < rel[i]) return fail(BP_E_INVALID, "bp_transcribe_host: sample offsets must be non-decreasing");
    total_frames += bp_num_frames(rel[i + 1] - rel[i]);
  }
  CK(m->st_note.reserve((size_t)total_frames * kPitches + 4));
  CK(m->st_onset.reserve((size_t)total_frames * kPitches + 4));
  CK(m->st_contour.reserve((size_t)total_frames * kContourBins + 4));
  PostI u;
  rc = reserve_internal(m, total_frames, &u);
  if (rc) return rc;

  // Sub-batches of files (about 4 internal chunks of windows each): all host->device copies are queued on the copy
  // stream up front, the compute stream waits for sub-batch k only, so the PCIe transfer of k+1.. overlaps the kernels.
  std::vector<int> cut{0};
  {
    // Sub-batches end on file boundaries and do not spill a few windows into an extra internal chunk: the first two
    // hold at most one chunk of windows (their copies are the ones that cannot be hidden), the next two at most two,
    // the rest at most four.
    int64_t w = 0;
    for (int i = 0; i < n_files; ++i) {
      const size_t k = cut.size() - 1;
      const int64_t limit = (k < 2 ? 1 : (k < 4 ? 2 : 4)) * (int64_t)m->chunk;
      const int64_t nw = bp_num_windows(rel[i + 1] - rel[i]);
      if (w > 0 && w + nw > limit) {
        cut.push_back(i);
        w = 0;
      }
      w += nw;
    }
    cut.push_back(n_files);
  }
  const size_t n_sub = cut.size() - 1;
  while (m->copy_ev.size() < n_sub) {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    m->copy_ev.push_back(e);
  }
  CK(cudaStreamSynchronize(st));  // st_audio may still be read by earlier work on the compute stream
  for (size_t k = 0; k < n_sub; ++k) {
    const int64_t s0 = rel[cut[k]], s1 = rel[cut[k + 1]];
    if (s1 > s0)
      CK(cudaMemcpyAsync(m->st_audio.p + s0, h_audio + h_sample_off[0] + s0, sizeof(float) * (s1 - s0),
                         cudaMemcpyHostToDevice, m->copy_stream));
    CK(cudaEventRecord(m->copy_ev[k], m->copy_stream));
  }
  h_frame_off[0] = 0;
  std::vector<int64_t> sub_off;
  for (size_t k = 0; k < n_sub; ++k) {
    const int f0 = cut[k], f1 = cut[k + 1];
    CK(cudaStreamWaitEvent(st, m->copy_ev[k], 0));
    sub_off.assign(f1 - f0 + 1, 0);
    const int64_t base = h_frame_off[f0];
    rc = run_inference_internal(m, m->st_audio.p, rel.data() + f0, f1 - f0, u, base, sub_off.data(), st);
    if (rc) return rc;
    for (int i = f0; i < f1; ++i) h_frame_off[i + 1] = base + sub_off[i - f0 + 1];
  }
  internal_to_rows(m, u, total_frames, m->st_note.p, m->st_onset.p, m->st_contour.p, st);
  CKL();
  rc = bp_decode_device(m, m->st_note.p, m->st_onset.p, m->st_contour.p, h_frame_off, n_files, params, notes, st);
  if (rc) return rc;
  if (total_frames > 0) {
    if (h_note) CK(cudaMemcpyAsync(h_note, m->st_note.p, sizeof(float) * total_frames * kPitches, cudaMemcpyDeviceToHost, st));
    if (h_onset) CK(cudaMemcpyAsync(h_onset, m->st_onset.p, sizeof(float) * total_frames * kPitches, cudaMemcpyDeviceToHost, st));
    if (h_contour)
      CK(cudaMemcpyAsync(h_contour, m->st_contour.p, sizeof(float) * total_frames * kContourBins, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  return BP_OK;
}

void bp_last_required(int64_t* notes, int64_t* bends) {
  if (notes) *notes = g_need_notes;
  if (bends) *bends = g_need_bends;
}

void* bp_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    fail(BP_E_CUDA, "bp_host_alloc: cudaHostAlloc failed");
    return nullptr;
  }
  return p;
}

void bp_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

namespace {
// Host threads that copy the files of one sub-batch after the other into pinned staging buffers, running ahead of the
// caller: worker t copies its share of sub-batch k as soon as the caller has released that buffer (`released` counts the
// sub-batches whose buffer may be overwritten) and reports it in done[k].  The threads live for one call.
struct Gatherer {
  const float* const* audio;
  const int64_t* rel;        // [n_files + 1] sample offsets
  const int* cut;            // [n_sub + 1] file index where each sub-batch starts
  int n_sub, n_threads;
  float* const* stage;       // 3 staging buffers
  std::atomic<int> released{0};
  std::vector<std::atomic<int>> done;
  std::vector<std::thread> threads;
  std::atomic<bool> abort{false};

  Gatherer(const float* const* a, const int64_t* r, const int* c, int ns, int nt, float* const* st)
      : audio(a), rel(r), cut(c), n_sub(ns), n_threads(nt), stage(st), done(ns) {
    for (auto& d : done) d.store(0);
    for (int t = 0; t < n_threads; ++t) threads.emplace_back([this, t] { run(t); });
  }
  ~Gatherer() {
    abort.store(true);
    for (auto& x : threads) x.join();
  }
  void run(int t) {
    for (int k = 0; k < n_sub; ++k) {
      for (int spins = 0; released.load(std::memory_order_acquire) <= k; ++spins) {  // buffer k % 3 still holds k - 3
        if (abort.load()) return;
        if (spins < 64)
          std::this_thread::yield();
        else
          std::this_thread::sleep_for(std::chrono::microseconds(50));  // do not fight the enqueueing thread for cores
      }
      const int f0 = cut[k], f1 = cut[k + 1];
      const int64_t base = rel[f0], total = rel[f1] - base;
      float* dst = stage[k % 3];
      // thread t copies samples [lo, hi) of the sub-batch: whole files where possible, split files otherwise
      const int64_t lo = base + total * t / n_threads, hi = base + total * (t + 1) / n_threads;
      if (hi > lo) {
        int i = (int)(std::upper_bound(rel + f0, rel + f1 + 1, lo) - rel) - 1;
        for (; i < f1 && rel[i] < hi; ++i) {
          const int64_t a = std::max(lo, rel[i]), b = std::min(hi, rel[i + 1]);
          if (b > a) std::memcpy(dst + (a - base), audio[i] + (a - rel[i]), sizeof(float) * (size_t)(b - a));
        }
      }
      done[k].fetch_add(1, std::memory_order_release);
    }
  }
  void release_upto(int k) { released.store(k, std::memory_order_release); }  // sub-batches < k may be gathered
  void wait(int k) {
    while (done[k].load(std::memory_order_acquire) < n_threads) std::this_thread::yield();
  }
};
}  // namespace

int bp_transcribe_files_host(bp_model_t* m, const float* const* audio, const int64_t* n_samples, int32_t n_files,
                             const bp_decode_params_t* params, float* h_note, float* h_onset, float* h_contour,
                             int64_t* h_frame_off, bp_notes_t* notes) {
  if (!m || !h_frame_off || n_files < 0 || (n_files > 0 && (!audio || !n_samples)))
    return fail(BP_E_INVALID, "bp_transcribe_files_host: bad argument");
  int rc = validate_params(params);
  if (rc) return rc;
  DeviceGuard g(m->device);
  cudaStream_t st = m->stream;
  std::vector<int64_t> rel(n_files + 1, 0);
  int64_t total_frames = 0;
  for (int i = 0; i < n_files; ++i) {
    if (n_samples[i] < 0 || (n_samples[i] > 0 && !audio[i]))
      return fail(BP_E_INVALID, "bp_transcribe_files_host: file " + std::to_string(i) + ": null audio or negative length");
    rel[i + 1] = rel[i] + n_samples[i];
    total_frames += bp_num_frames(n_samples[i]);
  }
  const int64_t total_samples = rel[n_files];
  CK(m->st_audio.reserve((size_t)std::max<int64_t>(total_samples, 1)));
  CK(m->st_note.reserve((size_t)total_frames * kPitches + 4));
  CK(m->st_onset.reserve((size_t)total_frames * kPitches + 4));
  CK(m->st_contour.reserve((size_t)total_frames * kContourBins + 4));
  PostI u;
  rc = reserve_internal(m, total_frames, &u);
  if (rc) return rc;
  // sub-batches of whole files, about two internal chunks of windows each (the first ones one chunk, so that the
  // kernels start early)
  std::vector<int> cut{0};
  {
    int64_t w = 0;
    for (int i = 0; i < n_files; ++i) {
      const size_t k = cut.size() - 1;
      const int64_t limit = (k < 2 ? 1 : 2) * (int64_t)m->chunk;
      const int64_t nw = bp_num_windows(n_samples[i]);
      if (w > 0 && w + nw > limit) {
        cut.push_back(i);
        w = 0;
      }
      w += nw;
    }
    cut.push_back(n_files);
  }
  const size_t n_sub = cut.size() - 1;
  size_t max_sub = 1;
  for (size_t k = 0; k < n_sub; ++k) max_sub = std::max<size_t>(max_sub, (size_t)(rel[cut[k + 1]] - rel[cut[k]]));
  if (max_sub > m->gather_cap) {
    for (float*& b : m->gather) {
      if (b) cudaFreeHost(b);
      b = nullptr;
    }
    m->gather_cap = 0;
    const size_t want = max_sub + max_sub / 8;
    for (float*& b : m->gather) CK(cudaHostAlloc(&b, want * sizeof(float), cudaHostAllocDefault));
    m->gather_cap = want;
  }
  if (!m->d2h_stream) CK(cudaStreamCreateWithFlags(&m->d2h_stream, cudaStreamNonBlocking));
  while (m->copy_ev.size() < n_sub || m->conv_ev.size() < n_sub) {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    (m->copy_ev.size() < n_sub ? m->copy_ev : m->conv_ev).push_back(e);
  }
  int n_threads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency() > 2 ? std::thread::hardware_concurrency() - 1 : 1u));
  if (const char* e = getenv("BP_B200_GATHER_THREADS")) n_threads = std::max(1, std::min(64, atoi(e)));
  const bool timing = getenv("BP_B200_TIMING") != nullptr;  // host-side breakdown of this call on stderr
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
  const auto t_call = now();
  double t_gather = 0, t_wait = 0, t_launch = 0;
  CK(cudaStreamSynchronize(st));  // st_audio / st_note.. may still be in use by earlier work on the compute stream
  h_frame_off[0] = 0;
  std::vector<int64_t> sub_off;
  Gatherer gatherer(audio, rel.data(), cut.data(), (int)n_sub, n_threads, m->gather);
  gatherer.release_upto(std::min<int>(3, (int)n_sub));  // the three buffers are free: the workers start at once
  for (size_t k = 0; k < n_sub; ++k) {
    const int f0 = cut[k], f1 = cut[k + 1];
    const int64_t s0 = rel[f0], s1 = rel[f1];
    float* stage = m->gather[k % 3];
    auto t0 = now();
    gatherer.wait((int)k);  // sub-batch k is in its staging buffer (gathered while earlier sub-batches were enqueued)
    t_gather += ms_since(t0);
    t0 = now();
    if (s1 > s0)
      CK(cudaMemcpyAsync(m->st_audio.p + s0, stage, sizeof(float) * (size_t)(s1 - s0), cudaMemcpyHostToDevice, m->copy_stream));
    CK(cudaEventRecord(m->copy_ev[k], m->copy_stream));
    CK(cudaStreamWaitEvent(st, m->copy_ev[k], 0));
    sub_off.assign(f1 - f0 + 1, 0);
    const int64_t base = h_frame_off[f0];
    rc = run_inference_internal(m, m->st_audio.p, rel.data() + f0, f1 - f0, u, base, sub_off.data(), st);
    if (rc) return rc;
    if (k >= 1) {  // buffer (k - 1) % 3 = (k + 2) % 3 is free once the upload of sub-batch k - 1 has left the host
      auto tw = now();
      CK(cudaEventSynchronize(m->copy_ev[k - 1]));
      t_wait += ms_since(tw);
      gatherer.release_upto((int)std::min<size_t>(k + 3, n_sub));
    }
    for (int i = f0; i < f1; ++i) h_frame_off[i + 1] = base + sub_off[i - f0 + 1];
    // this sub-batch's posteriorgrams: internal -> row-major, then to the host on their own stream while the next
    // sub-batches compute
    const int64_t nf = h_frame_off[f1] - base;
    if (nf > 0) {
      launch_pm_to_rows(u.note, u.stride, base, nf, kPitches, m->st_note.p + base * kPitches, st);
      launch_pm_to_rows(u.onset, u.stride, base, nf, kPitches, m->st_onset.p + base * kPitches, st);
      launch_cm_to_rows(u.contour, u.stride, base, nf, m->st_contour.p + base * kContourBins, st);
      m->launches += 3;
      CKL();
      CK(cudaEventRecord(m->conv_ev[k], st));
      if (h_note || h_onset || h_contour) {
        CK(cudaStreamWaitEvent(m->d2h_stream, m->conv_ev[k], 0));
        if (h_note)
          CK(cudaMemcpyAsync(h_note + base * kPitches, m->st_note.p + base * kPitches, sizeof(float) * nf * kPitches,
                             cudaMemcpyDeviceToHost, m->d2h_stream));
        if (h_onset)
          CK(cudaMemcpyAsync(h_onset + base * kPitches, m->st_onset.p + base * kPitches, sizeof(float) * nf * kPitches,
                             cudaMemcpyDeviceToHost, m->d2h_stream));
        if (h_contour)
          CK(cudaMemcpyAsync(h_contour + base * kContourBins, m->st_contour.p + base * kContourBins,
                             sizeof(float) * nf * kContourBins, cudaMemcpyDeviceToHost, m->d2h_stream));
      }
    }
    t_launch += ms_since(t0);
  }
  const auto t_dec = now();
  rc = bp_decode_device(m, m->st_note.p, m->st_onset.p, m->st_contour.p, h_frame_off, n_files, params, notes, st);
  const double ms_dec = ms_since(t_dec);
  const auto t_d2h = now();
  cudaError_t e1 = cudaStreamSynchronize(m->d2h_stream);
  if (timing)
    fprintf(stderr, "bp_transcribe_files_host: %d files, %zu sub-batches, %d gather threads: gather %.1f ms, staging waits %.1f ms, "
            "enqueue %.1f ms, decode (incl. waiting for the forward pass) %.1f ms, tail of the posteriorgram copies %.1f ms, total %.1f ms\n",
            n_files, n_sub, n_threads, t_gather, t_wait, t_launch, ms_dec, ms_since(t_d2h), ms_since(t_call));
  if (rc) return rc;
  CK(e1);
  CK(cudaStreamSynchronize(st));
  return BP_OK;
}

int64_t bp_resampled_length(int64_t n_frames, int32_t sample_rate) { return ingest_output_length(n_frames, sample_rate); }

int bp_load_pcm_device(bp_model_t* m, const void* d_pcm, int32_t sample_format, int64_t n_frames, int32_t channels,
                       int32_t sample_rate, float* d_audio, void* stream) {
  if (!m || n_frames < 0) return fail(BP_E_INVALID, "bp_load_pcm_device: bad argument");
  if (n_frames == 0) return BP_OK;
  if (!d_pcm || !d_audio) return fail(BP_E_INVALID, "bp_load_pcm_device: null buffer");
  DeviceGuard g(m->device);
  const int rc = launch_ingest(m->device, d_pcm, sample_format, n_frames, channels, sample_rate, d_audio,
                               static_cast<cudaStream_t>(stream));
  if (rc == -2)
    return fail(BP_E_INVALID, "bp_load_pcm_device: unsupported sample format / channel count / sample rate (format 0..3, "
                              "channels >= 1, rate ratio to 22 050 Hz below ~100)");
  if (rc) return fail(BP_E_CUDA, std::string("bp_load_pcm_device: ") + cudaGetErrorString(cudaGetLastError()));
  m->launches += 1;
  return BP_OK;
}

int bp_load_pcm_host(bp_model_t* m, const void* h_pcm, int32_t sample_format, int64_t n_frames, int32_t channels,
                     int32_t sample_rate, float* h_audio) {
  if (!m || n_frames < 0 || channels < 1 || sample_format < 0 || sample_format > 3)
    return fail(BP_E_INVALID, "bp_load_pcm_host: bad argument");
  if (n_frames == 0) return BP_OK;
  if (!h_pcm || !h_audio) return fail(BP_E_INVALID, "bp_load_pcm_host: null buffer");
  DeviceGuard g(m->device);
  cudaStream_t st = m->stream;
  static const int kBytes[4] = {4, 2, 4, 1};
  const size_t in_bytes = (size_t)n_frames * channels * kBytes[sample_format];
  const int64_t n_out = ingest_output_length(n_frames, sample_rate);
  CK(cudaStreamSynchronize(st));
  CK(m->st_pcm.reserve(in_bytes));
  CK(m->st_audio.reserve((size_t)n_out));
  CK(cudaMemcpyAsync(m->st_pcm.p, h_pcm, in_bytes, cudaMemcpyHostToDevice, st));
  const int rc = bp_load_pcm_device(m, m->st_pcm.p, sample_format, n_frames, channels, sample_rate, m->st_audio.p, st);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h_audio, m->st_audio.p, sizeof(float) * (size_t)n_out, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return BP_OK;
}

int64_t bp_debug_resample_filter(int32_t up, int32_t down, double* taps, int64_t capacity) {
  if (up < 1 || down < 1) return fail(BP_E_INVALID, "bp_debug_resample_filter: bad argument"), -1;
  const std::vector<double> h = ingest_filter(up, down);
  if (taps && capacity >= (int64_t)h.size()) std::memcpy(taps, h.data(), h.size() * sizeof(double));
  return (int64_t)h.size();
}

int bp_debug_tc_plan(int which, const float* w, int32_t* sizes, uint16_t* tiles, int32_t* tile_seq, uint32_t* slot_words,
                     int32_t* group_step_off, int32_t* group_ft) {
  if (!w || !sizes || which < 0 || which > 2) return fail(BP_E_INVALID, "bp_debug_tc_plan: bad argument");
  TcConvPlan pl;
  pl.build(which == 0 ? tc_contour_spec() : which == 1 ? tc_onset_spec() : tc_note_spec(), w);
  sizes[0] = pl.n_tiles;
  sizes[1] = (int32_t)pl.tile_seq.size();
  sizes[2] = pl.n_uses;
  sizes[3] = pl.n_groups;
  if (tiles) std::memcpy(tiles, pl.tiles.data(), pl.tiles.size() * 2);
  if (tile_seq) std::memcpy(tile_seq, pl.tile_seq.data(), pl.tile_seq.size() * 4);
  if (slot_words) {
    std::memcpy(slot_words, pl.slot_words[0].data(), pl.slot_words[0].size() * 4);
    std::memcpy(slot_words + pl.slot_words[0].size(), pl.slot_words[1].data(), pl.slot_words[1].size() * 4);
  }
  if (group_step_off) std::memcpy(group_step_off, pl.group_step_off.data(), pl.group_step_off.size() * 4);
  if (group_ft) std::memcpy(group_ft, pl.group_ft.data(), pl.group_ft.size() * 4);
  return BP_OK;
}

int bp_debug_tc_b2(int which, const float* w2, int32_t* sizes, uint16_t* tiles) {
  if (!w2 || !sizes || which < 0 || which > 2) return fail(BP_E_INVALID, "bp_debug_tc_b2: bad argument");
  std::vector<uint16_t> t;
  tc_build_b2(which, w2, t);
  const int n2 = which == 1 ? 16 : 32;
  sizes[0] = (int32_t)(t.size() / (2 * 16 * n2));
  sizes[1] = n2;
  sizes[2] = which == 0 ? 5 : which == 1 ? 3 : 7;
  sizes[3] = which == 0 ? 104 : which == 1 ? 32 : 64;
  sizes[4] = which == 0 ? 5 : which == 1 ? 4 : 8;
  if (tiles) std::memcpy(tiles, t.data(), t.size() * 2);
  return BP_OK;
}

int bp_model_profile(bp_model_t* m, int which) {
  if (!m) return fail(BP_E_INVALID, "bp_model_profile: null model");
  if (which < -1 || which > 6) return fail(BP_E_INVALID, "bp_model_profile: unknown kernel family");
  m->profile_which = which;
  m->prof_used = 0;
  m->prof_windows = 0;
  return BP_OK;
}

int bp_model_profile_read(bp_model_t* m, double* total_ms, int64_t* n_intervals, int64_t* n_windows) {
  if (!m || !total_ms || !n_intervals || !n_windows) return fail(BP_E_INVALID, "bp_model_profile_read: null argument");
  DeviceGuard g(m->device);
  CK(cudaDeviceSynchronize());
  double total = 0.0;
  for (size_t i = 0; i + 1 < m->prof_used; i += 2) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, m->prof_ev[i], m->prof_ev[i + 1]));
    total += ms;
  }
  *total_ms = total;
  *n_intervals = (int64_t)(m->prof_used / 2);
  *n_windows = m->prof_windows;
  return BP_OK;
}

int bp_debug_activation(bp_model_t* m, int which, float* h_out, int64_t n_windows) {
  if (!m || !h_out) return fail(BP_E_INVALID, "bp_debug_activation: null argument");
  if (n_windows <= 0 || n_windows > m->chunk || n_windows > m->last_forward_n)
    return fail(BP_E_INVALID, "bp_debug_activation: only valid for the windows of a single-chunk forward call");
  DeviceGuard g(m->device);
  const float* src = nullptr;
  size_t per = 0;
  switch (which) {
    case 0: src = m->y.p; per = (size_t)kFrames * kCqtBins; break;
    case 1: src = m->c1.p; per = (size_t)8 * kFrames * kContourBins; break;
    case 2: src = m->n1.p; per = (size_t)32 * kFrames * kPitches; break;
    case 3: src = m->o1.p; per = (size_t)32 * kFrames * kPitches; break;
    default: return fail(BP_E_INVALID, "bp_debug_activation: unknown activation id");
  }
  if (m->last_path >= 1 && which >= 2)
    return fail(BP_E_INVALID, "bp_debug_activation: the tensor-core paths never materialise the 32-channel activations "
                              "(use bp_model_set_path(m, 0))");
  if (m->last_path == 1 && which == 1)
    return fail(BP_E_INVALID, "bp_debug_activation: path 1 reduces the contour activations in the epilogue "
                              "(use bp_model_set_path(m, 2) or 0)");
  CK(cudaDeviceSynchronize());
  if (which == 0 && m->y_is_log) {  // the tensor-core paths normalise straight into the split operand: finish `y` now
    launch_lognorm(m->y.p, m->minmax.p, m->d_params + ParamLayout::bn, (int)m->last_forward_n, m->stream);
    CK(cudaStreamSynchronize(m->stream));
    m->y_is_log = false;
  }
  CK(cudaMemcpy(h_out, src, sizeof(float) * per * n_windows, cudaMemcpyDeviceToHost));
  if (which == 1 && m->last_path == 2) {  // the tensor-core path keeps this activation channels-last: return NCHW
    std::vector<float> tmp(h_out, h_out + per * n_windows);
    for (int64_t b = 0; b < n_windows; ++b)
      for (int t = 0; t < kFrames; ++t)
        for (int f = 0; f < kContourBins; ++f)
          for (int c = 0; c < 8; ++c)
            h_out[((b * 8 + c) * kFrames + t) * kContourBins + f] = tmp[((b * kFrames + t) * kContourBins + f) * 8 + c];
  }
  return BP_OK;
}

}  // extern "C"
