// Audio ingest on the device: PCM samples of any rate / channel count -> mono float32 at 22 050 Hz.
// Replaces `librosa.load(path, sr=22050, mono=True)` minus the container decode (reference:
// basic_pitch/inference.py:239): sample-format conversion, channel mean and a rational polyphase resampler in one kernel,
// so that a file crosses PCIe once, as the PCM it was stored as (2 bytes per sample for 16-bit audio).
//
// Resampler = the Kaiser-windowed FIR of basic_pitch_b200/audio_io.py (pass band 0.913 x Nyquist of the lower rate, 125 dB
// stop band: the shape of soxr's HQ preset that librosa uses), applied exactly like scipy.signal.resample_poly:
//   y[k] = sum_i x[i] * up * h[k * down + half - i * up],   n_out = ceil(n_in * up / down),  zero outside the signal
// in polyphase form: phase p = (k * down + half) % up, newest input i_max = (k * down + half) / up,
//   y[k] = sum_j hp[p][j] * x[i_max - j],   hp[p][j] = up * h[p + j * up].
// A CTA produces 256 consecutive outputs from an input span staged (converted and down-mixed) in shared memory.
#include <cmath>
#include <map>
#include <mutex>
#include <vector>

#include "kernels.cuh"

namespace bp {

namespace {
constexpr int kOutPerCta = 256;

struct Resampler {
  int up = 1, down = 1, half = 0, taps_per_phase = 0;
  float* d_hp = nullptr;  // [up][taps_per_phase]
};

double bessel_i0(double x) {  // power series, converges quickly for the arguments of a 125 dB Kaiser window (beta ~ 12.8)
  double sum = 1.0, term = 1.0;
  const double q = x * x / 4.0;
  for (int k = 1; k < 200; ++k) {
    term *= q / ((double)k * k);
    sum += term;
    if (term < 1e-18 * sum) break;
  }
  return sum;
}

// scipy.signal.kaiserord(125, width) + firwin(numtaps | 1, cutoff, window=("kaiser", beta)), as audio_io._resample_filter
std::vector<double> design_filter(int up, int down) {
  const int m = std::max(up, down);
  const double pass_edge = 0.913 / m, stop_edge = 1.0 / m, width = stop_edge - pass_edge;
  const double A = 125.0, beta = 0.1102 * (A - 8.7);
  int numtaps = (int)std::ceil((A - 7.95) / 2.285 / (M_PI * width) + 1.0);
  numtaps |= 1;
  const double cutoff = 0.5 * (pass_edge + stop_edge), alpha = 0.5 * (numtaps - 1);
  std::vector<double> h(numtaps);
  double sum = 0.0;
  const double i0b = bessel_i0(beta);
  for (int n = 0; n < numtaps; ++n) {
    const double t = n - alpha, x = cutoff * t;
    const double sinc = (x == 0.0) ? 1.0 : std::sin(M_PI * x) / (M_PI * x);
    const double r = t / alpha;
    const double w = bessel_i0(beta * std::sqrt(std::max(0.0, 1.0 - r * r))) / i0b;
    h[n] = cutoff * sinc * w;
    sum += h[n];
  }
  for (double& v : h) v /= sum;  // unit gain at DC
  return h;
}

std::mutex g_mu;
std::map<long long, Resampler> g_resamplers[64];  // per device, keyed by up << 32 | down

template <typename T>
__device__ __forceinline__ float to_float(T v);
template <>
__device__ __forceinline__ float to_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_float<short>(short v) { return (float)v / 32768.f; }
template <>
__device__ __forceinline__ float to_float<int>(int v) { return (float)v / 2147483648.f; }
template <>
__device__ __forceinline__ float to_float<unsigned char>(unsigned char v) { return ((float)v - 128.f) / 128.f; }

// mono sample i of the interleaved PCM: numpy's float32 mean over the channel axis (sequential sum, then / channels)
template <typename T>
__device__ __forceinline__ float mono_sample(const T* pcm, long long i, int channels) {
  const T* p = pcm + i * channels;
  if (channels == 1) return to_float<T>(p[0]);
  float s = to_float<T>(p[0]);
  for (int c = 1; c < channels; ++c) s = __fadd_rn(s, to_float<T>(p[c]));
  return __fdiv_rn(s, (float)channels);
}

template <typename T>
__global__ void __launch_bounds__(kOutPerCta) ingest_kernel(const T* __restrict__ pcm, long long n_in, int channels, int up,
                                                            int down, int half, int taps, const float* __restrict__ hp,
                                                            float* __restrict__ out, long long n_out, int span) {
  extern __shared__ float s_x[];
  const long long k0 = (long long)blockIdx.x * kOutPerCta;
  // inputs the CTA's outputs touch: i in [i_lo, i_lo + span)
  const long long i_lo = (k0 * down + half) / up - (taps - 1);
  for (int t = threadIdx.x; t < span; t += kOutPerCta) {
    const long long i = i_lo + t;
    s_x[t] = (i >= 0 && i < n_in) ? mono_sample<T>(pcm, i, channels) : 0.f;
  }
  __syncthreads();
  const long long k = k0 + threadIdx.x;
  if (k >= n_out) return;
  if (up == 1 && down == 1) {  // same rate: format conversion + down-mix only
    out[k] = s_x[(int)(k - i_lo)];
    return;
  }
  const long long t0 = k * down + half;
  const int p = (int)(t0 % up);
  const int newest = (int)(t0 / up - i_lo);  // index of x[i_max] in the staged span
  const float* h = hp + (size_t)p * taps;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;  // four partial sums: shorter dependency chains, smaller rounding error
  int j = 0;
  for (; j + 4 <= taps; j += 4) {
    a0 = fmaf(__ldg(h + j), s_x[newest - j], a0);
    a1 = fmaf(__ldg(h + j + 1), s_x[newest - j - 1], a1);
    a2 = fmaf(__ldg(h + j + 2), s_x[newest - j - 2], a2);
    a3 = fmaf(__ldg(h + j + 3), s_x[newest - j - 3], a3);
  }
  for (; j < taps; ++j) a0 = fmaf(__ldg(h + j), s_x[newest - j], a0);
  out[k] = (a0 + a1) + (a2 + a3);
}

long long gcd_ll(long long a, long long b) {
  while (b) {
    const long long t = a % b;
    a = b;
    b = t;
  }
  return a;
}
}  // namespace

// host-only: the designed low-pass (before the gain `up`), for tests against scipy's firwin
std::vector<double> ingest_filter(int up, int down) { return design_filter(up, down); }

long long ingest_output_length(long long n_frames, int sample_rate) {
  if (n_frames <= 0 || sample_rate <= 0) return 0;
  const long long g = gcd_ll(kSampleRate, sample_rate);
  const long long up = kSampleRate / g, down = sample_rate / g;
  return (n_frames * up + down - 1) / down;
}

// 0 on success; -1 CUDA error, -2 unsupported argument
int launch_ingest(int device, const void* d_pcm, int format, long long n_frames, int channels, int sample_rate, float* d_out,
                  cudaStream_t st) {
  if (n_frames <= 0) return 0;
  if (channels < 1 || sample_rate < 1 || format < 0 || format > 3 || device < 0 || device >= 64) return -2;
  const long long g = gcd_ll(kSampleRate, sample_rate);
  const int up = (int)(kSampleRate / g), down = (int)(sample_rate / g);
  Resampler rs;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto& cache = g_resamplers[device];
    const long long key = ((long long)up << 32) | (unsigned)down;
    auto it = cache.find(key);
    if (it == cache.end()) {
      rs.up = up, rs.down = down;
      if (up != 1 || down != 1) {
        const std::vector<double> h = design_filter(up, down);
        const int L = (int)h.size();
        rs.half = (L - 1) / 2;
        rs.taps_per_phase = (L + up - 1) / up;
        std::vector<float> hp((size_t)up * rs.taps_per_phase, 0.f);
        for (int p = 0; p < up; ++p)
          for (int j = 0; p + (long long)j * up < L; ++j) hp[(size_t)p * rs.taps_per_phase + j] = (float)(up * h[p + (size_t)j * up]);
        if (cudaMalloc(&rs.d_hp, hp.size() * sizeof(float)) != cudaSuccess) return -1;
        if (cudaMemcpy(rs.d_hp, hp.data(), hp.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
      } else {
        rs.taps_per_phase = 1;
      }
      it = cache.emplace(key, rs).first;
    }
    rs = it->second;
  }
  const long long n_out = ingest_output_length(n_frames, sample_rate);
  // staged inputs per CTA: the newest input of the last output minus the oldest of the first, plus one
  const int span = (int)(((long long)(kOutPerCta - 1) * down + up - 1) / up + rs.taps_per_phase + 1);
  const size_t smem = (size_t)span * sizeof(float);
  if (smem > 200 * 1024) return -2;  // rate ratios beyond ~100:1
  const unsigned grid = (unsigned)((n_out + kOutPerCta - 1) / kOutPerCta);
#define BP_INGEST(T)                                                                                                       \
  do {                                                                                                                     \
    if (smem > 48 * 1024) cudaFuncSetAttribute(ingest_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
    ingest_kernel<T><<<grid, kOutPerCta, smem, st>>>(static_cast<const T*>(d_pcm), n_frames, channels, rs.up, rs.down,     \
                                                     rs.half, rs.taps_per_phase, rs.d_hp, d_out, n_out, span);             \
  } while (0)
  switch (format) {
    case 0: BP_INGEST(float); break;
    case 1: BP_INGEST(short); break;
    case 2: BP_INGEST(int); break;
    default: BP_INGEST(unsigned char); break;
  }
#undef BP_INGEST
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace bp
