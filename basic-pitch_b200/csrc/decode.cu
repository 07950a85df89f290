// Note decode: posteriorgrams -> note events, bit-identical to the reference decode on the same input.
//
// Replaces reference: basic_pitch/note_creation.py
//   constrain_frequency          :314-343   (column range applied on every read; inputs untouched)
//   get_infered_onsets           :289-311   (float64, two file-global maxima)
//   output_to_notes_polyphonic   :360-511   (peak pick :398-404, onset loop :409-447, melodia :449-509)
//   get_pitch_bends              :182-219
// Semantics pinned in SURVEY.md Appendix B; oracle: oracle/decode_ref.py.
//
// Kernels
//   decode_prep_kernel   all cells of the batch in parallel: "remaining energy" E (column-major per file,
//                        frames with the pitch range applied), per-file max(onsets) and max(frame_diff)
//   decode_cand_kernel   all cells in parallel: float64 inferred onsets, strict time peaks, threshold ->
//                        one candidate bit per cell (warp ballot -> 32-bit words)
//   decode_seq_kernel    one CTA per file: candidates in (time desc, pitch desc) order through the greedy
//                        onset loop (warp-cooperative run-of-`energy_tol` scan + neighbour zeroing), then the
//                        melodia loop with per-column maxima kept in shared memory
//   note_finish_kernel   one warp per note: amplitude (NumPy pairwise float32 mean) and per-frame pitch-bend
//                        arg-max in float64
#include "kernels.cuh"

namespace bp {

__device__ __forceinline__ int find_file(const long long* __restrict__ off, int n_files, long long frame) {
  int lo = 0, hi = n_files;  // invariant: off[lo] <= frame < off[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= frame)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float constrained(const float* __restrict__ m, long long frame, int f, int lo, int hi) {
  return (f >= lo && f < hi) ? m[frame * kPitches + f] : 0.f;
}

// positive part of min(n[t]-n[t-1], n[t]-n[t-2]) in float64, zero history, rows 0,1 -> 0
__device__ __forceinline__ double frame_diff(const float* __restrict__ note, long long base, int t, int f, int lo,
                                             int hi) {
  if (t < 2) return 0.0;
  double a = (double)constrained(note, base + t, f, lo, hi);
  double d1 = a - (double)constrained(note, base + t - 1, f, lo, hi);
  double d2 = a - (double)constrained(note, base + t - 2, f, lo, hi);
  double d = d1 < d2 ? d1 : d2;
  return d < 0.0 ? 0.0 : d;
}

// ------------------------------------------------------------------------------------------------
__global__ void decode_prep_kernel(const float* __restrict__ note, const float* __restrict__ onset,
                                   const long long* __restrict__ frame_off, int n_files, float* __restrict__ energy,
                                   unsigned int* __restrict__ max_onset,          // [n_files] ordered-uint
                                   unsigned long long* __restrict__ max_fd,       // [n_files] bits of a double >= 0
                                   int lo, int hi) {
  const long long total = frame_off[n_files] * kPitches;
  const long long cell0 = (long long)blockIdx.x * blockDim.x;
  const long long cell = cell0 + threadIdx.x;
  __shared__ unsigned int s_mo[32];
  __shared__ unsigned long long s_fd[32];
  __shared__ int s_file_first, s_file_last;

  if (threadIdx.x == 0) {
    long long last_cell = cell0 + blockDim.x - 1;
    if (last_cell >= total) last_cell = total - 1;
    s_file_first = find_file(frame_off, n_files, cell0 / kPitches);
    s_file_last = find_file(frame_off, n_files, last_cell / kPitches);
  }
  __syncthreads();

  unsigned int mo = 0u;  // ordered encoding of -inf is > 0, so 0 is a safe identity for max
  unsigned long long fd_bits = 0ull;
  int file = s_file_first;
  if (cell < total) {
    const long long frame = cell / kPitches;
    const int f = (int)(cell - frame * kPitches);
    if (s_file_first != s_file_last) file = find_file(frame_off, n_files, frame);
    const long long base = frame_off[file];
    const int T = (int)(frame_off[file + 1] - base);
    const int t = (int)(frame - base);
    energy[base * kPitches + (long long)f * T + t] = constrained(note, frame, f, lo, hi);
    mo = float_to_ordered(constrained(onset, frame, f, lo, hi));
    fd_bits = (unsigned long long)__double_as_longlong(frame_diff(note, base, t, f, lo, hi));
  }
  if (s_file_first == s_file_last) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      mo = max(mo, __shfl_xor_sync(0xffffffffu, mo, o));
      fd_bits = max(fd_bits, __shfl_xor_sync(0xffffffffu, fd_bits, o));
    }
    if ((threadIdx.x & 31) == 0) {
      s_mo[threadIdx.x >> 5] = mo;
      s_fd[threadIdx.x >> 5] = fd_bits;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < (int)(blockDim.x >> 5); ++i) {
        mo = max(mo, s_mo[i]);
        fd_bits = max(fd_bits, s_fd[i]);
      }
      if (mo) atomicMax(max_onset + file, mo);
      atomicMax(max_fd + file, fd_bits);
    }
  } else if (cell < total) {
    atomicMax(max_onset + file, mo);
    atomicMax(max_fd + file, fd_bits);
  }
}

// float64 onset value used for peak picking (NaN when max(frame_diff) == 0, like the reference)
__device__ __forceinline__ double onset64(const float* __restrict__ note, const float* __restrict__ onset,
                                          long long base, int t, int f, int lo, int hi, int infer, double maxo,
                                          double maxfd) {
  double o = (double)constrained(onset, base + t, f, lo, hi);
  if (!infer) return o;
  double fd = frame_diff(note, base, t, f, lo, hi);
  double v = __ddiv_rn(__dmul_rn(maxo, fd), maxfd);
  if (v != v) return v;
  return o > v ? o : v;
}

__global__ void decode_cand_kernel(const float* __restrict__ note, const float* __restrict__ onset,
                                   const long long* __restrict__ frame_off, int n_files,
                                   const unsigned int* __restrict__ max_onset,
                                   const unsigned long long* __restrict__ max_fd, unsigned int* __restrict__ candbits,
                                   int lo, int hi, int infer, double onset_thresh) {
  const long long total = frame_off[n_files] * kPitches;
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool cand = false;
  if (cell < total) {
    const long long frame = cell / kPitches;
    const int f = (int)(cell - frame * kPitches);
    const int file = find_file(frame_off, n_files, frame);
    const long long base = frame_off[file];
    const int T = (int)(frame_off[file + 1] - base);
    const int t = (int)(frame - base);
    const double maxo = (double)ordered_to_float(max_onset[file]);
    const double maxfd = __longlong_as_double((long long)max_fd[file]);
    double val = 0.0;
    if (t >= 1 && t <= T - 2) {
      double c = onset64(note, onset, base, t, f, lo, hi, infer, maxo, maxfd);
      double p = onset64(note, onset, base, t - 1, f, lo, hi, infer, maxo, maxfd);
      double n = onset64(note, onset, base, t + 1, f, lo, hi, infer, maxo, maxfd);
      if (c > p && c > n) val = c;
    }
    cand = val >= onset_thresh;
  }
  unsigned int bits = __ballot_sync(0xffffffffu, cand);
  if ((threadIdx.x & 31) == 0 && cell < total) candbits[cell >> 5] = bits;
}

// ------------------------------------------------------------------------------------------------
// Sequential part, one CTA per file.
// ------------------------------------------------------------------------------------------------
constexpr int kSeqThreads = 128;

// Warp-cooperative scan used by both loops: starting at frame i (step +1 or -1) with run counter 0, walk while
// `in range` and run < tol, counting consecutive cells of column `col` that are below the threshold.
// Returns the exit index and the run length in *run.  `limit`: forward -> stop when i >= limit (= T-1);
// backward -> stop when i <= limit (= 0).
__device__ __forceinline__ int scan_run(const float* col, int i, int dir, int limit, int tol, double thresh,
                                        int* run) {
  const int lane = threadIdx.x & 31;
  int k = 0;
  while (true) {
    int idx = i + dir * lane;
    bool valid = dir > 0 ? (idx < limit) : (idx > limit);
    float e = valid ? col[idx] : 0.f;
    unsigned int mv = __ballot_sync(0xffffffffu, valid);
    unsigned int mb = __ballot_sync(0xffffffffu, valid && ((double)e < thresh));
    bool stop = false;
#pragma unroll 1
    for (int j = 0; j < 32; ++j) {
      if (!((mv >> j) & 1u) || k >= tol) {
        stop = true;
        break;
      }
      k = ((mb >> j) & 1u) ? k + 1 : 0;
      i += dir;
    }
    if (stop || k >= tol) break;
  }
  *run = k;
  return i;
}

__device__ __forceinline__ void zero_cols(float* E, int T, int f, int t_lo, int t_hi /* exclusive */) {
  const int lane = threadIdx.x & 31;
  for (int t = t_lo + lane; t < t_hi; t += 32) {
    E[(long long)f * T + t] = 0.f;
    if (f < kPitches - 1) E[(long long)(f + 1) * T + t] = 0.f;
    if (f > 0) E[(long long)(f - 1) * T + t] = 0.f;
  }
}

__device__ void column_max(const float* col, int T, float* out_v, int* out_t) {
  const int lane = threadIdx.x & 31;
  float bv = -INFINITY;
  int bt = 0x7fffffff;
  for (int t = lane; t < T; t += 32) {
    float v = col[t];
    if (v > bv) {
      bv = v;
      bt = t;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    int ot = __shfl_xor_sync(0xffffffffu, bt, o);
    if (ov > bv || (ov == bv && ot < bt)) {
      bv = ov;
      bt = ot;
    }
  }
  *out_v = bv;
  *out_t = bt;
}

__global__ void __launch_bounds__(kSeqThreads) decode_seq_kernel(
    const long long* __restrict__ frame_off, float* __restrict__ energy, const unsigned int* __restrict__ candbits,
    const long long* __restrict__ slot_off, int* __restrict__ note_count, int* __restrict__ note_start,
    int* __restrict__ note_end, int* __restrict__ note_pitch, int* __restrict__ overflow, DecodeParamsDev p) {
  const int file = blockIdx.x;
  const long long base = frame_off[file];
  const int T = (int)(frame_off[file + 1] - base);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long slot0 = slot_off[file];
  const int cap = (int)(slot_off[file + 1] - slot0);
  float* E = energy + base * kPitches;

  __shared__ float s_cmax[kPitches];
  __shared__ int s_carg[kPitches];
  __shared__ int s_pick[3];  // tm, f, done
  __shared__ int s_count;

  int count = 0;
  if (T <= 0) {
    if (threadIdx.x == 0) note_count[file] = 0;
    return;
  }

  // ---------------- onset loop (warp 0) ----------------
  if (warp == 0) {
    const long long c_lo = base * kPitches;             // first cell of this file
    const long long c_hi = c_lo + (long long)T * kPitches;  // one past the last
    long long w_hi = (c_hi - 1) >> 5, w_lo = c_lo >> 5;
    for (long long wbase = w_hi; wbase >= w_lo; wbase -= 32) {
      // lanes look at words wbase, wbase-1, ... (descending)
      long long w = wbase - lane;
      unsigned int bits = 0u;
      if (w >= w_lo) {
        bits = candbits[w];
        long long first = w << 5;
        if (first < c_lo) bits &= 0xffffffffu << (int)(c_lo - first);
        if (first + 32 > c_hi) bits &= 0xffffffffu >> (int)(first + 32 - c_hi);
      }
      unsigned int nonzero = __ballot_sync(0xffffffffu, bits != 0u);
      while (nonzero) {
        int src = __ffs(nonzero) - 1;  // lowest lane = highest word
        nonzero &= nonzero - 1;
        unsigned int wb = __shfl_sync(0xffffffffu, bits, src);
        long long wcell = (wbase - src) << 5;
        while (wb) {
          int bit = 31 - __clz(wb);
          wb &= ~(1u << bit);
          long long cell = wcell + bit - c_lo;
          int t0 = (int)(cell / kPitches);
          int f = (int)(cell - (long long)t0 * kPitches);
          if (t0 >= T - 1) continue;
          int k;
          int i = scan_run(E + (long long)f * T, t0 + 1, +1, T - 1, p.energy_tol, p.frame_thresh, &k);
          i -= k;
          if (i - t0 <= p.min_note_len) continue;
          zero_cols(E, T, f, t0, i);
          __syncwarp();
          if (lane == 0) {
            if (count < cap) {
              note_start[slot0 + count] = t0;
              note_end[slot0 + count] = i;
              note_pitch[slot0 + count] = f + 21;
            } else {
              *overflow = 1;
            }
          }
          ++count;
        }
      }
    }
    if (lane == 0) s_count = count;
  }
  __syncthreads();
  count = s_count;

  // ---------------- melodia loop ----------------
  if (p.melodia) {
    for (int f = warp; f < kPitches; f += kSeqThreads / 32) {
      float v;
      int t;
      column_max(E + (long long)f * T, T, &v, &t);
      if (lane == 0) {
        s_cmax[f] = v;
        s_carg[f] = t;
      }
    }
    while (true) {
      __syncthreads();
      if (warp == 0) {
        float bv = -INFINITY;
        int bt = 0x7fffffff, bf = 0x7fffffff;
        for (int f = lane; f < kPitches; f += 32) {
          float v = s_cmax[f];
          int t = s_carg[f];
          if (v > bv || (v == bv && (t < bt || (t == bt && f < bf)))) {
            bv = v;
            bt = t;
            bf = f;
          }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
          float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          int ot = __shfl_xor_sync(0xffffffffu, bt, o);
          int of = __shfl_xor_sync(0xffffffffu, bf, o);
          if (ov > bv || (ov == bv && (ot < bt || (ot == bt && of < bf)))) {
            bv = ov;
            bt = ot;
            bf = of;
          }
        }
        bool go = (double)bv > p.frame_thresh;
        if (go) {
          const int tm = bt, f = bf;
          float* col = E + (long long)f * T;
          if (lane == 0) col[tm] = 0.f;
          __syncwarp();
          int k;
          int i = scan_run(col, tm + 1, +1, T - 1, p.energy_tol, p.frame_thresh, &k);
          zero_cols(E, T, f, tm + 1, i);
          const int t_end = i - 1 - k;
          __syncwarp();
          i = scan_run(col, tm - 1, -1, 0, p.energy_tol, p.frame_thresh, &k);
          zero_cols(E, T, f, i + 1, tm);
          const int t_start = i + 1 + k;
          __syncwarp();
          if (t_end - t_start > p.min_note_len) {
            if (lane == 0) {
              if (count < cap) {
                note_start[slot0 + count] = t_start;
                note_end[slot0 + count] = t_end;
                note_pitch[slot0 + count] = f + 21;
              } else {
                *overflow = 1;
              }
            }
            ++count;
          }
          if (lane == 0) {
            s_pick[0] = tm;
            s_pick[1] = f;
          }
        }
        if (lane == 0) s_pick[2] = go ? 0 : 1;
      }
      __syncthreads();
      if (s_pick[2]) break;
      // refresh the maxima of the three touched columns
      {
        const int f = s_pick[1] - 1 + warp;
        if (warp < 3 && f >= 0 && f < kPitches) {
          float v;
          int t;
          column_max(E + (long long)f * T, T, &v, &t);
          if (lane == 0) {
            s_cmax[f] = v;
            s_carg[f] = t;
          }
        }
      }
    }
  }
  if (threadIdx.x == 0) note_count[file] = count;
}

void launch_decode_notes(const float* note, const float* onset, const DecodeBuffers& b, int n_files,
                         long long total_frames, const DecodeParamsDev& p, cudaStream_t st) {
  const long long cells = total_frames * kPitches;
  cudaMemsetAsync(b.max_onset, 0, sizeof(unsigned int) * n_files, st);
  cudaMemsetAsync(b.max_fd, 0, sizeof(unsigned long long) * n_files, st);
  if (cells > 0) {
    const int threads = 256;
    const unsigned int blocks = (unsigned int)((cells + threads - 1) / threads);
    decode_prep_kernel<<<blocks, threads, 0, st>>>(note, onset, b.frame_off, n_files, b.energy, b.max_onset, b.max_fd,
                                                   p.lo_col, p.hi_col);
    decode_cand_kernel<<<blocks, threads, 0, st>>>(note, onset, b.frame_off, n_files, b.max_onset, b.max_fd,
                                                    b.candbits, p.lo_col, p.hi_col, p.infer_onsets, p.onset_thresh);
  }
  decode_seq_kernel<<<n_files, kSeqThreads, 0, st>>>(b.frame_off, b.energy, b.candbits, b.slot_off, b.note_count,
                                                     b.note_start, b.note_end, b.note_pitch, b.overflow, p);
}

// ------------------------------------------------------------------------------------------------
// Amplitude + pitch bends.
// ------------------------------------------------------------------------------------------------
// float32 pairwise summation exactly as NumPy's add.reduce inner loop does it (blocks of 8 partial sums up to
// 128 elements, recursive halving above), so that amplitude == np.mean(frames[start:end, f]) bit for bit.
__device__ float np_pairwise_sum(const float* a, int n, int stride) {
  if (n < 8) {
    float res = 0.f;
    for (int i = 0; i < n; ++i) res = __fadd_rn(res, a[(long long)i * stride]);
    return res;
  } else if (n <= 128) {
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[(long long)j * stride];
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[(long long)(i + j) * stride]);
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, a[(long long)i * stride]);
    return res;
  } else {
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(np_pairwise_sum(a, n2, stride), np_pairwise_sum(a + (long long)n2 * stride, n - n2, stride));
  }
}

__global__ void note_finish_kernel(const float* __restrict__ note, const float* __restrict__ contour,
                                   const long long* __restrict__ note_base, const int* __restrict__ start,
                                   const int* __restrict__ end, const int* __restrict__ pitch, float* __restrict__ amp,
                                   const int* __restrict__ bend_off, int* __restrict__ bends, int n_notes,
                                   int with_bends, const double* __restrict__ gauss) {
  const int warps_per_block = blockDim.x >> 5;
  const int n = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= n_notes) return;
  const long long base = note_base[n];
  const int t0 = start[n], t1 = end[n], col = pitch[n] - 21;
  if (lane == 0) {
    const int len = t1 - t0;
    float s = np_pairwise_sum(note + (base + t0) * kPitches + col, len, kPitches);
    amp[n] = __fdiv_rn(s, (float)len);
  }
  if (!with_bends) return;
  // reference: note_creation.py:198-218 ; contour bin of the note = 3*(pitch-21)
  const int c = 3 * col;
  const int lo = max(c - 25, 0);
  const int hi = min(kContourBins, c + 26);
  const int g0 = max(0, 25 - c);       // first Gaussian tap used
  const int shift = 25 - g0;
  const int width = hi - lo;           // <= 51
  int* out = bends + bend_off[n];
  for (int t = t0; t < t1; ++t) {
    const float* row = contour + (base + t) * kContourBins + lo;
    double bv = -1.0;
    int bi = 0x7fffffff;
    for (int j = lane; j < width; j += 32) {
      double v = __dmul_rn((double)row[j], gauss[g0 + j]);
      if (v > bv) {
        bv = v;
        bi = j;
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) out[t - t0] = bi - shift;
  }
}

void launch_note_finish(const float* note, const float* contour, const long long* note_base, const int* start,
                        const int* end, const int* pitch, float* amp, const int* bend_off, int* bends, int n_notes,
                        int with_bends, const double* gauss, cudaStream_t st) {
  if (n_notes <= 0) return;
  const int threads = 128;
  const int blocks = (n_notes + 3) / 4;
  note_finish_kernel<<<blocks, threads, 0, st>>>(note, contour, note_base, start, end, pitch, amp, bend_off, bends,
                                                 n_notes, with_bends, gauss);
}

}  // namespace bp
