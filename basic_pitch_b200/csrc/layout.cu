// Layout conversions between the device-internal posteriorgram layouts (frame index fastest: what the fused conv
// epilogues store and the decode kernels read with coalesced accesses, see TcOut in kernels.cuh)
//   pitch-major   pm[pitch][frame]            note / onset
//   chunk-major   cm[8-bin chunk][frame][8]   contour
// and the row-major [frame][bins] arrays of the C ABI (= the reference's numpy arrays, reference:
// basic_pitch/inference.py:247-279 `unwrap_output`).  Both directions go through a shared-memory tile so that global
// reads and writes are coalesced on both sides.  With `ud` the source is a chunk of raw windows [B][172][bins] and only
// the centre frames of every window are moved, to their unwrapped position (the FP32 reference path, which produces
// row-major raw windows).
#include "kernels.cuh"

namespace bp {

namespace {

// Segment of consecutive frames a block works on: without `ud` the whole range, with `ud` the kept centre frames of
// window blockIdx.y.
struct Seg {
  long long src0, dst0;
  int n;
};
__device__ __forceinline__ Seg segment(const UnwrapDesc* __restrict__ ud, long long n_frames, long long dst_frame0) {
  if (!ud) return Seg{0, dst_frame0, (int)min(n_frames, (long long)0x7fffffff)};
  const UnwrapDesc u = ud[blockIdx.y];
  return Seg{(long long)blockIdx.y * kFrames + kOverlapHalf, u.dst_base, max(u.rows, 0)};
}

// rows [frame][width] -> pm[pitch][stride]; block (32, 8), tile 32 frames x 32 pitches
__global__ void rows_to_pm_kernel(const float* __restrict__ rows, long long n_frames, int width, float* __restrict__ pm,
                                  long long stride, long long dst_frame0, const UnwrapDesc* __restrict__ ud) {
  __shared__ float tile[32][33];
  const Seg sg = segment(ud, n_frames, dst_frame0);
  const long long i0 = (long long)blockIdx.x * 32;
  if (i0 >= sg.n) return;
  const int f0 = blockIdx.z * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const long long i = i0 + r;
    const int f = f0 + threadIdx.x;
    tile[r][threadIdx.x] = (i < sg.n && f < width) ? rows[(sg.src0 + i) * width + f] : 0.f;
  }
  __syncthreads();
  for (int c = threadIdx.y; c < 32; c += 8) {
    const long long i = i0 + threadIdx.x;
    const int f = f0 + c;
    if (i < sg.n && f < width) pm[(long long)f * stride + sg.dst0 + i] = tile[threadIdx.x][c];
  }
}

// pm[pitch][stride] (frames src_frame0 ..) -> rows [frame][width]
__global__ void pm_to_rows_kernel(const float* __restrict__ pm, long long stride, long long src_frame0, long long n_frames,
                                  int width, float* __restrict__ rows) {
  __shared__ float tile[32][33];
  const long long i0 = (long long)blockIdx.x * 32;
  const int f0 = blockIdx.z * 32;
  for (int c = threadIdx.y; c < 32; c += 8) {
    const long long i = i0 + threadIdx.x;
    const int f = f0 + c;
    tile[c][threadIdx.x] = (i < n_frames && f < width) ? pm[(long long)f * stride + src_frame0 + i] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const long long i = i0 + r;
    const int f = f0 + threadIdx.x;
    if (i < n_frames && f < width) rows[i * width + f] = tile[threadIdx.x][r];
  }
}

constexpr int kCmChunks = kContourBins / 8;  // 33
constexpr int kCmTileStride = 268;           // floats per staged row: 16-byte accesses of 8 lanes hit distinct banks

// rows [frame][264] -> cm[chunk][stride][8]; block 256, tile 32 frames x 33 chunks
__global__ void rows_to_cm_kernel(const float* __restrict__ rows, long long n_frames, float* __restrict__ cm, long long stride,
                                  long long dst_frame0, const UnwrapDesc* __restrict__ ud) {
  __shared__ __align__(16) float tile[32 * kCmTileStride];
  const Seg sg = segment(ud, n_frames, dst_frame0);
  const long long i0 = (long long)blockIdx.x * 32;
  if (i0 >= sg.n) return;
  const int nrows = (int)min((long long)32, sg.n - i0);
  const float4* src = reinterpret_cast<const float4*>(rows + (sg.src0 + i0) * kContourBins);  // 264 * 4 B = 66 float4 per row
  for (int e = threadIdx.x; e < nrows * (kContourBins / 4); e += blockDim.x) {
    const int r = e / (kContourBins / 4), k4 = e - r * (kContourBins / 4);
    *reinterpret_cast<float4*>(tile + r * kCmTileStride + k4 * 4) = src[e];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 32 * kCmChunks; e += blockDim.x) {
    const int c = e >> 5, r = e & 31;
    if (r < nrows) {
      float4* d = reinterpret_cast<float4*>(cm + ((long long)c * stride + sg.dst0 + i0 + r) * 8);
      const float4* s = reinterpret_cast<const float4*>(tile + r * kCmTileStride + c * 8);
      d[0] = s[0];
      d[1] = s[1];
    }
  }
}

__global__ void cm_to_rows_kernel(const float* __restrict__ cm, long long stride, long long src_frame0, long long n_frames,
                                  float* __restrict__ rows) {
  __shared__ __align__(16) float tile[32 * kCmTileStride];
  const long long i0 = (long long)blockIdx.x * 32;
  const int nrows = (int)min((long long)32, n_frames - i0);
  for (int e = threadIdx.x; e < 32 * kCmChunks; e += blockDim.x) {
    const int c = e >> 5, r = e & 31;
    if (r < nrows) {
      const float4* s = reinterpret_cast<const float4*>(cm + ((long long)c * stride + src_frame0 + i0 + r) * 8);
      float4* d = reinterpret_cast<float4*>(tile + r * kCmTileStride + c * 8);
      d[0] = s[0];
      d[1] = s[1];
    }
  }
  __syncthreads();
  float4* dst = reinterpret_cast<float4*>(rows + i0 * kContourBins);
  for (int e = threadIdx.x; e < nrows * (kContourBins / 4); e += blockDim.x) {
    const int r = e / (kContourBins / 4), k4 = e - r * (kContourBins / 4);
    dst[e] = *reinterpret_cast<const float4*>(tile + r * kCmTileStride + k4 * 4);
  }
}

}  // namespace

void launch_rows_to_pm(const float* rows, long long n_frames, int width, float* pm, long long stride, long long dst_frame0,
                       cudaStream_t st, const UnwrapDesc* ud, int n_windows) {
  if (ud ? n_windows <= 0 : n_frames <= 0) return;
  const dim3 grid(ud ? (kHopFrames + 31) / 32 : (unsigned)((n_frames + 31) / 32), ud ? n_windows : 1, (width + 31) / 32);
  rows_to_pm_kernel<<<grid, dim3(32, 8), 0, st>>>(rows, n_frames, width, pm, stride, dst_frame0, ud);
}

void launch_pm_to_rows(const float* pm, long long stride, long long src_frame0, long long n_frames, int width, float* rows,
                       cudaStream_t st) {
  if (n_frames <= 0) return;
  const dim3 grid((unsigned)((n_frames + 31) / 32), 1, (width + 31) / 32);
  pm_to_rows_kernel<<<grid, dim3(32, 8), 0, st>>>(pm, stride, src_frame0, n_frames, width, rows);
}

void launch_rows_to_cm(const float* rows, long long n_frames, float* cm, long long stride, long long dst_frame0,
                       cudaStream_t st, const UnwrapDesc* ud, int n_windows) {
  if (ud ? n_windows <= 0 : n_frames <= 0) return;
  const dim3 grid(ud ? (kHopFrames + 31) / 32 : (unsigned)((n_frames + 31) / 32), ud ? n_windows : 1);
  rows_to_cm_kernel<<<grid, 256, 0, st>>>(rows, n_frames, cm, stride, dst_frame0, ud);
}

void launch_cm_to_rows(const float* cm, long long stride, long long src_frame0, long long n_frames, float* rows,
                       cudaStream_t st) {
  if (n_frames <= 0) return;
  cm_to_rows_kernel<<<(unsigned)((n_frames + 31) / 32), 256, 0, st>>>(cm, stride, src_frame0, n_frames, rows);
}

}  // namespace bp
