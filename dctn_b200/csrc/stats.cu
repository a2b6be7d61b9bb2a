// Statistics kernels of the empirical-std initialisation / intermediate-representation logging path:
//
//   * value_stats: (sum, sum of squares) of a dense tensor accumulated in double into a device-resident pair — the
//     reduction `output.std(unbiased=False)` needs (dctn/eps.py:163-181) without a torch.cat of all slices and a second
//     pass over the concatenated tensor; dctn_eps_forward_stats() runs it on the forward's output while it is still in L2;
//   * window_stats: mean / variance of all K x K windows seen as rank-one tensors, WITHOUT expanding them — the identity
//     of dctn/rank_one_tensor.py:14-110 (sum of a rank-one tensor = product of the factor sums, squared Frobenius norm =
//     product of the factor squared norms) applied to the windows dctn/align.py:49-61 builds with torch.stack: here two
//     numbers per pixel (pixel_sums) and one product per window, nothing K*K times the input is ever materialised.
//
// All three kernels are HBM-bound streaming reductions: coalesced loads, per-thread double accumulators, a fixed-order
// block tree and a fixed-order final pass (deterministic; no atomics).
#include "common.cuh"
#include "eps_kernels.h"

namespace {

constexpr int ST_THREADS = 256;
constexpr int ST_MAX_BLOCKS = 148 * 4;

__device__ __forceinline__ void block_reduce2(double& a, double& b, double* sh /* [2][ST_THREADS/32] */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_down_sync(0xffffffffu, a, o);
    b += __shfl_down_sync(0xffffffffu, b, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[warp] = a; sh[ST_THREADS / 32 + warp] = b; }
  __syncthreads();
  if (warp == 0) {
    a = lane < ST_THREADS / 32 ? sh[lane] : 0.0;
    b = lane < ST_THREADS / 32 ? sh[ST_THREADS / 32 + lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_down_sync(0xffffffffu, a, o);
      b += __shfl_down_sync(0xffffffffu, b, o);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(ST_THREADS) value_stats_kernel(const T* __restrict__ v, long long n, double* __restrict__ partials) {
  __shared__ double sh[2 * ST_THREADS / 32];
  double s = 0.0, s2 = 0.0;
  const long long stride = (long long)gridDim.x * ST_THREADS;
  for (long long i = (long long)blockIdx.x * ST_THREADS + threadIdx.x; i < n; i += stride) {
    const double t = (double)v[i];
    s += t;
    s2 += t * t;
  }
  block_reduce2(s, s2, sh);
  if (threadIdx.x == 0) { partials[2 * blockIdx.x] = s; partials[2 * blockIdx.x + 1] = s2; }
}

// stats[0..1] += sum of the per-block partials, in block order (one thread: <= 592 additions)
__global__ void finish_stats_kernel(const double* __restrict__ partials, int nblocks, double* __restrict__ stats) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0, s2 = 0.0;
    for (int i = 0; i < nblocks; ++i) { s += partials[2 * i]; s2 += partials[2 * i + 1]; }
    stats[0] += s;
    stats[1] += s2;
  }
}

// per pixel (c, b, h, w): (sum_q x, sum_q x^2) -> ps[pixel] (double2)
template <typename T>
__global__ void __launch_bounds__(ST_THREADS) pixel_sums_kernel(const T* __restrict__ x, long long npix, int Q, double2* __restrict__ ps) {
  const long long stride = (long long)gridDim.x * ST_THREADS;
  for (long long i = (long long)blockIdx.x * ST_THREADS + threadIdx.x; i < npix; i += stride) {
    const T* px = x + i * Q;
    double s = 0.0, s2 = 0.0;
    for (int q = 0; q < Q; ++q) {
      const double t = (double)px[q];
      s += t;
      s2 += t * t;
    }
    ps[i] = make_double2(s, s2);
  }
}

// per window p: prod over the K*K*C pixels of the window of (s, s2); summed over all windows
__global__ void __launch_bounds__(ST_THREADS) window_products_kernel(const double2* __restrict__ ps, int C, int B, int H, int W, int K,
                                                                     double* __restrict__ partials) {
  __shared__ double sh[2 * ST_THREADS / 32];
  const int Ho = H - K + 1, Wo = W - K + 1;
  const long long P = (long long)B * Ho * Wo, chan = (long long)B * H * W;
  double s = 0.0, s2 = 0.0;
  const long long stride = (long long)gridDim.x * ST_THREADS;
  for (long long p = (long long)blockIdx.x * ST_THREADS + threadIdx.x; p < P; p += stride) {
    const long long b = p / (Ho * Wo);
    const int r = (int)(p - b * (Ho * Wo)), h = r / Wo, w = r - h * Wo;
    const long long org = (b * H + h) * (long long)W + w;
    double a = 1.0, a2 = 1.0;
    for (int c = 0; c < C; ++c)
      for (int dh = 0; dh < K; ++dh)
        for (int dw = 0; dw < K; ++dw) {
          const double2 t = ps[c * chan + org + (long long)dh * W + dw];
          a *= t.x;
          a2 *= t.y;
        }
    s += a;
    s2 += a2;
  }
  block_reduce2(s, s2, sh);
  if (threadIdx.x == 0) { partials[2 * blockIdx.x] = s; partials[2 * blockIdx.x + 1] = s2; }
}

inline int stats_blocks(long long n) {
  long long b = (n + ST_THREADS - 1) / ST_THREADS;
  if (b > ST_MAX_BLOCKS) b = ST_MAX_BLOCKS;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

size_t value_stats_workspace_bytes() { return (size_t)ST_MAX_BLOCKS * 2 * sizeof(double); }

template <typename T>
int launch_value_stats(const T* v, long long n, double* stats, void* ws, cudaStream_t st) {
  double* partials = (double*)ws;
  const int blocks = stats_blocks(n);
  value_stats_kernel<T><<<blocks, ST_THREADS, 0, st>>>(v, n, partials);
  finish_stats_kernel<<<1, 32, 0, st>>>(partials, blocks, stats);
  dctn_count_launch(2);
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}
template int launch_value_stats<float>(const float*, long long, double*, void*, cudaStream_t);
template int launch_value_stats<double>(const double*, long long, double*, void*, cudaStream_t);

size_t window_stats_workspace_bytes(int C, int B, int H, int W) {
  return value_stats_workspace_bytes() + (size_t)C * B * H * W * sizeof(double2);
}

template <typename T>
int launch_window_stats(const T* x, int C, int B, int H, int W, int Q, int K, double* stats, void* ws, cudaStream_t st) {
  double* partials = (double*)ws;
  double2* ps = (double2*)((char*)ws + value_stats_workspace_bytes());
  const long long npix = (long long)C * B * H * W;
  const long long P = (long long)B * (H - K + 1) * (W - K + 1);
  pixel_sums_kernel<T><<<stats_blocks(npix), ST_THREADS, 0, st>>>(x, npix, Q, ps);
  const int blocks = stats_blocks(P);
  window_products_kernel<<<blocks, ST_THREADS, 0, st>>>(ps, C, B, H, W, K, partials);
  finish_stats_kernel<<<1, 32, 0, st>>>(partials, blocks, stats);
  dctn_count_launch(3);
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}
template int launch_window_stats<float>(const float*, int, int, int, int, int, int, double*, void*, cudaStream_t);
template int launch_window_stats<double>(const double*, int, int, int, int, int, int, double*, void*, cudaStream_t);
