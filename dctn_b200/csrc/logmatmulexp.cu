// logmatmulexp: out[t][i] = log sum_r exp(A[t][r] + B[r][i])   (dctn/logmatmulexp.py:5-14)
//
// Numerically stable, max-shifted per OUTPUT ELEMENT (the shift is max_r(A[t][r] + B[r][i]), exactly what
// torch.logsumexp does on the reference's materialised (Theta,R,I) tensor) — a row/column pre-scaling
// "exp(A) @ exp(B)" GEMM would underflow on the scale-150 inputs of small_experiments/logmatmulexp_old.py:149-153.
// Nothing Theta*R*I-sized is ever stored: forward keeps a running (max, sum) per output in registers
// (online logsumexp), backward recomputes the softmax weights exp(A+B-out) from the saved output.
// Bound: Theta*R*I exponentials on the SFU (MUFU.EX2) pipe, not HBM.
#include "common.cuh"
#include "eps_kernels.h"

namespace {
constexpr int TS = 16;  // 16x16 output tile per CTA, r staged in chunks of 16

template <typename T> __device__ __forceinline__ T neg_inf();
template <> __device__ __forceinline__ float neg_inf<float>() { return -INFINITY; }
template <> __device__ __forceinline__ double neg_inf<double>() { return -(double)INFINITY; }
__device__ __forceinline__ float fexp(float v) { return expf(v); }
__device__ __forceinline__ double fexp(double v) { return exp(v); }
__device__ __forceinline__ float flog(float v) { return logf(v); }
__device__ __forceinline__ double flog(double v) { return log(v); }

template <typename T>
__global__ void __launch_bounds__(TS * TS) lme_fwd_kernel(const T* __restrict__ A, const T* __restrict__ B,
                                                          T* __restrict__ out, int Th, int R, int I) {
  __shared__ T As[TS][TS + 1];  // [t][r]
  __shared__ T Bs[TS][TS + 1];  // [r][i]
  const int tx = threadIdx.x % TS, ty = threadIdx.x / TS;
  const int t = blockIdx.y * TS + ty, i = blockIdx.x * TS + tx;
  T m = neg_inf<T>(), s = T(0);
  for (int r0 = 0; r0 < R; r0 += TS) {
    int ra = r0 + tx, rb = r0 + ty;
    As[ty][tx] = (t < Th && ra < R) ? A[(long long)t * R + ra] : neg_inf<T>();
    Bs[ty][tx] = (rb < R && i < I) ? B[(long long)rb * I + i] : neg_inf<T>();
    __syncthreads();
    // chunk max first (adds only), then one exp per term: 16 exps + at most 1 rescale per chunk
    T v[TS];
    T cm = neg_inf<T>();
#pragma unroll
    for (int r = 0; r < TS; ++r) {
      v[r] = As[ty][r] + Bs[r][tx];
      cm = v[r] > cm ? v[r] : cm;
    }
    if (cm > m) {
      // rescale the running sum to the new max (m == -inf -> s == 0, factor irrelevant)
      s = (m == neg_inf<T>()) ? T(0) : s * fexp(m - cm);
      m = cm;
    }
    if (m != neg_inf<T>() && m != -neg_inf<T>()) {
#pragma unroll
      for (int r = 0; r < TS; ++r) s += fexp(v[r] - m);
    } else if (m == -neg_inf<T>()) {
      s = T(1);  // +inf term dominates: result is +inf (log(1) + inf)
    }
    __syncthreads();
  }
  if (t < Th && i < I) out[(long long)t * I + i] = (m == neg_inf<T>()) ? m : m + flog(s);
}

// dA[t][r] = sum_i gout[t][i] * exp(A[t][r] + B[r][i] - out[t][i])
template <typename T>
__global__ void __launch_bounds__(TS * TS) lme_bwd_a_kernel(const T* __restrict__ A, const T* __restrict__ B,
                                                            const T* __restrict__ out, const T* __restrict__ gout,
                                                            T* __restrict__ dA, int Th, int R, int I) {
  __shared__ T Os[TS][TS + 1];  // out[t][i]
  __shared__ T Gs[TS][TS + 1];  // gout[t][i]
  __shared__ T Bs[TS][TS + 1];  // B[r][i]
  const int tx = threadIdx.x % TS, ty = threadIdx.x / TS;
  const int t = blockIdx.y * TS + ty, r = blockIdx.x * TS + tx;
  const T a = (t < Th && r < R) ? A[(long long)t * R + r] : T(0);
  T acc = T(0);
  for (int i0 = 0; i0 < I; i0 += TS) {
    int ii = i0 + tx;
    bool ok = (t < Th && ii < I);
    Os[ty][tx] = ok ? out[(long long)t * I + ii] : T(0);
    Gs[ty][tx] = ok ? gout[(long long)t * I + ii] : T(0);
    int rb = blockIdx.x * TS + ty;
    Bs[ty][tx] = (rb < R && ii < I) ? B[(long long)rb * I + ii] : neg_inf<T>();
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TS; ++k) {
      T g = Gs[ty][k];
      T w = fexp(a + Bs[tx][k] - Os[ty][k]);
      acc += (g != T(0)) ? g * w : T(0);
    }
    __syncthreads();
  }
  if (t < Th && r < R) dA[(long long)t * R + r] = acc;
}

// dB[r][i] = sum_t gout[t][i] * exp(A[t][r] + B[r][i] - out[t][i])
template <typename T>
__global__ void __launch_bounds__(TS * TS) lme_bwd_b_kernel(const T* __restrict__ A, const T* __restrict__ B,
                                                            const T* __restrict__ out, const T* __restrict__ gout,
                                                            T* __restrict__ dB, int Th, int R, int I) {
  __shared__ T Os[TS][TS + 1];  // out[t][i]
  __shared__ T Gs[TS][TS + 1];  // gout[t][i]
  __shared__ T As[TS][TS + 1];  // A[t][r]
  const int tx = threadIdx.x % TS, ty = threadIdx.x / TS;
  const int r = blockIdx.y * TS + ty, i = blockIdx.x * TS + tx;
  const T b = (r < R && i < I) ? B[(long long)r * I + i] : T(0);
  T acc = T(0);
  for (int t0 = 0; t0 < Th; t0 += TS) {
    int tt = t0 + ty;
    bool ok = (tt < Th && i < I);
    Os[ty][tx] = ok ? out[(long long)tt * I + i] : T(0);
    Gs[ty][tx] = ok ? gout[(long long)tt * I + i] : T(0);
    int ra = blockIdx.y * TS + tx;
    As[ty][tx] = (tt < Th && ra < R) ? A[(long long)tt * R + ra] : neg_inf<T>();
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TS; ++k) {
      T g = Gs[k][tx];
      T w = fexp(As[k][ty] + b - Os[k][tx]);
      acc += (g != T(0)) ? g * w : T(0);
    }
    __syncthreads();
  }
  if (r < R && i < I) dB[(long long)r * I + i] = acc;
}
}  // namespace

template <typename T>
int lme_forward(const T* A, const T* B, T* out, int Th, int R, int I, cudaStream_t st) {
  dim3 grid((I + TS - 1) / TS, (Th + TS - 1) / TS);
  lme_fwd_kernel<T><<<grid, TS * TS, 0, st>>>(A, B, out, Th, R, I);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}

template <typename T>
int lme_backward(const T* A, const T* B, const T* out, const T* gout, T* dA, T* dB, int Th, int R, int I,
                 cudaStream_t st) {
  if (dA) {
    dim3 grid((R + TS - 1) / TS, (Th + TS - 1) / TS);
    lme_bwd_a_kernel<T><<<grid, TS * TS, 0, st>>>(A, B, out, gout, dA, Th, R, I);
    dctn_count_launch();
    DCTN_CUDA_CHECK_RET(cudaGetLastError());
  }
  if (dB) {
    dim3 grid((I + TS - 1) / TS, (R + TS - 1) / TS);
    lme_bwd_b_kernel<T><<<grid, TS * TS, 0, st>>>(A, B, out, gout, dB, Th, R, I);
    dctn_count_launch();
    DCTN_CUDA_CHECK_RET(cudaGetLastError());
  }
  return 0;
}

template int lme_forward<float>(const float*, const float*, float*, int, int, int, cudaStream_t);
template int lme_forward<double>(const double*, const double*, double*, int, int, int, cudaStream_t);
template int lme_backward<float>(const float*, const float*, const float*, const float*, float*, float*, int, int, int, cudaStream_t);
template int lme_backward<double>(const double*, const double*, const double*, const double*, double*, double*, int, int, int, cudaStream_t);
