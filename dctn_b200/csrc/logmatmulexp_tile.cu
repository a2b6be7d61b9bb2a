// logmatmulexp for matrices whose inner dimension fits shared memory (the reference's use: chains of N x N matrices,
// N <= 300, small_experiments/logmatmulexp_benchmark/benchmark.py:21-52), ONE kernel per product:
//
//   out[t][i] = log sum_r exp(A[t][r] + B[r][i])                                             (dctn/logmatmulexp.py:5-14)
//             = m_t + n_i + log sum_r exp(A[t][r] - m_t) * exp(B[r][i] - n_i),   m_t = max_r A[t][r],  n_i = max_r B[r][i]
//
// The second form needs Theta*R + R*I exponentials and a plain fp32 (fp64) matrix product instead of Theta*R*I
// exponentials — at N = 256 that is 128 K instead of 16.8 M evaluations on the SFU pipe, which is what bounds the
// per-element form (csrc/logmatmulexp.cu).  It is NOT unconditionally stable: the dominant term of
// an output can be exp(-spread) below the row / column maxima and underflow (the scale-150 inputs of
// small_experiments/logmatmulexp_old.py:149-153).  Each CTA therefore checks its own operands: with
//   spread_A(t) = m_t - min_finite_r A[t][r],   spread_B(i) = n_i - min_finite_r B[r][i]
// the dominant product of every output of the tile is at least exp(-min(max_t spread_A, max_i spread_B)) (take r at the
// row maximum of A, or at the column maximum of B); below LIM = 60 (fp32) / 600 (fp64) both factors of the dominant
// product are normal numbers and every term that underflows is < e^-27 of it.  -inf entries (log 0) are exact zeros
// and do not count towards the spread.  A tile that fails the test (or holds +inf / NaN) takes the per-element
// max-shifted path — the same arithmetic as lme_fwd_kernel — from the operands it already has in shared memory, and
// raises a flag that makes the backward kernels do the same.  So the result is always the stable one; only its cost varies.
//
// Backward (nothing Theta*R*I-sized is kept; logmatmulexp_lowmem is the same function):
//   dA[t][r] = exp(A[t][r] - m_t) * sum_i G[t][i] * exp(B[r][i] - n_i),   G[t][i] = gout[t][i] * exp(m_t + n_i - out[t][i])
//   dB[r][i] = exp(B[r][i] - n_i) * sum_t exp(A[t][r] - m_t) * G[t][i]
// (per-element form exp(A + B - out) when the flag is up).  m and n are written by the forward kernel.
#include <cmath>

#include "../../include/dctn_b200.h"
#include "common.cuh"
#include "eps_kernels.h"

namespace {

constexpr int LT = 32;          // output tile LT x LT
constexpr int RT = 1;           // RT x RT outputs per thread.  One: a 256 x 256 product is only 64 tiles, so each tile gets 32 warps.
                                // Measured (profiles/r02g_cfg5_launches.txt): 256 threads x (2 x 2) forward 17.7 / backward 20.3 us,
                                // 1024 threads x (1 x 1) 17.5 / 17.7 us — the product loop is bound by shared-memory wavefronts
                                // (two loads per FMA), the rest is launch + one HBM round trip; a 4 x 4 register tile over
                                // k-major operands would cut the wavefronts 8x (not written)
constexpr int LTH = (LT / RT) * (LT / RT);
constexpr int NPART = LTH / 32; // threads per column in the column reductions
static_assert(LT == 32, "the staging loops map one lane to one tile column");
constexpr size_t LME_SMEM_LIMIT = 200 * 1024;

template <typename T> struct Lim;
template <> struct Lim<float> { static __device__ __forceinline__ float spread() { return 60.f; } };
template <> struct Lim<double> { static __device__ __forceinline__ double spread() { return 600.0; } };
__device__ __forceinline__ float xexp(float v) { return expf(v); }
__device__ __forceinline__ double xexp(double v) { return exp(v); }
__device__ __forceinline__ float xlog(float v) { return logf(v); }
__device__ __forceinline__ double xlog(double v) { return log(v); }
template <typename T> __device__ __forceinline__ T ninf() { return -(T)INFINITY; }
template <typename T> __device__ __forceinline__ bool finite_(T v) { return v - v == T(0); }
// One element global -> shared without a register round trip (LDGSTS): a CTA stages its whole operand tiles with every
// copy in flight at once and waits ONCE — the plain load/store loops paid one HBM latency per handful of elements
// (27 us for a 256 x 256 product on 64 CTAs; 17.8 us with the asynchronous copies, profiles/r02g_cfg5_launches.txt).
template <typename T>
__device__ __forceinline__ void cp_async_elem(T* dst_smem, const T* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  if (sizeof(T) == 4) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------- forward
template <typename T>
__global__ void __launch_bounds__(LTH) lme_tile_fwd_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ out,
                                                           T* __restrict__ rowmax, T* __restrict__ colmax, int* __restrict__ flag,
                                                           int Th, int R, int I) {
  extern __shared__ unsigned char lme_smem_raw[];
  T* sm = reinterpret_cast<T*>(lme_smem_raw);
  const int RP = R | 1;                      // odd row stride: rows of As are read at a stride by different threads
  T* As = sm;                                // [LT][RP]
  T* Bs = As + LT * RP;                      // [R][LT]
  T* mrow = Bs + (size_t)R * LT;             // [LT] row maxima, then [LT] column maxima
  T* ncol = mrow + LT;
  T* red = ncol + LT;                        // [8][LT] x 2 scratch of the column reduction
  __shared__ int s_unsafe, s_badA, s_badB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t0 = blockIdx.y * LT, i0 = blockIdx.x * LT;
  if (tid == 0) { s_unsafe = 0; s_badA = 0; s_badB = 0; }
  // a warp takes whole rows (A: LT/8 rows of R elements, B: R/8 rows of LT = 32 elements): coalesced, no index division
  for (int tl = warp; tl < LT; tl += LTH / 32) {
    const bool rowok = t0 + tl < Th;
    const T* src = A + (long long)(t0 + tl) * R;
    for (int r = lane; r < R; r += 32) {
      if (rowok) cp_async_elem(As + tl * RP + r, src + r);
      else As[tl * RP + r] = ninf<T>();
    }
  }
  {
    const bool colok = i0 + lane < I;
    for (int r = warp; r < R; r += LTH / 32) {
      if (colok) cp_async_elem(Bs + r * LT + lane, B + (long long)r * I + i0 + lane);
      else Bs[r * LT + lane] = ninf<T>();
    }
  }
  cp_async_wait_all();
  __syncthreads();
  // row maxima / finite minima of As: one warp per 4 rows
  for (int tl = warp; tl < LT; tl += LTH / 32) {
    T mx = ninf<T>(), mn = -ninf<T>();
    bool bad = false;
    for (int r = lane; r < R; r += 32) {
      const T v = As[tl * RP + r];
      if (v == ninf<T>()) continue;
      if (!finite_(v)) { bad = true; continue; }
      mx = v > mx ? v : mx;
      mn = v < mn ? v : mn;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const T a = __shfl_xor_sync(0xffffffffu, mx, o), b = __shfl_xor_sync(0xffffffffu, mn, o);
      mx = a > mx ? a : mx;
      mn = b < mn ? b : mn;
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
      mrow[tl] = mx;
      if (bad) s_unsafe = 1;
      else if (mx != ninf<T>() && mx - mn > Lim<T>::spread()) s_badA = 1;
    }
  }
  // column maxima / finite minima of Bs: 8 threads per column
  {
    const int c = tid & 31, part = tid >> 5;
    T mx = ninf<T>(), mn = -ninf<T>();
    bool bad = false;
    for (int r = part; r < R; r += LTH / 32) {
      const T v = Bs[r * LT + c];
      if (v == ninf<T>()) continue;
      if (!finite_(v)) { bad = true; continue; }
      mx = v > mx ? v : mx;
      mn = v < mn ? v : mn;
    }
    red[part * LT + c] = mx;
    red[(NPART + part) * LT + c] = mn;
    if (bad) s_unsafe = 1;
  }
  __syncthreads();
  if (tid < LT) {
    T mx = ninf<T>(), mn = -ninf<T>();
    for (int p = 0; p < NPART; ++p) {
      const T a = red[p * LT + tid], b = red[(NPART + p) * LT + tid];
      mx = a > mx ? a : mx;
      mn = b < mn ? b : mn;
    }
    ncol[tid] = mx;
    if (mx != ninf<T>() && mx - mn > Lim<T>::spread()) s_badB = 1;
  }
  __syncthreads();
  const bool exact = s_unsafe || (s_badA && s_badB);
  if (exact && tid == 0) atomicOr(flag, 1);
  if (blockIdx.x == 0 && tid < LT && t0 + tid < Th) rowmax[t0 + tid] = mrow[tid];
  if (blockIdx.y == 0 && tid < LT && i0 + tid < I) colmax[i0 + tid] = ncol[tid];
  const int ty = tid / (LT / RT), tx = tid % (LT / RT);   // outputs (RT ty + u, RT tx + v)
  T res[RT][RT];
  if (!exact) {
    for (int tl = warp; tl < LT; tl += LTH / 32) {
      const T m = mrow[tl];
      for (int r = lane; r < R; r += 32) As[tl * RP + r] = (m == ninf<T>()) ? T(0) : xexp(As[tl * RP + r] - m);
    }
    for (int idx = tid; idx < R * LT; idx += LTH) {
      const T n = ncol[idx & (LT - 1)];
      Bs[idx] = (n == ninf<T>()) ? T(0) : xexp(Bs[idx] - n);
    }
    __syncthreads();
    T acc[RT][RT] = {};
    const T* a0 = As + (RT * ty) * RP;
    const T* b = Bs + RT * tx;
#pragma unroll 8
    for (int r = 0; r < R; ++r) {
      T xv[RT], yv[RT];
#pragma unroll
      for (int u = 0; u < RT; ++u) { xv[u] = a0[u * RP + r]; yv[u] = b[r * LT + u]; }
#pragma unroll
      for (int u = 0; u < RT; ++u)
#pragma unroll
        for (int v = 0; v < RT; ++v) acc[u][v] += xv[u] * yv[v];
    }
#pragma unroll
    for (int u = 0; u < RT; ++u)
#pragma unroll
      for (int v = 0; v < RT; ++v) {
        const T m = mrow[RT * ty + u], n = ncol[RT * tx + v];
        res[u][v] = (m == ninf<T>() || n == ninf<T>() || acc[u][v] == T(0)) ? ninf<T>() : m + n + xlog(acc[u][v]);
      }
  } else {
    // per-element max-shifted logsumexp (the arithmetic of lme_fwd_kernel) from the raw operands in shared memory
#pragma unroll
    for (int u = 0; u < RT; ++u)
#pragma unroll
      for (int v = 0; v < RT; ++v) {
        const T* a = As + (RT * ty + u) * RP;
        const T* b = Bs + RT * tx + v;
        T m = ninf<T>();
        for (int r = 0; r < R; ++r) {
          const T w = a[r] + b[r * LT];
          m = w > m ? w : m;               // NaN never wins the comparison: handled below
        }
        T r0 = m;
        if (finite_(m)) {
          T s = T(0);
          for (int r = 0; r < R; ++r) s += xexp(a[r] + b[r * LT] - m);
          r0 = m + xlog(s);               // a NaN term makes s, hence the result, NaN — as torch.logsumexp does
        } else if (m == ninf<T>()) {
          for (int r = 0; r < R; ++r) {   // all terms -inf, or NaN among them
            const T w = a[r] + b[r * LT];
            if (w != w) r0 = w;
          }
        }
        res[u][v] = r0;
      }
  }
#pragma unroll
  for (int u = 0; u < RT; ++u)
#pragma unroll
    for (int v = 0; v < RT; ++v) {
      const int t = t0 + RT * ty + u, i = i0 + RT * tx + v;
      if (t < Th && i < I) out[(long long)t * I + i] = res[u][v];
    }
}

// ---------------------------------------------------------------------------------------------------------- backward
// WHICH = 0: dA tile (LT rows t x LT columns r), reduction over i;  WHICH = 1: dB tile (LT rows r x LT columns i), over t
template <typename T, int WHICH>
__global__ void __launch_bounds__(LTH) lme_tile_bwd_kernel(const T* __restrict__ A, const T* __restrict__ B, const T* __restrict__ out,
                                                           const T* __restrict__ gout, const T* __restrict__ rowmax,
                                                           const T* __restrict__ colmax, const int* __restrict__ flag,
                                                           T* __restrict__ dst, int Th, int R, int I) {
  extern __shared__ unsigned char lme_smem_raw[];
  T* sm = reinterpret_cast<T*>(lme_smem_raw);
  const int tid = threadIdx.x, ty = tid / (LT / RT), tx = tid % (LT / RT);
  const bool exact = *flag != 0;
  T acc[RT][RT] = {};
  if (WHICH == 0) {
    // rows: X[t][i] (t tile), Y[r][i] (r tile), K = i, both rows contiguous in K
    const int KP = I | 1;
    T* Gs = sm;                 // [LT][KP]  G (fast) or gout (exact)
    T* Os = Gs + LT * KP;       // [LT][KP]  out (exact only)
    T* Bs = Os + LT * KP;       // [LT][KP]  exp(B - n) (fast) or B (exact)
    const int t0 = blockIdx.y * LT, r0 = blockIdx.x * LT;
    const int warp = tid >> 5, lane = tid & 31;
    // raw gout / out / B rows staged with asynchronous copies (all in flight at once), then transformed in place
    for (int l = warp; l < LT; l += LTH / 32) {
      const int t = t0 + l, r = r0 + l;
      for (int i = lane; i < I; i += 32) {
        if (t < Th) {
          cp_async_elem(Gs + l * KP + i, gout + (long long)t * I + i);
          cp_async_elem(Os + l * KP + i, out + (long long)t * I + i);
        } else {
          Gs[l * KP + i] = T(0); Os[l * KP + i] = T(0);
        }
        if (r < R) cp_async_elem(Bs + l * KP + i, B + (long long)r * I + i);
        else Bs[l * KP + i] = ninf<T>();
      }
    }
    cp_async_wait_all();
    __syncthreads();
    if (!exact) {
      for (int l = warp; l < LT; l += LTH / 32) {
        const T m = (t0 + l < Th) ? rowmax[t0 + l] : ninf<T>();
        for (int i = lane; i < I; i += 32) {
          const T n = colmax[i];
          const T g = Gs[l * KP + i], o = Os[l * KP + i], b = Bs[l * KP + i];
          Gs[l * KP + i] = (g == T(0) || o == ninf<T>()) ? T(0) : g * xexp(m + n - o);
          Bs[l * KP + i] = (n == ninf<T>()) ? T(0) : xexp(b - n);
        }
      }
      __syncthreads();
    }
    if (!exact) {
      const T* g0 = Gs + (RT * ty) * KP;
      const T* b0 = Bs + (RT * tx) * KP;
#pragma unroll 8
      for (int i = 0; i < I; ++i) {
        T xv[RT], yv[RT];
#pragma unroll
        for (int u = 0; u < RT; ++u) { xv[u] = g0[u * KP + i]; yv[u] = b0[u * KP + i]; }
#pragma unroll
        for (int u = 0; u < RT; ++u)
#pragma unroll
          for (int v = 0; v < RT; ++v) acc[u][v] += xv[u] * yv[v];
      }
    }
#pragma unroll
    for (int u = 0; u < RT; ++u)
#pragma unroll
      for (int v = 0; v < RT; ++v) {
        const int t = t0 + RT * ty + u, r = r0 + RT * tx + v;
        if (t >= Th || r >= R) continue;
        const T a = A[(long long)t * R + r];
        T res;
        if (!exact) {
          const T m = rowmax[t];
          res = (m == ninf<T>()) ? T(0) : xexp(a - m) * acc[u][v];
        } else {
          const T* g = Gs + (RT * ty + u) * KP;
          const T* o = Os + (RT * ty + u) * KP;
          const T* b = Bs + (RT * tx + v) * KP;
          res = T(0);
          for (int i = 0; i < I; ++i) {
            const T gv = g[i];
            res += (gv != T(0)) ? gv * xexp(a + b[i] - o[i]) : T(0);
          }
        }
        dst[(long long)t * R + r] = res;
      }
  } else {
    // columns: X[t][r] (r tile), Y[t][i] (i tile), K = t
    T* As = sm;                       // [Th][LT]  exp(A - m) (fast) or A (exact)
    T* Gs = As + (size_t)Th * LT;     // [Th][LT]  G (fast) or gout (exact)
    T* Os = Gs + (size_t)Th * LT;     // [Th][LT]  out (exact only)
    const int r0 = blockIdx.y * LT, i0 = blockIdx.x * LT;
    const int warp = tid >> 5, lane = tid & 31;   // lane = column l of the tile (LT == 32), a warp takes rows t
    {
      const int r = r0 + lane, i = i0 + lane;
      for (int t = warp; t < Th; t += LTH / 32) {
        const int idx = t * LT + lane;
        if (r < R) cp_async_elem(As + idx, A + (long long)t * R + r);
        else As[idx] = ninf<T>();
        if (i < I) {
          cp_async_elem(Gs + idx, gout + (long long)t * I + i);
          cp_async_elem(Os + idx, out + (long long)t * I + i);
        } else {
          Gs[idx] = T(0); Os[idx] = T(0);
        }
      }
      cp_async_wait_all();
      __syncthreads();
      if (!exact) {
        const T n = (i < I) ? colmax[i] : ninf<T>();
        for (int t = warp; t < Th; t += LTH / 32) {
          const int idx = t * LT + lane;
          const T m = rowmax[t];
          const T a = As[idx], g = Gs[idx], o = Os[idx];
          As[idx] = (m == ninf<T>()) ? T(0) : xexp(a - m);
          Gs[idx] = (g == T(0) || o == ninf<T>()) ? T(0) : g * xexp(m + n - o);
        }
        __syncthreads();
      }
    }
    if (!exact) {
      const T* a = As + RT * ty;
      const T* g = Gs + RT * tx;
#pragma unroll 8
      for (int t = 0; t < Th; ++t) {
        T xv[RT], yv[RT];
#pragma unroll
        for (int u = 0; u < RT; ++u) { xv[u] = a[t * LT + u]; yv[u] = g[t * LT + u]; }
#pragma unroll
        for (int u = 0; u < RT; ++u)
#pragma unroll
          for (int v = 0; v < RT; ++v) acc[u][v] += xv[u] * yv[v];
      }
    }
#pragma unroll
    for (int u = 0; u < RT; ++u)
#pragma unroll
      for (int v = 0; v < RT; ++v) {
        const int r = r0 + RT * ty + u, i = i0 + RT * tx + v;
        if (r >= R || i >= I) continue;
        const T b = B[(long long)r * I + i];
        T res;
        if (!exact) {
          const T n = colmax[i];
          res = (n == ninf<T>()) ? T(0) : xexp(b - n) * acc[u][v];
        } else {
          res = T(0);
          for (int t = 0; t < Th; ++t) {
            const T gv = Gs[t * LT + RT * tx + v];
            res += (gv != T(0)) ? gv * xexp(As[t * LT + RT * ty + u] + b - Os[t * LT + RT * tx + v]) : T(0);
          }
        }
        dst[(long long)r * I + i] = res;
      }
  }
}

template <typename T> size_t fwd_smem(int R) { return ((size_t)LT * (R | 1) + (size_t)R * LT + 2 * LT + 2 * NPART * LT) * sizeof(T); }
template <typename T> size_t bwd_a_smem(int I) { return (size_t)3 * LT * (I | 1) * sizeof(T); }
template <typename T> size_t bwd_b_smem(int Th) { return (size_t)3 * Th * LT * sizeof(T); }

}  // namespace

// scratch kept from forward to backward: rowmax [Theta], colmax [I], flag
size_t lme_tile_workspace_bytes(int Th, int I, size_t es) { return ((size_t)Th + I) * es + 64; }

template <typename T>
bool lme_tile_supported(int Th, int R, int I) {
  return fwd_smem<T>(R) <= LME_SMEM_LIMIT && bwd_a_smem<T>(I) <= LME_SMEM_LIMIT && bwd_b_smem<T>(Th) <= LME_SMEM_LIMIT &&
         (long long)((Th + LT - 1) / LT) < 65536 && (long long)((R + LT - 1) / LT) < 65536;
}
template bool lme_tile_supported<float>(int, int, int);
template bool lme_tile_supported<double>(int, int, int);

template <typename T>
int lme_tile_forward(const T* A, const T* B, T* out, int Th, int R, int I, void* ws, cudaStream_t st) {
  T* rowmax = (T*)ws;
  T* colmax = rowmax + Th;
  int* flag = (int*)(colmax + I);
  DCTN_CUDA_CHECK_RET(cudaMemsetAsync(flag, 0, sizeof(int), st));
  const size_t smem = fwd_smem<T>(R);
  DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(lme_tile_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((I + LT - 1) / LT, (Th + LT - 1) / LT);
  lme_tile_fwd_kernel<T><<<grid, LTH, smem, st>>>(A, B, out, rowmax, colmax, flag, Th, R, I);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}
template int lme_tile_forward<float>(const float*, const float*, float*, int, int, int, void*, cudaStream_t);
template int lme_tile_forward<double>(const double*, const double*, double*, int, int, int, void*, cudaStream_t);

template <typename T>
int lme_tile_backward(const T* A, const T* B, const T* out, const T* gout, T* dA, T* dB, int Th, int R, int I, const void* ws,
                      cudaStream_t st) {
  const T* rowmax = (const T*)ws;
  const T* colmax = rowmax + Th;
  const int* flag = (const int*)(colmax + I);
  if (dA) {
    const size_t smem = bwd_a_smem<T>(I);
    auto k = lme_tile_bwd_kernel<T, 0>;
    DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((R + LT - 1) / LT, (Th + LT - 1) / LT);
    k<<<grid, LTH, smem, st>>>(A, B, out, gout, rowmax, colmax, flag, dA, Th, R, I);
    dctn_count_launch();
    DCTN_CUDA_CHECK_RET(cudaGetLastError());
  }
  if (dB) {
    const size_t smem = bwd_b_smem<T>(Th);
    auto k = lme_tile_bwd_kernel<T, 1>;
    DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((I + LT - 1) / LT, (R + LT - 1) / LT);
    k<<<grid, LTH, smem, st>>>(A, B, out, gout, rowmax, colmax, flag, dB, Th, R, I);
    dctn_count_launch();
    DCTN_CUDA_CHECK_RET(cudaGetLastError());
  }
  return 0;
}
template int lme_tile_backward<float>(const float*, const float*, const float*, const float*, float*, float*, int, int, int, const void*, cudaStream_t);
template int lme_tile_backward<double>(const double*, const double*, const double*, const double*, double*, double*, int, int, int, const void*, cudaStream_t);
