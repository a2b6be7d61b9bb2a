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
#include "../../include/dctn_b200.h"

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
  // linear grid (Theta can exceed the 65535-block limit of grid.y: one row per ConvSBS window)
  const unsigned nbx = (unsigned)(I + TS - 1) / TS, bx = blockIdx.x % nbx, by = blockIdx.x / nbx;
  const int t = by * TS + ty, i = bx * TS + tx;
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
    const int rc = R - r0;  // valid terms of this chunk (the padding is -inf: exp = 0, not worth an SFU op)
    if (m != neg_inf<T>() && m != -neg_inf<T>()) {
#pragma unroll
      for (int r = 0; r < TS; ++r)
        if (r < rc) s += fexp(v[r] - m);
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
  const unsigned nbx = (unsigned)(R + TS - 1) / TS, bx = blockIdx.x % nbx, by = blockIdx.x / nbx;
  const int t = by * TS + ty, r = bx * TS + tx;
  const T a = (t < Th && r < R) ? A[(long long)t * R + r] : T(0);
  T acc = T(0);
  for (int i0 = 0; i0 < I; i0 += TS) {
    int ii = i0 + tx;
    bool ok = (t < Th && ii < I);
    Os[ty][tx] = ok ? out[(long long)t * I + ii] : T(0);
    Gs[ty][tx] = ok ? gout[(long long)t * I + ii] : T(0);
    int rb = bx * TS + ty;
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
  const unsigned nbx = (unsigned)(I + TS - 1) / TS, bx = blockIdx.x % nbx, by = blockIdx.x / nbx;
  const int r = by * TS + ty, i = bx * TS + tx;
  const T b = (r < R && i < I) ? B[(long long)r * I + i] : T(0);
  T acc = T(0);
  for (int t0 = 0; t0 < Th; t0 += TS) {
    int tt = t0 + ty;
    bool ok = (tt < Th && i < I);
    Os[ty][tx] = ok ? out[(long long)tt * I + i] : T(0);
    Gs[ty][tx] = ok ? gout[(long long)tt * I + i] : T(0);
    int ra = by * TS + tx;
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

// ------------------------------------------------------------------------------------------------ batched, small matrices
// out[p][t][i] = log sum_r exp(A[p][t][r] + B[p][r][i]),  p < NB: one small product per batch element — the bond
// matrices of a ConvSBS ring (dctn/conv_sbs.py:258-304 contracts the same ring in linear space), one element per
// (image, window).  A CTA takes G consecutive batch elements: their A and B blocks are contiguous in HBM, so they are
// staged into shared memory with 128-bit loads, one thread computes one output from there, and the G*Th*I outputs of
// the group are again contiguous (coalesced store).  Bound: NB*Th*R*I exponentials (MUFU); HBM traffic is the
// algorithmic NB*(Th*R + R*I + Th*I) elements.
template <typename T>
__device__ __forceinline__ void stage_in(T* __restrict__ dst, const T* __restrict__ src, int n) {
  constexpr int V = 16 / sizeof(T);
  if ((n % V) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    const int4* s4 = reinterpret_cast<const int4*>(src);
    int4* d4 = reinterpret_cast<int4*>(dst);
    for (int k = threadIdx.x; k < n / V; k += blockDim.x) d4[k] = __ldg(s4 + k);
  } else {
    for (int k = threadIdx.x; k < n; k += blockDim.x) dst[k] = __ldg(src + k);
  }
}

// idx = (gi * NA + a) * NB + b walked in steps of blockDim.x without divisions (the decomposition of a thread's first
// index does not depend on the group, so it is computed once per kernel; two integer divisions per work item were a
// quarter of the instructions of the register-tiled forward).
struct Walk3 {
  int gi, a, b, sg, sa, sb, NA, NB;
  __device__ __forceinline__ void init(int idx, int step, int na, int nb) {
    NA = na; NB = nb;
    b = idx % nb; int q = idx / nb; a = q % na; gi = q / na;
    sb = step % nb; q = step / nb; sa = q % na; sg = q / na;
  }
  __device__ __forceinline__ void next() {
    b += sb; if (b >= NB) { b -= NB; ++a; }
    a += sa; if (a >= NA) { a -= NA; ++gi; }
    gi += sg;
  }
};

template <typename T>
__global__ void __launch_bounds__(256) lme_batched_fwd_kernel(const T* __restrict__ A, const T* __restrict__ B,
                                                              T* __restrict__ out, long long NB, int Th, int R, int I,
                                                              int G) {
  extern __shared__ int4 lme_smem4[];
  T* As = reinterpret_cast<T*>(lme_smem4);  // [G][Th][R]
  T* Bs = As + (size_t)G * Th * R;          // [G][R][I]   (G*Th*R is kept a multiple of 4 by the host)
  const int tr = Th * R, ri = R * I, ti = Th * I;
  Walk3 w0;
  w0.init(threadIdx.x, blockDim.x, Th, I);
  for (long long p0 = (long long)blockIdx.x * G; p0 < NB; p0 += (long long)gridDim.x * G) {
    const int g = (int)((NB - p0) < G ? (NB - p0) : G);
    stage_in(As, A + p0 * tr, g * tr);
    stage_in(Bs, B + p0 * ri, g * ri);
    __syncthreads();
    Walk3 w = w0;
    for (int idx = threadIdx.x; idx < g * ti; idx += blockDim.x, w.next()) {
      const int gi = w.gi, t = w.a, i = w.b;
      const T* a = As + gi * tr + t * R;
      const T* b = Bs + gi * ri + i;
      T m = neg_inf<T>();
      for (int r = 0; r < R; ++r) {
        T v = a[r] + b[r * I];
        m = v > m ? v : m;
      }
      T res = m;  // -inf (empty sum) and +inf propagate as they are
      if (m != neg_inf<T>() && m != -neg_inf<T>()) {
        T s = T(0);
        for (int r = 0; r < R; ++r) s += fexp(a[r] + b[r * I] - m);
        res = m + flog(s);
      }
      out[p0 * ti + idx] = res;
    }
    __syncthreads();
  }
}

// dA[p][t][r] = sum_i gout[p][t][i] * exp(A[p][t][r] + B[p][r][i] - out[p][t][i]);  dB likewise, summed over t.
template <typename T>
__global__ void __launch_bounds__(256) lme_batched_bwd_kernel(const T* __restrict__ A, const T* __restrict__ B,
                                                              const T* __restrict__ out, const T* __restrict__ gout,
                                                              T* __restrict__ dA, T* __restrict__ dB, long long NB,
                                                              int Th, int R, int I, int G) {
  extern __shared__ int4 lme_smem4[];
  const int tr = Th * R, ri = R * I, ti = Th * I;
  T* As = reinterpret_cast<T*>(lme_smem4);  // every block length G*x is a multiple of 4 elements (host)
  T* Bs = As + (size_t)G * tr;
  T* Os = Bs + (size_t)G * ri;
  T* Gs = Os + (size_t)G * ti;
  Walk3 wa0, wb0;
  wa0.init(threadIdx.x, blockDim.x, Th, R);
  wb0.init(threadIdx.x, blockDim.x, R, I);
  for (long long p0 = (long long)blockIdx.x * G; p0 < NB; p0 += (long long)gridDim.x * G) {
    const int g = (int)((NB - p0) < G ? (NB - p0) : G);
    stage_in(As, A + p0 * tr, g * tr);
    stage_in(Bs, B + p0 * ri, g * ri);
    stage_in(Os, out + p0 * ti, g * ti);
    stage_in(Gs, gout + p0 * ti, g * ti);
    __syncthreads();
    if (dA) {
      Walk3 w = wa0;
      for (int idx = threadIdx.x; idx < g * tr; idx += blockDim.x, w.next()) {
        const int gi = w.gi, t = w.a, r = w.b;
        const T a = As[idx];
        const T* b = Bs + gi * ri + r * I;
        const T* o = Os + gi * ti + t * I;
        const T* gg = Gs + gi * ti + t * I;
        T acc = T(0);
        for (int i = 0; i < I; ++i) {
          T gv = gg[i];
          if (gv != T(0)) acc += gv * fexp(a + b[i] - o[i]);
        }
        dA[p0 * tr + idx] = acc;
      }
    }
    if (dB) {
      Walk3 w = wb0;
      for (int idx = threadIdx.x; idx < g * ri; idx += blockDim.x, w.next()) {
        const int gi = w.gi, r = w.a, i = w.b;
        const T b = Bs[idx];
        const T* a = As + gi * tr + r;
        const T* o = Os + gi * ti + i;
        const T* gg = Gs + gi * ti + i;
        T acc = T(0);
        for (int t = 0; t < Th; ++t) {
          T gv = gg[t * I];
          if (gv != T(0)) acc += gv * fexp(a[t * R] + b - o[t * I]);
        }
        dB[p0 * ri + idx] = acc;
      }
    }
    __syncthreads();
  }
}

// ---- float32, I % 4 == 0: register-tiled variants.  One thread = one row t and four consecutive columns i: the A row is
// read once for four outputs, B comes in as 128-bit shared-memory loads, the R sums a+b stay in registers between the max
// pass and the exp pass (R is a template parameter), and exp is one MUFU.EX2 on (v - m) * log2(e).  The difference v - m
// is formed BEFORE the scaling so that the dominant term is exactly 2^0 (backward recomputes exp(a + b - out) and a
// forward that is off by one rounding of m*log2(e) would show up there on the scale-150 inputs).
__device__ __forceinline__ float ex2_approx(float v) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float lg2_approx(float v) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

template <int R>
__global__ void __launch_bounds__(256, (R <= 8 ? 5 : (R <= 12 ? 4 : 2))) lme_batched_fwd_vec_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                                  float* __restrict__ out, long long NB, int Th, int I,
                                                                  int G) {
  extern __shared__ int4 lme_smem4[];
  float* As = reinterpret_cast<float*>(lme_smem4);  // [G][Th][R]
  float* Bs = As + (size_t)G * Th * R;              // [G][R][I]
  const int tr = Th * R, ri = R * I, ti = Th * I, I4 = I >> 2, w_per = Th * I4;
  Walk3 w0;
  w0.init(threadIdx.x, blockDim.x, Th, I4);
  for (long long p0 = (long long)blockIdx.x * G; p0 < NB; p0 += (long long)gridDim.x * G) {
    const int g = (int)((NB - p0) < G ? (NB - p0) : G);
    stage_in(As, A + p0 * tr, g * tr);
    stage_in(Bs, B + p0 * ri, g * ri);
    __syncthreads();
    Walk3 w = w0;
    for (int idx = threadIdx.x; idx < g * w_per; idx += blockDim.x) {
      int gi, t, i4;
      if constexpr (R <= 8) {   // division-free walk; for larger R its registers cost more than the divisions
        gi = w.gi; t = w.a; i4 = w.b;
        w.next();
      } else {
        gi = idx / w_per;
        const int rem = idx - gi * w_per;
        t = rem / I4; i4 = rem - t * I4;
      }
      const float* a = As + gi * tr + t * R;
      const float4* b = reinterpret_cast<const float4*>(Bs + gi * ri) + i4;
      float4 v[R];
      float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float ar = a[r];
        const float4 br = b[r * I4];
        v[r] = make_float4(ar + br.x, ar + br.y, ar + br.z, ar + br.w);
        m.x = fmaxf(m.x, v[r].x); m.y = fmaxf(m.y, v[r].y); m.z = fmaxf(m.z, v[r].z); m.w = fmaxf(m.w, v[r].w);
      }
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        s.x += ex2_approx((v[r].x - m.x) * kLog2e); s.y += ex2_approx((v[r].y - m.y) * kLog2e);
        s.z += ex2_approx((v[r].z - m.z) * kLog2e); s.w += ex2_approx((v[r].w - m.w) * kLog2e);
      }
      // m = -inf (empty sum) or +inf: v - m is NaN, s is NaN; the result is m itself
      float4 res;
      // s is in [1, R]: lg2.approx is good to 2^-22 absolute there
      res.x = (fabsf(m.x) == INFINITY) ? m.x : fmaf(lg2_approx(s.x), kLn2, m.x);
      res.y = (fabsf(m.y) == INFINITY) ? m.y : fmaf(lg2_approx(s.y), kLn2, m.y);
      res.z = (fabsf(m.z) == INFINITY) ? m.z : fmaf(lg2_approx(s.z), kLn2, m.z);
      res.w = (fabsf(m.w) == INFINITY) ? m.w : fmaf(lg2_approx(s.w), kLn2, m.w);
      reinterpret_cast<float4*>(out + p0 * ti)[idx] = res;  // idx enumerates (gi, t, i4) in memory order
    }
    __syncthreads();
  }
}

// Backward, float32, I % 4 == 0.  dA: one thread per (t, r), 128-bit loads of the out / gout / B rows; dB: one thread per
// (r, four columns), walks the rows t.  Same exactness rule: a + b - out first, then the scaling.
__global__ void __launch_bounds__(256, 6) lme_batched_bwd_vec_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                                  const float* __restrict__ out,
                                                                  const float* __restrict__ gout, float* __restrict__ dA,
                                                                  float* __restrict__ dB, long long NB, int Th, int R,
                                                                  int I, int G) {
  extern __shared__ int4 lme_smem4[];
  const int tr = Th * R, ri = R * I, ti = Th * I, I4 = I >> 2;
  float* As = reinterpret_cast<float*>(lme_smem4);
  float* Bs = As + (size_t)G * tr;
  float* Os = Bs + (size_t)G * ri;
  float* Gs = Os + (size_t)G * ti;
  Walk3 wa0, wb0;
  wa0.init(threadIdx.x, blockDim.x, Th, R);
  wb0.init(threadIdx.x, blockDim.x, R, I4);
  for (long long p0 = (long long)blockIdx.x * G; p0 < NB; p0 += (long long)gridDim.x * G) {
    const int g = (int)((NB - p0) < G ? (NB - p0) : G);
    stage_in(As, A + p0 * tr, g * tr);
    stage_in(Bs, B + p0 * ri, g * ri);
    stage_in(Os, out + p0 * ti, g * ti);
    stage_in(Gs, gout + p0 * ti, g * ti);
    __syncthreads();
    if (dA) {
      Walk3 w = wa0;
      for (int idx = threadIdx.x; idx < g * tr; idx += blockDim.x, w.next()) {
        const int gi = w.gi, t = w.a, r = w.b;
        const float a = As[idx];
        const float4* b = reinterpret_cast<const float4*>(Bs + gi * ri + r * I);
        const float4* o = reinterpret_cast<const float4*>(Os + gi * ti + t * I);
        const float4* gg = reinterpret_cast<const float4*>(Gs + gi * ti + t * I);
        float acc = 0.f;
#pragma unroll 2
        for (int i = 0; i < I4; ++i) {
          const float4 bv = b[i], ov = o[i], gv = gg[i];
          if (gv.x != 0.f) acc += gv.x * ex2_approx(((a + bv.x) - ov.x) * kLog2e);
          if (gv.y != 0.f) acc += gv.y * ex2_approx(((a + bv.y) - ov.y) * kLog2e);
          if (gv.z != 0.f) acc += gv.z * ex2_approx(((a + bv.z) - ov.z) * kLog2e);
          if (gv.w != 0.f) acc += gv.w * ex2_approx(((a + bv.w) - ov.w) * kLog2e);
        }
        dA[p0 * tr + idx] = acc;
      }
    }
    if (dB) {
      const int w_per = R * I4;
      Walk3 w = wb0;
      for (int idx = threadIdx.x; idx < g * w_per; idx += blockDim.x, w.next()) {
        const int gi = w.gi, r = w.a, i4 = w.b, rem = r * I4 + i4;
        const float4 bv = reinterpret_cast<const float4*>(Bs + gi * ri)[rem];
        const float* a = As + gi * tr + r;
        const float4* o = reinterpret_cast<const float4*>(Os + gi * ti) + i4;
        const float4* gg = reinterpret_cast<const float4*>(Gs + gi * ti) + i4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
        for (int t = 0; t < Th; ++t) {
          const float av = a[t * R];
          const float4 ov = o[t * I4], gv = gg[t * I4];
          if (gv.x != 0.f) acc.x += gv.x * ex2_approx(((av + bv.x) - ov.x) * kLog2e);
          if (gv.y != 0.f) acc.y += gv.y * ex2_approx(((av + bv.y) - ov.y) * kLog2e);
          if (gv.z != 0.f) acc.z += gv.z * ex2_approx(((av + bv.z) - ov.z) * kLog2e);
          if (gv.w != 0.f) acc.w += gv.w * ex2_approx(((av + bv.w) - ov.w) * kLog2e);
        }
        reinterpret_cast<float4*>(dB + p0 * ri)[idx] = acc;
      }
    }
    __syncthreads();
  }
}
}  // namespace

template <typename T>
int lme_forward(const T* A, const T* B, T* out, int Th, int R, int I, cudaStream_t st) {
  const long long nblk = (long long)((I + TS - 1) / TS) * ((Th + TS - 1) / TS);
  if (nblk >= (1ll << 31)) return dctn_set_error(DCTN_ERR_UNSUPPORTED, "logmatmulexp: %lld output tiles exceed the grid limit", nblk);
  lme_fwd_kernel<T><<<(unsigned)nblk, TS * TS, 0, st>>>(A, B, out, Th, R, I);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}

template <typename T>
int lme_backward(const T* A, const T* B, const T* out, const T* gout, T* dA, T* dB, int Th, int R, int I,
                 cudaStream_t st) {
  if (dA) {
    const long long nblk = (long long)((R + TS - 1) / TS) * ((Th + TS - 1) / TS);
    if (nblk >= (1ll << 31)) return dctn_set_error(DCTN_ERR_UNSUPPORTED, "logmatmulexp backward: %lld tiles exceed the grid limit", nblk);
    lme_bwd_a_kernel<T><<<(unsigned)nblk, TS * TS, 0, st>>>(A, B, out, gout, dA, Th, R, I);
    dctn_count_launch();
    DCTN_CUDA_CHECK_RET(cudaGetLastError());
  }
  if (dB) {
    const long long nblk = (long long)((I + TS - 1) / TS) * ((R + TS - 1) / TS);
    lme_bwd_b_kernel<T><<<(unsigned)nblk, TS * TS, 0, st>>>(A, B, out, gout, dB, Th, R, I);
    dctn_count_launch();
    DCTN_CUDA_CHECK_RET(cudaGetLastError());
  }
  return 0;
}


// Group size: enough batch elements per CTA for ~512 outputs, within the shared-memory budget; G*x multiples of 4.
static int lme_batched_group(size_t per_elem_bytes, int work_per_elem, long long NB, size_t* smem) {
  const size_t budget = 96 * 1024;
  if (per_elem_bytes * 4 > budget) {
    if (per_elem_bytes > budget) return 0;
    *smem = per_elem_bytes;  // a single element per CTA: every block starts at the tensor base + multiple of its size
    return 1;
  }
  int G = (512 + work_per_elem - 1) / work_per_elem;
  G = (G + 3) & ~3;
  while ((size_t)G * per_elem_bytes > budget) G -= 4;
  if (G < 4) G = 4;
  if ((long long)G > NB) G = (int)((NB + 3) & ~3LL);
  *smem = (size_t)G * per_elem_bytes;
  return G;
}

template <int R>
static int launch_fwd_vec(const float* A, const float* B, float* out, long long NB, int Th, int I, int G, size_t smem,
                          cudaStream_t st) {
  // per device/context attribute: set on every launch (a process-wide flag would cover only the first GPU used)
  DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(lme_batched_fwd_vec_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  long long groups = (NB + G - 1) / G;
  int grid = (int)(groups < 148LL * 16 ? groups : 148LL * 16);
  lme_batched_fwd_vec_kernel<R><<<grid, 256, smem, st>>>(A, B, out, NB, Th, I, G);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}

// float32, I % 4 == 0, R <= 16, groups of 4 elements fit: the register-tiled kernels
static bool lme_vec_ok(size_t elem_size, int R, int I, int G) { return elem_size == 4 && (I & 3) == 0 && R <= 16 && (G & 3) == 0; }

template <typename T>
int lme_batched_forward(const T* A, const T* B, T* out, long long NB, int Th, int R, int I, cudaStream_t st) {
  size_t smem = 0;
  int G = lme_batched_group(sizeof(T) * ((size_t)Th * R + (size_t)R * I), (sizeof(T) == 4 && (I & 3) == 0 && R <= 16) ? Th * I / 4 : Th * I, NB, &smem);
  if (G && lme_vec_ok(sizeof(T), R, I, G)) {
    const float *a = (const float*)A, *b = (const float*)B;
    float* o = (float*)out;
    switch (R) {
#define LME_CASE(RR) case RR: return launch_fwd_vec<RR>(a, b, o, NB, Th, I, G, smem, st);
      LME_CASE(1) LME_CASE(2) LME_CASE(3) LME_CASE(4) LME_CASE(5) LME_CASE(6) LME_CASE(7) LME_CASE(8)
      LME_CASE(9) LME_CASE(10) LME_CASE(11) LME_CASE(12) LME_CASE(13) LME_CASE(14) LME_CASE(15) LME_CASE(16)
#undef LME_CASE
    }
  }
  if (G == 0)
    return dctn_set_error(DCTN_ERR_UNSUPPORTED, "logmatmulexp_batched: a (%d x %d) x (%d x %d) pair does not fit shared memory; use the 2-D entry per element", Th, R, R, I);
  // G == 1 with odd block sizes: stage_in falls back to scalar loads on its own (alignment test)
  DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(lme_batched_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  long long groups = (NB + G - 1) / G;
  int grid = (int)(groups < 148LL * 16 ? groups : 148LL * 16);
  lme_batched_fwd_kernel<T><<<grid, 256, smem, st>>>(A, B, out, NB, Th, R, I, G);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}

template <typename T>
int lme_batched_backward(const T* A, const T* B, const T* out, const T* gout, T* dA, T* dB, long long NB, int Th, int R,
                         int I, cudaStream_t st) {
  size_t smem = 0;
  int work = Th * R > R * I ? Th * R : R * I;
  int G = lme_batched_group(sizeof(T) * ((size_t)Th * R + (size_t)R * I + 2 * (size_t)Th * I), work, NB, &smem);
  if (G == 0)
    return dctn_set_error(DCTN_ERR_UNSUPPORTED, "logmatmulexp_batched backward: a (%d x %d) x (%d x %d) pair does not fit shared memory", Th, R, R, I);
  DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(lme_batched_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  long long groups = (NB + G - 1) / G;
  int grid = (int)(groups < 148LL * 16 ? groups : 148LL * 16);
  if (lme_vec_ok(sizeof(T), 1, I, G)) {
    DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(lme_batched_bwd_vec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    lme_batched_bwd_vec_kernel<<<grid, 256, smem, st>>>((const float*)A, (const float*)B, (const float*)out, (const float*)gout,
                                                        (float*)dA, (float*)dB, NB, Th, R, I, G);
    dctn_count_launch();
    DCTN_CUDA_CHECK_RET(cudaGetLastError());
    return 0;
  }
  lme_batched_bwd_kernel<T><<<grid, 256, smem, st>>>(A, B, out, gout, dA, dB, NB, Th, R, I, G);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}

template int lme_forward<float>(const float*, const float*, float*, int, int, int, cudaStream_t);
template int lme_forward<double>(const double*, const double*, double*, int, int, int, cudaStream_t);
template int lme_backward<float>(const float*, const float*, const float*, const float*, float*, float*, int, int, int, cudaStream_t);
template int lme_backward<double>(const double*, const double*, const double*, const double*, double*, double*, int, int, int, cudaStream_t);
template int lme_batched_forward<float>(const float*, const float*, float*, long long, int, int, int, cudaStream_t);
template int lme_batched_forward<double>(const double*, const double*, double*, long long, int, int, int, cudaStream_t);
template int lme_batched_backward<float>(const float*, const float*, const float*, const float*, float*, float*, long long, int, int, int, cudaStream_t);
template int lme_batched_backward<double>(const double*, const double*, const double*, const double*, double*, double*, long long, int, int, int, cudaStream_t);
