// Streaming thread-per-patch forward kernel for EPS layers with a tiny core (the HBM-bound regime of SURVEY.md
// section 8d: K=2 with Q=2..6, K=3 with Q=2; e.g. config 1 and the K=2 rows of the config-3 microbenchmark).
//
// One thread = one output patch.  The K*K*C factor vectors are read straight from x (pos2d unfold = index arithmetic;
// neighbouring threads read neighbouring pixels, so every warp load is one or two fully used cache lines and the
// overlap between patches is served by L1/L2: x crosses HBM once), the two Khatri-Rao halves are expanded in
// REGISTERS (template-unrolled, A = Q^MA and Bn = Q^MB values), and the core — transposed once per CTA into shared
// memory as [o][a][b] — is broadcast to all threads with 128-bit shared loads:
//     out[p][o] = sum_a kr1[a] * (sum_b kr2[b] * core[a][b][o])
// No Q^(K*K)-sized intermediate exists anywhere.  Algorithmic bytes per patch: Q floats of x (amortised) + O floats out.
//
// This file is the implementation shared by four translation units (eps_direct_{fwd,bwd}_{f32,f64}.cu): each defines
// DCTN_DIRECT_PART and includes it, so that the ~190 kernel instantiations compile in parallel (one TU took 7 minutes).
//   part 1 / 2: forward entry points, float / double        part 3 / 4: backward entry points, float / double
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "eps_kernels.h"

namespace {

constexpr int DTHREADS = 128;

template <int Q, int M> struct IPow { static constexpr int v = Q * IPow<Q, M - 1>::v; };
template <int Q> struct IPow<Q, 0> { static constexpr int v = 1; };

// kr[e] for e in [0, Q^M): product over factors j0..j0+M-1 (factor j0 slowest digit), built in registers
template <typename T, int Q, int M>
__device__ __forceinline__ void expand_kr(T (&kr)[IPow<Q, M>::v], const T* __restrict__ x, unsigned org, const EpsGeom& g, int j0) {
  kr[0] = T(1);
  int cur = 1;
#pragma unroll
  for (int j = 0; j < M; ++j) {
    T xv[Q];
    const T* px = x + org + g.foff[j0 + j];   // element offset is a multiple of Q: Q even -> 2-element vectors are aligned
    if constexpr (Q % 4 == 0 && sizeof(T) == 4) {
#pragma unroll
      for (int q = 0; q < Q; q += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(px + q));
        xv[q] = v.x; xv[q + 1] = v.y; xv[q + 2] = v.z; xv[q + 3] = v.w;
      }
    } else if constexpr (Q % 2 == 0) {
      using T2 = typename std::conditional<sizeof(T) == 4, float2, double2>::type;
#pragma unroll
      for (int q = 0; q < Q; q += 2) {
        const T2 v = __ldg(reinterpret_cast<const T2*>(px + q));
        xv[q] = v.x; xv[q + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int q = 0; q < Q; ++q) xv[q] = __ldg(px + q);
    }
#pragma unroll
    for (int e = IPow<Q, M>::v / Q - 1; e >= 0; --e) {   // only e < cur is live; later entries are overwritten before use
      if (e < cur) {
        const T base = kr[e];
#pragma unroll
        for (int q = Q - 1; q >= 0; --q) kr[e * Q + q] = base * xv[q];
      }
    }
    cur *= Q;
  }
}

template <typename T, int Q, int MA, int MB>
__global__ void __launch_bounds__(DTHREADS) direct_fwd_kernel(EpsGeom g, const T* __restrict__ x,
                                                              const T* __restrict__ core, T* __restrict__ out) {
  constexpr int A = IPow<Q, MA>::v, BN = IPow<Q, MB>::v;
  constexpr int V = 16 / sizeof(T);                 // elements per 128-bit shared load
  constexpr int BNP = (BN + V - 1) / V * V;         // padded b-run
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* cs = reinterpret_cast<T*>(smem_raw);           // [O][A][BNP]
  const int O = g.O;
  for (int idx = threadIdx.x; idx < O * A * BNP; idx += DTHREADS) {
    const int b = idx % BNP, r = idx / BNP, a = r % A, o = r / A;
    cs[idx] = (b < BN) ? core[((long long)a * BN + b) * O + o] : T(0);
  }
  __syncthreads();
  const unsigned hw = (unsigned)(g.Ho * g.Wo), Wo = (unsigned)g.Wo, P32 = (unsigned)g.P;   // P and |x| < 2^31 (host check)
  // PPT patches per thread per iteration: all their loads are issued before any arithmetic, so that enough bytes are in
  // flight per SM to cover HBM latency (B200 needs ~45 KB outstanding per SM for full bandwidth)
  constexpr int PPT = (A + BNP <= 16) ? 4 : ((A + BNP <= 40) ? 2 : 1);
  const unsigned stride = gridDim.x * DTHREADS;
  for (unsigned p0 = blockIdx.x * DTHREADS + threadIdx.x; p0 < P32; p0 += stride * PPT) {
    T kr1[PPT][A], kr2[PPT][BNP];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const unsigned p = p0 + k * stride;
      const unsigned pc = p < P32 ? p : p0;          // clamp: computed but not stored
      const unsigned b = pc / hw, r = pc - b * hw, h = r / Wo, w = r - h * Wo;
      const unsigned org = ((b * (unsigned)g.H + h) * (unsigned)g.W + w) * Q;
      expand_kr<T, Q, MA>(kr1[k], x, org, g, 0);
      T tmp[BN];
      expand_kr<T, Q, MB>(tmp, x, org, g, MA);
#pragma unroll
      for (int i = 0; i < BNP; ++i) kr2[k][i] = (i < BN) ? tmp[i] : T(0);
    }
    // two outputs per pass: every core value read from shared memory feeds 2*PPT FMAs; pairs are stored as one vector
    for (int o = 0; o < O; o += 2) {
      const bool two = o + 1 < O;
      const T* c0 = cs + (size_t)o * A * BNP;
      const T* c1 = two ? c0 + A * BNP : c0;
      T acc0[PPT], acc1[PPT];
#pragma unroll
      for (int k = 0; k < PPT; ++k) acc0[k] = acc1[k] = T(0);
#pragma unroll
      for (int a = 0; a < A; ++a) {
        T t0[PPT], t1[PPT];
#pragma unroll
        for (int k = 0; k < PPT; ++k) t0[k] = t1[k] = T(0);
#pragma unroll
        for (int bb = 0; bb < BNP; bb += V) {
          T u[V], v[V];
          if constexpr (sizeof(T) == 4) {
            const float4 uu = *reinterpret_cast<const float4*>(c0 + a * BNP + bb);
            const float4 vv = *reinterpret_cast<const float4*>(c1 + a * BNP + bb);
            u[0] = uu.x; u[1] = uu.y; u[2] = uu.z; u[3] = uu.w; v[0] = vv.x; v[1] = vv.y; v[2] = vv.z; v[3] = vv.w;
          } else {
            const double2 uu = *reinterpret_cast<const double2*>(c0 + a * BNP + bb);
            const double2 vv = *reinterpret_cast<const double2*>(c1 + a * BNP + bb);
            u[0] = uu.x; u[1] = uu.y; v[0] = vv.x; v[1] = vv.y;
          }
#pragma unroll
          for (int k = 0; k < PPT; ++k)
#pragma unroll
            for (int i = 0; i < V; ++i) {
              t0[k] = fma(kr2[k][bb + i], u[i], t0[k]);
              t1[k] = fma(kr2[k][bb + i], v[i], t1[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
          acc0[k] = fma(kr1[k][a], t0[k], acc0[k]);
          acc1[k] = fma(kr1[k][a], t1[k], acc1[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const unsigned p = p0 + k * stride;
        if (p < P32) {
          T* orow = out + (size_t)p * O;
          if (two && (O & 1) == 0) {
            using T2 = typename std::conditional<sizeof(T) == 4, float2, double2>::type;
            T2 v2; v2.x = acc0[k]; v2.y = acc1[k];
            *reinterpret_cast<T2*>(orow + o) = v2;     // p*O + o is even: aligned
          } else {
            orow[o] = acc0[k];
            if (two) orow[o + 1] = acc1[k];
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K = 2, C = 1 specialisation (the HBM-bound family).  One warp = RH consecutive output rows x up to 31 output columns:
//   * lane l loads pixel column w0 + l of the RH + 1 input rows ONCE (vector loads, fully coalesced); the right
//     neighbour of every pixel comes from lane l + 1 by __shfl_down — 1.25 loads per patch instead of 4;
//   * the outer product of a horizontally adjacent pixel pair, pp[r] = x[r][w] (x) x[r][w+1], is both the first
//     Khatri-Rao half of output row r and the second half of output row r - 1: computed once, used twice;
//   * out[r][o] = sum_a pp[r][a] * (sum_b pp[r+1][b] * core[a][b][o]), core broadcast from shared memory, each
//     128-bit shared load feeding 2 * RH FMAs per lane.
// PIX (Q = 2 only): x is the raw pixel image (B, H, W) and the feature map of the reference's data loader,
// phi(u) = scale * (sin^2(pi u / 2), cos^2(pi u / 2))  (dctn/dataset_loading.py:33-36), is evaluated on load: one float
// per pixel crosses HBM instead of two (dctn_eps_forward_from_pixels).
template <typename T, int Q, int OT, bool PIX = false>   // OT: compile-time Q_out (2..8), or 0 = runtime (any Q_out, pairs per pass)
__global__ void __launch_bounds__(DTHREADS) direct_k2_kernel(EpsGeom g, const T* __restrict__ x, const T* __restrict__ core,
                                                             T* __restrict__ out, T phi_scale) {
  constexpr int RH = 4;
  constexpr int A = Q * Q;                          // = Bn
  constexpr int V = 16 / sizeof(T);
  constexpr int BNP = (A + V - 1) / V * V;
  constexpr int OC = OT > 0 ? OT : 2;               // outputs computed per pass
  using T2 = typename std::conditional<sizeof(T) == 4, float2, double2>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* cs = reinterpret_cast<T*>(smem_raw);           // [O][A][BNP]
  const int O = OT > 0 ? OT : g.O;
  for (int idx = threadIdx.x; idx < O * A * BNP; idx += DTHREADS) {
    const int b = idx % BNP, r = idx / BNP, a = r % A, o = r / A;
    cs[idx] = (b < A) ? core[((long long)a * A + b) * O + o] : T(0);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned nhb = (unsigned)(g.Ho + RH - 1) / RH;   // row blocks per image
  const unsigned ntw = (unsigned)(g.Wo + 30) / 31;       // column tiles of 31 outputs (+1 halo lane)
  const unsigned ntask = (unsigned)g.B * nhb * ntw;
  const unsigned nwarps = gridDim.x * (DTHREADS / 32);
  const unsigned xrow = (unsigned)g.W * Q, orow_stride = (unsigned)g.Wo * O;   // 32-bit offsets (host checks the sizes)
  for (unsigned task = blockIdx.x * (DTHREADS / 32) + (threadIdx.x >> 5); task < ntask; task += nwarps) {
    const unsigned tw = task % ntw, t2 = task / ntw, hb = t2 % nhb, b = t2 / nhb;
    const unsigned h0 = hb * RH, wcol = tw * 31 + lane;  // input (and output) column of this lane
    // (1) this lane's pixel column, RH + 1 rows; out-of-range lanes/rows re-read a valid pixel (value never stored)
    const unsigned wc = wcol < (unsigned)g.W ? wcol : 0u;
    const T* px = x + ((b * (unsigned)g.H + h0) * (unsigned)g.W + wc) * (PIX ? 1 : Q);
    T xv[RH + 1][Q];
#pragma unroll
    for (int r = 0; r <= RH; ++r) {
      const T* pr = px + ((h0 + r < (unsigned)g.H) ? r * (PIX ? (unsigned)g.W : xrow) : 0u);
      if constexpr (PIX) {
        const T u = __ldg(pr) * T(0.5);
        T sn, cs2;
        if constexpr (sizeof(T) == 4) sincospif(u, &sn, &cs2); else sincospi(u, &sn, &cs2);
        xv[r][0] = phi_scale * sn * sn;
        xv[r][Q - 1] = phi_scale * cs2 * cs2;
      } else if constexpr (Q % 4 == 0 && sizeof(T) == 4) {
#pragma unroll
        for (int q = 0; q < Q; q += 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(pr + q));
          xv[r][q] = v.x; xv[r][q + 1] = v.y; xv[r][q + 2] = v.z; xv[r][q + 3] = v.w;
        }
      } else if constexpr (Q % 2 == 0) {
#pragma unroll
        for (int q = 0; q < Q; q += 2) {
          const T2 v = __ldg(reinterpret_cast<const T2*>(pr + q));
          xv[r][q] = v.x; xv[r][q + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int q = 0; q < Q; ++q) xv[r][q] = __ldg(pr + q);
      }
    }
    // (2) pair products with the right neighbour (lane + 1)
    T pp[RH + 1][BNP];
#pragma unroll
    for (int r = 0; r <= RH; ++r) {
      T xr[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) xr[q] = __shfl_down_sync(0xffffffffu, xv[r][q], 1);
#pragma unroll
      for (int i = 0; i < Q; ++i)
#pragma unroll
        for (int j = 0; j < Q; ++j) pp[r][i * Q + j] = xv[r][i] * xr[j];
#pragma unroll
      for (int i = A; i < BNP; ++i) pp[r][i] = T(0);
    }
    // (3) contraction with the core, OC outputs per pass
    const bool lane_ok = lane < 31 && wcol < (unsigned)g.Wo;
    T* obase = out + ((b * (unsigned)g.Ho + h0) * (unsigned)g.Wo + (lane_ok ? wcol : 0u)) * O;
    for (int o = 0; o < O; o += OC) {
      T acc[OC][RH];
#pragma unroll
      for (int c = 0; c < OC; ++c)
#pragma unroll
        for (int r = 0; r < RH; ++r) acc[c][r] = T(0);
#pragma unroll
      for (int a = 0; a < A; ++a) {
        T t[OC][RH];
#pragma unroll
        for (int c = 0; c < OC; ++c)
#pragma unroll
          for (int r = 0; r < RH; ++r) t[c][r] = T(0);
#pragma unroll
        for (int bb = 0; bb < BNP; bb += V) {
#pragma unroll
          for (int c = 0; c < OC; ++c) {
            const int oc = (OT > 0 || o + c < O) ? o + c : o;   // runtime-O tail: recompute output o (not stored)
            T u[V];
            if constexpr (sizeof(T) == 4) {
              const float4 uu = *reinterpret_cast<const float4*>(cs + ((size_t)oc * A + a) * BNP + bb);
              u[0] = uu.x; u[1] = uu.y; u[2] = uu.z; u[3] = uu.w;
            } else {
              const double2 uu = *reinterpret_cast<const double2*>(cs + ((size_t)oc * A + a) * BNP + bb);
              u[0] = uu.x; u[1] = uu.y;
            }
#pragma unroll
            for (int r = 0; r < RH; ++r)
#pragma unroll
              for (int i = 0; i < V; ++i) t[c][r] = fma(pp[r + 1][bb + i], u[i], t[c][r]);
          }
        }
#pragma unroll
        for (int c = 0; c < OC; ++c)
#pragma unroll
          for (int r = 0; r < RH; ++r) acc[c][r] = fma(pp[r][a], t[c][r], acc[c][r]);
      }
#pragma unroll
      for (int r = 0; r < RH; ++r) {
        if (lane_ok && h0 + r < (unsigned)g.Ho) {
          T* orow = obase + r * orow_stride + o;
          if constexpr (OT > 0 && OT % 2 == 0) {
#pragma unroll
            for (int c = 0; c < OC; c += 2) {
              T2 v2; v2.x = acc[c][r]; v2.y = acc[c + 1][r];
              *reinterpret_cast<T2*>(orow + c) = v2;          // (patch * O + c) is even: aligned
            }
          } else {
#pragma unroll
            for (int c = 0; c < OC; ++c)
              if (OT > 0 || o + c < O) orow[c] = acc[c][r];
          }
        }
      }
    }
  }
}

template <typename T, int Q, int OT, bool PIX = false>
int launch_direct_k2(const EpsGeom& g, const T* x, const T* core, T* out, cudaStream_t st, T phi_scale = T(0)) {
  constexpr int A = Q * Q;
  constexpr int V = 16 / sizeof(T);
  constexpr int BNP = (A + V - 1) / V * V;
  const size_t smem = (size_t)g.O * A * BNP * sizeof(T);
  auto k = direct_k2_kernel<T, Q, OT, PIX>;
  DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long ntask = (long long)g.B * ((g.Ho + 3) / 4) * ((g.Wo + 30) / 31);
  long long blocks = (ntask + DTHREADS / 32 - 1) / (DTHREADS / 32);
  const long long cap = 148ll * 16;
  if (blocks > cap) blocks = cap;
  k<<<(unsigned)blocks, DTHREADS, smem, st>>>(g, x, core, out, phi_scale);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}
template <typename T, int Q>
int dispatch_direct_k2(const EpsGeom& g, const T* x, const T* core, T* out, cudaStream_t st) {
  switch (g.O) {
    case 2: return launch_direct_k2<T, Q, 2>(g, x, core, out, st);
    case 3: return launch_direct_k2<T, Q, 3>(g, x, core, out, st);
    case 4: return launch_direct_k2<T, Q, 4>(g, x, core, out, st);
    case 6: return launch_direct_k2<T, Q, 6>(g, x, core, out, st);
    default: return launch_direct_k2<T, Q, 0>(g, x, core, out, st);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Backward for tiny cores (same register-resident Khatri-Rao halves as the forward).
//
// Core gradient: dcore[a][b][o] = sum_p kr1[p][a] kr2[p][b] gout[p][o].  Each thread accumulates PPT patches locally
// per (a, b, o) term, the warp reduces the term with shuffles, lane 0 adds it into this warp's shared-memory copy of
// dcore; the CTA then writes one partial per CTA and reduce_partials sums them in fixed order (deterministic).
template <typename T, int Q, int MA, int MB>
__global__ void __launch_bounds__(DTHREADS) direct_dcore_kernel(EpsGeom g, const T* __restrict__ x, const T* __restrict__ gout,
                                                                T* __restrict__ part) {
  constexpr int A = IPow<Q, MA>::v, BN = IPow<Q, MB>::v;
  constexpr int PPT = (A + BN <= 16) ? 4 : 2;
  constexpr int NW = DTHREADS / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* acc_s = reinterpret_cast<T*>(smem_raw);        // [NW][A*BN*O]
  const int O = g.O, DO = A * BN * O;
  for (int i = threadIdx.x; i < NW * DO; i += DTHREADS) acc_s[i] = T(0);
  __syncthreads();
  T* mine = acc_s + (threadIdx.x >> 5) * DO;
  const int lane = threadIdx.x & 31;
  const unsigned hw = (unsigned)(g.Ho * g.Wo), Wo = (unsigned)g.Wo, P32 = (unsigned)g.P;
  const unsigned stride = gridDim.x * DTHREADS;
  // all lanes of a warp iterate together (p0 differs only by lane), out-of-range patches contribute zeros
  for (unsigned p0 = blockIdx.x * DTHREADS + (threadIdx.x & ~31u); p0 < P32; p0 += stride * PPT) {
    T kr1[PPT][A], kr2[PPT][BN];
    const T* gp[PPT];
    bool ok[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const unsigned p = p0 + k * stride + lane;
      ok[k] = p < P32;
      const unsigned pc = ok[k] ? p : 0u;
      const unsigned b = pc / hw, r = pc - b * hw, h = r / Wo, w = r - h * Wo;
      const unsigned org = ((b * (unsigned)g.H + h) * (unsigned)g.W + w) * Q;
      expand_kr<T, Q, MA>(kr1[k], x, org, g, 0);
      expand_kr<T, Q, MB>(kr2[k], x, org, g, MA);
      gp[k] = gout + (size_t)pc * O;
    }
    for (int o = 0; o < O; ++o) {
      T gv[PPT];
#pragma unroll
      for (int k = 0; k < PPT; ++k) gv[k] = ok[k] ? __ldg(gp[k] + o) : T(0);
#pragma unroll
      for (int a = 0; a < A; ++a) {
        T ga[PPT];
#pragma unroll
        for (int k = 0; k < PPT; ++k) ga[k] = kr1[k][a] * gv[k];
#pragma unroll
        for (int b = 0; b < BN; ++b) {
          T v = T(0);
#pragma unroll
          for (int k = 0; k < PPT; ++k) v = fma(ga[k], kr2[k][b], v);
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
          if (lane == 0) mine[(a * BN + b) * O + o] += v;
        }
      }
    }
  }
  __syncthreads();
  T* dst = part + (size_t)blockIdx.x * DO;
  for (int i = threadIdx.x; i < DO; i += DTHREADS) {
    T v = T(0);
#pragma unroll
    for (int w = 0; w < NW; ++w) v += acc_s[w * DO + i];
    dst[i] = v;
  }
}

// Core gradient, tiled: dcore[a][n] = sum_p kr1[p][a] * GB[p][n],  n = (b, o),  GB[p][n] = kr2[p][b] * gout[p][o] — a GEMM
// with a tiny M x N output (M = A <= 36, N = Bn*O) and the patches as the reduction axis.  The shuffle kernel above
// pays 10 instructions per (a, b, o) term per 32..128 patches; here a CTA stages a chunk of PC patches as two tables in
// shared memory (phase 1: one thread per patch, both Khatri-Rao halves expanded in registers, 128-bit stores), then
// every thread owns a 4 x 4 register tile of dcore (x TPT) and a SLICE of the chunk's patches (phase 2: two 128-bit
// shared loads per 16 FMAs).  Tiny cores have fewer tiles than threads, so the CTA's threads are split into NS slices
// over the patches and the slices are summed in fixed order at the end (deterministic); one partial per CTA.
constexpr int DC_THREADS = 256;

template <typename T> __device__ __forceinline__ void ld4(const T* p, T (&v)[4]);
template <> __device__ __forceinline__ void ld4<float>(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void ld4<double>(const double* p, double (&v)[4]) {
  const double2 t0 = *reinterpret_cast<const double2*>(p), t1 = *reinterpret_cast<const double2*>(p + 2);
  v[0] = t0.x; v[1] = t0.y; v[2] = t1.x; v[3] = t1.y;
}

template <typename T> __device__ __forceinline__ void st4(T* p, T v0, T v1, T v2, T v3);
template <> __device__ __forceinline__ void st4<float>(float* p, float v0, float v1, float v2, float v3) {
  *reinterpret_cast<float4*>(p) = make_float4(v0, v1, v2, v3);
}
template <> __device__ __forceinline__ void st4<double>(double* p, double v0, double v1, double v2, double v3) {
  *reinterpret_cast<double2*>(p) = make_double2(v0, v1);
  *reinterpret_cast<double2*>(p + 2) = make_double2(v2, v3);
}

// Table layouts: K1[pl][a] with row stride SA, GB[pl][n] with n = o * BNP + b (BNP = Bn rounded up to 4) and row stride
// SN; both strides are ODD multiples of four elements, so the one-row-per-lane 128-bit stores of phase 1 and the
// 128-bit loads of phase 2 are free of bank conflicts.
template <typename T, int Q, int MA, int MB, int TPT>
__global__ void __launch_bounds__(DC_THREADS) direct_dcore_tiled_kernel(EpsGeom g, const T* __restrict__ x,
                                                                        const T* __restrict__ gout, T* __restrict__ part,
                                                                        int PC, int SA, int SN, int ntn_group, int NS) {
  constexpr int A = IPow<Q, MA>::v, BN = IPow<Q, MB>::v;
  constexpr int AP = (A + 3) & ~3, BNP = (BN + 3) & ~3;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* K1 = reinterpret_cast<T*>(smem_raw);   // [PC][SA]
  T* GB = K1 + (size_t)PC * SA;             // [PC][SN]: the column tiles [tn0, tn0 + ntn) of this CTA's group only
  const int O = g.O, N = BN * O;
  // blockIdx.y = group of 4-wide column tiles (cores with more than 512 tiles are cut along n; every group sees all
  // patches and writes its own columns of the CTA's partial)
  const int tn0 = blockIdx.y * ntn_group;
  const int ntn = min(ntn_group, ((O * BNP) >> 2) - tn0);
  const int ntiles = (AP >> 2) * ntn;
  const int tid = threadIdx.x;
  // tiles of this thread: tile index runs fastest over threads, then the slice
  int slice, ta[TPT], tn[TPT];
  bool live[TPT];
  if (TPT == 1) {
    const int tmax = (AP >> 2) * ntn_group;   // the slice split is that of a full group
    const int t = tid % tmax;
    slice = tid / tmax;
    live[0] = slice < NS && t < ntiles;
    const int tt = live[0] ? t : 0;
    ta[0] = tt / ntn; tn[0] = tt - ta[0] * ntn;
  } else {
    slice = 0;
#pragma unroll
    for (int k = 0; k < TPT; ++k) {
      const int t = tid + k * DC_THREADS;
      live[k] = t < ntiles;
      const int tt = live[k] ? t : 0;
      ta[k] = tt / ntn; tn[k] = tt - ta[k] * ntn;
    }
  }
  T acc[TPT][4][4];
#pragma unroll
  for (int k = 0; k < TPT; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[k][i][j] = T(0);

  const unsigned hw = (unsigned)(g.Ho * g.Wo), Wo = (unsigned)g.Wo;
  const long long nchunks = (g.P + PC - 1) / PC;
  for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
    // ---- phase 1: tables of the chunk
    for (int pl = tid; pl < PC; pl += DC_THREADS) {
      const long long p = c * PC + pl;
      T* k1 = K1 + (size_t)pl * SA;
      T* gb = GB + (size_t)pl * SN;
      const bool ok = p < g.P;
      const unsigned pc = ok ? (unsigned)p : 0u;
      const unsigned b = pc / hw, r = pc - b * hw, h = r / Wo, w = r - h * Wo;
      const unsigned org = ((b * (unsigned)g.H + h) * (unsigned)g.W + w) * Q;
      T kr1[AP], kr2[BNP];
      {
        T t1[A], t2[BN];
        expand_kr<T, Q, MA>(t1, x, org, g, 0);
        expand_kr<T, Q, MB>(t2, x, org, g, MA);
#pragma unroll
        for (int a = 0; a < AP; ++a) kr1[a] = (a < A && ok) ? t1[a] : T(0);
#pragma unroll
        for (int bb = 0; bb < BNP; ++bb) kr2[bb] = (bb < BN && ok) ? t2[bb] : T(0);
      }
#pragma unroll
      for (int a = 0; a < AP; a += 4) st4<T>(k1 + a, kr1[a], kr1[a + 1], kr1[a + 2], kr1[a + 3]);
      const T* gp = gout + (size_t)pc * O;
      for (int o = 0; o < O; ++o) {
        const T gv = __ldg(gp + o);
#pragma unroll
        for (int bb = 0; bb < BNP; bb += 4) {
          const int ct = ((o * BNP + bb) >> 2) - tn0;   // column tile inside this group?
          if (ct >= 0 && ct < ntn) st4<T>(gb + 4 * ct, kr2[bb] * gv, kr2[bb + 1] * gv, kr2[bb + 2] * gv, kr2[bb + 3] * gv);
        }
      }
    }
    __syncthreads();
    // ---- phase 2: this thread's slice of the patches into its register tiles
    if (TPT > 1 || live[0]) {
#pragma unroll 2
      for (int pl = slice; pl < PC; pl += NS) {
#pragma unroll
        for (int k = 0; k < TPT; ++k) {
          if (TPT > 1 && !live[k]) continue;
          T av[4], nv[4];
          ld4<T>(K1 + (size_t)pl * SA + 4 * ta[k], av);
          ld4<T>(GB + (size_t)pl * SN + 4 * tn[k], nv);
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[k][i][j] = fma(av[i], nv[j], acc[k][i][j]);
        }
      }
    }
    __syncthreads();
  }
  // ---- slices -> one partial per CTA (fixed summation order); dcore is stored [a][b][o]
  const int NPs = 4 * ntn;
  T* red = reinterpret_cast<T*>(smem_raw);  // [NS][AP][NPs]
#pragma unroll
  for (int k = 0; k < TPT; ++k) {
    if (!live[k]) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[((size_t)slice * AP + 4 * ta[k] + i) * NPs + 4 * tn[k] + j] = acc[k][i][j];
  }
  __syncthreads();
  T* dst = part + (size_t)blockIdx.x * A * N;
  for (int i = tid; i < A * NPs; i += DC_THREADS) {
    const int a = i / NPs, nl = i - a * NPs, n = 4 * tn0 + nl, o = n / BNP, bb = n - o * BNP;
    if (bb >= BN) continue;   // padding column
    T v = T(0);
    for (int sidx = 0; sidx < NS; ++sidx) v += red[((size_t)sidx * AP + a) * NPs + nl];
    dst[a * N + bb * O + o] = v;
  }
}

// cut the column tiles into groups of at most 512 register tiles (two per thread)
static void dcore_tiled_groups(int AP, int NP, int* ngroups, int* ntn_group) {
  const int nta = AP >> 2, ntn = NP >> 2;
  int per = (2 * DC_THREADS) / nta;            // column tiles per group
  if (per < 1) per = 1;
  if (per > ntn) per = ntn;
  *ngroups = (ntn + per - 1) / per;
  *ntn_group = (ntn + *ngroups - 1) / *ngroups;   // balanced groups
}
static bool dcore_tiled_fits(int A, int Bn, int O, size_t es) {
  const int AP = (A + 3) & ~3, NP = O * ((Bn + 3) & ~3);
  if ((AP >> 2) > 2 * DC_THREADS) return false;
  int ngroups, ntn_group;
  dcore_tiled_groups(AP, NP, &ngroups, &ntn_group);
  return ngroups <= 64 && (size_t)64 * (AP + 4 * ntn_group + 8) * es <= 160 * 1024 && (size_t)AP * 4 * ntn_group * es <= 160 * 1024;
}

// returns 1 when the tiled kernel was launched, 0 when the shape does not fit it, < 0 on error
template <typename T, int Q, int MA, int MB>
int try_launch_dcore_tiled(const EpsGeom& g, const T* x, const T* gout, T* part, int* blocks_out, cudaStream_t st) {
  constexpr int A = IPow<Q, MA>::v, BN = IPow<Q, MB>::v;
  constexpr int AP = (A + 3) & ~3, BNP = (BN + 3) & ~3;
  const int NP = g.O * BNP;
  int ngroups, ntn_group;
  dcore_tiled_groups(AP, NP, &ngroups, &ntn_group);
  const int ntiles = (AP >> 2) * ntn_group;            // tiles of a full group, <= 512
  const int TPT = ntiles > DC_THREADS ? 2 : 1;
  const int NS = TPT == 1 ? DC_THREADS / ntiles : 1;
  const int NPg = 4 * ntn_group;
  const int SA = ((AP >> 2) & 1) ? AP : AP + 4, SN = (ntn_group & 1) ? NPg : NPg + 4;   // odd multiples of 4
  int PC = 256;
  while (PC > 64 && (size_t)PC * (SA + SN) * sizeof(T) > 72 * 1024) PC >>= 1;
  size_t smem = (size_t)PC * (SA + SN) * sizeof(T);
  const size_t red = (size_t)NS * AP * NPg * sizeof(T);
  if (red > smem) smem = red;
  if (smem > 160 * 1024) return 0;
  const long long nchunks = (g.P + PC - 1) / PC;
  const int per_sm = smem > 100 * 1024 ? 1 : smem > 70 * 1024 ? 2 : smem > 50 * 1024 ? 3 : 4;
  long long blocks = (148ll * per_sm + ngroups - 1) / ngroups;   // partials = blocks; the grid is blocks x ngroups
  if (blocks > nchunks) blocks = nchunks;
  auto k1 = direct_dcore_tiled_kernel<T, Q, MA, MB, 1>;
  auto k2 = direct_dcore_tiled_kernel<T, Q, MA, MB, 2>;
  auto k = TPT == 1 ? k1 : k2;
  DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<dim3((unsigned)blocks, (unsigned)ngroups), DC_THREADS, smem, st>>>(g, x, gout, part, PC, SA, SN, ntn_group, NS);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  *blocks_out = (int)blocks;
  return 1;
}

// Core gradient for the smallest cores (A * Bn * O <= 96, i.e. K = 2, Q = 2, O <= 6 — config 1 and the HBM-bound rows
// of the config-3 grid): every thread keeps a FULL private copy of dcore in registers and streams its patches through
// it — no shared memory, no barrier and no shuffle in the loop; the copies are summed once at the end (shuffles inside
// the warp, shared memory across the warps, one partial per CTA, all in fixed order).
template <typename T, int Q, int MA, int MB, int O>
__global__ void __launch_bounds__(DTHREADS) direct_dcore_reg_kernel(EpsGeom g, const T* __restrict__ x,
                                                                    const T* __restrict__ gout, T* __restrict__ part) {
  constexpr int A = IPow<Q, MA>::v, BN = IPow<Q, MB>::v, DO = A * BN * O, NW = DTHREADS / 32;
  __shared__ T red[NW][DO];
  T acc[A * BN][O];
#pragma unroll
  for (int e = 0; e < A * BN; ++e)
#pragma unroll
    for (int o = 0; o < O; ++o) acc[e][o] = T(0);
  const unsigned hw = (unsigned)(g.Ho * g.Wo), Wo = (unsigned)g.Wo, P32 = (unsigned)g.P;
  const unsigned stride = gridDim.x * DTHREADS;
  // (b, h, w) of the thread's patch advance by the decomposed grid stride: no divisions in the loop
  const unsigned pfirst = blockIdx.x * DTHREADS + threadIdx.x;
  unsigned b = pfirst / hw, h = (pfirst - b * hw) / Wo, w = pfirst - b * hw - h * Wo;
  const unsigned sb = stride / hw, sh = (stride - sb * hw) / Wo, sw = stride - sb * hw - sh * Wo, Ho = (unsigned)g.Ho;
  for (unsigned p = pfirst; p < P32; p += stride) {
    const unsigned org = ((b * (unsigned)g.H + h) * (unsigned)g.W + w) * Q;
    w += sw; if (w >= Wo) { w -= Wo; ++h; }
    h += sh; if (h >= Ho) { h -= Ho; ++b; }
    b += sb;
    T kr1[A], kr2[BN], gv[O];
    expand_kr<T, Q, MA>(kr1, x, org, g, 0);
    expand_kr<T, Q, MB>(kr2, x, org, g, MA);
#pragma unroll
    for (int o = 0; o < O; ++o) gv[o] = __ldg(gout + (size_t)p * O + o);
#pragma unroll
    for (int a = 0; a < A; ++a)
#pragma unroll
      for (int bb = 0; bb < BN; ++bb) {
        const T kk = kr1[a] * kr2[bb];
#pragma unroll
        for (int o = 0; o < O; ++o) acc[a * BN + bb][o] = fma(kk, gv[o], acc[a * BN + bb][o]);
      }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int e = 0; e < A * BN; ++e)
#pragma unroll
    for (int o = 0; o < O; ++o) {
      T v = acc[e][o];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0) red[warp][e * O + o] = v;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < DO; i += DTHREADS) {
    T v = T(0);
#pragma unroll
    for (int w = 0; w < NW; ++w) v += red[w][i];
    part[(size_t)blockIdx.x * DO + i] = v;
  }
}

template <typename T, int Q, int MA, int MB, int O>
int launch_dcore_reg(const EpsGeom& g, const T* x, const T* gout, T* part, int* blocks_out, cudaStream_t st) {
  long long blocks = (g.P + DTHREADS - 1) / DTHREADS;
  if (blocks > 148ll * 8) blocks = 148ll * 8;
  direct_dcore_reg_kernel<T, Q, MA, MB, O><<<(unsigned)blocks, DTHREADS, 0, st>>>(g, x, gout, part);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  *blocks_out = (int)blocks;
  return 1;
}
template <typename T, int Q, int MA, int MB>
int try_launch_dcore_reg(const EpsGeom& g, const T* x, const T* gout, T* part, int* blocks_out, cudaStream_t st) {
  constexpr int AB = IPow<Q, MA>::v * IPow<Q, MB>::v;
  if constexpr (AB <= 16) {
    switch (g.O) {
      case 1: return launch_dcore_reg<T, Q, MA, MB, 1>(g, x, gout, part, blocks_out, st);
      case 2: return launch_dcore_reg<T, Q, MA, MB, 2>(g, x, gout, part, blocks_out, st);
      case 3: return launch_dcore_reg<T, Q, MA, MB, 3>(g, x, gout, part, blocks_out, st);
      case 4: return launch_dcore_reg<T, Q, MA, MB, 4>(g, x, gout, part, blocks_out, st);
      case 5: return launch_dcore_reg<T, Q, MA, MB, 5>(g, x, gout, part, blocks_out, st);
      case 6: return launch_dcore_reg<T, Q, MA, MB, 6>(g, x, gout, part, blocks_out, st);
      default: return 0;
    }
  }
  return 0;
}

// out[i] = sum_z part[z * count + i] for MANY partials of a SMALL tensor (one partial per CTA of the kernels above):
// the generic reduce_partials_kernel walks the partials serially per output (1184 dependent-latency loads for a
// 32-element core: 80 us, more than the gradient itself).  Here a CTA of 1024 threads takes 32 outputs x 32 slices of
// the partials (coalesced over the outputs), then sums the slices through shared memory — fixed order, deterministic.
template <typename T>
__global__ void __launch_bounds__(1024) reduce_partials_wide_kernel(const T* __restrict__ part, T* __restrict__ out,
                                                                    int count, int splits) {
  __shared__ T red[32][33];
  const int il = threadIdx.x & 31, zs = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + il;
  T v = T(0);
  if (i < count)
    for (int z = zs; z < splits; z += 32) v += part[(size_t)z * count + i];
  red[zs][il] = v;
  __syncthreads();
  if (zs == 0 && i < count) {
    T t = T(0);
#pragma unroll
    for (int k = 0; k < 32; ++k) t += red[k][il];
    out[i] = t;
  }
}
template <typename T>
int launch_reduce_partials_wide(const T* part, T* out, int count, int splits, cudaStream_t st) {
  reduce_partials_wide_kernel<T><<<(count + 31) / 32, 1024, 0, st>>>(part, out, count, splits);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}

// Input gradient, written per patch as dxp[p][j][q] (the shared gather kernel then sums the K*K overlapping patches of
// every pixel):  G[a][b] = sum_o gout[p][o] core[a][b][o];  W1[a] = sum_b G[a][b] kr2[b];  W2[b] = sum_a G[a][b] kr1[a];
// d x_j[q] = sum over the entries of its half with digit_j == q of W * (product of the other factors of that half).
template <typename T, int Q, int M>
__device__ __forceinline__ void leave_one_out(const T (&w)[IPow<Q, M>::v], const T (&xv)[M][Q], T* __restrict__ dst /*[M][Q]*/,
                                              int ds = 1 /* element stride of dst */) {
  constexpr int E = IPow<Q, M>::v;
#pragma unroll
  for (int t = 0; t < M; ++t) {
    T acc[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) acc[q] = T(0);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      T v = w[e];
      int ee = e, dig = 0;
#pragma unroll
      for (int u = M - 1; u >= 0; --u) {
        const int d = ee % Q;
        ee /= Q;
        if (u == t) dig = d; else v *= xv[u][d];
      }
#pragma unroll
      for (int q = 0; q < Q; ++q)
        if (q == dig) acc[q] += v;
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) dst[(t * Q + q) * ds] = acc[q];
  }
}

// The input-gradient kernels keep the core TRANSPOSED in shared memory, [o][a][BNP] (Bn padded to a multiple of four):
// for a fixed (o, a) the Bn values a thread needs are contiguous, so they arrive as broadcast 128-bit loads instead of
// one strided 32-bit load (plus its address arithmetic) per FMA.
template <typename T, int Q, int MB> struct DxCoreLayout { static constexpr int BNP = (IPow<Q, MB>::v + 3) & ~3; };
template <typename T, int A, int BN, int BNP>
__device__ __forceinline__ void stage_core_transposed(T* cs, const T* __restrict__ core, int O, int nthreads) {
  for (int idx = threadIdx.x; idx < O * A * BNP; idx += nthreads) {
    const int b = idx % BNP, r = idx / BNP, a = r % A, o = r / A;
    cs[idx] = (b < BN) ? core[((size_t)a * BN + b) * O + o] : T(0);
  }
}

// everything of one patch in registers; writes d x_j[q] of the patch's n factors to dst[(j*Q + q) * ds]
template <typename T, int Q, int MA, int MB>
__device__ __forceinline__ void patch_dx(const EpsGeom& g, const T* __restrict__ x, const T* cs /* core [O][A][BNP], shared */,
                                         const T* __restrict__ gp /* gout row */, unsigned org, T* __restrict__ dst, int ds) {
  constexpr int A = IPow<Q, MA>::v, BN = IPow<Q, MB>::v;
  const int O = g.O;
  T xa[MA][Q], xb[MB][Q];
#pragma unroll
  for (int j = 0; j < MA; ++j)
#pragma unroll
    for (int q = 0; q < Q; ++q) xa[j][q] = __ldg(x + org + g.foff[j] + q);
#pragma unroll
  for (int j = 0; j < MB; ++j)
#pragma unroll
    for (int q = 0; q < Q; ++q) xb[j][q] = __ldg(x + org + g.foff[MA + j] + q);
  T kr1[A], kr2[BN];
  expand_kr<T, Q, MA>(kr1, x, org, g, 0);
  expand_kr<T, Q, MB>(kr2, x, org, g, MA);
  T w1[A], w2[BN];
#pragma unroll
  for (int i = 0; i < BN; ++i) w2[i] = T(0);
  constexpr int BNP = DxCoreLayout<T, Q, MB>::BNP;
#pragma unroll
  for (int a = 0; a < A; ++a) {
    T grow[BNP];
#pragma unroll
    for (int i = 0; i < BNP; ++i) grow[i] = T(0);
    const T* ca = cs + a * BNP;
    for (int o = 0; o < O; ++o) {           // core row (o, a): 128-bit broadcast loads, one per four FMAs
      const T gv = __ldg(gp + o);
#pragma unroll
      for (int i = 0; i < BNP; i += 4) {
        T u[4];
        ld4<T>(ca + (size_t)o * (A * BNP) + i, u);
#pragma unroll
        for (int k = 0; k < 4; ++k) grow[i + k] = fma(gv, u[k], grow[i + k]);
      }
    }
    T s1 = T(0);
#pragma unroll
    for (int i = 0; i < BN; ++i) {
      s1 = fma(grow[i], kr2[i], s1);
      w2[i] = fma(grow[i], kr1[a], w2[i]);
    }
    w1[a] = s1;
  }
  leave_one_out<T, Q, MA>(w1, xa, dst, ds);
  leave_one_out<T, Q, MB>(w2, xb, dst + MA * Q * ds, ds);
}

template <typename T, int Q, int MA, int MB>
__global__ void __launch_bounds__(DTHREADS) direct_dx_kernel(EpsGeom g, const T* __restrict__ x, const T* __restrict__ core,
                                                             const T* __restrict__ gout, T* __restrict__ dxp) {
  constexpr int A = IPow<Q, MA>::v, BN = IPow<Q, MB>::v;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* cs = reinterpret_cast<T*>(smem_raw);           // core transposed: [O][A][BNP]
  const int O = g.O;
  stage_core_transposed<T, A, BN, DxCoreLayout<T, Q, MB>::BNP>(cs, core, O, DTHREADS);
  __syncthreads();
  const unsigned hw = (unsigned)(g.Ho * g.Wo), Wo = (unsigned)g.Wo, P32 = (unsigned)g.P;
  for (unsigned p = blockIdx.x * DTHREADS + threadIdx.x; p < P32; p += gridDim.x * DTHREADS) {
    const unsigned b = p / hw, r = p - b * hw, h = r / Wo, w = r - h * Wo;
    const unsigned org = ((b * (unsigned)g.H + h) * (unsigned)g.W + w) * Q;
    patch_dx<T, Q, MA, MB>(g, x, cs, gout + (size_t)p * O, org, dxp + (size_t)p * (MA + MB) * Q, 1);
  }
}

// Input gradient, FUSED per image: one CTA owns image b — its Ho*Wo patches write their per-factor contributions into
// shared memory ([j*Q + q][patch], odd row stride), and after one barrier the same CTA sums, for every pixel of the
// image, the <= K*K patches that contain it (fixed order) and stores dx coalesced.  The P x n x Q intermediate of the
// two-kernel path (dxp: written and re-read through HBM, 190 MB of the 265 MB it moves on the config-1 shape at
// B = 4096) never leaves the SM; HBM traffic is x + gout + dx.
constexpr int DXI_THREADS = 256;
template <typename T, int Q, int MA, int MB>
__global__ void __launch_bounds__(DXI_THREADS) direct_dx_image_kernel(EpsGeom g, const T* __restrict__ x,
                                                                      const T* __restrict__ core, const T* __restrict__ gout,
                                                                      T* __restrict__ dx, int NPP) {
  constexpr int A = IPow<Q, MA>::v, BN = IPow<Q, MB>::v, NF = MA + MB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* cs = reinterpret_cast<T*>(smem_raw);           // core transposed: [O][A][BNP]
  const int O = g.O;
  constexpr int BNP = DxCoreLayout<T, Q, MB>::BNP;
  T* dxs = cs + (size_t)O * A * BNP;                // [NF*Q][NPP]
  stage_core_transposed<T, A, BN, BNP>(cs, core, O, DXI_THREADS);
  __syncthreads();
  const int b = blockIdx.x, Wo = g.Wo, npatch = g.Ho * g.Wo;
  for (int pl = threadIdx.x; pl < npatch; pl += DXI_THREADS) {
    const int h = pl / Wo, w = pl - h * Wo;
    const unsigned org = (((unsigned)b * (unsigned)g.H + h) * (unsigned)g.W + w) * Q;
    patch_dx<T, Q, MA, MB>(g, x, cs, gout + ((size_t)b * npatch + pl) * O, org, dxs + pl, NPP);
  }
  __syncthreads();
  // gather: one thread per pixel (all Q values), warps over the rows of the C planes, lanes over the columns — no
  // divisions, the Q results of a pixel stored together
  const int per_c = g.H * g.W * Q, K = g.K, C = g.C, W = g.W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int row = warp; row < C * g.H; row += DXI_THREADS / 32) {
    const int c = row / g.H, h = row - c * g.H;
    T* drow = dx + ((size_t)c * g.B + b) * per_c + (size_t)h * W * Q;
    for (int w = lane; w < W; w += 32) {
      T sacc[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) sacc[q] = T(0);
      for (int dh = 0; dh < K; ++dh) {
        const int ph = h - dh;
        if (ph < 0 || ph >= g.Ho) continue;
        for (int dw = 0; dw < K; ++dw) {
          const int pw = w - dw;
          if (pw < 0 || pw >= Wo) continue;
          const T* src = dxs + (size_t)((dh * K + dw) * C + c) * Q * NPP + ph * Wo + pw;
#pragma unroll
          for (int q = 0; q < Q; ++q) sacc[q] += src[q * NPP];
        }
      }
#pragma unroll
      for (int q = 0; q < Q; ++q) drow[w * Q + q] = sacc[q];
    }
  }
}

// The per-image fused input gradient needs the image's per-patch contributions in shared memory.  A + Bn > 32 (K = 3,
// Q = 2): the per-patch registers leave too few 256-thread CTAs per SM, the two-kernel path wins (measured).
static bool dx_fused_fits(const EpsGeom& g, size_t es) {
  const size_t fsm = ((size_t)g.O * g.A * ((g.Bn + 3) & ~3) + (size_t)g.n * g.Q * ((g.Ho * g.Wo) | 1)) * es;
  return fsm <= 200 * 1024 && g.A + g.Bn <= 32 && !getenv("DCTN_B200_DX_UNFUSED");   // env: A/B switch to the two-kernel path
}

template <typename T, int Q, int MA, int MB>
int launch_direct_bwd(const EpsGeom& g, int kind, const T* x, const T* core, const T* gout, T* result, void* ws, cudaStream_t st) {
  constexpr int A = IPow<Q, MA>::v, BN = IPow<Q, MB>::v;
  long long blocks = (g.P + DTHREADS - 1) / DTHREADS;
  const long long cap = 148ll * 8;
  if (blocks > cap) blocks = cap;
  if (kind == 1) {
    const int DO = A * BN * g.O;
    if (!getenv("DCTN_B200_DCORE_SHUFFLE")) {   // A/B switch: the older warp-shuffle kernel
      int nb = 0;
      int rc = try_launch_dcore_reg<T, Q, MA, MB>(g, x, gout, (T*)ws, &nb, st);
      if (rc == 0) rc = try_launch_dcore_tiled<T, Q, MA, MB>(g, x, gout, (T*)ws, &nb, st);
      if (rc < 0) return rc;
      if (rc == 1) return launch_reduce_partials_wide<T>((const T*)ws, result, DO, nb, st);
    }
    const size_t smem = (size_t)(DTHREADS / 32) * DO * sizeof(T);
    auto k = direct_dcore_kernel<T, Q, MA, MB>;
    DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<(unsigned)blocks, DTHREADS, smem, st>>>(g, x, gout, (T*)ws);
    dctn_count_launch();
    DCTN_CUDA_CHECK_RET(cudaGetLastError());
    return launch_reduce_partials_wide<T>((const T*)ws, result, DO, (int)blocks, st);
  }
  if (dx_fused_fits(g, sizeof(T))) {
    const int npatch = g.Ho * g.Wo, NPP = npatch | 1;
    const size_t fsm = ((size_t)g.O * A * ((BN + 3) & ~3) + (size_t)(MA + MB) * Q * NPP) * sizeof(T);
    {
      auto kf = direct_dx_image_kernel<T, Q, MA, MB>;
      DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
      kf<<<(unsigned)g.B, DXI_THREADS, fsm, st>>>(g, x, core, gout, result, NPP);
      dctn_count_launch();
      DCTN_CUDA_CHECK_RET(cudaGetLastError());
      return 0;
    }
  }
  const size_t smem = (size_t)g.O * A * ((BN + 3) & ~3) * sizeof(T);
  auto k = direct_dx_kernel<T, Q, MA, MB>;
  DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<(unsigned)blocks, DTHREADS, smem, st>>>(g, x, core, gout, (T*)ws);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return launch_gather_dx<T>(g, (const T*)ws, result, st);
}

struct DirectShape { int Q, MA, MB; };
constexpr DirectShape kShapes[] = {{2, 2, 2}, {3, 2, 2}, {4, 2, 2}, {5, 2, 2}, {6, 2, 2}, {2, 5, 4}, {2, 1, 1}, {3, 1, 1}, {4, 1, 1}};

template <typename T, int Q, int MA, int MB>
int launch_direct(const EpsGeom& g, const T* x, const T* core, T* out, cudaStream_t st) {
  constexpr int A = IPow<Q, MA>::v, BN = IPow<Q, MB>::v;
  constexpr int V = 16 / sizeof(T);
  constexpr int BNP = (BN + V - 1) / V * V;
  const size_t smem = (size_t)g.O * A * BNP * sizeof(T);
  auto k = direct_fwd_kernel<T, Q, MA, MB>;
  DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long blocks = (g.P + DTHREADS - 1) / DTHREADS;
  const long long cap = 148ll * 16;                  // a few CTAs per SM, grid-stride over the rest
  if (blocks > cap) blocks = cap;
  k<<<(unsigned)blocks, DTHREADS, smem, st>>>(g, x, core, out);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}

}  // namespace

#if DCTN_DIRECT_PART == 1
bool direct_supported(const EpsGeom& g, int dtype) {
  const size_t es = dtype == 0 ? 4 : 8;
  if ((size_t)g.A * (g.Bn + 4) * g.O * es > 36 * 1024) return false;   // transposed core in shared memory, >= 6 CTAs per SM
  if (g.P * g.O >= (1ll << 31) || (long long)g.C * g.B * g.H * g.W * g.Q >= (1ll << 31)) return false;  // 32-bit index math
  for (const DirectShape& s : kShapes)
    if (s.Q == g.Q && s.MA == g.m && s.MB == g.n - g.m) return true;
  return false;
}

#endif
#if DCTN_DIRECT_PART == 1 || DCTN_DIRECT_PART == 2
template <typename T>
int direct_forward(const EpsGeom& g, const T* x, const T* core, T* out, cudaStream_t st) {
  if constexpr (std::is_same<T, float>::value)
    if (stream_k2q2_enabled() && stream_k2q2_supported(g, 0)) return stream_k2q2_forward(g, x, core, out, st);
  if (g.K == 2 && g.C == 1) {
    if (g.Q == 2) return dispatch_direct_k2<T, 2>(g, x, core, out, st);
    if (g.Q == 3) return dispatch_direct_k2<T, 3>(g, x, core, out, st);
  }
#define DCTN_DIRECT_CASE(q, ma, mb) \
  if (g.Q == q && g.m == ma && g.n - g.m == mb) return launch_direct<T, q, ma, mb>(g, x, core, out, st);
  DCTN_DIRECT_CASE(2, 2, 2) DCTN_DIRECT_CASE(3, 2, 2) DCTN_DIRECT_CASE(4, 2, 2) DCTN_DIRECT_CASE(5, 2, 2)
  DCTN_DIRECT_CASE(6, 2, 2) DCTN_DIRECT_CASE(2, 5, 4) DCTN_DIRECT_CASE(2, 1, 1) DCTN_DIRECT_CASE(3, 1, 1)
  DCTN_DIRECT_CASE(4, 1, 1)
#undef DCTN_DIRECT_CASE
  return dctn_set_error(-2, "direct forward kernel: no instance for Q=%d with %d+%d factors", g.Q, g.m, g.n - g.m);
}
// phi fused into the forward (raw pixels in): K = 2, C = 1, Q = 2 — the HBM-bound first layer of config 1
#if DCTN_DIRECT_PART == 1
bool direct_pixels_supported(const EpsGeom& g, int dtype) { return g.K == 2 && g.C == 1 && g.Q == 2 && direct_supported(g, dtype); }
#endif
template <typename T>
int direct_forward_pixels(const EpsGeom& g, const T* pixels, T scale, const T* core, T* out, cudaStream_t st) {
  if constexpr (std::is_same<T, float>::value)
    if (stream_k2q2_enabled() && stream_k2q2_supported(g, 0)) return stream_k2q2_forward_pixels(g, pixels, scale, core, out, st);
  switch (g.O) {
    case 2: return launch_direct_k2<T, 2, 2, true>(g, pixels, core, out, st, scale);
    case 4: return launch_direct_k2<T, 2, 4, true>(g, pixels, core, out, st, scale);
    case 6: return launch_direct_k2<T, 2, 6, true>(g, pixels, core, out, st, scale);
    default: return launch_direct_k2<T, 2, 0, true>(g, pixels, core, out, st, scale);
  }
}
#if DCTN_DIRECT_PART == 1
template int direct_forward_pixels<float>(const EpsGeom&, const float*, float, const float*, float*, cudaStream_t);
template int direct_forward<float>(const EpsGeom&, const float*, const float*, float*, cudaStream_t);
#else
template int direct_forward_pixels<double>(const EpsGeom&, const double*, double, const double*, double*, cudaStream_t);
template int direct_forward<double>(const EpsGeom&, const double*, const double*, double*, cudaStream_t);
#endif
#endif  // forward parts

// ---- backward entry points (core gradient: kind 1, input gradient: kind 2)
#if DCTN_DIRECT_PART == 3
static bool direct_shape_known(const EpsGeom& g) {
  if (g.P * g.O >= (1ll << 31) || (long long)g.C * g.B * g.H * g.W * g.Q >= (1ll << 31)) return false;  // 32-bit index math
  for (const DirectShape& s : kShapes)
    if (s.Q == g.Q && s.MA == g.m && s.MB == g.n - g.m) return true;
  return false;
}
bool direct_bwd_supported(const EpsGeom& g, int dtype, int kind) {
  const long long DO = (long long)g.A * g.Bn * g.O;
  if (kind == 1) {
    // the tiled core gradient does not keep the core on chip: any Q_out whose tables fit (e.g. CIFAR (2, 6 -> 24))
    if (!direct_shape_known(g)) return false;
    if (!getenv("DCTN_B200_DCORE_SHUFFLE") && dcore_tiled_fits(g.A, g.Bn, g.O, dtype == 0 ? 4 : 8)) return true;
    return direct_supported(g, dtype) && DO <= 2048;      // (fallback) one shuffle-reduction per core element per warp iteration
  }
  if (!direct_supported(g, dtype)) return false;
  return g.A + g.Bn <= 64 && g.P * g.n * g.Q < (1ll << 31);  // everything of a patch stays in registers
}
size_t direct_workspace_bytes(const EpsGeom& g, int dtype, int kind) {
  const size_t es = dtype == 0 ? 4 : 8;
  if (kind == 1) return (size_t)148 * 8 * g.A * g.Bn * g.O * es + 256;   // one partial per CTA (x) of the core-gradient kernels
  if (kind == 2) return dx_fused_fits(g, es) ? 256 : (size_t)g.P * g.n * g.Q * es + 256;
  return 256;
}
#endif
#if DCTN_DIRECT_PART == 3 || DCTN_DIRECT_PART == 4
template <typename T>
int direct_backward(const EpsGeom& g, int kind, const T* x, const T* core, const T* gout, T* result, void* ws, cudaStream_t st) {
  if constexpr (std::is_same<T, float>::value)
    if (stream_k2q2_enabled() && stream_k2q2_bwd_supported(g, 0, kind)) return stream_k2q2_backward(g, kind, x, core, gout, result, ws, st);
#define DCTN_DIRECT_CASE(q, ma, mb) \
  if (g.Q == q && g.m == ma && g.n - g.m == mb) return launch_direct_bwd<T, q, ma, mb>(g, kind, x, core, gout, result, ws, st);
  DCTN_DIRECT_CASE(2, 2, 2) DCTN_DIRECT_CASE(3, 2, 2) DCTN_DIRECT_CASE(4, 2, 2) DCTN_DIRECT_CASE(5, 2, 2)
  DCTN_DIRECT_CASE(6, 2, 2) DCTN_DIRECT_CASE(2, 5, 4) DCTN_DIRECT_CASE(2, 1, 1) DCTN_DIRECT_CASE(3, 1, 1)
  DCTN_DIRECT_CASE(4, 1, 1)
#undef DCTN_DIRECT_CASE
  return dctn_set_error(-2, "direct backward kernel: no instance for Q=%d with %d+%d factors", g.Q, g.m, g.n - g.m);
}
#if DCTN_DIRECT_PART == 3
template int direct_backward<float>(const EpsGeom&, int, const float*, const float*, const float*, float*, void*, cudaStream_t);
#else
template int direct_backward<double>(const EpsGeom&, int, const double*, const double*, const double*, double*, void*, cudaStream_t);
#endif
#endif  // backward parts
