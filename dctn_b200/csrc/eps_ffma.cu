// CUDA-core (FFMA/DFMA) kernel family for the EPS contraction: exact fp32 / fp64 arithmetic, any shape.
//
// All three contractions of the path are GEMMs whose "activation" operand is a Khatri-Rao product
// that is GENERATED in shared memory from two small per-patch tables (common.cuh), never read from HBM:
//   forward   T[p][(b,o)]  = sum_a KR1[p][a] * core[a][(b,o)]            then out[p][o] = sum_b KR2[p][b] T[p][b][o]
//   dcore     dcore[a][n]  = sum_p KR1[p][a] * (KR2[p][b] gout[p][o])    split over patch ranges (split-K)
//   dinput    dKR1[p][a]   = sum_n (KR2[p][b] gout[p][o]) * core[a][n]
//             dKR2[p][b]   = sum_o T[p][b][o] gout[p][o]                  (T recomputed as in forward)
//             then per patch a leave-one-out contraction turns dKR1/dKR2 into d x_j, and a gather sums
//             the K*K overlapping patch contributions of every input pixel (deterministic, no atomics).
// References: dctn/eps.py:19-40 (forward), autograd of the same for the two gradients.
#include "common.cuh"
#include "eps_kernels.h"

namespace {

constexpr int KC = 16;       // k-chunk of the tiled GEMMs
constexpr int NTHREADS = 256;

template <typename T, int V> struct VecLoad;
template <> struct VecLoad<float, 4> {
  static __device__ __forceinline__ void ld(float* dst, const float* src) {
    float4 v = *reinterpret_cast<const float4*>(src);
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
  }
};
template <> struct VecLoad<float, 2> {
  static __device__ __forceinline__ void ld(float* dst, const float* src) {
    float2 v = *reinterpret_cast<const float2*>(src);
    dst[0] = v.x; dst[1] = v.y;
  }
};
template <> struct VecLoad<double, 2> {
  static __device__ __forceinline__ void ld(double* dst, const double* src) {
    double2 v = *reinterpret_cast<const double2*>(src);
    dst[0] = v.x; dst[1] = v.y;
  }
};

template <typename T, int TM, int TN, int MT, int NT> struct Tile {
  static_assert((TM / MT) * (TN / NT) == NTHREADS, "tile/thread mismatch");
  static constexpr int VMAX = 16 / sizeof(T);               // elements per 128-bit shared load
  static constexpr int VM = MT < VMAX ? MT : VMAX;
  static constexpr int VN = NT < VMAX ? NT : VMAX;
  static constexpr int GM = MT / VM;
  static constexpr int GN = NT / VN;
  static constexpr int BPAD = 16 / sizeof(T);
  static constexpr int BS = TN + BPAD;                      // Bs row stride
  static constexpr int TXN = TN / NT;                       // threads along n
  static __device__ __forceinline__ int row(int ty, int i) { return (i / VM) * (TM / GM) + ty * VM + (i % VM); }
  static __device__ __forceinline__ int col(int tx, int j) { return (j / VN) * (TN / GN) + tx * VN + (j % VN); }
  static __device__ __forceinline__ void mma_chunk(const T* As, const T* Bs, int ty, int tx, T (&acc)[MT][NT]) {
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      T av[MT], bv[NT];
#pragma unroll
      for (int gi = 0; gi < GM; ++gi) VecLoad<T, VM>::ld(&av[gi * VM], &As[kk * TM + gi * (TM / GM) + ty * VM]);
#pragma unroll
      for (int gj = 0; gj < GN; ++gj) VecLoad<T, VN>::ld(&bv[gj * VN], &Bs[kk * BS + gj * (TN / GN) + tx * VN]);
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
  }
};

// ------------------------------------------------------------------------------------------------
// C[pl][c] = sum_k Gen[pl][k] * B(k, c),  Gen[pl][k] = tabH[pl][k / KL] * tabL[pl][k % KL]
// ------------------------------------------------------------------------------------------------
template <typename T> struct GenGemmArgs {
  EpsGeom g;
  const T* x;
  const T* gout;  // nullptr, or [P][O]: folded into the lo table (entry = e*O + o)
  long long p0;   // first patch of the chunk
  int np;         // patches in the chunk
  int jh0, cnth, KH;  // hi table: factors [jh0, jh0+cnth), KH = Q^cnth
  int cntl, KLb;      // lo table: factors [jh0+cnth, jh0+cnth+cntl), KLb = Q^cntl
  int KL, Kdim;       // KL = KLb * (gout ? O : 1); Kdim = KH*KL
  const T* Bm;
  long long ldb;      // B(k,c) = TRANSB ? Bm[c*ldb + k] : Bm[k*ldb + c]
  int Ncols;
  T* Cout;            // [np][ldc]
  long long ldc;
};

template <typename T, int TM, int TN, int MT, int NT, bool TRANSB>
__global__ void __launch_bounds__(NTHREADS) gen_gemm_kernel(const __grid_constant__ GenGemmArgs<T> a) {
  using TL = Tile<T, TM, TN, MT, NT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const EpsGeom& g = a.g;
  const int Q = g.Q, O = g.O;
  const int nf = a.cnth + a.cntl;
  const int xs_stride = (nf * Q) | 1;
  T* As = reinterpret_cast<T*>(smem_raw);
  T* Bs = As + KC * TM;
  T* tabH = Bs + KC * TL::BS;
  T* tabL = tabH + a.KH * TM;
  T* xs = tabL + a.KL * TM;
  T* gs = xs + TM * xs_stride;

  const int tid = threadIdx.x;
  const int tx = tid % TL::TXN, ty = tid / TL::TXN;
  const int pl0 = blockIdx.x * TM;           // first patch of this tile, relative to the chunk
  const long long pt0 = a.p0 + pl0;          // absolute
  const int n0 = blockIdx.y * TN;

  stage_x(xs, xs_stride, a.x, g, pt0, TM, a.jh0, nf);
  if (a.gout) {
    for (int idx = tid; idx < TM * O; idx += NTHREADS) {
      long long p = pt0 + idx / O;
      gs[idx] = (p < g.P) ? a.gout[p * O + (idx % O)] : T(0);
    }
  }
  __syncthreads();
  build_table(tabH, 1, TM, xs, xs_stride, 0, a.cnth, a.KH, Q, (const T*)nullptr, O, TM);
  build_table(tabL, 1, TM, xs, xs_stride, a.cnth, a.cntl, a.KLb, Q, a.gout ? gs : (const T*)nullptr, O, TM);
  __syncthreads();

  T acc[MT][NT];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j] = T(0);

  const int ml = tid % TM;
  const int kkA0 = tid / TM;
  constexpr int stepA = NTHREADS / TM;
  const int KL = a.KL, Kdim = a.Kdim;

  for (int k0 = 0; k0 < Kdim; k0 += KC) {
    {  // generate the A chunk
      int kh = k0 / KL;
      int kl = k0 - kh * KL + kkA0;
#pragma unroll
      for (int kk = kkA0; kk < KC; kk += stepA) {
        while (kl >= KL) { kl -= KL; ++kh; }
        T v = T(0);
        if (k0 + kk < Kdim) v = tabH[kh * TM + ml] * tabL[kl * TM + ml];
        As[kk * TM + ml] = v;
        kl += stepA;
      }
    }
    if (!TRANSB) {
      const int nl = tid % TN;
      constexpr int stepB = NTHREADS / TN;
      const int c = n0 + nl;
#pragma unroll
      for (int kk = tid / TN; kk < KC; kk += stepB) {
        int k = k0 + kk;
        T v = T(0);
        if (k < Kdim && c < a.Ncols) v = __ldg(&a.Bm[(long long)k * a.ldb + c]);
        Bs[kk * TL::BS + nl] = v;
      }
    } else {
      const int kk = tid % KC;
      constexpr int stepB = NTHREADS / KC;
      const int k = k0 + kk;
#pragma unroll
      for (int nl = tid / KC; nl < TN; nl += stepB) {
        int c = n0 + nl;
        T v = T(0);
        if (k < Kdim && c < a.Ncols) v = __ldg(&a.Bm[(long long)c * a.ldb + k]);
        Bs[kk * TL::BS + nl] = v;
      }
    }
    __syncthreads();
    TL::mma_chunk(As, Bs, ty, tx, acc);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < MT; ++i) {
    int pl = pl0 + TL::row(ty, i);
    if (pl >= a.np) continue;
    T* crow = a.Cout + (long long)pl * a.ldc;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      int c = n0 + TL::col(tx, j);
      if (c < a.Ncols) crow[c] = acc[i][j];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// split-K core gradient: part[z][a][n] = sum_{p in range z} KR1[p][a] * KR2[p][b(n)] * gout[p][o(n)]
// ------------------------------------------------------------------------------------------------
template <typename T> struct DcoreArgs {
  EpsGeom g;
  const T* x;
  const T* gout;
  T* part;              // [splits][A][N]
  long long per_split;  // patches per split (multiple of KC)
};

template <typename T, int TM, int TN, int MT, int NT>
__global__ void __launch_bounds__(NTHREADS) dcore_kernel(const __grid_constant__ DcoreArgs<T> a) {
  using TL = Tile<T, TM, TN, MT, NT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const EpsGeom& g = a.g;
  const int Q = g.Q, O = g.O;
  const int xs_stride = (g.n * Q) | 1;
  const int BLO = g.BL * O;
  T* As = reinterpret_cast<T*>(smem_raw);
  T* Bs = As + KC * TM;
  T* tAH = Bs + KC * TL::BS;      // [KC][AH]
  T* tAL = tAH + KC * g.AH;       // [KC][AL]
  T* tBH = tAL + KC * g.AL;       // [KC][BH]
  T* tBL = tBH + KC * g.BH;       // [KC][BL*O]
  T* xs = tBL + KC * BLO;         // [KC][xs_stride]
  T* gs = xs + KC * xs_stride;    // [KC][O]

  const int tid = threadIdx.x;
  const int tx = tid % TL::TXN, ty = tid / TL::TXN;
  const int a0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  long long pbeg = (long long)blockIdx.z * a.per_split;
  long long pend = pbeg + a.per_split;
  if (pend > g.P) pend = g.P;

  // the column each thread generates (fixed for the whole kernel)
  const int colA = tid % TM, kkA0 = tid / TM;
  constexpr int stepA = NTHREADS / TM;
  const int aidx = a0 + colA;
  const bool okA = aidx < g.A;
  const int ahi = okA ? aidx / g.AL : 0, alo = okA ? aidx % g.AL : 0;
  const int colB = tid % TN, kkB0 = tid / TN;
  constexpr int stepB = NTHREADS / TN;
  const int nidx = n0 + colB;
  const bool okB = nidx < g.N;
  const int bhi = okB ? nidx / BLO : 0, blo = okB ? nidx % BLO : 0;

  T acc[MT][NT];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j] = T(0);

  for (long long pc = pbeg; pc < pend; pc += KC) {
    // stage x and gout of KC patches (zeros past the end of this split's range)
    stage_x(xs, xs_stride, a.x, g, pc, KC, 0, g.n);
    for (int idx = tid; idx < KC * O; idx += NTHREADS) {
      long long p = pc + idx / O;
      gs[idx] = (p < pend) ? a.gout[p * O + (idx % O)] : T(0);
    }
    __syncthreads();
    build_table(tAH, g.AH, 1, xs, xs_stride, 0, g.a_nh, g.AH, Q, (const T*)nullptr, O, KC);
    build_table(tAL, g.AL, 1, xs, xs_stride, g.a_nh, g.a_nl, g.AL, Q, (const T*)nullptr, O, KC);
    build_table(tBH, g.BH, 1, xs, xs_stride, g.m, g.b_nh, g.BH, Q, (const T*)nullptr, O, KC);
    build_table(tBL, BLO, 1, xs, xs_stride, g.m + g.b_nh, g.b_nl, g.BL, Q, gs, O, KC);
    __syncthreads();
#pragma unroll
    for (int kk = kkA0; kk < KC; kk += stepA)
      As[kk * TM + colA] = okA ? tAH[kk * g.AH + ahi] * tAL[kk * g.AL + alo] : T(0);
#pragma unroll
    for (int kk = kkB0; kk < KC; kk += stepB)
      Bs[kk * TL::BS + colB] = okB ? tBH[kk * g.BH + bhi] * tBL[kk * BLO + blo] : T(0);
    __syncthreads();
    TL::mma_chunk(As, Bs, ty, tx, acc);
    // the next iteration's first barrier (after staging) protects As/Bs; xs/gs are not read here
  }

  T* part = a.part + (long long)blockIdx.z * g.A * g.N;
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    int ar = a0 + TL::row(ty, i);
    if (ar >= g.A) continue;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      int c = n0 + TL::col(tx, j);
      if (c < g.N) part[(long long)ar * g.N + c] = acc[i][j];
    }
  }
}

template <typename T>
__global__ void reduce_partials_kernel(const T* __restrict__ part, T* __restrict__ out, long long count, int splits) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < count; i += stride) {
    T s = T(0);
    for (int z = 0; z < splits; ++z) s += part[(long long)z * count + i];
    out[i] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// per-patch post kernels
// ------------------------------------------------------------------------------------------------
constexpr int PT = 32;  // patches per CTA in the post kernels

// out[p][o] = sum_b T[pl][b*O + o] * KR2[p][b]        (forward epilogue for the FFMA family)
template <typename T>
__global__ void __launch_bounds__(NTHREADS) fwd_post_kernel(EpsGeom g, const T* __restrict__ x,
                                                            const T* __restrict__ Tws, long long p0, int np,
                                                            T* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Q = g.Q, O = g.O;
  const int nf = g.n - g.m;
  const int xs_stride = (nf * Q) | 1;
  T* tH = reinterpret_cast<T*>(smem_raw);  // [PT][BH]
  T* tL = tH + PT * g.BH;                  // [PT][BL]
  T* xs = tL + PT * g.BL;
  const int pl0 = blockIdx.x * PT;
  stage_x(xs, xs_stride, x, g, p0 + pl0, PT, g.m, nf);
  __syncthreads();
  build_table(tH, g.BH, 1, xs, xs_stride, 0, g.b_nh, g.BH, Q, (const T*)nullptr, O, PT);
  build_table(tL, g.BL, 1, xs, xs_stride, g.b_nh, g.b_nl, g.BL, Q, (const T*)nullptr, O, PT);
  __syncthreads();
  for (int item = threadIdx.x; item < PT * O; item += NTHREADS) {
    int pl = item / O, o = item - pl * O;
    int plc = pl0 + pl;
    if (plc >= np) continue;
    const T* trow = Tws + (long long)plc * g.N + o;
    const T* th = tH + pl * g.BH;
    const T* tl = tL + pl * g.BL;
    T s = T(0);
    for (int bh = 0; bh < g.BH; ++bh) {
      T hv = th[bh];
      T s2 = T(0);
      const T* tr2 = trow + (long long)bh * g.BL * O;
      for (int bl = 0; bl < g.BL; ++bl) s2 = fma(tr2[bl * O], tl[bl], s2);
      s = fma(hv, s2, s);
    }
    out[(p0 + plc) * O + o] = s;
  }
}

// dKR2[pl][b] = sum_o T[pl][b*O + o] * gout[p][o]
template <typename T>
__global__ void dkr2_post_kernel(EpsGeom g, const T* __restrict__ gout, const T* __restrict__ Tws, long long p0,
                                 int np, T* __restrict__ dkr2) {
  long long total = (long long)np * g.Bn;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int pl = (int)(i / g.Bn);
    int b = (int)(i - (long long)pl * g.Bn);
    const T* t = Tws + (long long)pl * g.N + (long long)b * g.O;
    const T* gr = gout + (p0 + pl) * g.O;
    T s = T(0);
    for (int o = 0; o < g.O; ++o) s = fma(t[o], gr[o], s);
    dkr2[i] = s;
  }
}

// Leave-one-out stage: turns dKR (gradient w.r.t. one Khatri-Rao half, [np][E = EH*EL]) into
// d x_j for the factors j of that half:   dxp[p][j][q].
//   Wlo[el] = sum_eh dKR[eh*EL + el] * tabH[eh];   Whi[eh] = sum_el dKR[eh*EL + el] * tabL[el]
//   d x_j[q] (j at position t of a group with table entries e) = sum_{e: digit_t(e)=q} W[e] * prod_{t'!=t} x_{j'}[digit_t'(e)]
template <typename T>
__global__ void __launch_bounds__(NTHREADS) loo_kernel(EpsGeom g, const T* __restrict__ x, const T* __restrict__ dkr,
                                                       long long p0, int np, int j0, int cnth, int EH, int cntl,
                                                       int EL, T* __restrict__ dxp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Q = g.Q;
  const int nf = cnth + cntl;
  const int xs_stride = (nf * Q) | 1;
  const int E = EH * EL;
  T* tH = reinterpret_cast<T*>(smem_raw);  // [PT][EH]
  T* tL = tH + PT * EH;                    // [PT][EL]
  T* wH = tL + PT * EL;                    // [PT][EH]
  T* wL = wH + PT * EH;                    // [PT][EL]
  T* xs = wL + PT * EL;
  const int pl0 = blockIdx.x * PT;
  stage_x(xs, xs_stride, x, g, p0 + pl0, PT, j0, nf);
  __syncthreads();
  build_table(tH, EH, 1, xs, xs_stride, 0, cnth, EH, Q, (const T*)nullptr, 1, PT);
  build_table(tL, EL, 1, xs, xs_stride, cnth, cntl, EL, Q, (const T*)nullptr, 1, PT);
  __syncthreads();
  // Wlo: item = (pl, el), lanes run over el (contiguous reads)
  for (int item = threadIdx.x; item < PT * EL; item += NTHREADS) {
    int pl = item / EL, el = item - pl * EL;
    T s = T(0);
    if (pl0 + pl < np) {
      const T* d = dkr + (long long)(pl0 + pl) * E + el;
      const T* th = tH + pl * EH;
      for (int eh = 0; eh < EH; ++eh) s = fma(d[(long long)eh * EL], th[eh], s);
    }
    wL[item] = s;
  }
  for (int item = threadIdx.x; item < PT * EH; item += NTHREADS) {
    int pl = item / EH, eh = item - pl * EH;
    T s = T(0);
    if (pl0 + pl < np) {
      const T* d = dkr + (long long)(pl0 + pl) * E + (long long)eh * EL;
      const T* tl = tL + pl * EL;
      for (int el = 0; el < EL; ++el) s = fma(d[el], tl[el], s);
    }
    wH[item] = s;
  }
  __syncthreads();
  // final: item = (pl, t, q) over the nf factors of this half
  for (int item = threadIdx.x; item < PT * nf * Q; item += NTHREADS) {
    int pl = item / (nf * Q);
    int r = item - pl * nf * Q;
    int t = r / Q, q = r - t * Q;
    if (pl0 + pl >= np) continue;
    const bool in_hi = t < cnth;
    const int cnt = in_hi ? cnth : cntl;
    const int tt = in_hi ? t : t - cnth;             // position inside its group
    const int Eg = in_hi ? EH : EL;
    const T* w = (in_hi ? wH + pl * EH : wL + pl * EL);
    const T* xr = xs + pl * xs_stride + (in_hi ? 0 : cnth) * Q;
    // stride of digit tt inside the group index: Q^(cnt-1-tt)
    int dstride = 1;
    for (int u = 0; u < cnt - 1 - tt; ++u) dstride *= Q;
    T s = T(0);
    int others = Eg / Q;  // number of entries with digit tt fixed
    for (int oe = 0; oe < others; ++oe) {
      // expand oe (index over the other cnt-1 digits) into the full entry index with digit tt = q
      int lo_part = oe % dstride, hi_part = oe / dstride;
      int e = (hi_part * Q + q) * dstride + lo_part;
      T v = w[e];
      int ee = e;
      for (int u = cnt - 1; u >= 0; --u) {
        int d = ee % Q;
        ee /= Q;
        if (u != tt) v *= xr[u * Q + d];
      }
      s += v;
    }
    dxp[((p0 + pl0 + pl) * g.n + (j0 + t)) * Q + q] = s;
  }
}

// Shared-memory staged version of loo_kernel: the dKR rows of LPT patches are brought in ONCE with coalesced 128-bit
// loads (row-major [EH][EL] matrix per patch, stored with an odd row stride so that both the column-wise (Wlo) and the
// row-wise (Whi) passes are bank-conflict free); everything else is as in loo_kernel.
constexpr int LPT = 8;  // patches per CTA
template <typename T>
__global__ void __launch_bounds__(NTHREADS) loo_staged_kernel(EpsGeom g, const T* __restrict__ x, const T* __restrict__ dkr,
                                                              long long p0, int np, int j0, int cnth, int EH, int cntl,
                                                              int EL, T* __restrict__ dxp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Q = g.Q;
  const int nf = cnth + cntl;
  const int xs_stride = (nf * Q) | 1;
  const int E = EH * EL;
  const int ELS = EL | 1;                  // padded row stride of the staged matrix
  T* M = reinterpret_cast<T*>(smem_raw);   // [LPT][EH][ELS]
  T* tH = M + LPT * EH * ELS;              // [LPT][EH]
  T* tL = tH + LPT * EH;                   // [LPT][EL]
  T* wH = tL + LPT * EL;                   // [LPT][EH]
  T* wL = wH + LPT * EH;                   // [LPT][EL]
  T* xs = wL + LPT * EL;
  int* qd = reinterpret_cast<int*>(xs + LPT * xs_stride);   // qd[e] = e / Q, e < max(EH, EL, nf*Q) + 1
  {
    int nqd = EH > EL ? EH : EL;
    if (nf * Q > nqd) nqd = nf * Q;
    for (int e = threadIdx.x; e <= nqd; e += NTHREADS) qd[e] = e / Q;   // visible after the barrier below
  }
  const int pl0 = blockIdx.x * LPT;
  const int npl = (np - pl0 < LPT) ? (np - pl0) : LPT;
  // (1) stage the rows: contiguous npl*E elements of dkr starting at pl0*E
  {
    const T* src = dkr + (long long)pl0 * E;
    const int total = npl * E;
    constexpr int V = 16 / sizeof(T);
    if ((E % V) == 0) {
      for (int i = threadIdx.x * V; i < total; i += NTHREADS * V) {
        T v[V];
        if constexpr (sizeof(T) == 4) {
          const float4 f = *reinterpret_cast<const float4*>(src + i);
          v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
        } else {
          const double2 f = *reinterpret_cast<const double2*>(src + i);
          v[0] = f.x; v[1] = f.y;
        }
        // E % V == 0: the V elements stay inside one patch; one (patch, row, column) decode per vector, then a walk
        const int pl = i / E, e = i - pl * E;
        const int eh = e / EL;
        int el = e - eh * EL;
        T* dst = M + (pl * EH + eh) * ELS + el;
#pragma unroll
        for (int u = 0; u < V; ++u) {
          *dst++ = v[u];
          if (++el == EL) { el = 0; dst += ELS - EL; }
        }
      }
    } else {
      for (int i = threadIdx.x; i < total; i += NTHREADS) {
        const int e = i % E, pl = i / E;
        M[(pl * EH + e / EL) * ELS + e % EL] = src[i];
      }
    }
  }
  stage_x(xs, xs_stride, x, g, p0 + pl0, LPT, j0, nf);
  __syncthreads();
  build_table(tH, EH, 1, xs, xs_stride, 0, cnth, EH, Q, (const T*)nullptr, 1, LPT);
  build_table(tL, EL, 1, xs, xs_stride, cnth, cntl, EL, Q, (const T*)nullptr, 1, LPT);
  __syncthreads();
  for (int item = threadIdx.x; item < LPT * EL; item += NTHREADS) {
    const int pl = item / EL, el = item - pl * EL;
    T s = T(0);
    if (pl < npl) {
      const T* m = M + pl * EH * ELS + el;
      const T* th = tH + pl * EH;
      for (int eh = 0; eh < EH; ++eh) s = fma(m[eh * ELS], th[eh], s);
    }
    wL[item] = s;
  }
  for (int item = threadIdx.x; item < LPT * EH; item += NTHREADS) {
    const int pl = item / EH, eh = item - pl * EH;
    T s = T(0);
    if (pl < npl) {
      const T* m = M + (pl * EH + eh) * ELS;
      const T* tl = tL + pl * EL;
      for (int el = 0; el < EL; ++el) s = fma(m[el], tl[el], s);
    }
    wH[item] = s;
  }
  __syncthreads();
  // per-group stage without divisions: qd[e] = e / Q (built above), nested walks over the digits before / after position tt
  for (int item = threadIdx.x; item < LPT * nf * Q; item += NTHREADS) {
    const int pl = item / (nf * Q);
    const int r = item - pl * nf * Q;
    const int t = qd[r], q = r - t * Q;
    if (pl >= npl) continue;
    const bool in_hi = t < cnth;
    const int cnt = in_hi ? cnth : cntl;
    const int tt = in_hi ? t : t - cnth;
    const T* w = (in_hi ? wH + pl * EH : wL + pl * EL);
    const T* xr = xs + pl * xs_stride + (in_hi ? 0 : cnth) * Q;
    int dstride = 1, npre = 1;
    for (int u = 0; u < cnt - 1 - tt; ++u) dstride *= Q;
    for (int u = 0; u < tt; ++u) npre *= Q;
    T s = T(0);
    for (int hp = 0; hp < npre; ++hp) {
      const int e0 = (hp * Q + q) * dstride;
      for (int lp = 0; lp < dstride; ++lp) {
        T v = w[e0 + lp];
        int ee = e0 + lp;
        for (int u = cnt - 1; u >= 0; --u) {
          const int e1 = qd[ee], d = ee - e1 * Q;
          ee = e1;
          if (u != tt) v *= xr[u * Q + d];
        }
        s += v;
      }
    }
    dxp[((p0 + pl0 + pl) * g.n + (j0 + t)) * Q + q] = s;
  }
}

// dx[c][b][h][w][q] = sum over the patches that contain pixel (h, w) of dxp[p][j(dh,dw,c)][q]
template <typename T>
__global__ void gather_dx_kernel(EpsGeom g, const T* __restrict__ dxp, T* __restrict__ dx) {
  long long total = (long long)g.C * g.B * g.H * g.W * g.Q;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    int q = (int)(r % g.Q); r /= g.Q;
    int w = (int)(r % g.W); r /= g.W;
    int h = (int)(r % g.H); r /= g.H;
    int b = (int)(r % g.B); r /= g.B;
    int c = (int)r;
    T s = T(0);
    for (int dh = 0; dh < g.K; ++dh) {
      int ph = h - dh;
      if (ph < 0 || ph >= g.Ho) continue;
      for (int dw = 0; dw < g.K; ++dw) {
        int pw = w - dw;
        if (pw < 0 || pw >= g.Wo) continue;
        long long p = ((long long)b * g.Ho + ph) * g.Wo + pw;
        int j = (dh * g.K + dw) * g.C + c;
        s += dxp[(p * g.n + j) * g.Q + q];
      }
    }
    dx[i] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
template <typename T> struct Cfg;
template <> struct Cfg<float> {
  static constexpr int TM_BIG = 128, TM_SMALL = 64, TN = 128, MT_BIG = 8, MT_SMALL = 4, NT = 8;
};
template <> struct Cfg<double> {
  static constexpr int TM_BIG = 64, TM_SMALL = 32, TN = 64, MT_BIG = 4, MT_SMALL = 2, NT = 4;
};

constexpr size_t SMEM_LIMIT = 200 * 1024;

template <typename T, int TM, int TN>
size_t gen_gemm_smem(const GenGemmArgs<T>& a) {
  int nf = a.cnth + a.cntl;
  int xs_stride = (nf * a.g.Q) | 1;
  size_t el = (size_t)KC * TM + (size_t)KC * (TN + 16 / sizeof(T)) + (size_t)(a.KH + a.KL) * TM + (size_t)TM * xs_stride +
              (a.gout ? (size_t)TM * a.g.O : 0);
  return el * sizeof(T) + 16;
}

template <typename T, int TM, int TN, int MT, int NT>
int launch_gen_gemm_cfg(const GenGemmArgs<T>& a, bool transb, cudaStream_t st) {
  size_t smem = gen_gemm_smem<T, TM, TN>(a);
  dim3 grid((a.np + TM - 1) / TM, (a.Ncols + TN - 1) / TN);
  if (transb) {
    auto k = gen_gemm_kernel<T, TM, TN, MT, NT, true>;
    DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, NTHREADS, smem, st>>>(a);
  } else {
    auto k = gen_gemm_kernel<T, TM, TN, MT, NT, false>;
    DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, NTHREADS, smem, st>>>(a);
  }
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}

template <typename T>
int launch_gen_gemm(const GenGemmArgs<T>& a, bool transb, cudaStream_t st) {
  using C = Cfg<T>;
  if (gen_gemm_smem<T, C::TM_BIG, C::TN>(a) <= SMEM_LIMIT)
    return launch_gen_gemm_cfg<T, C::TM_BIG, C::TN, C::MT_BIG, C::NT>(a, transb, st);
  if (gen_gemm_smem<T, C::TM_SMALL, C::TN>(a) <= SMEM_LIMIT)
    return launch_gen_gemm_cfg<T, C::TM_SMALL, C::TN, C::MT_SMALL, C::NT>(a, transb, st);
  return dctn_set_error(-2, "EPS shape needs %zu bytes of shared memory for its Khatri-Rao tables (limit %zu)",
                        gen_gemm_smem<T, C::TM_SMALL, C::TN>(a), SMEM_LIMIT);
}

template <typename T>
GenGemmArgs<T> make_half1_args(const EpsGeom& g, const T* x, long long p0, int np) {
  // Gen = KR1 (first half, no gout): K = A
  GenGemmArgs<T> a{};
  a.g = g; a.x = x; a.gout = nullptr; a.p0 = p0; a.np = np;
  a.jh0 = 0; a.cnth = g.a_nh; a.KH = g.AH; a.cntl = g.a_nl; a.KLb = g.AL; a.KL = g.AL; a.Kdim = g.A;
  return a;
}

template <typename T>
size_t dcore_smem(const EpsGeom& g, int TM, int TN) {
  int xs_stride = (g.n * g.Q) | 1;
  size_t el = (size_t)KC * TM + (size_t)KC * (TN + 16 / sizeof(T)) +
              (size_t)KC * (g.AH + g.AL + g.BH + (size_t)g.BL * g.O) + (size_t)KC * xs_stride + (size_t)KC * g.O;
  return el * sizeof(T) + 16;
}

inline int pick_splits(const EpsGeom& g, int TM, int TN) {
  long long tiles = (long long)((g.A + TM - 1) / TM) * ((g.N + TN - 1) / TN);
  long long want = (2 * 148 + tiles - 1) / tiles;
  long long max_by_p = (g.P + 4 * KC - 1) / (4 * KC);
  long long s = want < max_by_p ? want : max_by_p;
  if (s < 1) s = 1;
  if (s > 1024) s = 1024;
  return (int)s;
}

template <typename T> long long post_patch_chunk(const EpsGeom& g, long long row_elems) {
  // patches per chunk so that the per-chunk scratch stays around 96 MB (L2-resident on B200)
  long long target = 96ll << 20;
  long long pc = target / (row_elems * (long long)sizeof(T));
  if (pc < 4096) pc = 4096;
  pc = (pc / 128) * 128;
  if (pc > g.P) pc = g.P;
  return pc;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// public launchers (declared in eps_kernels.h)
// ------------------------------------------------------------------------------------------------
template <typename T>
int launch_reduce_partials(const T* part, T* out, long long count, int splits, cudaStream_t st) {
  int blocks = (int)((count + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  reduce_partials_kernel<T><<<blocks, 256, 0, st>>>(part, out, count, splits);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}
template int launch_reduce_partials<float>(const float*, float*, long long, int, cudaStream_t);
template int launch_reduce_partials<double>(const double*, double*, long long, int, cudaStream_t);

// half == 0: factors [0, m) from dKR1 [np][A];  half == 1: factors [m, n) from dKR2 [np][Bn]
template <typename T>
int launch_loo(const EpsGeom& g, const T* x, const T* dkr, long long p0, int np, int half, T* dxp, cudaStream_t st) {
  const int j0 = half ? g.m : 0, cnth = half ? g.b_nh : g.a_nh, cntl = half ? g.b_nl : g.a_nl;
  const int EH = half ? g.BH : g.AH, EL = half ? g.BL : g.AL;
  const int nf = cnth + cntl;
  {
    const int nqd = (EH > EL ? (EH > nf * g.Q ? EH : nf * g.Q) : (EL > nf * g.Q ? EL : nf * g.Q)) + 1;
    const size_t ssm = ((size_t)LPT * EH * (EL | 1) + (size_t)LPT * 2 * (EH + EL) + (size_t)LPT * ((nf * g.Q) | 1)) * sizeof(T) +
                       (size_t)nqd * sizeof(int) + 16;
    if (ssm <= 100 * 1024) {   // at least two CTAs per SM
      DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(loo_staged_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));
      loo_staged_kernel<T><<<(np + LPT - 1) / LPT, NTHREADS, ssm, st>>>(g, x, dkr, p0, np, j0, cnth, EH, cntl, EL, dxp);
      dctn_count_launch();
      DCTN_CUDA_CHECK_RET(cudaGetLastError());
      return 0;
    }
  }
  size_t smem = ((size_t)PT * 2 * (EH + EL) + (size_t)PT * ((nf * g.Q) | 1)) * sizeof(T) + 16;
  if (smem > SMEM_LIMIT) return dctn_set_error(-2, "leave-one-out kernel needs %zu bytes of shared memory", smem);
  DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(loo_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  loo_kernel<T><<<(np + PT - 1) / PT, NTHREADS, smem, st>>>(g, x, dkr, p0, np, j0, cnth, EH, cntl, EL, dxp);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}
template int launch_loo<float>(const EpsGeom&, const float*, const float*, long long, int, int, float*, cudaStream_t);
template int launch_loo<double>(const EpsGeom&, const double*, const double*, long long, int, int, double*, cudaStream_t);

template <typename T>
int launch_gather_dx(const EpsGeom& g, const T* dxp, T* dx, cudaStream_t st) {
  long long total = (long long)g.C * g.B * g.H * g.W * g.Q;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  gather_dx_kernel<T><<<blocks, 256, 0, st>>>(g, dxp, dx);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}
template int launch_gather_dx<float>(const EpsGeom&, const float*, float*, cudaStream_t);
template int launch_gather_dx<double>(const EpsGeom&, const double*, double*, cudaStream_t);

template <typename T>
size_t ffma_workspace_bytes(const EpsGeom& g, int kind) {
  if (kind == 0) {
    long long pc = post_patch_chunk<T>(g, g.N);
    return (size_t)pc * g.N * sizeof(T);
  } else if (kind == 1) {
    using C = Cfg<T>;
    int splits = pick_splits(g, C::TM_BIG, C::TN);
    return (size_t)splits * g.A * g.N * sizeof(T);
  } else {
    long long pc = post_patch_chunk<T>(g, (long long)g.N + g.A + g.Bn);
    return ((size_t)pc * ((size_t)g.N + g.A + g.Bn) + (size_t)g.P * g.n * g.Q) * sizeof(T);
  }
}

template <typename T>
int ffma_forward(const EpsGeom& g, const T* x, const T* core, T* out, void* ws, cudaStream_t st) {
  T* Tws = reinterpret_cast<T*>(ws);
  long long pc = post_patch_chunk<T>(g, g.N);
  int nfb = g.n - g.m;
  size_t post_smem = ((size_t)PT * (g.BH + g.BL) + (size_t)PT * ((nfb * g.Q) | 1)) * sizeof(T) + 16;
  if (post_smem > SMEM_LIMIT) return dctn_set_error(-2, "forward post kernel needs %zu bytes of shared memory", post_smem);
  DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(fwd_post_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)post_smem));
  for (long long p0 = 0; p0 < g.P; p0 += pc) {
    int np = (int)((g.P - p0 < pc) ? (g.P - p0) : pc);
    GenGemmArgs<T> a = make_half1_args<T>(g, x, p0, np);
    a.Bm = core; a.ldb = g.N; a.Ncols = g.N; a.Cout = Tws; a.ldc = g.N;
    int rc = launch_gen_gemm<T>(a, false, st);
    if (rc) return rc;
    fwd_post_kernel<T><<<(np + PT - 1) / PT, NTHREADS, post_smem, st>>>(g, x, Tws, p0, np, out);
    dctn_count_launch();
    DCTN_CUDA_CHECK_RET(cudaGetLastError());
  }
  return 0;
}

template <typename T>
int ffma_backward_core(const EpsGeom& g, const T* x, const T* gout, T* dcore, void* ws, cudaStream_t st) {
  using C = Cfg<T>;
  constexpr int TM = C::TM_BIG, TN = C::TN;
  size_t smem = dcore_smem<T>(g, TM, TN);
  if (smem > SMEM_LIMIT) return dctn_set_error(-2, "core-gradient kernel needs %zu bytes of shared memory", smem);
  int splits = pick_splits(g, TM, TN);
  DcoreArgs<T> a{};
  a.g = g; a.x = x; a.gout = gout; a.part = reinterpret_cast<T*>(ws);
  long long per = (g.P + splits - 1) / splits;
  per = ((per + KC - 1) / KC) * KC;
  a.per_split = per;
  splits = (int)((g.P + per - 1) / per);
  auto k = dcore_kernel<T, TM, TN, C::MT_BIG, C::NT>;
  DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((g.A + TM - 1) / TM, (g.N + TN - 1) / TN, splits);
  k<<<grid, NTHREADS, smem, st>>>(a);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return launch_reduce_partials<T>(a.part, dcore, (long long)g.A * g.N, splits, st);
}

template <typename T>
int ffma_backward_input(const EpsGeom& g, const T* x, const T* core, const T* gout, T* dx, void* ws,
                        cudaStream_t st) {
  long long pc = post_patch_chunk<T>(g, (long long)g.N + g.A + g.Bn);
  T* Tws = reinterpret_cast<T*>(ws);
  T* dkr1 = Tws + (size_t)pc * g.N;
  T* dkr2 = dkr1 + (size_t)pc * g.A;
  T* dxp = dkr2 + (size_t)pc * g.Bn;
  const int nfa = g.m, nfb = g.n - g.m;
  size_t loo_smem_a = ((size_t)PT * 2 * (g.AH + g.AL) + (size_t)PT * ((nfa * g.Q) | 1)) * sizeof(T) + 16;
  size_t loo_smem_b = ((size_t)PT * 2 * (g.BH + g.BL) + (size_t)PT * ((nfb * g.Q) | 1)) * sizeof(T) + 16;
  size_t loo_smem = loo_smem_a > loo_smem_b ? loo_smem_a : loo_smem_b;
  if (loo_smem > SMEM_LIMIT) return dctn_set_error(-2, "leave-one-out kernel needs %zu bytes of shared memory", loo_smem);
  DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(loo_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)loo_smem));
  for (long long p0 = 0; p0 < g.P; p0 += pc) {
    int np = (int)((g.P - p0 < pc) ? (g.P - p0) : pc);
    // dKR1[p][a] = sum_n (KR2[p][b] gout[p][o]) core[a][n]
    {
      GenGemmArgs<T> a{};
      a.g = g; a.x = x; a.gout = gout; a.p0 = p0; a.np = np;
      a.jh0 = g.m; a.cnth = g.b_nh; a.KH = g.BH; a.cntl = g.b_nl; a.KLb = g.BL; a.KL = g.BL * g.O; a.Kdim = g.N;
      a.Bm = core; a.ldb = g.N; a.Ncols = g.A; a.Cout = dkr1; a.ldc = g.A;
      int rc = launch_gen_gemm<T>(a, true, st);
      if (rc) return rc;
      loo_kernel<T><<<(np + PT - 1) / PT, NTHREADS, loo_smem_a, st>>>(g, x, dkr1, p0, np, 0, g.a_nh, g.AH, g.a_nl, g.AL, dxp);
      dctn_count_launch();
      DCTN_CUDA_CHECK_RET(cudaGetLastError());
    }
    if (nfb > 0) {
      // T = KR1 @ core (as in forward), dKR2[p][b] = sum_o T[p][b][o] gout[p][o]
      GenGemmArgs<T> a = make_half1_args<T>(g, x, p0, np);
      a.Bm = core; a.ldb = g.N; a.Ncols = g.N; a.Cout = Tws; a.ldc = g.N;
      int rc = launch_gen_gemm<T>(a, false, st);
      if (rc) return rc;
      long long total = (long long)np * g.Bn;
      int blocks = (int)((total + 255) / 256);
      if (blocks > 148 * 16) blocks = 148 * 16;
      dkr2_post_kernel<T><<<blocks, 256, 0, st>>>(g, gout, Tws, p0, np, dkr2);
      dctn_count_launch();
      DCTN_CUDA_CHECK_RET(cudaGetLastError());
      loo_kernel<T><<<(np + PT - 1) / PT, NTHREADS, loo_smem_b, st>>>(g, x, dkr2, p0, np, g.m, g.b_nh, g.BH, g.b_nl, g.BL, dxp);
      dctn_count_launch();
      DCTN_CUDA_CHECK_RET(cudaGetLastError());
    }
  }
  long long total = (long long)g.C * g.B * g.H * g.W * g.Q;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  gather_dx_kernel<T><<<blocks, 256, 0, st>>>(g, dxp, dx);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}

template size_t ffma_workspace_bytes<float>(const EpsGeom&, int);
template size_t ffma_workspace_bytes<double>(const EpsGeom&, int);
template int ffma_forward<float>(const EpsGeom&, const float*, const float*, float*, void*, cudaStream_t);
template int ffma_forward<double>(const EpsGeom&, const double*, const double*, double*, void*, cudaStream_t);
template int ffma_backward_core<float>(const EpsGeom&, const float*, const float*, float*, void*, cudaStream_t);
template int ffma_backward_core<double>(const EpsGeom&, const double*, const double*, double*, void*, cudaStream_t);
template int ffma_backward_input<float>(const EpsGeom&, const float*, const float*, const float*, float*, void*, cudaStream_t);
template int ffma_backward_input<double>(const EpsGeom&, const double*, const double*, const double*, double*, void*, cudaStream_t);
