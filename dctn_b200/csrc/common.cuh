// Shared definitions for the dctn_b200 CUDA kernels (sm_100a).
//
// Notation: an EPS layer is (K, Q, O, C); n = K*K*C factors per patch,
// factor j = (dh*K + dw)*C + c (dctn/align.py:20-46) is the length-Q vector x[c, b, h+dh, w+dw, :].
// The core is flat [D = Q^n][O], factor 0 is the slowest index (dctn/eps.py:31-40).
// The factor list is split after the first m factors (the reference's own split, dctn/eps.py:25-30):
//   a in [0, A = Q^m)  indexes factors 0..m-1,   b in [0, Bn = Q^(n-m)) indexes factors m..n-1,
//   core[a][b][o],  N = Bn*O,  KR1[p][a] = prod_{j<m} x_j[p][a_j],  KR2[p][b] = prod_{j>=m} x_j[p][b_j].
// Each half is generated on chip from two small per-patch tables ("two-level Khatri-Rao"):
//   KR1[p][a] = TabAH[p][a / AL] * TabAL[p][a % AL],  TabAH over the first a_nh factors of the half.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DCTN_MAXN 32  // max factors per patch (K*K*C)

struct EpsGeom {
  int C, K, Q, O;
  int n, m;              // factors per patch; size of the first half
  int A, Bn, N;          // Q^m, Q^(n-m), Bn*O
  int a_nh, a_nl, AH, AL;  // first half:  hi group = factors [0, a_nh), lo group = [a_nh, m)
  int b_nh, b_nl, BH, BL;  // second half: hi group = factors [m, m+b_nh), lo group = [m+b_nh, n)
  int B, H, W, Ho, Wo;
  long long P;           // B*Ho*Wo patches
  long long foff[DCTN_MAXN];  // element offset of factor j relative to the patch-origin pixel
};

__device__ __forceinline__ long long patch_origin(const EpsGeom& g, long long p) {
  int hw = g.Ho * g.Wo;
  long long b = p / hw;
  int r = (int)(p - b * hw);
  int h = r / g.Wo;
  int w = r - h * g.Wo;
  return ((b * g.H + h) * (long long)g.W + w) * g.Q;
}

// Stage the factor vectors j in [jbeg, jbeg+nf) of `np` consecutive patches starting at p0 into
// shared memory: xs[pl*xs_stride + (j-jbeg)*Q + q].  Patches >= P are filled with zeros.
template <typename T>
__device__ __forceinline__ void stage_x(T* xs, int xs_stride, const T* __restrict__ x, const EpsGeom& g,
                                        long long p0, int np, int jbeg, int nf) {
  const int Q = g.Q;
  const int per_j = np * Q;
  const int total = nf * per_j;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    int jj = idx / per_j;
    int rem = idx - jj * per_j;
    int pl = rem / Q;
    int q = rem - pl * Q;
    long long p = p0 + pl;
    T v = T(0);
    if (p < g.P) v = x[patch_origin(g, p) + g.foff[jbeg + jj] + q];
    xs[pl * xs_stride + jj * Q + q] = v;
  }
}

// tab[eo*se + pl*sp] = (gs ? gs[pl*O + o] : 1) * prod_{t<cnt} xs[pl*xs_stride + (jrel0+t)*Q + digit_t(e)]
// where eo = e*O + o when gs != nullptr (else eo = e), and digit_0 is the slowest digit of e.
template <typename T>
__device__ __forceinline__ void build_table(T* tab, int sp, int se, const T* xs, int xs_stride, int jrel0,
                                            int cnt, int E, int Q, const T* gs, int O, int np) {
  const int EO = gs ? E * O : E;
  const int total = np * EO;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    int eo = idx / np;
    int pl = idx - eo * np;
    int e = eo, o = 0;
    if (gs) {
      e = eo / O;
      o = eo - e * O;
    }
    T v = gs ? gs[pl * O + o] : T(1);
    const T* xr = xs + pl * xs_stride + jrel0 * Q;
    for (int t = cnt - 1; t >= 0; --t) {
      int d = e % Q;
      e /= Q;
      v *= xr[t * Q + d];
    }
    tab[eo * se + pl * sp] = v;
  }
}

static inline int ipow_host(int q, int e) {
  long long r = 1;
  for (int i = 0; i < e; ++i) r *= q;
  return (int)r;
}

#define DCTN_CUDA_CHECK_RET(call)                                   \
  do {                                                              \
    cudaError_t _e = (call);                                        \
    if (_e != cudaSuccess) return dctn_set_cuda_error(_e, #call);   \
  } while (0)

int dctn_set_error(int code, const char* fmt, ...);
int dctn_set_cuda_error(cudaError_t e, const char* what);
void dctn_count_launch(int n = 1);
