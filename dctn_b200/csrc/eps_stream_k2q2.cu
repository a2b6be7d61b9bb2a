// Streaming kernels for the HBM-bound corner of the EPS path: K = 2, C = 1, Q_in = 2, float32 (config 1 and the K=2,Q=2 rows
// of the config-3 grid, BASELINE.json; reference: dctn/eps.py:19-40 on an input from dctn/dataset_loading.py:33-36).
//
// Forward.  Algorithmic bytes per patch: 2 floats of x (amortised over the four patches sharing a pixel) + Q_out floats of
// out; 16 * Q_out multiply-adds — the kernel has to issue well under ~60 instructions per patch to stay on the HBM
// roofline, so everything is organised around the instruction count (ncu of the previous kernel: 100 instructions per
// patch, 64 % of the issue slots busy, 37 % of the HBM rate):
//   * one warp = ONE task (RH output rows x 31 output columns of one image), no task loop: the block scheduler balances
//     the tail, the warp is provably convergent (plain SHFL, no WARPSYNC pairs), the task decode happens once;
//   * lane l loads pixel column w0 + l of the RH + 1 input rows (one 64-bit load per row, fully coalesced), the right
//     neighbour comes from lane l + 1 by shuffle; pp[r] = x[r][w] (x) x[r][w+1] is the first Khatri-Rao half of output row
//     r and the second half of row r - 1: computed once, used twice;
//   * out[r][o] = sum_a pp[r][a] * (sum_b pp[r+1][b] * core[a][b][o]) in packed fp32x2 arithmetic (FFMA2) over PAIRS OF
//     OUTPUTS (o, o+1): the core sits in shared memory as [a][pair][b] float2 — its natural (…, o) order — so one 128-bit
//     broadcast load brings two ready operands, and the pair products are kept duplicated (v, v);
//   * the x loads are issued before the core is staged and the CTA barrier, so they overlap it.
// PIX: x is the raw pixel image (B, H, W) and phi(u) = scale * (sin^2(pi u / 2), cos^2(pi u / 2)) is evaluated on load,
// once per loaded pixel (dctn_eps_forward_from_pixels).
#include <cstdlib>

#include "common.cuh"
#include "eps_kernels.h"

namespace {

typedef unsigned long long f32x2_t;   // two fp32 in one 64-bit register pair (low word = first value)
__device__ __forceinline__ f32x2_t pack2(float a, float b) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float2 unpack2(f32x2_t v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// phi of the reference's loader, scale * (sin^2, cos^2)(pi u / 2) = h * (1 - cos(pi u), 1 + cos(pi u)) with h = scale / 2:
// ONE polynomial instead of the two of sincospif.  cos(pi u) has period 1 in v = u / 2 and is even, so v is reduced to
// [-1/2, 1/2] (exact: v - rint(v)), tau = 2 |v| - 1/2 lies in [-1/2, 1/2] and cos(pi u) = -sin(pi tau), an odd polynomial
// of degree 9 (least-squares fit on Chebyshev nodes with p(1/4) = 2 imposed, the constant rounded so that the float32
// Horner evaluation returns exactly 2 there: pixels 0 and 1 — most of an MNIST image — give exactly (0, scale) and
// (scale, 0) like the reference).  Absolute error <= 1.3e-7 * scale over all u (checked against float64 on 7 M points);
// the relative accuracy sincospif has next to the zeros is given up.  14 instructions per pixel instead of 24.
__device__ __forceinline__ void phi_pixel(float u, float h, float& f0, float& f1) {
  float v = 0.5f * u;
  v = v - rintf(v);
  const float tau = fmaf(fabsf(v), 2.f, -0.5f);
  const float s = tau * tau;
  float r = fmaf(s, 0.0771312266588211f, -0.5980005860328674f);
  r = fmaf(r, s, 2.5500240325927734f);
  r = fmaf(r, s, -5.167706489562988f);
  r = fmaf(r, s, 3.1415927410125732f);
  const float sv = fminf(fmaxf(tau * r, -1.f), 1.f);      // sin(pi tau); clamped: the features stay non-negative
  f0 = fmaf(sv, h, h);
  f1 = fmaf(-sv, h, h);
}

constexpr int SK_THREADS = 128;

// grid.x = tasks of one image / warps per CTA, grid.y (x grid.z) = image; a dead warp (task past the image) recomputes task 0.
// "full" (warp-uniform): all RH + 1 input rows and RH output rows of the task exist — every block but possibly the last one
// of an image; only the loads and the stores branch on it, the shuffles stay in straight-line code.
template <int OP, int RH, bool PIX>   // OP: output pairs = ceil(Q_out / 2)
__global__ void __launch_bounds__(SK_THREADS) stream_k2q2_fwd_kernel(const float* __restrict__ x, const float* __restrict__ core,
                                                                      float* __restrict__ out, int H, int W, int O,
                                                                      unsigned nhb, unsigned tpi, unsigned pf_dist, float phi_scale) {
  __shared__ __align__(16) float2 cs[4 * OP * 4];   // [a][pair][b]
  const unsigned lane = threadIdx.x & 31;
  const unsigned tl = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const bool live = tl < tpi;
  const unsigned tk = live ? tl : 0u;
  // consecutive warps = consecutive row blocks of one column tile (the halo row of a block is the first row of the next)
  unsigned hb = tk, tw = 0;
  if (tpi != nhb) { tw = tk / nhb; hb = tk - tw * nhb; }
  const unsigned b = blockIdx.z * gridDim.y + blockIdx.y;
  const unsigned h0 = hb * RH, wcol = tw * 31 + lane;
  const unsigned Ho = H - 1, Wo = W - 1;
  const unsigned wc = wcol < (unsigned)W ? wcol : (unsigned)W - 1u;   // lanes past the row re-read its last pixel
  const bool full = h0 + RH <= Ho;
  // (1) this lane's pixel column, RH + 1 rows; rows past the image re-read the block's first row (value never stored)
  const float* px = x + (size_t)((b * (unsigned)H + h0) * (unsigned)W + wc) * (PIX ? 1 : 2);
  float xv[RH + 1][2];
  if (full) {
#pragma unroll
    for (int r = 0; r <= RH; ++r) {
      if constexpr (PIX) {
        xv[r][0] = __ldg(px + (unsigned)r * (unsigned)W);
      } else {
        const float2 v = __ldg(reinterpret_cast<const float2*>(px) + (unsigned)r * (unsigned)W);
        xv[r][0] = v.x; xv[r][1] = v.y;
      }
    }
  } else {
    const unsigned rows_in = (unsigned)H - h0;
#pragma unroll
    for (int r = 0; r <= RH; ++r) {
      const unsigned ro = (unsigned)r < rows_in ? (unsigned)r * (unsigned)W : 0u;
      if constexpr (PIX) {
        xv[r][0] = __ldg(px + ro);
      } else {
        const float2 v = __ldg(reinterpret_cast<const float2*>(px) + ro);
        xv[r][0] = v.x; xv[r][1] = v.y;
      }
    }
  }
  // L2 prefetch of an image `pf_dist` CTAs ahead: a warp has only (RH + 1) x 256 bytes of loads in flight and the SM's
  // resident warps cannot cover the ~45 KB per SM that HBM latency x bandwidth asks for; the prefetches cost two
  // instructions per CTA, hold no registers, and turn the demand loads above into L2 hits.
  if (pf_dist && blockIdx.x == 0) {
    const unsigned nimg = gridDim.y * gridDim.z, bp = b + pf_dist;
    if (bp < nimg) {
      const unsigned img_bytes = (unsigned)H * (unsigned)W * (PIX ? 4u : 8u);
      const char* base = reinterpret_cast<const char*>(x) + (size_t)bp * img_bytes;
      for (unsigned off = threadIdx.x * 128u; off < img_bytes; off += blockDim.x * 128u)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
    }
  }
  // (2) the core: [a][b][o] in HBM -> [a][pair][b] pairs (o, o+1); odd Q_out: the last pair is (o, 0).  Issued after the x
  // loads so that both are in flight together.
  for (int idx = threadIdx.x; idx < 16 * OP; idx += blockDim.x) {
    const int bq = idx & 3, op = (idx >> 2) % OP, a = idx / (4 * OP);
    const float* c = core + (a * 4 + bq) * O + 2 * op;
    cs[idx] = make_float2(__ldg(c), (2 * op + 1 < O) ? __ldg(c + 1) : 0.f);
  }
  if constexpr (PIX) {
#pragma unroll
    for (int r = 0; r <= RH; ++r) phi_pixel(xv[r][0], 0.5f * phi_scale, xv[r][0], xv[r][1]);
  }
  // (3) duplicated pair products with the right neighbour (lane + 1)
  f32x2_t pp[RH + 1][4];
#pragma unroll
  for (int r = 0; r <= RH; ++r) {
    const float r0 = __shfl_down_sync(0xffffffffu, xv[r][0], 1), r1 = __shfl_down_sync(0xffffffffu, xv[r][1], 1);
    const float p00 = xv[r][0] * r0, p01 = xv[r][0] * r1, p10 = xv[r][1] * r0, p11 = xv[r][1] * r1;
    pp[r][0] = pack2(p00, p00); pp[r][1] = pack2(p01, p01); pp[r][2] = pack2(p10, p10); pp[r][3] = pack2(p11, p11);
  }
  __syncthreads();
  // (4) contraction
  f32x2_t acc[RH][OP];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
    for (int op = 0; op < OP; ++op) {
      const ulonglong2 c01 = *reinterpret_cast<const ulonglong2*>(cs + (a * OP + op) * 4);
      const ulonglong2 c23 = *reinterpret_cast<const ulonglong2*>(cs + (a * OP + op) * 4 + 2);
#pragma unroll
      for (int r = 0; r < RH; ++r) {
        f32x2_t t = mul2(pp[r + 1][0], c01.x);
        t = fma2(pp[r + 1][1], c01.y, t);
        t = fma2(pp[r + 1][2], c23.x, t);
        t = fma2(pp[r + 1][3], c23.y, t);
        acc[r][op] = (a == 0) ? mul2(pp[r][0], t) : fma2(pp[r][a], t, acc[r][op]);
      }
    }
  }
  // (5) store: Q_out consecutive floats per patch, the patches of a row contiguous; all stores of a patch back to back so
  // that its sectors fill at once.  Measured and rejected: staging the tile in shared memory for 128-bit stores of whole
  // contiguous runs (15-20 % slower: the kernel is bound by issue slots and latency, not by store sectors) and storing each
  // output pair as soon as it is complete (9-row tasks for any Q_out: 18 % slower at Q_out = 6, the three partial writes of
  // a patch's 24 bytes arrive far apart).
  if (!live || lane == 31 || wcol >= Wo) return;
  float* ob = out + (size_t)((b * Ho + h0) * Wo + wcol) * (unsigned)O;
  const unsigned orow = Wo * (unsigned)O;
  auto store_row = [&](int r) {
    float* o = ob + (unsigned)r * orow;
    if ((O & 3) == 0) {            // 16 bytes per pair of pairs, patch * O is a multiple of 4: 128-bit stores
#pragma unroll
      for (int op = 0; op + 1 < OP; op += 2) {
        if (2 * op < O) {
          const float2 u = unpack2(acc[r][op]), v = unpack2(acc[r][op + 1]);
          *reinterpret_cast<float4*>(o + 2 * op) = make_float4(u.x, u.y, v.x, v.y);
        }
      }
    } else if ((O & 1) == 0) {     // patch * O + 2 * pair is even: 64-bit stores are aligned
#pragma unroll
      for (int op = 0; op < OP; ++op)
        if (2 * op < O) *reinterpret_cast<float2*>(o + 2 * op) = unpack2(acc[r][op]);
    } else {
#pragma unroll
      for (int op = 0; op < OP; ++op) {
        const float2 v = unpack2(acc[r][op]);
        if (2 * op < O) o[2 * op] = v.x;
        if (2 * op + 1 < O) o[2 * op + 1] = v.y;
      }
    }
  };
  if (full) {
#pragma unroll
    for (int r = 0; r < RH; ++r) store_row(r);
  } else {
    const unsigned rows_out = Ho - h0;
#pragma unroll
    for (int r = 0; r < RH; ++r)
      if ((unsigned)r < rows_out) store_row(r);
  }
}

template <int OP, int RH, bool PIX>
int launch_stream_fwd(const EpsGeom& g, const float* x, const float* core, float* out, float scale, cudaStream_t st) {
  const unsigned nhb = (unsigned)(g.Ho + RH - 1) / RH, ntw = (unsigned)(g.Wo + 30) / 31;
  const unsigned tpi = nhb * ntw;                             // tasks (= warps) per image
  const unsigned wpc = tpi < 4 ? tpi : 4;                     // 28 x 28 with 9-row blocks: one 3-warp CTA per image
  dim3 grid((tpi + wpc - 1) / wpc, (unsigned)g.B, 1);
  if (g.B > 65535) {                                          // images over grid.y x grid.z (B = y * z exactly)
    unsigned z = (unsigned)((g.B + 65534) / 65535);
    while (g.B % z) ++z;
    grid.y = (unsigned)g.B / z; grid.z = z;
    if (grid.y > 65535 || grid.z > 65535) return dctn_set_error(-2, "streaming forward: batch %d does not fit the grid", g.B);
  }
  static const int pf_env = [] { const char* e = getenv("DCTN_B200_STREAM_PF"); return e ? atoi(e) : 1024; }();
  const unsigned pf_dist = (unsigned)pf_env;
  stream_k2q2_fwd_kernel<OP, RH, PIX><<<grid, 32 * wpc, 0, st>>>(x, core, out, g.H, g.W, g.O, nhb, tpi, pf_dist, scale);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}

// rows per task: 9 divides the 27 output rows of a 28 x 28 image (1.11 loaded rows per output row, the core read once per
// 9 rows) and still fits the registers for one or two output pairs; three and four pairs take 4 rows
template <bool PIX>
int dispatch_stream_fwd(const EpsGeom& g, const float* x, const float* core, float* out, float scale, cudaStream_t st) {
  static const int rh_env = [] { const char* e = getenv("DCTN_B200_STREAM_RH"); return e ? atoi(e) : 0; }();
  const int op = (g.O + 1) / 2;
  const bool tall = rh_env ? rh_env > 4 : (g.Ho % 9 == 0 || g.Ho >= 45);
  switch (op) {
    case 1: return tall ? launch_stream_fwd<1, 9, PIX>(g, x, core, out, scale, st) : launch_stream_fwd<1, 4, PIX>(g, x, core, out, scale, st);
    case 2: return tall ? launch_stream_fwd<2, 9, PIX>(g, x, core, out, scale, st) : launch_stream_fwd<2, 4, PIX>(g, x, core, out, scale, st);
    case 3: return launch_stream_fwd<3, 4, PIX>(g, x, core, out, scale, st);
    case 4: return launch_stream_fwd<4, 4, PIX>(g, x, core, out, scale, st);
    default: return dctn_set_error(-2, "streaming forward: Q_out = %d not instantiated", g.O);
  }
}

}  // namespace

bool stream_k2q2_enabled() {
  static const bool on = [] { const char* e = getenv("DCTN_B200_STREAM"); return !e || atoi(e) != 0; }();
  return on;
}
bool stream_k2q2_supported(const EpsGeom& g, int dtype) {
  return dtype == 0 && g.K == 2 && g.C == 1 && g.Q == 2 && g.O >= 1 && g.O <= 8 && g.H >= 2 && g.W >= 2 &&
         g.P * g.O < (1ll << 31) && (long long)g.B * g.H * g.W * g.Q < (1ll << 31);
}
int stream_k2q2_forward(const EpsGeom& g, const float* x, const float* core, float* out, cudaStream_t st) {
  return dispatch_stream_fwd<false>(g, x, core, out, 0.f, st);
}
int stream_k2q2_forward_pixels(const EpsGeom& g, const float* pixels, float scale, const float* core, float* out, cudaStream_t st) {
  return dispatch_stream_fwd<true>(g, pixels, core, out, scale, st);
}
