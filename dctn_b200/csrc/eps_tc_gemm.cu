// tcgen05 GEMMs of the EPS forward and input-gradient passes (float32 in/out; arithmetic: split fp16 "3xFP16" (default),
// split TF32 "3xTF32", or single-pass TF32).
//
//   C[p][c] = sum_k Gen[p][k] * Bop[k][c]            p: 128 patches per CTA (TMEM lanes), c: BN columns per tile
//
//   * Gen (Khatri-Rao half, optionally times gout) is GENERATED per stage by 4 producer warps, one patch row per
//     thread (= one TMEM lane), from two-level tables built once per CTA, split into TF32 hi / lo parts and written
//     with tcgen05.st straight into TENSOR MEMORY: the MMAs read A from TMEM (TS form), so the generated operand
//     never touches shared memory — shared-memory bandwidth was the measured bottleneck of the SS form
//     (profiles/r01_gemm_fwd_ss_ncu.txt: smem pipe 91% busy, tensor pipe 51%);
//   * Bop (the core) is pre-packed once per call by pack_core_kernel into per-(tile, k-chunk) images that are
//     already split (hi / lo), K-major and swizzled, so that one elected thread streams each stage with two
//     cp.async.bulk copies (TMA engine, mbarrier complete_tx) — no tensor map needed;
//   * one elected thread issues tcgen05.mma kind::tf32 (M=128, N=BN, K=8): the dominant hi*hi products accumulate
//     in a MAIN TMEM accumulator, the two cross terms (hi*lo, lo*hi) in a separate SMALL accumulator.  The tensor
//     core rounds its fp32 accumulator toward zero on every MMA; keeping the small terms out of the main chain
//     cuts that bias by 3x (one rounding per k-step instead of three) at zero cost;
//   * 4 epilogue warps read both accumulators with tcgen05.ld and apply one of three fused epilogues:
//       MODE_FWD   out[p][o]  = sum_b C[p][(o,b)] * KR2[p][b]          (core packed as [a][(o,b)])      dctn/eps.py:19-40
//       MODE_DKR2  dKR2[p][b] = sum_o C[p][(b,o)] * gout[p][o]         (C = KR1 @ core, never stored)
//       MODE_STORE dKR1[p][a] = C[p][a]                                (Gen = KR2 x gout, Bop = core^T)
//
// 3xFP16 (F16 = true): every fp32 operand value v is represented as hi + lo * 2^-11 with hi = fp16(v) and
// lo = fp16((v - hi) * 2^11): 22 significant bits, exactly what the TF32 split gives, but kind::f16 MMAs run at twice the
// kind::tf32 rate and move half the operand bytes (a 128-byte swizzled row / a 32-column TMEM slab holds 64 K-values
// instead of 32).  hi*hi goes to the MAIN accumulator, hi*lo + lo*hi (scaled by 2^11) to the SMALL one; the epilogue
// adds small * 2^-11.  fp16 has a 5-bit exponent, so both operands are range-normalised with exact power-of-two
// scales that the epilogue undoes: each factor vector x_j (and the gout row) of a patch is scaled to max-abs in
// [0.5, 1) and the generated Khatri-Rao row by 2^15 (so its largest entry lies in [2^(15-#factors), 2^15)); the core
// by one global power of two that puts max|core| in [2^14, 2^15).  Entries far below the row / core maximum
// lose relative (not absolute) precision: |error| <= max(2^-22 |v|, 2^-36) at max = 2^15.
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>

#include "common.cuh"
#include "eps_kernels.h"
#include "tc_common.cuh"

// cycle probes for tuning (make NVFLAGS+=-DDCTN_TCG_TIMING, then run with DCTN_TCG_DEBUG=1); compiled out by default
#ifdef DCTN_TCG_TIMING
#define TCG_CLK() clock64()
#else
#define TCG_CLK() 0ll
#endif

namespace {

constexpr int GBM = 128;
constexpr int GBK = 32;         // K values per pipeline stage, TF32 (one 128-byte swizzled row = 32 fp32)
constexpr int GBK16 = 64;       // K values per pipeline stage, FP16 (one 128-byte swizzled row = 64 fp16)
constexpr int ARITH_F16X3 = 6;  // value of `passes` that selects the split-fp16 arithmetic (1, 3: TF32 passes)
constexpr int ASTAGES = 2;      // A-operand stages in TMEM (2 x (hi + lo) x 32 columns = 128 columns)
constexpr int MAX_BSTAGES = 4;  // B-operand stages in shared memory (as many as fit)
constexpr int MAX_BN = 192;     // accumulators: main + small = 2*BN columns, + 128 for A  <= 512 TMEM columns
constexpr int G_THREADS = 640;  // warp 0: bulk copies, warp 1: MMA, warp 2: TMEM alloc, warps 4-7 and 12-15: producers (each half of a
                                // stage's K range), warps 8-11 and 16-19: epilogue (two warps per TMEM lane quadrant, alternating
                                // 32-column batches: the epilogue of a tile is not overlapped with the next tile's MMAs — one
                                // accumulator set — so its length is paid in full, and one warp per scheduler runs its ~1000
                                // dependent instructions per batch at half the rate two warps reach)
constexpr int G_EPI_WARPS = 8;
enum { MODE_STORE = 0, MODE_FWD = 1, MODE_DKR2 = 2 };
constexpr size_t TCG_SMEM_LIMIT = 227 * 1024;

struct TcGemmArgs {
  EpsGeom g;
  const float* x;
  const float* gout;
  long long p0;  // first patch handled by this launch
  int np;        // number of patches
  // generated operand (struct GemmShape below): nf factors from jh0 (times gout if withG); value(h, r) =
  // tabKH[kh(h)] * tabKL[kl(h)] * tabR[r] with K ordered [r / RB][h][r % RB] (padded: h to Hpad, r to GP)
  int jh0, nf, withG, cnth, KH, cntl, KLb, cr, G, GP, RB, H, Hpad, Kp;
  int Ncols, ntiles, nk;
  int kseg, nseg;       // K stages per accumulation segment, segments per column tile (tc::seg_stages)
  int dbuf;             // 1: two accumulator sets (BN <= 96) alternate between virtual tiles: the epilogue of one overlaps the MMAs of the next
  long long seg_stride; // MODE_STORE: segment sg writes its own slice out + sg * seg_stride (summed by sum_slices_kernel)
  const float* packed;  // [ntiles][nk][2][BN*32]
  int BN;               // column-tile width: multiple of 16, <= MAX_BN
  int bstages;          // shared-memory stages for B
  int passes;           // MMA passes per product: 3 (split) or 1
  const uint32_t* core_absmax;  // F16: bits of max|core| (written by absmax_kernel), fixes the core's power-of-two scale
  float* out;           // MODE_STORE: [np][ldc]; MODE_FWD: out[P][O] (absolute patches); MODE_DKR2: [np][Bn]
  long long ldc;
  float* tsave;         // MODE_FWD, training: the accumulator rows T[p][(o, b)] are also stored here, [P][Ncols]
  long long* dbg;       // optional per-CTA cycle counters (DCTN_TCG_DEBUG): 8 per CTA
};

// ------------------------------------------------------------------------------------------------ core packing
// dst[((tile*nk + kc)*2 + part)*BN*32 + swizzled(row rr, k)] = part(core element (c = tile*BN + rr, k = kc*32 + ..))
//   MODE_STORE: element(c, k) = core[c*N + k]           (c = a, k = n)
//   MODE_DKR2 : element(c, k) = core[k*N + c]           (c = n, k = a)
//   MODE_FWD  : element(c, k) = core[(k*Bn + b)*O + o]  (c = o*Bn + b, k = a)
using tc::core_scale_exp;
using tc::split_f16x2;

__global__ void absmax_kernel(const float* __restrict__ v, long long n, uint32_t* __restrict__ out) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(v[i]));   // fmaxf drops NaNs; an inf core stays inf and disables the scaling
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));   // non-negative floats order like their bit patterns
}

// K order of the generated operand (GemmShape): packed position k' -> true k, or -1 for a padding row
struct KOrder {
  int H, Hpad, G, RB, Kp;
};
__device__ __forceinline__ int korder_true_k(const KOrder& ko, int kp) {
  if (kp >= ko.Kp) return -1;
  const int sec = ko.Hpad * ko.RB;
  const int ob = kp / sec, rem = kp - ob * sec;
  const int h = rem / ko.RB, r = ob * ko.RB + (rem - h * ko.RB);
  return (h < ko.H && r < ko.G) ? h * ko.G + r : -1;
}

template <bool F16>
__global__ void pack_core_kernel(const float* __restrict__ core, float* __restrict__ dst, EpsGeom g, int mode, int BN,
                                 int Ncols, KOrder ko, int ntiles, int nk, int passes, const uint32_t* __restrict__ absmax) {
  constexpr int KV = F16 ? 8 : 4;      // K values per 16-byte chunk
  constexpr int KS = F16 ? GBK16 : GBK;
  const long long total = (long long)ntiles * nk * BN * 8;  // one thread per 16-byte chunk
  float scale = 1.f;
  if (F16) scale = scalbnf(1.f, core_scale_exp(*absmax));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c16 = (int)(i & 7);
    long long r = i >> 3;
    const int rr = (int)(r % BN);
    r /= BN;
    const int kc = (int)(r % nk);
    const int tile = (int)(r / nk);
    const int c = tile * BN + rr;
    float v[KV];
#pragma unroll
    for (int u = 0; u < KV; ++u) v[u] = 0.f;
    if (c < Ncols) {
#pragma unroll
      for (int u = 0; u < KV; ++u) {
        const int k = korder_true_k(ko, kc * KS + c16 * KV + u);
        if (k >= 0) {
          long long idx;
          if (mode == MODE_STORE) idx = (long long)c * g.N + k;
          else if (mode == MODE_DKR2) idx = (long long)k * g.N + c;
          else {
            const int o = c / g.Bn, b = c - o * g.Bn;
            idx = ((long long)k * g.Bn + b) * g.O + o;
          }
          v[u] = __ldg(&core[idx]) * scale;
        }
      }
    }
    float* tile_base = dst + ((long long)(tile * nk + kc) * 2) * BN * 32;
    const int off = rr * 32 + ((c16 ^ (rr & 7)) << 2);  // in 4-byte units
    if (F16) {
      uint4 hi, lo;
      split_f16x2(v[0], v[1], hi.x, lo.x);
      split_f16x2(v[2], v[3], hi.y, lo.y);
      split_f16x2(v[4 % KV], v[5 % KV], hi.z, lo.z);
      split_f16x2(v[6 % KV], v[7 % KV], hi.w, lo.w);
      *(uint4*)(tile_base + off) = hi;
      if (passes == 3) *(uint4*)(tile_base + BN * 32 + off) = lo;
    } else {
      float4 hi, lo;
      tc::split_tf32(v[0], hi.x, lo.x);
      tc::split_tf32(v[1], hi.y, lo.y);
      tc::split_tf32(v[2], hi.z, lo.z);
      tc::split_tf32(v[3], hi.w, lo.w);
      *(float4*)(tile_base + off) = hi;
      if (passes == 3) *(float4*)(tile_base + BN * 32 + off) = lo;
    }
  }
}

// ------------------------------------------------------------------------------------------------ the GEMM
// Epilogue store of a 32-row x 32-column block whose rows live one per lane (v = this lane's row): staged, 16 columns at
// a time, through a [32][20] shared-memory tile so that a store instruction writes whole row segments (8 rows x 64
// contiguous bytes, or 2 rows x 64 bytes when the destination is not 16-byte aligned) instead of 32 rows x 16 (or 4)
// bytes — the row-strided form costs one L1 wavefront per row and instruction (measured: 4.2k cycles per batch for the
// odd-pitched dKR1 of the CIFAR (2, 23 -> 24) layer, 0.6k this way).
// dst0: row 0, first column of the block; values are multiplied by s1 * s2; accum: add to what is there.
constexpr int TST_FLOATS = 32 * 20;
__device__ __forceinline__ void store_tile(float* st, const float (&v)[32], float s1, float s2, float* dst0, long long ld,
                                           int nrows, int ncols, bool accum, int lane) {
  const bool vec = ncols == 32 && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(dst0) & 15) == 0;
#pragma unroll
  for (int hb = 0; hb < 2; ++hb) {
#pragma unroll
    for (int i = 0; i < 16; i += 4)
      *(float4*)(st + lane * 20 + i) = make_float4(v[16 * hb + i] * s1 * s2, v[16 * hb + i + 1] * s1 * s2,
                                                   v[16 * hb + i + 2] * s1 * s2, v[16 * hb + i + 3] * s1 * s2);
    __syncwarp();
    if (vec) {
      const int rsub = lane >> 2, c4 = (lane & 3) * 4;
#pragma unroll
      for (int r0 = 0; r0 < 32; r0 += 8) {
        const int r = r0 + rsub;
        if (r < nrows) {
          float4 x4 = *(const float4*)(st + r * 20 + c4);
          float4* p = (float4*)(dst0 + (long long)r * ld + 16 * hb + c4);
          if (accum) {
            const float4 o4 = *p;
            x4.x += o4.x; x4.y += o4.y; x4.z += o4.z; x4.w += o4.w;
          }
          *p = x4;
        }
      }
    } else {
      const int rsub = lane >> 4, c = lane & 15;
      float* p = dst0 + (long long)rsub * ld + 16 * hb + c;
#pragma unroll 4
      for (int r0 = 0; r0 < 32; r0 += 2, p += 2 * ld) {
        if (r0 + rsub < nrows && 16 * hb + c < ncols) {
          const float x1 = st[(r0 + rsub) * 20 + c];
          *p = accum ? *p + x1 : x1;
        }
      }
    }
    __syncwarp();
  }
}

// NV consecutive values of the generated operand of one patch row, from packed K position kp0 (a multiple of NV; NV
// divides 32, so the run stays inside one register-group section): the RB register-group values of the section are
// loaded once, then every table value (two loads) yields RB products
template <int NV, int RB>
__device__ __forceinline__ void gen_values(float (&v)[NV], int kp0, int Hpad, const uint32_t* __restrict__ hidx,
                                           const float* __restrict__ th, const float* __restrict__ tl, const float* __restrict__ tr) {
  const int sec = Hpad * RB;
  const int ob = kp0 / sec;
  const int h0 = (kp0 - ob * sec) / RB;
  float rg[RB];
#pragma unroll
  for (int r = 0; r < RB; ++r) rg[r] = tr[(ob * RB + r) * 128];
#pragma unroll
  for (int i = 0; i < NV / RB; ++i) {
    const uint32_t id = hidx[h0 + i];
    const float hv = th[(id & 0xFFFF) * 128] * tl[(id >> 16) * 128];
#pragma unroll
    for (int r = 0; r < RB; ++r) v[i * RB + r] = hv * rg[r];
  }
}

template <int MODE, bool F16>
__global__ void __launch_bounds__(G_THREADS, 1) tc_gemm_kernel(const __grid_constant__ TcGemmArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  const EpsGeom& g = a.g;
  constexpr int KS = F16 ? GBK16 : GBK;                // K values per pipeline stage
  const int Q = g.Q, O = g.O, BN = a.BN, NB = a.bstages;
  const uint32_t B_BYTES = (uint32_t)BN * 128;         // one part (hi or lo) of a B stage: BN rows of 128 bytes
  const uint32_t STAGE_BYTES = 2 * B_BYTES;            // multiple of 1024 (BN % 16 == 0)
  unsigned char* base = smem_dyn + ((1024u - (tc::smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* stages = base;
  float* tabKH = (float*)(base + NB * STAGE_BYTES);    // [KH + 1][128] (row KH is all zeros: padding h)
  float* tabKL = tabKH + (a.KH + 1) * 128;             // [KLb][128]
  float* tabR = tabKL + a.KLb * 128;                   // [GP][128] register group (rows >= G are zero)
  float* tabE = tabR + a.GP * 128;                     // MODE_FWD: [BH + BL][128]; MODE_DKR2: gout [O][128]
  const int nE = (MODE == MODE_FWD) ? (g.BH + g.BL) : (MODE == MODE_DKR2 ? O : 0);
  float* outs = tabE + nE * 128;                       // MODE_FWD: [2 epilogue groups][O][128]
  // index tables that make the inner loops branch-free (every load address is known up front -> full ILP):
  //   hidx[h]  = kh | kl << 16 for the table part of the generated operand (h >= H -> the all-zero row KH of tabKH)
  //   eidx[..] = MODE_FWD: bh | bl << 16 for b2 in [0, Bn + max(Bn, 32)), b = b2 % Bn (32 consecutive b never leave the table);
  //              MODE_DKR2: o | last << 8 for the BN columns of a tile
  uint32_t* hidx = (uint32_t*)(outs + ((MODE == MODE_FWD) ? 2 * O * 128 : 0));
  uint32_t* eidx = hidx + a.Hpad;
  const int neidx = (MODE == MODE_FWD) ? g.Bn + (g.Bn > 32 ? g.Bn : 32) : (MODE == MODE_DKR2 ? ((BN + 31) & ~31) : 0);
  // F16: power-of-two exponents of the per-patch normalisation: [0][pr] generated operand, [1][pr] epilogue factors
  int* rowexp = (int*)(eidx + ((neidx + 3) & ~3));       // 16-byte aligned (Hpad % 4 == 0): tstage is accessed as float4
  float* tstage = (float*)(rowexp + 256);              // MODE_STORE / MODE_FWD: [8 epilogue warps][32][20] (store_tile)
  uint64_t* bars = (uint64_t*)(tstage + ((MODE == MODE_DKR2) ? 0 : G_EPI_WARPS * TST_FLOATS));
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * MAX_BSTAGES + 2 * ASTAGES + 4);
  const uint32_t bar_fullB0 = tc::smem_u32(bars), bar_emptyB0 = bar_fullB0 + 8 * MAX_BSTAGES;
  const uint32_t bar_fullA0 = bar_emptyB0 + 8 * MAX_BSTAGES, bar_emptyA0 = bar_fullA0 + 8 * ASTAGES;
  const uint32_t bar_accfull = bar_emptyA0 + 8 * ASTAGES, bar_accempty = bar_accfull + 16;   // one pair per accumulator set
  // setup-only scratch aliased onto the (not yet used) B stages: x [n*Q][128] and gout [O][128]
  float* xs = (float*)stages;
  float* gsx = xs + g.n * Q * 128;
  int* fexp = (int*)(gsx + O * 128);                   // F16 only: [n + 1][128] exponents of the factors and of gout

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pl0 = blockIdx.x * GBM;                 // first patch of this CTA, relative to the launch
  const long long pt0 = a.p0 + pl0;                 // absolute
  constexpr uint32_t TMEM_COLS = 512;

  // ---------------- setup: barriers, TMEM, tables
  if (tid == 0) {
    for (int s = 0; s < MAX_BSTAGES; ++s) {
      tc::mbar_init(bar_fullB0 + 8 * s, 1);      // the expect_tx arrive of the copy warp (+ transaction bytes)
      tc::mbar_init(bar_emptyB0 + 8 * s, 1);     // tcgen05.commit
    }
    for (int s = 0; s < ASTAGES; ++s) {
      tc::mbar_init(bar_fullA0 + 8 * s, 8);      // 8 producer warps
      tc::mbar_init(bar_emptyA0 + 8 * s, 1);     // tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(bar_accfull + 8 * s, 1);
      tc::mbar_init(bar_accempty + 8 * s, G_EPI_WARPS);
    }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(tc::smem_u32(tmem_slot), TMEM_COLS);
  {
    // a thread owns one patch row (pr) and every fifth factor (G_THREADS = 5 x 128): ONE patch-origin computation, the Q
    // loads of a factor issued together and — F16 — its range normalisation in the same pass: the factor vector (and the
    // gout row) is scaled by a power of two so that its largest magnitude lies in [0.5, 1); exact, undone by the
    // epilogue through rowexp
    const int pr = tid & 127, slot = tid >> 7;
    const long long p = pt0 + pr;
    const bool valid = p < g.P;
    const long long org = valid ? patch_origin(g, p) : 0;
    const bool needg = a.withG || MODE == MODE_DKR2;
    for (int j = slot; j <= g.n; j += G_THREADS / 128) {
      const bool isg = j == g.n;
      if (isg && !needg) { if (F16) fexp[j * 128 + pr] = 0; continue; }
      const int cnt = isg ? O : Q;
      const float* src = isg ? a.gout + p * O : a.x + org + g.foff[isg ? 0 : j];
      float* dst = isg ? gsx + pr : xs + j * Q * 128 + pr;
      float m = 0.f;
      for (int q = 0; q < cnt; ++q) {
        const float v = valid ? __ldg(src + q) : 0.f;
        dst[q * 128] = v;
        m = fmaxf(m, fabsf(v));
      }
      if (F16) {
        int e = tc::norm_exp(m);
        if (isg && !a.withG) e = 0;     // MODE_DKR2 uses gout only in the epilogue (fp32): keep it as is
        if (e != 0) {
          const float s1 = __int_as_float((127 - e / 2) << 23), s2 = __int_as_float((127 - (e - e / 2)) << 23);
          for (int q = 0; q < cnt; ++q) dst[q * 128] = dst[q * 128] * s1 * s2;
        }
        fexp[j * 128 + pr] = e;
      }
    }
  }
  __syncthreads();
  if (F16) {
    if (tid < 128) {
      int ea = 0, eb = 0;
      for (int j = 0; j < g.n; ++j) {
        const bool in_gen = (j >= a.jh0 && j < a.jh0 + a.nf);
        if (in_gen) ea += fexp[j * 128 + tid];
        else eb += fexp[j * 128 + tid];
      }
      if (a.withG) ea += fexp[g.n * 128 + tid];
      rowexp[tid] = ea;
      rowexp[128 + tid] = eb;
    }
  }
  {
    // table entry e of a group of `cnt` factors starting at factor j0: prod_u x[j0+u][digit_u(e)] (digit 0 slowest)
    auto kr_entry = [&](int j0, int cnt, int e, int pr) -> float {
      float v = 1.f;
      for (int u = cnt - 1; u >= 0; --u) {
        const int d = e % Q;
        e /= Q;
        v *= xs[((j0 + u) * Q + d) * 128 + pr];
      }
      return v;
    };
    // F16: the generated row carries a factor 2^15 (largest entry in [2^(15 - #factors), 2^15))
    const float gen_scale = F16 ? 32768.f : 1.f;
    for (int idx = tid; idx < a.KH * 128; idx += G_THREADS) tabKH[idx] = gen_scale * kr_entry(a.jh0, a.cnth, idx >> 7, idx & 127);
    for (int idx = tid; idx < a.KLb * 128; idx += G_THREADS) tabKL[idx] = kr_entry(a.jh0 + a.cnth, a.cntl, idx >> 7, idx & 127);
    for (int idx = tid; idx < a.GP * 128; idx += G_THREADS) {
      const int pr = idx & 127, r = idx >> 7;
      float v = 0.f;
      if (r < a.G) {
        int e = r;
        float gv = 1.f;
        if (a.withG) {
          e = r / O;
          gv = gsx[(r - e * O) * 128 + pr];
        }
        v = gv * kr_entry(a.jh0 + a.cnth + a.cntl, a.cr, e, pr);
      }
      tabR[idx] = v;
    }
    if (MODE == MODE_FWD) {
      for (int idx = tid; idx < g.BH * 128; idx += G_THREADS) tabE[idx] = kr_entry(g.m, g.b_nh, idx >> 7, idx & 127);
      for (int idx = tid; idx < g.BL * 128; idx += G_THREADS)
        tabE[g.BH * 128 + idx] = kr_entry(g.m + g.b_nh, g.b_nl, idx >> 7, idx & 127);
      for (int idx = tid; idx < 2 * O * 128; idx += G_THREADS) outs[idx] = 0.f;
    }
    if (MODE == MODE_DKR2)
      for (int idx = tid; idx < O * 128; idx += G_THREADS) tabE[idx] = gsx[idx];
    if (tid < 128) tabKH[a.KH * 128 + tid] = 0.f;
    for (int h = tid; h < a.Hpad; h += G_THREADS)
      hidx[h] = (h < a.H) ? ((uint32_t)(h / a.KLb) | ((uint32_t)(h % a.KLb) << 16)) : (uint32_t)a.KH;
    if (MODE == MODE_FWD)
      for (int b2 = tid; b2 < neidx; b2 += G_THREADS) {
        const int b = b2 % g.Bn;
        eidx[b2] = (uint32_t)(b / g.BL) | ((uint32_t)(b % g.BL) << 16);
      }
    if (MODE == MODE_DKR2)
      for (int c = tid; c < neidx; c += G_THREADS) eidx[c] = (uint32_t)(c % O) | ((c % O == O - 1) ? 0x100u : 0u);
  }
  tc::tc_fence_before();
  __syncthreads();  // tables complete, xs/gsx scratch (aliasing the stages) dead from here on
  tc::tc_fence_after();
  // F16: exponent that turns an accumulator value into the true product: patch normalisation, 2^15 of the generated
  // row and the core's global scale
  const int core_exp = F16 ? core_scale_exp(__ldg(a.core_absmax)) : 0;
  // accumulator set s (only set 0 without dbuf): main at +2*BN*s, small at +2*BN*s + BN; then the A stages
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_a0 = tmem_base + (a.dbuf ? 4u : 2u) * (uint32_t)BN;   // stage s: hi at +64*s, lo at +64*s + 32
  const int total_it = a.ntiles * a.nk;

  if (warp == 0) {
    // =========================== bulk-copy issuer (B operand) ===========================
    if (tc::elect_one()) {
      const uint32_t bytes = B_BYTES * (a.passes == 3 ? 2u : 1u);
      int s = 0;
      uint32_t ph = 1;   // parity to wait for on the empty barrier: fresh barriers pass a wait on parity 1
      const float* src = a.packed;
      for (int i = 0; i < total_it; ++i) {
        tc::mbar_wait(bar_emptyB0 + 8 * s, ph);
        const uint32_t sb = tc::smem_u32(stages + s * STAGE_BYTES);
        tc::mbar_arrive_expect_tx(bar_fullB0 + 8 * s, bytes);
        tc::bulk_g2s(sb, src, B_BYTES, bar_fullB0 + 8 * s);
        if (a.passes == 3) tc::bulk_g2s(sb + B_BYTES, src + BN * 32, B_BYTES, bar_fullB0 + 8 * s);
        src += 2 * BN * 32;
        if (++s == NB) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = F16 ? tc::make_idesc_f16(GBM, BN) : tc::make_idesc_tf32(GBM, BN);
    long long dbg_waitA = 0, dbg_waitB = 0, dbg_waitAcc = 0, dbg_start = TCG_CLK();
    int sa = 0, sb_ = 0;
    uint32_t pha = 0, phb = 0;     // parities of the full barriers
    const uint64_t db_base = tc::make_sw128_kmajor_desc(tc::smem_u32(stages));
    const uint32_t stage_adv = STAGE_BYTES >> 4, part_adv = B_BYTES >> 4;  // descriptor address units (16 bytes)
    // one "virtual tile" per (column tile, K segment): see tc::seg_stages
    const int nvt = a.ntiles * a.nseg;
    for (int u = 0; u < nvt; ++u) {
      const int sg = u % a.nseg;
      const int kc0 = sg * a.kseg, kc1 = (kc0 + a.kseg < a.nk) ? kc0 + a.kseg : a.nk;
      long long ta = TCG_CLK();
      const int set = a.dbuf ? (u & 1) : 0, nuse = a.dbuf ? (u >> 1) : u;   // nuse: earlier uses of this accumulator set
      const uint32_t tmem_main = tmem_base + (uint32_t)(2 * BN * set), tmem_small = tmem_main + (uint32_t)BN;
      if (nuse > 0) tc::mbar_wait(bar_accempty + 8 * set, (uint32_t)((nuse - 1) & 1));  // epilogue has drained this set
      dbg_waitAcc += TCG_CLK() - ta;
      tc::tc_fence_after();
      for (int kc = kc0; kc < kc1; ++kc) {
        long long t0 = TCG_CLK();
        tc::mbar_wait(bar_fullB0 + 8 * sb_, phb);
        long long t1 = TCG_CLK();
        tc::mbar_wait(bar_fullA0 + 8 * sa, pha);
        long long t2 = TCG_CLK();
        dbg_waitB += t1 - t0; dbg_waitA += t2 - t1;
        tc::tc_fence_after();
        if (tc::elect_one()) {
          const uint64_t db_hi = db_base + (uint64_t)(sb_ * stage_adv);
          const uint64_t db_lo = db_hi + part_adv;
          const uint32_t a_hi = tmem_a0 + (uint32_t)(sa * 64), a_lo = a_hi + 32;
#pragma unroll
          for (int k = 0; k < 4; ++k) {                 // 4 MMAs of 32 bytes of K per row: 8 x tf32 or 16 x fp16
            const uint64_t adv = (uint64_t)(k * 2);     // 32 bytes >> 4 along the K-major smem rows
            const uint32_t acol = (uint32_t)(k * 8);    // 8 TMEM columns
            const uint32_t first = (kc == kc0 && k == 0) ? 0u : 1u;
            if (F16) {
              tc::umma_f16_ts(tmem_main, a_hi + acol, db_hi + adv, idesc, first);
              if (a.passes == 3) {
                tc::umma_f16_ts(tmem_small, a_hi + acol, db_lo + adv, idesc, first);
                tc::umma_f16_ts(tmem_small, a_lo + acol, db_hi + adv, idesc, 1u);
              }
            } else {
              tc::umma_tf32_ts(tmem_main, a_hi + acol, db_hi + adv, idesc, first);
              if (a.passes == 3) {
                tc::umma_tf32_ts(tmem_small, a_hi + acol, db_lo + adv, idesc, first);
                tc::umma_tf32_ts(tmem_small, a_lo + acol, db_hi + adv, idesc, 1u);
              }
            }
          }
          tc::umma_commit(bar_emptyA0 + 8 * sa);
          tc::umma_commit(bar_emptyB0 + 8 * sb_);
          if (kc == kc1 - 1) tc::umma_commit(bar_accfull + 8 * set);
        }
        __syncwarp();
        if (++sa == ASTAGES) { sa = 0; pha ^= 1; }
        if (++sb_ == NB) { sb_ = 0; phb ^= 1; }
      }
    }
    if (a.dbg && lane == 0) {
      long long* d = a.dbg + (long long)blockIdx.x * 8;
      d[0] = dbg_waitA; d[1] = dbg_waitB; d[2] = dbg_waitAcc; d[3] = TCG_CLK() - dbg_start;
    }
  } else if ((warp >= 4 && warp < 8) || (warp >= 12 && warp < 16)) {
    // =========================== A producers: one patch row (= TMEM lane) and one half of the stage per thread ===========================
    const int ph_ = warp >= 12 ? 1 : 0;        // which half of the stage's K range
    const int pr = (warp & 3) * 32 + lane;
    const float* th = tabKH + pr;
    const float* tl = tabKL + pr;
    const float* tr = tabR + pr;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    int sa = 0;
    uint32_t phe = 1;   // parity to wait for on the empty barrier
    long long dbg_pwait = 0, dbg_pst = 0, dbg_pgen = 0, tprev = TCG_CLK();
    for (int t = 0; t < a.ntiles; ++t) {
      for (int kc = 0; kc < a.nk; ++kc) {
        uint32_t hi[16], lo[16];   // this half of a 32-column TMEM slab: 16 tf32 values or 32 packed fp16 values
        constexpr int NV = KS / 2;  // generated values per thread and stage
        float v[NV];
        const int kp0 = kc * KS + NV * ph_;
        if (kp0 < a.Kp) {
          if (a.RB == 8) gen_values<NV, 8>(v, kp0, a.Hpad, hidx, th, tl, tr);
          else gen_values<NV, 4>(v, kp0, a.Hpad, hidx, th, tl, tr);
        } else {
#pragma unroll
          for (int j = 0; j < NV; ++j) v[j] = 0.f;
        }
        if (F16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) split_f16x2(v[(2 * j) % NV], v[(2 * j + 1) % NV], hi[j], lo[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float fh, fl;
            tc::split_tf32(v[j % NV], fh, fl);
            hi[j] = __float_as_uint(fh); lo[j] = __float_as_uint(fl);
          }
        }
        long long t0 = TCG_CLK();
        tc::mbar_wait(bar_emptyA0 + 8 * sa, phe);
        long long t1 = TCG_CLK();
        tc::tc_fence_after();
        const uint32_t dst = tmem_a0 + lane_base + (uint32_t)(sa * 64 + 16 * ph_);
        tc::tmem_st16_u(dst, hi);
        if (a.passes == 3) tc::tmem_st16_u(dst + 32, lo);
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(bar_fullA0 + 8 * sa);
        long long t2 = TCG_CLK();
        dbg_pwait += t1 - t0; dbg_pst += t2 - t1; dbg_pgen += t0 - tprev; tprev = t2;
        if (++sa == ASTAGES) { sa = 0; phe ^= 1; }
      }
    }
    if (a.dbg && warp == 4 && lane == 0) {
      long long* d = a.dbg + (long long)blockIdx.x * 8;
      d[4] = dbg_pwait; d[5] = dbg_pst; d[6] = dbg_pgen;
    }
  } else if ((warp >= 8 && warp < 12) || warp >= 16) {
    // =========================== epilogue ===========================
    // group 0 (warps 8-11) takes the even 32-column batches of a tile, group 1 (warps 16-19) the odd ones; MODE_DKR2's
    // reduction runs across batches: group 0 does all of it
    const int quad = warp & 3, grp = warp >= 16 ? 1 : 0;
    float* outs_g = outs + grp * O * 128;
    float* tst = tstage + (grp * 4 + quad) * TST_FLOATS;
    const int pr = quad * 32 + lane;
    const int pl = pl0 + pr;                 // relative to the launch
    const bool pvalid = pl < a.np;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    // running state of the fused reductions (all of it warp-uniform except s)
    float s = 0.f;
    int fo = 0, fb = 0;                      // MODE_FWD: current o and b of the next column
    int db = 0;                              // MODE_DKR2: next b to store
    const float* eH = tabE + pr;
    const float* eL = tabE + g.BH * 128 + pr;
    // F16: accumulator -> true value is a multiplication by 2^kexp, applied as two exact factors (|kexp| can exceed 127)
    float sc1 = 1.f, sc2 = 1.f, fsc1 = 1.f, fsc2 = 1.f;   // sc: generated operand only (T, dKR1, dKR2); fsc: + epilogue factors
    if (F16) {
      const int kexp = rowexp[pr] - 15 - core_exp;
      sc1 = scalbnf(1.f, kexp / 2); sc2 = scalbnf(1.f, kexp - kexp / 2);
      const int fexp_all = kexp + rowexp[128 + pr];
      fsc1 = scalbnf(1.f, fexp_all / 2); fsc2 = scalbnf(1.f, fexp_all - fexp_all / 2);
    }
    long long dbg_epi = 0;
    const int nvt = a.ntiles * a.nseg;
    for (int u = 0; u < nvt; ++u) {
      const int t = u / a.nseg;
      const bool accum = (u % a.nseg) != 0;   // a later K segment of the same column tile: add to what is there
      const int set = a.dbuf ? (u & 1) : 0, nuse = a.dbuf ? (u >> 1) : u;
      const uint32_t tmem_main = tmem_base + (uint32_t)(2 * BN * set), tmem_small = tmem_main + (uint32_t)BN;
      tc::mbar_wait(bar_accfull + 8 * set, (uint32_t)(nuse & 1));
      long long te0 = TCG_CLK();
      tc::tc_fence_after();
      const int n0 = t * BN;
      if (MODE == MODE_DKR2) {
        db = n0 / O;  // BN % O == 0 (checked on the host): every tile starts at o == 0
      }
#pragma unroll 1
      for (int cb = 0; cb < BN; cb += 32) {
        if (n0 + cb >= a.Ncols) break;   // padding columns of the last tile (warp-uniform): nothing to reduce or store
        if (MODE == MODE_DKR2 ? grp != 0 : ((cb >> 5) & 1) != grp) continue;
        if (MODE == MODE_FWD) {          // position of this batch's first column: output o, second-half index b
          fo = (n0 + cb) / g.Bn; fb = (n0 + cb) - fo * g.Bn;
        }
        float v[32];
        tc::tmem_ld32(tmem_main + lane_base + (uint32_t)cb, v);
        if (a.passes == 3) {
          float w[32];
          tc::tmem_ld32(tmem_small + lane_base + (uint32_t)cb, w);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = F16 ? fmaf(w[i], 1.f / 2048.f, v[i]) : v[i] + w[i];
        }
        if (MODE == MODE_STORE) {
          // every K segment has its own output slice (plain stores; a read-modify-write of the slice was measured at 4x
          // the whole kernel): sum_slices_kernel adds them with coalesced accesses
          const int row0 = pl0 + quad * 32;
          int ncols = a.Ncols - n0 - cb;
          if (BN - cb < ncols) ncols = BN - cb;
          if (ncols > 32) ncols = 32;
          store_tile(tst, v, F16 ? sc1 : 1.f, F16 ? sc2 : 1.f,
                     a.out + (long long)(u % a.nseg) * a.seg_stride + (long long)row0 * a.ldc + n0 + cb, a.ldc, a.np - row0, ncols,
                     false, lane);
        } else if (MODE == MODE_FWD) {
          if (a.tsave != nullptr) {   // keep T for the input gradient (dctn_eps_forward_train)
            const int row0 = pl0 + quad * 32;
            int ncols = a.Ncols - n0 - cb;
            if (BN - cb < ncols) ncols = BN - cb;
            if (ncols > 32) ncols = 32;
            store_tile(tst, v, F16 ? sc1 : 1.f, F16 ? sc2 : 1.f,
                       a.tsave + (pt0 + quad * 32) * (long long)a.Ncols + n0 + cb, a.Ncols, a.np - row0, ncols, accum, lane);
          }
          // columns cb..cb+31 are b = fb, fb+1, ... (wrapping to the next o at b == Bn; Bn >= 32: at most one wrap)
          int nvalid = BN - cb;
          if (a.Ncols - n0 - cb < nvalid) nvalid = a.Ncols - n0 - cb;
          if (nvalid > 32) nvalid = 32;
          const int wrap = g.Bn - fb;          // first column of this batch that belongs to the next o
          const uint32_t* bi = eidx + fb;      // doubled table: fb + 31 < 2*Bn
          float kr[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const uint32_t id = bi[i];
            kr[i] = eH[(id & 0xFFFF) * 128] * eL[(id >> 16) * 128];
          }
          if (g.Bn >= 32) {
            float s2 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float c = (i < nvalid) ? v[i] * kr[i] : 0.f;
              if (i < wrap) s += c; else s2 += c;
            }
            if (wrap <= nvalid) {
              outs_g[fo * 128 + pr] += s;
              s = s2; ++fo; fb = fb + nvalid - g.Bn;
            } else {
              fb += nvalid;
            }
          } else {
            // narrow second half (lopsided split, e.g. one factor: Bn = Q_in): several outputs per batch of 32 columns;
            // the run boundaries are warp-uniform
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (i < nvalid) {
                s = fmaf(v[i], kr[i], s);
                if (++fb == g.Bn) {
                  outs_g[fo * 128 + pr] += s;
                  s = 0.f; fb = 0; ++fo;
                }
              }
            }
          }
          if (fo < O) outs_g[fo * 128 + pr] += s;   // partial sum of this batch (the next batch recomputes its position)
          s = 0.f;
        } else {  // MODE_DKR2
          const uint4* cp = (const uint4*)(eidx + cb);
          uint32_t id[32];
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const uint4 u = cp[q4];
            id[4 * q4] = u.x; id[4 * q4 + 1] = u.y; id[4 * q4 + 2] = u.z; id[4 * q4 + 3] = u.w;
          }
          float gv[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) gv[i] = tabE[(id[i] & 0xFF) * 128 + pr];
          int nvalid = BN - cb;
          if (a.Ncols - n0 - cb < nvalid) nvalid = a.Ncols - n0 - cb;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i < nvalid) {
              s = fmaf(v[i], gv[i], s);
              if (id[i] & 0x100u) {
                if (pvalid) {
                  float* dst = a.out + (long long)pl * g.Bn + db;
                  const float r1 = F16 ? s * sc1 * sc2 : s;
                  *dst = accum ? *dst + r1 : r1;
                }
                s = 0.f; ++db;
              }
            }
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(bar_accempty + 8 * set);
      dbg_epi += TCG_CLK() - te0;
    }
    if (a.dbg && warp == 8 && lane == 0) a.dbg[(long long)blockIdx.x * 8 + 7] = dbg_epi;
    if (MODE == MODE_FWD) {
      asm volatile("bar.sync 1, %0;" ::"n"(32 * G_EPI_WARPS) : "memory");   // both groups' partial sums are in shared memory
      if (grp == 0 && pvalid) {
        float* orow = a.out + (pt0 + pr) * O;
        for (int o = 0; o < O; ++o) {
          const float r1 = outs[o * 128 + pr] + outs[(O + o) * 128 + pr];
          orow[o] = F16 ? r1 * fsc1 * fsc2 : r1;
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ host side
struct GemmShape {
  int jh0, nf, withG, Ncols, Kdim;   // generated operand: factors [jh0, jh0 + nf) (times gout if withG); true K extent
  // K order of the generated operand and of the packed core: k' = (ob * Hpad + h) * RB + rl, where
  //   r = ob * RB + rl < G indexes the REGISTER GROUP — the fastest-varying `cr` factors of the operand (times gout):
  //       a producer thread holds the RB values tabR[ob*RB ..] of its patch in registers for a whole section `ob`,
  //   h < H indexes the remaining factors, split into two table groups: value = tabKH[h / KLb] * tabKL[h % KLb].
  // A generated element costs one multiplication (plus 2 shared-memory loads per RB elements) instead of two or three
  // loads; the price is padding (h to Hpad, r to GP: those rows of the packed core are zero).
  int cr, G, GP, RB, cnth, KH, cntl, KLb, H, Hpad, Kp;
};

inline int ipow_i(int b, int e) { int r = 1; while (e-- > 0) r *= b; return r; }

inline GemmShape shape_for(const EpsGeom& g, int mode) {
  GemmShape s{};
  if (mode == MODE_STORE) {  // Gen = KR2 x gout over k = (b, o); columns = a
    s.jh0 = g.m; s.nf = g.n - g.m; s.Kdim = g.N; s.withG = 1; s.Ncols = g.A;
  } else {                   // Gen = KR1 over k = a; columns = n
    s.jh0 = 0; s.nf = g.m; s.Kdim = g.A; s.withG = 0; s.Ncols = g.N;
  }
  // register group: the choice with the least padded K; ties: wider blocks, then the larger group (smaller tables)
  long long best = -1;
  for (int cr = 0; cr <= s.nf; ++cr) {
    const long long G = (long long)ipow_i(g.Q, cr) * (s.withG ? g.O : 1);
    if (G > 64) break;
    const int H = ipow_i(g.Q, s.nf - cr);
    for (int RB = 8; RB >= 4; RB -= 4) {
      const int GP = ((int)G + RB - 1) / RB * RB, hq = 32 / RB, Hpad = (H + hq - 1) / hq * hq;
      const long long cost = (long long)GP * Hpad;
      if (best < 0 || cost < best || (cost == best && (RB > s.RB || (RB == s.RB && cr > s.cr)))) {
        best = cost; s.cr = cr; s.G = (int)G; s.GP = GP; s.RB = RB; s.H = H; s.Hpad = Hpad;
      }
    }
  }
  if (best < 0) {   // gout alone wider than 64 values: register blocks over gout only
    s.cr = 0; s.G = g.O; s.RB = 8; s.GP = (g.O + 7) / 8 * 8; s.H = ipow_i(g.Q, s.nf); s.Hpad = (s.H + 3) / 4 * 4;
  }
  const int nrem = s.nf - s.cr;
  s.cntl = nrem / 2; s.cnth = nrem - s.cntl;
  s.KH = ipow_i(g.Q, s.cnth); s.KLb = ipow_i(g.Q, s.cntl);
  s.Kp = s.GP * s.Hpad;      // multiple of 32
  return s;
}

// the K extent of a stage is 64 (fp16) or 32 (tf32); one tile-width choice serves both arithmetics
inline size_t gemm_fixed_smem(const EpsGeom& g, const GemmShape& s, int mode) {
  const int nE = (mode == MODE_FWD) ? (g.BH + g.BL) : (mode == MODE_DKR2 ? g.O : 0);
  const size_t ktab = (size_t)(s.KH + 1 + s.KLb + s.GP);
  const size_t neidx = (mode == MODE_FWD) ? (size_t)g.Bn + (g.Bn > 32 ? g.Bn : 32) : (mode == MODE_DKR2 ? (size_t)MAX_BN : 0);
  return 1024 + (ktab + nE + (mode == MODE_FWD ? 2 * g.O : 0)) * 128 * 4 + ((size_t)s.Hpad + neidx + 4) * 4 + 256 * 4 +
         (mode == MODE_DKR2 ? 0 : G_EPI_WARPS * TST_FLOATS * 4) + (2 * MAX_BSTAGES + 2 * ASTAGES + 4) * 8 + 16;
}
inline size_t bstage_bytes(int BN) { return 2 * (size_t)BN * 128; }   // hi + lo parts, BN rows of 128 bytes
inline int stage_k(int passes) { return passes == ARITH_F16X3 ? GBK16 : GBK; }
inline GemmShape shape_auto(const EpsGeom& g, int mode) { return shape_for(g, mode); }

// Two accumulator sets for the forward GEMM when K is shallow: its epilogue (KR2 reduction, the store of T) then costs
// more than the MMAs of a tile (CIFAR (2, 23 -> 24): 11.0k against 8.6k cycles per 160-column tile, and the T stores of
// all SMs arrive at HBM in the same burst); alternating sets hide it behind the next tile's MMAs.  The price: tiles of
// at most 96 columns (4 * BN + 128 <= 512 TMEM columns), i.e. the operand is generated more often, which a deep K does
// not repay.  DCTN_B200_DBUF_MAXK overrides the K bound (0 = never).
inline bool gemm_dbuf(const EpsGeom& g, int mode) {
  if (mode != MODE_FWD) return false;
  int maxk = 28 * GBK16;
  if (const char* e = getenv("DCTN_B200_DBUF_MAXK")) maxk = atoi(e);
  const GemmShape s = shape_auto(g, mode);
  return s.Kp <= maxk && s.Ncols <= 384;   // few tiles: the extra generation passes stay cheap
}

// number of shared-memory B stages that fit (0 = does not fit); the setup scratch (x and gout of 128 patches)
// is aliased onto the stages and must fit too
inline int pick_bstages(const EpsGeom& g, int mode, int BN) {
  const GemmShape s = shape_auto(g, mode);
  const size_t fixed = gemm_fixed_smem(g, s, mode);
  if (fixed >= TCG_SMEM_LIMIT) return 0;
  int nb = (int)((TCG_SMEM_LIMIT - fixed) / bstage_bytes(BN));
  if (nb > MAX_BSTAGES) nb = MAX_BSTAGES;
  if (const char* e = getenv("DCTN_B200_MAX_BSTAGES")) {   // experiments only
    const int v = atoi(e);
    if (v >= 2 && nb > v) nb = v;
  }
  if (nb < 2) return 0;
  if ((size_t)(g.n * g.Q + g.O + g.n + 1) * 128 * 4 > nb * bstage_bytes(BN)) return 0;   // x, gout, exponents
  return nb;
}

// pick the column-tile width (multiple of 16, <= MAX_BN): satisfies the epilogue's alignment (MODE_DKR2 needs whole
// (b, o) groups per tile), fits shared memory with >= 2 stages, least padded width, then widest
inline int pick_bn(const EpsGeom& g, int mode) {
  const GemmShape s = shape_auto(g, mode);
  int best = 0;
  long long best_cost = 0;
  for (int bn = gemm_dbuf(g, mode) ? 96 : MAX_BN; bn >= 64; bn -= 16) {
    if (mode == MODE_DKR2 && bn % g.O != 0) continue;
    if (pick_bstages(g, mode, bn) == 0) continue;
    long long cost = (long long)((s.Ncols + bn - 1) / bn) * bn;
    if (bn < 160) cost = cost * 9 / 8;  // narrow tiles: relatively more epilogue / barrier overhead per MMA
    if (!best || cost < best_cost) { best = bn; best_cost = cost; }
  }
  return best;
}

// size of the packed core image (TF32 layout; the fp16 image is half as large and uses the same buffer)
inline size_t packed_floats(const EpsGeom& g, int mode, int BN) {
  const GemmShape s = shape_auto(g, mode);
  long long ntiles = (s.Ncols + BN - 1) / BN, nk = (s.Kp + GBK - 1) / GBK;
  return (size_t)(ntiles * nk * 2 * BN * 32);
}
constexpr size_t WS_HEADER = 256;   // first bytes of every workspace: bits of max|core| (fp16 arithmetic)

// patches per launch of the input-gradient GEMMs: their outputs dKR1 / dKR2 go through a scratch buffer of at most
// 2 GiB (HBM write + read at ~6.5 TB/s costs far less than the launch gaps and partial waves of many small chunks)
inline long long dx_patch_chunk(const EpsGeom& g) {
  long long target = 2048ll << 20;
  long long pc = target / (((long long)g.A + g.Bn) * 4);
  if (pc < 4096) pc = 4096;
  // whole waves: one CTA per 128 patches, 148 CTAs resident at a time
  const long long wave = 148ll * GBM;
  if (pc >= wave) pc = (pc / wave) * wave;
  else pc = (pc / 128) * 128;
  if (pc > g.P) pc = g.P;
  return pc;
}

template <int MODE, bool F16>
int launch_gemm_inst(const TcGemmArgs& a, size_t smem, cudaStream_t st) {
  auto k = tc_gemm_kernel<MODE, F16>;
  DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<(a.np + GBM - 1) / GBM, G_THREADS, smem, st>>>(a);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}

// `passes`: 1 / 3 = TF32 passes, ARITH_F16X3 = split fp16 (absmax: the slot absmax_kernel filled before run_pack)
int run_gemm(const EpsGeom& g, int mode, int BN, const float* x, const float* gout, const float* packed, long long p0,
             int np, float* out, long long ldc, int passes, cudaStream_t st, const uint32_t* absmax, float* tsave = nullptr,
             long long seg_stride = 0) {
  const GemmShape s = shape_auto(g, mode);
  const bool f16 = passes == ARITH_F16X3;
  const int KS = stage_k(passes);
  TcGemmArgs a{};
  a.g = g; a.x = x; a.gout = gout; a.p0 = p0; a.np = np;
  a.jh0 = s.jh0; a.nf = s.nf; a.withG = s.withG; a.cnth = s.cnth; a.KH = s.KH; a.cntl = s.cntl; a.KLb = s.KLb;
  a.cr = s.cr; a.G = s.G; a.GP = s.GP; a.RB = s.RB; a.H = s.H; a.Hpad = s.Hpad; a.Kp = s.Kp;
  a.Ncols = s.Ncols; a.ntiles = (s.Ncols + BN - 1) / BN; a.nk = (s.Kp + KS - 1) / KS;
  // fp16 stages hold 64 K-values (4 MMAs), tf32 stages 32 (also 4 MMAs): the same number of accumulation steps
  a.kseg = tc::seg_stages(a.nk); a.nseg = (a.nk + a.kseg - 1) / a.kseg;
  a.seg_stride = seg_stride;
  a.dbuf = gemm_dbuf(g, mode) ? 1 : 0;
  a.packed = packed; a.BN = BN; a.bstages = pick_bstages(g, mode, BN); a.passes = f16 ? 3 : passes; a.out = out; a.ldc = ldc;
  a.core_absmax = absmax;
  a.tsave = tsave;
  a.dbg = nullptr;
#ifdef DCTN_TCG_TIMING   // cycle probes: timing builds only (allocates, synchronises, not thread-safe)
  static long long* dbg_buf = nullptr;
  const char* dbg_env = getenv("DCTN_TCG_DEBUG");
  const int ncta = (np + GBM - 1) / GBM;
  if (dbg_env && ncta <= 4096) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 4096 * 8 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 4096 * 8 * sizeof(long long), st);
    a.dbg = dbg_buf;
  }
#endif
  const size_t smem = gemm_fixed_smem(g, s, mode) + a.bstages * bstage_bytes(BN);
#ifdef DCTN_TCG_TIMING
  if (getenv("DCTN_DEBUG_SHAPE"))
    fprintf(stderr, "[tcg shape] mode=%d BN=%d NB=%d smem=%zu nf=%d cr=%d G=%d GP=%d RB=%d H=%d Hpad=%d KH=%d KLb=%d Kp=%d nk=%d ntiles=%d kseg=%d dbuf=%d\n",
            mode, BN, a.bstages, smem, a.nf, a.cr, a.G, a.GP, a.RB, a.H, a.Hpad, a.KH, a.KLb, a.Kp, a.nk, a.ntiles, a.kseg, a.dbuf);
#endif
  int rc;
  if (mode == MODE_STORE) rc = f16 ? launch_gemm_inst<MODE_STORE, true>(a, smem, st) : launch_gemm_inst<MODE_STORE, false>(a, smem, st);
  else if (mode == MODE_FWD) rc = f16 ? launch_gemm_inst<MODE_FWD, true>(a, smem, st) : launch_gemm_inst<MODE_FWD, false>(a, smem, st);
  else rc = f16 ? launch_gemm_inst<MODE_DKR2, true>(a, smem, st) : launch_gemm_inst<MODE_DKR2, false>(a, smem, st);
#ifdef DCTN_TCG_TIMING
  if (a.dbg && rc == 0) {
    static long long host[4096 * 8];
    cudaStreamSynchronize(st);
    cudaMemcpy(host, dbg_buf, (size_t)ncta * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
    double sum[8] = {0};
    for (int c = 0; c < ncta; ++c) for (int k = 0; k < 8; ++k) sum[k] += (double)host[c * 8 + k];
    const double nst = (double)a.ntiles * a.nk;
    fprintf(stderr, "[tcg dbg] mode=%d BN=%d NB=%d ntiles=%d nk=%d per-stage cycles: mma waitA %.0f waitB %.0f waitAcc(per tile) %.0f total %.0f | "
            "producer wait %.0f st %.0f gen %.0f | epilogue/tile %.0f\n", mode, BN, a.bstages, a.ntiles, a.nk,
            sum[0] / ncta / nst, sum[1] / ncta / nst, sum[2] / ncta / a.ntiles, sum[3] / ncta / nst,
            sum[4] / ncta / nst, sum[5] / ncta / nst, sum[6] / ncta / nst, sum[7] / ncta / a.ntiles);
  }
#endif
  return rc;
}

// K segments of the generic MODE_STORE GEMM (its output has one slice per segment)
inline int store_segments(const EpsGeom& g, int passes) {
  const int nk = (shape_auto(g, MODE_STORE).Kp + stage_k(passes) - 1) / stage_k(passes);
  const int kseg = tc::seg_stages(nk);
  return (nk + kseg - 1) / kseg;
}
// out[i] = sum_s slices[s * stride + i], i < count (fixed order; in place on slice 0)
__global__ void __launch_bounds__(256) sum_slices_kernel(float* __restrict__ slices, long long stride, long long count, int nslices) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    float s = slices[i];
    for (int k = 1; k < nslices; ++k) s += slices[(long long)k * stride + i];
    slices[i] = s;
  }
}

// max|core| -> ws header (fp16 arithmetic only); one call per entry point, before the packs
int run_absmax(const EpsGeom& g, const float* core, uint32_t* slot, int passes, cudaStream_t st) {
  if (passes != ARITH_F16X3) return 0;
  DCTN_CUDA_CHECK_RET(cudaMemsetAsync(slot, 0, sizeof(uint32_t), st));
  const long long n = (long long)g.A * g.N;
  int blocks = (int)((n + 1023) / 1024);
  if (blocks > 148 * 8) blocks = 148 * 8;
  absmax_kernel<<<blocks, 256, 0, st>>>(core, n, slot);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}

int run_pack(const EpsGeom& g, int mode, int BN, const float* core, float* dst, int passes, cudaStream_t st, const uint32_t* absmax) {
  const GemmShape s = shape_auto(g, mode);
  const bool f16 = passes == ARITH_F16X3;
  const int KS = stage_k(passes);
  const int ntiles = (s.Ncols + BN - 1) / BN, nk = (s.Kp + KS - 1) / KS;
  long long total = (long long)ntiles * nk * BN * 8;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  const KOrder ko{s.H, s.Hpad, s.G, s.RB, s.Kp};
  if (f16) pack_core_kernel<true><<<blocks, 256, 0, st>>>(core, dst, g, mode, BN, s.Ncols, ko, ntiles, nk, 3, absmax);
  else pack_core_kernel<false><<<blocks, 256, 0, st>>>(core, dst, g, mode, BN, s.Ncols, ko, ntiles, nk, passes, absmax);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}

inline bool common_ok(const EpsGeom& g) {
  if (g.P < 2048) return false;                       // tiny problems are launch-bound: CUDA-core family
  if (g.P >= (1ll << 31) / (g.Q > g.O ? g.Q : g.O)) return false;
  if (g.A < 64 || g.N < 64) return false;             // tiles would be mostly padding
  return true;
}

}  // namespace

// register-table kernels (eps_tc_fast.cu) serve the split-fp16 arithmetic when the shape allows; DCTN_B200_NO_FAST=1
// forces the generic table-lookup kernels (A/B comparisons, tests of the generic path)
static bool use_fast(const EpsGeom& g, int fmode, int passes) {
  if (passes != ARITH_F16X3 || !tcfast_supported(g, fmode)) return false;
  const char* e = getenv("DCTN_B200_NO_FAST");
  return !(e && e[0] == '1');
}

bool tcg_supported(const EpsGeom& g, int kind) {
  if (!common_ok(g)) return false;
  if (kind == 0) return pick_bn(g, MODE_FWD) != 0;
  if (kind == 2) return (g.n - g.m) > 0 && pick_bn(g, MODE_STORE) != 0 && pick_bn(g, MODE_DKR2) != 0;
  if (kind == 3) return (g.n - g.m) > 0 && pick_bn(g, MODE_FWD) != 0 && pick_bn(g, MODE_STORE) != 0;
  return false;
}

size_t tcg_workspace_bytes(const EpsGeom& g, int kind) {
  if (kind == 0) {
    size_t pf = packed_floats(g, MODE_FWD, pick_bn(g, MODE_FWD)), ff = tcfast_packed_floats(g, 1);
    return WS_HEADER + (pf > ff ? pf : ff) * 4 + 256;
  }
  if (kind == 2 || kind == 3) {
    const long long pc = dx_patch_chunk(g);
    size_t p1 = packed_floats(g, MODE_STORE, pick_bn(g, MODE_STORE)), f1 = tcfast_packed_floats(g, 0);
    if (tcfast_packed_floats(g, 2) > f1) f1 = tcfast_packed_floats(g, 2);
    if (tcfast_packed_floats(g, 3) > f1) f1 = tcfast_packed_floats(g, 3);
    const size_t nslice = (size_t)(store_segments(g, 3) > store_segments(g, ARITH_F16X3) ? store_segments(g, 3) : store_segments(g, ARITH_F16X3));
    size_t f = (p1 > f1 ? p1 : f1) + 64 + (size_t)pc * ((size_t)g.A * nslice + g.Bn) + (size_t)g.P * g.n * g.Q;
    if (kind == 2) f += packed_floats(g, MODE_DKR2, pick_bn(g, MODE_DKR2));
    return WS_HEADER + f * 4 + 1024;
  }
  return 0;
}

int tc_forward(const EpsGeom& g, const float* x, const float* core, float* out, void* ws, int passes, cudaStream_t st,
               float* tsave) {
  const int BN = pick_bn(g, MODE_FWD);
  if (!BN) return dctn_set_error(-2, "tcgen05 forward kernel does not support this shape");
  uint32_t* absmax = (uint32_t*)ws;
  float* packed = (float*)((char*)ws + WS_HEADER);
  int rc = run_absmax(g, core, absmax, passes, st);
  if (rc) return rc;
  if (use_fast(g, 1, passes)) {
    if ((rc = tcfast_pack(g, 1, core, packed, absmax, st))) return rc;
    return tcfast_gemm(g, 1, x, nullptr, packed, absmax, 0, (int)g.P, out, 0, tsave, st);
  }
  if ((rc = run_pack(g, MODE_FWD, BN, core, packed, passes, st, absmax))) return rc;
  return run_gemm(g, MODE_FWD, BN, x, nullptr, packed, 0, (int)g.P, out, 0, passes, st, absmax, tsave);
}

size_t tcg_saved_bytes(const EpsGeom& g) { return (size_t)g.P * (size_t)g.N * sizeof(float); }

namespace {
// dKR2[p][b] = sum_o T[p][o*Bn + b] * gout[p][o] from the rows saved by the training forward; one thread per
// (patch, 4 consecutive b): 128-bit coalesced reads of T (the only large stream: P*N floats, HBM-bound)
__global__ void __launch_bounds__(256) dkr2_from_saved_kernel(const float* __restrict__ T, const float* __restrict__ gout,
                                                              float* __restrict__ dkr2, long long p0, int np, int Bn, int O) {
  const int b4n = Bn >> 2;
  const long long total = (long long)np * b4n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int pl = (int)(i / b4n), b4 = (int)(i - (long long)pl * b4n);
    const long long p = p0 + pl;
    const float4* t = (const float4*)(T + p * (long long)Bn * O) + b4;
    const float* gr = gout + p * O;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int o = 0; o < O; ++o) {
      const float4 v = __ldcs(t + (long long)o * b4n);   // streamed once: do not keep in L2
      const float gv = __ldg(gr + o);
      s.x = fmaf(v.x, gv, s.x); s.y = fmaf(v.y, gv, s.y); s.z = fmaf(v.z, gv, s.z); s.w = fmaf(v.w, gv, s.w);
    }
    *((float4*)(dkr2 + (long long)pl * Bn) + b4) = s;
  }
}
__global__ void __launch_bounds__(256) dkr2_from_saved_scalar_kernel(const float* __restrict__ T, const float* __restrict__ gout,
                                                                     float* __restrict__ dkr2, long long p0, int np, int Bn, int O) {
  const long long total = (long long)np * Bn;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int pl = (int)(i / Bn), b = (int)(i - (long long)pl * Bn);
    const long long p = p0 + pl;
    const float* t = T + p * (long long)Bn * O + b;
    float s = 0.f;
    for (int o = 0; o < O; ++o) s = fmaf(t[(long long)o * Bn], __ldg(gout + p * O + o), s);
    dkr2[i] = s;
  }
}
// Second half of the input gradient from the saved T, complete: one warp per patch
//   dKR2[b]  = sum_o T[p][o*Bn + b] * gout[p][o]                 (coalesced 128-bit reads of the only large stream)
//   Whi[eh]  = sum_el dKR2[eh*BL + el] * TL[el],   Wlo[el] = sum_eh dKR2[eh*BL + el] * TH[eh]       (stage 1)
//   d x_j[q] = sum_{e: digit_t(e) = q} W[e] * prod_{t' != t} x_{j'}[digit_t'(e)]                       (stage 2)
// everything after the read of T lives in the warp's slice of shared memory; dxp[p][j][q] for the factors j >= m.
// Division-free digit arithmetic for the per-patch leave-one-out code: qd[e] = e / Q for e < nqd, built once per CTA.
// The digit loops of these kernels divided by the run-time Q two or three times per table entry and factor (~40
// instructions each): at K = 3, Q = 3 the per-group stage alone was ~3000 instructions per patch.
__host__ __device__ __forceinline__ int loo_nqd(int EH, int EL, int nfq, int extra) {
  int m = EH > EL ? EH : EL;
  if (nfq > m) m = nfq;
  if (extra > m) m = extra;
  return m + 1;
}
__device__ __forceinline__ void build_qd(int* qd, int nqd, int Q) {
  for (int e = threadIdx.x; e < nqd; e += blockDim.x) qd[e] = e / Q;
  __syncthreads();
}
// prod over the digits u of entry e (digit 0 slowest) of xg[u*Q + digit_u], leaving out position `skip` (-1: none)
__device__ __forceinline__ float kr_prod(const float* xg, const int* qd, int Q, int cnt, int e, int skip) {
  float v = 1.f;
  for (int u = cnt - 1; u >= 0; --u) {
    const int e1 = qd[e], d = e - e1 * Q;
    e = e1;
    if (u != skip) v *= xg[u * Q + d];
  }
  return v;
}
// per-group stage of the leave-one-out contraction for one warp:
//   d x_t[q] = sum_{e: digit_tt(e) = q} W[e] * prod_{u != tt} x_u[digit_u(e)],  W = wH (hi group) | wL (lo group)
__device__ __forceinline__ void loo_group_stage(const float* wH, const float* wL, const float* xs, const int* qd, int Q, int cnth,
                                                int cntl, int EH, int EL, int lane, float* __restrict__ dst /* [nf][Q] */) {
  const int nf = cnth + cntl;
  for (int item = lane; item < nf * Q; item += 32) {
    const int t = qd[item], q = item - t * Q;
    const bool in_hi = t < cnth;
    const int cnt = in_hi ? cnth : cntl, tt = in_hi ? t : t - cnth;
    const float* w = in_hi ? wH : wL;
    const float* xg = xs + (in_hi ? 0 : cnth) * Q;
    int dstride = 1, npre = 1;
    for (int u = 0; u < cnt - 1 - tt; ++u) dstride *= Q;
    for (int u = 0; u < tt; ++u) npre *= Q;
    float sacc = 0.f;
    for (int hp = 0; hp < npre; ++hp) {
      const int e0 = (hp * Q + q) * dstride;
      for (int lp = 0; lp < dstride; ++lp) sacc = fmaf(w[e0 + lp], kr_prod(xg, qd, Q, cnt, e0 + lp, tt), sacc);
    }
    dst[item] = sacc;
  }
}
// patch_origin with 32-bit divisions (the tcgen05 paths require P < 2^31)
__device__ __forceinline__ long long patch_origin32(const EpsGeom& g, long long p) {
  const unsigned hw = (unsigned)(g.Ho * g.Wo), pu = (unsigned)p;
  const unsigned b = pu / hw, r = pu - b * hw;
  const unsigned h = r / (unsigned)g.Wo, w = r - h * (unsigned)g.Wo;
  return (((long long)b * g.H + h) * (long long)g.W + w) * g.Q;
}

constexpr int LOO2_WARPS = 8;
template <bool VEC>
__global__ void __launch_bounds__(32 * LOO2_WARPS) loo2_from_saved_kernel(EpsGeom g, const float* __restrict__ x, const float* __restrict__ T,
                                                                          const float* __restrict__ gout, float* __restrict__ dxp,
                                                                          long long p0, int np) {
  extern __shared__ float l2_smem[];
  const int Q = g.Q, O = g.O, Bn = g.Bn, BH = g.BH, BL = g.BL, nf = g.n - g.m;
  const int BLS = BL | 1;                                   // padded row stride of the dKR2 matrix [BH][BL]
  const int per_warp = BH * BLS + 2 * (BH + BL) + nf * Q + O;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqd = loo_nqd(BH, BL, nf * Q, Bn);
  int* qd = (int*)(l2_smem + LOO2_WARPS * per_warp);
  build_qd(qd, nqd, Q);
  float* dk = l2_smem + warp * per_warp;
  float* tH = dk + BH * BLS;   // [BH] then tL [BL]
  float* tL = tH + BH;
  float* wH = tL + BL;         // [BH] then wL [BL]
  float* wL = wH + BH;
  float* xs = wL + BL;
  float* gs = xs + nf * Q;
  for (long long pl = (long long)blockIdx.x * LOO2_WARPS + warp; pl < np; pl += (long long)gridDim.x * LOO2_WARPS) {
    const long long p = p0 + pl;
    const long long o0 = patch_origin32(g, p);
    for (int i = lane; i < nf * Q; i += 32) {
      const int j = qd[i];
      xs[i] = __ldg(&x[o0 + g.foff[g.m + j] + (i - j * Q)]);
    }
    for (int o = lane; o < O; o += 32) gs[o] = __ldg(&gout[p * O + o]);
    __syncwarp();
    for (int e = lane; e < BH + BL; e += 32) {
      const bool hi = e < BH;
      tH[e] = kr_prod(xs + (hi ? 0 : g.b_nh) * Q, qd, Q, hi ? g.b_nh : g.b_nl, hi ? e : e - BH, -1);   // tL follows tH
    }
    const float* trow = T + p * (long long)Bn * O;
    if (VEC) {     // Bn % 4 == 0 and BL % 4 == 0: four consecutive b stay in one row of the [BH][BL] matrix
      // two chunks of 32 x float4 per pass with all their loads issued first: enough bytes in flight per warp to
      // keep HBM busy (one warp has only this patch's 6 KB to read)
      const int b4n = Bn >> 2;
      for (int b4 = lane; b4 < b4n; b4 += 64) {
        const bool two = b4 + 32 < b4n;
        float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
        for (int o0 = 0; o0 < O; o0 += 4) {
          float4 va[4], vb[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            va[k] = vb[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (o0 + k < O) {
              const float4* src = (const float4*)(trow + (long long)(o0 + k) * Bn) + b4;
              va[k] = __ldcs(src);
              if (two) vb[k] = __ldcs(src + 32);
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float gv = (o0 + k < O) ? gs[o0 + k] : 0.f;
            acc0.x = fmaf(va[k].x, gv, acc0.x); acc0.y = fmaf(va[k].y, gv, acc0.y); acc0.z = fmaf(va[k].z, gv, acc0.z); acc0.w = fmaf(va[k].w, gv, acc0.w);
            acc1.x = fmaf(vb[k].x, gv, acc1.x); acc1.y = fmaf(vb[k].y, gv, acc1.y); acc1.z = fmaf(vb[k].z, gv, acc1.z); acc1.w = fmaf(vb[k].w, gv, acc1.w);
          }
        }
        {
          const int b = b4 << 2;
          float* d = dk + (b / BL) * BLS + b % BL;
          d[0] = acc0.x; d[1] = acc0.y; d[2] = acc0.z; d[3] = acc0.w;
        }
        if (two) {
          const int b = (b4 + 32) << 2;
          float* d = dk + (b / BL) * BLS + b % BL;
          d[0] = acc1.x; d[1] = acc1.y; d[2] = acc1.z; d[3] = acc1.w;
        }
      }
    } else {
      // two chunks of 32 b and eight outputs per pass, all sixteen loads issued before the first use: a loop that
      // consumes each load at once waits one HBM latency per element (K = 3, Q = 3: 18 dependent loads per patch,
      // 657 us for 657 MB at B = 512)
      for (int b = lane; b < Bn; b += 64) {
        const bool two = b + 32 < Bn;
        float s0 = 0.f, s1 = 0.f;
        for (int o0 = 0; o0 < O; o0 += 8) {
          float va[8], vb[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            va[k] = vb[k] = 0.f;
            if (o0 + k < O) {
              const float* src = trow + (long long)(o0 + k) * Bn + b;
              va[k] = __ldcs(src);
              if (two) vb[k] = __ldcs(src + 32);
            }
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float gv = (o0 + k < O) ? gs[o0 + k] : 0.f;
            s0 = fmaf(va[k], gv, s0);
            s1 = fmaf(vb[k], gv, s1);
          }
        }
        dk[(b / BL) * BLS + b % BL] = s0;
        if (two) dk[((b + 32) / BL) * BLS + (b + 32) % BL] = s1;
      }
    }
    __syncwarp();
    for (int eh = lane; eh < BH; eh += 32) {
      float sacc = 0.f;
      for (int el = 0; el < BL; ++el) sacc = fmaf(dk[eh * BLS + el], tL[el], sacc);
      wH[eh] = sacc;
    }
    for (int el = lane; el < BL; el += 32) {
      float sacc = 0.f;
      for (int eh = 0; eh < BH; ++eh) sacc = fmaf(dk[eh * BLS + el], tH[eh], sacc);
      wL[el] = sacc;
    }
    __syncwarp();
    loo_group_stage(wH, wL, xs, qd, Q, g.b_nh, g.b_nl, BH, BL, lane, dxp + (p * g.n + g.m) * Q);
    __syncwarp();
  }
}

// Leave-one-out stage 2: from W[p] = (Whi[EH] | Wlo[EL]) of one half to d x_j for the factors j of that half,
//   d x_j[q] (j at position t of a group with table entries e) = sum_{e: digit_t(e) = q} W[e] * prod_{t' != t} x_{j'}[digit_t'(e)]
// one thread per (patch, factor, q); dxp[p][j][q]
__global__ void __launch_bounds__(256) loo_groups_kernel(EpsGeom g, const float* __restrict__ x, const float* __restrict__ W, int ldw,
                                                         long long p0, int np, int j0, int cnth, int EH, int cntl, int EL,
                                                         float* __restrict__ dxp) {
  const int Q = g.Q, nf = cnth + cntl;
  const long long total = (long long)np * nf * Q;
  for (long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x; item < total; item += (long long)gridDim.x * blockDim.x) {
    const long long pl = item / (nf * Q);
    const int r = (int)(item - pl * nf * Q);
    const int t = r / Q, q = r - t * Q;
    const long long p = p0 + pl;
    const long long o0 = patch_origin(g, p);
    const bool in_hi = t < cnth;
    const int cnt = in_hi ? cnth : cntl, tt = in_hi ? t : t - cnth, Eg = in_hi ? EH : EL;
    const int jb = j0 + (in_hi ? 0 : cnth);
    const float* w = W + pl * (long long)ldw + (in_hi ? 0 : EH);
    int dstride = 1;
    for (int u = 0; u < cnt - 1 - tt; ++u) dstride *= Q;
    float sacc = 0.f;
    const int others = Eg / Q;
    for (int oe = 0; oe < others; ++oe) {
      const int lo_part = oe % dstride, hi_part = oe / dstride;
      const int e = (hi_part * Q + q) * dstride + lo_part;
      float v = w[e];
      int ee = e;
      for (int u = cnt - 1; u >= 0; --u) {
        const int d = ee % Q;
        ee /= Q;
        if (u != tt) v *= __ldg(&x[o0 + g.foff[jb + u] + d]);
      }
      sacc += v;
    }
    dxp[(p * g.n + j0 + t) * Q + q] = sacc;
  }
}

inline int launch_loo_groups(const EpsGeom& g, const float* x, const float* W, int ldw, long long p0, int np, int j0, int cnth, int EH,
                             int cntl, int EL, float* dxp, cudaStream_t st) {
  const long long total = (long long)np * (cnth + cntl) * g.Q;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  loo_groups_kernel<<<blocks, 256, 0, st>>>(g, x, W, ldw, p0, np, j0, cnth, EH, cntl, EL, dxp);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}
// First-half leave-one-out for the generic GEMM path: dKR1 (np x A, possibly as several K-segment slices that are summed
// here, in slice order) -> d x_j for the factors of the first half, one WARP per patch.
//   Whi[eh] = sum_el dKR1[eh*EL + el] * TL[el],   Wlo[el] = sum_eh dKR1[eh*EL + el] * TH[eh],   then the per-group stage.
// A lane owns ROWS eh = lane, lane + 32, ... of the [EH][EL] matrix: Whi[eh] is a private sum, Wlo[el] a private partial
// per lane (EL <= ELB registers) reduced across the warp once per patch — three instructions per matrix element and no
// per-element index arithmetic.  The 32 rows a warp reads together are contiguous (32 * EL floats): every sector that
// the first load of a row brings into L1 is used by the following ones.  Replaces sum_slices_kernel +
// loo_staged_kernel, whose staging loop spent four integer divisions per element (CIFAR (2, 12 -> 24): 522 us for a
// 425 MB matrix; a first warp-per-patch version with a segmented shuffle reduction per row issued 3000 instructions per
// patch at K = 3, Q = 3 — profiles/r02f_loo_ncu.txt).
constexpr int LOO1_WARPS = 8;
template <int ELB>
__global__ void __launch_bounds__(32 * LOO1_WARPS) loo1_rows_kernel(EpsGeom g, const float* __restrict__ x, const float* __restrict__ dkr,
                                                                    long long slice_stride, int nslices, long long p0, int np,
                                                                    float* __restrict__ dxp) {
  extern __shared__ float l1_smem[];
  const int Q = g.Q, EH = g.AH, EL = g.AL, cnth = g.a_nh, cntl = g.a_nl, nf = g.m, E = g.A;
  const int per_warp = 2 * EH + 2 * EL + nf * Q;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* qd = (int*)(l1_smem + LOO1_WARPS * per_warp);
  build_qd(qd, loo_nqd(EH, EL, nf * Q, 0), Q);
  float* tH = l1_smem + warp * per_warp;
  float* tL = tH + EH;
  float* wH = tL + EL;
  float* wL = wH + EH;
  float* xs = wL + EL;
  for (long long pl = (long long)blockIdx.x * LOO1_WARPS + warp; pl < np; pl += (long long)gridDim.x * LOO1_WARPS) {
    const long long p = p0 + pl;
    const long long org = patch_origin32(g, p);
    for (int i = lane; i < nf * Q; i += 32) {
      const int j = qd[i];
      xs[i] = __ldg(&x[org + g.foff[j] + (i - j * Q)]);
    }
    __syncwarp();
    for (int e = lane; e < EH; e += 32) tH[e] = kr_prod(xs, qd, Q, cnth, e, -1);
    for (int e = lane; e < EL; e += 32) tL[e] = kr_prod(xs + cnth * Q, qd, Q, cntl, e, -1);
    __syncwarp();
    float tlr[ELB], accl[ELB];
#pragma unroll
    for (int el = 0; el < ELB; ++el) {
      tlr[el] = el < EL ? tL[el] : 0.f;
      accl[el] = 0.f;
    }
    const float* mat = dkr + pl * (long long)E;
    for (int eh0 = 0; eh0 < EH; eh0 += 32) {
      const int eh = eh0 + lane;
      const bool valid = eh < EH;
      const float* rowp = mat + (long long)(valid ? eh : 0) * EL;
      float d[ELB];
#pragma unroll
      for (int el = 0; el < ELB; ++el) d[el] = (valid && el < EL) ? __ldg(rowp + el) : 0.f;
      for (int sl = 1; sl < nslices; ++sl) {
        const float* r2 = rowp + (long long)sl * slice_stride;
        float t[ELB];
#pragma unroll
        for (int el = 0; el < ELB; ++el) t[el] = (valid && el < EL) ? __ldg(r2 + el) : 0.f;
#pragma unroll
        for (int el = 0; el < ELB; ++el) d[el] += t[el];
      }
      const float th = valid ? tH[eh] : 0.f;
      float whi = 0.f;
#pragma unroll
      for (int el = 0; el < ELB; ++el) {
        whi = fmaf(d[el], tlr[el], whi);
        accl[el] = fmaf(d[el], th, accl[el]);
      }
      if (valid) wH[eh] = whi;
    }
#pragma unroll
    for (int el = 0; el < ELB; ++el) {
      if (el < EL) {   // warp-uniform
        float v = accl[el];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) wL[el] = v;
      }
    }
    __syncwarp();
    loo_group_stage(wH, wL, xs, qd, Q, cnth, cntl, EH, EL, lane, dxp + p * g.n * Q);
    __syncwarp();
  }
}
inline size_t loo1_smem(const EpsGeom& g) {
  return ((size_t)LOO1_WARPS * (size_t)(2 * g.AH + 2 * g.AL + g.m * g.Q) + (size_t)loo_nqd(g.AH, g.AL, g.m * g.Q, 0)) * sizeof(float);
}
inline bool loo1_rows_ok(const EpsGeom& g) {
  if (const char* e = getenv("DCTN_B200_LOO1")) return e[0] == '1' && g.AL <= 32 && loo1_smem(g) <= 48 * 1024;
  // long rows only: the per-patch fixed cost of a warp (origin, tables, the per-group stage on 15-40 lanes) is ~2000
  // instructions — K = 3, Q = 3 (A = 243): 946 us against 676 us of the staged kernel; CIFAR (2, 23 -> 24) (A = 529, nine
  // slices): 341 against 312 us of sum_slices + staged; CIFAR (2, 12 -> 24) (A = 1728): 222 against 522 us
  return g.A >= 1024 && g.AL >= 1 && g.AL <= 32 && loo1_smem(g) <= 48 * 1024;
}
template <int ELB>
int launch_loo1_rows_inst(const EpsGeom& g, const float* x, const float* dkr, long long slice_stride, int nslices, long long p0, int np,
                          float* dxp, cudaStream_t st) {
  int blocks = (np + LOO1_WARPS - 1) / LOO1_WARPS;
  if (blocks > 148 * 8) blocks = 148 * 8;
  loo1_rows_kernel<ELB><<<blocks, 32 * LOO1_WARPS, loo1_smem(g), st>>>(g, x, dkr, slice_stride, nslices, p0, np, dxp);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}
inline int launch_loo1_rows(const EpsGeom& g, const float* x, const float* dkr, long long slice_stride, int nslices, long long p0, int np,
                            float* dxp, cudaStream_t st) {
  const int EL = g.AL;
  if (EL <= 4) return launch_loo1_rows_inst<4>(g, x, dkr, slice_stride, nslices, p0, np, dxp, st);
  if (EL <= 8) return launch_loo1_rows_inst<8>(g, x, dkr, slice_stride, nslices, p0, np, dxp, st);
  if (EL <= 12) return launch_loo1_rows_inst<12>(g, x, dkr, slice_stride, nslices, p0, np, dxp, st);
  if (EL <= 16) return launch_loo1_rows_inst<16>(g, x, dkr, slice_stride, nslices, p0, np, dxp, st);
  if (EL <= 24) return launch_loo1_rows_inst<24>(g, x, dkr, slice_stride, nslices, p0, np, dxp, st);
  return launch_loo1_rows_inst<32>(g, x, dkr, slice_stride, nslices, p0, np, dxp, st);
}

inline size_t loo2_smem(const EpsGeom& g) {
  return ((size_t)LOO2_WARPS * (size_t)(g.BH * (g.BL | 1) + 2 * (g.BH + g.BL) + (g.n - g.m) * g.Q + g.O) +
          (size_t)loo_nqd(g.BH, g.BL, (g.n - g.m) * g.Q, g.Bn)) * sizeof(float);
}
}  // namespace

// kind 2: recompute T (two GEMMs); kind 3: T saved by the training forward (one GEMM + one streaming pass over T)
static int backward_input_impl(const EpsGeom& g, const float* x, const float* core, const float* gout, const float* tsaved,
                               float* dx, void* ws, int passes, cudaStream_t st) {
  const int BN1 = pick_bn(g, MODE_STORE), BN2 = tsaved ? 0 : pick_bn(g, MODE_DKR2);
  if (!BN1 || (!tsaved && !BN2)) return dctn_set_error(-2, "tcgen05 input-gradient kernels do not support this shape");
  const long long pc = dx_patch_chunk(g);
  uint32_t* absmax = (uint32_t*)ws;
  float* packed1 = (float*)((char*)ws + WS_HEADER);
  const bool fast1 = use_fast(g, 0, passes);
  size_t pf1 = packed_floats(g, MODE_STORE, BN1);
  if (tcfast_packed_floats(g, 0) > pf1) pf1 = tcfast_packed_floats(g, 0);
  if (tcfast_packed_floats(g, 2) > pf1) pf1 = tcfast_packed_floats(g, 2);
  if (tcfast_packed_floats(g, 3) > pf1) pf1 = tcfast_packed_floats(g, 3);
  float* packed2 = packed1 + ((pf1 + 63) & ~(size_t)63);
  float* dkr1 = packed2 + (tsaved ? 0 : ((packed_floats(g, MODE_DKR2, BN2) + 63) & ~(size_t)63));
  const int nslice1 = fast1 ? 1 : store_segments(g, passes);     // generic MODE_STORE GEMM: one dKR1 slice per K segment
  const int nslice_ws = store_segments(g, 3) > store_segments(g, ARITH_F16X3) ? store_segments(g, 3) : store_segments(g, ARITH_F16X3);
  float* dkr2 = dkr1 + (size_t)pc * g.A * nslice_ws;
  float* dxp = dkr2 + (size_t)pc * g.Bn;
  int rc;
  if ((rc = run_absmax(g, core, absmax, passes, st))) return rc;
  int c1h = 0, E1H = 0, c1l = 0, E1L = 0;
  const int ldw1 = fast1 ? tcfast_loo_groups(g, &c1h, &E1H, &c1l, &E1L) : 0;
  const bool fused1x = fast1 && tcfast_supported(g, 3);          // both leave-one-out stages inside the GEMM kernel
  const bool fused1 = !fused1x && ldw1 > 0 && tcfast_supported(g, 2);   // first stage only: W through memory
  if (fast1) rc = tcfast_pack(g, fused1x ? 3 : fused1 ? 2 : 0, core, packed1, absmax, st);
  else rc = run_pack(g, MODE_STORE, BN1, core, packed1, passes, st, absmax);
  if (rc) return rc;
  if (!tsaved && (rc = run_pack(g, MODE_DKR2, BN2, core, packed2, passes, st, absmax))) return rc;
  // leave-one-out stage 1 fused into the producers of dKR (register-table GEMM epilogue / the pass over the saved T):
  // only W (hi-group + lo-group sums per patch) goes through memory instead of the P x A and P x Bn matrices
  const bool fused2 = tsaved != nullptr && loo2_smem(g) <= 96 * 1024;
  const bool vec2 = (g.Bn & 3) == 0 && (g.BL & 3) == 0;
  if (fused2) {
    if (vec2) DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(loo2_from_saved_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)loo2_smem(g)));
    else DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(loo2_from_saved_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)loo2_smem(g)));
  }
  for (long long p0 = 0; p0 < g.P; p0 += pc) {
    const int np = (int)((g.P - p0 < pc) ? (g.P - p0) : pc);
    if (fused1x) {
      if ((rc = tcfast_gemm(g, 3, x, gout, packed1, absmax, p0, np, dxp, 0, nullptr, st))) return rc;
    } else if (fused1) {
      if ((rc = tcfast_gemm(g, 2, x, gout, packed1, absmax, p0, np, dkr1, ldw1, nullptr, st))) return rc;
      if ((rc = launch_loo_groups(g, x, dkr1, ldw1, p0, np, 0, c1h, E1H, c1l, E1L, dxp, st))) return rc;
    } else {
      if (fast1) rc = tcfast_gemm(g, 0, x, gout, packed1, absmax, p0, np, dkr1, g.A, nullptr, st);
      else rc = run_gemm(g, MODE_STORE, BN1, x, gout, packed1, p0, np, dkr1, g.A, passes, st, absmax, nullptr, (long long)pc * g.A);
      if (rc) return rc;
      if (!fast1 && loo1_rows_ok(g)) {   // slices summed and both leave-one-out stages in one pass over dKR1
        if ((rc = launch_loo1_rows(g, x, dkr1, (long long)pc * g.A, nslice1, p0, np, dxp, st))) return rc;
      } else {
        if (nslice1 > 1) {
          const long long count = (long long)np * g.A;
          int blocks = (int)((count + 255) / 256);
          if (blocks > 148 * 16) blocks = 148 * 16;
          sum_slices_kernel<<<blocks, 256, 0, st>>>(dkr1, (long long)pc * g.A, count, nslice1);
          dctn_count_launch();
          DCTN_CUDA_CHECK_RET(cudaGetLastError());
        }
        if ((rc = launch_loo<float>(g, x, dkr1, p0, np, 0, dxp, st))) return rc;
      }
    }
    if (fused2) {
      int blocks = (np + LOO2_WARPS - 1) / LOO2_WARPS;
      if (blocks > 148 * 8) blocks = 148 * 8;
      if (vec2) loo2_from_saved_kernel<true><<<blocks, 32 * LOO2_WARPS, loo2_smem(g), st>>>(g, x, tsaved, gout, dxp, p0, np);
      else loo2_from_saved_kernel<false><<<blocks, 32 * LOO2_WARPS, loo2_smem(g), st>>>(g, x, tsaved, gout, dxp, p0, np);
      dctn_count_launch();
      DCTN_CUDA_CHECK_RET(cudaGetLastError());
      continue;
    }
    if (tsaved) {
      const bool vec = (g.Bn & 3) == 0;
      const long long items = (long long)np * (vec ? g.Bn / 4 : g.Bn);
      int blocks = (int)((items + 255) / 256);
      if (blocks > 148 * 16) blocks = 148 * 16;
      if (vec) dkr2_from_saved_kernel<<<blocks, 256, 0, st>>>(tsaved, gout, dkr2, p0, np, g.Bn, g.O);
      else dkr2_from_saved_scalar_kernel<<<blocks, 256, 0, st>>>(tsaved, gout, dkr2, p0, np, g.Bn, g.O);
      dctn_count_launch();
      DCTN_CUDA_CHECK_RET(cudaGetLastError());
    } else if ((rc = run_gemm(g, MODE_DKR2, BN2, x, gout, packed2, p0, np, dkr2, g.Bn, passes, st, absmax))) return rc;
    if ((rc = launch_loo<float>(g, x, dkr2, p0, np, 1, dxp, st))) return rc;
  }
  return launch_gather_dx<float>(g, dxp, dx, st);
}

int tc_backward_input(const EpsGeom& g, const float* x, const float* core, const float* gout, float* dx, void* ws,
                      int passes, cudaStream_t st) {
  return backward_input_impl(g, x, core, gout, nullptr, dx, ws, passes, st);
}
int tc_backward_input_saved(const EpsGeom& g, const float* x, const float* core, const float* gout, const float* tsaved,
                            float* dx, void* ws, int passes, cudaStream_t st) {
  return backward_input_impl(g, x, core, gout, tsaved, dx, ws, passes, st);
}
