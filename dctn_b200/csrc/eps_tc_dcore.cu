// tcgen05 core-gradient kernel, split-fp16 arithmetic (the default for large cores; float32 in/out).
//
//   dcore[a][(bh, blo)] = sum_p  KR1[p][a] * TBH[p][bh] * TLO[p][blo]      blo = (second-half lo-group entry, o)
//
// eps_tc.cu (split TF32) generates BOTH GEMM operands on chip, one operand row per thread.  With fp16 MMAs a 64-patch
// chunk is only 768 tensor cycles and that generation became the limiter twice over: shared-memory bandwidth (every
// thread reads two 256-byte table rows per chunk) and, once that was fixed, the fp32 -> fp16 conversions — F2FP issues
// at a quarter of the FMUL rate (tools/ubench/split_rates.cu: 12 cycles per packed pair and sub-partition).  So this
// kernel generates HALF as many elements and loads the rest:
//   * the reduction is regrouped as  sum_p (KR1[p][a] * TBH[p][bh]) * TLO[p][blo]:  for one (a-tile, bh) the A operand
//     A'[a][p] = TH[p][ah] * (TL[p][al] * TBH[p][bh]) is generated into TENSOR MEMORY (TS-form MMA, one row per thread,
//     packed fp32x2 arithmetic), and the B operand is the table TLO itself.  The second factor is pre-multiplied by the
//     table kernel (one row per (bh, al)): a producer thread reads TWO table rows per chunk, not three — the producers
//     were bound by shared-memory wavefronts (ncu: 60 % LSU at 46 % tensor pipe; a 128-bit shared load is four
//     wavefronts whatever its lanes share), 768 wavefronts per chunk and SM against 576 tensor cycles; now 512;
//   * TLO does not depend on the tile, so build_tables16_kernel splits it ONCE per call into the fp16 hi / lo K-major
//     swizzled image the MMA reads, and one elected thread streams it chunk by chunk with cp.async.bulk (TMA engine,
//     mbarrier complete_tx) straight into the B stages: no B generation, no generic-proxy writes, no proxy fences;
//   * CTA tile: 128 (a) x NT (blo), NT <= 128 columns; grid = a-tiles x BH x blo-tiles x splits of the patch range.
// Number format and accumulation are those of eps_tc_gemm.cu / eps_tc.cu: v = hi + lo * 2^-11 (22 significant bits),
// hi*hi into a MAIN TMEM accumulator, the cross terms into a SMALL one.  The tensor core rounds its accumulator toward
// zero, so the main chain is kept short: TWO main accumulators alternate between segments of SEG16 chunks, and while
// one fills, the producers add the other into fp32 registers ("promotion") — the MMAs never wait for it.  The small
// accumulator (weight 2^-11) runs through.  fp16 range: every factor vector and the gout row
// of a patch are scaled to max-abs in [0.5, 1) (exact powers of two); the reduction runs over patches, so the patch
// scale cannot be undone afterwards: patch_exp_kernel finds E_max = max_p E_p and 2^(E_p - E_max) <= 1 is folded into
// TBH.  A' and TLO carry 2^15 each; reduce_partials_scaled_kernel applies 2^(E_max - 30).
#include <climits>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "eps_kernels.h"
#include "tc_common.cuh"

#ifdef DCTN_TCG_TIMING
#define TC16_CLK() clock64()
#else
#define TC16_CLK() 0ll
#endif

namespace {

constexpr int BM = 128;            // a rows = TMEM lanes
constexpr int CH = 64;             // patches per chunk = one 128-byte fp16 K slab
constexpr int MAX_STAGES = 4;      // A' slabs in tensor memory (64 columns each: hi | lo): as many as fit next to the three
                                   // accumulators in the 512 TMEM columns
constexpr int MAX_BSTAGES = 6;     // B slabs in shared memory: their own, deeper ring (a bulk copy of a 24-32 KB slab from L2 has a
                                   // latency of ~2000 cycles: with the 2 stages the TMEM budget leaves at NT = 128 the MMAs waited
                                   // 1150 cycles per chunk for B)
constexpr int TSTAGES = 4;         // table buffers
constexpr int SEG16 = 12;          // chunks per promotion segment: 12 * 4 k-steps = 48 roundings of the main chain (the promotion
                                   // runs on the producers beside the MMAs: measured bias at 96 steps 1.5e-6 / 8.7e-6 for random / positive data)
constexpr int NPROD_WARPS = 16;    // warps 1..16 generate A' (4 per TMEM lane quadrant, 16 patches of the chunk each); warp 0 issues MMAs;
                                   // warp 17 streams the tables, warp 18 the B slabs.  Sixteen, not eight: a producer's chunk is a chain
                                   // of barrier wait -> table loads -> conversion chain -> tcgen05.st -> arrive, and with two warps per
                                   // sub-partition that latency was exposed (cycle probes: 1100 cycles per chunk against 840 of MMA)
constexpr int PPQ = NPROD_WARPS / 4;   // producer warps per lane quadrant = parts of the chunk's patch range
constexpr int PPW = CH / PPQ;          // patches of a chunk per producer warp
constexpr int NTHREADS = 32 * (3 + NPROD_WARPS);
constexpr int TS_ = CH + 4;        // table row stride in floats (272 bytes)
constexpr size_t SMEM_LIMIT = 227 * 1024;

struct Dc16Args {
  EpsGeom g;
  const float* tables;   // [ceil(P/64)][AH + AL + BH][TS_]
  const float* bimg;     // [ceil(P/64)][ntile][hi|lo][NT rows x 128 bytes], swizzled fp16
  float* part;           // [splits][A][N]
  long long per_split;   // multiple of CH
  int NT, ntile;         // blo columns per tile, number of blo tiles
  int bstages;           // B slabs in shared memory
  long long* dbg;
  int dbg_skip_gen;      // timing builds: producers signal without generating (isolates the MMA rate)
  int dbg_order;         // timing builds: 1 = every MMA into ONE accumulator, 2 = three accumulators round-robin (results are wrong)
};

// ---- pass 1: exps[p] = E_p, *emax = max_p E_p (one thread per patch; *emax starts below every possible value)
__global__ void __launch_bounds__(256) patch_exp_kernel(EpsGeom g, const float* __restrict__ x, const float* __restrict__ gout,
                                                        int* __restrict__ exps, int* __restrict__ emax) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int E = INT_MIN;
  if (p < g.P) {
    const long long o0 = patch_origin(g, p);
    E = 0;
    bool zero = false;   // an all-zero factor vector or gout row: the patch contributes nothing
    for (int j = 0; j < g.n; ++j) {
      float m = 0.f;
      for (int q = 0; q < g.Q; ++q) m = fmaxf(m, fabsf(__ldg(&x[o0 + g.foff[j] + q])));
      E += tc::norm_exp(m);
      zero |= (m == 0.f);
    }
    float m = 0.f;
    for (int o = 0; o < g.O; ++o) m = fmaxf(m, fabsf(__ldg(&gout[p * g.O + o])));
    E += tc::norm_exp(m);
    zero |= (m == 0.f);
    // Such a patch must not set E_max: its norm_exp(0) = 0 terms would put it far ABOVE patches whose gradients are
    // merely small (upstream gradients of a deep stack reach 1e-20 and underflow to exact zeros for some patches), and
    // every real patch would be scaled by 2^(E_p - E_max) into fp16 underflow — an all-zero core gradient.
    if (zero) E = INT_MIN;
    exps[p] = E;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) E = max(E, __shfl_xor_sync(0xffffffffu, E, o));
  if ((threadIdx.x & 31) == 0 && E != INT_MIN) atomicMax(emax, E);
}

// ---- pass 2, one CTA per chunk of 64 patches:
//   tables[chunk][entry][i]: entries [0, AH): 2^15 * first-half hi group,
//                            [AH + bh*AL + al]: first-half lo group entry al * 2^(E_p - E_max) * second-half hi group entry bh
//                            (all from normalised x)
//   bimg[chunk][tile][part][row][64 fp16]: split-fp16 image of 2^15 * (second-half lo group x gout), rows = blo
__global__ void __launch_bounds__(256) build_tables16_kernel(EpsGeom g, const float* __restrict__ x, const float* __restrict__ gout,
                                                             float* __restrict__ tables, uint32_t* __restrict__ bimg, int NT, int ntile,
                                                             const int* __restrict__ exps, const int* __restrict__ emax) {
  extern __shared__ float bt_smem[];
  const int Q = g.Q, O = g.O, NX = g.n * Q;
  float* xs = bt_smem;             // [NX][64]
  float* gs = xs + NX * CH;        // [O][64]
  float* wp = gs + O * CH;         // [64]
  float* tlo = wp + CH;            // [BL][64]: second-half lo group, evaluated once and reused for every o
  float* tal = tlo + g.BL * CH;    // [AL][64]: first-half lo group
  float* tbh = tal + g.AL * CH;    // [BH][64]: second-half hi group times the patch weight
  int* qd = (int*)(tbh + g.BH * CH);   // qd[e] = e / Q for e below the largest group size: digit walks without divisions
  {
    int nqd = g.AH > g.AL ? g.AH : g.AL;
    if (g.BH > nqd) nqd = g.BH;
    if (g.BL > nqd) nqd = g.BL;
    for (int e = threadIdx.x; e < nqd; e += blockDim.x) qd[e] = e / Q;
  }
  const long long p0 = (long long)blockIdx.x * CH;
  const int lq = (Q & (Q - 1)) == 0 ? 31 - __clz(Q) : -1;      // log2(Q) when Q is a power of two: digits by shifts
  {
    // a thread owns one patch (i) and every fourth factor / gout row: one patch-origin computation, the Q loads of a
    // factor issued together, normalisation (exact power-of-two scale, see the header) in the same pass
    const int i = threadIdx.x & (CH - 1), slot = threadIdx.x >> 6;   // 256 threads = 4 x 64
    const long long p = p0 + i;
    const bool valid = p < g.P;
    const long long org = valid ? patch_origin(g, p) : 0;
    for (int j = slot; j <= g.n; j += 4) {
      const bool isg = j == g.n;
      const int cnt = isg ? O : Q;
      const float* src = isg ? gout + p * O : x + org + g.foff[isg ? 0 : j];
      float* dst = isg ? gs + i : xs + j * Q * CH + i;
      float m = 0.f;
      for (int q = 0; q < cnt; ++q) {
        const float v = valid ? __ldg(src + q) : 0.f;
        dst[q * CH] = v;
        m = fmaxf(m, fabsf(v));
      }
      const int e = tc::norm_exp(m);
      if (e != 0) {
        const float s1 = __int_as_float((127 - e / 2) << 23), s2 = __int_as_float((127 - (e - e / 2)) << 23);
        for (int q = 0; q < cnt; ++q) dst[q * CH] = dst[q * CH] * s1 * s2;
      }
    }
  }
  if (threadIdx.x < CH) {
    const long long p = p0 + threadIdx.x;
    const int ep = (p < g.P) ? exps[p] : INT_MIN;     // INT_MIN: padding, or a patch that contributes nothing
    wp[threadIdx.x] = (ep != INT_MIN) ? scalbnf(1.f, max(ep - __ldg(emax), -200)) : 0.f;
  }
  __syncthreads();
  auto group = [&](int j0, int cnt, int e, int i) -> float {
    float v = 1.f;
    for (int u = cnt - 1; u >= 0; --u) {
      int d;
      if (lq >= 0) { d = e & (Q - 1); e >>= lq; }
      else { const int e1 = qd[e]; d = e - e1 * Q; e = e1; }
      v *= xs[((j0 + u) * Q + d) * CH + i];
    }
    return v;
  };
  const int ENT = g.AH + g.BH * g.AL;
  float* out = tables + (long long)blockIdx.x * ENT * TS_;
  for (int idx = threadIdx.x; idx < g.AH * CH; idx += blockDim.x) {
    const int i = idx & (CH - 1), t = idx >> 6;
    out[t * TS_ + i] = 32768.f * group(0, g.a_nh, t, i);
  }
  for (int idx = threadIdx.x; idx < ENT * (TS_ - CH); idx += blockDim.x) out[(idx / (TS_ - CH)) * TS_ + CH + idx % (TS_ - CH)] = 0.f;
  for (int idx = threadIdx.x; idx < g.BL * CH; idx += blockDim.x) tlo[idx] = 32768.f * group(g.m + g.b_nh, g.b_nl, idx >> 6, idx & (CH - 1));
  for (int idx = threadIdx.x; idx < g.AL * CH; idx += blockDim.x) tal[idx] = group(g.a_nh, g.a_nl, idx >> 6, idx & (CH - 1));
  for (int idx = threadIdx.x; idx < g.BH * CH; idx += blockDim.x) tbh[idx] = wp[idx & (CH - 1)] * group(g.m, g.b_nh, idx >> 6, idx & (CH - 1));
  __syncthreads();
  // first-half lo group x second-half hi group: one row per (bh, al), four patches per item (128-bit stores)
  {
    float* lg = out + (long long)g.AH * TS_;
    const int rows = g.BH * g.AL;
    for (int idx = threadIdx.x; idx < rows * (CH / 4); idx += blockDim.x) {
      const int i4 = idx & (CH / 4 - 1), row = idx >> 4;
      const int bh = row / g.AL, al = row - bh * g.AL;
      const float4 l4 = *(const float4*)(tal + al * CH + 4 * i4), b4 = *(const float4*)(tbh + bh * CH + 4 * i4);
      *(float4*)(lg + (long long)row * TS_ + 4 * i4) = make_float4(l4.x * b4.x, l4.y * b4.y, l4.z * b4.z, l4.w * b4.w);
    }
  }
  // B image: one item = (row, packed pair of patches)
  const int BLO = g.BL * O;
  uint32_t* img = bimg + (long long)blockIdx.x * ntile * 2 * NT * 32;
  for (int idx = threadIdx.x; idx < ntile * NT * 32; idx += blockDim.x) {
    const int i2 = idx & 31, row = idx >> 5;
    float v0 = 0.f, v1 = 0.f;
    if (row < BLO) {
      const int e = row / O, o = row - e * O;
      const float2 t2 = *(const float2*)(tlo + e * CH + 2 * i2), g2 = *(const float2*)(gs + o * CH + 2 * i2);
      v0 = g2.x * t2.x;
      v1 = g2.y * t2.y;
    }
    uint32_t hi, lo;
    tc::split_f16x2(v0, v1, hi, lo);
    const int tile = row / NT, rl = row - tile * NT;
    const int w = rl * 32 + ((((i2 >> 2) ^ (rl & 7)) & 7) << 2) + (i2 & 3);   // 32-bit word inside the swizzled tile
    img[(tile * 2 + 0) * NT * 32 + w] = hi;
    img[(tile * 2 + 1) * NT * 32 + w] = lo;
  }
}

// out[i] = 2^(E_max - 30) * sum_z part[z*count + i]  (fixed order: deterministic)
__global__ void reduce_partials_scaled_kernel(const float* __restrict__ part, float* __restrict__ out, long long count, int splits,
                                              const int* __restrict__ emax) {
  int k = __ldg(emax);
  k = (k < -100000) ? -600 : k - 30;       // E_max untouched: no patch contributes, every partial sum is zero
  const float s1 = scalbnf(1.f, k / 2), s2 = scalbnf(1.f, k - k / 2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[(long long)z * count + i];
    out[i] = s * s1 * s2;
  }
}

// 32 lanes x 16 consecutive fp32 columns of tensor memory
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 8 consecutive fp32 columns of tensor memory
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// NTH = NT / 32: accumulator columns per thread = NT / PPQ (compile time: the promotion registers)
template <int NTH>
__global__ void __launch_bounds__(NTHREADS, 1) tc_dcore16_kernel(const __grid_constant__ Dc16Args a) {
  extern __shared__ unsigned char smem_dyn[];
  constexpr int NT = 32 * NTH;
  constexpr int NCOL = NT / PPQ;                       // accumulator columns promoted by one thread
  constexpr int STAGES = ((512 - 3 * NT) / 64 < MAX_STAGES) ? (512 - 3 * NT) / 64 : MAX_STAGES;   // 4, 4, 3, 2 for NT = 32..128
  constexpr uint32_t PART_BYTES = NT * 128;            // hi or lo part of one B slab
  constexpr uint32_t STAGE_BYTES = 2 * PART_BYTES;
  const EpsGeom& g = a.g;
  const int BLO = g.BL * g.O;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int a0 = blockIdx.x * BM;
  const int bh = blockIdx.y / a.ntile, tile = blockIdx.y - bh * a.ntile;
  int a1 = a0 + BM - 1; if (a1 > g.A - 1) a1 = g.A - 1;
  const int ah0 = a0 / g.AL, nah = a1 / g.AL - ah0 + 1;
  // rows of one table buffer: [nah first-half hi | AL rows of (first-half lo x this CTA's bh) | zero row]
  const int rZ = nah + g.AL, TE = rZ + 1;

  unsigned char* base = smem_dyn + ((1024u - (tc::smem_u32(smem_dyn) & 1023u)) & 1023u);
  const int BST = a.bstages;
  unsigned char* stages = base;                                  // [BST][hi|lo][NT rows x 128 B]
  float* tabs = (float*)(base + BST * STAGE_BYTES);              // [TSTAGES][TE][TS_]
  uint64_t* bars = (uint64_t*)(tabs + TSTAGES * TE * TS_);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * MAX_STAGES + 2 * MAX_BSTAGES + 2 * TSTAGES + 4);
  const uint32_t bar_fullA0 = tc::smem_u32(bars), bar_emptyA0 = bar_fullA0 + 8 * MAX_STAGES;       // A' slabs (TMEM)
  const uint32_t bar_fullB0 = bar_emptyA0 + 8 * MAX_STAGES, bar_emptyB0 = bar_fullB0 + 8 * MAX_BSTAGES;   // B slabs (smem)
  const uint32_t bar_tfull0 = bar_emptyB0 + 8 * MAX_BSTAGES, bar_tempty0 = bar_tfull0 + 8 * TSTAGES;
  const uint32_t bar_accfull0 = bar_tempty0 + 8 * TSTAGES, bar_accempty0 = bar_accfull0 + 16;   // one pair per main accumulator

  long long pbeg = (long long)blockIdx.z * a.per_split;
  long long pend = pbeg + a.per_split;
  if (pend > g.P) pend = g.P;
  const int nchunks = (int)((pend - pbeg + CH - 1) / CH);
  const long long chunk0 = pbeg / CH;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(bar_fullA0 + 8 * s, NPROD_WARPS);
      tc::mbar_init(bar_emptyA0 + 8 * s, 1);             // tcgen05.commit
    }
    for (int s = 0; s < MAX_BSTAGES; ++s) {
      tc::mbar_init(bar_fullB0 + 8 * s, 1);              // expect_tx arrive of the streaming thread (+ bytes)
      tc::mbar_init(bar_emptyB0 + 8 * s, 1);             // tcgen05.commit
    }
    for (int s = 0; s < TSTAGES; ++s) {
      tc::mbar_init(bar_tfull0 + 8 * s, 1);
      tc::mbar_init(bar_tempty0 + 8 * s, NPROD_WARPS);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(bar_accfull0 + 8 * b, 1);
      tc::mbar_init(bar_accempty0 + 8 * b, NPROD_WARPS);
    }
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tc::smem_u32(tmem_slot), 512);
  for (int i = tid; i < TSTAGES * TS_; i += NTHREADS) tabs[((i / TS_) * TE + rZ) * TS_ + i % TS_] = 0.f;   // zero rows
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // columns: main accumulator 0 | main accumulator 1 | small accumulator | A' stages (64 columns each: hi | lo)
  const uint32_t tmem_main0 = tmem_base, tmem_small = tmem_base + 2 * NT, tmem_a0 = tmem_base + 3 * NT;

  if (warp == 0) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = tc::make_idesc_f16(BM, NT);
    const uint64_t db_base = tc::make_sw128_kmajor_desc(tc::smem_u32(stages));
    int s = 0, sb = 0;          // A' slab (TMEM) and B slab (smem) ring positions
    uint32_t ph = 0, phb = 0;
    long long dm_acc = 0, dm_b = 0, dm_a = 0, dm_start = TC16_CLK();
    // The full barriers of chunk c+1 are probed (not waited for) in the middle of issuing chunk c's asynchronous MMAs, so
    // that at the chunk boundary the next MMAs go out back to back when their operands are there (see eps_tc_fast.cu).
    bool ready = false;
    for (int c = 0; c < nchunks; ++c) {
      const bool seg_first = (c % SEG16) == 0;
      const bool seg_last = ((c + 1) % SEG16) == 0 || c == nchunks - 1;
      const int seg = c / SEG16, mb = seg & 1;          // segment and the main accumulator it uses
      const uint32_t tmem_main = tmem_main0 + (uint32_t)(mb * NT);
      long long m0 = TC16_CLK();
      if (seg_first && seg >= 2) {   // segment seg-2 (same accumulator) must have been promoted before it is overwritten
        tc::mbar_wait(bar_accempty0 + 8 * mb, (uint32_t)(((seg - 2) >> 1) & 1));
        tc::tc_fence_after();
      }
      long long m1 = TC16_CLK();
      dm_acc += m1 - m0;
      if (!ready) {
        tc::mbar_wait(bar_fullB0 + 8 * sb, phb);
        long long m2 = TC16_CLK();
        tc::mbar_wait(bar_fullA0 + 8 * s, ph);
        long long m3 = TC16_CLK();
        dm_b += m2 - m1; dm_a += m3 - m2;
        tc::tc_fence_after();
      }
      const uint64_t db_hi = db_base + (uint64_t)((sb * STAGE_BYTES) >> 4);
      const uint64_t db_lo = db_hi + (PART_BYTES >> 4);
      const uint32_t a_hi = tmem_a0 + (uint32_t)(s * 64), a_lo = a_hi + 32;
      auto issue = [&](int k) {                         // 16 patches (32 bytes of K per row) per MMA
        const uint64_t adv = (uint64_t)(k * 2);
        const uint32_t acol = (uint32_t)(k * 8);
        const uint32_t first = (seg_first && k == 0) ? 0u : 1u;
#ifdef DCTN_TCG_TIMING
        if (a.dbg_order == 1) {
          tc::umma_f16_ts(tmem_main0, a_hi + acol, db_hi + adv, idesc, 1u);
          tc::umma_f16_ts(tmem_main0, a_hi + acol, db_lo + adv, idesc, 1u);
          tc::umma_f16_ts(tmem_main0, a_lo + acol, db_hi + adv, idesc, 1u);
          return;
        }
        if (a.dbg_order == 2) {
          tc::umma_f16_ts(tmem_main0, a_hi + acol, db_hi + adv, idesc, 1u);
          tc::umma_f16_ts(tmem_main0 + NT, a_hi + acol, db_lo + adv, idesc, 1u);
          tc::umma_f16_ts(tmem_small, a_lo + acol, db_hi + adv, idesc, 1u);
          return;
        }
#endif
        tc::umma_f16_ts(tmem_main, a_hi + acol, db_hi + adv, idesc, first);
        tc::umma_f16_ts(tmem_small, a_hi + acol, db_lo + adv, idesc, (c == 0 && k == 0) ? 0u : 1u);
        tc::umma_f16_ts(tmem_small, a_lo + acol, db_hi + adv, idesc, 1u);
      };
      if (tc::elect_one()) { issue(0); issue(1); }
      __syncwarp();
      int s_n = s + 1, sb_n = sb + 1;
      uint32_t ph_n = ph, phb_n = phb;
      if (s_n == STAGES) { s_n = 0; ph_n ^= 1; }
      if (sb_n == BST) { sb_n = 0; phb_n ^= 1; }
      ready = c + 1 < nchunks && tc::mbar_test(bar_fullB0 + 8 * sb_n, phb_n) && tc::mbar_test(bar_fullA0 + 8 * s_n, ph_n);   // probe only
      if (ready) tc::tc_fence_after();
      if (tc::elect_one()) {
        issue(2); issue(3);
        tc::umma_commit(bar_emptyA0 + 8 * s);
        tc::umma_commit(bar_emptyB0 + 8 * sb);
        if (seg_last) tc::umma_commit(bar_accfull0 + 8 * mb);
      }
      __syncwarp();
      s = s_n; ph = ph_n; sb = sb_n; phb = phb_n;
    }
    if (a.dbg && lane == 0) {
      long long* d = a.dbg + ((long long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16;
      d[0] = dm_acc; d[1] = dm_b; d[2] = dm_a; d[3] = TC16_CLK() - dm_start; d[4] = nchunks;
    }
  } else if (warp == 1 + NPROD_WARPS) {
    // =========================== table streamer (for the producers) ===========================
    if (tc::elect_one()) {
      const int ENT = g.AH + g.BH * g.AL;
      const uint32_t row_b = TS_ * 4;
      const uint32_t tbytes = (uint32_t)(nah + g.AL) * row_b;
      int ts = 0;
      uint32_t tph = 1;
      for (int c = 0; c < nchunks; ++c) {
        tc::mbar_wait(bar_tempty0 + 8 * ts, tph);
        const float* src = a.tables + (chunk0 + c) * (long long)ENT * TS_;
        const uint32_t dst = tc::smem_u32(tabs + ts * TE * TS_);
        const uint32_t tbar = bar_tfull0 + 8 * ts;
        tc::mbar_arrive_expect_tx(tbar, tbytes);
        tc::bulk_g2s(dst, src + (long long)ah0 * TS_, (uint32_t)nah * row_b, tbar);
        tc::bulk_g2s(dst + (uint32_t)nah * row_b, src + (long long)(g.AH + bh * g.AL) * TS_, (uint32_t)g.AL * row_b, tbar);
        if (++ts == TSTAGES) { ts = 0; tph ^= 1; }
      }
    }
  } else if (warp == 2 + NPROD_WARPS) {
    // =========================== B streamer (for the MMAs): the pre-split image, one slab per chunk ===========================
    if (tc::elect_one()) {
      int s = 0;
      uint32_t ph = 1;
      for (int c = 0; c < nchunks; ++c) {
        tc::mbar_wait(bar_emptyB0 + 8 * s, ph);
        const float* bsrc = a.bimg + ((chunk0 + c) * a.ntile + tile) * (long long)(2 * NT * 32);
        tc::mbar_arrive_expect_tx(bar_fullB0 + 8 * s, STAGE_BYTES);
        tc::bulk_g2s(tc::smem_u32(stages + s * STAGE_BYTES), bsrc, STAGE_BYTES, bar_fullB0 + 8 * s);
        if (++s == BST) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // =========================== producers: A' rows into tensor memory ===========================
    const int pw = warp - 1;                 // 0..7
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
    const int hf = pw >> 2;                  // which PPW-patch part of the chunk this thread generates
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    int offH = rZ * TS_, offL = rZ * TS_;    // float offsets inside a table buffer; padding rows use the zero row
    {
      const int ai = a0 + quad * 32 + lane;
      if (ai < g.A) { offH = (ai / g.AL - ah0) * TS_; offL = (nah + ai % g.AL) * TS_; }
    }
    // promotion: row quad*32 + lane, columns [hf*NCOL, +NCOL)
    float racc[NCOL];
#pragma unroll
    for (int i = 0; i < NCOL; ++i) racc[i] = 0.f;
    int next_drain = 0;
    const int nseg = (nchunks + SEG16 - 1) / SEG16;
    // promotion of segment seg: its main accumulator; the last segment also takes the small accumulator (which ran through)
    auto drain = [&](int seg) {
      const int mb = seg & 1;
      tc::mbar_wait(bar_accfull0 + 8 * mb, (uint32_t)((seg >> 1) & 1));
      tc::tc_fence_after();
#pragma unroll
      for (int cb = 0; cb < NCOL; cb += 8) {
        float v[8];
        tmem_ld8(tmem_main0 + lane_base + (uint32_t)(mb * NT + hf * NCOL + cb), v);
#pragma unroll
        for (int i = 0; i < 8; ++i) racc[cb + i] += v[i];
        if (seg == nseg - 1) {
          tmem_ld8(tmem_small + lane_base + (uint32_t)(hf * NCOL + cb), v);
#pragma unroll
          for (int i = 0; i < 8; ++i) racc[cb + i] = fmaf(v[i], 1.f / 2048.f, racc[cb + i]);
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(bar_accempty0 + 8 * mb);
    };

    int s = 0, ts = 0;
    uint32_t phe = 1, tph = 0;
    long long dp[3] = {0, 0, 0};
    for (int c = 0; c < nchunks; ++c) {
      long long q0 = TC16_CLK();
      tc::mbar_wait(bar_tfull0 + 8 * ts, tph);
      long long q1 = TC16_CLK();
      const float* tb = tabs + ts * TE * TS_;
      const float4* th = (const float4*)(tb + offH) + hf * (PPW / 4);
      const float4* tl = (const float4*)(tb + offL) + hf * (PPW / 4);
      uint32_t hi[PPW / 2], lo[PPW / 2];
#ifdef DCTN_TCG_TIMING
      if (a.dbg_skip_gen) {
#pragma unroll
        for (int i = 0; i < PPW / 2; ++i) hi[i] = lo[i] = 0x3c003c00u;
      } else
#endif
#pragma unroll
      for (int q4 = 0; q4 < PPW / 4; ++q4) {
        const float4 h4 = th[q4], l4 = tl[q4];
        const tc::f32x2_t v01 = tc::mul2(tc::pack2(h4.x, h4.y), tc::pack2(l4.x, l4.y));
        const tc::f32x2_t v23 = tc::mul2(tc::pack2(h4.z, h4.w), tc::pack2(l4.z, l4.w));
        tc::split_f16x2_p(v01, hi[2 * q4], lo[2 * q4]);
        tc::split_f16x2_p(v23, hi[2 * q4 + 1], lo[2 * q4 + 1]);
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(bar_tempty0 + 8 * ts);   // table buffer read
      if (++ts == TSTAGES) { ts = 0; tph ^= 1; }
      tc::mbar_wait(bar_emptyA0 + 8 * s, phe);
      tc::tc_fence_after();
      const uint32_t dst = tmem_a0 + lane_base + (uint32_t)(s * 64 + hf * (PPW / 2));
      tc::tmem_st8_u(dst, hi);
      tc::tmem_st8_u(dst + 32, lo);
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(bar_fullA0 + 8 * s);
      long long q2 = TC16_CLK();
      if (++s == STAGES) { s = 0; phe ^= 1; }
      // segment seg was handed to the tensor core SEG16/2 chunks ago: it has long completed, so this never waits
      if ((c % SEG16) == SEG16 / 2 && c > SEG16) drain(next_drain++);
      long long q3 = TC16_CLK();
      dp[0] += q1 - q0; dp[1] += q2 - q1; dp[2] += q3 - q2;
    }
    if (a.dbg && lane == 0 && warp == 1) {
      long long* d = a.dbg + ((long long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + 5;
      for (int i = 0; i < 3; ++i) d[i] = dp[i];
    }
    while (next_drain < nseg) drain(next_drain++);

    // =========================== epilogue: registers -> partial tile ===========================
    const int arow = a0 + quad * 32 + lane;
    if (arow < g.A) {
      float* prow = a.part + ((long long)blockIdx.z * g.A + arow) * (long long)g.N + (long long)bh * BLO;
      const int c0 = tile * NT + hf * NCOL;   // first blo of this thread
#pragma unroll
      for (int cb = 0; cb < NCOL; cb += 4) {
        const int nb = c0 + cb;
        if (nb + 4 <= BLO && (BLO & 3) == 0 && (g.N & 3) == 0) {
          *(float4*)(prow + nb) = make_float4(racc[cb], racc[cb + 1], racc[cb + 2], racc[cb + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (nb + i < BLO) prow[nb + i] = racc[cb + i];
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ host side
// blo-tile width: multiple of 32, <= 128, least padding, then widest
inline int pick_nt(const EpsGeom& g) {
  const int BLO = g.BL * g.O;
  if (const char* e = getenv("DCTN_B200_DCORE_NT")) {   // tuning override
    const int nt = atoi(e);
    if (nt == 32 || nt == 64 || nt == 96 || nt == 128) return nt;
  }
  int best = 0;
  long long best_cost = 0;
  for (int nt = 128; nt >= 32; nt -= 32) {
    const long long cost = (long long)((BLO + nt - 1) / nt) * nt;
    if (!best || cost < best_cost) { best = nt; best_cost = cost; }
  }
  return best;
}
// The generated operand A' has A * BH rows and is regenerated for every blo tile, so the work of the producers (the
// limiter of this kernel) is proportional to BH * ntile.  The split of the second half into hi / lo groups is free here
// (n = bh * BLO + blo for any split): take the one with the least generation, ties to the larger hi group.
inline EpsGeom regroup(const EpsGeom& g0) {
  EpsGeom best = g0;
  long long best_cost = -1;
  const int cnt = g0.n - g0.m;
  int force_nh = -1;
  if (const char* e = getenv("DCTN_B200_DCORE_BNH")) force_nh = atoi(e);   // tuning override: hi-group size of the second half
  for (int nh = cnt; nh >= 0; --nh) {
    if (force_nh >= 0 && force_nh <= cnt && nh != force_nh) continue;
    EpsGeom g = g0;
    g.b_nh = nh; g.b_nl = cnt - nh;
    const long long BH = ipow_host(g0.Q, nh), BL = ipow_host(g0.Q, cnt - nh);
    if (BL * g0.O > (1 << 20)) continue;
    g.BH = (int)BH; g.BL = (int)BL;
    const int NT = pick_nt(g);
    // NT = 128 leaves room for only two A' stages in tensor memory and makes the MMAs of a chunk as long as its
    // generation: measured 1.4x per chunk (config 2, layer 2: 16 x 1 tiles of 96 beat 4 x 3 tiles of 128)
    // Both the generation and (the A' operand is read from tensor memory at ~70 cycles per MMA, which hides N <= 128)
    // the MMA time are proportional to the number of (bh, tile) pairs: wide tiles win.  NT = 128 leaves only two A'
    // slabs in tensor memory: measured 7 % faster than 16 x 96 at 12 x 128 (config 2, layer 2), not 25 %.
    const long long ntile = (BL * g0.O + NT - 1) / NT;
    if ((g0.P + CH - 1) / CH * ntile * 2 * NT * 128 > (640ll << 20)) continue;      // pre-split B image (workspace) cap
    const long long cost = BH * ntile * (NT == 128 ? 12 : 10);
    if (best_cost < 0 || cost < best_cost) { best = g; best_cost = cost; }
  }
  return best;
}

inline size_t dcore16_fixed_smem(int TE) {
  return 1024 + (size_t)TSTAGES * TE * TS_ * 4 + (2 * MAX_STAGES + 2 * MAX_BSTAGES + 2 * TSTAGES + 4) * 8 + 16;
}
// B slabs that fit beside the table buffers (at least as many as there are A' slabs in tensor memory)
inline int dcore16_bstages(int TE, int NT) {
  const size_t fixed = dcore16_fixed_smem(TE);
  int nb = fixed < SMEM_LIMIT ? (int)((SMEM_LIMIT - fixed) / ((size_t)2 * NT * 128)) : 0;
  if (nb > MAX_BSTAGES) nb = MAX_BSTAGES;
  return nb;
}
inline int dcore16_te(const EpsGeom& g) {
  int nah = (BM + g.AL - 1) / g.AL + 1; if (nah > g.AH) nah = g.AH;
  return nah + g.AL + 1;
}
inline size_t dcore16_smem(const EpsGeom& g, int NT) {
  int nah = (BM + g.AL - 1) / g.AL + 1; if (nah > g.AH) nah = g.AH;
  const int TE = nah + g.AL + 1;
  return dcore16_fixed_smem(TE) + (size_t)dcore16_bstages(TE, NT) * 2 * NT * 128;
}
// split the patch range so that the grid fills whole waves of one CTA per SM
inline void dcore16_split(const EpsGeom& g, int ntile, long long* per_split, int* splits) {
  const long long tiles = (long long)((g.A + BM - 1) / BM) * g.BH * ntile;
  const long long nchunk = (g.P + CH - 1) / CH;
  long long max_s = nchunk / (2 * SEG16);       // at least two promotion segments per CTA
  if (max_s < 1) max_s = 1;
  if (max_s > 32) max_s = 32;
  int best = 1;
  double best_eff = -1.0;
  for (int s = 1; s <= max_s; ++s) {
    const long long ctas = tiles * s;
    const double eff = (double)ctas / (double)(((ctas + 147) / 148) * 148);
    if (eff > best_eff + 0.02) { best = s; best_eff = eff; }   // smaller split counts win near-ties (fewer partial tiles)
  }
  long long per = (g.P + best - 1) / best;
  per = ((per + CH - 1) / CH) * CH;
  *per_split = per;
  *splits = (int)((g.P + per - 1) / per);
}
inline size_t table_floats(const EpsGeom& g) {
  return (size_t)((g.P + CH - 1) / CH) * (size_t)(g.AH + (size_t)g.BH * g.AL) * TS_;
}
inline size_t table_kernel_smem(const EpsGeom& g) {
  int nqd = g.AH > g.AL ? g.AH : g.AL;
  if (g.BH > nqd) nqd = g.BH;
  if (g.BL > nqd) nqd = g.BL;
  return (size_t)((g.n * g.Q + g.O + g.BL + g.AL + g.BH) * CH + CH + nqd) * sizeof(float);
}
inline size_t bimg_words(const EpsGeom& g, int NT, int ntile) { return (size_t)((g.P + CH - 1) / CH) * (size_t)ntile * 2 * NT * 32; }

}  // namespace

bool tc16_dcore_supported(const EpsGeom& g0) {
  const EpsGeom g = regroup(g0);
  if (g.P >= (1ll << 31) / (g.Q > g.O ? g.Q : g.O)) return false;   // 32-bit patch index math
  if (g.A < 64 || g.N < 64) return false;       // tiles would be mostly padding: the CUDA-core family is the better fit
  if (g.P < 4096) return false;                 // tiny reductions are launch-bound either way
  if (table_kernel_smem(g) > 200 * 1024) return false;   // table kernel staging
  return dcore16_smem(g, pick_nt(g)) <= SMEM_LIMIT;
}

size_t tc16_dcore_workspace_bytes(const EpsGeom& g0) {
  const EpsGeom g = regroup(g0);
  const int NT = pick_nt(g), ntile = (g.BL * g.O + NT - 1) / NT;
  long long per;
  int splits;
  dcore16_split(g, ntile, &per, &splits);
  return ((size_t)splits * g.A * g.N + 64 + table_floats(g) + 64 + bimg_words(g, NT, ntile) + 64 + (size_t)g.P + 64 + 64) * 4 + 256;
}

int tc16_backward_core(const EpsGeom& g0, const float* x, const float* gout, float* dcore, void* ws, cudaStream_t st) {
  const EpsGeom g = regroup(g0);
  const int NT = pick_nt(g), ntile = (g.BL * g.O + NT - 1) / NT;
  const size_t smem = dcore16_smem(g, NT);
  if (smem > SMEM_LIMIT) return dctn_set_error(-2, "tcgen05 core-gradient kernel needs %zu bytes of shared memory", smem);
  Dc16Args a{};
  a.g = g; a.part = (float*)ws; a.NT = NT; a.ntile = ntile;
  a.bstages = dcore16_bstages(dcore16_te(g), NT);
  if (a.bstages < 2) return dctn_set_error(-2, "tcgen05 core-gradient kernel: the tables of this shape leave no room for the B slabs");
  int splits;
  dcore16_split(g, ntile, &a.per_split, &splits);
  float* tables = a.part + (((size_t)splits * g.A * g.N + 63) & ~(size_t)63);
  uint32_t* bimg = (uint32_t*)(tables + ((table_floats(g) + 63) & ~(size_t)63));
  int* exps = (int*)(bimg + ((bimg_words(g, NT, ntile) + 63) & ~(size_t)63));
  int* emax = exps + ((g.P + 63) & ~63ll);
  a.tables = tables; a.bimg = (const float*)bimg;
  const size_t bsm = table_kernel_smem(g);
  if (bsm > 200 * 1024) return dctn_set_error(-2, "table kernel needs %zu bytes of shared memory", bsm);
  DCTN_CUDA_CHECK_RET(cudaMemsetAsync(emax, 0x80, sizeof(int), st));   // 0x80808080: below every possible E_p
  patch_exp_kernel<<<(unsigned)((g.P + 255) / 256), 256, 0, st>>>(g, x, gout, exps, emax);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(build_tables16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
  build_tables16_kernel<<<(unsigned)((g.P + CH - 1) / CH), 256, bsm, st>>>(g, x, gout, tables, bimg, NT, ntile, exps, emax);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  auto kern = NT == 32 ? tc_dcore16_kernel<1> : NT == 64 ? tc_dcore16_kernel<2> : NT == 96 ? tc_dcore16_kernel<3> : tc_dcore16_kernel<4>;
  DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((g.A + BM - 1) / BM, g.BH * ntile, splits);
  a.dbg = nullptr;
#ifdef DCTN_TCG_TIMING   // cycle probes: timing builds only (allocates, synchronises, not thread-safe)
  static long long* dbg_buf = nullptr;
  const long long ncta = (long long)grid.x * grid.y * grid.z;
  if (getenv("DCTN_TCG_DEBUG") && ncta <= 4096) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 4096 * 16 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 4096 * 16 * sizeof(long long), st);
    a.dbg = dbg_buf;
  }
  a.dbg_skip_gen = getenv("DCTN_B200_SKIP_GEN") != nullptr;
  a.dbg_order = getenv("DCTN_B200_MMA_ORDER") ? atoi(getenv("DCTN_B200_MMA_ORDER")) : 0;
#endif
  kern<<<grid, NTHREADS, smem, st>>>(a);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
#ifdef DCTN_TCG_TIMING
  if (a.dbg) {
    static long long host[4096 * 16];
    cudaStreamSynchronize(st);
    cudaMemcpy(host, dbg_buf, (size_t)ncta * 16 * sizeof(long long), cudaMemcpyDeviceToHost);
    double sum[16] = {0};
    for (int c = 0; c < ncta; ++c) for (int k = 0; k < 16; ++k) sum[k] += (double)host[c * 16 + k];
    const double nch = sum[4];
    fprintf(stderr, "[dcore16 dbg] NT=%d ctas=%lld splits=%d chunks/cta=%.0f per-chunk cycles: mma wait acc %.0f B %.0f A %.0f total %.0f | producer: wait-tables %.0f rows %.0f drain %.0f\n",
            NT, ncta, splits, nch / ncta, sum[0] / nch, sum[1] / nch, sum[2] / nch, sum[3] / nch, sum[5] / nch, sum[6] / nch, sum[7] / nch);
  }
#endif
  const long long count = (long long)g.A * g.N;
  int blocks = (int)((count + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  reduce_partials_scaled_kernel<<<blocks, 256, 0, st>>>(a.part, dcore, count, splits, emax);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}
