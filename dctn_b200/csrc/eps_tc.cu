// tcgen05 core-gradient kernel of the EPS contraction, split-TF32 arithmetic (float32; 3xTF32 or 1xTF32; variants tc3 / tc1).
// The default arithmetic, split fp16, has its own kernel in eps_tc_dcore.cu; this one is the validated reference
// arithmetic it is compared against.
//
//   dcore[a][n] = sum_p KR1[p][a] * KR2[p][b(n)] * gout[p][o(n)]          reduction over ALL patches
//
//   * CTA tile 128 (a) x 128 (n); the K loop runs over this CTA's patch range in chunks of 64 patches (two
//     32-wide slabs per pipeline stage, so every barrier round feeds 24 MMAs);
//   * BOTH operands are generated on chip from the two-level Khatri-Rao tables (common.cuh).  The tables themselves
//     (AH + AL + BH + BL*O floats per patch, e.g. 192) are computed ONCE per call by build_tables_kernel into an
//     L2-friendly scratch buffer laid out [chunk of 64 patches][entry][64 (+4 pad)]: every one of the (A/128)*(N/128)
//     output tiles needs them for the same patches, and rebuilding them per CTA was 60% of the producers' time
//     (profiles/r01_dcore_phase_cycles.txt).  One elected thread streams the rows a tile needs (4 contiguous ranges)
//     per chunk with cp.async.bulk (TMA engine, mbarrier complete_tx), double-buffered.  Every producer thread then
//     generates ONE operand row: its two table-row pointers are loop-invariant, so a row is 16 x LDS.128 + 32
//     multiplies + the TF32 hi/lo split per slab — no index arithmetic in the loop.
//     A rows (a) are written with tcgen05.st into TENSOR MEMORY (the MMA reads A from TMEM, TS form); B rows (n) go
//     to shared memory in the K-major SWIZZLE_128B layout.  Nothing Q^m-sized ever comes from HBM;
//   * warp 0 issues tcgen05.mma kind::tf32 (M=128, N=128, K=8): hi*hi products into a MAIN accumulator, the two
//     cross terms into a SMALL one.  The tensor core rounds its fp32 accumulator toward zero on every MMA
//     (measured: 768 accumulate steps shrink a result by 1.5e-5), so the main chain is also kept short: every
//     SEG_CHUNKS chunks the producers drain both accumulators with tcgen05.ld and add them into fp32 REGISTERS
//     (64 per thread, round-to-nearest) — "promotion"; residual bias ~2e-6 independent of the reduction length;
//   * split-K over patch ranges (about two waves of CTAs); partial tiles go to the workspace and a second kernel
//     sums them in a fixed order (deterministic).
// TMEM columns: main 0..127, small 128..255, A stage s at 256 + 128*s (slab0 hi | slab0 lo | slab1 hi | slab1 lo).
#include "common.cuh"
#include "eps_kernels.h"
#include <cstdio>
#include <cstdlib>

#include "tc_common.cuh"

// cycle probes for tuning (make NVFLAGS+=-DDCTN_TCG_TIMING, run with DCTN_TCG_DEBUG=1); compiled out by default
#ifdef DCTN_TCG_TIMING
#define TCD_CLK() clock64()
#else
#define TCD_CLK() 0ll
#endif

namespace {

constexpr int BM = 128;           // a rows  (= TMEM lanes)
constexpr int BN = 128;           // n rows of the B operand
constexpr int CH = 64;            // patches per pipeline stage
constexpr int SLABS = CH / 32;    // 32-wide K slabs per stage
constexpr int STAGES = 2;
constexpr int NPROD_WARPS = 8;    // warps 1..8; warp 0 issues MMAs; warp 9 streams the tables
constexpr int NTHREADS_TC = 32 * (2 + NPROD_WARPS);
constexpr int TSTAGES = 2;        // table buffers
constexpr int SEG_CHUNKS = 12;    // chunks per promotion segment: 12 * 2 slabs * 4 k-steps = 96 roundings of the main chain
constexpr int TS_ = CH + 4;       // table row stride in floats (272 B: 16-byte aligned, rows 4 banks apart)
constexpr uint32_t SLAB_BYTES = BN * 32 * 4;           // one part (hi or lo) of one slab of B: 16 KB
constexpr uint32_t STAGE_BYTES = SLABS * 2 * SLAB_BYTES;  // 64 KB
constexpr size_t TC_SMEM_LIMIT = 227 * 1024;

struct TcDcoreArgs {
  EpsGeom g;
  const float* tables;  // [ceil(P/64)][ENT][TS_] from build_tables_kernel, ENT = AH + AL + BH + BL*O
  float* part;          // [splits][A][N]
  long long per_split;  // multiple of CH
  int passes;           // 3 (fp32-accurate) or 1
  int three;            // 1: B rows are the 3-level product BH * BL * gout (separate BL and gout tables)
  long long* dbg;       // optional per-CTA cycle counters (16 per CTA)
};

__device__ __forceinline__ void producer_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * NPROD_WARPS) : "memory"); }

// which table entries a tile needs: all of the lo tables, a contiguous range of the hi tables
struct TileTables {
  int ah0, nah;   // first needed AH entry and count
  int bh0, nbh;   // first needed BH entry and count
  int te;         // total entries: nah + AL + nbh + BL*O
};
// number of entries of the last table section: BL*O (lo group x gout, two-level) or BL + O (three-level)
__host__ __device__ inline int last_section(const EpsGeom& g, int three) { return three ? g.BL + g.O : g.BL * g.O; }
__host__ __device__ inline TileTables tile_tables(const EpsGeom& g, int a0, int n0, int three) {
  TileTables t;
  const int BLO = g.BL * g.O;
  int a1 = a0 + BM - 1; if (a1 > g.A - 1) a1 = g.A - 1;
  int n1 = n0 + BN - 1; if (n1 > g.N - 1) n1 = g.N - 1;
  t.ah0 = a0 / g.AL; t.nah = a1 / g.AL - t.ah0 + 1;
  t.bh0 = n0 / BLO;  t.nbh = n1 / BLO - t.bh0 + 1;
  t.te = t.nah + g.AL + t.nbh + last_section(g, three);
  return t;
}
// upper bound of TileTables::te over all tiles
inline int max_tile_entries(const EpsGeom& g, int three) {
  const int BLO = g.BL * g.O;
  int nah = (BM + g.AL - 1) / g.AL + 1; if (nah > g.AH) nah = g.AH;
  int nbh = (BN + BLO - 1) / BLO + 1;   if (nbh > g.BH) nbh = g.BH;
  return nah + g.AL + nbh + last_section(g, three);
}

// tables[chunk][entry][i] for patch p = chunk*64 + i (zeros past P): entries [0,AH): first-half hi group,
// [AH, AH+AL): first-half lo group, then [.., +BH): second-half hi group, then either BL*O entries (second-half lo group
// x gout) or, three-level, BL entries (lo group) followed by O entries (gout).
__global__ void __launch_bounds__(256) build_tables_kernel(EpsGeom g, const float* __restrict__ x,
                                                           const float* __restrict__ gout, float* __restrict__ tables, int three) {
  extern __shared__ float bt_smem[];
  const int Q = g.Q, O = g.O, NX = g.n * Q;
  float* xs = bt_smem;             // [NX][64]
  float* gs = xs + NX * CH;        // [O][64]
  const long long p0 = (long long)blockIdx.x * CH;
  for (int idx = threadIdx.x; idx < (NX + O) * CH; idx += blockDim.x) {
    const int i = idx & (CH - 1), r = idx >> 6;
    const long long p = p0 + i;
    float v = 0.f;
    if (p < g.P) v = (r < NX) ? __ldg(&x[patch_origin(g, p) + g.foff[r / Q] + r % Q]) : __ldg(&gout[p * O + (r - NX)]);
    xs[idx] = v;
  }
  __syncthreads();
  const int ENT = g.AH + g.AL + g.BH + last_section(g, three);
  float* out = tables + (long long)blockIdx.x * ENT * TS_;
  for (int idx = threadIdx.x; idx < ENT * CH; idx += blockDim.x) {
    const int i = idx & (CH - 1), t = idx >> 6;
    int e, j0, cnt;
    float v = 1.f;
    if (t < g.AH) { e = t; j0 = 0; cnt = g.a_nh; }
    else if (t < g.AH + g.AL) { e = t - g.AH; j0 = g.a_nh; cnt = g.a_nl; }
    else if (t < g.AH + g.AL + g.BH) { e = t - g.AH - g.AL; j0 = g.m; cnt = g.b_nh; }
    else {
      const int k = t - (g.AH + g.AL + g.BH);
      j0 = g.m + g.b_nh; cnt = g.b_nl;
      if (!three) { e = k / O; v = gs[(k - e * O) * CH + i]; }     // lo group x gout
      else if (k < g.BL) { e = k; }                                 // lo group alone
      else { e = 0; cnt = 0; v = gs[(k - g.BL) * CH + i]; }         // gout row
    }
    for (int u = cnt - 1; u >= 0; --u) {
      const int d = e % Q;
      e /= Q;
      v *= xs[((j0 + u) * Q + d) * CH + i];
    }
    out[t * TS_ + i] = v;
  }
  // the 4 pad floats of every row are never read by the consumers (rows are read 64 wide) but are copied: define them
  for (int idx = threadIdx.x; idx < ENT * (TS_ - CH); idx += blockDim.x) out[(idx / (TS_ - CH)) * TS_ + CH + idx % (TS_ - CH)] = 0.f;
}

template <bool THREE>
__global__ void __launch_bounds__(NTHREADS_TC, 1) tc_dcore_kernel(const __grid_constant__ TcDcoreArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  const EpsGeom& g = a.g;
  const int O = g.O;
  const int BLO = g.BL * O;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int a0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const TileTables tt = tile_tables(g, a0, n0, THREE ? 1 : 0);
  const int TE = tt.te;
  const int LAST = last_section(g, THREE ? 1 : 0);

  // ---- carve shared memory
  unsigned char* base = smem_dyn + ((1024u - (tc::smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* stages = base;                                  // B operand: [STAGES][slab][hi|lo][128 rows x 128 B]
  float* tabs = (float*)(base + STAGES * STAGE_BYTES);           // [TSTAGES][TE][TS_]  table rows this tile needs
  float* zrow = tabs + TSTAGES * TE * TS_;                       // [TS_] zeros: padding operand rows multiply this
  float* onerow = zrow + TS_;                                    // [TS_] ones: third factor of rows that have none
  uint64_t* bars = (uint64_t*)(onerow + TS_);
  uint32_t* tmem_slot = (uint32_t*)(bars + 3 * STAGES + 2 * TSTAGES + 2);
  const uint32_t bar_fullA0 = tc::smem_u32(bars), bar_fullB0 = bar_fullA0 + 8 * STAGES;
  const uint32_t bar_empty0 = bar_fullB0 + 8 * STAGES;           // one per stage: frees both the TMEM A stage and the smem B stage
  const uint32_t bar_tfull0 = bar_empty0 + 8 * STAGES, bar_tempty0 = bar_tfull0 + 8 * TSTAGES;
  const uint32_t bar_accfull = bar_tempty0 + 8 * TSTAGES, bar_accempty = bar_accfull + 8;

  long long pbeg = (long long)blockIdx.z * a.per_split;
  long long pend = pbeg + a.per_split;
  if (pend > g.P) pend = g.P;
  const int nchunks = (int)((pend - pbeg + CH - 1) / CH);

  // ---- one-time setup
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(bar_fullA0 + 8 * s, 4);   // the 4 warps that write A rows
      tc::mbar_init(bar_fullB0 + 8 * s, 4);   // the 4 warps that write B rows
      tc::mbar_init(bar_empty0 + 8 * s, 1);   // tcgen05.commit
    }
    for (int s = 0; s < TSTAGES; ++s) {
      tc::mbar_init(bar_tfull0 + 8 * s, 1);             // expect_tx arrive of the table warp (+ bytes)
      tc::mbar_init(bar_tempty0 + 8 * s, NPROD_WARPS);  // every producer warp is done reading the buffer
    }
    tc::mbar_init(bar_accfull, 1);
    tc::mbar_init(bar_accempty, NPROD_WARPS);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tc::smem_u32(tmem_slot), 512);
  for (int i = tid; i < TS_; i += NTHREADS_TC) { zrow[i] = 0.f; onerow[i] = 1.f; }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_main = tmem_base, tmem_small = tmem_base + BN, tmem_a0 = tmem_base + 2 * BN;

  if (warp == 0) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = tc::make_idesc_tf32(BM, BN);
    const uint64_t db_base = tc::make_sw128_kmajor_desc(tc::smem_u32(stages));
    int s = 0;
    uint32_t ph = 0;
    long long dm_acc = 0, dm_b = 0, dm_a = 0, dm_start = TCD_CLK();
    for (int c = 0; c < nchunks; ++c) {
      const bool seg_first = (c % SEG_CHUNKS) == 0;
      const bool seg_last = ((c + 1) % SEG_CHUNKS) == 0 || c == nchunks - 1;
      long long m0 = TCD_CLK();
      if (seg_first && c > 0) {   // the previous segment must have been promoted before its accumulators are overwritten
        tc::mbar_wait(bar_accempty, (uint32_t)((c / SEG_CHUNKS - 1) & 1));
        tc::tc_fence_after();
      }
      long long m1 = TCD_CLK();
      tc::mbar_wait(bar_fullB0 + 8 * s, ph);
      long long m2 = TCD_CLK();
      tc::mbar_wait(bar_fullA0 + 8 * s, ph);
      long long m3 = TCD_CLK();
      dm_acc += m1 - m0; dm_b += m2 - m1; dm_a += m3 - m2;
      tc::tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int sl = 0; sl < SLABS; ++sl) {
          const uint64_t db_hi = db_base + (uint64_t)((s * STAGE_BYTES + sl * 2 * SLAB_BYTES) >> 4);
          const uint64_t db_lo = db_hi + (SLAB_BYTES >> 4);
          const uint32_t a_hi = tmem_a0 + (uint32_t)(s * 128 + sl * 64), a_lo = a_hi + 32;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)(k * 2);
            const uint32_t acol = (uint32_t)(k * 8);
            const uint32_t first = (seg_first && sl == 0 && k == 0) ? 0u : 1u;
            tc::umma_tf32_ts(tmem_main, a_hi + acol, db_hi + adv, idesc, first);
            if (a.passes == 3) {
              tc::umma_tf32_ts(tmem_small, a_hi + acol, db_lo + adv, idesc, first);
              tc::umma_tf32_ts(tmem_small, a_lo + acol, db_hi + adv, idesc, 1u);
            }
          }
        }
        tc::umma_commit(bar_empty0 + 8 * s);
        if (seg_last) tc::umma_commit(bar_accfull);
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
    if (a.dbg && lane == 0) {
      long long* d = a.dbg + ((long long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16;
      d[0] = dm_acc; d[1] = dm_b; d[2] = dm_a; d[3] = TCD_CLK() - dm_start; d[4] = nchunks;
    }
  } else if (warp == 1 + NPROD_WARPS) {
    // =========================== table streamer ===========================
    if (lane == 0) {
      const int ENT = g.AH + g.AL + g.BH + LAST;
      const uint32_t row_b = TS_ * 4;
      const uint32_t bytes = (uint32_t)TE * row_b;
      const long long chunk0 = pbeg / CH;     // per_split is a multiple of CH
      int ts = 0;
      uint32_t ph = 1;
      for (int c = 0; c < nchunks; ++c) {
        tc::mbar_wait(bar_tempty0 + 8 * ts, ph);
        const float* src = a.tables + (chunk0 + c) * (long long)ENT * TS_;
        const uint32_t dst = tc::smem_u32(tabs + ts * TE * TS_);
        const uint32_t bar = bar_tfull0 + 8 * ts;
        tc::mbar_arrive_expect_tx(bar, bytes);
        tc::bulk_g2s(dst, src + (long long)tt.ah0 * TS_, (uint32_t)tt.nah * row_b, bar);
        tc::bulk_g2s(dst + (uint32_t)tt.nah * row_b, src + (long long)g.AH * TS_, (uint32_t)g.AL * row_b, bar);
        tc::bulk_g2s(dst + (uint32_t)(tt.nah + g.AL) * row_b, src + (long long)(g.AH + g.AL + tt.bh0) * TS_, (uint32_t)tt.nbh * row_b, bar);
        tc::bulk_g2s(dst + (uint32_t)(tt.nah + g.AL + tt.nbh) * row_b, src + (long long)(g.AH + g.AL + g.BH) * TS_, (uint32_t)LAST * row_b, bar);
        if (++ts == TSTAGES) { ts = 0; ph ^= 1; }
      }
    }
  } else {
    // =========================== producers ===========================
    const int pw = warp - 1;                 // 0..7
    const bool is_a = pw < 4;                // warps 1..4 own the 128 A rows, warps 5..8 the 128 B rows
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
    const int row = is_a ? quad * 32 + lane : (pw - 4) * 32 + lane;
    // loop-invariant table rows of this thread's operand row (padding rows multiply the all-zero table row)
    int e_hi = TE, e_lo = TE, e_g = -1;   // e_g: gout table row (three-level B rows only)
    if (is_a) {
      const int ai = a0 + row;
      if (ai < g.A) { e_hi = ai / g.AL - tt.ah0; e_lo = tt.nah + ai % g.AL; }
    } else {
      const int ni = n0 + row;
      if (ni < g.N) {
        const int base = tt.nah + g.AL + tt.nbh, rem = ni % BLO;
        e_hi = tt.nah + g.AL + (ni / BLO - tt.bh0);
        if (THREE) { e_lo = base + rem / O; e_g = base + g.BL + rem % O; }
        else e_lo = base + rem;
      }
    }
    const float4* th4[TSTAGES];
    const float4* tl4[TSTAGES];
    const float4* tg4[TSTAGES];
#pragma unroll
    for (int b = 0; b < TSTAGES; ++b) {
      th4[b] = (const float4*)(e_hi < TE ? tabs + (b * TE + e_hi) * TS_ : zrow);
      tl4[b] = (const float4*)(e_lo < TE ? tabs + (b * TE + e_lo) * TS_ : zrow);
      tg4[b] = (const float4*)(e_g >= 0 ? tabs + (b * TE + e_g) * TS_ : onerow);
    }
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t brow_off = (uint32_t)(row * 128);
    const int bsw = row & 7;
    // promotion: this thread accumulates row (quad*32 + lane), columns [chalf*64, chalf*64 + 64)
    const int chalf = pw >> 2;
    float racc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) racc[i] = 0.f;
    int next_drain = 0;
    auto drain = [&](int seg) {
      tc::mbar_wait(bar_accfull, (uint32_t)(seg & 1));
      tc::tc_fence_after();
#pragma unroll
      for (int cb = 0; cb < 64; cb += 32) {
        float v[32];
        tc::tmem_ld32(tmem_main + lane_base + (uint32_t)(chalf * 64 + cb), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) racc[cb + i] += v[i];
        if (a.passes == 3) {
          tc::tmem_ld32(tmem_small + lane_base + (uint32_t)(chalf * 64 + cb), v);
#pragma unroll
          for (int i = 0; i < 32; ++i) racc[cb + i] += v[i];
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(bar_accempty);
    };

    int s = 0, ts = 0;
    uint32_t phe = 1;   // parity to wait for on the empty barrier (fresh barriers pass parity 1)
    uint32_t tph = 0;   // parity of the table-full barrier
    long long dp[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int c = 0; c < nchunks; ++c) {
      long long q0 = TCD_CLK();
      // (1) the table rows of this chunk have landed
      tc::mbar_wait(bar_tfull0 + 8 * ts, tph);
      long long q1 = TCD_CLK(), q2 = q1, q3 = q1, q4 = q1;
      const float4* th = th4[ts];
      const float4* tl = tl4[ts];
      const float4* tg = tg4[ts];
      // (2) one operand row per thread; wait until the MMAs that read this stage two chunks ago are done
      tc::mbar_wait(bar_empty0 + 8 * s, phe);
      long long q5 = TCD_CLK();
      tc::tc_fence_after();
#pragma unroll
      for (int sl = 0; sl < SLABS; ++sl) {
        float hi[32], lo[32];
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          float4 h4 = th[sl * 8 + q4];
          const float4 l4 = tl[sl * 8 + q4];
          if (THREE) {
            const float4 g4 = tg[sl * 8 + q4];
            h4.x *= g4.x; h4.y *= g4.y; h4.z *= g4.z; h4.w *= g4.w;
          }
          tc::split_tf32(h4.x * l4.x, hi[4 * q4 + 0], lo[4 * q4 + 0]);
          tc::split_tf32(h4.y * l4.y, hi[4 * q4 + 1], lo[4 * q4 + 1]);
          tc::split_tf32(h4.z * l4.z, hi[4 * q4 + 2], lo[4 * q4 + 2]);
          tc::split_tf32(h4.w * l4.w, hi[4 * q4 + 3], lo[4 * q4 + 3]);
        }
        if (is_a) {
          const uint32_t dst = tmem_a0 + lane_base + (uint32_t)(s * 128 + sl * 64);
          tc::tmem_st32(dst, hi);
          if (a.passes == 3) tc::tmem_st32(dst + 32, lo);
        } else {
          unsigned char* st = stages + s * STAGE_BYTES + sl * 2 * SLAB_BYTES + brow_off;
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const uint32_t off = (uint32_t)((q4 ^ bsw) << 4);
            *(float4*)(st + off) = make_float4(hi[4 * q4], hi[4 * q4 + 1], hi[4 * q4 + 2], hi[4 * q4 + 3]);
            if (a.passes == 3)
              *(float4*)(st + SLAB_BYTES + off) = make_float4(lo[4 * q4], lo[4 * q4 + 1], lo[4 * q4 + 2], lo[4 * q4 + 3]);
          }
        }
      }
      if (is_a) {
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(bar_fullA0 + 8 * s);
      } else {
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(bar_fullB0 + 8 * s);
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(bar_tempty0 + 8 * ts);   // this warp no longer reads the table buffer
      if (++ts == TSTAGES) { ts = 0; tph ^= 1; }
      long long q6 = TCD_CLK();
      dp[0] += q1 - q0; dp[1] += q2 - q1; dp[2] += q3 - q2; dp[3] += q4 - q3; dp[4] += q5 - q4; dp[5] += q6 - q5;
      if (++s == STAGES) { s = 0; phe ^= 1; }
      // the first chunk of a new segment is now queued behind the previous segment: promote that segment while the
      // tensor core finishes it, so the MMA warp finds work ready the moment the accumulators are handed back
      if ((c % SEG_CHUNKS) == 0 && c > 0) drain(next_drain++);
      dp[6] += TCD_CLK() - q6;
    }
    if (a.dbg && lane == 0 && (warp == 1 || warp == 5)) {
      long long* d = a.dbg + ((long long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + (warp == 1 ? 5 : 5);
      if (warp == 1) for (int i = 0; i < 7; ++i) d[i] = dp[i];
    }
    const int last_seg = (nchunks - 1) / SEG_CHUNKS;
    while (next_drain <= last_seg) drain(next_drain++);

    // =========================== epilogue: registers -> partial tile ===========================
    const int arow = a0 + quad * 32 + lane;
    if (arow < g.A) {
      float* prow = a.part + ((long long)blockIdx.z * g.A + arow) * (long long)g.N;
#pragma unroll
      for (int cb = 0; cb < 64; cb += 4) {
        const int nb = n0 + chalf * 64 + cb;
        if (nb + 4 <= g.N && (g.N & 3) == 0) {
          *(float4*)(prow + nb) = make_float4(racc[cb], racc[cb + 1], racc[cb + 2], racc[cb + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (nb + i < g.N) prow[nb + i] = racc[cb + i];
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

size_t dcore_tc_smem(const EpsGeom& g, int three) {
  const int TE = max_tile_entries(g, three);
  return 1024 + (size_t)STAGES * STAGE_BYTES + (size_t)(TSTAGES * TE + 2) * TS_ * 4 + (3 * STAGES + 2 * TSTAGES + 2) * 8 + 16;
}
// two-level B rows when their tables fit in shared memory, else three-level; -1: neither fits
inline int pick_three(const EpsGeom& g) {
  if (dcore_tc_smem(g, 0) <= TC_SMEM_LIMIT) return 0;
  if (dcore_tc_smem(g, 1) <= TC_SMEM_LIMIT) return 1;
  return -1;
}
inline size_t table_floats(const EpsGeom& g, int three) {
  const long long nchunks = (g.P + CH - 1) / CH;
  return (size_t)nchunks * (size_t)(g.AH + g.AL + g.BH + last_section(g, three)) * TS_;
}

inline void dcore_split(const EpsGeom& g, long long* per_split, int* splits) {
  // about two waves of one-CTA-per-SM (the kernel uses all of TMEM and most of shared memory)
  long long tiles = (long long)((g.A + BM - 1) / BM) * ((g.N + BN - 1) / BN);
  long long want = (2 * 148) / tiles;
  if (want < 1) want = 1;
  long long per = (g.P + want - 1) / want;
  per = ((per + CH - 1) / CH) * CH;
  const long long min_per = (long long)CH * SEG_CHUNKS * 2;  // do not split below a couple of segments
  if (per < min_per) per = min_per;
  *per_split = per;
  *splits = (int)((g.P + per - 1) / per);
}

}  // namespace

bool tc_supported(const EpsGeom& g, int kind) {
  if (kind != 1) return tcg_supported(g, kind);  // forward / input-gradient GEMMs live in eps_tc_gemm.cu
  if (g.P >= (1ll << 31) / (g.Q > g.O ? g.Q : g.O)) return false;  // 32-bit patch index math
  if (g.A < 64 || g.N < 64) return false;    // tiles would be mostly padding: the CUDA-core family is the better fit
  if (g.P < 4096) return false;              // tiny reductions are launch-bound either way
  return pick_three(g) >= 0 && tc16_dcore_supported(g);   // both arithmetics serve the same shapes
}

size_t tc_workspace_bytes(const EpsGeom& g, int kind) {
  if (kind == 1) {
    long long per;
    int splits;
    dcore_split(g, &per, &splits);
    const int three = pick_three(g);
    const size_t tf32 = ((size_t)splits * g.A * g.N + table_floats(g, three < 0 ? 0 : three)) * sizeof(float) + 256;
    const size_t f16 = tc16_dcore_workspace_bytes(g);
    return tf32 > f16 ? tf32 : f16;
  }
  return tcg_workspace_bytes(g, kind);
}

int tc_backward_core(const EpsGeom& g, const float* x, const float* gout, float* dcore, void* ws, int passes,
                     cudaStream_t st) {
  if (passes == 6) return tc16_backward_core(g, x, gout, dcore, ws, st);   // split fp16 (eps_tc_dcore.cu)
  const int three = pick_three(g);
  if (three < 0) return dctn_set_error(-2, "tcgen05 core-gradient kernel: tables of this shape do not fit in shared memory");
  const size_t smem = dcore_tc_smem(g, three);
  TcDcoreArgs a{};
  a.g = g; a.part = (float*)ws; a.passes = passes; a.three = three;
  int splits;
  dcore_split(g, &a.per_split, &splits);
  float* tables = a.part + (((size_t)splits * g.A * g.N + 63) & ~(size_t)63);
  a.tables = tables;
  {
    const size_t bsm = (size_t)(g.n * g.Q + g.O) * CH * sizeof(float);
    if (bsm > 200 * 1024) return dctn_set_error(-2, "table kernel needs %zu bytes of shared memory", bsm);
    DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(build_tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
    build_tables_kernel<<<(unsigned)((g.P + CH - 1) / CH), 256, bsm, st>>>(g, x, gout, tables, three);
    dctn_count_launch();
    DCTN_CUDA_CHECK_RET(cudaGetLastError());
  }
  auto kern = three ? tc_dcore_kernel<true> : tc_dcore_kernel<false>;
  DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((g.A + BM - 1) / BM, (g.N + BN - 1) / BN, splits);
  a.dbg = nullptr;
#ifdef DCTN_TCG_TIMING   // cycle probes: timing builds only (allocates, synchronises, not thread-safe)
  static long long* dbg_buf = nullptr;
  const int ncta = (int)(grid.x * grid.y * grid.z);
  if (getenv("DCTN_TCG_DEBUG") && ncta <= 4096) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 4096 * 16 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 4096 * 16 * sizeof(long long), st);
    a.dbg = dbg_buf;
  }
#endif
  kern<<<grid, NTHREADS_TC, smem, st>>>(a);
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
    fprintf(stderr, "[dcore dbg] ctas=%d chunks/cta=%.0f per-chunk cycles: mma wait acc %.0f B %.0f A %.0f total %.0f | producer: publish %.0f "
            "bar1 %.0f prefetch+tables %.0f bar2 %.0f wait-empty %.0f rows+signal %.0f drain %.0f\n", ncta, nch / ncta,
            sum[0] / nch, sum[1] / nch, sum[2] / nch, sum[3] / nch, sum[5] / nch, sum[6] / nch, sum[7] / nch, sum[8] / nch,
            sum[9] / nch, sum[10] / nch, sum[11] / nch);
  }
#endif
  return launch_reduce_partials<float>(a.part, dcore, (long long)g.A * g.N, splits, st);
}
