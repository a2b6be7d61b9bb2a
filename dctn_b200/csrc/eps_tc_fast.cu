// Register-table variant of the tcgen05 forward / input-gradient GEMMs (split-fp16 arithmetic only), for layers whose
// in_size Q is a power of two (config 2: Q = 2 and Q = 4).  Same pipeline as eps_tc_gemm.cu:
//
//   C[p][c] = sum_k Gen[p][k] * Bop[k][c]       128 patches per CTA = 128 TMEM lanes, column tiles of BN, K in stages of 64
//
// What changes is how the generated operand is produced.  The generic kernel looks two table entries up in shared
// memory for every generated element; with fp16 MMAs (twice the TF32 rate) those 128 LDS per thread and stage made
// SHARED-MEMORY BANDWIDTH the limiter — measured with the cycle probes: 1665 cycles of generation per stage against
// 1152 cycles of MMA, the rest of the pipe waiting (profiles/r01_f16_phase_cycles.txt).  Here
//   * the K index is ordered (e, kl) with kl in [0, KLR), KLR = Q^cl a power of two <= 16 (compile time): the lo-group
//     values of a thread's patch, TL[KLR], live in REGISTERS for the whole kernel, and a stage of 64 k needs only
//     64/KLR loads from the hi table (4 for KLR = 16, down from 128);
//   * for the input gradient the hi table carries the gout factor: e = (o, hi-group entry), i.e. the reduction index is
//     ordered (o, b) instead of (b, o); the packed core image is built in the same order (K order is free in a GEMM);
//   * the forward epilogue uses the same trick for KR2[p][b] = EH[b / ELR] * EL[b % ELR]: EL in registers.
// No index tables, no per-element shared-memory traffic; shared memory is left to the core stages (TMA writes, MMA
// reads).  Everything else — TS-form MMAs with the A operand in tensor memory, the main/small accumulator pair, the
// power-of-two range normalisation, bulk-copied pre-packed core — is as in eps_tc_gemm.cu.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "eps_kernels.h"
#include "tc_common.cuh"

#ifdef DCTN_TCG_TIMING
#define TCF_CLK() clock64()
#else
#define TCF_CLK() 0ll
#endif

namespace {

constexpr int FBM = 128;
constexpr int FKS = 64;          // K values per stage (one 128-byte row of fp16)
constexpr int F_ASTAGES = 3;     // at most: A stages in TMEM of (hi 32 + lo 32 columns); 3 when BN <= 160, else 2
constexpr int F_MAX_BSTAGES = 4;
constexpr int F_MAX_BN = 192;    // main + small accumulators (2 * BN) + 128 or 192 columns of A <= 512
constexpr int F_THREADS = 512;   // warp 0: bulk copies, warp 1: MMA, warp 2: TMEM alloc, warps 4-7: producers, warps 8-15: epilogue
                                 // (two warps per TMEM lane quadrant, alternating 32-column batches: the epilogue of a tile is
                                 // not overlapped with the next tile's MMAs, so its length is paid in full)
constexpr int F_EPI_WARPS = 8;
// FMODE_LOO: the input-gradient GEMM with the first stage of the leave-one-out contraction fused into the epilogue:
// instead of the P x A matrix dKR1 it writes, per patch, W[p] = (Whi[EHE] | Wlo[ELR]),
//   Whi[eh] = sum_el dKR1[eh*ELR + el] * TL1[el],   Wlo[el] = sum_eh dKR1[eh*ELR + el] * TH1[eh]
// (TH1 / TL1: hi / lo group tables of the FIRST half).  A 16x smaller output (config 2, layer 2: 80 instead of 1024
// floats per patch), and the separate pass that re-read dKR1 (1.28 ms) disappears.
// FMODE_LOOX: the second stage as well — the kernel writes d x_j (dxp[p][j][q]) for the first-half factors directly:
// each finished Whi[eh] is scattered into per-thread accumulators in shared memory, the lo group is contracted from the
// Wlo registers at the end; the first-half table comes from products of the (normalised) x kept in shared memory.
enum { FMODE_STORE = 0, FMODE_FWD = 1, FMODE_LOO = 2, FMODE_LOOX = 3 };
constexpr size_t F_SMEM_LIMIT = 227 * 1024;

struct FastArgs {
  EpsGeom g;
  const float* x;
  const float* gout;
  long long p0;
  int np;
  int jh0, cnth, cntl;   // generated group: hi factors [jh0, jh0+cnth), lo factors [jh0+cnth, jh0+cnth+cntl) -> TL registers
  int KHE;               // hi-table entries: Q^cnth (forward) or O * Q^cnth (input gradient, entry = o * Q^cnth + e)
  int withG;
  int Kdim;              // KHE * KLR
  int ej0, ecnth, ecntl; // forward epilogue: KR2 hi factors [ej0, ej0+ecnth), lo factors after them -> EL registers
  int EHE, ELR;          // Q^ecnth, Q^ecntl
  int Ncols, ntiles, nk, BN, bstages;
  int kseg, nseg;        // K stages per accumulation segment, segments per column tile (see tc::seg_stages)
  int dbuf;              // 1: two MAIN accumulators alternate between (virtual) tiles, the small one is parked (see the epilogue)
  int region_floats;     // size of the staging / register-table / parking region
  const float* packed;   // [ntiles][nk][hi|lo][BN rows x 128 bytes], swizzled
  const uint32_t* core_absmax;
  float* out;            // FMODE_STORE: [np][ldc]; FMODE_FWD: out[P][O]
  long long ldc;
  float* tsave;          // FMODE_FWD: optional T[P][Ncols]
  long long* dbg;
};

// ------------------------------------------------------------------------------------------------ core packing
// dst[((tile*nk + kc)*2 + part) * BN*128 bytes + swizzled(row rr, 16-byte chunk c16)] = 8 consecutive k of core column c
//   FMODE_FWD  : element(c = o*Bn + b, k = a)        = core[(a*Bn + b)*O + o]
//   FMODE_STORE: element(c = a, k = o*Bn + b)         = core[a*N + b*O + o]
__global__ void fast_pack_kernel(const float* __restrict__ core, float* __restrict__ dst, EpsGeom g, int mode, int BN, int Ncols,
                                 int Kdim, int ntiles, int nk, const uint32_t* __restrict__ absmax) {
  const long long total = (long long)ntiles * nk * BN * 8;
  const float scale = scalbnf(1.f, tc::core_scale_exp(*absmax));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c16 = (int)(i & 7);
    long long r = i >> 3;
    const int rr = (int)(r % BN);
    r /= BN;
    const int kc = (int)(r % nk);
    const int tile = (int)(r / nk);
    const int c = tile * BN + rr;
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = 0.f;
    if (c < Ncols) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = kc * FKS + c16 * 8 + u;
        if (k < Kdim) {
          long long idx;
          if (mode != FMODE_FWD) {
            const int o = k / g.Bn, b = k - o * g.Bn;
            idx = (long long)c * g.N + (long long)b * g.O + o;
          } else {
            const int o = c / g.Bn, b = c - o * g.Bn;
            idx = ((long long)k * g.Bn + b) * g.O + o;
          }
          v[u] = __ldg(&core[idx]) * scale;
        }
      }
    }
    uint4 hi, lo;
    tc::split_f16x2(v[0], v[1], hi.x, lo.x);
    tc::split_f16x2(v[2], v[3], hi.y, lo.y);
    tc::split_f16x2(v[4], v[5], hi.z, lo.z);
    tc::split_f16x2(v[6], v[7], hi.w, lo.w);
    float* tile_base = dst + ((long long)(tile * nk + kc) * 2) * BN * 32;
    const int off = rr * 32 + ((c16 ^ (rr & 7)) << 2);  // in 4-byte units
    *(uint4*)(tile_base + off) = hi;
    *(uint4*)(tile_base + BN * 32 + off) = lo;
  }
}

// product over `cnt` factors starting at j0 of the normalised x values of patch row pr, for table entry e (digit 0
// slowest); Q = 2^lq in this file's kernels: digits by shifts
__device__ __forceinline__ float kr_entry2(const float* xs, int lq, int j0, int cnt, int e, int pr) {
  float v = 1.f;
  const int Q = 1 << lq;
  for (int u = cnt - 1; u >= 0; --u) {
    v *= xs[(((j0 + u) << lq) + (e & (Q - 1))) * 128 + pr];
    e >>= lq;
  }
  return v;
}
// Group table by levels: level L holds, for every patch row, the products of the group's first L factors (digit 0
// slowest), E0 * Q^L entries; level 0 is `scale` (times g0[t] when the group is led by the gout factor).  The levels
// ping-pong between `dst` and `tmp` so that the last one lands in dst: one multiplication per entry and level instead
// of a full product per entry (the per-entry form took 14 k of a CTA's ~20 k setup cycles).  All threads of the CTA.
__device__ __forceinline__ void build_levels(float* dst, float* tmp, const float* xs, int lq, int j0, int cnt, const float* g0,
                                             int E0, float scale, int tid) {
  __syncthreads();                                   // the previous table's last level has been read out of tmp
  float* cur = (cnt & 1) ? tmp : dst;
  for (int idx = tid; idx < E0 * 128; idx += F_THREADS) cur[idx] = g0 ? scale * g0[idx] : scale;
  int E = E0;
  const int qm = (1 << lq) - 1;
  for (int L = 0; L < cnt; ++L) {
    __syncthreads();
    float* nxt = (cur == dst) ? tmp : dst;
    const float* xf = xs + (((j0 + L) << lq) << 7);
    const int En = E << lq;
    for (int idx = tid; idx < En * 128; idx += F_THREADS) {
      const int pr = idx & 127, t = idx >> 7;
      nxt[idx] = cur[((t >> lq) << 7) + pr] * xf[((t & qm) << 7) + pr];
    }
    cur = nxt;
    E = En;
  }
}
// 2^k as a float for -126 <= k <= 127
__device__ __forceinline__ float pow2i(int k) { return __int_as_float((k + 127) << 23); }
// power-of-two normalisation of `cnt` values in registers: returns e (max-abs * 2^-e in [0.5, 1)), scales in place (exact)
template <int MAXC>
__device__ __forceinline__ int normalise(float (&v)[MAXC], int cnt) {
  float m = 0.f;
#pragma unroll
  for (int q = 0; q < MAXC; ++q)
    if (q < cnt) m = fmaxf(m, fabsf(v[q]));
  int e = 0;
  const int ef = (__float_as_int(m) >> 23) & 0xFF;
  if (ef != 0 && ef != 0xFF) e = ef - 126;            // normal numbers: exponent field; 0 / inf / nan: leave unscaled
  else if (m > 0.f && m < 1.f) frexpf(m, &e);          // subnormal maximum (rare)
  if (e != 0) {
    const float s1 = pow2i(-(e / 2)), s2 = pow2i(-(e - e / 2));
#pragma unroll
    for (int q = 0; q < MAXC; ++q)
      if (q < cnt) v[q] = v[q] * s1 * s2;
  }
  return e;
}

// FMODE_LOOX epilogue, one batch of 32 accumulator columns = 32/ELRC runs of ELRC columns (one hi-group entry eh each):
//   Whi[eh] = sum_j v[run][j] * EL[j]  is scattered at once into the hi-group factor gradients dxa[t][digit_t(eh)] with the
//   leave-one-out product of the other hi factors;  Wlo[j] += v[run][j] * TH[eh],  TH[eh] = product of all hi factors
template <int ELRC>
__device__ __forceinline__ void loox_batch(const float (&v)[32], const float (&EL)[16], float (&WLO)[16], const float* xh, float* dxa,
                                           int eh0, int ecnth, int Q, int lq, int pr) {
#pragma unroll
  for (int r = 0; r < 32 / ELRC; ++r) {
    const int eh = eh0 + r;
    int dg[4];
    float xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      dg[u] = 0; xv[u] = 1.f;
      if (u < ecnth) {
        dg[u] = (eh >> (lq * (ecnth - 1 - u))) & (Q - 1);
        xv[u] = xh[((u * Q + dg[u]) << 7) + pr];
      }
    }
    const float p01 = xv[0] * xv[1], p23 = xv[2] * xv[3];
    const float th = p01 * p23;
    float whi = 0.f;
#pragma unroll
    for (int j = 0; j < ELRC; ++j) {
      whi = fmaf(v[r * ELRC + j], EL[j], whi);
      WLO[j] = fmaf(v[r * ELRC + j], th, WLO[j]);
    }
    const float lo3[4] = {xv[1] * p23, xv[0] * p23, p01 * xv[3], p01 * xv[2]};   // leave-one-out products
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (t < ecnth) dxa[((t * Q + dg[t]) << 7) + pr] += whi * lo3[t];
  }
}

// ------------------------------------------------------------------------------------------------ the GEMM
template <int MODE, int KLR>
__global__ void __launch_bounds__(F_THREADS, 1) tc_gemm_fast_kernel(const __grid_constant__ FastArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  const EpsGeom& g = a.g;
  const int Q = g.Q, O = g.O, BN = a.BN, NB = a.bstages;
  constexpr int RUNS = FKS / KLR;                      // hi-table entries per stage
  const uint32_t B_BYTES = (uint32_t)BN * 128;
  const uint32_t STAGE_BYTES = 2 * B_BYTES;
  unsigned char* base = smem_dyn + ((1024u - (tc::smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* stages = base;
  const int nHrows = a.nk * RUNS;                      // >= KHE; the tail rows are zero (K padding)
  float* tabH = (float*)(base + NB * STAGE_BYTES);     // [nHrows][128]
  float* tabEH = tabH + nHrows * 128;                  // FMODE_FWD / FMODE_LOO: [EHE][128] hi-group table of the epilogue's half
  // FMODE_LOOX instead: normalised x of the first half, its exponents, hi-group accumulators, Wlo
  const int mfirst = a.ecnth + a.ecntl;
  float* xh = tabEH;                                   // [mfirst*Q][128]
  int* fe = (int*)(xh + mfirst * Q * 128);             // [mfirst][128]
  float* dxa = (float*)(fe + mfirst * 128);            // [ecnth*Q][128]
  // (dxa, wlo and outs exist twice: one copy per epilogue half, merged at the end)
  float* wlo = (MODE == FMODE_LOOX) ? dxa + 2 * a.ecnth * Q * 128 : tabEH + a.EHE * 128;   // [2][16][128] (LOOX), [1][16][128] (LOO)
  const int eregion = (MODE == FMODE_LOOX) ? (mfirst * Q + mfirst + 2 * a.ecnth * Q + 32) * 128
                    : (MODE == FMODE_LOO ? (a.EHE + 16) * 128 : (MODE == FMODE_FWD ? a.EHE * 128 : 0));
  float* outs = tabEH + eregion;                       // FMODE_FWD: [2][O][128]
  // exponents of the per-patch normalisation: [0] generated group, [1] all other factors, [2] the epilogue's lo group
  int* rowexp = (int*)(outs + (MODE == FMODE_FWD ? 2 * O * 128 : 0));  // [3][128]
  // training forward: per epilogue warp a [32 rows][36] staging tile that turns the thread-per-row accumulator batches
  // into whole 128-byte row segments of T (a direct store from registers writes 32 half-used sectors per instruction)
  // training forward: per epilogue warp a [32 rows][20] staging tile (16 columns at a time).  It shares its memory with
  // regtab, the lo-group tables of the generated operand (rows 0..15) and of the epilogue's half (rows 16..31): those are
  // built by all threads with the other tables and copied into registers by the producer / epilogue threads right after
  // the setup barrier (a named barrier separates the copies from the first staging write)
  float* tstage = (float*)(rowexp + 384);
  float* regtab = tstage;                              // [32][128]
  constexpr int REGION_FLOATS = (F_EPI_WARPS * 32 * 20 > 32 * 128) ? F_EPI_WARPS * 32 * 20 : 32 * 128;
  // parking area of the small accumulator (double-buffered mode): [8 epilogue warps][32 words][32 lanes] of packed bf16
  // pairs.  The training forward keeps its staging tiles, so the area follows them; the other modes use the region only
  // for the register tables during the setup, and the area overlays it.
  uint32_t* park = (uint32_t*)((MODE == FMODE_FWD) ? tstage + REGION_FLOATS : tstage);
  uint64_t* bars = (uint64_t*)(tstage + a.region_floats);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * F_MAX_BSTAGES + 2 * F_ASTAGES + 6);
  const uint32_t bar_fullB0 = tc::smem_u32(bars), bar_emptyB0 = bar_fullB0 + 8 * F_MAX_BSTAGES;
  const uint32_t bar_fullA0 = bar_emptyB0 + 8 * F_MAX_BSTAGES, bar_emptyA0 = bar_fullA0 + 8 * F_ASTAGES;
  // accumulators: accfull[b] (MMAs of a tile in main accumulator b are complete), mainempty[b] (its epilogue is done),
  // smallempty (the small accumulator has been parked).  Single-buffered mode uses accfull[0] / mainempty[0] only.
  const uint32_t bar_accfull0 = bar_emptyA0 + 8 * F_ASTAGES, bar_mainempty0 = bar_accfull0 + 16, bar_smallempty = bar_mainempty0 + 16;
  const bool DB = a.dbuf != 0;
  // setup-only scratch aliased onto the B stages: x [n*Q][128], gout [O][128], exponents [(n+1)][128]
  float* xs = (float*)stages;
  float* gsx = xs + g.n * Q * 128;
  int* fexp = (int*)(gsx + O * 128);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long dbg_t_entry = TCF_CLK();
  const int pl0 = blockIdx.x * FBM;
  const long long pt0 = a.p0 + pl0;

  // ---------------- setup
  if (tid == 0) {
    for (int s = 0; s < F_MAX_BSTAGES; ++s) {
      tc::mbar_init(bar_fullB0 + 8 * s, 1);
      tc::mbar_init(bar_emptyB0 + 8 * s, 1);
    }
    for (int s = 0; s < F_ASTAGES; ++s) {
      tc::mbar_init(bar_fullA0 + 8 * s, 4);
      tc::mbar_init(bar_emptyA0 + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(bar_accfull0 + 8 * b, 1);
      tc::mbar_init(bar_mainempty0 + 8 * b, F_EPI_WARPS);
    }
    tc::mbar_init(bar_smallempty, F_EPI_WARPS);
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(tc::smem_u32(tmem_slot), 512);
  const int lq = 31 - __clz(Q);                        // Q is a power of two <= 16 (host check)
  {
    // every thread owns one patch row (pr) and every third factor: it loads the factor's Q values with vector loads,
    // normalises them (range normalisation, see eps_tc_gemm.cu) and stores values and exponent — one pass, all loads of
    // a thread independent of each other
    const int pr = tid & 127, slot = tid >> 7;         // F_THREADS = 4 * 128
    const long long p = pt0 + pr;
    const bool valid = p < g.P;
    const long long org = valid ? patch_origin(g, p) : 0;
    if (Q <= 4) {
      // up to four factors per round, every load of the round issued before the first use (one DRAM round trip per
      // round instead of one per factor: the loads were 5-8 k cycles of the setup)
      for (int j0 = slot; j0 < g.n; j0 += 4 * (F_THREADS / 128)) {
        float4 t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = j0 + k * (F_THREADS / 128);
          t[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid && j < g.n) {
            const float* px = a.x + org + g.foff[j];
            if (Q == 2) { const float2 u = __ldg((const float2*)px); t[k].x = u.x; t[k].y = u.y; }
            else t[k] = __ldg((const float4*)px);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = j0 + k * (F_THREADS / 128);
          if (j < g.n) {
            float v[4] = {t[k].x, t[k].y, t[k].z, t[k].w};
            fexp[j * 128 + pr] = normalise<4>(v, Q);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (q < Q) xs[((j << lq) + q) * 128 + pr] = v[q];
          }
        }
      }
    } else {
#pragma unroll 2
      for (int j = slot; j < g.n; j += F_THREADS / 128) {
        float v[16];
        const float* px = a.x + org + g.foff[j];
        if (!valid) {
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = 0.f;
        } else {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            if (q4 * 4 < Q) {
              const float4 t = __ldg((const float4*)px + q4);
              v[4 * q4] = t.x; v[4 * q4 + 1] = t.y; v[4 * q4 + 2] = t.z; v[4 * q4 + 3] = t.w;
            }
        }
        fexp[j * 128 + pr] = normalise<16>(v, Q);
#pragma unroll
        for (int q = 0; q < 16; ++q)
          if (q < Q) xs[((j << lq) + q) * 128 + pr] = v[q];
      }
    }
    if (slot == 0) {
      int e = 0;
      if (a.withG) {
        // gout rows are O floats (any O): scalar loads, two passes over the row in shared memory
        float m = 0.f;
        for (int o = 0; o < O; ++o) {
          const float t = valid ? __ldg(&a.gout[p * O + o]) : 0.f;
          gsx[o * 128 + pr] = t;
          m = fmaxf(m, fabsf(t));
        }
        e = tc::norm_exp(m);
        if (e != 0) {
          const float s1 = pow2i(-(e / 2)), s2 = pow2i(-(e - e / 2));
          for (int o = 0; o < O; ++o) gsx[o * 128 + pr] = gsx[o * 128 + pr] * s1 * s2;
        }
      }
      fexp[g.n * 128 + pr] = e;
    }
  }
  __syncthreads();
  const long long dbg_t_loaded = TCF_CLK();
  if (tid < 128) {
    int ea = 0, eb = 0, el = 0;
    for (int j = 0; j < g.n; ++j) {
      const bool in_gen = (j >= a.jh0 && j < a.jh0 + a.cnth + a.cntl);
      if (in_gen) ea += fexp[j * 128 + tid];
      else eb += fexp[j * 128 + tid];
      if (j >= a.ej0 + a.ecnth && j < a.ej0 + a.ecnth + a.ecntl) el += fexp[j * 128 + tid];
    }
    if (a.withG) ea += fexp[g.n * 128 + tid];
    rowexp[tid] = ea;
    rowexp[128 + tid] = eb;
    rowexp[256 + tid] = el;
  }
  {
    // zero rows: K padding of the hi table, unused rows of the lo-group tables
    for (int idx = tid + a.KHE * 128; idx < nHrows * 128; idx += F_THREADS) tabH[idx] = 0.f;
    for (int idx = tid; idx < 32 * 128; idx += F_THREADS) regtab[idx] = 0.f;
    float* ytmp = (float*)(fexp + (g.n + 1) * 128);     // level scratch, in the (still unused) B stages like xs
    // hi table: entry e (forward) or (o, e) (input gradient: the gout factor leads); carries the 2^15 of the generated row
    build_levels(tabH, ytmp, xs, lq, a.jh0, a.cnth, a.withG ? gsx : nullptr, a.withG ? O : 1, 32768.f, tid);
    if (MODE == FMODE_FWD || MODE == FMODE_LOO) build_levels(tabEH, ytmp, xs, lq, a.ej0, a.ecnth, nullptr, 1, 1.f, tid);
    build_levels(regtab, ytmp, xs, lq, a.jh0 + a.cnth, a.cntl, nullptr, 1, 1.f, tid);                       // rows [0, KLR)
    if (MODE != FMODE_STORE) build_levels(regtab + 16 * 128, ytmp, xs, lq, a.ej0 + a.ecnth, a.ecntl, nullptr, 1, 1.f, tid);   // rows [16, 16 + ELR)
    if (MODE == FMODE_LOOX) {
      for (int idx = tid; idx < mfirst * Q * 128; idx += F_THREADS) xh[idx] = xs[a.ej0 * Q * 128 + idx];
      for (int idx = tid; idx < mfirst * 128; idx += F_THREADS) fe[idx] = fexp[a.ej0 * 128 + idx];
      for (int idx = tid; idx < 2 * a.ecnth * Q * 128; idx += F_THREADS) dxa[idx] = 0.f;
    }
    if (MODE == FMODE_FWD)
      for (int idx = tid; idx < 2 * O * 128; idx += F_THREADS) outs[idx] = 0.f;
  }
  tc::tc_fence_before();
  __syncthreads();  // tables complete; the scratch aliasing the stages is dead from here on
  tc::tc_fence_after();
  const long long dbg_t_tables = TCF_CLK();
  // lo-group values of this thread's patch: registers for the rest of the kernel
  tc::f32x2_t TL2[KLR / 2];
  float EL[16];
  {
    const int pr = (warp & 3) * 32 + lane;
    if (warp >= 4 && warp < 8) {
#pragma unroll
      for (int j = 0; j < KLR; j += 2) TL2[j / 2] = tc::pack2(regtab[j * 128 + pr], regtab[(j + 1) * 128 + pr]);
    }
    if (MODE != FMODE_STORE && warp >= 8) {
#pragma unroll
      for (int j = 0; j < 16; ++j) EL[j] = regtab[(16 + j) * 128 + pr];
    }
  }
  if (warp >= 4) asm volatile("bar.sync 2, %0;" ::"n"(32 * (4 + F_EPI_WARPS)) : "memory");   // regtab consumed: tstage may be written
  const int core_exp = tc::core_scale_exp(__ldg(a.core_absmax));
  const long long dbg_t_setup = TCF_CLK();
  // tensor memory: main accumulator(s) | small accumulator | A stages
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t nacc = DB ? 3u : 2u;
  const uint32_t tmem_small = tmem_base + (nacc - 1u) * (uint32_t)BN;
  const uint32_t tmem_a0 = tmem_base + nacc * (uint32_t)BN;   // stage s: hi at +64*s, lo at +64*s + 32
  const int total_it = a.ntiles * a.nk;
  const int NA = ((int)nacc * BN + 64 * F_ASTAGES <= 512) ? F_ASTAGES : 2;   // A stages that fit beside the accumulators

  if (warp == 0) {
    // =========================== bulk-copy issuer (B operand) ===========================
    if (tc::elect_one()) {
      int s = 0;
      uint32_t ph = 1;
      const float* src = a.packed;
      for (int i = 0; i < total_it; ++i) {
        tc::mbar_wait(bar_emptyB0 + 8 * s, ph);
        const uint32_t sb = tc::smem_u32(stages + s * STAGE_BYTES);
        tc::mbar_arrive_expect_tx(bar_fullB0 + 8 * s, STAGE_BYTES);
        tc::bulk_g2s(sb, src, STAGE_BYTES, bar_fullB0 + 8 * s);   // hi and lo parts are contiguous
        src += 2 * BN * 32;
        if (++s == NB) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = tc::make_idesc_f16(FBM, BN);
    long long dbg_waitA = 0, dbg_waitB = 0, dbg_waitAcc = 0, dbg_start = TCF_CLK();
    int sa = 0, sb_ = 0;
    uint32_t pha = 0, phb = 0;
    const uint64_t db_base = tc::make_sw128_kmajor_desc(tc::smem_u32(stages));
    const uint32_t stage_adv = STAGE_BYTES >> 4, part_adv = B_BYTES >> 4;
    // The MMAs are asynchronous, so the issuing thread has a whole stage of tensor time to spare: the full barriers of
    // stage i+1 are PROBED in the middle of issuing stage i (after half of its MMAs); when they have completed, the next
    // stage's MMAs go out back to back at the stage boundary instead of after two barrier polls.  (A blocking wait at
    // that point was measured to hurt: it holds back the second half of the current stage whenever an operand is late.)
    bool ready = false;     // the current stage's full barriers have already been waited for
    // A column tile's K loop is cut into segments of at most a.kseg stages, each accumulated from zero and added to the
    // result by the epilogue in fp32 (round to nearest): the tensor core truncates its accumulator toward zero at every
    // MMA, a bias that grows with the length of the chain (tc::seg_stages).  One segment = one "virtual tile".
    const int nvt = a.ntiles * a.nseg;
    for (int u = 0; u < nvt; ++u) {
      const int sg = u % a.nseg;
      const int kc0 = sg * a.kseg, kc1 = (kc0 + a.kseg < a.nk) ? kc0 + a.kseg : a.nk;
      const int mb = DB ? (u & 1) : 0;
      const uint32_t tmem_main = tmem_base + (uint32_t)(mb * BN);
      long long ta = TCF_CLK();
      if (!DB) {
        if (u > 0) tc::mbar_wait(bar_mainempty0, (uint32_t)((u - 1) & 1));
      } else {
        if (u > 0) tc::mbar_wait(bar_smallempty, (uint32_t)((u - 1) & 1));                       // the previous tile's small part is parked
        if (u >= 2) tc::mbar_wait(bar_mainempty0 + 8 * mb, (uint32_t)(((u - 2) >> 1) & 1));      // tile u-2 has left this main accumulator
      }
      dbg_waitAcc += TCF_CLK() - ta;
      tc::tc_fence_after();
      for (int kc = kc0; kc < kc1; ++kc) {
        if (!ready) {
          long long t0 = TCF_CLK();
          tc::mbar_wait(bar_fullB0 + 8 * sb_, phb);
          long long t1 = TCF_CLK();
          tc::mbar_wait(bar_fullA0 + 8 * sa, pha);
          long long t2 = TCF_CLK();
          dbg_waitB += t1 - t0; dbg_waitA += t2 - t1;
          tc::tc_fence_after();
        }
        const uint64_t db_hi = db_base + (uint64_t)(sb_ * stage_adv);
        const uint64_t db_lo = db_hi + part_adv;
        const uint32_t a_hi = tmem_a0 + (uint32_t)(sa * 64), a_lo = a_hi + 32;
        auto issue = [&](int k) {
          const uint64_t adv = (uint64_t)(k * 2);
          const uint32_t acol = (uint32_t)(k * 8);
          const uint32_t first = (kc == kc0 && k == 0) ? 0u : 1u;
          tc::umma_f16_ts(tmem_main, a_hi + acol, db_hi + adv, idesc, first);
          tc::umma_f16_ts(tmem_small, a_hi + acol, db_lo + adv, idesc, first);
          tc::umma_f16_ts(tmem_small, a_lo + acol, db_hi + adv, idesc, 1u);
        };
        if (tc::elect_one()) { issue(0); issue(1); }
        __syncwarp();
        int sa_n = sa + 1, sb_n = sb_ + 1;
        uint32_t pha_n = pha, phb_n = phb;
        if (sa_n == NA) { sa_n = 0; pha_n ^= 1; }
        if (sb_n == NB) { sb_n = 0; phb_n ^= 1; }
        // look ahead WITHOUT blocking (the second half of this stage's MMAs must not be held back): if the next stage's
        // operands have landed, its MMAs follow this stage's back to back; otherwise the blocking wait happens at the top
        ready = !(u == nvt - 1 && kc == kc1 - 1) && tc::mbar_test(bar_fullB0 + 8 * sb_n, phb_n) && tc::mbar_test(bar_fullA0 + 8 * sa_n, pha_n);
        if (ready) tc::tc_fence_after();
        if (tc::elect_one()) {
          issue(2); issue(3);
          tc::umma_commit(bar_emptyA0 + 8 * sa);
          tc::umma_commit(bar_emptyB0 + 8 * sb_);
          if (kc == kc1 - 1) tc::umma_commit(bar_accfull0 + 8 * mb);
        }
        __syncwarp();
        sa = sa_n; pha = pha_n; sb_ = sb_n; phb = phb_n;
      }
    }
    if (a.dbg && lane == 0) {
      long long* d = a.dbg + (long long)blockIdx.x * 8;
      d[0] = dbg_waitA; d[1] = dbg_waitB; d[2] = dbg_waitAcc; d[3] = TCF_CLK() - dbg_start;
    }
  } else if (warp >= 4 && warp < 8) {
    // =========================== A producers: one patch row (= TMEM lane) per thread ===========================
    const int pr = (warp & 3) * 32 + lane;
    const float* th = tabH + pr;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    int sa = 0;
    uint32_t phe = 1;
    long long dbg_pwait = 0, dbg_pst = 0, dbg_pgen = 0, tprev = TCF_CLK();
    for (int t = 0; t < a.ntiles; ++t) {
      const float* thk = th;
      for (int kc = 0; kc < a.nk; ++kc) {
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int r = 0; r < RUNS; ++r) {
          const float h = thk[r * 128];
          const tc::f32x2_t h2 = tc::pack2(h, h);
#pragma unroll
          for (int j = 0; j < KLR; j += 2)
            tc::split_f16x2_p(tc::mul2(h2, TL2[j / 2]), hi[(r * KLR + j) >> 1], lo[(r * KLR + j) >> 1]);
        }
        thk += RUNS * 128;
        long long t0 = TCF_CLK();
        tc::mbar_wait(bar_emptyA0 + 8 * sa, phe);
        long long t1 = TCF_CLK();
        tc::tc_fence_after();
        const uint32_t dst = tmem_a0 + lane_base + (uint32_t)(sa * 64);
        tc::tmem_st32_u(dst, hi);
        tc::tmem_st32_u(dst + 32, lo);
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(bar_fullA0 + 8 * sa);
        long long t2 = TCF_CLK();
        dbg_pwait += t1 - t0; dbg_pst += t2 - t1; dbg_pgen += t0 - tprev; tprev = t2;
        if (++sa == NA) { sa = 0; phe ^= 1; }
      }
    }
    if (a.dbg && warp == 4 && lane == 0) {
      long long* d = a.dbg + (long long)blockIdx.x * 8;
      d[6] = dbg_pgen + 0 * (dbg_pwait + dbg_pst);
    }
  } else if (warp >= 8) {
    // =========================== epilogue ===========================
    const int quad = warp & 3;
    const int half = (warp - 8) >> 2;        // which of the two warps of this lane quadrant: takes every other 32-column batch
    const int pr = quad * 32 + lane;
    const int pl = pl0 + pr;
    const bool pvalid = pl < a.np;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    float* outsH = outs + half * O * 128;                 // this half's partial sums (FMODE_FWD)
    float* dxaH = dxa + half * a.ecnth * Q * 128;         // this half's hi-group accumulators (FMODE_LOOX)
    // accumulator -> true value: 2^kexp as two exact factors (|kexp| may exceed 127)
    const int kexp = rowexp[pr] - 15 - core_exp;
    const float sc1 = scalbnf(1.f, kexp / 2), sc2 = scalbnf(1.f, kexp - kexp / 2);
    const int fexp_all = kexp + rowexp[128 + pr];
    const float fsc1 = scalbnf(1.f, fexp_all / 2), fsc2 = scalbnf(1.f, fexp_all - fexp_all / 2);
    const float* eH = tabEH + pr;
    float s = 0.f;
    int cur_o = 0;
    // FMODE_LOO: the epilogue tables come from the normalised x while W is defined with the raw x: Whi (a sum against the
    // lo-group table) also undoes the lo group's exponents, Wlo the hi group's (rowexp[1] = first half = hi + lo group)
    float hsc1 = 1.f, hsc2 = 1.f, lsc1 = 1.f, lsc2 = 1.f;
    if (MODE == FMODE_LOO) {
      const int kh = kexp + rowexp[256 + pr], kl = kexp + rowexp[128 + pr] - rowexp[256 + pr];
      hsc1 = scalbnf(1.f, kh / 2); hsc2 = scalbnf(1.f, kh - kh / 2);
      lsc1 = scalbnf(1.f, kl / 2); lsc2 = scalbnf(1.f, kl - kl / 2);
    }
    float WLO[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) WLO[j] = 0.f;
    long long dbg_epi = 0;
    const int nvt = a.ntiles * a.nseg;
    for (int u = 0; u < nvt; ++u) {
      const int t = u / a.nseg;
      const bool accum = (u % a.nseg) != 0;      // a later K segment of the same column tile: add to what is there
      const int mb = DB ? (u & 1) : 0;
      const uint32_t tmem_main = tmem_base + (uint32_t)(mb * BN);
      tc::mbar_wait(bar_accfull0 + 8 * mb, (uint32_t)(DB ? ((u >> 1) & 1) : (u & 1)));
      long long te0 = TCF_CLK();
      tc::tc_fence_after();
      const int n0 = t * BN;
      uint32_t* mypark = park + (warp - 8) * (32 * 32) + lane;
      if (DB) {
        // Park the small accumulator: its columns of this warp go to shared memory as bf16 pairs (it carries the cross
        // terms, weight 2^-11: bf16 keeps 2^-20 of the result) and the accumulator is handed back at once — the next
        // tile's MMAs (into the OTHER main accumulator) run while this epilogue works through the main one.
#pragma unroll 1
        for (int cb = 32 * half; cb < BN; cb += 64) {
          if (n0 + cb >= a.Ncols) break;
          float w[32];
          tc::tmem_ld32(tmem_small + lane_base + (uint32_t)cb, w);
#pragma unroll
          for (int i = 0; i < 16; ++i) mypark[((cb >> 6) * 16 + i) * 32] = tc::pack_bf16x2(w[2 * i], w[2 * i + 1]);
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(bar_smallempty);
      }
#pragma unroll 1
      for (int cb = 0; cb < BN; cb += 32) {
        const int nb = n0 + cb;
        if (nb >= a.Ncols) break;             // Ncols % 32 == 0 (host check): batches are whole or empty
        if (((cb >> 5) & 1) != half) continue;
        float v[32];
        if (DB) {
          tc::tmem_ld32(tmem_main + lane_base + (uint32_t)cb, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t pw = mypark[((cb >> 6) * 16 + i) * 32];
            v[2 * i] = fmaf(__uint_as_float(pw << 16), 1.f / 2048.f, v[2 * i]);
            v[2 * i + 1] = fmaf(__uint_as_float(pw & 0xFFFF0000u), 1.f / 2048.f, v[2 * i + 1]);
          }
        } else {
          float w[32];
          tc::tmem_ld32x2(tmem_main + lane_base + (uint32_t)cb, tmem_small + lane_base + (uint32_t)cb, v, w);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaf(w[i], 1.f / 2048.f, v[i]);
        }
        if (MODE == FMODE_LOOX) {
          // columns nb..nb+31 are a = eh*ELR + el: 32/ELR runs of one eh each; Q is a power of two here
          const int eh0 = nb / a.ELR, lq = 31 - __clz(Q);
          switch (a.ELR) {
            case 16: loox_batch<16>(v, EL, WLO, xh, dxaH, eh0, a.ecnth, Q, lq, pr); break;
            case 8: loox_batch<8>(v, EL, WLO, xh, dxaH, eh0, a.ecnth, Q, lq, pr); break;
            case 4: loox_batch<4>(v, EL, WLO, xh, dxaH, eh0, a.ecnth, Q, lq, pr); break;
            default: loox_batch<2>(v, EL, WLO, xh, dxaH, eh0, a.ecnth, Q, lq, pr); break;
          }
        } else if (MODE == FMODE_LOO) {
          // columns nb..nb+31 are a = eh*ELR + el: 32/ELR runs of one eh each
          float* wrow = a.out + (long long)pl * a.ldc;
          const int eh0 = nb / a.ELR;
          const float* ehp = eH + eh0 * 128;
          switch (a.ELR) {
            case 16:
#pragma unroll
              for (int r = 0; r < 2; ++r) {
                const float th = ehp[r * 128];
                float whi = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) { whi = fmaf(v[r * 16 + j], EL[j], whi); WLO[j] = fmaf(v[r * 16 + j], th, WLO[j]); }
                if (pvalid) wrow[eh0 + r] = accum ? fmaf(whi * hsc1, hsc2, wrow[eh0 + r]) : whi * hsc1 * hsc2;
              }
              break;
            case 8:
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                const float th = ehp[r * 128];
                float whi = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) { whi = fmaf(v[r * 8 + j], EL[j], whi); WLO[j] = fmaf(v[r * 8 + j], th, WLO[j]); }
                if (pvalid) wrow[eh0 + r] = accum ? fmaf(whi * hsc1, hsc2, wrow[eh0 + r]) : whi * hsc1 * hsc2;
              }
              break;
            case 4:
#pragma unroll
              for (int r = 0; r < 8; ++r) {
                const float th = ehp[r * 128];
                float whi = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) { whi = fmaf(v[r * 4 + j], EL[j], whi); WLO[j] = fmaf(v[r * 4 + j], th, WLO[j]); }
                if (pvalid) wrow[eh0 + r] = accum ? fmaf(whi * hsc1, hsc2, wrow[eh0 + r]) : whi * hsc1 * hsc2;
              }
              break;
            default:  // 2
#pragma unroll
              for (int r = 0; r < 16; ++r) {
                const float th = ehp[r * 128];
                const float whi = fmaf(v[r * 2], EL[0], v[r * 2 + 1] * EL[1]);
                WLO[0] = fmaf(v[r * 2], th, WLO[0]);
                WLO[1] = fmaf(v[r * 2 + 1], th, WLO[1]);
                if (pvalid) wrow[eh0 + r] = accum ? fmaf(whi * hsc1, hsc2, wrow[eh0 + r]) : whi * hsc1 * hsc2;
              }
              break;
          }
        } else if (MODE == FMODE_STORE) {
          if (pvalid) {
            float* crow = a.out + (long long)pl * a.ldc + nb;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float4 r4 = make_float4(v[i] * sc1 * sc2, v[i + 1] * sc1 * sc2, v[i + 2] * sc1 * sc2, v[i + 3] * sc1 * sc2);
              if (accum) {
                const float4 o4 = *(const float4*)(crow + i);
                r4.x += o4.x; r4.y += o4.y; r4.z += o4.z; r4.w += o4.w;
              }
              *(float4*)(crow + i) = r4;
            }
          }
        } else {
          if (a.tsave != nullptr) {
            float* st = tstage + (warp - 8) * (32 * 20);
            const int rsub = lane >> 2, c4 = (lane & 3) * 4;
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {               // 16 columns at a time through the [32][20] tile
#pragma unroll
              for (int i = 0; i < 16; i += 4)
                *(float4*)(st + lane * 20 + i) = make_float4(v[16 * hb + i] * sc1 * sc2, v[16 * hb + i + 1] * sc1 * sc2,
                                                             v[16 * hb + i + 2] * sc1 * sc2, v[16 * hb + i + 3] * sc1 * sc2);
              __syncwarp();
              // one store instruction = 8 rows x 64 contiguous bytes
              float* tbase = a.tsave + (pt0 + quad * 32) * (long long)a.Ncols + nb + 16 * hb + c4;
#pragma unroll
              for (int r0 = 0; r0 < 32; r0 += 8) {
                const int r = r0 + rsub;
                if (pl0 + quad * 32 + r < a.np) {
                  float4 r4 = *(const float4*)(st + r * 20 + c4);
                  float4* tp = (float4*)(tbase + (long long)r * a.Ncols);
                  if (accum) {
                    const float4 o4 = *tp;
                    r4.x += o4.x; r4.y += o4.y; r4.z += o4.z; r4.w += o4.w;
                  }
                  *tp = r4;
                }
              }
              __syncwarp();
            }
          }
          // 32 columns of one o (Bn % 32 == 0): b = b0 .. b0+31, KR2[b] = EH[b / ELR] * EL[b % ELR]
          const int o = nb / g.Bn, b0 = nb - o * g.Bn;
          if (o != cur_o) {
            outsH[cur_o * 128 + pr] += s;
            s = 0.f;
            cur_o = o;
          }
          const float* ehp = eH + (b0 / a.ELR) * 128;
          float acc = 0.f;
          switch (a.ELR) {
            case 16:
#pragma unroll
              for (int r = 0; r < 2; ++r) {
                const float eh = ehp[r * 128];
#pragma unroll
                for (int j = 0; j < 16; ++j) acc = fmaf(v[r * 16 + j], eh * EL[j], acc);
              }
              break;
            case 8:
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                const float eh = ehp[r * 128];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc = fmaf(v[r * 8 + j], eh * EL[j], acc);
              }
              break;
            case 4:
#pragma unroll
              for (int r = 0; r < 8; ++r) {
                const float eh = ehp[r * 128];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc = fmaf(v[r * 4 + j], eh * EL[j], acc);
              }
              break;
            default:  // 2
#pragma unroll
              for (int r = 0; r < 16; ++r) {
                const float eh = ehp[r * 128];
                acc = fmaf(v[r * 2], eh * EL[0], acc);
                acc = fmaf(v[r * 2 + 1], eh * EL[1], acc);
              }
              break;
          }
          s += acc;
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(bar_mainempty0 + 8 * mb);
      dbg_epi += TCF_CLK() - te0;
    }
    if (a.dbg && warp == 8 && lane == 0) a.dbg[(long long)blockIdx.x * 8 + 7] = dbg_epi;
    // ---- merge the two halves: the second warp of a quadrant parks its partial state in shared memory, the first adds it
    if (half == 1) {
      if (MODE == FMODE_FWD) outsH[cur_o * 128 + pr] += s;
      if (MODE == FMODE_LOO || MODE == FMODE_LOOX) {
        float* w1 = wlo + (MODE == FMODE_LOOX ? 16 * 128 : 0);
#pragma unroll
        for (int j = 0; j < 16; ++j) w1[j * 128 + pr] = WLO[j];
      }
    }
    asm volatile("bar.sync 3, %0;" ::"n"(32 * F_EPI_WARPS) : "memory");
    if (half == 0 && (MODE == FMODE_LOO || MODE == FMODE_LOOX)) {
      const float* w1 = wlo + (MODE == FMODE_LOOX ? 16 * 128 : 0);
#pragma unroll
      for (int j = 0; j < 16; ++j) WLO[j] += w1[j * 128 + pr];
    }
    if (MODE == FMODE_LOOX && half == 0) {
      const int lq = 31 - __clz(Q);
      const int kb = kexp + rowexp[128 + pr];            // accumulator exponent + all first-half factor exponents
      float* drow = a.out + ((pt0 + pr) * g.n + a.ej0) * (long long)Q;
      // hi group: accumulated on the fly
      for (int t = 0; t < a.ecnth; ++t) {
        const int k = kb - fe[t * 128 + pr];
        const float s1 = scalbnf(1.f, k / 2), s2 = scalbnf(1.f, k - k / 2);
        if (pvalid)
          for (int q = 0; q < Q; ++q)
            drow[t * Q + q] = (dxa[((t * Q + q) << 7) + pr] + dxa[((a.ecnth * Q + t * Q + q) << 7) + pr]) * s1 * s2;
      }
      // lo group: from the Wlo registers (through this thread's column of shared memory, for dynamic indexing)
#pragma unroll
      for (int j = 0; j < 16; ++j) wlo[j * 128 + pr] = WLO[j];
      for (int tl = 0; tl < a.ecntl; ++tl) {
        const int u0 = a.ecnth + tl;
        const int k = kb - fe[u0 * 128 + pr];
        const float s1 = scalbnf(1.f, k / 2), s2 = scalbnf(1.f, k - k / 2);
        const int sh = lq * (a.ecntl - 1 - tl);
        for (int q = 0; q < Q; ++q) {
          float acc = 0.f;
          for (int el = 0; el < a.ELR; ++el) {
            if (((el >> sh) & (Q - 1)) != q) continue;
            float vv = wlo[el * 128 + pr];
            for (int t2 = 0; t2 < a.ecntl; ++t2)
              if (t2 != tl) vv *= xh[(((a.ecnth + t2) * Q + ((el >> (lq * (a.ecntl - 1 - t2))) & (Q - 1))) << 7) + pr];
            acc += vv;
          }
          if (pvalid) drow[u0 * Q + q] = acc * s1 * s2;
        }
      }
    }
    if (MODE == FMODE_LOO && half == 0 && pvalid) {
      float* wrow = a.out + (long long)pl * a.ldc + a.EHE;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < a.ELR) wrow[j] = WLO[j] * lsc1 * lsc2;
    }
    if (MODE == FMODE_FWD && half == 0) {
      outs[cur_o * 128 + pr] += s;
      if (pvalid) {
        float* orow = a.out + (pt0 + pr) * O;
        for (int o = 0; o < O; ++o) orow[o] = (outs[o * 128 + pr] + outs[(O + o) * 128 + pr]) * fsc1 * fsc2;
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, 512);
  if (a.dbg && tid == 0) {   // probes (timing build only): setup and whole-CTA cycles replace the producer wait / store slots
    a.dbg[(long long)blockIdx.x * 8 + 4] = dbg_t_setup - dbg_t_entry;
    a.dbg[(long long)blockIdx.x * 8 + 5] = TCF_CLK() - dbg_t_entry;
    long long* d2 = a.dbg + 4096 * 8 + (long long)blockIdx.x * 4;   // setup phases: x loads | tables | register copies
    d2[0] = dbg_t_loaded - dbg_t_entry; d2[1] = dbg_t_tables - dbg_t_loaded; d2[2] = dbg_t_setup - dbg_t_tables;
  }
}

// ------------------------------------------------------------------------------------------------ host side
struct FastShape {
  int ok;
  int jh0, cnth, cntl, KLR, KHE, withG, Kdim, Ncols;
  int ej0, ecnth, ecntl, EHE, ELR;
};

inline int ilog_pow2(int q) {  // log2(q) if q is a power of two >= 2, else -1
  if (q < 2 || (q & (q - 1))) return -1;
  int l = 0;
  while ((1 << l) < q) ++l;
  return l;
}
// lo-group size for `cnt` factors of size Q: the most factors with Q^cl <= 16, leaving at least one for the hi group
inline int lo_count(int Q, int cnt) {
  const int lq = ilog_pow2(Q);
  if (lq < 0 || cnt < 2) return 0;
  int cl = 4 / lq;
  if (cl > cnt - 1) cl = cnt - 1;
  return cl;
}

inline FastShape fast_shape(const EpsGeom& g, int mode) {
  FastShape s{};
  const int cntA = g.m, cntB = g.n - g.m;
  if (mode == FMODE_FWD) {
    const int cl = lo_count(g.Q, cntA), el = lo_count(g.Q, cntB);
    if (cl < 1 || el < 1) return s;
    s.jh0 = 0; s.cntl = cl; s.cnth = cntA - cl; s.KLR = ipow_host(g.Q, cl); s.KHE = ipow_host(g.Q, s.cnth); s.withG = 0;
    s.Kdim = g.A; s.Ncols = g.N;
    s.ej0 = g.m; s.ecntl = el; s.ecnth = cntB - el; s.ELR = ipow_host(g.Q, el); s.EHE = ipow_host(g.Q, s.ecnth);
    if (g.Bn % 32 != 0) return s;     // a 32-column epilogue batch must not straddle two o
  } else {
    const int cl = lo_count(g.Q, cntB);
    if (cl < 1) return s;
    s.jh0 = g.m; s.cntl = cl; s.cnth = cntB - cl; s.KLR = ipow_host(g.Q, cl); s.KHE = g.O * ipow_host(g.Q, s.cnth); s.withG = 1;
    s.Kdim = g.N; s.Ncols = g.A;
    if (g.A % 32 != 0) return s;
    if (mode == FMODE_LOO || mode == FMODE_LOOX) {   // epilogue groups of the first half
      const int el = lo_count(g.Q, cntA);
      if (el < 1) return s;
      s.ej0 = 0; s.ecntl = el; s.ecnth = cntA - el; s.ELR = ipow_host(g.Q, el); s.EHE = ipow_host(g.Q, s.ecnth);
      if (mode == FMODE_LOOX && s.ecnth > 4) return s;   // the on-the-fly scatter is unrolled for up to 4 hi factors
    }
  }
  if (s.KLR < 2 || s.KLR > 16) return s;
  s.ok = 1;
  return s;
}

// lo-group register tables [32][128], sharing their memory with the staging tiles of the T store [8 warps][32][20]
constexpr size_t REGION_BYTES = ((size_t)F_EPI_WARPS * 32 * 20 > 32 * 128 ? (size_t)F_EPI_WARPS * 32 * 20 : 32 * 128) * 4;
constexpr size_t PARK_BYTES = (size_t)F_EPI_WARPS * 32 * 32 * 4;   // double-buffered mode: the parked small accumulator
inline size_t fast_region_bytes(int mode, int dbuf) {
  if (!dbuf) return REGION_BYTES;
  return mode == FMODE_FWD ? REGION_BYTES + PARK_BYTES : (REGION_BYTES > PARK_BYTES ? REGION_BYTES : PARK_BYTES);
}
inline size_t fast_fixed_smem(const EpsGeom& g, const FastShape& s, int mode, int dbuf = 0) {
  const size_t nk = (size_t)(s.Kdim + FKS - 1) / FKS;
  const size_t nH = nk * (FKS / s.KLR);
  const size_t mfirst = (size_t)s.ecnth + s.ecntl;
  const size_t erows = (mode == FMODE_LOOX) ? mfirst * g.Q + mfirst + 2 * (size_t)s.ecnth * g.Q + 32
                     : (mode == FMODE_LOO ? (size_t)s.EHE + 16 : (mode == FMODE_FWD ? (size_t)s.EHE : 0));
  return 1024 + (nH + erows + (mode == FMODE_FWD ? 2 * (size_t)g.O : 0)) * 128 * 4 + 384 * 4 + fast_region_bytes(mode, dbuf) +
         (2 * F_MAX_BSTAGES + 2 * F_ASTAGES + 6) * 8 + 16;
}
inline size_t fast_stage_bytes(int BN) { return 2 * (size_t)BN * 128; }
inline int fast_bstages(const EpsGeom& g, const FastShape& s, int mode, int BN, int dbuf = 0) {
  const size_t fixed = fast_fixed_smem(g, s, mode, dbuf);
  if (fixed >= F_SMEM_LIMIT) return 0;
  int nb = (int)((F_SMEM_LIMIT - fixed) / fast_stage_bytes(BN));
  if (nb > F_MAX_BSTAGES) nb = F_MAX_BSTAGES;
  if (nb < 2) return 0;
  // setup scratch aliased onto the stages: x, gout, exponents and the level buffer of the table builder
  size_t ymax = (size_t)s.KHE > (size_t)s.EHE ? (size_t)s.KHE : (size_t)s.EHE;
  if (ymax < 16) ymax = 16;
  if ((size_t)(g.n * g.Q + g.O + g.n + 1 + ymax / g.Q + 1) * 128 * 4 > nb * fast_stage_bytes(BN)) return 0;
  return nb;
}
// column-tile width: multiple of 32 (whole epilogue batches), least padded, then widest
inline int fast_bn(const EpsGeom& g, const FastShape& s, int mode) {
  if (const char* e = getenv("DCTN_B200_FAST_BN")) {     // tuning override: "<mode>:<BN>[,<mode>:<BN>...]"
    for (const char* p = e; *p;) {
      int m = 0, bn = 0;
      if (sscanf(p, "%d:%d", &m, &bn) == 2 && m == mode && bn >= 32 && bn <= F_MAX_BN && bn % 32 == 0 && fast_bstages(g, s, mode, bn) != 0) return bn;
      while (*p && *p != ',') ++p;
      if (*p == ',') ++p;
    }
  }
  int best = 0;
  long long best_cost = 0;
  for (int bn = F_MAX_BN; bn >= 64; bn -= 32) {
    if (fast_bstages(g, s, mode, bn) == 0) continue;
    long long cost = (long long)((s.Ncols + bn - 1) / bn) * bn;
    if (bn < 160) cost = cost * 9 / 8;
    // two B stages cannot cover the latency of the bulk copies (measured, config 2 layer 2 input gradient: BN = 160 with 2
    // stages 2.86 ms, MMAs waiting 300-600 cycles per stage for B; BN = 128 with 3 stages 2.59 ms)
    if (fast_bstages(g, s, mode, bn) < 3) cost = cost * 5 / 4;
    if (!best || cost < best_cost) { best = bn; best_cost = cost; }
  }
  return best;
}

// Tile configuration: column-tile width and whether two main accumulators alternate (BN <= 128: 3 BN + 128 columns of
// tensor memory).  Double buffering hides the epilogue behind the next tile's MMAs — single-buffered, the MMAs waited
// 3-7 k cycles per tile for it (12-27 % of the kernel) — and is taken whenever it fits with three B stages.
struct FastCfg { int BN, dbuf, nb; };
inline FastCfg fast_cfg(const EpsGeom& g, const FastShape& s, int mode) {
  FastCfg c{fast_bn(g, s, mode), 0, 0};
  if (!c.BN) return c;
  c.nb = fast_bstages(g, s, mode, c.BN);
  static const int want = [] { const char* e = getenv("DCTN_B200_FAST_DBUF"); return e ? atoi(e) : 1; }();
  if (want) {
    for (int bn = 128; bn >= 64; bn -= 32) {
      if ((s.Ncols + bn - 1) / bn * bn > (s.Ncols + 127) / 128 * 128 + 32 && bn < 128) continue;   // narrower only when it pads less
      const int nb = fast_bstages(g, s, mode, bn, 1);
      if (nb >= 3) { c.BN = bn; c.dbuf = 1; c.nb = nb; break; }
    }
  }
  return c;
}

template <int MODE, int KLR>
int launch_fast(const FastArgs& a, size_t smem, cudaStream_t st) {
  auto k = tc_gemm_fast_kernel<MODE, KLR>;
  DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<(a.np + FBM - 1) / FBM, F_THREADS, smem, st>>>(a);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}
template <int MODE>
int launch_fast_klr(const FastArgs& a, int KLR, size_t smem, cudaStream_t st) {
  switch (KLR) {
    case 2: return launch_fast<MODE, 2>(a, smem, st);
    case 4: return launch_fast<MODE, 4>(a, smem, st);
    case 8: return launch_fast<MODE, 8>(a, smem, st);
    case 16: return launch_fast<MODE, 16>(a, smem, st);
  }
  return dctn_set_error(-2, "register-table GEMM: no instance for a lo group of %d entries", KLR);
}

}  // namespace

// mode: 0 = input-gradient GEMM (dKR1), 1 = forward, 2 = input-gradient GEMM with the fused leave-one-out stage
// (output W[np][ldc], ldc = tcfast_loo_width(g) = hi-group + lo-group entries of the first half)
int tcfast_loo_groups(const EpsGeom& g, int* cnth, int* EH, int* cntl, int* EL) {
  const FastShape s = fast_shape(g, FMODE_LOO);
  if (!s.ok) return 0;
  *cnth = s.ecnth; *EH = s.EHE; *cntl = s.ecntl; *EL = s.ELR;
  return s.EHE + s.ELR;
}
bool tcfast_supported(const EpsGeom& g, int mode) {
  const FastShape s = fast_shape(g, mode);
  return s.ok && fast_cfg(g, s, mode).BN != 0;
}

size_t tcfast_packed_floats(const EpsGeom& g, int mode) {
  const FastShape s = fast_shape(g, mode);
  if (!s.ok) return 0;
  const int BN = fast_cfg(g, s, mode).BN;
  if (!BN) return 0;
  const long long ntiles = (s.Ncols + BN - 1) / BN, nk = (s.Kdim + FKS - 1) / FKS;
  return (size_t)(ntiles * nk * 2 * BN * 32);
}

int tcfast_pack(const EpsGeom& g, int mode, const float* core, float* dst, const uint32_t* absmax, cudaStream_t st) {
  const FastShape s = fast_shape(g, mode);
  const int BN = fast_cfg(g, s, mode).BN;
  const int ntiles = (s.Ncols + BN - 1) / BN, nk = (s.Kdim + FKS - 1) / FKS;
  const long long total = (long long)ntiles * nk * BN * 8;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  fast_pack_kernel<<<blocks, 256, 0, st>>>(core, dst, g, mode, BN, s.Ncols, s.Kdim, ntiles, nk, absmax);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}

int tcfast_gemm(const EpsGeom& g, int mode, const float* x, const float* gout, const float* packed, const uint32_t* absmax,
                long long p0, int np, float* out, long long ldc, float* tsave, cudaStream_t st) {
  const FastShape s = fast_shape(g, mode);
  const FastCfg cfg = s.ok ? fast_cfg(g, s, mode) : FastCfg{0, 0, 0};
  const int BN = cfg.BN;
  if (!BN) return dctn_set_error(-2, "register-table GEMM does not support this shape");
  FastArgs a{};
  a.g = g; a.x = x; a.gout = gout; a.p0 = p0; a.np = np;
  a.jh0 = s.jh0; a.cnth = s.cnth; a.cntl = s.cntl; a.KHE = s.KHE; a.withG = s.withG; a.Kdim = s.Kdim;
  a.ej0 = s.ej0; a.ecnth = s.ecnth; a.ecntl = s.ecntl; a.EHE = s.EHE; a.ELR = s.ELR;
  a.Ncols = s.Ncols; a.ntiles = (s.Ncols + BN - 1) / BN; a.nk = (s.Kdim + FKS - 1) / FKS; a.BN = BN;
  a.kseg = tc::seg_stages(a.nk); a.nseg = (a.nk + a.kseg - 1) / a.kseg;
  a.bstages = cfg.nb;
  a.dbuf = cfg.dbuf;
  a.region_floats = (int)(fast_region_bytes(mode, cfg.dbuf) / 4);
  a.packed = packed; a.core_absmax = absmax; a.out = out; a.ldc = ldc; a.tsave = tsave;
  a.dbg = nullptr;
#ifdef DCTN_TCG_TIMING   // cycle probes: timing builds only (allocates, synchronises, not thread-safe)
  static long long* dbg_buf = nullptr;
  const int ncta = (np + FBM - 1) / FBM;
  if (getenv("DCTN_TCG_DEBUG") && ncta <= 4096) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 4096 * 12 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 4096 * 12 * sizeof(long long), st);
    a.dbg = dbg_buf;
  }
#endif
  const size_t smem = fast_fixed_smem(g, s, mode, cfg.dbuf) + a.bstages * fast_stage_bytes(BN);
  int rc = (mode == FMODE_FWD) ? launch_fast_klr<FMODE_FWD>(a, s.KLR, smem, st)
         : (mode == FMODE_LOO) ? launch_fast_klr<FMODE_LOO>(a, s.KLR, smem, st)
         : (mode == FMODE_LOOX) ? launch_fast_klr<FMODE_LOOX>(a, s.KLR, smem, st) : launch_fast_klr<FMODE_STORE>(a, s.KLR, smem, st);
#ifdef DCTN_TCG_TIMING
  if (a.dbg && rc == 0) {
    static long long host[4096 * 12];
    cudaStreamSynchronize(st);
    cudaMemcpy(host, dbg_buf, (size_t)4096 * 12 * sizeof(long long), cudaMemcpyDeviceToHost);
    double sum[8] = {0}, ph[3] = {0};
    for (int c = 0; c < ncta; ++c) for (int k = 0; k < 8; ++k) sum[k] += (double)host[c * 8 + k];
    for (int c = 0; c < ncta; ++c) for (int k = 0; k < 3; ++k) ph[k] += (double)host[4096 * 8 + c * 4 + k];
    fprintf(stderr, "[tcfast dbg] setup phases: x loads %.0f tables %.0f register copies %.0f\n", ph[0] / ncta, ph[1] / ncta, ph[2] / ncta);
    const double nst = (double)a.ntiles * a.nk;
    fprintf(stderr, "[tcfast dbg] mode=%d KLR=%d BN=%d NB=%d ntiles=%d nk=%d per-stage cycles: mma waitA %.0f waitB %.0f waitAcc(per tile) %.0f total %.0f | "
            "CTA setup %.0f whole %.0f | producer gen %.0f | epilogue/tile %.0f\n", mode, s.KLR, BN, a.bstages, a.ntiles, a.nk,
            sum[0] / ncta / nst, sum[1] / ncta / nst, sum[2] / ncta / a.ntiles, sum[3] / ncta / nst,
            sum[4] / ncta, sum[5] / ncta, sum[6] / ncta / nst, sum[7] / ncta / a.ntiles);
  }
#endif
  return rc;
}
