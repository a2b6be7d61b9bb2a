// tcgen05 GEMMs of the EPS forward and input-gradient passes (float32 in/out, 3xTF32 or 1xTF32 arithmetic).
//
//   C[p][c] = sum_k Gen[p][k] * Bop[k][c]            p: 128 patches per CTA (TMEM lanes), c: BN columns per tile
//
//   * Gen (Khatri-Rao half, optionally times gout) is GENERATED per stage by 4 producer warps, one patch row per
//     thread, from two-level tables built once per CTA, straight into the K-major SWIZZLE_128B layout, already
//     split into TF32 hi / lo parts;
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
#include "common.cuh"
#include "eps_kernels.h"
#include "tc_common.cuh"

namespace {

constexpr int GBM = 128;
constexpr int GBK = 32;
constexpr int GSTAGES = 2;
constexpr int G_THREADS = 384;  // warp 0: bulk copies, warp 1: MMA, warp 2: TMEM alloc, warps 4-7: producers, warps 8-11: epilogue
enum { MODE_STORE = 0, MODE_FWD = 1, MODE_DKR2 = 2 };
constexpr size_t TCG_SMEM_LIMIT = 227 * 1024;

struct TcGemmArgs {
  EpsGeom g;
  const float* x;
  const float* gout;
  long long p0;  // first patch handled by this launch
  int np;        // number of patches
  int jh0, cnth, KH, cntl, KLb, KL, Kdim, withG;  // generated operand (see GenGemmArgs in eps_ffma.cu)
  int Ncols, ntiles, nk;
  const float* packed;  // [ntiles][nk][2][BN*32]
  int passes;
  float* out;           // MODE_STORE: [np][ldc]; MODE_FWD: out[P][O] (absolute patches); MODE_DKR2: [np][Bn]
  long long ldc;
};

// ------------------------------------------------------------------------------------------------ core packing
// dst[((tile*nk + kc)*2 + part)*BN*32 + swizzled(row rr, k)] = part(core element (c = tile*BN + rr, k = kc*32 + ..))
//   MODE_STORE: element(c, k) = core[c*N + k]           (c = a, k = n)
//   MODE_DKR2 : element(c, k) = core[k*N + c]           (c = n, k = a)
//   MODE_FWD  : element(c, k) = core[(k*Bn + b)*O + o]  (c = o*Bn + b, k = a)
__global__ void pack_core_kernel(const float* __restrict__ core, float* __restrict__ dst, EpsGeom g, int mode, int BN,
                                 int Ncols, int Kdim, int ntiles, int nk, int passes) {
  const long long total = (long long)ntiles * nk * BN * 8;  // one thread per 16-byte chunk (4 consecutive k)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c16 = (int)(i & 7);
    long long r = i >> 3;
    const int rr = (int)(r % BN);
    r /= BN;
    const int kc = (int)(r % nk);
    const int tile = (int)(r / nk);
    const int c = tile * BN + rr;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < Ncols) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = kc * 32 + c16 * 4 + u;
        if (k < Kdim) {
          long long idx;
          if (mode == MODE_STORE) idx = (long long)c * g.N + k;
          else if (mode == MODE_DKR2) idx = (long long)k * g.N + c;
          else {
            const int o = c / g.Bn, b = c - o * g.Bn;
            idx = ((long long)k * g.Bn + b) * g.O + o;
          }
          v[u] = __ldg(&core[idx]);
        }
      }
    }
    float4 hi, lo;
    tc::split_tf32(v[0], hi.x, lo.x);
    tc::split_tf32(v[1], hi.y, lo.y);
    tc::split_tf32(v[2], hi.z, lo.z);
    tc::split_tf32(v[3], hi.w, lo.w);
    float* tile_base = dst + ((long long)(tile * nk + kc) * 2) * BN * 32;
    const int off = rr * 32 + ((c16 ^ (rr & 7)) << 2);  // in floats
    *(float4*)(tile_base + off) = hi;
    if (passes == 3) *(float4*)(tile_base + BN * 32 + off) = lo;
  }
}

// ------------------------------------------------------------------------------------------------ the GEMM
template <int BN>
struct GSmem {
  static constexpr uint32_t A_BYTES = GBM * GBK * 4;
  static constexpr uint32_t B_BYTES = BN * GBK * 4;
  static constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr uint32_t OFF_A_HI = 0, OFF_A_LO = A_BYTES, OFF_B_HI = 2 * A_BYTES, OFF_B_LO = 2 * A_BYTES + B_BYTES;
};

template <int BN, int MODE>
__global__ void __launch_bounds__(G_THREADS, 1) tc_gemm_kernel(const __grid_constant__ TcGemmArgs a) {
  using SM = GSmem<BN>;
  extern __shared__ unsigned char smem_dyn[];
  const EpsGeom& g = a.g;
  const int Q = g.Q, O = g.O;
  unsigned char* base = smem_dyn + ((1024u - (tc::smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* stages = base;
  float* tabKH = (float*)(base + GSTAGES * SM::STAGE_BYTES);  // [KH][128]
  float* tabKL = tabKH + a.KH * 128;                          // [KL][128]
  float* tabE = tabKL + a.KL * 128;                           // MODE_FWD: [BH + BL][128]; MODE_DKR2: gout [O][128]
  const int nE = (MODE == MODE_FWD) ? (g.BH + g.BL) : (MODE == MODE_DKR2 ? O : 0);
  float* outs = tabE + nE * 128;                              // MODE_FWD: [O][128]
  uint64_t* bars = (uint64_t*)(outs + ((MODE == MODE_FWD) ? O * 128 : 0));
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * GSTAGES + 2);
  const uint32_t bar_full0 = tc::smem_u32(bars), bar_empty0 = bar_full0 + 8 * GSTAGES;
  const uint32_t bar_accfull = bar_full0 + 16 * GSTAGES, bar_accempty = bar_accfull + 8;
  // setup-only scratch aliased onto the (not yet used) operand stages: x [n*Q][128] and gout [O][128]
  float* xs = (float*)stages;
  float* gsx = xs + g.n * Q * 128;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pl0 = blockIdx.x * GBM;                 // first patch of this CTA, relative to the launch
  const long long pt0 = a.p0 + pl0;                 // absolute
  constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256 : 512;

  // ---------------- setup: barriers, TMEM, tables
  if (tid == 0) {
    for (int s = 0; s < GSTAGES; ++s) {
      tc::mbar_init(bar_full0 + 8 * s, 4 + 1);   // 4 producer warps + the expect_tx arrive of the copy warp
      tc::mbar_init(bar_empty0 + 8 * s, 1);      // tcgen05.commit
    }
    tc::mbar_init(bar_accfull, 1);
    tc::mbar_init(bar_accempty, 4);              // 4 epilogue warps
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(tc::smem_u32(tmem_slot), TMEM_COLS);
  {
    const int NX = g.n * Q;
    for (int idx = tid; idx < NX * 128; idx += G_THREADS) {
      const int pr = idx & 127, jq = idx >> 7;
      const long long p = pt0 + pr;
      float v = 0.f;
      if (p < g.P) v = __ldg(&a.x[patch_origin(g, p) + g.foff[jq / Q] + jq % Q]);
      xs[jq * 128 + pr] = v;
    }
    if (a.withG || MODE == MODE_DKR2) {
      for (int idx = tid; idx < O * 128; idx += G_THREADS) {
        const int pr = idx & 127, o = idx >> 7;
        const long long p = pt0 + pr;
        gsx[o * 128 + pr] = (p < g.P) ? __ldg(&a.gout[p * O + o]) : 0.f;
      }
    }
  }
  __syncthreads();
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
    for (int idx = tid; idx < a.KH * 128; idx += G_THREADS) tabKH[idx] = kr_entry(a.jh0, a.cnth, idx >> 7, idx & 127);
    for (int idx = tid; idx < a.KL * 128; idx += G_THREADS) {
      const int pr = idx & 127, eo = idx >> 7;
      int e = eo;
      float gv = 1.f;
      if (a.withG) {
        e = eo / O;
        gv = gsx[(eo - e * O) * 128 + pr];
      }
      tabKL[idx] = gv * kr_entry(a.jh0 + a.cnth, a.cntl, e, pr);
    }
    if (MODE == MODE_FWD) {
      for (int idx = tid; idx < g.BH * 128; idx += G_THREADS) tabE[idx] = kr_entry(g.m, g.b_nh, idx >> 7, idx & 127);
      for (int idx = tid; idx < g.BL * 128; idx += G_THREADS)
        tabE[g.BH * 128 + idx] = kr_entry(g.m + g.b_nh, g.b_nl, idx >> 7, idx & 127);
      for (int idx = tid; idx < O * 128; idx += G_THREADS) outs[idx] = 0.f;
    }
    if (MODE == MODE_DKR2)
      for (int idx = tid; idx < O * 128; idx += G_THREADS) tabE[idx] = gsx[idx];
  }
  tc::tc_fence_before();
  __syncthreads();  // tables complete, xs/gsx scratch (aliasing the stages) dead from here on
  tc::tc_fence_after();
  const uint32_t tmem_main = *tmem_slot;
  const uint32_t tmem_small = tmem_main + BN;
  const int total_it = a.ntiles * a.nk;

  if (warp == 0) {
    // =========================== bulk-copy issuer (B operand) ===========================
    if (lane == 0) {
      const uint32_t bytes = SM::B_BYTES * (a.passes == 3 ? 2u : 1u);
      for (int i = 0; i < total_it; ++i) {
        const int s = i % GSTAGES;
        const uint32_t it = (uint32_t)(i / GSTAGES);
        tc::mbar_wait(bar_empty0 + 8 * s, (it & 1) ^ 1);
        const uint32_t sb = tc::smem_u32(stages + s * SM::STAGE_BYTES);
        const float* src = a.packed + (long long)i * 2 * BN * 32;
        tc::mbar_arrive_expect_tx(bar_full0 + 8 * s, bytes);
        tc::bulk_g2s(sb + SM::OFF_B_HI, src, SM::B_BYTES, bar_full0 + 8 * s);
        if (a.passes == 3) tc::bulk_g2s(sb + SM::OFF_B_LO, src + BN * 32, SM::B_BYTES, bar_full0 + 8 * s);
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = tc::make_idesc_tf32(GBM, BN);
    int i = 0;
    for (int t = 0; t < a.ntiles; ++t) {
      if (t > 0) tc::mbar_wait(bar_accempty, (uint32_t)((t - 1) & 1));  // epilogue has drained the previous tile
      tc::tc_fence_after();
      for (int kc = 0; kc < a.nk; ++kc, ++i) {
        const int s = i % GSTAGES;
        const uint32_t it = (uint32_t)(i / GSTAGES);
        tc::mbar_wait(bar_full0 + 8 * s, it & 1);
        tc::tc_fence_after();
        if (lane == 0) {
          const uint32_t sb = tc::smem_u32(stages + s * SM::STAGE_BYTES);
          const uint64_t da_hi = tc::make_sw128_kmajor_desc(sb + SM::OFF_A_HI);
          const uint64_t da_lo = tc::make_sw128_kmajor_desc(sb + SM::OFF_A_LO);
          const uint64_t db_hi = tc::make_sw128_kmajor_desc(sb + SM::OFF_B_HI);
          const uint64_t db_lo = tc::make_sw128_kmajor_desc(sb + SM::OFF_B_LO);
#pragma unroll
          for (int k = 0; k < GBK / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 2);
            const uint32_t first = (kc == 0 && k == 0) ? 0u : 1u;
            tc::umma_tf32(tmem_main, da_hi + adv, db_hi + adv, idesc, first);
            if (a.passes == 3) {
              tc::umma_tf32(tmem_small, da_hi + adv, db_lo + adv, idesc, first);
              tc::umma_tf32(tmem_small, da_lo + adv, db_hi + adv, idesc, 1u);
            }
          }
          tc::umma_commit(bar_empty0 + 8 * s);
          if (kc == a.nk - 1) tc::umma_commit(bar_accfull);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // =========================== A producers: one patch row per thread ===========================
    const int pr = (warp - 4) * 32 + lane;
    const float* th = tabKH + pr;
    const float* tl = tabKL + pr;
    const uint32_t rowoff = (uint32_t)(pr * 128);
    const int sw = pr & 7;
    int i = 0;
    for (int t = 0; t < a.ntiles; ++t) {
      for (int kc = 0; kc < a.nk; ++kc, ++i) {
        const int s = i % GSTAGES;
        const uint32_t it = (uint32_t)(i / GSTAGES);
        const int k0 = kc * GBK;
        int kh = k0 / a.KL;
        int kl = k0 - kh * a.KL;
        float v[GBK];
#pragma unroll
        for (int j = 0; j < GBK; ++j) {
          const bool ok = (k0 + j) < a.Kdim;
          const float hv = th[(ok ? kh : 0) * 128];
          const float lv = tl[(ok ? kl : 0) * 128];
          v[j] = ok ? hv * lv : 0.f;
          if (++kl == a.KL) { kl = 0; ++kh; }
        }
        tc::mbar_wait(bar_empty0 + 8 * s, (it & 1) ^ 1);
        unsigned char* st = stages + s * SM::STAGE_BYTES + rowoff;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 hi, lo;
          tc::split_tf32(v[4 * c + 0], hi.x, lo.x);
          tc::split_tf32(v[4 * c + 1], hi.y, lo.y);
          tc::split_tf32(v[4 * c + 2], hi.z, lo.z);
          tc::split_tf32(v[4 * c + 3], hi.w, lo.w);
          const uint32_t off = (uint32_t)((c ^ sw) << 4);
          *(float4*)(st + SM::OFF_A_HI + off) = hi;
          if (a.passes == 3) *(float4*)(st + SM::OFF_A_LO + off) = lo;
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(bar_full0 + 8 * s);
      }
    }
  } else if (warp >= 8) {
    // =========================== epilogue ===========================
    const int quad = warp & 3;
    const int pr = quad * 32 + lane;
    const int pl = pl0 + pr;                 // relative to the launch
    const bool pvalid = pl < a.np;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    // running state of the fused reductions (all of it warp-uniform except s)
    float s = 0.f;
    int fo = 0, fb = 0, fbh = 0, fbl = 0;    // MODE_FWD: current o, b, b / BL, b % BL
    int do_ = 0, db = 0;                     // MODE_DKR2: current o and b
    const float* eH = tabE + pr;
    const float* eL = tabE + g.BH * 128 + pr;
    for (int t = 0; t < a.ntiles; ++t) {
      tc::mbar_wait(bar_accfull, (uint32_t)(t & 1));
      tc::tc_fence_after();
      const int n0 = t * BN;
      if (MODE == MODE_FWD) {
        fo = n0 / g.Bn; fb = n0 - fo * g.Bn; fbh = fb / g.BL; fbl = fb - fbh * g.BL;
      } else if (MODE == MODE_DKR2) {
        do_ = 0; db = n0 / O;  // BN % O == 0 (checked on the host)
      }
#pragma unroll 1
      for (int cb = 0; cb < BN; cb += 32) {
        float v[32];
        tc::tmem_ld32(tmem_main + lane_base + (uint32_t)cb, v);
        if (a.passes == 3) {
          float w[32];
          tc::tmem_ld32(tmem_small + lane_base + (uint32_t)cb, w);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += w[i];
        }
        if (MODE == MODE_STORE) {
          if (pvalid) {
            float* crow = a.out + (long long)pl * a.ldc;
            const int nb = n0 + cb;
            if (cb + 32 <= BN && nb + 32 <= a.Ncols && (a.ldc & 3) == 0) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) *(float4*)(crow + nb + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (cb + i < BN && nb + i < a.Ncols) crow[nb + i] = v[i];
            }
          }
        } else if (MODE == MODE_FWD) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (cb + i < BN && n0 + cb + i < a.Ncols) {
              s = fmaf(v[i], eH[fbh * 128] * eL[fbl * 128], s);
              ++fb;
              if (++fbl == g.BL) { fbl = 0; ++fbh; }
              if (fb == g.Bn) {
                outs[fo * 128 + pr] += s;
                s = 0.f; fb = 0; fbh = 0; fbl = 0; ++fo;
              }
            }
          }
        } else {  // MODE_DKR2
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (cb + i < BN && n0 + cb + i < a.Ncols) {
              s = fmaf(v[i], tabE[do_ * 128 + pr], s);
              if (++do_ == O) {
                if (pvalid) a.out[(long long)pl * g.Bn + db] = s;
                s = 0.f; do_ = 0; ++db;
              }
            }
          }
        }
      }
      if (MODE == MODE_FWD) {  // flush the partial sum of this tile (the next tile recomputes its position)
        if (fo < O) outs[fo * 128 + pr] += s;
        s = 0.f;
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(bar_accempty);
    }
    if (MODE == MODE_FWD && pvalid) {
      float* orow = a.out + (pt0 + pr) * O;
      for (int o = 0; o < O; ++o) orow[o] = outs[o * 128 + pr];
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_main, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ host side
struct GemmShape {
  int jh0, cnth, KH, cntl, KLb, KL, Kdim, withG, Ncols;
};

inline GemmShape shape_for(const EpsGeom& g, int mode) {
  GemmShape s{};
  if (mode == MODE_STORE) {  // Gen = KR2 x gout over k = (b, o); columns = a
    s.jh0 = g.m; s.cnth = g.b_nh; s.KH = g.BH; s.cntl = g.b_nl; s.KLb = g.BL; s.KL = g.BL * g.O; s.Kdim = g.N;
    s.withG = 1; s.Ncols = g.A;
  } else {                   // Gen = KR1 over k = a; columns = n
    s.jh0 = 0; s.cnth = g.a_nh; s.KH = g.AH; s.cntl = g.a_nl; s.KLb = g.AL; s.KL = g.AL; s.Kdim = g.A;
    s.withG = 0; s.Ncols = g.N;
  }
  return s;
}

inline size_t gemm_smem(const EpsGeom& g, const GemmShape& s, int mode, int BN) {
  const int nE = (mode == MODE_FWD) ? (g.BH + g.BL) : (mode == MODE_DKR2 ? g.O : 0);
  size_t stage = 2 * (size_t)GBM * GBK * 4 + 2 * (size_t)BN * GBK * 4;
  size_t b = 1024 + GSTAGES * stage + (size_t)(s.KH + s.KL + nE + (mode == MODE_FWD ? g.O : 0)) * 128 * 4 + (2 * GSTAGES + 2) * 8 + 16;
  return b;
}

// setup scratch (x and gout of 128 patches) is aliased onto the operand stages: it must fit there
inline bool scratch_fits(const EpsGeom& g, int BN) {
  size_t stage = 2 * (size_t)GBM * GBK * 4 + 2 * (size_t)BN * GBK * 4;
  return (size_t)(g.n * g.Q + g.O) * 128 * 4 <= GSTAGES * stage;
}

// pick the column-tile width: fits shared memory, satisfies the epilogue's alignment, least padding, then widest
inline int pick_bn(const EpsGeom& g, int mode) {
  const GemmShape s = shape_for(g, mode);
  int best = 0;
  long long best_pad = 0;
  const int cands[4] = {256, 240, 192, 128};  // 240 = 16*15 serves Q_out = 3, 5, 6, 10, 12, 15, 24 in MODE_DKR2
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    if (mode == MODE_DKR2 && bn % g.O != 0) continue;
    if (gemm_smem(g, s, mode, bn) > TCG_SMEM_LIMIT || !scratch_fits(g, bn)) continue;
    long long pad = (long long)((s.Ncols + bn - 1) / bn) * bn;
    if (bn == 128) pad = pad * 5 / 4;  // N=128 MMAs are shared-memory-bandwidth bound: count them as 25% more expensive
    if (!best || pad < best_pad) { best = bn; best_pad = pad; }
  }
  return best;
}

inline size_t packed_floats(const EpsGeom& g, int mode, int BN) {
  const GemmShape s = shape_for(g, mode);
  long long ntiles = (s.Ncols + BN - 1) / BN, nk = (s.Kdim + GBK - 1) / GBK;
  return (size_t)(ntiles * nk * 2 * BN * 32);
}

inline long long dx_patch_chunk(const EpsGeom& g) {
  long long target = 96ll << 20;
  long long pc = target / (((long long)g.A + g.Bn) * 4);
  if (pc < 4096) pc = 4096;
  pc = (pc / 128) * 128;
  if (pc > g.P) pc = g.P;
  return pc;
}

template <int BN, int MODE>
int launch_gemm_inst(const TcGemmArgs& a, size_t smem, cudaStream_t st) {
  auto k = tc_gemm_kernel<BN, MODE>;
  DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<(a.np + GBM - 1) / GBM, G_THREADS, smem, st>>>(a);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}

int run_gemm(const EpsGeom& g, int mode, int BN, const float* x, const float* gout, const float* packed, long long p0,
             int np, float* out, long long ldc, int passes, cudaStream_t st) {
  const GemmShape s = shape_for(g, mode);
  TcGemmArgs a{};
  a.g = g; a.x = x; a.gout = gout; a.p0 = p0; a.np = np;
  a.jh0 = s.jh0; a.cnth = s.cnth; a.KH = s.KH; a.cntl = s.cntl; a.KLb = s.KLb; a.KL = s.KL; a.Kdim = s.Kdim; a.withG = s.withG;
  a.Ncols = s.Ncols; a.ntiles = (s.Ncols + BN - 1) / BN; a.nk = (s.Kdim + GBK - 1) / GBK;
  a.packed = packed; a.passes = passes; a.out = out; a.ldc = ldc;
  const size_t smem = gemm_smem(g, s, mode, BN);
#define DCTN_GEMM_CASE(bn, md) \
  if (BN == bn && mode == md) return launch_gemm_inst<bn, md>(a, smem, st);
  DCTN_GEMM_CASE(256, MODE_STORE) DCTN_GEMM_CASE(240, MODE_STORE) DCTN_GEMM_CASE(192, MODE_STORE) DCTN_GEMM_CASE(128, MODE_STORE)
  DCTN_GEMM_CASE(256, MODE_FWD) DCTN_GEMM_CASE(240, MODE_FWD) DCTN_GEMM_CASE(192, MODE_FWD) DCTN_GEMM_CASE(128, MODE_FWD)
  DCTN_GEMM_CASE(256, MODE_DKR2) DCTN_GEMM_CASE(240, MODE_DKR2) DCTN_GEMM_CASE(192, MODE_DKR2) DCTN_GEMM_CASE(128, MODE_DKR2)
#undef DCTN_GEMM_CASE
  return dctn_set_error(-2, "tcgen05 GEMM: no kernel instance for BN=%d mode=%d", BN, mode);
}

int run_pack(const EpsGeom& g, int mode, int BN, const float* core, float* dst, int passes, cudaStream_t st) {
  const GemmShape s = shape_for(g, mode);
  const int ntiles = (s.Ncols + BN - 1) / BN, nk = (s.Kdim + GBK - 1) / GBK;
  long long total = (long long)ntiles * nk * BN * 8;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_core_kernel<<<blocks, 256, 0, st>>>(core, dst, g, mode, BN, s.Ncols, s.Kdim, ntiles, nk, passes);
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

bool tcg_supported(const EpsGeom& g, int kind) {
  if (!common_ok(g)) return false;
  if (kind == 0) return pick_bn(g, MODE_FWD) != 0;
  if (kind == 2) return (g.n - g.m) > 0 && pick_bn(g, MODE_STORE) != 0 && pick_bn(g, MODE_DKR2) != 0;
  return false;
}

size_t tcg_workspace_bytes(const EpsGeom& g, int kind) {
  if (kind == 0) return packed_floats(g, MODE_FWD, pick_bn(g, MODE_FWD)) * 4 + 256;
  if (kind == 2) {
    const long long pc = dx_patch_chunk(g);
    size_t f = packed_floats(g, MODE_STORE, pick_bn(g, MODE_STORE)) + packed_floats(g, MODE_DKR2, pick_bn(g, MODE_DKR2)) +
               (size_t)pc * ((size_t)g.A + g.Bn) + (size_t)g.P * g.n * g.Q;
    return f * 4 + 1024;
  }
  return 0;
}

int tc_forward(const EpsGeom& g, const float* x, const float* core, float* out, void* ws, int passes, cudaStream_t st) {
  const int BN = pick_bn(g, MODE_FWD);
  if (!BN) return dctn_set_error(-2, "tcgen05 forward kernel does not support this shape");
  float* packed = (float*)ws;
  int rc = run_pack(g, MODE_FWD, BN, core, packed, passes, st);
  if (rc) return rc;
  return run_gemm(g, MODE_FWD, BN, x, nullptr, packed, 0, (int)g.P, out, 0, passes, st);
}

int tc_backward_input(const EpsGeom& g, const float* x, const float* core, const float* gout, float* dx, void* ws,
                      int passes, cudaStream_t st) {
  const int BN1 = pick_bn(g, MODE_STORE), BN2 = pick_bn(g, MODE_DKR2);
  if (!BN1 || !BN2) return dctn_set_error(-2, "tcgen05 input-gradient kernels do not support this shape");
  const long long pc = dx_patch_chunk(g);
  float* packed1 = (float*)ws;
  float* packed2 = packed1 + ((packed_floats(g, MODE_STORE, BN1) + 63) & ~(size_t)63);
  float* dkr1 = packed2 + ((packed_floats(g, MODE_DKR2, BN2) + 63) & ~(size_t)63);
  float* dkr2 = dkr1 + (size_t)pc * g.A;
  float* dxp = dkr2 + (size_t)pc * g.Bn;
  int rc;
  if ((rc = run_pack(g, MODE_STORE, BN1, core, packed1, passes, st))) return rc;
  if ((rc = run_pack(g, MODE_DKR2, BN2, core, packed2, passes, st))) return rc;
  for (long long p0 = 0; p0 < g.P; p0 += pc) {
    const int np = (int)((g.P - p0 < pc) ? (g.P - p0) : pc);
    if ((rc = run_gemm(g, MODE_STORE, BN1, x, gout, packed1, p0, np, dkr1, g.A, passes, st))) return rc;
    if ((rc = launch_loo<float>(g, x, dkr1, p0, np, 0, dxp, st))) return rc;
    if ((rc = run_gemm(g, MODE_DKR2, BN2, x, gout, packed2, p0, np, dkr2, g.Bn, passes, st))) return rc;
    if ((rc = launch_loo<float>(g, x, dkr2, p0, np, 1, dxp, st))) return rc;
  }
  return launch_gather_dx<float>(g, dxp, dx, st);
}
