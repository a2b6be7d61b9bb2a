// tcgen05 (5th-gen tensor core) kernel family for the EPS contraction, float32 only.
//
// fp32 accuracy on TF32 tensor cores: every fp32 operand v is split on chip into hi (top 19 bits, exactly a
// TF32 number) and lo = v - hi; a product a*b is issued as three MMAs  a_hi*b_hi + a_hi*b_lo + a_lo*b_hi
// accumulated in the same fp32 TMEM accumulator ("3xTF32", error ~2^-21 per product).  passes == 1 issues
// only a_hi*b_hi (plain TF32, rel. err ~1e-3, opt-in variant DCTN_VARIANT_TC1).
//
// Core gradient (dcore[a][n] = sum_p KR1[p][a] * KR2[p][b(n)] * gout[p][o(n)]), reduction over ALL patches:
//   * CTA tile 128 (a) x 256 (n), one TMEM accumulator (256 columns), K = patches in chunks of 32;
//   * BOTH operands are generated in shared memory by 8 producer warps from two-level Khatri-Rao tables
//     (common.cuh) directly in the K-major SWIZZLE_128B layout tcgen05.mma reads — nothing Q^m-sized
//     ever comes from HBM; only x (n*Q floats per patch) and gout (O floats per patch) are read;
//   * warp 0 issues the MMAs (one elected thread), tcgen05.commit frees the stage / publishes the accumulator;
//   * the tensor core's fp32 accumulator ROUNDS TOWARD ZERO on every MMA (measured: 768 accumulate steps
//     shrink the result by 1.5e-5), so the accumulation chain is kept short: the K loop is cut into segments
//     of SEG_CHUNKS chunks that alternate between two TMEM accumulators (2 x 256 columns = all of TMEM); while
//     the MMAs of segment i+1 run, the producer warps drain segment i with tcgen05.ld and add it into fp32
//     REGISTERS (128 per thread, round-to-nearest) — "promotion", ~2e-6 residual bias independent of K;
//   * split-K: one CTA per (tile, patch range), ~one wave of 148 CTAs; partial tiles go to the workspace and
//     a second kernel sums them in a fixed order (deterministic).
#include "common.cuh"
#include "eps_kernels.h"
#include "tc_common.cuh"

namespace {

constexpr int BM = 128;          // MMA M (rows of the accumulator = TMEM lanes)
constexpr int BK = 32;           // K elements per stage (128 bytes per row)
constexpr int STAGES = 2;
constexpr int NPROD_WARPS = 8;   // producer warps (warps 1..8); warp 0 issues MMAs
constexpr int NTHREADS_TC = 32 * (1 + NPROD_WARPS);
constexpr int SEG_CHUNKS = 8;    // chunks (of BK patches) accumulated in TMEM before promotion to registers

struct TcDcoreArgs {
  EpsGeom g;
  const float* x;
  const float* gout;
  float* part;          // [splits][A][N]
  long long per_split;  // multiple of BK
  int passes;           // 3 (fp32-accurate) or 1
};

__device__ __forceinline__ void producer_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * NPROD_WARPS) : "memory"); }

// packed digits (one byte each, digit 0 = slowest) of entry e of a group with cnt <= 4 factors
__device__ __forceinline__ uint32_t pack_digits(int e, int cnt, int Q) {
  uint32_t packed = 0;
  for (int t = cnt - 1; t >= 0; --t) {
    packed |= (uint32_t)(e % Q) << (8 * t);
    e /= Q;
  }
  return packed;
}

template <int BN>
struct DcoreSmem {
  static constexpr uint32_t A_BYTES = BM * BK * 4;   // 16 KB
  static constexpr uint32_t B_BYTES = BN * BK * 4;   // 32 KB for BN = 256
  static constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr uint32_t OFF_A_HI = 0, OFF_A_LO = A_BYTES, OFF_B_HI = 2 * A_BYTES, OFF_B_LO = 2 * A_BYTES + B_BYTES;
};

template <int BN>
__global__ void __launch_bounds__(NTHREADS_TC, 1) tc_dcore_kernel(const __grid_constant__ TcDcoreArgs a) {
  using SM = DcoreSmem<BN>;
  extern __shared__ unsigned char smem_dyn[];
  const EpsGeom& g = a.g;
  const int Q = g.Q, O = g.O;
  const int BLO = g.BL * O;
  const int TE = g.AH + g.AL + g.BH + BLO;   // table entries per patch
  const int NX = g.n * Q;

  // ---- carve shared memory (operand stages first, 1024-byte aligned for SWIZZLE_128B)
  unsigned char* base = smem_dyn + ((1024u - (tc::smem_u32(smem_dyn) & 1023u)) & 1023u);  // stays a shared-space pointer
  unsigned char* stages = base;
  float* tab = (float*)(base + STAGES * SM::STAGE_BYTES);   // [TE + 1][32]; row TE is all zeros (padding rows)
  float* xs = tab + (TE + 1) * 32;                           // [NX + 1][32]; row NX is all ones (unused factor slots)
  float* gs = xs + (NX + 1) * 32;                            // [O + 1][32];  row O is all ones (entries without gout)
  uint32_t* rowinfo = (uint32_t*)(gs + (O + 1) * 32);        // [BM + BN]: hi entry | lo entry << 16 (TE | TE<<16 = padding row)
  uint2* emeta = (uint2*)(rowinfo + BM + BN);                // [TE]: {4 xs rows packed, gs row} (8-byte aligned: all sizes above are multiples of 8)
  int* xoff = (int*)(emeta + TE);                            // [NX]: element offset of (factor j, component q) from the patch origin
  uint64_t* bars = (uint64_t*)(xoff + ((NX + 1) & ~1));
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 4);
  const uint32_t bar_full0 = tc::smem_u32(bars), bar_empty0 = bar_full0 + 8 * STAGES;
  const uint32_t bar_accfull0 = bar_full0 + 16 * STAGES, bar_accempty0 = bar_accfull0 + 16;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int a0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  long long pbeg = (long long)blockIdx.z * a.per_split;
  long long pend = pbeg + a.per_split;
  if (pend > g.P) pend = g.P;
  const int nchunks = (int)((pend - pbeg + BK - 1) / BK);

  // ---- one-time setup
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(bar_full0 + 8 * s, NPROD_WARPS);
      tc::mbar_init(bar_empty0 + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(bar_accfull0 + 8 * i, 1);
      tc::mbar_init(bar_accempty0 + 8 * i, NPROD_WARPS);
    }
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tc::smem_u32(tmem_slot), 2 * BN);
  // row -> (hi entry, lo entry); padding rows multiply the all-zero table row
  for (int r = tid; r < BM + BN; r += NTHREADS_TC) {
    uint32_t info = (uint32_t)TE | ((uint32_t)TE << 16);
    if (r < BM) {
      int ai = a0 + r;
      if (ai < g.A) info = (uint32_t)(ai / g.AL) | ((uint32_t)(g.AH + ai % g.AL) << 16);
    } else {
      int ni = n0 + (r - BM);
      if (ni < g.N) info = (uint32_t)(g.AH + g.AL + ni / BLO) | ((uint32_t)(g.AH + g.AL + g.BH + ni % BLO) << 16);
    }
    rowinfo[r] = info;
  }
  // table entry -> the (up to 4) xs rows whose product it is, plus its gout row; unused slots point at ones rows
  for (int t = tid; t < TE; t += NTHREADS_TC) {
    int e, j0, cnt, grow = O;
    if (t < g.AH) { e = t; j0 = 0; cnt = g.a_nh; }
    else if (t < g.AH + g.AL) { e = t - g.AH; j0 = g.a_nh; cnt = g.a_nl; }
    else if (t < g.AH + g.AL + g.BH) { e = t - g.AH - g.AL; j0 = g.m; cnt = g.b_nh; }
    else {
      int idx = t - (g.AH + g.AL + g.BH);
      e = idx / O; grow = idx - e * O; j0 = g.m + g.b_nh; cnt = g.b_nl;
    }
    const uint32_t dg = pack_digits(e, cnt, Q);
    uint32_t rows = 0;
    for (int u = 0; u < 4; ++u) {
      uint32_t row = (u < cnt) ? (uint32_t)((j0 + u) * Q) + ((dg >> (8 * u)) & 0xFF) : (uint32_t)NX;
      rows |= row << (8 * u);
    }
    emeta[t] = make_uint2(rows, (uint32_t)grow);
  }
  for (int jq = tid; jq < NX; jq += NTHREADS_TC) xoff[jq] = (int)g.foff[jq / Q] + jq % Q;
  if (tid < 32) {
    tab[TE * 32 + tid] = 0.f;
    xs[NX * 32 + tid] = 1.f;
    gs[O * 32 + tid] = 1.f;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = tc::make_idesc_tf32(BM, BN);
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % STAGES;
      const uint32_t it = (uint32_t)(c / STAGES);
      const int seg = c / SEG_CHUNKS, acc = seg & 1;
      const bool seg_first = (c % SEG_CHUNKS) == 0;
      const bool seg_last = ((c + 1) % SEG_CHUNKS) == 0 || c == nchunks - 1;
      // before overwriting an accumulator, its previous use (segment seg-2) must have been drained
      if (seg_first && seg >= 2) tc::mbar_wait(bar_accempty0 + 8 * acc, (uint32_t)(((seg >> 1) - 1) & 1));
      tc::mbar_wait(bar_full0 + 8 * s, it & 1);
      tc::tc_fence_after();
      if (lane == 0) {
        const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * BN);
        const uint32_t sb = tc::smem_u32(stages + s * SM::STAGE_BYTES);
        const uint64_t da_hi = tc::make_sw128_kmajor_desc(sb + SM::OFF_A_HI);
        const uint64_t da_lo = tc::make_sw128_kmajor_desc(sb + SM::OFF_A_LO);
        const uint64_t db_hi = tc::make_sw128_kmajor_desc(sb + SM::OFF_B_HI);
        const uint64_t db_lo = tc::make_sw128_kmajor_desc(sb + SM::OFF_B_LO);
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const uint64_t adv = (uint64_t)(k * 2);  // 32 bytes >> 4
          const uint32_t accum = (seg_first && k == 0) ? 0u : 1u;
          if (a.passes == 3) {
            // small terms first, then the dominant one
            tc::umma_tf32(tmem_acc, da_lo + adv, db_hi + adv, idesc, accum);
            tc::umma_tf32(tmem_acc, da_hi + adv, db_lo + adv, idesc, 1u);
            tc::umma_tf32(tmem_acc, da_hi + adv, db_hi + adv, idesc, 1u);
          } else {
            tc::umma_tf32(tmem_acc, da_hi + adv, db_hi + adv, idesc, accum);
          }
        }
        tc::umma_commit(bar_empty0 + 8 * s);                       // stage can be refilled once these MMAs have read it
        if (seg_last) tc::umma_commit(bar_accfull0 + 8 * acc);      // this segment's accumulator is complete
      }
      __syncwarp();
    }
  } else {
    // =========================== producers ===========================
    const int pw = warp - 1;  // 0..7
    const int kq = lane & 7;             // which 4 consecutive patches (16-byte chunk of a K-major row)
    const int rsub = lane >> 3;          // which of the 4 rows / entries a warp handles per iteration
    const int r0 = pw * 4 + rsub;        // first row / table entry owned by this thread (then every 32nd)
    const uint32_t off0 = (uint32_t)(r0 * 128 + ((kq ^ (r0 & 7)) << 4));  // its 16-byte chunk in a swizzled tile
    const int quad = warp & 3;           // TMEM lane quadrant this warp may access
    const int half = (warp - 1) >> 2;    // which half of the accumulator columns this warp promotes
    float racc[BN / 2];                  // fp32 running sum of this thread's row, BN/2 columns
#pragma unroll
    for (int i = 0; i < BN / 2; ++i) racc[i] = 0.f;
    int next_drain = 0;
    const unsigned hw = (unsigned)(g.Ho * g.Wo);
    // promote segment `seg` (complete in TMEM) into the register accumulators, then hand the buffer back
    auto drain = [&](int seg) {
      const int acc = seg & 1;
      tc::mbar_wait(bar_accfull0 + 8 * acc, (uint32_t)((seg >> 1) & 1));
      tc::tc_fence_after();
#pragma unroll
      for (int cb = 0; cb < BN / 2; cb += 32) {
        float v[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2) + cb), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) racc[cb + i] += v[i];
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(bar_accempty0 + 8 * acc);
    };
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % STAGES;
      const uint32_t it = (uint32_t)(c / STAGES);
      // (1) stage x and gout of this chunk's 32 patches: lane = patch (32-bit index math, offsets precomputed)
      {
        const unsigned p = (unsigned)pbeg + (unsigned)(c * BK + lane);
        const bool ok = p < (unsigned)pend;
        const unsigned pc = ok ? p : 0u;
        const unsigned b = pc / hw, r = pc - b * hw, h = r / (unsigned)g.Wo, w = r - h * (unsigned)g.Wo;
        const float* xp = a.x + (size_t)(((b * (unsigned)g.H + h) * (unsigned)g.W + w) * (unsigned)Q);
        const float* gp = a.gout + (size_t)pc * O;
#pragma unroll 4
        for (int jq = pw; jq < NX; jq += NPROD_WARPS) xs[jq * 32 + lane] = ok ? __ldg(xp + xoff[jq]) : 0.f;
        for (int o = pw; o < O; o += NPROD_WARPS) gs[o * 32 + lane] = ok ? __ldg(gp + o) : 0.f;
      }
      producer_bar_sync();
      // (2) two-level Khatri-Rao tables.  Thread = (entry, 4 consecutive patches): 128-bit shared accesses,
      //     branch-free (unused factor slots read the all-ones rows); two entries per iteration for ILP.
      {
        const float4* xs4 = (const float4*)xs;
        const float4* gs4 = (const float4*)gs;
        float4* tab4 = (float4*)tab;
        auto entry = [&](int t) -> float4 {
          const uint2 em = emeta[t];
          float4 v = gs4[em.y * 8 + kq];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 f = xs4[((em.x >> (8 * u)) & 0xFF) * 8 + kq];
            v.x *= f.x; v.y *= f.y; v.z *= f.z; v.w *= f.w;
          }
          return v;
        };
        int t = r0;
        for (; t + 32 < TE; t += 64) {
          const float4 v0 = entry(t), v1 = entry(t + 32);
          tab4[t * 8 + kq] = v0;
          tab4[(t + 32) * 8 + kq] = v1;
        }
        if (t < TE) tab4[t * 8 + kq] = entry(t);
      }
      producer_bar_sync();
      // (3) wait until the MMAs that read this stage (two chunks ago) are done, then generate the operand tiles:
      //     thread = (row, 4 consecutive patches) -> one 16-byte chunk of the swizzled row, hi and lo parts
      tc::mbar_wait(bar_empty0 + 8 * s, (it & 1) ^ 1);
      {
        // Each thread owns the same rows r0 + 32*i in every chunk: row r0 + 32*i keeps r & 7, so the swizzled byte
        // offset is (off0 + i*4096) and everything below except the table values is loop-invariant.
        unsigned char* st = stages + s * SM::STAGE_BYTES + off0;
        const float4* tab4 = (const float4*)tab + kq;
        const uint32_t* ri = rowinfo + r0;
        constexpr int NROW = (BM + BN) / 32;
        static_assert(BM % 32 == 0 && BN % 32 == 0 && NROW % 2 == 0, "row ownership pattern");
#pragma unroll
        for (int i = 0; i < NROW; i += 2) {
          float4 hi[2], lo[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const uint32_t info = ri[(i + u) * 32];
            const float4 th = tab4[(info & 0xFFFF) * 8], tl = tab4[(info >> 16) * 8];
            tc::split_tf32(th.x * tl.x, hi[u].x, lo[u].x);
            tc::split_tf32(th.y * tl.y, hi[u].y, lo[u].y);
            tc::split_tf32(th.z * tl.z, hi[u].z, lo[u].z);
            tc::split_tf32(th.w * tl.w, hi[u].w, lo[u].w);
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            constexpr uint32_t A_ROWS = BM / 32;
            const int ii = i + u;
            const uint32_t tile_hi = ii < (int)A_ROWS ? SM::OFF_A_HI + ii * 4096u : SM::OFF_B_HI + (ii - A_ROWS) * 4096u;
            const uint32_t tile_lo = ii < (int)A_ROWS ? SM::OFF_A_LO + ii * 4096u : SM::OFF_B_LO + (ii - A_ROWS) * 4096u;
            *(float4*)(st + tile_hi) = hi[u];
            if (a.passes == 3) *(float4*)(st + tile_lo) = lo[u];
          }
        }
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(bar_full0 + 8 * s);
      // Having been allowed to refill stage s means the MMAs of chunk c-STAGES are complete; one chunk into a
      // new segment (c % SEG_CHUNKS == STAGES-1) that covers the whole previous segment: promote it now, while
      // the tensor core works on the chunks just produced.
      if ((c % SEG_CHUNKS) == STAGES - 1 && c >= SEG_CHUNKS) drain(next_drain++);
    }
    const int last_seg = (nchunks - 1) / SEG_CHUNKS;
    while (next_drain <= last_seg) drain(next_drain++);
    // =========================== epilogue: registers -> partial tile ===========================
    const int arow = a0 + quad * 32 + lane;
    float* prow = a.part + ((long long)blockIdx.z * g.A + arow) * (long long)g.N;
    if (arow < g.A) {
#pragma unroll
      for (int cb = 0; cb < BN / 2; cb += 4) {
        const int nb = n0 + half * (BN / 2) + cb;
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
  if (warp == 0) tc::tmem_dealloc(tmem_base, 2 * BN);
}

template <int BN>
size_t dcore_tc_smem(const EpsGeom& g) {
  const int BLO = g.BL * g.O;
  const int TE = g.AH + g.AL + g.BH + BLO;
  size_t b = 1024 + (size_t)STAGES * DcoreSmem<BN>::STAGE_BYTES + (size_t)(TE + g.n * g.Q + g.O + 3) * 32 * 4 +
             (size_t)(BM + BN) * 4 + 8 + (size_t)TE * 8 + (size_t)(g.n * g.Q + 2) * 4 + (2 * STAGES + 4) * 8 + 16;
  return b;
}

constexpr size_t TC_SMEM_LIMIT = 227 * 1024;

inline void dcore_split(const EpsGeom& g, int BN, long long* per_split, int* splits) {
  // about one wave: one CTA per SM (the kernel uses all of TMEM and ~220 KB of shared memory)
  long long tiles = (long long)((g.A + BM - 1) / BM) * ((g.N + BN - 1) / BN);
  long long want = 148 / tiles;
  if (want < 1) want = 1;
  long long per = (g.P + want - 1) / want;
  per = ((per + BK - 1) / BK) * BK;
  const long long min_per = (long long)BK * SEG_CHUNKS * 4;  // do not split below a few segments
  if (per < min_per) per = min_per;
  *per_split = per;
  *splits = (int)((g.P + per - 1) / per);
}

}  // namespace

bool tc_supported(const EpsGeom& g, int kind) {
  if (kind != 1) return tcg_supported(g, kind);  // forward / input-gradient GEMMs live in eps_tc_gemm.cu
  if (g.a_nh > 4 || g.a_nl > 4 || g.b_nh > 4 || g.b_nl > 4) return false;  // packed digit bytes
  if (g.n * g.Q > 254 || g.O > 254 || g.AH + g.AL + g.BH + g.BL * g.O >= 65535) return false;
  if (g.P >= (1ll << 31) / (g.Q > g.O ? g.Q : g.O)) return false;  // 32-bit patch index math
  if (g.A < 64 || g.N < 128) return false;   // tiles would be mostly padding: the CUDA-core family is the better fit
  if (g.P < 4096) return false;              // tiny reductions are launch-bound either way
  return dcore_tc_smem<256>(g) <= TC_SMEM_LIMIT;
}

size_t tc_workspace_bytes(const EpsGeom& g, int kind) {
  if (kind == 1) {
    long long per;
    int splits;
    dcore_split(g, 256, &per, &splits);
    return (size_t)splits * g.A * g.N * sizeof(float);
  }
  return tcg_workspace_bytes(g, kind);
}

int tc_backward_core(const EpsGeom& g, const float* x, const float* gout, float* dcore, void* ws, int passes,
                     cudaStream_t st) {
  constexpr int BN = 256;
  size_t smem = dcore_tc_smem<BN>(g);
  if (smem > TC_SMEM_LIMIT) return dctn_set_error(-2, "tcgen05 core-gradient kernel needs %zu bytes of shared memory", smem);
  TcDcoreArgs a{};
  a.g = g; a.x = x; a.gout = gout; a.part = (float*)ws; a.passes = passes;
  int splits;
  dcore_split(g, BN, &a.per_split, &splits);
  auto k = tc_dcore_kernel<BN>;
  DCTN_CUDA_CHECK_RET(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((g.A + BM - 1) / BM, (g.N + BN - 1) / BN, splits);
  k<<<grid, NTHREADS_TC, smem, st>>>(a);
  dctn_count_launch();
  DCTN_CUDA_CHECK_RET(cudaGetLastError());
  return launch_reduce_partials<float>(a.part, dcore, (long long)g.A * g.N, splits, st);
}

