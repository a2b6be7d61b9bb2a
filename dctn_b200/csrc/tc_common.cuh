// Thin inline-PTX layer for the Blackwell (sm_100a) tensor-core path: mbarrier, tcgen05.mma (kind::tf32),
// TMEM allocation / loads, bulk async copies (TMA engine, UBLKCP) and the shared-memory matrix descriptor.
//
// Operand tiles live in shared memory in the canonical K-major SWIZZLE_128B layout: a tile of R rows x 32
// fp32 (= 128 bytes per row); row r starts at byte r*128, its eight 16-byte chunks are XOR-permuted with
// (r & 7); groups of 8 rows are 1024 bytes apart (SBO).  One tcgen05.mma kind::tf32 consumes K = 8 elements
// (32 bytes) per row, so a 32-wide tile is 4 MMA k-steps, each advancing the descriptor start address by 32 B.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded spin: a mis-programmed pipeline traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > (1u << 26)) __trap();
  }
}

// one non-blocking probe: true when the phase with this parity has completed
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}

// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One lane of the (converged) warp, chosen by the hardware.  Unlike `lane == 0`, the compiler knows that exactly one
// thread is active after the branch, so the uniform-datapath instructions inside it (tcgen05.mma, tcgen05.commit, bulk
// copies) are emitted straight-line; behind `if (lane == 0)` each of them sits in its own ELECT / BRA.U.ANY
// serialisation loop (cuobjdump), which cost ~100 cycles per MMA — more than a 128 x 96 x 16 MMA takes to execute.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole warp; writes the TMEM base address to shared memory at `dst`
__device__ __forceinline__ void tmem_alloc(uint32_t dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], fp32 operands read as TF32 (top 19 bits), fp32 accumulate. One thread issues.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// TS form: A operand read from TMEM (lane = row, one 32-bit column per K element), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16 TS form: fp16 operands (A: two consecutive K elements per 32-bit TMEM column, low half first; B: K-major
// shared memory, 64 elements per 128-byte row), fp32 accumulate, K = 16 per instruction
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// SS form of the above (both operands from shared memory)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on `bar` when all previously issued MMAs of this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread `lane` of warp w gets TMEM lane (32*(w%4) + lane)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Two 32-column loads (e.g. the main and the small accumulator of the same columns) in flight together, one wait
__device__ __forceinline__ void tmem_ld32x2(uint32_t taddr0, uint32_t taddr1, float (&v)[32], float (&w)[32]) {
  uint32_t r[32], q[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr0)
      : "memory");
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
        "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]),
        "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]),
        "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
      : "r"(taddr1)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) { v[i] = __uint_as_float(r[i]); w[i] = __uint_as_float(q[i]); }
}

// 32 lanes x 32 consecutive fp32 columns, registers -> TMEM (thread `lane` of warp w writes TMEM lane 32*(w%4)+lane)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
        "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
        "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])),
        "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])),
        "r"(__float_as_uint(v[23])), "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
        "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])),
        "r"(__float_as_uint(v[31]))
      : "memory");
}
// 16-column variant of tmem_st32
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
        "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
        "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15]))
      : "memory");
}
// 32 columns of raw 32-bit words (packed fp16 pairs)
__device__ __forceinline__ void tmem_st32_u(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
// 16 columns of raw 32-bit words
__device__ __forceinline__ void tmem_st16_u(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
// 8 columns of raw 32-bit words
__device__ __forceinline__ void tmem_st8_u(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- bulk async copy (TMA engine, no tensor map)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (SM100 "version 1"), K-major, SWIZZLE_128B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);  // start address  [0,14)
  d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset: 8 rows * 128 B [32,46)
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell) [46,48)
  d |= (uint64_t)2 << 61;                       // layout type SWIZZLE_128B [61,64)
  return d;
}
// Instruction descriptor: D=f32, A=B=tf32, both K-major, dense, shape M x N x 8.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Instruction descriptor: D=f32, A=B=f16, both K-major, dense, shape M x N x 16.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// byte offset of element (row r, column k in [0,32)) inside a SWIZZLE_128B K-major fp32 tile
__device__ __forceinline__ uint32_t sw128_offset(int r, int k) {
  return (uint32_t)(r * 128 + ((((k >> 2) ^ (r & 7)) & 7) << 4) + ((k & 3) << 2));
}

// fp32 -> (tf32-exact high part, remainder): hi has its low 13 mantissa bits cleared, lo = v - hi exactly
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  lo = v - hi;
}

// ---------------------------------------------------------------- split-fp16 number format
// Packed fp32x2 arithmetic (Blackwell FMUL2 / FFMA2): the operand generators are bound by the issue rate of the FP32
// pipe (measured: ~13 cycles per generated element with scalar code), and these halve the instruction count.
typedef unsigned long long f32x2_t;   // two fp32 in one 64-bit register pair (low word = first value)
__device__ __forceinline__ f32x2_t pack2(float a, float b) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ f32x2_t as_f32x2(const float2& v) { return pack2(v.x, v.y); }
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// v -> (fp16(v), fp16((v - fp16(v)) * 2^11)) for the two values of v, packed with the first value in the low half:
// F2FP, 2 x HADD2.F32 (unpack), FMUL2 (-2048 h), FFMA2 (2048 v - 2048 h), F2FP
__device__ __forceinline__ void split_f16x2_p(f32x2_t v, uint32_t& hi, uint32_t& lo) {
  float v0, v1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(v0), "=f"(v1) : "l"(v));
  const __half2 h = __floats2half2_rn(v0, v1);
  const f32x2_t hf = pack2(__low2float(h), __high2float(h));
  f32x2_t t, r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(hf), "l"(pack2(-2048.f, -2048.f)));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(v), "l"(pack2(2048.f, 2048.f)), "l"(t));
  float r0, r1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(r));
  const __half2 l = __floats2half2_rn(r0, r1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void split_f16x2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  split_f16x2_p(pack2(v0, v1), hi, lo);
}
// two fp32 -> packed bf16 pair (round to nearest), `lo` in the low half
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// exponent e with m = f * 2^e, f in [0.5, 1); 0 for m == 0 / inf / nan
__device__ __forceinline__ int norm_exp(float m) {
  int e = 0;
  if (m > 0.f && m < 3.0e38f) frexpf(m, &e);
  return e;
}
// the core is scaled by 2^(15 - norm_exp(max|core|)): largest magnitude in [2^14, 2^15)
__device__ __forceinline__ int core_scale_exp(uint32_t absmax_bits) {
  const float m = __uint_as_float(absmax_bits);
  if (!(m > 0.f) || !(m < 3.0e38f)) return 0;
  return 15 - norm_exp(m);
}

// Length of one accumulation segment of the GEMM kernels, in pipeline stages of 64 fp16 K-values (4 MMAs each).
// The tensor core truncates its fp32 accumulator toward zero on every MMA: measured against the float64 oracle
// (tools/bias_probe.py, profiles/r02_bias_probe.txt) the result shrinks by ~1.7e-8 per accumulation step for random-sign
// operands and ~9e-8 per step for all-positive ones (linear in the chain length), e.g. 4e-5 after the 794 steps of the
// CIFAR (2, 23 -> 24) input gradient.  Chains are therefore cut after 96 steps (24 stages) and the segments summed in
// fp32 by the epilogue (round to nearest), which bounds the bias at ~1.6e-6 / ~9e-6.  DCTN_B200_KSEG overrides (stages).
inline int seg_stages(int nk) {
  int kseg = 24;
  if (const char* e = getenv("DCTN_B200_KSEG")) {
    const int v = atoi(e);
    if (v > 0) kseg = v;
  }
  if (kseg > nk) kseg = nk;
  // equal segments: 25 stages -> 13 + 12 rather than 24 + 1
  const int nseg = (nk + kseg - 1) / kseg;
  return (nk + nseg - 1) / nseg;
}

}  // namespace tc
