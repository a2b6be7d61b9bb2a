// Microbenchmark: issue cost (cycles per warp-instruction per SM sub-partition) of the instructions the operand
// generators use, and of three fp32 -> split-fp16 variants.  nvcc -arch=sm_100a -O3 split_rates.cu -o split_rates
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint32_t f2fp(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }

template <int V> __device__ __forceinline__ void split(u64 v, uint32_t& hi, uint32_t& lo) {
  float v0, v1; unpack2(v, v0, v1);
  if (V == 0) {        // F2FP, 2 HADD2.F32, FMUL2, FFMA2, F2FP
    hi = f2fp(v0, v1);
    __half2 h = *reinterpret_cast<__half2*>(&hi);
    u64 t = mul2(pack2(__low2float(h), __high2float(h)), pack2(-2048.f, -2048.f));
    float r0, r1; unpack2(fma2(v, pack2(2048.f, 2048.f), t), r0, r1);
    lo = f2fp(r0, r1);
  } else if (V == 2) { // Veltkamp rounding in fp32x2, no unpack
    u64 t = mul2(v, pack2(8193.f, 8193.f));
    u64 d = add2(t, v ^ 0x8000000080000000ull);      // t - v
    u64 h = add2(t, d ^ 0x8000000080000000ull);      // t - d : v rounded to 11 significant bits
    float h0, h1; unpack2(h, h0, h1);
    hi = f2fp(h0, h1);
    u64 r = mul2(add2(v, h ^ 0x8000000080000000ull), pack2(2048.f, 2048.f));
    float r0, r1; unpack2(r, r0, r1);
    lo = f2fp(r0, r1);
  } else {             // truncation (TF32-style): LOP3 x2, FADD2, FMUL2, 2 F2FP
    u64 h = v & 0xFFFFE000FFFFE000ull;
    float h0, h1; unpack2(h, h0, h1);
    hi = f2fp(h0, h1);
    u64 r = mul2(add2(v, h ^ 0x8000000080000000ull), pack2(2048.f, 2048.f));
    float r0, r1; unpack2(r, r0, r1);
    lo = f2fp(r0, r1);
  }
}

template <int V> __global__ void k_split(const float* in, uint32_t* out, int iters, long long* cyc) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x * 8 + i];
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      uint32_t h, l;
      split<V>(pack2(a[i], a[i + 1]), h, l);
      acc ^= h + l;
      a[i] += 1.0f; a[i + 1] *= 1.0001f;     // 2 extra scalar FP ops per pair (FADD, FMUL)
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// op mix kernels: OP 0 FMUL2, 1 F2FP, 2 HADD2.F32, 3 LOP3, 4 FMUL, 5 FFMA2 imm
template <int OP> __global__ void k_op(const float* in, uint32_t* out, int iters, long long* cyc) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x * 8 + i];
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      if (OP == 0) { u64 r = mul2(pack2(a[i], a[i + 1]), pack2(1.0001f, 0.9999f)); unpack2(r, a[i], a[i + 1]); }
      if (OP == 1) { uint32_t h = f2fp(a[i], a[i + 1]); a[i] = __uint_as_float(__float_as_uint(a[i]) ^ (h & 1)); acc += h; }
      if (OP == 2) { __half2 h = *reinterpret_cast<__half2*>(&a[i]); a[i + 1] = __low2float(h) ; a[i] = __uint_as_float(__float_as_uint(a[i + 1]) + 1); }
      if (OP == 3) { uint32_t x = __float_as_uint(a[i]) & 0xFFFFE000u; a[i] = __uint_as_float(x ^ __float_as_uint(a[i + 1])); }
      if (OP == 4) { a[i] *= 1.0001f; a[i + 1] *= a[i]; }
      if (OP == 5) { u64 r = fma2(pack2(a[i], a[i + 1]), pack2(2048.f, 2048.f), pack2(a[i + 1], a[i])); unpack2(r, a[i], a[i + 1]); }
    }
  }
  long long t1 = clock64();
  for (int i = 0; i < 8; ++i) acc ^= __float_as_uint(a[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <typename K> void run(const char* name, K kern, int warps, float* in, uint32_t* out, long long* cyc, double per) {
  const int iters = 2000;
  kern<<<148, warps * 32>>>(in, out, 10, cyc);
  kern<<<148, warps * 32>>>(in, out, iters, cyc);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  // each SMSP runs warps/4 warps; per loop iteration each warp does 4 "units"
  printf("%-28s warps/SM=%2d: %7.2f cycles per unit per warp, %7.2f cycles per unit per SMSP (%s)\n", name, warps,
         (double)h / iters / 4, (double)h / iters / 4 / (warps / 4.0), cudaGetErrorString(cudaGetLastError()));
  (void)per;
}
int main() {
  float* in; uint32_t* out; long long* cyc;
  cudaMalloc(&in, 1024 * 8 * 4); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  float h[8192]; for (int i = 0; i < 8192; ++i) h[i] = 1.0f + i * 1e-3f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int warps : {4, 8, 16}) {
    run("split V0 (unpack)", k_split<0>, warps, in, out, cyc, 0);
    run("split V2 (veltkamp)", k_split<2>, warps, in, out, cyc, 0);
    run("split V3 (truncate)", k_split<3>, warps, in, out, cyc, 0);
    run("FMUL2 x1", k_op<0>, warps, in, out, cyc, 0);
    run("F2FP x1 (+LOP,IADD)", k_op<1>, warps, in, out, cyc, 0);
    run("HADD2.F32 x1 (+IADD)", k_op<2>, warps, in, out, cyc, 0);
    run("LOP3 x2", k_op<3>, warps, in, out, cyc, 0);
    run("FMUL x2", k_op<4>, warps, in, out, cyc, 0);
    run("FFMA2 imm x1", k_op<5>, warps, in, out, cyc, 0);
  }
  return 0;
}
