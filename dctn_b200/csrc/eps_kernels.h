// Internal launcher interface between api.cu and the kernel families.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "common.cuh"

// ---- CUDA-core family (eps_ffma.cu): any shape, float / double
template <typename T> size_t ffma_workspace_bytes(const EpsGeom& g, int kind);
template <typename T> int ffma_forward(const EpsGeom& g, const T* x, const T* core, T* out, void* ws, cudaStream_t st);
template <typename T> int ffma_backward_core(const EpsGeom& g, const T* x, const T* gout, T* dcore, void* ws, cudaStream_t st);
template <typename T> int ffma_backward_input(const EpsGeom& g, const T* x, const T* core, const T* gout, T* dx, void* ws, cudaStream_t st);

// out[i] = sum_z part[z*count + i], fixed order (deterministic split-K reduction)
template <typename T> int launch_reduce_partials(const T* part, T* out, long long count, int splits, cudaStream_t st);

// per-patch leave-one-out contraction dKR -> dxp[p][j][q] (half 0: first m factors, half 1: the rest) and the
// deterministic gather dxp -> dx
template <typename T> int launch_loo(const EpsGeom& g, const T* x, const T* dkr, long long p0, int np, int half, T* dxp, cudaStream_t st);
template <typename T> int launch_gather_dx(const EpsGeom& g, const T* dxp, T* dx, cudaStream_t st);

// ---- streaming thread-per-patch family for tiny cores (eps_direct.cu): HBM-bound shapes
bool direct_supported(const EpsGeom& g, int dtype);
template <typename T> int direct_forward(const EpsGeom& g, const T* x, const T* core, T* out, cudaStream_t st);
// phi(u) = scale * (sin^2(pi u / 2), cos^2(pi u / 2)) evaluated on load from the raw pixel image (B, H, W): K = 2, C = 1, Q = 2
bool direct_pixels_supported(const EpsGeom& g, int dtype);
template <typename T> int direct_forward_pixels(const EpsGeom& g, const T* pixels, T scale, const T* core, T* out, cudaStream_t st);
// eps_stream_k2q2.cu: K = 2, C = 1, Q_in = 2, float32 — the HBM-bound corner (one task per warp, packed fp32x2 arithmetic)
bool stream_k2q2_supported(const EpsGeom& g, int dtype);
bool stream_k2q2_enabled();   // DCTN_B200_STREAM=0 falls back to the generic small-core kernels (A/B measurements)
int stream_k2q2_forward(const EpsGeom& g, const float* x, const float* core, float* out, cudaStream_t st);
bool stream_k2q2_bwd_supported(const EpsGeom& g, int dtype, int kind);
int stream_k2q2_backward(const EpsGeom& g, int kind, const float* x, const float* core, const float* gout, float* result, void* ws, cudaStream_t st);
int stream_k2q2_forward_pixels(const EpsGeom& g, const float* pixels, float scale, const float* core, float* out, cudaStream_t st);
bool direct_bwd_supported(const EpsGeom& g, int dtype, int kind);
size_t direct_workspace_bytes(const EpsGeom& g, int dtype, int kind);
// kind 1: result = dcore (core unused), kind 2: result = dx
template <typename T> int direct_backward(const EpsGeom& g, int kind, const T* x, const T* core, const T* gout, T* result, void* ws, cudaStream_t st);

// ---- tcgen05 TF32 family (eps_tc.cu): float only, large cores
bool tc_supported(const EpsGeom& g, int kind);
size_t tc_workspace_bytes(const EpsGeom& g, int kind);
// tsave != nullptr: also store the GEMM rows T[P][N] (column order (o, b)) for tc_backward_input_saved
int tc_forward(const EpsGeom& g, const float* x, const float* core, float* out, void* ws, int passes, cudaStream_t st, float* tsave = nullptr);
size_t tcg_saved_bytes(const EpsGeom& g);
int tc_backward_input_saved(const EpsGeom& g, const float* x, const float* core, const float* gout, const float* tsaved, float* dx, void* ws, int passes, cudaStream_t st);
// eps_tc_gemm.cu: forward / input-gradient GEMMs on tcgen05 (A generated on chip, core streamed by bulk copies)
bool tcg_supported(const EpsGeom& g, int kind);
size_t tcg_workspace_bytes(const EpsGeom& g, int kind);
int tc_backward_core(const EpsGeom& g, const float* x, const float* gout, float* dcore, void* ws, int passes, cudaStream_t st);
int tc_backward_input(const EpsGeom& g, const float* x, const float* core, const float* gout, float* dx, void* ws, int passes, cudaStream_t st);

// eps_tc_dcore.cu: split-fp16 core gradient (A' = KR1 x second-half hi group generated into TMEM, second-half lo group x
// gout pre-split once and streamed by bulk copies)
bool tc16_dcore_supported(const EpsGeom& g);
size_t tc16_dcore_workspace_bytes(const EpsGeom& g);
int tc16_backward_core(const EpsGeom& g, const float* x, const float* gout, float* dcore, void* ws, cudaStream_t st);
// eps_tc_fast.cu: register-table variant of the two GEMMs above for power-of-two Q, split-fp16 arithmetic
// (mode 0: input-gradient GEMM dKR1, mode 1: forward, mode 2: input-gradient GEMM with the first leave-one-out stage
// fused: out = W[np][ldc], ldc = tcfast_loo_groups() = hi-group + lo-group entries of the first half)
int tcfast_loo_groups(const EpsGeom& g, int* cnth, int* EH, int* cntl, int* EL);
bool tcfast_supported(const EpsGeom& g, int mode);
size_t tcfast_packed_floats(const EpsGeom& g, int mode);
int tcfast_pack(const EpsGeom& g, int mode, const float* core, float* dst, const uint32_t* absmax, cudaStream_t st);
int tcfast_gemm(const EpsGeom& g, int mode, const float* x, const float* gout, const float* packed, const uint32_t* absmax,
                long long p0, int np, float* out, long long ldc, float* tsave, cudaStream_t st);

// ---- logmatmulexp (logmatmulexp.cu)
template <typename T> int lme_forward(const T* A, const T* B, T* out, int Th, int R, int I, cudaStream_t st);
template <typename T> int lme_backward(const T* A, const T* B, const T* out, const T* gout, T* dA, T* dB, int Th, int R, int I, cudaStream_t st);
// logmatmulexp_tile.cu: one fused kernel per product for matrices whose inner dimension fits shared memory; `ws` carries the
// row / column maxima and the "took the per-element path" flag from forward to backward
size_t lme_tile_workspace_bytes(int Th, int I, size_t es);
template <typename T> bool lme_tile_supported(int Th, int R, int I);
template <typename T> int lme_tile_forward(const T* A, const T* B, T* out, int Th, int R, int I, void* ws, cudaStream_t st);
template <typename T> int lme_tile_backward(const T* A, const T* B, const T* out, const T* gout, T* dA, T* dB, int Th, int R, int I, const void* ws, cudaStream_t st);
// batched small matrices: A [NB][Th][R], B [NB][R][I] -> out [NB][Th][I] (ConvSBS bond-matrix rings in log space)
template <typename T> int lme_batched_forward(const T* A, const T* B, T* out, long long NB, int Th, int R, int I, cudaStream_t st);
template <typename T> int lme_batched_backward(const T* A, const T* B, const T* out, const T* gout, T* dA, T* dB, long long NB, int Th, int R, int I, cudaStream_t st);

// ---- statistics (stats.cu): (sum, sum of squares) accumulated in double INTO stats[0..1]
size_t value_stats_workspace_bytes();
template <typename T> int launch_value_stats(const T* v, long long n, double* stats, void* ws, cudaStream_t st);
size_t window_stats_workspace_bytes(int C, int B, int H, int W);
template <typename T> int launch_window_stats(const T* x, int C, int B, int H, int W, int Q, int K, double* stats, void* ws, cudaStream_t st);
