/*
 * dctn_b200 — C ABI of the B200-native EPS contraction library (libdctn_b200.so).
 *
 * The reference (philip-bl/dctn) is pure Python: it has no FFI/plugin interface, its boundary for
 * this path is the Python function signature.  The entry points below are what a binding for the
 * reference's hot path binds instead of the opt_einsum/torch.einsum calls; each one cites the
 * reference call site it replaces.  INTEGRATION.md shows the ctypes stub and the
 * torch.autograd.Function a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless the name ends in
 *     `_host`; every tensor is dense row-major ("contiguous") and 16-byte aligned;
 *   - every launch goes on the cudaStream_t passed by the caller (as a void*), nothing synchronises;
 *   - the library allocates no device memory: scratch is a caller-provided workspace whose size is
 *     queried with dctn_eps_workspace_bytes();
 *   - functions return 0 on success and a negative dctn_status on failure, never throw;
 *     dctn_last_error() returns a thread-local message.  There is no CPU fallback.
 *
 * Tensor layouts (reference names)
 *   input  x    : (C, B, H, W, Q_in)        dctn/eps.py:20
 *   core        : (Q_in,)*(K*K*C) + (Q_out,) dctn/eps.py:22,66-70; flat view [D = Q_in^(K*K*C)][Q_out],
 *                 factor j = (dh*K + dw)*C + c is the j-th (slowest-first) index  dctn/align.py:20-46
 *   output      : (B, H-K+1, W-K+1, Q_out)   dctn/eps.py:39
 */
#ifndef DCTN_B200_H
#define DCTN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCTN_B200_VERSION 100 /* 0.1.0 */

typedef enum dctn_status {
  DCTN_OK = 0,
  DCTN_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, image smaller than the kernel */
  DCTN_ERR_UNSUPPORTED = -2,  /* shape outside what the kernels implement (message says which limit) */
  DCTN_ERR_WORKSPACE = -3,    /* workspace pointer null or smaller than dctn_eps_workspace_bytes() */
  DCTN_ERR_CUDA = -4          /* a CUDA runtime call failed (message holds cudaGetErrorString) */
} dctn_status;

typedef enum dctn_dtype { DCTN_F32 = 0, DCTN_F64 = 1 } dctn_dtype;

/* Kernel family selection ("static per-shape kernel plan", replaces the opt_einsum path cache
 * dctn/contraction_path_cache.py:19-35 and the hard-coded path dctn/eps.py:25-30). */
typedef enum dctn_variant {
  DCTN_VARIANT_AUTO = 0, /* library picks per shape (default) */
  DCTN_VARIANT_FFMA = 1, /* CUDA-core FMA kernels, exact fp32/fp64 arithmetic, any shape */
  DCTN_VARIANT_TC3 = 2,  /* tcgen05 TF32 tensor cores, 3-pass split (hi*hi + hi*lo + lo*hi): fp32-accurate */
  DCTN_VARIANT_TC1 = 3,  /* tcgen05 TF32 single pass (rel. err ~1e-3), opt-in only */
  DCTN_VARIANT_DIRECT = 4, /* thread-per-patch streaming kernel for tiny cores (HBM-bound shapes) */
  DCTN_VARIANT_TCH3 = 5  /* tcgen05 FP16 tensor cores, 3-pass split with power-of-two range normalisation: same 22
                            significant bits per operand as TC3 at twice the MMA rate; what AUTO uses on large cores */
} dctn_variant;

typedef enum dctn_ws_kind {
  DCTN_WS_FORWARD = 0,
  DCTN_WS_BACKWARD_CORE = 1,
  DCTN_WS_BACKWARD_INPUT = 2,
  DCTN_WS_BACKWARD_INPUT_SAVED = 3, /* dctn_eps_backward_input_saved() */
  DCTN_WS_FORWARD_STATS = 4         /* dctn_eps_forward_stats() */
} dctn_ws_kind;

typedef struct dctn_plan dctn_plan_t; /* opaque, owned by the library's plan cache, never freed by the caller */

int dctn_version(void);
const char* dctn_last_error(void);

/* Returns the cached plan for an EPS layer (kernel_size K, in_num_channels C, in_size Qin,
 * out_size Qout — dctn/eps.py:66-70 calc_eps_shape), or NULL (see dctn_last_error).  Thread-safe. */
const dctn_plan_t* dctn_eps_plan_get(int C, int K, int Qin, int Qout, int dtype, int variant);

/* Human-readable description of the plan (split, tile shapes, kernel family); static storage per plan. */
const char* dctn_eps_plan_describe(const dctn_plan_t* plan);

/* Which kernel family the library's per-shape dispatch uses for call `kind` (a dctn_ws_kind) at this size:
 * DCTN_FAMILY_*; a negative dctn_status when the plan's forced variant cannot serve the shape.  Diagnostic: tests and
 * benchmarks assert with it that the tcgen05 kernels, not a CUDA-core path, served a call. */
typedef enum dctn_family { DCTN_FAMILY_CUDA_CORE = 1, DCTN_FAMILY_TCGEN05 = 2, DCTN_FAMILY_STREAMING = 4 } dctn_family;
int dctn_eps_kernel_family(const dctn_plan_t* plan, int B, int H, int W, int kind);

/* Scratch bytes needed by the call `kind` on a (C,B,H,W,Qin) input. */
size_t dctn_eps_workspace_bytes(const dctn_plan_t* plan, int B, int H, int W, int kind);

/* out = eps(core, x).  Replaces dctn/eps.py:19-40 (align views + 4-step opt_einsum contraction);
 * eps_one_by_one (dctn/eps.py:43-63) maps to the same call. */
int dctn_eps_forward(const dctn_plan_t* plan, const void* x, const void* core, void* out,
                     int B, int H, int W, void* workspace, size_t workspace_bytes, void* stream);

/* dcore = d<gout, eps(core, x)>/dcore.  Replaces autograd through dctn/eps.py:31-40 w.r.t. `core`
 * (dense reduction over all patches).  dcore is overwritten (not accumulated). */
int dctn_eps_backward_core(const dctn_plan_t* plan, const void* x, const void* gout, void* dcore,
                           int B, int H, int W, void* workspace, size_t workspace_bytes, void* stream);

/* dx = d<gout, eps(core, x)>/dx.  Replaces autograd through dctn/eps.py:31-40 w.r.t. `input`
 * (needed for layers >= 2, dctn/epses_composition.py:139-141).  dx is overwritten. */
int dctn_eps_backward_input(const dctn_plan_t* plan, const void* x, const void* core, const void* gout,
                            void* dx, int B, int H, int W, void* workspace, size_t workspace_bytes,
                            void* stream);

/* Training forward: as dctn_eps_forward, and additionally keeps the forward GEMM's per-patch intermediate
 * T[p][(o, b)] = sum_a KR1[p][a] * core[a][b][o]  (P x Q_in^(n-m) * Q_out elements, n-m = floor(K*K*C/2) factors)
 * in the caller's `saved` buffer, so that the input gradient does not have to recompute it — the reference's
 * autograd keeps the same tensor (step 2 of dctn/eps.py:31-40) plus two larger ones.
 * dctn_eps_saved_bytes() returns the size of `saved`, or 0 when the kernel family serving this plan/shape has no
 * savable intermediate (then use dctn_eps_forward / dctn_eps_backward_input). */
size_t dctn_eps_saved_bytes(const dctn_plan_t* plan, int B, int H, int W);
int dctn_eps_forward_train(const dctn_plan_t* plan, const void* x, const void* core, void* out, void* saved,
                           size_t saved_bytes, int B, int H, int W, void* workspace, size_t workspace_bytes,
                           void* stream);
/* dx from the intermediate saved by dctn_eps_forward_train (same plan, same x, same core).  Workspace kind
 * DCTN_WS_BACKWARD_INPUT_SAVED. */
int dctn_eps_backward_input_saved(const dctn_plan_t* plan, const void* x, const void* core, const void* gout,
                                  const void* saved, size_t saved_bytes, void* dx, int B, int H, int W,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* Forward plus statistics: as dctn_eps_forward, and (sum, sum of squares) of `out` are ADDED in double precision to the
 * device-resident pair stats[0..1] (zero it before the first slice).  Replaces the `torch.cat` + `output.std()` pass of
 * make_eps_unit_empirical_output_std (dctn/eps.py:163-181, slices from transform_in_slices dctn/eps.py:126-137): the
 * biased variance is stats[1]/n - (stats[0]/n)^2.  Workspace kind DCTN_WS_FORWARD_STATS.  Deterministic. */
int dctn_eps_forward_stats(const dctn_plan_t* plan, const void* x, const void* core, void* out, double* stats,
                           int B, int H, int W, void* workspace, size_t workspace_bytes, void* stream);

/* Statistics of all K x K windows of x (C, B, H, W, Q) seen as rank-one tensors of K*K*C factors, without expanding
 * them: stats[0] += sum over windows of prod_j sum_q x_j[q], stats[1] += sum over windows of prod_j sum_q x_j[q]^2.
 * Replaces make_windows (dctn/align.py:49-61) + RankOneTensorsBatch.sum_over_batch / squared_fro_norm_over_batch
 * (dctn/rank_one_tensor.py:53-98) as used by log_intermediate_reps_stats (dctn/eps_plus_linear.py:176-186). */
size_t dctn_window_stats_workspace_bytes(int C, int B, int H, int W);
int dctn_window_stats(const void* x, int C, int B, int H, int W, int Q, int K, int dtype, double* stats,
                      void* workspace, size_t workspace_bytes, void* stream);

/* First layer fed with RAW pixels: out = eps(core, phi(pixels)), phi(u) = scale * (sin^2(pi u / 2), cos^2(pi u / 2)) —
 * the feature map the reference applies in its data loader (dctn/dataset_loading.py:33-36; scale = 2 nu,
 * new_runner.py:358-361) — evaluated inside the kernel, so one float per pixel crosses HBM instead of two.
 * pixels: (B, H, W) of the plan's dtype.  An ADDITIONAL entry: K = 2, C = 1, Q_in = 2 plans of
 * the streaming family only (the HBM-bound regime); DCTN_ERR_UNSUPPORTED otherwise.  No workspace. */
int dctn_eps_forward_from_pixels(const dctn_plan_t* plan, const void* pixels, double scale, const void* core,
                                 void* out, int B, int H, int W, void* stream);

/* out[t,i] = log sum_r exp(log_A[t,r] + log_B[r,i]).  Replaces dctn/logmatmulexp.py:5-14. */
int dctn_logmatmulexp_forward(const void* log_A, const void* log_B, void* out, int Theta, int R, int I,
                              int dtype, void* stream);

/* Gradients of the above given out (saved from forward) and gout; nothing Theta*R*I-sized is kept,
 * which is what logmatmulexp_lowmem (dctn/logmatmulexp.py:17-22) buys with checkpointing.
 * dA / dB may be NULL to skip. */
int dctn_logmatmulexp_backward(const void* log_A, const void* log_B, const void* out, const void* gout,
                               void* dA, void* dB, int Theta, int R, int I, int dtype, void* stream);

/* The same product and gradients by ONE fused kernel each, for matrices whose inner dimension fits shared memory (the
 * reference's use: chains of N x N matrices, N <= 300, small_experiments/logmatmulexp_benchmark/benchmark.py:21-52):
 * out = m_t + n_i + log(exp(A - m_t) @ exp(B - n_i)) with row / column maxima m, n — Theta*R + R*I exponentials instead
 * of Theta*R*I — guarded per tile by a dynamic-range test; tiles that fail it (e.g. the scale-150 inputs of
 * small_experiments/logmatmulexp_old.py:149-153) use the per-element max-shifted form, so the result is the stable one in
 * every case.  dctn_logmatmulexp_workspace_bytes() returns the scratch size, or 0 when the shape is served by the two
 * entries above only; the forward call's workspace must be handed unchanged to the backward call. */
size_t dctn_logmatmulexp_workspace_bytes(int Theta, int R, int I, int dtype);
int dctn_logmatmulexp_forward_ws(const void* log_A, const void* log_B, void* out, int Theta, int R, int I, int dtype,
                                 void* workspace, size_t workspace_bytes, void* stream);
int dctn_logmatmulexp_backward_ws(const void* log_A, const void* log_B, const void* out, const void* gout, void* dA,
                                  void* dB, int Theta, int R, int I, int dtype, const void* workspace,
                                  size_t workspace_bytes, void* stream);

/* Batched small-matrix form: log_A [batch][Theta][R], log_B [batch][R][I] -> out [batch][Theta][I], one product per
 * batch element.  ADDITIONAL entry: the reference's logmatmulexp asserts 2-D (dctn/logmatmulexp.py:8-10)
 * and contracts ConvSBS bond-matrix rings in linear space (dctn/conv_sbs.py:258-304); this is that ring step in log
 * space, one batch element per (image, window).  DCTN_ERR_UNSUPPORTED when one pair exceeds 96 KiB of shared memory.
 * out / dA / dB must be 16-byte aligned (128-bit stores). */
int dctn_logmatmulexp_batched_forward(const void* log_A, const void* log_B, void* out, long long batch, int Theta,
                                      int R, int I, int dtype, void* stream);
int dctn_logmatmulexp_batched_backward(const void* log_A, const void* log_B, const void* out, const void* gout,
                                       void* dA, void* dB, long long batch, int Theta, int R, int I, int dtype,
                                       void* stream);

/* Host-buffer convenience entry (end-to-end path): copies x and core from HOST memory, runs the
 * forward on `stream`, copies `out` back and synchronises the stream.  All three pointers are host
 * pointers; device scratch of dctn_eps_forward_host_device_bytes() bytes is passed by the caller. */
size_t dctn_eps_forward_host_device_bytes(const dctn_plan_t* plan, int B, int H, int W);
int dctn_eps_forward_host(const dctn_plan_t* plan, const void* x_host, const void* core_host,
                          void* out_host, int B, int H, int W, void* device_scratch,
                          size_t device_scratch_bytes, void* stream);

/* Counts kernels launched by this library on the calling process since load (bench.py's gpu_launches). */
unsigned long long dctn_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DCTN_B200_H */
