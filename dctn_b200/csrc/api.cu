// C ABI of libdctn_b200.so (include/dctn_b200.h): plan cache, geometry, per-shape dispatch.
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>

#include "../../include/dctn_b200.h"
#include "common.cuh"
#include "eps_kernels.h"

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";
static std::atomic<unsigned long long> g_launches{0};

int dctn_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int dctn_set_cuda_error(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return DCTN_ERR_CUDA;
}
void dctn_count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------------------ plan
struct dctn_plan {
  int C, K, Q, O, dtype, variant;
  int n, m;
  int a_nh, a_nl, b_nh, b_nl;
  long long A, Bn, N, D;
  std::string desc;
};

static std::mutex g_plan_mu;
static std::map<std::tuple<int, int, int, int, int, int>, dctn_plan*> g_plans;

static bool pow_fits(int q, int e, long long limit, long long* out) {
  long long r = 1;
  for (int i = 0; i < e; ++i) {
    r *= q;
    if (r > limit) return false;
  }
  *out = r;
  return true;
}

static void fill_geom(const dctn_plan* pl, int B, int H, int W, EpsGeom* g) {
  memset(g, 0, sizeof(*g));
  g->C = pl->C; g->K = pl->K; g->Q = pl->Q; g->O = pl->O;
  g->n = pl->n; g->m = pl->m;
  g->A = (int)pl->A; g->Bn = (int)pl->Bn; g->N = (int)pl->N;
  g->a_nh = pl->a_nh; g->a_nl = pl->a_nl; g->b_nh = pl->b_nh; g->b_nl = pl->b_nl;
  g->AH = ipow_host(pl->Q, pl->a_nh); g->AL = ipow_host(pl->Q, pl->a_nl);
  g->BH = ipow_host(pl->Q, pl->b_nh); g->BL = ipow_host(pl->Q, pl->b_nl);
  g->B = B; g->H = H; g->W = W; g->Ho = H - pl->K + 1; g->Wo = W - pl->K + 1;
  g->P = (long long)B * g->Ho * g->Wo;
  long long chan = (long long)B * H * W * pl->Q;
  for (int dh = 0; dh < pl->K; ++dh)
    for (int dw = 0; dw < pl->K; ++dw)
      for (int c = 0; c < pl->C; ++c) {
        int j = (dh * pl->K + dw) * pl->C + c;  // dctn/align.py:20-46
        g->foff[j] = c * chan + ((long long)dh * W + dw) * pl->Q;
      }
}

extern "C" int dctn_version(void) { return DCTN_B200_VERSION; }
extern "C" const char* dctn_last_error(void) { return g_err; }
extern "C" unsigned long long dctn_launch_count(void) { return g_launches.load(); }

extern "C" const dctn_plan_t* dctn_eps_plan_get(int C, int K, int Qin, int Qout, int dtype, int variant) {
  if (C < 1 || K < 1 || Qin < 1 || Qout < 1) {
    dctn_set_error(DCTN_ERR_BAD_ARG, "plan: C, K, Qin, Qout must be positive (got %d %d %d %d)", C, K, Qin, Qout);
    return nullptr;
  }
  if (dtype != DCTN_F32 && dtype != DCTN_F64) {
    dctn_set_error(DCTN_ERR_BAD_ARG, "plan: dtype must be DCTN_F32 or DCTN_F64");
    return nullptr;
  }
  if (variant < DCTN_VARIANT_AUTO || variant > DCTN_VARIANT_TCH3) {
    dctn_set_error(DCTN_ERR_BAD_ARG, "plan: unknown variant %d", variant);
    return nullptr;
  }
  auto key = std::make_tuple(C, K, Qin, Qout, dtype, variant);
  std::lock_guard<std::mutex> lk(g_plan_mu);
  auto it = g_plans.find(key);
  if (it != g_plans.end()) return it->second;

  int n = K * K * C;
  if (n > DCTN_MAXN) {
    dctn_set_error(DCTN_ERR_UNSUPPORTED, "plan: K*K*C = %d factors per patch exceeds the limit %d", n, DCTN_MAXN);
    return nullptr;
  }
  int m = (n + 1) / 2;  // the reference's split, dctn/eps.py:25-27
  long long A, Bn;
  const long long LIM = 1ll << 30;
  // The split is free (core[a][b][o] is one flat [Q^n][O] array for every m).  Mid-sized cores whose halves are too
  // narrow for the 128-row tensor-core tiles (A = Q^m < 64, e.g. CIFAR (2, 6 -> 24): A = Bn = 36) get a lopsided split
  // that is wide enough (A' >= 64 and N' = Q^(n-m') * O >= 64) instead of the CUDA-core kernels.  DCTN_B200_SPLIT_M forces m.
  if (dtype == DCTN_F32 && (variant == DCTN_VARIANT_AUTO || variant == DCTN_VARIANT_TCH3 || variant == DCTN_VARIANT_TC3 || variant == DCTN_VARIANT_TC1)) {
    long long a0, b0, d0;
    if (pow_fits(Qin, m, LIM, &a0) && pow_fits(Qin, n - m, LIM, &b0) && pow_fits(Qin, n, LIM, &d0) && d0 >= 1024 &&
        (a0 < 64 || b0 * Qout < 64)) {
      for (int mm = 1; mm < n; ++mm) {
        long long a1, b1;
        if (!pow_fits(Qin, mm, LIM, &a1) || !pow_fits(Qin, n - mm, LIM, &b1)) continue;
        if (a1 >= 64 && b1 * Qout >= 64 && a1 <= 4096) { m = mm; break; }
      }
    }
    // Four-factor layers (K = 2, one channel) whose Q_in is not a power of two run on the table-lookup GEMM kernels, where
    // the forward epilogue (Q_in^(n-m) * Q_out columns per patch) and the saved intermediate T (as many floats per patch)
    // dominate: three factors in the first half leave Q_in * Q_out columns instead of Q_in^2 * Q_out.  CIFAR (2, 12 -> 24)
    // at B = 64: forward 0.87 -> 0.51 ms, core gradient 0.33 -> 0.27, input gradient 0.82 -> 1.03; T 796 -> 66 MB.
    if (n == 4 && (Qin & (Qin - 1)) != 0 && pow_fits(Qin, 3, 4096, &a0) && pow_fits(Qin, n, LIM, &d0) && d0 >= 1024 && a0 >= 64 &&
        (long long)Qin * Qout >= 64)
      m = 3;
  }
  if (const char* e = getenv("DCTN_B200_SPLIT_M")) {
    const int mm = atoi(e);
    if (mm >= 1 && mm < n) m = mm;
  }
  if (!pow_fits(Qin, m, LIM, &A) || !pow_fits(Qin, n - m, LIM, &Bn) || Bn * Qout > LIM || A * Bn * Qout >= (1ll << 31)) {
    dctn_set_error(DCTN_ERR_UNSUPPORTED, "plan: core with Qin^(K*K*C) = %d^%d elements is too large", Qin, n);
    return nullptr;
  }
  if ((variant == DCTN_VARIANT_TC3 || variant == DCTN_VARIANT_TC1 || variant == DCTN_VARIANT_TCH3) && dtype != DCTN_F32) {
    dctn_set_error(DCTN_ERR_UNSUPPORTED, "plan: the tcgen05 tensor-core variants are float32 only");
    return nullptr;
  }
  dctn_plan* pl = new dctn_plan();
  pl->C = C; pl->K = K; pl->Q = Qin; pl->O = Qout; pl->dtype = dtype; pl->variant = variant;
  pl->n = n; pl->m = m; pl->A = A; pl->Bn = Bn; pl->N = Bn * Qout; pl->D = A * Bn;
  pl->a_nl = m / 2; pl->a_nh = m - pl->a_nl;
  pl->b_nl = (n - m) / 2; pl->b_nh = (n - m) - pl->b_nl;
  char buf[512];
  snprintf(buf, sizeof(buf),
           "EPS plan C=%d K=%d Qin=%d Qout=%d %s variant=%d: n=%d factors, split m=%d (A=%lld, Bn=%lld, N=%lld, D=%lld), "
           "tables A:(%d|%d) B:(%d|%d)",
           C, K, Qin, Qout, dtype == DCTN_F32 ? "f32" : "f64", variant, n, m, A, Bn, pl->N, pl->D, pl->a_nh,
           pl->a_nl, pl->b_nh, pl->b_nl);
  pl->desc = buf;
  g_plans[key] = pl;
  return pl;
}

extern "C" const char* dctn_eps_plan_describe(const dctn_plan_t* plan) { return plan ? plan->desc.c_str() : "(null plan)"; }

// ------------------------------------------------------------------------------------------------ dispatch
enum Family { FAM_FFMA = 1, FAM_TC = 2, FAM_DIRECT = 4 };

static int check_call(const dctn_plan_t* pl, int B, int H, int W) {
  if (!pl) return dctn_set_error(DCTN_ERR_BAD_ARG, "null plan");
  if (B < 1 || H < pl->K || W < pl->K)
    return dctn_set_error(DCTN_ERR_BAD_ARG, "input (B=%d, H=%d, W=%d) must have B >= 1 and H, W >= kernel_size %d", B, H, W, pl->K);
  return 0;
}

// arithmetic of the tensor-core family (the `passes` argument of the tc_* launchers): 1 / 3 = TF32 passes, 6 = split fp16.
// AUTO: split fp16 unless DCTN_B200_AUTO_ARITH=tf32.
static int tc_arith(const dctn_plan_t* pl, int kind) {
  static const int auto_arith = [] {
    const char* e = getenv("DCTN_B200_AUTO_ARITH");
    return (e && strcmp(e, "tf32") == 0) ? 3 : 6;
  }();
  int a = pl->variant == DCTN_VARIANT_TC1 ? 1 : pl->variant == DCTN_VARIANT_TC3 ? 3 : pl->variant == DCTN_VARIANT_TCH3 ? 6 : auto_arith;
  return a;
}

// which kernel family serves `kind` for this plan/geometry
static int pick_family(const dctn_plan_t* pl, const EpsGeom& g, int kind) {
  switch (pl->variant) {
    case DCTN_VARIANT_FFMA: return FAM_FFMA;
    case DCTN_VARIANT_TC3:
    case DCTN_VARIANT_TCH3:
    case DCTN_VARIANT_TC1: return tc_supported(g, kind) ? FAM_TC : -1;
    case DCTN_VARIANT_DIRECT:
      if (kind == DCTN_WS_FORWARD) return direct_supported(g, pl->dtype) ? FAM_DIRECT : -1;
      return direct_bwd_supported(g, pl->dtype, kind) ? FAM_DIRECT : -1;
    default: break;
  }
  if (kind == DCTN_WS_FORWARD && direct_supported(g, pl->dtype)) return FAM_DIRECT;
  if (kind != DCTN_WS_FORWARD && direct_bwd_supported(g, pl->dtype, kind)) return FAM_DIRECT;
  if (pl->dtype == DCTN_F32 && tc_supported(g, kind)) return FAM_TC;
  return FAM_FFMA;
}

// which kernel family serves call `kind` at this size: 1 = CUDA-core (FFMA/DFMA), 2 = tcgen05 tensor-core, 4 = streaming
// small-core; negative status when the plan's forced variant does not support the shape
extern "C" int dctn_eps_kernel_family(const dctn_plan_t* pl, int B, int H, int W, int kind) {
  int rc = check_call(pl, B, H, W);
  if (rc) return rc;
  if (kind < DCTN_WS_FORWARD || kind > DCTN_WS_BACKWARD_INPUT_SAVED) return dctn_set_error(DCTN_ERR_BAD_ARG, "kernel_family: unknown call kind %d", kind);
  EpsGeom g;
  fill_geom(pl, B, H, W, &g);
  int fam = pick_family(pl, g, kind == DCTN_WS_BACKWARD_INPUT_SAVED ? DCTN_WS_BACKWARD_INPUT : kind);
  if (fam < 0) return dctn_set_error(DCTN_ERR_UNSUPPORTED, "kernel_family: variant %d does not support this shape (%s)", pl->variant, pl->desc.c_str());
  return fam;
}

extern "C" size_t dctn_eps_workspace_bytes(const dctn_plan_t* pl, int B, int H, int W, int kind) {
  if (check_call(pl, B, H, W)) return 0;
  EpsGeom g;
  fill_geom(pl, B, H, W, &g);
  if (kind == DCTN_WS_FORWARD_STATS)   // the forward's workspace followed by the per-block partial sums
    return ((dctn_eps_workspace_bytes(pl, B, H, W, DCTN_WS_FORWARD) + 255) & ~(size_t)255) + value_stats_workspace_bytes();
  int fam = pick_family(pl, g, kind == DCTN_WS_BACKWARD_INPUT_SAVED ? DCTN_WS_BACKWARD_INPUT : kind);
  if (kind == DCTN_WS_BACKWARD_INPUT_SAVED && fam != FAM_TC) return 256;
  size_t bytes = 0;
  if (fam == FAM_FFMA) bytes = pl->dtype == DCTN_F32 ? ffma_workspace_bytes<float>(g, kind) : ffma_workspace_bytes<double>(g, kind);
  else if (fam == FAM_TC) bytes = tc_workspace_bytes(g, kind);
  else if (fam == FAM_DIRECT) bytes = direct_workspace_bytes(g, pl->dtype, kind);
  return bytes + 256;  // never zero, so callers can always pass a valid pointer
}

// the kernels use 128-bit loads and stores on every tensor
static int check_aligned(const void* p, const char* what) {
  if (((uintptr_t)p & 15) != 0) return dctn_set_error(DCTN_ERR_BAD_ARG, "%s (%p) must be 16-byte aligned", what, p);
  return 0;
}

static int check_ws(const dctn_plan_t* pl, int B, int H, int W, int kind, void* ws, size_t ws_bytes) {
  size_t need = dctn_eps_workspace_bytes(pl, B, H, W, kind);
  if (!ws || ws_bytes < need)
    return dctn_set_error(DCTN_ERR_WORKSPACE, "workspace of %zu bytes needed, got %zu (%p)", need, ws_bytes, ws);
  return 0;
}

extern "C" int dctn_eps_forward(const dctn_plan_t* pl, const void* x, const void* core, void* out, int B, int H,
                                int W, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_call(pl, B, H, W);
  if (rc) return rc;
  if (!x || !core || !out) return dctn_set_error(DCTN_ERR_BAD_ARG, "forward: null tensor pointer");
  if ((rc = check_aligned(x, "input")) || (rc = check_aligned(core, "core")) || (rc = check_aligned(out, "output")) || (rc = check_aligned(ws, "workspace"))) return rc;
  if ((rc = check_ws(pl, B, H, W, DCTN_WS_FORWARD, ws, ws_bytes))) return rc;
  EpsGeom g;
  fill_geom(pl, B, H, W, &g);
  cudaStream_t st = (cudaStream_t)stream;
  int fam = pick_family(pl, g, DCTN_WS_FORWARD);
  if (fam < 0) return dctn_set_error(DCTN_ERR_UNSUPPORTED, "forward: requested kernel variant %d does not support this shape (%s)", pl->variant, pl->desc.c_str());
  if (fam == FAM_DIRECT)
    return pl->dtype == DCTN_F32 ? direct_forward<float>(g, (const float*)x, (const float*)core, (float*)out, st)
                                 : direct_forward<double>(g, (const double*)x, (const double*)core, (double*)out, st);
  if (fam == FAM_TC)
    return tc_forward(g, (const float*)x, (const float*)core, (float*)out, ws, tc_arith(pl, DCTN_WS_FORWARD), st);
  return pl->dtype == DCTN_F32 ? ffma_forward<float>(g, (const float*)x, (const float*)core, (float*)out, ws, st)
                               : ffma_forward<double>(g, (const double*)x, (const double*)core, (double*)out, ws, st);
}

extern "C" int dctn_eps_backward_core(const dctn_plan_t* pl, const void* x, const void* gout, void* dcore, int B,
                                      int H, int W, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_call(pl, B, H, W);
  if (rc) return rc;
  if (!x || !gout || !dcore) return dctn_set_error(DCTN_ERR_BAD_ARG, "backward_core: null tensor pointer");
  if ((rc = check_aligned(x, "input")) || (rc = check_aligned(gout, "grad_output")) || (rc = check_aligned(dcore, "grad_core")) || (rc = check_aligned(ws, "workspace"))) return rc;
  if ((rc = check_ws(pl, B, H, W, DCTN_WS_BACKWARD_CORE, ws, ws_bytes))) return rc;
  EpsGeom g;
  fill_geom(pl, B, H, W, &g);
  cudaStream_t st = (cudaStream_t)stream;
  int fam = pick_family(pl, g, DCTN_WS_BACKWARD_CORE);
  if (fam < 0) return dctn_set_error(DCTN_ERR_UNSUPPORTED, "backward_core: requested kernel variant %d does not support this shape (%s)", pl->variant, pl->desc.c_str());
  if (fam == FAM_DIRECT)
    return pl->dtype == DCTN_F32 ? direct_backward<float>(g, 1, (const float*)x, nullptr, (const float*)gout, (float*)dcore, ws, st)
                                 : direct_backward<double>(g, 1, (const double*)x, nullptr, (const double*)gout, (double*)dcore, ws, st);
  if (fam == FAM_TC)
    return tc_backward_core(g, (const float*)x, (const float*)gout, (float*)dcore, ws, tc_arith(pl, DCTN_WS_BACKWARD_CORE), st);
  return pl->dtype == DCTN_F32 ? ffma_backward_core<float>(g, (const float*)x, (const float*)gout, (float*)dcore, ws, st)
                               : ffma_backward_core<double>(g, (const double*)x, (const double*)gout, (double*)dcore, ws, st);
}

extern "C" int dctn_eps_backward_input(const dctn_plan_t* pl, const void* x, const void* core, const void* gout,
                                       void* dx, int B, int H, int W, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_call(pl, B, H, W);
  if (rc) return rc;
  if (!x || !core || !gout || !dx) return dctn_set_error(DCTN_ERR_BAD_ARG, "backward_input: null tensor pointer");
  if ((rc = check_aligned(x, "input")) || (rc = check_aligned(core, "core")) || (rc = check_aligned(gout, "grad_output")) || (rc = check_aligned(dx, "grad_input")) || (rc = check_aligned(ws, "workspace"))) return rc;
  if ((rc = check_ws(pl, B, H, W, DCTN_WS_BACKWARD_INPUT, ws, ws_bytes))) return rc;
  EpsGeom g;
  fill_geom(pl, B, H, W, &g);
  cudaStream_t st = (cudaStream_t)stream;
  int fam = pick_family(pl, g, DCTN_WS_BACKWARD_INPUT);
  if (fam < 0) return dctn_set_error(DCTN_ERR_UNSUPPORTED, "backward_input: requested kernel variant %d does not support this shape (%s)", pl->variant, pl->desc.c_str());
  if (fam == FAM_DIRECT)
    return pl->dtype == DCTN_F32
               ? direct_backward<float>(g, 2, (const float*)x, (const float*)core, (const float*)gout, (float*)dx, ws, st)
               : direct_backward<double>(g, 2, (const double*)x, (const double*)core, (const double*)gout, (double*)dx, ws, st);
  if (fam == FAM_TC)
    return tc_backward_input(g, (const float*)x, (const float*)core, (const float*)gout, (float*)dx, ws, tc_arith(pl, DCTN_WS_BACKWARD_INPUT), st);
  return pl->dtype == DCTN_F32
             ? ffma_backward_input<float>(g, (const float*)x, (const float*)core, (const float*)gout, (float*)dx, ws, st)
             : ffma_backward_input<double>(g, (const double*)x, (const double*)core, (const double*)gout, (double*)dx, ws, st);
}

// ------------------------------------------------------------------------------------------------ statistics
// forward + (sum, sum of squares) of its output accumulated into stats[0..1] (device doubles): the reduction of the
// empirical-std initialisation (dctn/eps.py:163-181) slice by slice, no concatenated output, no second pass from HBM
extern "C" int dctn_eps_forward_stats(const dctn_plan_t* pl, const void* x, const void* core, void* out, double* stats,
                                      int B, int H, int W, void* ws, size_t ws_bytes, void* stream) {
  if (!stats) return dctn_set_error(DCTN_ERR_BAD_ARG, "forward_stats: null stats pointer");
  int rc = check_call(pl, B, H, W);
  if (rc) return rc;
  if ((rc = check_ws(pl, B, H, W, DCTN_WS_FORWARD_STATS, ws, ws_bytes))) return rc;
  const size_t fwd_bytes = (dctn_eps_workspace_bytes(pl, B, H, W, DCTN_WS_FORWARD) + 255) & ~(size_t)255;
  rc = dctn_eps_forward(pl, x, core, out, B, H, W, ws, fwd_bytes, stream);
  if (rc) return rc;
  const long long n = (long long)B * (H - pl->K + 1) * (W - pl->K + 1) * pl->O;
  void* part = (char*)ws + fwd_bytes;
  return pl->dtype == DCTN_F32 ? launch_value_stats<float>((const float*)out, n, stats, part, (cudaStream_t)stream)
                               : launch_value_stats<double>((const double*)out, n, stats, part, (cudaStream_t)stream);
}

extern "C" size_t dctn_window_stats_workspace_bytes(int C, int B, int H, int W) {
  if (C < 1 || B < 1 || H < 1 || W < 1) return 0;
  return window_stats_workspace_bytes(C, B, H, W) + 256;
}

extern "C" int dctn_window_stats(const void* x, int C, int B, int H, int W, int Q, int K, int dtype, double* stats, void* ws,
                                 size_t ws_bytes, void* stream) {
  if (!x || !stats) return dctn_set_error(DCTN_ERR_BAD_ARG, "window_stats: null pointer");
  if (C < 1 || B < 1 || Q < 1 || K < 1 || H < K || W < K)
    return dctn_set_error(DCTN_ERR_BAD_ARG, "window_stats: bad sizes C=%d B=%d H=%d W=%d Q=%d K=%d", C, B, H, W, Q, K);
  if (!ws || ws_bytes < dctn_window_stats_workspace_bytes(C, B, H, W) || ((uintptr_t)ws & 15))
    return dctn_set_error(DCTN_ERR_WORKSPACE, "window_stats: 16-byte aligned workspace of %zu bytes needed, got %zu", dctn_window_stats_workspace_bytes(C, B, H, W), ws_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DCTN_F32) return launch_window_stats<float>((const float*)x, C, B, H, W, Q, K, stats, ws, st);
  if (dtype == DCTN_F64) return launch_window_stats<double>((const double*)x, C, B, H, W, Q, K, stats, ws, st);
  return dctn_set_error(DCTN_ERR_BAD_ARG, "window_stats: bad dtype %d", dtype);
}

// ------------------------------------------------------------------------------------------------ forward from raw pixels
extern "C" int dctn_eps_forward_from_pixels(const dctn_plan_t* pl, const void* pixels, double scale, const void* core, void* out,
                                            int B, int H, int W, void* stream) {
  int rc = check_call(pl, B, H, W);
  if (rc) return rc;
  if (!pixels || !core || !out) return dctn_set_error(DCTN_ERR_BAD_ARG, "forward_from_pixels: null tensor pointer");
  if ((rc = check_aligned(pixels, "pixels")) || (rc = check_aligned(core, "core")) || (rc = check_aligned(out, "output"))) return rc;
  EpsGeom g;
  fill_geom(pl, B, H, W, &g);
  if (!direct_pixels_supported(g, pl->dtype))
    return dctn_set_error(DCTN_ERR_UNSUPPORTED, "forward_from_pixels: the fused feature map exists for K=2, C=1, Q_in=2 layers only (%s)", pl->desc.c_str());
  cudaStream_t st = (cudaStream_t)stream;
  return pl->dtype == DCTN_F32
             ? direct_forward_pixels<float>(g, (const float*)pixels, (float)scale, (const float*)core, (float*)out, st)
             : direct_forward_pixels<double>(g, (const double*)pixels, scale, (const double*)core, (double*)out, st);
}

// ------------------------------------------------------------------------------------------------ training forward
// Only the tcgen05 GEMM family has an intermediate worth keeping (its forward and the second half of its input
// gradient are the same GEMM); every other family reports 0 and the caller uses the plain entry points.
static bool saved_path(const dctn_plan_t* pl, const EpsGeom& g) {
  return pl->dtype == DCTN_F32 && pick_family(pl, g, DCTN_WS_FORWARD) == FAM_TC &&
         pick_family(pl, g, DCTN_WS_BACKWARD_INPUT) == FAM_TC && tc_supported(g, DCTN_WS_BACKWARD_INPUT_SAVED);
}

extern "C" size_t dctn_eps_saved_bytes(const dctn_plan_t* pl, int B, int H, int W) {
  if (check_call(pl, B, H, W)) return 0;
  EpsGeom g;
  fill_geom(pl, B, H, W, &g);
  return saved_path(pl, g) ? tcg_saved_bytes(g) : 0;
}

extern "C" int dctn_eps_forward_train(const dctn_plan_t* pl, const void* x, const void* core, void* out, void* saved,
                                      size_t saved_bytes, int B, int H, int W, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_call(pl, B, H, W);
  if (rc) return rc;
  if (!x || !core || !out || !saved) return dctn_set_error(DCTN_ERR_BAD_ARG, "forward_train: null tensor pointer");
  if ((rc = check_aligned(x, "input")) || (rc = check_aligned(core, "core")) || (rc = check_aligned(out, "output")) || (rc = check_aligned(saved, "saved")) || (rc = check_aligned(ws, "workspace"))) return rc;
  if ((rc = check_ws(pl, B, H, W, DCTN_WS_FORWARD, ws, ws_bytes))) return rc;
  EpsGeom g;
  fill_geom(pl, B, H, W, &g);
  if (!saved_path(pl, g))
    return dctn_set_error(DCTN_ERR_UNSUPPORTED, "forward_train: no savable intermediate for this plan/shape (dctn_eps_saved_bytes() == 0): %s", pl->desc.c_str());
  if (saved_bytes < tcg_saved_bytes(g))
    return dctn_set_error(DCTN_ERR_WORKSPACE, "forward_train: saved buffer of %zu bytes needed, got %zu", tcg_saved_bytes(g), saved_bytes);
  return tc_forward(g, (const float*)x, (const float*)core, (float*)out, ws, tc_arith(pl, DCTN_WS_FORWARD),
                    (cudaStream_t)stream, (float*)saved);
}

extern "C" int dctn_eps_backward_input_saved(const dctn_plan_t* pl, const void* x, const void* core, const void* gout,
                                             const void* saved, size_t saved_bytes, void* dx, int B, int H, int W,
                                             void* ws, size_t ws_bytes, void* stream) {
  int rc = check_call(pl, B, H, W);
  if (rc) return rc;
  if (!x || !core || !gout || !saved || !dx) return dctn_set_error(DCTN_ERR_BAD_ARG, "backward_input_saved: null tensor pointer");
  if ((rc = check_aligned(x, "input")) || (rc = check_aligned(core, "core")) || (rc = check_aligned(gout, "grad_output")) || (rc = check_aligned(saved, "saved")) || (rc = check_aligned(dx, "grad_input")) || (rc = check_aligned(ws, "workspace"))) return rc;
  EpsGeom g;
  fill_geom(pl, B, H, W, &g);
  if (!saved_path(pl, g))
    return dctn_set_error(DCTN_ERR_UNSUPPORTED, "backward_input_saved: no savable intermediate for this plan/shape: %s", pl->desc.c_str());
  if (saved_bytes < tcg_saved_bytes(g))
    return dctn_set_error(DCTN_ERR_WORKSPACE, "backward_input_saved: saved buffer of %zu bytes needed, got %zu", tcg_saved_bytes(g), saved_bytes);
  if ((rc = check_ws(pl, B, H, W, DCTN_WS_BACKWARD_INPUT_SAVED, ws, ws_bytes))) return rc;
  return tc_backward_input_saved(g, (const float*)x, (const float*)core, (const float*)gout, (const float*)saved, (float*)dx,
                                 ws, tc_arith(pl, DCTN_WS_BACKWARD_INPUT), (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------ logmatmulexp
extern "C" int dctn_logmatmulexp_forward(const void* A, const void* B, void* out, int Th, int R, int I, int dtype,
                                         void* stream) {
  if (!A || !B || !out) return dctn_set_error(DCTN_ERR_BAD_ARG, "logmatmulexp: null tensor pointer");
  if (Th < 1 || R < 1 || I < 1) return dctn_set_error(DCTN_ERR_BAD_ARG, "logmatmulexp: sizes must be positive (%d, %d, %d)", Th, R, I);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DCTN_F32) return lme_forward<float>((const float*)A, (const float*)B, (float*)out, Th, R, I, st);
  if (dtype == DCTN_F64) return lme_forward<double>((const double*)A, (const double*)B, (double*)out, Th, R, I, st);
  return dctn_set_error(DCTN_ERR_BAD_ARG, "logmatmulexp: bad dtype %d", dtype);
}

extern "C" int dctn_logmatmulexp_backward(const void* A, const void* B, const void* out, const void* gout, void* dA,
                                          void* dB, int Th, int R, int I, int dtype, void* stream) {
  if (!A || !B || !out || !gout) return dctn_set_error(DCTN_ERR_BAD_ARG, "logmatmulexp backward: null tensor pointer");
  if (Th < 1 || R < 1 || I < 1) return dctn_set_error(DCTN_ERR_BAD_ARG, "logmatmulexp: sizes must be positive (%d, %d, %d)", Th, R, I);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DCTN_F32)
    return lme_backward<float>((const float*)A, (const float*)B, (const float*)out, (const float*)gout, (float*)dA, (float*)dB, Th, R, I, st);
  if (dtype == DCTN_F64)
    return lme_backward<double>((const double*)A, (const double*)B, (const double*)out, (const double*)gout, (double*)dA, (double*)dB, Th, R, I, st);
  return dctn_set_error(DCTN_ERR_BAD_ARG, "logmatmulexp: bad dtype %d", dtype);
}

// one-kernel-per-product path (csrc/logmatmulexp_tile.cu); 0 bytes = shape not served, use the entries above
extern "C" size_t dctn_logmatmulexp_workspace_bytes(int Th, int R, int I, int dtype) {
  if (Th < 1 || R < 1 || I < 1) return 0;
  if (dtype == DCTN_F32) return lme_tile_supported<float>(Th, R, I) ? lme_tile_workspace_bytes(Th, I, 4) : 0;
  if (dtype == DCTN_F64) return lme_tile_supported<double>(Th, R, I) ? lme_tile_workspace_bytes(Th, I, 8) : 0;
  return 0;
}

extern "C" int dctn_logmatmulexp_forward_ws(const void* A, const void* B, void* out, int Th, int R, int I, int dtype, void* ws,
                                            size_t ws_bytes, void* stream) {
  if (!A || !B || !out) return dctn_set_error(DCTN_ERR_BAD_ARG, "logmatmulexp: null tensor pointer");
  const size_t need = dctn_logmatmulexp_workspace_bytes(Th, R, I, dtype);
  if (need == 0) return dctn_set_error(DCTN_ERR_UNSUPPORTED, "logmatmulexp_forward_ws: shape (%d, %d, %d) is served by dctn_logmatmulexp_forward only", Th, R, I);
  if (!ws || ws_bytes < need || ((uintptr_t)ws & 15)) return dctn_set_error(DCTN_ERR_WORKSPACE, "logmatmulexp: 16-byte aligned workspace of %zu bytes needed, got %zu", need, ws_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == DCTN_F32 ? lme_tile_forward<float>((const float*)A, (const float*)B, (float*)out, Th, R, I, ws, st)
                           : lme_tile_forward<double>((const double*)A, (const double*)B, (double*)out, Th, R, I, ws, st);
}

extern "C" int dctn_logmatmulexp_backward_ws(const void* A, const void* B, const void* out, const void* gout, void* dA, void* dB,
                                             int Th, int R, int I, int dtype, const void* ws, size_t ws_bytes, void* stream) {
  if (!A || !B || !out || !gout) return dctn_set_error(DCTN_ERR_BAD_ARG, "logmatmulexp backward: null tensor pointer");
  const size_t need = dctn_logmatmulexp_workspace_bytes(Th, R, I, dtype);
  if (need == 0) return dctn_set_error(DCTN_ERR_UNSUPPORTED, "logmatmulexp_backward_ws: shape (%d, %d, %d) is served by dctn_logmatmulexp_backward only", Th, R, I);
  if (!ws || ws_bytes < need || ((uintptr_t)ws & 15)) return dctn_set_error(DCTN_ERR_WORKSPACE, "logmatmulexp: the workspace of the forward call (%zu bytes) is needed, got %zu", need, ws_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == DCTN_F32
             ? lme_tile_backward<float>((const float*)A, (const float*)B, (const float*)out, (const float*)gout, (float*)dA, (float*)dB, Th, R, I, ws, st)
             : lme_tile_backward<double>((const double*)A, (const double*)B, (const double*)out, (const double*)gout, (double*)dA, (double*)dB, Th, R, I, ws, st);
}

extern "C" int dctn_logmatmulexp_batched_forward(const void* A, const void* B, void* out, long long batch, int Th, int R,
                                                 int I, int dtype, void* stream) {
  if (!A || !B || !out) return dctn_set_error(DCTN_ERR_BAD_ARG, "logmatmulexp_batched: null tensor pointer");
  if (batch < 1 || Th < 1 || R < 1 || I < 1)
    return dctn_set_error(DCTN_ERR_BAD_ARG, "logmatmulexp_batched: sizes must be positive (%lld, %d, %d, %d)", batch, Th, R, I);
  { int rc; if ((rc = check_aligned(out, "output"))) return rc; }   // the float32 register-tiled kernels store 128 bits at a time
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DCTN_F32) return lme_batched_forward<float>((const float*)A, (const float*)B, (float*)out, batch, Th, R, I, st);
  if (dtype == DCTN_F64) return lme_batched_forward<double>((const double*)A, (const double*)B, (double*)out, batch, Th, R, I, st);
  return dctn_set_error(DCTN_ERR_BAD_ARG, "logmatmulexp_batched: bad dtype %d", dtype);
}

extern "C" int dctn_logmatmulexp_batched_backward(const void* A, const void* B, const void* out, const void* gout,
                                                  void* dA, void* dB, long long batch, int Th, int R, int I, int dtype,
                                                  void* stream) {
  if (!A || !B || !out || !gout) return dctn_set_error(DCTN_ERR_BAD_ARG, "logmatmulexp_batched backward: null tensor pointer");
  if (batch < 1 || Th < 1 || R < 1 || I < 1)
    return dctn_set_error(DCTN_ERR_BAD_ARG, "logmatmulexp_batched: sizes must be positive (%lld, %d, %d, %d)", batch, Th, R, I);
  { int rc; if ((dA && (rc = check_aligned(dA, "grad_A"))) || (dB && (rc = check_aligned(dB, "grad_B")))) return rc; }
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DCTN_F32)
    return lme_batched_backward<float>((const float*)A, (const float*)B, (const float*)out, (const float*)gout, (float*)dA, (float*)dB, batch, Th, R, I, st);
  if (dtype == DCTN_F64)
    return lme_batched_backward<double>((const double*)A, (const double*)B, (const double*)out, (const double*)gout, (double*)dA, (double*)dB, batch, Th, R, I, st);
  return dctn_set_error(DCTN_ERR_BAD_ARG, "logmatmulexp_batched: bad dtype %d", dtype);
}

// ------------------------------------------------------------------------------------------------ host-buffer entry
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" size_t dctn_eps_forward_host_device_bytes(const dctn_plan_t* pl, int B, int H, int W) {
  if (check_call(pl, B, H, W)) return 0;
  size_t es = pl->dtype == DCTN_F32 ? 4 : 8;
  size_t xb = (size_t)pl->C * B * H * W * pl->Q * es;
  size_t cb = (size_t)pl->D * pl->O * es;
  size_t ob = (size_t)B * (H - pl->K + 1) * (W - pl->K + 1) * pl->O * es;
  return align256(xb) + align256(cb) + align256(ob) + dctn_eps_workspace_bytes(pl, B, H, W, DCTN_WS_FORWARD);
}

extern "C" int dctn_eps_forward_host(const dctn_plan_t* pl, const void* x_host, const void* core_host, void* out_host,
                                     int B, int H, int W, void* scratch, size_t scratch_bytes, void* stream) {
  int rc = check_call(pl, B, H, W);
  if (rc) return rc;
  if (!x_host || !core_host || !out_host) return dctn_set_error(DCTN_ERR_BAD_ARG, "forward_host: null host pointer");
  size_t need = dctn_eps_forward_host_device_bytes(pl, B, H, W);
  if (!scratch || scratch_bytes < need)
    return dctn_set_error(DCTN_ERR_WORKSPACE, "forward_host: device scratch of %zu bytes needed, got %zu", need, scratch_bytes);
  size_t es = pl->dtype == DCTN_F32 ? 4 : 8;
  size_t xb = (size_t)pl->C * B * H * W * pl->Q * es;
  size_t cb = (size_t)pl->D * pl->O * es;
  size_t ob = (size_t)B * (H - pl->K + 1) * (W - pl->K + 1) * pl->O * es;
  char* base = (char*)scratch;
  void* xd = base;
  void* cd = base + align256(xb);
  void* od = base + align256(xb) + align256(cb);
  void* ws = base + align256(xb) + align256(cb) + align256(ob);
  size_t wsb = scratch_bytes - (align256(xb) + align256(cb) + align256(ob));
  cudaStream_t st = (cudaStream_t)stream;
  DCTN_CUDA_CHECK_RET(cudaMemcpyAsync(xd, x_host, xb, cudaMemcpyHostToDevice, st));
  DCTN_CUDA_CHECK_RET(cudaMemcpyAsync(cd, core_host, cb, cudaMemcpyHostToDevice, st));
  rc = dctn_eps_forward(pl, xd, cd, od, B, H, W, ws, wsb, stream);
  if (rc) return rc;
  DCTN_CUDA_CHECK_RET(cudaMemcpyAsync(out_host, od, ob, cudaMemcpyDeviceToHost, st));
  DCTN_CUDA_CHECK_RET(cudaStreamSynchronize(st));
  return 0;
}
