// tcgen05 TF32 family — placeholder until the kernels land (reports "unsupported" so AUTO never picks it).
#include "common.cuh"
#include "eps_kernels.h"
bool tc_supported(const EpsGeom&, int) { return false; }
size_t tc_workspace_bytes(const EpsGeom&, int) { return 0; }
int tc_forward(const EpsGeom&, const float*, const float*, float*, void*, int, cudaStream_t) { return dctn_set_error(-2, "tcgen05 family not built"); }
int tc_backward_core(const EpsGeom&, const float*, const float*, float*, void*, int, cudaStream_t) { return dctn_set_error(-2, "tcgen05 family not built"); }
int tc_backward_input(const EpsGeom&, const float*, const float*, const float*, float*, void*, int, cudaStream_t) { return dctn_set_error(-2, "tcgen05 family not built"); }
