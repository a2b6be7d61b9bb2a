// Streaming thread-per-patch family — placeholder until the kernels land.
#include "common.cuh"
#include "eps_kernels.h"
bool direct_supported(const EpsGeom&, int) { return false; }
template <typename T> int direct_forward(const EpsGeom&, const T*, const T*, T*, cudaStream_t) { return dctn_set_error(-2, "direct family not built"); }
template int direct_forward<float>(const EpsGeom&, const float*, const float*, float*, cudaStream_t);
template int direct_forward<double>(const EpsGeom&, const double*, const double*, double*, cudaStream_t);
