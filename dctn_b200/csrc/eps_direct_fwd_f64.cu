// Streaming small-core family, forward, double instances (see eps_direct_impl.cuh).
#define DCTN_DIRECT_PART 2
#include "eps_direct_impl.cuh"
