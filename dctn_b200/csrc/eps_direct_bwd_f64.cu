// Streaming small-core family, backward, double instances (see eps_direct_impl.cuh).
#define DCTN_DIRECT_PART 4
#include "eps_direct_impl.cuh"
