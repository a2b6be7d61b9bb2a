// Streaming small-core family, forward, float instances (see eps_direct_impl.cuh).
#define DCTN_DIRECT_PART 1
#include "eps_direct_impl.cuh"
