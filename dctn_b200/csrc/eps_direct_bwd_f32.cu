// Streaming small-core family, backward, float instances (see eps_direct_impl.cuh).
#define DCTN_DIRECT_PART 3
#include "eps_direct_impl.cuh"
