// SDVRP instantiations of the persistent rollout kernel (see rollout_impl.cuh): the first sibling env behind the
// kernel's mask functor -- split deliveries (rl4co/envs/routing/sdvrp/env.py:55-116) with the
// SDVRPDynamicEmbedding (nn/env_embeddings/dynamic.py:60-78) folded into per-step scalars.
#include "rollout_impl.cuh"
namespace co {
int rollout_sdvrp(const co_rollout_args& A, cudaStream_t st) { return dispatch<CO_ENV_SDVRP>(A, st); }
}  // namespace co
