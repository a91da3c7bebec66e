// Persistent rollout kernel, prize-collecting TSP instantiation (depot rule on the collected prize in the mask functor,
// remaining-prize context, reward = saved penalties - tour length - all penalties).
#include "rollout_impl.cuh"
namespace co {
int rollout_pctsp(const co_rollout_args& A, cudaStream_t st) { return dispatch<CO_ENV_PCTSP>(A, st); }
}  // namespace co
