// CVRP instantiations of the persistent rollout kernel (see rollout_impl.cuh).
#include "rollout_impl.cuh"
namespace co {
int rollout_cvrp(const co_rollout_args& A, cudaStream_t st) { return dispatch<CO_ENV_CVRP>(A, st); }
}  // namespace co
