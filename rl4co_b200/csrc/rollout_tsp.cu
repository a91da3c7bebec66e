// TSP instantiations of the persistent rollout kernels: head-wise (rollout_hw_impl.cuh, default) and the
// round-1 kernel (rollout_impl.cuh, CO_ROLLOUT_IMPL=v3).
#include "rollout_hw_impl.cuh"
namespace co {
int rollout_tsp(const co_rollout_args& A, cudaStream_t st) { return hw::dispatch<CO_ENV_TSP>(A, st); }
int rollout_v3_tsp(const co_rollout_args& A, cudaStream_t st) { return dispatch<CO_ENV_TSP>(A, st); }
}  // namespace co
