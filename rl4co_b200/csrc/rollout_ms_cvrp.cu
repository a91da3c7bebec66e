// CVRP instantiations of the query-batched multistart rollout kernel (see rollout_ms_impl.cuh).
#include "rollout_ms_impl.cuh"
namespace co {
int rollout_ms_cvrp(const co_rollout_args& A, cudaStream_t st) { return dispatch_ms<CO_ENV_CVRP>(A, st); }
}  // namespace co
