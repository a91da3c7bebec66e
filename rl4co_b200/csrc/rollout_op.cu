// Persistent rollout kernel, orienteering instantiation (budget mask in the mask functor, prize sum as reward).
#include "rollout_impl.cuh"
namespace co {
int rollout_op(const co_rollout_args& A, cudaStream_t st) { return dispatch<CO_ENV_OP>(A, st); }
}  // namespace co
