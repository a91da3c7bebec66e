// Data path either side of the rollout: on-device instance generation and the dihedral-8
// augmentation as single streaming kernels (HBM-bound: 4 B written per generated float, 8 B read + 64 B
// written per node for the augmentation).
//   co_generate_uniform   <- rl4co/envs/common/utils.py:61-62 (get_sampler "uniform": torch.rand * (hi-lo) + lo),
//                            as used by tsp/generator.py:49-58 (locs) and cvrp/generator.py:114-123 (depot + locs)
//   co_generate_demand    <- rl4co/envs/routing/cvrp/generator.py:126-137: (floor(U[lo,hi)) + 1) / capacity with
//                            lo = min_demand - 1, hi = max_demand - 1
//   co_dihedral8          <- rl4co/data/transforms.py:16-38 (aug-major: row a*B + b)
// The random stream is Philox4x32-10 keyed by (seed, offset): same seed -> same instances on every run and on every
// GPU; it is NOT torch's CPU stream, so generated instances are distributionally, not bit-wise, the reference's.
#include "co_common.cuh"

namespace co {

__device__ __forceinline__ float u01(uint32_t r) {  // 24 random bits -> [0, 1), every value exactly representable
  return (float)(r >> 8) * (1.0f / 16777216.0f);
}

// 4 floats per Philox call, 4 calls per thread iteration -> 16-byte stores, grid-stride
__global__ void __launch_bounds__(256) generate_uniform_kernel(float* __restrict__ out, long n, uint64_t seed,
                                                                uint64_t offset, float lo, float hi) {
  const long n4 = n >> 2;
  const float span = hi - lo;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)), key);
    reinterpret_cast<float4*>(out)[i] =
        make_float4(u01(r.x) * span + lo, u01(r.y) * span + lo, u01(r.z) * span + lo, u01(r.w) * span + lo);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {  // tail (n not a multiple of 4)
    const long i = n4;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)), key);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    out[4 * n4 + threadIdx.x] = u01(w[threadIdx.x]) * span + lo;
  }
}

__global__ void __launch_bounds__(256) generate_demand_kernel(float* __restrict__ out, long n, uint64_t seed,
                                                               uint64_t offset, float lo, float hi, float capacity) {
  const float span = hi - lo;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < ((n + 3) >> 2); i += (long)gridDim.x * blockDim.x) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)), key);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long k = 4 * i + j;
      // (demand.int() + 1).float() / capacity : truncation towards zero of a non-negative value
      if (k < n) out[k] = ((float)((int)(u01(w[j]) * span + lo) + 1)) / capacity;
    }
  }
}

// one thread per (instance, node): reads (x, y) once, writes the 8 images; writes are float2-coalesced per image
__global__ void __launch_bounds__(256) dihedral8_kernel(const float2* __restrict__ xy, float2* __restrict__ out, long BN) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < BN; i += (long)gridDim.x * blockDim.x) {
    const float2 p = xy[i];
    const float x = p.x, y = p.y, rx = 1.0f - p.x, ry = 1.0f - p.y;
    out[0 * BN + i] = make_float2(x, y);
    out[1 * BN + i] = make_float2(rx, y);
    out[2 * BN + i] = make_float2(x, ry);
    out[3 * BN + i] = make_float2(rx, ry);
    out[4 * BN + i] = make_float2(y, x);
    out[5 * BN + i] = make_float2(ry, x);
    out[6 * BN + i] = make_float2(y, rx);
    out[7 * BN + i] = make_float2(ry, rx);
  }
}

}  // namespace co

using namespace co;

static inline int stream_grid(long work) {
  long g = (work + 255) / 256;
  const long cap = (long)device_info().sm_count * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

extern "C" int co_generate_uniform(float* out, long n, uint64_t seed, uint64_t offset, float lo, float hi, void* stream) {
  if (!out) return fail(CO_ERR_BAD_ARG, "co_generate_uniform: null pointer%s");
  if (n < 0 || !(hi >= lo)) return fail(CO_ERR_BAD_ARG, "co_generate_uniform: bad range or size%s");
  if ((uintptr_t)out & 15) return fail(CO_ERR_BAD_ARG, "co_generate_uniform: output must be 16-byte aligned%s");
  if (n == 0) return CO_OK;
  generate_uniform_kernel<<<stream_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(out, n, seed, offset, lo, hi);
  return check_launch("co_generate_uniform");
}

extern "C" int co_generate_demand(float* out, long n, uint64_t seed, uint64_t offset, int min_demand, int max_demand,
                                  float capacity, void* stream) {
  if (!out) return fail(CO_ERR_BAD_ARG, "co_generate_demand: null pointer%s");
  if (n < 0 || min_demand < 1 || max_demand < min_demand || !(capacity > 0.f))
    return fail(CO_ERR_BAD_ARG, "co_generate_demand: bad arguments%s");
  if (n == 0) return CO_OK;
  generate_demand_kernel<<<stream_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(
      out, n, seed, offset, (float)(min_demand - 1), (float)(max_demand - 1), capacity);
  return check_launch("co_generate_demand");
}

extern "C" int co_dihedral8(const float* locs, float* out, long B, int N, void* stream) {
  if (!locs || !out) return fail(CO_ERR_BAD_ARG, "co_dihedral8: null pointer%s");
  if (B < 0 || N < 1) return fail(CO_ERR_BAD_ARG, "co_dihedral8: bad shape%s");
  if (B == 0) return CO_OK;
  dihedral8_kernel<<<stream_grid(B * N), 256, 0, (cudaStream_t)stream>>>((const float2*)locs, (float2*)out, B * (long)N);
  return check_launch("co_dihedral8");
}
