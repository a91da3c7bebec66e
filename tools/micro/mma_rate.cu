// Micro-benchmark: issue-rate of tcgen05.mma for a few shapes/kinds from fixed smem tiles.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
template <int KIND>  // 0 = tf32, 1 = f16(bf16)
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int KIND, int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long* cycles_out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + N) * 128 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_d = tmem_base_s;
  // idesc: c=f32; a/b format tf32(2) or bf16(1); n_dim, m_dim=128
  const uint32_t fmt = KIND == 0 ? 2u : 1u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((128 >> 4) << 24);
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 128 * 128);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) mma<KIND>(tmem_d, make_desc(a + kk * 32), make_desc(b + kk * 32), idesc, 1);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
    t1 = clock64();
    if (blockIdx.x == 0) *cycles_out = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(256));
}

template <int KIND, int N>
void run(const char* name, int kelems) {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 2000, smem = (128 + N) * 128;
  cudaFuncSetAttribute(rate_kernel<KIND, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate_kernel<KIND, N><<<148, 128, smem>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  double per = (double)c / (iters * 4);
  printf("%-28s %s cycles/MMA %.1f  -> %.0f MAC/clk/SM\n", name, cudaGetErrorString(e), per, 128.0 * N * kelems / per);
}
int main() {
  run<0, 128>("tf32 128x128x8", 8);
  run<0, 256>("tf32 128x256x8", 8);
  run<1, 128>("bf16 128x128x16", 16);
  run<1, 256>("bf16 128x256x16", 16);
  return 0;
}
