// Micro-benchmark: tcgen05.ld throughput / latency (TMEM -> registers), 8 warps per CTA, 1 CTA/SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ uint32_t ld_sum(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ld_sum<16>(uint32_t taddr) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s ^= r[i];
  return s;
}
template <>
__device__ __forceinline__ uint32_t ld_sum<32>(uint32_t taddr) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s ^= r[i];
  return s;
}

template <int X>
__global__ void __launch_bounds__(256, 1) k(int iters, long long* out, uint32_t* sink, int active_warps) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = tmem_base_s + ((uint32_t)(32 * (warp & 3)) << 16) + (warp >> 2) * 256;
  uint32_t acc = 0;
  long long t0 = clock64();
  if (warp < active_warps)
    for (int it = 0; it < iters; ++it) acc ^= ld_sum<X>(base + ((it * X) & 255) % (256 - X + 1));
  long long t1 = clock64();
  if (acc == 0x12345) sink[0] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "n"(512));
}
template <int X>
void run(int warps) {
  long long* d; uint32_t* s; cudaMalloc(&d, 8); cudaMalloc(&s, 4);
  const int iters = 4000;
  k<X><<<148, 256>>>(iters, d, s, warps);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  double per = (double)c / iters;
  printf("x%d, %d warps: %s  %.1f cycles per ld+wait per warp -> %.0f B/clk/SM\n", X, warps, cudaGetErrorString(e), per,
         warps * 32.0 * X * 4 / per);
}
int main() {
  run<16>(1); run<16>(4); run<16>(8); run<32>(1); run<32>(4); run<32>(8);
  return 0;
}
