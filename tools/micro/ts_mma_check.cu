// Micro-check for next round's kernels: (1) does tcgen05.mma kind::tf32 accept its A operand from TMEM (TS form) with
// the "lane = row, column = k" layout written by tcgen05.st 32x32b, (2) what do a TS-form MMA (N = 16 / 32 / 128,
// K = 8) and a tcgen05.st.32x32b.x32 cost. One CTA, 128 threads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/ts_mma_check tools/micro/ts_mma_check.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

constexpr int KT = 64;  // K of the checked product (8 k-steps)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t sbo) {  // K-major, no swizzle, LBO 128 B
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(128 >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__host__ __device__ constexpr uint32_t idesc(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a_tmem), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit_wait(uint32_t bar, uint32_t parity) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// A [128][KT] and B [16][KT] hold tf32-representable values. out [128][16] = A B^T computed with A from TMEM.
// cyc[0..2] = cycles per TS-form MMA with N = 16 / 32 / 128 (K = 8), cyc[3] = cycles per tcgen05.st.x32 (4 warps busy)
__global__ void __launch_bounds__(128, 1) ts_kernel(const float* A, const float* B, float* out, long long* cyc, int iters) {
  extern __shared__ __align__(1024) unsigned char smem[];  // B tile: 16 (up to 128) rows x KT, K-major no-swizzle
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr uint32_t SBO = (KT / 4) * 128;  // 8-row groups are (KT/4) 16-byte chunks x 8 rows apart
  for (int i = tid; i < 128 * KT; i += 128) {  // rows >= 16 (only used by the timing loops) are zero
    const int n = i / KT, k = i % KT;
    const float v = n < 16 ? B[n * KT + k] : 0.f;
    *reinterpret_cast<float*>(smem + (n / 8) * SBO + (k / 4) * 128 + (n % 8) * 16 + (k % 4) * 4) = v;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = tmem + ((uint32_t)(32 * warp) << 16);

  // A row `tid` -> TMEM columns 0..KT-1 of lane `tid`
  for (int c = 0; c < KT; c += 8) tmem_st8(lane_base + c, A + tid * KT + c);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t phase = 0;
  if (tid == 0) {
    for (int ks = 0; ks < KT / 8; ++ks)  // D at columns 256..271
      mma_ts(tmem + 256, tmem + ks * 8, desc(s32(smem) + ks * 256, SBO), idesc(16), ks > 0);
    commit_wait(s32(&mbar), phase);
  }
  phase ^= 1;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t r[16];
  tmem_ld16(lane_base + 256, r);
  for (int c = 0; c < 16; ++c) out[tid * 16 + c] = __uint_as_float(r[c]);

  // ---- timing: TS-form MMAs
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (tid == 0) {
    const int ns[3] = {16, 32, 128};
    for (int v = 0; v < 3; ++v) {
      const uint32_t id = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ns[v] >> 3) << 17) | ((128u >> 4) << 24);
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) mma_ts(tmem + 256, tmem + ks * 8, desc(s32(smem) + ks * 256, SBO), id, 1);
      commit_wait(s32(&mbar), phase);
      phase ^= 1;
      cyc[v] = (clock64() - t0) / (8LL * iters);
    }
  }
  __syncthreads();
  // ---- timing: tcgen05.st.x32, all four warps
  uint32_t q[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) q[i] = tid * 32 + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    tmem_st32(lane_base + 64, q);
    tmem_st32(lane_base + 96, q);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  const long long t1 = clock64();
  if (tid == 0) cyc[3] = (t1 - t0) / (2LL * iters);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

static float to_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

int main() {
  std::vector<float> A(128 * KT), B(16 * KT), ref(128 * 16), out(128 * 16);
  srand(1);
  for (auto& v : A) v = to_tf32((float)rand() / RAND_MAX * 2.f - 1.f);
  for (auto& v : B) v = to_tf32((float)rand() / RAND_MAX * 2.f - 1.f);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 16; ++n) {
      double acc = 0;
      for (int k = 0; k < KT; ++k) acc += (double)A[m * KT + k] * B[n * KT + k];
      ref[m * 16 + n] = (float)acc;
    }
  float *dA, *dB, *dO;
  long long* dC;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, out.size() * 4); cudaMalloc(&dC, 4 * 8);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const int smem = 128 * KT * 4;
  cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  ts_kernel<<<1, 128, smem>>>(dA, dB, dO, dC, 256);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  long long cyc[4];
  cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(cyc, dC, sizeof(cyc), cudaMemcpyDeviceToHost);
  double err = 0, mag = 0;
  for (size_t i = 0; i < out.size(); ++i) { err = fmax(err, fabs((double)out[i] - ref[i])); mag = fmax(mag, fabs((double)ref[i])); }
  printf("TS-form tf32 MMA (A from TMEM, lane = row, column = k): max |err| = %.3e (max |ref| = %.3f) -> %s\n", err, mag,
         err < 1e-4 ? "layout confirmed" : "MISMATCH");
  printf("cycles per TS-form MMA 128xNx8: N=16 %lld, N=32 %lld, N=128 %lld; tcgen05.st.32x32b.x32: %lld cycles per warp-instruction (4 warps)\n",
         cyc[0], cyc[1], cyc[2], cyc[3]);
  return 0;
}
