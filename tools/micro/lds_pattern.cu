// LDS.128 wavefront cost for the broadcast patterns of the rollout kernel's logits phase: every lane of a warp reads
// one of TWO 16-byte chunks.  Prints cycles per warp-instruction with 8 warps issuing back to back (LSU-bound).
#include <cstdio>
#include <cstdint>
__global__ void k(int pattern, int stride_b, float* out, long long* cyc) {
  __shared__ __align__(16) float buf[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) buf[i] = (float)i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int sel = 0;
  if (pattern == 1) sel = lane & 1;        // alternating lanes
  else if (pattern == 2) sel = lane >> 4;  // half warps
  else if (pattern == 3) sel = (lane >> 3) & 1;  // quarter warps alternate
  else if (pattern == 4) sel = lane & 3;   // four chunks, alternating
  else if (pattern == 5) sel = lane >> 3;  // four chunks, quarter warps
  // chunk `sel` of instruction c: far-apart chunks (stride_b = 272 B, the padded layout of the kernel) advance by 16 B per
  // instruction; adjacent chunks (stride_b = 16) are interleaved, so an instruction's group of chunks advances as a whole
  const int groups = (pattern == 4 || pattern == 5) ? 4 : (pattern == 0 ? 1 : 2);
  const uint32_t step = (stride_b == 16) ? 16u * groups : 16u;
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(buf) + sel * stride_b;
  float4 acc = make_float4(0, 0, 0, 0);
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 64; ++it) {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr + c * step));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  long long t1 = clock64();
  __syncthreads();
  out[threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  struct { int p, s; const char* name; } cases[] = {
    {0, 0, "all lanes one chunk (pure broadcast)"},
    {1, 272, "two chunks 272 B apart, lane&1"}, {2, 272, "two chunks 272 B apart, lane>>4"}, {3, 272, "two chunks 272 B apart, (lane>>3)&1"},
    {1, 16, "two adjacent chunks (interleaved), lane&1"}, {2, 16, "two adjacent chunks (interleaved), lane>>4"},
    {4, 16, "four adjacent chunks, lane&3"}, {5, 16, "four adjacent chunks, lane>>3"},
    {4, 272, "four chunks 272 B apart, lane&3"}, {5, 272, "four chunks 272 B apart, lane>>3"},
  };
  for (auto& c : cases) {
    for (int warps : {1, 8}) {
      k<<<1, 32 * warps>>>(c.p, c.s, out, cyc);
      long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-52s warps %d: %.2f cycles per LDS.128 per warp (%.2f per SM-instruction)\n", c.name, warps, (double)h / 1024, (double)h / 1024 / warps);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
