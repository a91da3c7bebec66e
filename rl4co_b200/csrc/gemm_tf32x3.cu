// fp32-accurate GEMM on Blackwell tensor cores (tcgen05, kind::tf32, 3xTF32 split).
//
//   C[M, Nout] = epilogue( A[M, K] @ W[Nout, K]^T )        A, W, C: fp32 row-major
//   epilogue(v) = ((v + bias[n]) (+ residual[m, n])) (relu) * scale[n] + shift[n]
//
// Serves the dense projections either side of the rollout loop:
//   * FusedAttentionModelDecoder._precompute_cache (rl4co/models/zoo/am/decoder.py:201-228):
//     embeddings[B*N,128] @ Wcat[640,128]^T -> fused rollout cache;
//   * the AM encoder's Linear layers (rl4co/models/nn/attention.py:110-134, nn/mlp.py:45-60,
//     nn/graph/attnnet.py:45-52) with bias / ReLU / skip connection / eval-mode BatchNorm folded
//     into the epilogue.
//
// Numerics: each fp32 operand x is split into hi = rna_tf32(x) and lo = x - hi; the kernel
// accumulates hi*hi + lo*hi + hi*lo in fp32 TMEM accumulators (the dropped lo*lo term is
// ~2^-22 relative), i.e. fp32-class accuracy (~1e-6 relative) at tensor-core rate.
//
// Structure (one CTA per 128x128 output tile, 256 threads, 64 KB smem -> 3 CTAs/SM overlap each
// other's load / MMA / epilogue phases):
//   for each 32-wide k-block:  all warps LDG A / W_hi / W_lo -> split A -> STS in the UMMA
//   canonical K-major no-swizzle layout ((8 rows x 16 B) core matrices, LBO = 128 B, SBO = 1 KB)
//   -> fence.proxy.async -> one elected thread issues 12 tcgen05.mma (4 k-steps x 3 products)
//   -> tcgen05.commit -> mbarrier wait.  Epilogue: 4 warps tcgen05.ld 32x32b.x32 their TMEM
//   lanes, apply the fused epilogue, 128-bit stores.
#include "co_common.cuh"

namespace co {

constexpr int GM = 128, GN = 128, GK = 32;
constexpr int TILE_BYTES = GM * GK * 4;  // 16 KB per operand tile
constexpr uint32_t LBO = 128, SBO = 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  // cute::UMMA::SmemDescriptor (mma_sm100_desc.hpp): start>>4 [0,14), LBO>>4 [16,30),
  // SBO>>4 [32,46), version=1 [46,48), layout_type=SWIZZLE_NONE [61,64)
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(LBO >> 4) << 16;
  d |= (uint64_t)(SBO >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// cute::UMMA::InstrDescriptor: c_format=F32 (1<<4), a/b_format=TF32 (2<<7, 2<<10), K-major both,
// n_dim = N>>3 at [17,23), m_dim = M>>4 at [24,29)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((GN >> 3) << 17) | ((GM >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ float rna_tf32(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}
__device__ __forceinline__ float4 split_hi(float4 v) {
  return make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
}

__global__ void split_tf32_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, long n) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float v = w[i], h = rna_tf32(v);
    hi[i] = h;
    lo[i] = v - h;
  }
}

struct GemmArgs {
  const float* A; const float* Whi; const float* Wlo; float* C;
  const float* bias; const float* residual; const float* scale; const float* shift;
  int M, Nout, K, lda, ldc, ldr, relu, n_tiles;
};

__global__ void __launch_bounds__(256, 3) gemm_tf32x3_kernel(const GemmArgs g) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sAhi = smem;
  unsigned char* sAlo = smem + TILE_BYTES;
  unsigned char* sBhi = smem + 2 * TILE_BYTES;
  unsigned char* sBlo = smem + 3 * TILE_BYTES;
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tile = blockIdx.x % g.n_tiles, m_tile = blockIdx.x / g.n_tiles;  // N fastest: A tile reused from L2
  const int m0 = m_tile * GM, n0 = n_tile * GN;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(GN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_d = tmem_base_s;

  // load mapping: a warp covers 8 rows x 4 chunks(16 B) -> 512 contiguous smem bytes, 8 x 64 B global segments
  const int r8 = lane & 7, c4 = lane >> 3;
  uint32_t phase = 0;
  const int nkb = g.K / GK;
  for (int kb = 0; kb < nkb; ++kb) {
    const int k0 = kb * GK;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = warp + 8 * j;              // 32 (row-group, chunk-half) items per tile
      const int rg = q >> 1, chunk = 4 * (q & 1) + c4;
      const int row = 8 * rg + r8;
      const uint32_t soff = rg * SBO + chunk * LBO + r8 * 16;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), bh = a, bl = a;
      if (m0 + row < g.M) a = __ldg(reinterpret_cast<const float4*>(g.A + (size_t)(m0 + row) * g.lda + k0 + chunk * 4));
      if (n0 + row < g.Nout) {
        bh = __ldg(reinterpret_cast<const float4*>(g.Whi + (size_t)(n0 + row) * g.K + k0 + chunk * 4));
        bl = __ldg(reinterpret_cast<const float4*>(g.Wlo + (size_t)(n0 + row) * g.K + k0 + chunk * 4));
      }
      const float4 ah = split_hi(a);
      const float4 al = make_float4(a.x - ah.x, a.y - ah.y, a.z - ah.z, a.w - ah.w);
      *reinterpret_cast<float4*>(sAhi + soff) = ah;
      *reinterpret_cast<float4*>(sAlo + soff) = al;
      *reinterpret_cast<float4*>(sBhi + soff) = bh;
      *reinterpret_cast<float4*>(sBlo + soff) = bl;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> async-proxy (MMA) reads
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t ahi = smem_u32(sAhi), alo = smem_u32(sAlo), bhi = smem_u32(sBhi), blo = smem_u32(sBlo);
#pragma unroll
      for (int kk = 0; kk < GK / 8; ++kk) {     // UMMA_K = 8 tf32 = 2 core matrices along K
        const uint32_t off = kk * 2 * LBO;
        mma_tf32(tmem_d, make_desc(ahi + off), make_desc(bhi + off), (kb | kk) != 0);
        mma_tf32(tmem_d, make_desc(alo + off), make_desc(bhi + off), 1);
        mma_tf32(tmem_d, make_desc(ahi + off), make_desc(blo + off), 1);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    // wait until the MMAs have consumed this stage (also: accumulator complete after the last one)
    {
      uint32_t done = 0;
      const uint32_t bar = smem_u32(&mbar);
      while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(phase) : "memory");
      }
      phase ^= 1;
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");

  // ---- epilogue: warps 0..3 own TMEM lanes 32w..32w+31 = output rows
  if (warp < 4) {
    const int row = m0 + 32 * warp + lane;
    const bool row_ok = row < g.M;
#pragma unroll 1
    for (int cc = 0; cc < GN / 32; ++cc) {
      uint32_t r[32];
      const uint32_t taddr = tmem_d + ((uint32_t)(32 * warp) << 16) + cc * 32;
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
      const int nb = n0 + cc * 32;
      if (row_ok) {
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          const int n = nb + 4 * v;
          if (n < g.Nout) {
            float o[4] = {__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]), __uint_as_float(r[4 * v + 2]),
                          __uint_as_float(r[4 * v + 3])};
            if (g.bias) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + n));
              o[0] += b.x; o[1] += b.y; o[2] += b.z; o[3] += b.w;
            }
            if (g.residual) {
              const float4 x = __ldg(reinterpret_cast<const float4*>(g.residual + (size_t)row * g.ldr + n));
              o[0] += x.x; o[1] += x.y; o[2] += x.z; o[3] += x.w;
            }
            if (g.relu) {
#pragma unroll
              for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
            }
            if (g.scale) {
              const float4 s = __ldg(reinterpret_cast<const float4*>(g.scale + n));
              const float4 t = __ldg(reinterpret_cast<const float4*>(g.shift + n));
              o[0] = fmaf(o[0], s.x, t.x); o[1] = fmaf(o[1], s.y, t.y); o[2] = fmaf(o[2], s.z, t.z); o[3] = fmaf(o[3], s.w, t.w);
            }
            *reinterpret_cast<float4*>(g.C + (size_t)row * g.ldc + n) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(GN));
  }
}

}  // namespace co

using namespace co;

extern "C" int co_split_tf32(const float* w, float* hi, float* lo, long n, void* stream) {
  if (!w || !hi || !lo || n < 0) return fail(CO_ERR_BAD_ARG, "co_split_tf32: bad argument%s");
  if (n == 0) return CO_OK;
  split_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, hi, lo, n);
  return check_launch("co_split_tf32");
}

extern "C" int co_gemm_tf32x3(const float* A, const float* Whi, const float* Wlo, float* C, const float* bias,
                              const float* residual, const float* scale, const float* shift, int M, int Nout, int K,
                              int lda, int ldc, int ldr, int relu, void* stream) {
  if (!A || !Whi || !Wlo || !C) return fail(CO_ERR_BAD_ARG, "co_gemm_tf32x3: null pointer%s");
  if (M < 0 || Nout <= 0 || K <= 0) return fail(CO_ERR_BAD_ARG, "co_gemm_tf32x3: bad shape%s");
  if ((K % GK) || (Nout % 4) || (lda % 4) || (ldc % 4) || (residual && (ldr % 4)))
    return fail(CO_ERR_UNSUPPORTED, "co_gemm_tf32x3: K %% 32, Nout %% 4 and 16-byte row strides required%s");
  if ((scale == nullptr) != (shift == nullptr)) return fail(CO_ERR_BAD_ARG, "co_gemm_tf32x3: scale and shift go together%s");
  if (((uintptr_t)A | (uintptr_t)Whi | (uintptr_t)Wlo | (uintptr_t)C | (uintptr_t)bias | (uintptr_t)residual |
       (uintptr_t)scale | (uintptr_t)shift) & 15)
    return fail(CO_ERR_BAD_ARG, "co_gemm_tf32x3: pointers must be 16-byte aligned%s");
  if (M == 0) return CO_OK;
  GemmArgs g{A, Whi, Wlo, C, bias, residual, scale, shift, M, Nout, K, lda, ldc, ldr, relu, (Nout + GN - 1) / GN};
  const long tiles = (long)((M + GM - 1) / GM) * g.n_tiles;
  if (tiles > 0x7fffffffL) return fail(CO_ERR_UNSUPPORTED, "co_gemm_tf32x3: too many tiles%s");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * TILE_BYTES);
    if (e != cudaSuccess) return fail(CO_ERR_CUDA, "co_gemm_tf32x3: smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  gemm_tf32x3_kernel<<<(unsigned)tiles, 256, 4 * TILE_BYTES, (cudaStream_t)stream>>>(g);
  return check_launch("co_gemm_tf32x3");
}
