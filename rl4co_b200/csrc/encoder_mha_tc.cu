// Encoder self-attention core, tensor-core variant (N <= 128, 8 heads x 16, fp32 in / out):
//   S_h = Q_h K_h^T on tcgen05 (kind::tf32, 3xTF32 hi/lo split -> fp32-class scores) into TMEM,
//   softmax + P V on the FP32 pipe with ONE SCORE ROW PER THREAD read straight from TMEM.
// Same contract as co_encoder_mha (encoder_mha.cu): qkv [B*N, 384] -> out [B*N, 128], the
// F.scaled_dot_product_attention of rl4co/models/nn/attention.py:110-134.
//
// Persistent CTA per SM, 416 threads, warp-specialised:
//   warps 8..11 producers : V of the instance -> SMEM (double-buffered), then per head Q_h / K_h rows:
//                           LDG -> hi = cvt.rna.tf32, lo = x - hi -> STS into K-major (no-swizzle,
//                           8 x 16 B core matrices, LBO 128 B, SBO 512 B) operand tiles, 2-stage ring
//   warp  12    MMA issuer: per head 6 tcgen05.mma 128x128x8 (2 k-steps x {hi.hi, lo.hi, hi.lo}) into one
//                           of 4 TMEM score buffers (128 columns each); tcgen05.commit -> mbarriers
//   warps 0..7  consumers : warp w owns TMEM lanes 32*(w%4).. = query rows, heads h with h%2 == w/4;
//                           two passes over its 128-column score row (tcgen05.ld 32x32b.x32): row max,
//                           then p = ex2, l += p, o += p * V_h[j] (V broadcast from SMEM, FFMA2)
// Compared with the all-SIMT kernel the 8 FFMA2 per (row, key) of Q.K disappear, there are no
// padded query rows (one thread per real row) and no cross-lane reductions at all.
#include <stdlib.h>

#include "co_common.cuh"

namespace co {
namespace mhatc {

constexpr int TILE_B = 128 * 16 * 4;          // one operand tile [128 rows x 16 floats] = 8 KB
constexpr int STAGE_B = 4 * TILE_B;           // Qhi, Qlo, Khi, Klo
constexpr int V_B = 128 * E * 4;              // 64 KB
constexpr int SMEM_B = 2 * V_B + 2 * STAGE_B; // 192 KB
constexpr uint32_t LBO = 128, SBO = 512;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(LBO >> 4) << 16;
  d |= (uint64_t)(SBO >> 4) << 32;
  d |= (uint64_t)1 << 46;  // sm_100 descriptor version; layout_type 0 = SWIZZLE_NONE
  return d;
}
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((128 >> 3) << 17) | ((128 >> 4) << 24);
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(IDESC), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ float rna(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
}

// one key: p = 2^(s*QS - m*QS); l += p; o[0:16] += p * V_h[j]   (V row broadcast from SMEM)
__device__ __forceinline__ void pv_key(float sj, float mq, const float* vrow, float& l, float2 (&o)[8]) {
  constexpr float QS = 0.25f * 1.4426950408889634f;  // 1/sqrt(16) * log2(e)
  const float p = ex2f(fmaf(sj, QS, -mq));
  l += p;
  const float4* v4 = reinterpret_cast<const float4*>(vrow);
  const float4 v0 = v4[0], v1 = v4[1], v2 = v4[2], v3 = v4[3];
  const float2 pp = make_float2(p, p);
  o[0] = __ffma2_rn(pp, make_float2(v0.x, v0.y), o[0]); o[1] = __ffma2_rn(pp, make_float2(v0.z, v0.w), o[1]);
  o[2] = __ffma2_rn(pp, make_float2(v1.x, v1.y), o[2]); o[3] = __ffma2_rn(pp, make_float2(v1.z, v1.w), o[3]);
  o[4] = __ffma2_rn(pp, make_float2(v2.x, v2.y), o[4]); o[5] = __ffma2_rn(pp, make_float2(v2.z, v2.w), o[5]);
  o[6] = __ffma2_rn(pp, make_float2(v3.x, v3.y), o[6]); o[7] = __ffma2_rn(pp, make_float2(v3.z, v3.w), o[7]);
}

constexpr int CW = 8, PW = 4, PT = 32 * PW, MMAW = CW + PW;  // warps [0, CW) consumers, PW producers, one issuer
constexpr int THREADS = (CW + PW + 1) * 32;
constexpr int RGI = 16 / PW;  // row groups (8 rows) per producer warp

__global__ void __launch_bounds__(THREADS, 1) encoder_mha_tc_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                                            int B, int N) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sV = smem;               // [2][128][128] fp32
  unsigned char* sT = smem + 2 * V_B;     // [2 stages][Qhi, Qlo, Khi, Klo]
  __shared__ __align__(8) uint64_t bars[16];
  __shared__ uint32_t tmem_base_s;
  const uint32_t b0 = s32(bars);
  // 0-1 tfull, 2-3 tempty, 4-7 sfull, 8-11 sempty, 12-13 vfull, 14-15 vempty
  auto TFULL = [&](int s) { return b0 + 8 * s; };
  auto TEMPTY = [&](int s) { return b0 + 8 * (2 + s); };
  auto SFULL = [&](int u) { return b0 + 8 * (4 + u); };
  auto SEMPTY = [&](int u) { return b0 + 8 * (8 + u); };
  auto VFULL = [&](int v) { return b0 + 8 * (12 + v); };
  auto VEMPTY = [&](int v) { return b0 + 8 * (14 + v); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == MMAW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    auto init = [&](uint32_t bar, int cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(cnt)); };
    for (int s = 0; s < 2; ++s) {
      init(TFULL(s), PT); init(TEMPTY(s), 1); init(VFULL(s), PT); init(VEMPTY(s), 32 * CW);
    }
    for (int u = 0; u < 4; ++u) { init(SFULL(u), 1); init(SEMPTY(u), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;

  if (warp >= CW && warp < MMAW) {
    // ------------------------------------------------------------------ producers
    const int pt = tid - 32 * CW, pw = pt >> 5;
    const int r8 = lane & 7, c4 = lane >> 3;
    uint32_t hc = 0, ic = 0;
    // rows N .. 32*ceil(N/32) of both V buffers are read with p == 0: make them finite once
    for (int idx = N * 32 + pt; idx < ((N + 31) & ~31) * 32; idx += PT) {
      reinterpret_cast<float4*>(sV)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
      reinterpret_cast<float4*>(sV + V_B)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int b = blockIdx.x; b < B; b += gridDim.x, ++ic) {
      const float* base = qkv + (size_t)b * N * 3 * E;
      const int vb = ic & 1;
      bar_wait(VEMPTY(vb), ((ic >> 1) & 1) ^ 1);
      float* Vs = reinterpret_cast<float*>(sV + vb * V_B);
      const int n4 = N * 32;  // float4 slots of V
      for (int i0 = pt; i0 < n4; i0 += PT * 8) {  // 8 independent 16-B loads in flight per thread
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int idx = i0 + PT * u;
          if (idx < n4) v[u] = __ldg(reinterpret_cast<const float4*>(base + (size_t)(idx >> 5) * 3 * E + 2 * E) + (idx & 31));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int idx = i0 + PT * u;
          if (idx < n4) reinterpret_cast<float4*>(Vs)[idx] = v[u];
        }
      }
      bar_arrive(VFULL(vb));
      for (int h = 0; h < H; ++h, ++hc) {
        const int st = hc & 1;
        // a warp covers 8 rows x 4 16-B chunks of one head (64-B segments); all loads issued before the wait
        float4 q[RGI], k[RGI];
#pragma unroll
        for (int i = 0; i < RGI; ++i) {
          const int rg = pw + PW * i, row = 8 * rg + r8;
          q[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          k[i] = q[i];
          if (rg < 16 && row < N) {
            q[i] = __ldg(reinterpret_cast<const float4*>(base + (size_t)row * 3 * E + h * D) + c4);
            k[i] = __ldg(reinterpret_cast<const float4*>(base + (size_t)row * 3 * E + E + h * D) + c4);
          }
        }
        bar_wait(TEMPTY(st), ((hc >> 1) & 1) ^ 1);
        unsigned char* tq_hi = sT + st * STAGE_B;
#pragma unroll
        for (int i = 0; i < RGI; ++i) {
          const int rg = pw + PW * i;
          if (rg < 16) {
            const uint32_t soff = rg * SBO + c4 * LBO + r8 * 16;
            const float4 qh = make_float4(rna(q[i].x), rna(q[i].y), rna(q[i].z), rna(q[i].w));
            const float4 kh = make_float4(rna(k[i].x), rna(k[i].y), rna(k[i].z), rna(k[i].w));
            *reinterpret_cast<float4*>(tq_hi + soff) = qh;
            *reinterpret_cast<float4*>(tq_hi + TILE_B + soff) =
                make_float4(q[i].x - qh.x, q[i].y - qh.y, q[i].z - qh.z, q[i].w - qh.w);
            *reinterpret_cast<float4*>(tq_hi + 2 * TILE_B + soff) = kh;
            *reinterpret_cast<float4*>(tq_hi + 3 * TILE_B + soff) =
                make_float4(k[i].x - kh.x, k[i].y - kh.y, k[i].z - kh.z, k[i].w - kh.w);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        bar_arrive(TFULL(st));
      }
    }
  } else if (warp == MMAW) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t hc = 0;
      for (int b = blockIdx.x; b < B; b += gridDim.x) {
        for (int h = 0; h < H; ++h, ++hc) {
          const int st = hc & 1, u = hc & 3;
          bar_wait(TFULL(st), (hc >> 1) & 1);
          bar_wait(SEMPTY(u), ((hc >> 2) & 1) ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;");
          const uint32_t qhi = s32(sT + st * STAGE_B), qlo = qhi + TILE_B, khi = qhi + 2 * TILE_B, klo = qhi + 3 * TILE_B;
          const uint32_t d = tmem + u * 128;
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {  // K = 16 = 2 k-steps of 8 tf32 (2 core matrices each)
            const uint32_t off = kk * 2 * LBO;
            mma(d, desc(qhi + off), desc(khi + off), kk);
            mma(d, desc(qlo + off), desc(khi + off), 1);
            mma(d, desc(qhi + off), desc(klo + off), 1);
          }
          commit(TEMPTY(st));
          commit(SFULL(u));
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ consumers: one score row per thread
    const int q4 = warp & 3, hp = warp >> 2;
    const int row = 32 * q4 + lane;
    constexpr float QS = 0.25f * 1.4426950408889634f;  // 1/sqrt(16) * log2(e)
    const int nfull = N >> 5, rem = N & 31;
    uint32_t ic = 0;
    for (int b = blockIdx.x; b < B; b += gridDim.x, ++ic) {
      const int vb = ic & 1;
      bar_wait(VFULL(vb), (ic >> 1) & 1);
      const float* Vs = reinterpret_cast<const float*>(sV + vb * V_B);
      for (int h = hp; h < H; h += 2) {
        const uint32_t hc = ic * 8 + h;
        const int u = hc & 3;
        bar_wait(SFULL(u), (hc >> 2) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t trow = tmem + ((uint32_t)(32 * q4) << 16) + u * 128;
        uint32_t r[32];
        // pass 1: row max over the N real columns (columns >= N hold Q.0 = 0 and are skipped / masked)
        float m0 = -INFINITY, m1 = -INFINITY;
        for (int cc = 0; cc < nfull; ++cc) {
          tmem_ld32(trow + cc * 32, r);
#pragma unroll
          for (int jj = 0; jj < 32; jj += 2) {
            m0 = fmaxf(m0, __uint_as_float(r[jj]));
            m1 = fmaxf(m1, __uint_as_float(r[jj + 1]));
          }
        }
        if (rem) {
          tmem_ld32(trow + nfull * 32, r);
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) m0 = fmaxf(m0, jj < rem ? __uint_as_float(r[jj]) : -INFINITY);
        }
        const float mq = fmaxf(m0, m1) * QS;
        // pass 2: p = 2^((s - m) * QS), l += p, o += p * V_h[j]
        float2 o[8];
        float l = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] = make_float2(0.f, 0.f);
        for (int cc = 0; cc < nfull; ++cc) {
          tmem_ld32(trow + cc * 32, r);
          const float* vp = Vs + (size_t)(cc * 32) * E + h * D;
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) pv_key(__uint_as_float(r[jj]), mq, vp + jj * E, l, o);
        }
        if (rem) {
          tmem_ld32(trow + nfull * 32, r);
          const float* vp = Vs + (size_t)(nfull * 32) * E + h * D;
#pragma unroll
          for (int j4 = 0; j4 < 32; j4 += 4) {
            if (j4 < rem) {  // CTA-uniform; masked scores give p = 0 (V rows N..N+3 are zero-filled)
#pragma unroll
              for (int jj = j4; jj < j4 + 4; ++jj)
                pv_key(jj < rem ? __uint_as_float(r[jj]) : -INFINITY, mq, vp + jj * E, l, o);
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        bar_arrive(SEMPTY(u));
        if (row < N) {
          const float inv = 1.0f / l;
          float4* dst = reinterpret_cast<float4*>(out + ((size_t)b * N + row) * E + h * D);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            dst[c] = make_float4(o[2 * c].x * inv, o[2 * c].y * inv, o[2 * c + 1].x * inv, o[2 * c + 1].y * inv);
        }
      }
      bar_arrive(VEMPTY(vb));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == MMAW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

}  // namespace mhatc
}  // namespace co

namespace co {

int launch_encoder_mha_tc(const float* qkv, float* out, int B, int N, cudaStream_t stream) {
  static PerDeviceOnce once;
  bool& configured = once.flag();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mhatc::encoder_mha_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mhatc::SMEM_B);
    if (e != cudaSuccess) return fail(CO_ERR_CUDA, "co_encoder_mha: smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  int grid = device_info().sm_count;
  if (grid > B) grid = B;
  mhatc::encoder_mha_tc_kernel<<<grid, mhatc::THREADS, mhatc::SMEM_B, stream>>>(qkv, out, B, N);
  return check_launch("co_encoder_mha(tc)");
}

}  // namespace co
