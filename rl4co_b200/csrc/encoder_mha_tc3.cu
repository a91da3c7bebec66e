// Default encoder-attention kernel for N > 64 (CO_MHA_VARIANT=tc3): the pipeline of encoder_mha_tc2.cu, but P never
// touches shared memory.  Verified on the B200 against float64 SDPA (tests/test_gpu_parity.py,
// max |err| 9e-6 at scale 1.5 inputs) and measured at 7.4 ms per layer at 65 536 x 100 (tc: 10.2 ms).
//
// Encoder self-attention core (N <= 128, 8 heads x 16, fp32 in / out), contract of co_encoder_mha (encoder_mha.cu).
//   S_h = Q_h K_h^T                       6 SS-form MMAs 128x128x8 (3xTF32) -> TMEM columns [0, 128) of the head's parity
//   p_ij = 2^((S_ij - m_i) * QS)          one score row per thread; P_hi = top 19 bits goes back IN PLACE over S
//                                         (tcgen05.st.32x32b, lane = row, column = key), P_lo = p - P_hi into a ring
//                                         of two 32-column slots
//   O_h = P_hi [V_hi | V_lo] + P_lo V_hi  TS-form MMAs (A operand from TMEM; tools/micro/ts_mma_check.cu confirmed the
//                                         form and the lane = row / column = k layout on the B200) into 32 O columns
// TMEM per head parity (256 columns): S / P_hi [0,128), P_lo ring [128,192), O [192,224). One head per parity in
// flight; the issuer starts Q.K^T of the parity's next head as soon as the last P.V MMA of the current one retired,
// while the consumers are still busy reading O. SMEM: 2 QK stages x 32 KB + 4 V^T slots x 16 KB = 128 KB -- operand
// tiles only; per instance ~1 MB of shared-memory traffic instead of ~3.7 MB.
#include <stdlib.h>

#include "co_common.cuh"

namespace co {
namespace mhatc3 {

constexpr int THREADS = 448;
constexpr int QK_TILE = 128 * 16 * 4;   // [128 rows x 16 floats]
constexpr int QK_STAGE = 4 * QK_TILE;   // Qhi, Qlo, Khi, Klo
constexpr int VT_GRP = 4096;            // 8 d-rows x 128 keys x 4 B
constexpr int VT_SLOT = 4 * VT_GRP;     // hi d0-7, hi d8-15, lo d0-7, lo d8-15
constexpr int OFF_VT = 2 * QK_STAGE;
constexpr int SMEM_B = OFF_VT + 4 * VT_SLOT;  // 131 072
constexpr uint32_t COL_PLO = 128, COL_O = 192;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// K-major, SWIZZLE_NONE shared-memory descriptor (cute::UMMA::SmemDescriptor): LBO = step between core matrices
// along K (128 B), SBO = step between 8-row groups
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t sbo, uint32_t lbo = 128) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// cute::UMMA::InstrDescriptor: D = F32 [4,6), A = B = TF32 [7,10) [10,13), b_major [16] (1 = MN-major), N >> 3 [17,23), M >> 4 [24,29)
__host__ __device__ constexpr uint32_t idesc(int n, int b_mn = 0) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ bool bar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
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
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ float rna(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}
__device__ __forceinline__ float trunc_tf32(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
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
}
// tcgen05.wait::ld with the destination registers as in/out operands, so no use of them can be scheduled above it
__device__ __forceinline__ void tmem_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :: "memory");
}


// TS form: A operand from TMEM (lane = row, 8 consecutive 32-bit columns = one k-step), B from a SMEM descriptor
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a_tmem), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
        "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
        "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])),
        "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct HeadRegs {  // one head's Q / K / V slices for this producer thread: 4 row groups x one 16-B chunk
  float4 q[4], k[4], v[4];
};

__global__ void __launch_bounds__(THREADS, 1) encoder_mha_tc3_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                                      int B, int N) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[24];
  __shared__ uint32_t tmem_base_s;
  const uint32_t b0 = s32(bars), sbase = s32(smem);
  auto QFULL = [&](int p) { return b0 + 8 * p; };             // producers -> issuer p        (128)
  auto QEMPTY = [&](int p) { return b0 + 8 * (2 + p); };      // Q.K^T MMAs retired           (commit)
  auto VTFULL = [&](int s) { return b0 + 8 * (4 + s); };      // producers -> issuer          (128)
  auto VTEMPTY = [&](int s) { return b0 + 8 * (8 + s); };     // head's P.V MMAs retired      (commit)
  auto SFULL = [&](int p) { return b0 + 8 * (12 + p); };      // scores ready                 (commit)
  auto OFULL = [&](int p) { return b0 + 8 * (14 + p); };      // head's output accumulated, S / P_hi columns free (commit)
  auto PFULL = [&](int s) { return b0 + 8 * (16 + s); };      // consumers -> issuer          (128), s = 2 * parity + slot
  auto PEMPTY = [&](int s) { return b0 + 8 * (20 + s); };     // block's MMAs retired: P_lo slot free (commit)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 12) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    auto init = [&](uint32_t bar, int cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(cnt)); };
    for (int p = 0; p < 2; ++p) { init(QFULL(p), 128); init(QEMPTY(p), 1); init(SFULL(p), 1); init(OFULL(p), 1); }
    for (int s = 0; s < 4; ++s) { init(VTFULL(s), 128); init(VTEMPTY(s), 1); init(PFULL(s), 128); init(PEMPTY(s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  tc_before();
  __syncthreads();
  tc_after();
  const uint32_t tmem = tmem_base_s;
  const int ninst = (B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // instances of this CTA
  const int nc32 = (N + 31) >> 5;                                                 // 32-key blocks of P

  if (warp >= 8 && warp < 12) {
    // ------------------------------------------------------------------ producers (128 threads)
    const int pw = warp - 8, r8 = lane & 7, c4 = lane >> 3;
    const int total = ninst * H;
    auto load = [&](int hc, HeadRegs& R) {
      const int b = blockIdx.x + (hc >> 3) * gridDim.x, h = hc & 7;
      const float* base = qkv + (size_t)b * N * 3 * E + h * D;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = 8 * (pw + 4 * i) + r8;
        R.q[i] = R.k[i] = R.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < N) {
          const float4* src = reinterpret_cast<const float4*>(base + (size_t)row * 3 * E) + c4;
          R.q[i] = __ldg(src);
          R.k[i] = __ldg(src + E / 4);
          R.v[i] = __ldg(src + 2 * E / 4);
        }
      }
      // L2 prefetch of the head pair two pairs ahead (a 128-B line holds the slices of two heads)
      const int hp = hc + 4;
      if (!(hc & 1) && hp < total && c4 == 0) {
        const float* pb = qkv + (size_t)(blockIdx.x + (hp >> 3) * gridDim.x) * N * 3 * E + (hp & 7) * D;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = 8 * (pw + 4 * i) + r8;
          if (row < N) {
            prefetch_l2(pb + (size_t)row * 3 * E);
            prefetch_l2(pb + (size_t)row * 3 * E + E);
            prefetch_l2(pb + (size_t)row * 3 * E + 2 * E);
          }
        }
      }
    };
    auto store = [&](int hc, const HeadRegs& R) {
      const int p = hc & 1, np = hc >> 1;  // head parity, heads of this parity before this one
      bar_wait(QEMPTY(p), (np & 1) ^ 1);
      unsigned char* tq = smem + p * QK_STAGE;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t soff = (pw + 4 * i) * 512 + c4 * 128 + r8 * 16;
        const float4 q = R.q[i], k = R.k[i];
        const float4 qh = make_float4(rna(q.x), rna(q.y), rna(q.z), rna(q.w));
        const float4 kh = make_float4(rna(k.x), rna(k.y), rna(k.z), rna(k.w));
        *reinterpret_cast<float4*>(tq + soff) = qh;
        *reinterpret_cast<float4*>(tq + QK_TILE + soff) = make_float4(q.x - qh.x, q.y - qh.y, q.z - qh.z, q.w - qh.w);
        *reinterpret_cast<float4*>(tq + 2 * QK_TILE + soff) = kh;
        *reinterpret_cast<float4*>(tq + 3 * QK_TILE + soff) = make_float4(k.x - kh.x, k.y - kh.y, k.z - kh.z, k.w - kh.w);
      }
      fence_async();
      bar_arrive(QFULL(p));
      const int vs = 2 * p + (np & 1);
      bar_wait(VTEMPTY(vs), ((np >> 1) & 1) ^ 1);
      unsigned char* tv = smem + OFF_VT + vs * VT_SLOT;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = 8 * (pw + 4 * i) + r8;  // key
        const float v[4] = {R.v[i].x, R.v[i].y, R.v[i].z, R.v[i].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {  // element (d = 4 * c4 + e, key j) of V_h^T
          const int d = 4 * c4 + e;
          const uint32_t off = (d >> 3) * VT_GRP + (j >> 2) * 128 + (d & 7) * 16 + (j & 3) * 4;
          const float hi = rna(v[e]);
          *reinterpret_cast<float*>(tv + off) = hi;
          *reinterpret_cast<float*>(tv + 2 * VT_GRP + off) = v[e] - hi;
        }
      }
      fence_async();
      bar_arrive(VTFULL(vs));
    };
    HeadRegs ra, rb;
    load(0, ra);
    for (int hc = 0; hc < total; hc += 2) {  // total is even
      load(hc + 1, rb);
      store(hc, ra);
      if (hc + 2 < total) load(hc + 2, ra);
      store(hc + 1, rb);
    }
  } else if (warp >= 12) {
    // ------------------------------------------------------------------ MMA issuers (one thread per head parity)
    if (lane == 0) {
      const int p = warp - 12, total = ninst * 4;
      const uint32_t tp = tmem + p * 256;
      auto qk = [&](int n) {  // the parity's S columns are free: the caller waited for OFULL of head n - 1
        bar_wait(QFULL(p), n & 1);
        tc_after();
        const uint32_t qhi = sbase + p * QK_STAGE, qlo = qhi + QK_TILE, khi = qhi + 2 * QK_TILE, klo = qhi + 3 * QK_TILE;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const uint32_t off = kk * 256;
          mma(tp, desc(qhi + off, 512), desc(khi + off, 512), idesc(128), kk);
          mma(tp, desc(qlo + off, 512), desc(khi + off, 512), idesc(128), 1);
          mma(tp, desc(qhi + off, 512), desc(klo + off, 512), idesc(128), 1);
        }
        commit(QEMPTY(p));
        commit(SFULL(p));
      };
      qk(0);
      int psl = 0;
      uint32_t pph = 0;
      for (int n = 0; n < total; ++n) {
        const int vs = 2 * p + (n & 1);
        bar_wait(VTFULL(vs), (n >> 1) & 1);
        const uint32_t vt = sbase + OFF_VT + vs * VT_SLOT;
        for (int c32 = 0; c32 < nc32; ++c32) {
          const int ps = 2 * p + psl;
          bar_wait(PFULL(ps), pph);
          tc_after();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {  // 8 keys per k-step
            const uint64_t bd = desc(vt + (c32 * 8 + ks * 2) * 128, VT_GRP);
            mma_ts(tp + COL_O, tp + c32 * 32 + ks * 8, bd, idesc(32), (c32 | ks) != 0);   // P_hi [V_hi | V_lo]
            mma_ts(tp + COL_O, tp + COL_PLO + psl * 32 + ks * 8, bd, idesc(16), 1);       // P_lo  V_hi
          }
          commit(PEMPTY(ps));
          if (++psl == 2) { psl = 0; pph ^= 1; }
        }
        commit(OFULL(p));
        commit(VTEMPTY(vs));
        if (n + 1 < total) {
          bar_wait(OFULL(p), n & 1);  // every MMA that reads P_hi of head n has retired: S may be overwritten
          qk(n + 1);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ consumers: one score row per thread
    const int q4 = warp & 3, p = warp >> 2;
    const int row = 32 * q4 + lane;
    constexpr float QS = 0.25f * 1.4426950408889634f;  // 1/sqrt(16) * log2(e)
    const int nfull = N >> 5, rem = N & 31;
    const uint32_t trow = tmem + ((uint32_t)(32 * q4) << 16) + p * 256;
    const int total = ninst * 4;
    int psl = 0;
    uint32_t pph = 0;
    float l_prev = 1.f;
    uint32_t r[32];
    auto write_out = [&](int n) {  // epilogue of head n: O = (P_hi V_hi + P_lo V_hi) [0..15] + P_hi V_lo [16..31]
      bar_wait(OFULL(p), n & 1);
      tc_after();
      tmem_ld32(trow + COL_O, r);
      tmem_wait(r);
      if (row < N) {
        const int b = blockIdx.x + (n >> 2) * gridDim.x, h = 2 * (n & 3) + p;
        const float inv = 1.0f / l_prev;
        float4* dst = reinterpret_cast<float4*>(out + ((size_t)b * N + row) * E + h * D);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          dst[c] = make_float4((__uint_as_float(r[4 * c]) + __uint_as_float(r[16 + 4 * c])) * inv,
                               (__uint_as_float(r[4 * c + 1]) + __uint_as_float(r[17 + 4 * c])) * inv,
                               (__uint_as_float(r[4 * c + 2]) + __uint_as_float(r[18 + 4 * c])) * inv,
                               (__uint_as_float(r[4 * c + 3]) + __uint_as_float(r[19 + 4 * c])) * inv);
      }
    };
    for (int n = 0; n < total; ++n) {
      bar_wait(SFULL(p), n & 1);
      tc_after();
      // pass 1: row max over the N real columns
      float m0 = -INFINITY, m1 = -INFINITY;
      for (int cc = 0; cc < nfull; ++cc) {
        tmem_ld32(trow + cc * 32, r);
        tmem_wait(r);
#pragma unroll
        for (int jj = 0; jj < 32; jj += 2) {
          m0 = fmaxf(m0, __uint_as_float(r[jj]));
          m1 = fmaxf(m1, __uint_as_float(r[jj + 1]));
        }
      }
      if (rem) {
        tmem_ld32(trow + nfull * 32, r);
        tmem_wait(r);
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) m0 = fmaxf(m0, jj < rem ? __uint_as_float(r[jj]) : -INFINITY);
      }
      const float mq = fmaxf(m0, m1) * QS;
      // the previous head's output leaves the O columns before this head's first P.V MMA (which needs all 128
      // PFULL arrivals, each made after its thread's read) can overwrite them
      if (n > 0) write_out(n - 1);
      // pass 2: P_hi in place over S, P_lo into the ring, 32 keys per block
      float l = 0.f;
      for (int c32 = 0; c32 < nc32; ++c32) {
        tmem_ld32(trow + c32 * 32, r);
        tmem_wait(r);
        const int ps = 2 * p + psl;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float ph[16], pl[16];
          const int left = N - c32 * 32 - half * 16;  // keys >= N: p = 0 (their V^T columns are zero as well)
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const float pv = jj < left ? ex2f(fmaf(__uint_as_float(r[16 * half + jj]), QS, -mq)) : 0.f;
            l += pv;
            ph[jj] = trunc_tf32(pv);
            pl[jj] = pv - ph[jj];
          }
          tmem_st16(trow + c32 * 32 + half * 16, ph);
          if (half == 0) bar_wait(PEMPTY(ps), pph ^ 1);  // the MMAs that read this P_lo slot two blocks ago retired
          tmem_st16(trow + COL_PLO + psl * 32 + half * 16, pl);
        }
        tmem_st_wait();
        tc_before();
        bar_arrive(PFULL(ps));
        if (++psl == 2) { psl = 0; pph ^= 1; }
      }
      l_prev = l;
    }
    if (total > 0) write_out(total - 1);
  }
  tc_before();
  __syncthreads();
  if (warp == 12) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

}  // namespace mhatc3

int launch_encoder_mha_tc3(const float* qkv, float* out, int B, int N, cudaStream_t stream) {
  static PerDeviceOnce once;
  bool& configured = once.flag();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mhatc3::encoder_mha_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mhatc3::SMEM_B);
    if (e != cudaSuccess) return fail(CO_ERR_CUDA, "co_encoder_mha: smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  int grid = device_info().sm_count;
  if (grid > B) grid = B;
  mhatc3::encoder_mha_tc3_kernel<<<grid, mhatc3::THREADS, mhatc3::SMEM_B, stream>>>(qkv, out, B, N);
  return check_launch("co_encoder_mha(tc3)");
}

}  // namespace co
