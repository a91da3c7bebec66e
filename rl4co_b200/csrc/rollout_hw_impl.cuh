// Persistent whole-episode rollout kernel, "head-wise" layout (round 2) -- implementation header,
// instantiated per environment in rollout_tsp.cu / rollout_cvrp.cu.  Same contract and the same fused
// reference functions as rollout_impl.cuh (read its header for the file:line map):
//   ConstructivePolicy.forward loop     rl4co/models/common/constructive/base.py:219-251
//   AttentionModelDecoder.forward       rl4co/models/zoo/am/decoder.py:156-193
//   context embedding                   nn/env_embeddings/context.py:61-74,116-134,147-149
//   PointerAttention                    nn/attention.py:274-320
//   process_logits / DecodingStrategy   rl4co/utils/decoding.py:138-188,344-461
//   TSPEnv._step / CVRPEnv._step        envs/routing/tsp/env.py:60-86, cvrp/env.py:66-136
//   get_reward / get_log_likelihood     utils/ops.py:82-90, decoding.py:38-62
//
// What changed against rollout_impl.cuh (which measured 2 120 cycles per node selection, latency-bound with
// two block barriers and ~390 instructions per warp-step):
//   * the pointer logits are split BY HEAD as well:  logits[n] = sum_h o_h . L'_h[n]  (L' = logit_key @ W_out,
//     folded on the host side of the cache), so warp h -- which owns head h of glimpse_key / glimpse_val /
//     folded logit key for all nodes in registers (lane l owns nodes SPL*l .. SPL*l+SPL-1, 16 floats each per
//     tensor) -- goes from the context row all the way to its head's contribution to every raw logit with
//     NO block barrier: scores -> masked softmax -> value sum -> (shared-memory transpose) -> o_h broadcast
//     -> 8*SPL FFMA2 against L'_h;
//   * ONE block barrier per node selection: after it every warp redundantly sums the 8 per-head partials of its
//     lanes' nodes (8 vector LDS + adds in fixed order, so all warps get bit-identical logits), applies
//     tanh-clip / mask / temperature, and finds arg-max and log-sum-exp with warp-local REDUX / shuffles.
//     No second exchange: every warp knows the selected node and steps its replica of the environment;
//   * tanh through one EX2 + one RCP (abs. error ~2e-7, i.e. 2e-6 on a clipped logit) instead of tanhf;
//   * the TSP first-node half of the context projection is no longer a per-node table in the cache
//     (one row per episode was ever read): it is one 128x128 GEMV per episode from node_emb / w_first
//     (cache_width 4E); the old 5E layout with the table is still accepted;
//   * the next instance's cache rows are prefetched into L2 during the current episode.
// HBM traffic per instance is unchanged: one read of its cache rows + T*(8+4) B of outputs.
#pragma once
#include "rollout_impl.cuh"

namespace co {
namespace hw {

template <int SPL>
struct Cfg {
  static constexpr int NS = 32 * SPL;
  static constexpr int LOGSPL = SPL == 4 ? 2 : (SPL == 2 ? 1 : 0);
  static constexpr int MINB = SPL == 4 ? 1 : (SPL == 2 ? 2 : 3);
};

template <int SPL>
struct Smem {
  float ptab[(32 * SPL + 1) * E];   // current-node context table; last row = zeros
  float qfix[E];                    // per-episode fixed part of the query
  float wcap[E];                    // cvrp: remaining-capacity column of project_context
  float hfirst[E];                  // tsp: embedding of the first node (GEMV operand)
  float part[2][8][32 * SPL];       // per-head partial raw logits, double-buffered by step parity
  float gum[2][32 * SPL];           // sampling: -log q per node, double-buffered by step parity
  float tile[8][32 * TILE_LD];      // per-warp transpose tile for the value reduction
  float obuf[8][16];                // per-warp normalised head output (broadcast to the warp)
  float zbuf[32 * SPL];             // warp 0: clipped / masked logits of this step (chosen-node lookup)
  float dem[32 * SPL];
  float2 loc[32 * SPL];
  unsigned char order[32 * SPL];    // cvrp: customers sorted by demand (ascending)
  unsigned char rank_of[32 * SPL];  // cvrp: demand rank of each customer
};

template <int SPL> struct Vec;
template <> struct Vec<4> { using T = float4; };
template <> struct Vec<2> { using T = float2; };
template <> struct Vec<1> { using T = float; };

template <int SPL>
__device__ __forceinline__ void unpack(const typename Vec<SPL>::T& v, float (&x)[SPL]);
template <> __device__ __forceinline__ void unpack<4>(const float4& v, float (&x)[4]) { x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
template <> __device__ __forceinline__ void unpack<2>(const float2& v, float (&x)[2]) { x[0] = v.x; x[1] = v.y; }
template <> __device__ __forceinline__ void unpack<1>(const float& v, float (&x)[1]) { x[0] = v; }
template <int SPL>
__device__ __forceinline__ typename Vec<SPL>::T pack(const float (&x)[SPL]);
template <> __device__ __forceinline__ float4 pack<4>(const float (&x)[4]) { return make_float4(x[0], x[1], x[2], x[3]); }
template <> __device__ __forceinline__ float2 pack<2>(const float (&x)[2]) { return make_float2(x[0], x[1]); }
template <> __device__ __forceinline__ float pack<1>(const float (&x)[1]) { return x[0]; }

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int SPL, int ENV, int MODE>
__global__ void __launch_bounds__(256, Cfg<SPL>::MINB) rollout_kernel(const co_rollout_args A) {
  using C = Cfg<SPL>;
  constexpr int NS = C::NS, LOGSPL = C::LOGSPL;
  constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  using VecT = typename Vec<SPL>::T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<SPL>& sm = *reinterpret_cast<Smem<SPL>*>(smem_raw);

  const int tid = threadIdx.x, lane = tid & 31, h = tid >> 5;
  const int N = A.N, B_inst = A.B_inst, S = A.num_starts, T_max = A.T_max;
  const int B_traj = B_inst * S;
  const int CW = A.cache_width;          // 4E: [K | V | L' | cur-table];  5E (tsp): [K | V | L' | first-table | cur-table]
  const int CUR_BLK = CW / E - 1;
  const bool first_table = (ENV == CO_ENV_TSP) && (CW == 5 * E);
  const bool forced_start = (S > 1) && (A.flags & CO_ROLLOUT_FORCED_START);
  const bool philox = (A.noise == nullptr);
  const float clip = A.tanh_clipping, inv_temp = 1.0f / A.temperature;
  const float Zb = clip * inv_temp;       // z = clip*tanh(.)/T <= Zb: fixed log-softmax offset
  const float zscale = clip * inv_temp;
  float* tile = sm.tile[h];
  const int n0 = SPL * lane;              // first node owned by this lane

  float2 Kr[SPL][8], Vr[SPL][8], Lr[SPL][8];
  int par = 0;                            // step parity (double-buffered exchange areas)

  for (int b = blockIdx.x; b < B_inst; b += gridDim.x) {
    __syncthreads();  // previous instance no longer reads shared memory
    const float* crow = A.cache + (size_t)b * N * CW;
    // ---- one HBM read of the instance: registers <- head-h slices of glimpse_key / glimpse_val / folded logit key
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
      const int n = n0 + k;
      if (n < N) {
        const float4* ks = reinterpret_cast<const float4*>(crow + (size_t)n * CW + 0 * E + h * D);
        const float4* vs = reinterpret_cast<const float4*>(crow + (size_t)n * CW + 1 * E + h * D);
        const float4* ls = reinterpret_cast<const float4*>(crow + (size_t)n * CW + 2 * E + h * D);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 kv = __ldg(ks + c), vv = __ldg(vs + c), lv = __ldg(ls + c);
          Kr[k][2 * c] = make_float2(kv.x, kv.y); Kr[k][2 * c + 1] = make_float2(kv.z, kv.w);
          Vr[k][2 * c] = make_float2(vv.x, vv.y); Vr[k][2 * c + 1] = make_float2(vv.z, vv.w);
          Lr[k][2 * c] = make_float2(lv.x, lv.y); Lr[k][2 * c + 1] = make_float2(lv.z, lv.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          Kr[k][j] = make_float2(0.f, 0.f); Vr[k][j] = make_float2(0.f, 0.f); Lr[k][j] = make_float2(0.f, 0.f);
        }
      }
    }
    if (b + (int)gridDim.x < B_inst) {  // next instance's cache rows -> L2 while this episode runs
      const char* nxt = reinterpret_cast<const char*>(A.cache + (size_t)(b + gridDim.x) * N * CW);
      const int lines = (N * CW * 4 + 127) >> 7;
      for (int i = tid; i < lines; i += 256) prefetch_l2(nxt + ((size_t)i << 7));
    }
    // ---- shared memory <- context table, coordinates, demands
    for (int idx = tid; idx < N * (E / 4); idx += 256) {
      const int n = idx >> 5, c = idx & 31;
      reinterpret_cast<float4*>(sm.ptab + n * E)[c] =
          __ldg(reinterpret_cast<const float4*>(crow + (size_t)n * CW + CUR_BLK * E) + c);
    }
    if (tid < E) {
      sm.ptab[NS * E + tid] = 0.f;
      sm.wcap[tid] = (ENV == CO_ENV_CVRP) ? A.w_capacity[tid] : 0.f;
    }
    if (tid < NS) {
      sm.loc[tid] = (tid < N) ? reinterpret_cast<const float2*>(A.locs)[(size_t)b * N + tid] : make_float2(0.f, 0.f);
      sm.dem[tid] = (ENV == CO_ENV_CVRP && tid >= 1 && tid < N) ? A.demand[(size_t)b * (N - 1) + tid - 1] : 0.f;
    }
    const float cap = (ENV == CO_ENV_CVRP && A.vehicle_capacity) ? A.vehicle_capacity[b] : 1.0f;
    const float thr = cap + 1e-5f;  // fp32 add, as `td["vehicle_capacity"] + 1e-5`
    __syncthreads();
    if (ENV == CO_ENV_CVRP) {  // rank-sort customers by demand (ties by index) -> sm.order
      if (tid >= 1 && tid < N) {
        const float d = sm.dem[tid];
        int rank = 0;
        for (int m = 1; m < N; ++m) {
          const float dm = sm.dem[m];
          rank += (dm < d || (dm == d && m < tid)) ? 1 : 0;
        }
        sm.order[rank] = (unsigned char)tid;
        sm.rank_of[tid] = (unsigned char)rank;
      }
      __syncthreads();
    }
    float dmk[SPL];
#pragma unroll
    for (int k = 0; k < SPL; ++k) dmk[k] = sm.dem[n0 + k];
    // scores split by linearity: q.K = ptab[cur].K + qfix.K + rem * (wcap.K); the last two are
    // per-episode / per-instance constants held in registers (FK, WK)
    auto head_dot = [&](const float* vec, float (&out)[SPL]) {
      const float4* vp = reinterpret_cast<const float4*>(vec + h * D);
      float2 a2[SPL];
#pragma unroll
      for (int k = 0; k < SPL; ++k) a2[k] = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 x = vp[c];
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
          a2[k] = ffma2(make_float2(x.x, x.y), Kr[k][2 * c], a2[k]);
          a2[k] = ffma2(make_float2(x.z, x.w), Kr[k][2 * c + 1], a2[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < SPL; ++k) out[k] = a2[k].x + a2[k].y;
    };
    // tsp: qfix += project_context[:, :E] @ h[first]  (context.py:129-133), then FK = qfix . K
    auto add_first = [&](int a_first, float (&FKo)[SPL]) {
      if (first_table) {
        if (tid < E) sm.qfix[tid] += __ldg(crow + (size_t)a_first * CW + 3 * E + tid);
      } else {
        if (tid < E) sm.hfirst[tid] = __ldg(A.node_emb + ((size_t)b * N + a_first) * E + tid);
        __syncthreads();
        const int e = tid >> 1, half = tid & 1;
        const float4* wr = reinterpret_cast<const float4*>(A.w_first + (size_t)e * E + 64 * half);
        const float4* hv = reinterpret_cast<const float4*>(sm.hfirst + 64 * half);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
          const float4 w0 = __ldg(wr + c), x0 = hv[c], w1 = __ldg(wr + c + 1), x1 = hv[c + 1];
          s0 = fmaf(w0.x, x0.x, s0); s0 = fmaf(w0.y, x0.y, s0); s0 = fmaf(w0.z, x0.z, s0); s0 = fmaf(w0.w, x0.w, s0);
          s1 = fmaf(w1.x, x1.x, s1); s1 = fmaf(w1.y, x1.y, s1); s1 = fmaf(w1.z, x1.z, s1); s1 = fmaf(w1.w, x1.w, s1);
        }
        float sacc = s0 + s1;
        sacc += __shfl_xor_sync(FULL, sacc, 1);
        if (half == 0) sm.qfix[e] += sacc;
      }
      __syncthreads();
      head_dot(sm.qfix, FKo);
    };
    float WK[SPL], FK[SPL];
#pragma unroll
    for (int k = 0; k < SPL; ++k) { WK[k] = 0.f; FK[k] = 0.f; }
    if (ENV == CO_ENV_CVRP) head_dot(sm.wcap, WK);

    for (int s = 0; s < S; ++s) {
      const int traj = s * B_inst + b;  // start-major, rl4co/utils/ops.py:10-29
      // ---------------- reset (tsp/env.py:88-113, cvrp/env.py:98-124)
      uint32_t mybits = 0u;  // bit k = node n0+k visited (padding slots start as visited)
#pragma unroll
      for (int k = 0; k < SPL; ++k) mybits |= (n0 + k >= N) ? (1u << k) : 0u;
      int cur = (ENV == CO_ENV_TSP) ? NS : 0;  // NS -> zero row: step-0 placeholder context
      int first = 0, t = 0, dstep = 0, nvis = 0;
      uint32_t rmask[SPL];   // cvrp: visited customers as a bitmask over demand RANKS
#pragma unroll
      for (int k = 0; k < SPL; ++k) {
        const int lo = 32 * k, nc = N - 1;
        rmask[k] = (nc >= lo + 32) ? 0u : (nc <= lo ? 0xffffffffu : (0xffffffffu << (nc - lo)));
      }
      float used = 0.f, dist = 0.f, ll = 0.f;
      bool anyfeas = false, done = false, depot_seen = false;
      __syncthreads();  // previous trajectory finished with qfix
      if (tid < E) {
        float g = A.graph_ctx ? A.graph_ctx[(size_t)b * E + tid] : 0.f;
        if (ENV == CO_ENV_TSP && !forced_start) g += A.q_placeholder[tid];
        sm.qfix[tid] = g;
      }

      // one environment transition, replicated in every thread
      auto env_step = [&](int a) {
        mybits |= ((a >> LOGSPL) == lane) ? (1u << (a & (SPL - 1))) : 0u;
        if (h == 0 && (ENV == CO_ENV_CVRP || t != 0)) {  // incremental tour length: warp 0 only
          const float2 pa = sm.loc[a], pp = sm.loc[cur];   // (cur is the previous node here; tsp t = 0 has none)
          const float dx = pa.x - pp.x, dy = pa.y - pp.y;
          dist += sqrtf(dx * dx + dy * dy);
        }
        if (ENV == CO_ENV_TSP) {
          if (t == 0) first = a;
        } else {
          used = (used + sm.dem[a == 0 ? 1 : a]) * (a != 0 ? 1.0f : 0.0f);  // cvrp/env.py:70-76
          nvis += (a != 0 || !depot_seen) ? 1 : 0;
          depot_seen = depot_seen || (a == 0);
          if (a != 0) {
            const int r = sm.rank_of[a];
#pragma unroll
            for (int k = 0; k < SPL; ++k) rmask[k] |= ((r >> 5) == k) ? (1u << (r & 31)) : 0u;
          }
          // depot rule (cvrp/env.py:134): some unvisited customer still fits <=> the unvisited customer of
          // least demand fits (fp32 add is monotone in the demand)
          int pmin = NS;
#pragma unroll
          for (int k = SPL - 1; k >= 0; --k) {
            const uint32_t z = ~rmask[k];
            if (z) pmin = 32 * k + __ffs(z) - 1;
          }
          anyfeas = (pmin < N - 1) && !((sm.dem[sm.order[pmin < N - 1 ? pmin : 0]] + used) > thr);
        }
        cur = a; ++t;
        done = (ENV == CO_ENV_TSP) ? (t >= N) : (nvis >= N);  // cvrp: all nodes incl. the depot visited
      };

      if (forced_start) {  // multistart pre_decoder_hook, decoding.py:309-326 + ops.py:128-149
        const int a0 = (s % A.num_loc) + (ENV == CO_ENV_CVRP ? 1 : 0);
        if (tid == 0) { A.actions_out[(size_t)traj * T_max] = a0; A.logp_out[(size_t)traj * T_max] = 0.f; }
        env_step(a0);
        if (ENV == CO_ENV_TSP) {
          __syncthreads();      // qfix initialised
          add_first(a0, FK);    // ends with a barrier + head_dot
        } else {
          __syncthreads();
          head_dot(sm.qfix, FK);
        }
      } else {
        if (ENV == CO_ENV_CVRP) anyfeas = !((sm.dem[sm.order[0]] + used) > thr);
        __syncthreads();
        head_dot(sm.qfix, FK);
      }

      while (!done && t < T_max) {
        // early, latency-tolerant loads for this step
        int forced = 0;
        if (MODE == CO_MODE_EVALUATE) forced = (int)A.forced_actions[(size_t)traj * T_max + t];
        if (MODE == CO_MODE_SAMPLE && tid < N) {  // Gumbel perturbation -log q, q ~ Exp(1), one node per thread
          const float q = philox ? philox_exp1(A.seed, A.offset, traj, dstep, tid)
                                 : A.noise[((size_t)dstep * B_traj + traj) * N + tid];
          sm.gum[par][tid] = -logf(q);
        }

        // ---------------- glimpse + per-head logit partials: warp h = head h, warp-local up to the barrier
        bool fz[SPL];
        {
          const float4* pr = reinterpret_cast<const float4*>(sm.ptab + cur * E + h * D);
          const float rem = cap - used;  // context.py:147-149
          float2 sc2[SPL];
#pragma unroll
          for (int k = 0; k < SPL; ++k) sc2[k] = make_float2(0.f, 0.f);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 p = pr[c];
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
              sc2[k] = ffma2(make_float2(p.x, p.y), Kr[k][2 * c], sc2[k]);
              sc2[k] = ffma2(make_float2(p.z, p.w), Kr[k][2 * c + 1], sc2[k]);
            }
          }
          // scores in log2 units: s * (1/sqrt(head_dim)) * log2(e)
          float sc[SPL], m = -INFINITY;
#pragma unroll
          for (int k = 0; k < SPL; ++k) {
            fz[k] = feasible<ENV>(n0 + k, (mybits >> k) & 1u, dmk[k], used, thr, cur, anyfeas);
            float dot = (sc2[k].x + sc2[k].y) + FK[k];
            if (ENV == CO_ENV_CVRP) dot = fmaf(rem, WK[k], dot);
            sc[k] = fz[k] ? dot * (0.25f * LOG2E) : -INFINITY;
            m = fmaxf(m, sc[k]);
          }
          m = funkey(__reduce_max_sync(FULL, fkey(m)));
          float ev[SPL];
#pragma unroll
          for (int k = 0; k < SPL; ++k) ev[k] = fz[k] ? ex2(sc[k] - m) : 0.f;
          float esum = ev[0];
#pragma unroll
          for (int k = 1; k < SPL; ++k) esum += ev[k];
          // weighted value sum, two halves of 8 channels (keeps the live accumulators at 8 registers);
          // lane-sum of the 16 partial outputs through a padded shared-memory transpose
          float4* trow = reinterpret_cast<float4*>(tile + lane * TILE_LD);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            float2 acc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = __fmul2_rn(make_float2(ev[0], ev[0]), Vr[0][4 * hf + j]);
#pragma unroll
            for (int k = 1; k < SPL; ++k)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[j] = ffma2(make_float2(ev[k], ev[k]), Vr[k][4 * hf + j], acc[j]);
            trow[2 * hf] = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
            trow[2 * hf + 1] = make_float4(acc[2].x, acc[2].y, acc[3].x, acc[3].y);
          }
          esum = warp_sum(esum);
          __syncwarp();
          // lane (channel pair dp, row quarter q) sums 8 rows of its float2 column; quarter q reads its rows
          // rotated by 4*(q&1) so that the two quarters of a half-warp hit disjoint banks (row stride 20 floats)
          const int dp = lane & 7, q = lane >> 3;
          const float* tq = tile + (8 * q) * TILE_LD + 2 * dp;
          const int rot = 4 * (q & 1);
          float2 rs = *reinterpret_cast<const float2*>(tq + ((0 + rot) & 7) * TILE_LD);
#pragma unroll
          for (int r = 1; r < 8; ++r)
            rs = __fadd2_rn(rs, *reinterpret_cast<const float2*>(tq + ((r + rot) & 7) * TILE_LD));
          rs.x += __shfl_xor_sync(FULL, rs.x, 8);
          rs.y += __shfl_xor_sync(FULL, rs.y, 8);
          rs.x += __shfl_xor_sync(FULL, rs.x, 16);
          rs.y += __shfl_xor_sync(FULL, rs.y, 16);
          if (lane < 8) {  // esum >= 1 (the best-scoring node contributes exactly 1)
            const float inv = rcp_approx(esum);
            *reinterpret_cast<float2*>(&sm.obuf[h][2 * dp]) = make_float2(rs.x * inv, rs.y * inv);
          }
          __syncwarp();
          // head h's contribution to every raw logit: o_h . L'_h[n]
          const float4* ob = reinterpret_cast<const float4*>(sm.obuf[h]);
          float2 c2[SPL];
#pragma unroll
          for (int k = 0; k < SPL; ++k) c2[k] = make_float2(0.f, 0.f);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 o4 = ob[c];
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
              c2[k] = ffma2(make_float2(o4.x, o4.y), Lr[k][2 * c], c2[k]);
              c2[k] = ffma2(make_float2(o4.z, o4.w), Lr[k][2 * c + 1], c2[k]);
            }
          }
          float pl[SPL];
#pragma unroll
          for (int k = 0; k < SPL; ++k) pl[k] = c2[k].x + c2[k].y;
          *reinterpret_cast<VecT*>(&sm.part[par][h][n0]) = pack<SPL>(pl);
        }
        __syncthreads();  // the ONLY block barrier of a node selection: all per-head partials are in shared memory

        // ---------------- every warp: sum heads (fixed order), tanh clip, mask, log-sum-exp, selection
        float z[SPL];
        {
          float x[SPL];
          unpack<SPL>(*reinterpret_cast<const VecT*>(&sm.part[par][0][n0]), x);
          if (SPL == 1) {
#pragma unroll
            for (int hh = 1; hh < 8; ++hh) x[0] += sm.part[par][hh][n0];
          } else {  // packed adds (FADD2), heads in fixed order 0..7: bit-identical in every warp
            float2 xa[SPL / 2 + 1];
#pragma unroll
            for (int j = 0; j < SPL / 2; ++j) xa[j] = make_float2(x[2 * j], x[2 * j + 1]);
#pragma unroll
            for (int hh = 1; hh < 8; ++hh) {
              float v[SPL];
              unpack<SPL>(*reinterpret_cast<const VecT*>(&sm.part[par][hh][n0]), v);
#pragma unroll
              for (int j = 0; j < SPL / 2; ++j) xa[j] = __fadd2_rn(xa[j], make_float2(v[2 * j], v[2 * j + 1]));
            }
#pragma unroll
            for (int j = 0; j < SPL / 2; ++j) { x[2 * j] = xa[j].x; x[2 * j + 1] = xa[j].y; }
          }
#pragma unroll
          for (int k = 0; k < SPL; ++k) {
            // clip * tanh(x / sqrt(E)) / T  (decoding.py:169-177);  tanh(u) = 1 - 2 / (exp(2u) + 1)
            const float u = ex2(x[k] * (2.0f * LOG2E * 0.08838834764831845f));
            const float th = fmaf(-2.0f, rcp_approx(u + 1.0f), 1.0f);
            z[k] = fz[k] ? th * zscale : -INFINITY;
          }
        }
        float Ssum;
        {
          float es = ex2(fmaf(z[0], LOG2E, -Zb * LOG2E));  // exp(z - Zb) in (0,1]; 0 if masked
#pragma unroll
          for (int k = 1; k < SPL; ++k) es += ex2(fmaf(z[k], LOG2E, -Zb * LOG2E));
          Ssum = warp_sum(es);
        }
        int a;
        {
          float gk[SPL];
          if (MODE == CO_MODE_SAMPLE) unpack<SPL>(*reinterpret_cast<const VecT*>(&sm.gum[par][n0]), gk);
          float bk = -INFINITY;
          int bn = 0;
#pragma unroll
          for (int k = 0; k < SPL; ++k) {
            const float kf = (MODE == CO_MODE_SAMPLE) ? (fz[k] ? z[k] + gk[k] : -INFINITY) : z[k];
            if (k == 0 || kf > bk) { bk = kf; bn = k; }  // strict '>': lowest node wins ties (torch.argmax)
          }
          const unsigned key = fkey(bk);
          const unsigned wkey = __reduce_max_sync(FULL, key);
          const unsigned vote = __ballot_sync(FULL, key == wkey);
          const int src = __ffs(vote) - 1;  // lowest lane = lowest node range
          a = SPL * src + __shfl_sync(FULL, bn, src);
        }
        if (MODE == CO_MODE_EVALUATE) a = (forced < 0 || forced >= N) ? 0 : forced;
        if (h == 0) {  // outputs: warp 0 only
          *reinterpret_cast<VecT*>(&sm.zbuf[n0]) = pack<SPL>(z);
          __syncwarp();
          if (lane == 0) {  // log_softmax of the chosen node: (z - Zb) - log(sum exp(z - Zb))
            const float za = sm.zbuf[a];
            const float lp = (za - Zb) - lg2(Ssum) * LN2;
            A.logp_out[(size_t)traj * T_max + t] = lp;
            ll += lp;
            A.actions_out[(size_t)traj * T_max + t] = a;
          }
        }

        // ---------------- environment step
        const bool was_first = (ENV == CO_ENV_TSP) && (t == 0);
        env_step(a);
        ++dstep;
        par ^= 1;
        if (was_first) {  // context from now on: [h_first ; h_cur], context.py:129-133
          __syncthreads();  // every warp is past its reads of qfix-derived state; safe to rewrite qfix
          if (tid < E) sm.qfix[tid] = A.graph_ctx ? A.graph_ctx[(size_t)b * E + tid] : 0.f;
          __syncthreads();
          add_first(a, FK);
        }
      }

      // ---------------- epilogue: reward, log-likelihood, padding
      if (tid == 0) {
        const float2 pa = sm.loc[(ENV == CO_ENV_TSP) ? first : 0], pp = sm.loc[cur];
        const float dx = pa.x - pp.x, dy = pa.y - pp.y;
        A.reward_out[traj] = -(dist + sqrtf(dx * dx + dy * dy));
        A.loglik_out[traj] = ll;
        if (A.steps_out) A.steps_out[traj] = t;
        if (A.used_capacity_out) A.used_capacity_out[traj] = used;
        if (A.max_steps_out) atomicMax(A.max_steps_out, t);
      }
      // done instances keep selecting the depot with log-prob 0 until the batch finishes
      for (int c = t + tid; c < T_max; c += 256) {
        A.actions_out[(size_t)traj * T_max + c] = 0;
        A.logp_out[(size_t)traj * T_max + c] = 0.f;
      }
    }
  }
}

template <int SPL, int ENV, int MODE>
static int launch(const co_rollout_args& A, cudaStream_t st) {
  auto kern = rollout_kernel<SPL, ENV, MODE>;
  const size_t smem = sizeof(Smem<SPL>);
  static PerDeviceOnce once;
  static int ctas_per_sm = 1;
  bool& configured = once.flag();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(CO_ERR_CUDA, "co_rollout: smem attribute: %s", cudaGetErrorString(e));
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, 256, smem);
    if (e != cudaSuccess || ctas_per_sm < 1) return fail(CO_ERR_CUDA, "co_rollout: occupancy query failed%s");
    configured = true;
  }
  int grid = device_info().sm_count * ctas_per_sm;
  if (grid > A.B_inst) grid = A.B_inst;
  kern<<<grid, 256, smem, st>>>(A);
  return check_launch("co_rollout");
}

template <int ENV>
static int dispatch(const co_rollout_args& A, cudaStream_t st) {
  const int spl = A.N <= 32 ? 1 : (A.N <= 64 ? 2 : 4);
  const int mode = A.select_mode == CO_SELECT_GREEDY ? CO_MODE_GREEDY
                   : (A.select_mode == CO_SELECT_EVALUATE ? CO_MODE_EVALUATE : CO_MODE_SAMPLE);
#define CO_CASE(S_, M_) if (spl == S_ && mode == M_) return launch<S_, ENV, M_>(A, st)
  CO_CASE(1, CO_MODE_GREEDY); CO_CASE(2, CO_MODE_GREEDY); CO_CASE(4, CO_MODE_GREEDY);
  CO_CASE(1, CO_MODE_SAMPLE); CO_CASE(2, CO_MODE_SAMPLE); CO_CASE(4, CO_MODE_SAMPLE);
  CO_CASE(1, CO_MODE_EVALUATE); CO_CASE(2, CO_MODE_EVALUATE); CO_CASE(4, CO_MODE_EVALUATE);
#undef CO_CASE
  return fail(CO_ERR_BAD_ARG, "co_rollout: no kernel variant%s");
}

}  // namespace hw
}  // namespace co
