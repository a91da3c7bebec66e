// Query-batched persistent rollout for multistart / multisample decoding (POMO): the S
// trajectories of an instance share its K / V / logit-key (am/decoder.py:178-179 shares them the
// same way through `unbatchify`), so this variant advances Q = 4 trajectories per pass.
//
// Same structure and numerics as rollout_impl.cuh (read that header first); differences:
//   * every phase handles the Q trajectories of the group before the block barrier, written
//     stage-major in straight-line code so that their independent instruction streams interleave
//     (a warp issues in order: a per-trajectory loop would serialise the chains and gain nothing --
//     measured); the two barriers and the serial REDUX / MUFU / shuffle / LDS latencies are thereby
//     amortised over Q node selections (the single-trajectory kernel is latency-bound at 37 % issue use);
//   * the folded logit key lives in shared memory (padded rows, conflict-free LDS.128) instead of
//     registers -- it is read once per pass and reused by the Q queries -- which frees the
//     registers for the per-trajectory state and accumulators;
//   * trajectories that are done (CVRP: variable length) are still computed but their results are
//     discarded by predication; the group ends when all of its trajectories are done.
// Layout of results is unchanged: row j = s * B + b (start-major, rl4co/utils/ops.py:10-29).
#pragma once
#include "rollout_impl.cuh"

namespace co {

constexpr int MSQ = 4;  // trajectories per pass

template <int SPL>
struct CfgMS {  // logits phase of this kernel: thread (node, part) owns 16*SPL contiguous channels of its node
  static constexpr int NS = 32 * SPL;       // node slots
  static constexpr int PARTS = 8 / SPL;     // threads sharing one node in the logits phase
  static constexpr int NPW = 32 / PARTS;    // nodes per warp in the logits phase
  static constexpr int EPP = 16 * SPL;      // channels of logit_key per thread
  static constexpr int OPAD = EPP + 4;      // padded stride of a part's chunk in `o` (bank spread)
};

template <int SPL>
struct SmemMS {
  float ptab[(32 * SPL + 1) * E];          // current-node context table; last row = zeros
  float lkey[32 * SPL * (E + 4 * (8 / SPL))];  // folded logit key, row = PARTS chunks of (EPP + 4)
  float qfix[MSQ][E];                      // per-trajectory fixed part of the query
  float wcap[E];
  float o[MSQ][8 * (16 * SPL + 4) + 8];    // concatenated heads per trajectory (padded per part)
  float tile[8][2][32 * TILE_LD];          // per-warp ping-pong transpose tiles
  unsigned red_key[MSQ][8];
  int red_idx[MSQ][8];
  float red_sum[MSQ][8];
  float dem[32 * SPL];
  float2 loc[32 * SPL];
  unsigned char order[32 * SPL];
  unsigned char rank_of[32 * SPL];
  float ll_acc[MSQ];
};

template <int SPL, int ENV, int MODE>
__global__ void __launch_bounds__(256, 1) rollout_ms_kernel(const co_rollout_args A) {
  using C = CfgMS<SPL>;
  constexpr int NS = C::NS, PARTS = C::PARTS, NPW = C::NPW, EPP = C::EPP, OPAD = C::OPAD;
  constexpr int LROW = E + 4 * PARTS;                    // padded logit-key row (floats)
  constexpr int CW = (ENV == CO_ENV_TSP ? 5 : 4) * E;
  constexpr int CUR_BLK = (ENV == CO_ENV_TSP ? 4 : 3);
  constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  constexpr int Q = MSQ;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemMS<SPL>& sm = *reinterpret_cast<SmemMS<SPL>*>(smem_raw);

  const int tid = threadIdx.x, lane = tid & 31, h = tid >> 5;
  const int N = A.N, B_inst = A.B_inst, S = A.num_starts, T_max = A.T_max;
  const int B_traj = B_inst * S;
  const bool forced_start = (A.flags & CO_ROLLOUT_FORCED_START) != 0;
  const bool philox = (A.noise == nullptr);
  const int nL = h * NPW + lane / PARTS;
  const int part = lane % PARTS;
  const float clip = A.tanh_clipping, inv_temp = 1.0f / A.temperature;
  const float Zb = clip * inv_temp;

  float2 Kr[SPL][8], Vr[SPL][8];

  for (int b = blockIdx.x; b < B_inst; b += gridDim.x) {
    __syncthreads();
    const float* crow = A.cache + (size_t)b * N * CW;
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
      const int n = lane + 32 * k;
      if (n < N) {
        const float4* ks = reinterpret_cast<const float4*>(crow + (size_t)n * CW + 0 * E + h * D);
        const float4* vs = reinterpret_cast<const float4*>(crow + (size_t)n * CW + 1 * E + h * D);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 kv = __ldg(ks + c), vv = __ldg(vs + c);
          Kr[k][2 * c] = make_float2(kv.x, kv.y); Kr[k][2 * c + 1] = make_float2(kv.z, kv.w);
          Vr[k][2 * c] = make_float2(vv.x, vv.y); Vr[k][2 * c + 1] = make_float2(vv.z, vv.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { Kr[k][j] = make_float2(0.f, 0.f); Vr[k][j] = make_float2(0.f, 0.f); }
      }
    }
    // shared memory <- context table, folded logit key (padded rows), coordinates, demands
    for (int idx = tid; idx < NS * (E / 4); idx += 256) {
      const int n = idx >> 5, c = idx & 31;
      float4 pv = make_float4(0.f, 0.f, 0.f, 0.f), lv = pv;
      if (n < N) {
        pv = __ldg(reinterpret_cast<const float4*>(crow + (size_t)n * CW + CUR_BLK * E) + c);
        lv = __ldg(reinterpret_cast<const float4*>(crow + (size_t)n * CW + 2 * E) + c);
      }
      reinterpret_cast<float4*>(sm.ptab + n * E)[c] = pv;
      const int e = 4 * c;  // channel -> padded position: part chunk e / EPP shifted by 4 floats each
      *reinterpret_cast<float4*>(sm.lkey + n * LROW + e + 4 * (e / EPP)) = lv;
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
    const float thr = cap + 1e-5f;
    __syncthreads();
    if (ENV == CO_ENV_CVRP) {
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
    for (int k = 0; k < SPL; ++k) dmk[k] = sm.dem[lane + 32 * k];
    const float dL = sm.dem[nL];
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
    float WK[SPL];
#pragma unroll
    for (int k = 0; k < SPL; ++k) WK[k] = 0.f;
    if (ENV == CO_ENV_CVRP) head_dot(sm.wcap, WK);

    for (int g0 = 0; g0 < S; g0 += Q) {
      // ---------------- per-trajectory state (replicated in every thread)
      uint32_t mybits[Q], rmask[Q][ENV == CO_ENV_CVRP ? SPL : 1];
      int cur[Q], prev[Q], first[Q], tstep[Q], dstep[Q], nvis[Q];
      float used[Q], dist[Q], FK[Q][SPL];
      bool anyfeas[Q], fin[Q], depot_seen[Q];
#pragma unroll
      for (int j = 0; j < Q; ++j) {
        mybits[j] = (nL >= N) ? 0x100u : 0u;
#pragma unroll
        for (int k = 0; k < SPL; ++k) mybits[j] |= (lane + 32 * k >= N) ? (1u << k) : 0u;
        if (ENV == CO_ENV_CVRP) {
#pragma unroll
          for (int k = 0; k < SPL; ++k) {
            const int lo = 32 * k, nc = N - 1;
            rmask[j][k] = (nc >= lo + 32) ? 0u : (nc <= lo ? 0xffffffffu : (0xffffffffu << (nc - lo)));
          }
        }
        cur[j] = (ENV == CO_ENV_TSP) ? NS : 0;
        prev[j] = 0; first[j] = 0; tstep[j] = 0; dstep[j] = 0; nvis[j] = 0;
        used[j] = 0.f; dist[j] = 0.f;
        anyfeas[j] = false; depot_seen[j] = false;
        fin[j] = (g0 + j >= S);  // inactive tail trajectories of the last group
      }
      __syncthreads();  // previous group finished with qfix / ll_acc
      if (tid < E) {
        float gq = A.graph_ctx ? A.graph_ctx[(size_t)b * E + tid] : 0.f;
        if (ENV == CO_ENV_TSP && !forced_start) gq += A.q_placeholder[tid];
#pragma unroll
        for (int j = 0; j < Q; ++j) sm.qfix[j][tid] = gq;
      }
      if (tid < Q) sm.ll_acc[tid] = 0.f;

      auto env_step = [&](int j, int a) {
#pragma unroll
        for (int k = 0; k < SPL; ++k) mybits[j] |= (a == lane + 32 * k) ? (1u << k) : 0u;
        mybits[j] |= (a == nL) ? 0x100u : 0u;
        if (h == 0) {
          const float2 pa = sm.loc[a], pp = sm.loc[prev[j]];
          const float dx = pa.x - pp.x, dy = pa.y - pp.y;
          if (ENV == CO_ENV_CVRP || tstep[j] != 0) dist[j] += sqrtf(dx * dx + dy * dy);
        }
        if (ENV == CO_ENV_TSP) {
          if (tstep[j] == 0) first[j] = a;
        } else {
          used[j] = (used[j] + sm.dem[a == 0 ? 1 : a]) * (a != 0 ? 1.0f : 0.0f);
          nvis[j] += (a != 0 || !depot_seen[j]) ? 1 : 0;
          depot_seen[j] = depot_seen[j] || (a == 0);
          if (a != 0) {
            const int r = sm.rank_of[a];
#pragma unroll
            for (int k = 0; k < SPL; ++k) rmask[j][k] |= ((r >> 5) == k) ? (1u << (r & 31)) : 0u;
          }
          int pmin = NS;
#pragma unroll
          for (int k = SPL - 1; k >= 0; --k) {
            const uint32_t z = ~rmask[j][k];
            if (z) pmin = 32 * k + __ffs(z) - 1;
          }
          anyfeas[j] = (pmin < N - 1) && !((sm.dem[sm.order[pmin < N - 1 ? pmin : 0]] + used[j]) > thr);
        }
        prev[j] = a; cur[j] = a; ++tstep[j];
        if ((ENV == CO_ENV_TSP) ? (tstep[j] >= N) : (nvis[j] >= N)) fin[j] = true;
        if (tstep[j] >= T_max) fin[j] = true;
      };

#pragma unroll
      for (int j = 0; j < Q; ++j) {
        if (fin[j]) continue;
        const int s = g0 + j;
        const int traj = s * B_inst + b;
        if (forced_start) {  // decoding.py:309-326 + ops.py:128-149
          const int a0 = (s % A.num_loc) + (ENV == CO_ENV_CVRP ? 1 : 0);
          if (tid == 0) { A.actions_out[(size_t)traj * T_max] = a0; A.logp_out[(size_t)traj * T_max] = 0.f; }
          env_step(j, a0);
          if (ENV == CO_ENV_TSP && tid < E) sm.qfix[j][tid] += __ldg(crow + (size_t)a0 * CW + 3 * E + tid);
        } else if (ENV == CO_ENV_CVRP) {
          anyfeas[j] = !((sm.dem[sm.order[0]] + used[j]) > thr);
        }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < Q; ++j) head_dot(sm.qfix[j], FK[j]);

      while (!(fin[0] && fin[1] && fin[2] && fin[3])) {
        // Every phase is written stage-major over the trajectories in straight-line code (no
        // per-trajectory branches): a warp issues in order, so only interleaved independent
        // instruction streams hide the REDUX / MUFU / shuffle / LDS latencies of each chain.
        // Finished trajectories are computed too (their results are discarded by predication).
        // ---------------- glimpse (warp h = head h), two trajectories at a time (two transpose tiles)
#pragma unroll
        for (int jp = 0; jp < Q; jp += 2) {
          float sc[2][SPL], m[2], esum[2];
          uint32_t fzb[2];
          // stage A: scores of both trajectories, feasibility, row max
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int j = jp + u;
            const float4* pr = reinterpret_cast<const float4*>(sm.ptab + cur[j] * E + h * D);
            const float rem = cap - used[j];
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
            float mm = -INFINITY;
            fzb[u] = 0;
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
              const bool f = feasible<ENV>(lane + 32 * k, (mybits[j] >> k) & 1u, dmk[k], used[j], thr, cur[j], anyfeas[j]);
              fzb[u] |= f ? (1u << k) : 0u;
              float dot = (sc2[k].x + sc2[k].y) + FK[j][k];
              if (ENV == CO_ENV_CVRP) dot = fmaf(rem, WK[k], dot);
              sc[u][k] = f ? dot * (0.25f * LOG2E) : -INFINITY;
              mm = fmaxf(mm, sc[u][k]);
            }
            m[u] = mm;
          }
          const unsigned mk0 = __reduce_max_sync(FULL, fkey(m[0])), mk1 = __reduce_max_sync(FULL, fkey(m[1]));
          m[0] = funkey(mk0); m[1] = funkey(mk1);
          // stage B: exp, value accumulation, partial outputs to the trajectory's tile
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            float2 acc[8];
            float es = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
              const float e = ((fzb[u] >> k) & 1u) ? ex2(sc[u][k] - m[u]) : 0.f;
              es += e;
              const float2 e2 = make_float2(e, e);
#pragma unroll
              for (int c = 0; c < 8; ++c) acc[c] = ffma2(e2, Vr[k][c], acc[c]);
            }
            esum[u] = es;
            float4* trow = reinterpret_cast<float4*>(sm.tile[h][u] + lane * TILE_LD);
#pragma unroll
            for (int c = 0; c < 4; ++c) trow[c] = make_float4(acc[2 * c].x, acc[2 * c].y, acc[2 * c + 1].x, acc[2 * c + 1].y);
          }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {  // both row sums, level by level
            const float t0 = __shfl_xor_sync(FULL, esum[0], off), t1 = __shfl_xor_sync(FULL, esum[1], off);
            esum[0] += t0; esum[1] += t1;
          }
          __syncwarp();
          // stage C: lane sums through the transposed tiles, normalise, publish heads
          const int d = lane & 15, half = lane >> 4;
          float r2[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float* tile = sm.tile[h][u];
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int r = 0; r < 16; r += 4) {
              s0 += tile[(16 * half + ((r + 0 + 4 * half) & 15)) * TILE_LD + d];
              s1 += tile[(16 * half + ((r + 1 + 4 * half) & 15)) * TILE_LD + d];
              s2 += tile[(16 * half + ((r + 2 + 4 * half) & 15)) * TILE_LD + d];
              s3 += tile[(16 * half + ((r + 3 + 4 * half) & 15)) * TILE_LD + d];
            }
            r2[u] = (s0 + s1) + (s2 + s3);
          }
          const float x0 = __shfl_xor_sync(FULL, r2[0], 16), x1 = __shfl_xor_sync(FULL, r2[1], 16);
          if (lane < 16) {
            const int e = h * D + d;
            sm.o[jp][e + 4 * (e / EPP)] = __fdividef(r2[0] + x0, esum[0]);
            sm.o[jp + 1][e + 4 * (e / EPP)] = __fdividef(r2[1] + x1, esum[1]);
          }
          __syncwarp();  // the two tiles are reused by the next pair
        }
        __syncthreads();  // B1: heads of all trajectories complete

        // ---------------- pointer logits: thread (nL, part); logit key streamed once for the Q queries
        float z[Q];
        {
          float2 pa[Q], pb[Q];
#pragma unroll
          for (int j = 0; j < Q; ++j) { pa[j] = make_float2(0.f, 0.f); pb[j] = make_float2(0.f, 0.f); }
          const float4* lk = reinterpret_cast<const float4*>(sm.lkey + nL * LROW + part * (EPP + 4));
#pragma unroll
          for (int c = 0; c < EPP / 4; ++c) {
            const float4 l4 = lk[c];
#pragma unroll
            for (int j = 0; j < Q; ++j) {
              const float4 x = reinterpret_cast<const float4*>(sm.o[j] + part * OPAD)[c];
              pa[j] = ffma2(make_float2(x.x, x.y), make_float2(l4.x, l4.y), pa[j]);
              pb[j] = ffma2(make_float2(x.z, x.w), make_float2(l4.z, l4.w), pb[j]);
            }
          }
          float pl[Q], keyf[Q], ex[Q];
#pragma unroll
          for (int j = 0; j < Q; ++j) pl[j] = (pa[j].x + pa[j].y) + (pb[j].x + pb[j].y);
#pragma unroll
          for (int off = PARTS / 2; off > 0; off >>= 1) {
#pragma unroll
            for (int j = 0; j < Q; ++j) pl[j] += __shfl_xor_sync(FULL, pl[j], off);
          }
#pragma unroll
          for (int j = 0; j < Q; ++j) {
            const bool fzL = feasible<ENV>(nL, (mybits[j] >> 8) & 1u, dL, used[j], thr, cur[j], anyfeas[j]);
            const float lg = tanhf(pl[j] * 0.08838834764831845f) * clip;
            z[j] = fzL ? lg * inv_temp : -INFINITY;
            keyf[j] = z[j];
            if (MODE == CO_MODE_SAMPLE) {
              keyf[j] = -INFINITY;
              if (part == 0 && fzL) {
                const int traj = (g0 + j) * B_inst + b;
                const float q = philox ? philox_exp1(A.seed, A.offset, traj, dstep[j], nL)
                                       : A.noise[((size_t)dstep[j] * B_traj + (traj < B_traj ? traj : 0)) * N + nL];
                keyf[j] = z[j] - logf(q);
              }
            }
            ex[j] = (part == 0) ? ex2((z[j] - Zb) * LOG2E) : 0.f;
          }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {  // the Q partial sums, level by level
#pragma unroll
            for (int j = 0; j < Q; ++j) ex[j] += __shfl_xor_sync(FULL, ex[j], off);
          }
          unsigned wkey[Q], vote[Q];
#pragma unroll
          for (int j = 0; j < Q; ++j) wkey[j] = __reduce_max_sync(FULL, fkey(keyf[j]));
#pragma unroll
          for (int j = 0; j < Q; ++j) vote[j] = __ballot_sync(FULL, fkey(keyf[j]) == wkey[j]);
          if (lane == 0) {
#pragma unroll
            for (int j = 0; j < Q; ++j) {
              sm.red_key[j][h] = wkey[j];
              sm.red_idx[j][h] = h * NPW + (__ffs(vote[j]) - 1) / PARTS;
              sm.red_sum[j][h] = ex[j];
            }
          }
        }
        __syncthreads();  // B2: per-warp partials of all trajectories complete
        bool need_sync = false;
#pragma unroll
        for (int j = 0; j < Q; ++j) {
          const int traj = (g0 + j) * B_inst + b;
          int a;
          float Ssum;
          {
            const uint4 k0 = reinterpret_cast<const uint4*>(sm.red_key[j])[0], k1 = reinterpret_cast<const uint4*>(sm.red_key[j])[1];
            const int4 i0 = reinterpret_cast<const int4*>(sm.red_idx[j])[0], i1 = reinterpret_cast<const int4*>(sm.red_idx[j])[1];
            const float4 u0 = reinterpret_cast<const float4*>(sm.red_sum[j])[0], u1 = reinterpret_cast<const float4*>(sm.red_sum[j])[1];
            Ssum = ((u0.x + u0.y) + (u0.z + u0.w)) + ((u1.x + u1.y) + (u1.z + u1.w));
            unsigned bk = k0.x; a = i0.x;
            if (k0.y > bk) { bk = k0.y; a = i0.y; }
            if (k0.z > bk) { bk = k0.z; a = i0.z; }
            if (k0.w > bk) { bk = k0.w; a = i0.w; }
            if (k1.x > bk) { bk = k1.x; a = i1.x; }
            if (k1.y > bk) { bk = k1.y; a = i1.y; }
            if (k1.z > bk) { bk = k1.z; a = i1.z; }
            if (k1.w > bk) { bk = k1.w; a = i1.w; }
          }
          if (!fin[j]) {  // uniform across the block: the state is replicated
            const int t = tstep[j];
            if (MODE == CO_MODE_EVALUATE) {
              const int forced = (int)A.forced_actions[(size_t)traj * T_max + t];
              a = (forced < 0 || forced >= N) ? 0 : forced;
            }
            if (nL == a && part == 0) {
              const float lpL = (z[j] - Zb) - lg2(Ssum) * LN2;
              A.logp_out[(size_t)traj * T_max + t] = lpL;
              sm.ll_acc[j] += lpL;
            }
            if (tid == 0) A.actions_out[(size_t)traj * T_max + t] = a;
            const bool was_first = (ENV == CO_ENV_TSP) && (t == 0);
            env_step(j, a);
            ++dstep[j];
            if (was_first) {  // multisample without forced start: context becomes [h_first ; h_cur]
              if (tid < E) sm.qfix[j][tid] = (A.graph_ctx ? A.graph_ctx[(size_t)b * E + tid] : 0.f) +
                                             __ldg(crow + (size_t)a * CW + 3 * E + tid);
              need_sync = true;
            }
          }
        }
        if (need_sync) {  // uniform: tstep is replicated
          __syncthreads();
#pragma unroll
          for (int j = 0; j < Q; ++j) head_dot(sm.qfix[j], FK[j]);
        }
      }

      // ---------------- epilogue of the group
      __syncthreads();
#pragma unroll
      for (int j = 0; j < Q; ++j) {
        if (g0 + j >= S) continue;
        const int traj = (g0 + j) * B_inst + b;
        const int t = tstep[j];
        if (tid == 0) {
          const float2 pa = sm.loc[(ENV == CO_ENV_TSP) ? first[j] : 0], pp = sm.loc[prev[j]];
          const float dx = pa.x - pp.x, dy = pa.y - pp.y;
          A.reward_out[traj] = -(dist[j] + sqrtf(dx * dx + dy * dy));
          A.loglik_out[traj] = sm.ll_acc[j];
          if (A.steps_out) A.steps_out[traj] = t;
          if (A.used_capacity_out) A.used_capacity_out[traj] = used[j];
          if (A.max_steps_out) atomicMax(A.max_steps_out, t);
        }
        for (int c = t + tid; c < T_max; c += 256) {
          A.actions_out[(size_t)traj * T_max + c] = 0;
          A.logp_out[(size_t)traj * T_max + c] = 0.f;
        }
      }
    }
  }
}

template <int SPL, int ENV, int MODE>
static int launch_ms(const co_rollout_args& A, cudaStream_t st) {
  auto kern = rollout_ms_kernel<SPL, ENV, MODE>;
  const size_t smem = sizeof(SmemMS<SPL>);
  static PerDeviceOnce once;
  bool& configured = once.flag();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(CO_ERR_CUDA, "co_rollout(ms): smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  int grid = device_info().sm_count;
  if (grid > A.B_inst) grid = A.B_inst;
  kern<<<grid, 256, smem, st>>>(A);
  return check_launch("co_rollout(ms)");
}

template <int ENV>
static int dispatch_ms(const co_rollout_args& A, cudaStream_t st) {
  const int spl = A.N <= 32 ? 1 : (A.N <= 64 ? 2 : 4);
  const int mode = A.select_mode == CO_SELECT_GREEDY ? CO_MODE_GREEDY
                   : (A.select_mode == CO_SELECT_EVALUATE ? CO_MODE_EVALUATE : CO_MODE_SAMPLE);
#define CO_CASE(S_, M_) if (spl == S_ && mode == M_) return launch_ms<S_, ENV, M_>(A, st)
  CO_CASE(1, CO_MODE_GREEDY); CO_CASE(2, CO_MODE_GREEDY); CO_CASE(4, CO_MODE_GREEDY);
  CO_CASE(1, CO_MODE_SAMPLE); CO_CASE(2, CO_MODE_SAMPLE); CO_CASE(4, CO_MODE_SAMPLE);
  CO_CASE(1, CO_MODE_EVALUATE); CO_CASE(2, CO_MODE_EVALUATE); CO_CASE(4, CO_MODE_EVALUATE);
#undef CO_CASE
  return fail(CO_ERR_BAD_ARG, "co_rollout(ms): no kernel variant%s");
}

}  // namespace co
