// Persistent whole-episode rollout kernel (the north-star kernel).
//
// Replaces the `while not td["done"].all()` loop of ConstructivePolicy.forward
// (rl4co/models/common/constructive/base.py:219-251): per node selection it fuses
//   AttentionModelDecoder.forward      rl4co/models/zoo/am/decoder.py:156-193
//     context embedding                nn/env_embeddings/context.py:61-74,116-134,147-149
//     PointerAttention                 nn/attention.py:274-320
//   DecodingStrategy.step              rl4co/utils/decoding.py:138-188,344-461
//   TSPEnv._step / CVRPEnv._step       envs/routing/tsp/env.py:60-86, cvrp/env.py:66-136
// and at the end get_reward (ops.py:82-90) and get_log_likelihood (decoding.py:38-62).
//
// Design (B200): one CTA (256 threads = 8 warps) owns one instance for its whole episode.
//   * warp h holds head h of glimpse_key / glimpse_val for all nodes IN REGISTERS
//     (lane l owns nodes l, l+32, ..: 2 x SPL x 16 floats) -> the glimpse
//     (scores -> masked softmax -> weighted value sum) is warp-local: no block barrier,
//     only shuffles (a 16-value reduce-scatter + two all-reduces);
//   * logit_key is pre-multiplied by project_out on the host side of the cache
//     (logits = heads . (L @ W_out)[n]), and also lives in registers: thread (node, part)
//     owns 16*SPL contiguous channels of its node;
//   * the per-node context table (node_emb @ Wctx_cur^T) sits in shared memory, so the next
//     query is one row read + the per-episode fixed part;
//   * the visited set is a bitmask in registers (SPL x uint32), replicated in every thread;
//     capacity / current node / distance are replicated scalars: no state in HBM;
//   * two block barriers per step (heads ready; per-warp softmax partials ready) -- plus one
//     for sampling's second arg-max and one __syncthreads_or for the CVRP depot rule.
// HBM traffic per instance = one read of its cache rows + T*(8+4) B of outputs.
#include "co_common.cuh"

namespace co {

template <int SPL>
struct Cfg {
  static constexpr int NS = 32 * SPL;       // node slots
  static constexpr int PARTS = 8 / SPL;     // threads sharing one node in the logits phase
  static constexpr int NPW = 32 / PARTS;    // nodes per warp in the logits phase
  static constexpr int EPP = 16 * SPL;      // channels of logit_key per thread
  static constexpr int OPAD = EPP + 4;      // padded stride of a part's chunk in `o` (bank spread)
  static constexpr int MINB = SPL == 4 ? 1 : (SPL == 2 ? 2 : 3);
};

template <int SPL>
struct Smem {
  float ptab[(32 * SPL + 1) * E];  // current-node context table; last row = zeros
  float qfix[E];                   // per-episode fixed part of the query
  float wcap[E];                   // cvrp: remaining-capacity column of project_context
  float o[8 * (16 * SPL + 4) + 8]; // concatenated heads (padded per part)
  float4 red[8];
  float4 red2[8];
  float dem[32 * SPL];
  float2 loc[32 * SPL];
  float ll_acc;
};

// sum over the 32 lanes of v[d] for 16 values; lane l returns the total for d = (l >> 1) & 15
__device__ __forceinline__ float reduce_scatter16(const float (&v)[16], int lane) {
  float a[8], b[4], c[2];
  bool up = lane & 16;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float send = up ? v[j] : v[j + 8], keep = up ? v[j + 8] : v[j];
    a[j] = keep + __shfl_xor_sync(FULL, send, 16);
  }
  up = lane & 8;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float send = up ? a[j] : a[j + 4], keep = up ? a[j + 4] : a[j];
    b[j] = keep + __shfl_xor_sync(FULL, send, 8);
  }
  up = lane & 4;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    float send = up ? b[j] : b[j + 2], keep = up ? b[j + 2] : b[j];
    c[j] = keep + __shfl_xor_sync(FULL, send, 4);
  }
  up = lane & 2;
  float send = up ? c[0] : c[1], keep = up ? c[1] : c[0];
  float d = keep + __shfl_xor_sync(FULL, send, 2);
  d += __shfl_xor_sync(FULL, d, 1);
  return d;
}

template <int ENV>
__device__ __forceinline__ bool feasible(int n, bool visbit, float d, float used, float thr, int cur, bool anyfeas) {
  if (ENV == CO_ENV_TSP) return !visbit;
  // cvrp/env.py:126-136
  if (n == 0) return !(cur == 0 && anyfeas);
  return !visbit && !((d + used) > thr);
}

template <int SPL, int ENV>
__global__ void __launch_bounds__(256, Cfg<SPL>::MINB) rollout_kernel(const co_rollout_args A) {
  using C = Cfg<SPL>;
  constexpr int NS = C::NS, PARTS = C::PARTS, NPW = C::NPW, EPP = C::EPP, OPAD = C::OPAD;
  constexpr int CW = (ENV == CO_ENV_TSP ? 5 : 4) * E;  // cache row width
  constexpr int CUR_BLK = (ENV == CO_ENV_TSP ? 4 : 3); // block holding the current-node table
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<SPL>& sm = *reinterpret_cast<Smem<SPL>*>(smem_raw);

  const int tid = threadIdx.x, lane = tid & 31, h = tid >> 5;
  const int N = A.N, B_inst = A.B_inst, S = A.num_starts, T_max = A.T_max;
  const int B_traj = B_inst * S;
  const int mode = A.select_mode;
  const bool forced_start = (S > 1) && (A.flags & CO_ROLLOUT_FORCED_START);
  const int nL = h * NPW + lane / PARTS;  // node owned in the logits phase
  const int part = lane % PARTS;
  const float clip = A.tanh_clipping, inv_temp = 1.0f / A.temperature;

  float Kr[SPL][16], Vr[SPL][16], Lr[EPP];

  for (int b = blockIdx.x; b < B_inst; b += gridDim.x) {
    __syncthreads();  // previous instance no longer reads shared memory
    const float* crow = A.cache + (size_t)b * N * CW;
    // ---- one HBM read of the instance: registers <- glimpse_key/val head slices, folded logit key
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
      const int n = lane + 32 * k;
      if (n < N) {
        const float4* ks = reinterpret_cast<const float4*>(crow + (size_t)n * CW + 0 * E + h * D);
        const float4* vs = reinterpret_cast<const float4*>(crow + (size_t)n * CW + 1 * E + h * D);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float4 kv = __ldg(ks + c), vv = __ldg(vs + c);
          Kr[k][4 * c] = kv.x; Kr[k][4 * c + 1] = kv.y; Kr[k][4 * c + 2] = kv.z; Kr[k][4 * c + 3] = kv.w;
          Vr[k][4 * c] = vv.x; Vr[k][4 * c + 1] = vv.y; Vr[k][4 * c + 2] = vv.z; Vr[k][4 * c + 3] = vv.w;
        }
      } else {
#pragma unroll
        for (int d = 0; d < 16; ++d) { Kr[k][d] = 0.f; Vr[k][d] = 0.f; }
      }
    }
    if (nL < N) {
      const float4* ls = reinterpret_cast<const float4*>(crow + (size_t)nL * CW + 2 * E + part * EPP);
#pragma unroll
      for (int c = 0; c < EPP / 4; ++c) {
        float4 lv = __ldg(ls + c);
        Lr[4 * c] = lv.x; Lr[4 * c + 1] = lv.y; Lr[4 * c + 2] = lv.z; Lr[4 * c + 3] = lv.w;
      }
    } else {
#pragma unroll
      for (int c = 0; c < EPP; ++c) Lr[c] = 0.f;
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
    const float thr = cap + 1e-5f;
    __syncthreads();
    float dmk[SPL];
#pragma unroll
    for (int k = 0; k < SPL; ++k) dmk[k] = sm.dem[lane + 32 * k];
    const float dL = sm.dem[nL];

    for (int s = 0; s < S; ++s) {
      const int traj = s * B_inst + b;  // start-major, rl4co/utils/ops.py:10-29
      int64_t* act_row = A.actions_out + (size_t)traj * T_max;
      float* lp_row = A.logp_out + (size_t)traj * T_max;
      // ---------------- reset (tsp/env.py:88-113, cvrp/env.py:98-124)
      uint32_t vis[SPL];
#pragma unroll
      for (int k = 0; k < SPL; ++k) {
        const int lo = 32 * k;
        vis[k] = (N >= lo + 32) ? 0u : (N <= lo ? 0xffffffffu : (0xffffffffu << (N - lo)));
      }
      int cur = (ENV == CO_ENV_TSP) ? NS : 0;  // NS -> zero row: step-0 placeholder context
      int prev = 0, first = 0, t = 0, dstep = 0;
      float used = 0.f, dist = 0.f;
      bool anyfeas = false, done = false;
      __syncthreads();  // previous trajectory finished with qfix / ll_acc
      if (tid < E) {
        float g = A.graph_ctx ? A.graph_ctx[(size_t)b * E + tid] : 0.f;
        if (ENV == CO_ENV_TSP && !forced_start) g += A.q_placeholder[tid];
        sm.qfix[tid] = g;
      }
      if (tid == 0) sm.ll_acc = 0.f;

      // bit of node nL in the replicated visited mask (select chain: no dynamic register indexing)
      auto vis_bit_L = [&]() -> bool {
        uint32_t w = vis[0];
#pragma unroll
        for (int k = 1; k < SPL; ++k)
          if ((nL >> 5) == k) w = vis[k];
        return (w >> (nL & 31)) & 1u;
      };
      // one environment transition, replicated in every thread
      auto env_step = [&](int a) {
#pragma unroll
        for (int k = 0; k < SPL; ++k)
          if ((a >> 5) == k) vis[k] |= 1u << (a & 31);
        const float2 pa = sm.loc[a], pp = sm.loc[prev];
        const float dx = pa.x - pp.x, dy = pa.y - pp.y;
        if (ENV == CO_ENV_TSP) {
          if (t == 0) first = a; else dist += sqrtf(dx * dx + dy * dy);
        } else {
          dist += sqrtf(dx * dx + dy * dy);
          used = (used + sm.dem[a == 0 ? 1 : a]) * (a != 0 ? 1.0f : 0.0f);  // cvrp/env.py:70-76
        }
        prev = a; cur = a; ++t;
        bool all = true;
#pragma unroll
        for (int k = 0; k < SPL; ++k) all = all && (vis[k] == 0xffffffffu);
        done = all;
      };

      if (forced_start) {  // multistart pre_decoder_hook, decoding.py:309-326 + ops.py:128-149
        const int a0 = (s % A.num_loc) + (ENV == CO_ENV_CVRP ? 1 : 0);
        if (tid == 0) { act_row[0] = a0; lp_row[0] = 0.f; }
        env_step(a0);
        if (ENV == CO_ENV_TSP && tid < E) sm.qfix[tid] += __ldg(crow + (size_t)a0 * CW + 3 * E + tid);
      }
      if (ENV == CO_ENV_CVRP) {
        const bool f = (part == 0) && (nL >= 1) && feasible<ENV>(nL, vis_bit_L(), dL, used, thr, cur, false);
        anyfeas = __syncthreads_or(f);
      } else {
        __syncthreads();
      }

      while (!done && t < T_max) {
        // early, latency-tolerant loads for this step
        int forced = -1;
        float qn = 1.f;
        if (mode == CO_SELECT_EVALUATE) forced = (int)A.forced_actions[(size_t)traj * T_max + t];
        if (mode == CO_SELECT_SAMPLE_NOISE && part == 0 && nL < N)
          qn = A.noise[((size_t)dstep * B_traj + traj) * N + nL];

        // ---------------- glimpse: warp h = head h, fully warp-local
        {
          const float4* pr = reinterpret_cast<const float4*>(sm.ptab + cur * E + h * D);
          const float4* qf = reinterpret_cast<const float4*>(sm.qfix + h * D);
          const float4* wc = reinterpret_cast<const float4*>(sm.wcap + h * D);
          const float rem = cap - used;  // context.py:147-149
          float sc[SPL];
#pragma unroll
          for (int k = 0; k < SPL; ++k) sc[k] = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float4 p = pr[c], f = qf[c];
            float q0 = p.x + f.x, q1 = p.y + f.y, q2 = p.z + f.z, q3 = p.w + f.w;
            if (ENV == CO_ENV_CVRP) {
              float4 w = wc[c];
              q0 = fmaf(w.x, rem, q0); q1 = fmaf(w.y, rem, q1); q2 = fmaf(w.z, rem, q2); q3 = fmaf(w.w, rem, q3);
            }
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
              sc[k] = fmaf(q0, Kr[k][4 * c], sc[k]);
              sc[k] = fmaf(q1, Kr[k][4 * c + 1], sc[k]);
              sc[k] = fmaf(q2, Kr[k][4 * c + 2], sc[k]);
              sc[k] = fmaf(q3, Kr[k][4 * c + 3], sc[k]);
            }
          }
          float m = -INFINITY;
          bool fz[SPL];
#pragma unroll
          for (int k = 0; k < SPL; ++k) {
            fz[k] = feasible<ENV>(lane + 32 * k, (vis[k] >> lane) & 1u, dmk[k], used, thr, cur, anyfeas);
            sc[k] = fz[k] ? sc[k] * 0.25f : -INFINITY;  // 1/sqrt(head_dim)
            m = fmaxf(m, sc[k]);
          }
          m = warp_max(m);
          float acc[16], esum = 0.f;
#pragma unroll
          for (int d = 0; d < 16; ++d) acc[d] = 0.f;
#pragma unroll
          for (int k = 0; k < SPL; ++k) {
            const float e = fz[k] ? __expf(sc[k] - m) : 0.f;
            esum += e;
#pragma unroll
            for (int d = 0; d < 16; ++d) acc[d] = fmaf(e, Vr[k][d], acc[d]);
          }
          esum = warp_sum(esum);
          const float r = reduce_scatter16(acc, lane);
          if (!(lane & 1)) {
            const int e = h * D + ((lane >> 1) & 15);
            sm.o[e + 4 * (e / EPP)] = r / esum;
          }
        }
        __syncthreads();  // B1: heads complete

        // ---------------- pointer logits + tanh clip + mask: thread (nL, part)
        const bool fzL = feasible<ENV>(nL, vis_bit_L(), dL, used, thr, cur, anyfeas);
        float z;
        {
          const float4* ov = reinterpret_cast<const float4*>(sm.o + part * OPAD);
          float p = 0.f;
#pragma unroll
          for (int c = 0; c < EPP / 4; ++c) {
            const float4 x = ov[c];
            p = fmaf(x.x, Lr[4 * c], p); p = fmaf(x.y, Lr[4 * c + 1], p);
            p = fmaf(x.z, Lr[4 * c + 2], p); p = fmaf(x.w, Lr[4 * c + 3], p);
          }
#pragma unroll
          for (int off = PARTS / 2; off > 0; off >>= 1) p += __shfl_xor_sync(FULL, p, off);
          float lg = p / 11.313708498984761f;          // / sqrt(embed_dim), attention.py:291-293
          if (clip > 0.f) lg = tanhf(lg) * clip;       // decoding.py:169-170
          z = fzL ? lg * inv_temp : -INFINITY;         // decoding.py:173-177
        }
        {
          const float wmax = warp_max(z);
          const float ex = (part == 0 && z > -INFINITY) ? expf(z - wmax) : 0.f;
          const float wsum = warp_sum(ex);
          float bv = z; int bi = nL;
          warp_argmax(bv, bi);
          if (lane == 0) sm.red[h] = make_float4(wmax, wsum, __int_as_float(bi), 0.f);
        }
        __syncthreads();  // B2: per-warp partials complete
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < 8; ++w) M = fmaxf(M, sm.red[w].x);
        float Ssum = 0.f;
        int a = -1;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const float4 r = sm.red[w];
          if (r.x > -INFINITY) Ssum += r.y * expf(r.x - M);
          if (a < 0 && r.x == M) a = __float_as_int(r.z);
        }
        const float logS = logf(Ssum);
        const float lpL = (z - M) - logS;  // log_softmax of this thread's node
        if (mode == CO_SELECT_EVALUATE) {
          a = forced;
        } else if (mode == CO_SELECT_SAMPLE_NOISE || mode == CO_SELECT_SAMPLE_PHILOX) {
          // torch.multinomial(p, 1) == argmax(p / q), q ~ Exp(1)
          if (mode == CO_SELECT_SAMPLE_PHILOX) qn = philox_exp1(A.seed, A.offset, traj, dstep, nL);
          float key = (nL < N && part == 0) ? expf(lpL) / qn : -1.f;
          int ki = nL;
          warp_argmax(key, ki);
          if (lane == 0) sm.red2[h] = make_float4(key, __int_as_float(ki), 0.f, 0.f);
          __syncthreads();  // B2b
          float bk = -2.f;
#pragma unroll
          for (int w = 0; w < 8; ++w) {
            const float4 r = sm.red2[w];
            if (r.x > bk) { bk = r.x; a = __float_as_int(r.y); }
          }
        }
        if (a < 0 || a >= N) a = 0;  // malformed forced action: stay in range
        if (nL == a && part == 0) { lp_row[t] = lpL; sm.ll_acc += lpL; }
        if (tid == 0) act_row[t] = a;

        // ---------------- environment step
        const bool was_first = (ENV == CO_ENV_TSP) && (t == 0);
        env_step(a);
        ++dstep;
        if (was_first) {  // context from now on: [h_first ; h_cur], context.py:129-133
          if (tid < E) sm.qfix[tid] = (A.graph_ctx ? A.graph_ctx[(size_t)b * E + tid] : 0.f) +
                                      __ldg(crow + (size_t)a * CW + 3 * E + tid);
          __syncthreads();
        }
        if (ENV == CO_ENV_CVRP) {
          const bool f = (part == 0) && (nL >= 1) && feasible<ENV>(nL, vis_bit_L(), dL, used, thr, cur, false);
          anyfeas = __syncthreads_or(f);
        }
      }

      // ---------------- epilogue: reward, log-likelihood, padding
      __syncthreads();
      if (tid == 0) {
        const float2 pa = sm.loc[(ENV == CO_ENV_TSP) ? first : 0], pp = sm.loc[prev];
        const float dx = pa.x - pp.x, dy = pa.y - pp.y;
        A.reward_out[traj] = -(dist + sqrtf(dx * dx + dy * dy));
        A.loglik_out[traj] = sm.ll_acc;
        if (A.steps_out) A.steps_out[traj] = t;
        if (A.used_capacity_out) A.used_capacity_out[traj] = used;
        if (A.max_steps_out) atomicMax(A.max_steps_out, t);
      }
      // done instances keep selecting the depot with log-prob 0 until the batch finishes
      for (int c = t + tid; c < T_max; c += 256) { act_row[c] = 0; lp_row[c] = 0.f; }
    }
  }
}

template <int SPL, int ENV>
static int launch(const co_rollout_args& A, cudaStream_t st) {
  auto kern = rollout_kernel<SPL, ENV>;
  const size_t smem = sizeof(Smem<SPL>);
  static bool configured = false;
  static int ctas_per_sm = 1;
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

}  // namespace co

using namespace co;

extern "C" int co_cache_width(int env_kind) {
  return env_kind == CO_ENV_TSP ? 5 * E : (env_kind == CO_ENV_CVRP ? 4 * E : 0);
}
extern "C" int co_rollout_max_nodes(void) { return 128; }

extern "C" int co_rollout(const co_rollout_args* args, void* stream) {
  if (!args) return fail(CO_ERR_BAD_ARG, "co_rollout: null args%s");
  const co_rollout_args& A = *args;
  if (!A.cache || !A.locs || !A.actions_out || !A.logp_out || !A.reward_out || !A.loglik_out)
    return fail(CO_ERR_BAD_ARG, "co_rollout: null pointer%s");
  if (A.B_inst < 0 || A.N < 2 || A.num_starts < 1 || A.T_max < 1)
    return fail(CO_ERR_BAD_ARG, "co_rollout: bad shape%s B=%lld N=%lld", "", A.B_inst, A.N);
  if (A.N > co_rollout_max_nodes()) return fail(CO_ERR_UNSUPPORTED, "co_rollout: N=%s%lld > 128 nodes", "", A.N);
  if (!(A.temperature > 0.f)) return fail(CO_ERR_BAD_ARG, "co_rollout: temperature must be > 0%s");
  if (A.select_mode < 0 || A.select_mode > 3) return fail(CO_ERR_BAD_ARG, "co_rollout: bad select_mode%s");
  if (A.select_mode == CO_SELECT_EVALUATE && !A.forced_actions) return fail(CO_ERR_BAD_ARG, "co_rollout: forced_actions required%s");
  if (A.select_mode == CO_SELECT_SAMPLE_NOISE && !A.noise) return fail(CO_ERR_BAD_ARG, "co_rollout: noise required%s");
  if ((A.flags & CO_ROLLOUT_FORCED_START) && A.num_loc < 1) return fail(CO_ERR_BAD_ARG, "co_rollout: num_loc required for forced starts%s");
  if (A.env_kind == CO_ENV_TSP) {
    if (!A.q_placeholder) return fail(CO_ERR_BAD_ARG, "co_rollout: q_placeholder required for tsp%s");
    if (A.T_max < A.N) return fail(CO_ERR_BAD_ARG, "co_rollout: T_max < N%s");
  } else if (A.env_kind == CO_ENV_CVRP) {
    if (!A.demand || !A.w_capacity) return fail(CO_ERR_BAD_ARG, "co_rollout: demand / w_capacity required for cvrp%s");
  } else {
    return fail(CO_ERR_BAD_ARG, "co_rollout: unknown env kind%s");
  }
  if (A.B_inst == 0) return CO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int spl = A.N <= 32 ? 1 : (A.N <= 64 ? 2 : 4);
  if (A.env_kind == CO_ENV_TSP) {
    if (spl == 1) return launch<1, CO_ENV_TSP>(A, st);
    if (spl == 2) return launch<2, CO_ENV_TSP>(A, st);
    return launch<4, CO_ENV_TSP>(A, st);
  }
  if (spl == 1) return launch<1, CO_ENV_CVRP>(A, st);
  if (spl == 2) return launch<2, CO_ENV_CVRP>(A, st);
  return launch<4, CO_ENV_CVRP>(A, st);
}
