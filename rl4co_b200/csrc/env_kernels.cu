// Stateless environment kernels behind FusedTSPEnv / FusedCVRPEnv (step-at-a-time API).
// Pure HBM-streaming byte/float work: one warp per instance row, coalesced row reads.
//   co_tsp_step          <- rl4co/envs/routing/tsp/env.py:60-86
//   co_cvrp_action_mask  <- rl4co/envs/routing/cvrp/env.py:126-136
//   co_cvrp_step         <- rl4co/envs/routing/cvrp/env.py:66-96
//   co_tour_length       <- tsp/env.py:150-156, cvrp/env.py:138-147, rl4co/utils/ops.py:54-90
//   co_check_tours       <- tsp/env.py:158-164, cvrp/env.py:149-177
//   co_reward_stats      <- rl4co/models/rl/reinforce/baselines.py:75-81 (mean baseline)
//   co_sdvrp_step / co_sdvrp_action_mask <- rl4co/envs/routing/sdvrp/env.py:55-82,110-116 (sibling env: split deliveries)
#include "co_common.cuh"

namespace co {

constexpr int ROWS_PER_CTA = 8;  // 8 warps / CTA, one row each

// Copy one row of 0/1 bytes (bool mask / uint8 visited) with byte `set_idx` forced to `set_val`, 16 bytes per
// lane wherever the row allows it.  Rows start at arbitrary byte offsets (row * N), so the row is cut at the
// 16-byte boundaries of the address space: head and tail pieces move byte-wise, every interior word is one
// LDG.128 + one STG.128 (N = 100: 5-6 vector words + <= 30 bytes instead of 100 byte loads + 100 byte stores).
// Returns this lane's count of non-zero output bytes (values are 0 / 1).
__device__ __forceinline__ int copy_row_set_byte(const uint8_t* in, uint8_t* out, int N, int set_idx, uint8_t set_val,
                                                 int lane) {
  const uintptr_t ai = reinterpret_cast<uintptr_t>(in), ao = reinterpret_cast<uintptr_t>(out);
  int count = 0;
  if (((ai ^ ao) & 15) != 0 || N < 48) {  // different misalignment of source and destination, or a short row
    for (int n = lane; n < N; n += 32) {
      const uint8_t v = (n == set_idx) ? set_val : in[n];
      out[n] = v;
      count += v;
    }
    return count;
  }
  const int head = (int)((16 - (ai & 15)) & 15);          // bytes before the first aligned word
  const int words = (N - head) >> 4;                      // whole 16-byte words inside the row
  const int tail0 = head + (words << 4);                  // first byte after the last whole word
  for (int n = lane; n < head; n += 32) {
    const uint8_t v = (n == set_idx) ? set_val : in[n];
    out[n] = v;
    count += v;
  }
  for (int n = tail0 + lane; n < N; n += 32) {
    const uint8_t v = (n == set_idx) ? set_val : in[n];
    out[n] = v;
    count += v;
  }
  const uint4* iw = reinterpret_cast<const uint4*>(in + head);
  uint4* ow = reinterpret_cast<uint4*>(out + head);
  for (int w = lane; w < words; w += 32) {
    uint4 v = iw[w];
    const int rel = set_idx - (head + (w << 4));          // position of the forced byte inside this word
    if (rel >= 0 && rel < 16) {
      uint32_t* c = reinterpret_cast<uint32_t*>(&v);
      const int sh = 8 * (rel & 3);
      c[rel >> 2] = (c[rel >> 2] & ~(0xffu << sh)) | ((uint32_t)set_val << sh);
    }
    ow[w] = v;
    count += __popc(v.x & 0x01010101u) + __popc(v.y & 0x01010101u) + __popc(v.z & 0x01010101u) +
             __popc(v.w & 0x01010101u);
  }
  return count;
}

__global__ void __launch_bounds__(256) tsp_step_kernel(const int64_t* __restrict__ action,
                                                        const uint8_t* mask_in, uint8_t* mask_out,
                                                        int64_t* first_node, int64_t* current_node,
                                                        int64_t* i, uint8_t* done, int B, int N) {
  int row = blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= B) return;
  int a = (int)action[row];
  // action_mask.scatter(-1, action, 0) as a row copy with one byte cleared (16-byte words where possible)
  int left = copy_row_set_byte(mask_in + (size_t)row * N, mask_out + (size_t)row * N, N, a, 0, lane);
  left = __reduce_add_sync(FULL, left);
  if (lane == 0) {
    done[row] = (left == 0);
    int64_t iv = i[row];
    // reference: `action if td["i"].all() == 0 else first_node` is a batch-level test that is
    // true exactly on the first step of a lock-step batch; evaluated per instance here.
    if (iv == 0) first_node[row] = a;
    current_node[row] = a;
    i[row] = iv + 1;
  }
}

// shared row logic of get_action_mask; returns through mask_out. One warp per row.
__device__ __forceinline__ void cvrp_mask_row(const float* __restrict__ demand, float used, float cap,
                                              const uint8_t* visited, int cur, uint8_t* mask_out, int N,
                                              int lane) {
  const float thr = cap + 1e-5f;  // fp32 add, as `td["vehicle_capacity"] + 1e-5`
  int any_free = 0;
  for (int n = 1 + lane; n < N; n += 32) {
    bool exceeds = (demand[n - 1] + used) > thr;
    bool masked = (visited[n] != 0) || exceeds;
    mask_out[n] = masked ? 0 : 1;
    any_free |= masked ? 0 : 1;
  }
  any_free = __any_sync(FULL, any_free);
  if (lane == 0) mask_out[0] = ((cur == 0) && any_free) ? 0 : 1;
}

__global__ void __launch_bounds__(256) cvrp_mask_kernel(const float* __restrict__ demand,
                                                         const float* __restrict__ used,
                                                         const float* __restrict__ cap,
                                                         const uint8_t* __restrict__ visited,
                                                         const int64_t* __restrict__ current_node,
                                                         uint8_t* mask_out, int B, int N) {
  int row = blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= B) return;
  cvrp_mask_row(demand + (size_t)row * (N - 1), used[row], cap[row], visited + (size_t)row * N,
                (int)current_node[row], mask_out + (size_t)row * N, N, lane);
}

__global__ void __launch_bounds__(256) cvrp_step_kernel(const int64_t* __restrict__ action,
                                                         const float* __restrict__ demand,
                                                         const float* __restrict__ cap, const float* used_in,
                                                         float* used_out, const uint8_t* visited_in,
                                                         uint8_t* visited_out, int64_t* current_node,
                                                         uint8_t* done, uint8_t* mask_out, int B, int N) {
  int row = blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= B) return;
  int a = (int)action[row];
  const float* dem = demand + (size_t)row * (N - 1);
  int didx = min(max(a - 1, 0), N - 2);
  // (used + selected_demand) * (current_node != 0).float()
  float used = (used_in[row] + dem[didx]) * (a != 0 ? 1.0f : 0.0f);
  const uint8_t* vi = visited_in + (size_t)row * N;
  uint8_t* vo = visited_out + (size_t)row * N;
  // visited.scatter(-1, action, 1) and visited.sum(-1) (sums the uint8 values, which are 0 / 1)
  int cnt = copy_row_set_byte(vi, vo, N, a, 1, lane);
  cnt = __reduce_add_sync(FULL, cnt);
  __syncwarp();
  if (lane == 0) {
    used_out[row] = used;
    current_node[row] = a;
    done[row] = (cnt == N);
  }
  cvrp_mask_row(dem, used, cap[row], vo, a, mask_out + (size_t)row * N, N, lane);
}

// ---- SDVRP (rl4co/envs/routing/sdvrp/env.py): nodes may be revisited, the remaining demand is the dynamic state
// get_action_mask, sdvrp/env.py:110-116: mask_loc = (demand == 0) | (used >= capacity); depot rule as in CVRP
__device__ __forceinline__ void sdvrp_mask_row(const float* __restrict__ dwd, float used, float cap, int cur,
                                               uint8_t* mask_out, int N, int lane) {
  const bool full = used >= cap;
  int any_free = 0;
  for (int n = 1 + lane; n < N; n += 32) {
    const bool masked = (dwd[n] == 0.0f) || full;
    mask_out[n] = masked ? 0 : 1;
    any_free |= masked ? 0 : 1;
  }
  any_free = __any_sync(FULL, any_free);
  if (lane == 0) mask_out[0] = ((cur == 0) && any_free) ? 0 : 1;
}

__global__ void __launch_bounds__(256) sdvrp_mask_kernel(const float* __restrict__ dwd, const float* __restrict__ used,
                                                          const float* __restrict__ cap,
                                                          const int64_t* __restrict__ current_node, uint8_t* mask_out,
                                                          int B, int N) {
  int row = blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= B) return;
  sdvrp_mask_row(dwd + (size_t)row * N, used[row], cap[row], (int)current_node[row], mask_out + (size_t)row * N, N, lane);
}

// _step, sdvrp/env.py:55-82: delivered = min(demand[a], capacity - used); used = (used + delivered) * (a != 0);
// demand[a] -= delivered (scatter_add of -delivered); done = no positive demand left; then the mask
__global__ void __launch_bounds__(256) sdvrp_step_kernel(const int64_t* __restrict__ action, const float* dwd_in,
                                                          float* dwd_out, const float* __restrict__ cap,
                                                          const float* used_in, float* used_out, int64_t* current_node,
                                                          uint8_t* done, uint8_t* mask_out, int B, int N) {
  int row = blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= B) return;
  const int a = (int)action[row];
  const float* di = dwd_in + (size_t)row * N;
  float* dout = dwd_out + (size_t)row * N;
  const float c = cap[row], u = used_in[row];
  const float delivered = fminf(di[a], c - u);
  const float used = (u + delivered) * (a != 0 ? 1.0f : 0.0f);
  int positive = 0;
  for (int n = lane; n < N; n += 32) {
    float v = di[n];
    if (n == a) v = v + (-delivered);  // scatter_add(-1, a, -delivered)
    dout[n] = v;
    positive |= (v > 0.0f) ? 1 : 0;
  }
  positive = __any_sync(FULL, positive);
  __syncwarp();
  if (lane == 0) {
    used_out[row] = used;
    current_node[row] = a;
    done[row] = positive ? 0 : 1;
  }
  sdvrp_mask_row(dout, used, c, a, mask_out + (size_t)row * N, N, lane);
}

// reward = -(cyclic tour length); one warp per trajectory, lanes stride the T edges.
__global__ void __launch_bounds__(256) tour_length_kernel(const float2* __restrict__ locs,
                                                           const int64_t* __restrict__ actions,
                                                           float* __restrict__ reward, int B, int B_locs, int N,
                                                           int T, int with_depot) {
  int row = blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float2* x = locs + (size_t)(row % B_locs) * N;
  const int64_t* a = actions + (size_t)row * T;
  // ordered tour: [depot?] a_0 .. a_{T-1}; edge k joins tour[k] and tour[(k+1) % L]
  const int L = T + (with_depot ? 1 : 0);
  float acc = 0.f;
  // the coordinate row (N * 8 bytes) does not depend on the actions: pull its lines towards L1 now, so that the two
  // HBM round trips (actions, then gathered coordinates) overlap instead of following each other
  for (int ln = lane; ln * 16 < N; ln += 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(x + ln * 16));
  // every action is read from HBM once: lane l holds tour[k] for k = base + l; its successor comes from lane l + 1
  // (the chunk's last lane reads one element ahead), so each 32-edge chunk costs one coalesced 256-byte load
  for (int base = 0; base < L; base += 32) {
    const int k = base + lane;
    int n0 = 0;
    if (k < L) n0 = with_depot ? (k == 0 ? 0 : (int)a[k - 1]) : (int)a[k];
    int n1 = __shfl_down_sync(FULL, n0, 1);
    if (lane == 31 || k + 1 >= L) {
      const int k1 = (k + 1 >= L) ? 0 : k + 1;
      n1 = with_depot ? (k1 == 0 ? 0 : (int)a[k1 - 1]) : (int)a[k1];
    }
    if (k < L) {
      const float2 p0 = x[n0], p1 = x[n1];
      const float dx = p1.x - p0.x, dy = p1.y - p0.y;
      acc += sqrtf(dx * dx + dy * dy);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) reward[row] = -acc;
}

// validity: warp per trajectory.
__global__ void __launch_bounds__(256) check_tours_kernel(const int64_t* __restrict__ actions,
                                                           const float* __restrict__ demand,
                                                           const float* __restrict__ cap, int32_t* bad_count,
                                                           int B, int B_inst, int N, int T) {
  __shared__ uint32_t seen[ROWS_PER_CTA][64];  // up to 2048 nodes
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int row = blockIdx.x * ROWS_PER_CTA + w;
  if (row >= B) return;
  for (int k = lane; k < 64; k += 32) seen[w][k] = 0;
  __syncwarp();
  const int64_t* a = actions + (size_t)row * T;
  int bad = 0;
  const bool cvrp = demand != nullptr;
  for (int t = lane; t < T; t += 32) {
    long long n = a[t];
    if (n < 0 || n >= N) { bad = 1; continue; }
    if (cvrp && n == 0) continue;  // depot may repeat
    uint32_t bit = 1u << (n & 31);
    uint32_t old = atomicOr(&seen[w][n >> 5], bit);
    if (old & bit) bad = 1;  // visited twice
  }
  __syncwarp();
  // every (customer) node exactly once
  for (int n = (cvrp ? 1 : 0) + lane; n < N; n += 32)
    if (!((seen[w][n >> 5] >> (n & 31)) & 1u)) bad = 1;
  if (!cvrp && T != N) bad = 1;
  if (cvrp && lane == 0) {
    const float* dem = demand + (size_t)(row % B_inst) * (N - 1);
    const float c = cap ? cap[row % B_inst] : 1.0f;
    float used = 0.f;
    for (int t = 0; t < T; ++t) {
      long long n = a[t];
      if (n < 0 || n >= N) break;
      used += (n == 0) ? -c : dem[n - 1];
      if (used < 0.f) used = 0.f;
      if (!(used <= c + 1e-5f)) { bad = 1; break; }
    }
  }
  bad = __any_sync(FULL, bad);
  if (lane == 0 && bad) atomicAdd(bad_count, 1);
}

__global__ void __launch_bounds__(256) reward_stats_kernel(const float* __restrict__ reward, double* out2, int B) {
  double s = 0.0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < B; k += gridDim.x * blockDim.x) s += (double)reward[k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += part[k];
    atomicAdd(out2, t);
    if (blockIdx.x == 0) atomicAdd(out2 + 1, (double)B);
  }
}

}  // namespace co

using namespace co;

static inline int rows_grid(int B) { return (B + ROWS_PER_CTA - 1) / ROWS_PER_CTA; }

extern "C" int co_tsp_step(const int64_t* action, const uint8_t* mask_in, uint8_t* mask_out,
                           int64_t* first_node, int64_t* current_node, int64_t* i, uint8_t* done, int B,
                           int N, void* stream) {
  if (!action || !mask_in || !mask_out || !first_node || !current_node || !i || !done)
    return fail(CO_ERR_BAD_ARG, "co_tsp_step: null pointer%s");
  if (B < 0 || N <= 0) return fail(CO_ERR_BAD_ARG, "co_tsp_step: bad shape%s B=%lld N=%lld", "", B, N);
  if (B == 0) return CO_OK;
  tsp_step_kernel<<<rows_grid(B), 256, 0, (cudaStream_t)stream>>>(action, mask_in, mask_out, first_node,
                                                                  current_node, i, done, B, N);
  return check_launch("co_tsp_step");
}

extern "C" int co_cvrp_action_mask(const float* demand, const float* used_capacity,
                                   const float* vehicle_capacity, const uint8_t* visited,
                                   const int64_t* current_node, uint8_t* mask_out, int B, int N,
                                   void* stream) {
  if (!demand || !used_capacity || !vehicle_capacity || !visited || !current_node || !mask_out)
    return fail(CO_ERR_BAD_ARG, "co_cvrp_action_mask: null pointer%s");
  if (B < 0 || N < 2) return fail(CO_ERR_BAD_ARG, "co_cvrp_action_mask: bad shape%s B=%lld N=%lld", "", B, N);
  if (B == 0) return CO_OK;
  cvrp_mask_kernel<<<rows_grid(B), 256, 0, (cudaStream_t)stream>>>(demand, used_capacity, vehicle_capacity,
                                                                   visited, current_node, mask_out, B, N);
  return check_launch("co_cvrp_action_mask");
}

extern "C" int co_cvrp_step(const int64_t* action, const float* demand, const float* vehicle_capacity,
                            const float* used_in, float* used_out, const uint8_t* visited_in,
                            uint8_t* visited_out, int64_t* current_node, uint8_t* done, uint8_t* mask_out,
                            int B, int N, void* stream) {
  if (!action || !demand || !vehicle_capacity || !used_in || !used_out || !visited_in || !visited_out ||
      !current_node || !done || !mask_out)
    return fail(CO_ERR_BAD_ARG, "co_cvrp_step: null pointer%s");
  if (B < 0 || N < 2) return fail(CO_ERR_BAD_ARG, "co_cvrp_step: bad shape%s B=%lld N=%lld", "", B, N);
  if (B == 0) return CO_OK;
  cvrp_step_kernel<<<rows_grid(B), 256, 0, (cudaStream_t)stream>>>(action, demand, vehicle_capacity, used_in,
                                                                   used_out, visited_in, visited_out,
                                                                   current_node, done, mask_out, B, N);
  return check_launch("co_cvrp_step");
}

extern "C" int co_sdvrp_action_mask(const float* demand_with_depot, const float* used_capacity,
                                    const float* vehicle_capacity, const int64_t* current_node, uint8_t* mask_out,
                                    int B, int N, void* stream) {
  if (!demand_with_depot || !used_capacity || !vehicle_capacity || !current_node || !mask_out)
    return fail(CO_ERR_BAD_ARG, "co_sdvrp_action_mask: null pointer%s");
  if (B < 0 || N < 2) return fail(CO_ERR_BAD_ARG, "co_sdvrp_action_mask: bad shape%s B=%lld N=%lld", "", B, N);
  if (B == 0) return CO_OK;
  sdvrp_mask_kernel<<<rows_grid(B), 256, 0, (cudaStream_t)stream>>>(demand_with_depot, used_capacity, vehicle_capacity,
                                                                    current_node, mask_out, B, N);
  return check_launch("co_sdvrp_action_mask");
}

extern "C" int co_sdvrp_step(const int64_t* action, const float* demand_in, float* demand_out,
                             const float* vehicle_capacity, const float* used_in, float* used_out,
                             int64_t* current_node, uint8_t* done, uint8_t* mask_out, int B, int N, void* stream) {
  if (!action || !demand_in || !demand_out || !vehicle_capacity || !used_in || !used_out || !current_node || !done ||
      !mask_out)
    return fail(CO_ERR_BAD_ARG, "co_sdvrp_step: null pointer%s");
  if (B < 0 || N < 2) return fail(CO_ERR_BAD_ARG, "co_sdvrp_step: bad shape%s B=%lld N=%lld", "", B, N);
  if (B == 0) return CO_OK;
  sdvrp_step_kernel<<<rows_grid(B), 256, 0, (cudaStream_t)stream>>>(action, demand_in, demand_out, vehicle_capacity,
                                                                    used_in, used_out, current_node, done, mask_out, B, N);
  return check_launch("co_sdvrp_step");
}

extern "C" int co_tour_length(const float* locs, const int64_t* actions, float* reward, int B, int B_locs,
                              int N, int T, int with_depot, void* stream) {
  if (!locs || !actions || !reward) return fail(CO_ERR_BAD_ARG, "co_tour_length: null pointer%s");
  if (B < 0 || B_locs <= 0 || N <= 0 || T <= 0 || (B % B_locs) != 0)
    return fail(CO_ERR_BAD_ARG, "co_tour_length: bad shape%s B=%lld B_locs=%lld", "", B, B_locs);
  if (B == 0) return CO_OK;
  tour_length_kernel<<<rows_grid(B), 256, 0, (cudaStream_t)stream>>>((const float2*)locs, actions, reward, B,
                                                                     B_locs, N, T, with_depot);
  return check_launch("co_tour_length");
}

extern "C" int co_check_tours(const int64_t* actions, const float* demand, const float* vehicle_capacity,
                              int32_t* bad_count, int B, int B_inst, int N, int T, void* stream) {
  if (!actions || !bad_count) return fail(CO_ERR_BAD_ARG, "co_check_tours: null pointer%s");
  if (N > 2048) return fail(CO_ERR_UNSUPPORTED, "co_check_tours: N > 2048%s");
  if (B < 0 || B_inst <= 0 || T <= 0) return fail(CO_ERR_BAD_ARG, "co_check_tours: bad shape%s");
  if (B == 0) return CO_OK;
  check_tours_kernel<<<rows_grid(B), 256, 0, (cudaStream_t)stream>>>(actions, demand, vehicle_capacity,
                                                                     bad_count, B, B_inst, N, T);
  return check_launch("co_check_tours");
}

extern "C" int co_reward_stats(const float* reward, double* out2, int B, void* stream) {
  if (!reward || !out2) return fail(CO_ERR_BAD_ARG, "co_reward_stats: null pointer%s");
  if (B <= 0) return CO_OK;
  int grid = min((B + 255) / 256, 592);
  reward_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reward, out2, B);
  return check_launch("co_reward_stats");
}
