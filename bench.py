#!/usr/bin/env python
"""bench.py -- node-selections/sec of the AM rollout hot path (BASELINE.json metric).

Workload (config.workload): TSP-100 AttentionModel greedy rollout, 65 536 instances per GPU
(north_star's headline; weak scaling: every rank owns its own 65 536 instances, no data-path
collective, one 2-double NCCL all-reduce for the REINFORCE mean baseline per step when N>1).

  value : decode path with inputs RESIDENT in HBM (encoder output h + instance data):
          one step = FusedAttentionModelDecoder._precompute_cache (one tcgen05 3xTF32 GEMM)
                   + co_rollout (persistent kernel: context + glimpse + pointer + tanh/mask/
                     log-softmax + arg-max + env step + incremental tour length, all T steps)
  e2e   : the call a user makes -- policy(td, env, decode_type="greedy") -- from HOST buffers:
          pinned-host locs -> H2D, encoder (Linear layers on co_gemm_tf32x3, attention on
          co_encoder_mha), cache GEMM, co_rollout, D2H of actions + reward + log-likelihood.
  --impl reference : the reference's own algorithm on the host CPU cores (oracle port of the
          rl4co PyTorch path incl. its per-step K/V/L copies), same metric / config.

Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over
ranks.  Inputs (16.8 GB cache, 3.4 GB embeddings per rank) are far larger than the 126 MB L2,
so no explicit L2 flush is needed between iterations (stated in config).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "node-selections/sec TSP-100 AM rollout"
UNIT = "selections/s"
E = 128


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--env", default="tsp", choices=["tsp", "cvrp"])
    p.add_argument("--num-loc", type=int, default=100)
    p.add_argument("--batch", type=int, default=65536, help="instances per GPU")
    p.add_argument("--decode-type", default="greedy")
    p.add_argument("--cpu-batch", type=int, default=1024, help="bounded CPU sample (instances)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    return p.parse_args()


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.05:
                continue
            parts = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(parts[0])); mx = float(parts[1]); power.append(float(parts[2]))
            except Exception:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


# ------------------------------------------------------------------------------------ common
def make_policy_and_data(env_name, num_loc, batch, rank):
    from rl4co_b200.envs import get_env
    from rl4co_b200.policy import FusedAttentionModelPolicy

    torch.manual_seed(0)
    policy = FusedAttentionModelPolicy(env_name=env_name, embed_dim=128, num_heads=8, num_encoder_layers=3,
                                       normalization="batch", tanh_clipping=10.0).eval()
    env = get_env(env_name, generator_params=dict(num_loc=num_loc), check_solution=False)
    torch.manual_seed(1234 + rank)
    td_host = env.generator(batch)
    return policy, env, td_host


def algorithmic_bytes_per_instance(env_name, N, T):
    """read K,V,L + node table rows, graph ctx, coords (+demand); write
    actions (int64) + logp (f32) per step, reward + log-likelihood."""
    b = N * 4 * E * 4 + E * 4 + N * 8 + T * (8 + 4) + 4 + 4
    if env_name == "cvrp":
        b += (N - 1) * 4
    return b


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic(batch):
    """dram__bytes_read+write of the rollout kernel from the committed `ncu --set full` capture,
    scaled per instance to this launch (the kernel streams each instance exactly once)."""
    try:
        with open(os.path.join(ROOT, "profiles", "rollout_traffic.json")) as f:
            return json.load(f)["dram_bytes_per_instance"] * batch
    except Exception:
        return None


# ------------------------------------------------------------------------------------ CPU arm
def cpu_reference_run(env_name, num_loc, batch, decode_type, steps, warmup):
    """The reference's algorithm on the host cores: oracle port (the reference is Python and
    cannot travel to this box; the port is pinned to it by tests/golden)."""
    from oracle import am_rollout_oracle as O

    cores = os.cpu_count() or 1
    torch.manual_seed(0)
    from rl4co_b200.policy import FusedAttentionModelPolicy

    pol = FusedAttentionModelPolicy(env_name=env_name, num_encoder_layers=3).eval()
    W = {k: v.detach() for k, v in pol.state_dict().items()}
    # give the reference its best thread count: torch CPU ops on [B,N,128] tensors stop scaling
    # (and regress) well below a 100+-core host's full width
    probe = O.generate_instances(env_name, min(256, batch), num_loc)
    best = (None, float("inf"))
    with torch.inference_mode():
        for nt in sorted({min(c, cores) for c in (8, 16, 32, 64)}):  # the full width of a 100+-core host is far slower
            torch.set_num_threads(nt)
            st0 = O.env_reset(env_name, probe)
            h, _ = O.encoder_forward(W, env_name, st0, num_layers=3)
            O.rollout(W, env_name, probe, h, decode_type=decode_type, faithful_copies=True)  # warm
            t0 = time.perf_counter()
            st0 = O.env_reset(env_name, probe)
            h, _ = O.encoder_forward(W, env_name, st0, num_layers=3)
            O.rollout(W, env_name, probe, h, decode_type=decode_type, faithful_copies=True)
            dt = time.perf_counter() - t0
            if dt < best[1]:
                best = (nt, dt)
    torch.set_num_threads(best[0])
    torch.manual_seed(1234)
    inst = O.generate_instances(env_name, batch, num_loc)
    times, dec_times, nsel = [], [], 0
    with torch.inference_mode():
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            st0 = O.env_reset(env_name, inst)
            h, _ = O.encoder_forward(W, env_name, st0, num_layers=3)
            t1 = time.perf_counter()
            out = O.rollout(W, env_name, inst, h, decode_type=decode_type, faithful_copies=True)
            t2 = time.perf_counter()
            if it >= warmup:
                times.append(t2 - t0); dec_times.append(t2 - t1)
            nsel = out["actions"].numel()
    return {"selections": nsel, "policy_forward_s": sum(times) / len(times), "decode_only_s": sum(dec_times) / len(dec_times),
            "cores": torch.get_num_threads(), "host_cores": cores, "threads": torch.get_num_threads()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.env, args.num_loc, args.cpu_batch, args.decode_type, args.steps, min(args.warmup, 1))
    val = r["selections"] / r["policy_forward_s"]
    sample = (f"{args.env.upper()}-{args.num_loc} {args.decode_type} policy-forward (encoder+decode+reward), "
              f"B={args.cpu_batch} per step, torch CPU fp32, {r['threads']} threads (best of 8/16/32/64 on a {r['host_cores']}-core host)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": r["policy_forward_s"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.env.upper()}-{args.num_loc} AM {args.decode_type} rollout (reference algorithm, CPU)",
                   "batch_per_step": args.cpu_batch, "scope": "policy-forward"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample,
                         "decode_only_value": r["selections"] / r["decode_only_s"]},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch.distributed as dist

    from rl4co_b200 import native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    native.lib()

    env_name, n, B = args.env, args.num_loc, args.batch
    N = n + (1 if env_name == "cvrp" else 0)
    policy, env, td_host = make_policy_and_data(env_name, n, B, rank)
    policy = policy.to(dev)

    pinned = {k: td_host[k].pin_memory() for k in td_host.keys()}
    stats = torch.zeros(2, dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def baseline_allreduce(reward):
        """REINFORCE mean baseline over the GLOBAL batch: {sum, count} in f64 (north_star)."""
        stats.zero_()
        native.reward_stats(reward, stats)
        if world > 1:
            dist.all_reduce(stats)
        return stats

    # ---- resident inputs for `value`
    from rl4co_b200.tensordict import TensorDict

    with torch.inference_mode():
        td_dev = env.reset(TensorDict({k: v.to(dev) for k, v in pinned.items()}, batch_size=[B]))
        h, _ = policy.encoder(td_dev)
        h = h.contiguous()
    torch.cuda.synchronize()

    rollout_ev = []

    def decode_step(record=False):
        """hot path from resident h: cache GEMM + persistent rollout (+ baseline all-reduce)."""
        with torch.inference_mode():
            cached = policy.decoder._precompute_cache(h)
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            res = native.rollout(env_name, native.SELECT_GREEDY if "greedy" in args.decode_type else native.SELECT_SAMPLE_PHILOX,
                                 cached.rollout_cache, cached.graph_context_or_none, cached.q_placeholder,
                                 cached.w_capacity, td_dev["locs"], td_dev["demand"] if env_name == "cvrp" else None,
                                 td_dev["vehicle_capacity"].reshape(-1) if env_name == "cvrp" else None, B, N,
                                 tanh_clipping=10.0, seed=1)
            if record:
                e1.record()
                rollout_ev.append((e0, e1))
            baseline_allreduce(res["reward"])
        return res

    host_out = {}  # pinned result buffers, reused across steps (a pageable destination is copied through a bounce buffer)

    def e2e_step():
        with torch.inference_mode():
            td = TensorDict({k: v.to(dev, non_blocking=True) for k, v in pinned.items()}, batch_size=[B])
            td = env.reset(td)
            out = policy(td, env, phase="test", decode_type=args.decode_type)
            baseline_allreduce(out["reward"])
            for k in ("actions", "reward", "log_likelihood"):
                if k not in host_out or host_out[k].shape != out[k].shape:
                    host_out[k] = torch.empty(out[k].shape, dtype=out[k].dtype, pin_memory=True)
                host_out[k].copy_(out[k], non_blocking=True)
            torch.cuda.current_stream().synchronize()  # the results are on the host when the step returns
        return host_out

    # ---- warm-up
    for _ in range(max(args.warmup, 3)):
        res = decode_step()
    torch.cuda.synchronize()
    T_steps = res["steps"].sum().item()  # exact number of (decode -> select -> env.step) iterations
    sel_per_step_rank = float(T_steps)

    # ---- timed: value
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launch0 = native.LAUNCH_COUNT
    barrier()
    t_mark0 = sampler.mark() if sampler else 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        decode_step(record=True)
    ev1.record()
    barrier()
    t_mark1 = sampler.mark() if sampler else 0
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_mark0, t_mark1) if sampler else None
    n_launch = native.LAUNCH_COUNT - launch0  # libcorollout kernels: cache GEMM + rollout + reward stats per step
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    sel = torch.tensor([sel_per_step_rank], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(sel)
    ms_total = t.item()
    sel_total_per_step = sel.item()
    value = sel_total_per_step * args.steps / (ms_total * 1e-3)
    k_ms = sorted(a.elapsed_time(b) for a, b in rollout_ev)
    k_ms_avg = sum(k_ms) / len(k_ms)

    # ---- timed: e2e
    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            e2e_step()
        barrier()
        launch1 = native.LAUNCH_COUNT
        ev0.record()
        for _ in range(args.steps):
            host = e2e_step()
        ev1.record()
        barrier()
        e2e_launches = native.LAUNCH_COUNT - launch1
        t2 = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        h2d = sum(v.numel() * v.element_size() for v in pinned.values())
        d2h = sum(v.numel() * v.element_size() for v in host.values())
        e2e = {"value": sel_total_per_step * args.steps / (t2.item() * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": t2.item() / args.steps,
               "gpu_launches": e2e_launches,
               "scope": "policy(td_host, env): H2D + encoder + cache GEMM + rollout + D2H(actions,reward,ll)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_kind = measured_peaks()
    T_inst = N if env_name == "tsp" else None
    bytes_per_launch = algorithmic_bytes_per_instance(env_name, N, N if env_name == "tsp" else sel_per_step_rank / B) * B
    achieved = bytes_per_launch / (k_ms_avg * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "co::rollout_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)",
                "traffic": ncu_traffic(B) if (env_name == "tsp" and n == 100) else None, "kernel_ms": k_ms_avg, "algorithmic_bytes_per_launch": bytes_per_launch,
                "kernel_share_of_step": k_ms_avg * args.steps / ms_total,
                "note": "latency/issue-bound on-chip loop; HBM roofline shown as required, see DESIGN.md"}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(env_name, n, args.cpu_batch, args.decode_type, steps=2, warmup=1)
        cpu_baseline = {
            "value": r["selections"] / r["policy_forward_s"], "unit": UNIT, "cores": r["cores"], "kind": "port",
            "sample": f"{env_name.upper()}-{n} {args.decode_type} policy-forward, B={args.cpu_batch}, torch CPU fp32, "
                      f"{r['threads']} threads (best of 8/16/32/64 on a {r['host_cores']}-core host), mean of 2 after 1 warm-up",
            "decode_only_value": r["selections"] / r["decode_only_s"]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{env_name.upper()}-{n} AM {args.decode_type} rollout, batch {B} per GPU",
                   "global_batch": B * world, "nodes": N, "parallelism": f"dp{world} (instances sharded, no data-path collective)",
                   "value_scope": "precompute_cache GEMM + persistent rollout kernel from resident encoder output",
                   "l2_policy": "inputs (16.8 GB cache/rank) exceed L2; no flush needed",
                   "policy": "AttentionModelPolicy E=128 H=8 L=3 batch-norm random-init seed 0, eval"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": n_launch, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "selections_per_step": sel_total_per_step,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
