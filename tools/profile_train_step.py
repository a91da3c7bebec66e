"""Where one differentiable chunk of the C5 training step (CVRP-100 REINFORCE) spends its time on the GPU.
usage: python tools/profile_train_step.py [B] [N] [--topk K]
Prints CUDA-event timings of the stages and the torch.profiler table of the top CUDA kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rl4co_b200.envs import get_env
from rl4co_b200.policy import FusedAttentionModelPolicy
from rl4co_b200.reinforce import calculate_loss, evaluate_log_likelihood

args = [a for a in sys.argv[1:] if not a.startswith("--")]
B = int(args[0]) if len(args) > 0 else 8192
N = int(args[1]) if len(args) > 1 else 100
env_name = os.environ.get("ENV", "cvrp")
dev = torch.device("cuda:0")
torch.manual_seed(0)
torch.set_float32_matmul_precision(os.environ.get("PREC", "medium"))
pol = FusedAttentionModelPolicy(env_name=env_name).to(dev).train()
env = get_env(env_name, generator_params=dict(num_loc=N), check_solution=False)
td = env.reset(env.generator(B).to(dev))
opt = torch.optim.Adam(pol.parameters(), lr=1e-4)


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def one(prof=False):
    t = [ev()]
    enc = pol.encoder(td)
    t.append(ev())
    with torch.no_grad():
        out = pol(td, env, phase="train", decode_type="sampling", encoder_output=(enc[0].detach(), enc[1]), seed=1)
    t.append(ev())
    ll = evaluate_log_likelihood(pol, td, env, out["actions"], hidden=enc[0])
    t.append(ev())
    loss, _ = calculate_loss(out["reward"], ll, out["reward"].mean())
    opt.zero_grad(set_to_none=True)
    loss.backward()
    t.append(ev())
    opt.step()
    t.append(ev())
    torch.cuda.synchronize()
    names = ["encoder fwd (grad)", "sampling rollout (no grad)", "evaluate_ll fwd", "backward", "optimizer"]
    return {n: t[i].elapsed_time(t[i + 1]) for i, n in enumerate(names)}, out["actions"].shape


for _ in range(2):
    one()
res, shp = one()
print(f"env={env_name} B={B} N={N} actions{tuple(shp)} precision={torch.get_float32_matmul_precision()}")
for k, v in res.items():
    print(f"  {k:32s} {v:8.2f} ms")
print(f"  {'total':32s} {sum(res.values()):8.2f} ms")
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    one()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
