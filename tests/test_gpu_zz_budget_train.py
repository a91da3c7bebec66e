"""GPU: REINFORCE glue of the orienteering / prize-collecting / split-delivery envs -- the differentiable one-call
teacher-forced pass (reinforce.replay_budget_states + co_attn_fwd / co_attn_bwd; sdvrp: replay_split_delivery_states +
rank-one dynamic terms) equals what the persistent rollout kernel reported for the same actions, and one training step
runs end to end.  (Named to sort last: the kernels' own parity tests come first.)"""

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(env_name, n, B, seed=0):
    from rl4co_b200.envs import get_env
    from rl4co_b200.policy import FusedAttentionModelPolicy

    torch.manual_seed(seed)
    gp = dict(num_loc=n)
    if env_name == "op":
        gp["prize_type"] = "dist"
    env = get_env(env_name, generator_params=gp, check_solution=True)
    pol = FusedAttentionModelPolicy(env_name=env_name, num_encoder_layers=1).to(DEV).eval()
    td = env.reset(env.generator(B).to(DEV))
    return env, pol, td


@pytest.mark.parametrize("env_name,n", [("op", 20), ("pctsp", 20), ("op", 50), ("pctsp", 50), ("sdvrp", 20), ("sdvrp", 50)])
def test_differentiable_loglik_matches_kernel(env_name, n):
    from rl4co_b200.reinforce import evaluate_log_likelihood

    env, pol, td = _setup(env_name, n, 256)
    with torch.no_grad():
        out = pol(td, env, decode_type="sampling", seed=3, temperature=4.0, return_sum_log_likelihood=False)
    lp = evaluate_log_likelihood(pol, td, env, out["actions"], return_sum=False, temperature=4.0)
    assert torch.isfinite(lp).all()
    torch.testing.assert_close(lp, out["log_likelihood"], rtol=1e-4, atol=5e-5)
    assert lp.requires_grad


@pytest.mark.parametrize("env_name", ["op", "pctsp", "sdvrp"])
def test_reinforce_step_runs(env_name):
    from rl4co_b200.reinforce import get_reinforce_baseline, reinforce_step

    env, pol, td = _setup(env_name, 20, 128)
    opt = torch.optim.Adam(pol.parameters(), lr=1e-4)
    before = pol.decoder.pointer.project_out.weight.detach().clone()
    res = reinforce_step(pol, env, td, get_reinforce_baseline("exponential"),
                         optimizer=opt, seed=1)
    assert torch.isfinite(res["loss"]) and torch.isfinite(res["log_likelihood"]).all()
    assert not torch.equal(before, pol.decoder.pointer.project_out.weight)
