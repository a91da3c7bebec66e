"""GPU: REINFORCE / POMO glue -- the differentiable teacher-forced log-likelihood equals what the
rollout kernel reports, losses equal the oracle's, gradients flow, POMO reduction layout."""

import pytest
import torch

from oracle import am_rollout_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(env_name, n, B, seed=0, layers=1, **kw):
    from rl4co_b200.envs import get_env
    from rl4co_b200.policy import FusedAttentionModelPolicy

    torch.manual_seed(seed)
    env = get_env(env_name, generator_params=dict(num_loc=n), check_solution=True)
    pol = FusedAttentionModelPolicy(env_name=env_name, num_encoder_layers=layers, **kw).to(DEV).eval()
    td = env.reset(env.generator(B).to(DEV))
    return env, pol, td


@pytest.mark.parametrize("env_name,n", [("tsp", 20), ("cvrp", 20), ("tsp", 50), ("cvrp", 50)])
@pytest.mark.parametrize("decode_type", ["sampling", "multistart_sampling"])
def test_differentiable_loglik_matches_kernel(env_name, n, decode_type):
    from rl4co_b200.reinforce import evaluate_log_likelihood

    env, pol, td = _setup(env_name, n, 24)
    kw = {"num_starts": 5} if "multistart" in decode_type else {}
    with torch.no_grad():
        out = pol(td, env, decode_type=decode_type, seed=3, return_sum_log_likelihood=False, **kw)
    lp = evaluate_log_likelihood(pol, td, env, out["actions"], return_sum=False)
    torch.testing.assert_close(lp, out["log_likelihood"], rtol=1e-4, atol=5e-5)
    assert lp.requires_grad


def test_differentiable_loglik_is_finite_at_cvrp100():
    """exact capacity fits (integer demands / 50) must replay as feasible: regression for the
    fp32 prefix-sum replay that produced -inf log-probs."""
    from rl4co_b200.reinforce import evaluate_log_likelihood

    env, pol, td = _setup("cvrp", 100, 2048)
    with torch.no_grad():
        out = pol(td, env, decode_type="sampling", seed=5)
        ll = evaluate_log_likelihood(pol, td, env, out["actions"])
    assert torch.isfinite(ll).all()
    torch.testing.assert_close(ll, out["log_likelihood"], rtol=2e-4, atol=2e-3)


@pytest.mark.parametrize("env_name", ["tsp", "cvrp"])
def test_replay_states_match_oracle_masks(env_name):
    from rl4co_b200.reinforce import replay_states

    env, pol, td = _setup(env_name, 20, 16)
    with torch.no_grad():
        out = pol(td, env, decode_type="sampling", seed=1)
    acts = out["actions"]
    mask, prev, first, used = replay_states(env_name, td, acts)
    inst = {k: td[k].cpu() for k in ("locs", "demand") if k in td.keys()}
    st = O.tsp_reset(inst["locs"]) if env_name == "tsp" else None
    if env_name == "cvrp":
        st = O.cvrp_reset(td["locs"][:, 0].cpu(), td["locs"][:, 1:].cpu(), inst["demand"])
    for t in range(acts.shape[1]):
        assert torch.equal(mask[:, t].cpu(), st["action_mask"]), f"mask step {t}"
        if env_name == "cvrp":
            torch.testing.assert_close(used[:, t].cpu(), st["used_capacity"].reshape(-1), rtol=0, atol=1e-6)
        st = O.ENV_STEP[env_name](st, acts[:, t].cpu())


@pytest.mark.parametrize("env_name,baseline", [("tsp", "mean"), ("cvrp", "rollout"), ("tsp", "exponential"), ("cvrp", "no")])
def test_reinforce_step_trains(env_name, baseline):
    from rl4co_b200.reinforce import get_reinforce_baseline, reinforce_step

    env, pol, td = _setup(env_name, 20, 64, layers=2)
    bl = get_reinforce_baseline(baseline)
    bl.setup(pol)
    opt = torch.optim.Adam(pol.parameters(), lr=1e-3)
    before = [p.detach().clone() for p in pol.parameters()]
    res = reinforce_step(pol, env, td, bl, opt, seed=11)
    assert torch.isfinite(res["loss"]) and res["reward"].shape == (64,)
    changed = sum((a != b.detach()).any().item() for a, b in zip(before, pol.parameters()))
    assert changed > 5
    # loss value vs the oracle formula on the same numbers
    bl_val = res["bl_val"] if isinstance(res["bl_val"], torch.Tensor) else torch.tensor(float(res["bl_val"]), device=DEV)
    ref = O.reinforce_loss(res["reward"].cpu(), res["log_likelihood"].cpu(), bl_val.cpu())
    torch.testing.assert_close(res["reinforce_loss"].cpu(), ref, rtol=1e-5, atol=1e-6)


def test_mean_baseline_equals_reward_mean():
    from rl4co_b200.reinforce import MeanBaseline

    r = -torch.rand(10007, device=DEV) * 20
    v, _ = MeanBaseline().eval(None, r)
    assert abs(v.item() - r.double().mean().item()) < 1e-5


@pytest.mark.parametrize("env_name", ["tsp", "cvrp"])
def test_pomo_step_eval_and_train(env_name):
    from rl4co_b200.reinforce import pomo_step

    env, pol, td = _setup(env_name, 20, 6, use_graph_context=False)
    res = pomo_step(pol, env, td, num_augment=8, phase="test")
    S = env.get_num_starts(td)
    assert res["reward"].shape == (6, 8, S) and res["max_aug_reward"].shape == (6,)
    # oracle layout check: unbatchify(reward, (aug, start))
    flat = res["reward"].permute(2, 1, 0).reshape(-1).cpu()  # index s*(A*B) + a*B + b
    mr, mar = O.pomo_reduce(flat, 8, S)
    torch.testing.assert_close(mar, res["max_aug_reward"].cpu())
    # the identity augmentation (a = 0) reproduces the un-augmented multistart rollout
    plain = pomo_step(pol, env, td, num_augment=0, phase="test")
    torch.testing.assert_close(plain["reward"], res["reward"][:, 0], rtol=1e-5, atol=1e-5)
    opt = torch.optim.Adam(pol.parameters(), lr=1e-4)
    tr = pomo_step(pol, env, td, phase="train", optimizer=opt)
    assert torch.isfinite(tr["loss"])


@pytest.mark.parametrize("env_name", ["tsp", "cvrp"])
def test_policy_forward_under_autograd_is_differentiable(env_name):
    """`policy(td, env, phase="train")` with autograd on (what the reference's REINFORCE.shared_step
    does, reinforce.py:59-69) must return a log_likelihood that backpropagates into the policy."""
    env, pol, td = _setup(env_name, 20, 16)
    pol.train()
    out = pol(td, env, phase="train", decode_type="sampling", seed=2)
    assert out["log_likelihood"].requires_grad
    loss = -((out["reward"] - out["reward"].mean()) * out["log_likelihood"]).mean()
    loss.backward()
    g = pol.decoder.project_node_embeddings.weight.grad
    assert g is not None and torch.isfinite(g).all() and g.abs().sum() > 0
    with torch.no_grad():
        chk = pol(td, env, phase="train", actions=out["actions"])
    torch.testing.assert_close(out["log_likelihood"].detach(), chk["log_likelihood"], rtol=1e-4, atol=2e-4)


def test_low_temperature_takes_the_stepping_path():
    """2*clip/T beyond exp's fp32 range would flush the fused kernel's fixed-offset softmax to 0."""
    env, pol, td = _setup("tsp", 20, 8)
    with torch.no_grad():
        out = pol(td, env, decode_type="sampling", temperature=0.1)
    assert torch.isfinite(out["log_likelihood"]).all()


@pytest.mark.parametrize("env_name", ["tsp", "cvrp"])
def test_reinforce_step_chunked_matches_formula(env_name):
    """`micro_batch`: sampling per chunk, ONE baseline over the whole batch, per-chunk differentiable pass with
    gradient accumulation; the reported loss is the batch-mean REINFORCE loss of the sampled trajectories."""
    from rl4co_b200.reinforce import get_reinforce_baseline, reinforce_step

    env, pol, td = _setup(env_name, 20, 80, layers=1)
    bl = get_reinforce_baseline("mean")
    opt = torch.optim.Adam(pol.parameters(), lr=1e-3)
    before = [p.detach().clone() for p in pol.parameters()]
    res = reinforce_step(pol, env, td, bl, opt, seed=4, micro_batch=32, matmul_precision="highest")
    assert res["chunks"] == 3 and res["reward"].shape == (80,) and res["actions"].shape[0] == 80
    assert abs(float(res["bl_val"]) - res["reward"].mean().item()) < 1e-5  # baseline over the WHOLE batch
    ref = O.reinforce_loss(res["reward"].cpu(), res["log_likelihood"].cpu(), torch.as_tensor(float(res["bl_val"])))
    torch.testing.assert_close(res["reinforce_loss"].cpu(), ref, rtol=1e-4, atol=1e-6)
    changed = sum((a != b.detach()).any().item() for a, b in zip(before, pol.parameters()))
    assert changed > 5
    # phase 1 runs with BatchNorm momentum 0: it counts batches but leaves the running statistics alone, so they
    # move once per chunk per step (phase 2): 2 x chunks forward passes in train mode
    nbt = [m.num_batches_tracked.item() for m in pol.modules() if isinstance(m, torch.nn.BatchNorm1d)]
    assert nbt and all(v == 6 for v in nbt)
