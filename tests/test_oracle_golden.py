"""CPU: pin oracle/am_rollout_oracle.py against the golden vectors recorded from the
unmodified reference (tests/golden/make_golden.py)."""

import pytest
import torch

from oracle import am_rollout_oracle as O
from conftest import env_of

ENV_FIX = ["env_tsp20", "env_tsp50", "env_cvrp20", "env_cvrp50", "env_sdvrp20", "env_sdvrp50", "env_op20", "env_op50", "env_pctsp20", "env_pctsp50"]
AM_FIX = ["am_tsp20", "am_cvrp20", "am_tsp50", "am_cvrp50", "am_tsp100", "am_cvrp100"]
SD_FIX = ["am_sdvrp20", "am_sdvrp50", "am_op20", "am_op50", "am_pctsp20", "am_pctsp50"]  # sibling envs (sdvrp: dynamic embedding; op), single start


@pytest.mark.parametrize("name", ENV_FIX)
def test_env_mdp_bit_exact(golden, name):
    g = golden(name)
    env = env_of(name)
    st = O.env_reset(env, g.inst())
    actions = g["actions"]
    assert torch.equal(st["action_mask"], g["action_mask"][0])
    for t in range(actions.shape[1]):
        st = O.ENV_STEP[env](st, actions[:, t])
        assert torch.equal(st["action_mask"], g["action_mask"][t + 1]), f"mask step {t}"
        assert torch.equal(st["done"], g["done"][t]), f"done step {t}"
        assert torch.equal(st["current_node"].reshape(-1), g["current_node"][t])
        if env == "cvrp":
            assert torch.equal(st["visited"], g["visited"][t])
            assert torch.equal(st["used_capacity"], g["used_capacity"][t])
        if env == "sdvrp":
            assert torch.equal(st["demand_with_depot"], g["demand_with_depot"][t])
            assert torch.equal(st["used_capacity"], g["used_capacity"][t])
        if env == "op":
            assert torch.equal(st["visited"], g["visited"][t].bool())
            assert torch.equal(st["tour_length"], g["tour_length"][t])
            assert torch.equal(st["current_total_prize"], g["current_total_prize"][t])
        if env == "pctsp":
            assert torch.equal(st["visited"], g["visited"][t].bool())
            assert torch.equal(st["cur_total_prize"], g["cur_total_prize"][t])
            assert torch.equal(st["cur_total_penalty"], g["cur_total_penalty"][t])
    if env == "tsp":
        assert torch.equal(st["first_node"], g["first_node"])
        assert torch.equal(st["i"], g["i"])
        O.tsp_check_solution(actions)
    elif env == "sdvrp":
        O.sdvrp_check_solution(st, actions)
    elif env == "op":
        O.op_check_solution(st, actions)
    elif env == "pctsp":
        O.pctsp_check_solution(st, actions)
    else:
        O.cvrp_check_solution(st, actions)
    r = O.env_reward(env, st, actions)
    assert torch.equal(r, g["reward"])


@pytest.mark.parametrize("name", AM_FIX + SD_FIX)
def test_am_greedy(golden, name):
    g = golden(name)
    out = O.rollout(g.weights(), env_of(name), g.inst(), g["h"], "greedy", return_trace=True)
    assert torch.equal(out["actions"], g["greedy_actions"])
    torch.testing.assert_close(torch.stack(out["trace"]["logits"]), g["greedy_logits"], rtol=1e-6, atol=1e-6)
    assert torch.equal(torch.stack(out["trace"]["mask"]), g["greedy_masks"])
    torch.testing.assert_close(out["logprobs"], g["greedy_logprobs"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out["reward"], g["greedy_reward"], rtol=1e-6, atol=0)


@pytest.mark.parametrize("name", AM_FIX + SD_FIX)
def test_am_sampling_with_recorded_noise(golden, name):
    g = golden(name)
    q = g["sampling_noise"]
    out = O.rollout(g.weights(), env_of(name), g.inst(), g["h"], "sampling", noise=lambda t, shape: q[t])
    assert torch.equal(out["actions"], g["sampling_actions"])
    torch.testing.assert_close(out["logprobs"], g["sampling_logprobs"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out["reward"], g["sampling_reward"], rtol=1e-6, atol=0)


@pytest.mark.parametrize("name", AM_FIX + SD_FIX)
def test_am_teacher_forced(golden, name):
    g = golden(name)
    out = O.rollout(g.weights(), env_of(name), g.inst(), g["h"], actions=g["eval_actions"], return_trace=True)
    torch.testing.assert_close(torch.stack(out["trace"]["logits"]), g["eval_logits"], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(out["logprobs"], g["eval_logprobs"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out["reward"], g["eval_reward"], rtol=1e-6, atol=0)


@pytest.mark.parametrize("name", ["tsp20", "cvrp20"])
@pytest.mark.parametrize("tag,kw", [("beam3_best", dict(beam_width=3, select_best=True)),
                                    ("beam4_all", dict(beam_width=4, select_best=False)),
                                    ("beamN_best", dict(select_best=True))])
def test_am_beam_search(golden, name, tag, kw):
    """utils/decoding.py:464-600 as recorded from the reference (tests/golden/dec_*.npz; same policy and instances
    as am_*.npz -- the stored encoder output guards that)."""
    g, d = golden("am_" + name), golden("dec_" + name)
    assert torch.equal(g["h"], d["h"])
    out = O.rollout_beam_search(g.weights(), env_of(name), g.inst(), g["h"], **kw)
    assert torch.equal(out["actions"], d[tag + "_actions"])
    torch.testing.assert_close(out["logprobs"], d[tag + "_logprobs"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out["reward"], d[tag + "_reward"], rtol=1e-6, atol=0)


@pytest.mark.parametrize("name", ["tsp20", "cvrp20"])
@pytest.mark.parametrize("tag,kw", [("topk4", dict(top_k=4)), ("topp80", dict(top_p=0.8)),
                                    ("topk6_topp90", dict(top_k=6, top_p=0.9))])
def test_am_top_k_top_p_sampling_with_recorded_noise(golden, name, tag, kw):
    """utils/decoding.py:109-188 with the filters on, under the recorded-noise protocol."""
    g, d = golden("am_" + name), golden("dec_" + name)
    q = d[tag + "_noise"]
    out = O.rollout(g.weights(), env_of(name), g.inst(), g["h"], "sampling", noise=lambda t, shape: q[t], **kw)
    assert torch.equal(out["actions"], d[tag + "_actions"])
    torch.testing.assert_close(out["logprobs"], d[tag + "_logprobs"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out["reward"], d[tag + "_reward"], rtol=1e-6, atol=0)


@pytest.mark.parametrize("name", AM_FIX)
def test_am_multistart_and_pomo(golden, name):
    g = golden(name)
    mb = int(g["ms_batch"])
    inst = {k: v[:mb] for k, v in g.inst().items()}
    out = O.rollout(g.weights(), env_of(name), inst, g["h"][:mb], "multistart_greedy")
    assert torch.equal(out["actions"], g["ms_actions"])
    torch.testing.assert_close(out["logprobs"], g["ms_logprobs"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out["reward"], g["ms_reward"], rtol=1e-6, atol=0)
    out = O.rollout(g.weights(), env_of(name), inst, g["h"][:mb], "multistart_greedy", use_graph_context=False)
    assert torch.equal(out["actions"], g["pomo_actions"])
    torch.testing.assert_close(out["logprobs"], g["pomo_logprobs"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name,norm", [("enc_tsp20_batch", "batch"), ("enc_cvrp20_instance", "instance")])
def test_encoder(golden, name, norm):
    g = golden(name)
    env = env_of(name)
    st = O.env_reset(env, g.inst())
    h, init_h = O.encoder_forward(g.weights(), env, st, num_layers=1, normalization=norm)
    torch.testing.assert_close(init_h, g["init_h"], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(h, g["h"], rtol=1e-5, atol=1e-5)


def test_layout_conventions(golden):
    g = golden("layout")
    x = g["x"]
    assert torch.equal(O.dihedral_8_augmentation(x), g["dihedral8"])
    assert torch.equal(O.dihedral_8_augmentation(O.batchify(x, 8)[: x.shape[0]]), g["state_aug"])
    assert torch.equal(O.batchify(x, 4), g["batchify4"])
    r = torch.arange(3 * 8 * 4, dtype=torch.float32)
    assert torch.equal(O.unbatchify_multi(r, (8, 4)), g["unbatchify_8_4"])
    assert torch.equal(O.select_start_nodes(3, 5, 5, "tsp"), g["tsp_starts"])
    assert torch.equal(O.select_start_nodes(3, 5, 5, "cvrp"), g["cvrp_starts"])
    assert O.get_num_starts(5, "tsp") == int(g["tsp_num_starts"])
    assert O.get_num_starts(6, "cvrp") == int(g["cvrp_num_starts"])


def test_pomo_reduce_and_loss():
    torch.manual_seed(0)
    B, A, S = 3, 8, 5
    r = torch.randn(A * S * B)
    mr, mar = O.pomo_reduce(r, A, S)
    rr = r.view(S, A, B).permute(2, 1, 0)  # flat index = s*(A*B) + a*B + b
    assert torch.equal(mr, rr.max(-1)[0]) and torch.equal(mar, rr.max(-1)[0].max(-1)[0])
    ll = torch.randn(S * B)
    rew = torch.randn(S * B)
    bl = O.shared_baseline(rew, S)
    assert bl.shape == (B, 1)


def test_pomo_config_c4_fixture(golden):
    """BASELINE config C4 at its own scale (TSP-100, 6-layer instance-norm encoder, no graph context, dihedral-8,
    100 starts): the oracle port reproduces the unmodified reference's POMO forward -- weights are regenerated from
    parameter names (conftest.name_seeded_weights), so this also pins the state_dict naming."""
    from conftest import name_seeded_weights
    from rl4co_b200.policy import FusedAttentionModelPolicy

    g = golden("pomo_tsp100")
    n_aug, seed = int(g["num_augment"]), int(g["weight_seed"])
    pol = FusedAttentionModelPolicy(env_name="tsp", num_encoder_layers=6, normalization="instance",
                                    use_graph_context=False)
    W = name_seeded_weights(pol.state_dict(), seed)
    locs = g["inst::locs"]
    B, N = locs.shape[:2]
    aug = O.dihedral_8_augmentation(locs)  # StateAugmentation: batchify, keep the first B rows, x8 (transforms.py:40-48)
    assert torch.equal(aug, g["aug_locs"])
    with torch.inference_mode():
        st = O.env_reset("tsp", {"locs": aug})
        h, _ = O.encoder_forward(W, "tsp", st, num_layers=6, normalization="instance")
        torch.testing.assert_close(h[:2], g["h_first_rows"], rtol=1e-5, atol=1e-5)
        ref_actions = g["actions"].long()
        free = O.rollout(W, "tsp", {"locs": aug}, h, "multistart_greedy", use_graph_context=False, num_starts=N,
                         faithful_copies=False)
        same = (free["actions"] == ref_actions).all(1)
        assert same.float().mean() > 0.97, "free-running POMO trajectories differ beyond near-tie noise"
        # teacher-forced on the reference's actions (expanded batch; the forced first step carries log-prob 0)
        out = O.rollout(W, "tsp", {"locs": O.batchify(aug, N)}, O.batchify(h, N), actions=ref_actions,
                        use_graph_context=False, faithful_copies=False)
    torch.testing.assert_close(out["logprobs"][:, 1:].sum(1), g["logprobs_sum"], rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(out["logprobs"][: g["logprobs_rows"].shape[0], 1:], g["logprobs_rows"][:, 1:],
                               rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(out["reward"], g["reward"], rtol=1e-6, atol=0)
    max_r, max_aug = O.pomo_reduce(out["reward"], n_aug, N)[:2]
    torch.testing.assert_close(max_r, g["max_reward"])
    torch.testing.assert_close(max_aug, g["max_aug_reward"])


@pytest.mark.parametrize("name", ["am_op20", "am_op50", "am_pctsp20", "am_pctsp50"])
def test_op_encoder_from_instance(golden, name):
    """OPInitEmbedding (init.py:254-280: depot / (x, y, prize)) and PCTSPInitEmbedding (init.py:221-251: (x, y, expected
    prize, penalty)) + one encoder layer reproduce the recorded embeddings."""
    g = golden(name)
    env = env_of(name)
    st = O.env_reset(env, g.inst())
    h, _ = O.encoder_forward(g.weights(), env, st, num_layers=1, normalization="batch")
    torch.testing.assert_close(h, g["h"], rtol=1e-5, atol=1e-5)
