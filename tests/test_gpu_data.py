"""GPU: data path either side of the rollout -- on-device generation, dihedral-8 kernel, the
vectorised stepping kernels at odd row alignments, .npz / dataset / collate round trip."""

import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_generate_uniform_statistics_and_determinism():
    from rl4co_b200 import native

    a = native.generate_uniform((1 << 20, 3), DEV, seed=7, offset=0)
    b = native.generate_uniform((1 << 20, 3), DEV, seed=7, offset=0)
    c = native.generate_uniform((1 << 20, 3), DEV, seed=7, offset=1)
    d = native.generate_uniform((1 << 20, 3), DEV, seed=8, offset=0)
    assert torch.equal(a, b) and not torch.equal(a, c) and not torch.equal(a, d)
    assert a.min() >= 0 and a.max() < 1
    assert abs(a.mean().item() - 0.5) < 2e-3 and abs(a.var().item() - 1 / 12) < 2e-3
    hist = torch.histc(a, bins=64, min=0, max=1)
    exp = a.numel() / 64
    assert ((hist - exp) ** 2 / exp).sum().item() < 130  # chi2, 63 dof: 99.999 % quantile ~ 121
    x = native.generate_uniform((1001,), DEV, seed=3, offset=0, lo=-2.0, hi=5.0)  # n % 4 != 0 tail, range
    assert x.shape == (1001,) and x.min() >= -2 and x.max() < 5 and x[-1] != 0


def test_generate_demand_matches_reference_distribution():
    """cvrp/generator.py:126-137: integers 1..9 over the capacity, uniform."""
    from rl4co_b200 import native

    cap = 50.0
    d = native.generate_demand((4096, 100), DEV, seed=5, offset=1, min_demand=1, max_demand=10, capacity=cap)
    ints = (d * cap).round()
    assert torch.allclose(ints / cap, d) and ints.min() == 1 and ints.max() == 9
    counts = torch.bincount(ints.long().flatten(), minlength=10)[1:10].double()
    exp = d.numel() / 9
    assert ((counts - exp) ** 2 / exp).sum().item() < 40  # chi2, 8 dof


@pytest.mark.parametrize("env_name", ["tsp", "cvrp"])
def test_on_device_generator_feeds_the_policy(env_name):
    from rl4co_b200.envs import get_env
    from rl4co_b200.policy import FusedAttentionModelPolicy

    env = get_env(env_name, generator_params=dict(num_loc=20, device=DEV, seed=11), check_solution=True)
    td0 = env.generator(64)
    assert td0["locs"].is_cuda and td0.device is not None
    td1 = env.generator(64)
    assert not torch.equal(td0["locs"], td1["locs"])  # successive batches differ
    env2 = get_env(env_name, generator_params=dict(num_loc=20, device=DEV, seed=11))
    assert torch.equal(env2.generator(64)["locs"], td0["locs"])  # (seed, call index) reproduces the batch
    pol = FusedAttentionModelPolicy(env_name=env_name, num_encoder_layers=1).to(DEV).eval()
    with torch.inference_mode():
        out = pol(env.reset(td0), env, decode_type="greedy")  # check_solution=True validates the tours
    assert torch.isfinite(out["reward"]).all()


def test_dihedral8_kernel_bit_exact(golden):
    from rl4co_b200 import native
    from rl4co_b200.ops import StateAugmentation, dihedral_8_augmentation
    from rl4co_b200.tensordict import TensorDict

    g = golden("layout")
    x = g["x"].to(DEV)
    assert torch.equal(native.dihedral8(x).cpu(), g["dihedral8"])  # recorded from rl4co/data/transforms.py:16-38
    torch.manual_seed(0)
    big = torch.rand(777, 100, 2, device=DEV)
    assert torch.equal(native.dihedral8(big), dihedral_8_augmentation(big))
    td = StateAugmentation(num_augment=8)(TensorDict({"locs": x}, batch_size=[x.shape[0]]))
    assert torch.equal(td["locs"].cpu(), g["state_aug"])


@pytest.mark.parametrize("N", [5, 48, 50, 51, 100, 101, 127])
@pytest.mark.parametrize("inplace", [False, True])
def test_stepping_kernels_vector_path_matches_bytewise_semantics(N, inplace):
    """co_tsp_step / co_cvrp_step move mask / visited rows in 16-byte words where the row allows it: every
    alignment class of row starts (row * N mod 16) against a torch restatement of tsp/env.py:60-86 and
    cvrp/env.py:66-96."""
    from rl4co_b200 import native

    torch.manual_seed(N)
    B = 67
    mask = torch.rand(B, N, device=DEV) > 0.3
    act = torch.randint(0, N, (B,), device=DEV)
    m_in = mask.clone()
    m_out = m_in if inplace else torch.empty_like(m_in)
    first, cur, i = torch.zeros(B, dtype=torch.int64, device=DEV), torch.empty(B, dtype=torch.int64, device=DEV), \
        torch.ones(B, dtype=torch.int64, device=DEV)
    done = torch.empty(B, dtype=torch.bool, device=DEV)
    native.tsp_step(act, m_in, m_out, first, cur, i, done)
    ref = mask.clone().scatter(-1, act[:, None], False)
    assert torch.equal(m_out, ref) and torch.equal(done, ref.sum(-1) == 0) and torch.equal(cur, act)
    if N >= 2:
        vis = (torch.rand(B, N, device=DEV) > 0.5).to(torch.uint8)
        dem = torch.randint(1, 10, (B, N - 1), device=DEV).float() / 40
        used = torch.rand(B, 1, device=DEV) * 0.5
        cap = torch.ones(B, 1, device=DEV)
        v_in = vis.clone()
        v_out = v_in if inplace else torch.empty_like(v_in)
        used_out, cur2 = torch.empty_like(used), torch.empty(B, 1, dtype=torch.int64, device=DEV)
        done2, m2 = torch.empty(B, dtype=torch.bool, device=DEV), torch.empty(B, N, dtype=torch.bool, device=DEV)
        native.cvrp_step(act, dem, cap, used, used_out, v_in, v_out, cur2, done2, m2)
        v_ref = vis.clone().scatter(-1, act[:, None], 1)
        sel = dem.gather(1, (act - 1).clamp(0, N - 2)[:, None])
        u_ref = (used + sel) * (act != 0).float()[:, None]
        exceeds = dem + u_ref > cap + 1e-5
        mloc = (v_ref[:, 1:] != 0) | exceeds
        mdep = (act == 0)[:, None] & ((~mloc).sum(-1, keepdim=True) > 0)
        assert torch.equal(v_out, v_ref) and torch.equal(used_out, u_ref)
        assert torch.equal(done2, v_ref.sum(-1) == N) and torch.equal(m2, ~torch.cat((mdep, mloc), -1))


def test_npz_dataset_collate_round_trip(tmp_path):
    """data/utils.py:11-34, data/dataset.py:41-130 and CVRPEnv.load_data (cvrp/env.py:179-186)."""
    from torch.utils.data import DataLoader

    from rl4co_b200.data import TensorDictDataset, load_npz_to_tensordict, save_tensordict_to_npz
    from rl4co_b200.envs import get_env

    env = get_env("cvrp", generator_params=dict(num_loc=10))
    torch.manual_seed(0)
    td = env.generator(12)
    path = os.path.join(tmp_path, "cvrp10.npz")
    save_tensordict_to_npz(td, path)
    assert set(np.load(path).files) == {"locs", "depot", "demand", "capacity"}
    back = load_npz_to_tensordict(path)
    for k in td.keys():
        assert torch.equal(back[k], td[k])
    ds = TensorDictDataset(back)
    assert len(ds) == 12 and set(ds[3].keys()) == set(td.keys())
    batch = next(iter(DataLoader(ds, batch_size=5, collate_fn=TensorDictDataset.collate_fn)))
    assert batch.batch_size[0] == 5 and torch.equal(batch["locs"], td["locs"][:5])
    with torch.inference_mode():
        st = env.reset(batch.to(DEV))
    assert st["action_mask"].shape == (5, 11)
