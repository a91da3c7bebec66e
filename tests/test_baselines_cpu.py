"""CPU: REINFORCE baselines that are host-side glue against the reference's own classes
(rl4co/models/rl/reinforce/baselines.py, through the container stubs): Warmup blending, the rollout baseline's
epoch-end paired t-test swap, the critic baseline's value / loss, registry names."""

import copy
import importlib

import pytest
import torch
import torch.nn as nn

from oracle import ref_standin
from rl4co_b200 import reinforce as R
from rl4co_b200.data import TensorDictDataset
from rl4co_b200.tensordict import TensorDict

pytestmark = pytest.mark.skipif(not ref_standin.reference_available(), reason="no reference tree (oracle/_ref not staged)")


@pytest.fixture(scope="module")
def refbl():
    ref_standin.install()
    return importlib.import_module("rl4co.models.rl.reinforce.baselines")


class FakePolicy(nn.Module):
    """reward = -(sum x) + bias + wobble * sin(40 x0): a policy whose quality is one number, so that candidate /
    baseline comparisons are controlled."""

    def __init__(self, bias=0.0, wobble=0.0):
        super().__init__()
        self.bias = nn.Parameter(torch.tensor(float(bias)))
        self.wobble = wobble

    def forward(self, td, env=None, phase=None, decode_type=None, **kw):
        x = td["x"]
        return {"reward": -(x.sum(-1)) + self.bias + self.wobble * torch.sin(40 * x[:, 0])}


class FakeEnv:
    name = "fake"

    def reset(self, td):
        return td

    def dataset(self, batch_size=[], **kw):
        g = torch.Generator().manual_seed(123)
        return TensorDictDataset(TensorDict({"x": torch.rand(batch_size[0], 5, generator=g)}, batch_size=batch_size))


class RefDataset(torch.utils.data.Dataset):
    """the reference's DataLoader path needs a dataset with `collate_fn` (data/dataset.py:41-78)."""

    def __init__(self, ds):
        self.ds = ds
        self.collate_fn = ds.collate_fn

    def __len__(self):
        return len(self.ds)

    def __getitem__(self, i):
        return self.ds[i]


class RefEnv(FakeEnv):
    def dataset(self, batch_size=[], **kw):
        return RefDataset(super().dataset(batch_size))


@pytest.mark.parametrize("cand,expect_update", [(dict(bias=0.5, wobble=0.0), True),       # clearly better
                                                 (dict(bias=0.002, wobble=0.3), False),    # better on average, not significant
                                                 (dict(bias=-0.3, wobble=0.0), False)])    # worse
def test_rollout_baseline_ttest_swap_matches_reference(refbl, cand, expect_update):
    base = FakePolicy(0.0)
    mine, theirs = R.RolloutBaseline(bl_alpha=0.05), refbl.RolloutBaseline(bl_alpha=0.05)
    mine.setup(base, FakeEnv(), batch_size=16, device="cpu", dataset_size=64)
    theirs.setup(base, RefEnv(), batch_size=16, device="cpu", dataset_size=64)
    assert abs(float(mine.mean) - float(theirs.mean)) < 1e-6
    torch.testing.assert_close(torch.as_tensor(mine.bl_vals), torch.as_tensor(theirs.bl_vals))
    candidate = FakePolicy(**cand)
    updated = mine.epoch_callback(candidate, FakeEnv(), batch_size=16, device="cpu", epoch=0, dataset_size=64)
    theirs.epoch_callback(copy.deepcopy(candidate), RefEnv(), batch_size=16, device="cpu", epoch=0, dataset_size=64)
    assert updated == expect_update
    assert abs(float(mine.mean) - float(theirs.mean)) < 1e-6          # both swapped, or both kept the old policy
    assert float(mine.policy.bias) == float(theirs.policy.bias)
    td = TensorDict({"x": torch.rand(7, 5)}, batch_size=[7])
    torch.testing.assert_close(mine.eval(td, None, FakeEnv())[0], theirs.eval(td, None, RefEnv())[0])


def test_warmup_baseline_blends_like_reference(refbl, monkeypatch):
    # the exponential part's batch mean is `co_reward_stats` + all-reduce in the product (CUDA only, no CPU fallback);
    # this CPU test exercises the blending / epoch logic around it, so the mean is served by torch here
    monkeypatch.setattr(R, "global_mean_baseline", lambda r: r.double().mean().float())
    torch.manual_seed(0)
    mine = R.get_reinforce_baseline("rollout", n_epochs=2, exp_beta=0.8)
    theirs = refbl.get_reinforce_baseline("rollout", n_epochs=2, exp_beta=0.8)
    assert isinstance(mine, R.WarmupBaseline) and isinstance(mine.baseline, R.RolloutBaseline)
    base = FakePolicy(0.1)
    mine.setup(base, FakeEnv(), batch_size=16, device="cpu", dataset_size=32)
    theirs.setup(base, RefEnv(), batch_size=16, device="cpu", dataset_size=32)
    for epoch in range(3):
        for _ in range(2):  # two batches per epoch: the exponential part has memory
            td = TensorDict({"x": torch.rand(9, 5)}, batch_size=[9])
            reward = -torch.rand(9) * 5
            v_m, l_m = mine.eval(td, reward, FakeEnv())
            v_t, l_t = theirs.eval(td, reward, RefEnv())
            torch.testing.assert_close(torch.as_tensor(v_m).float().expand(9), torch.as_tensor(v_t).float().expand(9))
            assert float(l_m) == float(l_t) == 0.0
        mine.epoch_callback(base, FakeEnv(), batch_size=16, device="cpu", epoch=epoch, dataset_size=32)
        theirs.epoch_callback(base, RefEnv(), batch_size=16, device="cpu", epoch=epoch, dataset_size=32)
        assert mine.alpha == theirs.alpha
    assert mine.alpha == 1


def test_critic_baseline_matches_reference(refbl):
    class Enc(nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = nn.Linear(2, 128)

        def forward(self, td):
            h = self.lin(td["locs"])
            return h, h

    class Pol(nn.Module):
        def __init__(self):
            super().__init__()
            self.encoder = Enc()

    torch.manual_seed(1)
    pol = Pol()
    mine, theirs = R.CriticBaseline(), refbl.CriticBaseline()
    torch.manual_seed(2)
    mine.setup(pol, FakeEnv())
    torch.manual_seed(2)
    theirs.setup(pol, FakeEnv())
    assert [k for k, _ in mine.critic.state_dict().items()] == [k for k, _ in theirs.critic.state_dict().items()]
    td = TensorDict({"locs": torch.rand(6, 10, 2)}, batch_size=[6])
    reward = -torch.rand(6) * 4
    v_m, l_m = mine.eval(td, reward)
    v_t, l_t = theirs.eval(td, reward)
    torch.testing.assert_close(v_m, v_t)
    torch.testing.assert_close(l_m, l_t)
    assert l_m.requires_grad and not v_m.requires_grad


def test_registry_names(refbl):
    for name in ("no", "shared", "exponential", "mean", "critic", "rollout_only", "rollout", "warmup"):
        mine = R.get_reinforce_baseline(name)
        theirs = refbl.get_reinforce_baseline(name)
        assert type(mine).__name__ == type(theirs).__name__ or name == "mean"  # reference's Mean = Exponential(beta=0)
