"""On-disk / dataset formats either side of the rollout path.

  load_npz_to_tensordict / save_tensordict_to_npz   rl4co/data/utils.py:11-34
  TensorDictDataset (+ collate_fn)                  rl4co/data/dataset.py:41-130
  generate_tsp_data / generate_vrp_data / generate_dataset / generate_default_datasets
                                                    rl4co/data/generate_data.py:37-76,213-317
    (numpy-seeded validation / test sets: seed 4321 = "val", 1234 = "test"; same call order of the numpy global
     generator, so the files are bit-identical to the reference's for the tsp / vrp problems on this path)
"""

from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import Dataset

from .tensordict import TensorDict


def load_npz_to_tensordict(filename) -> TensorDict:
    x = np.load(filename)
    x_dict = {k: torch.from_numpy(np.asarray(v)) for k, v in dict(x).items()}
    batch_size = x_dict[list(x_dict.keys())[0]].shape[0]
    return TensorDict(x_dict, batch_size=batch_size)


def save_tensordict_to_npz(tensordict: TensorDict, filename, compress: bool = False):
    x_dict = {k: v.cpu().numpy() for k, v in tensordict.items()}
    (np.savez_compressed if compress else np.savez)(filename, **x_dict)


class TensorDictDataset(Dataset):
    """rl4co/data/dataset.py:41-78: list-of-dicts dataset, collated by stacking."""

    def __init__(self, td: TensorDict):
        self.data_len = td.batch_size[0]
        self.data = [{key: value[i] for key, value in td.items()} for i in range(self.data_len)]

    def __len__(self):
        return self.data_len

    def __getitem__(self, idx):
        return self.data[idx]

    @staticmethod
    def collate_fn(batch):
        return TensorDict({key: torch.stack([b[key] for b in batch]) for key in batch[0].keys()},
                          batch_size=torch.Size([len(batch)]))


# rl4co/data/generate_data.py:41-58 (Kool et al. capacities)
VRP_CAPACITIES = {10: 20.0, 15: 25.0, 20: 30.0, 30: 33.0, 40: 37.0, 50: 40.0, 60: 43.0, 75: 45.0,
                  100: 50.0, 125: 55.0, 150: 60.0, 200: 70.0, 500: 100.0, 1000: 150.0}


def generate_tsp_data(dataset_size, tsp_size):
    """generate_data.py:37-38"""
    return {"locs": np.random.uniform(size=(dataset_size, tsp_size, 2)).astype(np.float32)}


def generate_vrp_data(dataset_size, vrp_size, capacities=None):
    """generate_data.py:41-76: depot, locations, integer demands 1..9 (NOT yet divided by the capacity -- CVRPEnv.
    load_data does that, cvrp/env.py:179-186), capacity."""
    caps = dict(VRP_CAPACITIES)
    if capacities is not None:
        caps.update({k: v for k, v in capacities.items() if k in caps})
    return {
        "depot": np.random.uniform(size=(dataset_size, 2)).astype(np.float32),
        "locs": np.random.uniform(size=(dataset_size, vrp_size, 2)).astype(np.float32),
        "demand": np.random.randint(1, 10, size=(dataset_size, vrp_size)).astype(np.float32),
        "capacity": np.full(dataset_size, caps[vrp_size]).astype(np.float32),
    }


_GENERATORS = {"tsp": generate_tsp_data, "vrp": generate_vrp_data, "cvrp": generate_vrp_data, "sdvrp": generate_vrp_data}


def dataset_filename(data_dir, problem, graph_size, name, seed):
    """generate_data.py:268-279: data_dir/problem/problem{size}_{name}_seed{seed}.npz"""
    import os

    return os.path.join(data_dir, problem, f"{problem}{graph_size}_{name}_seed{seed}.npz")


def generate_dataset(filename=None, data_dir="data", name=None, problem="tsp", dataset_size=10000,
                     graph_sizes=(20, 50, 100), overwrite=False, seed=1234):
    """generate_data.py:213-311 for the problems on this path (tsp, vrp): one .npz per graph size, the numpy global
    generator re-seeded before each file.  Returns the list of file names."""
    import os

    graph_sizes = [graph_sizes] if isinstance(graph_sizes, int) else list(graph_sizes)
    filenames = [filename] if isinstance(filename, str) else filename
    problems = ["tsp", "vrp"] if problem == "all" else [problem]
    out, it = [], 0
    for prob in problems:
        if prob not in _GENERATORS:
            raise NotImplementedError(f"Environment type {prob} not implemented")
        for gs in graph_sizes:
            if filenames is None:
                fname = dataset_filename(data_dir, prob, gs, name, seed)
            else:
                fname = filenames[it] if filenames[it].endswith(".npz") else filenames[it] + ".npz"
                it += 1
            os.makedirs(os.path.dirname(fname) or ".", exist_ok=True)
            out.append(fname)
            if not overwrite and os.path.isfile(fname):
                continue
            np.random.seed(seed)
            np.savez(fname, **_GENERATORS[prob](dataset_size, gs))
    return out


def generate_default_datasets(data_dir, dataset_size=10000, graph_sizes=(20, 50, 100)):
    """generate_data.py:314-317: the validation (seed 4321) and test (seed 1234) sets of the paper."""
    return (generate_dataset(data_dir=data_dir, name="val", problem="all", seed=4321, dataset_size=dataset_size,
                             graph_sizes=graph_sizes)
            + generate_dataset(data_dir=data_dir, name="test", problem="all", seed=1234, dataset_size=dataset_size,
                               graph_sizes=graph_sizes))
