"""On-disk / dataset formats either side of the rollout path.

  load_npz_to_tensordict / save_tensordict_to_npz   rl4co/data/utils.py:11-34
  TensorDictDataset (+ collate_fn)                  rl4co/data/dataset.py:41-130
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
