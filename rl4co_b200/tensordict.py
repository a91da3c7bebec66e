"""Minimal batched tensor container with the slice of the `tensordict.TensorDict`
surface that rl4co's rollout path touches.

The real `tensordict` package is not installable on the build / GPU boxes (no
network), and rl4co's env / decoder API passes state around as a TensorDict
(reference: rl4co/envs/common/base.py:121-143, rl4co/utils/ops.py:10-51,
rl4co/models/zoo/am/decoder.py:178-179).  If the real package is importable we
re-export it so the drop-ins interoperate with a genuine rl4co install;
otherwise this duck-typed stand-in is used.

Only *batch* dimensions are affected by shape ops (expand / view / permute /
indexing), exactly like the real TensorDict: a tensor stored under a key has
shape ``batch_size + feature_shape``.
"""

from __future__ import annotations

from typing import Any, Iterable

import torch

try:  # pragma: no cover - real package absent in this image
    from tensordict import TensorDict as _RealTensorDict  # type: ignore

    _HAVE_REAL = True
except Exception:  # ModuleNotFoundError or a stub registered by the oracle
    _RealTensorDict = None
    _HAVE_REAL = False

__version__ = "0.6.0-b200-standin"


def _as_size(batch_size) -> torch.Size:
    if batch_size is None:
        return torch.Size([])
    if isinstance(batch_size, int):
        return torch.Size([batch_size])
    return torch.Size(list(batch_size))


class TensorDict:
    """dict[str, Tensor | TensorDict] with common leading batch dims."""

    def __init__(self, source: dict | None = None, batch_size=None, device=None, **_unused):
        self._batch_size = _as_size(batch_size)
        self._device = torch.device(device) if device is not None else None
        self._data: dict[str, Any] = {}
        if source is not None:
            for k, v in source.items():
                self.set(k, v)

    # ------------------------------------------------------------------ meta
    @property
    def batch_size(self) -> torch.Size:
        return self._batch_size

    @batch_size.setter
    def batch_size(self, value):
        self._batch_size = _as_size(value)

    @property
    def shape(self) -> torch.Size:
        return self._batch_size

    def size(self, dim: int | None = None):
        return self._batch_size if dim is None else self._batch_size[dim]

    def dim(self) -> int:
        return len(self._batch_size)

    ndim = property(dim)

    def numel(self) -> int:
        n = 1
        for s in self._batch_size:
            n *= s
        return n

    @property
    def device(self):
        if self._device is not None:
            return self._device
        devs = {v.device for v in self._data.values() if isinstance(v, (torch.Tensor, TensorDict)) and v.device is not None}
        if len(devs) == 1:
            return next(iter(devs))
        return None

    def is_empty(self) -> bool:
        return len(self._data) == 0

    # ------------------------------------------------------------ dict-like
    def keys(self, *_a, **_k):
        return self._data.keys()

    def values(self):
        return self._data.values()

    def items(self):
        return self._data.items()

    def __contains__(self, key) -> bool:
        return key in self._data

    def __iter__(self):
        return iter(self._data)

    def __len__(self) -> int:
        return self._batch_size[0] if len(self._batch_size) else 0

    def get(self, key: str, default: Any = ...):
        if key in self._data:
            return self._data[key]
        if default is ...:
            raise KeyError(f"key {key!r} not found in TensorDict with keys {list(self._data)}")
        return default

    def set(self, key: str, value, inplace: bool = False):
        if isinstance(value, dict):
            value = TensorDict(value, batch_size=self._batch_size)
        elif not isinstance(value, (torch.Tensor, TensorDict)):
            value = torch.as_tensor(value)
        self._data[key] = value
        return self

    def pop(self, key: str, default: Any = ...):
        if default is ...:
            return self._data.pop(key)
        return self._data.pop(key, default)

    def update(self, other: "dict | TensorDict", **_k):
        src = other.items() if isinstance(other, (dict, TensorDict)) else other
        for k, v in src:
            self.set(k, v)
        return self

    def select(self, *keys: str, strict: bool = True):
        return TensorDict(
            {k: self._data[k] for k in keys if strict or k in self._data},
            batch_size=self._batch_size,
            device=self._device,
        )

    def exclude(self, *keys: str):
        flat = set()
        for k in keys:
            flat.add(k[0] if isinstance(k, tuple) else k)
        return TensorDict(
            {k: v for k, v in self._data.items() if k not in flat},
            batch_size=self._batch_size,
            device=self._device,
        )

    def to_dict(self) -> dict:
        return {k: (v.to_dict() if isinstance(v, TensorDict) else v) for k, v in self._data.items()}

    # -------------------------------------------------------------- indexing
    def __getitem__(self, idx):
        if isinstance(idx, str):
            return self.get(idx)
        if isinstance(idx, tuple) and len(idx) and all(isinstance(i, str) for i in idx):
            out = self
            for i in idx:
                out = out.get(i)
            return out
        return self._map(lambda t: t[idx], new_batch=self._index_shape(idx))

    def __setitem__(self, idx, value):
        if isinstance(idx, str):
            self.set(idx, value)
            return
        # batch-index assignment from another TensorDict / dict
        src = value.items() if isinstance(value, (dict, TensorDict)) else None
        if src is None:
            raise TypeError("can only assign a TensorDict/dict to a batch index")
        for k, v in src:
            self._data[k][idx] = v

    def _index_shape(self, idx) -> torch.Size:
        probe = torch.empty(self._batch_size, dtype=torch.bool, device="meta")
        return probe[idx].shape

    def _map(self, fn, new_batch=None) -> "TensorDict":
        out = TensorDict({}, batch_size=self._batch_size if new_batch is None else new_batch, device=self._device)
        for k, v in self._data.items():
            out._data[k] = fn(v)
        return out

    def _feat(self, t) -> tuple:
        return tuple(t.shape[len(self._batch_size):])

    # ------------------------------------------------------------- tensor ops
    def clone(self, recurse: bool = True) -> "TensorDict":
        return self._map(lambda t: t.clone() if recurse else t)

    def copy(self) -> "TensorDict":
        return self.clone(recurse=False)

    def contiguous(self) -> "TensorDict":
        return self._map(lambda t: t.contiguous())

    def detach(self) -> "TensorDict":
        return self._map(lambda t: t.detach())

    def to(self, *args, **kwargs) -> "TensorDict":
        out = self._map(lambda t: t.to(*args, **kwargs))
        dev = kwargs.get("device", None)
        if dev is None:
            for a in args:
                if isinstance(a, (str, torch.device)):
                    dev = a
        out._device = torch.device(dev) if dev is not None else self._device
        return out

    def cpu(self):
        return self.to("cpu")

    def cuda(self, device=None):
        return self.to("cuda" if device is None else device)

    def pin_memory(self):
        return self._map(lambda t: t.pin_memory())

    def expand(self, *shape) -> "TensorDict":
        if len(shape) == 1 and not isinstance(shape[0], int):
            shape = tuple(shape[0])
        shape = tuple(shape)

        def _ex(t):
            if isinstance(t, TensorDict):
                return t.expand(*shape)
            return t.expand(*shape, *self._feat(t))

        return self._map(_ex, new_batch=torch.Size(shape))

    def view(self, *shape) -> "TensorDict":
        if len(shape) == 1 and not isinstance(shape[0], int):
            shape = tuple(shape[0])
        shape = tuple(shape)
        new_batch = torch.empty(self._batch_size, device="meta").view(*shape).shape

        def _vw(t):
            if isinstance(t, TensorDict):
                return t.view(*shape)
            return t.view(*new_batch, *self._feat(t))

        return self._map(_vw, new_batch=new_batch)

    def reshape(self, *shape) -> "TensorDict":
        if len(shape) == 1 and not isinstance(shape[0], int):
            shape = tuple(shape[0])
        shape = tuple(shape)
        new_batch = torch.empty(self._batch_size, device="meta").reshape(*shape).shape

        def _rs(t):
            if isinstance(t, TensorDict):
                return t.reshape(*shape)
            return t.reshape(*new_batch, *self._feat(t))

        return self._map(_rs, new_batch=new_batch)

    def permute(self, *dims) -> "TensorDict":
        if len(dims) == 1 and not isinstance(dims[0], int):
            dims = tuple(dims[0])
        dims = tuple(dims)
        nb = len(self._batch_size)
        assert sorted(dims) == list(range(nb)), "TensorDict.permute acts on batch dims only"
        new_batch = torch.Size([self._batch_size[d] for d in dims])

        def _pm(t):
            if isinstance(t, TensorDict):
                return t.permute(*dims)
            extra = tuple(range(nb, t.dim()))
            return t.permute(*dims, *extra)

        return self._map(_pm, new_batch=new_batch)

    def gather(self, dim: int, index: torch.Tensor) -> "TensorDict":
        nb = len(self._batch_size)
        if dim < 0:
            dim += nb

        def _ga(t):
            if isinstance(t, TensorDict):
                return t.gather(dim, index)
            feat = self._feat(t)
            idx = index.view(*index.shape, *([1] * len(feat))).expand(*index.shape, *feat)
            return t.gather(dim, idx)

        return self._map(_ga, new_batch=index.shape)

    def squeeze(self, dim: int) -> "TensorDict":
        nb = len(self._batch_size)
        if dim < 0:
            dim += nb
        new_batch = list(self._batch_size)
        assert new_batch[dim] == 1
        new_batch.pop(dim)
        return self._map(lambda t: t.squeeze(dim), new_batch=torch.Size(new_batch))

    def unsqueeze(self, dim: int) -> "TensorDict":
        nb = len(self._batch_size)
        if dim < 0:
            dim += nb + 1
        new_batch = list(self._batch_size)
        new_batch.insert(dim, 1)
        return self._map(lambda t: t.unsqueeze(dim), new_batch=torch.Size(new_batch))

    def __repr__(self) -> str:
        fields = ", ".join(
            f"{k}: {tuple(v.shape)} {str(getattr(v, 'dtype', 'td')).replace('torch.', '')}"
            for k, v in self._data.items()
        )
        return f"TensorDict({{{fields}}}, batch_size={list(self._batch_size)}, device={self.device})"


def cat(tds: Iterable[TensorDict], dim: int = 0) -> TensorDict:
    """torch.cat over the batch dimension for TensorDicts (collate helper)."""
    tds = list(tds)
    first = tds[0]
    out = {k: torch.cat([t[k] for t in tds], dim=dim) for k in first.keys()}
    new_batch = list(first.batch_size)
    new_batch[dim] = sum(t.batch_size[dim] for t in tds)
    return TensorDict(out, batch_size=new_batch)


if _HAVE_REAL:  # pragma: no cover
    TensorDict = _RealTensorDict  # type: ignore  # noqa: F811
