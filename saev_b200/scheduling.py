"""Host-side schedule objects of saev's loop (mirror of /root/reference/src/saev/utils/scheduling.py).

`WarmupCosine` produces the learning rate saev assigns to the param group AFTER each optimizer step
(train.py:449-451; the first step therefore runs with lr = 0.0, train.py:118) and `BatchLimiter` re-iterates a
loader until `n_samples` rows were seen (train.py:260).  Pure Python scalars: nothing here touches the GPU.
"""

from __future__ import annotations

import collections.abc
import math
import typing as tp


class Scheduler:
    def step(self) -> float:
        raise NotImplementedError(f"{self.__class__.__name__} must implement step().")

    def __repr__(self) -> str:
        raise NotImplementedError(f"{self.__class__.__name__} must implement __repr__().")


class Warmup(Scheduler):
    """scheduling.py:21-40: linear from `init` to `final` over `n_steps` calls, then `final`."""

    def __init__(self, init: float, final: float, n_steps: int):
        self.init, self.final, self.n_steps = init, final, n_steps
        self._step = 0

    def step(self) -> float:
        self._step += 1
        if self._step < self.n_steps:
            return self.init + (self.final - self.init) * (self._step / self.n_steps)
        return self.final

    def __repr__(self) -> str:
        return f"Warmup(init={self.init}, final={self.final}, n_steps={self.n_steps})"


class WarmupCosine(Scheduler):
    """scheduling.py:43-71: linear `init` -> `peak` over `n_warmup` calls, cosine `peak` -> `final` until
    `n_steps`, then `final`.  The k-th call evaluates the schedule at the 1-based step k."""

    def __init__(self, init: float, n_warmup: int, peak: float, n_steps: int, final: float):
        self.init, self.peak, self.final = init, peak, final
        self.n_warmup, self.n_steps = n_warmup, n_steps
        self._step = 0

    def step(self) -> float:
        self._step += 1
        s = self._step
        if s < self.n_warmup:
            return self.init + (self.peak - self.init) * (s / self.n_warmup)
        if s < self.n_steps:
            progress = (s - self.n_warmup) / (self.n_steps - self.n_warmup)
            return self.final + (self.peak - self.final) * (1 + math.cos(math.pi * progress)) / 2
        return self.final

    def __repr__(self) -> str:
        return (f"WarmupCosine(init={self.init}, peak={self.peak}, final={self.final}, n_warmup={self.n_warmup}, "
                f"n_steps={self.n_steps})")


def _infer_batch_size(batch: tp.Any, fallback: int) -> int:
    """scheduling.py:125-153: rows in a batch without assuming its schema: mapping -> len(first value),
    anything with __len__ -> len(batch), otherwise (or if that is not a positive int) `fallback`."""
    try:
        if isinstance(batch, collections.abc.Mapping):
            if len(batch) == 0:
                return fallback
            n = len(next(iter(batch.values())))
        else:
            n = len(batch)
        if isinstance(n, int) and n > 0:
            return n
    except Exception:
        pass
    return fallback


class BatchLimiter:
    """scheduling.py:83-122: yields batches from `dataloader`, restarting it as often as needed, until at least
    `n_samples` rows were produced.  Unknown attributes are forwarded to the wrapped loader."""

    def __init__(self, dataloader, n_samples: int):
        self.dataloader = dataloader
        self.n_samples = n_samples
        self.batch_size = dataloader.batch_size
        self.drop_last = dataloader.drop_last

    def __len__(self) -> int:
        return math.ceil(self.n_samples / self.batch_size)

    def __getattr__(self, name: str) -> tp.Any:
        try:
            return getattr(self.__dict__["dataloader"], name)
        except (AttributeError, KeyError):
            raise AttributeError(
                f"'{self.__class__.__name__}' object and its wrapped dataloader have no attribute '{name}'"
            ) from None

    def __iter__(self):
        self.n_seen = 0
        while True:
            for batch in self.dataloader:
                yield batch
                self.n_seen += _infer_batch_size(batch, fallback=self.batch_size)
                if self.n_seen >= self.n_samples:
                    return
            # the epoch's short last batch is not counted twice (scheduling.py:119-122)
            if not self.dataloader.drop_last:
                self.n_seen -= self.batch_size
