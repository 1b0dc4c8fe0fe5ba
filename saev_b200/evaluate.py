"""`evaluate()` — mirror of saev.framework.train.evaluate (/root/reference/src/saev/framework/train.py:510-618) on the
sparse forward state.

The reference reads the dense `fwd.f_x[B, d_sae]` of every validation batch twice (`f_x > 0` and `f_x`, column sums,
each moved to the host) and pulls three loss scalars with `.item()`.  Here every batch is one eval-mode forward plus
`saev_b200_eval_accumulate` (atomic per-atom adds from the [B, K] top-k lists, fp64 batch sums), all on the device
with no host sync inside the loop; the formulas that turn the accumulators into `EvalMetrics` are the reference's.

`saev_b200.install()` rebinds `saev.framework.train.evaluate` to this function (it is looked up as a module global at
call time, train.py:198) and hands it saev's own `EvalMetrics` class, so `worker_fn` logs and checkpoints as before.
"""

from __future__ import annotations

import dataclasses

import torch
from torch import Tensor

from . import data as _data
from . import scheduling as _sched

ALMOST_DEAD_LIM = 1e-7  # train.py:530-531
DENSE_LIM = 1e-2


@dataclasses.dataclass(frozen=True)
class EvalMetrics:
    """train.py:466-508 (the W&B table conversion is saev's business; in drop-in mode saev's own class is used)."""

    l0: float
    l1: float
    mse: float
    normalized_mse: float
    sse_sae: float
    sse_baseline: float
    n_dead: int
    n_almost_dead: int
    n_dense: int
    freqs: Tensor
    mean_values: Tensor
    almost_dead_threshold: float
    dense_threshold: float


def finish_metrics(state: dict, metrics_cls=EvalMetrics):
    """train.py:568-616 applied to the device accumulators of one SAE (`Engine.new_eval_state`)."""
    acc = state["acc"].cpu()
    n_tokens = int(acc[7].item())
    assert n_tokens > 0, "Validation dataloader yielded zero tokens; cannot compute normalized MSE."
    sum_vec = acc[8:]
    sse_baseline = (acc[0] - torch.dot(sum_vec, sum_vec) / n_tokens).item()
    assert sse_baseline > 0, f"Validation baseline variance non-positive: sse_baseline={sse_baseline:.6e}"
    n_fired, values = state["n_fired"].cpu(), state["values"].cpu()
    mean_values = values / n_fired
    freqs = n_fired / n_tokens
    sse_sae = acc[1].item()
    return metrics_cls(
        l0=(acc[4] / n_tokens).item(), l1=(acc[5] / n_tokens).item(), mse=(acc[6] / n_tokens).item(),
        normalized_mse=sse_sae / sse_baseline, sse_sae=sse_sae, sse_baseline=sse_baseline,
        n_dead=int((freqs == 0).sum().item()), n_almost_dead=int((freqs < ALMOST_DEAD_LIM).sum().item()),
        n_dense=int((freqs > DENSE_LIM).sum().item()), freqs=freqs, mean_values=mean_values,
        almost_dead_threshold=ALMOST_DEAD_LIM, dense_threshold=DENSE_LIM,
    )


@torch.no_grad()
def evaluate_batches(batches, saes, objectives, metrics_cls=EvalMetrics) -> list:
    """The loop of train.py:546-566 over an iterable of `{"act": Tensor[B, D]}` batches."""
    for m in (*saes, *objectives):
        m.eval()
    states: list[dict | None] = [None] * len(saes)
    dev = saes[0].W_dec.device
    for batch in batches:
        x = batch["act"].to(dev, non_blocking=True).contiguous()  # train.py:547
        for i, (sae, objective) in enumerate(zip(saes, objectives)):
            objective(sae, x)
            if states[i] is None:
                states[i] = sae.engine.new_eval_state()
            sae.engine.eval_accumulate(x, states[i])
    if any(s is None for s in states):
        raise AssertionError("Validation dataloader yielded zero tokens; cannot compute normalized MSE.")
    return [finish_metrics(s, metrics_cls) for s in states]


@torch.no_grad()
def evaluate(cfgs, saes, objectives, *, metrics_cls=EvalMetrics, loader_cls=None) -> list:
    """Same signature and result as saev.framework.train.evaluate (train.py:510-618)."""
    torch.cuda.empty_cache()
    if len({repr(c.val_data) for c in cfgs}) != 1:
        raise ValueError(f"Configs are not parallelizeable: {cfgs}.")
    cfg = cfgs[0]
    dataloader = (loader_cls or _data.ShuffledDataLoader)(cfg.val_data)
    n_val = min(dataloader.n_samples, cfg.n_val)
    limited = _sched.BatchLimiter(dataloader, n_val)
    try:
        return evaluate_batches(limited, saes, objectives, metrics_cls)
    finally:
        if hasattr(dataloader, "shutdown"):
            dataloader.shutdown()
