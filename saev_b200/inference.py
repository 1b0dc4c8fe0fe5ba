"""Forward-only pass over ordered activations -- mirror of saev.framework.inference.worker_fn's loop and artifacts
(/root/reference/src/saev/framework/inference.py:138-285) on the sparse forward state.

The reference runs `out = sae(acts)` per batch, keeps the DENSE `out.f_x[B, d_sae]`, copies it to the host (4.3 GB per
batch at the c3 shape) for `scipy.sparse.csr_array`, and reduces it three more times on the device (`mean_values`,
`sparsity`, `distributions`).  Here every batch is one eval-mode forward of the kernels (`saev_b200_forward`); the
per-atom statistics come from the [B, K] top-k lists (`saev_b200_eval_accumulate`), the CSR block from the same lists
(`sparse.topk_to_csr_parts`: only the lists cross PCIe), and the fp64 NMSE accumulators from the residual the
forward left behind -- the [B, d_sae] matrix is never formed.

    result = run(sae, batches, content_tokens_per_example=T, n_samples=N, n_dists=25, ignore_labels=[...])
    result.metrics            # Metrics.from_accumulators fields (saev/metrics.py:15-159) as a dict
    result.token_acts         # scipy.sparse.csr_array [n_tokens_total, d_sae]     (inference.py:236-251)
    result.mean_values, result.sparsity, result.distributions                     (inference.py:226-256)

`batches` is any iterable of ordered batches with saev's OrderedDataLoader schema (`act`, `example_idx`, `token_idx`,
optionally `token_labels`; ordered.py) -- the ordered reader itself is saev's (not on the training hot path);
`ordered_batches()` below is a minimal in-process reader of the same shard format for tests and small jobs.
"""

from __future__ import annotations

import dataclasses
import json
import math
import pathlib

import numpy as np
import torch
from torch import Tensor

from . import sparse as _sparse


@dataclasses.dataclass
class InferenceResult:
    metrics: dict
    token_acts: object | None      # scipy.sparse.csr_array or None (save=False)
    mean_values: Tensor | None     # [d_sae] sum of activations / number of firings          (inference.py:248)
    sparsity: Tensor | None        # [d_sae] firings / n_samples                              (inference.py:249)
    distributions: Tensor | None   # [n_samples, n_dists] first n_dists activations per token (inference.py:226)


def metrics_from_accumulators(*, sse_recon: float, sse_baseline: float, n_tokens: int, d_model: int) -> dict:
    """saev/metrics.py:99-131 (`Metrics.from_accumulators(...).to_dict()`)."""
    assert n_tokens > 0, f"n_tokens must be positive, got {n_tokens}."
    assert d_model > 0 and sse_recon >= 0.0, (d_model, sse_recon)
    assert sse_baseline > 0.0, f"sse_baseline must be > 0, got {sse_baseline}."
    n_elements = n_tokens * d_model
    return dict(mse_per_dim=sse_recon / n_elements, mse_per_token=sse_recon / n_tokens,
                normalized_mse=sse_recon / sse_baseline, baseline_mse_per_dim=sse_baseline / n_elements,
                baseline_mse_per_token=sse_baseline / n_tokens, sse_recon=sse_recon, sse_baseline=sse_baseline,
                n_tokens=n_tokens, d_model=d_model, n_elements=n_elements)


@torch.no_grad()
def run(sae, batches, *, content_tokens_per_example: int, n_samples: int, n_dists: int = 25, ignore_labels=(),
        save: bool = True) -> InferenceResult:
    """inference.py:171-285 for a `saev_b200.nn.SparseAutoencoder` (TopK) on a CUDA device.

    Masked tokens (`token_labels` in `ignore_labels`) are left out of every statistic and get an empty CSR row
    (:199-203, :234).  The pass checks, like the reference (:228-233), that the batches arrive in token order."""
    import scipy.sparse

    dev = sae.W_dec.device
    if dev.type != "cuda":
        raise RuntimeError("saev_b200.inference.run needs the SAE on a CUDA device (there is no CPU path)")
    sae.eval()
    ignore = torch.tensor(list(ignore_labels), dtype=torch.int64)
    S, D = sae.cfg.d_sae, sae.cfg.d_model
    state = None
    blocks, prev_i, n_tokens = [], -1, 0
    distributions = torch.zeros((n_samples, n_dists), device=dev) if save else None
    for batch in batches:
        x = batch["act"].to(dev).contiguous()
        bsz = x.shape[0]
        mask = torch.ones(bsz, dtype=torch.bool)
        if "token_labels" in batch:
            mask = torch.isin(batch["token_labels"].to(torch.int64).cpu(), ignore, invert=True)
        n_valid = int(mask.sum())
        n_tokens += n_valid
        # ordering checks of the reference (:228-233), on the host copies of the two index vectors
        ex, tok = batch["example_idx"].cpu().to(torch.int64), batch["token_idx"].cpu().to(torch.int64)
        if save:
            bidx = ex * content_tokens_per_example + tok
            assert int(bidx[0]) == prev_i + 1, "batches must arrive in token order (saev OrderedDataLoader)"
            assert bool((torch.arange(int(bidx[0]), int(bidx[-1]) + 1) == bidx).all()), "batch is not a contiguous token range"
            prev_i = int(bidx[-1])
        if n_valid > 0:
            xv = x if n_valid == bsz else x[mask.to(dev)].contiguous()
            out = sae(xv)  # eval forward of the kernels: top-k lists + residual stay on the device
            eng = sae.engine
            if state is None:
                state = eng.new_eval_state()
            eng.eval_accumulate(xv, state)  # fp64 sum x^2, sum x, sum r^2; per-atom firings / activation sums
            if save:
                idx, val = eng.topk_idx[:n_valid], eng.topk_val[:n_valid]
                # distributions[example_idx[mask], :] = f_x[mask, :n_dists]  (:226).  As written there the row index is
                # the EXAMPLE, so the tokens of one example overwrite each other; the sequential CPU assignment keeps
                # the last one -- reproduced here by writing only the last valid token of every example
                rows = ex[mask]
                last = torch.ones(n_valid, dtype=torch.bool)
                last[:-1] = rows[1:] != rows[:-1]
                dense_head = torch.zeros(n_valid, n_dists, device=dev)
                hit = (idx >= 0) & (idx < n_dists)
                r, k = hit.nonzero(as_tuple=True)
                dense_head[r, idx[r, k].long()] = val[r, k]
                distributions[rows[last].to(dev)] = dense_head[last.to(dev)]
                indptr, indices, data = _sparse.topk_to_csr_parts(idx, val, S)
                # re-insert the masked tokens as empty rows
                counts = torch.zeros(bsz, dtype=torch.int64)
                counts[mask] = (indptr[1:] - indptr[:-1]).cpu()
                full_ptr = torch.zeros(bsz + 1, dtype=torch.int64)
                full_ptr[1:] = torch.cumsum(counts, 0)
                blocks.append(scipy.sparse.csr_array((data.cpu().numpy(), indices.cpu().numpy(), full_ptr.numpy()),
                                                     shape=(bsz, S)))
        elif save:
            blocks.append(scipy.sparse.csr_array((bsz, S), dtype=np.float32))
    assert n_tokens > 0, "Inference dataloader yielded zero valid tokens; cannot compute metrics."
    acc = state["acc"].cpu()
    sum_vec = acc[8:]
    sse_baseline = float(acc[0] - torch.dot(sum_vec, sum_vec) / n_tokens)
    if sse_baseline <= 0.0:
        raise RuntimeError(f"Baseline variance is non-positive (sse_baseline={sse_baseline:.6e}); cannot compute normalized MSE.")
    metrics = metrics_from_accumulators(sse_recon=float(acc[1]), sse_baseline=sse_baseline, n_tokens=n_tokens, d_model=D)
    if not save:
        return InferenceResult(metrics, None, None, None, None)
    n_fired, values = state["n_fired"], state["values"]
    token_acts = scipy.sparse.vstack(blocks, format="csr")
    return InferenceResult(metrics, token_acts, (values / n_fired).cpu(), (n_fired / n_samples).cpu(), distributions.cpu())


def ordered_batches(shards_dir, layer: int, batch_size: int, *, labels: bool = False):
    """Minimal in-order reader of a saev shard directory (metadata.json, shards.json, acts%06d.bin; shards.py:43-180):
    yields `{"act", "example_idx", "token_idx"[, "token_labels"]}` over the content tokens of `layer`, example by
    example, token by token -- the order saev's OrderedDataLoader delivers (ordered.py:73-198).  Host-side, pageable."""
    from . import data as _data

    shards_dir = pathlib.Path(shards_dir)
    md = _data.Metadata.load(shards_dir)
    T, D, li = md.content_tokens_per_example, md.d_model, md.layers.index(layer)
    batch_size = batch_size // T * T  # whole examples per batch (inference.py:163-167)
    assert batch_size > 0
    info = json.loads((shards_dir / "shards.json").read_text())
    lab = None
    if labels:
        lab = np.memmap(shards_dir / "labels.bin", mode="r", dtype=np.uint8, shape=(md.n_examples, T))
    acts, exs = [], []
    n_buf = 0

    def flush():
        nonlocal acts, exs, n_buf
        a = torch.from_numpy(np.concatenate(acts).reshape(-1, D))
        e = torch.from_numpy(np.repeat(np.concatenate(exs), T).astype(np.int32))
        t = torch.arange(T, dtype=torch.int32).repeat(len(a) // T)
        out = {"act": a, "example_idx": e, "token_idx": t}
        if lab is not None:
            out["token_labels"] = torch.from_numpy(np.ascontiguousarray(lab[np.concatenate(exs)]).reshape(-1))
        acts, exs, n_buf = [], [], 0
        return out

    for s, entry in enumerate(info):
        n = int(entry["n_examples"])
        if n == 0:
            continue
        mm = np.memmap(shards_dir / entry["name"], mode="r", dtype=np.float32, shape=md.shard_shape)
        per = batch_size // T
        e0 = s * md.examples_per_shard
        i = 0
        while i < n:
            take = min(n - i, per - n_buf)
            acts.append(np.ascontiguousarray(mm[i : i + take, li, int(md.cls_token):, :]))
            exs.append(np.arange(e0 + i, e0 + i + take))
            n_buf += take
            i += take
            if n_buf == per:
                yield flush()
    if n_buf:
        yield flush()
