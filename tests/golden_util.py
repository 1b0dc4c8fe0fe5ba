"""Helpers shared by CPU (oracle) and GPU (product) tests for replaying tests/golden/*.npz."""
import pathlib

import numpy as np
import torch

from oracle import sae_oracle as orc

GOLDEN = pathlib.Path(__file__).resolve().parent / "golden"
CASES = ["tiny_topk_auxk", "tiny_topk_auxk_clamp", "tiny_topk_noaux_noproj", "tiny_relu_l1_auxk", "c1_topk",
         "c1_topk_auxk_live", "tiny_topk_matryoshka", "c1_topk_matryoshka", "tiny_relu_matryoshka", "tiny_batchtopk_auxk",
         "c1_batchtopk", "c1_batchtopk_matryoshka"]


def load_case(name):
    z = np.load(GOLDEN / f"{name}.npz")
    meta = {k[5:]: z[k].item() for k in z.files if k.startswith("meta_")}
    cfg = orc.OracleConfig(
        d_model=meta["D"], d_sae=meta["S"], activation=("topk", "relu", "batchtopk")[meta["act"]], top_k=max(meta["top_k"], 1),
        batch_momentum=meta.get("momentum", 0.1),
        l1_coeff=meta["l1_coeff"], aux=bool(meta["aux"]), k_aux=max(meta["k_aux"], 1), aux_alpha=meta["alpha"],
        dead_threshold_tokens=meta["dead_thr"], normalize_w_dec=bool(meta["normalize"]),
        remove_parallel_grads=bool(meta["remove_parallel"]), lr=meta["lr"], n_lr_warmup=meta["n_warmup"],
        n_steps=meta["sched_steps"], grad_clip=meta["grad_clip"],
    )
    return z, meta, cfg


def prefixes_of(z, step):
    """The Matryoshka cut points the reference drew for forward number `step` (None for single-prefix runs)."""
    if "prefixes" not in z.files or z["prefixes"].shape[1] <= 1:
        return None
    return [int(c) for c in z["prefixes"][step]]


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).flatten()
    b = torch.as_tensor(b, dtype=torch.float64).flatten()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))
