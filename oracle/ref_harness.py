"""Import and drive the UNMODIFIED reference (OSU-NLP-Group/saev) on the host CPU.  TEST / BASELINE INFRASTRUCTURE
ONLY: imported by tests/, `__graft_entry__.smoke()` and the CPU-baseline / `--impl reference` legs of bench.py, never
by anything under saev_b200/.

The reference is pure Python.  `__graft_entry__.build()` stages its package directory, unmodified, under the
git-ignored `oracle/_ref/saev` (it travels to the GPU box with the snapshot like the built .so; it is never part of
this repository's history); in the build container `/root/reference/src` is used directly when nothing is staged.
Three packages the image lacks are stubbed in `oracle/ref_stubs/` (orjson, open_clip, timm: import-time only).
"""

from __future__ import annotations

import os
import pathlib
import sys
import time

HERE = pathlib.Path(__file__).resolve().parent


def reference_src() -> pathlib.Path | None:
    """Directory that contains the reference's `saev` package, or None."""
    env = os.environ.get("SAEV_B200_REF_SRC", "")
    for cand in ([pathlib.Path(env)] if env else []) + [HERE / "_ref", pathlib.Path("/root/reference/src")]:
        if (cand / "saev" / "nn" / "modeling.py").is_file():
            return cand.resolve()
    return None


def import_reference():
    """Make `import saev` resolve to the reference.  Returns the source directory; raises ImportError if absent."""
    src = reference_src()
    if src is None:
        raise ImportError("the reference package is not staged (oracle/_ref/saev) and /root/reference is absent")
    for p in (str(src), str(HERE / "ref_stubs")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import saev.nn  # noqa: F401
    import saev.utils.scheduling  # noqa: F401

    return src


class ReferenceStep:
    """One SAE + objective + Adam(fused=True) + WarmupCosine of the reference, stepped with the statements of its
    training loop (src/saev/framework/train.py:334-362, 444-458; the logging block left out)."""

    def __init__(self, d_model: int, d_sae: int, top_k: int, *, relu: bool = False, k_aux: int = 512, lr: float = 4e-4,
                 n_lr_warmup: int = 500, n_steps: int = 10_000, grad_clip: float = 1.0, seed: int = 0,
                 device: str = "cpu"):
        import torch

        import_reference()
        import saev.nn
        import saev.nn.modeling as M
        import saev.nn.objectives as O
        import saev.utils.scheduling as sched

        torch.manual_seed(seed)
        act = M.Relu(sparsity=M.L1Sparsity(coeff=4e-4), aux=M.AuxK(k_aux=k_aux)) if relu else \
            M.TopK(top_k=top_k, aux=M.AuxK(k_aux=k_aux))
        cfg = M.SparseAutoencoderConfig(d_model=d_model, d_sae=d_sae, activation=act, reinit_blend=0.0)
        self.torch = torch
        self.sae = saev.nn.SparseAutoencoder(cfg).to(device)
        self.objective = saev.nn.get_objective(O.Matryoshka(n_prefixes=1)).to(device)
        self.pg = {"params": self.sae.parameters(), "lr": 0.0}  # train.py:118
        self.opt = torch.optim.Adam([self.pg], fused=True)      # train.py:294
        self.sched = sched.WarmupCosine(0.0, n_lr_warmup, lr, n_steps, 0.0)  # train.py:311-316
        self.grad_clip = grad_clip
        self.sae.train()
        self.objective.train()

    def step(self, x):
        torch = self.torch
        self.sae.normalize_w_dec()                                    # train.py:334-335
        loss, _ = self.objective(self.sae, x)                         # :341
        loss.loss.backward()                                          # :348
        self.sae.remove_parallel_grads()                              # :352
        torch.nn.utils.clip_grad_norm_(self.sae.parameters(), max_norm=self.grad_clip)  # :358-360
        self.opt.step()                                               # :444-446
        for pg in self.opt.param_groups:
            pg["lr"] = self.sched.step()                              # :449-451
        self.opt.zero_grad()                                          # :456-458
        return loss


def time_reference(d_model, d_sae, top_k, *, relu, rows, steps, warmup, threads=None):
    """(activations/s, ms/step, cores) of the reference's own modules on `rows` Gaussian rows per step."""
    import torch

    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = ReferenceStep(d_model, d_sae, top_k, relu=relu)
    g = torch.Generator().manual_seed(0)
    xs = [torch.randn(rows, d_model, generator=g) for _ in range(2)]
    for i in range(warmup):
        ref.step(xs[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        ref.step(xs[i % 2])
    dt = time.perf_counter() - t0
    return rows * steps / dt, dt / steps * 1e3, cores
