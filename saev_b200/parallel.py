"""Data-parallel training step: one process per GPU, full parameter + Adam replica per rank, batch rows
sharded across ranks, ONE gradient all-reduce per step (NCCL over NVLink) plus a 4*d_sae-byte MAX
all-reduce of the activity flags so every rank keeps an identical dead-latent tracker.

saev itself has no multi-GPU path; the semantics implemented here are "N ranks x B rows
behave exactly like one rank x (N*B) rows": every kernel divides by the GLOBAL batch (`tokens_global`), so
per-rank gradients/loss partials simply add up, `remove_parallel_grads` is linear and is applied before the
reduction, and the clip norm is taken on the reduced gradient.
"""

from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib
from .engine import Engine


class DataParallelTrainer:
    def __init__(self, engine: Engine, group=None):
        self.eng = engine
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self._w_dec_normalized = False

    def broadcast_params(self, src: int = 0) -> None:
        """Make every replica start from rank `src`'s parameters (model init is unseeded in saev)."""
        if self.world > 1:
            dist.broadcast(self.eng.params, src=src, group=self.group)
            self.eng.sync_weights()
        self._w_dec_normalized = False

    def step(self, x: torch.Tensor, lr: float, *, max_norm: float = 1.0, fused_renorm: bool = True) -> torch.Tensor:
        """One iteration of saev's loop body (train.py:332-460) on this rank's rows `x[B_local, D]`.
        Returns the device tensor of this rank's loss partials (see `global_losses`)."""
        eng = self.eng
        tokens_global = x.shape[0] * self.world  # equal per-rank batches (the loader guarantees it)
        if not self._w_dec_normalized:
            eng.normalize_w_dec()
        if self.world == 1:
            eng.forward(x, training=True, tokens_global=tokens_global)
        else:
            eng.forward(x, training=True, phase=_lib.PHASE_A, tokens_global=tokens_global)
            dist.all_reduce(eng.active_flags(), op=dist.ReduceOp.MAX, group=self.group)
            eng.forward(x, training=True, phase=_lib.PHASE_B, tokens_global=tokens_global)
        eng.backward(x, tokens_global=tokens_global)
        if self.world > 1:
            dist.all_reduce(eng.grads, op=dist.ReduceOp.SUM, group=self.group)
        eng.grad_sumsq(local=self.world == 1)
        renorm = fused_renorm and eng.cfg.normalize_w_dec
        eng.adam_step(lr, max_norm=max_norm, renorm_w_dec=renorm)
        self._w_dec_normalized = renorm
        return eng.losses

    def global_losses(self) -> dict:
        """Loss scalars of the global batch (host sync; call on log steps only)."""
        eng = self.eng
        vals = eng.losses.clone()
        if self.world > 1:
            n_dead = vals[5].clone()
            vals[5] = 0
            dist.all_reduce(vals, op=dist.ReduceOp.SUM, group=self.group)
            vals[5] = n_dead  # identical on every rank, not additive
        from .engine import LOSS_KEYS

        return dict(zip(LOSS_KEYS, vals.tolist()))
