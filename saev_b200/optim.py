"""Optimizer-side pieces of the drop-in boundary: gradient clipping and the Adam update.

saev's loop (/root/reference/src/saev/framework/train.py) does, per step,

    grad_norm = torch.nn.utils.clip_grad_norm_(sae.parameters(), max_norm=cfg.grad_clip)   # train.py:358-360
    opt.step()                      # torch.optim.Adam([{"params": sae.parameters(), "lr": 0.0}], fused=True)  :294,444-446
    pg["lr"] = sched.step()         # :449-451
    opt.zero_grad()                 # :456-458

`clip_grad_norm_` and `FusedAdam` below are what `saev_b200.install()` binds to those two names.  For the four
parameters of a `saev_b200.nn.SparseAutoencoder` whose gradients came out of `saev_b200_backward` they launch
`saev_b200_grad_sumsq` and `saev_b200_adam_step` (clip scale folded into the Adam kernel, so the gradient bucket
is read once); every other parameter is passed to the stock torch implementation untouched.
"""

from __future__ import annotations

import weakref

import torch
from torch import Tensor

_torch_clip_grad_norm_ = torch.nn.utils.clip_grad_norm_
_TorchAdam = torch.optim.Adam


def _owner(p):
    ref = getattr(p, "_b200_owner", None)
    return ref() if ref is not None else None


def _fused_owner(params):
    """The SparseAutoencoder all of `params` belong to, if they are exactly its four parameters and their .grad
    tensors are the views of the engine's flat gradient bucket (i.e. produced by the fused backward)."""
    if len(params) != 4:
        return None
    sae = _owner(params[0])
    if sae is None or sae.engine is None or not sae._grads_fused:
        return None
    if any(_owner(p) is not sae for p in params):
        return None
    eng = sae.engine
    ptrs = {eng.gW_enc_t.data_ptr(), eng.gb_enc.data_ptr(), eng.gW_dec.data_ptr(), eng.gb_dec.data_ptr()}
    if any(p.grad is None or p.grad.data_ptr() not in ptrs for p in params):
        return None
    return sae


def clip_grad_norm_(parameters, max_norm, norm_type: float = 2.0, error_if_nonfinite: bool = False, foreach=None):
    """torch.nn.utils.clip_grad_norm_ as called at train.py:358-360.  Fused case: returns ||g||_2 of the (already
    all-reduced) gradient bucket as a 0-d device tensor and records `max_norm`; the scale
    min(1, max_norm / (||g|| + 1e-6)) is applied inside the following `FusedAdam.step()` rather than written
    back to `.grad` (reading `.grad` between the two calls shows the un-clipped gradient)."""
    params = [parameters] if isinstance(parameters, Tensor) else list(parameters)
    sae = _fused_owner(params) if float(norm_type) == 2.0 else None
    if sae is None:
        return _torch_clip_grad_norm_(params, max_norm, norm_type=norm_type, error_if_nonfinite=error_if_nonfinite,
                                      foreach=foreach)
    eng = sae.engine
    eng.grad_sumsq(local=sae._dp_world == 1)
    total = eng.sumsq.sqrt().reshape(())
    if error_if_nonfinite and not bool(torch.isfinite(total)):
        raise RuntimeError("The total norm for gradients from `parameters` is non-finite, so it cannot be clipped.")
    opt = sae._fused_adam() if sae._fused_adam is not None else None
    if opt is None:
        # the update will NOT go through saev_b200_adam_step (cfg.optim="muon", train.py:296-306, or any stock
        # optimizer): scale the gradient bucket now, exactly as torch's clip_grad_norm_ does
        coef = torch.clamp(float(max_norm) / (total + 1e-6), max=1.0)
        eng.grads.mul_(coef)
        sae._pending_clip = None
    else:
        sae._pending_clip = float(max_norm)
    return total


class FusedAdam(_TorchAdam):
    """torch.optim.Adam whose update for a fused SparseAutoencoder runs in `saev_b200_adam_step` (Adam moments and
    the step count live in the engine's flat buffers; lr, betas and eps are read from the param group every step,
    so saev's `pg["lr"] = sched.step()` works unchanged).  Parameters that do not belong to a fused SAE fall
    through to torch's own Adam."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, **kw):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad, **kw)
        # a group holding exactly the four parameters of one SparseAutoencoder will be stepped by the fused kernel:
        # tell the module, so that clip_grad_norm_ may defer the clip scale to it
        for group in self.param_groups:
            ps = group["params"]
            sae = _owner(ps[0]) if len(ps) == 4 else None
            if sae is not None and all(_owner(p) is sae for p in ps) and len({id(p) for p in ps}) == 4:
                sae._fused_adam = weakref.ref(self)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        leftovers = False
        for group in self.param_groups:
            with_grad = [p for p in group["params"] if p.grad is not None]
            sae = _fused_owner(with_grad)
            if sae is None:
                leftovers = leftovers or bool(with_grad)
                continue
            if group.get("weight_decay", 0) != 0 or group.get("amsgrad", False) or group.get("maximize", False):
                raise NotImplementedError("saev_b200 FusedAdam: weight_decay / amsgrad / maximize are not supported "
                                          "for fused SparseAutoencoder parameters")
            eng = sae.engine
            max_norm = sae._pending_clip
            if max_norm is None:
                max_norm = 0.0  # clip_grad_norm_ was not called this step: plain Adam
            lr = group["lr"]
            lr = float(lr.item()) if isinstance(lr, Tensor) else float(lr)
            renorm = bool(sae.fuse_renorm and sae.cfg.normalize_w_dec)
            eng.adam_step(lr, max_norm=max_norm, betas=tuple(group["betas"]), eps=group["eps"], renorm_w_dec=renorm)
            sae._pending_clip = None
            sae._grads_fused = False
            sae._w_dec_normalized = renorm
            for p in with_grad:  # the bucket was consumed: hide these grads from the stock update below
                p._b200_consumed = True
        if leftovers:
            hidden = []
            for group in self.param_groups:
                for p in group["params"]:
                    if getattr(p, "_b200_consumed", False) and p.grad is not None:
                        hidden.append((p, p.grad))
                        p.grad = None
            try:
                super().step()
            finally:
                for p, g in hidden:
                    p.grad = g
        for group in self.param_groups:
            for p in group["params"]:
                if getattr(p, "_b200_consumed", False):
                    p._b200_consumed = False
        return loss
