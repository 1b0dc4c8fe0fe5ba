"""`saev_b200.install()` — make an UNMODIFIED saev checkout train through the B200 kernels.

saev's loop looks every hot-path name up by attribute at call time
(/root/reference/src/saev/framework/train.py): `nn.SparseAutoencoder` (:115), `nn.get_objective` (:119),
`saev.data.ShuffledDataLoader` (:259, :534), `torch.optim.Adam` (:294), `torch.nn.utils.clip_grad_norm_` (:358),
`evaluate` (:198, a module global of saev.framework.train).
`install()` rebinds exactly those names to the classes of this package and `uninstall()` restores them, so

    import saev.framework.train, saev_b200
    saev_b200.install()
    saev.framework.train.worker_fn([cfg])          # cfg.device = "cuda"

runs saev's own train()/evaluate()/make_saes() with the forward, backward, clip and Adam of every step inside
libsaev_b200.so.  Checkpoints are written by saev's `nn.dump` (state_dict keys W_dec, b_dec, W_enc, b_enc are
unchanged) and read back by `saev.nn.load`.

The model class that is bound is a subclass of BOTH `saev_b200.nn.SparseAutoencoder` (behaviour) and saev's
`SparseAutoencoder` (identity: saev's functions are beartype-checked against its own class, modeling.py:548).
"""

from __future__ import annotations

import importlib

import torch

from . import nn as _nn
from . import optim as _optim

_saved: dict[tuple[object, str], object] = {}
_dropin_cls = None


def _bind(obj, name: str, value) -> None:
    key = (obj, name)
    if key not in _saved:
        _saved[key] = getattr(obj, name)
    setattr(obj, name, value)


def dropin_class():
    """The SparseAutoencoder class `install()` binds to `saev.nn.SparseAutoencoder`."""
    global _dropin_cls
    if _dropin_cls is None:
        ref = importlib.import_module("saev.nn.modeling")
        _dropin_cls = type("SparseAutoencoder", (_nn.SparseAutoencoder, ref.SparseAutoencoder),
                           {"__doc__": _nn.SparseAutoencoder.__doc__, "__module__": __name__})
    return _dropin_cls


def install(*, model: bool = True, optimizer: bool = True, loader: bool = True, evaluate: bool = True,
            data_parallel: bool | str = "auto"):
    """Rebind saev's hot-path names to saev_b200 (idempotent).  Needs `saev` importable; raises ImportError
    otherwise.  `data_parallel="auto"` turns the gradient all-reduce on when torch.distributed is initialised
    with more than one rank."""
    saev_nn = importlib.import_module("saev.nn")
    saev_obj = importlib.import_module("saev.nn.objectives")
    if model:
        cls = dropin_class()
        dp = data_parallel
        if dp == "auto":
            dp = torch.distributed.is_available() and torch.distributed.is_initialized() and \
                torch.distributed.get_world_size() > 1

        def make(cfg):
            sae = cls(cfg)
            return sae.data_parallel() if dp else sae

        make.__doc__ = cls.__doc__
        _bind(saev_nn, "SparseAutoencoder", make if dp else cls)
        _bind(saev_nn, "get_objective", _nn.get_objective)
        _bind(saev_obj, "get_objective", _nn.get_objective)
    if optimizer:
        _bind(torch.optim, "Adam", _optim.FusedAdam)
        _bind(torch.nn.utils, "clip_grad_norm_", _optim.clip_grad_norm_)
    if loader:
        from . import data as _data

        saev_data = importlib.import_module("saev.data")
        _bind(saev_data, "ShuffledDataLoader", _data.ShuffledDataLoader)
    if evaluate and model:
        # train.py:198 looks `evaluate` up as a module global; ours accumulates the per-atom statistics from the sparse
        # forward state and returns saev's own EvalMetrics
        from . import evaluate as _eval

        ref_train = importlib.import_module("saev.framework.train")
        saev_data = importlib.import_module("saev.data")

        def evaluate_b200(cfgs, saes, objectives):
            return _eval.evaluate(cfgs, saes, objectives, metrics_cls=ref_train.EvalMetrics,
                                  loader_cls=saev_data.ShuffledDataLoader)

        evaluate_b200.__doc__ = _eval.evaluate.__doc__
        _bind(ref_train, "evaluate", evaluate_b200)
    return None


def uninstall() -> None:
    """Restore every name `install()` rebound."""
    for (obj, name), value in list(_saved.items()):
        setattr(obj, name, value)
    _saved.clear()


def installed() -> bool:
    return bool(_saved)
