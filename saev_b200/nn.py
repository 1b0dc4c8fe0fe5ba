"""Host-side mirror of `saev.nn` (model + objective API) on top of the CUDA engine.

Same names, argument meaning and error behaviour as the reference
(/root/reference/src/saev/nn/modeling.py, objectives.py) so that saev's training loop
(src/saev/framework/train.py:109-618) can use these classes unchanged:

    sae = SparseAutoencoder(cfg)                 # parameters W_dec, b_dec, W_enc, b_enc  (modeling.py:306-329)
    objective = get_objective(Matryoshka(n_prefixes=1))
    sae.normalize_w_dec()                         # train.py:334-335
    loss, fwd = objective(sae, acts)              # train.py:341   -> saev_b200_forward
    loss.loss.backward()                          # train.py:348   -> saev_b200_backward (writes .grad)
    sae.remove_parallel_grads()                   # train.py:352   (already folded into backward: no-op)
    clip_grad_norm_(sae.parameters(), c)          # train.py:358   -> saev_b200_grad_sumsq (see optim.py)
    opt.step()                                    # train.py:444   -> saev_b200_adam_step (optim.FusedAdam)

All arithmetic runs in libsaev_b200.so; modules can be constructed, saved and loaded on the CPU, but a
forward on a CPU tensor raises (there is no CPU fallback).

Differences from the reference that are deliberate and documented in DESIGN.md:
  * `W_enc` is an nn.Parameter of shape [d_model, d_sae] as in saev, but once on the GPU it is a transposed
    VIEW of the atom-major master copy W_enc_t[d_sae, d_model] the kernels use (state_dict keys and values
    are unchanged, checkpoints interoperate with saev.nn.load / saev.nn.dump).
  * `Output.h_x`, `Output.f_x`, `Output.x_hats` are materialised lazily (the fused path never writes the
    [B, d_sae] matrices); saev's logging block (train.py:365-442) reads them on log steps only.
  * Relu runs the dense path (five error-compensated bf16 split contractions on tcgen05); with Matryoshka prefixes the
    decoder, dh and W_dec-gradient contractions run once per prefix block on a window of the same operands.  BatchTopK (modeling.py:183-244) runs on the sparse path:
    per-row top-`capacity` lists (capacity = min(128, d_sae)), then a batch-wide selection kernel; a row that would
    need more than `capacity` slots is counted and `Loss.metrics()` raises (single rank only: the global selection
    does not shard).
"""

from __future__ import annotations

import dataclasses
import io
import json
import math
import pathlib
import typing as tp
import weakref

import torch
import torch.distributed as dist
from torch import Tensor

from . import __version__
from .engine import LOSS_KEYS, Engine, EngineConfig

SCHEMA_VERSION = 5  # modeling.py:20


# ----------------------------------------------------------------------------------------------
# configuration dataclasses (modeling.py:23-146, 259-284; objectives.py:13-25)
# ----------------------------------------------------------------------------------------------
@dataclasses.dataclass(frozen=True)
class NoSparsity:
    key: str = "no-sparsity"


@dataclasses.dataclass(frozen=True)
class L1Sparsity:
    key: str = "l1-sparsity"
    coeff: float = 1e-4


@dataclasses.dataclass(frozen=True)
class NoAux:
    key: str = "no-aux"


@dataclasses.dataclass(frozen=True)
class AuxK:
    key: str = "auxk"
    k_aux: int = 512
    alpha: float = 1 / 32


@dataclasses.dataclass(frozen=True)
class Relu:
    key: str = "relu"
    sparsity: tp.Any = L1Sparsity(coeff=4e-4)
    aux: tp.Any = NoAux()


@dataclasses.dataclass(frozen=True)
class TopK:
    key: str = "top-k"
    top_k: int = 32
    sparsity: tp.Any = NoSparsity()
    aux: tp.Any = AuxK()

    def __post_init__(self):
        assert self.top_k > 0, "top_k must be a positive integer."


@dataclasses.dataclass(frozen=True)
class BatchTopK:
    key: str = "batch-top-k"
    top_k: int = 32
    sparsity: tp.Any = NoSparsity()
    momentum: float = 0.1
    aux: tp.Any = AuxK()


@dataclasses.dataclass(frozen=True)
class SparseAutoencoderConfig:
    d_model: int = 1024
    d_sae: int = 1024 * 16
    activation: tp.Any = TopK()
    reinit_blend: float = 0.8
    reinit_enc_dec_tranpose: bool = True
    remove_parallel_grads: bool = True
    normalize_w_dec: bool = True


@dataclasses.dataclass(frozen=True)
class Matryoshka:
    n_prefixes: int = 10
    dead_threshold_tokens: int = 10_000_000


ObjectiveConfig = Matryoshka


def _kind(obj) -> str:
    """Class-name based dispatch so that saev's own config dataclasses are accepted as well as ours."""
    return type(obj).__name__


def sample_prefixes(d_sae: int, n_prefixes: int, min_prefix_length: int = 1, pareto_power: float = 0.5) -> Tensor:
    """objectives.py:158-201, same torch CPU ops on the same (global) generator, so a seeded run draws the cuts the
    reference would: n_prefixes - 1 lengths without replacement from a Pareto-shaped distribution over 1..d_sae-1,
    plus d_sae, sorted ascending."""
    if n_prefixes <= 1:
        return torch.tensor([d_sae], dtype=torch.int64)
    assert n_prefixes <= d_sae
    lengths = torch.arange(1, d_sae)
    pareto_cdf = 1 - ((min_prefix_length / lengths.float()) ** pareto_power)
    pareto_pdf = torch.cat([pareto_cdf[:1], pareto_cdf[1:] - pareto_cdf[:-1]])
    probability_dist = pareto_pdf / pareto_pdf.sum()
    sampled = torch.multinomial(probability_dist, num_samples=n_prefixes - 1, replacement=False)
    prefixes = torch.cat((lengths[sampled].detach().clone(), torch.tensor([d_sae])))
    prefixes, _ = torch.sort(prefixes, descending=False)
    return prefixes.to(torch.int64)


def engine_config(sae_cfg, obj_cfg, max_batch: int) -> EngineConfig:
    act = sae_cfg.activation
    kind = _kind(act)
    if kind not in ("TopK", "Relu", "BatchTopK"):
        raise TypeError(f"unknown activation config {act!r}")
    aux = getattr(act, "aux", None)
    sp = getattr(act, "sparsity", None)
    top_k, batch_k = getattr(act, "top_k", 1), 0
    if kind == "BatchTopK":
        # the sparse forward state holds `capacity` slots per row; BatchTopK.top_k is the AVERAGE per sample
        batch_k, top_k = top_k, batch_topk_capacity(top_k, sae_cfg.d_sae, int(getattr(obj_cfg, "n_prefixes", 1)))
    return EngineConfig(
        d_model=sae_cfg.d_model,
        d_sae=sae_cfg.d_sae,
        top_k=top_k,
        batch_k=batch_k,
        batch_momentum=float(getattr(act, "momentum", 0.1)),
        activation="relu" if kind == "Relu" else "topk",
        aux=_kind(aux) == "AuxK",
        k_aux=getattr(aux, "k_aux", 512),
        aux_alpha=getattr(aux, "alpha", 1 / 32),
        l1_coeff=float(getattr(sp, "coeff", 0.0)) if _kind(sp) == "L1Sparsity" else 0.0,
        dead_threshold_tokens=getattr(obj_cfg, "dead_threshold_tokens", 10_000_000),
        remove_parallel_grads=sae_cfg.remove_parallel_grads,
        normalize_w_dec=sae_cfg.normalize_w_dec,
        max_batch=max_batch,
        max_prefixes=max(1, min(int(getattr(obj_cfg, "n_prefixes", 1)), sae_cfg.d_sae)),
    )


def batch_topk_capacity(top_k: int, d_sae: int, n_prefixes: int = 1) -> int:
    """Slots per row of the sparse forward state under BatchTopK: the most the kernels hold (128: screen, re-score,
    repair, prefix decode), or d_sae when the dictionary is narrower.  SAEV_B200_BATCHTOPK_CAP lowers it (cheaper re-score when rows are known to be
    balanced; the tests use it to provoke truncation)."""
    import os

    cap = min(128, d_sae)
    env = os.environ.get("SAEV_B200_BATCHTOPK_CAP", "")
    if env:
        cap = max(1, min(cap, int(env)))
    if top_k > cap and cap < d_sae:
        raise NotImplementedError(f"BatchTopK(top_k={top_k}): the sparse forward state holds at most {cap} actives per "
                                  "row, which cannot even hold the average")
    return cap


# ----------------------------------------------------------------------------------------------
# lazy forward outputs
# ----------------------------------------------------------------------------------------------
class EncodeOut(tp.NamedTuple):
    h_x: Tensor
    f_x: Tensor


class Output:
    """Mirror of SparseAutoencoder.Output (modeling.py:299-304) with lazily materialised dense fields.
    Valid until the next forward of the same SAE (it reads the engine's top-k buffers)."""

    def __init__(self, sae: "SparseAutoencoder", x: Tensor, ticket: int):
        self._sae, self._x, self._ticket = sae, x, ticket
        self._cache: dict[str, Tensor] = {}

    def _check(self):
        if self._sae._ticket != self._ticket:
            raise RuntimeError("this Output refers to an earlier forward; its lazy fields are no longer available")

    @property
    def f_x(self) -> Tensor:
        if "f_x" not in self._cache:
            self._check()
            self._cache["f_x"] = self._sae.engine.dense_f_x(self._x.shape[0])
        return self._cache["f_x"]

    @property
    def x_hats(self) -> Tensor:
        if "x_hats" not in self._cache:
            self._check()
            self._cache["x_hats"] = self._sae.engine.x_hats(self._x)
        return self._cache["x_hats"]

    @property
    def h_x(self) -> Tensor:
        if "h_x" not in self._cache:
            self._check()
            eng = self._sae.engine
            # dense pre-activations through the 6-term (three-piece) split tensor-core product (fp32-class accuracy)
            self._cache["h_x"] = eng.gemm_nt(self._x, eng.W_enc_t, eng.b_enc, 6)
        return self._cache["h_x"]

    def __iter__(self):  # NamedTuple-style unpacking: h_x, f_x, x_hats
        return iter((self.h_x, self.f_x, self.x_hats))

    def f_x_csr(self, row_mask: Tensor | None = None):
        """`scipy.sparse.csr_array(fwd.f_x)` (inference.py:236) without the dense matrix: built from the top-k lists of
        the forward (TopK); the ReLU path, whose f_x is dense, converts it the reference's way."""
        self._check()
        eng = self._sae.engine
        B = self._x.shape[0]
        if eng.cfg.activation == "topk":
            from .sparse import topk_to_csr

            return topk_to_csr(eng.topk_idx[:B], eng.topk_val[:B], eng.S, row_mask)
        import scipy.sparse

        f = self.f_x if row_mask is None else self.f_x * row_mask.to(self.f_x.device)[:, None]
        return scipy.sparse.csr_array(f.cpu().numpy())

    def log_metrics(self) -> dict[str, float]:
        """What saev's log block (train.py:380-423) derives from `acts_BD`, `fwd.x_hats`, `fwd.f_x` and `sae.W_dec`
        -- explained_variance, dead_unit_pct, dictionary_coherence, avg_decoder_row_norm, sse_sae, sse_baseline,
        normalized_mse -- computed from the sparse forward state without the dense [B, S] / [S, S] matrices the
        reference forms there (one host sync, like the `.item()` calls it replaces)."""
        self._check()
        return self._sae.engine.log_metrics_dict(self._x)


class _Activation(torch.nn.Module):
    """Placeholder for `sae.activation` (objectives.py:149 reads `.cfg.sparsity`).  BatchTopK carries the reference's
    `threshold` buffer (modeling.py:213; state_dict key `activation.threshold`), aliased to the engine's device scalar
    once the engine exists so that the selection kernel's EMA update lands in it."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        if _kind(cfg) == "BatchTopK":
            self.register_buffer("threshold", torch.tensor(0.0))


# ----------------------------------------------------------------------------------------------
# model
# ----------------------------------------------------------------------------------------------
class SparseAutoencoder(torch.nn.Module):
    """Sparse auto-encoder with saev's parameterisation (modeling.py:288-445)."""

    EncodeOut = EncodeOut
    Output = Output

    def __init__(self, cfg):
        torch.nn.Module.__init__(self)  # explicit: dropin.dropin_class() also inherits from saev's class
        self.cfg = cfg
        S, D = cfg.d_sae, cfg.d_model
        # modeling.py:312-327: kaiming_uniform_ on [S, D] (bound sqrt(6 / D)), rows normalised, W_enc = W_dec.T
        W_dec = torch.nn.init.kaiming_uniform_(torch.empty(S, D))
        self.W_dec = torch.nn.Parameter(W_dec)
        self.b_dec = torch.nn.Parameter(torch.zeros(D))
        if cfg.normalize_w_dec:
            self.W_dec.data /= torch.norm(self.W_dec.data, dim=1, keepdim=True)
        # atom-major master copy, exposed with saev's [D, S] shape as a transposed view
        self.W_enc = torch.nn.Parameter(self.W_dec.data.clone().t())
        self.b_enc = torch.nn.Parameter(torch.zeros(S))
        self.activation = _Activation(cfg.activation)
        self.engine: Engine | None = None
        self._obj_cfg = Matryoshka(n_prefixes=1)
        self._max_batch = 0
        self._ticket = 0
        self._grads_fused = False
        self._w_dec_normalized = False
        self._pending_clip: float | None = None  # max_norm recorded by saev_b200.optim.clip_grad_norm_
        # Fold normalize_w_dec (train.py:334-335) into the tail of the Adam kernel.  Off by default: the reference
        # normalises at the START of the next step, so its checkpoints hold un-normalised rows.
        self.fuse_renorm = False
        # loss.loss.backward() passes an upstream gradient of exactly 1 (so does (loss_a + loss_b).backward()); anything
        # else (a scaled loss) is NOT honoured by the fused backward.  It is counted on the device and raised by the
        # next Loss.metrics() / check_grad_out() -- the places that synchronise anyway.
        self._check_grad_out = True
        self._bad_grad_out: Tensor | None = None
        self._dp_group = None
        self._dp_world = 1
        self._dp_synced = False      # replicas start from rank 0's parameters (first _bind after data_parallel())
        self._fused_adam = None      # weakref to the FusedAdam that owns exactly these four parameters (optim.py)
        self._seen_versions = None   # tensor version counters at the last bind: in-place writes by torch bump them
        ref = weakref.ref(self)
        for p in (self.W_dec, self.b_dec, self.W_enc, self.b_enc):
            p._b200_owner = ref

    def data_parallel(self, group=None) -> "SparseAutoencoder":
        """Opt in to data-parallel training over `group` (default process group): every rank holds a full replica,
        feeds its own rows, and the objective's backward all-reduces the flat gradient bucket once per step
        (saev itself is single-GPU; semantics are N ranks x B rows == 1 rank x N*B rows, see parallel.py)."""
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("data_parallel(): torch.distributed is not initialised")
        self._dp_group = group
        self._dp_world = dist.get_world_size(group)
        self._dp_synced = False
        return self

    # ---- engine binding --------------------------------------------------------------------
    def _bind(self, batch: int, obj_cfg=None) -> Engine:
        """Create (or grow) the CUDA engine and re-point the parameters at its flat buffers."""
        dev = self.W_dec.device
        if dev.type != "cuda":
            raise RuntimeError("saev_b200.nn.SparseAutoencoder runs on CUDA only: move the module to a B200 "
                               "(.to('cuda')); there is no CPU fallback")
        if obj_cfg is not None:
            self._obj_cfg = obj_cfg
        eng = self.engine
        need_new = (
            eng is None
            or eng.device != dev
            or batch > eng.cfg.max_batch
            or eng.cfg.dead_threshold_tokens != self._obj_cfg.dead_threshold_tokens
            or eng.cfg.max_prefixes < min(int(getattr(self._obj_cfg, "n_prefixes", 1)), self.cfg.d_sae)
        )
        if need_new:
            new = Engine(engine_config(self.cfg, self._obj_cfg, max(batch, self._max_batch)), device=dev)
            self._max_batch = new.cfg.max_batch
            with torch.no_grad():
                new.W_enc_t.copy_(self.W_enc.data.t())
                new.b_enc.copy_(self.b_enc.data)
                new.W_dec.copy_(self.W_dec.data)
                new.b_dec.copy_(self.b_dec.data)
                if eng is not None:
                    new.m.copy_(eng.m)
                    new.v.copy_(eng.v)
                    new.toks_since_active.copy_(eng.toks_since_active)
                    new.step_count = eng.step_count
            new.sync_weights()
            self.engine = eng = new
        # parameters must alias the engine buffers (they stop doing so after .to(), load_state_dict on a fresh
        # module, or a `.data =` assignment); copy the current values in and re-point
        if self.W_dec.data_ptr() != eng.W_dec.data_ptr() or self.W_enc.data_ptr() != eng.W_enc_t.data_ptr():
            with torch.no_grad():
                eng.W_enc_t.copy_(self.W_enc.data.t())
                eng.b_enc.copy_(self.b_enc.data)
                eng.W_dec.copy_(self.W_dec.data)
                eng.b_dec.copy_(self.b_dec.data)
            self.W_enc.data = eng.W_enc_t.t()
            self.b_enc.data = eng.b_enc
            self.W_dec.data = eng.W_dec
            self.b_dec.data = eng.b_dec
            eng.sync_weights()
            self._w_dec_normalized = False
        elif self._seen_versions != self._param_versions():
            # in-place writes through torch (load_state_dict's param.copy_, a stock optimizer such as Muon, W.mul_()):
            # same storage, new values -- the screen's fp16 copy and norm bounds are stale
            eng.sync_weights()
            self._w_dec_normalized = False
        if self._dp_world > 1 and not self._dp_synced:
            # saev's init is unseeded and datapoint init reads rank-local shards: without this every replica would start
            # from different weights and stay different (the summed gradient is the only thing the ranks share)
            dist.broadcast(eng.params, src=dist.get_global_rank(self._dp_group, 0) if self._dp_group is not None else 0,
                           group=self._dp_group)
            dist.broadcast(eng.m, src=dist.get_global_rank(self._dp_group, 0) if self._dp_group is not None else 0,
                           group=self._dp_group)
            dist.broadcast(eng.v, src=dist.get_global_rank(self._dp_group, 0) if self._dp_group is not None else 0,
                           group=self._dp_group)
            eng.sync_weights()
            self._w_dec_normalized = False
            self._dp_synced = True
        self._seen_versions = self._param_versions()
        if eng.cfg.batch_k > 0 and self.activation.threshold.data_ptr() != eng.threshold.data_ptr():
            with torch.no_grad():
                eng.threshold.copy_(self.activation.threshold)
            self.activation.threshold = eng.threshold
        return eng

    def _param_versions(self) -> tuple:
        return tuple(p._version for p in (self.W_enc, self.b_enc, self.W_dec, self.b_dec))

    def parameter_checksum(self) -> Tensor:
        """Device float64[4]: sums of the four parameters (data-parallel callers compare them across ranks)."""
        return torch.stack([p.detach().double().sum() for p in (self.W_enc, self.b_enc, self.W_dec, self.b_dec)])

    def check_grad_out(self) -> None:
        """Raise if any backward since the last check received an upstream gradient other than 1 (host sync)."""
        bad, self._bad_grad_out = self._bad_grad_out, None
        if bad is not None and int(bad.item()) != 0:
            raise NotImplementedError("saev_b200: the fused backward computes the gradients of loss.backward() only; "
                                      f"{int(bad.item())} backward call(s) were given a scaled upstream gradient, which it "
                                      "does not apply (scale the learning rate or the loss terms' coefficients instead)")

    def check_batch_topk(self) -> None:
        """BatchTopK: raise if any forward since the last check had a row that filled all of its `capacity` slots -- the
        reference may have kept more entries of such a row, so the step is not certified equal to it (host sync;
        SAEV_B200_BATCHTOPK_ALLOW_TRUNCATION=1 downgrades the error to a counter)."""
        import os

        eng = self.engine
        if eng is None or eng.cfg.batch_k <= 0:
            return
        n = int(eng.btk_truncated.item())
        if n and os.environ.get("SAEV_B200_BATCHTOPK_ALLOW_TRUNCATION", "") != "1":
            eng.btk_truncated.zero_()
            raise RuntimeError(f"saev_b200 BatchTopK: {n} row(s) needed more than the {eng.cfg.top_k} slots per row the "
                               "sparse forward state holds; the selection of those steps differs from the reference's "
                               "(lower BatchTopK.top_k, or set SAEV_B200_BATCHTOPK_ALLOW_TRUNCATION=1 to accept)")

    def weights_changed(self) -> None:
        """Call after writing W_enc outside the optimizer (e.g. datapoint init, train.py:141-185) so the bf16
        operand copy is rebuilt.  `_bind` detects re-pointed storage by itself; in-place writes need this."""
        if self.engine is not None:
            self.engine.sync_weights()
        self._w_dec_normalized = False

    # ---- reference API ---------------------------------------------------------------------
    def _eval_forward(self, x: Tensor):
        eng = self._bind(x.shape[0])
        eng.set_prefixes(None)  # SparseAutoencoder.forward decodes with the single full prefix (modeling.py:331-341)
        # BatchTopKActivation follows the module's own train/eval flag (modeling.py:219): batch selection in train mode
        eng.forward(x.contiguous(), training=False, batch_select=self.training)
        self._ticket += 1
        return Output(self, x, self._ticket)

    def forward(self, x: Tensor) -> Output:
        """modeling.py:331-341 (inference use: no dead tracking, no gradients)."""
        self._require_topk()
        return self._eval_forward(x)

    def encode(self, x: Tensor) -> EncodeOut:
        """modeling.py:343-349."""
        out = self.forward(x)
        return EncodeOut(h_x=out.h_x, f_x=out.f_x)

    def decode(self, f_x: Tensor, *, prefixes: Tensor | None = None) -> Tensor:
        """modeling.py:351-409 for the single-prefix case; dense f_x input is decoded by a (non-hot-path)
        dense product, Matryoshka prefixes are not supported."""
        if prefixes is not None and len(prefixes) > 1:
            raise NotImplementedError("Matryoshka prefix decoding (n_prefixes > 1) has no CUDA path in saev_b200")
        eng = self._bind(f_x.shape[0])
        # x_hat = f_x . W_dec + b_dec  ==  gemm_nt(f_x, W_dec^T) ; W_dec^T [D, S] is materialised once per call
        return eng.gemm_nt(f_x.contiguous(), eng.W_dec.t().contiguous(), eng.b_dec, 6)[:, None, :]

    @torch.no_grad()
    def normalize_w_dec(self):
        """modeling.py:411-417.  A no-op when the fused optimizer already renormalised the rows."""
        if not self.cfg.normalize_w_dec:
            return
        if self.W_dec.device.type != "cuda":
            self.W_dec.data /= torch.norm(self.W_dec.data, dim=1, keepdim=True)  # construction-time (CPU) init only
            return
        if self._w_dec_normalized:
            return
        if self.engine is None or self.W_dec.data_ptr() != self.engine.W_dec.data_ptr():
            # no engine yet (first call of the loop, train.py:334): a plain row normalisation of the parameter; the
            # engine is created by the forward that follows, at the real batch size
            self.W_dec.data /= torch.norm(self.W_dec.data, dim=1, keepdim=True)
            return
        self._bind(max(self._max_batch, 1)).normalize_w_dec()

    @torch.no_grad()
    def remove_parallel_grads(self):
        """modeling.py:419-445.  saev_b200_backward already removed the parallel component."""
        if not self.cfg.remove_parallel_grads or self.W_dec.grad is None:
            return
        if self._grads_fused:
            return
        raise RuntimeError("remove_parallel_grads(): gradients were not produced by the fused backward")

    def _require_topk(self):
        """TopK / BatchTopK (sparse path) and Relu (dense path) have CUDA paths."""
        if _kind(self.cfg.activation) not in ("TopK", "Relu", "BatchTopK"):
            raise NotImplementedError(f"activation {_kind(self.cfg.activation)} has no CUDA path in saev_b200")
        if _kind(self.cfg.activation) == "BatchTopK" and self._dp_world > 1:
            raise NotImplementedError("BatchTopK selects over the whole batch and does not shard across ranks "
                                      "(SURVEY 8e): run it on one GPU")


# ----------------------------------------------------------------------------------------------
# objective
# ----------------------------------------------------------------------------------------------
@dataclasses.dataclass(frozen=True)
class MatryoshkaLoss:
    """objectives.py:59-89; fields are 0-d views into the engine's device loss vector."""

    mse: Tensor
    sparsity: Tensor
    l0: Tensor
    l1: Tensor
    aux: Tensor
    n_dead: Tensor
    _total: Tensor = None
    _sae: tp.Any = None

    @property
    def loss(self) -> Tensor:
        return self._total

    def metrics(self) -> dict[str, object]:
        out = {
            "loss": self.loss.item(), "mse": self.mse.item(), "l0": self.l0.item(), "l1": self.l1.item(),
            "sparsity": self.sparsity.item(), "aux": self.aux.item(), "n_dead": self.n_dead,
        }
        if self._sae is not None:
            self._sae.check_grad_out()
            if self._sae.engine is not None and self._sae.engine.cfg.activation == "topk":
                # rows the tensor-core screen could not certify and the exact fp32 path re-did (cumulative); not a
                # reference key -- it shows up as loss/screen_repaired_rows in saev's log block (train.py:420)
                out["screen_repaired_rows"] = self._sae.engine.screen_stats()["repaired"]
            if self._sae.engine is not None and self._sae.engine.cfg.batch_k > 0:
                self._sae.check_batch_topk()
        return out


class _StepFunction(torch.autograd.Function):
    """Connects saev_b200_forward / saev_b200_backward to `loss.loss.backward()` (train.py:348)."""

    @staticmethod
    def forward(ctx, sae, x, tokens_global, W_dec, b_dec, W_enc, b_enc):
        eng = sae.engine
        if sae._dp_world > 1:
            from . import _lib

            eng.forward(x, training=True, phase=_lib.PHASE_A, tokens_global=tokens_global)
            dist.all_reduce(eng.active_flags(), op=dist.ReduceOp.MAX, group=sae._dp_group)
            eng.forward(x, training=True, phase=_lib.PHASE_B, tokens_global=tokens_global)
            # every rank's scalars are partial sums over the GLOBAL denominator; n_dead is identical on all ranks
            n_dead = eng.losses[5].clone()
            dist.all_reduce(eng.losses, op=dist.ReduceOp.SUM, group=sae._dp_group)
            eng.losses[5] = n_dead
        else:
            eng.forward(x, training=True, tokens_global=tokens_global)
        ctx.sae, ctx.x, ctx.tokens_global = sae, x, tokens_global
        return eng.losses[6].clone()

    @staticmethod
    def backward(ctx, grad_out):
        sae, eng = ctx.sae, ctx.sae.engine
        if sae._check_grad_out:  # no host sync here: counted on the device, raised by the next Loss.metrics()
            sae._bad_grad_out = (grad_out != 1).to(torch.int32) if sae._bad_grad_out is None else \
                sae._bad_grad_out + (grad_out != 1).to(torch.int32)
        eng.backward(ctx.x, tokens_global=ctx.tokens_global)
        if sae._dp_world > 1:
            dist.all_reduce(eng.grads, op=dist.ReduceOp.SUM, group=sae._dp_group)
        views = (("W_dec", eng.gW_dec), ("b_dec", eng.gb_dec), ("W_enc", eng.gW_enc_t.t()), ("b_enc", eng.gb_enc))
        for name, g in views:
            p = getattr(sae, name)
            if p.grad is None:
                p.grad = g  # alias the flat gradient bucket: no copy, the fused optimizer reads it in place
            elif p.grad.data_ptr() != g.data_ptr():
                p.grad.add_(g)
        sae._grads_fused = True
        return None, None, None, None, None, None, None


class MatryoshkaObjective(torch.nn.Module):
    """objectives.py:92-156."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.toks_since_active: Tensor | None = None

    def forward(self, sae: SparseAutoencoder, x: Tensor):
        sae._require_topk()
        if _kind(sae.cfg.activation) == "BatchTopK" and sae.training != self.training:
            raise NotImplementedError("BatchTopK: the SAE and the objective must be in the same train/eval mode")
        x = x.contiguous()
        eng = sae._bind(x.shape[0], self.cfg)
        # objectives.py:125: the cuts are drawn on the host from the global torch generator, every forward
        eng.set_prefixes(sample_prefixes(sae.cfg.d_sae, self.cfg.n_prefixes).tolist() if self.cfg.n_prefixes > 1 else None)
        sae._ticket += 1
        out = Output(sae, x, sae._ticket)
        if self.training:
            if self.toks_since_active is None:  # objectives.py:108-111: lazily created, not in the state_dict
                eng.toks_since_active.zero_()
            self.toks_since_active = eng.toks_since_active
            total = _StepFunction.apply(sae, x, x.shape[0] * sae._dp_world, sae.W_dec, sae.b_dec, sae.W_enc,
                                        sae.b_enc)
        else:
            eng.forward(x, training=False)
            total = eng.losses[6].clone()
        L = eng.losses.clone()
        loss = MatryoshkaLoss(mse=L[0], sparsity=L[2], l0=L[3], l1=L[4], aux=L[1], n_dead=L[5].to(torch.int64),
                              _total=total, _sae=sae)
        return loss, out


@torch.no_grad()
def datapoint_init(saes, dl, *, noise_device: str = "cuda") -> None:
    """The datapoint initialisation of saev's `make_saes` (train.py:121-185) with the arithmetic on the device
    (saev_b200_datapoint_init): pull max(d_sae, 65 536) rows from the loader, then per SAE
        W_enc[:, j] = W_dec[j] = normalise(blend * (x_perm[idx[j]] - mean) + (1 - blend) * kaiming[idx[j]]).
    The host draws (`randperm(n_samples)`, then one `randperm(d_sae)` per SAE) come from torch's global CPU generator in
    the reference's order, so a seeded run picks the same rows; `kaiming` is drawn on `noise_device` (the reference
    draws it wherever the loader's batches live: "cpu" with its own loader, "cuda" with this package's).
    `saes`: saev_b200.nn.SparseAutoencoder modules already on the GPU."""
    saes = list(saes)
    if all(sae.cfg.reinit_blend == 0 for sae in saes):
        return
    assert saes, "Need at least one SAE to initialize."
    d_sae = saes[0].cfg.d_sae
    assert all(d_sae == sae.cfg.d_sae for sae in saes), "All SAEs must have same .d_sae"
    n_samples = d_sae
    if hasattr(dl, "n_samples"):
        assert dl.n_samples >= n_samples, f"Need {n_samples} samples for datapoint init; dataloader has {dl.n_samples}."
    n_samples = max(n_samples, 65_536)  # train.py:142-145
    n_samples = min(n_samples, dl.n_samples)
    dev = saes[0].W_dec.device
    batches, n_seen = [], 0
    for batch in dl:
        act = batch["act"]
        batches.append(act.to(dev))
        n_seen += len(act)
        if n_seen >= n_samples:
            break
    assert n_seen >= n_samples, f"Datapoint init requested {n_samples} samples but saw {n_seen}."
    acts = torch.cat(batches, dim=0)[:n_samples].contiguous()
    del batches
    perm = torch.randperm(n_samples)  # train.py:161 `acts = acts[torch.randperm(n_samples)]`: folded into the row index
    D = acts.shape[1]
    noise = torch.empty(d_sae, D, device=noise_device)
    torch.nn.init.kaiming_uniform_(noise)  # train.py:165-166
    noise = noise.to(dev)
    for sae in saes:
        blend = sae.cfg.reinit_blend
        assert 0.0 <= blend <= 1.0, f"reinit_blend must be in [0, 1], got {blend}."
        idx = torch.randperm(d_sae)  # train.py:169
        eng = sae._bind(max(sae._max_batch, 1))
        eng.datapoint_init(acts, perm[idx].to(dev), noise, idx.to(dev), blend,
                           tie_transpose=sae.cfg.reinit_enc_dec_tranpose)
        sae._w_dec_normalized = bool(sae.cfg.normalize_w_dec)


def get_objective(cfg) -> MatryoshkaObjective:
    """objectives.py:204-220."""
    if _kind(cfg) == "Matryoshka":
        return MatryoshkaObjective(cfg)
    raise TypeError(f"unknown objective config {cfg!r}")


# ----------------------------------------------------------------------------------------------
# checkpoints (modeling.py:548-658, schema 5): one JSON header line + torch.save(state_dict)
# ----------------------------------------------------------------------------------------------
def _serialize_dataclass(obj) -> dict:
    """modeling.py:466-472: {"cls": class name, "params": {field: value | nested payload}}."""
    params = {}
    for f in dataclasses.fields(obj):
        v = getattr(obj, f.name)
        params[f.name] = _serialize_dataclass(v) if dataclasses.is_dataclass(v) else v
    return {"cls": type(obj).__name__, "params": params}


_BY_CLS = {c.__name__: c for c in (NoSparsity, L1Sparsity, NoAux, AuxK, Relu, TopK, BatchTopK)}


def _deserialize_dataclass(payload: dict):
    """modeling.py:486-505 (schema 5: no legacy nesting)."""
    cls = _BY_CLS.get(payload["cls"])
    if cls is None:
        raise ValueError(f"Unknown activation class '{payload['cls']}' in payload.")
    kw = {}
    for k, v in payload["params"].items():
        k = "key" if k == "kind" else k
        kw[k] = _deserialize_dataclass(v) if isinstance(v, dict) and "cls" in v and "params" in v else v
    return cls(**kw)


def dump(fpath, sae: SparseAutoencoder) -> None:
    """modeling.py:548-574: one JSON header line {schema, cfg, commit, lib} + torch.save(state_dict).  The header is the
    reference's schema 5, so `saev.nn.load` reads these files and `load` below reads saev's."""
    cfg = sae.cfg
    cfg_dict = {f.name: getattr(cfg, f.name) for f in dataclasses.fields(cfg)}
    cfg_dict["activation"] = _serialize_dataclass(cfg.activation)
    header = {"schema": SCHEMA_VERSION, "cfg": cfg_dict, "commit": "unknown", "lib": f"saev_b200-{__version__}"}
    fpath = pathlib.Path(fpath)
    fpath.parent.mkdir(exist_ok=True, parents=True)
    state = {k: v.detach().cpu().contiguous() for k, v in sae.state_dict().items()}
    with open(fpath, "wb") as fd:
        fd.write(json.dumps(header).encode() + b"\n")
        torch.save(state, fd)


def load(fpath, *, device="cpu") -> SparseAutoencoder:
    """modeling.py:577-658 for schema-5 files (what saev.nn.dump and `dump` above write)."""
    with open(fpath, "rb") as fd:
        header = json.loads(fd.readline())
        buffer = io.BytesIO(fd.read())
    if header.get("schema") != SCHEMA_VERSION:
        raise ValueError(f"saev_b200.nn.load reads schema {SCHEMA_VERSION} checkpoints; got {header.get('schema')!r} "
                         "(convert older files with saev.nn.load + saev.nn.dump)")
    cfg_dict = dict(header["cfg"])
    cfg_dict["activation"] = _deserialize_dataclass(cfg_dict["activation"])
    for legacy in ("n_reinit_samples", "seed"):  # modeling.py:449-453
        cfg_dict.pop(legacy, None)
    known = {f.name for f in dataclasses.fields(SparseAutoencoderConfig)}
    cfg = SparseAutoencoderConfig(**{k: v for k, v in cfg_dict.items() if k in known})
    model = SparseAutoencoder(cfg)
    model.load_state_dict(torch.load(buffer, weights_only=True, map_location="cpu"))
    return model.to(device)


def kaiming_bound(d_model: int) -> float:
    return math.sqrt(6.0 / d_model)
