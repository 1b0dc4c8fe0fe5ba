"""Device-side state + launch sequencing for one SAE on one B200.

`Engine` owns the flat parameter / gradient / Adam-moment buffers (order
[W_enc_t (S*D), b_enc (S), W_dec (S*D), b_dec (D)], see include/saev_b200.h), the kernel workspace and
the library handle, and exposes the steps of saev's training loop body
(/root/reference/src/saev/framework/train.py:332-460) as methods that only enqueue CUDA work:

    normalize_w_dec -> forward -> backward -> [all-reduce] -> grad_sumsq -> adam_step

PyTorch is used for memory, streams and torch.distributed only; all arithmetic happens in
libsaev_b200.so.  There is no CPU path: constructing an Engine without a CUDA device raises.
"""

from __future__ import annotations

import ctypes as C
import dataclasses

import torch

from . import _lib


@dataclasses.dataclass(frozen=True)
class EngineConfig:
    d_model: int
    d_sae: int
    top_k: int = 32
    activation: str = "topk"  # "topk" | "relu"
    aux: bool = True  # AuxK vs NoAux
    k_aux: int = 512
    aux_alpha: float = 1.0 / 32
    l1_coeff: float = 0.0
    dead_threshold_tokens: int = 10_000_000
    remove_parallel_grads: bool = True
    normalize_w_dec: bool = True
    max_batch: int = 16384
    aux_cols_cap: int = 0
    max_prefixes: int = 1  # Matryoshka.n_prefixes the workspace is sized for
    # BatchTopK (modeling.py:183-244): batch_k = BatchTopK.top_k (average actives per sample; 0 = plain TopK).  `top_k`
    # is then the per-row CAPACITY of the sparse forward state (<= 128): the batch-wide selection is exact as long as no
    # row owns more than `top_k` of the batch's winners (Engine.batch_topk_stats() counts the rows that might).
    batch_k: int = 0
    batch_momentum: float = 0.1


LOSS_KEYS = ("mse", "aux", "sparsity", "l0", "l1", "n_dead", "loss")


class Engine:
    def __init__(self, cfg: EngineConfig, device: torch.device | str = "cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("saev_b200.Engine needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("saev_b200.Engine only runs on CUDA devices")
        self.lib = _lib.load()
        S, D = cfg.d_sae, cfg.d_model
        self.S, self.D, self.K = S, D, cfg.top_k
        self.n_params = 2 * S * D + S + D
        with torch.cuda.device(self.device):
            c = _lib.Cfg(
                d_model=D,
                d_sae=S,
                act_kind=_lib.ACT_TOPK if cfg.activation == "topk" else _lib.ACT_RELU,
                top_k=cfg.top_k,
                aux_kind=_lib.AUX_AUXK if cfg.aux else _lib.AUX_NONE,
                k_aux=cfg.k_aux,
                aux_alpha=cfg.aux_alpha,
                l1_coeff=cfg.l1_coeff,
                dead_threshold_tokens=cfg.dead_threshold_tokens,
                remove_parallel_grads=int(cfg.remove_parallel_grads),
                max_batch=cfg.max_batch,
                aux_cols_cap=cfg.aux_cols_cap,
                max_prefixes=cfg.max_prefixes,
            )
            h = C.c_void_p()
            _lib.check(self.lib.saev_b200_create(C.byref(c), C.byref(h)))
            self.h = h
            dev = self.device
            self.params = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
            self.grads = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
            self.m = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
            self.v = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
            self.workspace = torch.zeros(self.lib.saev_b200_workspace_bytes(h), dtype=torch.uint8, device=dev)
            B, K = cfg.max_batch, max(cfg.top_k, 1)
            self.topk_idx = torch.empty(B, K, dtype=torch.int32, device=dev)
            self.topk_val = torch.empty(B, K, dtype=torch.float32, device=dev)
            self.resid = torch.empty(B, D, dtype=torch.float32, device=dev)
            self.losses = torch.zeros(8, dtype=torch.float32, device=dev)
            self.sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
            self.gnorm = torch.zeros(1, dtype=torch.float32, device=dev)
            self.toks_since_active = torch.zeros(S, dtype=torch.int64, device=dev)
            # BatchTopK: the EMA inference threshold (`activation.threshold` buffer of the reference, modeling.py:213)
            # and the selection counters of the last forward
            self.threshold = torch.zeros((), dtype=torch.float32, device=dev)
            self.btk_stats = torch.zeros(4, dtype=torch.int32, device=dev)
            self.btk_truncated = torch.zeros((), dtype=torch.int64, device=dev)  # cumulative btk_stats[1]
        self.W_enc_t, self.b_enc, self.W_dec, self.b_dec = self._views(self.params)
        self.gW_enc_t, self.gb_enc, self.gW_dec, self.gb_dec = self._views(self.grads)
        self.step_count = 0
        self._last_B = 0
        # decoder half of an Adam step that has been deferred to run beside the next forward's screen (train_step)
        self._dec_pending: dict | None = None
        self._side_stream: torch.cuda.Stream | None = None
        self._ev_grads = torch.cuda.Event()
        self._ev_dec = torch.cuda.Event()

    def _views(self, flat):
        S, D = self.S, self.D
        o1, o2, o3 = S * D, S * D + S, 2 * S * D + S
        return flat[:o1].view(S, D), flat[o1:o2], flat[o2:o3].view(S, D), flat[o3:]

    def __del__(self):
        h = getattr(self, "h", None)
        if h is not None and getattr(self, "lib", None) is not None:
            self.lib.saev_b200_destroy(h)
            self.h = None

    # ---- helpers -------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _ck(self, rc):
        _lib.check(rc, self.h)

    def _ws_tensor(self, fn, dtype, n):
        """Typed 1-D view of a region inside the workspace (zero-copy)."""
        p = fn(self.h, self.workspace.data_ptr())
        off = p - self.workspace.data_ptr()
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        return self.workspace[off : off + nbytes].view(dtype)

    def active_flags(self) -> torch.Tensor:
        return self._ws_tensor(self.lib.saev_b200_active_flags, torch.int32, self.S)

    def unsafe_rows(self) -> int:
        """Rows the tensor-core screen could not certify since the last sync_weights() (each was re-done by the exact
        fp32 path in the same forward, see `screen_stats()["repaired"]`).  Host sync."""
        return int(self._ws_tensor(self.lib.saev_b200_unsafe_rows, torch.int32, 1).item())

    def screen_stats(self) -> dict:
        """Cumulative diagnostics of the top-k screen since the last sync_weights(): rows that could not be
        certified, rows re-done by the exact path (equal once the forward has finished), of those the rows whose
        threshold guess was too high (`guess_failed`, expected: a few per 10^4 rows) and the rows whose observed screen
        error exceeded the bound (`bound_violations`, must be 0), candidates re-scored in fp32, list entries merged."""
        t = self._ws_tensor(self.lib.saev_b200_unsafe_rows, torch.int32, 12).tolist()
        return {"unsafe_rows": t[0], "repaired": t[7], "unrepaired": t[0] - t[7], "bound_violations": t[8],
                "guess_failed": t[11], "rescored": t[2] & 0xFFFFFFFF, "merged": t[9] & 0xFFFFFFFF}

    # ---- parameters ------------------------------------------------------------------------
    @torch.no_grad()
    def load_params(self, W_enc, b_enc, W_dec, b_dec) -> None:
        """Copy saev-layout parameters in (W_enc is [d_model, d_sae] as in saev; stored transposed)."""
        self._dec_pending = None
        self.W_enc_t.copy_(W_enc.to(self.device, torch.float32).t())
        self.b_enc.copy_(b_enc.to(self.device, torch.float32))
        self.W_dec.copy_(W_dec.to(self.device, torch.float32))
        self.b_dec.copy_(b_dec.to(self.device, torch.float32))
        self.sync_weights()

    @torch.no_grad()
    def init_params(self, seed: int | None = None) -> None:
        """saev's initialisation (modeling.py:306-329): W_dec = kaiming_uniform_([S, D]) (bound sqrt(6/D)),
        row-normalised; W_enc = W_dec.T; biases zero.  saev leaves the global RNG unseeded; `seed` makes
        it reproducible."""
        self._dec_pending = None
        gen = torch.Generator(device=self.device)
        if seed is not None:
            gen.manual_seed(seed)
        bound = (6.0 / self.D) ** 0.5
        self.W_dec.copy_((torch.rand(self.S, self.D, generator=gen, device=self.device) * 2 - 1) * bound)
        self.b_dec.zero_()
        self.b_enc.zero_()
        self._ck(self.lib.saev_b200_normalize_w_dec(self.h, self.W_dec.data_ptr(), self._stream()))
        self.W_enc_t.copy_(self.W_dec)
        self.m.zero_()
        self.v.zero_()
        self.toks_since_active.zero_()
        self.step_count = 0
        self.sync_weights()

    def sync_weights(self) -> None:
        with torch.cuda.device(self.device):
            self._ck(
                self.lib.saev_b200_sync_weights(self.h, self.W_enc_t.data_ptr(), self.b_enc.data_ptr(),
                                                self.workspace.data_ptr(), self._stream())
            )

    @torch.no_grad()
    def datapoint_init(self, acts: torch.Tensor, src_row: torch.Tensor, noise: torch.Tensor,
                       noise_row: torch.Tensor | None, blend: float, *, tie_transpose: bool = True) -> torch.Tensor:
        """saev's datapoint initialisation (train.py:141-185) on the device; see saev_b200_datapoint_init.
        Returns the column mean of `acts`."""
        assert acts.is_cuda and acts.dtype == torch.float32 and acts.is_contiguous() and acts.shape[1] == self.D
        assert src_row.is_cuda and src_row.dtype == torch.int64 and src_row.numel() == self.S
        assert noise.is_cuda and noise.dtype == torch.float32 and noise.is_contiguous() and noise.shape[1] == self.D
        if noise_row is not None:
            assert noise_row.is_cuda and noise_row.dtype == torch.int64 and noise_row.numel() == self.S
        mean = torch.empty(self.D, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._ck(self.lib.saev_b200_datapoint_init(
                self.h, acts.data_ptr(), acts.shape[0], src_row.data_ptr(), noise.data_ptr(), _lib.ptr(noise_row),
                float(blend), int(tie_transpose), int(self.cfg.normalize_w_dec), mean.data_ptr(), self.W_enc_t.data_ptr(),
                self.b_enc.data_ptr(), self.W_dec.data_ptr(), self.workspace.data_ptr(), self._stream()))
        return mean

    def normalize_w_dec(self) -> None:
        self.flush()
        if not self.cfg.normalize_w_dec:
            return
        with torch.cuda.device(self.device):
            self._ck(self.lib.saev_b200_normalize_w_dec(self.h, self.W_dec.data_ptr(), self._stream()))

    # ---- step pieces -------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, *, training: bool = True, phase: int = _lib.PHASE_ALL, tokens_global: int = 0,
                batch_select: bool | None = None):
        """`batch_select` (BatchTopK only): True = batch-wide top-(k B) selection + threshold EMA (the activation's train
        mode), False = JumpReLU with the stored threshold (its eval mode); default: follows `training`."""
        self.flush()  # (a decoder update deferred by train_step(overlap_decoder_update=True) runs first)
        if self.cfg.batch_k > 0:
            return self._forward_batch_topk(x, training=training, phase=phase, tokens_global=tokens_global,
                                            batch_select=training if batch_select is None else batch_select)
        return self._forward(x, training=training, phase=phase, tokens_global=tokens_global)

    def _forward_batch_topk(self, x: torch.Tensor, *, training: bool, phase: int, tokens_global: int, batch_select: bool):
        """BatchTopK forward: screen + exact re-score leave each row's `top_k` (= capacity) largest pre-activations, then
        saev_b200_batch_topk keeps the batch_k * B largest of the batch (training) or applies the JumpReLU threshold
        (eval), then decode / losses run on the thinned lists."""
        if phase != _lib.PHASE_ALL or (tokens_global not in (0, x.shape[0])):
            raise NotImplementedError("BatchTopK selects over the whole batch: single rank, unsplit forward only "
                                      "(the global top-k does not shard, SURVEY 8e)")
        self._forward(x, training=training, phase=_lib.PHASE_A_SCREEN | _lib.PHASE_A_RESCORE)
        with torch.cuda.device(self.device):
            self._ck(self.lib.saev_b200_batch_topk(
                self.h, x.shape[0], self.cfg.batch_k, int(batch_select), self.threshold.data_ptr(),
                float(self.cfg.batch_momentum), self.topk_idx.data_ptr(), self.topk_val.data_ptr(),
                self.btk_stats.data_ptr(), self.workspace.data_ptr(), self._stream()))
        self.btk_truncated += self.btk_stats[1]
        return self._forward(x, training=training, phase=_lib.PHASE_A_DECODE | _lib.PHASE_B)

    def batch_topk_stats(self) -> dict:
        """Counters of the last BatchTopK forward (host sync): entries kept, rows whose capacity may have truncated the
        selection (must be 0 for the result to be certified equal to the reference's), entries tied at the cut value;
        plus the cumulative truncated-row count."""
        t = self.btk_stats.tolist()
        return {"kept": t[0], "truncated_rows": t[1], "ties": t[2], "truncated_rows_total": int(self.btk_truncated.item())}

    def _forward(self, x: torch.Tensor, *, training: bool = True, phase: int = _lib.PHASE_ALL, tokens_global: int = 0):
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == self.D
        B = x.shape[0]
        self._last_B = B
        toks = self.toks_since_active.data_ptr() if training else None
        with torch.cuda.device(self.device):
            self._ck(
                self.lib.saev_b200_forward(
                    self.h, phase, x.data_ptr(), B, tokens_global or B,
                    self.W_enc_t.data_ptr(), self.b_enc.data_ptr(), self.W_dec.data_ptr(), self.b_dec.data_ptr(),
                    toks, int(training), self.topk_idx.data_ptr(), self.topk_val.data_ptr(), self.resid.data_ptr(),
                    self.losses.data_ptr(), self.workspace.data_ptr(), self._stream(),
                )
            )
        return self.losses

    def backward(self, x: torch.Tensor, *, tokens_global: int = 0) -> None:
        B = x.shape[0]
        with torch.cuda.device(self.device):
            self._ck(
                self.lib.saev_b200_backward(
                    self.h, x.data_ptr(), B, tokens_global or B,
                    self.W_enc_t.data_ptr(), self.b_enc.data_ptr(), self.W_dec.data_ptr(), self.b_dec.data_ptr(),
                    self.topk_idx.data_ptr(), self.topk_val.data_ptr(), self.resid.data_ptr(),
                    self.gW_enc_t.data_ptr(), self.gb_enc.data_ptr(), self.gW_dec.data_ptr(), self.gb_dec.data_ptr(),
                    self.workspace.data_ptr(), self._stream(),
                )
            )

    def backward_stage(self, x: torch.Tensor, stage: int, row_begin: int = 0, row_end: int = 0, *,
                       tokens_global: int = 0) -> None:
        """Staged backward (see saev_b200_backward_stage): stage 0 once, then stage 1 over row ranges that together
        cover [0, d_sae)."""
        B = x.shape[0]
        with torch.cuda.device(self.device):
            self._ck(
                self.lib.saev_b200_backward_stage(
                    self.h, stage, row_begin, row_end, x.data_ptr(), B, tokens_global or B,
                    self.W_enc_t.data_ptr(), self.b_enc.data_ptr(), self.W_dec.data_ptr(), self.b_dec.data_ptr(),
                    self.toks_since_active.data_ptr(), self.topk_idx.data_ptr(), self.topk_val.data_ptr(),
                    self.resid.data_ptr(), self.gW_enc_t.data_ptr(), self.gb_enc.data_ptr(), self.gW_dec.data_ptr(),
                    self.gb_dec.data_ptr(), self.workspace.data_ptr(), self._stream(),
                )
            )

    def grad_sumsq(self, *, local: bool = False) -> torch.Tensor:
        """||g||^2 of the gradient bucket.  `local=True`: the bucket still holds exactly what backward() wrote (single
        rank, no all-reduce in between) -> use the per-atom partials of the backward kernels (TopK path)."""
        if local and self.cfg.activation == "topk":
            with torch.cuda.device(self.device):
                self._ck(self.lib.saev_b200_grad_sumsq_local(self.h, self.gb_dec.data_ptr(), self.sumsq.data_ptr(),
                                                             self.workspace.data_ptr(), self._stream()))
            return self.sumsq
        with torch.cuda.device(self.device):
            self._ck(
                self.lib.saev_b200_grad_sumsq(
                    self.h, self.grads.data_ptr(), self.n_params, self.sumsq.data_ptr(), self.workspace.data_ptr(),
                    self._stream(),
                )
            )
        return self.sumsq

    # ---- sharded optimizer (data parallel) -------------------------------------------------------
    def set_optimizer_shard(self, row_begin: int, row_end: int) -> None:
        """adam_step() then only updates dictionary rows [row_begin, row_end) (+ both bias vectors); see
        include/saev_b200.h.  (0, 0) restores the full update."""
        self._ck(self.lib.saev_b200_set_optimizer_shard(self.h, row_begin, row_end))

    def set_reserved_sms(self, n_sms: int) -> None:
        """Keep `n_sms` SMs out of the top-k screen's persistent grid (room for concurrent NCCL kernels)."""
        self._ck(self.lib.saev_b200_set_reserved_sms(self.h, n_sms))

    def grad_sumsq_ranges(self, ranges) -> torch.Tensor:
        """||g||^2 over up to four [begin, end) element ranges of the flat gradient bucket."""
        n = len(ranges)
        b = (C.c_int64 * n)(*[r[0] for r in ranges])
        e = (C.c_int64 * n)(*[r[1] for r in ranges])
        with torch.cuda.device(self.device):
            self._ck(self.lib.saev_b200_grad_sumsq_ranges(self.h, self.grads.data_ptr(), n, b, e, self.sumsq.data_ptr(),
                                                          self.workspace.data_ptr(), self._stream()))
        return self.sumsq

    def shadow_weights(self) -> torch.Tensor:
        """The fp16 tensor-core operand copy of W_enc_t, [d_sae, d_model], inside the workspace."""
        return self._ws_tensor(self.lib.saev_b200_shadow_weights, torch.float16, self.S * self.D).view(self.S, self.D)

    def wnorm_rows(self) -> torch.Tensor:
        """Device float32[d_sae]: ||W_enc_t[j]||_2, the per-column input of the screen's error bound."""
        return self._ws_tensor(self.lib.saev_b200_wnorm_rows, torch.float32, self.S)

    def wnorm_scalar(self) -> torch.Tensor:
        """Device float32[3]: max_j ||W_enc_t[j]||^2, max_j |b_enc[j]|, max_j relative fp16 residual -- the
        dictionary-wide inputs of the screen's error bound (MAX-all-reduced by a sharded optimizer)."""
        return self._ws_tensor(self.lib.saev_b200_wnorm_scalar, torch.float32, 3)

    def adam_step(self, lr: float, *, max_norm: float = 1.0, grad_scale: float = 1.0, betas=(0.9, 0.999),
                  eps: float = 1e-8, renorm_w_dec: bool = False, parts: int = _lib.ADAM_ALL, step: int | None = None) -> None:
        """`parts`: encoder half, decoder half or both (see saev_b200_adam_step); `step`: the 1-based step count of
        this update (default: advance the engine's counter -- only the first half of a split update does that)."""
        if step is None:
            self.step_count += 1
            step = self.step_count
        with torch.cuda.device(self.device):
            self._ck(
                self.lib.saev_b200_adam_step(
                    self.h, self.W_enc_t.data_ptr(), self.b_enc.data_ptr(), self.W_dec.data_ptr(), self.b_dec.data_ptr(),
                    self.grads.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), lr, betas[0], betas[1], eps,
                    step, max_norm, grad_scale, self.sumsq.data_ptr(), int(renorm_w_dec),
                    self.gnorm.data_ptr(), parts, self.workspace.data_ptr(), self._stream(),
                )
            )

    def train_step(self, x: torch.Tensor, lr: float, *, max_norm: float = 1.0, fused_renorm: bool = False,
                   pre_normalized: bool = False, overlap_decoder_update: bool = False) -> torch.Tensor:
        """One iteration of train.py:332-460 on a single GPU (no logging block).

        `overlap_decoder_update` (needs the fused renorm): the decoder half of the Adam step -- W_dec, its moments and
        the row renorm, 28 B/param of pure HBM traffic -- is not run at the end of this step but on a side stream
        BESIDE the tensor-bound screen of the NEXT call (which reads only the encoder side); the re-score / decode of
        that call wait for it.  Same arithmetic, same results; call `flush()` before reading W_dec / b_dec from
        outside (checkpoint, evaluation)."""
        renorm = fused_renorm and self.cfg.normalize_w_dec
        if not overlap_decoder_update or (self.cfg.normalize_w_dec and not renorm) or self.cfg.activation != "topk" \
                or self.cfg.batch_k > 0:
            self.flush()
            if not pre_normalized:
                self.normalize_w_dec()
            self.forward(x, training=True)
            self.backward(x)
            self.grad_sumsq(local=True)
            self.adam_step(lr, max_norm=max_norm, renorm_w_dec=renorm)
            return self.losses
        if not pre_normalized and self._dec_pending is None:
            self.normalize_w_dec()
        self._forward(x, training=True, phase=_lib.PHASE_A_SCREEN)
        self._launch_pending_decoder_update()  # enqueued AFTER the screen kernel: its CTAs get the SMs first
        self._forward(x, training=True, phase=_lib.PHASE_A_REST | _lib.PHASE_B)
        self.backward(x)
        self.grad_sumsq(local=True)
        self.adam_step(lr, max_norm=max_norm, parts=_lib.ADAM_ENCODER)
        self._ev_grads.record(torch.cuda.current_stream(self.device))
        self._dec_pending = dict(lr=lr, max_norm=max_norm, renorm=renorm, step=self.step_count)
        return self.losses

    def _launch_pending_decoder_update(self) -> None:
        p, self._dec_pending = self._dec_pending, None
        if p is None:
            return
        main = torch.cuda.current_stream(self.device)
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=self.device)
        side = self._side_stream
        side.wait_event(self._ev_grads)
        with torch.cuda.stream(side):
            self.adam_step(p["lr"], max_norm=p["max_norm"], renorm_w_dec=p["renorm"], parts=_lib.ADAM_DECODER, step=p["step"])
            self._ev_dec.record(side)
        main.wait_event(self._ev_dec)

    def flush(self) -> None:
        """Run a deferred decoder update now (on the current stream)."""
        p, self._dec_pending = self._dec_pending, None
        if p is not None:
            self.adam_step(p["lr"], max_norm=p["max_norm"], renorm_w_dec=p["renorm"], parts=_lib.ADAM_DECODER, step=p["step"])

    # ---- lazy dense views ------------------------------------------------------------------
    def dense_f_x(self, B: int | None = None) -> torch.Tensor:
        B = B or self._last_B
        out = torch.empty(B, self.S, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._ck(
                self.lib.saev_b200_dense_f(
                    self.h, self.topk_idx.data_ptr(), self.topk_val.data_ptr(), B, out.data_ptr(),
                    self.workspace.data_ptr(), self._stream()
                )
            )
        return out

    def set_prefixes(self, prefixes=None) -> None:
        """Matryoshka cut points for the following forward/backward calls (sorted, last == d_sae); None = single prefix."""
        if prefixes is None or len(prefixes) <= 1:
            self._ck(self.lib.saev_b200_set_prefixes(self.h, None, 1))
            self._n_prefixes = 1
            return
        arr = (C.c_int32 * len(prefixes))(*[int(c) for c in prefixes])
        self._ck(self.lib.saev_b200_set_prefixes(self.h, arr, len(prefixes)))
        self._n_prefixes = len(prefixes)

    def x_hats(self, x: torch.Tensor) -> torch.Tensor:
        """[B, n_prefixes, d_model] reconstructions of the last forward (modeling.py:406)."""
        B, P = x.shape[0], getattr(self, "_n_prefixes", 1)
        out = torch.empty(B, P, self.D, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._ck(self.lib.saev_b200_x_hats(self.h, self.resid.data_ptr(), x.data_ptr(), B, out.data_ptr(),
                                               self.workspace.data_ptr(), self._stream()))
        return out

    def x_hat(self, x: torch.Tensor) -> torch.Tensor:
        B = x.shape[0]
        out = torch.empty(B, self.D, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._ck(self.lib.saev_b200_x_hat(self.h, self.resid.data_ptr(), x.data_ptr(), B, out.data_ptr(), self._stream()))
        return out

    def loss_dict(self) -> dict:
        self.flush()
        vals = self.losses.tolist()  # host sync, like Loss.metrics() in saev (objectives.py:80-89)
        return dict(zip(LOSS_KEYS, vals))

    # ---- per-stage device timing ------------------------------------------------------------
    def profile_enable(self, on: bool = True) -> None:
        self._ck(self.lib.saev_b200_profile_enable(self.h, int(on)))

    def profile_read(self) -> dict:
        """{stage: (total_ms, n_intervals)} since the last read; synchronises on the recorded events."""
        n = len(_lib.STAGES)
        ms = (C.c_float * n)()
        cnt = (C.c_int32 * n)()
        self._ck(self.lib.saev_b200_profile_read(self.h, ms, cnt))
        return {name: (float(ms[i]), int(cnt[i])) for i, name in enumerate(_lib.STAGES)}

    # ---- log-block metrics (saev train.py:365-442) -------------------------------------------
    def dictionary_coherence(self, W_dec: torch.Tensor | None = None) -> torch.Tensor:
        """max_{i<j} |cos(w_i, w_j)| over the decoder rows (train.py:415-421) without the [S, S] Gram matrix.
        Returns a device float32[4]: coherence, tensor-core screen maximum, i, j.  No host sync."""
        W = self.W_dec if W_dec is None else W_dec
        assert W.is_cuda and W.dtype == torch.float32 and W.is_contiguous() and tuple(W.shape) == (self.S, self.D)
        n = int(self.lib.saev_b200_coherence_scratch_bytes(self.h))
        scratch = torch.empty(n, dtype=torch.uint8, device=self.device)
        out = torch.empty(4, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._ck(self.lib.saev_b200_dictionary_coherence(self.h, W.data_ptr(), scratch.data_ptr(), n, out.data_ptr(),
                                                             self._stream()))
        return out

    LOG_KEYS = ("explained_variance", "dead_unit_pct", "dictionary_coherence", "avg_decoder_row_norm", "sse_sae",
                "sse_baseline", "normalized_mse")

    def log_metrics(self, x: torch.Tensor) -> torch.Tensor:
        """The per-SAE `metrics/*` of saev's log block (train.py:380-423) for the batch of the last forward, as a
        device float64[8] in LOG_KEYS order (+ the coherence screen maximum).  No host sync; nothing of size [B, S]
        or [S, S] is formed."""
        self.flush()
        B = x.shape[0]
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == self.D
        n = int(self.lib.saev_b200_log_scratch_bytes(self.h))
        scratch = torch.empty(n, dtype=torch.uint8, device=self.device)
        out = torch.empty(8, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            self._ck(self.lib.saev_b200_log_metrics(self.h, x.data_ptr(), self.resid.data_ptr(), B, self.W_dec.data_ptr(),
                                                    self.workspace.data_ptr(), scratch.data_ptr(), n, out.data_ptr(),
                                                    self._stream()))
        return out

    def log_metrics_dict(self, x: torch.Tensor) -> dict:
        return dict(zip(self.LOG_KEYS, self.log_metrics(x).tolist()))

    # ---- evaluate() accumulators (saev train.py:546-566) -------------------------------------
    def new_eval_state(self) -> dict:
        """Zeroed device accumulators for `eval_accumulate`: acc float64[8 + D], n_fired / values float32[S]."""
        return dict(acc=torch.zeros(8 + self.D, dtype=torch.float64, device=self.device),
                    n_fired=torch.zeros(self.S, dtype=torch.float32, device=self.device),
                    values=torch.zeros(self.S, dtype=torch.float32, device=self.device))

    def eval_accumulate(self, x: torch.Tensor, state: dict) -> None:
        """Fold the batch of the last forward into `state` (no host sync, no dense f_x)."""
        B = x.shape[0]
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == self.D
        with torch.cuda.device(self.device):
            self._ck(self.lib.saev_b200_eval_accumulate(
                self.h, x.data_ptr(), self.resid.data_ptr(), B, self.topk_idx.data_ptr(), self.topk_val.data_ptr(),
                self.losses.data_ptr(), state["acc"].data_ptr(), state["n_fired"].data_ptr(), state["values"].data_ptr(),
                self.workspace.data_ptr(), self._stream()))

    # ---- test hooks ------------------------------------------------------------------------
    def aux_selection(self, B: int | None = None) -> torch.Tensor:
        """int64 [B, k_use]: the atoms AuxK selected per row in the last training forward (ascending atom order), for
        parity tests that have to tell fp32-level ties from errors.  Host sync (reads n_dead)."""
        B = B or self._last_B
        mask_p, ld, dl_p, nd_p = C.c_void_p(), C.c_int64(), C.c_void_p(), C.c_void_p()
        self._ck(self.lib.saev_b200_aux_selection(self.h, self.workspace.data_ptr(), C.byref(mask_p), C.byref(ld),
                                                  C.byref(dl_p), C.byref(nd_p)))
        base = self.workspace.data_ptr()

        def view(p, nbytes, dtype):
            return self.workspace[p - base : p - base + nbytes].view(dtype)

        n_dead = int(view(nd_p.value, 4, torch.int32).item())
        dead_list = view(dl_p.value, 4 * n_dead, torch.int32).long()
        mask = view(mask_p.value, B * ld.value, torch.uint8).view(B, ld.value)[:, :n_dead] != 0
        k_use = int(mask[0].sum().item()) if B > 0 else 0
        assert bool((mask.sum(1) == k_use).all()), "AuxK selected a different number of latents in different rows"
        return dead_list[mask.nonzero()[:, 1].view(B, k_use)]

    def gemm_nt(self, A: torch.Tensor, Bt: torch.Tensor, bias: torch.Tensor | None, nterms: int) -> torch.Tensor:
        M, K = A.shape
        N = Bt.shape[0]
        out = torch.empty(M, N, dtype=torch.float32, device=self.device)
        scratch = torch.empty(3 * (M + N) * K, dtype=torch.bfloat16, device=self.device)
        with torch.cuda.device(self.device):
            self._ck(
                self.lib.saev_b200_gemm_nt(
                    self.h, A.data_ptr(), Bt.data_ptr(), _lib.ptr(bias), M, N, K, nterms, out.data_ptr(),
                    scratch.data_ptr(), self._stream(),
                )
            )
        return out
