"""Data-parallel training step: one process per GPU, full parameter replica per rank, batch rows sharded across
ranks, Adam state sharded by dictionary rows.

saev itself has no multi-GPU path; the semantics implemented here are "N ranks x B rows behave
exactly like one rank x (N*B) rows": every kernel divides by the GLOBAL batch (`tokens_global`), so per-rank
gradients / loss partials simply add up, `remove_parallel_grads` is linear and is applied before the reduction, and
the clip norm is taken on the reduced gradient.

Exchange per step (NCCL over NVLink):
  * 4*d_sae bytes MAX all-reduce of the activity flags between forward phases A and B, so that every rank keeps an
    identical dead-latent tracker;
  * the gradient bucket, once, in one of two ways:
      - default: all-reduce (sum), issued in `n_chunks` row chunks that OVERLAP the weight-gradient kernel: the
        staged backward (saev_b200_backward_stage) finishes rows [r0, r1) of both weight gradients, their all-reduce
        starts on NCCL's stream while the kernel for the next rows runs; only the last chunk and the two bias
        vectors are exposed.  Every rank then runs the same norm / clip / Adam pass.
      - `sharded=True` (d_sae divisible by the world size): the two weight-gradient regions are REDUCE-SCATTERED by
        dictionary rows, each rank runs the norm / clip / Adam / renorm kernel on its S/N rows only (Adam moments
        exist only for those rows), and the updated rows, their fp16 operand copy and the row-norm maximum are
        ALL-GATHERED.  Same bytes on the wire, the 28 B/param optimizer pass shrinks by N, but nothing overlaps.
"""

from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib
from .engine import LOSS_KEYS, Engine


class DataParallelTrainer:
    def __init__(self, engine: Engine, group=None, sharded: bool | None = None, n_chunks: int = 4,
                 gather_group=None, reserved_sms: int = 0, overlap_decoder_update: bool = False):
        """`gather_group` (sharded mode): a second process group (ideally created with few NCCL CTAs, e.g.
        `ProcessGroupNCCL.Options().config.max_ctas = 4`) on which the all-gathers of the fp32 rows run in the
        background, beside the top-k screen of the NEXT step, which then leaves `reserved_sms` SMs idle for them."""
        self.eng = engine
        import os

        # single rank: run the decoder half of Adam beside the next step's screen (Engine.train_step).  Measured at c3
        # (profiles/README.md): 4.33 instead of 4.37 ms/step -- the update has to fit into the ~6 K registers per SM
        # the resident screen CTA leaves free (one 64-thread block), which stretches it over the whole screen, so it
        # is off by default; SAEV_B200_OVERLAP_ADAM=1 turns it on.
        self.overlap_decoder_update = overlap_decoder_update or os.environ.get("SAEV_B200_OVERLAP_ADAM", "0") == "1"
        self._use_hp = os.environ.get("SAEV_B200_HP_STREAM", "1") != "0"
        self._split_decode = os.environ.get("SAEV_B200_DP_SPLIT_DECODE", "1") != "0" and hasattr(engine, "wnorm_rows")
        self._small = None  # staging for the two bias gradients (one all-reduce instead of two)
        self._hp_stream = None
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self._w_dec_normalized = False
        self.gather_group = gather_group
        self._pending = []  # background all-gathers of the previous step
        S = getattr(engine, "S", 0)
        can_shard = (self.world > 1 and S > 0 and S % self.world == 0 and hasattr(engine, "set_optimizer_shard")
                     and getattr(engine.cfg, "activation", "topk") == "topk")
        self.sharded = bool(sharded) and can_shard
        # row chunks of the overlapped all-reduce (multiples of 8 rows: one warp per row, 8 rows per block)
        self.chunks = []
        if (self.world > 1 and not self.sharded and n_chunks > 1 and hasattr(engine, "backward_stage")
                and getattr(engine.cfg, "activation", "topk") == "topk"):
            step_rows = -(-S // n_chunks)
            step_rows = (step_rows + 7) // 8 * 8
            self.chunks = [(r, min(S, r + step_rows)) for r in range(0, S, step_rows)]
        self.shard_chunks = []  # sharded + chunked: [(chunk_begin, chunk_end, own_begin, own_end)]
        if self.sharded:
            C = n_chunks if (n_chunks > 1 and S % (n_chunks * self.world * 8) == 0 and hasattr(engine, "backward_stage")) else 1
            D = engine.D
            SD = S * D
            if C > 1:
                per_chunk, per_rank = S // C, S // (C * self.world)
                for c in range(C):
                    c0 = c * per_chunk
                    self.shard_chunks.append((c0, c0 + per_chunk, c0 + self.rank * per_rank, c0 + (self.rank + 1) * per_rank))
                self._ranges = []
                for _, _, o0, o1 in self.shard_chunks:
                    self._ranges += [(o0 * D, o1 * D), (SD + S + o0 * D, SD + S + o1 * D)]
                self.j0, self.j1 = self.shard_chunks[0][2], self.shard_chunks[0][3]
            else:
                rows = S // self.world
                self.j0, self.j1 = self.rank * rows, (self.rank + 1) * rows
                self._ranges = [(self.j0 * D, self.j1 * D), (SD + S + self.j0 * D, SD + S + self.j1 * D)]
            engine.set_optimizer_shard(self.j0, self.j1)
            if self.rank == 0:  # the (all-reduced) bias gradients are counted once
                self._ranges += [(SD, SD + S), (2 * SD + S, 2 * SD + S + D)]
            if gather_group is not None and reserved_sms > 0:
                engine.set_reserved_sms(reserved_sms)

    def broadcast_params(self, src: int = 0) -> None:
        """Make every replica start from rank `src`'s parameters (model init is unseeded in saev)."""
        if self.world > 1:
            dist.broadcast(self.eng.params, src=src, group=self.group)
            self.eng.sync_weights()
        self._w_dec_normalized = False

    def step(self, x: torch.Tensor, lr: float, *, max_norm: float = 1.0, fused_renorm: bool = True) -> torch.Tensor:
        """One iteration of saev's loop body (train.py:332-460) on this rank's rows `x[B_local, D]`.
        Returns the device tensor of this rank's loss partials (see `global_losses`)."""
        eng, g = self.eng, self.group
        tokens_global = x.shape[0] * self.world  # equal per-rank batches (the loader guarantees it)
        if self.world == 1 and self.overlap_decoder_update and hasattr(eng, "train_step"):
            # single rank: the decoder half of Adam runs beside the next step's screen (Engine.train_step)
            # The step runs on a HIGH-priority stream of its own: the deferred decoder update sits on a default-priority
            # side stream, and the block scheduler must hand free SMs to the screen kernel's CTAs first.
            cur = torch.cuda.current_stream(x.device)
            if self._hp_stream is None and self._use_hp:
                self._hp_stream = torch.cuda.Stream(device=x.device, priority=-1)
            if self._hp_stream is not None:
                self._hp_stream.wait_stream(cur)
                with torch.cuda.stream(self._hp_stream):
                    out = eng.train_step(x, lr, max_norm=max_norm, fused_renorm=fused_renorm,
                                         pre_normalized=self._w_dec_normalized, overlap_decoder_update=True)
                cur.wait_stream(self._hp_stream)
            else:
                out = eng.train_step(x, lr, max_norm=max_norm, fused_renorm=fused_renorm,
                                     pre_normalized=self._w_dec_normalized, overlap_decoder_update=True)
            self._w_dec_normalized = fused_renorm and eng.cfg.normalize_w_dec
            return out
        if not self._w_dec_normalized:
            eng.normalize_w_dec()
        if self.world == 1:
            eng.forward(x, training=True, tokens_global=tokens_global)
        else:
            topk = getattr(eng.cfg, "activation", "topk") == "topk"
            if not topk:  # the dense (ReLU) path has no split phase A
                eng.forward(x, training=True, phase=_lib.PHASE_A, tokens_global=tokens_global)
                dist.all_reduce(eng.active_flags(), op=dist.ReduceOp.MAX, group=g)
                eng.forward(x, training=True, phase=_lib.PHASE_B, tokens_global=tokens_global)
            else:
                self._forward_topk(x, tokens_global)
        renorm = fused_renorm and eng.cfg.normalize_w_dec
        return self._backward_and_update(x, lr, max_norm, renorm, tokens_global)

    def _forward_topk(self, x, tokens_global):
        """Objective forward on this rank's rows, with the activity flags of all ranks folded in between phases."""
        eng, g = self.eng, self.group
        # the screen needs only the fp16 operand copy (gathered synchronously at the end of the last step); the fp32
        # rows may still be arriving on the gather group's stream
        eng.forward(x, training=True, phase=_lib.PHASE_A_SCREEN, tokens_global=tokens_global)
        if self._pending:
            self.finish()
        if self._split_decode:
            # the activity flags are final after the re-score: their MAX all-reduce runs beside the decode
            eng.forward(x, training=True, phase=_lib.PHASE_A_RESCORE, tokens_global=tokens_global)
            wk = dist.all_reduce(eng.active_flags(), op=dist.ReduceOp.MAX, group=g, async_op=True)
            eng.forward(x, training=True, phase=_lib.PHASE_A_DECODE, tokens_global=tokens_global)
            wk.wait()
        else:
            eng.forward(x, training=True, phase=_lib.PHASE_A_REST, tokens_global=tokens_global)
            dist.all_reduce(eng.active_flags(), op=dist.ReduceOp.MAX, group=g)
        eng.forward(x, training=True, phase=_lib.PHASE_B, tokens_global=tokens_global)

    def _backward_and_update(self, x, lr, max_norm, renorm, tokens_global):
        eng, g = self.eng, self.group
        if self.chunks:
            eng.backward_stage(x, 0, tokens_global=tokens_global)
            works = []
            for r0, r1 in self.chunks:
                eng.backward_stage(x, 1, r0, r1, tokens_global=tokens_global)
                works.append(dist.all_reduce(eng.gW_enc_t[r0:r1], op=dist.ReduceOp.SUM, group=g, async_op=True))
                works.append(dist.all_reduce(eng.gW_dec[r0:r1], op=dist.ReduceOp.SUM, group=g, async_op=True))
            works.append(dist.all_reduce(eng.gb_enc, op=dist.ReduceOp.SUM, group=g, async_op=True))
            works.append(dist.all_reduce(eng.gb_dec, op=dist.ReduceOp.SUM, group=g, async_op=True))
            for wk in works:
                wk.wait()  # stream-level wait (NCCL): the norm / Adam kernels below see the reduced bucket
            eng.grad_sumsq()
            eng.adam_step(lr, max_norm=max_norm, renorm_w_dec=renorm)
            self._w_dec_normalized = renorm
            return eng.losses
        if self.shard_chunks:
            return self._step_sharded_chunked(x, lr, max_norm, renorm, tokens_global)
        eng.backward(x, tokens_global=tokens_global)
        if not self.sharded:
            if self.world > 1:
                dist.all_reduce(eng.grads, op=dist.ReduceOp.SUM, group=g)
            eng.grad_sumsq(local=self.world == 1)
            eng.adam_step(lr, max_norm=max_norm, renorm_w_dec=renorm)
        else:
            j0, j1 = self.j0, self.j1
            dist.reduce_scatter_tensor(eng.gW_enc_t[j0:j1], eng.gW_enc_t, op=dist.ReduceOp.SUM, group=g)
            dist.reduce_scatter_tensor(eng.gW_dec[j0:j1], eng.gW_dec, op=dist.ReduceOp.SUM, group=g)
            self._all_reduce_bias_grads()
            eng.grad_sumsq_ranges(self._ranges)
            dist.all_reduce(eng.sumsq, op=dist.ReduceOp.SUM, group=g)
            eng.adam_step(lr, max_norm=max_norm, renorm_w_dec=renorm)  # rows [j0, j1) + both bias vectors
            shadow = eng.shadow_weights()
            dist.all_gather_into_tensor(shadow, shadow[j0:j1], group=g)
            wn = eng.wnorm_rows()
            dist.all_gather_into_tensor(wn, wn[j0:j1], group=g)
            dist.all_reduce(eng.wnorm_scalar(), op=dist.ReduceOp.MAX, group=g)
            if self.gather_group is not None:
                gg = self.gather_group
                self._pending = [dist.all_gather_into_tensor(eng.W_enc_t, eng.W_enc_t[j0:j1], group=gg, async_op=True),
                                 dist.all_gather_into_tensor(eng.W_dec, eng.W_dec[j0:j1], group=gg, async_op=True)]
            else:
                dist.all_gather_into_tensor(eng.W_enc_t, eng.W_enc_t[j0:j1], group=g)
                dist.all_gather_into_tensor(eng.W_dec, eng.W_dec[j0:j1], group=g)
        self._w_dec_normalized = renorm
        return eng.losses

    def _step_sharded_chunked(self, x, lr, max_norm, renorm, tokens_global):
        """Backward + exchange + optimizer of `step` for the sharded optimizer with interleaved chunk ownership."""
        eng, g = self.eng, self.group
        eng.backward_stage(x, 0, tokens_global=tokens_global)
        works = []
        for c0, c1, o0, o1 in self.shard_chunks:
            eng.backward_stage(x, 1, c0, c1, tokens_global=tokens_global)
            works.append(dist.reduce_scatter_tensor(eng.gW_enc_t[o0:o1], eng.gW_enc_t[c0:c1], op=dist.ReduceOp.SUM,
                                                    group=g, async_op=True))
            works.append(dist.reduce_scatter_tensor(eng.gW_dec[o0:o1], eng.gW_dec[c0:c1], op=dist.ReduceOp.SUM, group=g,
                                                    async_op=True))
        works.append(dist.all_reduce(eng.gb_enc, op=dist.ReduceOp.SUM, group=g, async_op=True))
        works.append(dist.all_reduce(eng.gb_dec, op=dist.ReduceOp.SUM, group=g, async_op=True))
        for wk in works:
            wk.wait()
        eng.grad_sumsq_ranges(self._ranges)
        dist.all_reduce(eng.sumsq, op=dist.ReduceOp.SUM, group=g)
        step = eng.step_count + 1
        for i, (_, _, o0, o1) in enumerate(self.shard_chunks):
            eng.set_optimizer_shard(o0, o1)
            parts = _lib.ADAM_ALL if i == 0 else (_lib.ADAM_ALL | _lib.ADAM_ROWS_ONLY | _lib.ADAM_KEEP_MAXIMA)
            eng.adam_step(lr, max_norm=max_norm, renorm_w_dec=renorm, parts=parts, step=None if i == 0 else step)
        shadow, wn = eng.shadow_weights(), eng.wnorm_rows()
        for c0, c1, o0, o1 in self.shard_chunks:
            dist.all_gather_into_tensor(shadow[c0:c1], shadow[o0:o1], group=g)
            dist.all_gather_into_tensor(wn[c0:c1], wn[o0:o1], group=g)
        dist.all_reduce(eng.wnorm_scalar(), op=dist.ReduceOp.MAX, group=g)
        gg = self.gather_group if self.gather_group is not None else g
        pend = []
        for c0, c1, o0, o1 in self.shard_chunks:
            pend.append(dist.all_gather_into_tensor(eng.W_enc_t[c0:c1], eng.W_enc_t[o0:o1], group=gg, async_op=True))
            pend.append(dist.all_gather_into_tensor(eng.W_dec[c0:c1], eng.W_dec[o0:o1], group=gg, async_op=True))
        if self.gather_group is not None:
            self._pending = pend
        else:
            for wk in pend:
                wk.wait()
        self._w_dec_normalized = renorm
        return eng.losses

    def _all_reduce_bias_grads(self) -> None:
        """gb_enc and gb_dec in ONE all-reduce (they are not adjacent in the bucket; every small collective costs
        ~30 us of latency at 8 ranks)."""
        eng, g = self.eng, self.group
        S, D = eng.gb_enc.numel(), eng.gb_dec.numel()
        if self._small is None:
            self._small = torch.empty(S + D, dtype=eng.gb_enc.dtype, device=eng.gb_enc.device)
        self._small[:S].copy_(eng.gb_enc)
        self._small[S:].copy_(eng.gb_dec)
        dist.all_reduce(self._small, op=dist.ReduceOp.SUM, group=g)
        eng.gb_enc.copy_(self._small[:S])
        eng.gb_dec.copy_(self._small[S:])

    def finish(self) -> None:
        """Make the current stream wait for the background all-gathers of the last step (call before anything reads
        the fp32 parameters: checkpointing, evaluation, the next step does it itself)."""
        for wk in self._pending:
            wk.wait()
        self._pending = []
        if hasattr(self.eng, "flush"):
            self.eng.flush()

    def global_losses(self) -> dict:
        """Loss scalars of the global batch (host sync; call on log steps only)."""
        eng = self.eng
        vals = eng.losses.clone()
        if self.world > 1:
            n_dead = vals[5].clone()
            vals[5] = 0
            dist.all_reduce(vals, op=dist.ReduceOp.SUM, group=self.group)
            vals[5] = n_dead  # identical on every rank, not additive
        return dict(zip(LOSS_KEYS, vals.tolist()))
