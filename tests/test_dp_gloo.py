"""World-size-2 test of the data-parallel step sequencing (saev_b200/parallel.py) over the gloo backend on CPU.

The CUDA engine cannot run here, so each rank drives `DataParallelTrainer` with a stand-in engine whose arithmetic is
the ORACLE's (test infrastructure), exposing the same surface as `saev_b200.engine.Engine` with the same contract
(kernels divide by the GLOBAL batch; phase A marks the atoms that fired, phase B consumes the all-reduced flags).
What is under test is the host logic: which buffers are all-reduced with which op and in which order, that the
clip norm is taken on the reduced gradient, and that "2 ranks x B/2 rows" then equals "1 rank x B rows" -- compared
against a plain single-process oracle run on the full batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sae_oracle as orc
from saev_b200 import _lib
from saev_b200.parallel import DataParallelTrainer

D, S, K, B = 16, 64, 4, 24
CFG = orc.OracleConfig(d_model=D, d_sae=S, top_k=K, aux=True, k_aux=8, dead_threshold_tokens=2 * B, lr=1e-2,
                       n_lr_warmup=2, n_steps=10)


class OracleEngine:
    """Engine look-alike (only what DataParallelTrainer touches) over flat parameter / gradient / moment buffers in the
    engine's order [W_enc_t, b_enc, W_dec, b_dec]; forward / backward math from oracle/sae_oracle.py."""

    def __init__(self, W_enc, b_enc, W_dec, b_dec):
        self.cfg = type("C", (), {"normalize_w_dec": True, "activation": "topk"})()
        self.S, self.D = S, D
        self.n = 2 * S * D + S + D
        self.params, self.grads = torch.zeros(self.n), torch.zeros(self.n)
        self.m, self.v = torch.zeros(self.n), torch.zeros(self.n)
        self.W_enc_t, self.b_enc, self.W_dec, self.b_dec = self._views(self.params)
        self.gW_enc_t, self.gb_enc, self.gW_dec, self.gb_dec = self._views(self.grads)
        self.W_enc_t.copy_(W_enc.t()); self.b_enc.copy_(b_enc); self.W_dec.copy_(W_dec); self.b_dec.copy_(b_dec)
        self.losses, self.sumsq = torch.zeros(8), torch.zeros(1)
        self._flags = torch.zeros(S, dtype=torch.int32)
        self.toks = torch.zeros(S, dtype=torch.int64)
        self._shadow = torch.zeros(S, D, dtype=torch.float16)
        self._wnorm = torch.zeros(1)
        self._wrows = torch.zeros(S)
        self.shard = None
        self.t = 0
        self.calls = []

    def _views(self, flat):
        o1, o2, o3 = S * D, S * D + S, 2 * S * D + S
        return flat[:o1].view(S, D), flat[o1:o2], flat[o2:o3].view(S, D), flat[o3:]

    def _state(self):
        return orc.OracleState.from_params(self.W_enc_t.t(), self.b_enc, self.W_dec, self.b_dec)

    def active_flags(self):
        return self._flags

    def shadow_weights(self):
        return self._shadow

    def wnorm_scalar(self):
        return self._wnorm

    def wnorm_rows(self):
        return self._wrows

    def sync_weights(self):
        self._shadow.copy_(self.W_enc_t.half())
        self._wnorm[0] = self.W_enc_t.pow(2).sum(1).max()
        self._wrows.copy_(self.W_enc_t.norm(dim=1))

    def set_optimizer_shard(self, j0, j1):
        self.shard = (j0, j1)

    def normalize_w_dec(self):
        self.W_dec.copy_(orc.normalize_w_dec(self.W_dec))

    def forward(self, x, *, training=True, phase=_lib.PHASE_ALL, tokens_global=0):
        Bl = x.shape[0]
        tg = tokens_global or Bl
        if phase & (_lib.PHASE_A | _lib.PHASE_A_REST | _lib.PHASE_A_RESCORE):  # (the screen / decode halves: no-ops here)
            self.calls.append("A")
            # the screen runs on the bf16 operand copy: it must be complete and current on every rank
            assert torch.equal(self._shadow, self.W_enc_t.half()), "stale fp16 operand rows"
            assert float(self._wnorm) == pytest.approx(float(self.W_enc_t.pow(2).sum(1).max()), rel=1e-6)
            assert torch.allclose(self._wrows, self.W_enc_t.norm(dim=1), rtol=1e-6), "stale per-column norms"
            st = self._st = self._state()
            h = orc.encode_pre(x, st.W_enc, st.b_enc)
            f, mask = orc.topk_activation(h, K)
            x_hat = orc.decode(f, st.W_dec, st.b_dec)
            self._fw = dict(h=h, f=f, mask=mask, x_hat=x_hat, r=x_hat - x)
            self._flags.copy_((f.abs() > 0).any(0).to(torch.int32))
        if phase & _lib.PHASE_B:
            self.calls.append("B")
            fw, st = self._fw, self._st
            self.toks = torch.where(self._flags > 0, torch.zeros_like(self.toks), self.toks + tg)
            dead = self.toks >= CFG.dead_threshold_tokens
            aux, fa, r_aux = orc.auxk(fw["h"], fw["r"], dead, st.W_dec, st.b_dec, CFG.k_aux, CFG.aux_alpha)
            scale = Bl / tg  # local means -> partial sums over the global denominator
            fw.update(f_aux=fa[0] if fa else None, mask_aux=fa[1] if fa else None, r_aux=r_aux)
            mse = fw["r"].pow(2).mean() * scale
            l1 = fw["f"].abs().sum(1).mean() * scale
            l0 = (fw["f"] != 0).float().sum(1).mean() * scale
            self.losses[:] = torch.stack([mse, aux * scale, torch.zeros(()), l0, l1, dead.sum().float(),
                                          mse + aux * scale, torch.zeros(())])
        return self.losses

    def backward(self, x, *, tokens_global=0):
        self.calls.append("bwd")
        fw, st, Bl = self._fw, self._st, x.shape[0]
        out = orc.ForwardOut(fw["h"], fw["f"], fw["mask"], fw["x_hat"], fw["r"], None, None, None, None, None, 0,
                             fw["f_aux"], fw["mask_aux"], fw["r_aux"])
        g = orc.backward(CFG, st, x, out)
        g["W_dec"] = orc.remove_parallel_grads(g["W_dec"], st.W_dec)
        scale = Bl / (tokens_global or Bl)
        flat = torch.cat([g["W_enc"].t().reshape(-1), g["b_enc"], g["W_dec"].reshape(-1), g["b_dec"]]) * scale
        self.grads.copy_(flat)

    def backward_stage(self, x, stage, row_begin=0, row_end=0, *, tokens_global=0):
        """Staged backward: stage 0 = bias gradient (+ AuxK rows), stage 1 = weight-gradient rows [row_begin, row_end)."""
        if stage == 0:
            self.backward(x, tokens_global=tokens_global)
            self.calls[-1] = "bwd0"
            self._staged = self.grads.clone()
            self.grads.fill_(float("nan"))  # anything a later stage forgets to produce poisons the result
            self.gb_dec.copy_(self._views(self._staged)[3])
        else:
            self.calls.append("bwd1")
            st = self._views(self._staged)
            self.gW_enc_t[row_begin:row_end] = st[0][row_begin:row_end]
            self.gb_enc[row_begin:row_end] = st[1][row_begin:row_end]
            self.gW_dec[row_begin:row_end] = st[2][row_begin:row_end]

    def grad_sumsq(self, *, local=False):
        assert not local, "with more than one rank the norm must be taken on the all-reduced bucket"
        self.calls.append("sumsq")
        self.sumsq[0] = self.grads.double().pow(2).sum()

    def grad_sumsq_ranges(self, ranges):
        self.calls.append("sumsq")
        self.sumsq[0] = sum(self.grads[b:e].double().pow(2).sum() for b, e in ranges)

    @property
    def step_count(self):
        return self.t

    def adam_step(self, lr, *, max_norm=1.0, renorm_w_dec=False, parts=_lib.ADAM_ALL, step=None, **kw):
        """torch Adam(fused) formulas (oracle.adam_step) on the rows of the shard + both bias vectors (a sharded
        optimizer with several row ranges calls once per range: the later calls carry ROWS_ONLY | KEEP_MAXIMA)."""
        rows_only = bool(parts & _lib.ADAM_ROWS_ONLY)
        if step is None:
            assert not rows_only
            self.calls.append("adam")
            self.t += 1
        else:
            assert step == self.t and rows_only and (parts & _lib.ADAM_KEEP_MAXIMA)
        j0, j1 = self.shard or (0, S)
        coef = min(1.0, max_norm / (float(self.sumsq.sqrt()) + 1e-6))
        bc1, bc2s = 1.0 - CFG.beta1**self.t, (1.0 - CFG.beta2**self.t) ** 0.5
        ps, gs, ms, vs = self._views(self.params), self._views(self.grads), self._views(self.m), self._views(self.v)
        for i, rows in enumerate((slice(j0, j1), slice(None), slice(j0, j1), slice(None))):
            if rows_only and i in (1, 3):
                continue
            p, g, m, v = ps[i][rows], gs[i][rows] * coef, ms[i][rows], vs[i][rows]
            m.copy_(m + (g - m) * (1.0 - CFG.beta1))
            v.copy_(v * CFG.beta2 + (1.0 - CFG.beta2) * g * g)
            p.sub_((lr / bc1) * (m / (v.sqrt() / bc2s + CFG.eps)))
        if renorm_w_dec:
            self.W_dec[j0:j1] = orc.normalize_w_dec(self.W_dec[j0:j1])
        self._shadow[j0:j1] = self.W_enc_t[j0:j1].half()
        mx = self.W_enc_t[j0:j1].pow(2).sum(1).max()
        self._wnorm[0] = torch.maximum(self._wnorm[0], mx) if parts & _lib.ADAM_KEEP_MAXIMA else mx
        self._wrows[j0:j1] = self.W_enc_t[j0:j1].norm(dim=1)


def _data():
    g = torch.Generator().manual_seed(3)
    params = orc.init_params(D, S, g)
    basis = torch.randn(3, D, generator=g)
    xs = [torch.randn(B, 3, generator=g) @ basis + 0.05 * torch.randn(B, D, generator=g) for _ in range(4)]
    return params, xs


def _worker(rank, world, port, out_dir, sharded, n_chunks):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        params, xs = _data()
        eng = OracleEngine(*params)
        tr = DataParallelTrainer(eng, sharded=sharded, n_chunks=n_chunks)
        assert tr.world == world and tr.rank == rank and tr.sharded == sharded
        assert bool(tr.chunks) == (n_chunks > 1 and not sharded) and bool(tr.shard_chunks) == (n_chunks > 1 and sharded)
        tr.broadcast_params(0)
        per = B // world
        rec = []
        lr = 0.0
        for step, x in enumerate(xs):
            tr.step(x[rank * per : (rank + 1) * per], lr, max_norm=CFG.grad_clip, fused_renorm=True)
            rec.append(tr.global_losses())
            lr = orc.warmup_cosine(step + 1, CFG.n_lr_warmup, CFG.lr, CFG.n_steps)
        # phase A -> flags all-reduce -> phase B -> backward (-> grads exchange) -> norm -> Adam, every step
        n_stage1 = len(tr.chunks) or len(tr.shard_chunks)
        bwd = ["bwd0"] + ["bwd1"] * n_stage1 if n_stage1 else ["bwd"]
        assert eng.calls == (["A", "B"] + bwd + ["sumsq", "adam"]) * len(xs)
        if sharded:  # Adam moments exist for the rank's own rows only
            own = torch.zeros(S, dtype=torch.bool)
            for o0, o1 in ([(c[2], c[3]) for c in tr.shard_chunks] or [eng.shard]):
                own[o0:o1] = True
            mW = eng._views(eng.m)[0]
            assert bool((mW[own] != 0).any()) and not bool((mW[~own] != 0).any())
        torch.save(dict(rec=rec, W_enc=eng.W_enc_t.t().clone(), W_dec=eng.W_dec.clone(), b_enc=eng.b_enc.clone(),
                        b_dec=eng.b_dec.clone(), toks=eng.toks), os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(120)
@pytest.mark.parametrize("sharded,n_chunks", [(False, 1), (False, 3), (True, 1), (True, 2)],
                         ids=["allreduce", "chunked-overlapped-allreduce", "sharded-optimizer",
                              "sharded-optimizer-chunked-reduce-scatter"])
def test_two_ranks_equal_one_rank_on_the_full_batch(tmp_path, sharded, n_chunks):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), sharded, n_chunks), nprocs=2, join=True)
    r0, r1 = (torch.load(tmp_path / f"rank{r}.pt") for r in range(2))
    # single-process reference on the full batches
    params, xs = _data()
    st = orc.OracleState.from_params(*params)
    ref = [orc.train_step(CFG, st, x) for x in xs]
    st.W_dec = orc.normalize_w_dec(st.W_dec)  # the trainer used the fused renorm at the end of each step
    assert any(r["n_dead"] > 0 and r["aux"] > 0 for r in ref), "the case must exercise AuxK"
    for k in ("W_enc", "W_dec", "b_enc", "b_dec"):
        assert torch.equal(r0[k], r1[k]), f"replicas diverged in {k}"
        assert torch.allclose(r0[k], getattr(st, k), rtol=1e-5, atol=1e-7), k
    assert torch.equal(r0["toks"], r1["toks"]) and torch.equal(r0["toks"], st.toks_since_active)
    for a, b, r in zip(r0["rec"], r1["rec"], ref):
        assert a == b
        for key in ("mse", "aux", "l0", "l1", "loss"):
            assert a[key] == pytest.approx(r[key], rel=1e-5, abs=1e-8), key
        assert int(a["n_dead"]) == r["n_dead"]
