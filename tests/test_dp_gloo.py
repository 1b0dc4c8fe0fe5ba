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
    """Engine look-alike (only what DataParallelTrainer touches); math from oracle/sae_oracle.py."""

    def __init__(self, W_enc, b_enc, W_dec, b_dec):
        self.cfg = type("C", (), {"normalize_w_dec": True})()
        self.st = orc.OracleState.from_params(W_enc, b_enc, W_dec, b_dec)
        self.n = 2 * S * D + S + D
        self.params = torch.zeros(self.n)
        self.grads = torch.zeros(self.n)
        self.losses = torch.zeros(8)
        self.sumsq = torch.zeros(1)
        self._flags = torch.zeros(S, dtype=torch.int32)
        self.toks = torch.zeros(S, dtype=torch.int64)
        self.calls = []

    def active_flags(self):
        return self._flags

    def sync_weights(self):
        pass

    def normalize_w_dec(self):
        self.st.W_dec = orc.normalize_w_dec(self.st.W_dec)

    def forward(self, x, *, training=True, phase=_lib.PHASE_ALL, tokens_global=0):
        st, Bl = self.st, x.shape[0]
        tg = tokens_global or Bl
        if phase & _lib.PHASE_A:
            self.calls.append("A")
            h = orc.encode_pre(x, st.W_enc, st.b_enc)
            f, mask = orc.topk_activation(h, K)
            x_hat = orc.decode(f, st.W_dec, st.b_dec)
            self._fw = dict(h=h, f=f, mask=mask, x_hat=x_hat, r=x_hat - x)
            self._flags.copy_((f.abs() > 0).any(0).to(torch.int32))
        if phase & _lib.PHASE_B:
            self.calls.append("B")
            fw = self._fw
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
        fw, st, Bl = self._fw, self.st, x.shape[0]
        out = orc.ForwardOut(fw["h"], fw["f"], fw["mask"], fw["x_hat"], fw["r"], None, None, None, None, None, 0,
                             fw["f_aux"], fw["mask_aux"], fw["r_aux"])
        g = orc.backward(CFG, st, x, out)
        g["W_dec"] = orc.remove_parallel_grads(g["W_dec"], st.W_dec)
        scale = Bl / (tokens_global or Bl)
        flat = torch.cat([g["W_enc"].t().reshape(-1), g["b_enc"], g["W_dec"].reshape(-1), g["b_dec"]]) * scale
        self.grads.copy_(flat)

    def grad_sumsq(self, *, local=False):
        assert not local, "with more than one rank the norm must be taken on the all-reduced bucket"
        self.calls.append("sumsq")
        self.sumsq[0] = self.grads.double().pow(2).sum()

    def adam_step(self, lr, *, max_norm=1.0, renorm_w_dec=False, **kw):
        self.calls.append("adam")
        st = self.st
        o1, o2, o3 = S * D, S * D + S, 2 * S * D + S
        g = dict(W_enc=self.grads[:o1].view(S, D).t(), b_enc=self.grads[o1:o2], W_dec=self.grads[o2:o3].view(S, D),
                 b_dec=self.grads[o3:])
        coef = min(1.0, max_norm / (float(self.sumsq.sqrt()) + 1e-6))
        st.lr = lr
        orc.adam_step(CFG, st, {k: v * coef for k, v in g.items()})
        if renorm_w_dec:
            st.W_dec = orc.normalize_w_dec(st.W_dec)


def _data():
    g = torch.Generator().manual_seed(3)
    params = orc.init_params(D, S, g)
    basis = torch.randn(3, D, generator=g)
    xs = [torch.randn(B, 3, generator=g) @ basis + 0.05 * torch.randn(B, D, generator=g) for _ in range(4)]
    return params, xs


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        params, xs = _data()
        eng = OracleEngine(*params)
        tr = DataParallelTrainer(eng)
        assert tr.world == world and tr.rank == rank
        tr.broadcast_params(0)
        per = B // world
        rec = []
        lr = 0.0
        for step, x in enumerate(xs):
            tr.step(x[rank * per : (rank + 1) * per], lr, max_norm=CFG.grad_clip, fused_renorm=True)
            rec.append(tr.global_losses())
            lr = orc.warmup_cosine(step + 1, CFG.n_lr_warmup, CFG.lr, CFG.n_steps)
        # phase A -> flags all-reduce -> phase B -> backward -> grads all-reduce -> norm -> Adam, every step
        assert eng.calls == ["A", "B", "bwd", "sumsq", "adam"] * len(xs)
        torch.save(dict(rec=rec, W_enc=eng.st.W_enc, W_dec=eng.st.W_dec, b_enc=eng.st.b_enc, b_dec=eng.st.b_dec,
                        toks=eng.toks), os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(120)
def test_two_ranks_equal_one_rank_on_the_full_batch(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
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
