"""GPU tests (run with `-m gpu`) of the pieces around the first training step: batch ownership of the loader
(saev's make_saes holds >= 4 batches before concatenating them), the device-side datapoint initialisation
against vectors the live reference produced (tests/golden/datapoint_init.npz, oracle/gen_golden_init.py), in-place
weight writes by stock torch code (load_state_dict, a non-fused optimizer), and the Muon split of train.py:296-306."""
import json
import pathlib

import numpy as np
import pytest
import torch

from saev_b200 import data, nn, optim
from tests.golden_util import rel_l2, t

pytestmark = pytest.mark.gpu
GOLDEN = pathlib.Path(__file__).resolve().parent / "golden"


def _write_dir(tmp_path, n_examples, T, D, ex_per_shard, seed=0):
    md = data.Metadata(family="fake-clip", ckpt="synthetic", layers=(0,), content_tokens_per_example=T, cls_token=False,
                       d_model=D, n_examples=n_examples, max_tokens_per_shard=ex_per_shard * T, data="", dataset="fake")
    root = tmp_path / "saev" / "shards"
    root.mkdir(parents=True)
    md.dump(root)
    rng = np.random.default_rng(seed)
    acts = rng.standard_normal((n_examples, 1, T, D), dtype=np.float32)
    with data.ShardWriter(root, md) as w:
        for s in range(0, n_examples, 7):  # ragged write_batch calls that straddle shard boundaries
            w.write_batch(torch.from_numpy(acts[s : s + 7]).cuda(), s)
    return root / md.hash, acts


def test_batches_stay_valid_while_the_caller_holds_them(tmp_path):
    """Default mode hands out OWNED tensors: hold every batch of an epoch (make_saes holds >= 4, train.py:147-160; the
    loader's ring has 3 slots), then check every row against the shard files.  With alias_ring=True the same pattern
    must show recycled buffers -- that mode is only for consumers that finish with a batch before asking for the next."""
    d, acts = _write_dir(tmp_path, n_examples=64, T=16, D=32, ex_per_shard=10)
    cfg = data.ShuffledConfig(shards=d, layer=0, batch_size=128, n_threads=2, buffer_size=2, seed=1, batch_timeout_s=20.0)
    dl = data.ShuffledDataLoader(cfg)
    held = [b for b in dl]
    assert len(held) == 8
    torch.cuda.synchronize()
    seen = set()
    for b in held:
        ex, tok = b["example_idx"].cpu().numpy(), b["token_idx"].cpu().numpy()
        assert np.array_equal(b["act"].cpu().numpy(), acts[ex, 0, tok]), "a held batch was overwritten"
        seen.update(zip(ex.tolist(), tok.tolist()))
    assert len(seen) == 64 * 16
    dl.shutdown()
    dl = data.ShuffledDataLoader(cfg, alias_ring=True)
    held = [b for b in dl]
    torch.cuda.synchronize()
    ptrs = {b["act"].data_ptr() for b in held}
    assert len(ptrs) <= 3, "alias_ring=True hands out views of the loader's ring"
    dl.shutdown()


class _FakeLoader:
    def __init__(self, batches):
        self.batches = batches
        self.n_samples = sum(len(b) for b in batches)
        self.drop_last, self.batch_size = False, len(batches[0])

    def __iter__(self):
        for b in self.batches:
            yield {"act": b}


def test_datapoint_init_reproduces_the_references_make_saes():
    """Same seed, same batches => the W_enc / W_dec the reference's make_saes left (train.py:141-185), from the
    device-side kernel (saev_b200_datapoint_init)."""
    z = np.load(GOLDEN / "datapoint_init.npz")
    D, S, seed = int(z["meta_D"]), int(z["meta_S"]), int(z["meta_seed"])
    batches = [t(b).cuda() for b in z["batches"]]
    torch.manual_seed(seed)
    saes = []
    for blend in z["blends"]:
        cfg = nn.SparseAutoencoderConfig(d_model=D, d_sae=S, activation=nn.TopK(top_k=8), reinit_blend=float(blend))
        saes.append(nn.SparseAutoencoder(cfg))  # (draws the same kaiming rows from the global generator as saev's class)
    saes = [s.cuda() for s in saes]
    nn.datapoint_init(saes, _FakeLoader(batches), noise_device="cpu")
    for i, sae in enumerate(saes):
        assert rel_l2(sae.W_dec.detach().cpu(), z[f"W_dec_{i}"]) < 1e-5, i
        assert rel_l2(sae.W_enc.detach().cpu(), z[f"W_enc_{i}"]) < 1e-5, i
        # the screen copy was rebuilt: a forward sees the new dictionary
        sae.eval()
        x = batches[0]
        out = sae(x)
        h = x.double() @ sae.W_enc.detach().double() + sae.b_enc.detach().double()
        assert rel_l2(out.f_x.double().sum(1).cpu(), h.topk(8, dim=1).values.sum(1).cpu()) < 1e-5


def test_in_place_writes_by_torch_are_noticed():
    """load_state_dict copies into the bound parameters in place (same storage): the screen's fp16 copy and norm
    bounds must follow, or the next forward screens with stale weights."""
    torch.manual_seed(0)
    cfg = nn.SparseAutoencoderConfig(d_model=64, d_sae=1024, activation=nn.TopK(top_k=8), reinit_blend=0.0)
    a, b = nn.SparseAutoencoder(cfg).cuda(), nn.SparseAutoencoder(cfg).cuda()
    x = torch.randn(256, 64, device="cuda")
    a.eval(), b.eval()
    a(x)  # binds the engine of `a` to its own weights
    a.load_state_dict(b.state_dict())  # in place
    fa, fb = a(x).f_x, b(x).f_x
    assert torch.equal(fa, fb)
    with torch.no_grad():
        a.W_enc.mul_(-1.0)  # any in-place op through torch
    h = -(x.double() @ b.W_enc.detach().double()) + b.b_enc.detach().double()
    assert rel_l2(a(x).f_x.double().sum(1).cpu(), h.topk(8, dim=1).values.sum(1).cpu()) < 1e-5


def test_muon_split_gets_an_eager_clip_and_fresh_screen_weights():
    """train.py:296-306 with cfg.optim == "muon": the two matrices go to torch.optim.Muon, the biases to Adam.  The
    patched clip_grad_norm_ must then scale .grad at once (no fused Adam will apply it), and the next forward must
    screen with the weights the stock optimizers wrote."""
    if not hasattr(torch.optim, "Muon"):
        pytest.skip("torch.optim.Muon not available")
    torch.manual_seed(1)
    cfg = nn.SparseAutoencoderConfig(d_model=64, d_sae=512, activation=nn.TopK(top_k=8, aux=nn.NoAux()), reinit_blend=0.0)
    sae = nn.SparseAutoencoder(cfg).cuda()
    obj = nn.get_objective(nn.Matryoshka(n_prefixes=1))
    sae.train(), obj.train()
    opts = [torch.optim.Muon([{"params": [sae.W_dec, sae.W_enc], "lr": 1e-2}]),
            optim.FusedAdam([{"params": [sae.b_dec, sae.b_enc], "lr": 1e-2}], fused=True)]
    x = torch.randn(512, 64, device="cuda") * 3
    for step in range(3):
        sae.normalize_w_dec()
        loss, _ = obj(sae, x)
        loss.loss.backward()
        sae.remove_parallel_grads()
        before = torch.cat([p.grad.flatten() for p in sae.parameters()]).norm()
        gn = optim.clip_grad_norm_(sae.parameters(), max_norm=0.05)
        after = torch.cat([p.grad.flatten() for p in sae.parameters()]).norm()
        assert float(gn) == pytest.approx(float(before), rel=1e-5)
        assert float(after) == pytest.approx(min(float(before), 0.05), rel=1e-4), "the clip was not applied"
        w0 = sae.W_enc.detach().clone()
        for o in opts:
            o.step()
        for o in opts:
            o.zero_grad()
        assert not torch.equal(w0, sae.W_enc.detach())
        # the forward after the stock update uses the updated encoder
        sae.eval()
        f = sae(x).f_x
        h = x.double() @ sae.W_enc.detach().double() + sae.b_enc.detach().double()
        assert rel_l2(f.double().sum(1).cpu(), h.topk(8, dim=1).values.sum(1).cpu()) < 1e-5, step
        sae.train()
