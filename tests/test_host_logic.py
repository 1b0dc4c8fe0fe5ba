"""CPU tests of the host-side logic around the CUDA path, against fixtures recorded from the LIVE reference by
oracle/gen_golden_host.py: lr schedules and BatchLimiter (scheduling.py), the shard chunk reader and the
draw/compaction planner of the native loader, the loader's sizes/validation, and the optimizer shims' fall-through
behaviour.  No compute entry point of the library is called here (there is no GPU)."""
import ctypes as C
import json
import pathlib
import zlib

import numpy as np
import pytest
import torch

from saev_b200 import _lib, data, optim, scheduling

GOLDEN = pathlib.Path(__file__).resolve().parent / "golden"
SCHED = json.loads((GOLDEN / "host_schedules.json").read_text())
CENSUS = json.loads((GOLDEN / "shards_census.json").read_text())
SHARDS = GOLDEN / CENSUS["dir"]


# ---- schedules ---------------------------------------------------------------------------------
@pytest.mark.parametrize("case", SCHED["warmup_cosine"], ids=lambda c: str(c["args"]))
def test_warmup_cosine_matches_reference(case):
    s = scheduling.WarmupCosine(*case["args"])
    assert repr(s) == case["repr"]
    assert [s.step() for _ in case["values"]] == case["values"]  # same float ops: bit-exact


@pytest.mark.parametrize("case", SCHED["warmup"], ids=lambda c: str(c["args"]))
def test_warmup_matches_reference(case):
    s = scheduling.Warmup(*case["args"])
    assert repr(s) == case["repr"]
    assert [s.step() for _ in case["values"]] == case["values"]


class _FakeLoader:
    def __init__(self, sizes, batch_size, drop_last):
        self.sizes, self.batch_size, self.drop_last = sizes, batch_size, drop_last
        self.extra_attr = "passthrough"

    def __iter__(self):
        for n in self.sizes:
            yield {"act": torch.zeros(n, 2)}


@pytest.mark.parametrize("case", SCHED["batch_limiter"], ids=lambda c: f"{c['sizes']}-{c['n_samples']}")
def test_batch_limiter_matches_reference(case):
    bl = scheduling.BatchLimiter(_FakeLoader(case["sizes"], case["batch_size"], case["drop_last"]), case["n_samples"])
    assert len(bl) == case["len"]
    assert [len(b["act"]) for b in bl] == case["yielded"]
    assert bl.n_seen == case["n_seen"]
    assert bl.extra_attr == case["extra_attr"]
    with pytest.raises(AttributeError, match="wrapped dataloader"):
        bl.no_such_attribute


def test_infer_batch_size_rules():
    f = scheduling._infer_batch_size
    assert f({"act": torch.zeros(3, 2)}, 9) == 3
    assert f({}, 9) == 9
    assert f(torch.zeros(5, 2), 9) == 5
    assert f(7, 9) == 9  # no __len__
    assert f({"act": torch.zeros(0, 2)}, 9) == 9  # empty -> fallback


# ---- native loader: host-only entry points --------------------------------------------------------
@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def _loader_cfg(md, layer, keep):
    order = np.arange(md.n_shards, dtype=np.int32)
    n_ex = np.array([n for _, n in data._load_shard_examples(SHARDS)], dtype=np.int32)
    keep.extend([order, n_ex, str(SHARDS).encode()])
    return _lib.LoaderCfg(
        shards_dir=keep[-1], examples_per_shard=md.examples_per_shard, n_layers=len(md.layers),
        tokens_per_example=md.tokens_per_example, d_model=md.d_model, layer_index=md.layers.index(layer),
        cls_token=int(md.cls_token), content_tokens=md.content_tokens_per_example,
        shard_order=order.ctypes.data_as(C.POINTER(C.c_int32)), shard_examples=n_ex.ctypes.data_as(C.POINTER(C.c_int32)),
        n_order=len(order), batch_size=4, pool_batches=2, n_threads=1, n_out_slots=3, chunk_examples=0,
        min_buffer_fill=0.0, reserved=0, n_rows_limit=-1, seed=0, labels=None, ignore_lut=None)


@pytest.mark.parametrize("case", CENSUS["cases"], ids=lambda c: f"layer{c['layer']}-ignore{c['ignore_labels']}")
def test_chunk_reader_reproduces_the_reference_loaders_rows(lib, case):
    """Union of all chunks == the multiset of rows saev's ShuffledDataLoader delivered for the same directory."""
    md = data.Metadata.load(SHARDS)
    keep = []
    c = _loader_cfg(md, case["layer"], keep)
    if case["ignore_labels"]:
        labels = np.fromfile(SHARDS / "labels.bin", dtype=np.uint8)
        lut = np.zeros(256, dtype=np.uint8)
        lut[case["ignore_labels"]] = 1
        keep.extend([labels, lut])
        c.labels, c.ignore_lut = labels.ctypes.data, lut.ctypes.data
    T, D = md.content_tokens_per_example, md.d_model
    rows = []
    for shard, (_, n_ex) in enumerate(data._load_shard_examples(SHARDS)):
        for ex0 in range(0, n_ex, 3):  # ragged chunks on purpose
            n = min(3, n_ex - ex0)
            act = np.full((n * T, D), np.nan, dtype=np.float32)
            meta = np.zeros((n * T, 2), dtype=np.int32)
            got = lib.saev_b200_loader_read_chunk(C.byref(c), shard, ex0, n, act.ctypes.data, meta.ctypes.data)
            assert 0 <= got <= n * T
            rows += [[int(meta[i, 0]), int(meta[i, 1]), zlib.crc32(act[i].tobytes())] for i in range(got)]
    assert sorted(rows) == case["rows"]
    assert len(rows) == case["n_samples"]


def test_chunk_reader_matches_the_written_activations(lib):
    md = data.Metadata.load(SHARDS)
    acts = np.load(GOLDEN / "shards_acts.npy")  # [example, layer, token (CLS first), d_model] as handed to ShardWriter
    keep = []
    c = _loader_cfg(md, 3, keep)
    T, D = md.content_tokens_per_example, md.d_model
    act = np.zeros((2 * T, D), dtype=np.float32)
    meta = np.zeros((2 * T, 2), dtype=np.int32)
    assert lib.saev_b200_loader_read_chunk(C.byref(c), 1, 1, 2, act.ctypes.data, meta.ctypes.data) == 2 * T
    np.testing.assert_array_equal(act.reshape(2, T, D), acts[5:7, 1, 1:, :])  # shard 1 holds examples 4..7
    assert meta[:, 0].tolist() == [5] * T + [6] * T and meta[:, 1].tolist() == list(range(T)) * 2
    assert lib.saev_b200_loader_read_chunk(C.byref(c), 7, 0, 1, act.ctypes.data, meta.ctypes.data) < 0  # no such shard
    assert b"cannot open" in lib.saev_b200_loader_last_error(None)


@pytest.mark.parametrize("fill,need", [(1, 1), (8, 8), (10, 3), (1000, 64), (65, 64), (4096, 1), (300, 299)])
def test_plan_draw_is_a_sample_without_replacement_and_repacks_the_pool(lib, fill, need):
    sel = np.zeros(need, dtype=np.int32)
    src = np.zeros(need, dtype=np.int32)
    dst = np.zeros(need, dtype=np.int32)
    n = C.c_int32()
    assert lib.saev_b200_loader_plan_draw(1234, fill, need, sel.ctypes.data, src.ctypes.data, dst.ctypes.data, C.byref(n)) == 0
    assert len(set(sel.tolist())) == need and sel.min() >= 0 and sel.max() < fill
    pool = np.arange(fill)
    drawn = pool[sel]
    m = n.value
    assert m == int((sel < fill - need).sum())
    assert set(dst[:m].tolist()) == set(sel[sel < fill - need].tolist())  # every hole is filled exactly once
    assert all(s >= fill - need and s not in set(sel.tolist()) for s in src[:m].tolist())
    assert len(set(src[:m].tolist())) == m
    pool[dst[:m]] = pool[src[:m]]
    assert sorted(pool[: fill - need].tolist() + drawn.tolist()) == list(range(fill))  # nothing lost, nothing doubled


def test_plan_draw_simulated_epoch_delivers_every_row_exactly_once(lib):
    """Numpy simulation of the device pool driven by the planner: ragged appends, draws, a short last batch."""
    rng = np.random.default_rng(0)
    total, batch, cap = 1003, 64, 256
    pool = np.full(cap, -1)
    fill, appended, out, seed = 0, 0, [], 0
    while len(out) < total:
        while appended < total and fill + 37 <= cap and (fill < batch or rng.random() < 0.5):
            n = min(37, total - appended)
            pool[fill : fill + n] = np.arange(appended, appended + n)
            fill += n
            appended += n
        need = min(batch, total - len(out))
        if fill < need:
            continue
        sel, src, dst = (np.zeros(need, dtype=np.int32) for _ in range(3))
        m = C.c_int32()
        seed += 1
        assert lib.saev_b200_loader_plan_draw(seed, fill, need, sel.ctypes.data, src.ctypes.data, dst.ctypes.data, C.byref(m)) == 0
        out += pool[sel].tolist()
        pool[dst[: m.value]] = pool[src[: m.value]]
        fill -= need
    assert sorted(out) == list(range(total))
    assert out != sorted(out)  # and it is shuffled


def test_plan_draw_rejects_bad_arguments(lib):
    buf = np.zeros(4, dtype=np.int32)
    n = C.c_int32()
    assert lib.saev_b200_loader_plan_draw(0, 3, 4, buf.ctypes.data, buf.ctypes.data, buf.ctypes.data, C.byref(n)) != 0
    assert lib.saev_b200_loader_plan_draw(0, 3, 0, buf.ctypes.data, buf.ctypes.data, buf.ctypes.data, C.byref(n)) != 0


# ---- loader front end ---------------------------------------------------------------------------
@pytest.mark.parametrize("case", CENSUS["cases"], ids=lambda c: f"layer{c['layer']}-bs{c['batch_size']}")
def test_loader_sizes_match_reference(case):
    cfg = data.ShuffledConfig(shards=SHARDS, layer=case["layer"], batch_size=case["batch_size"],
                              ignore_labels=case["ignore_labels"], seed=3)
    dl = data.ShuffledDataLoader(cfg)
    assert dl.n_samples == case["n_samples"]
    assert len(dl) == case["len"] == dl.n_batches
    assert dl.batch_size == case["batch_size"] and dl.drop_last is False
    assert dl.manager_pid == -1 and dl.reservoir is None  # nothing running before iteration (shuffled.py:452-456)
    assert dl.metadata.n_examples == 10 and dl.metadata.content_tokens_per_example == 5


def test_loader_rank_sharding_is_a_partition_of_the_seeded_permutation():
    cfg = data.ShuffledConfig(shards=SHARDS, layer=0, batch_size=4, seed=3)
    parts = [data.ShuffledDataLoader(cfg, rank=r, world_size=2) for r in range(2)]
    orders = [p._orders[p.rank].tolist() for p in parts]
    assert sorted(orders[0] + orders[1]) == [0, 1, 2]
    assert orders[0] + orders[1] != [] and set(orders[0]).isdisjoint(orders[1])
    full = np.random.default_rng(3).permutation(3).tolist()  # shuffled.py:326-328
    assert orders[0] == full[0::2] and orders[1] == full[1::2]
    # both ranks deliver the same number of rows per epoch (lock step)
    assert parts[0].n_samples == parts[1].n_samples == min(parts[0]._rows_per_rank)


def test_loader_validation_errors_mirror_the_reference(tmp_path):
    with pytest.raises(RuntimeError, match="Activations are not saved"):
        data.ShuffledDataLoader(data.ShuffledConfig(shards=tmp_path / "nope", layer=0))
    with pytest.raises(NotImplementedError, match="scale_norm"):
        data.ShuffledDataLoader(data.ShuffledConfig(shards=SHARDS, layer=0, scale_norm=True))
    with pytest.raises(NotImplementedError, match="content"):
        data.ShuffledDataLoader(data.ShuffledConfig(shards=SHARDS, layer="all"))
    with pytest.raises(ValueError, match="not in"):
        data.ShuffledDataLoader(data.ShuffledConfig(shards=SHARDS, layer=5))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_loader_has_no_host_only_mode():
    dl = data.ShuffledDataLoader(data.ShuffledConfig(shards=SHARDS, layer=0, batch_size=4))
    with pytest.raises(RuntimeError, match="CUDA"):
        next(iter(dl))


# ---- optimizer shims: everything that is not a fused SAE goes to stock torch ------------------------
def test_clip_and_adam_fall_through_for_ordinary_parameters():
    torch.manual_seed(0)
    a = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(3))]
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    for ps in (a, b):
        for i, p in enumerate(ps):
            p.grad = torch.full_like(p, 0.5 + i)
    n1 = optim.clip_grad_norm_(a, 1.0)
    n2 = torch.nn.utils.clip_grad_norm_(b, 1.0)
    assert torch.equal(n1, n2)
    o1, o2 = optim.FusedAdam([{"params": a, "lr": 1e-2}]), torch.optim.Adam([{"params": b, "lr": 1e-2}])
    for _ in range(3):
        o1.step()
        o2.step()
    for p, q in zip(a, b):
        assert torch.equal(p, q)


def test_install_rebinds_and_restores_saev_names():
    import sys

    ref_src = pathlib.Path("/root/reference/src")
    if not ref_src.exists():
        pytest.skip("needs the reference checkout (build container only)")
    stubs = str(pathlib.Path(__file__).resolve().parent.parent / "oracle" / "ref_stubs")
    sys.path[:0] = [stubs, str(ref_src)]
    try:
        import saev.data
        import saev.framework.train as ref_train
        import saev.nn
        import saev.nn.modeling as M

        import saev_b200
        from saev_b200 import dropin as inst

        orig = (saev.nn.SparseAutoencoder, saev.nn.get_objective, torch.optim.Adam, torch.nn.utils.clip_grad_norm_,
                saev.data.ShuffledDataLoader)
        orig_eval = ref_train.evaluate
        saev_b200.install()
        try:
            assert ref_train.evaluate is not orig_eval and ref_train.evaluate.__name__ == "evaluate_b200"
            assert torch.optim.Adam is optim.FusedAdam and torch.nn.utils.clip_grad_norm_ is optim.clip_grad_norm_
            assert saev.data.ShuffledDataLoader is data.ShuffledDataLoader
            sae = saev.nn.SparseAutoencoder(M.SparseAutoencoderConfig(d_model=16, d_sae=64, reinit_blend=0.0))
            assert isinstance(sae, M.SparseAutoencoder) and isinstance(sae, inst.dropin_class())
            assert list(sae.state_dict()) == ["W_dec", "b_dec", "W_enc", "b_enc"]  # checkpoint keys, modeling.py:312-327
            assert sae.W_enc.shape == (16, 64) and sae.W_dec.shape == (64, 16)
            with pytest.raises(RuntimeError, match="CUDA"):
                sae(torch.zeros(2, 16))  # no CPU fallback
        finally:
            saev_b200.uninstall()
        assert (saev.nn.SparseAutoencoder, saev.nn.get_objective, torch.optim.Adam, torch.nn.utils.clip_grad_norm_,
                saev.data.ShuffledDataLoader) == orig
        assert ref_train.evaluate is orig_eval
    finally:
        del sys.path[:2]


def test_finish_metrics_builds_saev_evalmetrics_from_accumulators():
    """saev_b200.evaluate.finish_metrics = train.py:568-616 on the accumulator layout of saev_b200_eval_accumulate;
    in drop-in mode the result is saev's own (beartype-checked) EvalMetrics."""
    from saev_b200 import evaluate as ev

    D, S, n_tok = 4, 6, 10
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n_tok, D, generator=g, dtype=torch.float64)
    acc = torch.zeros(8 + D, dtype=torch.float64)
    acc[0], acc[1], acc[4], acc[5], acc[6], acc[7] = (x * x).sum(), 3.5, 2.0 * n_tok, 7.0 * n_tok, 0.25 * n_tok, n_tok
    acc[8:] = x.sum(0)
    state = dict(acc=acc, n_fired=torch.tensor([0.0, 1.0, 10.0, 0.0, 3.0, 5.0]), values=torch.tensor([0.0, 2.0, 5.0, 0.0, -3.0, 1.0]))
    m = ev.finish_metrics(state)
    base = float((x * x).sum() - x.sum(0).dot(x.sum(0)) / n_tok)
    assert m.sse_baseline == pytest.approx(base) and m.normalized_mse == pytest.approx(3.5 / base)
    assert (m.l0, m.l1, m.mse) == pytest.approx((2.0, 7.0, 0.25))
    assert (m.n_dead, m.n_almost_dead, m.n_dense) == (2, 2, 4)
    assert torch.allclose(m.freqs, torch.tensor([0.0, 0.1, 1.0, 0.0, 0.3, 0.5]))
    assert torch.isnan(m.mean_values[0]) and m.mean_values[2] == 0.5 and m.mean_values[4] == -1.0
    ref_src = pathlib.Path("/root/reference/src")
    if ref_src.exists():
        import sys

        stubs = str(pathlib.Path(__file__).resolve().parent.parent / "oracle" / "ref_stubs")
        sys.path[:0] = [stubs, str(ref_src)]
        try:
            import saev.framework.train as ref_train

            r = ev.finish_metrics(state, ref_train.EvalMetrics)
            assert isinstance(r, ref_train.EvalMetrics) and r.n_dense == 4 and r.l1 == pytest.approx(7.0)
        finally:
            del sys.path[:2]


def test_topk_lists_to_csr_equals_scipy_on_dense():
    """saev_b200.sparse.topk_to_csr == scipy.sparse.csr_array(dense f_x) (inference.py:236), incl. empty slots,
    exact zeros among the selected values, negative values, and masked rows (:234)."""
    import numpy as np
    import scipy.sparse

    from saev_b200.sparse import topk_to_csr

    g = torch.Generator().manual_seed(4)
    B, K, S = 37, 8, 300
    idx = torch.stack([torch.randperm(S, generator=g)[:K] for _ in range(B)]).to(torch.int32)
    val = torch.randn(B, K, generator=g)
    val[3, 2] = 0.0  # an exact zero that TopK happened to select is not stored by scipy either
    idx[5, 6:] = -1  # empty slots (k > number of candidates)
    val[5, 6:] = 0.0
    mask = torch.ones(B, dtype=torch.bool)
    mask[[0, 11]] = False
    for m in (None, mask):
        dense = torch.zeros(B, S)
        for b in range(B):
            for k in range(K):
                if idx[b, k] >= 0:
                    dense[b, idx[b, k]] = val[b, k]
        if m is not None:
            dense[~m] = 0.0
        ref = scipy.sparse.csr_array(dense.numpy())
        got = topk_to_csr(idx, val, S, m)
        assert got.shape == ref.shape and got.nnz == ref.nnz
        assert np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices)
        assert np.array_equal(got.data, ref.data) and got.has_canonical_format


def test_evaluate_loop_against_the_oracle_with_a_stand_in_engine():
    """saev_b200.evaluate.evaluate (loader construction, BatchLimiter(n_val), per-batch accumulate, closing formulas)
    with an engine stand-in that fills the accumulator layout of saev_b200_eval_accumulate from the oracle's
    eval-mode forward; the result must equal oracle.evaluate (train.py:510-618) on the batches that were drawn."""
    import types

    from oracle import sae_oracle as orc
    from saev_b200 import evaluate as ev

    D, S, K = 16, 64, 4
    g = torch.Generator().manual_seed(12)
    W_enc, b_enc, W_dec, b_dec = orc.init_params(D, S, g)
    b_enc = 0.1 * torch.randn(S, generator=g)
    ocfg = orc.OracleConfig(d_model=D, d_sae=S, top_k=K)
    st = orc.OracleState.from_params(W_enc, b_enc, W_dec, b_dec)
    data = torch.randn(100, D, generator=g)

    class Loader:  # 100 rows in batches of 32 (short last batch), like ShuffledDataLoader with drop_last=False
        batch_size, drop_last, n_samples = 32, False, 100
        made = 0

        def __init__(self, cfg):
            Loader.made += 1

        def __iter__(self):
            for i in range(0, 100, 32):
                yield {"act": data[i:i + 32]}

        def shutdown(self):
            Loader.closed = True

    class Engine:
        def __init__(self):
            self.out = None

        def new_eval_state(self):
            return dict(acc=torch.zeros(8 + D, dtype=torch.float64), n_fired=torch.zeros(S), values=torch.zeros(S))

        def eval_accumulate(self, x, state):  # include/saev_b200.h: saev_b200_eval_accumulate
            o, acc, B = self.out, state["acc"], x.shape[0]
            x64 = x.double()
            acc[0] += (x64 * x64).sum()
            acc[1] += ((o.x_hat - x).double() ** 2).sum()
            acc[4] += float(o.l0) * B
            acc[5] += float(o.l1) * B
            acc[6] += float(o.mse) * B
            acc[7] += B
            acc[8:] += x64.sum(0)
            state["n_fired"] += (o.f > 0).sum(0)
            state["values"] += o.f.sum(0)

    class SAE(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.W_dec = torch.nn.Parameter(W_dec.clone())
            self.engine = Engine()

    class Objective(torch.nn.Module):
        def forward(self, sae, x):
            assert not self.training and not sae.training  # train.py:526-527
            sae.engine.out = orc.eval_forward(ocfg, st, x)
            return None, None

    cfg = types.SimpleNamespace(val_data="val", n_val=70, device="cpu", log_every=1)
    (m,) = ev.evaluate([cfg], [SAE()], [Objective()], loader_cls=Loader)
    assert Loader.made == 1 and Loader.closed
    ref = orc.evaluate(ocfg, st, [data[0:32], data[32:64], data[64:96]])  # BatchLimiter stops once >= 70 rows were seen
    for k in ("l0", "l1", "mse", "normalized_mse", "sse_sae", "sse_baseline"):
        assert getattr(m, k) == pytest.approx(ref[k], rel=1e-6), k
    assert (m.n_dead, m.n_almost_dead, m.n_dense) == (ref["n_dead"], ref["n_almost_dead"], ref["n_dense"])
    assert torch.equal(m.freqs, ref["freqs"])
    fired = ref["freqs"] > 0
    assert torch.allclose(m.mean_values[fired], ref["mean_values"][fired], rtol=1e-5)
    with pytest.raises(ValueError):
        ev.evaluate([cfg, types.SimpleNamespace(val_data="other", n_val=1)], [SAE()] * 2, [Objective()] * 2, loader_cls=Loader)


# ---- shard writer side (saev shards.py:112-135, 372-527) ----------------------------------------
def test_shard_writer_reproduces_the_reference_writers_directory(tmp_path):
    """The golden shard directory was written by saev's own Metadata.dump + ShardWriter (oracle/gen_golden_host.py)
    from tests/golden/shards_acts.npy with three write_batch calls; ours must produce the same directory name (the
    metadata hash), the same shards.json and byte-identical acts%06d.bin files."""
    md = data.Metadata.load(SHARDS)
    assert md.hash == SHARDS.name
    acts = torch.from_numpy(np.load(GOLDEN / "shards_acts.npy"))
    root = tmp_path / "saev" / "shards"
    root.mkdir(parents=True)
    md.dump(root)
    with data.ShardWriter(root, md) as w:
        w.write_batch(acts[:3], 0)
        w.write_batch(acts[3:9], 3)
        w.write_batch(acts[9:], 9)
    out = root / md.hash
    assert json.loads((out / "metadata.json").read_text()) == json.loads((SHARDS / "metadata.json").read_text())
    assert json.loads((out / "shards.json").read_text()) == json.loads((SHARDS / "shards.json").read_text())
    bins = sorted(p.name for p in SHARDS.glob("acts*.bin"))
    assert sorted(p.name for p in out.glob("acts*.bin")) == bins
    for name in bins:
        assert (out / name).read_bytes() == (SHARDS / name).read_bytes(), name


def test_shard_writer_rolls_over_like_the_reference_when_a_batch_fills_a_shard_exactly(tmp_path):
    """shards.py:434 uses `>=`: data that exactly fills the last shard leaves a trailing all-zero shard with
    n_examples 0; readers skip it through shards.json."""
    md = data.Metadata(family="fake-clip", ckpt="synthetic", layers=(0,), content_tokens_per_example=4, cls_token=False,
                       d_model=8, n_examples=6, max_tokens_per_shard=3 * 4, data="", dataset="fake")
    assert md.examples_per_shard == 3
    root = tmp_path / "saev" / "shards"
    root.mkdir(parents=True)
    md.dump(root)
    acts = torch.randn(6, 1, 4, 8)
    with data.ShardWriter(root, md) as w:
        w.write_batch(acts, 0)
    info = json.loads((root / md.hash / "shards.json").read_text())
    assert [e["n_examples"] for e in info] == [3, 3, 0]
    got = np.fromfile(root / md.hash / "acts000001.bin", dtype=np.float32).reshape(md.shard_shape)
    assert np.array_equal(got, acts[3:].numpy())


# ---- checkpoints interoperate with the reference (modeling.py:448-658) ---------------------------------------------
def _ref_modeling():
    from oracle import ref_harness

    try:
        ref_harness.import_reference()
    except ImportError:
        pytest.skip("reference package not available (neither /root/reference nor oracle/_ref)")
    import saev.nn.modeling as M

    return M


@pytest.mark.parametrize("act", ["topk", "relu", "batchtopk"])
def test_checkpoint_header_is_the_references_schema_5(act, tmp_path):
    from saev_b200 import nn as bnn

    M = _ref_modeling()
    mk = {"topk": lambda m: m.TopK(top_k=8, aux=m.AuxK(k_aux=64, alpha=0.125)),
          "relu": lambda m: m.Relu(sparsity=m.L1Sparsity(coeff=3e-4), aux=m.NoAux()),
          "batchtopk": lambda m: m.BatchTopK(top_k=4)}[act]
    ours = bnn.SparseAutoencoder(bnn.SparseAutoencoderConfig(d_model=16, d_sae=64, activation=mk(bnn), reinit_blend=0.0))
    theirs = M.SparseAutoencoder(M.SparseAutoencoderConfig(d_model=16, d_sae=64, activation=mk(M), reinit_blend=0.0))
    # 1. same header payload for the activation as the reference's serializer produces
    assert bnn._serialize_dataclass(ours.cfg.activation) == M._serialize_dataclass(theirs.cfg.activation)
    if act == "batchtopk":  # the EMA threshold is a registered buffer on both sides (modeling.py:213)
        assert list(ours.state_dict()) == list(theirs.state_dict())
        assert "activation.threshold" in ours.state_dict()
    # 2. ours -> reference loader
    bnn.dump(tmp_path / "ours.pt", ours)
    back = M.load(tmp_path / "ours.pt")
    assert type(back.cfg.activation).__name__ == type(theirs.cfg.activation).__name__
    assert back.cfg.activation == mk(M)
    for k, v in ours.state_dict().items():
        assert torch.equal(back.state_dict()[k], v), k
    # 3. reference -> our loader
    M.dump(tmp_path / "theirs.pt", theirs)
    mine = bnn.load(tmp_path / "theirs.pt")
    assert mine.cfg.activation == mk(bnn)
    assert mine.cfg.d_sae == 64 and mine.cfg.reinit_blend == 0.0
    for k, v in theirs.state_dict().items():
        assert torch.equal(mine.state_dict()[k], v), k


def test_fused_adam_registers_only_full_sae_groups():
    """clip_grad_norm_ may defer the clip scale to the optimizer only when a FusedAdam owns exactly the four parameters
    of the SAE (train.py:292-306: cfg.optim == "muon" hands the matrices to torch.optim.Muon and only the biases to
    Adam -- then the clip has to be applied to .grad at once)."""
    from saev_b200 import nn as bnn

    cfg = bnn.SparseAutoencoderConfig(d_model=8, d_sae=16, activation=bnn.TopK(top_k=2), reinit_blend=0.0)
    a, b = bnn.SparseAutoencoder(cfg), bnn.SparseAutoencoder(cfg)
    opt_a = optim.FusedAdam([{"params": a.parameters(), "lr": 0.0}])
    assert a._fused_adam is not None and a._fused_adam() is opt_a
    optim.FusedAdam([{"params": [b.b_enc, b.b_dec], "lr": 0.0}])  # the Muon split: biases only
    assert b._fused_adam is None


def test_batchtopk_maps_to_a_row_capacity_on_the_sparse_path(monkeypatch):
    """BatchTopK(top_k = k) runs as TopK lists of `capacity` slots per row + a batch-wide selection of k * B entries
    (engine_config): capacity = min(128, d_sae), SAEV_B200_BATCHTOPK_CAP lowers it, and an average that does not fit
    raises up front."""
    from saev_b200 import nn as bnn

    monkeypatch.delenv("SAEV_B200_BATCHTOPK_CAP", raising=False)
    cfg = bnn.SparseAutoencoderConfig(d_model=64, d_sae=4096, activation=bnn.BatchTopK(top_k=32, momentum=0.2))
    ec = bnn.engine_config(cfg, bnn.Matryoshka(n_prefixes=10), 512)
    assert (ec.activation, ec.top_k, ec.batch_k, ec.batch_momentum, ec.max_prefixes) == ("topk", 128, 32, 0.2, 10)
    assert bnn.batch_topk_capacity(8, 96) == 96  # a dictionary narrower than the capacity: every column is a slot
    monkeypatch.setenv("SAEV_B200_BATCHTOPK_CAP", "16")
    assert bnn.batch_topk_capacity(8, 4096) == 16
    with pytest.raises(NotImplementedError, match="cannot even hold the average"):
        bnn.batch_topk_capacity(32, 4096)
    monkeypatch.delenv("SAEV_B200_BATCHTOPK_CAP")
    with pytest.raises(NotImplementedError):
        bnn.batch_topk_capacity(200, 4096)
    # plain TopK and Relu are untouched by the BatchTopK fields
    ec = bnn.engine_config(bnn.SparseAutoencoderConfig(d_model=64, d_sae=4096, activation=bnn.TopK(top_k=32)),
                           bnn.Matryoshka(n_prefixes=1), 512)
    assert (ec.top_k, ec.batch_k) == (32, 0)
    ec = bnn.engine_config(bnn.SparseAutoencoderConfig(d_model=64, d_sae=4096, activation=bnn.Relu()),
                           bnn.Matryoshka(n_prefixes=4), 512)
    assert (ec.activation, ec.batch_k, ec.max_prefixes) == ("relu", 0, 4)
