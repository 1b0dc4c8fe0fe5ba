"""GPU tests of the shuffled loader (run with `-m gpu`): the census recorded from saev's own ShuffledDataLoader on
the golden shard directory (tests/golden/shards_census.json, written by oracle/gen_golden_host.py) must be
reproduced row for row -- every (example, content token) exactly once per epoch, same batch sizes, same
label filtering -- plus the pool semantics tests/test_reservoir_buffer.py pins for the reference's reservoir
(blocking get with timeout that consumes nothing, exact counters) and a larger synthetic directory that
exercises pool wrap-around, multiple I/O threads, several epochs and rank sharding."""
import json
import pathlib
import zlib

import numpy as np
import pytest
import torch

from saev_b200 import data

pytestmark = pytest.mark.gpu

GOLDEN = pathlib.Path(__file__).resolve().parent / "golden"
CENSUS = json.loads((GOLDEN / "shards_census.json").read_text())
SHARDS = GOLDEN / CENSUS["dir"]


def _rows(batch):
    act = batch["act"].cpu().numpy()
    return [[int(e), int(t), zlib.crc32(a.tobytes())] for a, e, t in
            zip(act, batch["example_idx"].cpu().tolist(), batch["token_idx"].cpu().tolist())]


@pytest.mark.parametrize("case", CENSUS["cases"], ids=lambda c: f"layer{c['layer']}-bs{c['batch_size']}")
@pytest.mark.parametrize("n_threads,zero_copy", [(1, False), (3, False), (2, True)])
def test_epoch_census_matches_reference_loader(case, n_threads, zero_copy):
    """(zero_copy=True: the shard files are mmap()ed + registered and copied with strided DMA -- this directory has two
    layers and a [CLS] token, so every example is its own segment; the label-filtered case falls back to staging.)"""
    cfg = data.ShuffledConfig(shards=SHARDS, layer=case["layer"], batch_size=case["batch_size"], n_threads=n_threads,
                              buffer_size=4, seed=3, ignore_labels=case["ignore_labels"], batch_timeout_s=10.0)
    dl = data.ShuffledDataLoader(cfg, chunk_examples=2 if zero_copy else 1, zero_copy=zero_copy)
    for epoch in range(2):  # the loader is re-iterable (BatchLimiter restarts it, scheduling.py:104-106)
        rows, sizes = [], []
        for batch in dl:
            assert batch["act"].is_cuda and batch["act"].dtype == torch.float32
            assert batch["example_idx"].dtype == torch.int32 and batch["token_idx"].dtype == torch.int32
            sizes.append(len(batch["act"]))
            rows += _rows(batch)
        assert sizes == case["batch_sizes"]
        assert sorted(rows) == case["rows"]
    dl.shutdown()


def _write_dir(tmp_path, n_examples, T, D, ex_per_shard, n_layers=1, cls=False, seed=0):
    d = tmp_path / "saev" / "shards" / "deadbeef"
    d.mkdir(parents=True)
    tokens = T + int(cls)
    md = dict(family="fake-clip", ckpt="synthetic", layers=list(range(n_layers)), content_tokens_per_example=T,
              cls_token=cls, d_model=D, n_examples=n_examples, max_tokens_per_shard=ex_per_shard * tokens * n_layers,
              data="", dataset="fake", pixel_agg="majority", dtype="float32", protocol="2.1")
    (d / "metadata.json").write_text(json.dumps(md))
    rng = np.random.default_rng(seed)
    acts = rng.standard_normal((n_examples, n_layers, tokens, D), dtype=np.float32)
    info = []
    for s in range(-(-n_examples // ex_per_shard)):
        part = acts[s * ex_per_shard : (s + 1) * ex_per_shard]
        full = np.zeros((ex_per_shard, n_layers, tokens, D), dtype=np.float32)  # shards are fixed-size, zero padded
        full[: len(part)] = part
        full.tofile(d / f"acts{s:06d}.bin")
        info.append({"name": f"acts{s:06d}.bin", "n_examples": len(part)})
    (d / "shards.json").write_text(json.dumps(info))
    return d, acts


def test_large_directory_exactly_once_with_small_pool(tmp_path):
    n_examples, T, D = 301, 16, 64
    d, acts = _write_dir(tmp_path, n_examples, T, D, ex_per_shard=37)
    cfg = data.ShuffledConfig(shards=d, layer=0, batch_size=256, n_threads=4, buffer_size=3, seed=11)
    dl = data.ShuffledDataLoader(cfg, chunk_examples=5)
    assert dl.n_samples == n_examples * T and len(dl) == -(-n_examples * T // 256)
    seen = np.zeros((n_examples, T), dtype=np.int64)
    order = []
    n_batches = 0
    for batch in dl:
        ex, tok = batch["example_idx"].cpu().numpy(), batch["token_idx"].cpu().numpy()
        np.testing.assert_array_equal(batch["act"].cpu().numpy(), acts[ex, 0, tok])  # rows carry the right payload
        np.add.at(seen, (ex, tok), 1)
        order += (ex * T + tok).tolist()
        n_batches += 1
        assert dl.manager_pid > 0 and 0.0 <= dl.reservoir.fill() <= 1.0
    assert n_batches == len(dl)
    assert (seen == 1).all()
    assert order != sorted(order)
    # consecutive rows of one example must be spread out: fewer than 5% of neighbours in a batch are file neighbours
    adjacent = np.mean(np.abs(np.diff(np.asarray(order))) == 1)
    assert adjacent < 0.05
    dl.shutdown()


def test_early_break_then_restart(tmp_path):
    """make_saes() breaks out of the loader after the datapoint-init rows (train.py:147-153) and train() then
    iterates the same loader again."""
    d, _ = _write_dir(tmp_path, 64, 8, 32, ex_per_shard=16)
    dl = data.ShuffledDataLoader(data.ShuffledConfig(shards=d, layer=0, batch_size=100, n_threads=2, buffer_size=2))
    it = iter(dl)
    first = next(it)["act"].clone()
    it.close()
    total = sum(len(b["act"]) for b in dl)
    assert total == 64 * 8 and first.shape == (100, 32)
    dl.shutdown()


def test_two_ranks_read_disjoint_shards(tmp_path):
    d, _ = _write_dir(tmp_path, 96, 4, 32, ex_per_shard=16)  # 6 equal shards
    cfg = data.ShuffledConfig(shards=d, layer=0, batch_size=64, n_threads=2, buffer_size=2, seed=5)
    got = []
    for r in range(2):
        dl = data.ShuffledDataLoader(cfg, rank=r, world_size=2)
        keys = set()
        for b in dl:
            keys |= set((b["example_idx"].cpu().numpy().astype(np.int64) * 4 + b["token_idx"].cpu().numpy()).tolist())
        got.append(keys)
        assert len(keys) == dl.n_samples == 96 * 4 // 2
        dl.shutdown()
    assert got[0].isdisjoint(got[1]) and len(got[0] | got[1]) == 96 * 4


def test_next_without_producers_reports_exhausted_not_a_hang(tmp_path):
    """The reference's reservoir get() blocks (with a timeout) while producers may still deliver
    (tests/test_reservoir_buffer.py:173-221, 375-410); with no epoch running the native call must return at once
    with n_rows = 0, and a following epoch still delivers everything."""
    import ctypes as C

    d, _ = _write_dir(tmp_path, 8, 4, 32, ex_per_shard=8)
    dl = data.ShuffledDataLoader(data.ShuffledConfig(shards=d, layer=0, batch_size=16, n_threads=1, buffer_size=2))
    dl._ensure_native()
    lib, h = dl._lib, dl._h
    act, ex, tok, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int32(-1)
    stream = torch.cuda.current_stream().cuda_stream
    # before an epoch was started there are no producers: the epoch reads as exhausted, not as a hang
    assert lib.saev_b200_loader_next(h, stream, 0.05, C.byref(act), C.byref(ex), C.byref(tok), C.byref(n)) == 0
    assert n.value == 0
    total = sum(len(b["act"]) for b in dl)
    assert total == 32
    dl.shutdown()
