"""Generate the HOST-side golden fixtures by running the LIVE reference (read-only at /root/reference/src).
TEST INFRASTRUCTURE ONLY; runs in the build container, the outputs are committed.

    python oracle/gen_golden_host.py

Writes
  tests/golden/host_schedules.json         saev.utils.scheduling.WarmupCosine / Warmup value sequences and
                                           BatchLimiter traces (scheduling.py:21-122)
  tests/golden/shards/saev/shards/<hash>/  a tiny shard directory written by the reference's own
                                           shards.ShardWriter (metadata.json, shards.json, acts%06d.bin,
                                           labels.bin) -- pins the on-disk format our loader reads
  tests/golden/shards_census.json          what the reference's ShuffledDataLoader delivers for that directory
                                           in one epoch: n_samples, len(), batch sizes, and the sorted list of
                                           (example_idx, token_idx, crc32 of the row) -- the exactly-once census
"""

import base64
import json
import pathlib
import pickle
import shutil
import sys
import zlib

import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / "ref_stubs"))
sys.path.insert(0, "/root/reference/src")

import saev.data  # noqa: E402
import saev.data.datasets  # noqa: E402
import saev.data.shards as shards  # noqa: E402
import saev.utils.scheduling as sched  # noqa: E402

GOLDEN = HERE.parent / "tests" / "golden"


def schedules():
    out = {"warmup_cosine": [], "warmup": [], "batch_limiter": []}
    for init, n_warmup, peak, n_steps, final, n_calls in [
        (0.0, 500, 4e-4, 10_000, 0.0, 40), (0.0, 4, 1e-2, 12, 0.0, 16), (0.1, 100, 0.9, 1000, 0.0, 1100),
        (0.0, 0, 1e-3, 5, 1e-5, 8), (0.0, 3, 1e-3, 3, 0.0, 6),
    ]:
        s = sched.WarmupCosine(init, n_warmup, peak, n_steps, final)
        out["warmup_cosine"].append({"args": [init, n_warmup, peak, n_steps, final], "repr": repr(s),
                                     "values": [s.step() for _ in range(n_calls)]})
    for init, final, n_steps, n_calls in [(0.0, 1.0, 10, 14), (0.5, 0.1, 3, 5)]:
        s = sched.Warmup(init, final, n_steps)
        out["warmup"].append({"args": [init, final, n_steps], "repr": repr(s), "values": [s.step() for _ in range(n_calls)]})

    class FakeLoader:
        def __init__(self, sizes, batch_size, drop_last):
            self.sizes, self.batch_size, self.drop_last = sizes, batch_size, drop_last
            self.extra_attr = "passthrough"

        def __iter__(self):
            for n in self.sizes:
                yield {"act": torch.zeros(n, 2)}

    for sizes, bs, drop_last, n_samples in [
        ([4, 4, 2], 4, False, 10), ([4, 4, 2], 4, False, 25), ([4, 4, 2], 4, False, 3), ([8, 8], 8, True, 40),
        ([16], 16, False, 16), ([5, 5, 5, 1], 5, False, 33),
    ]:
        bl = sched.BatchLimiter(FakeLoader(sizes, bs, drop_last), n_samples)
        seen = [len(b["act"]) for b in bl]
        out["batch_limiter"].append({"sizes": sizes, "batch_size": bs, "drop_last": drop_last, "n_samples": n_samples,
                                     "len": len(bl), "yielded": seen, "n_seen": bl.n_seen, "extra_attr": bl.extra_attr})
    (GOLDEN / "host_schedules.json").write_text(json.dumps(out, indent=1))


def write_shards():
    root = GOLDEN / "shards" / "saev" / "shards"
    if root.exists():
        shutil.rmtree(root)
    root.mkdir(parents=True)
    n_examples, T, D, layers = 10, 5, 8, (0, 3)
    md = shards.Metadata(
        family="fake-clip", ckpt="synthetic", layers=layers, content_tokens_per_example=T, cls_token=True, d_model=D,
        n_examples=n_examples, max_tokens_per_shard=4 * (T + 1) * len(layers),  # -> 4 examples per shard, 3 shards
        data=base64.b64encode(pickle.dumps(saev.data.datasets.FakeImg(n_examples=n_examples))).decode("utf8"),
        dataset=pathlib.Path("fake"),
    )
    assert md.examples_per_shard == 4 and md.n_shards == 3
    md.dump(root)
    gen = torch.Generator().manual_seed(7)
    acts = torch.randn(n_examples, len(layers), T + 1, D, generator=gen)
    with shards.ShardWriter(root, md) as w:
        w.write_batch(acts[:3], 0)
        w.write_batch(acts[3:9], 3)
        w.write_batch(acts[9:], 9)
    d = root / md.hash
    # labels.bin: uint8 [n_examples, content_tokens_per_example] (shuffled.py:474-483); background = 0
    labels = (torch.arange(n_examples * T).reshape(n_examples, T) % 3).to(torch.uint8).numpy()
    labels.tofile(d / "labels.bin")
    np.save(GOLDEN / "shards_acts.npy", acts.numpy())
    return d, md


def census(d, md):
    out = {"dir": str(d.relative_to(GOLDEN)), "cases": []}
    for layer, bs, ignore in [(0, 4, []), (3, 7, []), (3, 4, [0])]:
        cfg = saev.data.ShuffledConfig(shards=d, layer=layer, batch_size=bs, n_threads=2, buffer_size=4, seed=3,
                                       ignore_labels=ignore, batch_timeout_s=10.0)
        dl = saev.data.ShuffledDataLoader(cfg)
        rows, sizes = [], []
        for batch in dl:
            sizes.append(len(batch["act"]))
            assert batch["act"].dtype == torch.float32 and batch["example_idx"].dtype == torch.int32
            for a, e, t in zip(batch["act"], batch["example_idx"], batch["token_idx"]):
                rows.append([int(e), int(t), zlib.crc32(a.numpy().tobytes())])
        out["cases"].append({"layer": layer, "batch_size": bs, "ignore_labels": ignore, "n_samples": dl.n_samples,
                             "len": len(dl), "batch_sizes": sizes, "rows": sorted(rows)})
        dl.shutdown()
    (GOLDEN / "shards_census.json").write_text(json.dumps(out))


if __name__ == "__main__":
    schedules()
    d, md = write_shards()
    census(d, md)
    print("wrote host goldens under", GOLDEN)
