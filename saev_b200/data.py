"""Host-side mirror of `saev.data` for the training hot path: `ShuffledConfig`, `Metadata`, `ShuffledDataLoader`.

Same constructor, attributes and batch schema as the reference loader
(/root/reference/src/saev/data/shuffled.py:380-699): iterating yields dicts

    {"act": float32[B, d_model], "example_idx": int32[B], "token_idx": int32[B]}

with every (example, content token) of the requested layer delivered exactly once per epoch in a shuffled order
and a short last batch (`need = min(batch_size, remaining)`, shuffled.py:513).  The difference is where the work
happens: the shuffle pool is a device-resident buffer filled from pinned staging chunks by native I/O threads
(saev_b200/csrc/shard_loader.cu) and the batch tensors are CUDA tensors, so `batch["act"].to(device,
non_blocking=True)` (train.py:333) is a no-op.  By default every batch is an OWNED tensor (one device-to-device copy out
of the loader's ring of batch buffers, ~20 us for 67 MB), valid for as long as the caller keeps it -- saev's
`make_saes` holds four or more batches before concatenating them (train.py:147-160).  `alias_ring=True` hands out
zero-copy views of the ring instead; such a batch is only valid until the iteration has advanced twice more.

Multi-GPU: rank r of W reads the shards `order[r::W]` of the seeded shard permutation (disjoint files per rank)
and every rank delivers the same number of rows per epoch (the minimum over ranks), so data-parallel steps stay
in lock step.

There is no host-only mode: constructing the loader without a CUDA device raises.
"""

from __future__ import annotations

import ctypes as C
import dataclasses
import json
import math
import os
import pathlib
import typing as tp

import numpy as np
import torch

from . import _lib


@dataclasses.dataclass(frozen=True)
class ShuffledConfig:
    """Mirror of saev.data.shuffled.Config (shuffled.py:31-71); same field names and defaults."""

    shards: pathlib.Path = pathlib.Path("$SAEV_SCRATCH/saev/shards/abcdefg")
    tokens: str = "content"
    layer: int | str = -1
    batch_size: int = 1024 * 16
    drop_last: bool = False
    scale_norm: bool = False
    ignore_labels: list = dataclasses.field(default_factory=list)
    n_threads: int = 4
    buffer_size: int = 64
    min_buffer_fill: float = 0.0
    batch_timeout_s: float = 30.0
    seed: int = 17
    debug: bool = False
    log_every_s: float = 30.0
    use_tmpdir: bool = False


Config = ShuffledConfig


@dataclasses.dataclass(frozen=True, kw_only=True)
class Metadata:
    """The fields of saev's metadata.json the loader needs (shards.py:43-180).  When saev itself is importable
    the loader returns saev's own `Metadata` object instead, so `dataclasses.asdict(loader.metadata)` in
    train.py:267 sees the type it expects."""

    family: str
    ckpt: str
    layers: tuple
    content_tokens_per_example: int
    cls_token: bool
    d_model: int
    n_examples: int
    max_tokens_per_shard: int
    data: str = ""
    dataset: str = ""
    pixel_agg: str = "majority"
    dtype: str = "float32"
    protocol: str = "2.1"

    @classmethod
    def load(cls, shards_dir) -> "Metadata":
        with open(pathlib.Path(shards_dir) / "metadata.json") as fd:
            d = json.load(fd)
        d["layers"] = tuple(d["layers"])
        known = {f.name for f in dataclasses.fields(cls)}
        return cls(**{k: v for k, v in d.items() if k in known})

    def _json_dict(self) -> dict:
        d = dataclasses.asdict(self)
        d["layers"] = list(self.layers)
        d["dataset"] = str(self.dataset)
        return d

    @property
    def hash(self) -> str:
        """shards.py:126-135: first 8 hex digits of the SHA-256 of the sorted-key compact JSON of the metadata (what
        orjson.dumps(..., OPT_SORT_KEYS) produces), i.e. the name of the shard directory."""
        import hashlib

        blob = json.dumps(self._json_dict(), sort_keys=True, separators=(",", ":"), ensure_ascii=False).encode("utf8")
        return hashlib.sha256(blob).hexdigest()[:8]

    def dump(self, shards_root) -> None:
        """shards.py:112-124: write metadata.json into shards_root / hash."""
        shards_root = pathlib.Path(shards_root)
        assert shards_root.parts[-2:] == ("saev", "shards"), f"'{shards_root}' must end in saev/shards (disk.py:29-41)"
        (shards_root / self.hash).mkdir(exist_ok=True, parents=True)
        with open(shards_root / self.hash / "metadata.json", "w") as fd:
            json.dump(self._json_dict(), fd, indent=2)
            fd.write("\n")

    @property
    def tokens_per_example(self) -> int:
        return self.content_tokens_per_example + int(self.cls_token)

    @property
    def examples_per_shard(self) -> int:
        return self.max_tokens_per_shard // (self.tokens_per_example * len(self.layers))

    @property
    def n_shards(self) -> int:
        return math.ceil(self.n_examples / self.examples_per_shard)

    @property
    def shard_shape(self) -> tuple:
        return (self.examples_per_shard, len(self.layers), self.tokens_per_example, self.d_model)


class ShardWriter:
    """Writer side of the shard format the loader reads (mirror of saev's ShardWriter, shards.py:372-527): fp32
    memmaps `acts%06d.bin` of shape [examples_per_shard, n_layers, tokens_per_example, d_model] plus `shards.json`
    with the number of valid examples per file.  Same call protocol:

        md.dump(shards_root)
        with ShardWriter(shards_root, md) as w:
            w.write_batch(acts[n, n_layers, tokens, d_model], start_idx)

    `write_batch` accepts CPU or CUDA tensors (CUDA batches come down through one reused pinned buffer).  Like the
    reference, a batch that exactly fills a shard rolls over to a new (empty, all-zero) one (`>=` at shards.py:434),
    which readers skip through its `n_examples: 0` entry."""

    def __init__(self, shards_root, md: Metadata):
        shards_root = pathlib.Path(shards_root)
        assert shards_root.parts[-2:] == ("saev", "shards"), f"'{shards_root}' must end in saev/shards"
        self.md = md
        self.shards_dir = shards_root / md.hash
        self.shards_dir.mkdir(exist_ok=True, parents=True)
        self._shards: list[dict] = []
        self._pinned = None
        self.shard = -1
        self.acts = None
        self.filled = 0
        self.next_shard()

    def _host(self, t) -> np.ndarray:
        if isinstance(t, np.ndarray):
            return t
        if t.is_cuda:
            if self._pinned is None or self._pinned.numel() < t.numel():
                self._pinned = torch.empty(t.numel(), dtype=torch.float32).pin_memory()
            dst = self._pinned[: t.numel()].view(t.shape)
            dst.copy_(t.to(torch.float32), non_blocking=True)
            torch.cuda.current_stream(t.device).synchronize()
            return dst.numpy()
        return t.detach().to(torch.float32).numpy()

    def write_batch(self, activations, start_idx: int, patch_labels=None) -> None:
        if patch_labels is not None:
            raise NotImplementedError("saev_b200.data.ShardWriter does not write labels.bin (segmentation labels)")
        batch_size = len(activations)
        end_idx = start_idx + batch_size
        eps = self.md.examples_per_shard
        offset = eps * self.shard
        if end_idx >= offset + eps:  # shards.py:434
            n_fit = offset + eps - start_idx
            self.acts[start_idx - offset : start_idx - offset + n_fit] = self._host(activations[:n_fit])
            self.filled = start_idx - offset + n_fit
            self.next_shard()
            if n_fit < batch_size:
                self.write_batch(activations[n_fit:], start_idx + n_fit)
        else:
            assert 0 <= start_idx - offset and end_idx - offset <= eps, (start_idx, end_idx, offset, eps)
            self.acts[start_idx - offset : end_idx - offset] = self._host(activations)
            self.filled = end_idx - offset

    def flush(self) -> None:
        if self.acts is not None:
            self.acts.flush()
            self._shards.append({"name": os.path.basename(self.acts_path), "n_examples": int(self.filled)})
            with open(self.shards_dir / "shards.json", "w") as fd:
                json.dump(self._shards, fd, indent=2)
        self.acts = None

    def next_shard(self) -> None:
        self.flush()
        self.shard += 1
        self.acts_path = self.shards_dir / f"acts{self.shard:06}.bin"
        self.acts = np.memmap(self.acts_path, mode="w+", dtype=np.float32, shape=self.md.shard_shape)
        self.filled = 0

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.flush()


def _load_metadata(shards_dir: pathlib.Path):
    import sys

    if "saev.data.shards" in sys.modules:
        return sys.modules["saev.data.shards"].Metadata.load(shards_dir)
    return Metadata.load(shards_dir)


def _load_shard_examples(shards_dir: pathlib.Path) -> list[tuple[str, int]]:
    """shards.json: [{"name": "acts000000.bin", "n_examples": n}, ...] (shards.py:575-636)."""
    fpath = shards_dir / "shards.json"
    if not fpath.exists():
        raise FileNotFoundError(f"shards.json not found in '{shards_dir}' (expected next to metadata.json)")
    with open(fpath) as fd:
        return [(e["name"], int(e["n_examples"])) for e in json.load(fd)]


def _dist_rank_world() -> tuple[int, int]:
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class _PoolView:
    """What `loader.reservoir` exposes to saev's DataloaderMonitor (monitoring.py:74-81): fill(), qsize(),
    capacity — read from the native loader's counters."""

    def __init__(self, loader: "ShuffledDataLoader"):
        self._loader = loader

    def _stats(self):
        a, b, c, d = (C.c_int64() for _ in range(4))
        self._loader._lib.saev_b200_loader_stats(self._loader._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return a.value, b.value, c.value, d.value

    @property
    def capacity(self) -> int:
        return self._stats()[1]

    def qsize(self) -> int:
        return self._stats()[0]

    def fill(self) -> float:
        rows, cap, _, _ = self._stats()
        return rows / cap if cap else 0.0

    def close(self) -> None:
        pass


class ShuffledDataLoader:
    """Streaming shuffled loader over saev activation shards; see the module docstring."""

    def __init__(self, cfg, *, device: torch.device | str | None = None, rank: int | None = None,
                 world_size: int | None = None, n_out_slots: int = 3, chunk_examples: int = 0,
                 alias_ring: bool = False, zero_copy: bool | None = None):
        """`zero_copy`: None = automatic (shards on tmpfs are mmap()ed + registered with the driver and copied to the
        GPU straight out of the page cache; anything else is pread into pinned staging chunks), True / False force
        it.  Label filtering always takes the staging path."""
        self.cfg = cfg
        self.alias_ring = alias_ring
        self._zero_copy = {None: 0, False: 1, True: 2}[zero_copy]
        self._h = None
        self._lib = None
        self.reservoir = None
        shards_dir = pathlib.Path(os.path.expandvars(str(cfg.shards)))
        if not os.path.isdir(shards_dir):
            raise RuntimeError(f"Activations are not saved at '{cfg.shards}'.")  # shuffled.py:400-401
        if cfg.scale_norm:
            raise NotImplementedError("scale_norm not implemented.")  # shuffled.py:409-410
        self._shards_path = shards_dir
        self.metadata = _load_metadata(shards_dir)
        md = self.metadata
        if cfg.tokens != "content" or not isinstance(cfg.layer, int):
            # shuffled.py:300-304: the reference's manager raises the same way
            raise NotImplementedError("High-throughput loader only supports `content` and fixed `layer` mode for now.")
        if cfg.layer not in md.layers:
            raise ValueError(f"Layer {cfg.layer} not in {md.layers}")
        info = _load_shard_examples(shards_dir)
        missing = [n for n, _ in info if not (shards_dir / n).is_file() or (shards_dir / n).stat().st_size == 0]
        if missing:
            raise FileNotFoundError(f"Shard validation failed in '{shards_dir}': missing or empty {missing[:5]}")
        self._shard_examples_all = [n for _, n in info]
        self._labels = None
        self._ignore_lut = None
        if cfg.ignore_labels:
            labels_path = shards_dir / "labels.bin"
            if not labels_path.exists():
                raise FileNotFoundError(
                    f"ignore_labels filtering requested but labels.bin not found at {labels_path}")
            self._labels = np.memmap(labels_path, mode="r", dtype=np.uint8,
                                     shape=(md.n_examples, md.content_tokens_per_example))
            lut = np.zeros(256, dtype=np.uint8)
            lut[np.asarray(list(cfg.ignore_labels), dtype=np.int64)] = 1
            self._ignore_lut = lut
        r, w = _dist_rank_world()
        self.rank = r if rank is None else rank
        self.world_size = w if world_size is None else world_size
        self.device = torch.device(device) if device is not None else None
        self._n_out_slots = n_out_slots
        self._chunk_examples = chunk_examples
        # 1. global shuffle of the shard list (shuffled.py:326-328), then this rank's stride
        rng = np.random.default_rng(cfg.seed)
        order = rng.permutation(md.n_shards)
        self._orders = [order[k :: self.world_size] for k in range(self.world_size)]
        self._rows_per_rank = [self._count_rows(o) for o in self._orders]
        self._n_samples = min(self._rows_per_rank) if self.world_size > 1 else self._rows_per_rank[0]
        self._epoch = 0
        self._iterating = False

    # ---- sizes --------------------------------------------------------------------------------
    def _count_rows(self, order) -> int:
        md = self.metadata
        T = md.content_tokens_per_example
        if self._labels is None:
            return int(sum(self._shard_examples_all[s] for s in order)) * T
        total = 0
        for s in order:
            e0 = int(s) * md.examples_per_shard
            lab = self._labels[e0 : e0 + self._shard_examples_all[s]]
            total += int((self._ignore_lut[lab] == 0).sum())
        return total

    @property
    def n_batches(self) -> int:
        return len(self)

    @property
    def n_samples(self) -> int:
        return self._n_samples

    @property
    def batch_size(self) -> int:
        return self.cfg.batch_size

    @property
    def drop_last(self) -> bool:
        return self.cfg.drop_last

    @property
    def manager_pid(self) -> int:
        """saev reports its manager process here; the native I/O threads live in this process."""
        return os.getpid() if self._iterating else -1

    def __len__(self) -> int:
        return math.ceil(self.n_samples / self.cfg.batch_size)

    @property
    def zero_copy(self) -> bool:
        """Whether the native loader copies to the GPU straight out of the (registered) page cache."""
        self._ensure_native()
        return bool(self._lib.saev_b200_loader_zero_copy(self._h))

    # ---- native loader ------------------------------------------------------------------------
    def _ensure_native(self):
        if self._h is not None:
            return
        if not torch.cuda.is_available():
            raise RuntimeError("saev_b200.data.ShuffledDataLoader needs a CUDA device: its shuffle pool lives in HBM "
                               "(there is no host-only mode)")
        if self.device is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._lib = _lib.load()
        md = self.metadata
        order = np.ascontiguousarray(self._orders[self.rank], dtype=np.int32)
        n_ex = np.ascontiguousarray([self._shard_examples_all[s] for s in order], dtype=np.int32)
        self._keep = (order, n_ex, str(self._shards_path).encode())
        c = _lib.LoaderCfg(
            shards_dir=self._keep[2], examples_per_shard=md.examples_per_shard, n_layers=len(md.layers),
            tokens_per_example=md.tokens_per_example, d_model=md.d_model, layer_index=md.layers.index(self.cfg.layer),
            cls_token=int(md.cls_token), content_tokens=md.content_tokens_per_example,
            shard_order=order.ctypes.data_as(C.POINTER(C.c_int32)), shard_examples=n_ex.ctypes.data_as(C.POINTER(C.c_int32)),
            n_order=len(order), batch_size=self.cfg.batch_size, pool_batches=self.cfg.buffer_size,
            n_threads=self.cfg.n_threads, n_out_slots=self._n_out_slots, chunk_examples=self._chunk_examples,
            min_buffer_fill=float(self.cfg.min_buffer_fill), reserved=self._zero_copy, n_rows_limit=-1, seed=int(self.cfg.seed),
            labels=self._labels.ctypes.data if self._labels is not None else None,
            ignore_lut=self._ignore_lut.ctypes.data if self._ignore_lut is not None else None,
        )
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self._lib.saev_b200_loader_create(C.byref(c), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"libsaev_b200 loader error {rc}: {self._lib.saev_b200_loader_last_error(None).decode()}")
        self._h = h
        self.reservoir = _PoolView(self)

    def _wrap(self, ptr: int, n: int, shape, dtype) -> torch.Tensor:
        """Zero-copy CUDA tensor over a loader-owned device buffer (__cuda_array_interface__)."""
        typestr = {torch.float32: "<f4", torch.int32: "<i4"}[dtype]

        class _Buf:
            __cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 3,
                                        "strides": None}

        holder = _Buf()
        holder._owner = self  # the loader outlives the view
        return torch.as_tensor(holder, device=self.device)

    def __iter__(self) -> tp.Iterator[dict]:
        self._ensure_native()
        lib, h = self._lib, self._h
        D = self.metadata.d_model
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = lib.saev_b200_loader_start_epoch(h, self._n_samples, int(self.cfg.seed) + 0x9E3779B9 * self._epoch, stream)
        if rc != 0:
            raise RuntimeError(f"libsaev_b200 loader error {rc}: {lib.saev_b200_loader_last_error(h).decode()}")
        self._epoch += 1
        self._iterating = True
        act, ex, tok, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int32()
        try:
            while True:
                stream = torch.cuda.current_stream(self.device).cuda_stream
                rc = lib.saev_b200_loader_next(h, stream, float(self.cfg.batch_timeout_s), C.byref(act), C.byref(ex),
                                               C.byref(tok), C.byref(n))
                if rc == 131:
                    # shuffled.py:526-548: log and keep waiting while the producer side is alive
                    continue
                if rc != 0:
                    raise RuntimeError(f"loader crashed:\n{lib.saev_b200_loader_last_error(h).decode()}")
                if n.value == 0:
                    return
                batch = {
                    "act": self._wrap(act.value, n.value, (n.value, D), torch.float32),
                    "example_idx": self._wrap(ex.value, n.value, (n.value,), torch.int32),
                    "token_idx": self._wrap(tok.value, n.value, (n.value,), torch.int32),
                }
                if not self.alias_ring:
                    # owned copies, enqueued on the consumer stream right behind the loader's event wait (the ring
                    # slot is not recycled before the stream has passed this point)
                    batch = {k: v.clone() for k, v in batch.items()}
                yield batch
        finally:
            self._iterating = False
            if self._h is h:  # shutdown() may already have destroyed the native loader
                lib.saev_b200_loader_stop(h)

    def shutdown(self) -> None:
        """shuffled.py:555-575: stop the producer side and release the pool."""
        h, self._h = self._h, None
        self._iterating = False
        self.reservoir = None
        if h is not None and self._lib is not None:
            self._lib.saev_b200_loader_destroy(h)

    def __del__(self):
        try:
            self.shutdown()
        except Exception:
            pass
