#!/usr/bin/env python
"""Headline benchmark: activations/sec of the full SAE training step (saev train.py:332-460 loop body:
normalize -> objective forward (encode, TopK, decode, MSE + AuxK) -> backward -> remove_parallel_grads ->
clip -> Adam) on BASELINE.json's quoted configuration (d_model=1024, d_sae=65536, K=32, batch 16384 per GPU).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (N>1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

Prints ONE JSON line (rank 0).  Keys follow the driver contract; see DESIGN.md section 5 (Measurement).

Diagnostics that are NOT part of the headline line's workload (each changes `config` or adds a key, never `value`
of the default run): `torch_gpu_baseline` (N = 1: the dense torch restatement of the reference step on the same GPU,
TF32 as the reference enables it; `--no-torch-gpu-baseline` skips it), `final.check` (the loss of one more forward
recomputed with dense fp32 torch ops on the GPU), `--workload c1|c2|c5`, `--n-prefixes P` (Matryoshka), `--dense-features N` (N atoms
that fire on every row), `--dp-mode`, `--e2e ring|loader`.
"""

from __future__ import annotations

import argparse
import json
import os
import pathlib
import statistics
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (d_model, d_sae, top_k, batch per GPU)      top_k == 0: ReLU activation + L1Sparsity(4e-4) (dense path)
    "c3": (1024, 65536, 32, 16384),
    "c2": (768, 32768, 32, 4096),
    "c1": (128, 512, 16, 256),
    "c5": (1536, 131072, 0, 8192),
}
METRIC = "activations/sec"
FALLBACK_PEAKS = {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            flat = {}

            def walk(o):
                if isinstance(o, dict):
                    for k, v in o.items():
                        if isinstance(v, (int, float)):
                            flat.setdefault(k, float(v))
                        else:
                            walk(v)

            walk(d)
            if "bf16_tflops_sustained" in flat or "bf16_tflops" in flat:
                return flat, "measured"
        except Exception:
            pass
    return dict(FALLBACK_PEAKS), "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [int(s[0]) for s in self.samples if s[0].isdigit()]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": int(statistics.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "n_samples": len(sm)}


def oracle_cfg(orc, D, S, K):
    if K == 0:
        return orc.OracleConfig(d_model=D, d_sae=S, activation="relu", top_k=1, l1_coeff=4e-4, aux=True, k_aux=512,
                                lr=4e-4, n_lr_warmup=500, n_steps=10_000)
    return orc.OracleConfig(d_model=D, d_sae=S, top_k=K, aux=True, k_aux=512, lr=4e-4, n_lr_warmup=500, n_steps=10_000)


def cpu_reference_run(D, S, K, steps, warmup, batch, threads=None):
    """Times the reference's step on the host cores: the reference's OWN modules (saev.nn.SparseAutoencoder, its
    objective, torch Adam(fused=True), stepped with the statements of train.py:334-362, 444-458) when its package is
    staged (oracle/_ref, see oracle/ref_harness.py) -> kind "reference"; otherwise the CPU restatement
    oracle/sae_oracle.py (pinned against the live reference) -> kind "port".
    Returns (acts_per_s, ms_per_step, cores, kind)."""
    import torch

    try:
        from oracle import ref_harness

        rate, ms, cores = ref_harness.time_reference(D, S, max(K, 1), relu=K == 0, rows=batch, steps=steps, warmup=warmup,
                                                     threads=threads)
        return rate, ms, cores, "reference"
    except ImportError:
        pass
    from oracle import sae_oracle as orc

    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)
    W_enc, b_enc, W_dec, b_dec = orc.init_params(D, S, g)
    st = orc.OracleState.from_params(W_enc, b_enc, W_dec, b_dec)
    cfg = oracle_cfg(orc, D, S, K)
    xs = [torch.randn(batch, D, generator=g) for _ in range(2)]
    for i in range(warmup):
        orc.train_step(cfg, st, xs[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        orc.train_step(cfg, st, xs[i % 2])
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, cores, "port"


CPU_ROWS = 1024  # rows per CPU step, the same in the `cpu_baseline` leg of our line and in the `--impl reference` arm


def cpu_baseline_block(D, S, K, B, steps, warmup, rows):
    """The `cpu_baseline` object: the reference step on `rows` rows, plus a second point at rows / 2 that separates the
    per-step fixed cost (normalise, projection, clip, Adam over all 2 D S parameters: independent of the batch) from
    the per-row cost, so that the sample is not mistaken for "cost linear in rows"."""
    rate, ms, cores, kind = cpu_reference_run(D, S, K, steps=steps, warmup=warmup, batch=rows)
    out = {"value": rate, "unit": "activations/s", "cores": cores, "kind": kind, "ms_per_step": ms,
           "sample": f"{steps} timed steps x {rows} rows (+{warmup} warm-up) of the same d_model={D} d_sae={S} "
                     f"{'ReLU' if K == 0 else 'K=' + str(K)} step on torch CPU fp32, "
                     + ("the reference's own saev.nn / objective / Adam(fused) objects (oracle/_ref)" if kind == "reference"
                        else "oracle/sae_oracle.py")}
    try:
        _, ms_half, _, _ = cpu_reference_run(D, S, K, steps=1, warmup=1, batch=rows // 2)
        per_row = max(0.0, (ms - ms_half) / (rows - rows // 2))
        fixed = max(0.0, ms - per_row * rows)
        out["fixed_ms_per_step"] = fixed
        out["ms_per_row"] = per_row
        out["extrapolated_full_batch_value"] = B / ((fixed + per_row * B) * 1e-3) if fixed + per_row * B > 0 else None
    except Exception:  # noqa: BLE001
        pass
    return out


def torch_gpu_reference_run(D, S, K, B, steps, warmup):
    """The same restatement of the reference's step (dense torch ops, the [B, S] matrices materialised, TF32 matmuls
    as the reference enables them, train.py:257) on the SAME GPU: what a saev user gets on a B200 today (SURVEY 8d,
    baseline 2).  Timed with CUDA events after our own timed region; reported beside the CPU baseline, never as
    part of `value`."""
    import torch

    from oracle import sae_oracle as orc

    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    g = torch.Generator().manual_seed(0)
    W_enc, b_enc, W_dec, b_dec = orc.init_params(D, S, g)
    st = orc.OracleState.from_params(W_enc.cuda(), b_enc.cuda(), W_dec.cuda(), b_dec.cuda())
    cfg = oracle_cfg(orc, D, S, K)
    xs = [torch.randn(B, D, generator=g).cuda() for _ in range(2)]
    for i in range(warmup):
        orc.train_step(cfg, st, xs[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        orc.train_step(cfg, st, xs[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del st, xs
    torch.cuda.empty_cache()
    return B / ms * 1e3, ms


def run_reference(args, D, S, K, B, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the step on the host cores (rank 0 only), each
    step a bounded sample of CPU_ROWS rows -- the same sample size as the `cpu_baseline` leg of our own line."""
    if rank != 0:
        return
    rows = args.cpu_rows or CPU_ROWS
    blk = cpu_baseline_block(D, S, K, B, args.steps, args.warmup, rows)
    line = {
        "impl": "reference", "metric": METRIC, "value": blk["value"], "unit": "activations/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": blk["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, D, S, K, B), "cpu_rows_per_step": rows},
        "cpu_baseline": blk,
        "e2e": {"value": blk["value"], "unit": "activations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def final_check(eng, x, K):
    """Parity probe OUTSIDE the timed regions: one eval forward of the trained engine on `x`, and the same loss from
    dense fp32 torch ops (x W_enc + b -> top-k -> decode -> MSE; TF32 off) on the same parameters."""
    import torch

    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        eng.forward(x, training=False)
        ours = float(eng.losses[0])
        sse, n = 0.0, 0
        for r0 in range(0, x.shape[0], 2048):
            xs = x[r0:r0 + 2048]
            h = xs @ eng.W_enc_t.t() + eng.b_enc
            if K:
                v, i = h.topk(K, dim=1)
                xh = torch.zeros_like(xs)
                for c0 in range(0, K, 8):  # gather-decode in slices: [2048, 8, D] temporaries
                    xh += (eng.W_dec[i[:, c0:c0 + 8]] * v[:, c0:c0 + 8, None]).sum(1)
                xh += eng.b_dec
            else:
                xh = torch.relu(h) @ eng.W_dec + eng.b_dec
            sse += float(((xh - xs).double() ** 2).sum())
            n += xs.numel()
        ref = sse / n
        return {"mse_kernels": ours, "mse_torch_fp32": ref, "rel_err": abs(ours - ref) / max(abs(ref), 1e-30)}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32


def _fs_type(path):
    """File-system magic of `path` (statfs f_type), e.g. 0x01021994 = tmpfs."""
    import ctypes

    class _Statfs(ctypes.Structure):
        _fields_ = [("f_type", ctypes.c_long), ("pad", ctypes.c_long * 20)]

    buf = _Statfs()
    libc = ctypes.CDLL(None, use_errno=True)
    if libc.statfs(str(path).encode(), ctypes.byref(buf)) != 0:
        return None
    return buf.f_type & 0xFFFFFFFF


TMPFS_MAGIC = 0x01021994


def make_bench_shards(D, B, n_batches, rank, storage="tmpfs"):
    """Synthetic Gaussian activation shards in saev's on-disk format (fp32 [examples, layers=1, tokens, d_model] +
    metadata.json + shards.json, src/saev/data/shards.py:43-180, 575-636).  storage="tmpfs": under /dev/shm (page
    cache); "disk": on the first candidate directory that is NOT tmpfs, written through and evicted from the page cache
    (fsync + POSIX_FADV_DONTNEED) so that the loader really reads the device."""
    import shutil
    import tempfile

    import numpy as np
    import torch

    T = 256  # tokens per example (ViT-L/14 at 224 px)
    n_examples = -(-n_batches * B // T)
    per_shard = min(n_examples, max(1, (1 << 28) // (T * D * 4)))  # ~256 MB shards
    need = n_examples * T * D * 4
    base = None
    cands = ("/dev/shm", tempfile.gettempdir()) if storage == "tmpfs" else (tempfile.gettempdir(), "/var/tmp", str(ROOT / "gpurun_out"))
    for cand in cands:
        try:
            if not os.path.isdir(cand):
                continue
            if storage == "disk" and _fs_type(cand) == TMPFS_MAGIC:
                continue
            if shutil.disk_usage(cand).free > need * 1.2 + (1 << 30):
                base = cand
                break
        except OSError:
            continue
    if base is None:
        raise RuntimeError(f"no scratch space for the benchmark shards ({storage})")
    root = pathlib.Path(tempfile.mkdtemp(prefix=f"saev_b200_bench_r{rank}_", dir=base)) / "saev" / "shards" / "bench0000"
    root.mkdir(parents=True)
    md = dict(family="fake-clip", ckpt="synthetic", layers=[0], content_tokens_per_example=T, cls_token=False, d_model=D,
              n_examples=n_examples, max_tokens_per_shard=per_shard * T, data="", dataset="synthetic",
              pixel_agg="majority", dtype="float32", protocol="2.1")
    (root / "metadata.json").write_text(json.dumps(md))
    gen = torch.Generator().manual_seed(99 + rank)
    info = []
    for s0 in range(0, n_examples, per_shard):
        n = min(per_shard, n_examples - s0)
        block = torch.randn(per_shard * T, D, generator=gen)  # fixed-size shard files, tail rows unused
        name = f"acts{s0 // per_shard:06d}.bin"
        block.numpy().tofile(root / name)
        if storage == "disk":
            fd = os.open(root / name, os.O_RDONLY)
            try:
                os.fsync(fd)
                os.posix_fadvise(fd, 0, 0, os.POSIX_FADV_DONTNEED)
            finally:
                os.close(fd)
        info.append({"name": name, "n_examples": n})
    (root / "shards.json").write_text(json.dumps(info))
    return root


def loader_leg(tr, eng, dev, world, rank, D, B, steps, n_threads, lr_fn, gstep, barrier, storage, zero_copy=None):
    """`steps` training steps fed by saev_b200.data.ShuffledDataLoader from shard files (`storage`), every step's loss
    vector read back to pinned host memory; returns (activations/s over all ranks, note, bytes/step, steps done)."""
    import shutil

    import torch
    import torch.distributed as dist

    from saev_b200 import data as bdata
    from saev_b200.scheduling import BatchLimiter

    loss_host = torch.empty(steps, 8, dtype=torch.float32).pin_memory()
    n_batches = steps + 6
    shards_root = make_bench_shards(D, B, n_batches, rank, storage)
    try:
        lcfg = bdata.ShuffledConfig(shards=shards_root, layer=0, batch_size=B, n_threads=n_threads, buffer_size=4,
                                    seed=17 + rank, batch_timeout_s=60.0)
        loader = bdata.ShuffledDataLoader(lcfg, device=dev, rank=0, world_size=1, zero_copy=zero_copy)  # own directory per rank
        limiter = BatchLimiter(loader, steps * B + 2 * B)
        it = iter(limiter)
        zc = loader.zero_copy
        for _ in range(2):  # pool fill + first batches are warm-up
            tr.step(next(it)["act"], lr_fn(gstep))
            gstep += 1
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            batch = next(it)
            tr.step(batch["act"], lr_fn(gstep))
            loss_host[i].copy_(eng.losses, non_blocking=True)
            gstep += 1
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        it.close()
        loader.shutdown()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        value = world * B * steps / (float(t.item()) * 1e-3)
        note = (f"ShuffledDataLoader over {n_batches} batches of fp32 shards under {shards_root.parents[3]} "
                f"({'tmpfs' if _fs_type(shards_root) == TMPFS_MAGIC else 'disk, page cache dropped'}; {n_threads} I/O threads, "
                + ("mmap + cudaHostRegister: DMA out of the page cache" if zc else "pread -> pinned staging")
                + " -> HBM pool -> gather)")
        return value, note, gstep, wall
    finally:
        shutil.rmtree(shards_root.parents[2], ignore_errors=True)


def workload_name(w, D, S, K, B, n_prefixes=1, k_aux=512, dead_threshold=10_000_000):
    act = f"TopK k={K}" if K else "ReLU + L1Sparsity(4e-4)"
    mat = f"Matryoshka({n_prefixes} prefixes) " if n_prefixes > 1 else ""
    thr = f"{dead_threshold // 1_000_000}M" if dead_threshold % 1_000_000 == 0 else f"{dead_threshold // 1000}k"
    return (f"{w}: d_model={D} d_sae={S} {act} batch={B}/GPU, objective {mat}MSE+AuxK(k_aux={k_aux}, alpha=1/32, "
            f"dead_threshold={thr} tokens), remove_parallel_grads, clip 1.0, Adam, lr warmup; Gaussian activations")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows per CPU-baseline step (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-gpu-baseline", action="store_true",
                    help="skip timing the dense torch restatement of the reference step on this GPU (TF32; N = 1 only)")
    ap.add_argument("--preheat-s", type=float, default=2.0,
                    help="seconds of untimed steps before the timed region (clocks / power settle; beyond --warmup)")
    ap.add_argument("--no-aux", action="store_true")
    ap.add_argument("--dp-mode", default="auto", choices=["auto", "chunked", "plain", "sharded", "sharded-overlap", "sharded-chunked"],
                    help="gradient exchange for N > 1 (see saev_b200/parallel.py)")
    ap.add_argument("--dense-features", type=int, default=0,
                    help="diagnostic (not the headline workload): give this many atoms a large encoder bias so that they "
                         "fire on every row, like the dense features of a real run")
    ap.add_argument("--n-prefixes", type=int, default=1,
                    help="Matryoshka prefixes (1 = the north-star objective; 10 = saev's default objective)")
    ap.add_argument("--dead-threshold", type=int, default=10_000_000,
                    help="Matryoshka.dead_threshold_tokens (the c5 AuxK sweep: 10M / 1M / 100k)")
    ap.add_argument("--k-aux", type=int, default=512, help="AuxK.k_aux (the c5 AuxK sweep: 256 / 512 / 1024)")
    ap.add_argument("--dead-atoms", type=int, default=0,
                    help="diagnostic: give this many atoms a large negative encoder bias so that they never fire and die "
                         "once dead_threshold_tokens have passed (AuxK then has work to do)")
    ap.add_argument("--gather-ctas", type=int, default=16)
    ap.add_argument("--reserved-sms", type=int, default=16)
    ap.add_argument("--e2e", default="loader", choices=["loader", "ring"], help="end-to-end input path")
    ap.add_argument("--loader-threads", type=int, default=0, help="I/O threads per rank (0 = cores / ranks - 1, 2..8)")
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="N > 1: do not pin each rank to the CPUs local to its GPU (saev_b200/numa.py)")
    ap.add_argument("--no-disk-leg", action="store_true",
                    help="skip the extra end-to-end leg that streams the shards from a non-tmpfs directory (N = 1 only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    D, S, K, B = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, D, S, K, B, rank, world)
        return

    import torch
    import torch.distributed as dist

    from saev_b200 import _lib
    from saev_b200.engine import Engine, EngineConfig
    from saev_b200.parallel import DataParallelTrainer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; saev_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = {"cpus": 0, "how": "not requested"}
    if world > 1 and not args.no_numa_bind:
        from saev_b200.numa import bind_process_to_gpu

        numa = bind_process_to_gpu(local_rank)  # before any pinned allocation / shard file is written
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL kernels on a high-priority stream: the chunked gradient all-reduce then really runs beside the
        # weight-gradient kernel instead of queueing behind its blocks (measured: 0.33 ms/step at 2 GPUs)
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)

    eng = Engine(EngineConfig(d_model=D, d_sae=S, top_k=max(K, 1), activation="topk" if K else "relu",
                              l1_coeff=0.0 if K else 4e-4, aux=not args.no_aux, k_aux=args.k_aux, aux_alpha=1 / 32,
                              dead_threshold_tokens=args.dead_threshold, max_batch=B,
                              max_prefixes=max(1, args.n_prefixes)),
                 device=dev)
    eng.init_params(seed=0)
    if args.dense_features > 0:
        eng.b_enc[:: max(1, S // args.dense_features)][: args.dense_features] = 8.0
    if args.dead_atoms > 0:
        eng.b_enc[1 :: max(1, S // args.dead_atoms)][: args.dead_atoms] = -1e3
    if args.dense_features > 0 or args.dead_atoms > 0:
        eng.sync_weights()
    if K == 0 and world > 1:
        args.dp_mode = "plain"  # the dense (ReLU) path exchanges the whole gradient bucket with one all-reduce
    if args.dp_mode == "auto":
        # measured on B200 (profiles/README.md, round 2): the row-sharded optimizer whose fp32 all-gathers run beside the
        # next step's screen wins at every rank count (N=2: 5.04 vs 5.26 ms chunked all-reduce; N=8: 4.7-4.9 vs 5.2-5.9 ms)
        args.dp_mode = "sharded-overlap"
    gather_group = None
    if world > 1 and args.dp_mode in ("sharded-overlap", "sharded-chunked"):
        gopts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        gopts.config.max_ctas = args.gather_ctas
        gopts.config.min_ctas = 1
        gather_group = dist.new_group(backend="nccl", pg_options=gopts)
    tr = DataParallelTrainer(eng, sharded=args.dp_mode.startswith("sharded"),
                             n_chunks=4 if args.dp_mode in ("chunked", "sharded-chunked") else 1, gather_group=gather_group,
                             reserved_sms=args.reserved_sms)
    tr.broadcast_params(0)

    # synthetic activations: NB rotating batches per rank, resident in HBM.  Many distinct batches on purpose: with only
    # a few the SAE memorises them within the pre-heat (hundreds of steps), its pre-activations pile up just under each
    # row's top-k threshold and the screen admits twice as many candidates as it does on fresh data (measured: 95 vs 45
    # per row) -- a cost a real run, which never sees a batch twice, does not pay.
    NB = 32
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x_dev = [torch.randn(B, D, generator=gen, device=dev) for _ in range(NB)]
    peak_lr, n_warm = 4e-4, 500

    def lr_at(step):  # WarmupCosine with a long horizon: linear warm-up region (scheduling.py:58-68)
        return peak_lr * min(step, n_warm) / n_warm

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    lib = _lib.load()
    if args.n_prefixes > 1:
        # saev draws new cuts every step on the host (objectives.py:125); one fixed draw keeps the benchmark repeatable
        import torch as _t

        from saev_b200.nn import sample_prefixes

        _t.manual_seed(0)
        eng.set_prefixes(sample_prefixes(S, args.n_prefixes).tolist())
    gstep = 0
    for _ in range(args.warmup):
        tr.step(x_dev[gstep % NB], lr_at(gstep))
        gstep += 1
    # pre-heat: a 20-step timed region is ~0.1 s, far too short for clocks and power to settle; run the same steps
    # untimed for a fixed wall time first (every rank the same number of steps: the count comes from rank 0)
    if args.preheat_s > 0:
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        tr.step(x_dev[gstep % NB], lr_at(gstep))
        gstep += 1
        torch.cuda.synchronize(dev)
        per = max(1e-4, time.perf_counter() - t0)
        n_pre = torch.tensor([max(1, min(5000, int(args.preheat_s / per)))], device=dev)
        if world > 1:
            dist.broadcast(n_pre, src=0)
        for _ in range(int(n_pre.item())):
            tr.step(x_dev[gstep % NB], lr_at(gstep))
            gstep += 1

    # ---------------- timed region 1: inputs resident in HBM ----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = lib.saev_b200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        tr.step(x_dev[gstep % NB], lr_at(gstep))
        gstep += 1
    e1.record()
    barrier()
    launches = lib.saev_b200_launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    # per-stage device times (the roofline's kernel time): the same steps once more with the library's CUDA-event
    # stage timers on -- kept out of the headline region, whose launches they would space out
    eng.profile_enable(True)
    n_prof = min(args.steps, 10)
    for _ in range(n_prof):
        tr.step(x_dev[gstep % NB], lr_at(gstep))
        gstep += 1
    barrier()
    stages = eng.profile_read()
    eng.profile_enable(False)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    tr.finish()
    final_losses = tr.global_losses()
    screen = eng.screen_stats()

    # ---------------- timed region 2: end to end ----------------
    # Preferred: the loader a saev user would construct -- activation shards on local storage (tmpfs) are read by the
    # native I/O threads into PINNED staging chunks, copied to the HBM shuffle pool on the loader's stream, gathered
    # into batches on the device, and every step's loss vector is read back to pinned host memory.  Fallback (no
    # writable scratch space): a two-slot pinned ring fed from host tensors.
    e2e_path, e2e_note = "loader", ""
    e2e_value = None
    loss_host = torch.empty(args.steps, 8, dtype=torch.float32).pin_memory()
    e2e_disk = None
    if args.loader_threads <= 0:  # do not oversubscribe the host: every rank also runs a feeder + the Python loop
        args.loader_threads = max(2, min(8, (os.cpu_count() or 8) // world - 1))
    try:
        if args.e2e != "loader":
            raise RuntimeError("ring requested")
        e2e_value, e2e_note, gstep, _ = loader_leg(tr, eng, dev, world, rank, D, B, args.steps, args.loader_threads, lr_at,
                                                   gstep, barrier, "tmpfs")
    except Exception as err:  # noqa: BLE001
        e2e_path, e2e_note = "pinned-ring", f"loader path unavailable ({type(err).__name__}: {str(err)[:80]})"
    if e2e_value is not None and world == 1 and not args.no_disk_leg:
        # the same leg from real storage (HBM-resident / page cache / disk reported separately)
        try:
            v, note, gstep, wall = loader_leg(tr, eng, dev, world, rank, D, B, args.steps, args.loader_threads, lr_at, gstep,
                                              barrier, "disk", zero_copy=False)
            e2e_disk = {"value": v, "unit": "activations/s", "read_GBps": v * D * 4 / 1e9, "note": note}
        except Exception as err:  # noqa: BLE001
            e2e_disk = {"error": f"{type(err).__name__}: {str(err)[:120]}"}
    if e2e_value is None:
        NH = min(NB, 4)  # pinned host copies for the fallback ring
        x_host = [torch.empty(B, D, dtype=torch.float32).pin_memory() for _ in range(NH)]
        for h, d in zip(x_host, x_dev):
            h.copy_(d)
        stage_bufs = [torch.empty(B, D, device=dev) for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        copied = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        main_stream = torch.cuda.current_stream(dev)

        def prefetch(i):
            slot = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[slot])
                stage_bufs[slot].copy_(x_host[i % NH], non_blocking=True)
                copied[slot].record(copy_stream)

        for s in range(2):
            consumed[s].record(main_stream)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        prefetch(0)
        for i in range(args.steps):
            if i + 1 < args.steps:
                prefetch(i + 1)
            slot = i % 2
            main_stream.wait_event(copied[slot])
            tr.step(stage_bufs[slot], lr_at(gstep))
            consumed[slot].record(main_stream)
            loss_host[i].copy_(eng.losses, non_blocking=True)
            gstep += 1
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_value = world * B * args.steps / (float(t.item()) * 1e-3)

    if rank == 0:
        peaks, peaks_src = load_peaks()
        peak_tf = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops")
        peak_burst = peaks.get("bf16_tflops") or peak_tf
        gemm_ms, gemm_n = stages["encode_gemm"]
        gemm_avg_ms = gemm_ms / max(1, gemm_n)
        flops = 2.0 * B * D * S  # algorithmic FLOPs of the encoder contraction per launch
        achieved_tf = flops / (gemm_avg_ms * 1e-3) / 1e12 if gemm_avg_ms > 0 else 0.0
        traffic = None
        tp = ROOT / "profiles" / "encode_gemm_traffic.json"
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get(args.workload)
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": "activations/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": workload_name(args.workload, D, S, K, B, args.n_prefixes, args.k_aux, args.dead_threshold),
                "global_batch": world * B,
                **({"dense_features": args.dense_features} if args.dense_features > 0 else {}),
                **({"dead_atoms": args.dead_atoms} if args.dead_atoms > 0 else {}),
                "parallelism": f"dp{world}" + (f" ({args.dp_mode} gradient exchange)" if world > 1 else ""),
                **({"numa_bind": numa} if world > 1 else {}),
                "preheat_s": args.preheat_s,
                "precision": ("fp16 tcgen05 screen of the encoder contraction (deterministic error bound) + exact fp32 "
                              "re-score of the candidates; every value that reaches the loss / gradients / parameters "
                              "is fp32") if K else
                             ("dense path: all five contractions as " +
                              ("3-term (two-piece, ~2^-17 relative)" if os.environ.get("SAEV_B200_DENSE_TERMS", "6")[:1] == "3"
                               else "6-term (three-piece, fp32-class)") +
                              " bf16 split products on tcgen05" +
                              (" (single-CTA kernel)" if os.environ.get("SAEV_B200_DENSE_PAIR", "1")[:1] == "0"
                               else " (CTA-pair kernel, pieces staged once per k-block)") +
                              ", fp32 accumulation; everything else fp32"),
                "l2_policy": f"per-step working set (params+grads+Adam moments {eng.n_params * 16 / 1e9:.2f} GB, "
                             f"{NB} rotating input batches of {B * D * 4 / 1e6:.0f} MB) is far larger than the 126 MB L2; "
                             "no explicit flush",
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "activations/s",
                    "h2d_bytes_per_step": B * D * 4 + (B * 8 if e2e_path == "loader" else 0), "d2h_bytes_per_step": 32,
                    "path": e2e_path, "note": e2e_note, **({"from_disk": e2e_disk} if e2e_disk is not None else {})},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "encode_gemm2_kernel (tcgen05 cta_group::2 encoder contraction + "
                         "top-k screen)" if K else "encode_gemm_kernel<2> (tcgen05 encoder contraction, 3-term split, ReLU)",
                         "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic,
                         "peak_source": f"{peaks_src} (16-bit dense tensor peak, sustained: the kernel is timed inside a "
                                        f"long pre-heated step loop)",
                         "frac_of_burst_peak": achieved_tf / peak_burst if peak_burst else None, "peak_burst": peak_burst,
                         "avg_launch_ms": gemm_avg_ms,
                         "step_frac_of_encoder_roofline": (value / world) * 2.0 * D * S / (peak_tf * 1e12)},
            "stage_ms_per_step": {k: v[0] / n_prof for k, v in stages.items()},
            "final": {"mse": final_losses["mse"], "loss": final_losses["loss"], "n_dead": final_losses["n_dead"],
                      "unsafe_rows": eng.unsafe_rows(), "screen": screen},
        }
        if world == 1:
            try:
                line["final"]["check"] = final_check(eng, x_dev[0], K)
            except Exception as e:  # noqa: BLE001
                line["final"]["check"] = {"error": f"{type(e).__name__}: {e}"[:200]}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_block(D, S, K, B, steps=3, warmup=1, rows=args.cpu_rows or CPU_ROWS)
        if world == 1 and not args.no_torch_gpu_baseline:
            try:
                rate, ms = torch_gpu_reference_run(D, S, K, B, steps=5, warmup=2)
                line["torch_gpu_baseline"] = {"value": rate, "unit": "activations/s", "ms_per_step": ms, "kind": "port",
                                              "sample": f"5 timed steps x {B} rows (+2 warm-up); oracle/sae_oracle.py "
                                                        "on this GPU: dense torch ops, TF32 matmuls (train.py:257)"}
            except Exception as e:  # reported, never fatal for the bench line
                line["torch_gpu_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
