"""Multi-GPU parity check, run under torchrun: N ranks x (B/N) rows through DataParallelTrainer (sharded optimizer and
plain all-reduce) must give what ONE engine gives on the full B rows (the same kernels, global denominators).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/gpu_dp_check.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist

from saev_b200.engine import Engine, EngineConfig
from saev_b200.parallel import DataParallelTrainer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
D, S, K, B = 256, 4096, 16, 512 * world
cfg = EngineConfig(d_model=D, d_sae=S, top_k=K, aux=True, k_aux=64, dead_threshold_tokens=2 * B, max_batch=B)
g = torch.Generator(device="cpu").manual_seed(0)
basis = torch.randn(8, D, generator=g)
xs = [(torch.randn(B, 8, generator=g) @ basis + 0.05 * torch.randn(B, D, generator=g)).to(dev) for _ in range(4)]


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-30))


ref = Engine(cfg, device=dev)
ref.init_params(seed=7)
ref0 = {k: getattr(ref, k).clone() for k in ("W_enc_t", "b_enc", "W_dec", "b_dec")}
lrs = [0.0, 5e-4, 1e-3, 1e-3]
ref_losses = []
for x, lr in zip(xs, lrs):
    ref.train_step(x, lr, fused_renorm=True, pre_normalized=False)
    ref_losses.append(ref.loss_dict())
ok = True
per = B // world
gopts = dist.ProcessGroupNCCL.Options()
gopts.config.max_ctas = 4
bg = dist.new_group(backend="nccl", pg_options=gopts)
for sharded, n_chunks, gg in ((False, 4, None), (False, 1, None), (True, 1, None), (True, 1, bg), (True, 4, None),
                              (True, 4, bg)):
    eng = Engine(cfg, device=dev)
    eng.init_params(seed=7)
    tr = DataParallelTrainer(eng, sharded=sharded, n_chunks=n_chunks, gather_group=gg, reserved_sms=4 if gg else 0)
    tr.broadcast_params(0)
    assert tr.sharded == sharded and bool(tr.chunks) == (n_chunks > 1 and not sharded)
    assert bool(tr.shard_chunks) == (n_chunks > 1 and sharded)
    sharded = f"{sharded}/chunks={n_chunks}/background-gather={gg is not None}"
    for i, (x, lr) in enumerate(zip(xs, lrs)):
        tr.step(x[rank * per:(rank + 1) * per].contiguous(), lr, fused_renorm=True)
        gl = tr.global_losses()
        for k in ("mse", "aux", "l0", "loss"):
            if abs(gl[k] - ref_losses[i][k]) > 2e-5 * max(abs(ref_losses[i][k]), 1e-6) + 1e-7:
                ok = False
                print(f"[rank {rank}] sharded={sharded} step {i} {k}: {gl[k]} vs {ref_losses[i][k]}")
        if int(gl["n_dead"]) != int(ref_losses[i]["n_dead"]):
            ok = False
            print(f"[rank {rank}] sharded={sharded} step {i} n_dead {gl['n_dead']} vs {ref_losses[i]['n_dead']}")
    tr.finish()
    errs = {n: rel(a, b) for n, a, b in (("W_enc_t", eng.W_enc_t, ref.W_enc_t), ("b_enc", eng.b_enc, ref.b_enc),
                                         ("W_dec", eng.W_dec, ref.W_dec), ("b_dec", eng.b_dec, ref.b_dec))}
    sh = rel(eng.shadow_weights().float(), ref.shadow_weights().float())
    st = eng.screen_stats()
    if max(errs.values()) > 2e-5 or sh > 1e-6 or st["unrepaired"] != 0 or st["bound_violations"] != 0:
        ok = False
    if rank == 0:
        print(f"sharded={sharded}: param rel-L2 vs single-GPU {errs}, fp16 operand {sh:.2e}, n_dead(last)={ref_losses[-1]['n_dead']}")
# ---- the drop-in surface under data parallelism: saev_b200.nn objective + optim shims, `sae.data_parallel()` ----
from saev_b200 import nn as bnn, optim as boptim

torch.manual_seed(7)
sae_cfg = bnn.SparseAutoencoderConfig(d_model=D, d_sae=S, activation=bnn.TopK(top_k=K, aux=bnn.AuxK(k_aux=64)),
                                      reinit_blend=0.0)
sae = bnn.SparseAutoencoder(sae_cfg)
with torch.no_grad():  # same start as the engines above
    sae.W_dec.copy_(ref0["W_dec"]); sae.b_dec.copy_(ref0["b_dec"]); sae.W_enc.copy_(ref0["W_enc_t"].t()); sae.b_enc.copy_(ref0["b_enc"])
obj = bnn.get_objective(bnn.Matryoshka(n_prefixes=1, dead_threshold_tokens=2 * B))
opt = boptim.FusedAdam([{"params": sae.parameters(), "lr": 0.0}], fused=True)
sae.train(); sae = sae.to(dev); obj.train(); sae.data_parallel()
for i, (x, lr) in enumerate(zip(xs, lrs)):
    opt.param_groups[0]["lr"] = lr
    sae.normalize_w_dec()
    loss, fwd = obj(sae, x[rank * per:(rank + 1) * per].contiguous())
    loss.loss.backward()
    sae.remove_parallel_grads()
    gn = boptim.clip_grad_norm_(sae.parameters(), max_norm=1.0)
    m = loss.metrics()
    for k in ("mse", "aux", "l0", "loss"):
        if abs(m[k] - ref_losses[i][k]) > 2e-5 * max(abs(ref_losses[i][k]), 1e-6) + 1e-7:
            ok = False
            print(f"[rank {rank}] nn-API step {i} {k}: {m[k]} vs {ref_losses[i][k]}")
    opt.step()
    opt.zero_grad()
sae.normalize_w_dec()  # the engines above renormalised inside the Adam kernel
errs = {n: rel(a, b) for n, a, b in (("W_enc", sae.W_enc.t(), ref.W_enc_t), ("b_enc", sae.b_enc, ref.b_enc),
                                     ("W_dec", sae.W_dec, ref.W_dec), ("b_dec", sae.b_dec, ref.b_dec))}
if max(errs.values()) > 2e-5:
    ok = False
if rank == 0:
    print(f"nn API + data_parallel(): param rel-L2 vs single-GPU {errs}")
flag = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(flag)
if rank == 0:
    print("DP CHECK", "OK" if int(flag) == 0 else "FAILED")
dist.destroy_process_group()
sys.exit(0 if int(flag) == 0 else 1)
