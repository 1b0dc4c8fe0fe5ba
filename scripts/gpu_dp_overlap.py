"""Timing experiment (torchrun, >= 2 GPUs): how much of the gradient all-reduce hides behind the staged backward."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, torch.distributed as dist
from saev_b200.engine import Engine, EngineConfig
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
hp = os.environ.get("HP", "0") == "1"
opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=hp)
dist.init_process_group("nccl", device_id=dev, pg_options=opts)
D, S, K, B = 1024, 65536, 32, 16384
eng = Engine(EngineConfig(d_model=D, d_sae=S, top_k=K, aux=True, max_batch=B), device=dev)
eng.init_params(seed=0)
x = torch.randn(B, D, device=dev)
eng.forward(x, training=True, tokens_global=B * world)
def timeit(fn, n=10):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def chunks(n):
    step = (-(-S // n) + 7) // 8 * 8
    return [(r, min(S, r + step)) for r in range(0, S, step)]
def bwd_only(n):
    eng.backward_stage(x, 0, tokens_global=B * world)
    for r0, r1 in chunks(n): eng.backward_stage(x, 1, r0, r1, tokens_global=B * world)
def comm_only(n):
    ws = []
    for r0, r1 in chunks(n):
        ws.append(dist.all_reduce(eng.gW_enc_t[r0:r1], async_op=True)); ws.append(dist.all_reduce(eng.gW_dec[r0:r1], async_op=True))
    for w in ws: w.wait()
def overlapped(n):
    eng.backward_stage(x, 0, tokens_global=B * world)
    ws = []
    for r0, r1 in chunks(n):
        eng.backward_stage(x, 1, r0, r1, tokens_global=B * world)
        ws.append(dist.all_reduce(eng.gW_enc_t[r0:r1], async_op=True)); ws.append(dist.all_reduce(eng.gW_dec[r0:r1], async_op=True))
    for w in ws: w.wait()
def mono():
    eng.backward(x, tokens_global=B * world); dist.all_reduce(eng.grads)
res = {"mono": timeit(mono)}
for n in (1, 2, 4, 8):
    res[f"bwd{n}"] = timeit(lambda: bwd_only(n)); res[f"comm{n}"] = timeit(lambda: comm_only(n)); res[f"ovl{n}"] = timeit(lambda: overlapped(n))
if rank == 0: print(f"world={world} high_priority={hp}: " + ", ".join(f"{k}={v:.3f}" for k, v in res.items()))
dist.destroy_process_group()
