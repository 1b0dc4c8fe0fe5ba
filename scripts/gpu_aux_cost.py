"""How much the AuxK path costs once latents are dead (c3 shape): step time vs number of dead latents."""
import sys
sys.path.insert(0, ".")
import torch
from saev_b200.engine import Engine, EngineConfig
D, S, K, B = 1024, 65536, 32, 16384
for target_dead in (0, 2048, 8192, 32768):
    eng = Engine(EngineConfig(d_model=D, d_sae=S, top_k=K, aux=True, k_aux=512, dead_threshold_tokens=B, max_batch=B))
    eng.init_params(seed=0)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, D, device="cuda", generator=g)
    if target_dead:
        # push the bias of `target_dead` latents far down: they never make the top-k, so they are dead after one step
        eng.b_enc[:target_dead] = -100.0
    eng.train_step(x, 0.0)
    eng.train_step(x, 1e-4)
    torch.cuda.synchronize()
    eng.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 5
    for _ in range(n):
        eng.train_step(x, 1e-4)
    e1.record(); torch.cuda.synchronize()
    st = eng.profile_read()
    ld = eng.loss_dict()
    print(f"n_dead={int(ld['n_dead'])} aux={ld['aux']:.4g}: {e0.elapsed_time(e1)/n:.2f} ms/step; loss stage {st['loss'][0]/n:.2f} ms, bias_aux stage {st['bias_aux'][0]/n:.2f} ms")
    del eng
