"""Full-size check of the screen + re-score selection against an exact top-k (torch fp64, row chunks)."""
import sys
sys.path.insert(0, ".")
import torch
from saev_b200.engine import Engine, EngineConfig
D, S, K = 1024, 65536, 32
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
train = len(sys.argv) > 2
eng = Engine(EngineConfig(d_model=D, d_sae=S, top_k=K, max_batch=B, aux=False))
eng.init_params(seed=0)
if not train:
    eng.b_enc.copy_(0.02 * torch.randn(S, device="cuda"))
g = torch.Generator(device="cuda").manual_seed(5)
for it in range(3 if train else 1):
    x = torch.randn(B, D, device="cuda", generator=g)
    before = eng.unsafe_rows()
    eng.forward(x, training=train)
    torch.cuda.synchronize()
    print(f"iter {it}: B={B} training={train}: unsafe rows flagged: {eng.unsafe_rows() - before} of {B}")
    bad = 0; worst = 0.0; missing_rank = []
    for r0 in range(0, B, 2048):
        xs = x[r0:r0 + 2048]
        h = xs.double() @ eng.W_enc_t.double().t() + eng.b_enc.double()
        hv, hi = h.topk(K, dim=1)
        oi = eng.topk_idx[r0:r0 + 2048].long(); ov = eng.topk_val[r0:r0 + 2048].double()
        so, _ = oi.sort(dim=1); sr, _ = hi.sort(dim=1)
        rows_bad = (so != sr).any(dim=1)
        bad += int(rows_bad.sum())
        ovs, _ = ov.sort(dim=1, descending=True)
        worst = max(worst, float((ovs - hv).abs().max()))
        # values at our indices must equal the exact pre-activations there
        worst = max(worst, float((h.gather(1, oi) - ov).abs().max()))
    print(f"   rows whose index set differs from exact fp64 top-k: {bad}; max |value error| {worst:.3e}")
    if train:
        eng.backward(x); eng.grad_sumsq(); eng.adam_step(1e-4)
