import sys
sys.path.insert(0, ".")
import torch
from saev_b200.engine import Engine, EngineConfig
from saev_b200.parallel import DataParallelTrainer
D, S, K, B = 1024, 65536, 32, 16384
eng = Engine(EngineConfig(d_model=D, d_sae=S, top_k=K, max_batch=B, aux=True))
eng.init_params(seed=0)
tr = DataParallelTrainer(eng)
g = torch.Generator(device="cuda").manual_seed(5)
xs = [torch.randn(B, D, device="cuda", generator=g) for _ in range(4)]
def exact_check(x):
    bad = 0; worst = 0.0
    for r0 in range(0, B, 2048):
        h = x[r0:r0 + 2048].double() @ eng.W_enc_t.double().t() + eng.b_enc.double()
        hv, hi = h.topk(K, dim=1)
        oi = eng.topk_idx[r0:r0 + 2048].long(); ov = eng.topk_val[r0:r0 + 2048].double()
        bad += int((oi.sort(dim=1)[0] != hi.sort(dim=1)[0]).any(dim=1).sum())
        worst = max(worst, float((h.gather(1, oi) - ov).abs().max()))
        gap = (hv[:, -1] - h.topk(K + 8, dim=1).values[:, -1])
    return bad, worst, float(gap.mean())
for it in range(71):
    before = eng.unsafe_rows()
    x = xs[it % 4]
    if it % 10 == 0:
        eng.normalize_w_dec(); eng.forward(x, training=False); torch.cuda.synchronize()
        bad, worst, gap = exact_check(x)
        print(f"step {it}: eval-forward exactness: rows differing {bad}, max value err {worst:.2e}, mean gap(32..40) {gap:.4f}; |W_enc_t| rms {float(eng.W_enc_t.pow(2).mean().sqrt()):.5f} b_enc rms {float(eng.b_enc.pow(2).mean().sqrt()):.2e}", flush=True)
        before = eng.unsafe_rows()
    tr.step(x, 4e-4 * it / 500)
    torch.cuda.synchronize()
    if it % 10 == 0:
        print(f"step {it}: unsafe {eng.unsafe_rows() - before}  mse {eng.loss_dict()['mse']:.6f}", flush=True)
