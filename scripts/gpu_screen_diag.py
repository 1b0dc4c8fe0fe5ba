"""Diagnostic: how wide is the screen's admission band as training proceeds (c3 shape)?  Prints encoder row-norm
quantiles, the k-th largest pre-activation, the per-column error bound and the number of columns whose upper bound
reaches the k-th largest lower bound, after 0 / 50 / 200 / 450 training steps."""
import sys
sys.path.insert(0, ".")
import torch
from saev_b200.engine import Engine, EngineConfig

D, S, K, B = 1024, 65536, 32, 16384
eng = Engine(EngineConfig(d_model=D, d_sae=S, top_k=K, max_batch=B, aux=True))
eng.init_params(seed=0)
g = torch.Generator(device="cuda").manual_seed(1234)
xs = [torch.randn(B, D, device="cuda", generator=g) for _ in range(4)]
step = 0


def report():
    x = xs[0][:1024]
    W, b = eng.W_enc_t, eng.b_enc
    wn = W.norm(dim=1)
    q = torch.tensor([0.0, 0.01, 0.5, 0.99, 1.0], device="cuda")
    h = x @ W.t() + b
    xn = x.norm(dim=1, keepdim=True)
    a1 = 1.01 * (2 * 2**-11 + 2**-22 + (2 * (D // 16 + 16) + (D // 32 + 8)) * 2**-23 + 2**-21)
    E = xn * wn[None, :] * a1
    lo, up = h - E, h + E
    L = lo.topk(K, dim=1).values[:, -1:]
    n_up = (up >= L).sum(1).float()
    kth = h.topk(K, dim=1).values[:, -1]
    sel = h.topk(K, dim=1).indices
    wn_sel = wn[sel]
    st = eng.screen_stats()
    print(f"step {step}: wnorm q(0,.01,.5,.99,1)={[round(float(v),3) for v in torch.quantile(wn, q)]} "
          f"|b| max {float(b.abs().max()):.3f}  kth mean {float(kth.mean()):.3f}  "
          f"norm of selected atoms mean {float(wn_sel.mean()):.3f}  E(sel) mean {float((xn*wn_sel*a1).mean()):.4f}  "
          f"cols with u>=L: mean {float(n_up.mean()):.1f} max {int(n_up.max())}  "
          f"h std over cols {float(h.std(dim=1).mean()):.3f}  cumulative stats {st}", flush=True)


report()
for n in (50, 150, 250):
    for _ in range(n):
        lr = 4e-4 * min(step, 500) / 500
        eng.train_step(xs[step % 4], lr, fused_renorm=True, pre_normalized=step > 0)
        step += 1
    report()
