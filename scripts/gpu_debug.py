"""Verbose GPU bring-up checks (not part of the test-suite): run under gpurun."""
import sys, time, traceback
sys.path.insert(0, ".")
import torch
from saev_b200.engine import Engine, EngineConfig
from oracle import sae_oracle as O

torch.manual_seed(0)
dev = "cuda"

def rel(a, b):
    a = a.double().cpu(); b = b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))

def check_gemm():
    eng = Engine(EngineConfig(d_model=128, d_sae=512, top_k=16, max_batch=256, aux=False))
    for (M, N, K) in [(128, 256, 64), (128, 256, 128), (200, 600, 128), (256, 512, 1024), (300, 1000, 776)]:
        A = torch.randn(M, K, device=dev); Bt = torch.randn(N, K, device=dev); bias = torch.randn(N, device=dev)
        ref = (A.double() @ Bt.double().T + bias.double()).float()
        for nt in (1, 3):
            out = eng.gemm_nt(A, Bt, bias, nt)
            torch.cuda.synchronize()
            if nt == 1:
                ref1 = (A.bfloat16().double() @ Bt.bfloat16().double().T + bias.double()).float()
                print(f"gemm M{M} N{N} K{K} nterms=1: rel vs bf16-ref {rel(out, ref1):.2e}  vs fp32 {rel(out, ref):.2e}", flush=True)
            else:
                print(f"gemm M{M} N{N} K{K} nterms=3: rel vs fp32 {rel(out, ref):.2e}", flush=True)

def check_step(D, S, K, B, steps=4, aux=False, thr=10_000_000):
    cfg = EngineConfig(d_model=D, d_sae=S, top_k=K, max_batch=B, aux=aux, dead_threshold_tokens=thr, k_aux=16)
    eng = Engine(cfg)
    g = torch.Generator().manual_seed(1)
    W_enc, b_enc, W_dec, b_dec = O.init_params(D, S, g)
    b_enc = 0.01 * torch.randn(S, generator=g); b_dec = 0.01 * torch.randn(D, generator=g)
    ocfg = O.OracleConfig(d_model=D, d_sae=S, top_k=K, aux=aux, k_aux=16, dead_threshold_tokens=thr, lr=1e-3, n_lr_warmup=2, n_steps=100)
    st = O.OracleState.from_params(W_enc, b_enc, W_dec, b_dec)
    eng.load_params(W_enc, b_enc, W_dec, b_dec)
    lr = 0.0
    for s in range(steps):
        x = torch.randn(B, D, generator=g)
        xd = x.to(dev)
        eng.normalize_w_dec()
        eng.forward(xd, training=True)
        torch.cuda.synchronize()
        res = O.train_step(ocfg, st, x)
        ld = eng.loss_dict()
        print(f"[D{D} S{S} K{K} B{B}] step {s}: mse {ld['mse']:.6f} vs {res['mse']:.6f}  l0 {ld['l0']} l1 {ld['l1']:.5f} vs {res['l1']:.5f} n_dead {ld['n_dead']} vs {res['n_dead']} aux {ld['aux']:.6f} vs {res['aux']:.6f} unsafe {eng.unsafe_rows()}", flush=True)
        f_ref = res['out'].f
        f_ours = eng.dense_f_x(B).cpu()
        print("   f rel", rel(f_ours, f_ref), " resid rel", rel(eng.resid[:B], res['out'].r), flush=True)
        eng.backward(xd)
        eng.grad_sumsq()
        torch.cuda.synchronize()
        # oracle grads are post-clip; ours pre-clip: compare after applying clip coef
        gn = float(eng.sumsq.sqrt())
        coef = min(1.0, ocfg.grad_clip / (gn + 1e-6))
        print(f"   gnorm {gn:.6f} vs {res['grad_norm']:.6f}", flush=True)
        print("   gW_enc rel", rel(eng.gW_enc_t.t() * coef, res['grads']['W_enc']),
              "gb_enc", rel(eng.gb_enc * coef, res['grads']['b_enc']),
              "gW_dec", rel(eng.gW_dec * coef, res['grads']['W_dec']),
              "gb_dec", rel(eng.gb_dec * coef, res['grads']['b_dec']), flush=True)
        eng.adam_step(lr, max_norm=ocfg.grad_clip)
        torch.cuda.synchronize()
        lr = st.lr
        print("   params: W_enc", rel(eng.W_enc_t.t(), st.W_enc), "b_enc", rel(eng.b_enc, st.b_enc),
              "W_dec", rel(eng.W_dec, st.W_dec), "b_dec", rel(eng.b_dec, st.b_dec), flush=True)

def timing(D, S, K, B, iters=5):
    cfg = EngineConfig(d_model=D, d_sae=S, top_k=K, max_batch=B, aux=False)
    eng = Engine(cfg)
    g = torch.Generator().manual_seed(1)
    W_enc, b_enc, W_dec, b_dec = O.init_params(D, S, g)
    eng.load_params(W_enc, b_enc, W_dec, b_dec)
    x = torch.randn(B, D, device=dev)
    names = ["normalize", "forward", "backward", "sumsq", "adam"]
    fns = [eng.normalize_w_dec, lambda: eng.forward(x, training=True), lambda: eng.backward(x), eng.grad_sumsq,
           lambda: eng.adam_step(1e-4)]
    for _ in range(2):
        for f in fns: f()
    torch.cuda.synchronize()
    tot = {n: 0.0 for n in names}
    for _ in range(iters):
        for n, f in zip(names, fns):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize()
            tot[n] += e0.elapsed_time(e1) / iters
    step = sum(tot.values())
    print(f"timing D{D} S{S} K{K} B{B}: " + " ".join(f"{n}={t:.3f}ms" for n, t in tot.items()) + f" | step {step:.3f} ms => {B/step*1e3:.0f} act/s; unsafe rows {eng.unsafe_rows()}", flush=True)

for name, fn in [("gemm", check_gemm),
                 ("c1", lambda: check_step(128, 512, 16, 256)),
                 ("mid", lambda: check_step(768, 4096, 32, 1000, steps=3)),
                 ("t2", lambda: timing(768, 32768, 32, 4096)),
                 ("t3", lambda: timing(1024, 65536, 32, 16384))]:
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    try:
        t0 = time.time(); fn(); print(f"== {name} done in {time.time()-t0:.1f}s", flush=True)
    except Exception:
        traceback.print_exc(); print(f"== {name} FAILED", flush=True)
