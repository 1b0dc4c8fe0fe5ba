"""Time the log-block kernels at the c3 / c5 dictionary shapes and compare the coherence with a blocked torch fp32
evaluation on the same GPU (what the reference's one-line Gram-matrix formula costs is reported too when it fits)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from saev_b200.engine import Engine, EngineConfig  # noqa: E402


def torch_blocked(W, block=8192):
    Wn = W / W.norm(dim=1, keepdim=True)
    best = torch.zeros((), device=W.device)
    S = W.shape[0]
    cols = torch.arange(S, device=W.device)[None, :]
    for a in range(0, S, block):
        G = (Wn[a:a + block] @ Wn.T).abs()
        rows = torch.arange(a, min(a + block, S), device=W.device)[:, None]
        best = torch.maximum(best, G.masked_fill_(cols <= rows, 0).max())
    return best


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, out


for name, D, S, B in (("c3", 1024, 65536, 16384), ("c5", 1536, 131072, 8192)):
    eng = Engine(EngineConfig(d_model=D, d_sae=S, top_k=32, activation="topk", aux=True, k_aux=512, max_batch=B))
    eng.init_params(seed=0)
    eng.W_dec.mul_(1 + 0.1 * torch.rand(S, 1, device="cuda"))
    x = torch.randn(B, D, device="cuda")
    eng.forward(x, training=True)
    ms_c, out = timed(lambda: eng.dictionary_coherence())
    ms_l, m = timed(lambda: eng.log_metrics(x))
    torch.backends.cuda.matmul.allow_tf32 = False
    ms_t, ref = timed(lambda: torch_blocked(eng.W_dec), n=1)
    torch.backends.cuda.matmul.allow_tf32 = True
    ms_t32, ref32 = timed(lambda: torch_blocked(eng.W_dec), n=1)
    print(f"{name}: coherence kernel {ms_c:.2f} ms -> {float(out[0]):.7f} (screen {float(out[1]):.7f}, pair {int(out[2])},{int(out[3])}); "
          f"all log metrics {ms_l:.2f} ms; torch blocked fp32 {ms_t:.1f} ms -> {float(ref):.7f}; "
          f"torch blocked TF32 {ms_t32:.1f} ms -> {float(ref32):.7f}", flush=True)
    print("   metrics:", dict(zip(eng.LOG_KEYS, m.tolist())), flush=True)
    del eng
    torch.cuda.empty_cache()
