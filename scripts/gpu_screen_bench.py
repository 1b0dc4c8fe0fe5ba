"""Timing experiments on the screen kernel alone (c3 shape, weights after `--train` steps).
SAEV_B200_SCREEN_DEBUG: 0 = the real kernel, 1 = nothing admitted (scan only), 2 = empty epilogue (TMA + MMA pipeline
only), 3 = thresholds kept from the previous launch on the SAME batch (perfect warm start)."""
import os
import sys
sys.path.insert(0, ".")
import torch
from saev_b200 import _lib
from saev_b200.engine import Engine, EngineConfig

D, S, K, B = 1024, 65536, 32, 16384
train = int(sys.argv[1]) if len(sys.argv) > 1 else 150
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "1", "2", "3", "0"]
eng = Engine(EngineConfig(d_model=D, d_sae=S, top_k=K, max_batch=B, aux=True))
eng.init_params(seed=0)
g = torch.Generator(device="cuda").manual_seed(1234)
xs = [torch.randn(B, D, device="cuda", generator=g) for _ in range(4)]
for step in range(train):
    eng.train_step(xs[step % 4], 4e-4 * min(step, 500) / 500, fused_renorm=True, pre_normalized=step > 0)
torch.cuda.synchronize()
for mode in modes:
    dbg, _, trig = mode.partition(":")
    os.environ["SAEV_B200_SCREEN_DEBUG"] = dbg
    if trig:
        os.environ["SAEV_B200_TRIGGER"] = trig
    same = dbg == "3"
    os.environ["SAEV_B200_SCREEN_DEBUG"] = "0"
    eng.forward(xs[0], training=True, phase=_lib.PHASE_A_SCREEN)  # (mode 3: leaves the thresholds of xs[0] behind)
    os.environ["SAEV_B200_SCREEN_DEBUG"] = dbg
    for _ in range(3):
        eng.forward(xs[0], training=True, phase=_lib.PHASE_A_SCREEN)
    eng.profile_enable(True)
    for i in range(20):
        eng.forward(xs[0 if same else i % 4], training=True, phase=_lib.PHASE_A_SCREEN)
    st = eng.profile_read()
    eng.profile_enable(False)
    ms = st["encode_gemm"][0] / st["encode_gemm"][1]
    cnt = eng._ws_tensor(eng.lib.saev_b200_unsafe_rows, torch.int32, 1)  # sync
    print(f"debug={mode}: screen kernel {ms:.3f} ms  = {2.0 * B * D * S / ms / 1e9:.0f} TF/s", flush=True)
