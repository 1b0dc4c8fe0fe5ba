"""Full-size check of the screen + re-score selection against an exact fp32 top-k (torch, TF32 off)."""
import sys
sys.path.insert(0, ".")
import torch
from saev_b200.engine import Engine, EngineConfig
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
torch.set_float32_matmul_precision("highest")
D, S, K, B = 1024, 65536, 32, 2048
eng = Engine(EngineConfig(d_model=D, d_sae=S, top_k=K, max_batch=B, aux=False))
eng.init_params(seed=0)
eng.b_enc.copy_(0.02 * torch.randn(S, device="cuda"))
x = torch.randn(B, D, device="cuda")
eng.forward(x, training=False)
torch.cuda.synchronize()
print("unsafe rows flagged:", eng.unsafe_rows(), "of", B)
h = (x.double() @ eng.W_enc_t.double().t() + eng.b_enc.double())
hv, hi = h.topk(K, dim=1)
ours_i = eng.topk_idx[:B].long(); ours_v = eng.topk_val[:B].double()
oi, _ = ours_i.sort(dim=1); ri, _ = hi.sort(dim=1)
bad_rows = (oi != ri).any(dim=1)
print("rows whose index set differs from exact fp64 top-k:", int(bad_rows.sum()))
# value parity on sorted values
ov, _ = ours_v.sort(dim=1, descending=True)
print("max |val diff| (sorted values):", float((ov - hv).abs().max()), " rel:", float(((ov - hv).abs() / hv.abs()).max()))
# margins: gap between exact k-th and (k+1)-th .. and 40th
h41 = h.topk(48, dim=1).values
gap_32_40 = (h41[:, 31] - h41[:, 39])
approx = (x.bfloat16().double() @ eng.W_enc_t.bfloat16().double().t() + eng.b_enc.double())
err = (approx - h).abs()
print("bf16 screen error: mean %.5f max %.5f ; h std %.4f" % (float(err.mean()), float(err.max()), float(h.std())))
print("gap(32nd - 40th exact): mean %.4f  min %.5f; frac rows gap < 4*max_err_row: %.3f" % (
    float(gap_32_40.mean()), float(gap_32_40.min()), float((gap_32_40 < 4 * err.max(dim=1).values).double().mean())))
# does the approx top-40 contain the exact top-32 ?
a40 = approx.topk(40, dim=1).indices
contained = torch.stack([torch.isin(hi[r], a40[r]).all() for r in range(B)])
print("rows whose exact top-32 is inside the bf16 top-40:", int(contained.sum()), "of", B)
