"""CPU study behind DESIGN.md's outlier note: how the admission margin of the bf16 top-k screen (and with it the number
of candidates the exact re-score has to touch) reacts to massive-activation dimensions, and what carrying those
dimensions as bf16 (hi, lo) pairs in an extra k-block would buy.  Pure torch on the CPU; not part of the product."""
import torch

torch.manual_seed(0)
B, D, S, K = 256, 1024, 16384, 32
MARGIN_C = 12.0 * 0.81649658 * 2.0 ** -9  # prep_x_kernel: 2 * 6 sigma, sigma = 2^-9 sqrt(2/3) ||x||_inf ||w||_2
W = torch.randn(S, D)
W = W / W.norm(dim=1, keepdim=True)


def bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def study(x, outlier_dims=()):
    h = x.double() @ W.double().T  # exact
    keep = torch.ones(D, dtype=torch.bool)
    keep[list(outlier_dims)] = False
    xs, Ws = bf(x), bf(W)
    screen = (xs[:, keep].double() @ Ws[:, keep].double().T)
    if outlier_dims:  # outlier dims as (hi, lo) x (hi, lo) pieces: hi.hi + hi.lo + lo.hi  (~2^-17 relative)
        od = list(outlier_dims)
        xh, wh = xs[:, od], Ws[:, od]
        xl, wl = bf(x[:, od] - xh), bf(W[:, od] - wh)
        screen = screen + (xh.double() @ wh.double().T + xh.double() @ wl.double().T + xl.double() @ wh.double().T)
    err = (screen - h).abs().max(dim=1).values
    xinf = x[:, keep].abs().max(dim=1).values
    margin = MARGIN_C * xinf  # ||w||_2 = 1
    kth = screen.topk(K, dim=1).values[:, -1]
    cand = (screen > (kth - margin.double())[:, None]).sum(dim=1).float()
    covered = bool((err <= margin / 2).all())
    return float(cand.mean()), float(cand.max()), covered, float((err / (margin / 2)).max())


x = torch.randn(B, D)
print("gaussian                         : cand/row mean %.1f max %.0f, errors within E_b: %s (max ratio %.2f)" % study(x))
for scale in (10.0, 30.0, 100.0):
    xo = x.clone()
    xo[:, [7, 300]] *= scale  # two massive-activation dimensions
    print("2 dims x%-5g current margin      : cand/row mean %.1f max %.0f, errors within E_b: %s (max ratio %.2f)" % ((scale,) + study(xo)))
    print("2 dims x%-5g dims as (hi,lo) pair : cand/row mean %.1f max %.0f, errors within E_b: %s (max ratio %.2f)" % ((scale,) + study(xo, (7, 300))))
