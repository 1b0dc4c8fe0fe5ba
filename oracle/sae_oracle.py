"""CPU oracle for the saev SAE training step.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain dense fp32 PyTorch-CPU tensor ops with *explicit* gradient
formulas (no autograd), the arithmetic of one step of the reference training loop

    /root/reference/src/saev/framework/train.py:332-460
    /root/reference/src/saev/nn/objectives.py:101-156, 224-237
    /root/reference/src/saev/nn/modeling.py:25-103, 150-179, 343-445
    /root/reference/src/saev/utils/scheduling.py:43-71

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it; the product package `saev_b200` never does (it has no CPU path at all).

Parity status: PINNED.  `oracle/gen_golden.py` runs the *live* reference (imported from
/root/reference/src in the build container) on fixed seeds and writes `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks every function here against those vectors and against the
hand-computed known answers the reference's own tests hold (tests/test_auxk.py,
tests/test_nn_activations.py, tests/test_nn_objectives.py, tests/test_nn_modeling.py).

All tensors are fp32, row-major.  Shapes: x[B,D]  W_enc[D,S]  b_enc[S]  W_dec[S,D]  b_dec[D].
"""

from __future__ import annotations

import dataclasses
import math

import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# configuration (mirrors modeling.py:109-146, 259-284; objectives.py:13-25; train.py:50-78)
# --------------------------------------------------------------------------------------
@dataclasses.dataclass(frozen=True)
class OracleConfig:
    d_model: int
    d_sae: int
    activation: str = "topk"  # "topk" | "relu" | "batchtopk"   modeling.py:111-146
    batch_momentum: float = 0.1  # BatchTopK.momentum          modeling.py:140
    top_k: int = 32  # modeling.py:123
    l1_coeff: float = 0.0  # L1Sparsity.coeff (0 => NoSparsity)  modeling.py:25-42
    aux: bool = True  # AuxK vs NoAux                  modeling.py:50-103
    k_aux: int = 512  # modeling.py:72
    aux_alpha: float = 1.0 / 32  # modeling.py:73
    dead_threshold_tokens: int = 10_000_000  # objectives.py:24
    normalize_w_dec: bool = True  # modeling.py:283
    remove_parallel_grads: bool = True  # modeling.py:281
    lr: float = 4e-4  # train.py:73
    n_lr_warmup: int = 500  # train.py:75
    n_steps: int = 1000  # len(BatchLimiter)  scheduling.py:94-95
    grad_clip: float = 1.0  # train.py:77
    beta1: float = 0.9  # torch.optim.Adam defaults, train.py:294
    beta2: float = 0.999
    eps: float = 1e-8


@dataclasses.dataclass
class OracleState:
    W_enc: Tensor
    b_enc: Tensor
    W_dec: Tensor
    b_dec: Tensor
    m: dict
    v: dict
    t: int = 0  # Adam step count
    lr: float = 0.0  # param_group lr; 0.0 before the first step (train.py:118)
    sched_step: int = 0  # WarmupCosine._step
    toks_since_active: Tensor | None = None  # objectives.py:99,108-111 (lazily created)
    threshold: float = 0.0  # BatchTopKActivation.threshold buffer (modeling.py:213): EMA of the smallest positive survivor

    @staticmethod
    def from_params(W_enc, b_enc, W_dec, b_dec) -> "OracleState":
        ps = dict(W_enc=W_enc.clone(), b_enc=b_enc.clone(), W_dec=W_dec.clone(), b_dec=b_dec.clone())
        return OracleState(
            **ps,
            m={k: torch.zeros_like(p) for k, p in ps.items()},
            v={k: torch.zeros_like(p) for k, p in ps.items()},
        )

    def params(self) -> dict:
        return dict(W_enc=self.W_enc, b_enc=self.b_enc, W_dec=self.W_dec, b_dec=self.b_dec)


def init_params(d_model: int, d_sae: int, generator: torch.Generator | None = None):
    """modeling.py:306-329: W_dec = kaiming_uniform_([S,D]) row-normalised; W_enc = W_dec.T.clone();
    biases zero.  kaiming_uniform_ with a=0 on a [S,D] tensor: fan_in = D, bound = sqrt(6/D)."""
    bound = math.sqrt(6.0 / d_model)
    W_dec = (torch.rand(d_sae, d_model, generator=generator) * 2 - 1) * bound
    W_dec = W_dec / torch.norm(W_dec, dim=1, keepdim=True)
    W_enc = W_dec.T.clone()
    return W_enc, torch.zeros(d_sae), W_dec, torch.zeros(d_model)


# --------------------------------------------------------------------------------------
# forward pieces
# --------------------------------------------------------------------------------------
def normalize_w_dec(W_dec: Tensor) -> Tensor:
    """modeling.py:411-417: W_dec /= ||W_dec||_2 along dim=1 (rows = dictionary atoms)."""
    return W_dec / torch.norm(W_dec, dim=1, keepdim=True)


def encode_pre(x: Tensor, W_enc: Tensor, b_enc: Tensor) -> Tensor:
    """modeling.py:344-347: h = x @ W_enc + b_enc."""
    return x @ W_enc + b_enc


def topk_activation(h: Tensor, top_k: int):
    """modeling.py:174-179: idx = topk(h, min(k,S), sorted=False); f = scatter(zeros, idx, 1) * h.
    No ReLU: negative pre-activations can be selected.  Returns (f, mask)."""
    k = min(top_k, h.shape[1])
    _, idx = torch.topk(h, k, dim=-1, sorted=False)
    mask = torch.zeros_like(h).scatter(-1, idx, 1.0)
    return mask * h, mask


def batch_topk_activation(h: Tensor, top_k: int, training: bool, threshold: float, momentum: float, flat_idx=None):
    """modeling.py:214-244.  Training: the (top_k * B) largest entries of the flattened [B, S] matrix survive
    (`torch.topk(x_flat, k, sorted=False)` -> scatter mask), and threshold <- (1 - m) threshold + m * min(positive
    survivors).  Eval: JumpReLU, x if x > threshold else 0 (threshold <= 0: x if x > 0).  Returns (f, mask, threshold).
    `flat_idx`: the kernels' selection as flat indices (adopted after a tie check, cf. adopt_selection)."""
    if not training:
        thr = max(float(threshold), 0.0)
        mask = (h > thr).to(h.dtype)
        return mask * h, mask, float(threshold)
    B, S = h.shape
    k = min(top_k * B, S * B)
    flat = h.flatten()
    _, idx = torch.topk(flat, k, sorted=False)
    if flat_idx is not None:
        idx = adopt_selection(flat[None, :], idx[None, :], flat_idx[None, :], what="batch top-k")[0]
    mask = torch.zeros_like(flat).scatter(-1, idx, 1.0).reshape(h.shape)
    f = mask * h
    pos = f[f > 0]
    if pos.numel() > 0:  # (the reference's `pos.min()` raises on an empty selection)
        threshold = float(torch.tensor(float(threshold), dtype=torch.float32) * (1 - momentum) + momentum * pos.min())
    return f, mask, threshold


def relu_activation(h: Tensor):
    """modeling.py:155-156: relu(h).  Returns (f, mask) with mask = h > 0 (autograd of relu)."""
    return torch.relu(h), (h > 0).to(h.dtype)


def dead_tracker_update(toks: Tensor | None, f: Tensor, dead_threshold_tokens: int):
    """objectives.py:107-120 (training only): active = any_b(|f|>0); toks += B; toks[active] = 0;
    dead = toks >= threshold.  Returns (toks_new int64[S], dead bool[S])."""
    B, S = f.shape
    if toks is None:
        toks = torch.zeros(S, dtype=torch.int64, device=f.device)
    active = (f.abs() > 0).any(dim=0)
    toks = toks + B
    toks = torch.where(active, torch.zeros_like(toks), toks)
    return toks, toks >= dead_threshold_tokens


def decode(f: Tensor, W_dec: Tensor, b_dec: Tensor) -> Tensor:
    """modeling.py:386-406 with the default single prefix [d_sae]: x_hat = f @ W_dec + b_dec."""
    return f @ W_dec + b_dec


def decode_prefixes(f: Tensor, W_dec: Tensor, b_dec: Tensor, prefixes) -> Tensor:
    """modeling.py:364-406 (Matryoshka): blocks [0,p1), [p1,p2), ...; block i contributes f[:, blk] @ W_dec[blk]
    (+ b_dec on block 0); x_hats[:, i] = cumulative sum over blocks <= i.  Returns [B, P, D]."""
    cuts = [0] + [int(c) for c in prefixes]
    outs = []
    for i, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
        o = f[:, a:b] @ W_dec[a:b]
        outs.append(o + b_dec if i == 0 else o)
    return torch.cumsum(torch.stack(outs, dim=-2), dim=-2)


def sample_prefixes(d_sae: int, n_prefixes: int, min_prefix_length: int = 1, pareto_power: float = 0.5) -> Tensor:
    """objectives.py:158-201: n_prefixes - 1 cut points drawn without replacement (torch.multinomial on the GLOBAL CPU
    generator) from a Pareto-shaped distribution over 1..d_sae-1, plus d_sae itself, sorted ascending."""
    if n_prefixes <= 1:
        return torch.tensor([d_sae], dtype=torch.int64)
    assert n_prefixes <= d_sae
    lengths = torch.arange(1, d_sae)
    cdf = 1 - ((min_prefix_length / lengths.float()) ** pareto_power)
    pdf = torch.cat([cdf[:1], cdf[1:] - cdf[:-1]])
    idx = torch.multinomial(pdf / pdf.sum(), num_samples=n_prefixes - 1, replacement=False)
    out = torch.cat((lengths[idx].detach().clone(), torch.tensor([d_sae])))
    return torch.sort(out, descending=False)[0].to(torch.int64)


def mean_squared_err(x_hat: Tensor, x: Tensor) -> Tensor:
    """objectives.py:224-237 followed by .mean() (objectives.py:133-138):
    u = max|x| clamped at 1e-12; ((x_hat/u - x/u)^2) * u^2, mean over all elements."""
    upper = x.abs().max().clamp(min=1e-12)
    d = x_hat / upper - x / upper
    return (d * d * upper * upper).mean()


def adopt_selection(h: Tensor, own_idx: Tensor, given_idx: Tensor, tie_tol: float = 4e-6, what: str = "top-k") -> Tensor:
    """torch.topk breaks fp32-level ties between the k-th and (k+1)-th largest value of a row in an
    implementation-defined way; at 10^8 (row, column) pairs per batch SOME row has such a tie, and the kernels (other
    summation order) may keep the other column -- one swapped entry is 4e-3 of ||gW_enc|| at batch 4096.  This checks
    that `given_idx` (the kernels' selection, int64 [B, k]) differs from torch's `own_idx` in at most a handful of rows
    and ONLY by such ties (exchanged values agree to tie_tol * max|h_row|), then returns it for the oracle to use, so
    that everything downstream of the selection can still be compared tightly."""
    assert given_idx.shape == own_idx.shape, (what, given_idx.shape, own_idx.shape)
    so, sr = given_idx.sort(dim=1).values, own_idx.sort(dim=1).values
    rows = (so != sr).any(dim=1).nonzero().flatten()
    assert rows.numel() <= max(2, h.shape[0] // 250), f"{what}: {rows.numel()} rows differ from torch.topk"
    for r in rows.tolist():
        a, b = set(so[r].tolist()), set(sr[r].tolist())
        assert len(a) == so.shape[1], f"{what}: duplicate column in row {r}"
        only_given, only_own = sorted(a - b), sorted(b - a)
        finite = h[r][torch.isfinite(h[r])]
        scale = float(finite.abs().max())
        gap = (h[r, only_given].sort().values - h[r, only_own].sort().values).abs().max()
        assert float(gap) <= tie_tol * scale, (what, r, only_given, only_own, float(gap), scale)
    return given_idx


def auxk(h: Tensor, r: Tensor, dead: Tensor, W_dec: Tensor, b_dec: Tensor, k_aux: int, alpha: float,
         aux_idx: Tensor | None = None):
    """modeling.py:89-103.  e = (x - x_hat).detach() = -r; masked = h.masked_fill(~dead, -inf);
    k_use = min(k_aux, n_dead); top_i = masked.topk(k_use); f_aux = scatter(h at top_i);
    x_aux = decode(f_aux) (b_dec IS added); aux = alpha * mean((x_aux - e)^2).
    Returns (aux scalar, f_aux[B,S] or None, r_aux[B,D] or None).  `aux_idx`: see adopt_selection()."""
    n_dead = int(dead.sum())
    k_use = min(k_aux, n_dead)
    if k_use == 0:
        return torch.zeros(()), None, None
    e = -r
    masked = h.masked_fill(~dead, float("-inf"))
    _, top_i = masked.topk(k_use, dim=-1)
    if aux_idx is not None:
        top_i = adopt_selection(masked, top_i, aux_idx, what="AuxK top-k_aux")
    f_aux = torch.zeros_like(h)
    f_aux.scatter_(-1, top_i, h.gather(-1, top_i))
    x_aux = decode(f_aux, W_dec, b_dec)
    r_aux = x_aux - e
    mask_aux = torch.zeros_like(h).scatter_(-1, top_i, 1.0)
    return alpha * r_aux.pow(2).mean(), (f_aux, mask_aux), r_aux


@dataclasses.dataclass
class ForwardOut:
    h: Tensor
    f: Tensor
    mask: Tensor
    x_hat: Tensor
    r: Tensor
    mse: Tensor
    sparsity: Tensor
    l0: Tensor
    l1: Tensor
    aux: Tensor
    n_dead: int
    f_aux: Tensor | None
    mask_aux: Tensor | None
    r_aux: Tensor | None
    x_hats: Tensor | None = None  # [B, P, D] when Matryoshka prefixes are in use
    prefixes: list | None = None

    @property
    def loss(self) -> Tensor:
        """objectives.py:76-78."""
        return self.mse + self.sparsity + self.aux


def forward(cfg: OracleConfig, st: OracleState, x: Tensor, training: bool = True, prefixes=None,
            topk_idx: Tensor | None = None, aux_idx=None) -> ForwardOut:
    """objectives.py:101-156.  `prefixes` = the sorted cut points sample_prefixes() drew for this step (last one =
    d_sae); None / a single cut is the Matryoshka(n_prefixes=1) case.
    `topk_idx` (int64 [B, k]) / `aux_idx` (int64 [B, k_use], or a callable returning it, evaluated only when AuxK is
    live): the kernels' selections; adopted after adopt_selection() has verified that they differ from torch.topk's
    only at fp32-level ties."""
    h = encode_pre(x, st.W_enc, st.b_enc)
    if cfg.activation == "topk" and topk_idx is not None:
        _, own = torch.topk(h, min(cfg.top_k, h.shape[1]), dim=-1)
        topk_idx = adopt_selection(h, own, topk_idx, what="top-k")
        mask = torch.zeros_like(h).scatter(-1, topk_idx, 1.0)
        f = mask * h
    elif cfg.activation == "topk":
        f, mask = topk_activation(h, cfg.top_k)
    elif cfg.activation == "relu":
        f, mask = relu_activation(h)
    elif cfg.activation == "batchtopk":
        f, mask, st.threshold = batch_topk_activation(h, cfg.top_k, training, st.threshold, cfg.batch_momentum,
                                                      flat_idx=topk_idx)
    else:
        raise ValueError(cfg.activation)

    dead = None
    if training:
        st.toks_since_active, dead = dead_tracker_update(st.toks_since_active, f, cfg.dead_threshold_tokens)

    x_hats = None
    if prefixes is not None and len(prefixes) > 1:
        prefixes = [int(c) for c in prefixes]
        x_hats = decode_prefixes(f, st.W_dec, st.b_dec, prefixes)
        x_hat = x_hats[:, -1, :]  # AuxK and the logged reconstruction use the last (full) prefix, modeling.py:96
        r = x_hat - x
        mse = mean_squared_err(x_hats, x[:, None, :].expand_as(x_hats))  # objectives.py:133-138
    else:
        prefixes = None
        x_hat = decode(f, st.W_dec, st.b_dec)
        r = x_hat - x
        mse = mean_squared_err(x_hat, x)

    # modeling.py:30-31, 40-42
    l1 = f.abs().sum(dim=1).mean(dim=0)
    sparsity = l1 * cfg.l1_coeff if cfg.l1_coeff != 0.0 else torch.zeros(())
    # objectives.py:150-151
    l0 = (f != 0).float().sum(dim=1).mean(dim=0)

    aux, fa, r_aux = torch.zeros(()), None, None
    n_dead = 0
    if training and dead is not None:
        n_dead = int(dead.sum())
        if cfg.aux:
            if callable(aux_idx):
                aux_idx = aux_idx(n_dead)
            aux, fa, r_aux = auxk(h, r, dead, st.W_dec, st.b_dec, cfg.k_aux, cfg.aux_alpha, aux_idx=aux_idx)
    f_aux, mask_aux = fa if fa is not None else (None, None)
    return ForwardOut(h, f, mask, x_hat, r, mse, sparsity, l0, l1, aux, n_dead, f_aux, mask_aux, r_aux, x_hats, prefixes)


# --------------------------------------------------------------------------------------
# backward (what autograd computes for train.py:347-348), projection, clip, Adam, schedule
# --------------------------------------------------------------------------------------
def backward(cfg: OracleConfig, st: OracleState, x: Tensor, out: ForwardOut) -> dict:
    """Gradients of loss = mse + sparsity + aux w.r.t. the four parameters.
    G = 2 r/(B D);  G_a = 2 alpha r_a/(B D)
    gW_dec = f^T G + f_a^T G_a ; gb_dec = sum_b G + sum_b G_a
    dh = mask*(G W_dec^T) + mask_a*(G_a W_dec^T) (+ coeff*sign(f)/B on the active set for L1)
    gW_enc = x^T dh ; gb_enc = sum_b dh."""
    B, D = x.shape
    if out.x_hats is not None:
        # Matryoshka: loss_mse = mean over B*P*D of (x_hats - x)^2.  d/d x_hats[:, i] = G_i = 2 r_i / (B P D); column j
        # (in block c(j)) feeds every prefix i >= c(j), so it sees the suffix sum Gs_c = sum_{i >= c} G_i.
        P = out.x_hats.shape[1]
        Gi = (out.x_hats - x[:, None, :]) * (2.0 / (B * P * D))
        Gs = torch.flip(torch.cumsum(torch.flip(Gi, dims=[1]), dim=1), dims=[1])  # [B, P, D] suffix sums
        cuts = [0] + out.prefixes
        gW_dec = torch.zeros_like(st.W_dec)
        df = torch.zeros_like(out.f)
        for c, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
            gW_dec[a:b] = out.f[:, a:b].T @ Gs[:, c]
            df[:, a:b] = Gs[:, c] @ st.W_dec[a:b].T
        gb_dec = Gs[:, 0].sum(dim=0)
    else:
        G = out.r * (2.0 / (B * D))
        gW_dec = out.f.T @ G
        gb_dec = G.sum(dim=0)
        df = G @ st.W_dec.T
    if cfg.l1_coeff != 0.0:
        df = df + (cfg.l1_coeff / B) * torch.sign(out.f)
    dh = out.mask * df
    if out.f_aux is not None:
        Ga = out.r_aux * (2.0 * cfg.aux_alpha / (B * D))
        gW_dec = gW_dec + out.f_aux.T @ Ga
        gb_dec = gb_dec + Ga.sum(dim=0)
        dh = dh + out.mask_aux * (Ga @ st.W_dec.T)
    return dict(W_enc=x.T @ dh, b_enc=dh.sum(dim=0), W_dec=gW_dec, b_dec=gb_dec)


def remove_parallel_grads(gW_dec: Tensor, W_dec: Tensor) -> Tensor:
    """modeling.py:419-445: g_j -= (<g_j,w_j>/||w_j||^2) w_j; rows with ||w_j||^2 == 0 untouched."""
    par = (gW_dec * W_dec).sum(dim=1)
    nsq = (W_dec * W_dec).sum(dim=1)
    scales = torch.where(nsq > 0, par / torch.where(nsq > 0, nsq, torch.ones_like(nsq)), torch.zeros_like(par))
    return gW_dec - scales[:, None] * W_dec


PARAM_ORDER = ("W_dec", "b_dec", "W_enc", "b_enc")  # nn.Module registration order, modeling.py:312-327


def clip_grad_norm(grads: dict, max_norm: float):
    """torch.nn.utils.clip_grad_norm_ as called at train.py:358-360:
    n = ||(all grads)||_2 ; coef = min(1, max_norm/(n+1e-6)) ; g *= coef.  Returns (grads, n)."""
    norms = torch.stack([torch.linalg.vector_norm(grads[k]) for k in PARAM_ORDER])
    total = torch.linalg.vector_norm(norms)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return {k: g * coef for k, g in grads.items()}, total


def adam_step(cfg: OracleConfig, st: OracleState, grads: dict) -> None:
    """torch.optim.Adam(fused=True) defaults (train.py:294,444-446): t += 1;
    m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ;
    p -= (lr/(1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps).  lr is the value assigned at the END of
    the previous step (0.0 on the first step)."""
    st.t += 1
    bc1 = 1.0 - cfg.beta1**st.t
    bc2 = 1.0 - cfg.beta2**st.t
    step_size = st.lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    for k, p in st.params().items():
        g = grads[k]
        st.m[k] = st.m[k] + (g - st.m[k]) * (1.0 - cfg.beta1)  # lerp form used by torch's fused kernel
        st.v[k] = st.v[k] * cfg.beta2 + (1.0 - cfg.beta2) * g * g
        denom = st.v[k].sqrt() / bc2_sqrt + cfg.eps
        setattr(st, k, p - step_size * (st.m[k] / denom))


def warmup_cosine(step: int, n_warmup: int, peak: float, n_steps: int, init: float = 0.0, final: float = 0.0) -> float:
    """scheduling.py:58-68 evaluated at the already-incremented 1-based `_step`."""
    if step < n_warmup:
        return init + (peak - init) * (step / n_warmup)
    if step < n_steps:
        progress = (step - n_warmup) / (n_steps - n_warmup)
        return final + (peak - final) * (1 + math.cos(math.pi * progress)) / 2
    return final


def train_step(cfg: OracleConfig, st: OracleState, x: Tensor, prefixes=None, topk_idx: Tensor | None = None,
               aux_idx=None) -> dict:
    """One iteration of train.py:332-460 (log block excluded).  Mutates `st`; returns scalars and,
    for parity tests, the clipped gradients.  `topk_idx`, `aux_idx`: see forward()."""
    if cfg.normalize_w_dec:
        st.W_dec = normalize_w_dec(st.W_dec)  # train.py:334-335
    out = forward(cfg, st, x, training=True, prefixes=prefixes, topk_idx=topk_idx, aux_idx=aux_idx)  # train.py:341
    grads = backward(cfg, st, x, out)  # train.py:348
    if cfg.remove_parallel_grads:
        grads["W_dec"] = remove_parallel_grads(grads["W_dec"], st.W_dec)  # train.py:352
    grads, gnorm = clip_grad_norm(grads, cfg.grad_clip)  # train.py:358-360
    adam_step(cfg, st, grads)  # train.py:444-446
    st.sched_step += 1
    st.lr = warmup_cosine(st.sched_step, cfg.n_lr_warmup, cfg.lr, cfg.n_steps)  # train.py:449-451
    return dict(
        loss=float(out.loss), mse=float(out.mse), aux=float(out.aux), sparsity=float(out.sparsity),
        l0=float(out.l0), l1=float(out.l1), n_dead=out.n_dead, grad_norm=float(gnorm), grads=grads, out=out,
    )


def eval_forward(cfg: OracleConfig, st: OracleState, x: Tensor, prefixes=None) -> ForwardOut:
    """Objective in eval mode (train.py:526-527,559): no dead tracking, aux = 0."""
    return forward(cfg, st, x, training=False, prefixes=prefixes)


# --------------------------------------------------------------------------------------
# log block (train.py:365-442): metrics computed every `log_every` steps from the batch, the forward outputs and W_dec
# --------------------------------------------------------------------------------------
def dictionary_coherence(W_dec: Tensor, block: int = 1024) -> Tensor:
    """train.py:415-421: W_norm = W / W.norm(dim=1, keepdim=True); (W_norm @ W_norm.T).abs().triu(1).max().
    Evaluated block-row by block-row so that the [S, S] matrix never exists; same fp32 arithmetic per element."""
    S = W_dec.shape[0]
    Wn = W_dec / W_dec.norm(dim=1, keepdim=True)
    best = torch.zeros((), dtype=W_dec.dtype, device=W_dec.device)
    cols = torch.arange(S, device=W_dec.device)[None, :]
    for a in range(0, S, block):
        G = (Wn[a:a + block] @ Wn.T).abs()
        rows = torch.arange(a, min(a + block, S), device=W_dec.device)[:, None]
        best = torch.maximum(best, torch.where(cols > rows, G, torch.zeros_like(G)).max())
    return best


def log_block_metrics(x: Tensor, x_hat: Tensor, f_x: Tensor, W_dec: Tensor) -> dict:
    """train.py:380-423 for one SAE: `x` = acts_BD, `x_hat` = fwd.x_hats[:, -1, :], `f_x` = fwd.f_x."""
    x64 = x.to(torch.float64)
    n = x64.shape[0]
    sum_vec = x64.sum(dim=0)
    sse_baseline = float(torch.sum(x64 * x64) - torch.dot(sum_vec, sum_vec) / n)  # train.py:381-391
    residual = x - x_hat
    sse_sae = float(torch.sum(residual.to(torch.float64) ** 2))  # train.py:401-403
    return dict(
        explained_variance=float(1 - residual.var() / x.var()),  # train.py:407 (unbiased var over all elements)
        dead_unit_pct=float(((f_x.abs() > 1e-12).sum(0) == 0).float().mean()),  # train.py:410
        dictionary_coherence=float(dictionary_coherence(W_dec)),
        avg_decoder_row_norm=float(W_dec.norm(dim=1).mean()),  # train.py:420
        sse_sae=sse_sae,
        sse_baseline=sse_baseline,
        normalized_mse=sse_sae / sse_baseline,  # train.py:404-406
    )


# --------------------------------------------------------------------------------------
# evaluate() (train.py:510-618)
# --------------------------------------------------------------------------------------
def evaluate(cfg: OracleConfig, st: OracleState, batches, almost_dead_lim: float = 1e-7, dense_lim: float = 1e-2) -> dict:
    """train.py:536-616 for one SAE over an iterable of x[B, D] batches (eval-mode objective forward per batch)."""
    S, D = st.W_dec.shape
    n_fired = torch.zeros(S)
    values = torch.zeros(S)
    l0_sum = l1_sum = mse_sum = 0.0
    sse_sae = torch.zeros((), dtype=torch.float64)
    sum_sq = torch.zeros((), dtype=torch.float64)
    sum_vec = torch.zeros(D, dtype=torch.float64)
    n_tokens = 0
    for x in batches:
        x64 = x.to(torch.float64)
        sum_sq += torch.sum(x64 * x64)  # train.py:548
        sum_vec += x64.sum(dim=0)  # train.py:549
        n_tokens += x.shape[0]
        out = eval_forward(cfg, st, x)
        sse_sae += torch.sum((x - out.x_hat).to(torch.float64) ** 2)  # train.py:558-559
        n_fired += (out.f > 0).sum(dim=0)  # train.py:560-562
        values += out.f.sum(dim=0)  # train.py:563
        l0_sum += float(out.l0) * x.shape[0]  # train.py:564-566
        l1_sum += float(out.l1) * x.shape[0]
        mse_sum += float(out.mse) * x.shape[0]
    sse_baseline = float(sum_sq - torch.dot(sum_vec, sum_vec) / n_tokens)  # train.py:570-571
    freqs = n_fired / n_tokens
    return dict(
        l0=l0_sum / n_tokens, l1=l1_sum / n_tokens, mse=mse_sum / n_tokens, normalized_mse=float(sse_sae) / sse_baseline,
        sse_sae=float(sse_sae), sse_baseline=sse_baseline, n_dead=int((freqs == 0).sum()),
        n_almost_dead=int((freqs < almost_dead_lim).sum()), n_dense=int((freqs > dense_lim).sum()), freqs=freqs,
        mean_values=values / n_fired,
    )
