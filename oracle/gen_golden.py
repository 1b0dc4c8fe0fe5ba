"""Generate tests/golden/*.npz by running the LIVE reference (read-only at /root/reference/src)
on fixed seeds.  TEST INFRASTRUCTURE ONLY; runs in the build container (the reference does not
travel to the GPU box), the resulting vectors are committed.

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz

The loop body below is the text of /root/reference/src/saev/framework/train.py:332-460 with the
logging block removed, driving the reference's own `saev.nn.SparseAutoencoder`,
`saev.nn.get_objective`, `torch.optim.Adam(fused=True)` and
`saev.utils.scheduling.WarmupCosine` objects -- nothing from this repository is on that path.
"""

import pathlib
import sys

import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / "ref_stubs"))
sys.path.insert(0, "/root/reference/src")

import saev.nn  # noqa: E402
import saev.nn.modeling as M  # noqa: E402
import saev.nn.objectives as O  # noqa: E402
import saev.utils.scheduling  # noqa: E402

GOLDEN = HERE.parent / "tests" / "golden"


def planted_batches(gen, n_steps, B, D, n_atoms, k_active, noise):
    """x = z A + noise with a planted sparse dictionary, so MSE falls and latents die."""
    A = torch.randn(n_atoms, D, generator=gen)
    A = A / A.norm(dim=1, keepdim=True)
    xs = []
    for _ in range(n_steps):
        z = torch.zeros(B, n_atoms)
        idx = torch.stack([torch.randperm(n_atoms, generator=gen)[:k_active] for _ in range(B)])
        z.scatter_(1, idx, torch.rand(B, k_active, generator=gen) * 3 + 0.5)
        xs.append(z @ A + noise * torch.randn(B, D, generator=gen))
    return torch.stack(xs)


def run_case(name, *, D, S, B, n_steps, activation, dead_thr, lr, n_warmup, sched_steps, grad_clip=1.0,
             data="gauss", seed=0, save_grads_every=1, normalize=True, remove_parallel=True, b_enc_shift=0.0,
             n_prefixes=1):
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    sae_cfg = M.SparseAutoencoderConfig(
        d_model=D, d_sae=S, activation=activation, reinit_blend=0.0,
        normalize_w_dec=normalize, remove_parallel_grads=remove_parallel,
    )
    sae = saev.nn.SparseAutoencoder(sae_cfg)
    objective = saev.nn.get_objective(O.Matryoshka(n_prefixes=n_prefixes, dead_threshold_tokens=dead_thr))
    # record the prefix cuts the reference draws (torch.multinomial on the global RNG, objectives.py:125,189-191)
    drawn = []
    orig_sample = O.sample_prefixes

    def recording_sample(*a, **k):
        out = orig_sample(*a, **k)
        drawn.append(out.clone().numpy())
        return out

    O.sample_prefixes = recording_sample
    if b_enc_shift != 0.0:  # push the upper half of the latents negative so that they die (ReLU case)
        with torch.no_grad():
            sae.b_enc[S // 2 :] = b_enc_shift
    init = {k: v.detach().clone().numpy() for k, v in sae.state_dict().items()}

    if data == "gauss":
        xs = torch.randn(n_steps, B, D, generator=gen)
    else:
        xs = planted_batches(gen, n_steps, B, D, n_atoms=max(8, S // 8), k_active=4, noise=0.05)

    # train.py:118,294,311-316
    pg = {"params": sae.parameters(), "lr": 0.0}
    opt = torch.optim.Adam([pg], fused=True)
    sched = saev.utils.scheduling.WarmupCosine(0.0, n_warmup, lr, sched_steps, 0.0)
    sae.train()
    objective.train()

    rec = {k: [] for k in ("loss", "mse", "aux", "sparsity", "l0", "l1", "n_dead", "grad_norm", "lr", "threshold")}
    grads_rec, xhat_rec, grad_steps = {k: [] for k, _ in sae.named_parameters()}, [], []
    for step in range(n_steps):
        acts = xs[step]
        sae.normalize_w_dec()  # train.py:334-335
        loss, fwd = objective(sae, acts)  # train.py:341
        loss.loss.backward()  # train.py:348
        sae.remove_parallel_grads()  # train.py:352
        gn = torch.nn.utils.clip_grad_norm_(sae.parameters(), max_norm=grad_clip)  # train.py:358-360
        rec["lr"].append(opt.param_groups[0]["lr"])
        for key in ("loss", "mse", "aux", "sparsity", "l0", "l1"):
            rec[key].append(float(getattr(loss, key)))
        rec["n_dead"].append(int(loss.n_dead))
        rec["grad_norm"].append(float(gn))
        # BatchTopKActivation.threshold after this step's EMA update (modeling.py:237-242); 0 for the other activations
        rec["threshold"].append(float(getattr(sae.activation, "threshold", torch.tensor(0.0))))
        if step % save_grads_every == 0 or step == n_steps - 1:
            grad_steps.append(step)
            for k, p in sae.named_parameters():
                grads_rec[k].append(p.grad.detach().clone().numpy())
            xhat_rec.append(fwd.x_hats[:, -1, :].detach().clone().numpy())
        opt.step()  # train.py:444-446
        opt.param_groups[0]["lr"] = sched.step()  # train.py:449-451
        opt.zero_grad()  # train.py:456-458

    out = {f"init_{k}": v for k, v in init.items()}
    out.update({f"final_{k}": v.detach().numpy() for k, v in sae.state_dict().items()})
    for p_name, p in sae.named_parameters():
        st = opt.state[p]
        out[f"m_{p_name}"] = st["exp_avg"].numpy()
        out[f"v_{p_name}"] = st["exp_avg_sq"].numpy()
    out["toks_since_active"] = (
        objective.toks_since_active.numpy() if objective.toks_since_active is not None else np.zeros(S, np.int64)
    )
    out["xs"] = xs.numpy()
    out["grad_steps"] = np.array(grad_steps)
    for k in grads_rec:
        out[f"grads_{k}"] = np.stack(grads_rec[k])
    out["x_hat"] = np.stack(xhat_rec)
    for k, v in rec.items():
        out[f"rec_{k}"] = np.array(v)

    # eval-mode forward on the last batch (train.py:526-527,559)
    sae.eval()
    objective.eval()
    with torch.no_grad():
        eloss, efwd = objective(sae, xs[-1])
    out["eval_mse"] = np.array(float(eloss.mse))
    out["eval_l0"] = np.array(float(eloss.l0))
    out["eval_l1"] = np.array(float(eloss.l1))
    out["eval_x_hat"] = efwd.x_hats[:, -1, :].numpy()
    O.sample_prefixes = orig_sample
    out["prefixes"] = np.stack(drawn)  # [n_steps + 1 (eval), n_prefixes]
    out["eval_x_hats_all"] = efwd.x_hats.numpy()

    meta = dict(D=D, S=S, B=B, n_steps=n_steps, dead_thr=dead_thr, lr=lr, n_warmup=n_warmup,
                sched_steps=sched_steps, grad_clip=grad_clip, normalize=int(normalize),
                remove_parallel=int(remove_parallel), n_prefixes=n_prefixes, seed=seed)
    if isinstance(activation, M.TopK):
        meta.update(act=0, top_k=activation.top_k)
    elif isinstance(activation, M.BatchTopK):
        meta.update(act=2, top_k=activation.top_k, momentum=activation.momentum)
    else:
        meta.update(act=1, top_k=0)
    sp = activation.sparsity
    meta["l1_coeff"] = sp.coeff if isinstance(sp, M.L1Sparsity) else 0.0
    ax = activation.aux
    meta.update(aux=int(isinstance(ax, M.AuxK)), k_aux=getattr(ax, "k_aux", 0), alpha=getattr(ax, "alpha", 0.0))
    for k, v in meta.items():
        out[f"meta_{k}"] = np.array(v)
    GOLDEN.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLDEN / f"{name}.npz", **out)
    print(f"{name}: mse {rec['mse'][0]:.5f}->{rec['mse'][-1]:.5f} aux[-1]={rec['aux'][-1]:.5g} "
          f"n_dead[-1]={rec['n_dead'][-1]} gnorm[-1]={rec['grad_norm'][-1]:.4f} "
          f"size={(GOLDEN / f'{name}.npz').stat().st_size / 1e6:.2f}MB")


def main():
    import sys

    only = set(sys.argv[1:])
    global run_case
    _run = run_case

    def run_case(name, **kw):  # `python oracle/gen_golden.py <name> ...` regenerates only the named cases
        if not only or name in only:
            _run(name, **kw)

    # (1) survey appendix-A case: AuxK live from step 3, clipping active.
    run_case("tiny_topk_auxk", D=32, S=256, B=64, n_steps=12,
             activation=M.TopK(top_k=8, aux=M.AuxK(k_aux=16, alpha=1 / 32)),
             dead_thr=3 * 64, lr=1e-2, n_warmup=4, sched_steps=12, data="planted")
    # (2) same but k_aux larger than n_dead will ever be  -> k_use = n_dead clamp (test_auxk.py:71-81)
    run_case("tiny_topk_auxk_clamp", D=32, S=128, B=64, n_steps=8,
             activation=M.TopK(top_k=4, aux=M.AuxK(k_aux=512, alpha=1 / 32)),
             dead_thr=2 * 64, lr=5e-3, n_warmup=3, sched_steps=8, data="planted", seed=3)
    # (3) TopK without aux, no normalisation/projection flags
    run_case("tiny_topk_noaux_noproj", D=32, S=256, B=64, n_steps=6,
             activation=M.TopK(top_k=8, aux=M.NoAux()),
             dead_thr=10_000_000, lr=1e-2, n_warmup=2, sched_steps=6, normalize=False, remove_parallel=False, seed=5)
    # (4) cfg-5 family: ReLU + L1 + AuxK
    run_case("tiny_relu_l1_auxk", D=32, S=256, B=64, n_steps=10,
             activation=M.Relu(sparsity=M.L1Sparsity(coeff=4e-4), aux=M.AuxK(k_aux=16, alpha=1 / 32)),
             dead_thr=3 * 64, lr=1e-2, n_warmup=4, sched_steps=10, data="planted", seed=7, b_enc_shift=-1.5)
    # (5) BASELINE.json configs[0]: D=128 S=512 K=16 B=256, Gaussian activations, defaults otherwise
    run_case("c1_topk", D=128, S=512, B=256, n_steps=8,
             activation=M.TopK(top_k=16), dead_thr=10_000_000, lr=4e-4, n_warmup=500, sched_steps=1000,
             save_grads_every=7, seed=11)
    # (6) c1 with the dead threshold lowered so AuxK goes live on Gaussian data is unlikely (every latent
    #     fires); use planted data with a wider dictionary instead.
    run_case("c1_topk_auxk_live", D=128, S=512, B=256, n_steps=8,
             activation=M.TopK(top_k=16, aux=M.AuxK(k_aux=64, alpha=1 / 32)),
             dead_thr=2 * 256, lr=2e-3, n_warmup=3, sched_steps=8, data="planted", save_grads_every=7, seed=13)
    # (7) the reference's DEFAULT objective family: Matryoshka with several random prefix cuts per step
    #     (objectives.py:22,124-138; modeling.py:377-406), AuxK live, clipping active
    run_case("tiny_topk_matryoshka", D=32, S=256, B=64, n_steps=10,
             activation=M.TopK(top_k=8, aux=M.AuxK(k_aux=16, alpha=1 / 32)),
             dead_thr=3 * 64, lr=1e-2, n_warmup=4, sched_steps=10, data="planted", seed=17, n_prefixes=5)
    run_case("c1_topk_matryoshka", D=128, S=512, B=256, n_steps=6,
             activation=M.TopK(top_k=16, aux=M.AuxK(k_aux=64, alpha=1 / 32)),
             dead_thr=2 * 256, lr=2e-3, n_warmup=3, sched_steps=6, data="planted", save_grads_every=5, seed=19,
             n_prefixes=10)
    # (7b) ReLU + L1 + AuxK under Matryoshka prefixes (the dense path with per-block contractions)
    run_case("tiny_relu_matryoshka", D=32, S=256, B=64, n_steps=8,
             activation=M.Relu(sparsity=M.L1Sparsity(coeff=4e-4), aux=M.AuxK(k_aux=16, alpha=1 / 32)),
             dead_thr=3 * 64, lr=1e-2, n_warmup=4, sched_steps=8, data="planted", seed=37, b_enc_shift=-1.5, n_prefixes=4)
    # (8) BatchTopK (modeling.py:183-244): batch-wide top-(k B) selection in training, EMA threshold, JumpReLU in the
    #     eval forward; planted data so that the per-row counts differ and latents die (AuxK live)
    run_case("tiny_batchtopk_auxk", D=32, S=256, B=64, n_steps=10,
             activation=M.BatchTopK(top_k=8, aux=M.AuxK(k_aux=16, alpha=1 / 32)),
             dead_thr=3 * 64, lr=1e-2, n_warmup=4, sched_steps=10, data="planted", seed=23)
    run_case("c1_batchtopk", D=128, S=512, B=256, n_steps=6,
             activation=M.BatchTopK(top_k=16, momentum=0.4, aux=M.AuxK(k_aux=64, alpha=1 / 32)),
             dead_thr=2 * 256, lr=2e-3, n_warmup=3, sched_steps=6, data="planted", save_grads_every=5, seed=29)
    # BatchTopK under the reference's default objective family (Matryoshka prefixes)
    run_case("c1_batchtopk_matryoshka", D=128, S=512, B=256, n_steps=6,
             activation=M.BatchTopK(top_k=8, momentum=0.5, aux=M.AuxK(k_aux=64, alpha=1 / 32)),
             dead_thr=2 * 256, lr=2e-3, n_warmup=3, sched_steps=6, data="planted", save_grads_every=5, seed=31,
             n_prefixes=5)


if __name__ == "__main__":
    main()
