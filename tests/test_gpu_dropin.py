"""GPU tests of the reference-facing Python surface (run with `-m gpu`): the golden runs recorded from the LIVE
reference (tests/golden/*.npz) are replayed through `saev_b200.nn` + `saev_b200.optim` + `saev_b200.scheduling`
with the statements of saev's loop body (/root/reference/src/saev/framework/train.py:332-460) -- the same calls
`saev_b200.install()` routes an unmodified saev checkout through.  Tolerance 2e-5 relative (north-star bar: 1e-4)."""
import pytest
import torch

from saev_b200 import nn, optim, scheduling
from tests.golden_util import CASES, load_case, rel_l2, t

pytestmark = pytest.mark.gpu
TOL = 2e-5
TOL_DENSE = 2e-5  # ReLU: three-piece bf16 split contractions, see test_gpu_golden.py


def _make(z, meta, cfg):
    aux = nn.AuxK(k_aux=cfg.k_aux, alpha=cfg.aux_alpha) if cfg.aux else nn.NoAux()
    if cfg.activation == "relu":
        sp = nn.L1Sparsity(coeff=cfg.l1_coeff) if cfg.l1_coeff else nn.NoSparsity()
        act = nn.Relu(sparsity=sp, aux=aux)
    elif cfg.activation == "batchtopk":
        act = nn.BatchTopK(top_k=cfg.top_k, momentum=cfg.batch_momentum, aux=aux)
    else:
        act = nn.TopK(top_k=cfg.top_k, aux=aux)
    sae_cfg = nn.SparseAutoencoderConfig(
        d_model=cfg.d_model, d_sae=cfg.d_sae, activation=act, reinit_blend=0.0,
        remove_parallel_grads=cfg.remove_parallel_grads, normalize_w_dec=cfg.normalize_w_dec)
    sae = nn.SparseAutoencoder(sae_cfg)
    keys = ("W_dec", "b_dec", "W_enc", "b_enc") + (("activation.threshold",) if cfg.activation == "batchtopk" else ())
    sae.load_state_dict({k: t(z[f"init_{k}"]) for k in keys})
    objective = nn.get_objective(nn.Matryoshka(n_prefixes=max(1, meta.get("n_prefixes", 1)),
                                               dead_threshold_tokens=cfg.dead_threshold_tokens))
    return sae, objective


@pytest.mark.parametrize("name", CASES)
def test_loop_body_through_the_nn_api(name):
    z, meta, cfg = load_case(name)
    TOL = TOL_DENSE if cfg.activation == "relu" else globals()["TOL"]
    if meta.get("n_prefixes", 1) > 1:
        # replay the reference's RNG stream: manual_seed, then the kaiming draw of SparseAutoencoder.__init__, then one
        # sample_prefixes() per forward (objectives.py:125) -- our mirror must consume the generator identically
        torch.manual_seed(meta["seed"])
    sae, objective = _make(z, meta, cfg)
    param_group = {"params": sae.parameters(), "lr": 0.0}  # train.py:118
    opt = optim.FusedAdam([param_group], fused=True)  # train.py:294 (constructed while the module is on the CPU)
    sched = scheduling.WarmupCosine(0.0, cfg.n_lr_warmup, cfg.lr, cfg.n_steps, 0.0)
    sae.train()
    sae = sae.to("cuda")
    objective.train()
    objective = objective.to("cuda")
    xs = t(z["xs"])
    grad_steps = list(z["grad_steps"])
    for step in range(meta["n_steps"]):
        acts = xs[step].to("cuda", non_blocking=True)
        sae.normalize_w_dec()
        loss, fwd = objective(sae, acts)
        loss.loss.backward()
        sae.remove_parallel_grads()
        grad_norm = optim.clip_grad_norm_(sae.parameters(), max_norm=cfg.grad_clip)
        m = loss.metrics()
        for key in ("mse", "aux", "sparsity", "l0", "l1", "loss"):
            assert m[key] == pytest.approx(float(z[f"rec_{key}"][step]), rel=TOL, abs=1e-7), (step, key)
        assert int(m["n_dead"]) == int(z["rec_n_dead"][step])
        assert grad_norm.item() == pytest.approx(float(z["rec_grad_norm"][step]), rel=TOL)
        assert opt.param_groups[0]["lr"] == pytest.approx(float(z["rec_lr"][step]), rel=1e-12, abs=0)
        if cfg.activation == "batchtopk":  # the EMA buffer of BatchTopKActivation (modeling.py:213,237-242)
            assert float(sae.activation.threshold) == pytest.approx(float(z["rec_threshold"][step]), rel=1e-5), step
        if step in grad_steps:
            i = grad_steps.index(step)
            coef = min(1.0, cfg.grad_clip / (grad_norm.item() + 1e-6))
            for k in ("W_enc", "b_enc", "W_dec", "b_dec"):
                g = getattr(sae, k).grad
                assert g.shape == getattr(sae, k).shape
                assert rel_l2((g * coef).cpu(), z[f"grads_{k}"][i]) < TOL, (step, k)
            assert fwd.x_hats.shape == (meta["B"], max(1, meta.get("n_prefixes", 1)), cfg.d_model)
            assert rel_l2(fwd.x_hats[:, -1, :].cpu(), z["x_hat"][i]) < 10 * TOL
            assert fwd.f_x.shape == (meta["B"], cfg.d_sae)
            if cfg.activation == "topk":
                assert int((fwd.f_x != 0).sum(1).max()) <= cfg.top_k
            elif cfg.activation == "batchtopk":  # exactly k * B survivors over the whole batch (modeling.py:206)
                assert int((fwd.f_x != 0).sum()) == cfg.top_k * meta["B"]
            else:
                assert bool((fwd.f_x >= 0).all())
        opt.step()
        opt.param_groups[0]["lr"] = sched.step()
        opt.zero_grad()
        assert sae.W_dec.grad is None
    sd = sae.state_dict()
    assert list(sd) == ["W_dec", "b_dec", "W_enc", "b_enc"] + (["activation.threshold"] if cfg.activation == "batchtopk" else [])
    for k in sd:
        assert rel_l2(sd[k].cpu(), z[f"final_{k}"]) < TOL, k
    assert torch.equal(objective.toks_since_active.cpu(), t(z["toks_since_active"]))
    # eval mode (train.py:526-527, 559)
    sae.eval()
    objective.eval()
    loss, fwd = objective(sae, xs[-1].to("cuda"))
    assert loss.mse.item() == pytest.approx(float(z["eval_mse"]), rel=TOL)
    assert loss.l0.item() == pytest.approx(float(z["eval_l0"]), rel=TOL)  # (BatchTopK: JumpReLU with the EMA threshold)
    assert loss.aux.item() == 0.0 and int(loss.n_dead) == 0
    loss.metrics()  # BatchTopK: raises if a row ran out of slots


def test_batchtopk_checkpoint_keeps_the_threshold(tmp_path):
    """modeling.py:213: `threshold` is a registered buffer, so it travels in the state_dict / checkpoint."""
    z, meta, cfg = load_case("tiny_batchtopk_auxk")
    sae, objective = _make(z, meta, cfg)
    sae, objective = sae.to("cuda").train(), objective.to("cuda").train()
    objective(sae, t(z["xs"])[0].to("cuda"))
    thr = float(sae.activation.threshold)
    assert thr == pytest.approx(float(z["rec_threshold"][0]), rel=1e-5)
    nn.dump(tmp_path / "sae.pt", sae)
    back = nn.load(tmp_path / "sae.pt")
    assert back.cfg == sae.cfg and float(back.activation.threshold) == thr


def test_batchtopk_row_out_of_slots_is_reported(monkeypatch):
    """With 16 slots per row (SAEV_B200_BATCHTOPK_CAP) the planted batches have rows that own more of the batch's
    k * B winners than they can hold: the selection kernel counts them and Loss.metrics() raises instead of training
    on a selection that differs from the reference's."""
    monkeypatch.setenv("SAEV_B200_BATCHTOPK_CAP", "16")
    z, meta, cfg = load_case("tiny_batchtopk_auxk")
    sae, objective = _make(z, meta, cfg)
    sae, objective = sae.to("cuda").train(), objective.to("cuda").train()
    loss, _ = objective(sae, t(z["xs"])[0].to("cuda"))
    assert sae.engine.cfg.top_k == 16 and sae.engine.batch_topk_stats()["truncated_rows"] > 0
    with pytest.raises(RuntimeError, match="slots per row"):
        loss.metrics()


def test_checkpoint_round_trip(tmp_path):
    z, meta, cfg = load_case("c1_topk")
    sae, _ = _make(z, meta, cfg)
    sae = sae.to("cuda")
    sae(torch.randn(8, cfg.d_model, device="cuda"))  # binds the engine: parameters now alias the flat buffers
    nn.dump(tmp_path / "sae.pt", sae)
    back = nn.load(tmp_path / "sae.pt")
    assert back.cfg == sae.cfg
    for k, v in sae.state_dict().items():
        assert torch.equal(back.state_dict()[k], v.cpu()), k


def test_unsupported_configs_fail_loudly():
    with pytest.raises(NotImplementedError, match="cannot even hold the average"):  # more than 128 actives per row
        nn.SparseAutoencoder(nn.SparseAutoencoderConfig(d_model=64, d_sae=512, activation=nn.BatchTopK(top_k=200),
                                                        reinit_blend=0.0)).to("cuda")(torch.randn(4, 64, device="cuda"))
