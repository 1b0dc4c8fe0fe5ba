"""GPU parity (run with `-m gpu` on a B200): replay the golden runs recorded from the LIVE reference
(tests/golden/*.npz, written by oracle/gen_golden.py) through the CUDA path via the C ABI, step by step:
losses, n_dead, grad norm, the four clipped gradients, x_hat, final parameters, Adam moments, the dead
tracker state and an eval-mode forward.

Tolerance: the north-star bar is 1e-4 relative (fp32); the CUDA path differs from the reference only in
fp32 summation order, so the assertions use TOL = 2e-5 for vectors and scalars (measured ~1e-7..1e-6)."""
import pytest
import torch

from tests.golden_util import CASES, load_case, rel_l2, t

pytestmark = pytest.mark.gpu

TOL = 2e-5
# the dense ReLU path computes its five contractions as bf16 split products (~2^-17 relative per product) instead of
# fp32 FMAs: still an order of magnitude inside the 1e-4 bar, but not at summation-order level
TOL_DENSE = 6e-5
TOPK_CASES = [c for c in CASES if "relu" not in c]


def _engine_for(meta, cfg, B):
    from saev_b200.engine import Engine, EngineConfig

    return Engine(
        EngineConfig(
            d_model=cfg.d_model, d_sae=cfg.d_sae, top_k=cfg.top_k, activation=cfg.activation, aux=cfg.aux,
            k_aux=cfg.k_aux, aux_alpha=cfg.aux_alpha, l1_coeff=cfg.l1_coeff,
            dead_threshold_tokens=cfg.dead_threshold_tokens, remove_parallel_grads=cfg.remove_parallel_grads,
            normalize_w_dec=cfg.normalize_w_dec, max_batch=B,
        )
    )


@pytest.mark.parametrize("name", CASES)
def test_cuda_path_replays_reference_run(name):
    z, meta, cfg = load_case(name)
    TOL = TOL_DENSE if cfg.activation == "relu" else globals()["TOL"]
    B = meta["B"]
    eng = _engine_for(meta, cfg, B)
    eng.load_params(t(z["init_W_enc"]), t(z["init_b_enc"]), t(z["init_W_dec"]), t(z["init_b_dec"]))
    xs = t(z["xs"]).cuda()
    grad_steps = list(z["grad_steps"])
    for step in range(meta["n_steps"]):
        x = xs[step].contiguous()
        lr = float(z["rec_lr"][step])
        eng.normalize_w_dec()
        eng.forward(x, training=True)
        eng.backward(x)
        eng.grad_sumsq()
        ld = eng.loss_dict()
        for key in ("mse", "aux", "sparsity", "l0", "l1", "loss"):
            assert ld[key] == pytest.approx(float(z[f"rec_{key}"][step]), rel=TOL, abs=1e-7), (step, key)
        assert int(ld["n_dead"]) == int(z["rec_n_dead"][step]), step
        gn = float(eng.sumsq.sqrt())
        assert gn == pytest.approx(float(z["rec_grad_norm"][step]), rel=TOL), step
        if step in grad_steps:
            i = grad_steps.index(step)
            coef = min(1.0, cfg.grad_clip / (gn + 1e-6))  # reference grads were recorded after clip_grad_norm_
            ours = dict(W_enc=eng.gW_enc_t.t(), b_enc=eng.gb_enc, W_dec=eng.gW_dec, b_dec=eng.gb_dec)
            for k, g in ours.items():
                assert rel_l2((g * coef).cpu(), z[f"grads_{k}"][i]) < TOL, (step, k)
            assert rel_l2(eng.x_hat(x).cpu(), z["x_hat"][i]) < TOL, step
        eng.adam_step(lr, max_norm=cfg.grad_clip)
    ours = dict(W_enc=eng.W_enc_t.t(), b_enc=eng.b_enc, W_dec=eng.W_dec, b_dec=eng.b_dec)
    S, D = eng.S, eng.D
    m = dict(zip(("W_enc", "b_enc", "W_dec", "b_dec"), eng._views(eng.m)))
    v = dict(zip(("W_enc", "b_enc", "W_dec", "b_dec"), eng._views(eng.v)))
    m["W_enc"], v["W_enc"] = m["W_enc"].t(), v["W_enc"].t()
    for k in ours:
        assert rel_l2(ours[k].cpu(), z[f"final_{k}"]) < TOL, k
        assert rel_l2(m[k].cpu(), z[f"m_{k}"]) < TOL, k
        assert rel_l2(v[k].cpu(), z[f"v_{k}"]) < 10 * TOL, k
    assert torch.equal(eng.toks_since_active.cpu(), t(z["toks_since_active"]))
    # eval-mode forward (train.py:526-527,559): no dead tracking, aux = 0
    x = xs[-1].contiguous()
    eng.forward(x, training=False)
    ld = eng.loss_dict()
    assert ld["mse"] == pytest.approx(float(z["eval_mse"]), rel=TOL)
    assert ld["l0"] == pytest.approx(float(z["eval_l0"]), rel=TOL)
    assert ld["aux"] == 0.0 and ld["n_dead"] == 0.0
    assert rel_l2(eng.x_hat(x).cpu(), z["eval_x_hat"]) < TOL
    if cfg.activation == "topk":
        assert eng.unsafe_rows() == 0


def test_fused_renorm_equals_start_of_step_normalize():
    """Hoisting normalize_w_dec (train.py:334-335) into the Adam tail gives the same next-step state."""
    z, meta, cfg = load_case("c1_topk")
    B = meta["B"]
    a, b = _engine_for(meta, cfg, B), _engine_for(meta, cfg, B)
    for e in (a, b):
        e.load_params(t(z["init_W_enc"]), t(z["init_b_enc"]), t(z["init_W_dec"]), t(z["init_b_dec"]))
    xs = t(z["xs"]).cuda()
    for step in range(4):
        x = xs[step].contiguous()
        a.train_step(x, 1e-3, fused_renorm=False)
        b.train_step(x, 1e-3, fused_renorm=True, pre_normalized=step > 0)
    a.normalize_w_dec()
    assert rel_l2(b.W_dec.cpu(), a.W_dec.cpu()) < 1e-6
    assert rel_l2(b.W_enc_t.cpu(), a.W_enc_t.cpu()) < 1e-6
