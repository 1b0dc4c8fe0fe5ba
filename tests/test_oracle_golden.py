"""Pin the CPU oracle (oracle/sae_oracle.py) against
(a) golden vectors produced by the live reference (oracle/gen_golden.py -> tests/golden/*.npz), and
(b) the hand-computed known answers in the reference's own tests (cited per test).
Runs on CPU (`-m "not gpu"`)."""
import math

import pytest
import torch

from oracle import sae_oracle as orc
from tests.golden_util import CASES, load_case, prefixes_of, rel_l2, t

TOL = 2e-5  # fp32, same op order up to reduction order; measured ~1e-7..3e-6


@pytest.mark.parametrize("name", CASES)
def test_oracle_replays_reference_run(name):
    z, meta, cfg = load_case(name)
    st = orc.OracleState.from_params(t(z["init_W_enc"]), t(z["init_b_enc"]), t(z["init_W_dec"]), t(z["init_b_dec"]))
    xs = t(z["xs"])
    grad_steps = list(z["grad_steps"])
    for step in range(meta["n_steps"]):
        assert st.lr == pytest.approx(z["rec_lr"][step], rel=1e-12, abs=0)
        out = orc.train_step(cfg, st, xs[step], prefixes=prefixes_of(z, step))
        for key in ("mse", "aux", "sparsity", "l0", "l1", "grad_norm", "loss"):
            assert out[key] == pytest.approx(float(z[f"rec_{key}"][step]), rel=TOL, abs=1e-7), (step, key)
        assert out["n_dead"] == int(z["rec_n_dead"][step]), step
        if cfg.activation == "batchtopk":  # EMA of the smallest positive survivor (modeling.py:237-242)
            assert st.threshold == pytest.approx(float(z["rec_threshold"][step]), rel=1e-6), step
        if step in grad_steps:
            i = grad_steps.index(step)
            for k in ("W_enc", "b_enc", "W_dec", "b_dec"):
                assert rel_l2(out["grads"][k], z[f"grads_{k}"][i]) < TOL, (step, k)
            assert rel_l2(out["out"].x_hat, z["x_hat"][i]) < TOL
    for k in ("W_enc", "b_enc", "W_dec", "b_dec"):
        assert rel_l2(getattr(st, k), z[f"final_{k}"]) < TOL, k
        assert rel_l2(st.m[k], z[f"m_{k}"]) < TOL, k
        assert rel_l2(st.v[k], z[f"v_{k}"]) < 10 * TOL, k
    assert torch.equal(st.toks_since_active, t(z["toks_since_active"]))
    ev = orc.eval_forward(cfg, st, xs[-1], prefixes=prefixes_of(z, meta["n_steps"]))
    if ev.x_hats is not None:
        assert rel_l2(ev.x_hats, z["eval_x_hats_all"]) < TOL
    assert float(ev.mse) == pytest.approx(float(z["eval_mse"]), rel=TOL)
    assert float(ev.l0) == pytest.approx(float(z["eval_l0"]), rel=TOL)
    assert rel_l2(ev.x_hat, z["eval_x_hat"]) < TOL


def test_sample_prefixes_reproduces_the_reference_draws():
    """Same torch CPU ops on the same global generator state => the same cuts the reference drew (objectives.py:158-201);
    the golden run seeded torch with meta['seed'] and drew once per forward."""
    from saev_b200 import nn as bnn

    for name in ("tiny_topk_matryoshka", "c1_topk_matryoshka"):
        z, meta, cfg = load_case(name)
        for fn in (orc.sample_prefixes, bnn.sample_prefixes):
            # reference order: manual_seed(seed); SparseAutoencoder(cfg) consumes kaiming_uniform_([S, D]) first
            torch.manual_seed(meta["seed"])
            torch.nn.init.kaiming_uniform_(torch.empty(meta["S"], meta["D"]))
            draws = [fn(meta["S"], meta["n_prefixes"]).tolist() for _ in range(meta["n_steps"] + 1)]
            assert draws == z["prefixes"].tolist(), (name, fn.__module__)
    assert orc.sample_prefixes(64, 1).tolist() == [64]


# ---- known answers restated from the reference's tests ---------------------------------------
def _identity_state(d):
    eye = torch.eye(d)
    return orc.OracleState.from_params(eye.clone(), torch.zeros(d), eye.clone(), torch.zeros(d))


def test_auxk_topk_value_matches_manual():
    """/root/reference/tests/test_auxk.py:39-50: aux = (3^2+4^2)/4 with identity weights, x=0, x_hat=0."""
    st = _identity_state(4)
    h = torch.tensor([[1.0, 2.0, 3.0, 4.0]])
    aux, _, _ = orc.auxk(h, torch.zeros(1, 4), torch.ones(4, dtype=torch.bool), st.W_dec, st.b_dec, 2, 1.0)
    assert float(aux) == pytest.approx((9 + 16) / 4)


def test_auxk_alpha_and_clamp_and_zero_dead():
    """test_auxk.py:25-36 (n_dead=0 -> 0), :53-68 (alpha scales), :71-81 (k clamps to n_dead -> 25/4)."""
    st = _identity_state(4)
    h = torch.tensor([[1.0, 2.0, 3.0, 4.0]])
    a1, _, _ = orc.auxk(h, torch.zeros(1, 4), torch.ones(4, dtype=torch.bool), st.W_dec, st.b_dec, 2, 1.0)
    a2, _, _ = orc.auxk(h, torch.zeros(1, 4), torch.ones(4, dtype=torch.bool), st.W_dec, st.b_dec, 2, 0.5)
    assert float(a2) == pytest.approx(0.5 * float(a1))
    h = torch.tensor([[0.0, 0.0, 5.0, 0.0]])
    dead = torch.tensor([False, True, True, False])
    a3, _, _ = orc.auxk(h, torch.zeros(1, 4), dead, st.W_dec, st.b_dec, 8, 1.0)
    assert float(a3) == pytest.approx(25 / 4)
    a0, fa, _ = orc.auxk(h, torch.zeros(1, 4), torch.zeros(4, dtype=torch.bool), st.W_dec, st.b_dec, 2, 1.0)
    assert float(a0) == 0.0 and fa is None


def test_auxk_selects_on_preacts_among_dead_only():
    """test_auxk.py:84-113,198-238: only the top dead latent gets gradient; live latents none."""
    st = _identity_state(4)
    h = torch.tensor([[1.0, 2.0, 3.0, 0.5]])
    dead = torch.tensor([False, True, True, False])
    _, (fa, ma), _ = orc.auxk(h, torch.zeros(1, 4), dead, st.W_dec, st.b_dec, 1, 1.0)
    assert ma.tolist() == [[0.0, 0.0, 1.0, 0.0]]
    assert fa.tolist() == [[0.0, 0.0, 3.0, 0.0]]


def test_topk_known_answers():
    """/root/reference/tests/test_nn_activations.py:29-38 (basic), :41-52 (ties keep exactly k),
    :67-77 (negatives can be selected: no ReLU)."""
    h = torch.tensor([[1.0, 5.0, 3.0, 2.0], [4.0, 1.0, 6.0, 2.0]])
    f, mask = orc.topk_activation(h, 2)
    assert f.tolist() == [[0.0, 5.0, 3.0, 0.0], [4.0, 0.0, 6.0, 0.0]]
    f, mask = orc.topk_activation(torch.full((3, 7), 2.0), 3)
    assert (mask.sum(dim=1) == 3).all()
    f, _ = orc.topk_activation(torch.tensor([[-5.0, -1.0, -3.0, -2.0]]), 2)
    assert f.tolist() == [[0.0, -1.0, 0.0, -2.0]]
    f, _ = orc.topk_activation(torch.tensor([[1.0, 2.0]]), 5)  # k > d_sae clamps (modeling.py:175)
    assert f.tolist() == [[1.0, 2.0]]


def test_batch_topk_known_answers():
    """/root/reference/tests/test_nn_activations.py:171-233: basic, k exceeds the element count, single row, uneven
    distribution across rows, ties (exactly k * B survive); :220-224 eval = JumpReLU with the stored threshold."""
    bt = lambda h, k: orc.batch_topk_activation(torch.tensor(h), k, True, 0.0, 0.1)  # noqa: E731
    assert bt([[5.0, 1.0, 3.0], [2.0, 4.0, 1.0]], 2)[0].tolist() == [[5.0, 0.0, 3.0], [2.0, 4.0, 0.0]]
    assert bt([[5.0, 1.0, 3.0], [2.0, 4.0, 1.0]], 8)[0].tolist() == [[5.0, 1.0, 3.0], [2.0, 4.0, 1.0]]
    assert bt([[1.0, 2.0, 3.0, 4.0]], 2)[0].tolist() == [[0.0, 0.0, 3.0, 4.0]]
    assert bt([[10.0, 20.0, 30.0], [1.0, 2.0, 3.0]], 2)[0].tolist() == [[10.0, 20.0, 30.0], [0.0, 0.0, 3.0]]
    f, mask, thr = bt([[2.0, 2.0, 2.0], [2.0, 2.0, 2.0]], 2)
    assert int(mask.sum()) == 4 and set(f[f != 0].tolist()) == {2.0}
    assert thr == pytest.approx(0.1 * 2.0)  # threshold <- 0.9 * 0 + 0.1 * min positive survivor
    h = torch.tensor([[0.5, -1.0, 2.0], [0.0625, 0.375, -0.25]])
    assert orc.batch_topk_activation(h, 2, False, 0.25, 0.1)[0].tolist() == [[0.5, 0.0, 2.0], [0.0, 0.375, 0.0]]
    assert orc.batch_topk_activation(h, 2, False, 0.0, 0.1)[0].tolist() == [[0.5, 0.0, 2.0], [0.0625, 0.375, 0.0]]


def test_mean_squared_err_known_answers():
    """/root/reference/tests/test_nn_objectives.py:13-52: equals mean((x_hat-x)^2); overflow-safe for 1e20."""
    x = torch.tensor([[1.0, 2.0], [3.0, 4.0]])
    xh = torch.tensor([[1.5, 2.0], [2.0, 6.0]])
    assert float(orc.mean_squared_err(xh, x)) == pytest.approx((0.25 + 0 + 1 + 4) / 4)
    big = torch.tensor([[1e20, -1e20]])
    assert math.isfinite(float(orc.mean_squared_err(big * 1.0001, big)))


def test_dead_counter_semantics():
    """test_auxk.py:310-353 / objectives.py:115-120: counter += B, reset where active, dead at >= thr."""
    f = torch.tensor([[1.0, 0.0, 0.0], [0.0, 0.0, 0.0]])
    toks, dead = orc.dead_tracker_update(None, f, 4)
    assert toks.tolist() == [0, 2, 2] and dead.tolist() == [False, False, False]
    toks, dead = orc.dead_tracker_update(toks, f, 4)
    assert toks.tolist() == [0, 4, 4] and dead.tolist() == [False, True, True]
    f2 = torch.tensor([[0.0, -0.5, 0.0], [0.0, 0.0, 0.0]])  # |f| > 0 counts (negatives revive)
    toks, dead = orc.dead_tracker_update(toks, f2, 4)
    assert toks.tolist() == [2, 0, 6] and dead.tolist() == [False, False, True]


def test_remove_parallel_grads_orthogonal():
    """/root/reference/tests/test_nn_modeling.py:323-338: after projection <g_j, w_j> = 0."""
    g = torch.randn(16, 8, generator=torch.Generator().manual_seed(0))
    w = torch.randn(16, 8, generator=torch.Generator().manual_seed(1))
    w[3] = 0  # zero-norm row untouched
    gp = orc.remove_parallel_grads(g, w)
    assert (gp * w).sum(dim=1).abs().max() < 1e-5
    assert torch.equal(gp[3], g[3])


def test_warmup_cosine_matches_reference_formula():
    """scheduling.py:58-68."""
    vals = [orc.warmup_cosine(s, 4, 1.0, 12) for s in range(1, 14)]
    assert vals[0] == 0.25 and vals[2] == 0.75 and vals[3] == 1.0
    assert vals[11] == 0.0 and vals[12] == 0.0
    assert vals[7] == pytest.approx((1 + math.cos(math.pi * 4 / 8)) / 2)


def test_log_block_matches_reference_capture():
    """oracle.log_block_metrics against the metric dicts the live reference logged (train.py:380-423), captured with
    their inputs by oracle/gen_golden_log.py."""
    import numpy as np

    from tests.golden_util import GOLDEN

    z = np.load(GOLDEN / "log_block.npz")
    assert int(z["n"]) >= 3
    for i in range(int(z["n"])):
        got = orc.log_block_metrics(*(torch.from_numpy(z[f"{k}_{i}"]) for k in ("x", "x_hat", "f_x", "W_dec")))
        for k, ref in zip([str(k) for k in z["keys"]], z[f"metrics_{i}"]):
            assert got[k] == pytest.approx(float(ref), rel=1e-6, abs=1e-9), (i, k)


def test_dictionary_coherence_blocked_equals_direct():
    g = torch.Generator().manual_seed(3)
    W = torch.randn(300, 24, generator=g) * torch.rand(300, 1, generator=g)
    Wn = W / W.norm(dim=1, keepdim=True)
    direct = (Wn @ Wn.T).abs().triu(1).max()  # train.py:417-418 verbatim
    assert float(orc.dictionary_coherence(W, block=64)) == pytest.approx(float(direct), rel=1e-6)
    assert float(orc.dictionary_coherence(W[:1])) == 0.0


def test_evaluate_matches_reference_capture():
    """oracle.evaluate against the EvalMetrics the live reference returned (train.py:510-618) for the same SAE
    parameters and validation set (oracle/gen_golden_log.py)."""
    import numpy as np

    from tests.golden_util import GOLDEN

    z = np.load(GOLDEN / "evaluate.npz")
    acts = torch.from_numpy(z["acts"])
    D, S = z["param_W_enc"].shape
    cfg = orc.OracleConfig(d_model=D, d_sae=S, top_k=int(z["top_k"]))
    st = orc.OracleState.from_params(*(torch.from_numpy(z[f"param_{k}"]) for k in ("W_enc", "b_enc", "W_dec", "b_dec")))
    bs = int(z["batch_size"])
    got = orc.evaluate(cfg, st, [acts[i:i + bs] for i in range(0, acts.shape[0], bs)])
    for k in ("l0", "l1", "mse", "normalized_mse", "sse_sae", "sse_baseline"):
        assert got[k] == pytest.approx(float(z[k]), rel=1e-5), k  # batch composition differs (random loader order)
    for k in ("n_dead", "n_almost_dead", "n_dense"):
        assert got[k] == int(z[k]), k
    assert torch.equal(got["freqs"], torch.from_numpy(z["freqs"]))
    ref_mean = torch.from_numpy(z["mean_values"])
    fired = got["freqs"] > 0
    assert torch.allclose(got["mean_values"][fired], ref_mean[fired], rtol=1e-4, atol=1e-6)
    assert torch.isnan(got["mean_values"][~fired]).all() and torch.isnan(ref_mean[~fired]).all()  # 0 / 0 as in the reference
