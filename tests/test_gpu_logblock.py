"""Log-block metrics (saev train.py:365-442) of the CUDA path against the oracle / the reference capture."""

import numpy as np
import pytest
import torch

from tests.golden_util import GOLDEN

pytestmark = pytest.mark.gpu

COH_TOL = 1e-4  # absolute; two-piece bf16 screen (~3e-5 on unit rows) + exact re-evaluation of the winning pairs


def _engine(D, S, K, B):
    from saev_b200.engine import Engine, EngineConfig

    return Engine(EngineConfig(d_model=D, d_sae=S, top_k=K, activation="topk", aux=True, k_aux=64, max_batch=B))


@pytest.mark.parametrize("S,D", [(4096, 256), (1000, 128), (300, 24), (20000, 64)])
def test_dictionary_coherence_matches_oracle(S, D):
    """Random rows of unequal norm (the kernel normalises like train.py:416-417), ragged S, one planted near-duplicate
    pair far from the first rows so that the winner sits in a late tile."""
    from oracle import sae_oracle as orc

    g = torch.Generator().manual_seed(S + D)
    W = torch.randn(S, D, generator=g) * (0.5 + torch.rand(S, 1, generator=g))
    i, j = S // 2 + 3, S - 5
    W[j] = 1.7 * (W[i] + 0.35 * torch.randn(D, generator=g))
    eng = _engine(D, S, 8, 64)
    out = eng.dictionary_coherence(W.cuda()).cpu()
    ref = float(orc.dictionary_coherence(W))
    ref64 = float(orc.dictionary_coherence(W.double()))
    assert abs(float(out[0]) - ref64) < 1e-6, (float(out[0]), ref64)  # the winning pair is re-evaluated exactly
    assert abs(float(out[0]) - ref) < COH_TOL
    assert abs(float(out[1]) - ref64) < COH_TOL  # the tensor-core screen itself
    assert (int(out[2]), int(out[3])) == (i, j)


def test_dictionary_coherence_negative_and_tiny():
    """|.| is taken (anti-parallel rows count, train.py:418); a single row has an empty upper triangle -> 0."""
    from oracle import sae_oracle as orc

    g = torch.Generator().manual_seed(5)
    W = torch.randn(512, 64, generator=g)
    W[400] = -W[17] + 0.05 * torch.randn(64, generator=g)
    eng = _engine(64, 512, 8, 64)
    out = eng.dictionary_coherence(W.cuda()).cpu()
    assert abs(float(out[0]) - float(orc.dictionary_coherence(W.double()))) < 1e-6
    assert (int(out[2]), int(out[3])) == (17, 400)


def test_log_metrics_match_reference_capture():
    """Replays the inputs the live reference had in its log block (oracle/gen_golden_log.py): the SAE weights are
    loaded into the engine, the captured batch is run forward, and the metric kernel is compared with the numbers the
    reference logged.  x_hat comes from our own forward (parity of that is test_gpu_golden's job)."""
    from oracle import sae_oracle as orc
    from saev_b200.engine import Engine

    z = np.load(GOLDEN / "log_block.npz")
    keys = [str(k) for k in z["keys"]]
    for i in range(int(z["n"])):
        x, W_dec, x_hat, f_x = (torch.from_numpy(z[f"{k}_{i}"]) for k in ("x", "W_dec", "x_hat", "f_x"))
        B, D = x.shape
        S = W_dec.shape[0]
        eng = _engine(D, S, 4, B)
        # engine state whose forward reproduces the captured x_hat / f_x: decoder = captured W_dec, and an encoder
        # that makes exactly the captured latents win (h = 1e3 * onehot pattern is not needed -- feed f_x directly)
        eng.load_params(torch.zeros(D, S), torch.zeros(S), W_dec, torch.zeros(D))
        xd = x.cuda()
        eng.resid[:B].copy_((x_hat - x).cuda())
        fired = (f_x.abs() > 0).any(0).to(torch.int32).cuda()
        eng.active_flags().copy_(fired)
        got = eng.log_metrics_dict(xd)
        ref = dict(zip(keys, z[f"metrics_{i}"]))
        for k in ("sse_sae", "sse_baseline", "normalized_mse"):
            assert got[k] == pytest.approx(float(ref[k]), rel=1e-6), (i, k)  # fp64 accumulation of fp32 inputs
        assert got["explained_variance"] == pytest.approx(float(ref["explained_variance"]), abs=2e-6), i  # fp32 .var() in the reference
        assert got["dead_unit_pct"] == pytest.approx(float(ref["dead_unit_pct"]), abs=1e-7), i
        assert got["avg_decoder_row_norm"] == pytest.approx(float(ref["avg_decoder_row_norm"]), rel=1e-6), i
        assert got["dictionary_coherence"] == pytest.approx(float(ref["dictionary_coherence"]), abs=COH_TOL), i
        orc_m = orc.log_block_metrics(x, x_hat, f_x, W_dec)
        assert got["dictionary_coherence"] == pytest.approx(orc_m["dictionary_coherence"], abs=COH_TOL)


def test_log_metrics_after_real_forward():
    """End to end on the CUDA path: forward a batch, then metrics from the engine's own residual / activity flags
    against the oracle evaluated on the oracle's forward of the same state."""
    from oracle import sae_oracle as orc

    D, S, K, B = 128, 2048, 16, 512
    g = torch.Generator().manual_seed(9)
    W_enc, b_enc, W_dec, b_dec = orc.init_params(D, S, g)
    x = torch.randn(B, 12, generator=g) @ torch.randn(12, D, generator=g) / 3 + 0.2 * torch.randn(B, D, generator=g)
    eng = _engine(D, S, K, B)
    eng.load_params(W_enc, b_enc, W_dec, b_dec)
    eng.forward(x.cuda(), training=True)
    got = eng.log_metrics_dict(x.cuda())
    cfg = orc.OracleConfig(d_model=D, d_sae=S, top_k=K, aux=True, k_aux=64)
    out = orc.forward(cfg, orc.OracleState.from_params(W_enc, b_enc, W_dec, b_dec), x, training=True)
    ref = orc.log_block_metrics(x, out.x_hat, out.f, W_dec)
    for k in ("sse_sae", "sse_baseline", "normalized_mse", "avg_decoder_row_norm"):
        assert got[k] == pytest.approx(ref[k], rel=2e-5), k
    assert got["explained_variance"] == pytest.approx(ref["explained_variance"], abs=2e-5)
    assert got["dead_unit_pct"] == pytest.approx(ref["dead_unit_pct"], abs=1e-7)
    assert got["dictionary_coherence"] == pytest.approx(ref["dictionary_coherence"], abs=COH_TOL)


def _check_eval(m, ref, rel):
    for k in ("l0", "l1", "mse", "normalized_mse", "sse_sae", "sse_baseline"):
        assert getattr(m, k) == pytest.approx(float(ref[k]), rel=rel), k
    for k in ("n_dead", "n_almost_dead", "n_dense"):
        assert getattr(m, k) == int(ref[k]), k
    ref_f, ref_mv = torch.as_tensor(ref["freqs"]), torch.as_tensor(ref["mean_values"])
    assert torch.equal(m.freqs, ref_f)
    fired = ref_f > 0
    assert torch.allclose(m.mean_values[fired], ref_mv[fired], rtol=1e-4, atol=1e-6)
    assert torch.isnan(m.mean_values[~fired]).all()


def test_evaluate_matches_reference_capture():
    """saev_b200.evaluate over the validation set the live reference evaluated (oracle/gen_golden_log.py), through the
    nn mirror: same SAE parameters, all rows once -> the EvalMetrics the reference returned (train.py:510-618)."""
    from saev_b200 import evaluate as ev
    from saev_b200 import nn

    z = np.load(GOLDEN / "evaluate.npz")
    D, S = z["param_W_enc"].shape
    sae = nn.SparseAutoencoder(nn.SparseAutoencoderConfig(d_model=D, d_sae=S, activation=nn.TopK(top_k=int(z["top_k"])),
                                                          reinit_blend=0.0))
    sae.load_state_dict({k: torch.from_numpy(z[f"param_{k}"]) for k in ("W_dec", "b_dec", "W_enc", "b_enc")})
    sae = sae.to("cuda")
    obj = nn.get_objective(nn.Matryoshka(n_prefixes=1))
    acts, bs = torch.from_numpy(z["acts"]), int(z["batch_size"])
    batches = [{"act": acts[i:i + bs]} for i in range(0, acts.shape[0], bs)]
    (m,) = ev.evaluate_batches(batches, [sae], [obj])
    _check_eval(m, z, rel=2e-5)


@pytest.mark.parametrize("act", ["topk", "relu"])
def test_eval_accumulate_matches_oracle(act):
    """Engine-level, ragged last batch, both activations (ReLU reads the dense bf16 pieces of the workspace)."""
    from oracle import sae_oracle as orc
    from saev_b200 import evaluate as ev
    from saev_b200.engine import Engine, EngineConfig

    D, S, K, B = 64, 768, 8, 200
    g = torch.Generator().manual_seed(21)
    W_enc, b_enc, W_dec, b_dec = orc.init_params(D, S, g)
    b_enc = 0.05 * torch.randn(S, generator=g) - (0.3 if act == "relu" else 0.0)
    xs = [torch.randn(n, D, generator=g) for n in (B, B, 77)]
    cfg = orc.OracleConfig(d_model=D, d_sae=S, activation=act, top_k=K, l1_coeff=4e-4 if act == "relu" else 0.0)
    ref = orc.evaluate(cfg, orc.OracleState.from_params(W_enc, b_enc, W_dec, b_dec), xs)
    eng = Engine(EngineConfig(d_model=D, d_sae=S, top_k=K, activation=act, l1_coeff=cfg.l1_coeff, aux=False, max_batch=B))
    eng.load_params(W_enc, b_enc, W_dec, b_dec)
    st = eng.new_eval_state()
    for x in xs:
        xd = x.cuda()
        eng.forward(xd, training=False)
        eng.eval_accumulate(xd, st)
    m = ev.finish_metrics(st)
    if act == "relu":  # pre-activations within fp32 rounding of 0 may flip sign: counts can differ by a few
        assert abs(m.n_dead - ref["n_dead"]) <= 1 and abs(m.n_dense - ref["n_dense"]) <= 2
        assert (m.freqs - ref["freqs"]).abs().max() <= 2.5 / 477
        for k in ("l0", "l1", "mse", "normalized_mse", "sse_sae", "sse_baseline"):
            assert getattr(m, k) == pytest.approx(ref[k], rel=1e-3 if k == "l0" else 2e-5), k
    else:
        _check_eval(m, ref, rel=2e-5)


def test_f_x_csr_equals_scipy_on_dense_f_x():
    """Output.f_x_csr() (from the [B, K] lists) against scipy.sparse.csr_array of the dense f_x, as the reference's
    inference dump builds it (inference.py:236), with masked rows (:234)."""
    import scipy.sparse

    from saev_b200 import nn

    D, S, K, B = 64, 1024, 8, 200
    torch.manual_seed(2)
    sae = nn.SparseAutoencoder(nn.SparseAutoencoderConfig(d_model=D, d_sae=S, activation=nn.TopK(top_k=K), reinit_blend=0.0))
    sae = sae.to("cuda").eval()
    x = torch.randn(B, D, device="cuda")
    out = sae(x)
    mask = torch.rand(B) > 0.2
    dense = out.f_x.cpu().clone()
    ref = scipy.sparse.csr_array(dense.numpy())
    got = out.f_x_csr()
    assert got.nnz == ref.nnz == B * K
    assert np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices)
    assert np.array_equal(got.data, ref.data)
    dense[~mask] = 0.0
    ref_m = scipy.sparse.csr_array(dense.numpy())
    got_m = out.f_x_csr(mask)
    assert np.array_equal(got_m.indptr, ref_m.indptr) and np.array_equal(got_m.indices, ref_m.indices)
    assert np.array_equal(got_m.data, ref_m.data)
