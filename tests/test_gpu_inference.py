"""GPU test (run with `-m gpu`) of the forward-only inference pass (saev_b200/inference.py) against the arithmetic
of saev.framework.inference.worker_fn (/root/reference/src/saev/framework/inference.py:171-285) restated with the
oracle's dense forward on the CPU: metrics, per-atom statistics, the `distributions` block and the CSR matrix, with
and without ignored token labels, over the golden shard directory the reference's ShardWriter wrote."""
import json
import pathlib

import numpy as np
import pytest
import scipy.sparse
import torch

from oracle import sae_oracle as orc
from saev_b200 import inference, nn

pytestmark = pytest.mark.gpu
GOLDEN = pathlib.Path(__file__).resolve().parent / "golden"
CENSUS = json.loads((GOLDEN / "shards_census.json").read_text())
SHARDS = GOLDEN / CENSUS["dir"]


def _reference_pass(st, cfg, batches, T, n_samples, n_dists, ignore):
    """inference.py:196-285 with dense tensors (f_x from the oracle's eval forward)."""
    S = st.W_dec.shape[0]
    sparsity, mean_values = torch.zeros(S), torch.zeros(S)
    dist = torch.zeros(n_samples, n_dists)
    blocks = []
    sse = torch.zeros((), dtype=torch.float64)
    sum_sq = torch.zeros((), dtype=torch.float64)
    sum_vec = torch.zeros(st.W_dec.shape[1], dtype=torch.float64)
    n_tokens = 0
    for b in batches:
        x = b["act"]
        out = orc.eval_forward(cfg, st, x)
        f_x = out.f.clone()
        mask = torch.ones(len(x), dtype=torch.bool)
        if ignore:
            mask = torch.isin(b["token_labels"].long(), torch.tensor(ignore), invert=True)
        n_tokens += int(mask.sum())
        xm = x[mask].double()
        diff = xm - out.x_hat[mask].double()
        sse += (diff * diff).sum()
        sum_sq += (xm * xm).sum()
        sum_vec += xm.sum(0)
        dist[b["example_idx"][mask].long(), :] = f_x[mask, :n_dists]
        mean_values += f_x[mask].sum(0)
        sparsity += (f_x[mask] > 0).sum(0)
        f_x[~mask, :] = 0.0
        blocks.append(scipy.sparse.csr_array(f_x.numpy()))
    sse_baseline = float(sum_sq - torch.dot(sum_vec, sum_vec) / n_tokens)
    return dict(sse_recon=float(sse), sse_baseline=sse_baseline, n_tokens=n_tokens, mean_values=mean_values / sparsity,
                sparsity=sparsity / n_samples, dist=dist, csr=scipy.sparse.vstack(blocks, format="csr"))


@pytest.mark.parametrize("ignore", [[], [0]], ids=["all-tokens", "ignore-label-0"])
def test_inference_pass_matches_the_reference_arithmetic(ignore):
    md = json.loads((SHARDS / "metadata.json").read_text())
    T, D, n_examples = md["content_tokens_per_example"], md["d_model"], md["n_examples"]
    S, K, n_dists = 64, 4, 9
    torch.manual_seed(4)
    cfg = nn.SparseAutoencoderConfig(d_model=D, d_sae=S, activation=nn.TopK(top_k=K), reinit_blend=0.0)
    sae = nn.SparseAutoencoder(cfg)
    with torch.no_grad():
        sae.b_enc.copy_(0.3 * torch.randn(S))
        sae.b_dec.copy_(0.1 * torch.randn(D))
    st = orc.OracleState.from_params(sae.W_enc.detach().clone(), sae.b_enc.detach().clone(), sae.W_dec.detach().clone(),
                                     sae.b_dec.detach().clone())
    ocfg = orc.OracleConfig(d_model=D, d_sae=S, top_k=K, aux=False)
    batches = list(inference.ordered_batches(SHARDS, 3, 3 * T, labels=True))  # 3 examples per batch, last one short
    assert sum(len(b["act"]) for b in batches) == n_examples * T
    n_samples = n_examples * T
    ref = _reference_pass(st, ocfg, batches, T, n_samples, n_dists, ignore)
    res = inference.run(sae.cuda(), batches, content_tokens_per_example=T, n_samples=n_samples, n_dists=n_dists,
                        ignore_labels=ignore)
    m = res.metrics
    assert m["n_tokens"] == ref["n_tokens"] and m["d_model"] == D and m["n_elements"] == ref["n_tokens"] * D
    assert m["sse_recon"] == pytest.approx(ref["sse_recon"], rel=2e-5)
    assert m["sse_baseline"] == pytest.approx(ref["sse_baseline"], rel=1e-9)
    assert m["normalized_mse"] == pytest.approx(ref["sse_recon"] / ref["sse_baseline"], rel=2e-5)
    assert m["mse_per_dim"] == pytest.approx(ref["sse_recon"] / (ref["n_tokens"] * D), rel=2e-5)
    assert torch.allclose(res.sparsity, ref["sparsity"])
    fired = ref["sparsity"] > 0
    assert torch.allclose(res.mean_values[fired], ref["mean_values"][fired], rtol=2e-5, atol=1e-6)
    assert bool(torch.isnan(res.mean_values[~fired]).all())  # 0 / 0, as in the reference
    assert torch.allclose(res.distributions, ref["dist"], rtol=2e-5, atol=1e-6)
    a, b = res.token_acts, ref["csr"]
    assert a.shape == b.shape == (n_samples, S)
    assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
    assert np.allclose(a.data, b.data, rtol=2e-5, atol=1e-6)
    # metrics-only mode
    res2 = inference.run(sae, batches, content_tokens_per_example=T, n_samples=n_samples, ignore_labels=ignore, save=False)
    assert res2.token_acts is None and res2.metrics == pytest.approx(m, rel=1e-12)  # (fp64 atomics: order of addition)
