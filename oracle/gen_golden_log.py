"""Generate tests/golden/log_block.npz and tests/golden/evaluate.npz: inputs and outputs of the reference's LOG BLOCK
(/root/reference/src/saev/framework/train.py:365-442) and of its `evaluate()` (train.py:510-618), captured from a live
run of the unmodified `saev.framework.train.worker_fn` on CPU.  TEST INFRASTRUCTURE ONLY; runs in the build container.

    python oracle/gen_golden_log.py

The block is inline code of `train()`, not a callable, so it is pinned by observation: `ParallelWandbRun.log` is
wrapped, and at every call made from `train()` the wrapper reads the caller's locals (`acts_BD`, `saes`, `fwds`) and
the metric dicts the reference just computed from them.  `train.evaluate` is wrapped the same way: its `EvalMetrics`
result is stored with the parameters of the SAE it was given and the full validation set (n_val >= n_samples, so the
result does not depend on the loader's random order).  Nothing from this repository is on that path.
"""

import base64
import pathlib
import pickle
import sys
import tempfile

import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / "ref_stubs"))
sys.path.insert(0, "/root/reference/src")

import saev.data  # noqa: E402
import saev.data.datasets  # noqa: E402
import saev.data.shards as shards  # noqa: E402
import saev.framework.train as train  # noqa: E402
import saev.nn  # noqa: E402
import saev.utils.wandb  # noqa: E402
from saev.nn.modeling import TopK  # noqa: E402

GOLDEN = HERE.parent / "tests" / "golden"
KEYS = ("explained_variance", "dead_unit_pct", "dictionary_coherence", "avg_decoder_row_norm", "sse_sae",
        "sse_baseline", "normalized_mse")


def main():
    captured = []
    orig_log = saev.utils.wandb.ParallelWandbRun.log

    def capturing_log(self, metrics, *, step):
        fr = sys._getframe(1)
        loc = fr.f_locals
        if fr.f_code.co_name == "train" and "acts_BD" in loc and "fwds" in loc:
            sae, fwd = loc["saes"][0], loc["fwds"][0]
            captured.append(dict(
                step=step, x=loc["acts_BD"].detach().clone(), W_dec=sae.W_dec.detach().clone(),
                x_hat=fwd.x_hats[:, -1, :].detach().clone(), f_x=fwd.f_x.detach().clone(),
                metrics={k: float(metrics[0][f"metrics/{k}"]) for k in KEYS},
            ))
        return orig_log(self, metrics, step=step)

    saev.utils.wandb.ParallelWandbRun.log = capturing_log
    evals = []
    orig_eval = train.evaluate

    def capturing_eval(cfgs, saes, objectives):
        out = orig_eval(cfgs, saes, objectives)
        evals.append((out[0], {k: v.detach().clone() for k, v in saes[0].state_dict().items()}))
        return out

    train.evaluate = capturing_eval
    with tempfile.TemporaryDirectory() as tmp:
        tmp = pathlib.Path(tmp)
        root = tmp / "saev" / "shards"
        root.mkdir(parents=True)
        n_examples, T, D = 48, 8, 32
        md = shards.Metadata(
            family="fake-clip", ckpt="synthetic", layers=(0,), content_tokens_per_example=T, cls_token=False, d_model=D,
            n_examples=n_examples, max_tokens_per_shard=16 * T,
            data=base64.b64encode(pickle.dumps(saev.data.datasets.FakeImg(n_examples=n_examples))).decode("utf8"),
            dataset=pathlib.Path("fake"),
        )
        md.dump(root)
        gen = torch.Generator().manual_seed(11)
        basis = torch.randn(6, D, generator=gen)
        acts = torch.randn(n_examples, 1, T, 6, generator=gen) @ basis + 0.3 * torch.randn(n_examples, 1, T, D, generator=gen)
        with shards.ShardWriter(root, md) as w:
            w.write_batch(acts, 0)
        d = root / md.hash
        cfg = train.Config(
            n_train=n_examples * T, n_val=100_000, device="cpu", track=False, log_every=2, lr=3e-3, n_lr_warmup=2,
            runs_root=tmp / "saev" / "runs", objective=saev.nn.objectives.Matryoshka(n_prefixes=1),
            train_data=saev.data.ShuffledConfig(shards=d, layer=0, batch_size=48),
            val_data=saev.data.ShuffledConfig(shards=d, layer=0, batch_size=48),
            sae=saev.nn.SparseAutoencoderConfig(d_model=D, d_sae=4 * D, activation=TopK(top_k=4), reinit_blend=0.0),
        )
        (tmp / "saev" / "runs").mkdir(parents=True)
        train.worker_fn([cfg])
    assert len(captured) >= 3, len(captured)
    captured = captured[:4]
    out = {"n": np.int64(len(captured))}
    for i, c in enumerate(captured):
        for k in ("x", "W_dec", "x_hat", "f_x"):
            out[f"{k}_{i}"] = c[k].numpy()
        out[f"metrics_{i}"] = np.array([c["metrics"][k] for k in KEYS], dtype=np.float64)
    out["keys"] = np.array(KEYS)
    np.savez_compressed(GOLDEN / "log_block.npz", **out)
    for c in captured:
        print(c["step"], c["metrics"])
    em, sd = evals[0]
    ev = {f"param_{k}": v.numpy() for k, v in sd.items()}
    ev["acts"] = acts[:, 0].reshape(-1, D).numpy()  # every (example, token) row of layer 0 = the validation set
    for f in ("l0", "l1", "mse", "normalized_mse", "sse_sae", "sse_baseline"):
        ev[f] = np.float64(getattr(em, f))
    for f in ("n_dead", "n_almost_dead", "n_dense"):
        ev[f] = np.int64(getattr(em, f))
    ev["freqs"], ev["mean_values"] = em.freqs.numpy(), em.mean_values.numpy()
    ev["top_k"], ev["batch_size"] = np.int64(4), np.int64(48)
    np.savez_compressed(GOLDEN / "evaluate.npz", **ev)
    print("evaluate:", {f: float(ev[f]) for f in ("l0", "l1", "mse", "normalized_mse", "n_dead", "n_almost_dead", "n_dense")})


if __name__ == "__main__":
    main()
