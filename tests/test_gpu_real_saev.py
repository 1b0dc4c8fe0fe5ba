"""Opt-in end-to-end drop-in test: the UNMODIFIED saev training entry point (`saev.framework.train.worker_fn`,
train.py:204-240 -> train() :243-508 -> evaluate() :510-618) on a B200 with `saev_b200.install()` active.

The reference checkout does not exist on the GPU box, so this only runs when `SAEV_B200_REF_SRC` names a directory
that contains the `saev` package (`scripts/stage_reference.sh` stages it under the git-ignored `baseline/_ref/`);
otherwise it is skipped.  Import stubs for the three packages the image lacks come from `oracle/ref_stubs/`.
"""

import base64
import os
import pathlib
import pickle
import sys
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu

REF = os.environ.get("SAEV_B200_REF_SRC", "")


@pytest.mark.skipif(not (REF and pathlib.Path(REF, "saev").is_dir()), reason="SAEV_B200_REF_SRC not staged")
def test_unmodified_worker_fn_trains_and_evaluates_through_the_kernels():
    stubs = str(pathlib.Path(__file__).resolve().parent.parent / "oracle" / "ref_stubs")
    sys.path[:0] = [stubs, str(pathlib.Path(REF).resolve())]
    try:
        import saev.data
        import saev.data.datasets
        import saev.data.shards as shards
        import saev.framework.train as train
        import saev.nn
        from saev.nn.modeling import TopK

        import saev_b200
        from saev_b200 import _lib

        with tempfile.TemporaryDirectory() as tmp:
            tmp = pathlib.Path(tmp)
            root = tmp / "saev" / "shards"
            root.mkdir(parents=True)
            n_examples, T, D = 256, 16, 64
            md = shards.Metadata(
                family="fake-clip", ckpt="synthetic", layers=(0,), content_tokens_per_example=T, cls_token=False,
                d_model=D, n_examples=n_examples, max_tokens_per_shard=64 * T,
                data=base64.b64encode(pickle.dumps(saev.data.datasets.FakeImg(n_examples=n_examples))).decode("utf8"),
                dataset=pathlib.Path("fake"),
            )
            md.dump(root)
            g = torch.Generator().manual_seed(3)
            basis = torch.randn(12, D, generator=g)
            acts = torch.randn(n_examples, 1, T, 12, generator=g) @ basis / 3 + 0.2 * torch.randn(n_examples, 1, T, D, generator=g)
            with shards.ShardWriter(root, md) as w:
                w.write_batch(acts, 0)
            d = root / md.hash
            (tmp / "saev" / "runs").mkdir(parents=True)
            cfg = train.Config(
                n_train=4 * n_examples * T, n_val=n_examples * T, device="cuda", track=False, log_every=4, lr=2e-3,
                n_lr_warmup=4, runs_root=tmp / "saev" / "runs", objective=saev.nn.objectives.Matryoshka(n_prefixes=1),
                train_data=saev.data.ShuffledConfig(shards=d, layer=0, batch_size=512),
                val_data=saev.data.ShuffledConfig(shards=d, layer=0, batch_size=512),
                sae=saev.nn.SparseAutoencoderConfig(d_model=D, d_sae=8 * D, activation=TopK(top_k=8), reinit_blend=0.0),
            )
            evals = []
            saev_b200.install()
            try:
                bound_eval = train.evaluate

                def recording_eval(cfgs, saes, objectives):
                    out = bound_eval(cfgs, saes, objectives)
                    evals.extend(out)
                    return out

                train.evaluate = recording_eval
                launches0 = _lib.load().saev_b200_launch_count()
                run_ids = train.worker_fn([cfg])
                launches1 = _lib.load().saev_b200_launch_count()
            finally:
                saev_b200.uninstall()
            assert len(run_ids) == 1 and launches1 > launches0, "the step must have run inside libsaev_b200.so"
            ckpt = tmp / "saev" / "runs" / run_ids[0] / "checkpoint" / "sae.pt"
            sae = saev.nn.load(ckpt)  # the reference's own loader rebuilds its own class from our checkpoint
            assert type(sae).__module__.startswith("saev.") and sae.W_dec.shape == (8 * D, D)
            (m,) = evals
            assert isinstance(m, train.EvalMetrics)
            assert 0.0 < m.normalized_mse < 0.9, m.normalized_mse  # it learned something on the planted data
            assert m.l0 == pytest.approx(8.0)
            # the reference's own forward on the trained weights agrees with what our evaluate() measured
            x = acts[:, 0].reshape(-1, D)
            ref_out = sae(x)
            nmse = float(((ref_out.x_hats[:, -1, :] - x) ** 2).sum() / ((x - x.mean(0)) ** 2).sum())
            assert nmse == pytest.approx(m.normalized_mse, rel=1e-3)
    finally:
        del sys.path[:2]
