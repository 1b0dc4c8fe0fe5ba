"""Golden vectors for the datapoint initialisation (saev src/saev/framework/train.py:109-189 `make_saes`), produced by
the LIVE reference on fixed seeds.  TEST INFRASTRUCTURE ONLY (runs in the build container; the vectors are committed).

    python oracle/gen_golden_init.py        # rewrites tests/golden/datapoint_init.npz

A fake loader hands the reference CPU batches; the reference draws `randperm(n_samples)`, the kaiming rows and one
`randperm(d_sae)` per SAE from torch's global CPU generator.  The file stores the batches, the seed and the resulting
W_enc / W_dec of two SAEs (blend 0.8 tied, blend 0.5 tied) -- `saev_b200.nn.datapoint_init(noise_device="cpu")` must
reproduce them from the same seed.
"""

import pathlib
import sys

import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import ref_harness  # noqa: E402

ref_harness.import_reference()
import saev.framework.train as train  # noqa: E402
import saev.nn.modeling as M  # noqa: E402
import saev.nn.objectives as O  # noqa: E402


class FakeLoader:
    def __init__(self, batches):
        self.batches = batches
        self.n_samples = sum(len(b) for b in batches)
        self.drop_last = False
        self.batch_size = len(batches[0])

    def __iter__(self):
        for b in self.batches:
            yield {"act": b}


def main():
    D, S, B, n_batches, seed = 32, 192, 80, 4, 11
    g = torch.Generator().manual_seed(5)
    basis = torch.randn(6, D, generator=g)
    batches = [torch.randn(B, 6, generator=g) @ basis + 0.3 * torch.randn(B, D, generator=g) + 0.7 for _ in range(n_batches)]
    cfgs = [
        (M.SparseAutoencoderConfig(d_model=D, d_sae=S, activation=M.TopK(top_k=8), reinit_blend=0.8), O.Matryoshka(n_prefixes=1)),
        (M.SparseAutoencoderConfig(d_model=D, d_sae=S, activation=M.TopK(top_k=8), reinit_blend=0.5), O.Matryoshka(n_prefixes=1)),
    ]
    torch.manual_seed(seed)
    saes, _, _ = train.make_saes(cfgs, FakeLoader(batches))
    out = {"batches": torch.stack(batches).numpy(), "meta_seed": seed, "meta_D": D, "meta_S": S,
           "blends": np.array([c[0].reinit_blend for c in cfgs], dtype=np.float64)}
    for i, sae in enumerate(saes):
        out[f"W_enc_{i}"] = sae.W_enc.detach().numpy()
        out[f"W_dec_{i}"] = sae.W_dec.detach().numpy()
    path = HERE.parent / "tests" / "golden" / "datapoint_init.npz"
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
