"""GPU parity (run with `-m gpu` on a B200): replay the golden runs recorded from the LIVE reference
(tests/golden/*.npz, written by oracle/gen_golden.py) through the CUDA path via the C ABI, step by step:
losses, n_dead, grad norm, the four clipped gradients, x_hat, final parameters, Adam moments, the dead
tracker state and an eval-mode forward.

Tolerance: the north-star bar is 1e-4 relative (fp32); the CUDA path differs from the reference only in
fp32 summation order, so the assertions use TOL = 2e-5 for vectors and scalars (measured ~1e-7..1e-6)."""
import pytest
import torch

from tests.golden_util import CASES, load_case, prefixes_of, rel_l2, t

pytestmark = pytest.mark.gpu

TOL = 2e-5
# the dense ReLU path computes its five contractions as three-piece bf16 split products (hi + lo + lo2 = 24 bits,
# six tensor-core terms) instead of fp32 FMAs: same tolerance as the sparse path
TOL_DENSE = 2e-5
TOPK_CASES = [c for c in CASES if "relu" not in c]


def _engine_for(meta, cfg, B):
    from saev_b200.engine import Engine, EngineConfig
    from saev_b200.nn import batch_topk_capacity

    act, top_k, batch_k = cfg.activation, cfg.top_k, 0
    if act == "batchtopk":  # sparse path: `top_k` = slots per row, `batch_k` = BatchTopK.top_k
        act, batch_k, top_k = "topk", cfg.top_k, batch_topk_capacity(cfg.top_k, cfg.d_sae, meta.get("n_prefixes", 1))
    return Engine(
        EngineConfig(
            d_model=cfg.d_model, d_sae=cfg.d_sae, top_k=top_k, activation=act, aux=cfg.aux, batch_k=batch_k,
            batch_momentum=cfg.batch_momentum,
            k_aux=cfg.k_aux, aux_alpha=cfg.aux_alpha, l1_coeff=cfg.l1_coeff,
            dead_threshold_tokens=cfg.dead_threshold_tokens, remove_parallel_grads=cfg.remove_parallel_grads,
            normalize_w_dec=cfg.normalize_w_dec, max_batch=B, max_prefixes=max(1, meta.get("n_prefixes", 1)),
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
        eng.set_prefixes(prefixes_of(z, step))
        eng.forward(x, training=True)
        eng.backward(x)
        eng.grad_sumsq()
        ld = eng.loss_dict()
        for key in ("mse", "aux", "sparsity", "l0", "l1", "loss"):
            assert ld[key] == pytest.approx(float(z[f"rec_{key}"][step]), rel=TOL, abs=1e-7), (step, key)
        assert int(ld["n_dead"]) == int(z["rec_n_dead"][step]), step
        if cfg.activation == "batchtopk":
            st = eng.batch_topk_stats()
            assert st["kept"] == cfg.top_k * B and st["truncated_rows"] == 0, (step, st)
            assert float(eng.threshold) == pytest.approx(float(z["rec_threshold"][step]), rel=1e-5), step
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
    eng.set_prefixes(prefixes_of(z, meta["n_steps"]))
    eng.forward(x, training=False)
    ld = eng.loss_dict()
    if prefixes_of(z, 0) is not None:
        assert rel_l2(eng.x_hats(x).cpu(), z["eval_x_hats_all"]) < 10 * TOL  # r_i recovered as a difference of suffix sums
    assert ld["mse"] == pytest.approx(float(z["eval_mse"]), rel=TOL)
    assert ld["l0"] == pytest.approx(float(z["eval_l0"]), rel=TOL)
    assert ld["aux"] == 0.0 and ld["n_dead"] == 0.0
    assert rel_l2(eng.x_hat(x).cpu(), z["eval_x_hat"]) < TOL
    if cfg.activation == "batchtopk":
        assert eng.batch_topk_stats()["truncated_rows_total"] == 0
    if cfg.activation != "relu":
        _assert_screen_clean(eng)


def _identity_engine(S, k, B, dec_scale=1.0):
    """D = S, W_enc = W_dec = I, zero biases: the pre-activations ARE the inputs, so the hand-computed cases of the
    reference's activation tests can be fed through the fused forward."""
    from saev_b200.engine import Engine, EngineConfig

    eng = Engine(EngineConfig(d_model=S, d_sae=S, top_k=S, batch_k=k, activation="topk", aux=False, normalize_w_dec=False,
                              remove_parallel_grads=False, max_batch=B))
    eye = torch.eye(S)
    eng.load_params(eye, torch.zeros(S), dec_scale * eye, torch.zeros(S))
    return eng


def test_batchtopk_known_answers():
    """/root/reference/tests/test_nn_activations.py:171-233 through the CUDA path (columns padded with values below every
    entry of the case): basic, uneven distribution across rows, ties (exactly k * B survive), k exceeding the element
    count, and the gradient mask (:237-252: d f / d h = 1 on the survivors, 0 elsewhere -> gb_enc counts them)."""
    S = 8
    pad = lambda rows, fill: torch.tensor([r + [fill] * (S - len(r)) for r in rows]).cuda()  # noqa: E731
    eng = _identity_engine(S, 2, 2)
    eng.forward(pad([[5.0, 1.0, 3.0], [2.0, 4.0, 1.0]], 0.0), training=True)
    assert eng.dense_f_x(2)[:, :3].tolist() == [[5.0, 0.0, 3.0], [2.0, 4.0, 0.0]]
    assert float(eng.threshold) == pytest.approx(0.1 * 2.0)
    eng.forward(pad([[10.0, 20.0, 30.0], [1.0, 2.0, 3.0]], 0.0), training=True)
    assert eng.dense_f_x(2)[:, :3].tolist() == [[10.0, 20.0, 30.0], [0.0, 0.0, 3.0]]
    x = pad([[2.0, 2.0, 2.0], [2.0, 2.0, 2.0]], 0.0)
    eng.forward(x, training=True)
    f = eng.dense_f_x(2)
    assert int((f != 0).sum()) == 4 and set(f[f != 0].tolist()) == {2.0}
    st = eng.batch_topk_stats()
    assert st["kept"] == 4 and st["ties"] == 6 and st["truncated_rows"] == 0
    # gradient mask (:237-252): with W_dec = 2 I the residual on a survivor is its own value, so
    # d loss / d h[b, j] = 2 / (B D) * 2 * h[b, j] on the survivors and 0 on every other entry
    eng = _identity_engine(S, 2, 2, dec_scale=2.0)
    x = pad([[5.0, 1.0, 3.0], [2.0, 4.0, 1.0]], 0.0)
    eng.forward(x, training=True)
    eng.backward(x)
    assert eng.gb_enc.tolist() == pytest.approx([0.25 * (5 + 2), 0.25 * 4, 0.25 * 3, 0, 0, 0, 0, 0])
    # k exceeds the number of elements: everything survives (:184-193)
    eng = _identity_engine(S, 8, 2)
    x = pad([[5.0, 1.0, 3.0], [2.0, 4.0, 1.0]], 0.5)
    eng.forward(x, training=True)
    assert torch.equal(eng.dense_f_x(2), x)
    # eval: JumpReLU with the stored threshold (:220-224); threshold <= 0 -> ReLU
    eng = _identity_engine(S, 2, 2)
    h = pad([[0.5, -1.0, 2.0], [0.0625, 0.375, -0.25]], -3.0)
    eng.threshold.fill_(0.25)
    eng.forward(h, training=False)
    assert eng.dense_f_x(2)[:, :3].tolist() == [[0.5, 0.0, 2.0], [0.0, 0.375, 0.0]]
    eng.threshold.fill_(0.0)
    eng.forward(h, training=False)
    assert eng.dense_f_x(2)[:, :3].tolist() == [[0.5, 0.0, 2.0], [0.0625, 0.375, 0.0]]


@pytest.mark.parametrize("name", ["c1_topk_auxk_live", "tiny_topk_auxk_clamp"])
def test_two_pass_decode_replays_reference_run(name, monkeypatch):
    """By default d loss / d h is computed inside the weight-gradient kernel; SAEV_B200_FUSE_DH=0 (read at create) brings
    back the decode kernel's second gather pass (also what d_model > 1024 and Matryoshka use).  Same golden run."""
    monkeypatch.setenv("SAEV_B200_FUSE_DH", "0")
    test_cuda_path_replays_reference_run(name)


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


def _assert_screen_clean(eng):
    """Every row the tensor-core screen could not certify was redone by the exact path; the only expected cause is a
    threshold guess that verification rejected (a row or so per step by construction) -- no candidate-list overflow,
    no observed error above the deterministic bound."""
    st = eng.screen_stats()
    assert st["unrepaired"] == 0 and st["bound_violations"] == 0 and st["unsafe_rows"] == st["guess_failed"], st


# ---- mid-size parity against the oracle on seeded inputs (sizes the CPU oracle finishes in seconds) ----
@pytest.mark.parametrize("act,D,S,K,B", [("topk", 256, 4096, 32, 1024), ("topk", 768, 8192, 32, 640), ("relu", 192, 2048, 0, 520),
                                          ("topk", 128, 1000, 16, 300), ("topk", 256, 4096, 128, 520)])
@pytest.mark.parametrize("aux_path", ["tc", "sgemm", "auto"])
def test_midsize_steps_match_oracle(act, D, S, K, B, aux_path, monkeypatch):
    _midsize_run(act, D, S, K, B, aux_path, monkeypatch)


@pytest.mark.parametrize("act,D,S,K,B", [("relu", 192, 2048, 0, 520), ("topk", 256, 4096, 32, 640)])
def test_midsize_matryoshka_steps_match_oracle(act, D, S, K, B, monkeypatch):
    """Matryoshka prefixes with cuts that are not multiples of any tile size (objectives.py:124-138): the dense path
    runs the decoder / dh / W_dec-gradient contractions per prefix block on windows [cut_{c-1}, cut_c) of the operands
    (a 1-column block, a block inside one k-block, blocks straddling tiles); the sparse path emits every prefix from
    one decode."""
    _midsize_run(act, D, S, K, B, "auto", monkeypatch, prefixes=[1, 37, 100, 777, S - 3, S])


def test_dense_features_take_the_block_per_atom_path(monkeypatch):
    """Three atoms with a large encoder bias fire on every row of a 1400-row batch: their lists (> 512 entries) are
    handled by wgrad_heavy_kernel (one block per atom) instead of one warp; same oracle, same tolerances."""
    _midsize_run("topk", 256, 4096, 32, 1400, "auto", monkeypatch, dense_atoms=(5, 2049, 4095))


def test_dense_features_with_two_pass_decode(monkeypatch):
    monkeypatch.setenv("SAEV_B200_FUSE_DH", "0")
    _midsize_run("topk", 256, 4096, 32, 1400, "auto", monkeypatch, dense_atoms=(5, 2049, 4095))


def _midsize_run(act, D, S, K, B, aux_path, monkeypatch, dense_atoms=(), k_aux=64, n_steps=4, rank=24,
                 ref_norm_rel=1e-3, tie_tolerant=False, prefixes=None):
    """Four steps (AuxK live from step 2, L1 on for ReLU, ragged batch sizes, d_sae not a multiple of the tile).

    The library has two AuxK implementations (tensor-core split contractions / fp32 tiles) and picks one per step from
    a lagged dead-latent count; SAEV_B200_AUX (read at create) pins either so that both are held to the oracle."""
    from oracle import sae_oracle as orc
    from saev_b200.engine import Engine, EngineConfig

    if aux_path == "auto":
        monkeypatch.delenv("SAEV_B200_AUX", raising=False)
    else:
        monkeypatch.setenv("SAEV_B200_AUX", aux_path)

    g = torch.Generator().manual_seed(D + S)
    W_enc, b_enc, W_dec, b_dec = orc.init_params(D, S, g)
    b_enc = 0.02 * torch.randn(S, generator=g)
    b_dec = 0.02 * torch.randn(D, generator=g)
    for j in dense_atoms:
        b_enc[j] = 4.0
    l1 = 4e-4 if act == "relu" else 0.0
    ocfg = orc.OracleConfig(d_model=D, d_sae=S, activation=act, top_k=max(K, 1), l1_coeff=l1, aux=True, k_aux=k_aux,
                            dead_threshold_tokens=2 * B, lr=1e-3, n_lr_warmup=2, n_steps=10)
    st = orc.OracleState.from_params(W_enc, b_enc, W_dec, b_dec)
    eng = Engine(EngineConfig(d_model=D, d_sae=S, top_k=max(K, 1), activation=act, aux=True, k_aux=k_aux, l1_coeff=l1,
                              dead_threshold_tokens=2 * B, max_batch=B, max_prefixes=len(prefixes) if prefixes else 1))
    eng.load_params(W_enc, b_enc, W_dec, b_dec)
    eng.set_prefixes(prefixes)
    basis = torch.randn(max(rank, 1), D, generator=g)
    tol = TOL_DENSE if act == "relu" else TOL
    # ReLU case: the inputs carry a large common offset (x - 0.5), i.e. every contraction cancels heavily (the fp32
    # oracle itself is only good to 4e-6 here; the two-piece split, SAEV_B200_DENSE_TERMS=3, is at 1.7e-4)
    tol_g = TOL
    lr = 0.0
    for step in range(n_steps):  # dead latents appear at step 2; in "auto" step 3 is the first on the tensor-core path
        if rank > 0:
            x = torch.randn(B, rank, generator=g) @ basis / 4 + 0.1 * torch.randn(B, D, generator=g)
        else:  # full-rank Gaussian rows: nearly every atom fires, only a handful die
            x = torch.randn(B, D, generator=g)
        if act == "relu":
            x = x - 0.5  # push many pre-activations negative so that latents die
        xd = x.cuda()
        eng.normalize_w_dec()
        eng.forward(xd, training=True)
        eng.backward(xd)
        eng.grad_sumsq()
        forced, forced_aux = None, None
        if act == "topk" and tie_tolerant:
            # the oracle adopts the kernels' selections after checking that they differ from torch.topk's only at
            # fp32-level ties (oracle.adopt_selection)
            forced = eng.topk_idx[:B].cpu().long()
            forced_aux = lambda n_dead: eng.aux_selection(B).cpu()  # noqa: E731
        ref = orc.train_step(ocfg, st, x, topk_idx=forced, aux_idx=forced_aux, prefixes=prefixes)
        ld = eng.loss_dict()
        if prefixes:
            assert rel_l2(eng.x_hats(xd).cpu(), ref["out"].x_hats) < 10 * tol, step
        for key in ("mse", "aux", "sparsity", "l1", "loss"):
            assert ld[key] == pytest.approx(ref[key], rel=tol, abs=1e-7), (step, key)
        assert abs(ld["l0"] - ref["l0"]) <= (1e-3 if act == "relu" else 0) * max(ref["l0"], 1) + 1e-6, step
        assert abs(int(ld["n_dead"]) - ref["n_dead"]) <= (2 if act == "relu" else 0), step
        # The reference's clip_grad_norm_ accumulates the norm of ~10^7 elements in fp32 on the CPU, which is itself
        # off by up to ~4e-4 at these sizes (measured against an fp64 run of the oracle); the kernel accumulates in
        # fp64.  So: un-clip the oracle's gradients, compare those tightly, compare our norm tightly with the fp64
        # norm of the oracle's gradients, and the two norms with each other only to the reference's own accuracy.
        gn = float(eng.sumsq.sqrt())
        coef_ref = min(1.0, ocfg.grad_clip / (ref["grad_norm"] + 1e-6))
        ref_grads = {k: v / coef_ref for k, v in ref["grads"].items()}
        gn_ref64 = float(sum(v.double().pow(2).sum() for v in ref_grads.values()).sqrt())
        assert gn == pytest.approx(gn_ref64, rel=tol_g), step
        assert gn == pytest.approx(ref["grad_norm"], rel=ref_norm_rel), step
        ours = dict(W_enc=eng.gW_enc_t.t(), b_enc=eng.gb_enc, W_dec=eng.gW_dec, b_dec=eng.gb_dec)
        errs = {k: rel_l2(gr.cpu(), ref_grads[k]) for k, gr in ours.items()}
        assert max(errs.values()) < tol_g, (step, errs)
        eng.adam_step(lr, max_norm=ocfg.grad_clip)
        lr = st.lr
    for name, p in (("W_enc", eng.W_enc_t.t()), ("b_enc", eng.b_enc), ("W_dec", eng.W_dec), ("b_dec", eng.b_dec)):
        assert rel_l2(p.cpu(), getattr(st, name)) < tol, name
    if act == "topk":
        _assert_screen_clean(eng)
    return eng


@pytest.mark.parametrize("nterms,tol", [(1, 6e-3), (3, 4e-5), (6, 8e-6)])
def test_split_contraction_accuracy(nterms, tol):
    """The tcgen05 contraction alone against fp64, on operands with a large common offset (heavy cancellation):
    error relative to the result for the 1-, 3- and 6-term bf16 splits."""
    from saev_b200.engine import Engine, EngineConfig

    eng = Engine(EngineConfig(d_model=64, d_sae=256, top_k=8, max_batch=64))
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(300, 520, device="cuda", generator=g) - 0.5
    Bt = torch.randn(777, 520, device="cuda", generator=g) * 0.1 + 0.05
    bias = torch.randn(777, device="cuda", generator=g)
    ref = A.double() @ Bt.double().t() + bias.double()
    out = eng.gemm_nt(A, Bt, bias, nterms)
    assert rel_l2(out.cpu(), ref.cpu()) < tol


@pytest.mark.parametrize("fuse", ["1", "0"])
def test_staged_backward_equals_one_shot_backward(fuse, monkeypatch):
    """saev_b200_backward_stage (stage 0, then the weight-gradient rows in chunks -- what the overlapped data-parallel
    exchange drives) must leave the same gradient bucket as saev_b200_backward, including atoms on the block-per-slice
    path (dense features) and the AuxK rows of dead atoms, which the row chunks must leave alone."""
    from oracle import sae_oracle as orc
    from saev_b200.engine import Engine, EngineConfig

    monkeypatch.setenv("SAEV_B200_FUSE_DH", fuse)
    D, S, K, B = 256, 4096, 32, 1400
    g = torch.Generator().manual_seed(77)
    W_enc, b_enc, W_dec, b_dec = orc.init_params(D, S, g)
    b_enc = 0.02 * torch.randn(S, generator=g)
    for j in (3, 1111, 4090):
        b_enc[j] = 4.0
    basis = torch.randn(24, D, generator=g)
    engs = []
    for _ in range(2):
        e = Engine(EngineConfig(d_model=D, d_sae=S, top_k=K, activation="topk", aux=True, k_aux=64,
                                dead_threshold_tokens=2 * B, max_batch=B))
        e.load_params(W_enc, b_enc, W_dec, b_dec)
        engs.append(e)
    a, b = engs
    chunks = [(0, 1000), (1000, 1008), (1008, 3000), (3000, S)]
    for step in range(3):  # dead latents (AuxK live) from step 2
        x = (torch.randn(B, 24, generator=g) @ basis / 4 + 0.1 * torch.randn(B, D, generator=g)).cuda()
        for e in (a, b):
            e.normalize_w_dec()
            e.forward(x, training=True)
        a.backward(x)
        b.backward_stage(x, 0)
        for r0, r1 in chunks:
            b.backward_stage(x, 1, r0, r1)
        assert float(a.losses[5]) == float(b.losses[5])
        for name in ("gW_enc_t", "gb_enc", "gW_dec", "gb_dec"):
            ga, gb = getattr(a, name), getattr(b, name)
            assert rel_l2(gb.cpu(), ga.cpu()) < 1e-6, (step, name)  # summation order inside an atom's list may differ
        for e in (a, b):
            e.grad_sumsq()
            e.adam_step(1e-3, max_norm=1.0)
    assert int(a.losses[5]) > 0, "the case must exercise AuxK"


def test_overlapped_decoder_update_is_the_same_step(monkeypatch):
    """Engine.train_step(overlap_decoder_update=True) defers the decoder half of Adam to a side stream beside the next
    forward's screen: same kernels on the same data, so parameters, moments and losses must agree with the in-order
    step to summation-order noise (atoms with long lists are accumulated with floating-point REDs), including across a
    flush() in the middle and with AuxK live.  (One AuxK implementation is pinned: the automatic choice between the
    two depends on when a lagged device-to-host read lands, i.e. on host timing.)"""
    from saev_b200.engine import Engine, EngineConfig

    monkeypatch.setenv("SAEV_B200_AUX", "sgemm")
    D, S, K, B = 256, 4096, 16, 1024
    g = torch.Generator().manual_seed(5)
    basis = torch.randn(12, D, generator=g)
    xs = [(torch.randn(B, 12, generator=g) @ basis / 3 + 0.1 * torch.randn(B, D, generator=g)).cuda() for _ in range(7)]
    engs = []
    for _ in range(2):
        e = Engine(EngineConfig(d_model=D, d_sae=S, top_k=K, aux=True, k_aux=64, dead_threshold_tokens=2 * B, max_batch=B))
        e.init_params(seed=3)
        engs.append(e)
    a, b = engs
    losses = []
    for i, x in enumerate(xs):
        lr = 1e-3 * min(i, 3) / 3
        a.train_step(x, lr, fused_renorm=True, pre_normalized=i > 0)
        b.train_step(x, lr, fused_renorm=True, pre_normalized=i > 0, overlap_decoder_update=True)
        if i == 3:
            b.flush()  # e.g. a checkpoint in the middle of training
            assert rel_l2(b.W_dec.cpu(), a.W_dec.cpu()) < 1e-6
        if i % 2 == 0:  # (loss_dict() flushes: keep steps whose decoder update really runs beside the next screen)
            losses.append((a.loss_dict(), b.loss_dict()))
    for la, lb in losses:
        for k in la:
            assert la[k] == pytest.approx(lb[k], rel=1e-6, abs=1e-9), k
    assert int(a.losses[5]) > 0, "the case must exercise AuxK"
    b.flush()
    for name in ("params", "m", "v"):
        assert rel_l2(getattr(b, name).cpu(), getattr(a, name).cpu()) < 1e-6, name
    assert torch.equal(a.toks_since_active, b.toks_since_active)
