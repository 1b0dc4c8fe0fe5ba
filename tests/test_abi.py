"""CPU checks of the drop-in boundary: the built library loads without a GPU, exports every symbol that
include/saev_b200.h declares, the ctypes prototypes cover exactly those symbols, and compute entry points
fail loudly (no silent CPU fallback) when there is no CUDA device."""
import ctypes as C
import pathlib
import re
import subprocess

import pytest
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "saev_b200.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(saev_b200_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from saev_b200 import _lib

    if not _lib.LIB_PATH.exists():
        subprocess.run([str(ROOT / "build.sh")], check=True)
    return _lib.load()


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("saev_b200_create", "saev_b200_forward", "saev_b200_backward", "saev_b200_adam_step",
                 "saev_b200_grad_sumsq", "saev_b200_normalize_w_dec", "saev_b200_ring_submit"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/saev_b200.h but not exported"


def test_ctypes_prototypes_match_header(lib):
    from saev_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert lib.saev_b200_abi_version() == _lib.ABI_VERSION


def test_cfg_struct_layout_matches_header():
    from saev_b200 import _lib

    # int32 x6, float x2, int64, int32 x4  -> 8-byte aligned, 56 bytes
    assert C.sizeof(_lib.Cfg) == 56
    assert _lib.Cfg.dead_threshold_tokens.offset == 32


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    from saev_b200 import _lib
    from saev_b200.engine import Engine, EngineConfig

    with pytest.raises(RuntimeError, match="CUDA"):
        Engine(EngineConfig(d_model=128, d_sae=512, top_k=16, max_batch=256))
    cfg = _lib.Cfg(d_model=128, d_sae=512, act_kind=0, top_k=16, aux_kind=0, k_aux=0, aux_alpha=0.0, l1_coeff=0.0,
                   dead_threshold_tokens=1, remove_parallel_grads=1, max_batch=256, aux_cols_cap=0, max_prefixes=0)
    h = C.c_void_p()
    rc = lib.saev_b200_create(C.byref(cfg), C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CUDA device" in lib.saev_b200_last_error(None)


def test_create_rejects_bad_configs(lib):
    from saev_b200 import _lib

    def rc(**kw):
        base = dict(d_model=128, d_sae=512, act_kind=0, top_k=16, aux_kind=0, k_aux=0, aux_alpha=0.0, l1_coeff=0.0,
                    dead_threshold_tokens=1, remove_parallel_grads=1, max_batch=256, aux_cols_cap=0, max_prefixes=0)
        base.update(kw)
        h = C.c_void_p()
        return lib.saev_b200_create(C.byref(_lib.Cfg(**base)), C.byref(h)), lib.saev_b200_last_error(None)

    assert rc(d_model=100)[0] == 2  # not a multiple of 8
    assert rc(top_k=0)[0] == 2
    assert rc(top_k=1024)[0] == 2  # > d_sae
    assert rc(d_sae=510)[0] == 2
    code, msg = rc(top_k=200, d_sae=4096)  # > 128: the screen kernel's row capacity
    assert code == 3 and b"top_k" in msg


def test_product_package_never_imports_the_oracle():
    for py in (ROOT / "saev_b200").rglob("*.py"):
        src = py.read_text()
        assert "import oracle" not in src and "from oracle" not in src, py
