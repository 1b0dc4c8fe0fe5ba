"""ctypes binding of libsaev_b200.so (the C ABI declared in include/saev_b200.h).

There is deliberately no CPU or PyTorch fallback: if the shared library is missing or a CUDA device is
absent, the functions here raise.  `load()` only dlopens the library (works without a GPU, used by the
CPU test-suite to check the exported symbols); every compute entry needs a B200.
"""

from __future__ import annotations

import ctypes as C
import pathlib

LIB_PATH = pathlib.Path(__file__).resolve().parent / "lib" / "libsaev_b200.so"

ABI_VERSION = 11

ACT_TOPK, ACT_RELU = 0, 1
AUX_NONE, AUX_AUXK = 0, 1
PHASE_A, PHASE_B, PHASE_ALL = 1, 2, 3
PHASE_A_SCREEN, PHASE_A_REST, PHASE_A_RESCORE, PHASE_A_DECODE = 4, 8, 16, 32
ADAM_ENCODER, ADAM_DECODER, ADAM_ALL, ADAM_ROWS_ONLY, ADAM_KEEP_MAXIMA = 1, 2, 3, 4, 8
STAGES = ("prep", "encode_gemm", "rescore", "decode", "loss", "csc", "wgrad", "bias_aux", "sumsq", "adam")


class Cfg(C.Structure):
    """Mirror of `saev_b200_cfg` (include/saev_b200.h)."""

    _fields_ = [
        ("d_model", C.c_int32),
        ("d_sae", C.c_int32),
        ("act_kind", C.c_int32),
        ("top_k", C.c_int32),
        ("aux_kind", C.c_int32),
        ("k_aux", C.c_int32),
        ("aux_alpha", C.c_float),
        ("l1_coeff", C.c_float),
        ("dead_threshold_tokens", C.c_int64),
        ("remove_parallel_grads", C.c_int32),
        ("max_batch", C.c_int32),
        ("aux_cols_cap", C.c_int32),
        ("max_prefixes", C.c_int32),
    ]


class LoaderCfg(C.Structure):
    """Mirror of `saev_b200_loader_cfg` (include/saev_b200.h)."""

    _fields_ = [
        ("shards_dir", C.c_char_p),
        ("examples_per_shard", C.c_int32),
        ("n_layers", C.c_int32),
        ("tokens_per_example", C.c_int32),
        ("d_model", C.c_int32),
        ("layer_index", C.c_int32),
        ("cls_token", C.c_int32),
        ("content_tokens", C.c_int32),
        ("shard_order", C.POINTER(C.c_int32)),
        ("shard_examples", C.POINTER(C.c_int32)),
        ("n_order", C.c_int32),
        ("batch_size", C.c_int32),
        ("pool_batches", C.c_int32),
        ("n_threads", C.c_int32),
        ("n_out_slots", C.c_int32),
        ("chunk_examples", C.c_int32),
        ("min_buffer_fill", C.c_float),
        ("reserved", C.c_int32),
        ("n_rows_limit", C.c_int64),
        ("seed", C.c_uint64),
        ("labels", C.c_void_p),
        ("ignore_lut", C.c_void_p),
    ]


_p = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_f = C.c_float

# name -> (restype, argtypes); must list every symbol include/saev_b200.h declares
SIGNATURES = {
    "saev_b200_abi_version": (C.c_int, []),
    "saev_b200_launch_count": (C.c_uint64, []),
    "saev_b200_last_error": (C.c_char_p, [_p]),
    "saev_b200_create": (C.c_int, [C.POINTER(Cfg), C.POINTER(_p)]),
    "saev_b200_destroy": (C.c_int, [_p]),
    "saev_b200_workspace_bytes": (C.c_size_t, [_p]),
    "saev_b200_sync_weights": (C.c_int, [_p, _p, _p, _p, _p]),
    "saev_b200_datapoint_init": (C.c_int, [_p, _p, _i64, _p, _p, _p, _f, _i32, _i32, _p, _p, _p, _p, _p, _p]),
    "saev_b200_normalize_w_dec": (C.c_int, [_p, _p, _p]),
    "saev_b200_forward": (
        C.c_int,
        [_p, C.c_int, _p, _i32, _i64, _p, _p, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _p],
    ),
    "saev_b200_batch_topk": (C.c_int, [_p, _i32, _i32, _i32, _p, _f, _p, _p, _p, _p, _p]),
    "saev_b200_active_flags": (_p, [_p, _p]),
    "saev_b200_unsafe_rows": (_p, [_p, _p]),
    "saev_b200_backward": (C.c_int, [_p, _p, _i32, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "saev_b200_backward_stage": (
        C.c_int, [_p, _i32, _i32, _i32, _p, _i32, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "saev_b200_grad_sumsq": (C.c_int, [_p, _p, _i64, _p, _p, _p]),
    "saev_b200_grad_sumsq_local": (C.c_int, [_p, _p, _p, _p, _p]),
    "saev_b200_set_optimizer_shard": (C.c_int, [_p, _i32, _i32]),
    "saev_b200_set_reserved_sms": (C.c_int, [_p, _i32]),
    "saev_b200_grad_sumsq_ranges": (C.c_int, [_p, _p, _i32, C.POINTER(_i64), C.POINTER(_i64), _p, _p, _p]),
    "saev_b200_shadow_weights": (_p, [_p, _p]),
    "saev_b200_wnorm_scalar": (_p, [_p, _p]),
    "saev_b200_wnorm_rows": (_p, [_p, _p]),
    "saev_b200_adam_step": (
        C.c_int,
        [_p, _p, _p, _p, _p, _p, _p, _p, _f, _f, _f, _f, _i64, _f, _f, _p, _i32, _p, _i32, _p, _p],
    ),
    "saev_b200_densify": (C.c_int, [_p, _p, _p, _i32, _p, _p]),
    "saev_b200_dense_f": (C.c_int, [_p, _p, _p, _i32, _p, _p, _p]),
    "saev_b200_x_hat": (C.c_int, [_p, _p, _p, _i32, _p, _p]),
    "saev_b200_set_prefixes": (C.c_int, [_p, C.POINTER(_i32), _i32]),
    "saev_b200_x_hats": (C.c_int, [_p, _p, _p, _i32, _p, _p, _p]),
    "saev_b200_coherence_scratch_bytes": (C.c_size_t, [_p]),
    "saev_b200_dictionary_coherence": (C.c_int, [_p, _p, _p, C.c_size_t, _p, _p]),
    "saev_b200_log_scratch_bytes": (C.c_size_t, [_p]),
    "saev_b200_log_metrics": (C.c_int, [_p, _p, _p, _i32, _p, _p, _p, C.c_size_t, _p, _p]),
    "saev_b200_eval_accumulate": (C.c_int, [_p, _p, _p, _i32, _p, _p, _p, _p, _p, _p, _p, _p]),
    "saev_b200_aux_selection": (C.c_int, [_p, _p, C.POINTER(_p), C.POINTER(_i64), C.POINTER(_p), C.POINTER(_p)]),
    "saev_b200_gemm_nt": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p]),
    "saev_b200_profile_enable": (C.c_int, [_p, _i32]),
    "saev_b200_profile_read": (C.c_int, [_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "saev_b200_ring_create": (C.c_int, [_i32, C.c_size_t, C.POINTER(_p)]),
    "saev_b200_ring_destroy": (C.c_int, [_p]),
    "saev_b200_ring_host_ptr": (_p, [_p, _i32]),
    "saev_b200_ring_submit": (C.c_int, [_p, _i32, _p, C.c_size_t]),
    "saev_b200_ring_wait": (C.c_int, [_p, _i32, _p]),
    "saev_b200_ring_host_sync": (C.c_int, [_p, _i32]),
    "saev_b200_loader_create": (C.c_int, [C.POINTER(LoaderCfg), C.POINTER(_p)]),
    "saev_b200_loader_destroy": (C.c_int, [_p]),
    "saev_b200_loader_start_epoch": (C.c_int, [_p, _i64, C.c_uint64, _p]),
    "saev_b200_loader_next": (
        C.c_int, [_p, _p, C.c_double, C.POINTER(_p), C.POINTER(_p), C.POINTER(_p), C.POINTER(_i32)]),
    "saev_b200_loader_stats": (
        C.c_int, [_p, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "saev_b200_loader_zero_copy": (C.c_int, [_p]),
    "saev_b200_loader_stop": (C.c_int, [_p]),
    "saev_b200_loader_last_error": (C.c_char_p, [_p]),
    "saev_b200_loader_read_chunk": (_i64, [C.POINTER(LoaderCfg), _i32, _i32, _i32, _p, _p]),
    "saev_b200_loader_plan_draw": (C.c_int, [C.c_uint64, _i64, _i32, _p, _p, _p, C.POINTER(_i32)]),
}

_lib = None


class LibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen libsaev_b200.so and attach prototypes.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise LibraryError(
            f"{LIB_PATH} is missing: build it with ./build.sh (or __graft_entry__.build()). "
            "saev_b200 has no CPU / PyTorch fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.saev_b200_abi_version() != ABI_VERSION:
        raise LibraryError("libsaev_b200.so ABI version mismatch: rebuild with ./build.sh")
    _lib = lib
    return lib


def check(rc: int, handle=None) -> None:
    if rc != 0:
        msg = load().saev_b200_last_error(handle)
        raise RuntimeError(f"libsaev_b200 error {rc}: {msg.decode() if msg else '?'}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
