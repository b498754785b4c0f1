"""ctypes binding of libcarel_head.so (the C ABI declared in include/carel_head.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, the
caller gets an exception.  Build with ``python -m carel_vae_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("CAREL_HEAD_LIB", _PKG / "libcarel_head.so"))

_lib = None

c_f32p = C.c_void_p
c_ptr = C.c_void_p
i64 = C.c_int64

_SIGNATURES = {
    "carel_version": (C.c_int, []),
    "carel_last_error": (C.c_char_p, []),
    "carel_launch_count": (i64, []),
    "carel_timing_enable": (None, [C.c_int]),
    "carel_set_reserved_sms": (None, [C.c_int]),
    "carel_get_reserved_sms": (C.c_int, []),
    "carel_set_side_branches": (None, [C.c_int]),
    "carel_get_side_branches": (C.c_int, []),
    "carel_timing_report": (C.c_int, [C.c_char_p, C.c_size_t]),
    "carel_gemm_bf16": (C.c_int, [c_ptr, i64, C.c_int, c_ptr, i64, C.c_int, i64, i64, i64, c_ptr, i64,
                                  c_ptr, C.c_int, c_ptr]),
    "carel_gemm_bf16_2sm": (C.c_int, [c_ptr, i64, C.c_int, c_ptr, i64, C.c_int, i64, i64, i64, c_ptr, i64,
                                      c_ptr, C.c_int, c_ptr]),
    "carel_mmd_workspace_bytes": (C.c_size_t, [i64, i64, i64, C.c_int]),
    "carel_mmd_fwd": (C.c_int, [c_ptr, i64, c_ptr, i64, i64, i64, i64, C.POINTER(C.c_float), C.c_int,
                                C.c_float, C.c_int, i64, i64, c_ptr, c_ptr, c_ptr, C.c_size_t, C.c_int, c_ptr]),
    "carel_mmd_finalize": (C.c_int, [c_ptr, i64, i64, c_ptr, c_ptr]),
    "carel_mmd_bwd": (C.c_int, [c_ptr, i64, i64, i64, C.POINTER(C.c_float), C.c_int, C.c_float, i64, i64,
                                c_ptr, i64, c_ptr, i64, c_ptr, C.c_size_t, C.c_int, c_ptr]),
    "carel_heads_workspace_bytes": (C.c_size_t, [i64, i64]),
    "carel_heads_fwd": (C.c_int, [c_ptr, i64, i64, i64, i64, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), c_ptr,
                                  c_ptr, c_ptr, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                                  C.POINTER(C.c_float), c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "carel_heads_bwd": (C.c_int, [c_ptr, c_ptr, i64, i64, i64, i64, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                  c_ptr, c_ptr, c_ptr, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                                  C.POINTER(C.c_float), c_ptr, i64, C.POINTER(C.c_void_p), c_ptr, C.c_size_t, c_ptr]),
    "carel_wsum_fwd": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_float), C.c_int, c_ptr, c_ptr]),
    "carel_wsum_bwd": (C.c_int, [c_ptr, C.POINTER(C.c_float), C.c_int, c_ptr, c_ptr]),
    "carel_club_fwd": (C.c_int, [c_ptr, c_ptr, c_ptr, i64, c_ptr, i64, i64, c_ptr, c_ptr, c_ptr]),
    "carel_club_bwd": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, i64, c_ptr, i64, i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "carel_latent_workspace_bytes": (C.c_size_t, [i64, i64, i64]),
    "carel_latent_fwd": (C.c_int, [c_ptr, i64, i64, i64, i64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), c_ptr,
                                   c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "carel_latent_bwd": (C.c_int, [c_ptr, i64, c_ptr, c_ptr, c_ptr, c_ptr, i64, i64, i64, c_ptr, i64,
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), c_ptr, C.c_size_t, c_ptr]),
    "carel_bow_workspace_bytes": (C.c_size_t, [i64, i64, i64, i64]),
    "carel_bow_pack_ell": (C.c_int, [c_ptr, i64, i64, i64, i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "carel_bow_pack_ell_host": (C.c_int, [c_ptr, i64, i64, i64, i64, c_ptr, c_ptr, c_ptr, C.c_int]),
    "carel_bow_fwd": (C.c_int, [c_ptr, i64, c_ptr, c_ptr, c_ptr, i64, c_ptr, c_ptr, c_ptr, i64, i64, i64, i64,
                                C.c_float, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "carel_bow_bwd": (C.c_int, [c_ptr, c_ptr, i64, i64, i64, i64, i64, C.c_float, c_ptr, i64, c_ptr, c_ptr,
                                c_ptr, C.c_size_t, c_ptr]),
    "carel_bow_bwd_parts": (C.c_int, [c_ptr, c_ptr, i64, i64, i64, i64, i64, C.c_float, c_ptr, i64, c_ptr, c_ptr,
                                      c_ptr, C.c_size_t, C.c_int, c_ptr]),
    "carel_hsic_workspace_bytes": (C.c_size_t, [i64, i64, i64]),
    "carel_hsic_fwd": (C.c_int, [c_ptr, i64, c_ptr, i64, i64, i64, i64, C.c_float, C.c_float, i64, i64, c_ptr, c_ptr,
                                 C.c_size_t, c_ptr]),
    "carel_hsic_ws_layout": (C.c_int, [i64, i64, i64, C.POINTER(i64)]),
    "carel_hsic_finalize": (C.c_int, [i64, i64, i64, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "carel_hsic_bwd": (C.c_int, [c_ptr, i64, i64, i64, C.c_float, C.c_float, i64, i64, c_ptr, i64, c_ptr, i64, c_ptr,
                                 C.c_size_t, c_ptr]),
    "carel_club_mlp_workspace_bytes": (C.c_size_t, [i64, i64]),
    "carel_club_mlp_fwd": (C.c_int, [c_ptr, i64, i64, i64, C.POINTER(C.c_void_p), c_ptr, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "carel_club_mlp_bwd": (C.c_int, [c_ptr, c_ptr, c_ptr, i64, c_ptr, i64, i64, C.POINTER(C.c_void_p), c_ptr, i64,
                                     C.POINTER(C.c_void_p), c_ptr, C.c_size_t, c_ptr]),
    "carel_disc_workspace_bytes": (C.c_size_t, [i64]),
    "carel_disc_fwd": (C.c_int, [c_ptr, i64, i64, i64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), c_ptr, c_ptr, C.c_float,
                                 C.c_float, C.c_float, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "carel_disc_bwd": (C.c_int, [C.POINTER(C.c_void_p), c_ptr, i64, i64, i64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                 c_ptr, c_ptr, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_void_p), c_ptr, C.c_size_t,
                                 c_ptr]),
    "carel_pair_preds_workspace_bytes": (C.c_size_t, [i64, i64, i64]),
    "carel_pair_preds": (C.c_int, [c_ptr, i64, i64, i64, i64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), c_ptr, c_ptr,
                                   c_ptr, c_ptr, c_ptr, c_ptr, C.c_size_t, C.c_int, c_ptr]),
    # ---- fp32 precision tier ----
    "carel_f32_gemm": (C.c_int, [c_ptr, i64, C.c_int, c_ptr, i64, C.c_int, i64, i64, i64, c_ptr, i64, c_ptr, C.c_float,
                                 C.c_int, C.c_int, c_ptr]),
    "carel_latent_f32_workspace_bytes": (C.c_size_t, [i64, i64, i64]),
    "carel_latent_f32_fwd": (C.c_int, [c_ptr, i64, i64, i64, i64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), c_ptr,
                                       c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "carel_latent_f32_bwd": (C.c_int, [c_ptr, i64, c_ptr, c_ptr, i64, C.POINTER(C.c_void_p), c_ptr, c_ptr, c_ptr, i64, i64,
                                       i64, c_ptr, i64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), c_ptr, C.c_size_t,
                                       c_ptr]),
    "carel_mmd_f32_workspace_bytes": (C.c_size_t, [i64, i64, i64]),
    "carel_mmd_f32_fwd": (C.c_int, [c_ptr, i64, c_ptr, i64, i64, i64, i64, C.POINTER(C.c_float), C.c_int, C.c_float, i64,
                                    i64, c_ptr, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "carel_mmd_f32_bwd": (C.c_int, [c_ptr, c_ptr, i64, c_ptr, i64, i64, i64, i64, C.POINTER(C.c_float), C.c_int, C.c_float,
                                    i64, i64, c_ptr, i64, c_ptr, i64, c_ptr, C.c_size_t, c_ptr]),
    "carel_hsic_f32_workspace_bytes": (C.c_size_t, [i64, i64, i64]),
    "carel_hsic_f32_fwd": (C.c_int, [c_ptr, i64, c_ptr, i64, i64, i64, i64, C.c_float, C.c_float, c_ptr, c_ptr, C.c_size_t,
                                     c_ptr]),
    "carel_hsic_f32_bwd": (C.c_int, [c_ptr, c_ptr, i64, c_ptr, i64, i64, i64, i64, C.c_float, C.c_float, c_ptr, i64, c_ptr,
                                     i64, c_ptr, C.c_size_t, c_ptr]),
    "carel_bow_f32_workspace_bytes": (C.c_size_t, [i64, i64, i64, C.c_int]),
    "carel_bow_f32_fwd": (C.c_int, [c_ptr, i64, c_ptr, c_ptr, c_ptr, i64, c_ptr, c_ptr, c_ptr, i64, i64, i64, i64, C.c_float,
                                    c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "carel_bow_f32_bwd": (C.c_int, [c_ptr, c_ptr, i64, c_ptr, c_ptr, i64, i64, i64, i64, C.c_float, c_ptr, i64, c_ptr, c_ptr,
                                    c_ptr, C.c_size_t, c_ptr]),
    "carel_pairwise_workspace_bytes": (C.c_size_t, [i64, i64, i64]),
    "carel_pairwise": (C.c_int, [c_ptr, i64, c_ptr, i64, i64, i64, i64, C.c_int, C.POINTER(C.c_float),
                                 C.c_int, C.c_float, c_ptr, i64, c_ptr, C.c_size_t, c_ptr]),
}


class CarelError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raise loudly if it is not there."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise CarelError(
                f"{LIB_PATH} not found: the CUDA extension is required (no CPU / PyTorch fallback). "
                "Build it with `python -m carel_vae_b200.build`.")
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            if not hasattr(handle, name):
                continue  # declared in a later build step; callers fail loudly when they need it
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def declare(name: str, restype, argtypes) -> None:
    """Register the signature of an entry point (used by the op modules)."""
    _SIGNATURES[name] = (restype, argtypes)
    if _lib is not None and hasattr(_lib, name):
        fn = getattr(_lib, name)
        fn.restype = restype
        fn.argtypes = argtypes


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().carel_last_error()
        raise CarelError(f"{what} failed: {msg.decode() if msg else 'unknown error'}")


def float_array(values):
    arr = (C.c_float * len(values))(*[float(v) for v in values])
    return arr


def launch_count() -> int:
    return int(lib().carel_launch_count())


def set_reserved_sms(n: int) -> None:
    """Leave `n` SMs free in every kernel launched from now on (0 = use all): see carel_set_reserved_sms."""
    lib().carel_set_reserved_sms(int(n))


def reserved_sms() -> int:
    return int(lib().carel_get_reserved_sms())


def set_side_branches(mask) -> None:
    """Bit mask (True = 3): 1 carel_bow_fwd's sparse-target branch, 2 the two O(nnz) gathers of carel_bow_bwd side by side,
    each on the library's second stream (see carel_set_side_branches)."""
    lib().carel_set_side_branches(3 if mask is True else int(mask))


def side_branches() -> int:
    return int(lib().carel_get_side_branches())


def timing_enable(on: bool) -> None:
    lib().carel_timing_enable(1 if on else 0)


def timing_report() -> dict:
    """{tag: (total_ms, launches)} of the tensor-core kernels launched since timing was enabled."""
    buf = C.create_string_buffer(1 << 16)
    lib().carel_timing_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        tag, ms, cnt = line.split()
        out[tag] = (float(ms), int(cnt))
    return out


# ------------------------------------------------------------------------------------------------
# precision tier of the ops (north_star: 1e-3 relative in bf16, 1e-5 in fp32):
#   "bf16"   (default, the product path) 16-bit tensor-core operands, fp32 accumulation
#   "fp32"   fp32 FFMA arithmetic on the CUDA cores over materialised intermediates (csrc/sgemm.cuh): every op of the head
#   "bf16x3" MMD only: operands split into two bf16 terms, three products in one K-concatenated tensor-core GEMM
#            (~1e-5 accurate at 3x the tensor work; fused backward for dim <= 112)
# ------------------------------------------------------------------------------------------------
_PRECISION = {"bf16": 0, "bf16x3": 1, "fp32": 2}
_default_precision = "bf16"


def set_precision(mode: str) -> None:
    global _default_precision
    if mode not in _PRECISION:
        raise ValueError(f"precision must be one of {list(_PRECISION)}")
    _default_precision = mode


def get_precision() -> str:
    return _default_precision


def precision_code(mode=None) -> int:
    return _PRECISION[_default_precision if mode is None else mode]
