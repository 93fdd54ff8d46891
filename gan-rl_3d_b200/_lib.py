"""ctypes binding of librlg_b200.so (include/rlg_b200.h).

The product path has no CPU implementation: if the library is missing or a call fails this module raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "lib", "librlg_b200.so")
EXP_LIB_PATH = os.path.join(_PKG_DIR, "lib", "librlg_b200_exp.so")

_lib: Optional[ctypes.CDLL] = None

c_float_p = ctypes.POINTER(ctypes.c_float)
c_int32_p = ctypes.POINTER(ctypes.c_int32)


class RlgLayer(ctypes.Structure):
    """struct rlg_layer (include/rlg_b200.h)."""
    _fields_ = [("w", ctypes.c_void_p), ("b", ctypes.c_void_p),
                ("c_in", ctypes.c_int32), ("c_out", ctypes.c_int32)]


class RlgBnLayer(ctypes.Structure):
    """struct rlg_bn_layer (include/rlg_b200.h)."""
    _fields_ = [("w", ctypes.c_void_p), ("b", ctypes.c_void_p), ("gamma", ctypes.c_void_p), ("beta", ctypes.c_void_p),
                ("running_mean", ctypes.c_void_p), ("running_var", ctypes.c_void_p),
                ("eps", ctypes.c_float), ("momentum", ctypes.c_float), ("c_in", ctypes.c_int32), ("c_out", ctypes.c_int32)]


class RlgBnGrads(ctypes.Structure):
    """struct rlg_bn_grads (include/rlg_b200.h)."""
    _fields_ = [("dw", ctypes.c_void_p), ("db", ctypes.c_void_p), ("dgamma", ctypes.c_void_p), ("dbeta", ctypes.c_void_p)]


class RlgPreparePlan(ctypes.Structure):
    """struct rlg_prepare_plan (include/rlg_b200.h)."""
    _fields_ = [(n, ctypes.c_void_p) for n in ("item", "method", "n_keep", "keep_idx", "center", "q_index", "q_gamma", "rot",
                                                "scale", "jitter", "pad_idx")]


class RlgError(RuntimeError):
    def __init__(self, fn: str, code: int, msg: str):
        super().__init__(f"{fn} failed with code {code}: {msg}")
        self.code = code


# every symbol include/rlg_b200.h declares; tests/test_abi.py checks the .so exports each of them
EXPORTS = {
    "rlg_version": (ctypes.c_int, []),
    "rlg_last_error": (ctypes.c_char_p, []),
    "rlg_device_sm_count": (ctypes.c_int, []),
    "rlg_chamfer_ws_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "rlg_chamfer_fwd": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint, ctypes.c_void_p]),
    "rlg_chamfer_bwd": (ctypes.c_int, [ctypes.c_void_p] * 8 + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 2
                        + [ctypes.c_uint, ctypes.c_void_p]),
    "rlg_chamfer_loss_fwd": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float,
                                            ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint, ctypes.c_void_p]),
    "rlg_chamfer_loss_bwd": (ctypes.c_int, [ctypes.c_void_p] * 7 + [ctypes.c_float] * 2 + [ctypes.c_int] * 3
                             + [ctypes.c_void_p] * 2 + [ctypes.c_uint, ctypes.c_void_p]),
    "rlg_chamfer_bwd_ws_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "rlg_chamfer_bwd_det": (ctypes.c_int, [ctypes.c_void_p] * 8 + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 3
                            + [ctypes.c_size_t, ctypes.c_uint, ctypes.c_void_p]),
    "rlg_chamfer_loss_bwd_det": (ctypes.c_int, [ctypes.c_void_p] * 7 + [ctypes.c_float] * 2 + [ctypes.c_int] * 3
                                 + [ctypes.c_void_p] * 3 + [ctypes.c_size_t, ctypes.c_uint, ctypes.c_void_p]),
    "rlg_encoder_ws_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(RlgLayer), ctypes.c_int]),
    "rlg_encoder_fwd": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(RlgLayer),
                                       ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "rlg_encoder_pack_bytes": (ctypes.c_size_t, [ctypes.POINTER(RlgLayer), ctypes.c_int]),
    "rlg_encoder_pack_bf16": (ctypes.c_int, [ctypes.POINTER(RlgLayer), ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                                             ctypes.c_void_p]),
    "rlg_encoder_fwd_bf16": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(RlgLayer),
                                            ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                            ctypes.c_void_p]),
    "rlg_encoder_gemm_pack_bytes": (ctypes.c_size_t, [ctypes.POINTER(RlgLayer), ctypes.c_int, ctypes.c_int]),
    "rlg_encoder_gemm_pack": (ctypes.c_int, [ctypes.POINTER(RlgLayer), ctypes.c_int, ctypes.c_int, c_float_p, ctypes.c_void_p,
                                             ctypes.c_size_t, ctypes.c_void_p]),
    "rlg_encoder_gemm_ws_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(RlgLayer), ctypes.c_int,
                                                    ctypes.c_int]),
    "rlg_encoder_gemm_fwd": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(RlgLayer), ctypes.c_int,
                                            ctypes.c_int, c_float_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "rlg_encoder_train_saved_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(RlgBnLayer), ctypes.c_int]),
    "rlg_encoder_train_ws_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(RlgBnLayer), ctypes.c_int]),
    "rlg_encoder_train_saved_layout": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(RlgBnLayer), ctypes.c_int,
                                                      ctypes.POINTER(ctypes.c_size_t), ctypes.c_int]),
    "rlg_encoder_train_fwd": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(RlgBnLayer), ctypes.c_int,
                                             ctypes.c_uint, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                             ctypes.c_size_t, ctypes.c_void_p]),
    "rlg_encoder_train_bwd": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(RlgBnLayer), ctypes.c_int,
                                             ctypes.c_uint, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                             ctypes.POINTER(RlgBnGrads), ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "rlg_batch_prepare": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(RlgPreparePlan),
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "rlg_fp32_peak": (ctypes.c_int, [c_float_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
}

CHAMFER_WS_CLEAN = 1
CHAMFER_ALGO_SIMPLE = 2
CHAMFER_TILE_ONLY = 4
CHAMFER_ALGO_TENSOR = 16
CHAMFER_TRACK_TWO = 32
CHAMFER_FILTER_ONLY = 64


def chamfer_reserve_sms(n: int) -> int:
    """RLG_CHAMFER_RESERVE_SMS(n)"""
    return (int(n) & 0xff) << 16


# experiments build only (librlg_b200_exp.so, tools/): kernel variants in bits 8-11, first-generation tensor sweep
X_CHAMFER_TENSOR_V1 = 128
CHAMFER_BWD_ACCUMULATE = 1
ENC_BF16 = 1
ENC_FP32X = 2
ENC_BATCH_STATS = 1


def load() -> ctypes.CDLL:
    """dlopen the in-tree library.  Raises (never falls back) if it has not been built.
    RLG_EXPERIMENTS_LIB=1 (tools/ only) loads the experiments build instead: same ABI plus timing variants."""
    global _lib
    if _lib is not None:
        return _lib
    path = EXP_LIB_PATH if os.environ.get("RLG_EXPERIMENTS_LIB") == "1" else LIB_PATH
    if os.environ.get("RLG_EXPERIMENTS_LIB", "").endswith(".so"):       # tools/: an A-B variant built by build.build_variant
        path = os.environ["RLG_EXPERIMENTS_LIB"]
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python gan-rl_3d_b200/build.py` (or __graft_entry__.build()). "
            "There is no CPU fallback for the B200 hot path.")
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in EXPORTS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(fn: str, rc: int) -> None:
    if rc != 0:
        msg = load().rlg_last_error()
        raise RlgError(fn, rc, msg.decode("utf-8", "replace") if msg else "")


# ---- optional NVTX ranges around the library calls: RLG_NVTX=1, or set_nvtx(True) -----------------
_NVTX = os.environ.get("RLG_NVTX", "0") == "1"


def set_nvtx(enabled: bool) -> None:
    global _NVTX
    _NVTX = bool(enabled)


def nvtx_push(name: str) -> None:
    """Open an NVTX range around a library call (nsys / ncu --nvtx timelines); a no-op unless switched on."""
    if _NVTX:
        import torch
        torch.cuda.nvtx.range_push(name)


def nvtx_pop() -> None:
    if _NVTX:
        import torch
        torch.cuda.nvtx.range_pop()
