"""ctypes binding of include/cake_b200.h (libcake_b200.so).

There is no CPU fallback: importing this module without the built library, or calling into it
without a B200, raises.  The oracle under oracle/ is test infrastructure and is never imported here.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_float, c_int, c_size_t, c_uint32, c_uint64, c_void_p

from .config import CConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
# CAKE_B200_LIB: profiling / A-B builds of the same library (bench_tools/); the product library is the in-tree default
SO_PATH = os.environ.get("CAKE_B200_LIB") or os.path.join(_HERE, "libcake_b200.so")

# every symbol include/cake_b200.h declares: (name, restype, argtypes)
_VP, _I = c_void_p, c_int
SYMBOLS = [
    ("cake_b200_last_error", c_char_p, []),
    ("cake_b200_version", c_char_p, []),
    ("cake_b200_ctx_create", _I, [_I, POINTER(CConfig), POINTER(_VP)]),
    ("cake_b200_ctx_destroy", None, [_VP]),
    ("cake_b200_sync", _I, [_VP]),
    ("cake_b200_stream", _VP, [_VP]),
    ("cake_b200_launch_count", _I, [_VP, POINTER(c_uint64)]),
    ("cake_b200_dev_alloc", _I, [_VP, c_size_t, POINTER(_VP)]),
    ("cake_b200_dev_free", _I, [_VP, _VP]),
    ("cake_b200_block_load", _I, [_VP, _I] + [_VP] * 14 + [POINTER(_VP)]),
    ("cake_b200_block_free", None, [_VP]),
    ("cake_b200_block_layer", _I, [_VP]),
    ("cake_b200_cache_create", _I, [_VP, _I, _I, POINTER(_VP)]),
    ("cake_b200_cache_clear", _I, [_VP]),
    ("cake_b200_cache_free", None, [_VP]),
    ("cake_b200_cache_len", _I, [_VP, _I]),
    ("cake_b200_cache_read", _I, [_VP, _I, _I, _VP, c_size_t]),
    ("cake_b200_cache_fill_synthetic", _I, [_VP, POINTER(_I), _I, _I, c_uint32]),
    ("cake_b200_forward_batch", _I, [_VP, POINTER(_VP), POINTER(_I), _I, _VP, _VP, _VP, _I, _I, _I]),
    ("cake_b200_forward_batch_host", _I, [_VP, POINTER(_VP), POINTER(_I), _I, _VP, _VP, _VP, _I, _I, _I]),
    ("cake_b200_head_load", _I, [_VP, _VP, _VP, _VP]),
    ("cake_b200_embed", _I, [_VP, POINTER(c_uint32), _I, _I, _VP]),
    ("cake_b200_logits", _I, [_VP, _VP, _I, _I, _VP, POINTER(c_uint32)]),
    ("cake_b200_repeat_penalty_argmax", _I, [_VP, _VP, c_float, POINTER(c_uint32), _I, POINTER(c_uint32)]),
    ("cake_b200_comm_unique_id", _I, [_VP]),
    ("cake_b200_comm_init", _I, [_VP, _VP, _I, _I]),
    ("cake_b200_send", _I, [_VP, _VP, c_size_t, _I]),
    ("cake_b200_recv", _I, [_VP, _VP, c_size_t, _I]),
    ("cake_b200_ring_export", _I, [_VP, _VP]),
    ("cake_b200_ring_import", _I, [_VP, _VP]),
    ("cake_b200_decode_build", _I, [_VP, POINTER(_VP), POINTER(_I), _I, _VP, _I, _I]),
    ("cake_b200_decode_begin", _I, [_VP, c_uint32, _I]),
    ("cake_b200_decode_run", _I, [_VP, _I]),
    ("cake_b200_decode_tokens", _I, [_VP, POINTER(c_uint32), _I]),
    ("cake_b200_decode_step_host", _I, [_VP, c_uint32, POINTER(c_uint32)]),
    ("cake_b200_decode_logits", _I, [_VP, _VP, c_size_t]),
    ("cake_b200_bench_kernel", _I, [_VP, POINTER(_VP), POINTER(_I), _I, _VP, _I, _I, POINTER(c_float)]),
    ("cake_b200_decode_trace", _I, [_VP, POINTER(c_uint64), _I]),
]


class CSampling(ctypes.Structure):
    """``cake_b200_sampling`` (include/cake_b200.h)."""
    _fields_ = [("kind", c_int), ("top_k", c_int), ("temperature", c_float), ("top_p", c_float), ("seed", c_uint64)]


SYMBOLS += [
    ("cake_b200_sample", _I, [_VP, _VP, POINTER(CSampling), c_float, POINTER(c_uint32), _I, c_uint64, POINTER(c_float), POINTER(c_uint32)]),
    ("cake_b200_decode_set_sampling", _I, [_VP, POINTER(CSampling)]),
    ("cake_b200_load_stats", _I, [_VP, POINTER(ctypes.c_double), POINTER(ctypes.c_double)]),
]


class CBlockVariant(ctypes.Structure):
    """``cake_b200_block_variant`` (include/cake_b200.h)."""
    _fields_ = [("sliding_window", c_int), ("use_rope", c_int), ("post_attention_norm", _VP), ("post_feedforward_norm", _VP)]


SYMBOLS += [("cake_b200_block_set_variant", _I, [_VP, POINTER(CBlockVariant)])]


class CakeB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cake_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                f"{SO_PATH} is missing: build it with `python -m cake_b200.build` "
                "(libcake_b200 has no CPU or PyTorch fallback)")
        L = ctypes.CDLL(SO_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, res, args in SYMBOLS:
            if os.environ.get("CAKE_B200_LIB") and not hasattr(L, name):
                continue  # A/B against an older build of the library (bench_tools/ab.sh); the product library must export everything
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise CakeB200Error(rc, lib().cake_b200_last_error().decode(errors="replace"))


def ptr(t) -> int:
    """Raw address of a torch tensor / numpy array / int / None."""
    if t is None:
        return 0
    if isinstance(t, int):
        return t
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def ptr_array(handles):
    return (c_void_p * len(handles))(*handles)


def int_array(vals):
    return (c_int * len(vals))(*vals)


__all__ = ["lib", "check", "ptr", "ptr_array", "int_array", "CakeB200Error", "SYMBOLS", "SO_PATH", "byref",
           "c_uint32", "c_uint64", "c_void_p", "c_int"]
