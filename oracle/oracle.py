"""ctypes wrapper around oracle/_build/libcake_oracle.so (the C restatement in cake_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of cake_oracle.c.  Imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs, never by cake_b200/.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np
import torch

from cake_b200.config import CConfig, Config, DTYPES

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcake_oracle.so")


class OraConfig(ctypes.Structure):
    _fields_ = list(CConfig._fields_) + [("silu_mode", ctypes.c_int)]


class OraLayer(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in
                ("q", "k", "v", "o", "gate", "up", "down", "ln1", "ln2",
                 "q_bias", "k_bias", "v_bias", "q_norm", "k_norm", "post_attn", "post_ffn")] + \
               [("window", ctypes.c_int), ("no_rope", ctypes.c_int)]


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "cake_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        # The oracle links the system libgomp (torch bundles its own copy): pin its threads, otherwise the two
        # OpenMP runtimes fight over cores (measured 1.2 GB/s unbound vs 68 GB/s bound on 8 cores).
        os.environ.setdefault("OMP_PROC_BIND", "true")
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")
        L = ctypes.CDLL(_SO)
        vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        L.ora_model_create.restype = vp
        L.ora_model_create.argtypes = [ctypes.POINTER(OraConfig)]
        L.ora_model_set_layer.argtypes = [vp, ci, ctypes.POINTER(OraLayer)]
        L.ora_model_set_head.argtypes = [vp, vp, vp, vp]
        L.ora_model_free.argtypes = [vp]
        L.ora_model_cos.restype = ctypes.POINTER(cf)
        L.ora_model_cos.argtypes = [vp]
        L.ora_model_sin.restype = ctypes.POINTER(cf)
        L.ora_model_sin.argtypes = [vp]
        L.ora_cache_create.restype = vp
        L.ora_cache_create.argtypes = [ctypes.POINTER(OraConfig), ci]
        L.ora_cache_clear.argtypes = [vp]
        L.ora_cache_free.argtypes = [vp]
        L.ora_cache_len.restype = ci
        L.ora_cache_len.argtypes = [vp, ci]
        L.ora_cache_k.restype = ctypes.POINTER(cf)
        L.ora_cache_k.argtypes = [vp, ci]
        L.ora_cache_v.restype = ctypes.POINTER(cf)
        L.ora_cache_v.argtypes = [vp, ci]
        L.ora_cache_set_len.argtypes = [vp, ci, ci]
        L.ora_block_forward.restype = ci
        L.ora_block_forward.argtypes = [vp, ci, vp, vp, ci, ci, vp]
        L.ora_forward.restype = ci
        L.ora_forward.argtypes = [vp, vp, vp, ci, ci, vp]
        L.ora_forward_layers.restype = ci
        L.ora_forward_layers.argtypes = [vp, vp, ci, ci, vp, ci, ci]
        L.ora_embed.argtypes = [vp, vp, ci, vp]
        L.ora_logits.argtypes = [vp, vp, ci, vp]
        L.ora_argmax.restype = ctypes.c_uint32
        L.ora_argmax.argtypes = [vp, ci]
        L.ora_repeat_penalty.argtypes = [vp, ci, cf, vp, ci, ci]
        L.ora_rms_norm.argtypes = [vp, ci, ci, vp, cf, vp, ci]
        L.ora_linear.argtypes = [vp, ci, ci, vp, ci, vp, vp, ci, ci]
        L.ora_rope.argtypes = [vp, ci, ci, ci, vp, vp, ci, ci]
        L.ora_causal_mask.argtypes = [ci, vp]
        L.ora_silu_mul.restype = cf
        L.ora_silu_mul.argtypes = [cf, cf, ci, ci]
        L.ora_gelu_mul.restype = cf
        L.ora_gelu_mul.argtypes = [cf, cf, ci]
        L.ora_round.restype = cf
        L.ora_round.argtypes = [cf, ci]
        L.ora_num_threads.restype = ci
        L.ora_set_dot_mode.argtypes = [ci]
        L.ora_dot_mode.restype = ci
        L.ora_set_num_threads.argtypes = [ci]
        _lib = L
    return _lib


def _f32(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        a = a.detach().float().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a) -> int:
    if a is None:
        return 0
    if isinstance(a, torch.Tensor):
        return a.data_ptr()
    return a.ctypes.data


def round_to(a, dtype: str) -> np.ndarray:
    """Round an f32 array to the model dtype and back (values carried in float)."""
    t = torch.from_numpy(_f32(a))
    from cake_b200.synth import TORCH_DTYPES
    return t.to(TORCH_DTYPES[dtype]).float().numpy()


def set_dot_mode(mode: int) -> None:
    """Summation-order probe of the linear layers (0 default, 1 second f32 order, 2 f64 accumulate): see cake_oracle.c."""
    lib().ora_set_dot_mode(mode)


# ---- op-level entry points (used by the known-answer tests) -----------------------------------
def rms_norm(x, w: torch.Tensor, eps: float, dtype: str) -> np.ndarray:
    x = _f32(x)
    n = x.shape[-1]
    out = np.empty_like(x)
    w = w.contiguous()
    lib().ora_rms_norm(_ptr(x), x.size // n, n, _ptr(w), eps, _ptr(out), DTYPES[dtype])
    return out


def linear(x, W: torch.Tensor, bias: Optional[torch.Tensor], dtype: str) -> np.ndarray:
    x = _f32(x)
    K = x.shape[-1]
    S = x.size // K
    N = W.shape[0]
    W = W.contiguous()
    out = np.empty(x.shape[:-1] + (N,), dtype=np.float32)
    lib().ora_linear(_ptr(x), S, K, _ptr(W), N, _ptr(bias), _ptr(out), N, DTYPES[dtype])
    return out


def causal_mask(seq: int) -> np.ndarray:
    out = np.empty((seq, seq), dtype=np.uint8)
    lib().ora_causal_mask(seq, _ptr(out))
    return out


def silu_mul(g: float, u: float, dtype: str, mode: int = 0) -> float:
    return lib().ora_silu_mul(g, u, DTYPES[dtype], mode)


def gelu_mul(g: float, u: float, dtype: str) -> float:
    return lib().ora_gelu_mul(g, u, DTYPES[dtype])


def argmax(logits) -> int:
    l = _f32(logits)
    return int(lib().ora_argmax(_ptr(l), l.size))


def repeat_penalty(logits, penalty: float, ctx, dtype: str) -> np.ndarray:
    l = _f32(logits).copy()
    c = np.ascontiguousarray(ctx, dtype=np.uint32)
    lib().ora_repeat_penalty(_ptr(l), l.size, penalty, _ptr(c), c.size, DTYPES[dtype])
    return l


class OracleCache:
    def __init__(self, model: "OracleModel", cap: int):
        self.model = model
        self.cap = cap
        self.h = lib().ora_cache_create(ctypes.byref(model.ccfg), cap)

    def clear(self):
        lib().ora_cache_clear(self.h)

    def len(self, layer: int) -> int:
        return lib().ora_cache_len(self.h, layer)

    def kv(self, layer: int):
        """(K, V) numpy views of shape (n_kv, cap, hd) (whole capacity; valid rows are [:len])."""
        c = self.model.cfg
        shape = (c.num_key_value_heads, self.cap, c.hd)
        n = int(np.prod(shape))
        k = np.ctypeslib.as_array(lib().ora_cache_k(self.h, layer), shape=(n,)).reshape(shape)
        v = np.ctypeslib.as_array(lib().ora_cache_v(self.h, layer), shape=(n,)).reshape(shape)
        return k, v

    def set_len(self, layer: int, n: int):
        lib().ora_cache_set_len(self.h, layer, n)

    def __del__(self):
        try:
            lib().ora_cache_free(self.h)
        except Exception:
            pass


class OracleModel:
    """Restatement of TextModelBase + Transformer blocks over an HF-named state dict."""

    def __init__(self, cfg: Config, weights: dict, dtype: str = "bf16", max_seq: Optional[int] = None,
                 silu_mode: int = 0, layers: Optional[range] = None):
        self.cfg, self.dtype = cfg, dtype
        cc = CConfig.from_config(cfg, dtype, max_seq)
        self.ccfg = OraConfig(*[getattr(cc, f) for f, _ in CConfig._fields_], silu_mode)
        self.h = lib().ora_model_create(ctypes.byref(self.ccfg))
        self._keep = []
        p = cfg.model_prefix

        def W(name):
            t = weights.get(name)
            if t is None:
                return None
            t = t.detach().cpu().contiguous()
            self._keep.append(t)
            return t

        from cake_b200.loader import BLOCK_TENSORS, EXTRA_TENSORS, block_tensors  # same tensor naming / fused-checkpoint rules as the loader
        for i in (layers if layers is not None else range(cfg.num_hidden_layers)):
            views = block_tensors(weights, cfg, cfg.layer_name(i))
            var = cfg.layer_variant(i, self.ccfg.max_seq)
            ptrs = []
            for short in BLOCK_TENSORS + EXTRA_TENSORS:
                t = views[short]
                if t is not None:
                    t = t.detach().cpu().contiguous()
                    self._keep.append(t)
                ptrs.append(_ptr(t))
            ol = OraLayer(*ptrs, var["window"], int(var["no_rope"]))
            lib().ora_model_set_layer(self.h, i, ctypes.byref(ol))
        from cake_b200.loader import rms_norm_weight
        emb, lnf, head = W(f"{p}.embed_tokens.weight"), W(f"{p}.norm.weight"), W("lm_head.weight")
        if lnf is not None and cfg.residual_rms_norm:
            lnf = rms_norm_weight(lnf, cfg).contiguous()
            self._keep.append(lnf)
        lib().ora_model_set_head(self.h, _ptr(emb), _ptr(lnf), _ptr(head))

    def new_cache(self, cap: Optional[int] = None) -> OracleCache:
        return OracleCache(self, cap or self.ccfg.max_seq)

    def rope_tables(self):
        rot = int(self.cfg.hd * self.cfg.partial_rotary_factor)
        n = self.ccfg.max_seq * (rot // 2)
        shape = (self.ccfg.max_seq, rot // 2)
        return (np.ctypeslib.as_array(lib().ora_model_cos(self.h), shape=(n,)).reshape(shape),
                np.ctypeslib.as_array(lib().ora_model_sin(self.h), shape=(n,)).reshape(shape))

    def block_forward(self, layer: int, x, index_pos: int, cache: OracleCache) -> np.ndarray:
        x = _f32(x).reshape(-1, self.cfg.hidden_size)
        out = np.empty_like(x)
        rc = lib().ora_block_forward(self.h, layer, cache.h, _ptr(x), x.shape[0], index_pos, _ptr(out))
        if rc:
            raise RuntimeError(f"oracle block_forward rc={rc}")
        return out

    def forward_layers(self, x, l0: int, l1: int, index_pos: int, cache: OracleCache) -> np.ndarray:
        x = _f32(x).reshape(-1, self.cfg.hidden_size).copy()
        rc = lib().ora_forward_layers(self.h, cache.h, l0, l1, _ptr(x), x.shape[0], index_pos)
        if rc:
            raise RuntimeError(f"oracle forward_layers rc={rc}")
        return x

    def embed(self, ids) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        x = np.empty((ids.size, self.cfg.hidden_size), dtype=np.float32)
        lib().ora_embed(self.h, _ptr(ids), ids.size, _ptr(x))
        return x

    def logits(self, x) -> np.ndarray:
        x = _f32(x).reshape(-1, self.cfg.hidden_size)
        out = np.empty(self.cfg.vocab_size, dtype=np.float32)
        lib().ora_logits(self.h, _ptr(x), x.shape[0], _ptr(out))
        return out

    def forward(self, ids, index_pos: int, cache: OracleCache) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        out = np.empty(self.cfg.vocab_size, dtype=np.float32)
        rc = lib().ora_forward(self.h, cache.h, _ptr(ids), ids.size, index_pos, _ptr(out))
        if rc:
            raise RuntimeError(f"oracle forward rc={rc}")
        return out

    def generate(self, prompt_ids, n_new: int, cache: Optional[OracleCache] = None):
        """text_model.rs:397-495 + master.rs:131-155 greedy loop. Returns (tokens, per-step logits)."""
        cache = cache or self.new_cache()
        cache.clear()
        toks, all_logits = [], []
        ids, pos = list(prompt_ids), 0
        for _ in range(n_new):
            lg = self.forward(ids, pos, cache)
            pos += len(ids)
            t = argmax(lg)
            toks.append(t)
            all_logits.append(lg)
            ids = [t]
        return toks, all_logits

    def __del__(self):
        try:
            lib().ora_model_free(self.h)
        except Exception:
            pass
