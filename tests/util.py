"""Shared helpers for the parity tests: build the same synthetic checkpoint for the oracle and the
CUDA path, and compare in units of the model dtype's ulp."""
import numpy as np
import torch

from cake_b200.config import Config
from cake_b200.synth import TORCH_DTYPES, make_checkpoint

ULP = {"bf16": 2.0 ** -7, "f16": 2.0 ** -10}  # spacing relative to the binade (1 ulp of D)


def ulp_at_scale(ref: np.ndarray, dtype: str) -> float:
    """1 ulp of D at the magnitude of the largest reference element.  Block outputs are sums of terms of
    that magnitude (residual + projection), so a 1-ulp flip of an intermediate shows up as an absolute
    error of this size even on outputs that happen to be small."""
    scale = max(float(np.abs(ref).max()), 2.0 ** -6)
    return ULP[dtype] * 2.0 ** np.floor(np.log2(scale))


def max_ulp_err(a: np.ndarray, ref: np.ndarray, dtype: str) -> float:
    """max |a-ref| in ulps of D at the tensor's scale (see ulp_at_scale)."""
    a, ref = np.asarray(a, np.float32), np.asarray(ref, np.float32)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    assert np.isfinite(a).all()
    return float(np.abs(a - ref).max() / ulp_at_scale(ref, dtype))


def mean_ulp_err(a: np.ndarray, ref: np.ndarray, dtype: str) -> float:
    a, ref = np.asarray(a, np.float32), np.asarray(ref, np.float32)
    return float(np.abs(a - ref).mean() / ulp_at_scale(ref, dtype))


def to_np(t: torch.Tensor) -> np.ndarray:
    return t.detach().float().cpu().numpy()


def rand_x(shape, dtype: str, seed: int, scale: float = 1.0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(TORCH_DTYPES[dtype])


def medium_config(**kw) -> Config:
    """hidden 512, 8 heads / 2 kv heads of 64, inter 1024: big enough to exercise every kernel's
    tiling (multi-CTA row split, K split across warps, several attention splits), small enough for the
    oracle to run in well under a second."""
    base = dict(hidden_size=512, intermediate_size=1024, vocab_size=1024, num_hidden_layers=3,
                num_attention_heads=8, num_key_value_heads=2, rms_norm_eps=1e-5, rope_theta=500000.0,
                max_seq_len=512, head_dim=64)
    base.update(kw)
    return Config(**base)


def checkpoint(cfg: Config, dtype: str, seed: int = 1234, std: float = 0.05, peaked: bool = False):
    return make_checkpoint(cfg, dtype, seed=seed, std=std, peaked=peaked)


def f32_to_bits(a, dtype: str) -> np.ndarray:
    """f32 values that are exactly representable in D -> their 16-bit patterns (what travels on the wire)."""
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(TORCH_DTYPES[dtype])
    return t.view(torch.uint16).numpy()


def bits_to_f32(bits, dtype: str) -> np.ndarray:
    t = torch.from_numpy(np.ascontiguousarray(bits, dtype=np.uint16).copy())
    return t.view(TORCH_DTYPES[dtype]).float().numpy()
