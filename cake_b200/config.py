"""Model configuration for the block-forward path.

Mirrors the reference's generalized ``Config`` (cake-core/src/models/common/config.rs:87-150) for the
fields the dense Llama-family block path reads, and the HF ``config.json`` mappings of the architectures
whose blocks are exactly that path: ``LlamaConfig::into_config`` (models/llama3/config.rs:62-98),
``Qwen3Config`` (qwen3/config.rs:55-93), ``Qwen2Config`` (qwen2/config.rs:69-105, q/k/v bias),
``MistralConfig`` (mistral/config.rs:56-93, without an active sliding window), ``Falcon3Config``
(falcon3/config.rs:53-90) and ``Phi4Config`` (phi4/config.rs:63-100: pre-fused qkv / gate_up tensors, partial
rotary).  Field names follow the reference.
"""
from __future__ import annotations

import ctypes
import json
from dataclasses import dataclass, field, asdict
from typing import Optional

import numpy as np

DTYPES = {"bf16": 0, "f16": 1, "f32": 2}


@dataclass
class RopeScaling:
    """config.rs RopeScaling (llama3 type only; cache.rs:49-80)."""
    factor: float = 1.0
    low_freq_factor: float = 1.0
    high_freq_factor: float = 4.0
    original_max_position_embeddings: int = 0
    rope_type: Optional[str] = None


@dataclass
class Config:
    hidden_size: int
    intermediate_size: int
    vocab_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_key_value_heads: int
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    rope_scaling: Optional[RopeScaling] = None
    tie_word_embeddings: bool = False
    max_seq_len: int = 4096
    use_qkv_bias: bool = False
    model_prefix: str = "model"
    head_dim: Optional[int] = None
    partial_rotary_factor: float = 1.0
    use_qk_norm: bool = False
    eos_token_id: list = field(default_factory=list)
    sliding_window: Optional[int] = None
    fused_qkv_proj: bool = False      # Phi-3/4: one 'qkv_proj' tensor = cat(q, k, v) (attention.rs:90-94)
    fused_gate_up_proj: bool = False  # Phi-3/4: one 'gate_up_proj' tensor = cat(gate, up) (mlp.rs:38-40)
    use_gelu_mlp: bool = False        # config.rs:133: gelu_tanh(gate) * up (mlp.rs:25-26)
    embed_scale: Optional[float] = None   # config.rs:136: embeddings scaled before the blocks (text_model.rs:274-276)
    residual_rms_norm: bool = False   # config.rs:113: norm weights stored as deltas; (1 + w) in f32 at LOAD time (config.rs:155-173)
    pre_reshape_qk_norm: bool = False  # config.rs:116 (OLMo2): QK-norm over the whole q / k projection, before the head reshape
    global_layers: list = field(default_factory=list)  # config.rs:126-130: per-layer schedule, True = global (full context)
    block_kind: str = "llama"         # which block.rs the layers are: "llama" (transformer.rs), "olmo2", "gemma3", "exaone4"

    def layer_variant(self, i: int, max_seq: Optional[int] = None) -> dict:
        """Norm placement and attention mode of layer ``i``, as the reference's per-architecture block.rs sets them up:
          llama   (common/transformer.rs:103-135)  pre-norms, the config's window, RoPE
          olmo2   (olmo2/block.rs:62-90)           NO pre-norms, post-attention / post-feedforward norms, RoPE
          gemma3  (gemma3/block.rs:60-135)         all four norms; global layers: full context + RoPE, local: window, NO RoPE
          exaone4 (exaone4/block.rs:50-110)        pre-norms; global layers: full context, NO RoPE, local: window + RoPE
        ``window``: -1 = the config's, 0 = none, > 0 = this layer's (passed as 0 when it can never bite)."""
        def win(w):
            return int(w) if (w and w < (max_seq or self.max_seq_len)) else 0
        is_global = bool(self.global_layers[i]) if i < len(self.global_layers) else False
        if self.block_kind == "olmo2":
            return dict(pre_norms=False, post_norms=True, window=-1, no_rope=False)
        if self.block_kind == "gemma3":
            return dict(pre_norms=True, post_norms=True, window=0 if is_global else win(self.sliding_window), no_rope=not is_global)
        if self.block_kind == "exaone4":
            return dict(pre_norms=True, post_norms=False, window=0 if is_global else win(self.sliding_window), no_rope=is_global)
        if self.block_kind == "exaone4_hf":
            # NOT a reference block: the structure HF transformers gives EXAONE 4.0 (post-norms like OLMo2 + the hybrid
            # schedule).  Test-only: it pins the oracle's per-layer NoPE / window handling to an independent implementation
            # (tests/golden/hf_exaone4_hf_tiny.npz); the reference's own EXAONE4 block is pre-norm (exaone4/block.rs:96-110).
            return dict(pre_norms=False, post_norms=True, window=0 if is_global else win(self.sliding_window), no_rope=is_global)
        return dict(pre_norms=True, post_norms=False, window=-1, no_rope=False)

    @property
    def standard_blocks(self) -> bool:
        return self.block_kind == "llama"

    @property
    def hd(self) -> int:
        # attention.rs:85
        return self.head_dim if self.head_dim else self.hidden_size // self.num_attention_heads

    @property
    def size_q(self) -> int:
        return self.hd * self.num_attention_heads

    @property
    def size_kv(self) -> int:
        return self.hd * self.num_key_value_heads

    def layer_name(self, i: int) -> str:
        # text_model.rs:205  "{prefix}.layers.{i}"
        return f"{self.model_prefix}.layers.{i}"

    # ---- HF config.json mappings -------------------------------------------------------------
    # arch -> (default rope_theta, default max_position_embeddings, fixed flags): the serde defaults and the
    # constants each ``into_config`` hard-wires.  Strings the reference does not know fall back to Llama
    # (cake/mod.rs:81-109); architectures it knows but whose block is not this path are refused here.
    _ARCHS = {
        "LlamaForCausalLM": (500000.0, 4096, dict()),
        "Qwen2ForCausalLM": (1000000.0, 32768, dict(use_qkv_bias=True)),
        "Qwen3ForCausalLM": (1000000.0, 40960, dict(use_qk_norm=True, _head_dim=True)),
        "MistralForCausalLM": (1000000.0, 131072, dict(_head_dim=True, _sliding_window=True)),
        "FalconForCausalLM": (500000.0, 131072, dict(_head_dim=True)),
        "Phi3ForCausalLM": (1000000.0, 131072, dict(_head_dim=True, _partial=True, _fused=True)),
        "Phi4ForCausalLM": (1000000.0, 131072, dict(_head_dim=True, _partial=True, _fused=True)),
    }
    _OTHER_BLOCKS = ("Qwen3_5ForConditionalGeneration", "Qwen3MoeForCausalLM", "Qwen3_5MoeForConditionalGeneration",
                     "LuxTTSForTextToSpeech")
    _SIBLING_BLOCKS = {"Gemma3ForCausalLM": "gemma3", "OLMo2ForCausalLM": "olmo2", "Olmo2ForCausalLM": "olmo2",
                       "ExaoneForCausalLM": "exaone4"}  # cake/mod.rs:99-105

    @staticmethod
    def detect_arch(d: dict) -> str:
        """config.rs:175-190 detect_text_model_arch: the first *string* entry of ``architectures``, else ""."""
        archs = d.get("architectures")
        if isinstance(archs, list):
            for a in archs:
                if isinstance(a, str):
                    return a
        return ""

    @staticmethod
    def from_hf(d: dict) -> "Config":
        arch = Config.detect_arch(d)
        if arch in Config._OTHER_BLOCKS:
            raise ValueError(f"architecture {arch!r} is outside the block-forward path built here")
        if arch in Config._SIBLING_BLOCKS:
            return Config._from_hf_sibling(d, Config._SIBLING_BLOCKS[arch])
        rope_default, max_default, flags = Config._ARCHS.get(arch, Config._ARCHS["LlamaForCausalLM"])
        rs = d.get("rope_scaling")
        rope = None
        if rs:
            rope = RopeScaling(
                # serde defaults of the reference's RopeScaling: every missing number is 0.0 (config.rs)
                factor=float(rs.get("factor", 0.0)),
                low_freq_factor=float(rs.get("low_freq_factor", 0.0)),
                high_freq_factor=float(rs.get("high_freq_factor", 0.0)),
                original_max_position_embeddings=int(rs.get("original_max_position_embeddings", 0)),
                rope_type=rs.get("rope_type"),
            )
        eos = d.get("eos_token_id")
        eos = [] if eos is None else (list(eos) if isinstance(eos, (list, tuple)) else [eos])
        return Config(
            hidden_size=d["hidden_size"],
            intermediate_size=d["intermediate_size"],
            vocab_size=d["vocab_size"],
            num_hidden_layers=d["num_hidden_layers"],
            num_attention_heads=d["num_attention_heads"],
            num_key_value_heads=d.get("num_key_value_heads") or d["num_attention_heads"],
            rms_norm_eps=d["rms_norm_eps"],
            rope_theta=float(d.get("rope_theta") or rope_default),
            rope_scaling=rope,
            tie_word_embeddings=bool(d.get("tie_word_embeddings", False)),
            max_seq_len=int(d.get("max_position_embeddings") or max_default),
            eos_token_id=eos,
            use_qkv_bias=bool(flags.get("use_qkv_bias", False)),
            use_qk_norm=bool(flags.get("use_qk_norm", False)),
            head_dim=d.get("head_dim") if flags.get("_head_dim") else None,
            sliding_window=d.get("sliding_window") if flags.get("_sliding_window") else None,
            partial_rotary_factor=float(d.get("partial_rotary_factor") or 1.0) if flags.get("_partial") else 1.0,
            fused_qkv_proj=bool(flags.get("_fused")), fused_gate_up_proj=bool(flags.get("_fused")),
        )

    @staticmethod
    def _from_hf_sibling(d: dict, kind: str) -> "Config":
        """models/{gemma3,olmo2,exaone4}/config.rs ``into_config``: the serde defaults and the flags each hard-wires."""
        n = int(d["num_hidden_layers"])
        rs = d.get("rope_scaling")
        rope = RopeScaling(factor=float(rs.get("factor", 0.0)), low_freq_factor=float(rs.get("low_freq_factor", 0.0)),
                           high_freq_factor=float(rs.get("high_freq_factor", 0.0)),
                           original_max_position_embeddings=int(rs.get("original_max_position_embeddings", 0)),
                           rope_type=rs.get("rope_type")) if rs else None
        eos = d.get("eos_token_id")
        eos = [] if eos is None else (list(eos) if isinstance(eos, (list, tuple)) else [eos])
        base = dict(hidden_size=d["hidden_size"], intermediate_size=d["intermediate_size"], vocab_size=d["vocab_size"],
                    num_hidden_layers=n, num_attention_heads=d["num_attention_heads"],
                    num_key_value_heads=d.get("num_key_value_heads") or d["num_attention_heads"],
                    rms_norm_eps=d["rms_norm_eps"], rope_scaling=rope, eos_token_id=eos, head_dim=d.get("head_dim"),
                    use_qk_norm=True, block_kind=kind)
        if kind == "olmo2":    # olmo2/config.rs:52-90
            return Config(**base, rope_theta=float(d.get("rope_theta") or 500000.0),
                          tie_word_embeddings=bool(d.get("tie_word_embeddings", False)),
                          max_seq_len=int(d.get("max_position_embeddings") or 4096), pre_reshape_qk_norm=True)
        if kind == "exaone4":  # exaone4/config.rs:62-110: every `global_layer_period`-th layer is global
            period = int(d.get("global_layer_period") or 4)
            return Config(**base, rope_theta=float(d.get("rope_theta") or 500000.0),
                          tie_word_embeddings=bool(d.get("tie_word_embeddings", False)),
                          max_seq_len=int(d.get("max_position_embeddings") or 131072),
                          sliding_window=int(d.get("sliding_window") or 4096),
                          global_layers=[(i + 1) % period == 0 for i in range(n)])
        # gemma3/config.rs:76-125: explicit schedule (true = global) or every `sliding_window_pattern`-th layer
        sched = d.get("sliding_window_attention_schedule") or []
        pattern = int(d.get("sliding_window_pattern") or 6)
        gl = [bool(sched[i]) if i < len(sched) else False for i in range(n)] if sched else [(i + 1) % pattern == 0 for i in range(n)]
        return Config(**base, rope_theta=float(d.get("rope_theta") or 10000.0), tie_word_embeddings=True,
                      max_seq_len=int(d.get("max_position_embeddings") or 131072),
                      sliding_window=int(d.get("sliding_window") or 1024), global_layers=gl, residual_rms_norm=True,
                      use_gelu_mlp=True, embed_scale=float(np.sqrt(np.float32(d["hidden_size"]))))

    @staticmethod
    def from_path(path: str) -> "Config":
        with open(path) as f:
            return Config.from_hf(json.load(f))

    def to_hf(self, arch: str = "LlamaForCausalLM") -> dict:
        d = dict(
            architectures=[arch], hidden_size=self.hidden_size, intermediate_size=self.intermediate_size,
            vocab_size=self.vocab_size, num_hidden_layers=self.num_hidden_layers,
            num_attention_heads=self.num_attention_heads, num_key_value_heads=self.num_key_value_heads,
            rms_norm_eps=self.rms_norm_eps, rope_theta=self.rope_theta,
            tie_word_embeddings=self.tie_word_embeddings, max_position_embeddings=self.max_seq_len,
        )
        if self.head_dim:
            d["head_dim"] = self.head_dim
        if self.sliding_window:
            d["sliding_window"] = self.sliding_window
        if self.partial_rotary_factor != 1.0:
            d["partial_rotary_factor"] = self.partial_rotary_factor
        if self.rope_scaling:
            d["rope_scaling"] = asdict(self.rope_scaling)
        # the keys the sibling architectures' config.rs read (models/{gemma3,exaone4}/config.rs)
        if self.block_kind == "gemma3":
            d["sliding_window_attention_schedule"] = [bool(g) for g in self.global_layers]
        elif self.block_kind == "exaone4":
            hits = [i + 1 for i, g in enumerate(self.global_layers) if g]
            period = hits[0] if hits else self.num_hidden_layers + 1
            if [(i + 1) % period == 0 for i in range(self.num_hidden_layers)] != [bool(g) for g in self.global_layers]:
                raise ValueError("EXAONE4 can only express an every-n-th global-layer schedule (global_layer_period)")
            d["global_layer_period"] = period
        return d

    def is_eos(self, tok: int) -> bool:
        # config.rs:6-19 EosTokenId::is_eos
        return tok in self.eos_token_id


class CConfig(ctypes.Structure):
    """``cake_b200_config`` from include/cake_b200.h (and the oracle's ora_config prefix)."""
    _fields_ = [
        ("hidden", ctypes.c_int), ("inter", ctypes.c_int), ("n_heads", ctypes.c_int),
        ("n_kv_heads", ctypes.c_int), ("head_dim", ctypes.c_int), ("n_layers", ctypes.c_int),
        ("vocab", ctypes.c_int), ("max_seq", ctypes.c_int),
        ("rms_eps", ctypes.c_float), ("rope_theta", ctypes.c_float), ("partial_rotary", ctypes.c_float),
        ("qkv_bias", ctypes.c_int), ("qk_norm", ctypes.c_int), ("tie_embeddings", ctypes.c_int),
        ("rope_llama3", ctypes.c_int),
        ("rope_factor", ctypes.c_float), ("rope_low", ctypes.c_float), ("rope_high", ctypes.c_float),
        ("rope_orig_max", ctypes.c_int),
        ("dtype", ctypes.c_int),
        ("sliding_window", ctypes.c_int), ("use_gelu_mlp", ctypes.c_int), ("embed_scale", ctypes.c_float),
        ("pre_reshape_qk_norm", ctypes.c_int),
    ]

    @staticmethod
    def from_config(c: Config, dtype: str = "bf16", max_seq: Optional[int] = None) -> "CConfig":
        rs = c.rope_scaling
        llama3 = bool(rs and rs.rope_type == "llama3" and rs.original_max_position_embeddings > 0)
        # cache.rs:173-205 limit = min(window, max_seq_len); a window that never bites is passed as 0
        win = int(c.sliding_window) if (c.sliding_window and c.sliding_window < (max_seq or c.max_seq_len)) else 0
        return CConfig(
            c.hidden_size, c.intermediate_size, c.num_attention_heads, c.num_key_value_heads, c.hd,
            c.num_hidden_layers, c.vocab_size, max_seq or c.max_seq_len,
            c.rms_norm_eps, c.rope_theta, c.partial_rotary_factor,
            int(c.use_qkv_bias), int(c.use_qk_norm), int(c.tie_word_embeddings),
            int(llama3),
            rs.factor if llama3 else 1.0, rs.low_freq_factor if llama3 else 1.0,
            rs.high_freq_factor if llama3 else 4.0, rs.original_max_position_embeddings if llama3 else 0,
            DTYPES[dtype],
            win, int(c.use_gelu_mlp), float(c.embed_scale or 0.0), int(c.pre_reshape_qk_norm),
        )


# Named shapes used by tests / bench.
def llama3_8b(max_seq: int = 8192) -> Config:
    return Config(4096, 14336, 128256, 32, 32, 8, rms_norm_eps=1e-5, rope_theta=500000.0, max_seq_len=max_seq,
                  eos_token_id=[128001, 128009])


def llama3_70b(max_seq: int = 8192) -> Config:
    return Config(8192, 28672, 128256, 80, 64, 8, rms_norm_eps=1e-5, rope_theta=500000.0, max_seq_len=max_seq,
                  eos_token_id=[128001, 128009])


def qwen3_0_6b(max_seq: int = 4096) -> Config:
    return Config(1024, 3072, 151936, 28, 16, 8, rms_norm_eps=1e-6, rope_theta=1000000.0, max_seq_len=max_seq,
                  head_dim=128, use_qk_norm=True, tie_word_embeddings=True, eos_token_id=[151645])


def reference_test_config(**kw) -> Config:
    """tests/unit_tests/helpers.rs:8-44 ``test_config()``."""
    base = dict(hidden_size=64, intermediate_size=128, vocab_size=256, num_hidden_layers=4,
                num_attention_heads=4, num_key_value_heads=2, rms_norm_eps=1e-6, rope_theta=10000.0,
                max_seq_len=64)
    base.update(kw)
    return Config(**base)
