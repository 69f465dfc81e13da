"""Synthetic HF-layout checkpoints (there are no real weights offline).

Tensor names and shapes are the ones the reference loads:
  transformer.rs:89-90   {layer}.input_layernorm.weight / post_attention_layernorm.weight  (H)
  attention.rs:96-129    {layer}.self_attn.{q,k,v,o}_proj.weight [out,in] (+ .bias, q_norm/k_norm.weight)
  mlp.rs:43-49           {layer}.mlp.{gate,up,down}_proj.weight
  text_model.rs:159-193  {prefix}.embed_tokens.weight, {prefix}.norm.weight, lm_head.weight
"""
from __future__ import annotations

import torch

from .config import Config

TORCH_DTYPES = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}


def layer_tensor_shapes(cfg: Config) -> dict:
    H, I = cfg.hidden_size, cfg.intermediate_size
    s = {
        "input_layernorm.weight": (H,),
        "post_attention_layernorm.weight": (H,),
        "self_attn.q_proj.weight": (cfg.size_q, H),
        "self_attn.k_proj.weight": (cfg.size_kv, H),
        "self_attn.v_proj.weight": (cfg.size_kv, H),
        "self_attn.o_proj.weight": (H, cfg.size_q),
        "mlp.gate_proj.weight": (I, H),
        "mlp.up_proj.weight": (I, H),
        "mlp.down_proj.weight": (H, I),
    }
    if cfg.use_qkv_bias:
        s["self_attn.q_proj.bias"] = (cfg.size_q,)
        s["self_attn.k_proj.bias"] = (cfg.size_kv,)
        s["self_attn.v_proj.bias"] = (cfg.size_kv,)
    if cfg.use_qk_norm:  # attention.rs:121-122: over head_dim, or over the whole projection (OLMo2)
        pre = getattr(cfg, "pre_reshape_qk_norm", False)
        s["self_attn.q_norm.weight"] = (cfg.size_q if pre else cfg.hd,)
        s["self_attn.k_norm.weight"] = (cfg.size_kv if pre else cfg.hd,)
    kind = getattr(cfg, "block_kind", "llama")
    if kind in ("olmo2", "exaone4_hf"):     # olmo2/block.rs:49-52: two post-norms, no input norms
        del s["input_layernorm.weight"]
        s["post_feedforward_layernorm.weight"] = (H,)
    elif kind == "gemma3":  # gemma3/block.rs:84-91: four norms
        s["pre_feedforward_layernorm.weight"] = (H,)
        s["post_feedforward_layernorm.weight"] = (H,)
    return s


def residual_deltas(sd: dict) -> dict:
    """Checkpoints of ``residual_rms_norm`` models (Gemma3) store norm weights as deltas around 0 (forward weight = 1 + w,
    config.rs:155-173): turn a synthetic checkpoint whose norm vectors sit around 1 into that convention."""
    return {k: ((v.float() - 1.0).to(v.dtype) if (k.endswith("norm.weight") or k.endswith("layernorm.weight")) else v)
            for k, v in sd.items()}


def _fill(shape, name: str, gen: torch.Generator, device, dtype, std: float):
    t = torch.randn(shape, generator=gen, device=device, dtype=torch.float32) * std
    if name.endswith("norm.weight") or name.endswith("layernorm.weight"):
        t = t + 1.0
    return t.to(dtype)


def make_layer(cfg: Config, i: int, dtype: str = "bf16", seed: int = 1234, device="cpu", std: float = 0.02) -> dict:
    """Weights of one block, keyed by full HF name.  Deterministic per (seed, layer, device type).  With
    ``cfg.fused_qkv_proj`` / ``cfg.fused_gate_up_proj`` (Phi-3/4 checkpoints) the same values are stored as
    ``self_attn.qkv_proj.weight = cat(q, k, v)`` / ``mlp.gate_up_proj.weight = cat(gate, up)``."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed * 1000003 + i)
    out = {}
    for short, shape in layer_tensor_shapes(cfg).items():
        out[short] = _fill(shape, short, gen, device, TORCH_DTYPES[dtype], std)
    if cfg.fused_qkv_proj:
        out["self_attn.qkv_proj.weight"] = torch.cat([out.pop(f"self_attn.{p}_proj.weight") for p in "qkv"], 0)
    if cfg.fused_gate_up_proj:
        out["mlp.gate_up_proj.weight"] = torch.cat([out.pop("mlp.gate_proj.weight"), out.pop("mlp.up_proj.weight")], 0)
    return {f"{cfg.layer_name(i)}.{short}": t for short, t in out.items()}


def make_head(cfg: Config, dtype: str = "bf16", seed: int = 1234, device="cpu", std: float = 0.02,
              peaked: bool = False) -> dict:
    """embed / final norm / lm_head.  ``peaked``: lm_head rows are scaled copies of the embedding rows,
    which gives large top-1/top-2 logit margins (used for greedy token-id parity)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed * 1000003 + 999983)
    td = TORCH_DTYPES[dtype]
    p = cfg.model_prefix
    out = {
        f"{p}.embed_tokens.weight": _fill((cfg.vocab_size, cfg.hidden_size), "embed", gen, device, td, std),
        f"{p}.norm.weight": _fill((cfg.hidden_size,), "norm.weight", gen, device, td, std),
    }
    if not cfg.tie_word_embeddings:
        if peaked:
            out["lm_head.weight"] = (out[f"{p}.embed_tokens.weight"].float() * 4.0).to(td)
        else:
            out["lm_head.weight"] = _fill((cfg.vocab_size, cfg.hidden_size), "lm_head", gen, device, td, std)
    return out


def make_checkpoint(cfg: Config, dtype: str = "bf16", seed: int = 1234, device="cpu", std: float = 0.02,
                    peaked: bool = False) -> dict:
    sd = make_head(cfg, dtype, seed, device, std, peaked)
    for i in range(cfg.num_hidden_layers):
        sd.update(make_layer(cfg, i, dtype, seed, device, std))
    return sd


class LazyCheckpoint(dict):
    """var_builder that materialises ONE block at a time on ``device`` — a 15 GB (8B) or 141 GB (70B) synthetic
    checkpoint never exists whole; ``Forwarder::load`` pulls a layer's tensors, the previous layer is dropped.
    Head tensors (embed / norm / lm_head) are created on first use and kept until ``drop_head()``.
    ``host_copy=True`` also keeps a CPU copy of everything handed out in ``self.host`` (what a checker needs to see
    the same weights).  Values are those of ``make_layer`` / ``make_head`` for (seed, device type)."""

    def __init__(self, cfg: Config, dtype: str = "bf16", seed: int = 1234, device="cuda", std: float = 0.02,
                 host_copy: bool = False, peaked: bool = False):
        super().__init__()
        self.cfg, self.dtype, self.seed, self.device, self.std, self.peaked = cfg, dtype, seed, device, std, peaked
        self.host = {} if host_copy else None

    def _ensure(self, k: str):
        if k in self:
            return
        if ".layers." in k:
            i = int(k.split(".layers.")[1].split(".")[0])
            self.drop_layers()
            new = make_layer(self.cfg, i, self.dtype, self.seed, self.device, self.std)
        elif k.startswith(self.cfg.model_prefix + ".") or k == "lm_head.weight":
            new = make_head(self.cfg, self.dtype, self.seed, self.device, self.std, self.peaked)
        else:
            return
        self.update(new)
        if self.host is not None:
            for kk, t in new.items():
                self.host[kk] = t.cpu()

    def get(self, k, d=None):
        self._ensure(k)
        return dict.get(self, k, d)

    def __getitem__(self, k):
        self._ensure(k)
        return dict.__getitem__(self, k)

    def drop_layers(self):
        for kk in [kk for kk in self if ".layers." in kk]:
            del self[kk]

    def drop_head(self):
        for kk in [kk for kk in self if ".layers." not in kk]:
            del self[kk]
