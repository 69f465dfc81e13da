"""Generates tests/golden/hf_*.npz — run in the authoring container (needs `transformers`).

Each fixture holds a seeded synthetic checkpoint's *seed* (weights are regenerated from
cake_b200.synth, not stored), the input ids, and the fp32 logits of HuggingFace transformers
(LlamaForCausalLM / Qwen3ForCausalLM / Qwen2ForCausalLM / MistralForCausalLM — an independent implementation
of the architectures the reference's llama3/, qwen3/, qwen2/, mistral/ and falcon3/ wrappers load) for every
position.  tests/test_oracle_golden.py
checks the oracle (f32 mode) against them; the GPU parity tests then compare against the oracle.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cake_b200.config import RopeScaling, reference_test_config  # noqa: E402
from cake_b200.synth import make_checkpoint  # noqa: E402

CASES = {
    "llama_tiny": ("LlamaForCausalLM", dict()),
    "llama3_rope_scaled": ("LlamaForCausalLM", dict(rope_theta=500000.0,
                           rope_scaling=RopeScaling(8.0, 1.0, 4.0, 16, "llama3"))),
    "qwen3_tiny": ("Qwen3ForCausalLM", dict(head_dim=32, use_qk_norm=True, tie_word_embeddings=True)),
    # sibling dense architectures whose block is the same path
    "qwen2_tiny": ("Qwen2ForCausalLM", dict(use_qkv_bias=True, rope_theta=1000000.0, tie_word_embeddings=True)),
    "mistral_tiny": ("MistralForCausalLM", dict(head_dim=32, rope_theta=1000000.0)),
    # Falcon3 checkpoints are Llama blocks with an explicit head_dim (falcon3/config.rs:53-90); HF runs them as Llama
    "falcon3_tiny": ("FalconForCausalLM", dict(head_dim=24, rope_theta=500000.0, num_key_value_heads=1)),
    # Phi-3/4: pre-fused qkv_proj / gate_up_proj tensors, rotary on the first half of each head only (phi4/config.rs:63-100)
    # an ACTIVE sliding window (mistral/config.rs + cache.rs:173-205): HF masks per query (last 4 keys); cake trims the
    # cache per call — identical once the prompt is no longer than the window and the rest is decoded token by token,
    # which is how tests/test_oracle_golden.py replays this fixture
    "mistral_window": ("MistralForCausalLM", dict(head_dim=32, rope_theta=1000000.0, sliding_window=4)),
    # use_gelu_mlp (mlp.rs:25-26; Gemma-style gelu_tanh gate) through HF's Llama with hidden_act=gelu_pytorch_tanh
    "llama_gelu_tiny": ("LlamaForCausalLM", dict(use_gelu_mlp=True)),
    # the sibling block structures (models/{olmo2,gemma3}/block.rs):
    # OLMo2: no pre-norms, post-attention / post-feedforward norms, QK-norm over the whole projection before the reshape
    "olmo2_tiny": ("Olmo2ForCausalLM", dict(use_qk_norm=True, pre_reshape_qk_norm=True, block_kind="olmo2")),
    # Gemma3: sandwich norms with (1 + w) weights, gelu_tanh MLP, scaled embeddings, per-head QK-norm, tied head.  Every layer
    # is GLOBAL here: the reference's local layers run without RoPE (gemma3/block.rs:62-66) while HF's rotate with a second
    # base frequency, so only the global-layer arithmetic has an independent implementation to be pinned to
    "gemma3_global_tiny": ("Gemma3ForCausalLM", dict(head_dim=32, use_qk_norm=True, block_kind="gemma3", global_layers=[True] * 4,
                                                     sliding_window=16, residual_rms_norm=True, use_gelu_mlp=True,
                                                     tie_word_embeddings=True, embed_scale=8.0)),
    # the hybrid schedule — local layers: sliding window + RoPE, global layers: full context, NO RoPE — on HF's EXAONE 4.0,
    # whose block is post-norm (block_kind "exaone4_hf", test-only; the reference's own EXAONE4 block is pre-norm,
    # exaone4/block.rs:96-110, i.e. the Llama structure already pinned above + this per-layer attention mode)
    "exaone4_hf_tiny": ("Exaone4ForCausalLM", dict(head_dim=32, use_qk_norm=True, block_kind="exaone4_hf", sliding_window=5,
                                                   global_layers=[False, True, False, True])),
    "phi3_tiny": ("Phi3ForCausalLM", dict(partial_rotary_factor=0.5, rope_theta=1000000.0, fused_qkv_proj=True,
                                          fused_gate_up_proj=True)),
}
SEED, STD, N_IDS = 3, 0.1, 12


def case_config(name):
    arch, kw = CASES[name]
    return arch, reference_test_config(**kw)


def case_checkpoint(cfg, seed, std):
    """The synthetic checkpoint of a case, in the convention its architecture stores norm weights in."""
    from cake_b200.synth import residual_deltas
    sd = make_checkpoint(cfg, "f32", seed=seed, std=std)
    return residual_deltas(sd) if cfg.residual_rms_norm else sd


def hf_sibling(arch, cfg):
    import transformers as tf
    common = dict(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size, vocab_size=cfg.vocab_size,
                  num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                  num_key_value_heads=cfg.num_key_value_heads, rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta,
                  max_position_embeddings=cfg.max_seq_len, tie_word_embeddings=cfg.tie_word_embeddings)
    types = ["full_attention" if g else "sliding_attention" for g in cfg.global_layers]
    if arch == "Olmo2ForCausalLM":
        return tf.Olmo2ForCausalLM(tf.Olmo2Config(**common, attention_bias=False))
    if arch == "Exaone4ForCausalLM":
        return tf.Exaone4ForCausalLM(tf.Exaone4Config(**common, head_dim=cfg.hd, sliding_window=cfg.sliding_window, layer_types=types,
                                                       attn_implementation="eager"))
    return tf.Gemma3ForCausalLM(tf.Gemma3TextConfig(**common, head_dim=cfg.hd, sliding_window=cfg.sliding_window, layer_types=types,
                                                    query_pre_attn_scalar=cfg.hd, hidden_activation="gelu_pytorch_tanh",
                                                    attn_implementation="eager"))


def hf_logits(arch, cfg, sd, ids):
    from transformers import (LlamaConfig, LlamaForCausalLM, MistralConfig, MistralForCausalLM, Phi3Config, Phi3ForCausalLM,
                              Qwen2Config, Qwen2ForCausalLM, Qwen3Config, Qwen3ForCausalLM)
    d = cfg.to_hf(arch)
    d.pop("architectures")
    window = d.pop("sliding_window", None)
    if arch in ("Olmo2ForCausalLM", "Exaone4ForCausalLM", "Gemma3ForCausalLM"):
        m = hf_sibling(arch, cfg)
        if arch == "Gemma3ForCausalLM":
            assert abs(float(m.model.embed_tokens.embed_scale) - cfg.embed_scale) < 1e-6  # HF: sqrt(hidden) = 8 at hidden 64
    elif arch in ("LlamaForCausalLM", "FalconForCausalLM"):
        m = LlamaForCausalLM(LlamaConfig(**d, attention_bias=False, mlp_bias=False,
                                         hidden_act="gelu_pytorch_tanh" if cfg.use_gelu_mlp else "silu"))
    elif arch == "Qwen2ForCausalLM":
        m = Qwen2ForCausalLM(Qwen2Config(**d, use_sliding_window=False))
    elif arch == "MistralForCausalLM":
        m = MistralForCausalLM(MistralConfig(**d, sliding_window=window, attn_implementation="eager"))
    elif arch == "Phi3ForCausalLM":
        m = Phi3ForCausalLM(Phi3Config(**d, pad_token_id=0, original_max_position_embeddings=d["max_position_embeddings"]))
    else:
        m = Qwen3ForCausalLM(Qwen3Config(**d))
    sd = {k: v.float() for k, v in sd.items()}
    if cfg.tie_word_embeddings:
        sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        return m.eval()(ids).logits[0].numpy()


def main():
    out = os.path.dirname(os.path.abspath(__file__))
    for name in (sys.argv[1:] or CASES):
        arch, cfg = case_config(name)
        sd = case_checkpoint(cfg, SEED, STD)
        ids = torch.randint(0, cfg.vocab_size, (1, N_IDS), generator=torch.Generator().manual_seed(11))
        lg = hf_logits(arch, cfg, sd, ids)
        np.savez_compressed(os.path.join(out, f"hf_{name}.npz"), ids=ids[0].numpy().astype(np.uint32),
                            logits=lg.astype(np.float32), seed=SEED, std=STD)
        print(name, lg.shape)


if __name__ == "__main__":
    main()
