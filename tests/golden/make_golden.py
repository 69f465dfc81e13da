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
    "phi3_tiny": ("Phi3ForCausalLM", dict(partial_rotary_factor=0.5, rope_theta=1000000.0, fused_qkv_proj=True,
                                          fused_gate_up_proj=True)),
}
SEED, STD, N_IDS = 3, 0.1, 12


def case_config(name):
    arch, kw = CASES[name]
    return arch, reference_test_config(**kw)


def hf_logits(arch, cfg, sd, ids):
    from transformers import (LlamaConfig, LlamaForCausalLM, MistralConfig, MistralForCausalLM, Phi3Config, Phi3ForCausalLM,
                              Qwen2Config, Qwen2ForCausalLM, Qwen3Config, Qwen3ForCausalLM)
    d = cfg.to_hf(arch)
    d.pop("architectures")
    window = d.pop("sliding_window", None)
    if arch in ("LlamaForCausalLM", "FalconForCausalLM"):
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
        sd = make_checkpoint(cfg, "f32", seed=SEED, std=STD)
        ids = torch.randint(0, cfg.vocab_size, (1, N_IDS), generator=torch.Generator().manual_seed(11))
        lg = hf_logits(arch, cfg, sd, ids)
        np.savez_compressed(os.path.join(out, f"hf_{name}.npz"), ids=ids[0].numpy().astype(np.uint32),
                            logits=lg.astype(np.float32), seed=SEED, std=STD)
        print(name, lg.shape)


if __name__ == "__main__":
    main()
