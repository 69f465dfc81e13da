"""config.json -> block configuration for the dense architectures whose block is the path built here
(SURVEY.md §8f-4): per-architecture serde defaults and hard-wired flags of the reference's into_config
(llama3/config.rs:62-98, qwen2/config.rs:69-105, qwen3/config.rs:55-93, mistral/config.rs:56-93,
falcon3/config.rs:53-90), arch detection (common/config.rs:175-190, cake/mod.rs:81-109) — the Python mirror and
the compiled C++ host must agree."""
import json
import os
import subprocess

import pytest

from cake_b200.build import build_host
from cake_b200.config import CConfig, Config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUN = os.path.join(ROOT, "cake_b200", "host", "cake_run")
BASE = dict(hidden_size=64, intermediate_size=128, vocab_size=256, num_hidden_layers=2, num_attention_heads=4,
            rms_norm_eps=1e-6)

CASES = {
    # name: (config.json extras, expected resolved fields)
    "llama_defaults": (dict(architectures=["LlamaForCausalLM"]),
                       dict(rope_theta=500000.0, max_seq=4096, qkv_bias=0, qk_norm=0, head_dim=16, kv_heads=4)),
    "llama_ignores_head_dim": (dict(architectures=["LlamaForCausalLM"], head_dim=32), dict(head_dim=16)),
    "no_architectures_is_llama": (dict(), dict(rope_theta=500000.0, max_seq=4096, qkv_bias=0)),
    "unknown_string_is_llama": (dict(architectures=["SomethingElseLM"]), dict(rope_theta=500000.0, qk_norm=0)),
    "first_string_entry_wins": (dict(architectures=[42, "Qwen2ForCausalLM", "LlamaForCausalLM"]), dict(qkv_bias=1)),
    "qwen2": (dict(architectures=["Qwen2ForCausalLM"], num_key_value_heads=2, sliding_window=4096),
              dict(rope_theta=1000000.0, max_seq=32768, qkv_bias=1, qk_norm=0, head_dim=16, kv_heads=2)),
    "qwen3": (dict(architectures=["Qwen3ForCausalLM"], head_dim=32, tie_word_embeddings=True),
              dict(rope_theta=1000000.0, max_seq=40960, qkv_bias=0, qk_norm=1, head_dim=32, tie=1)),
    "mistral": (dict(architectures=["MistralForCausalLM"], head_dim=32, sliding_window=None, rope_theta=10000.0),
                dict(rope_theta=10000.0, max_seq=131072, head_dim=32, qkv_bias=0)),
    "falcon3": (dict(architectures=["FalconForCausalLM"], head_dim=32, num_key_value_heads=1, eos_token_id=[7, 9]),
                dict(rope_theta=500000.0, max_seq=131072, head_dim=32, kv_heads=1, n_eos=2)),
    "phi4_mini": (dict(architectures=["Phi3ForCausalLM"], partial_rotary_factor=0.75, head_dim=32, num_key_value_heads=2),
                  dict(rope_theta=1000000.0, max_seq=131072, head_dim=32, kv_heads=2, partial_rotary=0.75, fused=1, qkv_bias=0)),
    "phi4": (dict(architectures=["Phi4ForCausalLM"]), dict(partial_rotary=1.0, fused=1, head_dim=16)),
    "llama_ignores_partial_rotary": (dict(architectures=["LlamaForCausalLM"], partial_rotary_factor=0.5),
                                     dict(partial_rotary=1.0, fused=0)),
}


def _python_view(d: dict) -> dict:
    c = Config.from_hf(d)
    cc = CConfig.from_config(c, "bf16")
    return dict(rope_theta=float(cc.rope_theta), max_seq=cc.max_seq, qkv_bias=cc.qkv_bias, qk_norm=cc.qk_norm,
                head_dim=cc.head_dim, kv_heads=cc.n_kv_heads, tie=cc.tie_embeddings, n_eos=len(c.eos_token_id),
                partial_rotary=float(cc.partial_rotary), fused=int(c.fused_qkv_proj and c.fused_gate_up_proj))


def _cpp_kv(tmp_path, d: dict) -> dict:
    build_host()
    with open(tmp_path / "config.json", "w") as f:
        json.dump(d, f)
    r = subprocess.run([RUN, str(tmp_path), "--show-config"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return dict(p.split("=", 1) for p in r.stdout.split())


def _cpp_view(tmp_path, d: dict) -> dict:
    kv = _cpp_kv(tmp_path, d)
    return {k: (float(v) if k in ("rope_theta", "rms_eps", "partial_rotary", "embed_scale") else v if k in ("arch", "block_kind", "globals") else int(v))
            for k, v in kv.items()}


@pytest.mark.parametrize("name", list(CASES))
def test_python_and_cpp_resolve_the_same_block_config(tmp_path, name):
    extra, want = CASES[name]
    d = {**BASE, **extra}
    py, cpp = _python_view(d), _cpp_view(tmp_path, d)
    for k, v in want.items():
        assert py[k] == v, (name, "python", k, py[k], v)
        assert cpp[k] == v, (name, "c++", k, cpp[k], v)
    for k in py:
        assert py[k] == cpp[k], (name, k, py[k], cpp[k])


@pytest.mark.parametrize("arch", ["Qwen3MoeForCausalLM", "Qwen3_5ForConditionalGeneration", "LuxTTSForTextToSpeech"])
def test_other_block_types_are_refused_by_name(tmp_path, arch):
    """MoE / linear-attention / TTS blocks are outside the path in both hosts."""
    d = {**BASE, "architectures": [arch]}
    with pytest.raises(ValueError, match="outside the block-forward path"):
        Config.from_hf(d)
    build_host()
    with open(tmp_path / "config.json", "w") as f:
        json.dump(d, f)
    r = subprocess.run([RUN, str(tmp_path), "--show-config"], capture_output=True, text=True)
    assert r.returncode == 1 and "outside the block-forward path" in r.stderr


SIBLING_JSON = {
    "olmo2": {"architectures": ["Olmo2ForCausalLM"]},
    "olmo2_caps": {"architectures": ["OLMo2ForCausalLM"], "head_dim": 32, "tie_word_embeddings": True},
    "exaone4": {"architectures": ["ExaoneForCausalLM"], "num_hidden_layers": 8, "sliding_window": 16, "max_position_embeddings": 64},
    "exaone4_defaults": {"architectures": ["ExaoneForCausalLM"], "num_hidden_layers": 5, "global_layer_period": 2},
    "gemma3_pattern": {"architectures": ["Gemma3ForCausalLM"], "num_hidden_layers": 12, "sliding_window": 16, "max_position_embeddings": 64,
                       "hidden_size": 1152, "num_attention_heads": 4, "head_dim": 256},
    "gemma3_schedule": {"architectures": ["Gemma3ForCausalLM"], "num_hidden_layers": 3, "sliding_window_attention_schedule": [True, False, True]},
}


@pytest.mark.parametrize("name", list(SIBLING_JSON))
def test_python_and_cpp_resolve_the_same_sibling_block_config(tmp_path, name):
    """models/{olmo2,gemma3,exaone4}/config.rs into_config in both hosts: serde defaults, hard-wired flags, per-layer schedule."""
    d = {**BASE, **SIBLING_JSON[name]}
    c = Config.from_hf(d)
    cc = CConfig.from_config(c, "bf16")
    kv = _cpp_kv(tmp_path, d)
    var = [c.layer_variant(i) for i in range(c.num_hidden_layers)]
    local = [v["window"] for v in var if v["window"] > 0]
    want = dict(block_kind=c.block_kind, rope_theta=float(cc.rope_theta), max_seq=cc.max_seq, qk_norm=cc.qk_norm, tie=cc.tie_embeddings,
                head_dim=cc.head_dim, kv_heads=cc.n_kv_heads, pre_reshape_qk_norm=cc.pre_reshape_qk_norm, gelu=cc.use_gelu_mlp,
                residual_rms_norm=int(c.residual_rms_norm), sliding_window=cc.sliding_window if c.block_kind == "olmo2" else 0,
                layer_window=local[0] if local else 0,
                globals="".join("1" if g else "0" for g in c.global_layers) or "-")
    for k, v in want.items():
        got = kv[k]
        assert (float(got) == float(v)) if isinstance(v, float) else (str(got) == str(v)), (name, k, got, v)
    assert abs(float(kv["embed_scale"]) - float(cc.embed_scale)) < 1e-4


def test_sibling_block_structures_resolve_like_the_reference():
    """models/{olmo2,gemma3,exaone4}/config.rs into_config + block.rs load: flags, per-layer schedule, norm placement."""
    o = Config.from_hf({**BASE, "architectures": ["Olmo2ForCausalLM"]})
    assert (o.block_kind, o.use_qk_norm, o.pre_reshape_qk_norm, o.rope_theta, o.sliding_window) == ("olmo2", True, True, 500000.0, None)
    assert o.layer_variant(0) == dict(pre_norms=False, post_norms=True, window=-1, no_rope=False)
    assert CConfig.from_config(o, "bf16").pre_reshape_qk_norm == 1

    e = Config.from_hf({**BASE, "architectures": ["ExaoneForCausalLM"], "num_hidden_layers": 8, "sliding_window": 16,
                        "max_position_embeddings": 64})
    assert e.block_kind == "exaone4" and e.use_qk_norm and not e.pre_reshape_qk_norm
    assert e.global_layers == [False, False, False, True, False, False, False, True]       # exaone4/config.rs is_global_layer, period 4
    assert e.layer_variant(0) == dict(pre_norms=True, post_norms=False, window=16, no_rope=False)   # local: window + RoPE
    assert e.layer_variant(3) == dict(pre_norms=True, post_norms=False, window=0, no_rope=True)     # global: full context, no RoPE
    assert e.layer_variant(0, max_seq=16)["window"] == 0                                            # a window that can never bite

    g = Config.from_hf({**BASE, "architectures": ["Gemma3ForCausalLM"], "num_hidden_layers": 12, "sliding_window": 16,
                        "max_position_embeddings": 64, "hidden_size": 1152})
    assert g.block_kind == "gemma3" and g.residual_rms_norm and g.use_gelu_mlp and g.tie_word_embeddings and g.use_qk_norm
    assert abs(g.embed_scale - 1152 ** 0.5) < 1e-5 and g.rope_theta == 10000.0
    assert [i for i, x in enumerate(g.global_layers) if x] == [5, 11]                       # gemma3/config.rs test_gemma3_pattern
    assert g.layer_variant(0) == dict(pre_norms=True, post_norms=True, window=16, no_rope=True)     # local: window, NO RoPE
    assert g.layer_variant(5) == dict(pre_norms=True, post_norms=True, window=0, no_rope=False)     # global: full context + RoPE
    g2 = Config.from_hf({**BASE, "architectures": ["Gemma3ForCausalLM"], "num_hidden_layers": 3,
                         "sliding_window_attention_schedule": [True, False, True]})
    assert g2.global_layers == [True, False, True]


def test_active_sliding_window_reaches_the_library(tmp_path):
    """mistral/config.rs: sliding_window is carried into the generalized config; cache.rs:173-205 caps the visible KV at
    min(window, max_seq_len) — a window that can never bite is passed as 0 (full context)."""
    d = {**BASE, "architectures": ["MistralForCausalLM"], "sliding_window": 4096, "max_position_embeddings": 32768}
    c = Config.from_hf(d)
    assert c.sliding_window == 4096
    assert CConfig.from_config(c, "bf16").sliding_window == 4096
    assert CConfig.from_config(c, "bf16", max_seq=4096).sliding_window == 0
    build_host()
    with open(tmp_path / "config.json", "w") as f:
        json.dump(d, f)
    r = subprocess.run([RUN, str(tmp_path), "--show-config"], capture_output=True, text=True)
    assert r.returncode == 0 and "sliding_window=4096" in r.stdout
    r = subprocess.run([RUN, str(tmp_path), "--show-config", "--max-seq", "4096"], capture_output=True, text=True)
    assert r.returncode == 0 and "max_seq=4096" in r.stdout and "sliding_window=0" in r.stdout


@pytest.mark.parametrize("drop", ["hidden_size", "intermediate_size", "vocab_size", "num_hidden_layers", "num_attention_heads",
                                  "rms_norm_eps"])
def test_fields_without_a_serde_default_are_required(tmp_path, drop):
    """Every *Config struct of the reference declares these without `#[serde(default)]`: a config.json without one of
    them does not load (serde: "missing field")."""
    d = {**BASE, "architectures": ["LlamaForCausalLM"]}
    d.pop(drop)
    with pytest.raises(KeyError, match=drop):
        Config.from_hf(d)
    build_host()
    with open(tmp_path / "config.json", "w") as f:
        json.dump(d, f)
    r = subprocess.run([RUN, str(tmp_path), "--show-config"], capture_output=True, text=True)
    assert r.returncode == 1 and f"missing field `{drop}`" in r.stderr


@pytest.mark.parametrize("text", ["{", '{"hidden_size": }', "[1,2", "", '{"hidden_size": 1e999}'])
def test_malformed_config_is_an_error_not_a_crash(tmp_path, text):
    build_host()
    (tmp_path / "config.json").write_text(text)
    r = subprocess.run([RUN, str(tmp_path), "--show-config"], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("error:")
