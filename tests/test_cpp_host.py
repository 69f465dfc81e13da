"""The compiled host side (cake_b200/host/cake_host.hpp + cake_run): C++ mirror of cake's Forwarder /
TextModelBase / Master over the C ABI, loading HF safetensors by the reference's tensor names."""
import json
import os
import subprocess

import numpy as np
import pytest
import torch

from cake_b200.build import build
from tests.util import checkpoint, medium_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUN = os.path.join(ROOT, "cake_b200", "host", "cake_run")


def _write_model(tmp_path, cfg, sd, sharded=False):
    from safetensors.torch import save_file
    with open(tmp_path / "config.json", "w") as f:
        json.dump({**cfg.to_hf("LlamaForCausalLM"), "eos_token_id": [1023]}, f)
    sd = {k: v.contiguous() for k, v in sd.items()}
    if not sharded:
        save_file(sd, str(tmp_path / "model.safetensors"))
    else:  # utils/mod.rs:333-384: model.safetensors.index.json + shards
        keys = sorted(sd)
        half = len(keys) // 2
        parts = {"model-00001-of-00002.safetensors": keys[:half], "model-00002-of-00002.safetensors": keys[half:]}
        wm = {}
        for fn, ks in parts.items():
            save_file({k: sd[k] for k in ks}, str(tmp_path / fn))
            wm.update({k: fn for k in ks})
        with open(tmp_path / "model.safetensors.index.json", "w") as f:
            json.dump({"metadata": {}, "weight_map": wm}, f)


def test_cake_run_builds_and_fails_cleanly_without_gpu(tmp_path):
    build()
    assert os.path.exists(RUN)
    cfg = medium_config(num_hidden_layers=1)
    _write_model(tmp_path, cfg, checkpoint(cfg, "bf16", seed=1))
    r = subprocess.run([RUN, str(tmp_path), "--prompt-ids", "1,2,3", "-n", "2"], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
    else:  # config + safetensors parsed, then the library refuses: no CPU fallback
        assert r.returncode == 1 and "ctx_create" in r.stderr, r.stderr
    r = subprocess.run([RUN, str(tmp_path / "nope"), "--prompt-ids", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "can't read" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("sharded", [False, True])
def test_cake_run_tokens_equal_python_master(tmp_path, sharded):
    from cake_b200.model import Context, Master, TextModelBase
    build()
    cfg = medium_config()
    sd = checkpoint(cfg, "bf16", seed=61, peaked=True)
    _write_model(tmp_path, cfg, sd, sharded)
    prompt = np.random.default_rng(2).integers(0, cfg.vocab_size - 1, 10).tolist()
    r = subprocess.run([RUN, str(tmp_path), "--prompt-ids", ",".join(map(str, prompt)), "-n", "12"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    toks = [int(t) for t in r.stdout.splitlines()[0].split(":")[1].split()]
    ctx = Context(cfg, sd, "bf16", device=0)
    ref = Master(TextModelBase.load(ctx)).generate_text(prompt, 12)["tokens"]
    ctx.close()
    assert toks == ref
    from oracle import oracle as O   # and both equal the oracle's greedy tokens
    assert toks == list(O.OracleModel(cfg, sd, "bf16", max_seq=cfg.max_seq_len).generate(prompt, 12)[0])


@pytest.mark.parametrize("damage", ["truncate", "offsets", "header_len", "not_json"])
def test_damaged_safetensors_is_an_error_not_a_crash(tmp_path, damage):
    """The mmapped reader validates header length, JSON, and every tensor's offsets against its shape and the file size."""
    import struct
    build()
    cfg = medium_config(num_hidden_layers=1)
    _write_model(tmp_path, cfg, checkpoint(cfg, "bf16", seed=1))
    p = tmp_path / "model.safetensors"
    raw = p.read_bytes()
    (n,) = struct.unpack("<Q", raw[:8])
    if damage == "truncate":
        p.write_bytes(raw[: len(raw) // 2])
    elif damage == "offsets":
        hdr = json.loads(raw[8:8 + n])
        k = next(k for k in hdr if k != "__metadata__")
        hdr[k]["data_offsets"][1] += 2
        blob = json.dumps(hdr).encode()
        p.write_bytes(struct.pack("<Q", len(blob)) + blob + raw[8 + n:])
    elif damage == "header_len":
        p.write_bytes(struct.pack("<Q", 1 << 40) + raw[8:])
    else:
        p.write_bytes(struct.pack("<Q", 16) + b"this is not json" + raw[8 + n:])
    r = subprocess.run([RUN, str(tmp_path), "--prompt-ids", "1,2", "-n", "1"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and r.stderr.startswith("error:") and "ctx_create" not in r.stderr, r.stderr


def _fnv(b: bytes) -> str:
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


@pytest.mark.parametrize("flavour", ["llama", "qwen2_bias", "qwen3_qknorm_tied", "phi_fused", "sharded_worker"])
def test_compiled_host_hands_the_library_the_right_bytes(tmp_path, flavour):
    """cake_run / cake_worker against a recording stand-in for libcake_b200.so (tests/fake_b200): the compiled path
    config.json -> mmapped VarBuilder -> TextModelBase::load / Transformer::load must give ctx_create the resolved
    config and head_load / block_load pointers to exactly the checkpoint's bytes, in header order, null where unused —
    for plain, biased, QK-norm + tied, pre-fused (Phi) checkpoints and for a worker's shard subset.  No GPU."""
    import torch
    from cake_b200.loader import save_checkpoint
    from tests.fake_b200.make_fake import build as build_fake
    build()
    fake_dir = tmp_path / "fake"
    fake_dir.mkdir()
    build_fake(str(fake_dir))
    kw = dict(num_hidden_layers=3, hidden_size=128, intermediate_size=256, vocab_size=256, num_attention_heads=4,
              num_key_value_heads=2, head_dim=32)
    arch, extra = {"llama": ("LlamaForCausalLM", {}),
                   "qwen2_bias": ("Qwen2ForCausalLM", dict(use_qkv_bias=True)),
                   "qwen3_qknorm_tied": ("Qwen3ForCausalLM", dict(use_qk_norm=True, tie_word_embeddings=True)),
                   "phi_fused": ("Phi3ForCausalLM", dict(fused_qkv_proj=True, fused_gate_up_proj=True, partial_rotary_factor=0.75)),
                   "sharded_worker": ("LlamaForCausalLM", {})}[flavour]
    cfg = medium_config(**kw, **extra)
    plain = medium_config(**kw, **{k: v for k, v in extra.items() if not k.startswith("fused")})
    sd, sd_plain = checkpoint(cfg, "bf16", seed=9), checkpoint(plain, "bf16", seed=9)
    model = tmp_path / "model"
    save_checkpoint(str(model), cfg, sd, arch=arch, shard_bytes=200_000)
    log = tmp_path / "log.txt"
    env = {**os.environ, "LD_LIBRARY_PATH": str(fake_dir), "FAKE_B200_LOG": str(log)}
    if flavour == "sharded_worker":
        worker = os.path.join(ROOT, "cake_b200", "host", "cake_worker")
        p = subprocess.Popen([worker, str(model), "--layers", "model.layers.1-2", "--address", "127.0.0.1:0", "--max-seq", "64"],
                             stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
        assert p.stdout.readline().startswith("listening on "), p.stderr.read()
        p.kill()
        p.wait(5)
        layers = [1, 2]
    else:
        r = subprocess.run([RUN, str(model), "--prompt-ids", "1,2", "-n", "1", "--max-seq", "64"], capture_output=True, text=True, env=env, timeout=60)
        assert r.returncode == 1 and "fake libcake_b200" in r.stderr, r.stderr   # loads everything, then the stand-in refuses to compute
        layers = [0, 1, 2]
    lines = log.read_text().split("\n")
    ctx = [ln.split() for ln in lines if ln.startswith("ctx ")][0]
    want_ctx = [0, cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.hd, cfg.num_hidden_layers,
                cfg.vocab_size, 64]
    assert [int(x) for x in ctx[1:10]] == want_ctx
    assert float(ctx[12]) == cfg.partial_rotary_factor and [int(x) for x in ctx[13:17]] == [int(cfg.use_qkv_bias), int(cfg.use_qk_norm), int(cfg.tie_word_embeddings), 0]

    def h(name):
        t = sd_plain.get(name)
        return "0" * 16 if t is None else _fnv(t.contiguous().view(torch.uint8).numpy().tobytes())

    if flavour != "sharded_worker":
        head = [ln.split() for ln in lines if ln.startswith("head ")][0]
        assert head[1:] == [h("model.embed_tokens.weight"), h("model.norm.weight"), "0" * 16 if cfg.tie_word_embeddings else h("lm_head.weight")]
    order = ["self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight", "self_attn.o_proj.weight",
             "mlp.gate_proj.weight", "mlp.up_proj.weight", "mlp.down_proj.weight", "input_layernorm.weight",
             "post_attention_layernorm.weight", "self_attn.q_proj.bias", "self_attn.k_proj.bias", "self_attn.v_proj.bias",
             "self_attn.q_norm.weight", "self_attn.k_norm.weight"]
    blocks = {int(ln.split()[1]): ln.split()[2:] for ln in lines if ln.startswith("block ")}
    assert sorted(blocks) == layers
    for i in layers:
        assert blocks[i] == [h(f"model.layers.{i}.{s}") for s in order], f"layer {i}"


def _emulation(tmp_path):
    from tests.fake_b200.make_fake import build as build_fake
    build()
    d = tmp_path / "emu"
    d.mkdir(exist_ok=True)
    build_fake(str(d), oracle=True)
    return {**os.environ, "LD_LIBRARY_PATH": str(d)}


@pytest.mark.parametrize("flavour,penalty", [("llama", 1.0), ("llama", 1.3), ("qwen3_tied", 1.0), ("phi_fused", 1.0)])
def test_cake_run_host_loop_reproduces_the_oracle_on_the_cpu(tmp_path, flavour, penalty):
    """The compiled TextModelBase / Master loop (prefill through the block walk, then the host-stepped decode loop, or
    the logits + repeat-penalty path) over an oracle-backed emulation of the C ABI (tests/fake_b200): its token ids must
    be the oracle's own — i.e. positions, cache handling, sampling rule and tok/s bookkeeping are driven correctly."""
    from cake_b200.loader import save_checkpoint
    from oracle import oracle as O
    env = _emulation(tmp_path)
    kw = dict(num_hidden_layers=3, hidden_size=128, intermediate_size=256, vocab_size=256, num_attention_heads=4,
              num_key_value_heads=2, head_dim=32)
    arch, extra = {"llama": ("LlamaForCausalLM", {}),
                   "qwen3_tied": ("Qwen3ForCausalLM", dict(use_qk_norm=True, tie_word_embeddings=True)),
                   "phi_fused": ("Phi3ForCausalLM", dict(fused_qkv_proj=True, fused_gate_up_proj=True, partial_rotary_factor=0.5))}[flavour]
    cfg = medium_config(**kw, **extra)
    sd = checkpoint(cfg, "bf16", seed=13, peaked=not cfg.tie_word_embeddings)
    model = tmp_path / "model"
    save_checkpoint(str(model), cfg, sd, arch=arch, shard_bytes=300_000)
    prompt = np.random.default_rng(3).integers(0, cfg.vocab_size - 1, 9).tolist()
    n = 12
    r = subprocess.run([RUN, str(model), "--prompt-ids", ",".join(map(str, prompt)), "-n", str(n), "--max-seq", "64",
                        "--repeat-penalty", str(penalty)], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stderr
    got = [int(t) for t in r.stdout.splitlines()[0].split(":")[1].split()]
    om = O.OracleModel(cfg, sd, "bf16", max_seq=64)
    cache, ids, pos, want = om.new_cache(), list(prompt), 0, []
    for _ in range(n):   # text_model.rs:397-495 with the repeat-penalty rule of :435-452 (generated tokens only)
        lg = om.forward(ids, pos, cache)
        pos += len(ids)
        if penalty != 1.0:
            lg = O.repeat_penalty(O.round_to(lg, "bf16"), penalty, want[-128:], "bf16")
        want.append(O.argmax(lg))
        ids = [want[-1]]
    assert got == want
    assert "tok/s:" in r.stdout


@pytest.mark.parametrize("ckpt_dtype,model_dtype", [("f32", "bf16"), ("f16", "bf16"), ("bf16", "f16")])
def test_cake_run_converts_checkpoints_of_another_dtype_on_load(tmp_path, ckpt_dtype, model_dtype):
    """The reference's VarBuilder casts every tensor to the run's --dtype on load (utils/mod.rs:250-267 + candle's
    VarBuilder::get): an F32 / F16 checkpoint under --dtype bf16 (and BF16 under f16) must behave like the checkpoint
    rounded to the model dtype, as the Python host's `.to(dtype)` does."""
    import torch
    from cake_b200.loader import save_checkpoint
    from cake_b200.synth import TORCH_DTYPES
    from oracle import oracle as O
    env = _emulation(tmp_path)
    cfg = medium_config(num_hidden_layers=3, hidden_size=128, intermediate_size=256, vocab_size=256, num_attention_heads=4,
                        num_key_value_heads=2, head_dim=32, use_qk_norm=True)
    sd = checkpoint(cfg, ckpt_dtype, seed=29, peaked=True)
    model = tmp_path / "model"
    save_checkpoint(str(model), cfg, sd, arch="Qwen3ForCausalLM", shard_bytes=400_000)
    prompt = [5, 17, 200, 3, 77]
    r = subprocess.run([RUN, str(model), "--prompt-ids", ",".join(map(str, prompt)), "-n", "10", "--max-seq", "64", "--dtype", model_dtype],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stderr
    got = [int(t) for t in r.stdout.splitlines()[0].split(":")[1].split()]
    sd_d = {k: v.to(TORCH_DTYPES[model_dtype]) for k, v in sd.items()}
    om = O.OracleModel(cfg, sd_d, model_dtype, max_seq=64)
    assert got == list(om.generate(prompt, 10)[0])


SIBLING_CKPT = {
    "olmo2": ("Olmo2ForCausalLM", dict(block_kind="olmo2", use_qk_norm=True, pre_reshape_qk_norm=True)),
    "gemma3": ("Gemma3ForCausalLM", dict(block_kind="gemma3", use_qk_norm=True, residual_rms_norm=True, use_gelu_mlp=True,
                                         tie_word_embeddings=True, embed_scale=float(np.sqrt(np.float32(128))), sliding_window=5,
                                         global_layers=[False, True, False, True], rope_theta=10000.0)),
    "exaone4": ("ExaoneForCausalLM", dict(block_kind="exaone4", use_qk_norm=True, sliding_window=5, global_layers=[False, True, False, True])),
}


@pytest.mark.parametrize("kind,dtype", [("olmo2", "bf16"), ("gemma3", "bf16"), ("gemma3", "f16"), ("exaone4", "bf16")])
def test_cake_run_steps_the_sibling_block_structures_on_the_cpu(tmp_path, kind, dtype):
    """models/{olmo2,gemma3,exaone4}: config.json -> per-layer variants, the architectures' own tensor names, (1 + w) norm
    weights computed at load in the model dtype (config.rs:155-173), embed_scale, and host-stepped decode (no decode graph
    for these blocks) — the compiled host over the oracle-backed emulation must emit the oracle's tokens."""
    from cake_b200.loader import save_checkpoint
    from cake_b200.synth import residual_deltas
    from oracle import oracle as O
    env = _emulation(tmp_path)
    arch, extra = SIBLING_CKPT[kind]
    cfg = medium_config(num_hidden_layers=4, hidden_size=128, intermediate_size=256, vocab_size=256, num_attention_heads=4,
                        num_key_value_heads=2, head_dim=32, **extra)
    sd = checkpoint(cfg, dtype, seed=17, peaked=not cfg.tie_word_embeddings)
    if cfg.residual_rms_norm:
        sd = residual_deltas(sd)
    model = tmp_path / "model"
    save_checkpoint(str(model), cfg, sd, arch=arch, shard_bytes=300_000)
    prompt = np.random.default_rng(5).integers(0, cfg.vocab_size - 1, 4).tolist()   # fits the local window of 5
    n = 12
    r = subprocess.run([RUN, str(model), "--prompt-ids", ",".join(map(str, prompt)), "-n", str(n), "--max-seq", "64", "--dtype", dtype],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stderr
    got = [int(t) for t in r.stdout.splitlines()[0].split(":")[1].split()]
    om = O.OracleModel(cfg, sd, dtype, max_seq=64)
    assert got == list(om.generate(prompt, n)[0])


def test_cake_worker_sessions_and_errors_on_the_cpu(tmp_path):
    """cake_worker's B200Backend (one KV cache per connection, forwards under a lock, errors reported per request) over
    the oracle-backed emulation: activations must equal the oracle's block outputs bit for bit, a second connection
    starts from an empty cache, an out-of-order position is a WorkerError and the session stays usable."""
    from cake_b200.loader import save_checkpoint
    from cake_b200.wire import RawTensor, WireClient
    from oracle import oracle as O
    from tests.util import bits_to_f32, f32_to_bits, rand_x
    env = _emulation(tmp_path)
    cfg = medium_config(num_hidden_layers=4, hidden_size=128, intermediate_size=256, vocab_size=256, num_attention_heads=4,
                        num_key_value_heads=2, head_dim=32)
    sd = checkpoint(cfg, "bf16", seed=19)
    model = tmp_path / "model"
    save_checkpoint(str(model), cfg, sd, shard_bytes=200_000)
    worker = os.path.join(ROOT, "cake_b200", "host", "cake_worker")
    p = subprocess.Popen([worker, str(model), "--layers", "model.layers.2-3", "--address", "127.0.0.1:0", "--max-seq", "32", "--connections", "2"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    try:
        line = p.stdout.readline().strip()
        assert line.startswith("listening on "), p.stderr.read()
        addr = line[len("listening on "):]
        om = O.OracleModel(cfg, sd, "bf16", max_seq=32)
        x = rand_x((1, 6, cfg.hidden_size), "bf16", seed=2)
        raw = lambda t: RawTensor.from_numpy_bits(f32_to_bits(t.float().numpy(), "bf16"), "bf16")  # noqa: E731
        for conn in range(2):   # second connection: own, empty cache -> the same prefill at position 0 is accepted again
            c = WireClient(addr, cfg.layer_name(2), timeout=30)
            assert c.info.dtype == "BF16" and c.info.device == "cuda"
            oc = om.new_cache()
            batch = [(cfg.layer_name(i), 0, i) for i in (2, 3)]
            y = c.forward_batch(raw(x[:, :5]), batch)
            ref = om.block_forward(3, om.block_forward(2, x[0, :5].float().numpy(), 0, oc), 0, oc)
            assert y.shape == [1, 5, cfg.hidden_size] and np.array_equal(bits_to_f32(y.to_numpy_bits(), "bf16")[0], ref)
            with pytest.raises(RuntimeError, match=r"forward pass failed for layer model.layers.2 \(block_idx=2\).*cache length"):
                c.forward_batch(raw(x[:, 5:6]), [(cfg.layer_name(2), 9, 2)])
            with pytest.raises(RuntimeError, match="could not find layer model.layers.0"):
                c.forward_batch(raw(x[:, 5:6]), [(cfg.layer_name(0), 5, 0)])
            y = c.forward_batch(raw(x[:, 5:6]), [(cfg.layer_name(i), 5, i) for i in (2, 3)])    # the session is intact
            ref = om.block_forward(3, om.block_forward(2, x[0, 5:6].float().numpy(), 5, oc), 5, oc)
            assert np.array_equal(bits_to_f32(y.to_numpy_bits(), "bf16")[0], ref)
            c.goodbye()                                                                           # clears this session's cache
            y0 = c.forward_mut(raw(x[:, :1]), 0, 2)
            oc.clear()
            assert np.array_equal(bits_to_f32(y0.to_numpy_bits(), "bf16")[0], om.block_forward(2, x[0, :1].float().numpy(), 0, oc))
            c.close()
        assert p.wait(20) == 0
    finally:
        if p.poll() is None:
            p.kill()
            p.wait(5)
