"""Weight loading: cake_b200/loader.py against the reference's loading rules
(utils/mod.rs:160-384, cake/mod.rs:335-357) — checked with the `safetensors` library as an independent
writer/reader, with the oracle (tokens from a checkpoint on disk == tokens from the in-memory state dict),
and on the GPU through Context/TextModelBase."""
import json
import os
import struct

import numpy as np
import pytest
import torch

from cake_b200.config import Config
from cake_b200.loader import (INDEX_NAME, CheckpointError, SafetensorsFile, VarBuilder, detect_model_prefix,
                              open_model, save_checkpoint)
from cake_b200.parallel import box_topology
from tests.util import checkpoint, medium_config


def _bits(t: torch.Tensor) -> np.ndarray:
    return t.contiguous().view(torch.uint8).numpy()


def test_reads_what_the_safetensors_library_writes(tmp_path):
    from safetensors.torch import save_file
    g = torch.Generator().manual_seed(3)
    sd = {"a.weight": torch.randn(5, 7, generator=g).to(torch.bfloat16), "b.bias": torch.randn(9, generator=g).half(),
          "c": torch.randn(2, 3, 4, generator=g), "ids": torch.arange(6, dtype=torch.int64).view(2, 3),
          "empty": torch.empty(0, 4, dtype=torch.bfloat16)}
    p = str(tmp_path / "model.safetensors")
    save_file(sd, p, metadata={"format": "pt"})
    f = SafetensorsFile(p)
    assert sorted(f.names()) == sorted(sd) and f.metadata == {"format": "pt"}
    for k, v in sd.items():
        got = f.get(k)
        assert got.dtype == v.dtype and tuple(got.shape) == tuple(v.shape)
        assert np.array_equal(_bits(got), _bits(v)), k
    # views alias the mapping (no copy): two reads of one tensor share storage
    assert f.get("c").data_ptr() == f.get("c").data_ptr()


def test_library_reads_what_save_checkpoint_writes(tmp_path):
    from safetensors import safe_open
    cfg = medium_config(num_hidden_layers=2)
    sd = checkpoint(cfg, "bf16", seed=5)
    files = save_checkpoint(str(tmp_path), cfg, sd, shard_bytes=600_000)
    assert len(files) > 1 and os.path.exists(tmp_path / INDEX_NAME)
    wm = json.load(open(tmp_path / INDEX_NAME))["weight_map"]
    assert set(wm) == set(sd)
    for name, fn in wm.items():
        with safe_open(str(tmp_path / fn), framework="pt") as f:
            assert np.array_equal(_bits(f.get_tensor(name)), _bits(sd[name])), name
    assert Config.from_path(str(tmp_path / "config.json")).hidden_size == cfg.hidden_size


def test_unaligned_header_is_copied_not_misread(tmp_path):
    t = torch.arange(10, dtype=torch.float32)
    header = json.dumps({"x": {"dtype": "F32", "shape": [10], "data_offsets": [0, 40]}}).encode()
    if len(header) % 4 == 0:
        header += b" "
    with open(tmp_path / "u.safetensors", "wb") as f:
        f.write(struct.pack("<Q", len(header)) + header + t.numpy().tobytes())
    assert torch.equal(SafetensorsFile(str(tmp_path / "u.safetensors")).get("x"), t)


@pytest.mark.parametrize("blob,needle", [
    (b"\x01\x02", "shorter than"),
    (struct.pack("<Q", 1 << 40) + b"{}", "does not fit"),
    (struct.pack("<Q", 4) + b"nope", "not JSON"),
    (struct.pack("<Q", 2) + b"[]", "not a JSON object"),
])
def test_malformed_files_raise_named_errors(tmp_path, blob, needle):
    p = tmp_path / "bad.safetensors"
    p.write_bytes(blob)
    with pytest.raises(CheckpointError, match=needle) as e:
        SafetensorsFile(str(p))
    assert "bad.safetensors" in str(e.value)


def test_bad_offsets_and_missing_files(tmp_path):
    header = json.dumps({"x": {"dtype": "F32", "shape": [10], "data_offsets": [0, 44]}}).encode()
    (tmp_path / "o.safetensors").write_bytes(struct.pack("<Q", len(header)) + header + b"\0" * 44)
    with pytest.raises(CheckpointError, match="offsets"):
        SafetensorsFile(str(tmp_path / "o.safetensors"))
    with pytest.raises(CheckpointError, match="can't open"):
        VarBuilder.from_index(str(tmp_path / "nowhere" / INDEX_NAME))
    (tmp_path / INDEX_NAME).write_text('{"metadata": {}}')
    with pytest.raises(CheckpointError, match="no weight map"):
        VarBuilder.from_index(str(tmp_path / INDEX_NAME))
    (tmp_path / INDEX_NAME).write_text('{"weight_map": 3}')
    with pytest.raises(CheckpointError, match="not a map"):
        VarBuilder.from_index(str(tmp_path / INDEX_NAME))


def _sharded_by_layer(tmp_path, cfg, sd):
    """One shard per layer + one for embed/norm/head, like HF's layer-ordered shards."""
    from safetensors.torch import save_file
    wm = {}
    for i in range(cfg.num_hidden_layers):
        fn = f"model-{i + 1:05d}.safetensors"
        part = {k: v for k, v in sd.items() if k.startswith(cfg.layer_name(i) + ".")}
        save_file(part, str(tmp_path / fn))
        wm.update({k: fn for k in part})
    rest = {k: v for k, v in sd.items() if k not in wm}
    save_file(rest, str(tmp_path / "model-head.safetensors"))
    wm.update({k: "model-head.safetensors" for k in rest})
    json.dump({"metadata": {}, "weight_map": wm}, open(tmp_path / INDEX_NAME, "w"))
    json.dump(cfg.to_hf(), open(tmp_path / "config.json", "w"))


def test_shard_selection_follows_the_node_role(tmp_path):
    cfg = medium_config(num_hidden_layers=4)
    sd = checkpoint(cfg, "bf16", seed=9)
    _sharded_by_layer(tmp_path, cfg, sd)
    index = str(tmp_path / INDEX_NAME)
    topo = box_topology(cfg, 2)  # gpu1 owns layers 2-3
    assert topo["gpu1"]["layers"] == [cfg.layer_name(2), cfg.layer_name(3)]
    every = VarBuilder.from_index(index)
    assert len(every.shard_paths()) == 5 and set(every) == set(sd)
    # master: shards holding only worker layers are skipped (utils/mod.rs:298-311)
    _, master = open_model(str(tmp_path), topo)
    names = {os.path.basename(p) for p in master.shard_paths()}
    assert names == {"model-00001.safetensors", "model-00002.safetensors", "model-head.safetensors"}
    assert f"{cfg.layer_name(0)}.mlp.up_proj.weight" in master and f"{cfg.layer_name(2)}.mlp.up_proj.weight" not in master
    with pytest.raises(KeyError, match="cannot find tensor"):
        master[f"{cfg.layer_name(3)}.mlp.up_proj.weight"]
    assert master.get("not.there") is None
    # worker: only the shards with its layers (utils/mod.rs:352-364)
    _, worker = open_model(str(tmp_path), topo, worker="gpu1")
    assert {os.path.basename(p) for p in worker.shard_paths()} == {"model-00003.safetensors", "model-00004.safetensors"}
    assert worker.nbytes() == sum(v.numel() * 2 for k, v in sd.items() if ".layers.2." in k or ".layers.3." in k)
    with pytest.raises(CheckpointError, match="not in the topology"):
        open_model(str(tmp_path), topo, worker="gpu7")
    # a prefix must match whole path components: "model.layers.1" does not select "model.layers.10"
    assert VarBuilder.for_specific_layers(index, []).shard_paths() == every.shard_paths()


def test_layer_prefix_needs_a_dot_boundary(tmp_path):
    cfg = medium_config(num_hidden_layers=11)
    _sharded_by_layer(tmp_path, cfg, checkpoint(cfg, "bf16", seed=2))
    vb = VarBuilder.for_specific_layers(str(tmp_path / INDEX_NAME), [cfg.layer_name(1)])
    assert [os.path.basename(p) for p in vb.shard_paths()] == ["model-00002.safetensors"]


def test_model_prefix_is_detected_from_the_index(tmp_path):
    cfg = medium_config(num_hidden_layers=2)
    cfg.model_prefix = "language_model.model"
    sd = checkpoint(cfg, "bf16", seed=4)
    save_checkpoint(str(tmp_path), cfg, sd, shard_bytes=10_000_000)
    assert detect_model_prefix(str(tmp_path / INDEX_NAME), "model") == "language_model.model"
    got, vb = open_model(str(tmp_path))
    assert got.model_prefix == "language_model.model" and got.layer_name(1) == "language_model.model.layers.1"
    assert f"{got.model_prefix}.embed_tokens.weight" in vb
    assert detect_model_prefix(str(tmp_path / "missing.json"), "model") == "model"


def test_single_file_checkpoint_without_index(tmp_path):
    cfg = medium_config(num_hidden_layers=1)
    sd = checkpoint(cfg, "f16", seed=8)
    assert save_checkpoint(str(tmp_path), cfg, sd) == ["model.safetensors"]
    got, vb = open_model(str(tmp_path))
    assert set(vb) == set(sd) and vb["lm_head.weight"].dtype == torch.float16
    with pytest.raises(CheckpointError, match="config.json"):
        open_model(str(tmp_path / "nope"))


def test_oracle_tokens_from_disk_equal_tokens_from_memory(tmp_path):
    """The loader feeds exactly the bytes of the state dict: the oracle generates the same ids from both."""
    from oracle.oracle import OracleModel
    cfg = medium_config(num_hidden_layers=2)
    sd = checkpoint(cfg, "bf16", seed=61, peaked=True)
    save_checkpoint(str(tmp_path), cfg, sd, shard_bytes=2_000_000)
    got, vb = open_model(str(tmp_path))
    prompt = np.random.default_rng(2).integers(0, cfg.vocab_size - 1, 9).tolist()
    a = OracleModel(cfg, sd, "bf16", max_seq=64).generate(prompt, 8)
    b = OracleModel(got, vb, "bf16", max_seq=64).generate(prompt, 8)
    assert list(a[0]) == list(b[0])


@pytest.mark.gpu
def test_gpu_master_loads_from_disk(tmp_path):
    from cake_b200.model import Context, Master, TextModelBase
    cfg = medium_config()
    sd = checkpoint(cfg, "bf16", seed=61, peaked=True)
    save_checkpoint(str(tmp_path), cfg, sd, shard_bytes=2_000_000)
    prompt = np.random.default_rng(2).integers(0, cfg.vocab_size - 1, 10).tolist()

    def run(config, weights):
        ctx = Context(config, weights, "bf16", max_seq=64)
        try:
            return Master(TextModelBase.load(ctx)).generate_text(prompt, 12)["tokens"]
        finally:
            ctx.close()

    got, vb = open_model(str(tmp_path))
    from_disk, from_memory = run(got, vb), run(cfg, sd)
    assert len(from_disk) == 12 and from_disk == from_memory
    # ... and both are the ORACLE's greedy tokens for the checkpoint read back from disk (not only self-consistent)
    from oracle.oracle import OracleModel
    want = list(OracleModel(got, vb, "bf16", max_seq=64).generate(prompt, 12)[0])
    assert from_disk == want


def test_wrong_shapes_are_refused_before_anything_is_copied():
    """candle's `vb.get(shape, name)` refuses a tensor whose shape is not the one the layer declares; the loader must do
    the same before it hands a pointer to the library (which would read rows x cols elements through it)."""
    from cake_b200.model import B200Transformer, TextModelBase

    cfg = medium_config(num_hidden_layers=1)
    sd = checkpoint(cfg, "bf16", seed=3)

    class Ctx:  # only what load() touches before the first library call
        config, var_builder, torch_dtype, topology = cfg, sd, torch.bfloat16, {}

    name = cfg.layer_name(0)
    good = sd[f"{name}.mlp.down_proj.weight"]
    sd[f"{name}.mlp.down_proj.weight"] = good[:, :-8].contiguous()
    with pytest.raises(ValueError, match=r"shape mismatch for model.layers.0.mlp.down_proj.weight, expected: \[512, 1024\], got: \[512, 1016\]"):
        B200Transformer.load(name, Ctx)
    sd[f"{name}.mlp.down_proj.weight"] = good
    del sd[f"{name}.self_attn.o_proj.weight"]
    with pytest.raises(KeyError, match="self_attn.o_proj.weight not found"):
        B200Transformer.load(name, Ctx)
    sd["lm_head.weight"] = sd["lm_head.weight"][:-1].contiguous()
    with pytest.raises(ValueError, match="shape mismatch for lm_head.weight"):
        TextModelBase.load(Ctx)


def _fnv1a(b: bytes) -> int:
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.parametrize("layers", [None, "model.layers.2-3"])
def test_cpp_varbuilder_maps_the_same_tensors_as_the_python_loader(tmp_path, layers):
    """The compiled mmapped reader (cake_host.hpp VarBuilder) and cake_b200/loader.py must see the same shards, names,
    dtypes, shapes and bytes — for the whole checkpoint and for a worker's layer subset (utils/mod.rs:334-384)."""
    import subprocess
    from cake_b200.build import build_host
    build_host()
    worker = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cake_b200", "host", "cake_worker")
    cfg = medium_config(num_hidden_layers=4, hidden_size=128, intermediate_size=256, vocab_size=256, num_attention_heads=4,
                        num_key_value_heads=2, head_dim=32)
    sd = checkpoint(cfg, "bf16", seed=12)
    _sharded_by_layer(tmp_path, cfg, sd)
    args = [worker, str(tmp_path), "--list-tensors"] + (["--layers", layers] if layers else [])
    r = subprocess.run(args, capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().split("\n")
    index = str(tmp_path / INDEX_NAME)
    vb = VarBuilder.for_specific_layers(index, [cfg.layer_name(2), cfg.layer_name(3)]) if layers else VarBuilder.from_index(index)
    assert lines[0] == f"files {len(vb.shard_paths())} prefix model"
    got = {}
    for ln in lines[1:]:
        name, dtype, shape, nbytes, h = ln.split()
        got[name] = (dtype, shape, int(nbytes), int(h, 16))
    assert set(got) == set(vb)
    for name in vb:
        t = vb[name]
        raw = t.contiguous().view(torch.uint8).numpy().tobytes()
        assert got[name] == ("BF16", "[" + "x".join(str(d) for d in t.shape) + "]", len(raw), _fnv1a(raw)), name


def test_prefused_checkpoints_split_into_views_and_give_the_same_model(tmp_path):
    """Phi-3/4 checkpoints store qkv_proj / gate_up_proj pre-fused (attention.rs:90-94, mlp.rs:38-40).  The loader hands
    the library row views of the fused tensor (no copy), and the model is the same as the unfused one."""
    from cake_b200.loader import block_tensors
    from oracle.oracle import OracleModel
    kw = dict(num_hidden_layers=2, hidden_size=128, intermediate_size=256, vocab_size=256, num_attention_heads=4,
              num_key_value_heads=2, head_dim=32, partial_rotary_factor=0.75)
    plain = medium_config(**kw)
    fused = medium_config(**kw, fused_qkv_proj=True, fused_gate_up_proj=True)
    sd_p, sd_f = checkpoint(plain, "bf16", seed=4, peaked=True), checkpoint(fused, "bf16", seed=4, peaked=True)
    n = fused.layer_name(1)
    assert f"{n}.self_attn.qkv_proj.weight" in sd_f and f"{n}.self_attn.q_proj.weight" not in sd_f
    assert f"{n}.mlp.gate_up_proj.weight" in sd_f and f"{n}.mlp.up_proj.weight" not in sd_f
    save_checkpoint(str(tmp_path), fused, sd_f, arch="Phi3ForCausalLM", shard_bytes=300_000)
    cfg, vb = open_model(str(tmp_path))
    assert cfg.fused_qkv_proj and cfg.fused_gate_up_proj and cfg.partial_rotary_factor == 0.75 and cfg.hd == 32
    views, ref = block_tensors(vb, cfg, n), block_tensors(sd_p, plain, n)
    for k, t in ref.items():
        assert (t is None) == (views[k] is None), k
        if t is not None:
            assert torch.equal(views[k], t), k
    w = vb[f"{n}.self_attn.qkv_proj.weight"]
    assert views["self_attn.q_proj.weight"].data_ptr() == w.data_ptr()                       # views, not copies
    assert views["self_attn.k_proj.weight"].data_ptr() == w.data_ptr() + cfg.size_q * cfg.hidden_size * 2
    assert views["self_attn.v_proj.weight"].is_contiguous() and views["mlp.up_proj.weight"].is_contiguous()
    prompt = np.random.default_rng(1).integers(0, cfg.vocab_size - 1, 6).tolist()
    a = OracleModel(plain, sd_p, "bf16", max_seq=32).generate(prompt, 6)[0]
    b = OracleModel(cfg, vb, "bf16", max_seq=32).generate(prompt, 6)[0]
    assert list(a) == list(b)
    # a fused tensor of the wrong height is refused
    bad = dict(sd_f)
    bad[f"{n}.mlp.gate_up_proj.weight"] = bad[f"{n}.mlp.gate_up_proj.weight"][:-2].contiguous()
    with pytest.raises(ValueError, match="shape mismatch for model.layers.1.mlp.gate_up_proj.weight"):
        block_tensors(bad, fused, n)


@pytest.mark.parametrize("flavour", ["llama", "qwen2_bias", "qwen3_qknorm", "phi_fused"])
def test_block_load_receives_the_fourteen_tensors_in_header_order(monkeypatch, flavour):
    """B200Transformer.load hands cake_b200_block_load (include/cake_b200.h) q,k,v,o,gate,up,down,ln1,ln2, three biases
    and two QK-norm weights — in that order, null where the config does not use them, and pointing at exactly the
    checkpoint's bytes (row views for pre-fused checkpoints).  A recording stand-in for the library reads the memory
    behind every pointer; no GPU involved."""
    import ctypes
    from cake_b200 import model as M
    from cake_b200.synth import layer_tensor_shapes
    kw = dict(num_hidden_layers=2, hidden_size=128, intermediate_size=256, vocab_size=256, num_attention_heads=4,
              num_key_value_heads=2, head_dim=32)
    extra = {"llama": {}, "qwen2_bias": dict(use_qkv_bias=True), "qwen3_qknorm": dict(use_qk_norm=True),
             "phi_fused": dict(fused_qkv_proj=True, fused_gate_up_proj=True, partial_rotary_factor=0.5)}[flavour]
    cfg = medium_config(**kw, **extra)
    plain = medium_config(**kw, **{k: v for k, v in extra.items() if not k.startswith("fused")})
    sd, sd_plain = checkpoint(cfg, "bf16", seed=77), checkpoint(plain, "bf16", seed=77)
    seen = {}

    class FakeLib:
        def cake_b200_block_load(self, ctx_h, layer, *rest):
            seen["layer"], seen["ptrs"] = layer, list(rest[:14])
            return 0

    monkeypatch.setattr(M, "lib", lambda: FakeLib())

    class Ctx:
        config, var_builder, torch_dtype, h, _children = cfg, sd, torch.bfloat16, ctypes.c_void_p(1), []

    name = cfg.layer_name(1)
    blk = M.B200Transformer.load(name, Ctx)
    assert blk.layer_name() == name and seen["layer"] == 1 and len(Ctx._children) == 1
    order = ["self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight", "self_attn.o_proj.weight",
             "mlp.gate_proj.weight", "mlp.up_proj.weight", "mlp.down_proj.weight", "input_layernorm.weight",
             "post_attention_layernorm.weight", "self_attn.q_proj.bias", "self_attn.k_proj.bias", "self_attn.v_proj.bias",
             "self_attn.q_norm.weight", "self_attn.k_norm.weight"]
    shapes = layer_tensor_shapes(plain)
    for short, p in zip(order, seen["ptrs"]):
        if short not in shapes:
            assert not p, f"{short} must be null for {flavour}"
            continue
        want = sd_plain[f"{name}.{short}"].contiguous().view(torch.uint8).numpy().tobytes()
        assert p and ctypes.string_at(p, len(want)) == want, short
    blk.h = None  # nothing to free in the stand-in
