"""Weight loading: HF safetensors checkpoints -> host tensors that alias the mmapped shards.

Mirrors the reference's loading path for the block-forward scope:

  utils/mod.rs:160-206     model.safetensors | model.safetensors.index.json -> shard list
  utils/mod.rs:250-267     load_var_builder_from_index            (every shard)
  utils/mod.rs:272-329     load_var_builder_for_local_layers      (master: skip shards that hold only worker layers)
  utils/mod.rs:334-384     load_var_builder_for_specific_layers   (worker: only shards that hold its layers)
  utils/mod.rs:209-247     prefetch_safetensors                   (here: madvise(WILLNEED) on the mapping)
  cake/mod.rs:335-357      weight-prefix auto-detection from the index ("….layers.0.")

The safetensors container is parsed here (8-byte little-endian header length, JSON header, raw little-endian
tensor bytes) rather than through a library so that the tensors handed to ``cake_b200_block_load`` are *views of
the mapping*: the H2D copy into the fused qkv / gate_up layouts reads the page cache directly and nothing is staged
twice.  A ``VarBuilder`` is a read-only mapping ``name -> torch.Tensor`` (what ``Context.var_builder`` expects).
"""
from __future__ import annotations

import json
import mmap
import os
import struct
from collections.abc import Mapping
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import torch

from .config import Config

INDEX_NAME = "model.safetensors.index.json"
SINGLE_NAME = "model.safetensors"
MAX_HEADER = 100 * 1024 * 1024  # the format's own limit

_DTYPES = {
    "BF16": torch.bfloat16, "F16": torch.float16, "F32": torch.float32, "F64": torch.float64,
    "I64": torch.int64, "I32": torch.int32, "I16": torch.int16, "I8": torch.int8, "U8": torch.uint8,
    "BOOL": torch.bool, "F8_E4M3": torch.float8_e4m3fn, "F8_E5M2": torch.float8_e5m2,
}


class CheckpointError(RuntimeError):
    """Malformed or incomplete checkpoint (the reference bails with an anyhow error naming the file)."""


class SafetensorsFile:
    """One mmapped ``.safetensors`` shard."""

    def __init__(self, path: str):
        self.path = path
        try:
            size = os.path.getsize(path)
            f = open(path, "rb")
        except OSError as e:
            raise CheckpointError(f"can't open {path}: {e}") from e
        with f:
            head = f.read(8)
            if len(head) != 8:
                raise CheckpointError(f"{path}: shorter than a safetensors header")
            (n,) = struct.unpack("<Q", head)
            if n > MAX_HEADER or 8 + n > size:
                raise CheckpointError(f"{path}: header length {n} does not fit the file ({size} bytes)")
            try:
                header = json.loads(f.read(n))
            except (UnicodeDecodeError, json.JSONDecodeError) as e:
                raise CheckpointError(f"{path}: header is not JSON: {e}") from e
            if not isinstance(header, dict):
                raise CheckpointError(f"{path}: header is not a JSON object")
            # private copy-on-write mapping: never written, but a writable buffer lets torch alias it silently
            self._mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_COPY) if size else None
        self.metadata = header.pop("__metadata__", None)
        self._base = 8 + n
        self._entries: Dict[str, Tuple[torch.dtype, Tuple[int, ...], int, int]] = {}
        data_len = size - self._base
        for name, e in header.items():
            try:
                dt, shape, (b, en) = _DTYPES[e["dtype"]], tuple(int(d) for d in e["shape"]), e["data_offsets"]
            except (KeyError, TypeError, ValueError) as err:
                raise CheckpointError(f"{path}: bad entry for tensor {name!r}: {err!r}") from err
            numel = 1
            for d in shape:
                numel *= d
            if not (0 <= b <= en <= data_len) or en - b != numel * dt.itemsize:
                raise CheckpointError(f"{path}: tensor {name!r} has offsets [{b},{en}) for shape {shape}")
            self._entries[name] = (dt, shape, b, numel)

    def names(self) -> Iterable[str]:
        return self._entries.keys()

    def __contains__(self, name: str) -> bool:
        return name in self._entries

    def nbytes(self, name: str) -> int:
        dt, _, _, numel = self._entries[name]
        return numel * dt.itemsize

    def get(self, name: str) -> torch.Tensor:
        dt, shape, b, numel = self._entries[name]
        if numel == 0:
            return torch.empty(shape, dtype=dt)
        off = self._base + b
        if off % dt.itemsize:  # header not padded to the element size: alias is impossible, copy once
            raw = bytearray(self._mm[off:off + numel * dt.itemsize])
            return torch.frombuffer(raw, dtype=dt, count=numel).view(shape)
        return torch.frombuffer(self._mm, dtype=dt, count=numel, offset=off).view(shape)

    def prefetch(self) -> None:
        """Ask the kernel to read the shard ahead (the reference reads every shard once into a scratch Vec)."""
        if self._mm is not None and hasattr(self._mm, "madvise"):
            self._mm.madvise(mmap.MADV_WILLNEED)


def _read_weight_map(index_path: str) -> Dict[str, str]:
    try:
        with open(index_path) as f:
            j = json.load(f)
    except OSError as e:
        raise CheckpointError(f"can't open {index_path}: {e}") from e
    except json.JSONDecodeError as e:
        raise CheckpointError(f"can't parse {index_path}: {e}") from e
    wm = j.get("weight_map") if isinstance(j, dict) else None
    if wm is None:
        raise CheckpointError(f"no weight map in {index_path}")
    if not isinstance(wm, dict):
        raise CheckpointError(f"weight map in {index_path} is not a map")
    return {k: v for k, v in wm.items() if isinstance(v, str)}


def _under(name: str, prefixes: Iterable[str]) -> bool:
    return any(name.startswith(p + ".") for p in prefixes)


class VarBuilder(Mapping):
    """Read-only ``name -> tensor`` view over a set of shards (candle's mmapped VarBuilder, as the path uses it)."""

    def __init__(self, files: Sequence[SafetensorsFile], prefetch: bool = True):
        self.files = list(files)
        self._where: Dict[str, SafetensorsFile] = {}
        for f in self.files:
            for n in f.names():
                self._where[n] = f
            if prefetch:
                f.prefetch()

    # -- the three reference entry points --------------------------------------------------------
    @classmethod
    def from_index(cls, index_path: str, **kw) -> "VarBuilder":
        """utils/mod.rs:250-267: every shard of the index, or ``model.safetensors`` next to it when there is none."""
        parent = os.path.dirname(index_path) or "."
        if os.path.exists(index_path):
            shards = sorted(set(_read_weight_map(index_path).values()))
        else:
            shards = [SINGLE_NAME]
        return cls([SafetensorsFile(os.path.join(parent, s)) for s in shards], **kw)

    @classmethod
    def for_local_layers(cls, index_path: str, worker_layers: Iterable[str], **kw) -> "VarBuilder":
        """utils/mod.rs:272-329 (master): keep a shard iff it holds at least one tensor that is NOT under a worker layer."""
        worker_layers = set(worker_layers)
        if not os.path.exists(index_path) or not worker_layers:
            return cls.from_index(index_path, **kw)
        parent = os.path.dirname(index_path) or "."
        need = sorted({shard for name, shard in _read_weight_map(index_path).items() if not _under(name, worker_layers)})
        return cls([SafetensorsFile(os.path.join(parent, s)) for s in need], **kw)

    @classmethod
    def for_specific_layers(cls, index_path: str, layer_prefixes: Iterable[str], **kw) -> "VarBuilder":
        """utils/mod.rs:334-384 (worker): keep a shard iff it holds a tensor under one of ``layer_prefixes``."""
        layer_prefixes = list(layer_prefixes)
        if not os.path.exists(index_path) or not layer_prefixes:
            return cls.from_index(index_path, **kw)
        parent = os.path.dirname(index_path) or "."
        need = sorted({shard for name, shard in _read_weight_map(index_path).items() if _under(name, layer_prefixes)})
        return cls([SafetensorsFile(os.path.join(parent, s)) for s in need], **kw)

    # -- Mapping ---------------------------------------------------------------------------------
    def __getitem__(self, name: str) -> torch.Tensor:
        f = self._where.get(name)
        if f is None:
            raise KeyError(f"cannot find tensor {name}")  # candle's VarBuilder error text
        return f.get(name)

    def __iter__(self) -> Iterator[str]:
        return iter(self._where)

    def __len__(self) -> int:
        return len(self._where)

    def __contains__(self, name) -> bool:
        return name in self._where

    def shard_paths(self) -> List[str]:
        return [f.path for f in self.files]

    def nbytes(self) -> int:
        return sum(f.nbytes(n) for n, f in self._where.items())


BLOCK_TENSORS = ("self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight", "self_attn.o_proj.weight",
                 "mlp.gate_proj.weight", "mlp.up_proj.weight", "mlp.down_proj.weight", "input_layernorm.weight",
                 "post_attention_layernorm.weight", "self_attn.q_proj.bias", "self_attn.k_proj.bias", "self_attn.v_proj.bias",
                 "self_attn.q_norm.weight", "self_attn.k_norm.weight")


# the two extra RmsNorm vectors of the sandwich-norm blocks (olmo2/block.rs:49-52, gemma3/block.rs:84-91), by logical role
EXTRA_TENSORS = ("post_attn_norm", "post_ffn_norm")


def rms_norm_weight(t: torch.Tensor, cfg: Config) -> torch.Tensor:
    """config.rs:155-173 load_rms_norm_weight: with ``residual_rms_norm`` the checkpoint stores deltas and the forward
    weight is (1 + w), added in f32 and cast back to the model dtype AT LOAD TIME — the kernels never see the flag."""
    if not cfg.residual_rms_norm:
        return t
    return (t.to(torch.float32) + 1.0).to(t.dtype)


def block_tensors(vb, cfg: Config, name: str) -> Dict[str, Optional[torch.Tensor]]:
    """The tensors `Transformer::load` reads for layer ``name`` (transformer.rs:79-101, attention.rs:76-149,
    mlp.rs:34-59), by the short names of BLOCK_TENSORS, each checked against the shape the layer declares (candle's
    ``vb.get(shape, name)``).  Pre-fused checkpoints (Phi-3/4: ``cfg.fused_qkv_proj`` / ``cfg.fused_gate_up_proj``,
    attention.rs:90-94, mlp.rs:38-40) are split into row views of the fused tensor — q, k, v and gate, up are
    contiguous row ranges of it, so no copy is made.  Biases / QK-norm weights are None unless the config uses them."""
    H, I, sq, skv, hd = cfg.hidden_size, cfg.intermediate_size, cfg.size_q, cfg.size_kv, cfg.hd

    def get(short: str, shape) -> torch.Tensor:
        t = vb.get(f"{name}.{short}")
        if t is None:
            raise KeyError(f"tensor {name}.{short} not found")  # candle VarBuilder error
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"shape mismatch for {name}.{short}, expected: {list(shape)}, got: {list(t.shape)}")
        return t

    out: Dict[str, Optional[torch.Tensor]] = {k: None for k in BLOCK_TENSORS}
    if cfg.fused_qkv_proj:
        w = get("self_attn.qkv_proj.weight", (sq + 2 * skv, H))
        out["self_attn.q_proj.weight"], out["self_attn.k_proj.weight"], out["self_attn.v_proj.weight"] = \
            w[:sq], w[sq:sq + skv], w[sq + skv:]
    else:
        out["self_attn.q_proj.weight"] = get("self_attn.q_proj.weight", (sq, H))
        out["self_attn.k_proj.weight"] = get("self_attn.k_proj.weight", (skv, H))
        out["self_attn.v_proj.weight"] = get("self_attn.v_proj.weight", (skv, H))
    out["self_attn.o_proj.weight"] = get("self_attn.o_proj.weight", (H, sq))
    if cfg.fused_gate_up_proj:
        w = get("mlp.gate_up_proj.weight", (2 * I, H))
        out["mlp.gate_proj.weight"], out["mlp.up_proj.weight"] = w[:I], w[I:]
    else:
        out["mlp.gate_proj.weight"] = get("mlp.gate_proj.weight", (I, H))
        out["mlp.up_proj.weight"] = get("mlp.up_proj.weight", (I, H))
    out["mlp.down_proj.weight"] = get("mlp.down_proj.weight", (H, I))
    # Norm vectors by ROLE.  out["input_layernorm.weight"] is the pre-attention norm and out["post_attention_layernorm.weight"]
    # the pre-MLP norm (the Llama naming, transformer.rs:84-90); the sibling blocks name theirs differently:
    #   gemma3 (block.rs:84-91): input_layernorm | post_attention_layernorm = POST-attention | pre_feedforward_layernorm = pre-MLP
    #                            | post_feedforward_layernorm
    #   olmo2  (block.rs:49-52): no pre-norms; post_attention_layernorm and post_feedforward_layernorm are post-norms
    kind = getattr(cfg, "block_kind", "llama")
    out.update({k: None for k in EXTRA_TENSORS})
    norm = lambda short: rms_norm_weight(get(short, (H,)), cfg)
    if kind in ("olmo2", "exaone4_hf"):
        out["post_attn_norm"], out["post_ffn_norm"] = norm("post_attention_layernorm.weight"), norm("post_feedforward_layernorm.weight")
    elif kind == "gemma3":
        out["input_layernorm.weight"] = norm("input_layernorm.weight")
        out["post_attn_norm"] = norm("post_attention_layernorm.weight")
        out["post_attention_layernorm.weight"] = norm("pre_feedforward_layernorm.weight")
        out["post_ffn_norm"] = norm("post_feedforward_layernorm.weight")
    else:
        out["input_layernorm.weight"] = norm("input_layernorm.weight")
        out["post_attention_layernorm.weight"] = norm("post_attention_layernorm.weight")
    if cfg.use_qkv_bias:  # attention.rs:96-107 (never together with a fused qkv_proj)
        out["self_attn.q_proj.bias"] = get("self_attn.q_proj.bias", (sq,))
        out["self_attn.k_proj.bias"] = get("self_attn.k_proj.bias", (skv,))
        out["self_attn.v_proj.bias"] = get("self_attn.v_proj.bias", (skv,))
    if cfg.use_qk_norm:   # attention.rs:120-129
        pre = getattr(cfg, "pre_reshape_qk_norm", False)  # attention.rs:121-122: norm dim = size_q / size_kv (OLMo2) or head_dim
        out["self_attn.q_norm.weight"] = rms_norm_weight(get("self_attn.q_norm.weight", (sq if pre else hd,)), cfg)
        out["self_attn.k_norm.weight"] = rms_norm_weight(get("self_attn.k_norm.weight", (skv if pre else hd,)), cfg)
    return out


def detect_model_prefix(index_path: str, configured: str) -> str:
    """cake/mod.rs:335-357: the text before the first ``.layers.0.`` key of the index wins over the configured prefix."""
    if not os.path.exists(index_path):
        return configured
    try:
        wm = _read_weight_map(index_path)
    except CheckpointError:
        return configured  # the reference ignores an unreadable index at this point
    for key in wm:
        pos = key.find(".layers.0.")
        if pos >= 0:
            return key[:pos]
    return configured


def worker_layer_names(topology: Optional[dict]) -> List[str]:
    """Topology::all_worker_layers (topology.rs:174-185) for an already-expanded topology
    ``{worker: {"layers": [...]}}`` (see parallel.parse_topology)."""
    out: List[str] = []
    for node in (topology or {}).values():
        out.extend(node.get("layers", []))
    return out


def open_model(path: str, topology: Optional[dict] = None, worker: Optional[str] = None,
               prefetch: bool = True) -> Tuple[Config, VarBuilder]:
    """Context::from_args' model part (cake/mod.rs:280-392): ``config.json`` -> Config (prefix auto-detected),
    then the VarBuilder variant the node's role calls for:

    * master (``worker is None``): everything except shards that hold only worker layers;
    * worker ``worker``: only the shards that hold its layers.
    """
    cfg_path = os.path.join(path, "config.json")
    if not os.path.exists(cfg_path):
        raise CheckpointError(f"can't open {cfg_path}")
    cfg = Config.from_path(cfg_path)
    index = os.path.join(path, INDEX_NAME)
    cfg.model_prefix = detect_model_prefix(index, cfg.model_prefix)
    if worker is None:
        vb = VarBuilder.for_local_layers(index, worker_layer_names(topology), prefetch=prefetch)
    else:
        if not topology or worker not in topology:
            raise CheckpointError(f"worker {worker!r} is not in the topology")
        vb = VarBuilder.for_specific_layers(index, topology[worker].get("layers", []), prefetch=prefetch)
    return cfg, vb


def save_checkpoint(path: str, cfg: Config, tensors: Dict[str, torch.Tensor], arch: str = "LlamaForCausalLM",
                    shard_bytes: Optional[int] = None) -> List[str]:
    """Write an HF-layout checkpoint (config.json + model.safetensors, or shards + index when ``shard_bytes`` is
    given).  Test / benchmark utility (the synthetic checkpoints of SURVEY.md §8d); tensors are written in name
    order, a new shard starts when the current one would exceed ``shard_bytes``."""
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(cfg.to_hf(arch), f)
    rev = {v: k for k, v in _DTYPES.items()}
    groups: List[List[str]] = [[]]
    used = 0
    for name in sorted(tensors):
        nb = tensors[name].numel() * tensors[name].element_size()
        if shard_bytes and groups[-1] and used + nb > shard_bytes:
            groups.append([])
            used = 0
        groups[-1].append(name)
        used += nb
    files = [SINGLE_NAME] if len(groups) == 1 and not shard_bytes else \
        [f"model-{i + 1:05d}-of-{len(groups):05d}.safetensors" for i in range(len(groups))]
    weight_map = {}
    for fname, names in zip(files, groups):
        header, off = {}, 0
        for n in names:
            t = tensors[n]
            nb = t.numel() * t.element_size()
            header[n] = {"dtype": rev[t.dtype], "shape": list(t.shape), "data_offsets": [off, off + nb]}
            off += nb
            weight_map[n] = fname
        blob = json.dumps(header, separators=(",", ":")).encode()
        blob += b" " * (-len(blob) % 8)  # keep tensor data 8-byte aligned, as the safetensors writers do
        with open(os.path.join(path, fname), "wb") as f:
            f.write(struct.pack("<Q", len(blob)))
            f.write(blob)
            for n in names:
                t = tensors[n].detach().cpu().contiguous()
                f.write(t.view(torch.uint8).numpy().tobytes() if t.numel() else b"")
    if files != [SINGLE_NAME]:
        with open(os.path.join(path, INDEX_NAME), "w") as f:
            json.dump({"metadata": {"total_size": sum(t.numel() * t.element_size() for t in tensors.values())},
                       "weight_map": weight_map}, f)
    return files
