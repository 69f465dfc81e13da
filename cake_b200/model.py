"""Host-side mirror of the reference's interface for the block-forward path, over the C ABI.

Same names, argument meaning and error behaviour as the reference so that tests read like its own:

  Forwarder            cake-core/src/cake/mod.rs:510-556   (load / forward / forward_mut / forward_batch /
                                                            goodbye / layer_name / ident)
  B200Transformer      models/common/transformer.rs:14-150 (`type Shardable`, the local block)
  Cache                models/common/cache.rs:9-254        (clear / as_new; K,V live on the GPU)
  Context              cake/mod.rs:41-65                   (config, dtype, device, var_builder, cache, topology)
  TextModelBase        models/common/text_model.rs:133-530 (forward / prepare_prompt / next_token / reset / goodbye)
  Master               cake/sharding/master.rs:14-191      (generate_text and its tok/s definition)

torch is used for device memory and (in parallel.py) torch.distributed rendezvous only.  There is no
PyTorch or CPU fallback: every forward goes through libcake_b200.so.
"""
from __future__ import annotations

import ctypes
import time
import weakref
from abc import ABC, abstractmethod
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import capi
from .capi import byref, c_uint32, c_void_p, check, int_array, lib, ptr, ptr_array
from .config import CConfig, Config
from .synth import TORCH_DTYPES


class Cache:
    """KV cache of one session (cache.rs:9).  RoPE tables live in the ctx (computed once per model)."""

    def __init__(self, ctx: "Context", batch: int = 1, max_seq: Optional[int] = None):
        self.ctx, self.batch = ctx, batch
        self.max_seq = max_seq or ctx.max_seq
        self.h = c_void_p()
        self._lib = lib()  # a handle is only ever given back to the library that made it
        check(self._lib.cake_b200_cache_create(ctx.h, batch, self.max_seq, byref(self.h)))
        ctx._children.append(weakref.ref(self))

    def close(self):
        if self.h:
            self._lib.cake_b200_cache_free(self.h)
            self.h = None

    def clear(self) -> None:  # cache.rs:247-253
        check(lib().cake_b200_cache_clear(self.h))

    def as_new(self) -> "Cache":  # cache.rs:241-245
        return Cache(self.ctx, self.batch, self.max_seq)

    def with_kv_cache(self) -> bool:
        return True

    def len(self, block_idx: int) -> int:
        return lib().cake_b200_cache_len(self.h, block_idx)

    def kv(self, block_idx: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(K, V) copies of shape (batch, n_kv, len, hd) on the host, dtype D."""
        c = self.ctx.config
        n = self.len(block_idx)
        out = []
        for which in (0, 1):
            t = torch.empty((self.batch, c.num_key_value_heads, n, c.hd), dtype=TORCH_DTYPES[self.ctx.dtype])
            check(lib().cake_b200_cache_read(self.h, block_idx, which, ptr(t), t.numel() * t.element_size()))
            out.append(t)
        return out[0], out[1]

    def fill_synthetic(self, block_idx: Sequence[int], length: int, seed: int = 7) -> None:
        check(lib().cake_b200_cache_fill_synthetic(self.h, int_array(list(block_idx)), len(block_idx), length, seed))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """cake/mod.rs:41-65: shared state handed to Forwarder::load / forward."""

    def __init__(self, config: Config, var_builder: Dict[str, torch.Tensor], dtype: str = "bf16", device: int = 0,
                 max_seq: Optional[int] = None, topology: Optional[dict] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("cake_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.config, self.var_builder, self.dtype, self.device = config, var_builder, dtype, device
        self.max_seq = max_seq or config.max_seq_len
        self.topology = topology or {}
        self.ccfg = CConfig.from_config(config, dtype, self.max_seq)
        self.h = c_void_p()
        self._children = []  # weakrefs to caches / blocks: freed before the ctx (they point into it)
        torch.cuda.set_device(device)
        check(lib().cake_b200_ctx_create(device, byref(self.ccfg), byref(self.h)))
        self.cache: Optional[Cache] = Cache(self)
        self.torch_dtype = TORCH_DTYPES[dtype]
        # torch sees the library's stream as an ExternalStream: tensors handed to the C ABI are allocated
        # on it, so the caching allocator orders their reuse after the kernels that touch them.
        self.torch_stream = torch.cuda.ExternalStream(self.stream_ptr, device=device)

    def sync(self) -> None:
        check(lib().cake_b200_sync(self.h))

    @property
    def stream_ptr(self) -> int:
        return lib().cake_b200_stream(self.h)

    def launch_count(self) -> int:
        n = ctypes.c_uint64()
        check(lib().cake_b200_launch_count(self.h, byref(n)))
        return n.value

    def empty(self, *shape) -> torch.Tensor:
        with torch.cuda.stream(self.torch_stream):
            return torch.empty(shape, dtype=self.torch_dtype, device=f"cuda:{self.device}")

    def to_device(self, t: torch.Tensor) -> torch.Tensor:
        """Host tensor (values in D) -> device tensor on the library's stream."""
        with torch.cuda.stream(self.torch_stream):
            return t.to(self.torch_dtype).to(f"cuda:{self.device}", non_blocking=False)

    def close(self):
        if self.h:
            self.sync()
            for ref in self._children:
                obj = ref()
                if obj is not None:
                    obj.close()
            self._children = []
            self.cache = None
            lib().cake_b200_ctx_destroy(self.h)
            self.h = None


class Forwarder(ABC):
    """cake/mod.rs:510-556."""

    @classmethod
    @abstractmethod
    def load(cls, name: str, ctx: Context) -> "Forwarder": ...

    @abstractmethod
    def forward(self, x: torch.Tensor, index_pos: int, block_idx: int, ctx: Context) -> torch.Tensor: ...

    def forward_mut(self, x, index_pos, block_idx, ctx):
        return self.forward(x, index_pos, block_idx, ctx)

    def forward_batch(self, x, batch: List[Tuple[str, int, int]], ctx):
        raise NotImplementedError  # mod.rs:538-543 unimplemented!() by default

    def goodbye(self) -> None:
        return None

    @abstractmethod
    def layer_name(self) -> str: ...

    def ident(self) -> str:
        return "local"


def _layer_index(name: str) -> int:
    return int(name.rsplit(".", 1)[1])


class B200Transformer(Forwarder):
    """The local transformer block (`type Shardable = Transformer`, llama.rs:52 / qwen3/model.rs:24),
    backed by a cake_b200_block handle."""

    def __init__(self, name: str, handle: c_void_p, ctx: Context):
        self.name, self.h = name, handle
        self._lib = lib()
        ctx._children.append(weakref.ref(self))

    def close(self):
        if self.h:
            self._lib.cake_b200_block_free(self.h)
            self.h = None

    @classmethod
    def load(cls, name: str, ctx: Context) -> "B200Transformer":
        from .loader import BLOCK_TENSORS, block_tensors
        views = block_tensors(ctx.var_builder, ctx.config, name)
        keep = []
        for short in BLOCK_TENSORS:  # the argument order of cake_b200_block_load
            t = views[short]
            if t is not None:
                if t.dtype != ctx.torch_dtype:
                    t = t.to(ctx.torch_dtype)
                t = t.contiguous()   # row views of a fused tensor are contiguous already
            keep.append(t)
        h = c_void_p()
        check(lib().cake_b200_block_load(ctx.h, _layer_index(name), *[ptr(t) for t in keep], byref(h)))
        blk = cls(name, h, ctx)
        # the sibling block structures (models/{olmo2,gemma3,exaone4}/block.rs): post-norms and the per-layer attention mode
        var = ctx.config.layer_variant(_layer_index(name), getattr(ctx, "max_seq", None))
        if var["post_norms"] or var["window"] >= 0 or var["no_rope"]:
            from .capi import CBlockVariant
            extra = []
            for k in ("post_attn_norm", "post_ffn_norm"):
                t = views.get(k)
                if t is not None:
                    t = t.to(ctx.torch_dtype).contiguous()
                extra.append(t)
            cv = CBlockVariant(var["window"], 0 if var["no_rope"] else 1, ptr(extra[0]), ptr(extra[1]))
            check(lib().cake_b200_block_set_variant(h, byref(cv)))
        return blk

    def forward(self, x, index_pos, block_idx, ctx):
        return self.forward_batch(x, [(self.name, index_pos, block_idx)], ctx, blocks=[self])

    def forward_batch(self, x, batch, ctx, blocks: Optional[List["B200Transformer"]] = None):
        """x: (b, s, H) device tensor in D.  `batch` = [(layer_name, index_pos, block_idx)] for
        consecutive local layers (text_model.rs:298-321); all entries share index_pos."""
        blocks = blocks or [self]
        assert len(blocks) == len(batch)
        b, s, _ = x.shape
        with torch.cuda.stream(ctx.torch_stream):
            x = x.contiguous()
            y = torch.empty_like(x)
        cache = ctx.cache
        if cache is None:
            raise RuntimeError("No cache specified")  # transformer.rs:120 expect()
        check(lib().cake_b200_forward_batch(ctx.h, ptr_array([blk.h for blk in blocks]),
                                            int_array([bi for _, _, bi in batch]), len(blocks), cache.h,
                                            ptr(x), ptr(y), b, s, batch[0][1]))
        return y

    def forward_host(self, x_host: np.ndarray, index_pos: int, block_idx: int, ctx: Context) -> np.ndarray:
        """Forwarder::forward with HOST buffers (uint16 bit patterns of D), copies inside the call."""
        b, s, H = x_host.shape
        y = np.empty_like(x_host)
        check(lib().cake_b200_forward_batch_host(ctx.h, ptr_array([self.h]), int_array([block_idx]), 1, ctx.cache.h,
                                                 ptr(x_host), ptr(y), b, s, index_pos))
        return y

    def layer_name(self) -> str:
        return self.name

    def __str__(self):
        return f"{self.name} (local)"  # transformer.rs:72-76

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class Token:
    id: int
    text: Optional[str]
    is_end_of_stream: bool


class TextModelBase:
    """text_model.rs:133-530 for token-id prompts (tokenizer / chat template stay in the caller)."""

    def __init__(self, ctx: Context, blocks: List[Forwarder], repeat_penalty: float = 1.0, repeat_last_n: int = 128,
                 temperature: float = 0.0, top_k: Optional[int] = None, top_p: Optional[float] = None, seed: int = 299792458):
        """temperature / top_k / top_p / seed: the reference's Args that create_logits_processor reads
        (text_model.rs:102-118): temperature <= 0 -> ArgMax, else GumbelSoftmax | TopK | TopP | TopKThenTopP."""
        self.ctx, self.blocks = ctx, blocks
        self.sampling = sampling_from_args(temperature, top_k, top_p, seed)
        self.tokens: List[int] = []
        self.index_pos = 0
        self.generated = 0
        self.prompt_len = 0
        self.repeat_penalty, self.repeat_last_n = repeat_penalty, repeat_last_n
        self._graph_ready = False

    @classmethod
    def load(cls, ctx: Context, block_cls=B200Transformer, make_remote: Optional[Callable] = None, **kw):
        """text_model.rs:150-264: embed, lm_head (tied -> embed), ln_f, then one block per layer —
        local ones via Forwarder::load, layers the topology assigns elsewhere via `make_remote`."""
        cfg, vb, p = ctx.config, ctx.var_builder, ctx.config.model_prefix
        def get(tname: str, shape):
            t = vb[tname]
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"shape mismatch for {tname}, expected: {list(shape)}, got: {list(t.shape)}")
            return t.to(ctx.torch_dtype).contiguous()

        emb = get(f"{p}.embed_tokens.weight", (cfg.vocab_size, cfg.hidden_size))
        from .loader import rms_norm_weight
        lnf = rms_norm_weight(get(f"{p}.norm.weight", (cfg.hidden_size,)), cfg)   # text_model.rs:186-192
        head = None if cfg.tie_word_embeddings else get("lm_head.weight", (cfg.vocab_size, cfg.hidden_size))
        check(lib().cake_b200_head_load(ctx.h, ptr(emb), ptr(lnf), ptr(head)))
        blocks: List[Forwarder] = []
        for i in range(cfg.num_hidden_layers):
            name = cfg.layer_name(i)
            owner = _topology_owner(ctx.topology, name)
            if owner is None:
                blocks.append(block_cls.load(name, ctx))
            else:
                if make_remote is None:
                    raise RuntimeError(f"layer {name} is assigned to worker {owner!r} but no transport was given")
                blocks.append(make_remote(owner, name, ctx))
        return cls(ctx, blocks, **kw)

    # -- forward (text_model.rs:266-368) ---------------------------------------------------------
    def forward(self, ids, idx: int) -> torch.Tensor:
        """ids: (batch, seq) token ids at absolute position idx -> logits (batch, vocab) device tensor in D."""
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        b, s = ids.shape
        ctx = self.ctx
        x = ctx.empty(b, s, ctx.config.hidden_size)
        check(lib().cake_b200_embed(ctx.h, ids.ctypes.data_as(ctypes.POINTER(c_uint32)), b, s, ptr(x)))
        n, i = len(self.blocks), 0
        while i < n:
            ident = self.blocks[i].ident()
            j = i
            group = []
            while j < n and self.blocks[j].ident() == ident:  # contiguous run on the same node (:298-321)
                group.append((self.blocks[j].layer_name(), idx, j))
                j += 1
            if ident == "local":
                x = self.blocks[i].forward_batch(x, group, ctx, blocks=self.blocks[i:j])
            else:
                x = self.blocks[i].forward_batch(x, group, ctx)
            i = j
        logits = ctx.empty(b, ctx.config.vocab_size)
        check(lib().cake_b200_logits(ctx.h, ptr(x), b, s, ptr(logits), None))
        return logits

    # -- generation state (text_model.rs:371-530) ------------------------------------------------
    def prepare_prompt(self, token_ids: Sequence[int]) -> None:
        self.tokens = list(token_ids)
        self.ctx.cache.clear()
        self.index_pos = 0
        self.prompt_len = len(self.tokens)
        self._graph_ready = False

    def next_token(self, index: int) -> Token:
        if index > 0:
            ctx_tokens, ctx_index = self.tokens[-1:], self.index_pos
        else:
            ctx_tokens, ctx_index = self.tokens, 0
        logits = self.forward([ctx_tokens], ctx_index)[0]
        self.index_pos += len(ctx_tokens)
        out = c_uint32()
        if self.repeat_penalty == 1.0:
            pen = []
        else:  # text_model.rs:435-452: only generated tokens, last repeat_last_n of them
            gen = self.tokens[self.prompt_len:]
            pen = gen[max(0, len(gen) - self.repeat_last_n):]
        arr = (c_uint32 * max(1, len(pen)))(*pen)
        if self.sampling.kind == 0:
            check(lib().cake_b200_repeat_penalty_argmax(self.ctx.h, ptr(logits), self.repeat_penalty, arr, len(pen), byref(out)))
        else:  # the sampled token is drawn on the device: only 4 bytes come back (never the 256 KB of logits)
            check(lib().cake_b200_sample(self.ctx.h, ptr(logits), byref(self.sampling), self.repeat_penalty, arr, len(pen),
                                         self.generated, None, byref(out)))
        self.last_logits = logits
        tok = int(out.value)
        self.generated += 1
        self.tokens.append(tok)
        return Token(tok, None, self.ctx.config.is_eos(tok))

    def reset(self) -> None:
        self.tokens.clear()
        self.ctx.cache.clear()
        self.index_pos = self.generated = self.prompt_len = 0
        self._graph_ready = False

    def goodbye(self) -> None:
        for b in self.blocks:
            b.goodbye()

    # -- the graph-captured greedy decode loop (all layers local) ----------------------------------
    def local_handles(self):
        assert all(b.ident() == "local" for b in self.blocks)
        return ptr_array([b.h for b in self.blocks]), int_array(list(range(len(self.blocks))))

    def decode_build(self, rank: int = 0, world: int = 1, blocks: Optional[List[B200Transformer]] = None,
                     block_idx: Optional[List[int]] = None) -> None:
        if rank == 0 and (self.sampling.kind != 0 or getattr(self, "_sampling_set", False)):
            check(lib().cake_b200_decode_set_sampling(self.ctx.h, byref(self.sampling)))
            self._sampling_set = True
        if blocks is None:
            hs, idx = self.local_handles()
            n = len(self.blocks)
        else:
            hs, idx, n = ptr_array([b.h for b in blocks]), int_array(block_idx), len(blocks)
        check(lib().cake_b200_decode_build(self.ctx.h, hs, idx, n, self.ctx.cache.h, rank, world))
        self._graph_ready = True

    def decode_greedy(self, first_token: int, n_steps: int, sync: bool = True) -> List[int]:
        """n_steps greedy tokens fed back on the device (no host work between steps)."""
        L = lib()
        check(L.cake_b200_decode_begin(self.ctx.h, first_token, self.index_pos))
        check(L.cake_b200_decode_run(self.ctx.h, n_steps))
        self.index_pos += n_steps
        if not sync:
            return []
        out = (c_uint32 * n_steps)()
        check(L.cake_b200_decode_tokens(self.ctx.h, out, n_steps))
        toks = [int(t) for t in out]
        self.tokens.extend(toks)
        self.generated += n_steps
        return toks


def sampling_from_args(temperature: float, top_k: Optional[int], top_p: Optional[float], seed: int) -> "capi.CSampling":
    """text_model.rs:102-118 create_logits_processor."""
    if temperature is None or temperature <= 0.0:
        kind = 0
    elif top_k is None and top_p is None:
        kind = 5
    elif top_p is None:
        kind = 2
    elif top_k is None:
        kind = 3
    else:
        kind = 4
    return capi.CSampling(kind, int(top_k or 0), float(temperature or 0.0), float(top_p or 0.0), int(seed))


def _topology_owner(topology: dict, layer_name: str) -> Optional[str]:
    """topology.rs:44-56 Node::is_text_model_layer_owner over expanded layer lists."""
    for worker, node in (topology or {}).items():
        if layer_name in node.get("layers", []):
            return worker
    return None


class Master:
    """sharding/master.rs:14-191 for token-id prompts."""

    def __init__(self, model: TextModelBase):
        self.model = model

    def generate_text(self, prompt_ids: Sequence[int], sample_len: int,
                      stream: Optional[Callable[[Token], None]] = None) -> dict:
        """master.rs:109-168.  tok/s = (generated - 1) / time since the first token (:131-134,160-166)."""
        m = self.model
        m.prepare_prompt(prompt_ids)
        start = time.perf_counter()
        out: List[int] = []
        for index in range(sample_len):
            if index == 1:
                start = time.perf_counter()  # master.rs:132-134: timer restarts after the first token
            tok = m.next_token(index)
            if tok.is_end_of_stream:
                break
            out.append(tok.id)
            if stream:
                stream(tok)
        dt = time.perf_counter() - start
        n = m.generated
        return {"tokens": out, "generated": n, "tok_s": (n - 1) / dt if n > 1 and dt > 0 else 0.0}

    def goodbye(self):
        self.model.goodbye()
