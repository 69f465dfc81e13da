"""Host-side logic of the sharded path on CPU: topology/range parsing (reference: cake-core/src/cake/
sharding/topology.rs:13,134-172 and tests/unit_tests/test_topology.rs) and the master<->worker protocol
with world_size 2 over gloo (reference: tests/protocol.rs — two tasks over loopback; here two processes).
The CUDA blocks are replaced by a stand-in Forwarder so that no GPU is needed; the transport is GlooTransport."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cake_b200.config import llama3_8b, reference_test_config
from cake_b200.model import Forwarder, _topology_owner
from cake_b200.parallel import Client, GlooTransport, Worker, box_topology, expand_layers, layer_split, parse_topology


def test_layer_range_expansion():
    assert expand_layers(["model.layers.0-2"]) == ["model.layers.0", "model.layers.1", "model.layers.2"]
    assert expand_layers(["model.layers.7", "model.layers.9-9"]) == ["model.layers.7", "model.layers.9"]
    assert expand_layers(["model.language_model.layers.10-11"]) == ["model.language_model.layers.10", "model.language_model.layers.11"]
    with pytest.raises(ValueError, match="end must be >= start"):
        expand_layers(["model.layers.5-2"])


def test_parse_topology_and_ownership():
    topo = parse_topology({"w1": {"host": "10.0.0.2:10128", "layers": ["model.layers.4-7"]},
                           "w2": {"host": "10.0.0.3:10128", "layers": ["model.layers.8-8"]}})
    assert len(topo["w1"]["layers"]) == 4
    assert _topology_owner(topo, "model.layers.5") == "w1"
    assert _topology_owner(topo, "model.layers.8") == "w2"
    assert _topology_owner(topo, "model.layers.0") is None           # unassigned layers stay on the master
    assert _topology_owner(topo, "model.layers.80") is None          # "model.layers.8" must not match "…80"


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_layer_split_is_contiguous_and_complete(world):
    cfg = llama3_8b()
    parts = layer_split(cfg.num_hidden_layers, world)
    assert [i for r in parts for i in r] == list(range(cfg.num_hidden_layers))
    assert max(len(r) for r in parts) - min(len(r) for r in parts) <= 1
    topo = box_topology(cfg, world)
    assert set(topo) == {f"gpu{r}" for r in range(1, world)}
    for r in range(1, world):
        assert topo[f"gpu{r}"]["layers"] == [f"model.layers.{i}" for i in parts[r]]


def test_70b_split_is_ten_layers_per_gpu():
    from cake_b200.config import llama3_70b
    assert [len(r) for r in layer_split(llama3_70b().num_hidden_layers, 8)] == [10] * 8


# ---- world_size 2 over gloo -----------------------------------------------------------------------
class _Cache:
    def __init__(self):
        self.cleared = 0

    def clear(self):
        self.cleared += 1


class _Ctx:
    """Context stand-in: config + cache + empty()."""

    def __init__(self, cfg):
        self.config, self.cache, self.topology = cfg, _Cache(), {}

    def empty(self, *shape):
        return torch.empty(shape, dtype=torch.float32)

    def sync(self):
        pass


class _AffineBlock(Forwarder):
    """y = 2x + (layer index + 1): order-sensitive, so a wrong layer order or a skipped layer is caught."""

    def __init__(self, name):
        self.name, self.idx = name, int(name.rsplit(".", 1)[1])

    @classmethod
    def load(cls, name, ctx):
        return cls(name)

    def forward(self, x, index_pos, block_idx, ctx):
        assert block_idx == self.idx
        return 2 * x + (self.idx + 1)

    def forward_batch(self, x, batch, ctx, blocks=None):
        for (name, pos, idx), blk in zip(batch, blocks or [self]):
            assert name == blk.name
            x = blk.forward(x, pos, idx, ctx)
        return x

    def layer_name(self):
        return self.name


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = reference_test_config()          # 4 layers -> rank 0 owns 0-1, rank 1 owns 2-3
    ctx = _Ctx(cfg)
    tr = GlooTransport()
    try:
        if rank == 1:
            w = Worker(ctx, 1, world, block_cls=_AffineBlock, transport=tr)
            assert w.names == ["model.layers.2", "model.layers.3"]
            w.serve()
            q.put(("worker", w.served, ctx.cache.cleared))
        else:
            ctx.topology = box_topology(cfg, world)
            blocks = []
            for i in range(cfg.num_hidden_layers):
                name = cfg.layer_name(i)
                owner = _topology_owner(ctx.topology, name)
                blocks.append(_AffineBlock.load(name, ctx) if owner is None else Client(owner, name, ctx, transport=tr))
            assert [b.ident() for b in blocks] == ["local", "local", "gpu1", "gpu1"]

            def forward(x, pos):  # the block walk of text_model.rs:284-332
                i = 0
                while i < len(blocks):
                    j = i
                    while j < len(blocks) and blocks[j].ident() == blocks[i].ident():
                        j += 1
                    batch = [(blocks[k].layer_name(), pos, k) for k in range(i, j)]
                    x = blocks[i].forward_batch(x, batch, ctx, blocks=blocks[i:j]) if blocks[i].ident() == "local" \
                        else blocks[i].forward_batch(x, batch, ctx)
                    i = j
                return x

            x = torch.arange(8, dtype=torch.float32).reshape(1, 1, 8)
            ref = x
            for i in range(4):
                ref = 2 * ref + (i + 1)
            outs = [forward(x, p) for p in range(3)]                  # 3 sequential ops (protocol.rs: 10 sequential)
            big = forward(torch.ones(1, 64, 8), 3)                     # "large tensor" case
            for b in blocks:                                           # TextModelBase::goodbye (text_model.rs:517-528):
                b.goodbye()                                            # one Goodbye per worker, sent by its first Client
            after = forward(x, 0)
            # worker.rs:490-503 / protocol.rs "layer not found": the worker reports the error and keeps serving
            from cake_b200.parallel import WorkerError
            ghost = Client("gpu1", "model.layers.99", ctx, transport=tr)
            try:
                ghost.forward_batch(x, [("model.layers.99", 0, 99)], ctx)
                reported = False
            except WorkerError as e:
                reported = "model.layers.99" in str(e)
            after2 = forward(x, 1)
            dist.broadcast_object_list([("shutdown",)], src=0)
            q.put(("master", reported and all(torch.equal(o, ref) for o in outs + [after, after2]), tuple(big.shape)))
    finally:
        dist.destroy_process_group()


def test_master_worker_protocol_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict()
    for _ in range(2):
        item = q.get(timeout=120)
        res[item[0]] = item[1:]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["master"] == (True, (1, 64, 8))
    assert res["worker"] == (6, 1)   # 6 batches served (the failed one is not counted), cache cleared once by Goodbye


# ---- the same protocol with the real block class over the emulated library -------------------------------------------
_REAL_KINDS = {
    "llama": dict(),
    # a sibling block structure through the same master / worker plumbing: per-layer variants are set on whichever rank
    # loads the layer (gemma3/block.rs: local layers 0, 2 = window 5 without RoPE; global layers 1, 3 = full context + RoPE)
    "gemma3": dict(block_kind="gemma3", use_qk_norm=True, residual_rms_norm=True, use_gelu_mlp=True, tie_word_embeddings=True,
                   embed_scale=11.313708498984761, sliding_window=5, global_layers=[False, True, False, True]),
}


def _real_cfg_sd(kind):
    from cake_b200.synth import residual_deltas
    from tests.util import checkpoint, medium_config
    cfg = medium_config(num_hidden_layers=4, hidden_size=128, intermediate_size=256, vocab_size=256, num_attention_heads=4,
                        num_key_value_heads=2, head_dim=32, **_REAL_KINDS[kind])
    sd = checkpoint(cfg, "bf16", seed=23, peaked=not cfg.tie_word_embeddings)
    return cfg, (residual_deltas(sd) if cfg.residual_rms_norm else sd)


def _run_real_blocks(rank, world, port, q, so_path, kind="llama"):
    """rank 0: TextModelBase whose layers 2-3 are `Client`s (GlooTransport); rank 1: `Worker` with B200Transformer blocks.
    The C ABI behind both is the oracle-backed emulation (tests/fake_b200), contexts are tests/cpu_ctx.CpuContext."""
    import numpy as np
    from cake_b200 import capi
    from cake_b200.model import B200Transformer, Master, TextModelBase
    from tests.cpu_ctx import CpuContext
    from tests.util import checkpoint, medium_config
    capi.SO_PATH, capi._lib = so_path, None
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, sd = _real_cfg_sd(kind)
    ctx = CpuContext(cfg, sd, "bf16", max_seq=64)
    tr = GlooTransport()
    try:
        if rank == 1:
            w = Worker(ctx, 1, world, block_cls=B200Transformer, transport=tr)
            w.serve()
            q.put(("worker", w.served, ctx.cache.len(2)))
        else:
            ctx.topology = box_topology(cfg, world)
            model = TextModelBase.load(ctx, make_remote=lambda wk, name, c: Client(wk, name, c, transport=tr))
            assert [b.ident() for b in model.blocks] == ["local", "local", "gpu1", "gpu1"]
            prompt = np.random.default_rng(9).integers(0, cfg.vocab_size - 1, 7).tolist()
            toks = Master(model).generate_text(prompt, 6)["tokens"]
            dist.broadcast_object_list([("goodbye",)], src=0)
            ctx.cache.clear()
            toks2 = Master(model).generate_text(prompt[:4], 3)["tokens"]
            dist.broadcast_object_list([("shutdown",)], src=0)
            q.put(("master", toks, toks2, prompt))
    finally:
        dist.destroy_process_group()
        ctx.close()


@pytest.mark.parametrize("kind", list(_REAL_KINDS))
def test_sharded_master_and_worker_with_real_blocks_over_gloo(tmp_path, kind):
    from oracle import oracle as O
    from tests.fake_b200.make_fake import build as build_fake
    from tests.util import checkpoint, medium_config
    so = build_fake(str(tmp_path), oracle=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run_real_blocks, args=(r, 2, port, q, so, kind)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict()
    for _ in range(2):
        item = q.get(timeout=180)
        res[item[0]] = item[1:]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    toks, toks2, prompt = res["master"]
    cfg, sd = _real_cfg_sd(kind)
    om = O.OracleModel(cfg, sd, "bf16", max_seq=64)
    assert toks == list(om.generate(prompt, 6)[0]) and toks2 == list(om.generate(prompt[:4], 3)[0])
    assert res["worker"] == (6 + 3, 4 + 2)   # one batch per forward; the worker's cache holds the second prompt: 4 + 2 fed-back tokens
