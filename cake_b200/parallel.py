"""Layer sharding across the GPUs of one NVSwitch box: the reference's master<->worker roles with the
tensor hand-off carried by in-stream NCCL send/recv instead of TCP.

  Topology / layer ranges   cake-core/src/cake/sharding/topology.rs:13,119-172  ("model.layers.0-5")
  default contiguous split  cake/sharding/default.rs:10-172 (here: equal split, the GPUs are identical)
  Client (remote block)     cake/sharding/client.rs:13-188   (forward_batch ships x, gets x back)
  Worker loop               cake/sharding/worker.rs:358-575  (Batch -> run its blocks -> Tensor)

One process per GPU (torchrun).  torch.distributed is used for rendezvous and for the tiny control
messages (the `(seq, index_pos)` fields of the reference's Batch message and Goodbye); activations never
touch the host.
"""
from __future__ import annotations

import os
import re
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .capi import check, lib, ptr
from .config import Config
from .model import B200Transformer, Context, Forwarder, TextModelBase

_LAYER_RANGE = re.compile(r"^(.+[^\d])(\d+)-(\d+)$", re.M)  # topology.rs:13


def expand_layers(layers: Sequence[str]) -> List[str]:
    """topology.rs:143-168: 'model.layers.0-5' -> model.layers.0 .. model.layers.5 (inclusive)."""
    out: List[str] = []
    for name in layers:
        m = _LAYER_RANGE.match(name)
        if m:
            base, start, stop = m.group(1), int(m.group(2)), int(m.group(3))
            if stop < start:
                raise ValueError(f"invalid range expression {name}, end must be >= start")
            out.extend(f"{base}{n}" for n in range(start, stop + 1))
        else:
            out.append(name)
    return out


def parse_topology(topo: Dict[str, dict]) -> Dict[str, dict]:
    """Topology::from_path semantics on an already-loaded YAML mapping: name -> {host, layers[...]}."""
    return {name: {**node, "layers": expand_layers(node.get("layers", []))} for name, node in topo.items()}


def load_topology(path: str) -> Dict[str, dict]:
    import yaml
    with open(path) as f:
        return parse_topology(yaml.safe_load(f))


def layer_split(n_layers: int, world: int) -> List[range]:
    """Contiguous equal split (rank r owns [r*n/N, (r+1)*n/N)); rank 0 = master, also owns embed/head."""
    return [range(r * n_layers // world, (r + 1) * n_layers // world) for r in range(world)]


def box_topology(cfg: Config, world: int) -> Dict[str, dict]:
    """The reference-style topology of an N-GPU box: ranks 1..N-1 are workers; unassigned layers stay on
    the master (cake/mod.rs:385-392)."""
    topo = {}
    for r, rg in enumerate(layer_split(cfg.num_hidden_layers, world)):
        if r == 0 or len(rg) == 0:
            continue
        topo[f"gpu{r}"] = {"host": f"nvlink://{r}", "layers": [f"{cfg.model_prefix}.layers.{rg.start}-{rg.stop - 1}"]}
    return parse_topology(topo)


def rank_of(worker: str) -> int:
    return int(worker.replace("gpu", ""))


# ------------------------------------------------------------------------------------------------
def init_comm(ctx: Context, rank: int, world: int) -> None:
    """Create the library's NCCL communicator; the 128-byte unique id travels over torch.distributed."""
    import torch.distributed as dist
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        check(lib().cake_b200_comm_unique_id(ptr(uid)))
    if dist.get_backend() == "nccl":
        d = uid.cuda(ctx.device)
        dist.broadcast(d, 0)
        torch.cuda.synchronize()
        uid = d.cpu()
    else:
        dist.broadcast(uid, 0)
    check(lib().cake_b200_comm_init(ctx.h, ptr(uid), rank, world))
    if os.environ.get("CAKE_B200_RING", "p2p") == "p2p":
        # fused hand-off: exchange the IPC handles of the inboxes; rank r writes the inbox of rank (r+1) % world
        mine = torch.zeros(64, dtype=torch.uint8)
        check(lib().cake_b200_ring_export(ctx.h, ptr(mine)))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(mine.numpy().tobytes()))
        nxt = torch.frombuffer(bytearray(handles[(rank + 1) % world]), dtype=torch.uint8)
        check(lib().cake_b200_ring_import(ctx.h, ptr(nxt)))
        dist.barrier()


class NcclTransport:
    """Activation hand-off through the library's communicator: in-stream ncclSend/ncclRecv, no host copy."""

    def __init__(self, ctx: Context):
        self.ctx = ctx

    def empty_like(self, x):
        with torch.cuda.stream(self.ctx.torch_stream):
            return torch.empty_like(x)

    def send(self, x: torch.Tensor, peer: int) -> None:
        check(lib().cake_b200_send(self.ctx.h, ptr(x), x.numel() * x.element_size(), peer))

    def recv(self, y: torch.Tensor, peer: int) -> None:
        check(lib().cake_b200_recv(self.ctx.h, ptr(y), y.numel() * y.element_size(), peer))


class GlooTransport:
    """CPU stand-in used by the world_size-2 gloo tests of the host logic (tests/test_parallel_cpu.py)."""

    def empty_like(self, x):
        return torch.empty_like(x)

    def send(self, x, peer):
        import torch.distributed as dist
        dist.send(x.contiguous(), peer)

    def recv(self, y, peer):
        import torch.distributed as dist
        dist.recv(y, peer)


class WorkerError(RuntimeError):
    """Message::WorkerError (message.rs:240-247) surfaced on the master."""


class Client(Forwarder):
    """A block that lives on another GPU (client.rs:13).  forward_batch ships the activation to the
    worker rank and receives the result — NCCL send/recv on the ctx stream, no host copy."""

    def __init__(self, worker: str, name: str, ctx, transport=None):
        self.worker, self.name, self.ctx, self.peer = worker, name, ctx, rank_of(worker)
        self.transport = transport or NcclTransport(ctx)
        seen = ctx.__dict__.setdefault("_client_workers", set())
        self._primary = worker not in seen
        seen.add(worker)

    @classmethod
    def load(cls, name, ctx):  # pragma: no cover - constructed through TextModelBase.load(make_remote=...)
        raise NotImplementedError

    def forward(self, x, index_pos, block_idx, ctx):
        return self.forward_batch(x, [(self.name, index_pos, block_idx)], ctx)

    def forward_batch(self, x, batch, ctx):
        import torch.distributed as dist
        b, s, _ = x.shape
        # control header == the (layer_name, index_pos, block_idx) list of Message::Batch (message.rs:191-247)
        dist.broadcast_object_list([("batch", self.peer, b, s, list(x.shape), batch)], src=0)
        x = x.contiguous()
        y = self.transport.empty_like(x)
        self.transport.send(x, self.peer)
        self.transport.recv(y, self.peer)
        # the worker always answers: the tensor, then a status (None | message) == Message::Tensor | WorkerError
        # (worker.rs:490-520); an error never tears the session down (client.rs:120-135 turns it into Err)
        status = [None]
        dist.broadcast_object_list(status, src=self.peer)
        if status[0] is not None:
            raise WorkerError(f"worker {self.worker}: {status[0]}")
        return y

    def goodbye(self):
        """client.rs:176-187: Goodbye clears the worker's KV cache.  The reference keeps one connection per remote
        layer and says goodbye on each; here a worker has one session, so only the first Client of a worker sends it."""
        import torch.distributed as dist
        if self._primary:
            dist.broadcast_object_list([("goodbye", self.peer)], src=0)
        return None

    def layer_name(self):
        return self.name

    def ident(self):
        return self.worker

    def __str__(self):
        return f"{self.name}@{self.worker}"


class Worker:
    """worker.rs:79-597 on rank > 0: owns the blocks of its layer range and a per-session cache."""

    def __init__(self, ctx, rank: int, world: int, block_cls=B200Transformer, transport=None):
        self.ctx, self.rank, self.world = ctx, rank, world
        names = box_topology(ctx.config, world).get(f"gpu{rank}", {"layers": []})["layers"]
        self.blocks: Dict[str, Forwarder] = {n: block_cls.load(n, ctx) for n in names}
        self.names = names
        self.transport = transport or NcclTransport(ctx)
        self.served = 0

    def block_list(self) -> Tuple[List[B200Transformer], List[int]]:
        blks = [self.blocks[n] for n in self.names]
        return blks, [int(n.rsplit(".", 1)[1]) for n in self.names]

    def serve(self) -> None:
        """Handle control messages until 'shutdown' (worker.rs:358-575 loop)."""
        import torch.distributed as dist
        ctx = self.ctx
        while True:
            msg = [None]
            dist.broadcast_object_list(msg, src=0)
            op = msg[0]
            if op[0] == "shutdown":
                return
            if op[0] == "goodbye":  # worker.rs:364-371 (addressed to one worker, or to all when no rank is given)
                if len(op) < 2 or op[1] is None or op[1] == self.rank:
                    ctx.cache.clear()
            elif op[0] == "batch":
                _, peer, b, s, shape, batch = op
                if peer != self.rank:
                    dist.broadcast_object_list([None], src=peer)  # the addressed worker's status message
                    continue
                x = ctx.empty(*shape)
                self.transport.recv(x, 0)
                err = None
                try:  # worker.rs:490-503: any failure is reported as WorkerError and the worker keeps serving
                    missing = [name for name, _, _ in batch if name not in self.blocks]
                    if missing:
                        raise KeyError(f"could not find layer {missing[0]}")
                    blks = [self.blocks[name] for name, _, _ in batch]
                    y = blks[0].forward_batch(x, batch, ctx, blocks=blks)
                except Exception as e:  # noqa: BLE001 - reported to the master verbatim
                    err = f"{type(e).__name__}: {e}"
                    y = x  # payload of the error reply is ignored by the client
                self.transport.send(y, 0)
                dist.broadcast_object_list([err], src=self.rank)
                if err is None:
                    self.served += 1
            elif op[0] == "decode":  # graph-captured ring decode: n steps without further control traffic
                _, n_steps, index_pos = op
                blks, idx = self.block_list()
                if not getattr(self, "_graph", False):
                    from .capi import int_array, ptr_array
                    check(lib().cake_b200_decode_build(ctx.h, ptr_array([b.h for b in blks]), int_array(idx), len(blks),
                                                       ctx.cache.h, self.rank, self.world))
                    self._graph = True
                check(lib().cake_b200_decode_begin(ctx.h, 0, index_pos))
                check(lib().cake_b200_decode_run(ctx.h, n_steps))
                ctx.sync()
            elif op[0] == "fill":  # bench: synthetic cache fill on every shard
                _, length, seed = op
                _, idx = self.block_list()
                ctx.cache.fill_synthetic(idx, length, seed)
                ctx.sync()


class ShardedMaster:
    """Rank 0: TextModelBase whose non-local layers are Clients (text_model.rs:204-227), plus the ring
    decode loop."""

    def __init__(self, ctx: Context, world: int, **kw):
        import torch.distributed as dist  # noqa: F401
        self.ctx, self.world = ctx, world
        ctx.topology = box_topology(ctx.config, world)
        self.model = TextModelBase.load(ctx, make_remote=lambda w, n, c: Client(w, n, c), **kw)
        self.local = [b for b in self.model.blocks if b.ident() == "local"]
        self.local_idx = [i for i, b in enumerate(self.model.blocks) if b.ident() == "local"]

    def fill_synthetic(self, length: int, seed: int = 7) -> None:
        import torch.distributed as dist
        dist.broadcast_object_list([("fill", length, seed)], src=0)
        self.ctx.cache.fill_synthetic(self.local_idx, length, seed)
        self.ctx.sync()
        self.model.index_pos = length

    def decode_build(self) -> None:
        self.model.decode_build(0, self.world, blocks=self.local, block_idx=self.local_idx)

    def decode_greedy(self, first_token: int, n_steps: int) -> List[int]:
        import torch.distributed as dist
        dist.broadcast_object_list([("decode", n_steps, self.model.index_pos)], src=0)
        return self.model.decode_greedy(first_token, n_steps)

    def goodbye(self) -> None:
        import torch.distributed as dist
        dist.broadcast_object_list([("goodbye", None)], src=0)
        self.ctx.cache.clear()

    def shutdown(self) -> None:
        import torch.distributed as dist
        dist.broadcast_object_list([("shutdown",)], src=0)
