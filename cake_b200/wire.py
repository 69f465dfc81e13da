"""cake's TCP wire protocol, so that a B200 box can serve an unmodified remote cake master.

What is mirrored (reference file:line):

  proto/mod.rs:3-10            PROTO_MAGIC = 0x0104F4C7, MESSAGE_MAX_SIZE = 512 MiB
  proto/message.rs:7-36        dtype <-> u8 tag
  proto/message.rs:39-48       RawTensor { data: Vec<u8>, dtype: u8, shape: Vec<usize> }
  proto/message.rs:171-188     WorkerInfo
  proto/message.rs:190-247     enum Message (variant order = wire tag)
  proto/message.rs:334-394     framing: 8-byte header (magic u32 | payload length u32, big endian) + payload
  auth.rs:1-118                mutual HMAC-SHA256 challenge-response before any framing
  worker.rs:298-575            Worker side of one master connection (Hello -> WorkerInfo, SingleOp / Batch ->
                               Tensor | WorkerError, Goodbye -> cache clear + WorkerInfo)
  client.rs:24-188             Client (master side), a `Forwarder`

The payload encoding is the `speedy` crate's (pinned 0.8 in the reference's Cargo.lock; not vendored) with the
`BigEndian` context, restated from its published format: integers big endian; `usize` as u64; `u128` as 16 bytes;
`bool` as u8; `String`/`Vec<T>` as a u32 element count followed by the elements; tuples and struct fields in
declaration order; enum variants as a u32 tag in declaration order followed by the variant's fields.

The compute behind a `WireWorker` is a backend object (``B200Backend`` = the C ABI's ``cake_b200_forward_batch_host``;
tests plug a CPU stand-in); nothing in this module touches CUDA itself.
"""
from __future__ import annotations

import hashlib
import hmac
import os
import platform
import socket
import struct
import threading
import time
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

PROTO_MAGIC = 0x0104F4C7
MESSAGE_MAX_SIZE = 512 * 1024 * 1024
NONCE_SIZE = HMAC_SIZE = 32

# proto/message.rs:7-36
DTYPE_TAGS = {"u8": 0, "u32": 1, "i64": 2, "bf16": 3, "f16": 4, "f32": 5, "f64": 6, "f8e4m3": 7}
TAG_DTYPES = {v: k for k, v in DTYPE_TAGS.items()}
DTYPE_SIZES = {"u8": 1, "u32": 4, "i64": 8, "bf16": 2, "f16": 2, "f32": 4, "f64": 8, "f8e4m3": 1}


class ProtocolError(RuntimeError):
    pass


# ------------------------------------------------------------------------------------------ speedy (BigEndian)
class _Writer:
    def __init__(self):
        self.parts: List[bytes] = []

    def u8(self, v):
        self.parts.append(struct.pack(">B", v))

    def u32(self, v):
        self.parts.append(struct.pack(">I", v))

    def u64(self, v):
        self.parts.append(struct.pack(">Q", v))

    def u128(self, v):
        self.parts.append(int(v).to_bytes(16, "big"))

    def boolean(self, v):
        self.u8(1 if v else 0)

    def blob(self, b):
        self.u32(len(b))
        self.parts.append(bytes(b))

    def string(self, s: str):
        self.blob(s.encode("utf-8"))

    def bytes(self) -> bytes:
        return b"".join(self.parts)


class _Reader:
    def __init__(self, buf: bytes):
        self.buf, self.pos = memoryview(buf), 0

    def _take(self, n: int) -> memoryview:
        if n < 0 or self.pos + n > len(self.buf):
            raise ProtocolError(f"truncated message: need {n} bytes at offset {self.pos}, have {len(self.buf) - self.pos}")
        v = self.buf[self.pos:self.pos + n]
        self.pos += n
        return v

    def u8(self):
        return self._take(1)[0]

    def u32(self):
        return struct.unpack(">I", self._take(4))[0]

    def u64(self):
        return struct.unpack(">Q", self._take(8))[0]

    def u128(self):
        return int.from_bytes(self._take(16), "big")

    def boolean(self):
        return self.u8() != 0

    def blob(self) -> bytes:
        return bytes(self._take(self.u32()))

    def string(self) -> str:
        try:
            return self.blob().decode("utf-8")
        except UnicodeDecodeError as e:
            raise ProtocolError(f"invalid utf-8 in string: {e}") from e

    def done(self):
        if self.pos != len(self.buf):
            raise ProtocolError(f"{len(self.buf) - self.pos} trailing bytes after message")


# ------------------------------------------------------------------------------------------ RawTensor
@dataclass
class RawTensor:
    """proto/message.rs:39-48.  ``data`` holds the elements in native (little-endian) byte order, row-major."""
    data: bytes
    dtype: int
    shape: List[int]

    @property
    def dtype_name(self) -> str:
        if self.dtype not in TAG_DTYPES:
            raise ProtocolError(f"unknown dtype tag: {self.dtype}")  # message.rs:34
        return TAG_DTYPES[self.dtype]

    def numel(self) -> int:
        n = 1
        for d in self.shape:
            n *= d
        return n

    def validate(self) -> "RawTensor":
        if len(self.data) != self.numel() * DTYPE_SIZES[self.dtype_name]:
            raise ProtocolError(f"tensor of shape {self.shape} and dtype {self.dtype_name} cannot have {len(self.data)} bytes")
        return self

    @staticmethod
    def from_numpy_bits(a: np.ndarray, dtype: str) -> "RawTensor":
        """``a`` carries the elements' bit patterns (uint16 for bf16/f16, or the natural numpy dtype)."""
        a = np.ascontiguousarray(a)
        if a.dtype.itemsize != DTYPE_SIZES[dtype]:
            raise ValueError(f"{a.dtype} does not carry {dtype} elements")
        return RawTensor(a.tobytes(), DTYPE_TAGS[dtype], list(a.shape))

    def to_numpy_bits(self) -> np.ndarray:
        self.validate()
        npdt = {"u8": np.uint8, "u32": np.uint32, "i64": np.int64, "bf16": np.uint16, "f16": np.uint16,
                "f32": np.float32, "f64": np.float64, "f8e4m3": np.uint8}[self.dtype_name]
        return np.frombuffer(self.data, dtype=npdt).reshape(self.shape)

    def _write(self, w: _Writer):
        w.blob(self.data)
        w.u8(self.dtype)
        w.u32(len(self.shape))
        for d in self.shape:
            w.u64(d)

    @staticmethod
    def _read(r: _Reader) -> "RawTensor":
        data, dtype = r.blob(), r.u8()
        return RawTensor(data, dtype, [r.u64() for _ in range(r.u32())])


# ------------------------------------------------------------------------------------------ messages
@dataclass
class WorkerInfo:  # proto/message.rs:171-188
    version: str = ""
    dtype: str = ""
    os: str = ""
    arch: str = ""
    device: str = ""
    device_idx: int = 0
    latency: int = 0

    def _write(self, w: _Writer):
        for s in (self.version, self.dtype, self.os, self.arch, self.device):
            w.string(s)
        w.u64(self.device_idx)
        w.u128(self.latency)

    @staticmethod
    def _read(r: _Reader) -> "WorkerInfo":
        return WorkerInfo(r.string(), r.string(), r.string(), r.string(), r.string(), r.u64(), r.u128())


Op = Tuple[str, int, int]  # (layer_name, index_pos, block_idx)


@dataclass
class Message:
    """proto/message.rs:190-247.  ``kind`` is the variant name; the fields a variant does not have stay None."""
    kind: str
    info: Optional[WorkerInfo] = None            # WorkerInfo
    x: Optional[RawTensor] = None                # SingleOp, Batch, Tensor
    layer_name: Optional[str] = None             # SingleOp
    index_pos: int = 0                           # SingleOp
    block_idx: int = 0                           # SingleOp
    batch: Optional[List[Op]] = None             # Batch
    layers: Optional[List[str]] = None           # LayerAssignment
    model_hash: str = ""                         # LayerAssignment
    needs_data: bool = False                     # LayerAssignmentAck
    filename: str = ""                           # ModelDataChunk, ModelDataResume
    offset: int = 0                              # ModelDataChunk, ModelDataResume
    total_size: int = 0                          # ModelDataChunk
    compressed: bool = False                     # ModelDataChunk
    checksum: int = 0                            # ModelDataChunk
    data: bytes = b""                            # ModelDataChunk
    message: str = ""                            # WorkerError

    KINDS = ("Hello", "WorkerInfo", "SingleOp", "Batch", "Tensor", "Goodbye", "LayerAssignment",
             "LayerAssignmentAck", "ModelDataChunk", "ModelDataDone", "ModelDataResume", "WorkerReady", "WorkerError")

    # -- constructors (message.rs:249-283) --------------------------------------------------------
    @staticmethod
    def hello() -> "Message":
        return Message("Hello")

    @staticmethod
    def goodbye() -> "Message":
        return Message("Goodbye")

    @staticmethod
    def single_op(layer_name: str, x: RawTensor, index_pos: int, block_idx: int) -> "Message":
        return Message("SingleOp", layer_name=layer_name, x=x, index_pos=index_pos, block_idx=block_idx)

    @staticmethod
    def from_batch(x: RawTensor, batch: Sequence[Op]) -> "Message":
        return Message("Batch", x=x, batch=[(str(n), int(p), int(i)) for n, p, i in batch])

    @staticmethod
    def from_tensor(x: RawTensor) -> "Message":
        return Message("Tensor", x=x)

    @staticmethod
    def worker_error(message: str) -> "Message":
        return Message("WorkerError", message=message)

    # -- payload (message.rs:287-331) ---------------------------------------------------------------
    def to_bytes(self) -> bytes:
        w = _Writer()
        try:
            tag = self.KINDS.index(self.kind)
        except ValueError:
            raise ProtocolError(f"unknown message kind {self.kind!r}") from None
        w.u32(tag)
        k = self.kind
        if k == "WorkerInfo":
            self.info._write(w)
        elif k == "SingleOp":
            w.string(self.layer_name)
            self.x._write(w)
            w.u64(self.index_pos)
            w.u64(self.block_idx)
        elif k == "Batch":
            self.x._write(w)
            w.u32(len(self.batch))
            for name, pos, idx in self.batch:
                w.string(name)
                w.u64(pos)
                w.u64(idx)
        elif k == "Tensor":
            self.x._write(w)
        elif k == "LayerAssignment":
            w.u32(len(self.layers))
            for name in self.layers:
                w.string(name)
            w.string(self.model_hash)
        elif k == "LayerAssignmentAck":
            w.boolean(self.needs_data)
        elif k == "ModelDataChunk":
            w.string(self.filename)
            w.u64(self.offset)
            w.u64(self.total_size)
            w.boolean(self.compressed)
            w.u32(self.checksum)
            w.blob(self.data)
        elif k == "ModelDataResume":
            w.string(self.filename)
            w.u64(self.offset)
        elif k == "WorkerError":
            w.string(self.message)
        return w.bytes()

    @staticmethod
    def from_bytes(raw: bytes) -> "Message":
        r = _Reader(raw)
        tag = r.u32()
        if tag >= len(Message.KINDS):
            raise ProtocolError(f"unknown message tag {tag}")
        k = Message.KINDS[tag]
        m = Message(k)
        if k == "WorkerInfo":
            m.info = WorkerInfo._read(r)
        elif k == "SingleOp":
            m.layer_name = r.string()
            m.x = RawTensor._read(r)
            m.index_pos, m.block_idx = r.u64(), r.u64()
        elif k == "Batch":
            m.x = RawTensor._read(r)
            m.batch = [(r.string(), r.u64(), r.u64()) for _ in range(r.u32())]
        elif k == "Tensor":
            m.x = RawTensor._read(r)
        elif k == "LayerAssignment":
            m.layers = [r.string() for _ in range(r.u32())]
            m.model_hash = r.string()
        elif k == "LayerAssignmentAck":
            m.needs_data = r.boolean()
        elif k == "ModelDataChunk":
            m.filename, m.offset, m.total_size = r.string(), r.u64(), r.u64()
            m.compressed, m.checksum, m.data = r.boolean(), r.u32(), r.blob()
        elif k == "ModelDataResume":
            m.filename, m.offset = r.string(), r.u64()
        elif k == "WorkerError":
            m.message = r.string()
        r.done()
        return m

    # -- framing (message.rs:334-394) -----------------------------------------------------------------
    def frame(self) -> bytes:
        payload = self.to_bytes()
        if len(payload) > MESSAGE_MAX_SIZE:
            raise ProtocolError(f"request size {len(payload)} > MESSAGE_MAX_SIZE")
        return struct.pack(">II", PROTO_MAGIC, len(payload)) + payload

    def to_writer(self, sock: socket.socket) -> int:
        buf = self.frame()
        sock.sendall(buf)
        return len(buf)

    @staticmethod
    def from_reader(sock: socket.socket) -> Tuple[int, "Message"]:
        magic, size = struct.unpack(">II", _recv_exact(sock, 8))
        if magic != PROTO_MAGIC:
            raise ProtocolError(f"invalid magic value: {magic}")
        if size > MESSAGE_MAX_SIZE:
            raise ProtocolError(f"request size {size} > MESSAGE_MAX_SIZE")
        return size, Message.from_bytes(_recv_exact(sock, size))


def _recv_exact(sock: socket.socket, n: int) -> bytes:
    buf = bytearray(n)
    view, got = memoryview(buf), 0
    while got < n:
        k = sock.recv_into(view[got:], n - got)
        if k == 0:
            raise ConnectionError("connection closed by peer")
        got += k
    return bytes(buf)


# ------------------------------------------------------------------------------------------ auth.rs
def compute_hmac(key: bytes, data: bytes) -> bytes:
    return hmac.new(key, data, hashlib.sha256).digest()


def authenticate_as_master(sock: socket.socket, key: str) -> None:
    """auth.rs:54-84."""
    kb = key.encode()
    nonce = os.urandom(NONCE_SIZE)
    sock.sendall(nonce)
    resp = _recv_exact(sock, HMAC_SIZE + NONCE_SIZE)
    if not hmac.compare_digest(resp[:HMAC_SIZE], compute_hmac(kb, nonce)):
        raise ProtocolError("worker authentication failed: invalid HMAC")
    sock.sendall(compute_hmac(kb, resp[HMAC_SIZE:]))


def authenticate_as_worker(sock: socket.socket, key: str) -> None:
    """auth.rs:89-118."""
    kb = key.encode()
    master_nonce = _recv_exact(sock, NONCE_SIZE)
    nonce = os.urandom(NONCE_SIZE)
    sock.sendall(compute_hmac(kb, master_nonce) + nonce)
    if not hmac.compare_digest(_recv_exact(sock, HMAC_SIZE), compute_hmac(kb, nonce)):
        raise ProtocolError("master authentication failed: invalid HMAC")


# ------------------------------------------------------------------------------------------ worker side
class B200Backend:
    """The compute behind a WireWorker on a B200: the worker's blocks through ``cake_b200_forward_batch_host``
    (host buffers in, host buffers out; the H2D/D2H copies happen inside the C ABI call).

    Every master connection gets its own session = its own KV cache (the reference clones a fresh cache per
    connection, worker.rs:60-75 — and a cake master opens one connection per remote *layer*, text_model.rs:211-227,
    so many sessions are open at once while only the first of a contiguous run carries traffic).  K/V pages of a cache
    are allocated per layer on first use, so idle sessions cost nothing.  Forwards of different sessions share the
    ctx's stream and scratch and are serialised by a lock."""

    def __init__(self, ctx, blocks: Dict[str, "object"]):
        self.ctx, self.blocks = ctx, dict(blocks)
        self.dtype = ctx.dtype
        self.lock = threading.Lock()

    def info(self) -> Tuple[str, int]:
        return "cuda", int(self.ctx.device)

    def new_session(self) -> "_B200Session":
        return _B200Session(self)


class _B200Session:
    def __init__(self, be: B200Backend):
        from .model import Cache
        self.be = be
        with be.lock:
            self.cache = Cache(be.ctx)

    def clear_cache(self) -> None:
        with self.be.lock:
            self.cache.clear()

    def close(self) -> None:
        with self.be.lock:
            self.be.ctx.sync()
            self.cache.close()

    def forward_ops(self, x: RawTensor, ops: Sequence[Op]) -> RawTensor:
        from .capi import check, int_array, lib, ptr_array
        be, ctx = self.be, self.be.ctx
        for name, _, _ in ops:
            if name not in be.blocks:
                raise LookupError(f"could not find layer {name}")  # worker.rs:513
        if x.dtype_name != be.dtype:
            raise ValueError(f"forward pass failed for layer {ops[0][0]} (block_idx={ops[0][2]}): activation dtype "
                             f"{x.dtype_name} != model dtype {be.dtype}")
        a = x.to_numpy_bits()
        if a.ndim != 3 or a.shape[2] != ctx.config.hidden_size:
            raise ValueError(f"forward pass failed for layer {ops[0][0]} (block_idx={ops[0][2]}): unexpected shape {x.shape}")
        a = np.ascontiguousarray(a)
        # consecutive ops that share index_pos go down in one call (they always do: text_model.rs:298-321)
        i = 0
        with be.lock:
            while i < len(ops):
                j = i
                while j < len(ops) and ops[j][1] == ops[i][1]:
                    j += 1
                y = np.empty_like(a)
                hs = [be.blocks[n].h for n, _, _ in ops[i:j]]
                rc = lib().cake_b200_forward_batch_host(ctx.h, ptr_array(hs), int_array([b for _, _, b in ops[i:j]]),
                                                        j - i, self.cache.h, a.ctypes.data, y.ctypes.data,
                                                        a.shape[0], a.shape[1], ops[i][1])
                try:
                    check(rc)
                except Exception as e:
                    raise RuntimeError(f"forward pass failed for layer {ops[i][0]} (block_idx={ops[i][2]}): {e}") from e
                a, i = y, j
        return RawTensor.from_numpy_bits(a, be.dtype)


class WireWorker:
    """One cake worker endpoint (worker.rs:79-597) in front of a backend.  One thread per master connection (the
    reference spawns a task per connection); a backend that offers ``new_session()`` gives every connection its own
    KV cache, a plain backend object (tests) is shared by all connections."""

    VERSION = "cake-b200"

    def __init__(self, backend, host: str = "127.0.0.1", port: int = 0, cluster_key: Optional[str] = None):
        self.backend, self.cluster_key = backend, cluster_key
        self.sock = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self.sock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self.sock.bind((host, port))
        self.sock.listen(64)
        self.address = "%s:%d" % self.sock.getsockname()
        self.served = 0
        self.connections = 0
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self._conn_threads: List[threading.Thread] = []
        self._conns: List[socket.socket] = []

    def to_info(self, latency_ms: int) -> WorkerInfo:
        """worker.rs:47-57 (dtype is the Debug form of candle's DType: "BF16", "F16")."""
        dev, idx = self.backend.info()
        return WorkerInfo(self.VERSION, self.backend.dtype.upper(), platform.system().lower(), platform.machine(), dev, idx,
                          int(latency_ms))

    # worker.rs:298-575
    def handle_master_client(self, conn: socket.socket) -> None:
        conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        if self.cluster_key is not None:
            authenticate_as_worker(conn, self.cluster_key)
        t0 = time.perf_counter()
        _, first = Message.from_reader(conn)
        latency = (time.perf_counter() - t0) * 1e3
        if first.kind == "LayerAssignment":  # master re-running setup against a running worker (:316-329)
            Message("LayerAssignmentAck", needs_data=False).to_writer(conn)
            Message("WorkerReady").to_writer(conn)
            return
        if first.kind != "Hello":
            raise ProtocolError(f"unexpected first message (expected Hello): {first.kind}")
        sess = self.backend.new_session() if hasattr(self.backend, "new_session") else self.backend
        try:
            if sess is self.backend:
                sess.clear_cache()
            Message("WorkerInfo", info=self.to_info(latency)).to_writer(conn)
            self._serve_session(conn, sess)
        finally:
            if sess is not self.backend:
                sess.close()

    def _serve_session(self, conn: socket.socket, sess) -> None:
        while True:
            t0 = time.perf_counter()
            try:
                _, msg = Message.from_reader(conn)
            except (ConnectionError, OSError):
                return  # the reference's `while let Ok(..)` ends the same way
            read_ms = (time.perf_counter() - t0) * 1e3
            if msg.kind == "Goodbye":  # :363-383
                sess.clear_cache()
                Message("WorkerInfo", info=self.to_info(read_ms)).to_writer(conn)
                continue
            if msg.kind == "SingleOp":
                x, ops = msg.x, [(msg.layer_name, msg.index_pos, msg.block_idx)]
            elif msg.kind == "Batch":
                x, ops = msg.x, msg.batch
            else:
                raise ProtocolError(f"unhandled message in loop: {msg.kind}")
            try:
                if not ops:
                    raise ValueError("empty batch")
                y = sess.forward_ops(x.validate(), ops)
            except Exception as e:  # :490-520: report, keep the connection
                Message.worker_error(str(e)).to_writer(conn)
                continue
            Message.from_tensor(y).to_writer(conn)
            self.served += 1

    def _connection(self, conn: socket.socket) -> None:
        with conn:
            conn.settimeout(None)
            try:
                self.handle_master_client(conn)
            except (ProtocolError, ConnectionError, OSError):
                pass  # worker.rs:586-593: log and keep accepting

    def serve_forever(self) -> None:
        self.sock.settimeout(0.2)
        while not self._stop.is_set():
            try:
                conn, _ = self.sock.accept()
            except socket.timeout:
                continue
            except OSError:
                return
            self.connections += 1
            t = threading.Thread(target=self._connection, args=(conn,), daemon=True)
            self._conns.append(conn)
            self._conn_threads.append(t)
            t.start()

    def start(self) -> "WireWorker":
        self._thread = threading.Thread(target=self.serve_forever, daemon=True)
        self._thread.start()
        return self

    def stop(self) -> None:
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=5)
        self.sock.close()
        for c in self._conns:  # unblock connection threads still waiting for a message
            try:
                c.shutdown(socket.SHUT_RDWR)
            except OSError:
                pass
        for t in self._conn_threads:
            t.join(timeout=5)


# ------------------------------------------------------------------------------------------ master side
class WireClient:
    """client.rs:13-188: a remote block run reached over cake's TCP protocol.  Has the `Forwarder` methods
    (`forward_mut`, `forward_batch`, `goodbye`, `layer_name`, `ident`) over RawTensor activations."""

    def __init__(self, address: str, layer_name: str, cluster_key: Optional[str] = None, timeout: Optional[float] = None):
        self.address, self.name = address, layer_name
        host, port = address.rsplit(":", 1)
        try:
            self.sock = socket.create_connection((host, int(port)), timeout=timeout)
        except OSError as e:
            raise ConnectionError(f"can't connect to {address}: {e}") from e
        self.sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        if cluster_key is not None:
            authenticate_as_master(self.sock, cluster_key)
        resp = self.request(Message.hello())
        if resp.kind != "WorkerInfo":
            raise ProtocolError(f"unexpected worker info message: {resp.kind}")
        self.info = resp.info

    def request(self, req: Message) -> Message:
        req.to_writer(self.sock)
        return Message.from_reader(self.sock)[1]

    def forward_request(self, req: Message) -> RawTensor:
        msg = self.request(req)
        if msg.kind == "Tensor":
            return msg.x
        if msg.kind == "WorkerError":
            raise RuntimeError(f"worker {self.address} reported error: {msg.message}")
        raise ProtocolError(f"unexpected response {msg.kind}")

    def forward_mut(self, x: RawTensor, index_pos: int, block_idx: int) -> RawTensor:
        return self.forward_request(Message.single_op(self.name, x, index_pos, block_idx))

    def forward_batch(self, x: RawTensor, batch: Sequence[Op]) -> RawTensor:
        return self.forward_request(Message.from_batch(x, batch))

    def goodbye(self) -> None:
        self.request(Message.goodbye())

    def layer_name(self) -> str:
        return self.name

    def ident(self) -> str:
        return self.address

    def close(self) -> None:
        try:
            self.sock.close()
        except OSError:
            pass

    def __str__(self):
        i = self.info
        return f"{self.name}@{self.address} [{i.device}<{i.device_idx}> {i.os}-{i.arch} latency={i.latency}ms]"


class WireRemote:
    """Adapter that lets ``TextModelBase.load(make_remote=...)`` use a WireClient for layers a topology assigns to a
    TCP worker: device tensor -> RawTensor -> wire -> device tensor (the reference's `Client` does the same D2H/H2D,
    message.rs:96-145)."""

    def __init__(self, client: WireClient, name: str, ctx):
        self.client, self.name, self.ctx = client, name, ctx

    def _raw(self, x) -> RawTensor:
        import torch
        t = x.detach().contiguous().cpu()
        return RawTensor.from_numpy_bits(t.view(torch.uint16).numpy() if t.element_size() == 2 else t.numpy(), self.ctx.dtype)

    def _dev(self, raw: RawTensor):
        import torch
        t = torch.from_numpy(raw.to_numpy_bits().copy())
        return self.ctx.to_device(t.view(self.ctx.torch_dtype) if t.element_size() == 2 else t)

    def forward(self, x, index_pos, block_idx, ctx):
        return self.forward_batch(x, [(self.name, index_pos, block_idx)], ctx)

    forward_mut = forward

    def forward_batch(self, x, batch, ctx):
        ctx.sync()
        return self._dev(self.client.forward_batch(self._raw(x), batch))

    def goodbye(self) -> None:
        self.client.goodbye()

    def layer_name(self) -> str:
        return self.name

    def ident(self) -> str:
        return self.client.address
