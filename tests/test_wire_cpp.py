"""The compiled worker endpoint (cake_b200/host/cake_wire.hpp + cake_worker.cc) against the Python implementation of
the same protocol (cake_b200/wire.py): two independent codecs must agree byte for byte on every message, the HMAC
handshake must interoperate in both outcomes, and the worker loop must behave like worker.rs:298-575.
Scenarios as in tests/test_wire.py (the reference's tests/protocol.rs mock-worker cases)."""
import os
import socket
import subprocess

import numpy as np
import pytest

from cake_b200.build import build_host
from cake_b200.parallel import expand_layers
from cake_b200.wire import DTYPE_TAGS, Message, ProtocolError, RawTensor, WireClient, WorkerInfo, _recv_exact

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "cake_b200", "host", "cake_worker")


class Proc:
    def __init__(self, *args):
        build_host()
        self.p = subprocess.Popen([WORKER, *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        line = self.p.stdout.readline().strip()
        if not line.startswith("listening on "):
            err = self.p.stderr.read()
            self.p.wait(5)
            raise RuntimeError(f"cake_worker did not start: {line!r} {err!r}")
        self.address = line[len("listening on "):]

    def close(self, timeout=10) -> int:
        try:
            return self.p.wait(timeout)
        except subprocess.TimeoutExpired:
            self.p.kill()
            self.p.wait(5)
            return -9

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if self.p.poll() is None:
            self.p.kill()
            self.p.wait(5)


def raw_f16(shape) -> RawTensor:
    n = int(np.prod(shape))
    a = (np.arange(n, dtype=np.float32) * 0.01).astype(np.float16).reshape(shape)
    return RawTensor.from_numpy_bits(a.view(np.uint16), "f16")


@pytest.mark.parametrize("expr", ["model.layers.0-2", "model.layers.7,model.layers.9-9", "model.language_model.layers.10-11",
                                  "model.layers.3", "a.b-c.layers.1-2", "model.layers.5-2"])
def test_layer_range_expansion_matches_python(expr):
    build_host()
    r = subprocess.run([WORKER, "--expand", expr], capture_output=True, text=True)
    try:
        want = expand_layers(expr.split(","))
    except ValueError:
        assert r.returncode == 1 and "end must be >= start" in r.stderr
        return
    assert r.returncode == 0 and r.stdout.split() == want


@pytest.mark.parametrize("key", [None, "test-cluster-key-123"])
def test_python_master_against_cpp_worker(key):
    args = ["--echo", "--address", "127.0.0.1:0", "--connections", "1"] + (["--cluster-key", key] if key else [])
    with Proc(*args) as w:
        c = WireClient(w.address, "model.layers.0", cluster_key=key, timeout=10)
        assert c.info.device == "cpu" and c.info.dtype == "F16" and c.info.version == "cake-b200" and c.info.os == "linux"
        t = raw_f16([1, 128])
        y = c.forward_mut(t, 0, 0)
        assert y.dtype == DTYPE_TAGS["f16"] and y.shape == [1, 128] and y.data == t.data
        t = raw_f16([1, 64])
        assert c.forward_batch(t, [("model.layers.0", 0, 0), ("model.layers.1", 0, 1), ("model.layers.2", 0, 2)]).data == t.data
        for i in range(10):   # tests/protocol.rs:262-290: a different shape every time
            ti = raw_f16([1, 32 + i])
            yi = c.forward_mut(ti, i, 0)
            assert yi.data == ti.data and yi.shape == [1, 32 + i]
        c.goodbye()
        assert c.forward_mut(t, 0, 0).data == t.data
        big = raw_f16([1, 5120])
        assert c.forward_mut(big, 0, 0).data == big.data
        with pytest.raises(RuntimeError, match="reported error: could not find layer model.layers.99"):
            c.forward_batch(t, [("model.layers.99", 0, 99)])
        bad = RawTensor(t.data[:-2], t.dtype, t.shape)  # shape/bytes mismatch is reported, not fatal
        with pytest.raises(RuntimeError, match="reported error"):
            c.forward_mut(bad, 0, 0)
        assert c.forward_mut(t, 0, 0).data == t.data
        c.close()
        assert w.close() == 0


def test_wrong_key_is_rejected_then_right_key_is_served():
    with Proc("--echo", "--address", "127.0.0.1:0", "--connections", "2", "--cluster-key", "correct-key") as w:
        with pytest.raises((ProtocolError, ConnectionError)):
            WireClient(w.address, "model.layers.0", cluster_key="wrong-key", timeout=10)
        c = WireClient(w.address, "model.layers.0", cluster_key="correct-key", timeout=10)
        assert c.forward_mut(raw_f16([1, 8]), 0, 0).shape == [1, 8]
        c.close()
        assert w.close() == 0


def test_every_message_survives_cpp_decode_and_encode_byte_for_byte():
    x = raw_f16([1, 64])
    msgs = [
        Message.hello(), Message.goodbye(), Message("ModelDataDone"), Message("WorkerReady"),
        Message("WorkerInfo", info=WorkerInfo("0.1.0", "F16", "linux", "aarch64", "cuda", 1, (1 << 70) + 42)),
        Message.single_op("model.layers.5", x, 10, 5),
        Message.from_batch(x, [(f"model.layers.{i}", 3, i) for i in range(32)]),
        Message.from_tensor(x), Message.from_tensor(RawTensor(b"", DTYPE_TAGS["bf16"], [0, 4])),
        Message("LayerAssignment", layers=["model.layers.0", "model.layers.1"], model_hash="abc123"),
        Message("LayerAssignment", layers=[], model_hash=""),
        Message("LayerAssignmentAck", needs_data=True), Message("LayerAssignmentAck", needs_data=False),
        Message("ModelDataChunk", filename="model.safetensors", offset=1 << 33, total_size=1 << 34, compressed=True,
                checksum=0xDEADBEEF, data=b"\x00\x01\x02" * 100),
        Message("ModelDataResume", filename="m.safetensors", offset=12345),
        Message.worker_error("forward pass failed for layer model.layers.3 (block_idx=3): boom — ünïcode"),
    ]
    with Proc("--echo", "--reflect", "--address", "127.0.0.1:0", "--connections", "1") as w:
        host, port = w.address.rsplit(":", 1)
        s = socket.create_connection((host, int(port)), timeout=10)
        Message.hello().to_writer(s)
        assert Message.from_reader(s)[1].kind == "WorkerInfo"
        for m in msgs:
            s.sendall(m.frame())
            magic_size = _recv_exact(s, 8)
            assert magic_size[:4] == bytes([0x01, 0x04, 0xF4, 0xC7])
            payload = _recv_exact(s, int.from_bytes(magic_size[4:], "big"))
            assert payload == m.to_bytes(), m.kind
            assert Message.from_bytes(payload) == m
        s.close()
        assert w.close() == 0


def test_bad_magic_ends_the_connection_but_not_the_worker():
    with Proc("--echo", "--address", "127.0.0.1:0", "--connections", "3") as w:
        host, port = w.address.rsplit(":", 1)
        s = socket.create_connection((host, int(port)), timeout=10)
        s.sendall(bytes([0xDE, 0xAD, 0xBE, 0xEF, 0, 0, 0, 4, 0, 0, 0, 0]))
        try:
            assert s.recv(16) == b""   # closed without an answer (a reset if our payload was still unread)
        except ConnectionResetError:
            pass
        s.close()
        s = socket.create_connection((host, int(port)), timeout=10)   # LayerAssignment as first message (worker.rs:316-329)
        Message("LayerAssignment", layers=["model.layers.1"], model_hash="h").to_writer(s)
        ack, ready = Message.from_reader(s)[1], Message.from_reader(s)[1]
        assert ack.kind == "LayerAssignmentAck" and ack.needs_data is False and ready.kind == "WorkerReady"
        s.close()
        c = WireClient(w.address, "model.layers.0", timeout=10)
        assert c.forward_mut(raw_f16([1, 8]), 0, 0).shape == [1, 8]
        c.close()
        assert w.close() == 0


def test_worker_needs_a_gpu_for_real_layers(tmp_path):
    import torch
    from cake_b200.loader import save_checkpoint
    from tests.util import checkpoint, medium_config
    if torch.cuda.is_available():
        pytest.skip("only meaningful without a GPU")
    cfg = medium_config(num_hidden_layers=2)
    save_checkpoint(str(tmp_path), cfg, checkpoint(cfg, "bf16", seed=1), shard_bytes=1_000_000)
    r = subprocess.run([WORKER, str(tmp_path), "--layers", "model.layers.1-1", "--address", "127.0.0.1:0"],
                       capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "ctx_create" in r.stderr   # config + shard subset parsed, then no CPU fallback


@pytest.mark.gpu
def test_gpu_cpp_worker_equals_local_blocks(tmp_path):
    """cake_worker serves layers 2-3 of a checkpoint on disk; the bits must equal the same blocks run in-process."""
    import torch
    from cake_b200.loader import save_checkpoint
    from cake_b200.model import B200Transformer, Context
    from tests.util import checkpoint, medium_config, rand_x
    cfg = medium_config(num_hidden_layers=4)
    sd = checkpoint(cfg, "bf16", seed=17)
    save_checkpoint(str(tmp_path), cfg, sd, shard_bytes=2_000_000)
    x = rand_x((1, 5, cfg.hidden_size), "bf16", seed=3)
    batch = [(cfg.layer_name(i), 0, i) for i in (2, 3)]
    ctx = Context(cfg, sd, "bf16", max_seq=64)
    try:
        blks = [B200Transformer.load(n, ctx) for n, _, _ in batch]
        y_local = blks[0].forward_batch(ctx.to_device(x), batch, ctx, blocks=blks)
        ctx.sync()
        y_local = y_local.cpu().view(torch.uint16).numpy()
        x1 = rand_x((1, 1, cfg.hidden_size), "bf16", seed=4)   # then one decode step at position 5
        step = [(cfg.layer_name(i), 5, i) for i in (2, 3)]
        y1_local = blks[0].forward_batch(ctx.to_device(x1), step, ctx, blocks=blks)
        ctx.sync()
        # both results against the oracle's layers 2-3 (prefill of 5, then the decode step at position 5)
        from oracle import oracle as O
        from tests.util import max_ulp_err
        om = O.OracleModel(cfg, sd, "bf16", max_seq=64)
        oc = om.new_cache(64)
        assert max_ulp_err(torch.from_numpy(y_local.copy()).view(torch.bfloat16)[0].float().numpy(),
                           om.forward_layers(x[0].float().numpy(), 2, 4, 0, oc), "bf16") <= 3.0
        assert max_ulp_err(y1_local[0].float().cpu().numpy(), om.forward_layers(x1[0].float().numpy(), 2, 4, 5, oc), "bf16") <= 3.0
        y1_local = y1_local.cpu().view(torch.uint16).numpy()
    finally:
        ctx.close()
    with Proc(str(tmp_path), "--layers", "model.layers.2-3", "--max-seq", "64", "--address", "127.0.0.1:0", "--connections", "2") as w:
        c = WireClient(w.address, cfg.layer_name(2), timeout=120)
        other = WireClient(w.address, cfg.layer_name(3), timeout=120)   # a second open connection = its own KV cache
        assert c.info.device == "cuda" and c.info.dtype == "BF16"
        y = c.forward_batch(RawTensor.from_numpy_bits(x.view(torch.uint16).numpy(), "bf16"), batch)
        assert y.shape == [1, 5, cfg.hidden_size] and np.array_equal(y.to_numpy_bits(), y_local)
        y1 = c.forward_batch(RawTensor.from_numpy_bits(x1.view(torch.uint16).numpy(), "bf16"), step)
        assert np.array_equal(y1.to_numpy_bits(), y1_local)
        with pytest.raises(RuntimeError, match="could not find layer"):
            c.forward_batch(RawTensor.from_numpy_bits(x1.view(torch.uint16).numpy(), "bf16"), [(cfg.layer_name(0), 6, 0)])
        with pytest.raises(RuntimeError, match="forward pass failed for layer"):   # position out of order: reported, not fatal
            c.forward_batch(RawTensor.from_numpy_bits(x1.view(torch.uint16).numpy(), "bf16"), [(cfg.layer_name(2), 9, 2)])
        # the other session's cache is still empty: the same prefill at position 0 is accepted there and gives the same bits
        y2 = other.forward_batch(RawTensor.from_numpy_bits(x.view(torch.uint16).numpy(), "bf16"), batch)
        assert np.array_equal(y2.to_numpy_bits(), y_local)
        for cl in (c, other):
            cl.goodbye()
            cl.close()
        assert w.close(30) == 0


def test_cpp_worker_serves_many_open_connections_like_a_cake_master():
    """text_model.rs:211-227 opens one Client per remote layer before any traffic: all Hellos must be answered while
    the earlier connections stay open; only the first connection of the run carries the batch."""
    with Proc("--echo", "--address", "127.0.0.1:0", "--connections", "16") as w:
        clients = [WireClient(w.address, f"model.layers.{i}", timeout=10) for i in range(16, 32)]
        t = raw_f16([1, 1, 64])
        batch = [(f"model.layers.{i}", 7, i) for i in range(16, 32)]
        for _ in range(3):
            assert clients[0].forward_batch(t, batch).data == t.data
        assert clients[5].forward_mut(t, 0, 21).data == t.data
        for c in clients:
            c.goodbye()
            c.close()
        assert w.close() == 0


TOPOLOGY = """# two workers, the syntax of cake's topology.yml
gpu1:
  host: "10.0.0.2:10128"   # first box
  description: 'B200 #1'
  layers:
    - "model.layers.8-15"
    - model.layers.20
gpu2:
  host: 10.0.0.3:10128
  layers: ["model.layers.16-17", 'model.layers.31']
  vram_bytes: 1000
"""


def test_topology_file_gives_cpp_and_python_workers_the_same_layers(tmp_path):
    import sys
    from cake_b200.parallel import load_topology
    build_host()
    p = tmp_path / "topology.yml"
    p.write_text(TOPOLOGY)
    topo = load_topology(str(p))
    assert topo["gpu2"]["layers"] == ["model.layers.16", "model.layers.17", "model.layers.31"]
    for name in ("gpu1", "gpu2"):
        r = subprocess.run([WORKER, "--topology", str(p), "--name", name], capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout.split() == topo[name]["layers"], r.stderr
        r = subprocess.run([sys.executable, "-m", "cake_b200.worker", "--topology", str(p), "--name", name],
                           capture_output=True, text=True, cwd=ROOT)
        assert r.returncode == 0 and r.stdout.split() == topo[name]["layers"], r.stderr
    r = subprocess.run([WORKER, "--topology", str(p), "--name", "gpu9"], capture_output=True, text=True)
    assert r.returncode == 1 and "could not find topology node" in r.stderr
    r = subprocess.run([sys.executable, "-m", "cake_b200.worker", "--topology", str(p), "--name", "gpu9"],
                       capture_output=True, text=True, cwd=ROOT)
    assert r.returncode != 0 and "could not find topology node" in r.stderr


def test_python_worker_cli_speaks_to_the_python_client():
    import sys
    p = subprocess.Popen([sys.executable, "-m", "cake_b200.worker", "--echo", "--address", "127.0.0.1:0"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT)
    try:
        line = p.stdout.readline().strip()
        assert line.startswith("listening on "), p.stderr.read()
        c = WireClient(line[len("listening on "):], "model.layers.0", timeout=10)
        t = raw_f16([1, 1, 32])
        assert c.forward_mut(t, 0, 0).data == t.data and c.info.device == "cpu"
        c.goodbye()
        c.close()
    finally:
        p.kill()
        p.wait(5)


def test_hostile_frames_are_answered_or_dropped_never_fatal():
    """Shapes whose element count wraps around u64, absurd element counts and truncated payloads must not take the
    worker down: the op is reported as an error (or the connection dropped) and the next connection is served."""
    with Proc("--echo", "--address", "127.0.0.1:0", "--connections", "3") as w:
        c = WireClient(w.address, "model.layers.0", timeout=10)
        wrap = RawTensor(b"", DTYPE_TAGS["f16"], [1 << 63, 2, 64])           # product == 0 (mod 2**64)
        c.sock.sendall(Message.single_op("model.layers.0", wrap, 0, 0).frame())
        assert Message.from_reader(c.sock)[1].kind == "WorkerError"
        assert c.forward_mut(raw_f16([1, 8]), 0, 0).shape == [1, 8]           # connection still usable
        c.close()
        host, port = w.address.rsplit(":", 1)
        s = socket.create_connection((host, int(port)), timeout=10)
        Message.hello().to_writer(s)
        assert Message.from_reader(s)[1].kind == "WorkerInfo"
        # Batch that announces 2**32-1 ops but carries none: decoding fails, the connection is dropped
        payload = Message.from_batch(raw_f16([1, 8]), []).to_bytes()[:-4] + b"\xff\xff\xff\xff"
        s.sendall(bytes([0x01, 0x04, 0xF4, 0xC7]) + len(payload).to_bytes(4, "big") + payload)
        try:
            assert s.recv(16) == b""
        except ConnectionResetError:
            pass
        s.close()
        c = WireClient(w.address, "model.layers.0", timeout=10)
        assert c.forward_mut(raw_f16([1, 8]), 0, 0).shape == [1, 8]
        c.close()
        assert w.close() == 0


def test_random_messages_agree_between_the_two_codecs():
    """Property check: whatever message the Python codec can build, the C++ codec decodes and re-encodes to the very
    same bytes (200 seeded random messages over all variants, including empty strings/lists and 64/128-bit extremes)."""
    rng = np.random.default_rng(2024)

    def rstr(maxlen=24):
        n = int(rng.integers(0, maxlen))
        return "".join(chr(int(c)) for c in rng.choice([*range(32, 127), 0xE9, 0x4E2D, 0x1F370], n))

    def ru64():  # index into a Python list: numpy would turn 2**64-1 into a float
        vals = [0, 1, 255, 2**31, 2**32 - 1, 2**63, 2**64 - 1, int(rng.integers(0, 2**62))]
        return vals[int(rng.integers(0, len(vals)))]

    def rtensor():
        dt = str(rng.choice(["bf16", "f16", "f32", "u8", "u32", "i64", "f64"]))
        shape = [int(d) for d in rng.integers(0, 5, int(rng.integers(0, 4)))]
        n = int(np.prod(shape)) if shape else 1
        size = {"bf16": 2, "f16": 2, "f32": 4, "u8": 1, "u32": 4, "i64": 8, "f64": 8}[dt]
        return RawTensor(rng.bytes(n * size), DTYPE_TAGS[dt], shape)

    def rmsg():
        k = int(rng.integers(0, 13))
        kind = Message.KINDS[k]
        if kind == "WorkerInfo":
            return Message(kind, info=WorkerInfo(rstr(), rstr(), rstr(), rstr(), rstr(), ru64(), [0, 5, 2**64, 2**128 - 1][int(rng.integers(0, 4))]))
        if kind == "SingleOp":
            return Message.single_op(rstr(), rtensor(), ru64(), ru64())
        if kind == "Batch":
            return Message.from_batch(rtensor(), [(rstr(), ru64(), ru64()) for _ in range(int(rng.integers(0, 6)))])
        if kind == "Tensor":
            return Message.from_tensor(rtensor())
        if kind == "LayerAssignment":
            return Message(kind, layers=[rstr() for _ in range(int(rng.integers(0, 5)))], model_hash=rstr())
        if kind == "LayerAssignmentAck":
            return Message(kind, needs_data=bool(rng.integers(0, 2)))
        if kind == "ModelDataChunk":
            return Message(kind, filename=rstr(), offset=ru64(), total_size=ru64(), compressed=bool(rng.integers(0, 2)),
                           checksum=int(rng.integers(0, 2**32)), data=rng.bytes(int(rng.integers(0, 300))))
        if kind == "ModelDataResume":
            return Message(kind, filename=rstr(), offset=ru64())
        if kind == "WorkerError":
            return Message.worker_error(rstr(80))
        return Message(kind)

    with Proc("--echo", "--reflect", "--address", "127.0.0.1:0", "--connections", "1") as w:
        host, port = w.address.rsplit(":", 1)
        s = socket.create_connection((host, int(port)), timeout=10)
        Message.hello().to_writer(s)
        assert Message.from_reader(s)[1].kind == "WorkerInfo"
        seen = set()
        for _ in range(200):
            m = rmsg()
            seen.add(m.kind)
            s.sendall(m.frame())
            head = _recv_exact(s, 8)
            payload = _recv_exact(s, int.from_bytes(head[4:], "big"))
            assert payload == m.to_bytes(), m
        assert len(seen) == 13
        s.close()
        assert w.close() == 0
