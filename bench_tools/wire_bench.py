"""TCP round-trip benchmark of the wire endpoint, the reference's own `bench_tcp_roundtrip` /
`bench_tcp_roundtrip_batch` (cake-core/tests/protocol.rs:435-540): loopback, echo worker, SingleOp / 16-layer Batch of a
[1, hidden] f16 activation, 50 warm-up + 500 timed round trips per size.  No GPU involved — this is the protocol cost
that the NVLink hand-off removes from the intra-box path and that remains when a remote cake master drives the box.

    python bench_tools/wire_bench.py            # C++ worker (cake_worker --echo) and Python worker, Python client
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cake_b200.build import build_host  # noqa: E402
from cake_b200.wire import Message, RawTensor, WireClient  # noqa: E402

SIZES = [64, 512, 2048, 4096, 5120, 8192]
WARMUP, ITERS = 50, 500


def spawn(cmd):
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, cwd=ROOT)
    line = p.stdout.readline().strip()
    assert line.startswith("listening on "), line
    return p, line[len("listening on "):]


def roundtrips(addr, hidden, batch_layers=0):
    c = WireClient(addr, "model.layers.0", timeout=10)
    a = (np.arange(hidden, dtype=np.float32) * 0.01).astype(np.float16).reshape(1, 1, hidden)
    x = RawTensor.from_numpy_bits(a.view(np.uint16), "f16")
    msg = Message.from_batch(x, [(f"model.layers.{i}", 0, i) for i in range(batch_layers)]) if batch_layers \
        else Message.single_op("model.layers.0", x, 0, 0)
    frame = msg.frame()   # encode once: measures socket + worker + decode of the answer, like the reference's loop
    for _ in range(WARMUP):
        c.sock.sendall(frame)
        Message.from_reader(c.sock)
    t0 = time.perf_counter()
    for _ in range(ITERS):
        c.sock.sendall(frame)
        Message.from_reader(c.sock)
    dt = time.perf_counter() - t0
    c.close()
    return dt / ITERS * 1e6


def main():
    build_host()
    out = {}
    for label, cmd in (("cpp_worker", [os.path.join(ROOT, "cake_b200", "host", "cake_worker"), "--echo", "--address", "127.0.0.1:0"]),
                       ("python_worker", [sys.executable, "-m", "cake_b200.worker", "--echo", "--address", "127.0.0.1:0"])):
        p, addr = spawn(cmd)
        try:
            out[label] = {"single_op_us": {f"[1,{h}]": round(roundtrips(addr, h), 1) for h in SIZES},
                          "batch16_us": {f"[1,{h}]": round(roundtrips(addr, h, 16), 1) for h in SIZES}}
        finally:
            p.kill()
            p.wait(5)
    print(json.dumps({"bench": "tcp round trip, loopback, echo worker, python client", "iters": ITERS, "host_cores": os.cpu_count(), **out}, indent=1))


if __name__ == "__main__":
    main()
