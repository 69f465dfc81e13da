"""ThreadSanitizer run of the threaded worker endpoint (cake_worker: one thread and one KV-cache session per master
connection, forwards serialised by a mutex) over the oracle-backed emulation of the C ABI (tests/fake_b200):
six concurrent connections, each three rounds of prefill chunks + goodbye.  Expect `tsan warnings: 0`.
(OMP_NUM_THREADS=1: libgomp is not instrumented, its worker threads would be reported as racing with their own master.)

    python bench_tools/worker_tsan.py
"""
import os, subprocess, sys, threading, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(tempfile.gettempdir(), "cake_worker_tsan")
subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-o", BIN, os.path.join(ROOT, "cake_b200", "host", "cake_worker.cc"),
                       "-I", os.path.join(ROOT, "include"), "-L", os.path.join(ROOT, "cake_b200"), "-lcake_b200", "-pthread"])

sys.path.insert(0, ROOT)
import numpy as np  # noqa: F401
sys.path.insert(0, ROOT)
from cake_b200.loader import save_checkpoint
from cake_b200.wire import RawTensor, WireClient
from tests.fake_b200.make_fake import build as build_fake
from tests.util import checkpoint, medium_config, rand_x, f32_to_bits, bits_to_f32
from oracle import oracle as O
tmp = tempfile.mkdtemp()
build_fake(tmp, oracle=True)
cfg = medium_config(num_hidden_layers=4, hidden_size=128, intermediate_size=256, vocab_size=256, num_attention_heads=4, num_key_value_heads=2, head_dim=32)
sd = checkpoint(cfg, "bf16", seed=19)
save_checkpoint(tmp + "/model", cfg, sd, shard_bytes=200_000)
env = {**os.environ, "LD_LIBRARY_PATH": tmp, "TSAN_OPTIONS": "halt_on_error=0", "OMP_NUM_THREADS": "1"}
NCONN = 6
p = subprocess.Popen([BIN, tmp + "/model", "--layers", "model.layers.2-3", "--address", "127.0.0.1:0", "--max-seq", "32", "--connections", str(NCONN)],
                     stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
addr = p.stdout.readline().strip()[len("listening on "):]
om = O.OracleModel(cfg, sd, "bf16", max_seq=32)
x = rand_x((1, 12, cfg.hidden_size), "bf16", seed=2)
raw = lambda t: RawTensor.from_numpy_bits(f32_to_bits(t.float().numpy(), "bf16"), "bf16")
errs = []
def run(k):
    try:
        c = WireClient(addr, cfg.layer_name(2), timeout=60)
        oc = om.new_cache()
        for rep in range(3):
            for t in range(0, 12, 3):
                y = c.forward_batch(raw(x[:, t:t+3]), [(cfg.layer_name(i), t, i) for i in (2, 3)])
            c.goodbye()
        c.close()
    except Exception as e:
        errs.append(repr(e))
ths = [threading.Thread(target=run, args=(k,)) for k in range(NCONN)]
[t.start() for t in ths]; [t.join() for t in ths]
rc = p.wait(60)
err = p.stderr.read()
print("client errors:", errs); print("worker rc:", rc); print("tsan warnings:", err.count("WARNING: ThreadSanitizer")); print(err[-1500:] if "ThreadSanitizer" in err else "")
