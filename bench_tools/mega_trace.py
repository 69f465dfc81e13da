"""Profiling aid: per-phase timeline of the persistent decode megakernel on the Llama-3-8B shape at KV 2048, from the
kernel's own stamps (profiling build -DMK_TRACE=1 -> cake_b200/libcake_b200_trace.so; the product library is untouched).
   python bench_tools/mega_trace.py [n_layers=8]
Prints (a) CTA 0's phase / barrier durations per layer and (b) for layer 1, over all 148 CTAs: staging, compute,
starved (waiting for data) and producer-blocked (ring full) times per phase — who limits each phase."""
import ctypes, os, sys
os.environ["CAKE_B200_MEGA_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cake_b200 import capi
from cake_b200.build import build_trace
capi.SO_PATH = build_trace()
from cake_b200.capi import lib, check
from cake_b200.config import llama3_8b
from cake_b200.model import Context, TextModelBase, Cache
from cake_b200.synth import make_head, make_layer

nl = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = llama3_8b(max_seq=4096); cfg.num_hidden_layers = nl
sd = make_head(cfg, "bf16", device="cuda")
for i in range(nl): sd.update(make_layer(cfg, i, "bf16", device="cuda"))
ctx = Context(cfg, sd, "bf16", max_seq=4096)
ctx.cache = Cache(ctx, 1, 2304)
m = TextModelBase.load(ctx)
ctx.cache.fill_synthetic(list(range(nl)), 2048, 7); ctx.sync(); m.index_pos = 2048
m.decode_build()
m.decode_greedy(5, 8)
L = lib(); L.cake_b200_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]; L.cake_b200_debug_trace.restype = ctypes.c_int
BASE, NT = 4096, 4096 + 160 * 5 * 16
buf = (ctypes.c_ulonglong * NT)()
check(L.cake_b200_debug_trace(ctx.h, buf, NT))
raw = np.array(buf[:], dtype=np.int64)
n = 1 + nl * 10
t = (raw[:n] - raw[0]) / 1e3
names = ["qkv", "bar", "attn", "bar", "o", "bar", "gate_up", "bar", "down", "bar"]
d = np.diff(t)
per = {}
for l in range(nl):
    seg = d[l * 10:(l + 1) * 10]
    for k, v in zip(names, seg):
        per.setdefault(k, []).append(v)
    if l < 3 or l == nl - 1:
        print(f"layer {l}: " + " ".join(f"{k}={v:.1f}" for k, v in zip(names, seg)))
for k, v in per.items():
    print(f"{k:8s} mean {np.mean(v[1:]):7.2f} us  (layers 1..{nl - 1})")
lay = sum(np.mean(v[1:]) * (5 if k == "bar" else 1) for k, v in per.items())
print(f"layer total (CTA 0, layers 1..) {lay:.1f} us")

rec = raw[BASE:BASE + 148 * 5 * 16].reshape(148, 5, 16).astype(np.float64)
clk_ghz = 1.9
pn = ["qkv", "attn", "o", "gate_up", "down"]
t0 = rec[:, 0, 0].min()
print("\nlayer 1, all 148 CTAs (us; median / max over CTAs).  starved = warp 0 waiting for a full stage; blocked = producer waiting for a free stage")
for p in range(5):
    r = rec[:, p, :]
    act = r[:, 0] > 0
    g = lambda a: f"{np.median(a[act]):6.2f}/{np.max(a[act]):6.2f}"
    start, staged, last, arrive, passed = [(r[:, i] - t0) / 1e3 for i in range(5)]
    line = (f"{pn[p]:8s} enter@{np.median(start[act]):7.2f}  stage {g(staged - start)}  consume {g(last - staged)}  epilogue {g(arrive - last)}  "
            f"barrier {g(passed - arrive)}  starved {g(r[:, 5] / clk_ghz / 1e3)}  stages {np.median(r[act, 6]):.0f}")
    pf, ps, pp = [(r[:, i] - t0) / 1e3 for i in (7, 8, 9)]
    line += (f" | producer first@{np.median(pf[act]):7.2f} static+{g(ps - pf)} pool+{g(pp - ps)} blocked {g(r[:, 10] / clk_ghz / 1e3)} pool_groups {np.median(r[act, 11]):.0f}"
             f" (max {np.max(r[act, 11]):.0f})")
    print(line)
    if p == 1:
        a12, a13, a14, a15 = [(r[:, i] - t0) / 1e3 for i in (12, 13, 14, 15)]
        act2 = r[:, 12] > 0
        gg = lambda a: f"{np.median(a[act2]):6.2f}/{np.max(a[act2]):6.2f}"
        print(f"         attention: prep {gg(a12 - start)} tiles {gg(a13 - a12)} partials {gg(a14 - a13)} ticket {gg(a15 - a14)} merge+rest {gg(arrive - a15)}")
print("phase span (first CTA enters -> last CTA passes the barrier), us:",
      {pn[p]: round(float(((rec[:, p, 4].max() - rec[:, p, 0].min()) / 1e3)), 2) for p in range(5)})
