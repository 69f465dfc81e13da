"""Profiling aid: per-phase timing of the persistent decode megakernel (CTA 0's %globaltimer stamps).
   CAKE_B200_MEGA_TRACE=1 python bench_tools/mega_trace.py [n_layers]"""
import ctypes, os, sys
os.environ["CAKE_B200_MEGA_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
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
L = lib(); L.cake_b200_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
n = 1 + nl * 9 + 4
buf = (ctypes.c_ulonglong * n)()
check(L.cake_b200_debug_trace(ctx.h, buf, n))
t = np.array(buf[:], dtype=np.int64); t = (t - t[0]) / 1e3
# stamps: start, then per layer: [qkv done, barrier done, attn done, barrier done, o done, b, gu done, b, down done, b]
names = ["qkv", "bar", "attn", "bar", "o", "bar", "gate_up", "bar", "down", "bar"]
d = np.diff(t)
per = {}
for l in range(nl):
    seg = d[l * 10:(l + 1) * 10]
    for k, v in zip(names, seg):
        per.setdefault(k if k != "bar" else "bar", []).append(v)
    if l < 3 or l == nl - 1:
        print(f"layer {l}: " + " ".join(f"{k}={v:.1f}" for k, v in zip(names, seg)))
for k, v in per.items():
    print(f"{k:8s} mean {np.mean(v):7.2f} us  (n={len(v)})")
print("total us", t[min(len(t) - 1, nl * 10)], "head+tail", d[nl * 10:nl * 10 + 3] if len(d) > nl * 10 else None)

n2 = 2048 + 8
buf2 = (ctypes.c_ulonglong * n2)()
check(L.cake_b200_debug_trace(ctx.h, buf2, n2))
a = np.array(buf2[2048:2056], dtype=np.int64)
print("attention sub-phases (layer 1, CTA 0) us:", dict(zip(["qprep", "tile_wait", "scores", "softmax", "pv", "partial_wr", "fence+ticket"], np.round(np.diff(a) / 1e3, 2))))

n3 = 2560 + 148
buf3 = (ctypes.c_ulonglong * n3)()
check(L.cake_b200_debug_trace(ctx.h, buf3, n3))
fin = np.array(buf3[2304:2304 + 148], dtype=np.int64) / 1e3
rel = np.array(buf3[2560:2560 + 148], dtype=np.int64) / 1e3
fin -= fin.min(); 
print("gate_up(l=1) finish spread over CTAs us: p50 %.2f p90 %.2f max %.2f ; sorted tail:" % (np.percentile(fin, 50), np.percentile(fin, 90), fin.max()), np.round(np.sort(fin)[-6:], 2))
print("barrier release after last finisher us: min %.2f max %.2f" % ((rel - rel.min()).min() , (rel.max() - (np.array(buf3[2304:2304+148],dtype=np.int64)/1e3).max())))
print("slowest CTAs:", np.argsort(fin)[-8:])
