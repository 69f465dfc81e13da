"""How far may two CORRECT implementations of the path differ?  (CPU only; test-infrastructure experiment.)
The reference pins "f32 accumulate, round to D" but not the summation ORDER.  This runs the oracle on the full
Llama-3-8B shape (random-init, seed 1234 on the CPU generator) for one decode token at KV length 2048 with three
summation orders of the linear layers (cake_oracle.c dot modes 0/1/2) and reports the logit differences in bf16 ulps
at the logits' top binade, per depth.   python bench_tools/order_sensitivity.py [n_layers ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cake_b200.config import llama3_8b
from cake_b200.synth import make_head, make_layer
from oracle import oracle as O
from tests.util import max_ulp_err

depths = [int(a) for a in sys.argv[1:]] or [4, 8, 16, 32]
L = 2048
cfg = llama3_8b(max_seq=L + 8)
nl = max(depths)
cfg.num_hidden_layers = nl
t0 = time.time()
sd = make_head(cfg, "bf16", seed=1234)
for i in range(nl):
    sd.update(make_layer(cfg, i, "bf16", seed=1234))
print(f"weights for {nl} layers in {time.time() - t0:.0f}s", flush=True)
om = O.OracleModel(cfg, sd, "bf16", max_seq=L + 8)
rng = np.random.default_rng(0)
kv = [O.round_to(rng.standard_normal((cfg.num_key_value_heads, L, cfg.hd), dtype=np.float32), "bf16") for _ in range(2)]
x0 = om.embed([17])
res = {}
for mode in (0, 1, 2):
    O.set_dot_mode(mode)
    oc = om.new_cache(L + 8)
    for l in range(nl):
        k, v = oc.kv(l)
        k[:, :L] = kv[0]; v[:, :L] = kv[1]
        oc.set_len(l, L)
    x = x0.copy()
    t0 = time.time()
    prev = 0
    for d in sorted(depths):
        x = om.forward_layers(x, prev, d, L, oc)
        prev = d
        res[(mode, d)] = (x.copy(), om.logits(x))
    print(f"mode {mode}: {time.time() - t0:.1f}s", flush=True)
O.set_dot_mode(0)
print("depth | hidden: mode1 vs 0, mode2 vs 0 (ulp) | logits: mode1 vs 0, mode2 vs 0, mode2 vs 1 (ulp) | top-1 equal")
for d in sorted(depths):
    h0, l0 = res[(0, d)]; h1, l1 = res[(1, d)]; h2, l2 = res[(2, d)]
    print(f"{d:5d} | {max_ulp_err(h1, h0, 'bf16'):6.2f} {max_ulp_err(h2, h0, 'bf16'):6.2f} | "
          f"{max_ulp_err(l1, l0, 'bf16'):6.2f} {max_ulp_err(l2, l0, 'bf16'):6.2f} {max_ulp_err(l2, l1, 'bf16'):6.2f} | "
          f"{O.argmax(l0) == O.argmax(l1) == O.argmax(l2)}")
