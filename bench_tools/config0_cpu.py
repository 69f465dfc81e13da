"""BASELINE.json configs[0] on the CPU: Qwen3-0.6B-shaped synthetic checkpoint (f16, the reference's default dtype),
one prompt, 32 greedy tokens, through the oracle port of the reference's CPU path (the cake binary itself cannot be
built here).  tok/s as the reference defines it: (generated - 1) / time since the first token (master.rs:160-166).

    python bench_tools/config0_cpu.py [n_tokens=32] [prompt_len=16]

The GPU counterpart of this configuration is tests/test_gpu_fullsize.py::test_qwen3_0_6b_config0_f16_greedy_against_oracle.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("OMP_PROC_BIND", "close")
os.environ.setdefault("OMP_WAIT_POLICY", "active")

import numpy as np  # noqa: E402

from cake_b200.config import qwen3_0_6b  # noqa: E402
from cake_b200.synth import make_checkpoint  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    plen = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    cfg = qwen3_0_6b(max_seq=256)
    t0 = time.perf_counter()
    sd = make_checkpoint(cfg, "f16", seed=1234, std=0.02, peaked=False)
    m = O.OracleModel(cfg, sd, "f16", max_seq=256)
    build_s = time.perf_counter() - t0
    prompt = np.random.default_rng(7).integers(0, cfg.vocab_size, plen).tolist()
    cache = m.new_cache()
    ids, pos, toks, start = list(prompt), 0, [], None
    for i in range(n):
        if i == 1:
            start = time.perf_counter()   # master.rs:132-134: the timer restarts after the first token (prefill excluded)
        lg = m.forward(ids, pos, cache)
        pos += len(ids)
        t = O.argmax(lg)
        toks.append(t)
        ids = [t]
    dt = time.perf_counter() - start
    print(json.dumps({"config": "Qwen3-0.6B shape, f16, CPU (oracle port of the reference path), greedy", "generated": n,
                      "prompt_len": plen, "tok_s": (n - 1) / dt, "ms_per_token": dt / (n - 1) * 1e3,
                      "threads": O.lib().ora_num_threads(), "host_cores": os.cpu_count(), "setup_s": round(build_s, 1),
                      "tokens_head": toks[:8]}))


if __name__ == "__main__":
    main()
