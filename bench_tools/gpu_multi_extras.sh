#!/usr/bin/env bash
# Rehearsal of the driver's N=8 line at N GPUs: the 8B leg followed IN THE SAME PROCESSES by the 70B (config3) leg.
N=${1:-2}; out=gpurun_out; mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 32 --warmup 4 --extras all > $out/bench_extras_n$N.json 2> $out/bench_extras_n$N.err
python - $out/bench_extras_n$N.json $out/bench_extras_n$N.err <<'PY'
import json, sys
txt = open(sys.argv[1]).read().strip().splitlines()
print("stdout lines:", len(txt))
try:
    d = json.loads(txt[-1])
    print("8B N", d["n_gpus"], "tok/s", round(d["value"], 2), "sha", d.get("tokens_sha"), "parity", (d.get("parity") or {}).get("ok"))
    c3 = d.get("config3")
    print("config3:", {k: c3[k] for k in ("value", "n_gpus", "ms_per_step", "frac_of_single_stream_ceiling", "tokens_sha", "error") if k in c3} if c3 else None)
except Exception as e:
    print("FAILED", e); print(open(sys.argv[2]).read()[-2500:])
PY
