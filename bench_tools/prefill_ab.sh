#!/usr/bin/env bash
# tcgen05 GEMM tile A/B on BASELINE configs[4] (bs=32 x 4096 prefill): 128x256 (default) vs 128x128 tiles
out=gpurun_out; mkdir -p $out
for bn in 256 128; do
  CAKE_B200_TC_BN=$bn timeout 600 python bench.py --workload prefill --steps 2 > $out/prefill_bn$bn.json 2> $out/prefill_bn$bn.err
  python - $bn $out/prefill_bn$bn.json $out/prefill_bn$bn.err <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2])); print(f"BN={sys.argv[1]}: {d['value']:.0f} tok/s  {d['ms_per_step']:.1f} ms  {d['roofline']['achieved']} TFLOP/s  frac {d['roofline']['frac']} finite {d['finite']}")
except Exception as e:
    print("BN", sys.argv[1], "FAILED", e); print(open(sys.argv[3]).read()[-1500:])
PY
done
