#!/usr/bin/env bash
# tcgen05 GEMM A/B on BASELINE configs[4] (bs=32 x 4096 prefill): specs "BN MC", e.g. "256 1" "256 0" "128 0"
out=gpurun_out; mkdir -p $out
for spec in "$@"; do
  set -- $spec; bn=$1; mc=$2
  CAKE_B200_TC_BN=$bn CAKE_B200_TC_MC=$mc timeout 400 python bench.py --workload prefill --steps 2 > $out/prefill_bn${bn}_mc$mc.json 2> $out/prefill_bn${bn}_mc$mc.err
  python - "$spec" $out/prefill_bn${bn}_mc$mc.json $out/prefill_bn${bn}_mc$mc.err <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2])); print(f"BN MC = {sys.argv[1]}: {d['value']:.0f} tok/s  {d['ms_per_step']:.1f} ms  {d['roofline']['achieved']} TFLOP/s  frac {d['roofline']['frac']} finite {d['finite']}")
except Exception as e:
    print("BN MC", sys.argv[1], "FAILED", e); print(open(sys.argv[3]).read()[-1500:])
PY
done
