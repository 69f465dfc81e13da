#!/usr/bin/env bash
# prefill attention A/B on BASELINE configs[4] (bs=32 x 4096): specs "<CAKE_B200_FA> <CAKE_B200_FA_TAU>", e.g. "tc 5.545" "tc 0" "mma 0"
out=gpurun_out; mkdir -p $out
for spec in "$@"; do
  set -- $spec; fa=$1; tau=$2
  CAKE_B200_FA=$fa CAKE_B200_FA_TAU=$tau timeout 400 python bench.py --workload prefill --steps 2 > $out/prefill_fa_${fa}_$tau.json 2> $out/prefill_fa_${fa}_$tau.err
  python - "$spec" $out/prefill_fa_${fa}_$tau.json $out/prefill_fa_${fa}_$tau.err <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2])); print(f"FA TAU = {sys.argv[1]}: {d['value']:.0f} tok/s  {d['ms_per_step']:.1f} ms  {d['roofline']['achieved']} TFLOP/s  frac {d['roofline']['frac']} finite {d['finite']}")
except Exception as e:
    print("FA", sys.argv[1], "FAILED", e); print(open(sys.argv[3]).read()[-1500:])
PY
done
