#!/usr/bin/env bash
# 70B (fits one B200: 141 GB bf16) decode A/B between library builds: bash bench_tools/ab70.sh lib1 lib2 ...
out=gpurun_out; mkdir -p $out
i=0
for lib in "$@"; do
  i=$((i+1)); envs=""
  [ "$lib" != "-" ] && envs="CAKE_B200_LIB=$PWD/cake_b200/$lib"
  env $envs timeout 600 python bench.py --model 70b --steps 16 --warmup 3 --e2e-steps 4 --extras none --no-isolated --no-parity --no-cpu > $out/ab70_$i.json 2> $out/ab70_$i.err
  python - "$lib" $out/ab70_$i.json $out/ab70_$i.err <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2])); print(f"70B {sys.argv[1]:28s} tok/s {d['value']:.3f}  ms {d['ms_per_step']:.3f} frac {d['roofline']['frac']}  of ceiling {d['value']/d['token_roofline']['roofline_tok_s_single_stream']:.4f} sha {d.get('tokens_sha')}")
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[3]).read()[-1200:])
PY
done
