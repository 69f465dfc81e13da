#!/usr/bin/env bash
# A/B runs of library variants (profiling aid): bash bench_tools/ab.sh "<lib|-> [ENV=V ...]" ...
set -u
out=gpurun_out; mkdir -p $out
if ! timeout 150 python -X faulthandler -c "import faulthandler, sys; faulthandler.dump_traceback_later(100, exit=True); import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; then echo "SMOKE FAILED"; tail -15 $out/smoke.log; exit 1; fi
i=0
for spec in "$@"; do
  i=$((i+1))
  set -- $spec
  lib=$1; shift
  envs="$*"
  [ "$lib" != "-" ] && envs="$envs CAKE_B200_LIB=$PWD/cake_b200/$lib"
  env $envs timeout 300 python bench.py --extras none --no-isolated --no-parity --no-cpu --e2e-steps 4 > $out/ab_$i.json 2> $out/ab_$i.err
  python - "$spec" $out/ab_$i.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2])); print(f"{sys.argv[1]:60s} tok/s {d['value']:.2f}  frac {d['roofline']['frac']}  sha {d['tokens_sha']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
