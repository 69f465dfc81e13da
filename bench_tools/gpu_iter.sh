#!/usr/bin/env bash
# One kernel iteration on the GPU box, hang-safe: a 2-minute smoke gate first; nothing else runs if it fails.
#   gpurun --timeout 1500 -- 'bash bench_tools/gpu_iter.sh [tests] [bench] [trace] [alltests]'
set -u
out=gpurun_out; mkdir -p $out
if ! timeout 150 python -X faulthandler -c "import faulthandler, sys; faulthandler.dump_traceback_later(100, exit=True); import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; then
  echo "SMOKE FAILED / HUNG"; tail -15 $out/smoke.log; exit 1
fi
tail -1 $out/smoke.log
for what in "$@"; do
  case $what in
    tests) timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_l8_full.py -x -q -s > $out/pytest_core.log 2>&1; grep -E "passed|failed|ulp|Error" $out/pytest_core.log | tail -25;;
    alltests) timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -5 $out/pytest_gpu.log;;
    bench) timeout 400 python bench.py --extras none --no-isolated > $out/bench_iter.json 2> $out/bench_iter.err
           python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_iter.json"))
    print("tok/s", round(d["value"], 2), "ms", round(d["ms_per_step"], 4), "frac", d["roofline"]["frac"], "e2e", round(d["e2e"]["value"], 2),
          "parity_ok", d["parity"]["ok"], [s["max_logit_err_ulp"] for s in d["parity"]["steps"]], "sha", d["tokens_sha"], "cpu", d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("bench failed:", e); print(open("gpurun_out/bench_iter.err").read()[-1500:])
PY
           ;;
    trace) timeout 300 python bench_tools/mega_trace.py 8 > $out/mega_trace_iter.txt 2>&1; tail -16 $out/mega_trace_iter.txt;;
  esac
done
