#!/usr/bin/env bash
# Multi-GPU iteration (gpurun --gpus N): 2-GPU parity tests, then the sharded bench for 8B and 70B.
N=${1:-2}; out=gpurun_out; mkdir -p $out
if ! timeout 150 python -X faulthandler -c "import faulthandler; faulthandler.dump_traceback_later(100, exit=True); import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; then echo "SMOKE FAILED"; tail -15 $out/smoke.log; exit 1; fi
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -s > $out/pytest_multi.log 2>&1; grep -E "passed|failed|skipped|70B widths|Error" $out/pytest_multi.log | tail -6
for model in 8b 70b; do
  steps=128; [ $model = 70b ] && steps=32
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --model $model --steps $steps --warmup 8 --extras none > $out/bench_${model}_n$N.json 2> $out/bench_${model}_n$N.err
  python - $out/bench_${model}_n$N.json $out/bench_${model}_n$N.err <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(d["metric"], "N", d["n_gpus"], "tok/s", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "p50/p99/max ms", round(d["e2e"]["ms_per_step_p50"], 3), round(d["e2e"]["ms_per_step_p99"], 3), round(d["e2e"]["ms_per_step_max"], 3),
          "sha", d.get("tokens_sha"), "parity", (d.get("parity") or {}).get("ok"), "trace", json.dumps(d.get("steps_trace"))[:600])
except Exception as e:
    print("FAILED", e); print(open(sys.argv[2]).read()[-2000:])
PY
done
