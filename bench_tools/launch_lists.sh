#!/usr/bin/env bash
# ncu launch lists (per-launch gpu__time_duration, cold cache, serialised) of the decode bench and of one prefill layer
out=gpurun_out; mkdir -p $out
K='regex:decode_mega|gemm_tc|attn_|rmsnorm|rope_|embed_|sample_|gemv|argmax|set_'
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 300 --csv --log-file $out/launches_decode.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu --no-parity --no-isolated --extras none > $out/launches_decode.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 60 --csv --log-file $out/launches_prefill.csv \
  python bench_tools/prefill_bench.py 1 32 4096 > $out/launches_prefill.log 2>&1
python - <<'PY'
import csv, collections
for name in ("decode", "prefill"):
    try:
        rows = [r for r in csv.reader(open(f"gpurun_out/launches_{name}.csv")) if len(r) > 5 and r[0].isdigit()]
    except Exception as e:
        print(name, "failed", e); continue
    agg = collections.OrderedDict()
    for r in rows:
        us = float(r[-1].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[-2], 1e-3)
        k = r[4][:70]; agg.setdefault(k, []).append(us)
    print(f"== {name}: {len(rows)} launches")
    for k, v in agg.items(): print(f"  {k:70s} n={len(v):3d} mean {sum(v)/len(v):10.1f} us  total {sum(v):10.1f} us")
PY
