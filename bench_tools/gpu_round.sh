#!/usr/bin/env bash
# One gpurun call that collects everything a kernel iteration needs (run from the repo root on the GPU box):
#   gpurun --timeout 1500 -- 'bash bench_tools/gpu_round.sh [tests|bench|probes|ncu|all]'
# Results land in gpurun_out/ (merged back by gpurun).  Every step runs under its own `timeout` so that a hung kernel
# costs minutes, not the call's whole limit.
set -u
what="${1:-all}"
out=gpurun_out
mkdir -p "$out"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -ccbin /usr/bin/g++ -I cake_b200/csrc"

step() { echo "=== $1" | tee -a "$out/round.log"; }

if [[ "$what" == tests || "$what" == all ]]; then
  step "pytest -m gpu"
  timeout 900 python -m pytest tests -m gpu -x -q > "$out/pytest_gpu.log" 2>&1; echo "exit $?" >> "$out/pytest_gpu.log"
  tail -3 "$out/pytest_gpu.log" | tee -a "$out/round.log"
fi
if [[ "$what" == bench || "$what" == all ]]; then
  step "bench N=1"
  timeout 600 python bench.py --gpus 1 --steps 128 --warmup 8 > "$out/bench_n1.json" 2> "$out/bench_n1.err"
  tail -c 600 "$out/bench_n1.json" | tee -a "$out/round.log"
  step "per-phase trace of the decode megakernel"
  timeout 300 python bench_tools/mega_trace.py > "$out/mega_trace.txt" 2>&1; tail -12 "$out/mega_trace.txt" | tee -a "$out/round.log"
fi
if [[ "$what" == probes || "$what" == all ]]; then
  step "phase-boundary probes"
  $NVCC $ARCH -o /tmp/barrier_probe bench_tools/barrier_probe.cu && timeout 120 /tmp/barrier_probe 2000 > "$out/barrier_probe.txt" 2>&1
  $NVCC $ARCH -o /tmp/l2_prefetch_probe bench_tools/l2_prefetch_probe.cu && { timeout 120 /tmp/l2_prefetch_probe 32 8; timeout 120 /tmp/l2_prefetch_probe 24 4;
    for idle in 0 100 1000 3000; do echo "--- HBM idle ${idle} us"; timeout 120 /tmp/l2_prefetch_probe 32 $idle | grep "mode 0"; done; } > "$out/l2_prefetch_probe.txt" 2>&1
  $NVCC $ARCH -o /tmp/attn_tile_probe bench_tools/attn_tile_probe.cu && timeout 60 /tmp/attn_tile_probe 200 > "$out/attn_tile_probe.txt" 2>&1
  cat "$out/barrier_probe.txt" "$out/l2_prefetch_probe.txt" "$out/attn_tile_probe.txt" | tee -a "$out/round.log"
fi
if [[ "$what" == ncu || "$what" == all ]]; then
  step "ncu launch list + full capture of the decode kernel (numbers under ncu are never bench values)"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$out/launches.csv" \
      python bench.py --gpus 1 --steps 2 --warmup 1 --e2e-steps 2 --no-cpu > "$out/ncu_bench.log" 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_mega_kernel -c 1 -o "$out/mega_full" -f \
      python bench.py --gpus 1 --steps 2 --warmup 1 --e2e-steps 2 --no-cpu >> "$out/ncu_bench.log" 2>&1
  ls -la "$out" | tee -a "$out/round.log"
fi
echo done | tee -a "$out/round.log"
