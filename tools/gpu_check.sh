#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, bench, the ncu launch list of a bench step and one full capture of the
# dominant kernels.  Usage (repo root, under gpurun):  bash tools/gpu_check.sh [quick]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -3
echo "== pytest gpu" ; timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
echo "== bench" ; timeout 400 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-400
tail -3 gpurun_out/bench.err
if [ "${1:-}" != "quick" ]; then
  echo "== ncu launch list (3 warm-up steps skipped by -s, then ~2 steps)"
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 80 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 5 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
  python tools/summarize_launches.py gpurun_out/launches.csv | tee gpurun_out/launch_summary.txt | head -24
  echo "== ncu full capture: scan kernels"
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:gru_scan -s 4 -c 2 -o gpurun_out/scan_full \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  tail -2 gpurun_out/ncu_full.log
fi
