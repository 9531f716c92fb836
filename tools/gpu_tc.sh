#!/bin/bash
# tcgen05 bring-up visit: stand-alone kernel tests (bounded waits), then the fast pytest subset.
set -u
mkdir -p gpurun_out
for t in tools/_bin/*_test; do
  echo "== $t"; timeout 300 $t 2>&1 | tail -40 | tee gpurun_out/$(basename $t).txt
done
if [ "${1:-}" = "pytest" ]; then
  echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
fi
if [ "${1:-}" = "bench" ] || [ "${2:-}" = "bench" ]; then
  echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json; tail -3 gpurun_out/bench.err
fi
