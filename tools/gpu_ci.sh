#!/bin/bash
# GPU-box routine: tests, bench (own arm + reference arm).  Logs land in gpurun_out/.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
python bench.py --steps ${STEPS:-10} --warmup 3 2> gpurun_out/bench_err.log | tee gpurun_out/bench.json
tail -5 gpurun_out/bench_err.log
python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/bench_err.log | tee gpurun_out/bench_ref.json
