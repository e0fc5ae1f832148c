#!/bin/bash
# One gpurun call (1 GPU): the final-state GPU test suite and the default bench line.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 480 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
cut -c1-400 $O/bench_default.json
