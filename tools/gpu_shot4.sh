#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python -m pytest tests/test_gpu_vio.py -x -q > $O/pytest_vio.log 2>&1; echo "vio rc=$?" >> $O/pytest_vio.log; tail -2 $O/pytest_vio.log
timeout 300 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
cut -c1-300 $O/bench_default.json
timeout 200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -2 $O/pytest_gpu.log
