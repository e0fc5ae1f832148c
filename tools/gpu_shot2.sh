#!/bin/bash
# One gpurun --gpus 2 call: the bench under torchrun (replicas, 8 streams per GPU, configs[4] stream feature-sharded) and the
# sharded end-to-end test on the configs[4] stream.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
O=gpurun_out

timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 2 --steps 100 --warmup 10 > $O/bench_gpus2.json 2> $O/bench_gpus2.err; echo "bench2 rc=$?"
tail -c 3000 $O/bench_gpus2.json
RVIO_TEST_CONFIG=4 RVIO_TEST_FRAMES=58 RVIO_TEST_ORACLE=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29612 tests/dist_sharded_vio.py > $O/sharded_vio_c4.log 2>&1; echo "sharded c4 rc=$?"
tail -3 $O/sharded_vio_c4.log
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q > $O/pytest_multi_gpu2.log 2>&1; echo "multi rc=$?"
tail -3 $O/pytest_multi_gpu2.log
