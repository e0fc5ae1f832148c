#!/bin/bash
# One gpurun call: GPU parity tests, PDL A/B, the default bench, the reference arm, an ncu launch list.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
timeout 300 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
RVIO_B200_PDL=1 timeout 200 python -m pytest tests/test_gpu_vio.py tests/test_gpu_tracker.py tests/test_gpu_edge_cases.py tests/test_gpu_updater.py -x -q > $O/pytest_gpu_pdl.log 2>&1; echo "pytest pdl rc=$?" >> $O/pytest_gpu_pdl.log
tail -3 $O/pytest_gpu_pdl.log
for p in 0 1 0 1; do
  RVIO_B200_PDL=$p timeout 150 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-extra-legs --batch-streams 0 > $O/ab_pdl${p}_$RANDOM.json 2> $O/ab_err.log
done
python - <<'P' > $O/ab_summary.txt
import json, glob
r = {0: [], 1: []}
for f in sorted(glob.glob('gpurun_out/ab_pdl*_*.json')):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        r[int(j['pdl'])].append((j['value'], j['e2e']['value'], j['e2e']['strict']['value'], j['e2e'].get('h2d_frame_us')))
    except Exception as e:
        print('bad', f, e)
print(r)
best = max((sum(v[0] for v in vals) / len(vals), k) for k, vals in r.items() if vals)
print('BEST', best[1])
P
cat $O/ab_summary.txt
BEST=$(grep BEST $O/ab_summary.txt | awk '{print $2}'); BEST=${BEST:-0}
RVIO_B200_PDL=$BEST timeout 420 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
cut -c1-600 $O/bench_default.json
timeout 150 python bench.py --impl reference --steps 60 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
cut -c1-300 $O/bench_reference.json
RVIO_B200_PDL=$BEST timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --batch-streams 0 --no-extra-legs > $O/ncu_bench.log 2>&1; echo "ncu rc=$?"
