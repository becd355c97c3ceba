#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + full capture of the hot kernels.
# Usage (from the repo root, on the GPU box): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
# launch list of the bench command (one metric, no clock control): shares of the step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
# full capture of the dominant kernels (smaller resident stack: ncu replays each launch ~40x)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_kernel --launch-skip 8 -c 3 \
  -o $OUT/score_full -f python bench.py --steps 1 --warmup 1 --particles 8192 --no-e2e --no-cpu-baseline > $OUT/ncu_score.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:insert_kernel -c 2 \
  -o $OUT/insert_full -f python bench.py --steps 1 --warmup 1 --particles 8192 --no-e2e --no-cpu-baseline > $OUT/ncu_insert.log 2>&1
ls -la $OUT
