#!/bin/bash
# ncu --set full capture of one kernel of the bench step (small resident stack). Usage: tools/ncu_kernel.sh <regex> <tag> [skip] [count]
K=$1; TAG=$2; SKIP=${3:-6}; CNT=${4:-2}
mkdir -p gpurun_out/$TAG
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip $SKIP -c $CNT \
  -o gpurun_out/$TAG/$TAG -f python bench.py --steps 1 --warmup 1 --particles 4096 --no-e2e --no-cpu-baseline > gpurun_out/$TAG/ncu.log 2>&1
tail -3 gpurun_out/$TAG/ncu.log
