#!/bin/bash
# One GPU-box visit: machine facts, parity tests, smoke, bench (both arms, every config), ncu launch list + full
# captures of the hot kernels.  Usage (from the repo root, on the GPU box): bash tools/gpu_round.sh <tag> [sections]
# sections (default "facts tests smoke bench ref configs launches ncu"): any subset, space separated.
TAG=${1:-r02}
SECTIONS=${2:-"facts tests smoke bench ref configs launches ncu"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
has() { [[ " $SECTIONS " == *" $1 "* ]]; }
if has facts; then
  { nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv; nproc; free -g; cat /sys/fs/cgroup/memory.max 2>/dev/null;
    lscpu | head -25; nvidia-smi topo -m; numactl -H 2>/dev/null | head; } > $OUT/facts.txt 2>&1
fi
if has tests; then
  ( time timeout 1500 python -m pytest tests -m gpu -q --durations=15 ) > $OUT/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $OUT/pytest_gpu.log
fi
if has smoke; then timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; fi
if has bench; then timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; fi
if has ref; then ( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err; fi
if has configs; then
  for C in C1 C3 C4 C5; do
    timeout 900 python bench.py --config $C --steps 3 --warmup 3 > $OUT/bench_$C.json 2> $OUT/bench_$C.err
  done
fi
if has launches; then
  # launch list of the bench command (one metric, no clock control): shares of the step
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-strong > $OUT/bench_under_ncu.log 2>&1
fi
if has ncu; then
  # full capture of the dominant kernels (smaller resident stack: ncu replays each launch ~40x)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_kernel --launch-skip 8 -c 3 \
    -o $OUT/score_full -f python bench.py --steps 1 --warmup 1 --particles 8192 --no-e2e --no-cpu-baseline --no-strong > $OUT/ncu_score.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:insert_kernel -c 2 \
    -o $OUT/insert_full -f python bench.py --steps 1 --warmup 1 --particles 8192 --no-e2e --no-cpu-baseline --no-strong > $OUT/ncu_insert.log 2>&1
fi
if has ncu_prep; then
  # the preprocessing passes (fused path: the second and later chunks)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fft_|image_stats|zero_slots" --launch-skip 12 -c 8 \
    -o $OUT/prep_full -f python bench.py --steps 1 --warmup 1 --particles 16384 --no-e2e --no-cpu-baseline --no-strong > $OUT/ncu_prep.log 2>&1
fi
if has multi; then
  # needs gpurun --gpus N: 2-GPU equality test of the product entry, H2D scaling of the platform, bench at every N
  NG=$(nvidia-smi -L | wc -l)
  ( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > $OUT/pytest_multi.log 2>&1
  for N in 1 2 4 8; do
    [ $N -le $NG ] || continue
    if [ $N -eq 1 ]; then
      timeout 300 python tools/h2d_scaling.py > $OUT/h2d_$N.json 2>&1
    else
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_scaling.py > $OUT/h2d_$N.json 2>&1
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_N$N.json 2> $OUT/bench_N$N.err
    fi
  done
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $NG --steps 5 --warmup 3 > $OUT/bench_ref_N$NG.json 2> $OUT/bench_ref_N$NG.err
fi
if has variants; then
  # A/B of kernel build variants on the same box: alternative libraries through CSPB_LIB (pyp_b200/_lib.py)
  for L in $(cd pyp_b200 && ls libcspb200*.so); do
    [ -f pyp_b200/$L ] || continue
    CSPB_LIB=$PWD/pyp_b200/$L timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-strong > $OUT/variant_$L.json 2> $OUT/variant_$L.err
  done
fi
if has e2e_debug; then
  CSPB_PIPE_DEBUG=1 timeout 600 python tools/time_e2e.py 32768 > $OUT/time_e2e.log 2>&1
fi
if has ncu_grad; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_grad_kernel --launch-skip 6 -c 3 \
    -o $OUT/grad_full -f python bench.py --steps 1 --warmup 1 --particles 8192 --no-e2e --no-cpu-baseline --no-strong > $OUT/ncu_grad.log 2>&1
fi
if has ncu_c4; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_kernel --launch-skip 8 -c 2 \
    -o $OUT/score_c4_full -f python bench.py --config C4 --steps 1 --warmup 1 --particles 512 --no-e2e --no-cpu-baseline > $OUT/ncu_score_c4.log 2>&1
fi
ls -la $OUT
