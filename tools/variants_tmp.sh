B="python bench.py --particles 16384 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline"
P='import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["stage_ms_per_step"]["score_kernels"],2), round(d["ms_per_step"],2))'
for v in mb5 mb4; do CSPB_LIB=$PWD/pyp_b200/libcspb200_$v.so $B 2>/dev/null | tail -1 | python -c "$P" $v; done
$B 2>/dev/null | tail -1 | python -c "$P" default_mb6
$B 2>/dev/null | tail -1 | python -c "$P" default_mb6
