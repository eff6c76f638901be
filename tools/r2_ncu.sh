#!/bin/bash
# ncu captures of the round's kernels on one B200 (run under gpurun; reports land in gpurun_out/, summarised in the
# build container with tools/ncu_summary.py).  Every capture follows a plain run of the same command that exited 0.
#   bash tools/r2_ncu.sh [tag]
set -u
TAG=${1:-r2f}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on"

python tools/r2_probe.py asm dmg p1 pa tab vec > $OUT/${TAG}_probe.json 2> $OUT/${TAG}_probe.err || { echo "probe failed"; tail -5 $OUT/${TAG}_probe.err; exit 1; }
cat $OUT/${TAG}_probe.json

cap() {  # name, kernel regex, launch-skip, probe args...
   local name=$1 rx=$2 skip=$3
   shift 3
   $NCU -k "regex:$rx" --launch-skip $skip -c 1 -f -o $OUT/${TAG}_$name python tools/r2_probe.py "$@" > $OUT/${TAG}_$name.log 2>&1 || echo "capture $name failed"
}
cap asm_p2_n1448 'assemble_fast_kernel' 20 asm
cap dmg_gather_100pct 'assemble_fast_kernel' 4 dmg100
cap dmg_prepass_100pct 'cell_setup_damage_kernel' 3 dmg100
cap asm_p1_n2896 'assemble_fast_kernel' 20 p1
cap pa_q2_n2048 'pa_tile_kernel' 5 pa
cap tabulate_p2 'tabulate_kernel' 4 tab
cap residual_cells 'cell_residual_kernel' 4 vec
cap residual_gather 'vector_gather_kernel' 4 vec

# config 4 (n = 5792): launch list of the bench step, then the two dominant kernels
B="python bench.py --steps 2 --warmup 3 --skip-extras --no-cpu-baseline --e2e-steps 1"
$B > $OUT/${TAG}_bench_short.json 2> $OUT/${TAG}_bench_short.err || { echo "bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file $OUT/${TAG}_launches.csv $B > $OUT/${TAG}_launches.log 2>&1
$NCU -k 'regex:assemble_fast_kernel' --launch-skip 2 -c 1 -f -o $OUT/${TAG}_asm_n5792 $B > $OUT/${TAG}_asm_n5792.log 2>&1
$NCU -k 'regex:spmv_tma_kernel' --launch-skip 5 -c 1 -f -o $OUT/${TAG}_spmv_n5792 $B > $OUT/${TAG}_spmv_n5792.log 2>&1
ls -la $OUT | tail -20
