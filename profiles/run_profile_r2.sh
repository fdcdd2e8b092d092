#!/bin/bash
# Round-2 profile capture (run on the GPU box from the repo root, e.g. under gpurun):
#   1. launch list of the bench command itself (first 700 launches: upload + structure build + the first train steps of
#      the e2e leg, kernels launched one by one)                      -> gpurun_out/r2_launches.csv
#   2. DRAM bytes + duration of every launch of one C2 train step     -> gpurun_out/r2_traffic_ncu.csv
#   3. --set full of the top kernels at the widest layer (D = 78)     -> gpurun_out/r2_top_<name>.ncu-rep + raw csv
# Numbers taken under ncu are cold-cache and serialised: use the shares, not the absolutes.
set -u
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/r2_prof_l.log 2>&1
STEP="python profiles/prof_step.py 8192 1"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
    --csv --log-file $OUT/r2_traffic_ncu.csv $STEP > $OUT/r2_prof_t.log 2>&1
cap() {  # name regex skip
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o $OUT/r2_top_$1 $STEP > $OUT/r2_prof_$1.log 2>&1
  ncu -i $OUT/r2_top_$1.ncu-rep --page raw --csv > $OUT/r2_top_$1_raw.csv 2>/dev/null
}
# the step runs the five layers in order forward (D = 14 .. 78) and in reverse order backward: launch 22 of 25 forward
# iterations and launch 2 of the backward ones belong to the widest layer
cap fwd78   'rows_tma_kernel<.int.0>'   22
cap dx78    'rows_tma_kernel<.int.1>'   2
cap dw78    'dw_tma_kernel'             2
cap agg78   'agg_stats_kernel<.int.2, .bool.0, .bool.1' 22
cap dz78    'dz_kernel'                 2
ls -la $OUT | tail -24
