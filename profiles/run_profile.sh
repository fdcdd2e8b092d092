#!/bin/bash
# Round-1 profile capture (run on the GPU box from the repo root, e.g. under gpurun):
#   1. launch list of one C2 train step          -> gpurun_out/r1_final_launches.csv
#   2. DRAM bytes of every GEMM launch           -> gpurun_out/r1_gemm_traffic_ncu.csv
#   3. --set full of the top kernels (widest layer) -> gpurun_out/top_<name>.ncu-rep + raw csv
# Numbers taken under ncu are cold-cache and serialised: use the shares, not the absolutes.
set -u
OUT=gpurun_out
mkdir -p $OUT
STEP="python profiles/prof_step.py 8192 1"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r1_final_launches.csv $STEP > $OUT/prof_l.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
    -k regex:gemm --csv --log-file $OUT/r1_gemm_traffic_ncu.csv $STEP > $OUT/prof_t.log 2>&1
cap() {  # name regex skip
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o $OUT/top_$1 $STEP > $OUT/prof_$1.log 2>&1
  ncu -i $OUT/top_$1.ncu-rep --page raw --csv > $OUT/top_$1_raw.csv 2>/dev/null
}
cap fwd80   'gemm_rows_tc_kernel<.int.80, .bool.1'        2
cap dx80    'gemm_rows_tc_kernel<.int.80, .bool.0'        3
cap dx64x2  'gemm_rows_tc_kernel<.int.64, .bool.0, .int.2' 2
cap dw      'gemm_dw_tc_kernel'                           2
cap agg     'agg_stats_kernel<.int.2, .bool.0'            22
cap dz      'dz_kernel'                                   2
ls -la $OUT | tail -20
