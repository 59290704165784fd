#!/bin/bash
OUT=gpurun_out/r2q
mkdir -p $OUT
COMMON="--no-e2e --no-cpu --no-parity --no-peaks --others none"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:safe_solve_kernel -s 1 -c 1 -o $OUT/c4_safe_solve_tri -f \
    python bench.py --workload c4 --steps 1 --warmup 3 $COMMON > $OUT/ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tc_mu_kernel -s 9 -c 3 -o $OUT/c5_tc_mu_big -f \
    python bench.py --steps 2 --warmup 3 $COMMON > $OUT/ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_nzb2_kernel -s 2 -c 2 -o $OUT/c3_spmm_nzb2 -f \
    python bench.py --workload c3 --steps 2 --warmup 3 $COMMON > $OUT/ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:pycmf -c 800 --csv \
    --log-file $OUT/launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 3 $COMMON > $OUT/launches_c3.log 2>&1; echo "launch list c3 rc=$?"
ls -la $OUT
