#!/bin/bash
# ncu evidence for round 2: launch list of the default bench command (our kernels), one --set full capture per dominant kernel
OUT=gpurun_out/r2p
mkdir -p $OUT
COMMON="--no-e2e --no-cpu --no-parity --no-peaks"
# 1. launch list of the headline workload (same command as the bench, fewer steps): per-launch durations of our kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:pycmf -c 600 --csv \
    --log-file $OUT/launches_c5.csv python bench.py --steps 2 --warmup 3 --others none $COMMON > $OUT/launches_c5.log 2>&1; echo "launch list c5 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:pycmf -c 800 --csv \
    --log-file $OUT/launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 3 --others none $COMMON > $OUT/launches_c3.log 2>&1; echo "launch list c3 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:pycmf -c 800 --csv \
    --log-file $OUT/launches_c2.csv python bench.py --workload c2 --steps 2 --warmup 3 --others none $COMMON > $OUT/launches_c2.log 2>&1; echo "launch list c2 rc=$?"
# 2. full captures
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tc_mu_kernel -s 2 -c 2 -o $OUT/c5_tc_mu -f \
    python bench.py --steps 2 --warmup 3 --others none $COMMON > $OUT/ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_nzb_kernel -s 2 -c 2 -o $OUT/c3_spmm_nzb -f \
    python bench.py --workload c3 --steps 2 --warmup 3 --others none $COMMON > $OUT/ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:safe_solve_kernel -s 1 -c 1 -o $OUT/c4_safe_solve -f \
    python bench.py --workload c4 --steps 1 --warmup 3 --others none $COMMON > $OUT/ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dmma_resid_kernel -s 2 -c 2 -o $OUT/c2_f64_dmma_resid -f \
    python bench.py --workload c2 --dtype float64 --steps 2 --warmup 3 --others none $COMMON > $OUT/ncu_c2f64.log 2>&1; echo "ncu c2 f64 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dmma_gemm_kernel -s 2 -c 2 -o $OUT/c5_f64_dmma_gemm -f \
    python bench.py --workload c5 --scale 0.05 --dtype float64 --steps 2 --warmup 3 --others none $COMMON > $OUT/ncu_c5f64.log 2>&1; echo "ncu c5 f64 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_pass_kernel -s 2 -c 2 -o $OUT/c2_tc_pass -f \
    python bench.py --workload c2 --steps 2 --warmup 3 --others none $COMMON > $OUT/ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
ls -la $OUT
