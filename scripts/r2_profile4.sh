#!/bin/bash
OUT=gpurun_out/r2z
mkdir -p $OUT
COMMON="--no-e2e --no-cpu --no-parity --no-peaks --others none"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:row_grad_hess_mma_kernel -s 2 -c 2 -o $OUT/c4_hess_mma -f \
    python bench.py --workload c4 --steps 1 --warmup 3 $COMMON > $OUT/ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:pycmf -c 600 --csv \
    --log-file $OUT/launches_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 $COMMON > $OUT/launches_c4.log 2>&1; echo "launch list c4 rc=$?"
