#!/bin/bash
OUT=gpurun_out/r2r
mkdir -p $OUT
COMMON="--no-e2e --no-cpu --no-parity --no-peaks --others none"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:safe_solve_kernel -s 1 -c 1 -o $OUT/c4_safe_solve_tri2 -f \
    python bench.py --workload c4 --steps 1 --warmup 3 $COMMON > $OUT/ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
timeout 300 python -m pytest tests/test_gpu_api.py -q -k "transform or fit_with" > $OUT/pytest_api.log 2>&1; echo "api rc=$?"; tail -3 $OUT/pytest_api.log
