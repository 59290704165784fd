#!/bin/bash
# Standard battery run on the GPU box through gpurun; writes everything under gpurun_out/<tag>/
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log
python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
python bench.py --steps ${STEPS:-50} --warmup 5 > $OUT/bench.json 2> $OUT/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_bench.log 2>&1
tail -5 $OUT/smoke.log; tail -15 $OUT/pytest_gpu.log; cat $OUT/bench.json; tail -3 $OUT/bench.err; cat $OUT/bench_ref.json; tail -3 $OUT/bench_ref.err
