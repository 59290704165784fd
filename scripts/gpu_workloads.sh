#!/bin/bash
# Benches of the non-default workloads (parity-test configs of BASELINE.json) on one GPU; output under gpurun_out/<tag>/
TAG=${1:-wl}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() { name=$1; shift; timeout ${TMO:-240} python bench.py --no-cpu --no-e2e "$@" > $OUT/$name.json 2> $OUT/$name.err; echo "$name rc=$?"; tail -c 1500 $OUT/$name.json; tail -2 $OUT/$name.err; }
run c1 --workload c1 --steps 100 --warmup 5
run c5_s01 --workload c5 --scale 0.1 --steps 5 --warmup 3
run c3_s0125 --workload c3 --scale 0.125 --steps 10 --warmup 3
run c4_s002 --workload c4 --scale 0.02 --steps 3 --warmup 3
