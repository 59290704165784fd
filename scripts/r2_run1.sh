#!/bin/bash
OUT=gpurun_out/r2a
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $OUT/pytest.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err ) 2> $OUT/bench.time; echo "bench rc=$?"
tail -3 $OUT/bench.time
tail -c 6000 $OUT/bench.json; tail -5 $OUT/bench.err
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/ref.json 2> $OUT/ref.err ) 2> $OUT/ref.time; echo "ref rc=$?"
tail -3 $OUT/ref.time; tail -c 1500 $OUT/ref.json; tail -3 $OUT/ref.err
