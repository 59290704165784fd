#!/bin/bash
OUT=gpurun_out/r2b
mkdir -p $OUT
python scripts/debug_c2_parity.py 2>&1 | grep "float32" > $OUT/c2_debug.log; cat $OUT/c2_debug.log
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -15 $OUT/pytest.log
for u in 8 4; do
timeout 300 python bench.py --workload c3 --steps 20 --warmup 3 --no-e2e --no-cpu --no-parity --no-peaks --opt spmm_unroll=$u > $OUT/c3_u$u.json 2> $OUT/c3_u$u.err; echo "c3 u$u rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/c3_u$u.json")); print(d["value"], d["ms_per_step"], d["roofline"]["families_ms_per_step"], d["roofline"].get("l2_gather"))
PY
done
