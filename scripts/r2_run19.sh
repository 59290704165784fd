#!/bin/bash
OUT=gpurun_out/r2w
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -k "hessian_mma or configs or sampled" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $OUT/pytest.log | cut -c1-300
run() { name=$1; shift
timeout 600 python bench.py "$@" --steps 10 --warmup 3 --no-cpu --no-peaks --others none --no-e2e --no-parity > $OUT/$name.json 2> $OUT/$name.err; echo "$name rc=$?"; tail -2 $OUT/$name.err | grep -v -i warn
python - <<PY
import json
d=json.loads([l for l in open("$OUT/$name.json") if l.startswith("{")][-1])
print("$name", d["value"], d["ms_per_step"], d["roofline"]["families_ms_per_step"])
PY
}
run c4 --workload c4
run c4_big --workload c4 --scale 0.05 --col-scale 0.1 --steps 3
