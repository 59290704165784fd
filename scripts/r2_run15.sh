#!/bin/bash
OUT=gpurun_out/r2s
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -k "safe_solve or configs or sampled or api" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $OUT/pytest.log | cut -c1-300
run() { name=$1; shift
timeout 600 python bench.py "$@" --steps 10 --warmup 3 --no-cpu --no-peaks --others none > $OUT/$name.json 2> $OUT/$name.err; echo "$name rc=$?"; tail -2 $OUT/$name.err | grep -v -i warn
python - <<PY
import json
d=json.loads([l for l in open("$OUT/$name.json") if l.startswith("{")][-1])
print("$name", d["value"], d["ms_per_step"], d["roofline"]["families_ms_per_step"], (d["parity"] or {}).get("pass"), (d["parity"] or {}).get("factor_rel_fro"), (d["parity"] or {}).get("error"))
PY
}
run c4 --workload c4 --no-e2e
run c3 --workload c3 --no-e2e --no-parity
