#!/bin/bash
OUT=gpurun_out/r2h
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -6 $OUT/pytest.log
run() { name=$1; shift
timeout 600 python bench.py "$@" --steps 20 --warmup 3 --no-cpu --no-peaks --others none > $OUT/$name.json 2> $OUT/$name.err; echo "$name rc=$?"; tail -2 $OUT/$name.err
python - <<PY
import json
d=json.loads([l for l in open("$OUT/$name.json") if l.startswith("{")][-1])
print("$name", d["value"], d["ms_per_step"], "e2e", (d["e2e"] or {}).get("value"), d["roofline"]["families_ms_per_step"], (d["parity"] or {}).get("pass"), (d["parity"] or {}).get("objective_max_rel_err"), (d["parity"] or {}).get("factor_rel_fro"), (d["parity"] or {}).get("error"))
PY
}
run c2_f64 --workload c2 --dtype float64
run c3 --workload c3
run c1 --workload c1
run c2 --workload c2
run c4 --workload c4 --no-e2e
