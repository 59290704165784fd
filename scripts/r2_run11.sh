#!/bin/bash
OUT=gpurun_out/r2k
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x -k "safe_solve or golden or midsize or sampled or configs or jacobi" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -15 $OUT/pytest.log
run() { name=$1; shift
timeout 600 python bench.py "$@" --steps 20 --warmup 3 --no-cpu --no-peaks --others none > $OUT/$name.json 2> $OUT/$name.err; echo "$name rc=$?"; tail -2 $OUT/$name.err
python - <<PY
import json
d=json.loads([l for l in open("$OUT/$name.json") if l.startswith("{")][-1])
print("$name", d["value"], d["ms_per_step"], "e2e", (d["e2e"] or {}).get("value"), d["roofline"]["families_ms_per_step"], (d["parity"] or {}).get("pass"), (d["parity"] or {}).get("objective_max_rel_err"), (d["parity"] or {}).get("factor_rel_fro"), (d["parity"] or {}).get("error"))
PY
}
run c4 --workload c4 --no-e2e
run c4_jacobi --workload c4 --no-e2e --opt solve_path=0
