#!/bin/bash
OUT=gpurun_out/r2l
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -k "safe_solve or spmm or configs or midsize or sampled or golden" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -12 $OUT/pytest.log | cut -c1-400
run() { name=$1; shift
timeout 600 python bench.py "$@" --steps 20 --warmup 3 --no-cpu --no-peaks --others none > $OUT/$name.json 2> $OUT/$name.err; echo "$name rc=$?"; tail -2 $OUT/$name.err
python - <<PY
import json
d=json.loads([l for l in open("$OUT/$name.json") if l.startswith("{")][-1])
print("$name", d["value"], d["ms_per_step"], "e2e", (d["e2e"] or {}).get("value"), d["roofline"]["families_ms_per_step"], (d["parity"] or {}).get("pass"), (d["parity"] or {}).get("objective_max_rel_err"), (d["parity"] or {}).get("factor_rel_fro"), (d["parity"] or {}).get("error"))
PY
}
run c4 --workload c4 --no-e2e
run c3 --workload c3 --no-e2e
run c3_u8 --workload c3 --no-e2e --opt spmm_unroll=8
run c3_nolean --workload c3 --no-e2e --opt spmm_lean=0
