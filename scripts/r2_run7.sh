#!/bin/bash
OUT=gpurun_out/r2g
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -6 $OUT/pytest.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err ) 2> $OUT/bench.time; echo "bench rc=$?"
tail -3 $OUT/bench.time; tail -3 $OUT/bench.err
python - <<PY
import json
d=json.loads([l for l in open("$OUT/bench.json") if l.startswith("{")][-1])
print("C5", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["families_ms_per_step"], d["parity"]["pass"])
for x in d["others"]:
    print(x["workload"], x["value"], x["ms_per_step"], "e2e", x["e2e"]["value"], x["roofline"]["families_ms_per_step"], x["parity"].get("pass"), x["parity"].get("objective_max_rel_err"), x["parity"].get("factor_rel_fro"), x["parity"].get("error"))
PY
timeout 600 python bench.py --workload c2 --dtype float64 --steps 20 --warmup 3 --no-cpu --no-peaks --others none > $OUT/c2_f64.json 2> $OUT/c2_f64.err; echo "rc=$?"; tail -2 $OUT/c2_f64.err
python - <<PY
import json
d=json.loads([l for l in open("$OUT/c2_f64.json") if l.startswith("{")][-1])
print("C2 f64", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["families_ms_per_step"], d["parity"])
PY
