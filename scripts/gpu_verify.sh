#!/bin/bash
OUT=gpurun_out/r2n
mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $OUT/pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $OUT/smoke.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err ) 2> $OUT/bench.time; echo "bench rc=$?"
tail -3 $OUT/bench.time; grep -v Warn $OUT/bench.err | tail -3
python - <<PY
import json
d=json.loads([l for l in open("$OUT/bench.json") if l.startswith("{")][-1])
print("C5", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline"]["families_ms_per_step"], d["parity"]["pass"], d["cpu_baseline"]["value"])
for x in d["others"]:
    print(x["workload"], x["value"], x["ms_per_step"], "e2e", (x["e2e"] or {}).get("value"), (x["cpu_baseline"] or {}).get("value"), x["roofline"]["families_ms_per_step"], x["parity"].get("pass"), x["parity"].get("objective_max_rel_err"), x["parity"].get("factor_rel_fro"), x["parity"].get("error"))
PY
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/ref.json 2> $OUT/ref.err ) 2> $OUT/ref.time; echo "ref rc=$?"; tail -3 $OUT/ref.time
