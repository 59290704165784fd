#!/bin/bash
OUT=gpurun_out/r2y
mkdir -p $OUT
N=${1:-8}
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err ) 2> $OUT/bench_n$N.time; echo "bench rc=$?"
tail -3 $OUT/bench_n$N.time; grep -v "OMP_NUM\|\*\*\*\*" $OUT/bench_n$N.err | tail -5
python - <<PY
import json
d=json.loads([l for l in open("$OUT/bench_n$N.json") if l.startswith("{")][-1])
print("C5", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["families_ms_per_step"], d["parity"]["pass"])
for x in d["others"]:
    print(x["workload"], x["value"], x["ms_per_step"], "e2e", (x["e2e"] or {}).get("value"), x["roofline"]["families_ms_per_step"], x["parity"].get("pass"), x["parity"].get("objective_max_rel_err"), x["parity"].get("factor_rel_fro"), x["parity"].get("error"))
PY
