#!/bin/bash
OUT=gpurun_out/r2t
mkdir -p $OUT
N=${1:-2}
timeout 900 python -m pytest tests -m gpu -q -k "golden or midsize or sampled or multi or configs" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $OUT/pytest.log | cut -c1-300
timeout 300 python bench.py --workload c2 --steps 100 --warmup 5 --no-cpu --no-peaks --others none --no-parity > $OUT/c2_n1.json 2> $OUT/c2_n1.err; echo "c2 n1 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload c2 --steps 100 --warmup 5 --no-cpu --no-peaks --others none > $OUT/c2_n$N.json 2> $OUT/c2_n$N.err; echo "c2 n$N rc=$?"
python - <<PY
import json
for f in ("$OUT/c2_n1.json","$OUT/c2_n$N.json"):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, d["value"], d["ms_per_step"], (d["e2e"] or {}).get("value"), (d["parity"] or {}).get("pass"))
PY
