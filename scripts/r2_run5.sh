#!/bin/bash
OUT=gpurun_out/r2e
mkdir -p $OUT
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_api.py -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -12 $OUT/pytest.log
for sl in 2; do
PYCMF_B200_V_SLABS=$sl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --others c3 --no-e2e --no-parity --no-peaks > $OUT/bench_n${N}_sl$sl.json 2> $OUT/bench_n${N}_sl$sl.err; echo "bench slabs=$sl rc=$?"
tail -3 $OUT/bench_n${N}_sl$sl.err
python - <<PY
import json
d=json.loads([l for l in open("$OUT/bench_n${N}_sl$sl.json") if l.startswith("{")][-1])
print("C5", d["value"], d["ms_per_step"], d["roofline"]["families_ms_per_step"])
for x in d["others"]:
    print(x["workload"], x["value"], x["ms_per_step"], x["roofline"]["families_ms_per_step"])
PY
done
