#!/bin/bash
OUT=gpurun_out/r2f
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py -q -k "api or dmma or spmm or gemm or golden" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
tail -12 $OUT/pytest.log
for dp in 1 0; do
timeout 600 python bench.py --workload c5 --scale 0.05 --dtype float64 --dense-path $dp --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity > $OUT/c5_f64_dp$dp.json 2> $OUT/c5_f64_dp$dp.err; echo "rc=$?"
tail -2 $OUT/c5_f64_dp$dp.err
python - <<PY
import json
d=json.loads([l for l in open("$OUT/c5_f64_dp$dp.json") if l.startswith("{")][-1])
print("C5 f64 scale .05 dense_path $dp", d["value"], d["ms_per_step"], d["roofline"]["families_ms_per_step"], d["peaks"]["measured_in_run"])
PY
done
