mkdir -p gpurun_out/r3r
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "safe_solve or midsize or golden or jacobi" > gpurun_out/r3r/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3r/pytest.log
timeout 240 python bench.py --no-cpu --no-e2e --workload c4 --scale 0.02 --steps 3 --warmup 3 > gpurun_out/r3r/c4_s002.json 2> gpurun_out/r3r/c4.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r3r/c4_s002.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['families_ms_per_step'], d['config']['objective_last'])
PY
