mkdir -p gpurun_out/r3j
timeout 200 python scripts/e2e_breakdown.py 50 > gpurun_out/r3j/e2e.txt 2>&1; tail -6 gpurun_out/r3j/e2e.txt
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/r3j/bench.json 2> gpurun_out/r3j/bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r3j/bench.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e'])"; tail -3 gpurun_out/r3j/bench.err
