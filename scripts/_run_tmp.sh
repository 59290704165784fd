mkdir -p gpurun_out/r3t
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r3t/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3t/pytest_multi.log
N=4
run() { name=$1; port=$2; shift; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@" > gpurun_out/r3t/${name}_n$N.json 2> gpurun_out/r3t/${name}_n$N.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r3t/${name}_n$N.json').read().strip().splitlines()[-1])
    print(d['n_gpus'], d['value'], d['ms_per_step'], d.get('e2e') and d['e2e']['value'], d['roofline']['families_ms_per_step'], d['config']['objective_last'])
except Exception as e:
    print('no json', e)
PY
tail -2 gpurun_out/r3t/${name}_n$N.err | cut -c1-300; }
run c5_s02 29522 --workload c5 --scale 0.2 --steps 5 --warmup 3 --no-e2e
run c3_s0125 29523 --workload c3 --scale 0.125 --steps 10 --warmup 3 --no-e2e
