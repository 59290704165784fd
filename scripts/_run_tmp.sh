mkdir -p gpurun_out/r3u
N=8
run() { name=$1; port=$2; shift; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@" > gpurun_out/r3u/${name}_n$N.json 2> gpurun_out/r3u/${name}_n$N.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r3u/${name}_n$N.json').read().strip().splitlines()[-1])
    print(d['n_gpus'], d['value'], d['ms_per_step'], d.get('e2e') and d['e2e']['value'], d['roofline']['families_ms_per_step'])
except Exception as e:
    print('no json', e)
PY
tail -2 gpurun_out/r3u/${name}_n$N.err | cut -c1-300; }
run c2 29531 --steps 50 --warmup 5
run c5_s04 29532 --workload c5 --scale 0.4 --steps 5 --warmup 3 --no-e2e
