mkdir -p gpurun_out/r3k
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r3k/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r3k/pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r3k/bench_c2_n2.json 2> gpurun_out/r3k/bench_c2_n2.err; tail -c 700 gpurun_out/r3k/bench_c2_n2.json | head -c 700; echo; tail -2 gpurun_out/r3k/bench_c2_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c5 --scale 0.2 --steps 5 --warmup 3 --no-e2e > gpurun_out/r3k/bench_c5_n2.json 2> gpurun_out/r3k/bench_c5_n2.err; head -c 400 gpurun_out/r3k/bench_c5_n2.json; echo; tail -2 gpurun_out/r3k/bench_c5_n2.err
timeout 300 python bench.py --workload c5 --scale 0.2 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r3k/bench_c5_n1.json 2> gpurun_out/r3k/bench_c5_n1.err; head -c 400 gpurun_out/r3k/bench_c5_n1.json; echo; tail -2 gpurun_out/r3k/bench_c5_n1.err
