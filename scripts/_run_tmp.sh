mkdir -p gpurun_out/r3e
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "spmm or golden or midsize or sqerr" > gpurun_out/r3e/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r3e/pytest.log
for sc in 0.125 ; do
timeout 240 python bench.py --no-cpu --no-e2e --workload c3 --scale $sc --steps 10 --warmup 3 > gpurun_out/r3e/c3_$sc.json 2> gpurun_out/r3e/c3_$sc.err; tail -c 900 gpurun_out/r3e/c3_$sc.json; tail -3 gpurun_out/r3e/c3_$sc.err
done
timeout 240 python bench.py --no-cpu --no-e2e --workload c3 --scale 0.125 --steps 10 --warmup 3 --opt spmm_path=0 > gpurun_out/r3e/c3_generic.json 2> gpurun_out/r3e/c3_generic.err; tail -c 400 gpurun_out/r3e/c3_generic.json
