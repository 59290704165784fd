mkdir -p gpurun_out/r3n
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "tc_mu_wide or midsize" > gpurun_out/r3n/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3n/pytest.log
timeout 200 python scripts/tc_mu_bench.py 20000 50000 256,128,64 > gpurun_out/r3n/tc_mu_bench.txt 2>&1; tail -4 gpurun_out/r3n/tc_mu_bench.txt
timeout 200 python scripts/tc_mu_trace.py 256 > gpurun_out/r3n/trace256.txt 2>&1; grep "steady\|converter" gpurun_out/r3n/trace256.txt
