mkdir -p gpurun_out/r3v
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3v/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/r3v/smoke.log
timeout 200 python -m pytest tests -m gpu -q -x > gpurun_out/r3v/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r3v/pytest_gpu.log
