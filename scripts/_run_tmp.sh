bash scripts/gpu_battery.sh r3s > gpurun_out/r3s_battery.log 2>&1
tail -30 gpurun_out/r3s_battery.log | cut -c1-1500
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_pass_kernel -s 4 -c 2 -o gpurun_out/r3s/tc_pass python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r3s/ncu_tc_pass.log 2>&1; echo "ncu rc=$?"
