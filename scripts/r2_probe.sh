#!/bin/bash
OUT=gpurun_out/p1
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
free -g > $OUT/host.txt; nproc >> $OUT/host.txt; lscpu | head -20 >> $OUT/host.txt
timeout 60 scripts/tc_mu_pair.bin > $OUT/pair.log 2>&1; echo "pair rc=$?" >> $OUT/pair.log
timeout 120 scripts/lanczos_solve.bin 128 4096 > $OUT/lanczos128.log 2>&1; echo "rc=$?" >> $OUT/lanczos128.log
timeout 120 scripts/lanczos_solve.bin 64 8192 > $OUT/lanczos64.log 2>&1; echo "rc=$?" >> $OUT/lanczos64.log
timeout 120 python scripts/measure_peaks.py 2.0 > $OUT/peaks.json 2> $OUT/peaks.err
timeout 400 python bench.py --workload c5 --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/c5_full.json 2> $OUT/c5_full.err; echo "c5 rc=$?"
timeout 300 python bench.py --workload c3 --scale 0.125 --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/c3_shard.json 2> $OUT/c3_shard.err; echo "c3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm -s 4 -c 2 -o $OUT/c3_spmm -f \
    python bench.py --workload c3 --scale 0.125 --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_c3.log 2>&1; echo "ncu rc=$?"
python - > $OUT/pin.txt 2>&1 <<'PY'
import time, torch
t0=time.time(); a=torch.empty(8*(1<<30), dtype=torch.uint8, pin_memory=True); print("pin 8GiB", time.time()-t0)
d=torch.empty(8*(1<<30), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0=time.time(); d.copy_(a, non_blocking=True); torch.cuda.synchronize(); print("h2d 8GiB", time.time()-t0)
t0=time.time(); a.copy_(d, non_blocking=True); torch.cuda.synchronize(); print("d2h 8GiB", time.time()-t0)
PY
cat $OUT/pair.log $OUT/lanczos128.log $OUT/lanczos64.log $OUT/peaks.json $OUT/host.txt $OUT/pin.txt
tail -c 2500 $OUT/c5_full.json; tail -3 $OUT/c5_full.err; tail -c 2500 $OUT/c3_shard.json; tail -3 $OUT/c3_shard.err; tail -5 $OUT/ncu_c3.log
