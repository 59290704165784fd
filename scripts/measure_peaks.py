"""cuBLAS GEMM peaks on this GPU (TF32, fp32 without TF32, fp64) -- the tensor rooflines MEASURED_PEAKS.json does not carry.
Prints one JSON object; run on the GPU box (bench.py measures the same inline)."""
import json
import sys
import time

import torch


def gemm_tflops(dtype, n, tf32, reps=6, sustained_s=0.0):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    c = torch.empty(n, n, device="cuda", dtype=dtype)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out = {"burst": 2.0 * n ** 3 / best / 1e9}
    if sustained_s > 0:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0, cnt = time.perf_counter(), 0
        e0.record()
        while time.perf_counter() - t0 < sustained_s:
            for _ in range(10):
                torch.matmul(a, b, out=c)
            cnt += 10
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        out["sustained"] = 2.0 * n ** 3 * cnt / e0.elapsed_time(e1) / 1e9
    torch.backends.cuda.matmul.allow_tf32 = False
    return out


if __name__ == "__main__":
    s = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    res = {"tf32_tflops": gemm_tflops(torch.float32, 8192, True, sustained_s=s),
           "fp32_tflops": gemm_tflops(torch.float32, 8192, False),
           "fp64_tflops": gemm_tflops(torch.float64, 4096, False, sustained_s=s),
           "bf16_tflops": gemm_tflops(torch.bfloat16, 8192, False, sustained_s=s)}
    print(json.dumps(res))
