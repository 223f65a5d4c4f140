"""Measure the FP64 roofline denominators that MEASURED_PEAKS.json lacks:
cuBLAS DGEMM / ZGEMM 8192^3, burst (best of 10) and sustained (back to back for ~4 s).
Writes gpurun_out/fp64_peaks.json.  torch is only the cuBLAS caller here."""
import json, time, sys, os
import torch

def bench(dtype, n, flop_per_mac):
    a = torch.randn(n, n, dtype=dtype, device="cuda")
    b = torch.randn(n, n, dtype=dtype, device="cuda")
    c = torch.empty_like(a)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    flops = flop_per_mac * n ** 3
    burst = flops / best * 1e-9
    # sustained
    reps = max(4, int(4000.0 / best))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record(); e1.synchronize()
    sus = flops * reps / e0.elapsed_time(e1) * 1e-9
    return burst, sus

if __name__ == "__main__":
    out = {"gpu": torch.cuda.get_device_name(0)}
    d = bench(torch.float64, 8192, 2)
    out["dgemm_tflops"], out["dgemm_tflops_sustained"] = d
    z = bench(torch.complex128, 4096, 8)
    out["zgemm_tflops"], out["zgemm_tflops_sustained"] = z
    out["how"] = "torch.matmul f64 8192^3 (2N^3) and c128 4096^3 (8N^3): best of 10 (burst), back-to-back ~4 s (sustained)"
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/fp64_peaks.json", "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))
