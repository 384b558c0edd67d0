"""Measure cuBLAS DGEMM / SGEMM (with and without TF32) throughput on this GPU, the same way
MEASURED_PEAKS.json was made for bf16: 8192^3, best of 10 (burst) and back-to-back for 4 s (sustained)."""
import json, time, torch

def bench(dtype, n=8192, tf32=False):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    c = torch.empty_like(a)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(10):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    flops = 2.0 * n ** 3
    burst = flops / best * 1e-9
    # sustained
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    reps = max(3, int(4000.0 / best))
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record(); e1.synchronize()
    sus = flops * reps / e0.elapsed_time(e1) * 1e-9
    return round(burst, 2), round(sus, 2)

out = {"gpu": torch.cuda.get_device_name(0)}
out["dgemm_tflops"], out["dgemm_tflops_sustained"] = bench(torch.float64)
out["sgemm_tflops"], out["sgemm_tflops_sustained"] = bench(torch.float32, tf32=False)
out["tf32gemm_tflops"], out["tf32gemm_tflops_sustained"] = bench(torch.float32, tf32=True)
print(json.dumps(out))
