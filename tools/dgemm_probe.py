"""cuBLAS DGEMM throughput via torch (library peak used as the fp64 'tensor' roofline denominator)."""
import json, torch
for n in (2048, 4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2): c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(json.dumps({"probe": "cublas_dgemm", "n": n, "tflops": round(2 * n ** 3 / best / 1e9, 2), "ms": round(best, 3)}))
