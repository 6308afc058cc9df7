"""cuBLAS DGEMM throughput on this GPU (the 'library' FP64 roof; SURVEY 7 asks for it)."""
import json, torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2): torch.matmul(a, b)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps({"dgemm_8192_tflops": 2 * n ** 3 / best * 1e-9, "ms": best}))
