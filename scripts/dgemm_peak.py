"""cuBLAS DGEMM peak through torch (fp64 roofline denominator; MEASURED_PEAKS.json has none)."""
import json
import torch

def measure(n=8192, reps=6):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / best * 1e-9

if __name__ == "__main__":
    print(json.dumps({"dgemm_tflops_8192": measure(8192), "dgemm_tflops_4096": measure(4096)}))
