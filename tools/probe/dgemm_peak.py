"""Measure the FP64 GEMM peak (cuBLAS via torch.matmul) on this B200: burst and sustained."""
import json, time, torch
def main():
    dev = torch.device("cuda:0")
    res = {}
    for n in (4096, 8192):
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        b = torch.randn(n, n, dtype=torch.float64, device=dev)
        for _ in range(2): a @ b
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[f"dgemm_{n}_burst_tflops"] = 2 * n**3 / best * 1e-9
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        reps = 6 if n == 8192 else 40
        e0.record()
        for _ in range(reps): c = a @ b
        e1.record(); torch.cuda.synchronize()
        res[f"dgemm_{n}_sustained_tflops"] = 2 * n**3 * reps / e0.elapsed_time(e1) * 1e-9
    print(json.dumps(res))
if __name__ == "__main__":
    main()
