"""Peer numbers from cuBLAS (via torch.matmul, float64) — NOT on the product path.
(1) Dgemm 8192^3 burst / sustained = the library FP64 peak on this box;
(2) X'X at the C2 shape (n=1e6, p=500, column-major X) = what a library-only Gram build costs.
Prints one JSON object."""
import json, time, torch

def tm(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

out = {}
N = 8192
a = torch.rand(N, N, dtype=torch.float64, device="cuda"); b = torch.rand(N, N, dtype=torch.float64, device="cuda")
ms = tm(lambda: torch.matmul(a, b), reps=10)
out["dgemm_8192_burst_tflops"] = 2 * N**3 / ms * 1e-9
t0 = time.time(); k = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 3.0:
    torch.matmul(a, b); k += 1
    if k % 4 == 0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
out["dgemm_8192_sustained_tflops"] = 2 * N**3 * k / e0.elapsed_time(e1) * 1e-9
del a, b
n, p = 1_000_000, 500
Xt = torch.rand(p, n, dtype=torch.float64, device="cuda")   # row-major p×n == column-major n×p
ms = tm(lambda: torch.matmul(Xt, Xt.T), reps=5)
out["cublas_XtX_c2_ms"] = ms
out["cublas_XtX_c2_hw_tflops"] = 2 * n * p * p / ms * 1e-9
out["cublas_XtX_c2_algo_tflops"] = n * p * (p + 1) / ms * 1e-9
print(json.dumps(out))
