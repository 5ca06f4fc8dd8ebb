# cuBLAS DGEMM rate through torch (denominator only; never on the product path).
import torch, time, json
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(3): c = a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): c = a @ b
e1.record(); torch.cuda.synchronize()
sus = e0.elapsed_time(e1) / 20
print(json.dumps({"dgemm_n": n, "burst_tflops": 2*n**3/best*1e-9, "sustained_tflops": 2*n**3/sus*1e-9}))
# cusolver potrf for context
for N in (4096, 16384):
    x = torch.randn(N, N, dtype=torch.float64, device="cuda"); k = x @ x.T + N*torch.eye(N, dtype=torch.float64, device="cuda")
    for _ in range(2): l = torch.linalg.cholesky(k)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); l = torch.linalg.cholesky(k); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"cusolver_potrf_n": N, "ms": ms, "tflops": N**3/3/ms*1e-9}))
