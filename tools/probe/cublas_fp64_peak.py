"""cuBLAS DGEMM/ZGEMM on device-resident operands: the FP64 roofline denominator (BASELINE.md B-cuBLAS).
torch.matmul(float64) dispatches to cublasDgemm. Not on the product path; run once per pod."""
import json, sys, time, torch
dev = torch.device("cuda:0")
def bench(n, dtype, reps=5):
    a = torch.rand(n, n, device=dev, dtype=dtype); b = torch.rand(n, n, device=dev, dtype=dtype)
    c = torch.empty(n, n, device=dev, dtype=dtype)
    torch.matmul(a, b, out=c); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    mult = 8 if dtype.is_complex else 2
    return {"probe": "cublas", "dtype": str(dtype), "n": n, "ms": best, "tflops": mult * n**3 / best * 1e-9}
sizes = [int(x) for x in sys.argv[1:]] or [4096, 8192, 16384]
for n in sizes:
    print(json.dumps(bench(n, torch.float64)), flush=True)
print(json.dumps(bench(8192, torch.complex128, 3)), flush=True)
print(json.dumps(bench(8192, torch.float32, 3)), flush=True)
