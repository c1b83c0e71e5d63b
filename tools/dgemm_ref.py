"""Developer tool: what does cuBLAS DGEMM reach on the c3 GEMM shapes (vs 8192^3)?"""
import torch
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for (M, N, K) in [(8192, 8192, 8192), (97556, 500, 544), (97556, 544, 500), (24389, 500, 544), (97556, 2000, 544), (97556, 512, 512)]:
    A = torch.randn(M, K, dtype=torch.float64, device="cuda")
    B = torch.randn(N, K, dtype=torch.float64, device="cuda")
    Bn = B.t().contiguous()
    ms = t(lambda: torch.matmul(A, B.t()))
    ms2 = t(lambda: torch.matmul(A, Bn))
    print(f"M={M} N={N} K={K}: TN {2*M*N*K/ms/1e9:.2f} TFLOP/s ({ms:.3f} ms)   NN {2*M*N*K/ms2/1e9:.2f} TFLOP/s")
