"""CUDA-event timing of the tcgen05 GEMMs at the C4 training shapes (TFLOP/s per shape)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ss_asr_b200 import _lib
lib = _lib.load()
dev = 'cuda'
st = _lib.stream()

def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

print('NN: C[M,N] = A[M,K] B[N,K]^T')
for (M, N, K) in [(131072, 2048, 80), (65536, 2048, 1024), (65536, 1024, 2048), (32768, 2048, 1024), (16384, 2048, 1024),
                  (256, 1024, 768), (256, 1024, 512), (256, 512, 1024), (10496, 128, 512)]:
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    B = torch.randn(N, K, device=dev).to(torch.bfloat16)
    C = torch.empty(M, N, device=dev)
    bias = torch.randn(N, device=dev)
    f = lambda: _lib.check(lib.ssasr_gemm_bf16_tc(M, N, K, A.data_ptr(), K, 0, B.data_ptr(), K, 0, C.data_ptr(), N, bias.data_ptr(), 0, st), 'gemm')
    ms = timeit(f)
    ref = A.float() @ B.float().t() + bias
    err = float((C - ref).abs().max() / ref.abs().max())
    print(f'  M={M:7d} N={N:5d} K={K:5d}: {ms * 1e3:8.1f} us  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s  out {M * N * 4 / ms / 1e6:7.1f} GB/s  relerr {err:.1e}')
    del A, B, C, ref
print('TN: C[M,N] = A[K,M]^T B[K,N]')
for (M, N, K) in [(2048, 1024, 65536), (2048, 80, 131072), (1024, 256, 131072), (1024, 256, 16384), (2048, 1024, 16384)]:
    A = torch.randn(K, M, device=dev).to(torch.bfloat16)
    B = torch.randn(K, N, device=dev).to(torch.bfloat16)
    C = torch.empty(M, N, device=dev)
    f = lambda: _lib.check(lib.ssasr_gemm_bf16_tc_tn(M, N, K, A.data_ptr(), M, 0, B.data_ptr(), N, 0, C.data_ptr(), N, 0, st), 'gemm_tn')
    ms = timeit(f)
    ref = A.float().t() @ B.float()
    err = float((C - ref).abs().max() / ref.abs().max())
    print(f'  M={M:7d} N={N:5d} K={K:6d}: {ms * 1e3:8.1f} us  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s  relerr {err:.1e}')
    del A, B, C, ref
