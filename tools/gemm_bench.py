#!/usr/bin/env python
"""Throughput of the tcgen05 GEMM (test hook moegan_umma_gemm, fp32 out) on large and on MoE-sized shapes, with and
without cta_group::2 (MOEGAN_DISABLE_CTA2=1 in the environment), next to torch.matmul (cuBLAS) on the same shapes."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("moe-gan_cpsc541_b200")
C = pkg._cabi; lib = C.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n * 1e3  # us

print("cta2 disabled" if os.environ.get("MOEGAN_DISABLE_CTA2") == "1" else "cta2 enabled")
for kind, M, N, K in [(0, 8192, 8192, 8192), (0, 16384, 4096, 4096), (2, 4096, 4096, 16384), (0, 8192, 2048, 512), (0, 8192, 512, 2048),
                      (1, 8192, 2048, 512), (1, 8192, 512, 2048), (2, 512, 2048, 8192), (2, 2048, 512, 8192)]:
    if kind == 0: A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16(); ref = lambda: A @ B.T
    elif kind == 1: A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(K, N, device="cuda").bfloat16(); ref = lambda: A @ B
    else: A = torch.randn(K, M, device="cuda").bfloat16(); B = torch.randn(K, N, device="cuda").bfloat16(); ref = lambda: A.T @ B
    Cout = torch.empty(M, N, device="cuda")
    f = lambda: C.check(lib.moegan_umma_gemm(kind, 0, M, N, K, C.ptr(A), C.ptr(B), C.ptr(Cout), st), "gemm")
    us = timeit(f); us_ref = timeit(ref)
    err = float((Cout - ref().float()).abs().max() / ref().float().abs().max())
    fl = 2.0 * M * N * K
    print(f"kind {kind} M={M:6d} N={N:5d} K={K:6d}: ours {us:8.1f} us {fl / us / 1e6:7.1f} TF/s | cuBLAS {us_ref:8.1f} us {fl / us_ref / 1e6:7.1f} TF/s | err {err:.1e}")
