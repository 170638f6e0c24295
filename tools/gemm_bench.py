"""GEMM micro-benchmark on the four block shapes of config 2 (M = 8192), single-CTA vs CTA-pair kernel.
usage: python tools/gemm_bench.py   (on the GPU box)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lumina_t2x_b200 import _lib

lib = _lib.load()
M = 8192
shapes = [("qkv", 3456, 2304, 0), ("wo", 2304, 2304, 0), ("w13_swiglu", 12288, 2304, 1), ("w2", 2304, 6144, 0)]
g = torch.Generator(device="cuda").manual_seed(0)
for name, N, K, sw in shapes:
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    Cc = torch.empty(M, N // 2 if sw else N, device="cuda", dtype=torch.bfloat16)
    ref = None
    for pair in (0, 1):
        ms = C.c_float(0)
        rc = lib.ndit_op_gemm_bench(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(Cc.data_ptr()), M, N, K, sw, pair, 20,
                                    C.byref(ms), None)
        torch.cuda.synchronize()
        if ref is None:
            ref = Cc.clone()
        d = (Cc.float() - ref.float()).abs().max().item()
        print(f"{name:12s} N={N:6d} K={K:5d} allow_pair={pair} ran_pair={rc} {ms.value * 1e3:8.1f} us {2.0 * M * N * K / ms.value / 1e9:8.1f} TFLOP/s maxdiff_vs_single={d:.4f}",
              flush=True)
