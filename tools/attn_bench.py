"""Attention micro-benchmark over build variants (config-2 shape: B=2, N=4096, T=128, H=32, Hkv=8).
usage: python tools/attn_bench.py <lib.so>[:gen] [...]   (run on the GPU box; gen = 1 / 3 selects the kernel generation through
NDIT_ATTN_GEN, which each library reads once at its first call)"""
import ctypes as C
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

B, N, T, H, Hkv, hd = 2, 4096, 128, 32, 8, 72
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * N, (H + 2 * Hkv) * hd, device="cuda", generator=g).to(torch.bfloat16)
kvy = torch.randn(B * T, 2 * Hkv * hd, device="cuda", generator=g).to(torch.bfloat16)
ymask = torch.zeros(B, T, dtype=torch.uint8, device="cuda")
ymask[0, :] = 1
ymask[1, :8] = 1
gate = torch.tanh(0.5 * torch.randn(H, device="cuda", generator=g)).to(torch.bfloat16).float()
out = torch.empty(B * N, H * hd, device="cuda", dtype=torch.bfloat16)
ss, sc = math.sqrt(math.log(N, 4096) / hd), 1 / math.sqrt(hd)
flops = 4 * B * N * N * H * hd + 4 * B * N * T * H * hd
ref = None
for spec in sys.argv[1:]:
    path, _, gen = spec.partition(":")
    os.environ["NDIT_ATTN_GEN"] = gen or "0"
    lib = C.CDLL(os.path.abspath(path))
    ms = C.c_float(0)
    p = lambda t: C.c_void_p(t.data_ptr())
    f = lib.ndit_op_attention_bench
    f.argtypes = [C.c_void_p] * 5 + [C.c_int32] * 5 + [C.c_float, C.c_float, C.c_int32, C.POINTER(C.c_float), C.c_void_p]
    rc = f(p(qkv), p(kvy), p(ymask), p(gate), p(out), B, N, T, H, Hkv, ss, sc, 20, C.byref(ms), None)
    torch.cuda.synchronize()
    if ref is None:
        ref = out.clone()
    diff = (out.float() - ref.float()).abs().max().item()
    print(f"{os.path.basename(path) + (' gen' + gen if gen else ''):40s} rc={rc} {ms.value * 1e3:8.1f} us  {flops / ms.value / 1e9:7.1f} TFLOP/s  maxdiff_vs_first={diff:.4f}", flush=True)
    if hasattr(lib, "ndit_debug_attn_timing"):
        buf = (C.c_longlong * (2 * 64 * 8))()
        lib.ndit_debug_attn_timing(buf)
        t = torch.tensor(list(buf), dtype=torch.float64).view(2, 64, 8)
        names = ["wait_S", "ldS+free", "max+rescale", "wait_Pfree", "exp+stP", "fence+arrive"]
        for x in range(2):
            d = (t[x, 4:30, 1:7] - t[x, 4:30, 0:6])
            per_block = (t[x, 5:31, 0] - t[x, 4:30, 0]).mean().item()
            print(f"   tile {x}: cycles/block {per_block:7.0f} | " + "  ".join(f"{n} {v:6.0f}" for n, v in zip(names, d.mean(0).tolist())), flush=True)
        off = (t[1, 4:30, 4] - t[0, 4:30, 4])
        print(f"   exp-phase start offset tile1-tile0 (same block index): mean {off.mean().item():7.0f}  min {off.min().item():7.0f}  max {off.max().item():7.0f}", flush=True)
    if hasattr(lib, "ndit_debug_attn_hr_timing") and gen == "3":
        buf = (C.c_longlong * (3 * 64 * 8))()
        lib.ndit_debug_attn_hr_timing(buf)
        t = torch.tensor(list(buf), dtype=torch.float64).view(3, 64, 8)
        names = ["wait_S", "ldS+free", "max+post", "nonspec+token", "exps", "xchg+wait_PV+stP", "stwait+arrive"]
        for x in range(2):
            d = (t[x, 4:30, 1:8] - t[x, 4:30, 0:7])
            per_block = (t[x, 5:31, 0] - t[x, 4:30, 0]).mean().item()
            print(f"   tile {x}: cycles/block {per_block:7.0f} | " + "  ".join(f"{n} {v:6.0f}" for n, v in zip(names, d.mean(0).tolist())), flush=True)
        off = (t[1, 4:30, 4] - t[0, 4:30, 4])
        print(f"   exp-phase start offset tile1-tile0 (same block index): mean {off.mean().item():7.0f}  min {off.min().item():7.0f}  max {off.max().item():7.0f}", flush=True)
        h1 = (t[1, 4:30, 4] - t[0, 4:30, 5]).mean().item()
        h2 = (t[0, 5:31, 4] - t[1, 4:30, 5]).mean().item()
        print(f"   token handover (exps issued by one tile -> other tile past its token wait): A->B {h1:6.0f}  B->A {h2:6.0f}", flush=True)
        for x in range(2):
            wake = (t[2, 4:30, 2 * x] - t[x, 4:30, 7]).mean().item()       # softmax quadrant-0 arrive -> MMA warp past p_full wait
            iss = (t[2, 4:30, 2 * x + 1] - t[2, 4:30, 2 * x]).mean().item()
            print(f"   MMA warp tile {x}: p_full arrive(q0) -> issuing {wake:6.0f}   issue of P V {iss:6.0f}", flush=True)
