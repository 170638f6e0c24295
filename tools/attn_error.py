"""Achieved error of the fused attention kernels against an fp32 PyTorch evaluation of the same op (the reference's rounding
points: bf16 self output, bf16 gated cross output, bf16 sum), per test case of tests/test_ops_gpu.py::test_attention and at the
config-2 shape.  Prints max|err| / max|ref| per kernel generation; the test tolerance is set from these numbers."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from lumina_t2x_b200 import _lib
from test_ops_gpu import ATTN_CASES, _attn_ref, ptr

lib = _lib.load()
cases = list(ATTN_CASES) + [(2, 4096, 128, 32, 8, [128, 8])]
for use_ref, name in ((1, "refkernel"), (3, "gen1"), (2, "gen3")):
    worst = 0.0
    for (B, N, T, H, Hkv, valid) in cases:
        hd = 72
        g = torch.Generator(device="cuda").manual_seed(N + T)
        qkv = torch.randn(B * N, (H + 2 * Hkv) * hd, device="cuda", generator=g).to(torch.bfloat16)
        kvy = torch.randn(B * T, 2 * Hkv * hd, device="cuda", generator=g).to(torch.bfloat16)
        ymask = torch.zeros(B, T, dtype=torch.uint8, device="cuda")
        for b, n in enumerate(valid):
            ymask[b, :n] = 1
        gate_tanh = torch.tanh(0.5 * torch.randn(H, device="cuda", generator=g)).to(torch.bfloat16).float()
        ss, sc = math.sqrt(math.log(N, 64) / hd), 1 / math.sqrt(hd)
        out = torch.empty(B * N, H * hd, device="cuda", dtype=torch.bfloat16)
        if N > 2048 and use_ref == 1:
            continue
        rc = lib.ndit_op_attention(ptr(qkv), ptr(kvy), ptr(ymask), ptr(gate_tanh), ptr(out), B, N, T, H, Hkv, ss, sc, use_ref, None)
        torch.cuda.synchronize()
        assert rc == 0
        ref = _attn_ref(qkv, kvy, ymask, gate_tanh, B, N, T, H, Hkv, ss, sc)
        e = ((out.float() - ref.float()).abs().max() / ref.float().abs().max()).item()
        worst = max(worst, e)
        print(f"{name:10s} B={B} N={N:5d} T={T:4d} H={H:3d} Hkv={Hkv}: max|err|/max|ref| = {e:.3e}", flush=True)
    print(f"{name:10s} worst = {worst:.3e}", flush=True)
