"""Head-to-head: this repo's fused attention kernel vs flash-attn 2 (flash_attn_varlen_func, the kernel the reference calls at
lumina_next_t2i/models/model.py:392-403) on the config-2 and config-3 shapes, same box, CUDA events.
usage: python tools/attn_vs_fa2.py   (on the GPU box)"""
import ctypes as C
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lumina_t2x_b200 import _lib

lib = _lib.load()
H, Hkv, hd, T = 32, 8, 72, 128
for B, N in ((2, 4096), (2, 16384)):
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn(B * N, (H + 2 * Hkv) * hd, device="cuda", generator=g).to(torch.bfloat16)
    kvy = torch.randn(B * T, 2 * Hkv * hd, device="cuda", generator=g).to(torch.bfloat16)
    ymask = torch.zeros(B, T, dtype=torch.uint8, device="cuda")
    ymask[0, :] = 1
    ymask[1, :8] = 1
    gate = torch.zeros(H, device="cuda")          # gate 0: the cross segment contributes nothing -> outputs comparable with FA2
    out = torch.empty(B * N, H * hd, device="cuda", dtype=torch.bfloat16)
    ss, sc = math.sqrt(math.log(N, 4096) / hd), 1 / math.sqrt(hd)
    flops_self = 4 * B * N * N * H * hd
    flops_all = flops_self + 4 * B * N * T * H * hd
    p = lambda t: C.c_void_p(t.data_ptr())
    ms = C.c_float(0)
    iters = 20 if N <= 4096 else 5
    rc = lib.ndit_op_attention_bench(p(qkv), p(kvy), p(ymask), p(gate), p(out), B, N, T, H, Hkv, ss, sc, iters, C.byref(ms), None)
    torch.cuda.synchronize()
    assert rc == 0
    print(f"B={B} N={N}: engine attention_fused (self + T=128 cross): {ms.value * 1e3:9.1f} us  {flops_all / ms.value / 1e9:7.1f} TFLOP/s", flush=True)
    try:
        from flash_attn import flash_attn_varlen_func
        q = qkv[:, : H * hd].reshape(B * N, H, hd).contiguous()
        k = qkv[:, H * hd: (H + Hkv) * hd].reshape(B * N, Hkv, hd).contiguous()
        v = qkv[:, (H + Hkv) * hd:].reshape(B * N, Hkv, hd).contiguous()
        cu = torch.arange(0, (B + 1) * N, N, device="cuda", dtype=torch.int32)
        f = lambda: flash_attn_varlen_func(q, k, v, cu_seqlens_q=cu, cu_seqlens_k=cu, max_seqlen_q=N, max_seqlen_k=N, dropout_p=0.0,
                                           causal=False, softmax_scale=ss)
        o = f()
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            f()
        e1.record()
        torch.cuda.synchronize()
        fms = e0.elapsed_time(e1) / iters
        d = (o.reshape(B * N, H * hd).float() - out.float()).abs().max().item()
        print(f"B={B} N={N}: flash_attn_varlen_func 2.x (self only):          {fms * 1e3:9.1f} us  {flops_self / fms / 1e9:7.1f} TFLOP/s  "
              f"max|engine - fa2| = {d:.4f} (|out| max {out.float().abs().max().item():.3f})", flush=True)
    except Exception as ex:
        print(f"flash-attn unavailable: {type(ex).__name__}: {ex}", flush=True)
    # torch SDPA (what PyTorch picks on this GPU: cuDNN / flash backends)
    qh = qkv[:, : H * hd].reshape(B, N, H, hd).transpose(1, 2)
    kh = qkv[:, H * hd: (H + Hkv) * hd].reshape(B, N, Hkv, hd).transpose(1, 2)
    vh = qkv[:, (H + Hkv) * hd:].reshape(B, N, Hkv, hd).transpose(1, 2)
    f2 = lambda: torch.nn.functional.scaled_dot_product_attention(qh, kh, vh, scale=ss, enable_gqa=True)
    try:
        for _ in range(3):
            f2()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            f2()
        e1.record()
        torch.cuda.synchronize()
        sms = e0.elapsed_time(e1) / iters
        print(f"B={B} N={N}: torch SDPA (enable_gqa, self only):              {sms * 1e3:9.1f} us  {flops_self / sms / 1e9:7.1f} TFLOP/s", flush=True)
    except Exception as ex:
        print(f"SDPA failed: {type(ex).__name__}: {ex}", flush=True)
