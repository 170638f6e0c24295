"""Timing of the VAE-decode end (nvae_decode) at the sdxl-vae architecture, next to the stock path the reference would run
(the oracle's torch restatement on the GPU under autocast(bf16): cuDNN convolutions, ATen group_norm, SDPA) and a full-size parity
check against the same restatement in fp32 on the GPU (TF32 off).   python tools/vae_bench.py [latent_side] [iters] [--no-stock]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vae_oracle as VO  # noqa: E402
from lumina_t2x_b200.vae import AutoencoderKL  # noqa: E402


def stock_decode(cfg, W, z):
    """The reference's call: fp32 module under autocast(bf16) (sample.py:173, :238), restated functionally with torch ops."""
    import torch.nn.functional as F
    G = cfg.norm_num_groups

    def conv(x, pre, pad):
        return F.conv2d(x, W[pre + ".weight"], W[pre + ".bias"], padding=pad)

    def gn(x, pre):
        return F.group_norm(x, G, W[pre + ".weight"], W[pre + ".bias"], eps=1e-6)

    def res(x, pre):
        h = conv(F.silu(gn(x, pre + ".norm1")), pre + ".conv1", 1)
        h = conv(F.silu(gn(h, pre + ".norm2")), pre + ".conv2", 1)
        if pre + ".conv_shortcut.weight" in W:
            x = conv(x, pre + ".conv_shortcut", 0)
        return x + h

    with torch.autocast("cuda", dtype=torch.bfloat16):
        x = conv(conv(z, "post_quant_conv", 0), "decoder.conv_in", 1)
        x = res(x, "decoder.mid_block.resnets.0")
        a = "decoder.mid_block.attentions.0"
        B, C, H, Wd = x.shape
        u = gn(x, a + ".group_norm").view(B, C, H * Wd).transpose(1, 2)
        q, k, v = (F.linear(u, W[f"{a}.{n}.weight"], W[f"{a}.{n}.bias"]) for n in ("to_q", "to_k", "to_v"))
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        o = F.linear(o, W[a + ".to_out.0.weight"], W[a + ".to_out.0.bias"])
        x = o.transpose(1, 2).reshape(B, C, H, Wd) + x
        x = res(x, "decoder.mid_block.resnets.1")
        for i in range(4):
            for j in range(cfg.layers_per_block + 1):
                x = res(x, f"decoder.up_blocks.{i}.resnets.{j}")
            if i < 3:
                x = conv(F.interpolate(x, scale_factor=2.0, mode="nearest"), f"decoder.up_blocks.{i}.upsamplers.0.conv", 1)
        return conv(F.silu(gn(x, "decoder.conv_norm_out")), "decoder.conv_out", 1)


def main():
    side = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 128
    iters = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 10
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = VO.VaeCfg()
    W = VO.synthetic_weights(cfg, seed=0)
    m = AutoencoderKL()
    m.load_state_dict(W, strict=True)
    m = m.cuda()
    z = torch.randn(1, 4, side, side, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16).cuda()
    out = m.decode(z).sample
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        m.decode(z)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    flops = 10.3e12 * (side / 128) ** 2
    rec = {"workload": f"AutoencoderKL.decode sdxl-vae architecture, 1 x 4 x {side} x {side} latent -> {8 * side} x {8 * side}", "iters": iters,
           "engine_ms_median": ts[len(ts) // 2], "engine_ms_min": ts[0], "engine_tflops": flops / (ts[len(ts) // 2] * 1e-3) / 1e12}
    if "--no-stock" not in sys.argv:
        Wc = {k: v.cuda() for k, v in W.items()}
        with torch.no_grad():
            ref = stock_decode(cfg, Wc, z.float())
            torch.cuda.synchronize()
            st = []
            for i in range(max(3, iters // 2)):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                stock_decode(cfg, Wc, z.float())
                b.record()
                torch.cuda.synchronize()
                st.append(a.elapsed_time(b))
            st.sort()
            ref32 = VO.decode(cfg, Wc, z.float(), "fp32")
        rel = lambda x, y: ((x.float() - y.float()).abs().max() / y.float().abs().max()).item()   # noqa: E731
        rms = lambda x, y: ((x.float() - y.float()).pow(2).mean().sqrt() / y.float().pow(2).mean().sqrt()).item()   # noqa: E731
        rec.update({"stock_autocast_ms_median": st[len(st) // 2], "speedup_vs_stock": st[len(st) // 2] / ts[len(ts) // 2],
                    "rel_linf_engine_vs_fp32": rel(out, ref32), "rel_linf_stock_vs_fp32": rel(ref, ref32),
                    "rel_rms_engine_vs_fp32": rms(out, ref32), "rel_rms_stock_vs_fp32": rms(ref, ref32)})
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
