"""Device time of one forward_with_cfg for the BASELINE configs other than the bench workload (config 2):
config 1 (class-conditional 600M, 256^2), config 3 (Lumina-Next 2B GQA, 2048^2), config 4 (Flag-DiT 5B, 1024^2),
config 5 (MoE 600M "both", 512^2).  Random weights of the reference architecture created on the device; CUDA-event
timing after warm-up.  Prints one JSON line per config.  Usage: python tools/config_timing.py [1 3 4 5]"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lumina_t2x_b200 import models  # noqa: E402
from lumina_t2x_b200.models import lumina_t2i, moe  # noqa: E402


def randomize(m):
    with torch.no_grad():
        for k, p in m.named_parameters():
            if p.dim() == 2:
                p.normal_(std=(0.5 if "adaLN" in k else 1.0) / math.sqrt(p.shape[1]))
            elif "norm" in k and k.endswith("weight"):
                p.fill_(1.0)
            else:
                p.normal_(std=0.02)
    return m.eval().to("cuda", dtype=torch.bfloat16)


def timed(fn, warm=3, iters=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def flops(B, N, D, H, Hkv, hd, F, L, T=0, ffn_mult=1.0):
    gemm = 2 * B * N * (D * (H + 2 * Hkv) * hd + D * D + ffn_mult * 3 * D * F) * L
    attn = 4 * B * N * N * H * hd * L + 4 * B * N * T * H * hd * L
    return gemm, attn


def main():
    which = [int(a) for a in sys.argv[1:]] or [1, 3, 4, 5]
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    for c in which:
        with dev:
            if c == 1:
                m = randomize(models.DiT_Llama_600M_patch2(input_size=32, num_classes=1000, qk_norm=True))
                z = torch.randn(2, 4, 32, 32, device="cuda", generator=g).bfloat16()
                y = torch.tensor([207, 1000], device="cuda")
                fn = lambda: m.forward_with_cfg(z, torch.zeros(2, device="cuda"), y, 4.0)  # noqa: E731
                fl = flops(2, 256, 1536, 32, 32, 48, 4096, 16)
                name = "config1: DiT_Llama_600M_patch2 256x256"
            elif c == 3:
                m = randomize(models.NextDiT_2B_GQA_patch2(qk_norm=True, cap_feat_dim=2048, max_tokens=16384))
                z = torch.randn(2, 4, 256, 256, device="cuda", generator=g).bfloat16()
                cap = torch.randn(2, 128, 2048, device="cuda", generator=g).bfloat16()
                mask = torch.ones(2, 128, dtype=torch.int64, device="cuda")
                fn = lambda: m.forward_with_cfg(z, torch.full((2,), 0.5, device="cuda"), cap, mask, 4.0, scale_factor=2.0,  # noqa: E731
                                                scale_watershed=0.3, base_seqlen=4096, proportional_attn=True)
                fl = flops(2, 16384, 2304, 32, 8, 72, 6144, 24, T=128)
                name = "config3: NextDiT_2B_GQA_patch2 2048x2048"
            elif c == 4:
                m = randomize(lumina_t2i.DiT_Llama_5B_patch2(qk_norm=True, cap_feat_dim=4096, max_tokens=4160))
                z = torch.randn(2, 4, 128, 128, device="cuda", generator=g).bfloat16()
                cap = torch.randn(2, 128, 4096, device="cuda", generator=g).bfloat16()
                mask = torch.ones(2, 128, dtype=torch.int64, device="cuda")
                fn = lambda: m.forward_with_cfg(z, torch.full((2,), 0.5, device="cuda"), cap, mask, 4.0, proportional_attn=True,  # noqa: E731
                                                base_seqlen=64 * 64 + 64 * 2, ntk_factor=1.0)
                fl = flops(2, 4160, 3072, 32, 32, 96, 8192, 32, T=128)
                name = "config4: Flag-DiT DiT_Llama_5B_patch2 1024x1024"
            elif c == 6:
                # compositional generation at the config-2 shape: 4 region captions (2 x 2 split) + the negative one, T = 128
                from lumina_t2x_b200.models import compositional
                m = randomize(compositional.NextDiT_2B_GQA_patch2(qk_norm=True, cap_feat_dim=2048, max_tokens=4096, max_cap_len=128))
                z = torch.randn(2, 4, 128, 128, device="cuda", generator=g).bfloat16()
                cap = torch.randn(5, 128, 2048, device="cuda", generator=g).bfloat16()
                mask = torch.ones(5, 128, dtype=torch.int64, device="cuda")
                gcap = torch.randn(1, 128, 2048, device="cuda", generator=g).bfloat16()
                gmask = torch.ones(1, 128, dtype=torch.int64, device="cuda")
                fn = lambda: m.forward_with_cfg(z, torch.full((2,), 0.5, device="cuda"), cap, mask, 4.0, base_seqlen=4096,  # noqa: E731
                                                proportional_attn=True, global_cap_feats=gcap, global_cap_mask=gmask, h_split_num=2, w_split_num=2)
                fl = flops(2, 4096, 2304, 32, 8, 72, 6144, 24, T=128)     # algorithmic: one caption per token
                name = "compositional: NextDiT_2B_GQA_patch2 1024x1024, 2x2 regions (5 caption rows)"
            else:
                m = randomize(moe.DiT_Llama_600M_patch2_Both(input_size=64, num_classes=1000, qk_norm=True))
                z = torch.randn(2, 4, 64, 64, device="cuda", generator=g).bfloat16()
                y = torch.tensor([207, 1000], device="cuda")
                fn = lambda: m.forward_with_cfg(z, torch.full((2,), 0.5, device="cuda"), y, 4.0)  # noqa: E731
                fl = flops(2, 1024, 1536, 32, 32, 48, 4096, 16, ffn_mult=4.0)   # top-2 of 4, twice (algorithmic)
                name = "config5: MoE DiT_Llama_600M_patch2_Both 512x512"
        out = fn()
        assert torch.isfinite(out.float()).all()
        n0 = m.launch_count()
        ms = timed(fn)
        per = (m.launch_count() - n0) // 8
        rec = {"config": name, "ms_per_forward": round(ms, 3), "kernel_launches_per_forward": per,
               "algorithmic_tflop": round((fl[0] + fl[1]) / 1e12, 3),
               "tflops": round((fl[0] + fl[1]) / 1e9 / ms, 1), "params_B": round(m.parameter_count() / 1e9, 3)}
        if c == 6:
            # the same engine with plain captions (one per row): what the region-masked caption segment costs on top
            from lumina_t2x_b200.models.nextdit import NextDiT as _Base
            pfn = lambda: _Base.forward_with_cfg(m, z, torch.full((2,), 0.5, device="cuda"), cap[:2], mask[:2], 4.0, base_seqlen=4096,  # noqa: E731
                                                 proportional_attn=True)
            pfn()
            rec["ms_per_forward_plain_captions_same_engine"] = round(timed(pfn), 3)
        if c in (1, 5):
            # whole 30-point Euler solve through transport.Sampler (29 model calls): direct launches vs the captured CUDA graph
            from lumina_t2x_b200 import transport
            sfn = transport.Sampler(transport.create_transport("Linear", "velocity", None, None, None)).sample_ode(
                sampling_method="euler", num_steps=30, atol=1e-6, rtol=1e-3, reverse=False, time_shifting_factor=1.0)
            outs = {}
            for gopt in (0, 1):
                m.set_option("graph", gopt)
                solve = lambda: sfn(z, m.forward_with_cfg, y=y, cfg_scale=4.0)[-1]  # noqa: E731
                outs[gopt] = solve()
                r0 = m.graph_replay_count()
                rec["ms_per_model_call_in_solve_graph%d" % gopt] = round(timed(solve, warm=2, iters=5) / 29, 3)
                rec["graph_replays_graph%d" % gopt] = m.graph_replay_count() - r0
            rec["graph_equals_direct"] = bool(torch.equal(outs[0], outs[1]))
        print(json.dumps(rec), flush=True)
        del m
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
