#!/usr/bin/env python
"""Benchmark of the Next-DiT denoising hot path (BASELINE.json metric: latents/sec).

  python bench.py --gpus N --steps K --warmup W            # this repo's engine
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU (oracle port)

A "step" is one complete ODE solve of one latent (= one cond/uncond pair) per GPU on the configuration
BASELINE.json quotes the metric on (configs[1]): Lumina-Next-T2I 2B GQA, 1024x1024 image = latent
[2,4,128,128] = 4096 tokens, T=128 synthetic caption tokens, 30-step Euler (29 forward_with_cfg calls),
CFG=2, bf16, random-init weights, synthetic data.  One JSON line is printed by rank 0.

  value     device-resident throughput: inputs (noise, caption features) already in HBM, CUDA-event timed.
  e2e       same solve through the host-buffer C-ABI call ndit_sample_host: pinned host inputs are copied
            H2D and the final latent D2H inside the timed region (wall clock around synchronous calls).
  roofline  for the dominant kernel (the tcgen05 GEMM): algorithmic FLOPs / CUDA-event time of its launches,
            measured with per-launch events inside one extra solve (engine option "profile").
  cpu_baseline  the oracle (CPU port of the reference algorithm, fp32) on a bounded sample.
  gpu_eager_baseline  the same forward as eager PyTorch bf16 on the same GPU (the reference's stock CUDA path restated;
                the reference itself cannot travel to the GPU box).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

LATENT = 128          # 1024 / 8
T_CAP = 128
NUM_STEPS = 30        # grid points -> 29 model calls (transport/integrators.py:97)
CFG_SCALE = 2.0
WORKLOAD = ("Lumina-Next-T2I 2B GQA (NextDiT_2B_GQA_patch2, qk_norm, cap_feat_dim 2048), 1024x1024 -> latent "
            "[2,4,128,128] (4096 tokens x cond/uncond), T=128 caption tokens, 30-step Euler (29 model calls), CFG=2")


def flops_per_forward(D=2304, L=24, H=32, Hkv=8, hd=72, F=6144, Cc=2048, B=2, N=4096, T=128):
    """Algorithmic FLOPs of one forward_with_cfg (BASELINE.md section 2 formulas)."""
    M = B * N
    gemm = L * 2 * M * (2 * D * D + 2 * D * Hkv * hd + 3 * D * F)
    attn = L * 4 * B * N * N * H * hd
    cross = L * (4 * B * N * T * H * hd + 4 * B * T * Cc * Hkv * hd)
    return gemm, attn, cross


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return dict(bf16=float(d.get("bf16_tflops_sustained") or d["bf16_tflops"]), hbm=float(d["hbm_gbs"]), src="measured (MEASURED_PEAKS.json, sustained)")
        except Exception:
            pass
    return dict(bf16=1400.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum) from the committed
    `ncu --set full` capture of one gemm2_bf16_tn_kernel launch (the wo projection, 8192x2304x2304: 86 MB algorithmic,
    profiles/r01_ncu_gemm2_pair_full.txt -> profiles/ncu_traffic.json via tools/ncu_summarize.py); None if absent."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p))["gemm_first_pair"]
    except Exception:
        return None


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc = gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        try:
            self.proc.terminate()
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            return None
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            pass
    return local_rank


# ------------------------------------------------------------------------------------------ CPU baseline (oracle)
def cpu_sample_once(full_layers=24, sample_layers=2):
    """One bounded sample of the reference algorithm on the host: forward_with_cfg at the full token count
    with 0 and `sample_layers` transformer blocks; per-block time extrapolated to `full_layers` blocks and
    29 model calls.  Returns seconds per latent (extrapolated) and the measured pieces."""
    from oracle import nextdit_oracle as O
    out = {}
    torch.set_grad_enabled(False)
    for nl in (0, sample_layers):
        cfg = O.NextDiTConfig(n_layers=nl)
        W = {k: v.float() for k, v in O.synthetic_weights(cfg, seed=0).items()}
        z, cap, mask = O.synthetic_inputs(cfg, (LATENT, LATENT), T_CAP, 8, seed=1)
        t = torch.full((2,), 0.3)
        t0 = time.perf_counter()
        O.forward_with_cfg(cfg, W, z.float(), t, cap.float(), mask, CFG_SCALE, 1.0, 1.0, 4096, True, precision="fp32")
        out[nl] = time.perf_counter() - t0
    per_block = (out[sample_layers] - out[0]) / sample_layers
    fwd = out[0] + full_layers * per_block
    return (NUM_STEPS - 1) * fwd, out[0], per_block


def run_reference(args, rank):
    if rank != 0:
        return
    cores = torch.get_num_threads()
    for _ in range(min(args.warmup, 1)):
        cpu_sample_once()
    vals = []
    for _ in range(args.steps):
        vals.append(cpu_sample_once()[0])
    sec = statistics.mean(vals)
    sample = ("oracle port (oracle/nextdit_oracle.py, fp32, torch CPU) of the reference forward_with_cfg at the full 2x4096-token "
              "shape with 0 and 2 of the 24 blocks; per-block time x24 + embed/final, x29 model calls")
    line = {"metric": "latents/sec", "value": 1.0 / sec, "unit": "latents/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "note": "reference algorithm on host CPU; extrapolated from a bounded sample"},
            "cpu_baseline": {"value": 1.0 / sec, "unit": "latents/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": 1.0 / sec, "unit": "latents/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def gpu_eager_baseline(m, z_dev, cap_dev, mask_dev, device, engine_out, iters=3):
    """Same-GPU "stock CUDA path" baseline: the reference forward_with_cfg as eager PyTorch bf16
    (cuBLAS Linears, fused SDPA attention, one ATen kernel per op) - oracle/nextdit_oracle.py::forward_with_cfg_eager -
    on the weights of the benchmarked model.  Device time of `iters` model calls, x29 calls per latent."""
    from oracle import nextdit_oracle as O
    cfg = O.config_2b_gqa()
    W = {k: v.detach() for k, v in m.state_dict().items()}
    t = torch.full((2,), 0.5, device=device)
    mask = mask_dev.to(torch.int64)

    def call():
        return O.forward_with_cfg_eager(cfg, W, z_dev, t, cap_dev, mask, CFG_SCALE, 1.0, 1.0, 4096, True)

    with torch.no_grad():
        out = call()
        call()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            call()
        e1.record()
        torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / iters
    rel = ((out.float() - engine_out.float()).abs().max() / engine_out.float().abs().max()).item()
    return {"value": 1e3 / (ms * (NUM_STEPS - 1)), "unit": "latents/s", "ms_per_model_call": ms,
            "kind": "eager PyTorch bf16 restatement of the reference forward_with_cfg (cuBLAS + SDPA flash) on the same GPU, "
                    "device time of one model call x29",
            "rel_linf_vs_engine_one_call": rel}


# ------------------------------------------------------------------------------------------ engine arm
def build_flagship(device):
    from lumina_t2x_b200 import models
    torch.manual_seed(0)
    with torch.device(device):
        m = models.NextDiT_2B_GQA_patch2(qk_norm=True, cap_feat_dim=2048, max_tokens=4096, max_cap_len=T_CAP, max_batch=2)
    g = torch.Generator(device=device).manual_seed(0)
    with torch.no_grad():
        for k, p in m.named_parameters():
            # the reference zero-initialises adaLN / final layer / cap_embedder / gates (output would be 0): re-draw them
            if p.dim() == 2 and ("adaLN" in k or "final_layer.linear" in k or "cap_embedder.1" in k):
                p.copy_(torch.randn(p.shape, generator=g, device=device) * (0.5 / math.sqrt(p.shape[1])))
            elif k.endswith("attention.gate"):
                p.copy_(0.5 * torch.randn(p.shape, generator=g, device=device))
            elif p.dim() == 1 and k.endswith("bias"):
                p.copy_(0.02 * torch.randn(p.shape, generator=g, device=device))
    return m.eval().to(device, dtype=torch.bfloat16)


def run_engine(args, rank, local_rank, world):
    from lumina_t2x_b200 import _lib
    dist_on = world > 1
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if dist_on:
        import torch.distributed as dist
        # NCCL may print its version banner on stdout when the communicator is created: keep stdout clean for the
        # single JSON line by pointing fd 1 at stderr until the first collective has run.
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize(device)
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    m = build_flagship(device)
    lib, h = m._engine(device)
    stream = torch.cuda.current_stream(device)
    sp = C.c_void_p(stream.cuda_stream)

    g = torch.Generator().manual_seed(1000 + rank)
    z_host = torch.randn(1, 4, LATENT, LATENT, generator=g).to(torch.bfloat16).repeat(2, 1, 1, 1).contiguous().pin_memory()
    cap_host = torch.randn(2, T_CAP, 2048, generator=g).to(torch.bfloat16).contiguous().pin_memory()
    mask_host = torch.zeros(2, T_CAP, dtype=torch.uint8)
    mask_host[0, :] = 1
    mask_host[1, :8] = 1
    mask_host = mask_host.pin_memory()
    final_host = torch.empty_like(z_host).pin_memory()
    z_dev, cap_dev, mask_dev = z_host.to(device), cap_host.to(device), mask_host.to(device)
    final_dev = torch.empty_like(z_dev)
    grid = torch.linspace(0.0, 1.0, NUM_STEPS)          # transport/integrators.py:97-99, time_shifting_factor = 1
    grid = grid / (grid + 1.0 - 1.0 * grid)
    garr = (C.c_float * NUM_STEPS)(*[float(v) for v in grid])
    step = _lib.NditStepParams(CFG_SCALE, 1.0, 1.0, 1, 4096)

    def step_dev():
        _lib.check(lib.ndit_sample(h, C.c_void_p(z_dev.data_ptr()), 2, LATENT, LATENT, garr, NUM_STEPS, _lib.NDIT_EULER, C.byref(step),
                                   None, C.c_void_p(final_dev.data_ptr()), sp), h)

    def step_e2e():
        _lib.check(lib.ndit_sample_host(h, C.c_void_p(z_host.data_ptr()), C.c_void_p(cap_host.data_ptr()), C.c_void_p(mask_host.data_ptr()),
                                        2, LATENT, LATENT, T_CAP, garr, NUM_STEPS, _lib.NDIT_EULER, C.byref(step),
                                        C.c_void_p(final_host.data_ptr()), sp), h)

    def barrier():
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize(device)

    def max_over_ranks(x: float) -> float:
        if not dist_on:
            return x
        t = torch.tensor([x], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    _lib.check(lib.ndit_set_caption(h, C.c_void_p(cap_dev.data_ptr()), C.c_void_p(mask_dev.data_ptr()), 2, T_CAP, sp), h)
    for _ in range(args.warmup):
        step_dev()
    gathered = torch.empty((world,) + tuple(final_dev.shape), dtype=final_dev.dtype, device=device) if dist_on else None

    # ---- device-resident timing
    clocks = ClockSampler(physical_gpu_index(local_rank))
    barrier()
    clocks.start()
    l0 = lib.ndit_launch_count(h)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
    if dist_on:      # the one collective of the path: gather the final latents of all ranks
        dist.all_gather_into_tensor(gathered, final_dev)
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = int(lib.ndit_launch_count(h) - l0)
    clk = clocks.stop()
    assert torch.isfinite(final_dev.float()).all(), "non-finite latents"

    # ---- end to end through the host-buffer C ABI
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    h2d = z_host.numel() * 2 + cap_host.numel() * 2 + mask_host.numel()
    d2h = final_host.numel() * 2

    # ---- per-kernel-class CUDA-event profile of one more solve
    _lib.check(lib.ndit_set_option(h, b"profile", 1), h)
    step_dev()
    ms_c = (C.c_float * 7)()
    n_c = (C.c_int64 * 7)()
    _lib.check(lib.ndit_profile_read(h, ms_c, n_c, 7), h)
    _lib.check(lib.ndit_set_option(h, b"profile", 0), h)
    names = ["gemm_qkv", "gemm_wo", "gemm_w13_swiglu", "gemm_w2", "attention", "rowwise", "conditioning"]
    prof = {n: {"ms": float(ms_c[i]), "launches": int(n_c[i])} for i, n in enumerate(names)}

    if rank != 0:
        if dist_on:
            dist.destroy_process_group()
        return
    calls = NUM_STEPS - 1
    gemm_f, attn_f, cross_f = flops_per_forward()
    pk = peaks()
    total_tf = (gemm_f + attn_f + cross_f) * calls / 1e12
    lat_per_s = world * args.steps / (ms / 1e3)
    gemm_ms = sum(prof[k]["ms"] for k in names[:4])
    gemm_launches = sum(prof[k]["launches"] for k in names[:4])
    gemm_tflops = gemm_f * calls / 1e12 / (gemm_ms / 1e3) if gemm_ms > 0 else 0.0
    attn_tflops = (attn_f + cross_f) * calls / 1e12 / (prof["attention"]["ms"] / 1e3) if prof["attention"]["ms"] > 0 else 0.0
    prof_total = sum(v["ms"] for v in prof.values())
    line = {
        "metric": "latents/sec", "value": lat_per_s, "unit": "latents/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "latents_per_gpu_per_step": 1, "parallelism": f"dp{world} (independent latents per GPU, weights replicated, "
                   "one all-gather of the final latents)", "l2": "inputs larger than L2: 3.3 GB of bf16 weights stream per model call (L2 126 MB)",
                   "timestep_dtype": "bf16 (torchdiffeq casts t to the state dtype)"},
        "clocks": clk,
        "e2e": {"value": world * args.steps / e2e_s, "unit": "latents/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "ndit_sample_host (C ABI, pinned host buffers, H2D + D2H inside the timed region)"},
        "gpu_launches": launches,
        "algorithmic_tflop_per_latent": total_tf,
        "tflops_whole_path": total_tf * lat_per_s / world,
        "frac_of_bf16_peak_whole_path": total_tf * lat_per_s / world / pk["bf16"],
        "roofline": {"kernel": "gemm_bf16_tn_kernel (tcgen05, all four projections of the block)", "bound": "tensor",
                     "achieved": gemm_tflops, "peak": pk["bf16"], "unit": "TFLOP/s", "frac": gemm_tflops / pk["bf16"],
                     "traffic": ncu_traffic(), "peak_source": pk["src"], "launches_timed": gemm_launches,
                     "share_of_step": gemm_ms / prof_total if prof_total else None},
        "kernels": {**prof, "attention_tflops": attn_tflops, "attention_frac_of_peak": attn_tflops / pk["bf16"]},
    }
    if world == 1 and not args.no_gpu_eager_baseline:
        try:
            t05 = torch.full((2,), 0.5, device=device)
            eng = m.forward_with_cfg(z_dev, t05, cap_dev, mask_dev, CFG_SCALE, 1.0, 1.0, 4096, True)
            line["gpu_eager_baseline"] = gpu_eager_baseline(m, z_dev, cap_dev, mask_dev, device, eng)
        except Exception as ex:      # a baseline must never take the benchmark line down
            line["gpu_eager_baseline"] = {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}
    if world == 1 and not args.no_cpu_baseline:
        sec, _, _ = cpu_sample_once()
        line["cpu_baseline"] = {"value": 1.0 / sec, "unit": "latents/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "oracle (fp32 torch CPU port of the reference algorithm): full-shape forward_with_cfg with 0 and 2 of 24 "
                                          "blocks, per-block time extrapolated x24, x29 model calls"}
    print(json.dumps(line), flush=True)
    if dist_on:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            # convenience: re-launch under torchrun when started as a plain python process
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                   "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
    run_engine(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
