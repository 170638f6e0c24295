"""Runs a few full-size forward_with_cfg calls (config 2 shapes) for ncu captures.
usage: python tools/one_forward.py [n_calls]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
m = bench.build_flagship(dev)
g = torch.Generator().manual_seed(1)
z = torch.randn(1, 4, 128, 128, generator=g).to(torch.bfloat16).repeat(2, 1, 1, 1).to(dev)
cap = torch.randn(2, 128, 2048, generator=g).to(torch.bfloat16).to(dev)
mask = torch.zeros(2, 128, dtype=torch.int32)
mask[0, :] = 1
mask[1, :8] = 1
mask = mask.to(dev)
t = torch.full((2,), 0.3, device=dev)
for _ in range(n):
    out = m.forward_with_cfg(z, t, cap, mask, 2.0, 1.0, 1.0, 4096, True)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
