"""Times ndit_op_resid_rms_mod on the config-2 shape (8192 x 2304) and prints a checksum; run once per NDIT_RESID4 value.
usage: NDIT_RESID4=0|1 python tools/resid_bench.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lumina_t2x_b200 import _lib

lib = _lib.load()
M, rows, D = 8192, 4096, 2304
g = torch.Generator(device="cuda").manual_seed(5)
X0 = torch.randn(M, D, device="cuda", generator=g).to(torch.bfloat16)
o = torch.randn(M, D, device="cuda", generator=g).to(torch.bfloat16)
w_post = (1 + 0.1 * torch.randn(D, device="cuda", generator=g)).to(torch.bfloat16)
w_pre = (1 + 0.1 * torch.randn(D, device="cuda", generator=g)).to(torch.bfloat16)
gt = torch.tanh(torch.randn(M // rows, D, device="cuda", generator=g)).to(torch.bfloat16).contiguous()
op = (1 + 0.3 * torch.randn(M // rows, D, device="cuda", generator=g)).to(torch.bfloat16).contiguous()
u = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
p = lambda t: C.c_void_p(t.data_ptr())


def call(X):
    rc = lib.ndit_op_resid_rms_mod(p(X), p(o), p(w_post), p(gt), p(w_pre), p(op), None, p(u), M, rows, D, 1e-5, None)
    assert rc == 0, lib.ndit_last_error(None)


X = X0.clone()
call(X)
torch.cuda.synchronize()
torch.save({"X": X.cpu(), "u": u.cpu()}, f"gpurun_out/resid_out_{os.environ.get('NDIT_RESID4', 'default')}.pt")
Xs = [X0.clone() for _ in range(8)]      # separate residual buffers so the data does not stay in L2 between calls
for Xi in Xs:
    call(Xi)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    for Xi in Xs:
        call(Xi)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 40
print(f"NDIT_RESID4={os.environ.get('NDIT_RESID4', 'default')}: {us:.1f} us per launch, {4 * M * D * 2 / us / 1e6:.2f} TB/s algorithmic", flush=True)
