"""The C ABI without Python: examples/ndit_host_demo.c (plain C, links libndit_b200.so and libcudart only) builds a small NextDiT
from counter-based pseudo-random weights, runs a 5-point Euler solve from host buffers and prints a checksum.  This test
compiles and runs it, rebuilds the same weights and inputs in Python, runs the reference-API mirror (models.NextDiT +
transport.Sampler) and requires bit-identical latents."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _fnv(key: bytes) -> int:
    h = 0xcbf29ce484222325
    for b in key:
        h = ((h ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


def _urand(seed: int, n: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = np.uint64((seed * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF) + np.arange(n, dtype=np.uint64)
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return ((x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) * 2.0 - 1.0).astype(np.float32)


def _tensor(key: str, shape, center: float, amp: float) -> torch.Tensor:
    n = int(np.prod(shape))
    v = np.float32(center) + np.float32(amp) * _urand(_fnv(key.encode()), n)
    return torch.from_numpy(v.astype(np.float32)).to(torch.bfloat16).view(*shape)


def _recipe(key: str):
    if key.endswith(".bias") or key == "pad_token":
        return 0.0, 0.02
    if key.endswith("attention.gate"):
        return 0.0, 0.5
    if key.endswith("norm.weight") or key.endswith("norm1.weight") or key.endswith("norm2.weight") or key == "cap_embedder.0.weight":
        return 1.0, 0.1
    table = {"x_embedder.weight": 0.25, "t_embedder.mlp.0.weight": 0.06, "t_embedder.mlp.2.weight": 0.04, "cap_embedder.1.weight": 0.06,
             "final_layer.linear.weight": 0.04, "final_layer.adaLN_modulation.1.weight": 0.02}
    if key in table:
        return 0.0, table[key]
    for suf, amp in (("wq.weight", 0.04), ("wk.weight", 0.04), ("wv.weight", 0.04), ("wo.weight", 0.04), ("wk_y.weight", 0.06),
                     ("wv_y.weight", 0.06), ("w1.weight", 0.04), ("w3.weight", 0.04), ("w2.weight", 0.025),
                     ("adaLN_modulation.1.weight", 0.02)):
        if key.endswith(suf):
            return 0.0, amp
    raise KeyError(key)


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_plain_c_host_matches_python_mirror(tmp_path):
    exe = str(tmp_path / "ndit_host_demo")
    libdir = os.path.join(ROOT, "lumina_t2x_b200")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "examples", "ndit_host_demo.c"), "-o", exe, "-L", libdir, "-lndit_b200", "-L", os.path.join(cuda, "lib64"),
           "-lcudart", f"-Wl,-rpath,{libdir}", "-lm"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=300).stdout
    m = re.search(r"params=(\d+) launches=(\d+) sum_abs=([0-9.]+) fnv64=([0-9a-f]{16})", out)
    assert m, out
    c_params, c_launches, c_sum, c_fnv = int(m.group(1)), int(m.group(2)), float(m.group(3)), int(m.group(4), 16)

    from lumina_t2x_b200 import models, transport
    net = models.NextDiT(dim=576, n_layers=2, n_heads=8, n_kv_heads=2, qk_norm=True, cap_feat_dim=256, max_tokens=256, max_cap_len=32)
    sd = {k: _tensor(k, tuple(v.shape), *_recipe(k)) for k, v in net.state_dict().items()}
    net.load_state_dict(sd, strict=True)
    net = net.eval().to("cuda", dtype=torch.bfloat16)
    assert net.parameter_count() == c_params
    z1 = _tensor("z", (1, 4, 32, 32), 0.0, 1.7)
    z = torch.cat([z1, z1], 0).cuda()
    cap = _tensor("cap", (2, 16, 256), 0.0, 1.5).cuda()
    mask = torch.zeros(2, 16, dtype=torch.int64)
    mask[0, :] = 1
    mask[1, :4] = 1
    fn = transport.Sampler(transport.create_transport("Linear", "velocity")).sample_ode(sampling_method="euler", num_steps=5)
    lat = fn(z, net.forward_with_cfg, cap_feats=cap, cap_mask=mask.cuda(), cfg_scale=4.0, proportional_attn=True, base_seqlen=64)[-1]
    bits = lat.contiguous().view(torch.int16).cpu().numpy().astype(np.uint16).ravel()
    h = 0xcbf29ce484222325
    for b in bits.tolist():
        h = ((h ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    assert abs(lat.float().abs().sum().item() - c_sum) <= 1e-3 * c_sum
    assert h == c_fnv, (hex(h), hex(c_fnv))
    assert c_launches > 0
