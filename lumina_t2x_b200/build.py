"""Build libndit_b200.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build()."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libndit_b200.so")
SOURCES = ["engine.cu", "gemm_tcgen05.cu", "attention_tcgen05.cu", "attention_hr_tcgen05.cu", "rowwise.cu", "tensormap.cu", "text_encoder.cu", "vae_decoder.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, variant: str = "", defs: str = "") -> str:
    """variant/defs: experimental builds (extra -D flags) go to build_<variant>/ and libndit_b200_<variant>.so."""
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build" + ("_" + variant if variant else ""))
    lib_path = LIB if not variant else os.path.join(HERE, f"libndit_b200_{variant}.so")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "ndit.h"))
    headers.append(os.path.join(os.path.dirname(HERE), "include", "ndit_text.h"))
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        extra = (defs or os.environ.get("NDIT_NVCC_DEFS", "")).split()     # e.g. "-DAT_TIMING -DAT_POLY_PER8=0"
        r = subprocess.run([nvcc] + NVCC_FLAGS + extra + ["-c", s, "-o", o], capture_output=True, text=True)
        return s, r

    with ThreadPoolExecutor(max_workers=4) as ex:
        for s, r in ex.map(compile_one, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(f"--- {os.path.basename(s)}\n{r.stdout}{r.stderr}\n")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {s}")
            with open(os.path.join(objdir, os.path.basename(s) + ".ptxas.log"), "w") as f:
                f.write(r.stdout + r.stderr)
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(lib_path, objs):
        r = subprocess.run([nvcc, "-shared", "-o", lib_path, "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"] + objs,
                           capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
