"""Build liblsenerf_b200.so (the C-ABI library of sm_100a kernels) in-tree with nvcc.

    python -m lsenerf_b200.build [--force] [--verbose] [--variant NAME]

Each .cu is compiled to an object with ``-gencode arch=compute_100a,code=sm_100a -lineinfo`` (objects cached by
mtime) and linked into ``lsenerf_b200/liblsenerf_b200.so``.  ``march.cu`` is compiled with ``-fmad=false``: the
DDA recurrences must round exactly like the CPU oracle.  No torch headers are involved: the ABI is plain C.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "liblsenerf_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "lsenerf_b200.h")

SOURCES = ["abi.cu", "raygen.cu", "march.cu", "hashgrid.cu", "field.cu", "field_simt.cu", "field_tc.cu", "field_tc_bwd.cu", "field_hash.cu",
           "composite.cu", "loss.cu", "mapper_mlp.cu", "adam_collective.cu", "microbench.cu", "microbench_mma.cu"]
EXTRA = {"march.cu": ["-fmad=false"]}
COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"] + os.environ.get("NVCC_EXTRA", "").split()

def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _deps_mtime() -> float:
    m = os.path.getmtime(HEADER)
    for f in os.listdir(CSRC):
        if f.endswith(".cuh"):
            m = max(m, os.path.getmtime(os.path.join(CSRC, f)))
    return m


def _compile(src: str, force: bool, verbose: bool, obj_dir: str = OBJ) -> str:
    obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
    spath = os.path.join(CSRC, src)
    newest = max(os.path.getmtime(spath), _deps_mtime())
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
        return obj
    cmd = [_nvcc(), *COMMON, *EXTRA.get(src, []), "-c", spath, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(obj + ".log", "w") as fh:
        fh.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{log}")
    if verbose:
        print(log)
    return obj


def build(force: bool = False, verbose: bool = False, variant: str = "") -> str:
    """variant: build into lsenerf_b200/_obj_<variant>/ -> liblsenerf_b200_<variant>.so next to the default library, for A/B runs of
    an experimental build (`NVCC_EXTRA="-DFOO=1" python -m lsenerf_b200.build --variant foo`, then `LSE_B200_LIB=.../liblsenerf_b200_foo.so`)."""
    obj_dir = OBJ + (f"_{variant}" if variant else "")
    lib = LIB.replace(".so", f"_{variant}.so") if variant else LIB
    os.makedirs(obj_dir, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose, obj_dir), SOURCES))
    if force or not os.path.exists(lib) or any(os.path.getmtime(o) > os.path.getmtime(lib) for o in objs):
        cmd = [_nvcc(), "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return lib


if __name__ == "__main__":
    _variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else ""
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, variant=_variant))
