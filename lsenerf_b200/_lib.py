"""ctypes binding of liblsenerf_b200.so (the C ABI declared in include/lsenerf_b200.h).

The prototypes are parsed from the header itself so the Python side can never drift from the ABI.  There is no
CPU fallback: ``lib()`` raises if the shared library has not been built (``python -m lsenerf_b200.build``), and every
op that reaches it requires CUDA tensors.
"""
from __future__ import annotations

import ctypes
import os
import re

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
# LSE_B200_LIB: another build of the same library (lsenerf_b200.build --variant ...), for A/B runs of build-time experiments
LIB_PATH = os.environ.get("LSE_B200_LIB") or os.path.join(HERE, "liblsenerf_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "lsenerf_b200.h")


class LseFieldParams(ctypes.Structure):
    _fields_ = [("table", ctypes.c_void_p), ("scalings", ctypes.c_void_p), ("levels", ctypes.c_int32),
                ("log2_T", ctypes.c_int32),
                ("W1", ctypes.c_void_p), ("b1", ctypes.c_void_p), ("W2", ctypes.c_void_p), ("b2", ctypes.c_void_p),
                ("W3", ctypes.c_void_p), ("b3", ctypes.c_void_p), ("W4", ctypes.c_void_p), ("b4", ctypes.c_void_p),
                ("W5", ctypes.c_void_p), ("b5", ctypes.c_void_p),
                ("emb", ctypes.c_void_p), ("emb_rows", ctypes.c_int32), ("emb_mode", ctypes.c_int32),
                ("sh_shifted", ctypes.c_int32), ("use_tensor_cores", ctypes.c_int32),
                ("relu_mask_density", ctypes.c_void_p), ("relu_mask_color", ctypes.c_void_p), ("rgb_fwd", ctypes.c_void_p), ("color_act", ctypes.c_void_p)]


class LseFieldGrads(ctypes.Structure):
    _fields_ = [("d_table", ctypes.c_void_p),
                ("dW1", ctypes.c_void_p), ("db1", ctypes.c_void_p), ("dW2", ctypes.c_void_p), ("db2", ctypes.c_void_p),
                ("dW3", ctypes.c_void_p), ("db3", ctypes.c_void_p), ("dW4", ctypes.c_void_p), ("db4", ctypes.c_void_p),
                ("dW5", ctypes.c_void_p), ("db5", ctypes.c_void_p), ("d_emb", ctypes.c_void_p)]


class LseEpilogueCfg(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("use_mapping", "map_mode", "rgb_mapper", "evs_mapper", "ev_one_dim",
                                               "training", "deblur", "want_ev")]


class LseMlpMapper(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4")] + [("dim", ctypes.c_int32), ("pad", ctypes.c_int32)]


class LseMlpMapperGrads(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("dW1", "db1", "dW2", "db2", "dW3", "db3", "dW4", "db4")]


_STRUCTS = {"LseFieldParams": LseFieldParams, "LseFieldGrads": LseFieldGrads, "LseEpilogueCfg": LseEpilogueCfg,
            "LseMlpMapper": LseMlpMapper, "LseMlpMapperGrads": LseMlpMapperGrads}
_SCALARS = {"int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "float": ctypes.c_float,
            "double": ctypes.c_double, "lse_stream_t": ctypes.c_void_p}


def parse_header(path: str = HEADER_PATH):
    """-> {name: (restype, [argtypes])} for every function the header declares."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int64_t|int|const char \*)\s*(lse_\w+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "const char *": ctypes.c_char_p}[ret]
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.replace("[host]", "").split())
                if "*" in a:
                    base = a.replace("const", "").split("*")[0].strip()
                    argtypes.append(ctypes.POINTER(_STRUCTS[base]) if base in _STRUCTS else ctypes.c_void_p)
                else:
                    argtypes.append(_SCALARS[a.split()[0]])
        protos[name] = (restype, argtypes)
    return protos


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m lsenerf_b200.build` "
                               "(the CUDA extension is mandatory; there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in parse_header().items():
            fn = getattr(L, name)  # AttributeError here == header/library mismatch
            fn.restype, fn.argtypes = restype, argtypes
        if L.lse_abi_version() != 1:
            raise RuntimeError("liblsenerf_b200.so ABI version mismatch")
        _lib = L
    return _lib


def P(t):
    """device pointer of a tensor (None -> NULL); the tensor must be CUDA and contiguous."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("lsenerf_b200 ops run on CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("lsenerf_b200 ops need contiguous tensors")
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_prof = None


def profile_start():
    """Bracket every ABI call with CUDA events on the launching stream (bench.py roofline; never on during timing)."""
    global _prof
    _prof = []


def profile_stop():
    """-> {abi_function: (calls, total_ms)}; synchronises."""
    global _prof
    rec, _prof = _prof or [], None
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in rec:
        c, t = out.get(name, (0, 0.0))
        out[name] = (c + 1, t + e0.elapsed_time(e1))
    return out


def call(name: str, *args):
    if _prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib(), name)(*args)
        e1.record()
        _prof.append((name, e0, e1))
    else:
        rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib().lse_last_error().decode()}")


def launch_count() -> int:
    return int(lib().lse_launch_count())
