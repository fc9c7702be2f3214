"""CPU: the C-ABI library loads without a GPU and exports every symbol include/lsenerf_b200.h declares."""
import ctypes
import os

import pytest

from lsenerf_b200 import _lib


def test_header_parses():
    protos = _lib.parse_header()
    assert len(protos) >= 28
    for must in ("lse_pose_interp_fwd", "lse_march_count", "lse_hashgrid_fwd", "lse_hashgrid_bwd", "lse_field_density_fwd",
                 "lse_field_color_bwd", "lse_composite_fwd", "lse_epilogue_fwd", "lse_loss_fwd", "lse_adam_step"):
        assert must in protos


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        from lsenerf_b200 import build
        build.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in _lib.parse_header():
        assert hasattr(L, name), name
    lib = _lib.lib()
    assert lib.lse_abi_version() == 1
    assert isinstance(_lib.launch_count(), int)


def test_no_cpu_fallback():
    """ops refuse CPU tensors loudly instead of computing on the host."""
    import torch
    from lsenerf_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.hash_indices(torch.rand(4, 3), torch.ones(16), 19)


def test_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.LseEpilogueCfg) == 8 * 4
    assert ctypes.sizeof(_lib.LseFieldGrads) == 12 * 8
    assert ctypes.sizeof(_lib.LseFieldParams) == 2 * 8 + 2 * 4 + 10 * 8 + 8 + 4 * 4 + 4 * 8
