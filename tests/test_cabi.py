"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/lsvs_b200.h declares."""
import ctypes

import pytest


def test_build_and_symbols():
    import __graft_entry__ as g
    g.build()
    from lsvs_b200 import native
    lib = native.lib()
    syms = native.declared_symbols()
    assert len(syms) >= 7
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/lsvs_b200.h but not exported"
    assert lib.lsvs_version() >= 1


def test_argument_errors_without_gpu():
    """Shape errors are reported before any CUDA call (reference: AssertionError / ValueError)."""
    from lsvs_b200 import native
    lib = native.lib()
    rc = lib.lsvs_sim3_apply_points(None, None, None, None, ctypes.c_int(1), ctypes.c_longlong(4), None)
    assert rc == -1 and b"null" in lib.lsvs_last_error()
    with pytest.raises(ValueError):
        native.check(rc, "sim3_apply_points")


def test_no_cpu_fallback():
    import torch
    from aligned_vggt.utils import alignment as A
    from lsvs_b200 import native
    with pytest.raises(native.NativeError):
        A.apply_sim3_alignment_on_point_maps(torch.zeros(1, 1, 2, 2, 3), torch.eye(4)[None], torch.ones(1))
