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


def test_new_entry_points_reject_null_arguments():
    """lsvs_peer_* / lsvs_pose_chain_gt validate their arguments before touching CUDA."""
    from lsvs_b200 import native
    lib = native.lib()
    assert lib.lsvs_peer_put(None, None, ctypes.c_size_t(16), None) == -1
    assert lib.lsvs_peer_signal(None, ctypes.c_uint(1), None) == -1
    assert lib.lsvs_peer_wait(None, ctypes.c_uint(1), None, ctypes.c_double(1.0), None) == -1
    assert lib.lsvs_peer_alloc(ctypes.c_size_t(0), None) == -1
    assert lib.lsvs_peer_free(None) == 0 and lib.lsvs_peer_close(None) == 0
    z = ctypes.c_int(1)
    rc = lib.lsvs_pose_chain_gt(None, None, None, None, z, z, z, z, z, z, None, ctypes.c_int(3), ctypes.c_int(3), None, None, None, None)
    assert rc == -1 and b"gt_poses" in lib.lsvs_last_error()


def test_no_cpu_fallback():
    import torch
    from aligned_vggt.utils import alignment as A
    from lsvs_b200 import native
    with pytest.raises(native.NativeError):
        A.apply_sim3_alignment_on_point_maps(torch.zeros(1, 1, 2, 2, 3), torch.eye(4)[None], torch.ones(1))


def test_documents_name_only_exported_entry_points():
    """Every lsvs_* entry point that INTEGRATION.md / DESIGN.md / README.md mention is declared in include/lsvs_b200.h."""
    import os
    import re
    from lsvs_b200 import native
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    declared = set(native.declared_symbols())
    families = ("lsvs_peer_", "lsvs_sim3_apply_", "lsvs_dpt_", "lsvs_engine_")       # documents also use family prefixes / wildcards
    for doc in ("INTEGRATION.md", "DESIGN.md", "README.md"):
        text = open(os.path.join(root, doc)).read()
        for name in set(re.findall(r"\blsvs_[a-z0-9_]+\b", text)):
            if name in declared or name.startswith("lsvs_b200") or name in ("lsvs_engine", "lsvs_bf16", "lsvs_gemm_epilogue", "lsvs_engine_config"):
                continue
            assert any(name == f or name == f.rstrip("_") or (name.startswith(f) and any(d.startswith(name) for d in declared)) for f in families) \
                or any(d.startswith(name) for d in declared), f"{doc} mentions {name}, which include/lsvs_b200.h does not declare"
