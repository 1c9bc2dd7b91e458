"""ctypes binding of include/lsvs_b200.h (the C-ABI drop-in boundary).

There is no CPU fallback: if liblsvs_b200.so is missing or a call fails, this raises.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("LSVS_B200_LIB") or os.path.join(_HERE, "lib", "liblsvs_b200.so")  # override: A/B measurement builds
HEADER = os.path.join(_HERE, "..", "include", "lsvs_b200.h")

_lib = None


class NativeError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(f"{LIB_PATH} not built: run `python large-scale-vit-slam_b200/lsvs_b200/build.py` "
                              "(there is no CPU fallback for this path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.lsvs_last_error.restype = ctypes.c_char_p
        _lib.lsvs_launch_count.restype = ctypes.c_ulonglong
    return _lib


def declared_symbols():
    """Every function name declared in include/lsvs_b200.h."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lsvs_[a-z0-9_]+)\s*\(", src)))


def check(rc, what=""):
    if rc != 0:
        msg = lib().lsvs_last_error().decode(errors="replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise NativeError(f"{what} failed ({rc}): {msg}")


def ptr(t):
    """Device/host pointer of a torch tensor (or None) as c_void_p."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(lib().lsvs_launch_count())
