"""CPU emulation of the lsvs_peer_* entry points (include/lsvs_b200.h) over named shared host memory, so that the Python side of
the mailbox transport (lsvs_b200.scheduler.PeerTransport: layout, slot reuse, sequence flags, short chunks) runs in the CPU
suite under gloo.  "Device memory" = a multiprocessing.shared_memory block, an IPC handle = its name, a put = memmove, a signal =
a 32-bit store, a wait = polling on the host (stricter than the stream-ordered wait of the CUDA kernels: it blocks the caller)."""
import ctypes
import time
from multiprocessing import shared_memory


class NativeError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        self._blocks = {}      # address -> SharedMemory (kept alive)

    def _addr(self, shm):
        return ctypes.addressof(ctypes.c_char.from_buffer(shm.buf))

    def lsvs_peer_alloc(self, nbytes, ref):
        shm = shared_memory.SharedMemory(create=True, size=int(nbytes))
        shm.buf[:int(nbytes)] = bytes(int(nbytes))
        a = self._addr(shm)
        self._blocks[a] = shm
        ref._obj.value = a
        return 0

    def lsvs_peer_export(self, ptr, buf):
        name = self._blocks[ptr.value].name.encode()
        assert len(name) < 64
        for i in range(64):
            buf[i] = name[i] if i < len(name) else 0
        return 0

    def lsvs_peer_open(self, buf, ref):
        name = bytes(buf).split(b"\0")[0].decode()
        shm = shared_memory.SharedMemory(name=name)
        a = self._addr(shm)
        self._blocks[a] = shm
        ref._obj.value = a
        return 0

    def lsvs_peer_put(self, dst, src, nbytes, stream):
        ctypes.memmove(int(dst), int(src), int(nbytes))
        return 0

    def lsvs_peer_signal(self, flag, value, status, stream):
        if status and ctypes.c_int32.from_address(int(status)).value != 0:
            ctypes.c_uint32.from_address(int(flag) + 4).value = 1   # poison instead of the sequence number
            return 0
        ctypes.c_uint32.from_address(int(flag)).value = int(value)
        return 0

    def lsvs_peer_wait(self, flag, value, status, timeout_s, stream):
        st = ctypes.c_int32.from_address(int(status))
        if st.value != 0:
            return 0
        t0 = time.time()
        while ((ctypes.c_uint32.from_address(int(flag)).value - int(value)) & 0xFFFFFFFF) >= 0x80000000:
            if ctypes.c_uint32.from_address(int(flag) + 4).value != 0:
                st.value = 2
                return 0
            if time.time() - t0 > timeout_s:
                st.value = 1
                return 0
            time.sleep(0.0005)
        return 0

    def _drop(self, ptr, unlink):
        shm = self._blocks.pop(ptr.value, None)
        if shm is not None:
            try:
                shm.close()
                if unlink:
                    shm.unlink()
            except (BufferError, FileNotFoundError):
                pass
        return 0

    def lsvs_peer_close(self, ptr):
        return self._drop(ptr, unlink=False)

    def lsvs_peer_free(self, ptr):
        return self._drop(ptr, unlink=True)


_LIB = _Lib()


def lib():
    return _LIB


def stream_ptr():
    return None


def check(rc, what=""):
    if rc != 0:
        raise NativeError(f"{what} failed ({rc})")
