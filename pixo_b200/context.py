"""Context — owner of a pixo_b200_ctx (one CUDA stream + reusable device/pinned scratch)."""
from __future__ import annotations

import ctypes as C
import threading

from . import _lib


class Context:
    def __init__(self, device: int = 0):
        lib = _lib.load()
        h = C.c_void_p()
        rc = lib.pixo_b200_ctx_create(device, C.byref(h))
        if rc != 0:
            raise _lib.PixoError(rc, (lib.pixo_b200_last_error(None) or b"").decode())
        self._h = h
        self.device = device

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().pixo_b200_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream_ptr: int | None):
        _lib.check(self._h, _lib.load().pixo_b200_ctx_set_stream(self._h, cuda_stream_ptr or None))

    def sync(self):
        _lib.check(self._h, _lib.load().pixo_b200_ctx_sync(self._h))

    def set_host_threads(self, n: int):
        _lib.check(self._h, _lib.load().pixo_b200_ctx_set_host_threads(self._h, n))

    def set_scan_capacity(self, bytes_per_frame: int = 0, gpu_retry: bool = True):
        """Device scan buffer per frame (0 = heuristic) and whether an overflowing frame is re-coded on
        the GPU (test hook: a tiny capacity with gpu_retry=False forces the host entropy coder)."""
        _lib.check(self._h, _lib.load().pixo_b200_ctx_set_scan_capacity(self._h, int(bytes_per_frame), int(gpu_retry)))

    @property
    def host_fallbacks(self) -> int:
        """Frames finished by the host entropy coder since the context was created."""
        return int(_lib.load().pixo_b200_ctx_host_fallbacks(self._h))

    @property
    def launch_count(self) -> int:
        return int(_lib.load().pixo_b200_ctx_launch_count(self._h))


_tls = threading.local()


def default_context() -> Context:
    """One lazily-created context per host thread (the library's threading contract)."""
    ctx = getattr(_tls, "ctx", None)
    if ctx is None or ctx.handle is None:
        ctx = Context(0)
        _tls.ctx = ctx
    return ctx
