"""One kj_ctx per process per GPU (one process per GPU; multi-GPU = torch.distributed in dist.py)."""
from __future__ import annotations

import ctypes as C
import threading

from . import _abi

_lock = threading.Lock()
_contexts = {}


class Context:
    def __init__(self, device: int = 0, stream=None):
        L = _abi.lib()
        h = C.c_void_p()
        rc = L.kj_init(device, stream, C.byref(h))
        if rc < 0:
            raise _abi.KjError(rc, (L.kj_last_error(None) or b"").decode())
        self.handle = h
        self.device = device
        self._L = L

    def close(self):
        if self.handle:
            self._L.kj_destroy(self.handle)
            self.handle = None

    @property
    def launches(self) -> int:
        return int(self._L.kj_launch_count(self.handle))

    def set_rounding_mode(self, mode: int):
        _abi.check(self._L.kj_set_rounding_mode(self.handle, mode), self.handle)

    def enable_timers(self, on: bool = True):
        self._L.kj_enable_timers(self.handle, 1 if on else 0)

    def set_stage_chunk(self, nbytes: int):
        """Size of the host -> device staging chunks of add_host / add_file."""
        _abi.check(self._L.kj_set_stage_chunk(self.handle, int(nbytes)), self.handle)

    def reset_timers(self):
        self._L.kj_reset_timers(self.handle)

    def scan_kernel_stats(self):
        """(average ms per launch, launches, bytes owned by those launches)"""
        n = C.c_uint64()
        ms = self._L.kj_scan_kernel_ms(self.handle, C.byref(n))
        return float(ms), int(n.value), int(self._L.kj_scan_kernel_bytes(self.handle))

    def verify_kernel_ms(self) -> float:
        return float(self._L.kj_verify_kernel_ms(self.handle))


def default_context(device: int | None = None) -> Context:
    """The process-wide context of `device` (default: LOCAL_RANK, else 0)."""
    import os
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    with _lock:
        ctx = _contexts.get(device)
        if ctx is None:
            ctx = Context(device)
            _contexts[device] = ctx
        return ctx
