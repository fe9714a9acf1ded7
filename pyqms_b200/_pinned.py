"""Reusable page-locked host buffers for packed hit arrays ("pinned-host result rings").

Hits leave the device by DMA straight into these blocks; a block goes back to the pool when the last numpy array
that views it is garbage collected, so a loop over runs reuses the same few gigabytes instead of faulting in fresh
pages per run (3.5 -> 14 GB/s with staged copies into new arrays, link speed into a reused pinned block)."""
import ctypes as C

import numpy as np

from . import _capi

_LOCK_LATER_BYTES = 32 << 20     # arrays from this size on are page-locked from their second use
_LIMIT_BYTES = 12 << 30          # blocks kept for reuse; beyond it freed blocks are released to the OS


class _Block:
    """One cudaHostAlloc block; views of it keep it alive, the last one returns it to the pool."""

    __slots__ = ("ptr", "nbytes", "pool", "__weakref__")

    def __init__(self, ptr, nbytes, pool):
        self.ptr, self.nbytes, self.pool = ptr, nbytes, pool

    def __del__(self):
        pool = self.pool
        if pool is not None and self.ptr:
            try:
                pool._give_back(self.ptr, self.nbytes)
            except Exception:
                pass


class _View:
    """numpy-facing window on a block (``np.asarray(view)`` keeps ``view`` -- and with it the block -- as its base)."""

    def __init__(self, block, dtype, count):
        self._block = block
        self.__array_interface__ = {"shape": (int(count),), "typestr": np.dtype(dtype).str, "data": (block.ptr, False), "version": 3}


class PinnedPool:
    def __init__(self):
        self._free = {}           # size class -> [ptr]
        self._held = 0
        self._seen = set()        # size classes asked for before
        self._lib = _capi.load()

    @staticmethod
    def _size_class(nbytes):
        size = 1 << 16
        while size < nbytes:
            size += size >> 1 if size >= (1 << 20) else size      # x2 below 1 MiB, x1.5 above
        return size

    def array(self, dtype, count):
        """Uninitialised 1-D array of ``count`` elements in page-locked memory."""
        nbytes = self._size_class(max(1, int(count)) * np.dtype(dtype).itemsize)
        if nbytes >= _LOCK_LATER_BYTES and nbytes not in self._seen:
            # page-locking costs ~0.4 s per GB (measured: cudaHostAlloc 2.3 GB/s), more than a staged copy into a
            # pageable array (~14 GB/s): a size is only locked once it comes back, i.e. in a loop over runs
            self._seen.add(nbytes)
            return np.empty(int(count), dtype=dtype)
        stack = self._free.get(nbytes)
        if stack:
            ptr = stack.pop()
            self._held -= nbytes
        else:
            out = C.c_void_p()
            rc = self._lib.qms_host_alloc(C.byref(out), nbytes)
            if rc != _capi.QMS_OK:
                self.release()                                  # give cached blocks back and try once more
                rc = self._lib.qms_host_alloc(C.byref(out), nbytes)
            if rc != _capi.QMS_OK:
                raise _capi.QmsError(rc, (self._lib.qms_last_error(None) or b"").decode())
            ptr = out.value
        return np.asarray(_View(_Block(ptr, nbytes, self), dtype, count))

    def _give_back(self, ptr, nbytes):
        if self._held + nbytes > _LIMIT_BYTES:
            self._lib.qms_host_free(ptr)
            return
        self._free.setdefault(nbytes, []).append(ptr)
        self._held += nbytes

    def release(self):
        for stack in self._free.values():
            for ptr in stack:
                self._lib.qms_host_free(ptr)
        self._free.clear()
        self._held = 0

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass
