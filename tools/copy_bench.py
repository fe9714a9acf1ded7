#!/usr/bin/env python
"""Host-memory facts the result path is designed around (run on the GPU box): cost of page-locking, DMA rates."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from pyqms_b200 import _capi  # noqa: E402

lib = _capi.load()
GB = 1 << 30
for size in (GB // 4, GB, 4 * GB):
    p = C.c_void_p()
    t0 = time.perf_counter()
    assert lib.qms_host_alloc(C.byref(p), size) == 0
    t1 = time.perf_counter()
    dev = torch.empty(size, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    host = torch.from_numpy(np.ctypeslib.as_array((C.c_uint8 * size).from_address(p.value)))
    t2 = time.perf_counter()
    host.copy_(dev)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    host.copy_(dev)
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    dev.copy_(host)
    torch.cuda.synchronize()
    t5 = time.perf_counter()
    lib.qms_host_free(p)
    t6 = time.perf_counter()
    fresh = np.empty(size, dtype=np.uint8)
    t7 = time.perf_counter()
    fresh[::4096] = 1
    t8 = time.perf_counter()
    print("%.2f GB: cudaHostAlloc %.3f s (%.1f GB/s), D2H first %.1f GB/s, D2H again %.1f GB/s, H2D %.1f GB/s, free %.3f s, numpy first touch %.1f GB/s"
          % (size / GB, t1 - t0, size / GB / (t1 - t0), size / GB / (t3 - t2), size / GB / (t4 - t3), size / GB / (t5 - t4), t6 - t5,
             size / GB / (t8 - t7)))
