"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests).

* library build (P1): envelopes are sharded by index range over the ranks; every rank computes its
  shard into the global slot layout (identical on all ranks) and the shards are assembled with
  all-gathers of the per-slot / per-envelope device arrays.  Index finalisation
  then runs redundantly so every replica is identical.
* matching (P2): spectra are sharded by contiguous ranges, the library is replicated, there is no
  data-path collective; rank 0 may gather the packed hits in rank order = spectrum order.
"""
import ctypes as C
import sys

import numpy as np


def _dist():
    # a process group can only exist if the caller imported torch.distributed; importing torch here would add
    # seconds to a single-process library build for nothing
    dist = sys.modules.get("torch.distributed")
    if dist is None:
        return None
    return dist if dist.is_available() and dist.is_initialized() else None


def world_size():
    d = _dist()
    return d.get_world_size() if d else 1


def rank():
    d = _dist()
    return d.get_rank() if d else 0


def bind_to_gpu_numa_node(device):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that the pinned spectrum buffers
    (first touch) and the staging copies live next to the GPU's PCIe root.  Returns the node or None."""
    import os
    try:
        import torch
        props = torch.cuda.get_device_properties(device)
        bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        return None
    return None


def shard_range(n, r, world):
    """Contiguous block partition of range(n): rank r gets [begin, end); sizes differ by at most 1."""
    base, extra = divmod(int(n), int(world))
    begin = r * base + min(r, extra)
    return begin, begin + base + (1 if r < extra else 0)


def shard_slices(boundaries, world):
    """[(begin, end)] per rank over ``len(boundaries) - 1`` units."""
    n = len(boundaries) - 1
    return [shard_range(n, r, world) for r in range(world)]


def _device_tensor(ptr_value, n_elems, dtype, device):
    """Zero-copy torch view of a device buffer owned by the qms context (``__cuda_array_interface__``)."""
    import torch

    class _Wrap:
        pass
    np_dtype = np.dtype(dtype)
    w = _Wrap()
    w.__cuda_array_interface__ = {"shape": (int(n_elems),), "typestr": np_dtype.str, "data": (int(ptr_value), False),
                                  "version": 2}
    return torch.as_tensor(w, device="cuda:%d" % device)


def allgather_ranges(full, ranges):
    """In-place assembly of a 1-D tensor whose slice ``ranges[r]`` was filled by rank r.

    Shard sizes differ by construction, so each rank pads its slice to the widest shard, the padded
    pieces are all-gathered (NCCL on GPU tensors, gloo on CPU tensors) and every foreign piece is copied to
    its global offset.  Returns the number of bytes this rank received."""
    import torch
    import torch.distributed as dist
    me, world = dist.get_rank(), dist.get_world_size()
    widest = max(e - b for b, e in ranges)
    if widest == 0:
        return 0
    mine = torch.zeros(widest, dtype=full.dtype, device=full.device)
    b, e = ranges[me]
    mine[: e - b] = full[b:e]
    pieces = [torch.empty(widest, dtype=full.dtype, device=full.device) for _ in range(world)]
    dist.all_gather(pieces, mine)
    received = 0
    for r, (rb, re) in enumerate(ranges):
        if r != me and re > rb:
            full[rb:re] = pieces[r][: re - rb]
            received += (re - rb) * full.element_size()
    return received


def allgather_envelopes(ctx, n_env, world):
    """Assemble the envelope shards of all ranks in place (NCCL all-gather over NVLink).

    Every rank has written its shard [begin_r, end_r) of envelopes into the shared slot layout; the per-slot
    arrays (mass, relabun, abun, pos, c_flag) and per-envelope arrays (len, n_c) are gathered one by one."""
    import torch
    n_e, n_s = C.c_int64(), C.c_int64()
    ctx.call("qms_envelope_layout", C.byref(n_e), C.byref(n_s), None)
    slot_off = np.zeros(n_e.value + 1, dtype=np.int64)
    ctx.call("qms_envelope_layout", C.byref(n_e), C.byref(n_s), slot_off.ctypes.data_as(C.c_void_p))
    ptrs = [C.c_void_p() for _ in range(7)]
    ctx.call("qms_envelope_buffers", *[C.byref(p) for p in ptrs])
    ctx.call("qms_synchronize")
    env_ranges = [shard_range(n_env, r, world) for r in range(world)]
    slot_ranges = [(int(slot_off[b]), int(slot_off[e])) for b, e in env_ranges]
    # every rank must hold the same layout, or the collectives below would wait for each other forever
    import torch.distributed as dist
    digest = torch.tensor([n_e.value, n_s.value, int(np.bitwise_xor.reduce(slot_off * np.arange(1, len(slot_off) + 1)))],
                          dtype=torch.int64, device="cuda:%d" % ctx.device)
    lo, hi = digest.clone(), digest.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if not bool(torch.equal(lo, hi)):
        raise RuntimeError("ranks disagree on the envelope layout (different molecule lists / parameters per rank?): %s vs %s"
                           % (lo.tolist(), hi.tolist()))
    specs = [(ptrs[0], np.float64, n_s.value, slot_ranges), (ptrs[1], np.float64, n_s.value, slot_ranges),
             (ptrs[2], np.int32, n_s.value, slot_ranges), (ptrs[3], np.int32, n_s.value, slot_ranges),
             (ptrs[4], np.uint8, n_s.value, slot_ranges), (ptrs[5], np.int32, n_e.value, env_ranges),
             (ptrs[6], np.int32, n_e.value, env_ranges)]
    received = 0
    for pointer, dtype, total, ranges in specs:
        if total:
            received += allgather_ranges(_device_tensor(pointer.value, total, dtype, ctx.device), ranges)
    torch.cuda.synchronize(ctx.device)
    return received


def gather_packed_hits(out, spectrum_begin, world):
    """Gather per-rank packed hit dicts on rank 0 in rank (= spectrum) order; other ranks get None.

    ``out`` is the dict of ``IsotopologueLibrary.match_packed`` for this rank's spectrum shard whose first
    spectrum has global index ``spectrum_begin``.  Works with any backend (object gather of numpy arrays)."""
    d = _dist()
    nh, nw = out["n_hits"], out["n_windows"]
    local = {"spec": out["spec"][:nh] + np.int32(spectrum_begin), "entry": out["entry"][:nh], "score": out["score"][:nh],
             "scaling": out["scaling"][:nh], "win_len": np.diff(out["win_off"][:nh + 1]), "mmz": out["mmz"][:nw],
             "mi": out["mi"][:nw]}
    if d is None or world == 1:
        parts = [local]
    else:
        parts = [None] * world if d.get_rank() == 0 else None
        d.gather_object(local, parts, dst=0)
        if d.get_rank() != 0:
            return None
    merged = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    merged["win_off"] = np.concatenate([[0], np.cumsum(merged.pop("win_len"))]).astype(np.int64)
    merged["n_hits"], merged["n_windows"] = len(merged["spec"]), len(merged["mmz"])
    return merged
