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


def numa_report(device):
    """What the box says about the GPU's place in the host: the sysfs NUMA node of its PCIe function (-1 = the platform
    does not say, the usual answer of a single-socket VM), the NUMA nodes that are online and the CPUs this process may run
    on.  Never raises; ``bench.py`` prints it next to the copy rates."""
    import os
    report = {"gpu_node": None, "nodes_online": None, "cpus_allowed": None}
    try:
        report["cpus_allowed"] = len(os.sched_getaffinity(0))
        with open("/sys/devices/system/node/online") as fh:
            report["nodes_online"] = fh.read().strip()
    except Exception:
        pass
    try:
        import torch
        props = torch.cuda.get_device_properties(device)
        bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as fh:
            report["gpu_node"] = int(fh.read().strip())
    except Exception:
        pass
    return report


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


SHARD_MIN_ENVELOPES = 4_000_000


def shard_build_by_default(n_env, world):
    """Whether ``IsotopologueLibrary(distributed=None)`` shards the envelope build over the ranks.

    Measured on B200: the envelope kernel builds ~650 M peptide envelopes per second (0.55 ms for the 365 k envelopes of
    the 500 k-peptide library), an envelope record is ~400 bytes, i.e. the all-gather moves a shard about as fast as a
    rank computes it, and a communicator costs a fraction of a second to create -- below a few million envelopes the
    replicated build is the faster one on every rank (round 1: 1.26 s sharded vs 0.014 s replicated for 38 envelopes)."""
    return world > 1 and n_env >= SHARD_MIN_ENVELOPES


def _same_on_all_ranks(values, device):
    """Raise if the ranks disagree on ``values`` (the collectives of the build would wait for each other forever)."""
    import torch
    import torch.distributed as dist
    mine = torch.tensor(values, dtype=torch.int64, device=device)
    pieces = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(pieces, mine)
    if any(not bool(torch.equal(piece, mine)) for piece in pieces):
        raise RuntimeError("ranks disagree on the envelope layout (different molecule lists / parameters per rank?): %s"
                           % [piece.tolist() for piece in pieces])


def allgather_envelopes(ctx, n_env, world):
    """Assemble the envelope shards of all ranks in place.  Returns {"bytes_received", "ms", "how"}.

    Under NCCL this is the library's own collective: ``qms_comm_init`` (rank 0's id travels through a 128-byte
    broadcast) and ``qms_allgather_library`` -- one ``ncclAllGather`` of fixed-stride envelope records over NVLink
    (csrc/comm.cu).  Other backends (gloo in CPU tests has no device memory to gather) use padded torch all-gathers
    of the seven arrays."""
    import torch
    import torch.distributed as dist
    n_e, n_s = C.c_int64(), C.c_int64()
    ctx.call("qms_envelope_layout", C.byref(n_e), C.byref(n_s), None)
    slot_off = np.zeros(n_e.value + 1, dtype=np.int64)
    ctx.call("qms_envelope_layout", C.byref(n_e), C.byref(n_s), slot_off.ctypes.data_as(C.c_void_p))
    device = "cuda:%d" % ctx.device
    digest = [n_e.value, n_s.value, int(np.bitwise_xor.reduce(slot_off * np.arange(1, len(slot_off) + 1)))]
    if dist.get_backend() == "nccl":
        _same_on_all_ranks(digest, device)
        token = torch.zeros(128, dtype=torch.uint8, device=device)
        if dist.get_rank() == 0:
            raw = (C.c_ubyte * 128)()
            ctx.check(ctx.lib.qms_comm_unique_id(raw, 128))
            token = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
        dist.broadcast(token, src=0)
        raw = (C.c_ubyte * 128).from_buffer_copy(token.cpu().numpy().tobytes())
        ctx.call("qms_comm_init", raw, dist.get_rank(), world)
        received, ms = C.c_int64(), C.c_double()
        ctx.call("qms_allgather_library", C.byref(received), C.byref(ms))
        return {"bytes_received": received.value, "ms": ms.value, "how": "one ncclAllGather of fixed-stride records (qms_allgather_library)"}
    ptrs = [C.c_void_p() for _ in range(7)]
    ctx.call("qms_envelope_buffers", *[C.byref(p) for p in ptrs])
    ctx.call("qms_synchronize")
    env_ranges = [shard_range(n_env, r, world) for r in range(world)]
    slot_ranges = [(int(slot_off[b]), int(slot_off[e])) for b, e in env_ranges]
    _same_on_all_ranks(digest, device)
    specs = [(ptrs[0], np.float64, n_s.value, slot_ranges), (ptrs[1], np.float64, n_s.value, slot_ranges),
             (ptrs[2], np.int32, n_s.value, slot_ranges), (ptrs[3], np.int32, n_s.value, slot_ranges),
             (ptrs[4], np.uint8, n_s.value, slot_ranges), (ptrs[5], np.int32, n_e.value, env_ranges),
             (ptrs[6], np.int32, n_e.value, env_ranges)]
    received = 0
    for pointer, dtype, total, ranges in specs:
        if total:
            received += allgather_ranges(_device_tensor(pointer.value, total, dtype, ctx.device), ranges)
    torch.cuda.synchronize(ctx.device)
    return {"bytes_received": received, "ms": None, "how": "seven padded torch all-gathers"}


_HIT_FIELDS = (("spec", np.int32, "h"), ("entry", np.int32, "h"), ("score", np.float64, "h"), ("scaling", np.float64, "h"),
               ("win_off", np.int64, "h1"), ("peak", np.int64, "w"), ("peak32", np.int32, "w"), ("peak16", np.uint16, "w"),
               ("mmz", np.float64, "w"), ("mi", np.float64, "w"), ("spec_off", np.int64, "s1"), ("row_off", np.int64, "s1"))
# what a segment holds, by what the matching pass brings back from the device (IsotopologueLibrary._run_matcher):
# the matched values, or the int32 peak rows, or the compact columns (+ the shard's row offsets, which widen them)
_MODE_FIELDS = {"values": ("spec", "entry", "score", "scaling", "win_off", "peak", "mmz", "mi"),
                "rows32": ("spec", "entry", "score", "scaling", "win_off", "peak32"),
                "compact": ("entry", "score", "scaling", "peak16", "spec_off", "row_off")}
_MODES = ("rows32", "values", "compact")


def _mode_name(values):
    return values if isinstance(values, str) else ("values" if values else "rows32")


def _segment_layout(cap_h, cap_w, values, cap_s=0):
    """Byte offsets of the hit arrays inside a segment with room for cap_h hits, cap_w hit windows and cap_s spectra."""
    offset, layout = 64, {}
    wanted = _MODE_FIELDS[_mode_name(values)]
    for name, dtype, kind in _HIT_FIELDS:
        if name not in wanted:
            continue
        count = {"h": cap_h, "h1": cap_h + 1, "w": cap_w, "s1": cap_s + 1}[kind]
        layout[name] = (offset, np.dtype(dtype), count)
        offset = (offset + count * np.dtype(dtype).itemsize + 63) // 64 * 64
    return layout, offset


class SharedHitArena:
    """Where the hits of this rank land when the ranks of one box match shards of a run: a POSIX shared-memory segment
    per rank, page-locked in the owning process (the library copies the hits into it by DMA, sub-batch by sub-batch,
    while it matches) and mapped by rank 0.  After ``collect`` rank 0 holds the hits of all ranks, in rank = spectrum
    order, as one packed batch per rank -- no pickling, no copy through a pipe, and every GPU uses its own PCIe link
    (the north star's "results are gathered to host in spectrum order").

        arena = SharedHitArena()
        lib.use_hit_arena(arena)
        res = lib.match_many(shard_peaks, offsets=shard_offsets, ...)
        batches = arena.collect(res.packed_batches()[-1], spectrum_begin)     # rank 0: [PackedHits per rank]; else None

    The arrays of a collected batch are views of the segments: they stay valid until the owning rank matches again."""

    def __init__(self, tag=None, register=True):
        import os
        self.rank, self.world = rank(), world_size()
        self.tag = tag or "qms_%d_%s" % (os.getppid(), os.environ.get("MASTER_PORT", "0"))
        self.register = register
        self.generation = 0
        self._own = None           # (name, mmap, bytes, cap_h, cap_w, mode, cap_s)
        self.entry_windows = None  # windows per index entry (IsotopologueLibrary.use_hit_arena): widens compact columns
        self._peers = {}           # rank 0: (rank, generation) -> mmap
        self.stats = {}

    # ---- segment of this rank ------------------------------------------------------------------------------
    def _name(self, r, generation):
        return "%s_r%d_g%d" % (self.tag, r, generation)

    def _drop_own(self):
        import os
        if self._own is None:
            return
        name, mapping, nbytes, *_ = self._own
        if self.register:
            from . import _capi
            address = C.addressof(C.c_char.from_buffer(mapping))
            _capi.load().qms_host_unregister(address)
        try:
            mapping.close()
        except BufferError:
            pass                    # arrays handed out earlier still view it; the pages go with them
        try:
            os.unlink("/dev/shm/" + name)
        except OSError:
            pass
        self._own = None

    def arrays(self, cap_h, cap_w, values, cap_s=0):
        """numpy views (capacity cap_h hits / cap_w windows / cap_s spectra) inside this rank's segment, growing it when
        needed.  ``values``: True (with the matched values), False (int32 peak rows) or "compact" (see _MODE_FIELDS)."""
        import mmap
        import os
        own = self._own
        values = _mode_name(values)
        if own is None or own[3] < cap_h or own[4] < cap_w or own[5] != values or own[6] < cap_s:
            self._drop_own()
            if own is not None:
                cap_h, cap_w, cap_s = max(cap_h, own[3]), max(cap_w, own[4]), max(cap_s, own[6])
            cap_h, cap_w = cap_h + cap_h // 4 + 64, cap_w + cap_w // 4 + 512
            layout, nbytes = _segment_layout(cap_h, cap_w, values, cap_s)
            room = os.statvfs("/dev/shm")
            if room.f_bavail * room.f_frsize < nbytes + (16 << 20):
                raise OSError("/dev/shm has %d bytes free, the hit segment of rank %d needs %d (use the page-locked pool instead: "
                              "lib.use_hit_arena(None))" % (room.f_bavail * room.f_frsize, self.rank, nbytes))
            self.generation += 1
            name = self._name(self.rank, self.generation)
            fd = os.open("/dev/shm/" + name, os.O_CREAT | os.O_RDWR | os.O_TRUNC, 0o600)
            try:
                os.ftruncate(fd, nbytes)
                mapping = mmap.mmap(fd, nbytes)
            finally:
                os.close(fd)
            if self.register:
                from . import _capi
                lib = _capi.load()
                address = C.addressof(C.c_char.from_buffer(mapping))
                rc = lib.qms_host_register(address, nbytes)
                if rc != _capi.QMS_OK:
                    raise _capi.QmsError(rc, (lib.qms_last_error(None) or b"").decode())
            self._own = (name, mapping, nbytes, cap_h, cap_w, values, cap_s)
        name, mapping, nbytes, cap_h, cap_w, values, cap_s = self._own
        layout, _ = _segment_layout(cap_h, cap_w, values, cap_s)
        return {field: np.frombuffer(mapping, dtype=dtype, count=count, offset=offset) for field, (offset, dtype, count) in layout.items()
                if field != "row_off"}

    # ---- rank 0: the hits of all ranks ---------------------------------------------------------------------------
    def collect(self, out, spectrum_begin=0):
        """All ranks call it after matching their shard; rank 0 gets one packed batch per rank, in rank order
        (``batch["spec"]`` counts from the first spectrum of the rank's shard, ``batch["spec_base"]`` is that
        spectrum's number in the run), the other ranks get None."""
        import mmap
        import os
        import time
        from .isotopologue_library import PackedHits
        t0 = time.perf_counter()
        own = self._own
        have = dict.keys(out)
        mode = "values" if "mmz" in have else ("compact" if "peak16" in have else "rows32")
        n_spec = len(out["spec_off"]) - 1 if mode == "compact" else 0
        probe = dict.__getitem__(out, "entry")
        in_segment = own is not None and own[5] == mode and out["n_hits"] > 0 and \
            C.addressof(C.c_char.from_buffer(own[1])) <= probe.ctypes.data < C.addressof(C.c_char.from_buffer(own[1])) + own[2]
        if own is None or not in_segment:                    # matched before the arena was attached, or nothing to show
            fresh = self.arrays(max(1, out["n_hits"]), max(1, out["n_windows"]), mode, n_spec)
            for field, array in fresh.items():
                source = out[field]
                array[:len(source)] = source
            own = self._own
        if mode == "compact":                                # the shard's row offsets, counted from its first row
            layout, _ = _segment_layout(own[3], own[4], mode, own[6])
            offset, dtype, count = layout["row_off"]
            rows = np.frombuffer(own[1], dtype=dtype, count=n_spec + 1, offset=offset)
            rows[:] = np.asarray(out._offsets[:n_spec + 1], dtype=np.int64) - int(out._offsets[0])
        meta = [out["n_hits"], out["n_windows"], own[3], own[4], self.generation, _MODES.index(own[5]), int(spectrum_begin), n_spec, own[6]]
        d = _dist()
        if d is None or self.world == 1:
            gathered = [meta]
        else:
            import torch
            on_gpu = d.get_backend() == "nccl"
            mine = torch.tensor(meta, dtype=torch.int64, device="cuda" if on_gpu else "cpu")
            pieces = [torch.empty_like(mine) for _ in range(self.world)]
            d.all_gather(pieces, mine)                       # 72 bytes per rank: the only collective of the matching path
            gathered = [piece.tolist() for piece in pieces]
        if self.rank != 0:
            return None
        batches = []
        for r, (n_hits, n_windows, cap_h, cap_w, generation, mode, begin, n_spec, cap_s) in enumerate(gathered):
            if r == self.rank:
                mapping = own[1]
            else:
                mapping = self._peers.get((r, generation))
                if mapping is None:
                    for key in [k for k in self._peers if k[0] == r]:
                        try:
                            self._peers.pop(key).close()
                        except BufferError:
                            pass
                    fd = os.open("/dev/shm/" + self._name(r, generation), os.O_RDWR)
                    try:
                        mapping = mmap.mmap(fd, os.fstat(fd).st_size)
                    finally:
                        os.close(fd)
                    self._peers[(r, generation)] = mapping
            layout, _ = _segment_layout(cap_h, cap_w, _MODES[mode], cap_s)
            views = {}
            for field, (offset, dtype, count) in layout.items():
                used = {"spec": n_hits, "entry": n_hits, "score": n_hits, "scaling": n_hits, "win_off": n_hits + 1,
                        "spec_off": n_spec + 1, "row_off": n_spec + 1}.get(field, n_windows)
                views[field] = np.frombuffer(mapping, dtype=dtype, count=used, offset=offset)
            batch = PackedHits(None, views.pop("row_off", None), self.entry_windows)
            batch.update(views)
            batch["n_hits"], batch["n_windows"], batch["spec_base"], batch["rank"] = n_hits, n_windows, begin, r
            batches.append(batch)
        self.stats = {"collect_ms": 1e3 * (time.perf_counter() - t0), "hits": [int(g[0]) for g in gathered],
                      "bytes_per_rank": [int(g[0]) * (20 if g[5] == 2 else 32) + int(g[1]) * (4, 24, 2)[g[5]] for g in gathered]}
        return batches

    def close(self):
        self._drop_own()
        for mapping in self._peers.values():
            try:
                mapping.close()
            except BufferError:
                pass
        self._peers = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gather_hits_to_rank0(out, spectrum_begin, arena, stats=None):
    """``arena.collect`` plus bookkeeping for bench.py: rank 0 returns the list of per-rank batches."""
    batches = arena.collect(out, spectrum_begin)
    if stats is not None:
        stats["own"] = out
        if batches is not None:
            stats["summary"] = dict(arena.stats, ranks=len(batches), total_hits=int(sum(b["n_hits"] for b in batches)))
    return batches if batches is not None else out


def gather_packed_hits(out, spectrum_begin, world, row_begin=0):
    """Gather per-rank packed hit dicts on rank 0 in rank (= spectrum) order; other ranks get None.

    ``out`` is the dict of ``IsotopologueLibrary.match_packed`` for this rank's spectrum shard whose first
    spectrum has global index ``spectrum_begin`` and whose first peak row has global index ``row_begin``.  Works with any
    backend (object gather of numpy arrays: fine for small results and CPU tests; ``SharedHitArena`` is the way for runs)."""
    d = _dist()
    nh, nw = out["n_hits"], out["n_windows"]
    rows = np.asarray(out["peak"][:nw], dtype=np.int64)
    local = {"spec": out["spec"][:nh] + np.int32(spectrum_begin), "entry": out["entry"][:nh], "score": out["score"][:nh],
             "scaling": out["scaling"][:nh], "win_len": np.diff(out["win_off"][:nh + 1]), "mmz": out["mmz"][:nw],
             "mi": out["mi"][:nw], "peak": np.where(rows >= 0, rows + np.int64(row_begin), -1)}
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
