#!/usr/bin/env python
"""Benchmark of the matching hot path (spectra/s through match_all) on B200 -- see DESIGN.md section 6.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload proteome|bsa|bsa_pct]

Default workload = BASELINE.json configs[3], the north-star configuration: a 500 000-peptide tryptic proteome
library (charges 1-5, carbamidomethyl fixed label; 1.59 M index entries, 10.3 M windows) against the 200 000-spectrum
synthetic run, which is made of 8 shards of 25 000 spectra (seed 20260104 + shard).  One *step* = one pass of the
matcher over ONE shard per GPU: N=1 matches shard 0, N GPUs match shards 0..N-1 (rank r owns shard r, library
replicated, no data-path collective; N=8 is the whole named run) -- weak scaling.
Prints ONE JSON line (rank 0).  `value` = spectra/s with the shard resident in HBM, device-timed (CUDA events on the
library's stream, max over ranks); `e2e` = the same through the public host API (`match_many`) with pinned HOST
buffers, H2D of the spectra and D2H of the packed hits inside the timed region and, for N>1, the hits of all ranks
gathered on rank 0 in spectrum order; `roofline` = the kernel that dominates the step against the measured HBM peak
with the algorithmic bytes of SURVEY.md 8d; `cpu_baseline` = the reference's CPU path on a bounded sample.
`--impl reference` times the UNMODIFIED reference (baseline/_ref, pure Python) on all host cores, one process per core
over disjoint spectra of the same shard, against a 5 000-peptide sample of the same library (the reference needs
~20 min and ~40 GB per process to build the full one) and extrapolates by its measured per-bin cost.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from pyqms_b200.synthetic import BSA_MOLECULES, CAM_FIXED_LABEL, random_tryptic_peptides, synth_run  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[3] / SURVEY.md 8d "C4": the north-star configuration
    "proteome": dict(lib=dict(charges=[1, 2, 3, 4, 5], fixed_labels=CAM_FIXED_LABEL),
                     run=dict(n_spectra=25000, seed=20260104, entry_fraction=0.02), shards=8,
                     name="human-scale tryptic proteome (500k peptides x z1-5, CAM) vs the 200k-spectrum run, one 25k-spectrum shard per GPU"),
    # configs[1] "C2"
    "bsa": dict(lib=dict(molecules=BSA_MOLECULES, charges=[1, 2, 3, 4], metabolic_labels={"15N": [0, 0.99]},
                         fixed_labels=CAM_FIXED_LABEL),
                run=dict(n_spectra=30000, seed=20260101), shards=1, name="BSA 19 molecules 14N+15N z1-4 CAM vs 30k-spectrum synthetic run"),
    # configs[2] "C3"
    "bsa_pct": dict(lib=dict(molecules=BSA_MOLECULES, charges=[1, 2, 3, 4], fixed_labels=CAM_FIXED_LABEL,
                             metabolic_labels={"15N": [round(0.01 * i, 2) for i in range(100)]}),
                    run=dict(n_spectra=30000, seed=20260101), shards=1, name="BSA x 100 15N percentiles vs 30k-spectrum run"),
}
REFERENCE_SAMPLE_PEPTIDES = 5000       # library sample the CPU arms can build
FIXTURE = ROOT / "bench_data" / "proteome500k_shard0_planted.npz"


def static_config(args):
    """`config` of the JSON line: a function of the command line only, so both arms print the same dict."""
    wl = WORKLOADS[args.workload]
    cfg = {"workload": wl["name"], "spectra_per_gpu": wl["run"]["n_spectra"], "run_seed": wl["run"]["seed"],
           "input_kind": "numpy (N,2) float64", "l2_policy": "inputs (hundreds of MB of peaks per step) larger than L2"}
    if args.workload == "proteome":
        cfg["library_peptides"] = args.proteome_size
    return cfg


def library_kwargs(args, n_peptides=None):
    kw = dict(WORKLOADS[args.workload]["lib"])
    if args.workload == "proteome":
        peptides = random_tryptic_peptides(args.proteome_size, seed=20260104)
        kw["molecules"] = peptides if n_peptides is None else peptides[:n_peptides]
    return kw


def load_fixture(args):
    """c-peak table of the entries planted into shard 0 (written by tools/make_bench_fixture.py from the full library):
    lets a process without the library (the reference arm) draw the identical spectra."""
    if args.workload != "proteome" or args.proteome_size != 500000 or not FIXTURE.exists():
        return None
    z = np.load(FIXTURE)
    return {"n_entries": int(z["n_entries"]), "lens": z["lens"].astype(np.int64), "mz": z["mz"], "ci": z["ci"].astype(np.float64)}


def make_shard(args, shard, entry_peaks=None):
    """(peaks, offsets, how) of one shard of the run.  Shard 0 comes from the committed planted table when it matches
    the library; other shards (and other workloads) plant from the library's own entries."""
    run = dict(WORKLOADS[args.workload]["run"])
    run["seed"] = run["seed"] + shard
    table = load_fixture(args) if shard == 0 else None
    if table is not None and (entry_peaks is None or table["n_entries"] == entry_peaks[2]):
        peaks, offsets = synth_run(None, None, planted=table, **run)
        return peaks, offsets, "planted table bench_data/%s" % FIXTURE.name
    if entry_peaks is None:
        return None, None, "no planted table"
    mz, ci = entry_peaks[0]()
    peaks, offsets = synth_run(mz, ci, **run)
    return peaks, offsets, "planted from the library's entries"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.stop = threading.Event()
        self.thread = threading.Thread(target=self._loop, daemon=True)

    def _loop(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *exc):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------- CPU legs
def cpu_library(args, n_peptides):
    """(library, kind): the unmodified reference from baseline/_ref when it is there, else the oracle port."""
    sys.setrecursionlimit(20000)
    kw = library_kwargs(args, n_peptides)
    ref_dir = ROOT / "baseline" / "_ref"
    if (ref_dir / "pyqms" / "__init__.py").exists() and not args.cpu_port:
        import warnings
        warnings.filterwarnings("ignore")
        if str(ref_dir) not in sys.path:
            sys.path.insert(0, str(ref_dir))
        import pyqms                                       # the reference itself (+ the parser stand-in)
        return pyqms.IsotopologueLibrary(verbose=False, **kw), "reference"
    from oracle.qms_oracle import OracleLibrary
    return OracleLibrary(**kw), "port"


def cpu_match(lib, kind, spectrum, spec_id, results):
    if kind == "reference":
        return lib.match_all(mz_i_list=spectrum, file_name="run", spec_id=spec_id, spec_rt=spec_id * 0.2, results=results)
    return lib.match_all(spectrum, "run", spec_id, spec_id * 0.2, results)


def n_bins(lib, kind):
    return len(lib.match_sets)


def full_library_bins(args, entries_full=None):
    """Match bins of the full library (entries / MAX_MOLECULES_PER_MATCH_BIN): the reference's per-spectrum cost is
    linear in them (it re-slices and re-hashes the spectrum per bin, IL:1835-1839)."""
    if entries_full is None:
        table = load_fixture(args)
        entries_full = table["n_entries"] if table is not None else None
    return None if entries_full is None else (entries_full + 19) // 20


_WORKER = {}


def _cpu_worker_run(spectra):
    lib, kind = _WORKER["lib"], _WORKER["kind"]
    t0 = time.perf_counter()
    res = None if kind == "reference" else {}
    for s, spec in spectra:
        res = cpu_match(lib, kind, spec, s, res)
    hits = sum(len(v["data"]) if kind == "reference" else len(v) for v in res.values()) if res is not None else 0
    return time.perf_counter() - t0, hits


def cpu_time_sample(lib, kind, peaks, offsets, budget_s, max_spectra):
    """Time the CPU path on leading spectra of the shard until ~budget_s elapsed (single core)."""
    res, n, t0 = (None if kind == "reference" else {}), 0, time.perf_counter()
    while n < min(max_spectra, len(offsets) - 1):
        res = cpu_match(lib, kind, peaks[offsets[n]:offsets[n + 1]], n, res)
        n += 1
        if n >= 20 and time.perf_counter() - t0 > budget_s:
            break
    hits = sum(len(v["data"]) if kind == "reference" else len(v) for v in res.values()) if res is not None else 0
    return n, time.perf_counter() - t0, hits


def run_reference_arm(args):
    """--impl reference: the reference's own CPU code on all host cores, one process per core over disjoint spectra."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = min(max(1, (os.cpu_count() or 1)), 64)
    sample = REFERENCE_SAMPLE_PEPTIDES if args.workload == "proteome" else None
    t0 = time.perf_counter()
    lib, kind = cpu_library(args, sample)
    build_s = time.perf_counter() - t0
    peaks, offsets, how = make_shard(args, 0)
    if peaks is None:                                     # no planted table: plant from the sample library's own entries
        from tests.golden.spectra import entry_peaks as ref_entry_peaks
        mz, ci = ref_entry_peaks(lib)
        run = dict(WORKLOADS[args.workload]["run"])
        run["n_spectra"] = min(run["n_spectra"], args.ref_spectra_per_core * cores)
        peaks, offsets = synth_run(mz, ci, **run)
        how = "planted from the sample library's entries (no planted table for this configuration)"
    n = min(len(offsets) - 1, args.ref_spectra_per_core * cores)
    shards = [[(s, peaks[offsets[s]:offsets[s + 1]]) for s in range(i, n, cores)] for i in range(cores)]
    _WORKER["lib"], _WORKER["kind"] = lib, kind           # inherited by the forked workers (no rebuild per process)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        times = []
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pool.map(_cpu_worker_run, shards, chunksize=1)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
    total = sum(times)
    measured = n * len(times) / total
    bins_sample = n_bins(lib, kind)
    bins_full = full_library_bins(args) if sample is not None else bins_sample
    scale = bins_sample / bins_full if bins_full else 1.0
    value = measured * scale
    line = {
        "impl": "reference", "metric": "spectra/sec through match_all", "value": value, "unit": "spectra/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": static_config(args),
        "cpu_baseline": {"value": value, "unit": "spectra/s", "cores": cores, "kind": kind,
                         "sample": "%d spectra of shard 0 per step (%s), %d processes x %d spectra, against %s; measured %.1f spectra/s "
                                   "on %d match bins%s" % (
                                       n, how, cores, len(shards[0]),
                                       "the first %d peptides of the library" % sample if sample else "the full library",
                                       measured, bins_sample,
                                       ", scaled by %d/%d bins to the full library (the reference's cost per spectrum is linear in "
                                       "the bins, BASELINE.md section 3)" % (bins_sample, bins_full) if scale != 1.0 else ""),
                         "measured_on_sample": measured, "bins_sample": bins_sample, "bins_full": bins_full,
                         "library_build_s": build_s},
        "e2e": {"value": value, "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------- our arm
def device_buffers(torch, QmsHits, cap_h, cap_w):
    dev = {"spec": torch.empty(cap_h, dtype=torch.int32, device="cuda"), "entry": torch.empty(cap_h, dtype=torch.int32, device="cuda"),
           "score": torch.empty(cap_h, dtype=torch.float64, device="cuda"), "scaling": torch.empty(cap_h, dtype=torch.float64, device="cuda"),
           "win_off": torch.empty(cap_h + 1, dtype=torch.int64, device="cuda"), "mmz": torch.empty(cap_w, dtype=torch.float64, device="cuda"),
           "mi": torch.empty(cap_w, dtype=torch.float64, device="cuda"), "peak": torch.empty(cap_w, dtype=torch.int64, device="cuda")}
    return dev, QmsHits(cap_h, cap_w, *[dev[k].data_ptr() for k in ("spec", "entry", "score", "scaling", "win_off", "mmz", "mi", "peak")])


def sparse_probe_figure(pyqms_b200, torch, local, peak_gbs, n_spectra=30000, steps=10):
    """The HBM-bound regime (configs[1], BSA library: 99 % of the peaks are rejected by the cover bitmap): roofline of the
    streaming probe kernel, reported beside the headline as `roofline_sparse`."""
    import ctypes as C
    from pyqms_b200._capi import QmsHits
    wl = WORKLOADS["bsa"]
    lib = pyqms_b200.IsotopologueLibrary(verbose=False, device=local, distributed=False, **wl["lib"])
    mz, ci = lib.entry_peaks()
    peaks, offsets = synth_run(mz, ci, n_spectra=n_spectra, seed=wl["run"]["seed"])
    ctx, xi = lib._ctx, lib.params["MZ_SCORE_PERCENTILE"]
    d_peaks = torch.from_numpy(peaks).cuda()
    dev, hits = device_buffers(torch, QmsHits, 1 << 16, 1 << 19)
    nh, nw = C.c_int64(), C.c_int64()
    for it in range(3 + steps):
        if it == 3:
            ctx.reset_counters()
        ctx.check(ctx.lib.qms_match_batch(ctx.handle, d_peaks.data_ptr(), offsets.ctypes.data, len(offsets) - 1, 0, xi, C.byref(hits),
                                          C.byref(nh), C.byref(nw)))
    c = ctx.counters()
    alg = 16.0 * c["peaks"] / steps
    ms = c["ms_probe"] / steps
    achieved = alg / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    step_ms = (c["ms_probe"] + c["ms_gate"] + c["ms_score"] + c["ms_compact"]) / steps
    return {"workload": "%s, first %d spectra" % (wl["name"], n_spectra), "bound": "hbm", "kernel": "spectrum_probe_warp_kernel",
            "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
            "algorithmic_bytes_per_launch": alg, "kernel_ms": ms, "kernels_ms_per_step": step_ms,
            "spectra_per_s_device_kernels": (len(offsets) - 1) / (step_ms * 1e-3) if step_ms > 0 else None}


def run_ours(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import pyqms_b200
    from pyqms_b200._capi import QmsHits

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    from pyqms_b200.distributed import bind_to_gpu_numa_node, numa_report
    numa_node = bind_to_gpu_numa_node(local)               # before any pinned allocation
    numa_facts = numa_report(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=600))
    wl = WORKLOADS[args.workload]
    lib_kwargs = library_kwargs(args)                      # (drawing 500 000 synthetic peptides is not part of the build)
    t0 = time.perf_counter()
    lib = pyqms_b200.IsotopologueLibrary(verbose=False, device=local, **lib_kwargs)     # sharded build + all-gather when world > 1
    build_s = time.perf_counter() - t0
    build_counters = lib.device_counters()
    shard = rank % wl["shards"]
    t0 = time.perf_counter()
    peaks, offsets, run_how = make_shard(args, shard, (lib.entry_peaks, None, int(lib._n_entries)))
    synth_s = time.perf_counter() - t0
    n_spec = len(offsets) - 1
    ctx = lib._ctx
    xi = lib.params["MZ_SCORE_PERCENTILE"]
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    # ---- resident inputs + device output buffers
    d_peaks = torch.from_numpy(peaks).cuda()
    cap_h, cap_w = 1 << 16, 1 << 19
    while True:
        dev, hits = device_buffers(torch, QmsHits, cap_h, cap_w)
        nh, nw = C.c_int64(), C.c_int64()
        rc = ctx.lib.qms_match_batch(ctx.handle, d_peaks.data_ptr(), offsets.ctypes.data, n_spec, 0, xi, C.byref(hits), C.byref(nh),
                                     C.byref(nw))
        if rc == -3:
            cap_h, cap_w = nh.value + 1024, nw.value + 8192
            del dev
            continue
        ctx.check(rc)
        break

    def device_step():
        # peaks resident in HBM; the row offsets stay a host array (the library needs the row count on the host)
        ctx.check(ctx.lib.qms_match_batch(ctx.handle, d_peaks.data_ptr(), offsets.ctypes.data, n_spec, 0, xi, C.byref(hits),
                                          C.byref(nh), C.byref(nw)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        device_step()
    ctx.reset_counters()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local) as clocks:
        start.record(stream)
        for _ in range(args.steps):
            device_step()
        stop.record(stream)
        barrier()
        # keep sampling long enough for at least a few nvidia-smi samples under load on short runs
        extra_until = time.perf_counter() + (0.0 if len(clocks.rows) >= 3 else 0.8)
        while time.perf_counter() < extra_until:
            device_step()
        torch.cuda.synchronize()
    ms = start.elapsed_time(stop)
    counters = ctx.counters()
    steps_counted = counters["spectra"] / n_spec
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n_spec * args.steps / (ms_max / 1e3)
    device_hits = int(nh.value)
    del dev, d_peaks
    torch.cuda.empty_cache()

    # ---- e2e through the public host API: pinned host spectra -> packed hits on the host (rank 0 holds all ranks' hits)
    pin_peaks = torch.from_numpy(peaks).pin_memory()
    pin_offsets = torch.from_numpy(offsets).pin_memory()
    h_peaks, h_offsets = pin_peaks.numpy(), pin_offsets.numpy()
    spec_ids = list(range(n_spec))
    gather_stats = {}
    arena = None
    if world > 1 and not os.environ.get("QMS_BENCH_NO_ARENA"):
        # hits of every rank land, by DMA, in a page-locked shared-memory segment that rank 0 maps (no data-path collective);
        # the ranks agree first that /dev/shm has room for all segments (a container may come with 64 MB of it)
        from pyqms_b200.distributed import SharedHitArena, gather_hits_to_rank0
        need = world * (80 * max(device_hits, 1) + (64 << 20))            # ~58 bytes per hit of this workload + growth head room
        try:
            room = os.statvfs("/dev/shm")
            enough = room.f_bavail * room.f_frsize > need
        except OSError:
            enough = False
        agreed = torch.tensor([1 if enough else 0], dtype=torch.int64, device="cuda")
        dist.all_reduce(agreed, op=dist.ReduceOp.MIN)
        if int(agreed.item()):
            arena = SharedHitArena()
            lib.use_hit_arena(arena)
        else:
            gather_stats["summary"] = {"unavailable": "/dev/shm has no room for %d bytes of hit segments; every rank keeps its hits" % need}

    def e2e_step():
        # the call a user makes for a whole run: host (N,2) array in, pyqms Results out (hits stay packed until read)
        res = lib.match_many(h_peaks, file_name="run", spec_ids=spec_ids, spec_rts=spec_ids, offsets=h_offsets)
        out = res.packed_batches()[-1]
        if arena is not None:
            out = gather_hits_to_rank0(out, spectrum_begin=rank * n_spec, arena=arena, stats=gather_stats)
        return out

    for _ in range(max(1, min(args.warmup, 3))):
        out = e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_each = []
    for _ in range(args.steps):
        t1 = time.perf_counter()
        out = e2e_step()
        e2e_each.append(time.perf_counter() - t1)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n_spec * args.steps / float(t.item())
    h2d = int(peaks.nbytes + offsets.nbytes)
    own = out if arena is None else gather_stats.get("own", out)
    have = dict.keys(own)                         # the arrays the device copied out (columns widened on the host do not count)
    shipped = ("entry", "score", "scaling", "peak16", "spec_off") if "peak16" in have else \
        ("spec", "entry", "score", "scaling", "win_off", "peak32") if "peak32" in have else \
        ("spec", "entry", "score", "scaling", "win_off", "peak", "mmz", "mi")
    d2h = int(sum(dict.__getitem__(own, k).nbytes for k in shipped))
    e2e_counters = ctx.counters()
    e2e_calls = max(args.steps + max(1, min(args.warmup, 3)), 1)
    ms_h2d_step = (e2e_counters["ms_h2d"] - counters["ms_h2d"]) / e2e_calls
    ms_d2h_step = (e2e_counters["ms_d2h"] - counters["ms_d2h"]) / e2e_calls

    # ---- roofline of the kernel that dominates the step (SURVEY.md 8d algorithmic bytes)
    peaks_json = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    peak_gbs = float(peaks_json.get("hbm_gbs", 6650.0))
    per_step = {k: counters[k] / max(steps_counted, 1) for k in counters}
    launches_per_step = per_step["kernel_launches"]
    dense = counters["dense_batches"] > 0
    step_alg_bytes = (16.0 * per_step["peaks"] + 16.0 * per_step["candidates"] + 28.0 * per_step["candidate_windows"]
                      + 24.0 * per_step["hits"] + 16.0 * per_step["hit_windows"])
    if dense:
        # match_spectrum_kernel reads every peak row once, every candidate's entry header and window rows
        kernel, kernel_ms = "match_spectrum_kernel", per_step["ms_gate"]
        alg_bytes = 16.0 * per_step["peaks"] + 16.0 * per_step["candidates"] + 28.0 * per_step["candidate_windows"]
    else:
        kernel, kernel_ms = "spectrum_probe_warp_kernel", per_step["ms_probe"]
        alg_bytes = 16.0 * per_step["peaks"]
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                "frac": achieved / peak_gbs, "traffic": None, "peak_source": "MEASURED_PEAKS.json" if peaks_json else "fallback",
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_ms, "kernel_share_of_step": kernel_ms / (ms_max / args.steps),
                "step_algorithmic_bytes": step_alg_bytes, "step_achieved_gbs": step_alg_bytes / (ms_max / args.steps * 1e-3) / 1e9,
                "step_ms_breakdown": {k: per_step[k] for k in ("ms_probe", "ms_gate", "ms_score", "ms_compact", "ms_d2h", "ms_h2d")},
                "note": ("not HBM-bound: the kernel gates ~100 (peak, window) incidences per peak row from L2-resident library tables "
                         "and shared-memory spectrum tables; see the ncu figures" if dense else None)}
    ncu_file = ROOT / "profiles" / "kernel_ncu.json"
    if ncu_file.exists():
        try:
            facts = json.loads(ncu_file.read_text()).get(kernel)
            if facts:
                roofline["traffic"] = facts.get("dram_bytes_per_launch")
                roofline["ncu"] = facts
        except Exception:
            pass

    # ---- N > 1: the same library once more with the envelope build sharded over the ranks and assembled by one
    # ncclAllGather (qms_allgather_library) -- not the default for a library of this size, reported for comparison
    sharded_variant = None
    if world > 1 and not args.no_extras:
        del out
        t0 = time.perf_counter()
        sharded = pyqms_b200.IsotopologueLibrary(verbose=False, device=local, distributed=True, **lib_kwargs)
        sharded_variant = dict(sharded._build_stats, seconds_wall=time.perf_counter() - t0,
                               ms_envelopes_this_rank=sharded.device_counters()["ms_envelopes"],
                               identical_index=bool(sharded._n_entries == lib._n_entries and sharded._n_windows == lib._n_windows))
        del sharded

    line = None
    if rank == 0:
        extras = {}
        if not args.no_extras:
            try:
                extras["roofline_sparse"] = sparse_probe_figure(pyqms_b200, torch, local, peak_gbs)
            except Exception as exc:                       # the headline must not die with an extra
                extras["roofline_sparse"] = {"error": repr(exc)}
            extras["build"] = build_metric(args, pyqms_b200, local)
        # ---- CPU baseline: the reference (or the port) on a bounded sample of the same shard, one core
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sample = REFERENCE_SAMPLE_PEPTIDES if args.workload == "proteome" else None
            t0 = time.perf_counter()
            cpu_lib, kind = cpu_library(args, sample)
            cpu_build = time.perf_counter() - t0
            n, secs, nhits = cpu_time_sample(cpu_lib, kind, peaks, offsets, args.cpu_budget, 20000)
            bins_sample = n_bins(cpu_lib, kind)
            bins_full = (int(lib._n_entries) + 19) // 20
            scale = bins_sample / bins_full if sample else 1.0
            cpu = {"value": n / secs * scale, "unit": "spectra/s", "cores": 1, "kind": kind,
                   "sample": "first %d spectra of the same shard in %.1f s (%d hits) against %s: %.2f spectra/s on %d match bins%s" % (
                       n, secs, nhits, "the first %d peptides of the library" % sample if sample else "the full library", n / secs,
                       bins_sample, ", scaled by %d/%d bins to the full library (cost per spectrum linear in the bins, "
                       "BASELINE.md section 3)" % (bins_sample, bins_full) if sample else ""),
                   "measured_on_sample": n / secs, "library_build_s": cpu_build}
        line = {
            "metric": "spectra/sec through match_all", "value": value, "unit": "spectra/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": static_config(args),
            "e2e": {"value": e2e_value, "unit": "spectra/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * float(t.item()) / args.steps, "gather": gather_stats.get("summary"),
                    "ms_h2d_per_step": ms_h2d_step, "ms_d2h_per_step": ms_d2h_step,
                    "h2d_gbs_per_gpu": h2d / ms_h2d_step / 1e6 if ms_h2d_step > 0 else None,
                    "d2h_gbs_per_gpu": d2h / ms_d2h_step / 1e6 if ms_d2h_step > 0 else None,
                    "sub_batches_per_step": (e2e_counters["dense_batches"] - counters["dense_batches"]) / e2e_calls,
                    "kernel_ms_per_step": sum(e2e_counters[k] - counters[k] for k in ("ms_probe", "ms_gate", "ms_score", "ms_compact")) / e2e_calls,
                    "ms_per_step_each": [round(1e3 * x, 1) for x in e2e_each], "hits_in_shared_memory": arena is not None,
                    "note": "copies of neighbouring sub-batches overlap the kernels (copy streams); the GB/s are per GPU with all "
                            "GPUs of the run copying at once: the host's memory system is shared"},
            "gpu_launches": int(round(launches_per_step * args.steps)),
            "clocks": clocks.summary(), "roofline": roofline, "cpu_baseline": cpu,
            "workload_stats": {"shard": shard, "spectra": n_spec, "peaks": int(len(peaks)), "run_source": run_how, "synth_s": synth_s,
                               "library_entries": int(lib._n_entries), "library_windows": int(lib._n_windows),
                               "hits_per_step": device_hits, "numa_node_rank0": numa_node, "numa_rank0": numa_facts, "matcher": "spectrum-resident" if dense else "row-streaming"},
            "library_build": {"seconds_wall": build_s, "envelopes": int(build_counters["envelopes"]),
                              "ms_trees": build_counters["ms_trees"], "ms_envelopes": build_counters["ms_envelopes"],
                              "ms_index": build_counters["ms_index"],
                              "envelopes_per_s_device": (build_counters["envelopes"] / (build_counters["ms_envelopes"] * 1e-3)
                                                         if build_counters["ms_envelopes"] > 0 else None),
                              "envelopes_per_s_wall": build_counters["envelopes"] / build_s,
                              "distributed": getattr(lib, "_build_stats", None), "sharded_variant": sharded_variant},
            "counters_per_step": {k: per_step[k] for k in ("peaks", "live_peaks", "incidences", "candidates", "candidate_windows",
                                                            "combinations", "hits", "hit_windows")},
        }
        line.update(extras)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def build_metric(args, pyqms_b200, local):
    """Second half of the metric: isotopologue envelopes built per second (BASELINE.json configs[4] shape, bounded)."""
    from pyqms_b200.synthetic import averagine_formulas
    formulas = averagine_formulas(args.build_formulas, seed=20260105)
    # one untimed warm-up build (first-touch allocations of the 60 kDa cell table, lazy module loading), then timed builds
    walls, runs, sweep = [], [], None
    for _ in range(1 + max(1, args.build_repeats)):
        del sweep
        t0 = time.perf_counter()
        sweep = pyqms_b200.IsotopologueLibrary(molecules=formulas, charges=[1], verbose=False, device=local, distributed=False,
                                               params={"UPPER_MZ_LIMIT": 60000, "LOWER_MZ_LIMIT": 50})
        walls.append(time.perf_counter() - t0)
        runs.append((walls[-1], sweep.device_counters()))
    wall, c = sorted(runs[1:], key=lambda r: r[0])[(len(runs) - 1) // 2]      # the median build and its stage times
    out = {"metric": "isotopologue envelopes built/sec", "workload": "%d averagine formulas 100 Da - 50 kDa, natural abundance" % args.build_formulas,
           "envelopes": int(c["envelopes"]), "value_device": c["envelopes"] / (c["ms_envelopes"] * 1e-3),
           "value_e2e_host_strings_to_index": c["envelopes"] / wall, "unit": "envelopes/s", "ms_envelope_kernel": c["ms_envelopes"],
           "ms_trees": c["ms_trees"], "ms_index": c["ms_index"], "fp64_gflops": c["envelope_flops"] / (c["ms_envelopes"] * 1e-3) / 1e9,
           "wall_s": wall, "wall_s_all": walls, "timing": "median of %d builds after 1 warm-up build" % (len(walls) - 1)}
    # FP64 roofline of the envelope kernel: DFMA-chain and DMUL+DADD-chain peaks measured on this GPU
    import ctypes
    fused, unfused = ctypes.c_double(), ctypes.c_double()
    sweep._ctx.call("qms_measure_fp64_peak", ctypes.byref(fused), ctypes.byref(unfused))
    achieved = out["fp64_gflops"] / 1e3
    out["roofline"] = {"bound": "fp64", "kernel": "envelope_convolve_kernel", "achieved": achieved, "unit": "TFLOP/s",
                       "peak": unfused.value, "frac": achieved / unfused.value if unfused.value else None,
                       "peak_fused_dfma": fused.value, "frac_of_fused": achieved / fused.value if fused.value else None,
                       "peak_source": "qms_measure_fp64_peak: dependency-chain kernels on this GPU; the kernel may not fuse (parity)"}
    if not args.no_cpu_baseline:
        from oracle.qms_oracle import OracleLibrary
        sys.setrecursionlimit(20000)
        # the reference's build is O(N^2) big-int work in the largest atom count (552 s for one 50 kDa formula,
        # SURVEY.md section 6): the CPU sample is bounded to formulas below ~3 kDa (<= 130 carbons)
        small = [f for f in formulas[:20000] if int(f[2:].split("H")[0]) <= 130][:args.build_cpu_formulas]
        t0 = time.perf_counter()
        ora = OracleLibrary(molecules=small, charges=[1], params={"UPPER_MZ_LIMIT": 60000, "LOWER_MZ_LIMIT": 50})
        secs = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": len(ora) / secs, "unit": "envelopes/s", "cores": 1, "kind": "port",
                               "sample": "first %d formulas <= 130 C (~3 kDa) of the same sweep (%.1f s)" % (len(small), secs)}
    del sweep
    return out


def main():
    # exactly one line may reach stdout (the JSON line): libraries that print to fd 1 (NCCL's version banner)
    # are diverted to stderr for the whole run, the JSON line is written to the saved descriptor
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="proteome", choices=list(WORKLOADS))
    ap.add_argument("--proteome-size", type=int, default=500000)
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU matching time for cpu_baseline")
    ap.add_argument("--cpu-port", action="store_true", help="CPU legs with the oracle port even if baseline/_ref is there")
    ap.add_argument("--ref-spectra-per-core", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the sparse-library probe figure and the build metric")
    ap.add_argument("--build-formulas", type=int, default=100000)
    ap.add_argument("--build-repeats", type=int, default=3)
    ap.add_argument("--build-cpu-formulas", type=int, default=2000)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
