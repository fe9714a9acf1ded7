#!/usr/bin/env python
"""Benchmark of the matching hot path (spectra/s through match_all) on B200 -- see DESIGN.md section 6.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload bsa|bsa_pct|proteome]

One *step* = one pass of the matcher over one synthetic LC-MS run (the `bsa` workload: BASELINE.json
configs[1], 19 BSA molecules, 14N+15N, charges 1-4, carbamidomethyl fixed label, 30 000 spectra).
Prints ONE JSON line (rank 0).  `value` = spectra/s with the run resident in HBM, device-timed (CUDA events
on the library's stream, max over ranks); `e2e` = the same through the public host API with pinned HOST
buffers (H2D of the spectra and D2H of the packed hits inside the timed region); `roofline` = the probe
kernel against the measured HBM peak; `cpu_baseline` = the CPU oracle port (the reference's algorithm,
pure Python like the reference) on a bounded sample of the same run on this box's host cores.
`--impl reference` times that CPU path alone, on all host cores (one process per core over spectrum shards).
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
    # BASELINE.json configs[1] / SURVEY.md 8d "C2"
    "bsa": dict(lib=dict(molecules=BSA_MOLECULES, charges=[1, 2, 3, 4], metabolic_labels={"15N": [0, 0.99]},
                         fixed_labels=CAM_FIXED_LABEL),
                run=dict(n_spectra=30000, seed=20260101), name="BSA 19 molecules 14N+15N z1-4 CAM vs 30k-spectrum synthetic run"),
    # configs[2] "C3"
    "bsa_pct": dict(lib=dict(molecules=BSA_MOLECULES, charges=[1, 2, 3, 4], fixed_labels=CAM_FIXED_LABEL,
                             metabolic_labels={"15N": [round(0.01 * i, 2) for i in range(100)]}),
                    run=dict(n_spectra=30000, seed=20260101), name="BSA x 100 15N percentiles vs 30k-spectrum run"),
    # configs[3] "C4" scaled by --proteome-size (default 500k peptides is the named size)
    "proteome": dict(lib=dict(charges=[1, 2, 3, 4, 5], fixed_labels=CAM_FIXED_LABEL),
                     run=dict(n_spectra=25000, seed=20260104, entry_fraction=0.02),
                     name="random tryptic proteome z1-5 CAM vs 25k spectra per GPU"),
}


def make_run(entry_mz, entry_ci, run, rank):
    kw = dict(run)
    kw["seed"] = kw["seed"] + rank
    return synth_run(entry_mz, entry_ci, **kw)


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


# ------------------------------------------------------------------------------------- CPU legs (oracle)
def _oracle_library(workload, proteome_size):
    from oracle.qms_oracle import OracleLibrary
    sys.setrecursionlimit(20000)
    kw = dict(WORKLOADS[workload]["lib"])
    if workload == "proteome":
        kw["molecules"] = random_tryptic_peptides(proteome_size, seed=20260104)
    return OracleLibrary(**kw)


def _oracle_entry_peaks(lib):
    mzs, cis = [], []
    for lo, hi, z, lp, formula in lib.formulas_sorted_by_mz:
        env = lib[formula]["env"][lp]
        keep = [n for n, p in enumerate(env["c_peak_pos"]) if p is not None]
        mzs.append(np.array([env[z]["mz"][n] for n in keep]))
        cis.append(np.array([env["abun"][n] for n in keep], dtype=np.float64))
    return mzs, cis


_WORKER = {}


def _cpu_worker_init(workload, proteome_size):
    _WORKER["lib"] = _oracle_library(workload, proteome_size)


def _cpu_worker_run(spectra):
    lib = _WORKER["lib"]
    t0 = time.perf_counter()
    res = {}
    for s, spec in enumerate(spectra):
        lib.match_all(spec, "run", s, s * 0.2, res)
    return time.perf_counter() - t0, sum(len(v) for v in res.values())


def cpu_time_sample(lib, peaks, offsets, budget_s, max_spectra):
    """Time the oracle on leading spectra of the run until ~budget_s elapsed (single core)."""
    res, n, t0 = {}, 0, time.perf_counter()
    while n < min(max_spectra, len(offsets) - 1):
        lib.match_all(peaks[offsets[n]:offsets[n + 1]], "run", n, n * 0.2, res)
        n += 1
        if n >= 20 and time.perf_counter() - t0 > budget_s:
            break
    return n, time.perf_counter() - t0, sum(len(v) for v in res.values())


def run_reference_arm(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference is pure Python and cannot
    travel to the GPU box) on all host cores, one process per core over disjoint spectrum shards."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = max(1, (os.cpu_count() or 1))
    cores = min(cores, 64)
    lib = _oracle_library(args.workload, args.proteome_size)
    mzs, cis = _oracle_entry_peaks(lib)
    per_proc = args.ref_spectra_per_core
    run = dict(WORKLOADS[args.workload]["run"])
    run["n_spectra"] = min(run["n_spectra"], per_proc * cores)
    peaks, offsets = make_run(mzs, cis, run, 0)
    n = len(offsets) - 1
    shards = [[peaks[offsets[s]:offsets[s + 1]] for s in range(i, n, cores)] for i in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_worker_init, initargs=(args.workload, args.proteome_size)) as pool:
        times = []
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pool.map(_cpu_worker_run, shards)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = n * len(times) / total
    line = {
        "impl": "reference", "metric": "spectra/sec through match_all", "value": value, "unit": "spectra/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload]["name"], "library_entries": len(lib.formulas_sorted_by_mz)},
        "cpu_baseline": {"value": value, "unit": "spectra/s", "cores": cores, "kind": "port",
                         "sample": "%d spectra of the run per step, %d processes x %d spectra" % (n, cores, len(shards[0]))},
        "e2e": {"value": value, "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------- our arm
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
    from pyqms_b200.distributed import bind_to_gpu_numa_node
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None    # before any pinned allocation
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=300))
    wl = WORKLOADS[args.workload]
    kw = dict(wl["lib"])
    if args.workload == "proteome":
        kw["molecules"] = random_tryptic_peptides(args.proteome_size, seed=20260104)
    t0 = time.perf_counter()
    lib = pyqms_b200.IsotopologueLibrary(verbose=False, device=local, **kw)     # sharded build + all-gather when world > 1
    build_s = time.perf_counter() - t0
    build_counters = lib.device_counters()
    entry_mz, entry_ci = lib.entry_peaks()
    peaks, offsets = make_run(entry_mz, entry_ci, wl["run"], rank)
    n_spec = len(offsets) - 1
    ctx = lib._ctx
    xi = lib.params["MZ_SCORE_PERCENTILE"]
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    # ---- resident inputs + device output buffers
    d_peaks = torch.from_numpy(peaks).cuda()
    d_offsets = torch.from_numpy(offsets).cuda()
    cap_h, cap_w = 1 << 16, 1 << 19
    while True:
        dev = {"spec": torch.empty(cap_h, dtype=torch.int32, device="cuda"), "entry": torch.empty(cap_h, dtype=torch.int32, device="cuda"),
               "score": torch.empty(cap_h, dtype=torch.float64, device="cuda"), "scaling": torch.empty(cap_h, dtype=torch.float64, device="cuda"),
               "win_off": torch.empty(cap_h + 1, dtype=torch.int64, device="cuda"), "mmz": torch.empty(cap_w, dtype=torch.float64, device="cuda"),
               "mi": torch.empty(cap_w, dtype=torch.float64, device="cuda"), "peak": torch.empty(cap_w, dtype=torch.int64, device="cuda")}
        hits = QmsHits(cap_h, cap_w, *[dev[k].data_ptr() for k in ("spec", "entry", "score", "scaling", "win_off", "mmz", "mi", "peak")])
        nh, nw = C.c_int64(), C.c_int64()
        rc = ctx.lib.qms_match_batch(ctx.handle, d_peaks.data_ptr(), offsets.ctypes.data, n_spec, 0, xi, C.byref(hits), C.byref(nh),
                                     C.byref(nw))
        if rc == -3:
            cap_h, cap_w = nh.value + 1024, nw.value + 8192
            continue
        ctx.check(rc)
        break

    def device_step():
        # peaks resident in HBM; the 240 KB of row offsets stay a host array (the library needs the row count on the host)
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

    # ---- e2e through the public host API: pinned host spectra -> packed hits on the host
    pin_peaks = torch.from_numpy(peaks).pin_memory()
    pin_offsets = torch.from_numpy(offsets).pin_memory()
    h_peaks, h_offsets = pin_peaks.numpy(), pin_offsets.numpy()
    spec_ids = list(range(n_spec))

    def e2e_step():
        # the call a user makes for a whole run: host (N,2) array in, pyqms Results out (hits stay packed until read)
        res = lib.match_many(h_peaks, file_name="run", spec_ids=spec_ids, spec_rts=spec_ids, offsets=h_offsets)
        return res.packed_batches()[-1]

    for _ in range(max(1, args.warmup)):
        out = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n_spec * args.steps / float(t.item())
    h2d = int(peaks.nbytes + offsets.nbytes)
    d2h = int(out["n_hits"] * (4 + 4 + 8 + 8 + 8) + 8 + out["n_windows"] * 8)   # spec, entry, score, scaling, win_off + peak rows

    # ---- roofline of the dominant kernel (spectrum_probe_gate): SURVEY.md 8d algorithmic bytes
    launches_per_step = counters["kernel_launches"] / max(steps_counted, 1)
    probe_ms = counters["ms_probe"] / max(steps_counted, 1)
    # the probe kernel streams every (m/z, intensity) row once: 16 B per peak (first term); the
    # candidate and hit terms belong to the gate / score / compaction kernels and are reported with the step
    alg_bytes = 16.0 * counters["peaks"] / max(steps_counted, 1)
    step_alg_bytes = (16.0 * counters["peaks"] + 16.0 * counters["candidates"] + 28.0 * counters["candidate_windows"]
                      + 24.0 * counters["hits"] + 16.0 * counters["hit_windows"]) / max(steps_counted, 1)
    peaks_json = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    peak_gbs = float(peaks_json.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (probe_ms * 1e-3) / 1e9 if probe_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "spectrum_probe_warp_kernel", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                "frac": achieved / peak_gbs, "traffic": None, "peak_source": "MEASURED_PEAKS.json" if peaks_json else "fallback",
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": probe_ms,
                "step_algorithmic_bytes": step_alg_bytes, "step_achieved_gbs": step_alg_bytes / (ms_max / args.steps * 1e-3) / 1e9,
                "step_ms_breakdown": {k: counters[k] / max(steps_counted, 1) for k in ("ms_probe", "ms_gate", "ms_score", "ms_compact", "ms_d2h", "ms_h2d")}}
    traffic_file = ROOT / "profiles" / "probe_traffic.json"
    if traffic_file.exists():
        try:
            roofline["traffic"] = json.loads(traffic_file.read_text()).get(args.workload)
        except Exception:
            pass

    # ---- second half of the metric: isotopologue envelopes built per second (BASELINE.json configs[4] shape, bounded)
    build_metric = None
    if rank == 0 and not args.no_build_metric:
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
        build_metric = {"metric": "isotopologue envelopes built/sec", "workload": "%d averagine formulas 100 Da - 50 kDa, natural abundance" % args.build_formulas,
                        "envelopes": int(c["envelopes"]), "value_device": c["envelopes"] / (c["ms_envelopes"] * 1e-3),
                        "value_e2e_host_strings_to_index": c["envelopes"] / wall, "unit": "envelopes/s", "ms_envelope_kernel": c["ms_envelopes"],
                        "ms_trees": c["ms_trees"], "ms_index": c["ms_index"], "fp64_gflops": c["envelope_flops"] / (c["ms_envelopes"] * 1e-3) / 1e9,
                        "wall_s": wall, "wall_s_all": walls, "timing": "median of %d builds after 1 warm-up build" % (len(walls) - 1)}
        # FP64 roofline of the envelope kernel: DFMA-chain and DMUL+DADD-chain peaks measured on this GPU
        import ctypes
        fused, unfused = ctypes.c_double(), ctypes.c_double()
        sweep._ctx.call("qms_measure_fp64_peak", ctypes.byref(fused), ctypes.byref(unfused))
        achieved = build_metric["fp64_gflops"] / 1e3
        build_metric["roofline"] = {"bound": "fp64", "kernel": "envelope_convolve_kernel", "achieved": achieved, "unit": "TFLOP/s",
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
            build_metric["cpu_baseline"] = {"value": len(ora) / secs, "unit": "envelopes/s", "cores": 1, "kind": "port",
                                            "sample": "first %d formulas <= 130 C (~3 kDa) of the same sweep (%.1f s)" % (len(small), secs)}
        del sweep

    line = None
    if rank == 0:
        # ---- CPU baseline: oracle port on a bounded sample of the same run, one core
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            ora = _oracle_library(args.workload, min(args.proteome_size, 5000))
            n, secs, nhits = cpu_time_sample(ora, peaks, offsets, args.cpu_budget, 20000)
            cpu = {"value": n / secs, "unit": "spectra/s", "cores": 1, "kind": "port",
                   "sample": "first %d spectra of the same run (%.1f s, %d hits), oracle/qms_oracle.py single process" % (n, secs, nhits)}
        line = {
            "metric": "spectra/sec through match_all", "value": value, "unit": "spectra/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "spectra_per_gpu": n_spec, "peaks_per_gpu": int(len(peaks)),
                       "library_entries": int(lib._n_entries), "library_windows": int(lib._n_windows),
                       "l2_policy": "inputs (%d MB of peaks per step) larger than L2" % (peaks.nbytes >> 20),
                       "input_kind": "numpy (N,2) float64", "hits_per_step": int(nh.value), "numa_node_rank0": numa_node},
            "e2e": {"value": e2e_value, "unit": "spectra/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(round(launches_per_step * args.steps)),
            "clocks": clocks.summary(), "roofline": roofline, "cpu_baseline": cpu,
            "library_build": {"seconds_wall": build_s, "envelopes": int(build_counters["envelopes"]),
                              "ms_trees": build_counters["ms_trees"], "ms_envelopes": build_counters["ms_envelopes"],
                              "ms_index": build_counters["ms_index"],
                              "envelopes_per_s_device": (build_counters["envelopes"] / (build_counters["ms_envelopes"] * 1e-3)
                                                         if build_counters["ms_envelopes"] > 0 else None)},
            "build": build_metric,
            "counters_per_step": {k: counters[k] / max(steps_counted, 1) for k in
                                  ("peaks", "live_peaks", "incidences", "candidates", "combinations", "hits")},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    # exactly one line may reach stdout (the JSON line): libraries that print to fd 1 (NCCL's version banner)
    # are diverted to stderr for the whole run, the JSON line is written to the saved descriptor
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="bsa", choices=list(WORKLOADS))
    ap.add_argument("--proteome-size", type=int, default=500000)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU oracle time for cpu_baseline")
    ap.add_argument("--ref-spectra-per-core", type=int, default=250)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-build-metric", action="store_true")
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
