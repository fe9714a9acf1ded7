#!/usr/bin/env python
"""Scale measurements for the larger BASELINE.json configs (run on the GPU box; writes gpurun_out/scale_*.json).

    python tools/scale_run.py proteome --peptides 500000 --spectra 25000
    python tools/scale_run.py envelopes --formulas 1000000
    python tools/scale_run.py bsa_pct
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import pyqms_b200  # noqa: E402
from pyqms_b200.synthetic import (BSA_MOLECULES, CAM_FIXED_LABEL, averagine_formulas, random_tryptic_peptides,  # noqa: E402
                                  synth_run)


def timed_build(**kw):
    t0 = time.perf_counter()
    lib = pyqms_b200.IsotopologueLibrary(verbose=False, **kw)
    wall = time.perf_counter() - t0
    c = lib.device_counters()
    return lib, {"wall_s": wall, "envelopes": c["envelopes"], "entries": int(lib._n_entries), "windows": int(lib._n_windows),
                 "ms_trees": c["ms_trees"], "ms_envelopes": c["ms_envelopes"], "ms_index": c["ms_index"],
                 "envelope_flops": c["envelope_flops"],
                 "envelopes_per_s_device": c["envelopes"] / (c["ms_envelopes"] * 1e-3) if c["ms_envelopes"] else None,
                 "gflops_fp64": c["envelope_flops"] / (c["ms_envelopes"] * 1e-3) / 1e9 if c["ms_envelopes"] else None,
                 "device_bytes": lib._ctx.device_bytes()}


def timed_match(lib, peaks, offsets, reps=3):
    lib.match_packed(peaks, offsets)          # warm-up (allocations)
    lib._ctx.reset_counters()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = lib.match_packed(peaks, offsets)
    wall = (time.perf_counter() - t0) / reps
    c = lib.device_counters()
    n = len(offsets) - 1
    per = {k: c[k] / reps for k in ("peaks", "live_peaks", "incidences", "gated_incidences", "candidates", "candidate_windows", "combinations", "hits",
                                    "ms_probe", "ms_gate", "ms_score", "ms_compact", "ms_h2d", "ms_d2h")}
    device_ms = per["ms_probe"] + per["ms_gate"] + per["ms_score"] + per["ms_compact"]
    return {"spectra": n, "wall_s_e2e_host_buffers": wall, "spectra_per_s_e2e": n / wall, "device_ms": device_ms,
            "spectra_per_s_device": n / (device_ms * 1e-3), "per_step": per, "hits": int(out["n_hits"])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["proteome", "envelopes", "bsa_pct"])
    ap.add_argument("--peptides", type=int, default=500000)
    ap.add_argument("--spectra", type=int, default=25000)
    ap.add_argument("--formulas", type=int, default=1000000)
    ap.add_argument("--max-mass", type=float, default=50000.0)
    args = ap.parse_args()
    out = {"what": args.what}
    if args.what == "proteome":
        t0 = time.perf_counter()
        peptides = random_tryptic_peptides(args.peptides, seed=20260104)
        out["generate_peptides_s"] = time.perf_counter() - t0
        lib, out["build"] = timed_build(molecules=peptides, charges=[1, 2, 3, 4, 5], fixed_labels=CAM_FIXED_LABEL)
        mz, ci = lib.entry_peaks()
        peaks, offsets = synth_run(mz, ci, n_spectra=args.spectra, seed=20260104, entry_fraction=0.02)
        out["run"] = {"spectra": args.spectra, "peaks": int(len(peaks))}
        out["match"] = timed_match(lib, peaks, offsets)
    elif args.what == "bsa_pct":
        lib, out["build"] = timed_build(molecules=BSA_MOLECULES, charges=[1, 2, 3, 4], fixed_labels=CAM_FIXED_LABEL,
                                        metabolic_labels={"15N": [round(0.01 * i, 2) for i in range(100)]})
        mz, ci = lib.entry_peaks()
        peaks, offsets = synth_run(mz, ci, n_spectra=30000, seed=20260101)
        out["run"] = {"spectra": 30000, "peaks": int(len(peaks))}
        out["match"] = timed_match(lib, peaks, offsets)
    else:
        t0 = time.perf_counter()
        formulas = averagine_formulas(args.formulas, seed=20260105, max_mass=args.max_mass)
        out["generate_formulas_s"] = time.perf_counter() - t0
        lib, out["build"] = timed_build(molecules=formulas, charges=[1], params={"UPPER_MZ_LIMIT": 60000, "LOWER_MZ_LIMIT": 50})
    path = ROOT / "gpurun_out" / ("scale_%s.json" % args.what)
    path.parent.mkdir(exist_ok=True)
    path.write_text(json.dumps(out, indent=1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
