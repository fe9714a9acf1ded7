"""Seeded spectra for the parity cases (works on any library object with the reference dict layout)."""
import numpy as np

from pyqms_b200.synthetic import synth_run
from tests.golden.cases import QUICK_START_PEAKS


def entry_peaks(lib):
    """Per index entry: theoretical c-peak m/z and integer abundance, from the reference dict layout."""
    mzs, cis = [], []
    for lo, hi, z, lp, formula in lib.formulas_sorted_by_mz:
        env = lib[formula]["env"][lp]
        keep = [n for n, p in enumerate(env["c_peak_pos"]) if p is not None]
        mzs.append(np.array([env[z]["mz"][n] for n in keep]))
        cis.append(np.array([env["abun"][n] for n in keep], dtype=np.float64))
    return mzs, cis


def draw_case_spectra(lib, run, name):
    run = dict(run)
    dense = run.pop("dense", False)
    crowd = run.pop("crowd", 0)
    offset_ppm = run.pop("offset_ppm", 0.0)
    mzs, cis = entry_peaks(lib)
    lo = float(lib.params["LOWER_MZ_LIMIT"]) - 30
    hi = float(lib.params["UPPER_MZ_LIMIT"]) + 30
    peaks, offsets = synth_run(mzs, cis, mz_lo=lo, mz_hi=hi, min_peaks=20, **run)
    if dense:
        # crowd every planted peak with near-duplicates: same 0.001-Da bucket and neighbouring buckets
        rng = np.random.default_rng(run["seed"] + 1000)
        spec = np.repeat(np.arange(len(offsets) - 1), np.diff(offsets))
        lib_mz = np.concatenate(mzs)
        near = np.abs(peaks[:, 0][:, None] - lib_mz[None, :]).min(axis=1) < 0.01
        idx = np.flatnonzero(near)
        extra_spec, extra = [], []
        for rep in range(3):
            pick = idx[rng.random(idx.size) < 0.5]
            jitter = rng.choice([2e-4, 5e-4, 1.1e-3, 2.3e-3, -1.2e-3, -4e-4], size=pick.size)
            extra.append(np.column_stack([peaks[pick, 0] + jitter,
                                          peaks[pick, 1] * rng.uniform(0.3, 1.5, size=pick.size)]))
            extra_spec.append(spec[pick])
        spec = np.concatenate([spec] + extra_spec)
        allp = np.concatenate([peaks] + extra)
        order = np.lexsort((allp[:, 1], allp[:, 0], spec))
        peaks = np.ascontiguousarray(allp[order])
        offsets = np.zeros(len(offsets), dtype=np.int64)
        np.cumsum(np.bincount(spec, minlength=len(offsets) - 1), out=offsets[1:])
    if crowd:
        # `crowd` peaks in distinct 1/P-Da buckets inside every window of every entry, in every spectrum: crowd ** n_c
        # candidate combinations per (spectrum, entry) pair (IL:2004-2034)
        rng = np.random.default_rng(run["seed"] + 2000)
        rel = float(lib.params["REL_MZ_RANGE"])
        precision = float(lib.params["INTERNAL_PRECISION"])
        n_spec = len(offsets) - 1
        rows = [[] for _ in range(n_spec)]
        for mz_list, ci_list in zip(mzs, cis):
            for mz, ci in zip(mz_list, ci_list):
                half = int(mz * rel * precision) - 1
                for s in range(n_spec):
                    picks = rng.choice(np.arange(-half, half + 1), size=crowd, replace=False)
                    for b in picks:
                        rows[s].append(((round(mz * precision) + int(b)) / precision,
                                        float(ci * 40.0 * rng.uniform(0.5, 1.5))))
        spec = np.repeat(np.arange(n_spec), np.diff(offsets))
        extra_spec = np.concatenate([np.full(len(r), s) for s, r in enumerate(rows)])
        allp = np.concatenate([peaks] + [np.array(r).reshape(-1, 2) for r in rows])
        spec = np.concatenate([spec, extra_spec])
        order = np.lexsort((allp[:, 1], allp[:, 0], spec))
        peaks = np.ascontiguousarray(allp[order])
        offsets = np.zeros(n_spec + 1, dtype=np.int64)
        np.cumsum(np.bincount(spec, minlength=n_spec), out=offsets[1:])
    if name == "c1_basic":   # prepend the docs' quick-start spectrum as spectrum 0
        qs = np.array(QUICK_START_PEAKS)
        peaks = np.concatenate([qs, peaks])
        offsets = np.concatenate([[0], offsets + len(qs)])
    return peaks, offsets
