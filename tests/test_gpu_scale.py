"""Full-size parity through size-independent properties (BASELINE.json configs at their named shape, or a
bounded version for the proteome): ordering, idempotence, batch-split invariance, sub-library independence,
and exact agreement with the CPU oracle on a random sample of spectra."""
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from pyqms_b200.synthetic import (BSA_MOLECULES, CAM_FIXED_LABEL, averagine_formulas, random_tryptic_peptides,  # noqa: E402
                                  synth_run)
from tests.helpers import assert_results_equal, results_to_rows  # noqa: E402


def packed_rows(out, lo=None, hi=None):
    nh = out["n_hits"]
    rows = []
    for h in range(nh):
        s = int(out["spec"][h])
        if lo is not None and not (lo <= s < hi):
            continue
        a, b = int(out["win_off"][h]), int(out["win_off"][h + 1])
        rows.append((s, int(out["entry"][h]), float(out["score"][h]), float(out["scaling"][h]),
                     tuple(np.nan_to_num(out["mmz"][a:b], nan=-1.0)), tuple(np.nan_to_num(out["mi"][a:b], nan=-1.0))))
    return rows


def test_bsa_full_run_properties():
    """configs[1]: BSA 14N+15N z1-4 vs the 30 000-spectrum run of bench.py."""
    import pyqms_b200
    from oracle.qms_oracle import OracleLibrary
    kw = dict(molecules=BSA_MOLECULES, charges=[1, 2, 3, 4], metabolic_labels={"15N": [0, 0.99]}, fixed_labels=CAM_FIXED_LABEL)
    lib = pyqms_b200.IsotopologueLibrary(verbose=False, **kw)
    mz, ci = lib.entry_peaks()
    peaks, offsets = synth_run(mz, ci, n_spectra=30000, seed=20260101)
    out = lib.match_packed(peaks, offsets)
    nh = out["n_hits"]
    assert nh > 5000
    key = out["spec"][:nh].astype(np.int64) << 32 | out["entry"][:nh]
    assert np.all(key[1:] > key[:-1])                                   # (spectrum, entry) strictly ascending
    assert np.all(out["score"][:nh] >= lib.params["M_SCORE_THRESHOLD"])
    again = lib.match_packed(peaks, offsets)
    assert packed_rows(out) == packed_rows(again)                       # idempotent
    cut = 17321                                                         # batch split invariance
    first = lib.match_packed(peaks[:offsets[cut]], offsets[:cut + 1])
    second = lib.match_packed(peaks[offsets[cut]:], offsets[cut:] - offsets[cut])
    shifted = [(r[0] + cut,) + r[1:] for r in packed_rows(second)]
    assert packed_rows(first) + shifted == packed_rows(out)
    # exact agreement with the oracle on 300 random spectra (hit sets, matched peaks; score 1e-6)
    ora = OracleLibrary(**kw)
    sample = sorted(np.random.default_rng(1).choice(30000, size=300, replace=False).tolist())
    want = {}
    for s in sample:
        ora.match_all(peaks[offsets[s]:offsets[s + 1]], "run", s, s * 0.2, want)
    got = lib.match_many([peaks[offsets[s]:offsets[s + 1]] for s in sample], file_name="run", spec_ids=sample,
                         spec_rts=[s * 0.2 for s in sample])
    assert_results_equal(results_to_rows(got), results_to_rows(want), score_tol=1e-6, cmz_rel=1e-12, ri_rel=1e-9)
    assert sum(len(v) for v in want.values()) > 50


def test_proteome_sublibrary_independence():
    """configs[3] shape, bounded: 20 000 tryptic peptides z1-5 CAM vs 1 500 spectra.  Matching of an entry does
    not depend on the other entries (numpy input), so the hits of a 300-peptide sub-library computed by the CPU
    oracle must equal the full-library GPU hits restricted to those formulas."""
    import pyqms_b200
    from oracle.qms_oracle import OracleLibrary
    peptides = random_tryptic_peptides(20000, seed=20260104)
    lib = pyqms_b200.IsotopologueLibrary(molecules=peptides, charges=[1, 2, 3, 4, 5], fixed_labels=CAM_FIXED_LABEL, verbose=False)
    assert lib._n_entries > 60000
    mz, ci = lib.entry_peaks()
    peaks, offsets = synth_run(mz, ci, n_spectra=1500, seed=20260104, entry_fraction=0.02, elution_sigma=3.0)
    subset = peptides[::67][:300]
    ora = OracleLibrary(molecules=subset, charges=[1, 2, 3, 4, 5], fixed_labels=CAM_FIXED_LABEL)
    sample = list(range(0, 1500, 10))
    want = {}
    for s in sample:
        ora.match_all(peaks[offsets[s]:offsets[s + 1]], "run", s, float(s), want)
    got = lib.match_many([peaks[offsets[s]:offsets[s + 1]] for s in sample], file_name="run", spec_ids=sample,
                         spec_rts=[float(s) for s in sample])
    formulas = set(ora.keys())
    got_rows = [r for r in results_to_rows(got) if r[0][1] in formulas]
    assert_results_equal(got_rows, results_to_rows(want), score_tol=1e-6, cmz_rel=1e-12, ri_rel=1e-9, key_order=False)
    assert sum(len(v) for v in want.values()) > 20
    # dense-library regime sanity: every spectrum row is live, hits stay ordered
    out = lib.match_packed(peaks, offsets)
    nh = out["n_hits"]
    key = out["spec"][:nh].astype(np.int64) << 32 | out["entry"][:nh]
    assert np.all(key[1:] > key[:-1])


def test_spectrum_resident_matcher_equals_row_streaming_matcher():
    """The two matchers of the library (match.cu / match_spectrum.cu) on a dense library: identical packed hits,
    bit for bit, including spectra too large for the shared-memory bucket array (global spill) and empty spectra."""
    import pyqms_b200
    peptides = random_tryptic_peptides(20000, seed=20260104)
    lib = pyqms_b200.IsotopologueLibrary(molecules=peptides, charges=[1, 2, 3, 4, 5], fixed_labels=CAM_FIXED_LABEL, verbose=False)
    mz, ci = lib.entry_peaks()
    peaks, offsets = synth_run(mz, ci, n_spectra=600, seed=77, entry_fraction=0.05, elution_sigma=3.0)
    big, big_off = synth_run(mz, ci, n_spectra=3, seed=78, mean_peaks=9000, sigma_peaks=0.05, max_peaks=12000, entry_fraction=0.01)
    empty = np.zeros(2, dtype=np.int64)
    peaks = np.concatenate([peaks, big])
    offsets = np.concatenate([offsets, offsets[-1] + empty[1:], offsets[-1] + big_off[1:]])      # one empty spectrum in between
    assert np.diff(offsets).max() > 4096 and np.diff(offsets).min() == 0
    outs = {}
    for mode in (0, 1):
        lib._ctx.set_tuning("dense_mode", mode)
        before = lib.device_counters()["dense_batches"]
        outs[mode] = lib.match_packed(peaks, offsets)
        assert lib.device_counters()["dense_batches"] - before == mode
    lib._ctx.set_tuning("dense_mode", 2)
    a, b = outs[0], outs[1]
    assert a["n_hits"] == b["n_hits"] > 1000 and a["n_windows"] == b["n_windows"]
    for k in ("spec", "entry", "win_off", "peak"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("score", "scaling", "mmz", "mi"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    # the pipelined run: sub-batches of ~60 000 rows, copies of neighbouring sub-batches overlapping the kernels;
    # first with result arrays that are too small (QMS_E_CAPACITY -> qms_fetch_hits), then with room, then lazily
    lib._ctx.set_tuning("sub_batch_rows", 60000)
    try:
        for attempt in range(2):
            if attempt == 0:
                lib._hit_caps = (64, 512)
            before = lib.device_counters()["dense_batches"]
            c = lib.match_packed(peaks, offsets)
            assert lib.device_counters()["dense_batches"] - before > 5
            assert c["n_hits"] == b["n_hits"] and c["n_windows"] == b["n_windows"]
            for k in ("spec", "entry", "win_off", "peak"):
                assert np.array_equal(b[k], c[k]), (attempt, k)
            for k in ("score", "scaling", "mmz", "mi"):
                assert np.array_equal(b[k], c[k], equal_nan=True), (attempt, k)
        n = len(offsets) - 1
        res = lib.match_many(peaks, file_name="run", spec_ids=list(range(n)), spec_rts=list(range(n)), offsets=offsets)
        lazy = res.packed_batches()[-1]
        assert "mmz" not in dict.keys(lazy)                  # the values did not cross the bus ...
        assert "peak16" in dict.keys(lazy) and "spec" not in dict.keys(lazy)      # ... and the index columns came in their compact forms
        for k in ("spec", "entry", "win_off", "peak", "score", "scaling"):
            assert np.array_equal(b[k], lazy[k]), k
        assert np.array_equal(b["mmz"], lazy["mmz"], equal_nan=True) and np.array_equal(b["mi"], lazy["mi"], equal_nan=True)   # ... and are gathered on demand
        # the compact columns on every way out of the device: fetched after a capacity miss, from the row-streaming matcher,
        # and -- switched off -- the int32 rows instead
        for variant in ("fetch", "row-streaming", "rows32"):
            if variant == "fetch":
                lib._hit_caps = (64, 512)
            lib._ctx.set_tuning("dense_mode", 0 if variant == "row-streaming" else 2)
            lib._compact_hits = variant != "rows32"
            got = lib._run_matcher(peaks, offsets, 0, lib.params["MZ_SCORE_PERCENTILE"], values=False)
            assert ("peak16" in dict.keys(got)) == (variant != "rows32") and ("peak32" in dict.keys(got)) == (variant == "rows32")
            for k in ("spec", "entry", "win_off", "peak", "score", "scaling"):
                assert np.array_equal(b[k], got[k]), (variant, k)
            assert np.array_equal(b["mmz"], got["mmz"], equal_nan=True), variant
            if variant != "rows32":
                compact_rows = dict.__getitem__(got, "peak16").copy()
        # hit_nc8 through the C-ABI, and the limits of the compact forms
        import ctypes as C
        from pyqms_b200._capi import QMS_E_LIMIT, QmsHits
        nh, nw = b["n_hits"], b["n_windows"]
        nc8, peak16, spec_off = np.zeros(nh, np.uint8), np.zeros(nw, np.uint16), np.full(n + 1, -1, np.int64)
        hits = QmsHits(nh, nw, None, None, None, None, None, None, None, None, None, nc8.ctypes.data, peak16.ctypes.data, spec_off.ctypes.data)
        got_h, got_w = C.c_int64(), C.c_int64()
        lib._ctx.check(lib._ctx.lib.qms_fetch_hits(lib._ctx.handle, C.byref(hits), C.byref(got_h), C.byref(got_w)))
        assert got_h.value == nh and np.array_equal(nc8, np.diff(b["win_off"])) and np.array_equal(peak16, compact_rows)
        assert np.array_equal(spec_off, np.searchsorted(b["spec"], np.arange(n + 1)))
        long_off = np.array([0, len(peaks)], dtype=np.int64)                      # one spectrum of > 65 534 rows
        order = np.argsort(peaks[:, 0], kind="stable")
        rc = lib._ctx.lib.qms_match_batch(lib._ctx.handle, np.ascontiguousarray(peaks[order]).ctypes.data, long_off.ctypes.data, 1, 0,
                                          C.c_double(lib.params["MZ_SCORE_PERCENTILE"]), C.byref(hits), C.byref(got_h), C.byref(got_w))
        assert len(peaks) > 65534 and rc == QMS_E_LIMIT
    finally:
        lib._compact_hits = True
        lib._ctx.set_tuning("dense_mode", 2)
        lib._ctx.set_tuning("sub_batch_rows", 0)


def test_bsa_percentile_grid_run_properties():
    """configs[2] at its named library size (19 BSA molecules x 100 label percentiles, 7 600 entries whose windows pile up
    on the same m/z: the combination-heavy regime) against 6 000 spectra of the run: both matchers agree bit for bit, hits
    are ordered, a split run equals the whole."""
    import pyqms_b200
    lib = pyqms_b200.IsotopologueLibrary(molecules=BSA_MOLECULES, charges=[1, 2, 3, 4], fixed_labels=CAM_FIXED_LABEL,
                                         metabolic_labels={"15N": [round(0.01 * i, 2) for i in range(100)]}, verbose=False)
    assert lib._n_entries == 7600
    mz, ci = lib.entry_peaks()
    peaks, offsets = synth_run(mz, ci, n_spectra=6000, seed=20260101)
    outs = {}
    for mode in (0, 1):
        lib._ctx.set_tuning("dense_mode", mode)
        outs[mode] = lib.match_packed(peaks, offsets)
    lib._ctx.set_tuning("dense_mode", 2)
    a, b = outs[0], outs[1]
    assert a["n_hits"] == b["n_hits"] > 100000
    for k in ("spec", "entry", "win_off", "peak"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("score", "scaling", "mmz", "mi"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    key = a["spec"].astype(np.int64) << 32 | a["entry"]
    assert np.all(key[1:] > key[:-1])
    cut = 2517
    first = lib.match_packed(peaks[:offsets[cut]], offsets[:cut + 1])
    second = lib.match_packed(peaks[offsets[cut]:], offsets[cut:] - offsets[cut])
    assert np.array_equal(np.concatenate([first["score"], second["score"]]), a["score"])
    assert np.array_equal(np.concatenate([first["entry"], second["entry"]]), a["entry"])


def test_proteome_at_its_named_size():
    """configs[3] at full size: the 500 000-peptide library against shard 0 of the 200 000-spectrum run (25 000 spectra,
    39.4 M peaks; the spectra come from the committed planted-peak table, like in bench.py).  Size-independent properties:
    hits strictly ordered by (spectrum, entry), every score above the threshold, idempotence, batch-split invariance
    (device-resident single batch = pipelined host run = two halves), and exact agreement with the CPU oracle for a
    300-peptide sub-library on 40 spectra (matching an entry does not depend on the other entries, numpy input)."""
    import sys
    from pathlib import Path
    import pyqms_b200
    from oracle.qms_oracle import OracleLibrary
    root = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root))
    import bench
    table = np.load(bench.FIXTURE)
    planted = {"n_entries": int(table["n_entries"]), "lens": table["lens"].astype(np.int64), "mz": table["mz"], "ci": table["ci"].astype(np.float64)}
    peaks, offsets = synth_run(None, None, n_spectra=25000, seed=20260104, entry_fraction=0.02, planted=planted)
    assert len(peaks) == 39371082
    peptides = random_tryptic_peptides(500000, seed=20260104)
    lib = pyqms_b200.IsotopologueLibrary(molecules=peptides, charges=[1, 2, 3, 4, 5], fixed_labels=CAM_FIXED_LABEL, verbose=False)
    assert lib._n_entries == planted["n_entries"] == 1585058
    whole = lib.match_packed(peaks, offsets)                                  # pipelined host run (sub-batches)
    assert lib.device_counters()["dense_batches"] >= 3
    nh = whole["n_hits"]
    assert nh == 25096897
    key = whole["spec"].astype(np.int64) << 32 | whole["entry"]
    assert np.all(key[1:] > key[:-1])
    assert np.all(whole["score"] >= lib.params["M_SCORE_THRESHOLD"])
    assert np.all(np.diff(whole["win_off"]) >= 2) and whole["win_off"][-1] == whole["n_windows"]
    rows = whole["peak"][whole["peak"] >= 0]
    assert np.array_equal(peaks[rows, 0], whole["mmz"][whole["peak"] >= 0])   # matched values = the caller's rows
    cut = 11003                                                               # two halves = the whole
    first = lib.match_packed(peaks[:offsets[cut]], offsets[:cut + 1])
    second = lib.match_packed(peaks[offsets[cut]:], offsets[cut:] - offsets[cut])
    assert first["n_hits"] + second["n_hits"] == nh
    for k in ("entry", "score", "scaling"):
        assert np.array_equal(np.concatenate([first[k], second[k]]), whole[k]), k
    assert np.array_equal(np.concatenate([first["spec"], second["spec"] + cut]), whole["spec"])
    del first, second
    # oracle on a sub-library: same hits for its formulas
    sample = list(range(0, 25000, 625))
    subset = peptides[::1667][:300]
    ora = OracleLibrary(molecules=subset, charges=[1, 2, 3, 4, 5], fixed_labels=CAM_FIXED_LABEL)
    want = {}
    for s in sample:
        ora.match_all(peaks[offsets[s]:offsets[s + 1]], "run", s, float(s), want)
    got = lib.match_many([peaks[offsets[s]:offsets[s + 1]] for s in sample], file_name="run", spec_ids=sample,
                         spec_rts=[float(s) for s in sample])
    formulas = set(ora.keys())
    got_rows = [r for r in results_to_rows(got) if r[0][1] in formulas]
    assert_results_equal(got_rows, results_to_rows(want), score_tol=1e-6, cmz_rel=1e-12, ri_rel=1e-9, key_order=False)
    assert sum(len(v) for v in want.values()) > 5


def test_envelope_sweep_against_oracle():
    """configs[4] shape, bounded: averagine formulas 100 Da - 12 kDa, FP64 abundance parity 1e-9 vs the oracle."""
    sys.setrecursionlimit(20000)
    import pyqms_b200
    from oracle.qms_oracle import OracleLibrary
    formulas = averagine_formulas(60, seed=20260105, min_mass=100.0, max_mass=12000.0)
    kw = dict(molecules=formulas, charges=[1, 4, 9], params={"UPPER_MZ_LIMIT": 20000, "LOWER_MZ_LIMIT": 10})
    lib = pyqms_b200.IsotopologueLibrary(verbose=False, **kw)
    ora = OracleLibrary(**kw)
    assert sorted(lib.keys()) == sorted(ora.keys())
    lp = (("N", "0.000"),)
    worst = 0.0
    for formula in ora.keys():
        a, b = lib[formula]["env"][lp], ora[formula]["env"][lp]
        assert a["abun"] == b["abun"] and a["c_peak_pos"] == b["c_peak_pos"] and a["n_c_peaks"] == b["n_c_peaks"], formula
        for x, y in zip(a["relabun"], b["relabun"]):
            worst = max(worst, abs(x - y) / y)
        for z in (1, 4, 9):
            assert [None if s is None else (min(s), max(s)) for s in a[z]["tmzs"]] == \
                [None if s is None else (min(s), max(s)) for s in b[z]["tmzs"]], (formula, z)
            for x, y in zip(a[z]["mz"], b[z]["mz"]):
                assert abs(x - y) <= 1e-12 * y
    assert worst < 1e-9
