#!/usr/bin/env python
"""Randomised differential run: CUDA path vs the CPU oracle over random libraries, params and spectra (both input kinds).

    python tools/fuzz_parity.py [n_configs] [first_seed] [seconds_per_config=20]
    python tools/fuzz_parity.py --matchers [n_configs] [first_seed]      # the two matchers of the library against each other

Prints one line per configuration and exits non-zero at the first mismatch.  (Test infrastructure: imports oracle/.)
The oracle enumerates candidate combinations like the reference and can take minutes on a dense spectrum with wide
windows: every configuration runs under an alarm and is skipped when the oracle exceeds its time limit -- do not
hold a GPU box with an unbounded run.
"""
import random
import signal
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.setrecursionlimit(20000)
import pyqms_b200  # noqa: E402
from oracle.qms_oracle import OracleLibrary  # noqa: E402
from pyqms_b200.synthetic import averagine_formulas, random_tryptic_peptides, synth_run  # noqa: E402
from tests.golden.spectra import entry_peaks  # noqa: E402
from tests.helpers import assert_results_equal, results_to_rows  # noqa: E402


def random_config(seed):
    rng = random.Random(seed)
    molecules = random_tryptic_peptides(rng.randint(5, 60), seed=seed)
    if rng.random() < 0.5:
        molecules += averagine_formulas(rng.randint(3, 20), seed=seed, min_mass=200.0, max_mass=rng.choice([2000.0, 6000.0]))
    if rng.random() < 0.4:
        molecules += [m + "#Oxidation:1" for m in molecules[:5] if m[0] != "+"]
    charges = sorted(rng.sample([1, 2, 3, 4, 5], rng.randint(1, 4)))
    labels = rng.choice([None, {"15N": [0, 0.99]}, {"15N": [0, 0.5, 0.994]}, {"13C": [0, 0.98], "15N": [0, 0.99]}])
    fixed = rng.choice([None, {"C": ["C(2)H(3)14N(1)O(1)"]}])
    params = {"M_SCORE_THRESHOLD": rng.choice([0.3, 0.5, 0.7]), "REL_MZ_RANGE": rng.choice([5e-6, 1e-5, 5e-5]),
              "REL_I_RANGE": rng.choice([0.2, 0.5]), "MINIMUM_NUMBER_OF_MATCHED_ISOTOPOLOGUES": rng.choice([1, 2, 3]),
              "REQUIRED_PERCENTILE_PEAK_OVERLAP": rng.choice([0.2, 0.5, 0.8]), "MZ_SCORE_PERCENTILE": rng.choice([0.2, 0.4, 0.9]),
              "MAX_MOLECULES_PER_MATCH_BIN": rng.choice([1, 7, 20]), "MACHINE_OFFSET_IN_PPM": rng.choice([0.0, -2.0, 4.0]),
              "INTERNAL_PRECISION": rng.choice([1000, 1000, 10000])}
    wide = params["REL_MZ_RANGE"] >= 5e-5              # wide windows x dense spectra = combinatorial candidates for the oracle
    run = dict(n_spectra=rng.randint(4, 16), seed=seed + 1, mean_peaks=rng.choice([80, 300] if wide else [80, 300, 1200]),
               elution_sigma=rng.choice([1.0, 3.0]),
               entry_fraction=rng.choice([0.2, 0.6, 1.0]), ppm_sd=rng.choice([1.0, 3.0]), dropout=rng.choice([0.0, 0.2, 0.4]))
    return dict(molecules=molecules, charges=charges, metabolic_labels=labels, fixed_labels=fixed, params=params), run, rng.random() < 0.5


class TooSlow(Exception):
    pass


def _alarm(signum, frame):
    raise TooSlow()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    limit = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    signal.signal(signal.SIGALRM, _alarm)
    for seed in range(first, first + n):
        signal.alarm(limit)
        try:
            check(seed)
        except TooSlow:
            print("seed %d: skipped, the oracle needs more than %d s" % (seed, limit), flush=True)
        finally:
            signal.alarm(0)


def check(seed):
    kwargs, run, as_lists = random_config(seed)
    t0 = time.time()
    oracle = OracleLibrary(**kwargs)
    if len(oracle.formulas_sorted_by_mz) == 0:
        print("seed %d: empty library, skipped" % seed)
        return
    lib = pyqms_b200.IsotopologueLibrary(verbose=False, **kwargs)
    dense = seed % 2 == 1                                # odd seeds: the spectrum-resident matcher wherever the library allows it
    lib._ctx.set_tuning("dense_mode", 1 if dense else 0)
    peaks, offsets = synth_run(*entry_peaks(oracle), **run)
    want, got = {}, None
    spectra = [peaks[offsets[s]:offsets[s + 1]] for s in range(len(offsets) - 1)]
    if as_lists:
        spectra = [[(float(a), float(b)) for a, b in spec] for spec in spectra]
    for s, spec in enumerate(spectra):
        oracle.match_all(spec, "f", s, float(s), want)
        got = lib.match_all(mz_i_list=spec, file_name="f", spec_id=s, spec_rt=float(s), results=got)
    tie_prone = any(len(v) > 2 for v in (kwargs["metabolic_labels"] or {}).values()) or len(kwargs["metabolic_labels"] or {}) > 1
    assert_results_equal(results_to_rows(got), results_to_rows(want), score_tol=1e-6, cmz_rel=1e-12, ri_rel=1e-9, key_order=not tie_prone)
    hits = sum(len(v) for v in want.values())
    print("seed %d: ok  entries %d, %d spectra (%s), %d keys, %d hits, %d spectrum-resident batches, %.1f s" % (
        seed, len(oracle.formulas_sorted_by_mz), len(spectra), "lists" if as_lists else "numpy", len(want), hits,
        lib.device_counters()["dense_batches"], time.time() - t0), flush=True)


def check_matchers(seed):
    """Row-streaming against spectrum-resident matcher on a random dense library (no oracle: hundreds to thousands of
    peptides, label sets, window widths, crowded spectra, sub-batched host runs): packed hits must agree bit for bit."""
    rng = random.Random(seed)
    molecules = random_tryptic_peptides(rng.randint(300, 6000), seed=seed)
    charges = sorted(rng.sample([1, 2, 3, 4, 5], rng.randint(2, 5)))
    labels = rng.choice([None, None, {"15N": [0, 0.99]}, {"15N": [0, 0.2, 0.5, 0.8, 0.99]}])
    fixed = rng.choice([None, {"C": ["C(2)H(3)14N(1)O(1)"]}])
    params = {"M_SCORE_THRESHOLD": rng.choice([0.2, 0.5, 0.7]), "REL_MZ_RANGE": rng.choice([5e-6, 1e-5, 3e-5]),
              "MINIMUM_NUMBER_OF_MATCHED_ISOTOPOLOGUES": rng.choice([1, 2, 3]), "REQUIRED_PERCENTILE_PEAK_OVERLAP": rng.choice([0.2, 0.5, 0.8]),
              "MZ_SCORE_PERCENTILE": rng.choice([0.0, 0.4, 1.0]), "INTERNAL_PRECISION": rng.choice([1000, 1000, 10000]),
              "MIN_REL_PEAK_INTENSITY_FOR_MATCHING": rng.choice([0.01, 0.05])}
    lib = pyqms_b200.IsotopologueLibrary(molecules=molecules, charges=charges, metabolic_labels=labels, fixed_labels=fixed,
                                         params=params, verbose=False)
    mz, ci = lib.entry_peaks()
    peaks, offsets = synth_run(mz, ci, n_spectra=rng.randint(20, 300), seed=seed + 1, mean_peaks=rng.choice([300, 1500, 5000]),
                               max_peaks=rng.choice([5000, 20000]), elution_sigma=rng.choice([1.0, 4.0]),
                               entry_fraction=rng.choice([0.01, 0.1, 0.5]), ppm_sd=rng.choice([1.0, 3.0]), dropout=rng.choice([0.0, 0.3]))
    from pyqms_b200._capi import QmsError
    outs, refused = [], 0
    for mode, sub_rows in ((0, 0), (1, 0), (1, rng.choice([20000, 100000]))):
        lib._ctx.set_tuning("dense_mode", mode)
        lib._ctx.set_tuning("sub_batch_rows", sub_rows)
        before = lib.device_counters()["dense_batches"]
        try:
            outs.append(lib.match_packed(peaks, offsets))
        except QmsError as error:                       # wide label grids x crowded spectra: 2^32 combinations in one pair
            assert error.code == -5 and "candidate combinations" in str(error), error
            refused += 1
        used = lib.device_counters()["dense_batches"] - before
    if refused:
        assert refused == 3, "the matchers disagree on the combination limit"
        print("seed %d: both matchers refuse the batch (a pair above the combination limit)" % seed, flush=True)
        return
    a = outs[0]
    for b in outs[1:]:
        assert a["n_hits"] == b["n_hits"] and a["n_windows"] == b["n_windows"], (seed, a["n_hits"], b["n_hits"])
        for k in ("spec", "entry", "win_off", "peak"):
            assert np.array_equal(a[k], b[k]), (seed, k)
        for k in ("score", "scaling", "mmz", "mi"):
            assert np.array_equal(a[k], b[k], equal_nan=True), (seed, k)
    print("seed %d: ok  entries %d, windows %d, %d spectra, %d peaks, %d hits, %d spectrum-resident sub-batches in the last run" % (
        seed, lib._n_entries, lib._n_windows, len(offsets) - 1, len(peaks), a["n_hits"], used), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--matchers":
        count = int(sys.argv[2]) if len(sys.argv) > 2 else 20
        start = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
        for config_seed in range(start, start + count):
            check_matchers(config_seed)
    else:
        main()
