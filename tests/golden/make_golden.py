#!/usr/bin/env python
"""Generate the golden fixtures by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Imports /root/reference/pyqms with the parser stand-in of tests/golden/_standin,
builds every library of tests/golden/cases.py, draws the seeded synthetic spectra, runs the
reference's match_all in list mode and numpy mode, and writes
    tests/golden/<case>.json.gz   library dump + match results (floats via repr: exact)
    tests/golden/<case>.npz       the spectra that were matched
    tests/golden/kat.json         score_matches known answers + the 16 real matches of
                                  tests/data/test_BSA_pyqms_results.pkl (re-scored by the reference)
"""
import gzip
import json
import pickle
import sys
import warnings
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / "_standin"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(HERE.parent.parent))
sys.setrecursionlimit(20000)
warnings.filterwarnings("ignore")
import pyqms  # noqa: E402  the reference
from tests.golden.cases import CASES, QUICK_START_PEAKS  # noqa: E402
from tests.golden.spectra import draw_case_spectra  # noqa: E402


def f(x):
    return None if x is None else float(x)


def dump_library(lib, skip_trees=False, tree_bounds=False):
    formulas = list(lib.keys())
    lps = list(lib.labled_percentiles)
    env = {}
    for formula in formulas:
        rows = []
        for lp in lps:
            e = lib[formula]["env"][lp]
            rows.append({
                "mass": e["mass"], "abun": e["abun"], "relabun": e["relabun"], "c_peak_pos": e["c_peak_pos"],
                "n_c_peaks": e["n_c_peaks"],
                "z": {str(z): {"mz": e[z]["mz"],
                               "win": [None if s is None else [min(s), max(s), len(s)] for s in e[z]["tmzs"]],
                               "n_atmzs": len(e[z]["atmzs"])} for z in lib.charges}})
        env[formula] = rows
    trees = {}
    for el, by_label in ({} if skip_trees else lib.element_trees).items():
        trees[el] = {}
        for label, levels in by_label.items():
            out = {}
            for n, node in levels.items():
                lo, hi = node["minPos"], node["maxPos"]
                if lo is None:
                    out[str(n)] = None
                    continue
                span = range(min(lo, hi), max(lo, hi) + 1)
                out[str(n)] = [lo, hi, [f(node["env"][p]["abun"]) for p in span],
                               [f(node["env"][p]["mass"]) for p in span]]
            trees[el][label] = out
    bounds = {}
    if tree_bounds:      # minPos / maxPos of every level (what the pruning threshold decides), without the envelopes
        for el, by_label in lib.element_trees.items():
            bounds[el] = {label: {str(n): [node["minPos"], node["maxPos"]] for n, node in levels.items()}
                          for label, levels in by_label.items()}
    return {
        "tree_bounds": bounds,
        "formulas": formulas, "cc": {k: dict(lib[k]["cc"]) for k in formulas},
        "labled_percentiles": [[list(p) for p in lp] for lp in lps],
        "molecule_to_formula": lib.lookup["molecule to formula"],
        "fixed_label_variations": {k: sorted(v) for k, v in lib.lookup["molecule fixed label variations"].items()},
        "env": env, "trees": trees,
        "index": [[lo, hi, z, lps.index(lp), formulas.index(fo)] for lo, hi, z, lp, fo in lib.formulas_sorted_by_mz],
        "match_set_mz_range": lib.match_set_mz_range, "n_match_sets": len(lib.match_sets),
        "match_sets": [[ms["ids"], ms["mz_range"], len(ms["tmzs"])] for ms in lib.match_sets.values()],
        "isotopic_distributions": {el: {lab: [list(t) for t in d] for lab, d in v.items()}
                                   for el, v in lib.isotopic_distributions.items()
                                   if el in lib.element_trees},
        "fixed_label_levels": lib.params["FIXED_LABEL_ISOTOPE_ENRICHMENT_LEVELS"],
    }


def dump_results(res):
    out = []
    for key, val in res.items():
        rows = []
        for m in val["data"]:
            rows.append([m.spec_id, f(m.rt), f(m.score), f(m.scaling_factor),
                         [[f(a), f(b), f(c), f(d), int(e)] for a, b, c, d, e in m.peaks]])
        out.append([[key.file_name, key.formula, key.charge, [list(p) for p in key.label_percentiles]], rows])
    return out


def run_matches(lib, peaks, offsets, as_list):
    res = None
    for s in range(len(offsets) - 1):
        spec = peaks[offsets[s]:offsets[s + 1]]
        if as_list:
            spec = [(float(a), float(b)) for a, b in spec]
        res = lib.match_all(mz_i_list=spec, file_name="run", spec_id=s, spec_rt=s * 0.2, results=res)
    return res


def main():
    only = sys.argv[1:]
    for name, case in CASES.items():
        if only and name not in only:
            continue
        lib = pyqms.IsotopologueLibrary(verbose=False, **case["lib"])
        peaks, offsets = draw_case_spectra(lib, case["run"], name)
        blob = {"library": dump_library(lib, case.get("skip_trees", False), case.get("tree_bounds", False)),
                "results_numpy": dump_results(run_matches(lib, peaks, offsets, False)),
                "results_list": dump_results(run_matches(lib, peaks, offsets, True))}
        with gzip.open(HERE / (name + ".json.gz"), "wt") as fh:
            json.dump(blob, fh, separators=(",", ":"))
        np.savez_compressed(HERE / (name + ".npz"), peaks=peaks, offsets=offsets)
        print(name, "entries", len(lib.formulas_sorted_by_mz), "peaks", len(peaks),
              "hits numpy/list", sum(len(r[1]) for r in blob["results_numpy"]),
              sum(len(r[1]) for r in blob["results_list"]))
    if only:
        return
    # known answers of score_matches (IL:2441-2498), xi = 0.5 and 0.4
    lib = pyqms.IsotopologueLibrary(molecules=["ELVISLIVES"], charges=[2], verbose=False)
    rng = np.random.default_rng(99)
    kats = []
    hand = [  # tests/score_matches_tests.py:8-25 inputs
        [(1, 1, 0.9, 1, 1), (0.5, 1, 0.2, 0.5, 1)], [(1, 2, 0.9, 1, 1), (0.5, 2, 0.2, 0.5, 1)],
        [(1, 3, 0.9, 1, 1), (0.123, 123, 0.0, 0.5, 1)], [(1.1, 3, 0.9, 1, 1), (0.55, 3, 0.2, 0.5, 1)],
        [(1, 0, 0.9, 1, 1), (0.5, 0, 0.2, 0.5, 1)], [(1, 5, 0.9, 0.5, 1), (0.5, 5, 0.2, 0.2, 1)],
        [(1, 6, 1.0, 1, 1), (0.5, 6, 1.0, 0.5, 1), (None, None, 0.0, 0.5, 0.2)]]
    for peaks in hand:
        for xi in (0.5, 0.4):
            s, sc = lib.score_matches(peaks, xi)
            kats.append([peaks, xi, f(s), f(sc)])
    for _ in range(200):
        n = int(rng.integers(2, 9))
        cmz = 400 + np.cumsum(rng.uniform(0.3, 1.0, n))
        ri = np.sort(rng.uniform(0.01, 1, n))[::-1]
        ri[0] = 1.0
        ci = np.rint(ri * 60000).astype(int)
        amount = 10 ** rng.uniform(0, 3)
        peaks = []
        for j in range(n):
            if rng.random() < 0.25:
                peaks.append((None, None, float(ri[j]), float(cmz[j]), int(ci[j])))
            else:
                peaks.append((float(cmz[j] * (1 + rng.normal(0, 3e-6))), float(ci[j] * amount * (1 + rng.normal(0, 0.2))),
                              float(ri[j]), float(cmz[j]), int(ci[j])))
        if all(p[0] is None for p in peaks):
            continue
        xi = float(rng.choice([0.4, 0.5, 0.0, 1.0]))
        s, sc = lib.score_matches(peaks, xi)
        kats.append([peaks, xi, f(s), f(sc)])
    # the reference's pickled real-data Results fixture: 16 matches, re-scored by the reference
    real = pickle.load(open("/root/reference/tests/data/test_BSA_pyqms_results.pkl", "rb"))
    lib5 = pyqms.IsotopologueLibrary(molecules=["SHCIAEVEK"], charges=[1, 2, 3, 4, 5], metabolic_labels={"15N": [0]},
                                     fixed_labels={"C": ["C(2)H(3)14N(1)O(1)"]}, verbose=False)
    real_rows = []
    for key, val in real.items():
        for m in val["data"]:
            s, sc = lib5.score_matches(m.peaks, real.params["MZ_SCORE_PERCENTILE"])
            assert s == m.score and sc == m.scaling_factor      # reference reproduces its own fixture
            real_rows.append([[key.file_name, key.formula, key.charge, [list(p) for p in key.label_percentiles]],
                              m.spec_id, f(m.score), f(m.scaling_factor),
                              [[f(a), f(b), f(c), f(d), int(e)] for a, b, c, d, e in m.peaks]])
    (HERE / "kat.json").write_text(json.dumps({"score_matches": kats, "bsa_pickle_matches": real_rows,
                                               "bsa_pickle_xi": real.params["MZ_SCORE_PERCENTILE"]},
                                              separators=(",", ":")))
    print("kat.json:", len(kats), "score KATs,", len(real_rows), "real matches")


if __name__ == "__main__":
    main()
