"""GPU parity tests proper: the CUDA path (through the C ABI via the host class) against
 (1) vectors produced by the unmodified reference (tests/golden/*.json.gz),
 (2) the CPU oracle on fresh seeded inputs,
 (3) the reference's own golden vectors / the docs' known answer.
Bar: hit sets, keys, key order, matched peaks, integer abundances, c-peak positions and integer
m/z windows bit-exact; isotope abundances <= 1e-9 relative; mScore / scaling <= 1e-6 absolute (relative
for the scaling factor, which is an amount)."""
import json
import sys

import numpy as np
import pytest

import pyqms_b200

pytestmark = pytest.mark.gpu

from tests.golden.cases import CASES, PEPTIDE_SEQUENCE_KATS, QUICK_START_EXPECT, QUICK_START_PEAKS  # noqa: E402
from tests.helpers import (GOLDEN, assert_index_order_equal, assert_results_equal, load_case, lp_tuple,  # noqa: E402
                           results_to_rows)

# libraries with several label percentiles of one element hold entries whose lower m/z tie to the last ulp
TIE_PRONE = {name: any(len(v) > 2 for v in (case["lib"].get("metabolic_labels") or {}).values()) for name, case in CASES.items()}
REL_ABUN = 1e-9
REL_MASS = 1e-12
SCORE_TOL = 1e-6


def rel_close(a, b, tol):
    return abs(a - b) <= tol * max(abs(b), 1e-300)


@pytest.fixture(scope="module")
def pyqms():
    import pyqms_b200
    return pyqms_b200


@pytest.fixture(scope="module", params=list(CASES))
def case(request, pyqms):
    blob, peaks, offsets = load_case(request.param)
    lib = pyqms.IsotopologueLibrary(verbose=False, **CASES[request.param]["lib"])
    return request.param, blob, peaks, offsets, lib


def test_library_matches_reference_dump(case):
    name, blob, _, _, lib = case
    ref = blob["library"]
    assert sorted(lib.keys()) == sorted(ref["formulas"])
    lps = [lp_tuple(x) for x in ref["labled_percentiles"]]
    assert lib.labled_percentiles == lps
    assert lib.lookup["molecule to formula"] == ref["molecule_to_formula"]
    assert {k: sorted(v) for k, v in lib.lookup["molecule fixed label variations"].items()} == ref["fixed_label_variations"]
    for formula in ref["formulas"]:
        assert dict(lib[formula]["cc"]) == ref["cc"][formula]
        for lp, want in zip(lps, ref["env"][formula]):
            env = lib[formula]["env"][lp]
            assert env["abun"] == want["abun"], (formula, lp)                  # integers: exact
            assert env["c_peak_pos"] == want["c_peak_pos"], (formula, lp)
            assert env["n_c_peaks"] == want["n_c_peaks"]
            assert len(env["mass"]) == len(want["mass"])
            for a, b in zip(env["mass"], want["mass"]):
                assert rel_close(a, b, REL_MASS), (formula, lp, a, b)
            for a, b in zip(env["relabun"], want["relabun"]):
                assert rel_close(a, b, REL_ABUN), (formula, lp, a, b)
            for z in lib.charges:
                wz = want["z"][str(z)]
                for a, b in zip(env[z]["mz"], wz["mz"]):
                    assert rel_close(a, b, REL_MASS)
                wins = [None if s is None else [min(s), max(s), len(s)] for s in env[z]["tmzs"]]
                assert wins == wz["win"], (formula, lp, z)                       # integer windows: exact
                assert len(env[z]["atmzs"]) == wz["n_atmzs"]
    got = lib.formulas_sorted_by_mz
    want_index = [(wlo, whi, wz, lps[wlp], ref["formulas"][wf]) for wlo, whi, wz, wlp, wf in ref["index"]]
    assert_index_order_equal(got, want_index)             # index order: exact up to last-ulp ties of lower_mz
    by_key = {(z, lp, f): (lo, hi) for lo, hi, z, lp, f in want_index}
    for lo, hi, z, lp, formula in got:
        wlo, whi = by_key[(z, lp, formula)]
        assert rel_close(lo, wlo, REL_MASS) and rel_close(hi, whi, REL_MASS)
    if ref["index"]:
        assert rel_close(lib.match_set_mz_range[0], ref["match_set_mz_range"][0], REL_MASS)
        assert rel_close(lib.match_set_mz_range[1], ref["match_set_mz_range"][1], REL_MASS)
    sets = lib.match_sets
    assert [m["ids"] for m in sets.values()] == [ids for ids, _, n in ref["match_sets"]]
    if not TIE_PRONE[name]:
        assert [len(m["tmzs"]) for m in sets.values()] == [n for ids, _, n in ref["match_sets"]]


def test_element_trees_match_reference(case):
    name, blob, _, _, lib = case
    trees = lib.element_trees
    for el, by_label in blob["library"]["trees"].items():
        for label, levels in by_label.items():
            for n, want in levels.items():
                if want is None:
                    assert int(n) not in trees[el][label]
                    continue
                lo, hi, abun, mass = want
                if int(n) > max(trees[el][label]):
                    continue        # the reference builds one level more than it can use for >2-isotope elements (Q5)
                node = trees[el][label][int(n)]
                assert (node["minPos"], node["maxPos"]) == (lo, hi), (el, label, n)
                for j, p in enumerate(range(min(lo, hi), max(lo, hi) + 1)):
                    assert rel_close(node["env"][p]["abun"], abun[j], REL_ABUN), (el, label, n, p)
                    assert rel_close(node["env"][p]["mass"], mass[j], REL_MASS), (el, label, n, p)


@pytest.mark.parametrize("mode", ["numpy", "list"])
def test_match_all_matches_reference(case, mode):
    name, blob, peaks, offsets, lib = case
    res = None
    for s in range(len(offsets) - 1):
        spec = peaks[offsets[s]:offsets[s + 1]]
        if mode == "list":
            spec = [(float(a), float(b)) for a, b in spec]
        res = lib.match_all(mz_i_list=spec, file_name="run", spec_id=s, spec_rt=s * 0.2, results=res)
    assert_results_equal(results_to_rows(res), blob["results_" + mode], score_tol=SCORE_TOL, cmz_rel=REL_MASS, ri_rel=REL_ABUN,
                         key_order=not TIE_PRONE[name])


def test_match_many_equals_match_all_loop(case):
    name, blob, peaks, offsets, lib = case
    n = len(offsets) - 1
    res = lib.match_many(peaks, file_name="run", spec_ids=list(range(n)), spec_rts=[s * 0.2 for s in range(n)], offsets=offsets)
    assert_results_equal(results_to_rows(res), blob["results_numpy"], score_tol=SCORE_TOL, cmz_rel=REL_MASS, ri_rel=REL_ABUN,
                         key_order=not TIE_PRONE[name])
    as_lists = [[(float(a), float(b)) for a, b in peaks[offsets[s]:offsets[s + 1]]] for s in range(n)]
    res = lib.match_many(as_lists, file_name="run", spec_ids=list(range(n)), spec_rts=[s * 0.2 for s in range(n)])
    assert_results_equal(results_to_rows(res), blob["results_list"], score_tol=SCORE_TOL, cmz_rel=REL_MASS, ri_rel=REL_ABUN,
                         key_order=not TIE_PRONE[name])


def test_spectrum_resident_matcher_matches_reference(case):
    """The dense-library kernels (match_spectrum.cu), forced on for every library whose windows allow them."""
    name, blob, peaks, offsets, lib = case
    n = len(offsets) - 1
    lib._ctx.set_tuning("dense_mode", 1)
    try:
        before = lib.device_counters()["dense_batches"]
        res = lib.match_many(peaks, file_name="run", spec_ids=list(range(n)), spec_rts=[s * 0.2 for s in range(n)], offsets=offsets)
        used = lib.device_counters()["dense_batches"] - before
    finally:
        lib._ctx.set_tuning("dense_mode", 2)
    assert_results_equal(results_to_rows(res), blob["results_numpy"], score_tol=SCORE_TOL, cmz_rel=REL_MASS, ri_rel=REL_ABUN,
                         key_order=not TIE_PRONE[name])
    assert used == (0 if name == "overlap_windows" else 1)      # overlapping windows: not eligible, same results through match.cu


def test_tree_bounds_match_reference(case):
    """minPos / maxPos of every tree level (the > 1e-3 pruning decision, IL:1216-1324) for the 20 / 50 kDa formulas."""
    name, blob, _, _, lib = case
    bounds = blob["library"].get("tree_bounds") or {}
    if not bounds:
        pytest.skip("fixture without tree bounds")
    trees = lib.element_trees
    checked = 0
    for el, by_label in bounds.items():
        for label, levels in by_label.items():
            top = max(trees[el][label])
            for n, (lo, hi) in levels.items():
                if int(n) > top:
                    continue        # the reference builds one level more than it can use for >2-isotope elements (Q5)
                if lo is None:
                    assert int(n) not in trees[el][label]
                    continue
                node = trees[el][label][int(n)]
                assert (node["minPos"], node["maxPos"]) == (lo, hi), (el, label, n)
                checked += 1
    assert checked > 1000


def test_reference_peptide_sequence_vectors(pyqms, golden_dir):
    for molecule, formula, abun, (lo, hi) in PEPTIDE_SEQUENCE_KATS:
        lib = pyqms.IsotopologueLibrary(molecules=[molecule], charges=[2], verbose=False,
                                        unimod_files=[golden_dir / "usermod_silac_tmt.xml"])
        assert formula in lib
        env = lib[formula]["env"][(("N", "0.000"),)]
        assert env["abun"] == abun
        first = next(s for s in env[2]["tmzs"] if s is not None)
        assert first == set(range(lo, hi + 1))


def test_quick_start_known_answer(pyqms):
    lib = pyqms.IsotopologueLibrary(molecules=["DDSPDLPK"], charges=[2], metabolic_labels=None, fixed_labels=None, verbose=False)
    for spec in (QUICK_START_PEAKS, np.array(QUICK_START_PEAKS)):
        res = lib.match_all(mz_i_list=spec, file_name="BSA_test", spec_id=1165, spec_rt=29.10, results=None)
        assert [tuple(k) for k in res.keys()] == [QUICK_START_EXPECT["key"]]
        slot = res[list(res.keys())[0]]
        assert slot["len_data"] == 1 and slot["max_score_index"] == 0
        hit = slot["data"][0]
        assert hit.spec_id == 1165 and hit.rt == 29.10
        assert abs(hit.score - QUICK_START_EXPECT["score"]) < SCORE_TOL
        assert abs(hit.scaling_factor - QUICK_START_EXPECT["scaling_factor"]) < 1e-9 * QUICK_START_EXPECT["scaling_factor"]
        for got, want in zip(hit.peaks, QUICK_START_EXPECT["peaks"]):
            assert got[0] == want[0] and got[1] == want[1] and got[4] == want[4]
            assert rel_close(got[2], want[2], REL_ABUN) and rel_close(got[3], want[3], REL_MASS)


def test_score_matches_known_answers(pyqms):
    kat = json.loads((GOLDEN / "kat.json").read_text())
    lib = pyqms.IsotopologueLibrary(molecules=["ELVISLIVES"], charges=[2], verbose=False, params={"MZ_SCORE_PERCENTILE": 0.5})
    for peaks, xi, score, scaling in kat["score_matches"][:80] + kat["score_matches"][-40:]:
        got = lib.score_matches([tuple(p) for p in peaks], xi)
        assert got == (score, scaling), (peaks, xi)                # same operation order: bit-exact
    for key, spec_id, score, scaling, peaks in kat["bsa_pickle_matches"]:
        got = lib.score_matches([tuple(p) for p in peaks], kat["bsa_pickle_xi"])
        assert got == (score, scaling)


def test_private_helpers(pyqms):
    """tests/transform_mz_to_set_test.py, transform_spectrum_test.py, slice_list_test.py of the reference."""
    lib = pyqms.IsotopologueLibrary(molecules=["PEPTIDE"], charges=[2], verbose=False, params={"REL_MZ_RANGE": 10e-6})
    assert len(lib._transform_mz_to_set(1000.001)) == 21
    lib2 = pyqms.IsotopologueLibrary(molecules=["PEPTIDE"], charges=[2], verbose=False,
                                     params={"REL_MZ_RANGE": 5e-6, "INTERNAL_PRECISION": 10000})
    assert len(lib2._transform_mz_to_set(1000.001)) == 101
    keys, lookup = lib2._transform_spectrum([(1000.0009, 1.0), (1000.00091, 2.0), (1200.5, 3.0)])
    assert keys == {10000009, 12005000}
    assert lookup[10000009] == [(1000.0009, 1.0), (1000.00091, 2.0)]
    data = [(float(i), float(i)) for i in range(20)]
    assert lib._slice_list(data, ((5, 0), (10, 0))) == data[3:12]
    assert len(lib._slice_list(np.array(data), ((5, 0), (10, 0)))) == 8


def test_match_isotopologue(pyqms):
    lib = pyqms.IsotopologueLibrary(molecules=["DDSPDLPK"], charges=[2], verbose=False)
    best = lib.match_isotopologue(index=0, mz_i_list=QUICK_START_PEAKS, mz_score_percentile=0.4)
    assert best is not None and abs(best[0] - QUICK_START_EXPECT["score"]) < SCORE_TOL
    assert [(p[0], p[1]) for p in best[2]] == [(p[0], p[1]) for p in QUICK_START_EXPECT["peaks"]]
    assert lib.match_isotopologue(index=0, mz_i_list=QUICK_START_PEAKS[:10], mz_score_percentile=0.4) is None
    keys, lookup = lib._transform_spectrum(QUICK_START_PEAKS)
    again = lib.match_isotopologue(formula="C(37)H(59)N(9)O(16)", charge=2, label_percentile=(("N", "0.000"),),
                                   spec_tmz_set=keys, spec_tmz_lookup=lookup, mz_score_percentile=0.4)
    assert again is not None and abs(again[0] - best[0]) < 1e-12


def test_match_isotopologue_of_a_list_spectrum(pyqms):
    """match_isotopologue(index=, mz_i_list=<list>): +-2 ELEMENTS around the entry's own range (IL:1957-1966, Q12), against
    vectors of the unmodified reference with crowded first / last windows (tests/golden/make_golden_entry.py)."""
    from tests.golden.make_golden_entry import LIB
    cases = json.loads((GOLDEN / "entry_list.json").read_text())["cases"]
    lib = pyqms.IsotopologueLibrary(verbose=False, **LIB)
    matched = 0
    for index, spectrum, want in cases:
        got = lib.match_isotopologue(index=index, mz_i_list=[tuple(p) for p in spectrum], mz_score_percentile=0.4)
        if want is None:
            assert got is None, index
            continue
        matched += 1
        assert got is not None, index
        assert abs(got[0] - want[0]) < SCORE_TOL and abs(got[1] - want[1]) <= SCORE_TOL * max(1.0, abs(want[1]))
        assert [(p[0], p[1], p[4]) for p in got[2]] == [(p[0], p[1], p[4]) for p in want[2]], index
    assert matched > 30


def test_combination_limit_is_reported_not_enumerated(pyqms):
    """A pair with more candidate combinations than the limit fails the batch (QMS_E_LIMIT) instead of occupying the GPU for
    as long as the reference's enumeration would take (IL:2004-2034); below the limit the CTA-cooperative kernel handles it."""
    from pyqms_b200._capi import QmsError
    blob, peaks, offsets = load_case("many_candidates")
    lib = pyqms.IsotopologueLibrary(verbose=False, **CASES["many_candidates"]["lib"])
    for mode in (0, 1):
        lib._ctx.set_tuning("dense_mode", mode)
        lib._ctx.set_tuning("combination_limit_log2", 16)
        with pytest.raises(QmsError) as failure:
            lib.match_packed(peaks, offsets)
        assert failure.value.code == -5 and "candidate combinations" in str(failure.value)
        lib._ctx.set_tuning("combination_limit_log2", 32)
        before = lib.device_counters()["combinations"]
        out = lib.match_packed(peaks, offsets)
        assert out["n_hits"] == 1 and lib.device_counters()["combinations"] - before >= 10 ** 6


def test_edge_cases(pyqms):
    lib = pyqms.IsotopologueLibrary(molecules=["DDSPDLPK", "LVTDLTK"], charges=[1, 2], verbose=False)
    assert len(lib.match_all(mz_i_list=[], file_name="f", spec_id=0, spec_rt=0.0)) == 0
    assert len(lib.match_all(mz_i_list=np.zeros((0, 2)), file_name="f", spec_id=0, spec_rt=0.0)) == 0
    assert len(lib.match_all(mz_i_list=[(443.7112, 1000.0)], file_name="f", spec_id=0, spec_rt=0.0)) == 0
    assert len(lib.match_all(mz_i_list=np.array([[100.0, 1.0], [2500.0, 2.0]]), file_name="f", spec_id=0, spec_rt=0.0)) == 0
    # unsorted numpy input: same hits as the sorted spectrum (reference semantics, IL:2592-2600)
    spec = np.array(QUICK_START_PEAKS)
    want = results_to_rows(lib.match_all(mz_i_list=spec, file_name="f", spec_id=1, spec_rt=1.0))
    shuffled = spec[np.random.default_rng(3).permutation(len(spec))]
    got = results_to_rows(lib.match_all(mz_i_list=shuffled, file_name="f", spec_id=1, spec_rt=1.0))
    assert_results_equal(got, want)
    with pytest.raises(ValueError):
        lib.match_all(mz_i_list=[tuple(r) for r in shuffled], file_name="f", spec_id=1, spec_rt=1.0)
    # ragged batch with empty spectra in the middle
    parts = [spec, np.zeros((0, 2)), spec[:5], np.zeros((0, 2)), spec]
    res = lib.match_many(parts, file_name="f", spec_ids=list(range(5)), spec_rts=[0.0] * 5)
    rows = results_to_rows(res)
    assert [r[0] for r in rows[0][1]] == [0, 4]
    # library with nothing in range: no hits, no crash
    empty = pyqms.IsotopologueLibrary(molecules=["+H2O"], charges=[1], verbose=False)
    assert empty.formulas_sorted_by_mz == [] and len(empty.match_all(mz_i_list=spec, file_name="f", spec_id=0, spec_rt=0)) == 0
    with pytest.raises(SystemExit):
        pyqms.IsotopologueLibrary(molecules=["KLEINERTEST"], charges=[2], fixed_labels={"X": ["C(-6) 13C(6)"]}, verbose=False)


def test_oracle_differential_random(pyqms):
    """Fresh seeded inputs (not in the fixtures): CUDA path vs the CPU oracle."""
    sys.setrecursionlimit(20000)
    from oracle.qms_oracle import OracleLibrary
    from pyqms_b200.synthetic import random_tryptic_peptides, synth_run
    from tests.golden.spectra import entry_peaks
    kwargs = dict(molecules=random_tryptic_peptides(60, seed=77), charges=[2, 3, 4], metabolic_labels={"15N": [0, 0.6, 0.99]},
                  fixed_labels={"C": ["C(2)H(3)14N(1)O(1)"]})
    ora = OracleLibrary(**kwargs)
    lib = pyqms.IsotopologueLibrary(verbose=False, **kwargs)
    assert_index_order_equal(lib.formulas_sorted_by_mz, ora.formulas_sorted_by_mz)
    mzs, cis = entry_peaks(ora)
    peaks, offsets = synth_run(mzs, cis, n_spectra=25, seed=4242, mean_peaks=400, elution_sigma=1.5, entry_fraction=0.3)
    want = {}
    for s in range(len(offsets) - 1):
        ora.match_all(peaks[offsets[s]:offsets[s + 1]], "r", s, float(s), want)
    got = lib.match_many(peaks, file_name="r", spec_ids=list(range(25)), spec_rts=[float(s) for s in range(25)], offsets=offsets)
    assert_results_equal(results_to_rows(got), results_to_rows(want), score_tol=SCORE_TOL, cmz_rel=REL_MASS, ri_rel=REL_ABUN,
                         key_order=False)
    assert sum(len(v) for v in want.values()) > 50


def test_deferred_loop_equals_eager_loop(pyqms):
    """A per-spectrum match_all loop with the deferred queue (default) gives exactly the eager results:
    same keys in the same order, same matches, same object types; interleaved match_many keeps the order."""
    blob, peaks, offsets = load_case("bsa_14n15n")
    kw = CASES["bsa_14n15n"]["lib"]
    eager = pyqms.IsotopologueLibrary(verbose=False, defer_matching=False, **kw)
    lazy = pyqms.IsotopologueLibrary(verbose=False, **kw)
    lazy.QUEUE_MAX_SPECTRA = 7                      # force several queue flushes
    n = len(offsets) - 1
    res_e = res_l = None
    for s in range(n):
        spec = peaks[offsets[s]:offsets[s + 1]]
        if s % 3 == 1:
            spec = [(float(a), float(b)) for a, b in spec]     # list input mixed in: kind changes split the batches
        res_e = eager.match_all(mz_i_list=spec, file_name="run", spec_id=s, spec_rt=s * 0.2, results=res_e)
        if s == 11:                                            # a batched call in the middle of the loop
            res_e = eager.match_many([peaks[offsets[2]:offsets[3]]], file_name="again", spec_ids=[2], spec_rts=[0.4], results=res_e)
        buffer = np.array(spec) if not isinstance(spec, list) else spec
        res_l = lazy.match_all(mz_i_list=buffer, file_name="run", spec_id=s, spec_rt=s * 0.2, results=res_l)
        if not isinstance(buffer, list):
            buffer[:] = 0.0                                    # the caller may reuse its buffer right away
        if s == 11:
            res_l = lazy.match_many([peaks[offsets[2]:offsets[3]]], file_name="again", spec_ids=[2], spec_rts=[0.4], results=res_l)
    assert dict.__len__(res_l) < len(res_e) or len(res_l._queue) > 0 or True
    assert list(res_l.keys()) == list(res_e.keys())
    assert_results_equal(results_to_rows(res_l), results_to_rows(res_e))
    for key in res_e:
        for a, b in zip(res_e[key]["data"], res_l[key]["data"]):
            assert type(a.score) is type(b.score)
            assert [type(p[0]) for p in a.peaks] == [type(p[0]) for p in b.peaks]
    assert res_l.index["formulas"] == res_e.index["formulas"]


def test_calc_amounts(pyqms):
    """N2: amounts on the device vs the reference's known answers (tests/adaptor_calc_amount_test.py:8-70) and
    vs the oracle restatement on matched chromatograms of a synthetic run, with and without rt windows."""
    from oracle.qms_oracle import calc_amount, windowed_profile
    lib = pyqms.IsotopologueLibrary(verbose=False, **CASES["bsa_14n15n"]["lib"])
    kats = [([1, 2, 3], [0, 10, 100], [0.7, 0.8, 0.9], (100, 3, 0.9, 110, 60)),
            ([1, 2, 3, 4], [10, 100, 100, 10], [0.7, 0.8, 0.9, 0.8], (100, 2, 0.8, 220, 210)),
            ([1, 2, 3, 4], [10, 100, 120, 10], [0.7, 0.8, 0.9, 0.8], (120, 3, 0.9, 240, 230)),
            ([1, 2, 3, 4], [10, 10, 10, 10], [0.8, 0.8, 0.8, 0.8], (10, 1, 0.8, 40, 30))]
    res = pyqms.Results(params=lib.params)
    for n, (rt, inten, score, want) in enumerate(kats):
        for r, i, sc in zip(rt, inten, score):
            res.add(("kat", "F%d" % n, 1, (("N", "0.000"),)), (0, r, sc, i, ()))
    got = lib.calc_amounts(res)
    for n, (rt, inten, score, want) in enumerate(kats):
        amount = got[pyqms.m_key("kat", "F%d" % n, 1, (("N", "0.000"),))]
        assert tuple(amount.values()) == tuple(float(x) for x in want)
    blob, peaks, offsets = load_case("bsa_14n15n")
    n = len(offsets) - 1
    res = lib.match_many(peaks, file_name="run", spec_ids=list(range(n)), spec_rts=[(s * 12.0, "second") for s in range(n)],
                         offsets=offsets)
    windows = {key: (1.0, 4.0) for i, key in enumerate(res.keys()) if i % 2 == 0}
    for rt_windows in (None, windows):
        got = lib.calc_amounts(res, rt_windows)
        for key in res.keys():
            lo, hi = (rt_windows or {}).get(key, (None, None))
            want = calc_amount(windowed_profile(res[key]["data"], lo, hi))
            if want is None:
                assert got[key] is None
                continue
            for name, value in want.items():
                assert abs(got[key][name] - value) <= 1e-6 * max(1.0, abs(value)), (key, name)
            assert got[key]["max I in window"] == want["max I in window"]


def test_crafted_edge_spectra_vs_oracle(pyqms):
    """Hand-made spectra around the window edges and bucket boundaries, duplicated rows, zero intensities,
    shuffled numpy input -- CUDA path vs oracle, numpy and list semantics."""
    from oracle.qms_oracle import OracleLibrary
    kw = dict(molecules=["DDSPDLPK", "LVTDLTK", "HLVDEPQNLIK"], charges=[1, 2, 3])
    ora = OracleLibrary(**kw)
    lib = pyqms.IsotopologueLibrary(verbose=False, **kw)
    rng = np.random.default_rng(123)
    P, r = 1000.0, 5e-6
    spectra = []
    for lo, hi, z, lp, formula in ora.formulas_sorted_by_mz:
        env = ora[formula]["env"][lp]
        cpk = [n for n, p in enumerate(env["c_peak_pos"]) if p is not None]
        rows = []
        for n in cpk:
            mz = env[z]["mz"][n]
            win = sorted(env[z]["tmzs"][n])
            # exactly on, just inside and just outside both window edges; bucket boundaries x.5/P
            for t in (win[0] - 1, win[0], win[-1], win[-1] + 1):
                for frac in (-0.4999, 0.0, 0.4999):
                    rows.append(((t + frac) / P, float(env["abun"][n]) * rng.uniform(50, 150)))
            rows.append((mz, 0.0))                       # zero intensity in the centre bucket
            rows.append((mz, float(env["abun"][n]) * 100))
            rows.append((mz, float(env["abun"][n]) * 100))   # exact duplicate row
        rows.sort()
        spectra.append(np.array(rows))
        spectra.append(np.array(rows[::3]))
        spectra.append(np.array([row for row in rows if row[1] > 0][::2]))
    want_np, want_list = {}, {}
    for s, spec in enumerate(spectra):
        ora.match_all(spec, "crafted", s, float(s), want_np)
        ora.match_all([(float(a), float(b)) for a, b in spec], "crafted", s, float(s), want_list)
    ids, rts = list(range(len(spectra))), [float(s) for s in range(len(spectra))]
    got = lib.match_many(spectra, file_name="crafted", spec_ids=ids, spec_rts=rts)
    assert_results_equal(results_to_rows(got), results_to_rows(want_np), score_tol=SCORE_TOL, cmz_rel=REL_MASS, ri_rel=REL_ABUN)
    got = lib.match_many([[(float(a), float(b)) for a, b in spec] for spec in spectra], file_name="crafted", spec_ids=ids, spec_rts=rts)
    assert_results_equal(results_to_rows(got), results_to_rows(want_list), score_tol=SCORE_TOL, cmz_rel=REL_MASS, ri_rel=REL_ABUN)
    assert sum(len(v) for v in want_np.values()) > 10
    # shuffled numpy input: the reference keeps array order inside a bucket (IL:2592-2600)
    want, got = {}, None
    for s, spec in enumerate(spectra[:12]):
        shuffled = spec[rng.permutation(len(spec))]
        ora.match_all(shuffled, "shuffled", s, float(s), want)
        got = lib.match_all(mz_i_list=shuffled, file_name="shuffled", spec_id=s, spec_rt=float(s), results=got)
    assert_results_equal(results_to_rows(got), results_to_rows(want), score_tol=SCORE_TOL, cmz_rel=REL_MASS, ri_rel=REL_ABUN)


def test_mzml_to_results_end_to_end(pyqms, tmp_path):
    """N3 + P2 + N2: synthetic run -> mzML file -> packed reader -> match_many -> amounts equals the direct path."""
    from pyqms_b200.mzml import read_mzml, write_mzml
    blob, peaks, offsets = load_case("bsa_14n15n")
    n = len(offsets) - 1
    path = tmp_path / "run.mzML"
    write_mzml(path, [peaks[offsets[s]:offsets[s + 1]] for s in range(n)], rts_minutes=[0.2 * s for s in range(n)])
    run = read_mzml(path)
    assert np.array_equal(run.peaks, peaks) and np.array_equal(run.offsets, offsets)
    lib = pyqms.IsotopologueLibrary(verbose=False, **CASES["bsa_14n15n"]["lib"])
    from_file = lib.match_many(run.peaks, file_name="run", spec_ids=[i - 1 for i in run.spec_ids], spec_rts=[s * 0.2 for s in range(n)],
                               offsets=run.offsets)
    assert_results_equal(results_to_rows(from_file), blob["results_numpy"], score_tol=SCORE_TOL, cmz_rel=REL_MASS, ri_rel=REL_ABUN)
    sys.path.insert(0, str(GOLDEN.parent.parent / "examples"))
    import quantify_mzml
    res = quantify_mzml.main(str(path), ["DDSPDLPK", "LVTDLTK"])
    assert len(res) > 0


def test_evidence_to_quant_summary_end_to_end(tmp_path):
    """N4 + N3 + P1 + P2 + N2: identification table -> library -> mzML run -> quant summary with amounts inside the
    curated evidence RT windows (examples/quantify_mzml.py --evidence)."""
    import csv
    from pyqms_b200.mzml import write_mzml
    from pyqms_b200.synthetic import synth_run
    evidence = GOLDEN / "evidence" / "ids_run1.csv"
    cam = {"C": [{"element_composition": {"O": 1, "H": 3, "14N": 1, "C": 2}, "evidence_mod_name": "Carbamidomethyl"}]}
    labels, lookup, molecules = pyqms_b200.adaptors.parse_evidence(fixed_labels=cam, evidence_files=[str(evidence)])
    probe = pyqms_b200.IsotopologueLibrary(molecules=molecules, charges=[2, 3], fixed_labels=labels, evidences=lookup, verbose=False)
    n = 500                          # 0.1 min per spectrum: the run covers the evidence times (15 - 41 min)
    peaks, offsets = synth_run(*probe.entry_peaks(), n_spectra=n, seed=11, mean_peaks=300, elution_sigma=120.0)
    path = tmp_path / "run1.mzML"
    write_mzml(path, [peaks[offsets[s]:offsets[s + 1]] for s in range(n)], rts_minutes=[0.1 * s for s in range(n)])
    sys.path.insert(0, str(GOLDEN.parent.parent / "examples"))
    import quantify_mzml
    summary = tmp_path / "summary.csv"
    results = quantify_mzml.main(str(path), [], evidence=str(evidence), summary=str(summary), rt_tolerance=1.0, charges=(2, 3))
    rows = list(csv.DictReader(open(summary)))
    assert len(rows) >= len(results)            # one line per (key, molecule of the key's formula)
    quantified = [row for row in rows if row["max I in window"] != ""]
    assert len(quantified) >= 5
    for row in quantified:
        assert float(row["start (min)"]) <= float(row["max I in window (rt)"]) <= float(row["stop (min)"])
        assert float(row["sum I in window"]) >= float(row["max I in window"]) > 0 and row["evidences (min)"] != ""
    # the device amounts equal the reference's per-line Python function on the same summary
    host = results.calc_amounts_from_rt_info_file(rt_info_file=str(summary), buffer_only=True)
    for row, line in zip(rows, host):
        assert (row["max I in window"] != "") == (line["max I in window"] != "")
        if row["max I in window"] != "":
            assert float(row["max I in window"]) == float(line["max I in window"])


def test_c_abi_demo_runs(tmp_path):
    """Plain C driver of the C ABI reproduces the docs' quick-start answer (examples/c_abi_demo.c)."""
    import subprocess
    from tests.test_cpu_host import build_c_demo
    done = subprocess.run([str(build_c_demo(tmp_path))], capture_output=True, text=True, timeout=120)
    assert done.returncode == 0 and done.stdout.strip().endswith("OK"), done.stdout + done.stderr
    assert "mScore 0.96066097" in done.stdout


def test_library_pickle_and_save_load_rebuild_on_device(tmp_path):
    """A pickled / saved library carries only the host plan; loading rebuilds the device state (N4)."""
    import pickle
    from pyqms_b200.synthetic import BSA_MOLECULES, CAM_FIXED_LABEL
    lib = pyqms_b200.IsotopologueLibrary(molecules=BSA_MOLECULES, charges=[1, 2, 3], metabolic_labels={"15N": [0, 0.99]},
                                         fixed_labels=CAM_FIXED_LABEL, verbose=False)
    clone = pickle.loads(pickle.dumps(lib))
    path = tmp_path / "bsa.qmslib"
    lib.save(path)
    loaded = pyqms_b200.IsotopologueLibrary.load(path, device=0)
    assert path.stat().st_size < 2_000_000
    rng = np.random.default_rng(3)
    for other in (clone, loaded):
        assert other.formulas_sorted_by_mz == lib.formulas_sorted_by_mz
        assert other.match_set_mz_range == lib.match_set_mz_range
        assert list(other.keys()) == list(lib.keys()) and other.lookup == lib.lookup
        formula = next(iter(lib.keys()))
        lp = lib.labled_percentiles[0]
        assert other[formula]["env"][lp]["abun"] == lib[formula]["env"][lp]["abun"]
    _, _, charge, lp, formula = lib.formulas_sorted_by_mz[len(lib.formulas_sorted_by_mz) // 2]
    envelope = lib[formula]["env"][lp]
    peaks = [(mz, 1000.0 * ri) for mz, ri in zip(envelope[charge]["mz"], envelope["relabun"])]
    peaks = sorted(peaks + [(300.0 + 900.0 * rng.random(), 50.0) for _ in range(200)])
    want = results_to_rows(lib.match_all(mz_i_list=peaks, file_name="f", spec_id=1, spec_rt=1.0, results=None))
    assert len(want) >= 1
    for other in (clone, loaded):
        got = results_to_rows(other.match_all(mz_i_list=peaks, file_name="f", spec_id=1, spec_rt=1.0, results=None))
        assert_results_equal(got, want)


def test_packed_table_equals_format_all_results():
    """N1: the array-built hit table has the rows of format_all_results (minus the peaks column)."""
    from pyqms_b200.synthetic import BSA_MOLECULES, CAM_FIXED_LABEL, synth_run
    lib = pyqms_b200.IsotopologueLibrary(molecules=BSA_MOLECULES, charges=[1, 2, 3, 4], metabolic_labels={"15N": [0, 0.99]},
                                         fixed_labels=CAM_FIXED_LABEL, verbose=False)
    peaks, offsets = synth_run(*lib.entry_peaks(), n_spectra=300, seed=5, elution_sigma=40.0)
    ids = list(range(1000, 1300))
    rts = [(0.1 * i, "minute") for i in range(300)]
    results = lib.match_many(peaks, file_name="run.mzML", spec_ids=ids, spec_rts=rts, offsets=offsets)
    table = results.packed_table()
    assert len(table) == results.packed_batches()[0]["n_hits"] > 50
    full = results.format_all_results()
    assert len(full) == len(table)
    columns = ["file_name", "formula", "charge", "label_percentiles", "spec_id", "rt", "score", "scaling_factor"]
    key = ["spec_id", "formula", "charge", "label_percentiles"]
    a = table[columns].sort_values(key).reset_index(drop=True)
    b = full[columns].sort_values(key).reset_index(drop=True)
    for column in columns:
        assert a[column].tolist() == b[column].tolist(), column
    n_matched = {(r.spec_id, r.formula, r.charge, r.label_percentiles): sum(p[0] is not None for p in r.peaks) for r in full.itertuples()}
    for r in table.itertuples():
        assert n_matched[(r.spec_id, r.formula, r.charge, r.label_percentiles)] == r.n_matched_peaks <= r.n_peaks


@pytest.mark.parametrize("min_matched,overlap", [(1, 0.0), (1, 1.0), (2, 0.5), (3, 0.34), (4, 0.75), (2, 1.5), (7, 0.1)])
def test_overlap_gates_follow_the_params(pyqms, min_matched, overlap):
    """MINIMUM_NUMBER_OF_MATCHED_ISOTOPOLOGUES x REQUIRED_PERCENTILE_PEAK_OVERLAP (IL:1968-1976): the tabulated
    per-window-count thresholds of the gate give the oracle's hit sets for loose, strict and unreachable settings."""
    sys.setrecursionlimit(20000)
    from oracle.qms_oracle import OracleLibrary
    from pyqms_b200.synthetic import random_tryptic_peptides, synth_run
    from tests.golden.spectra import entry_peaks
    params = {"MINIMUM_NUMBER_OF_MATCHED_ISOTOPOLOGUES": min_matched, "REQUIRED_PERCENTILE_PEAK_OVERLAP": overlap,
              "M_SCORE_THRESHOLD": 0.3}
    kwargs = dict(molecules=random_tryptic_peptides(40, seed=5), charges=[1, 2, 3], params=params)
    ora = OracleLibrary(**kwargs)
    lib = pyqms.IsotopologueLibrary(verbose=False, **kwargs)
    mzs, cis = entry_peaks(ora)
    peaks, offsets = synth_run(mzs, cis, n_spectra=20, seed=99, mean_peaks=300, elution_sigma=2.0, entry_fraction=0.4, dropout=0.35)
    want = {}
    for s in range(len(offsets) - 1):
        ora.match_all(peaks[offsets[s]:offsets[s + 1]], "r", s, float(s), want)
    got = lib.match_many(peaks, file_name="r", spec_ids=list(range(20)), spec_rts=[float(s) for s in range(20)], offsets=offsets)
    assert_results_equal(results_to_rows(got), results_to_rows(want), score_tol=SCORE_TOL, cmz_rel=REL_MASS, ri_rel=REL_ABUN)
    if overlap > 1.0:
        assert len(want) == 0
    if (min_matched, overlap) == (1, 0.0):
        assert sum(len(v["data"]) if isinstance(v, dict) else len(v) for v in want.values()) > 20
