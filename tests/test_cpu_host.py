"""CPU-only checks (-m "not gpu"): the C-ABI library loads and exports every declared symbol, the host
planning layer (string work) matches the reference's golden dumps, sharding helpers, synthetic generators."""
import ctypes
import re
import sys
from pathlib import Path

import numpy as np
import pytest

from tests.golden.cases import CASES
from tests.helpers import load_case, lp_tuple

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "qms_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qms_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pyqms_b200 import _capi
    if not _capi.LIB_PATH.exists():                 # fresh checkout: compile first (nvcc cross-compiles without a GPU)
        sys.path.insert(0, str(ROOT))
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(str(_capi.LIB_PATH))          # built by __graft_entry__.build() / make -C pyqms_b200/csrc
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(_capi.EXPORTS) == names           # the ctypes table binds exactly the declared ABI


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import pyqms_b200
    from pyqms_b200._capi import QmsError
    with pytest.raises(QmsError, match="no CPU fallback"):
        pyqms_b200.IsotopologueLibrary(molecules=["PEPTIDE"], charges=[2], verbose=False)


def test_saved_library_holds_the_plan_and_needs_a_gpu_to_load(tmp_path):
    import pickle
    import torch
    import pyqms_b200
    from pyqms_b200._capi import QmsError
    lib = pyqms_b200.IsotopologueLibrary(molecules=["PEPTIDEK", "DDSPDLPK#Oxidation:1"], charges=[1, 2], verbose=False,
                                         _plan_only=True)
    lib.save(tmp_path / "plan.qmslib")
    blob = (tmp_path / "plan.qmslib").read_bytes()
    restore, (state, formulas) = lib.__reduce__()
    assert "_ctx" not in state and "_views" not in state and "_terms" in state and "_pool_tables" in state
    assert formulas == list(lib.keys()) and len(state["_cc_counts"]) == len(formulas)
    if not torch.cuda.is_available():
        with pytest.raises(QmsError, match="no CPU fallback"):
            pickle.loads(blob)
        with pytest.raises(QmsError, match="no CPU fallback"):
            pyqms_b200.adaptors.calc_amount_function({"rt": [1.0], "i": [2.0], "scores": [0.5]})


def test_product_never_imports_the_oracle():
    for path in (ROOT / "pyqms_b200").rglob("*.py"):
        assert "oracle" not in path.read_text(), path


@pytest.mark.parametrize("name", list(CASES))
def test_host_planning_matches_reference(name):
    """Formulas, compositions, label tuples, fixed-label expansion, pools and envelope terms (no device)."""
    import pyqms_b200
    blob, _, _ = load_case(name)
    ref = blob["library"]
    lib = pyqms_b200.IsotopologueLibrary(verbose=False, _plan_only=True, **CASES[name]["lib"])
    assert sorted(lib.keys()) == sorted(ref["formulas"])
    assert lib.labled_percentiles == [lp_tuple(x) for x in ref["labled_percentiles"]]
    assert lib.lookup["molecule to formula"] == ref["molecule_to_formula"]
    assert {k: sorted(v) for k, v in lib.lookup["molecule fixed label variations"].items()} == ref["fixed_label_variations"]
    for formula in ref["formulas"]:
        assert dict(lib[formula]["cc"]) == ref["cc"][formula]
    assert lib.params["FIXED_LABEL_ISOTOPE_ENRICHMENT_LEVELS"] == ref["fixed_label_levels"]
    # every element tree the reference built has a pool, with the reference's isotope table layout
    for element, by_label in ref["isotopic_distributions"].items():
        for label, table in by_label.items():
            pool = lib._pool_ids[(element, label)]
            off = lib._pool_tables[0]
            assert off[pool + 1] - off[pool] == len(table)
            assert list(lib._pool_tables[3][off[pool]:off[pool + 1]]) == [row[2] for row in table]
    term_off = lib._terms[0]
    assert len(term_off) == len(ref["formulas"]) * len(ref["labled_percentiles"]) + 1
    # envelope terms: (pool, level) in the reference's element order -> levels must exist in its trees
    for e in range(len(term_off) - 1 if ref["trees"] else 0):      # (tree dumps are skipped for the 22 kDa fixture)
        for t in range(term_off[e], term_off[e + 1]):
            element, label = lib._pools[lib._terms[1][t]]
            assert str(lib._terms[2][t]) in ref["trees"][element][label], (element, label, lib._terms[2][t])


def test_chemistry_parser():
    from pyqms_b200.chemistry import ChemicalComposition, parse_formula
    assert parse_formula("C(2)H(3)14N(1)O(1)") == {"C": 2, "H": 3, "14N": 1, "O": 1}
    assert parse_formula("+H2O") == {"H": 2, "O": 1}
    assert parse_formula("C(-6) 13C(6) N(-4) 15N(4)") == {"C": -6, "13C": 6, "N": -4, "15N": 4}
    import pyqms_b200
    aa = {k: ChemicalComposition(formula="+" + v) for k, v in pyqms_b200.knowledge_base.aa_compositions.items()}
    cc = ChemicalComposition(aa_compositions=aa)
    cc.use(deprecated_format="PEPTIDE")
    assert cc.hill_notation_unimod() == "C(34)H(53)N(7)O(15)"          # IL docstring, IL:100-104
    cc.use(deprecated_format="TESTPEPTIDE#Oxidation:1")
    assert cc.hill_notation_unimod() == "C(50)H(79)N(11)O(25)"          # tests/peptide_sequence_test.py:24
    cc.use(deprecated_format="PEPTIDE+PO3")
    assert cc["P"] == 1 and cc["O"] == 18
    with pytest.raises(KeyError):
        cc.use(deprecated_format="PEPTIDE#NoSuchModification:1")
    assert cc["Zz"] == 0


def test_results_container():
    import pyqms_b200
    res = pyqms_b200.Results(params=pyqms_b200.params)
    key = ("f.mzML", "C(2)H(4)", 2, (("N", "0.000"),))
    res.add(key, (1, 0.1, 0.6, 2.0, ((1.0, 2.0, 1.0, 1.0, 5),)))
    k = res.add(key, (2, 0.2, 0.9, 3.0, ((None, None, 1.0, 1.0, 5),)))
    assert isinstance(k, pyqms_b200.m_key) and k.charge == 2
    slot = res[k]
    assert slot["len_data"] == 2 and slot["max_score"] == 0.9 and slot["max_score_index"] == 1
    assert isinstance(slot["data"][0], pyqms_b200.match) and slot["data"][1].spec_id == 2
    assert res.index["formulas"] == {"C(2)H(4)"} and res.index["charges"] == {2}
    assert res.max_score()[0] == 0.9 and res.max_score()[2] == 1          # [score, key, index, match] (results.py:360-398)
    assert res.max_score(charges=[3]) == [0, None, None, None]
    assert len(list(res.extract_results(charges=[3]))) == 0
    assert len(list(res.extract_results(score_threshold=0.8))) == 1
    frame = res.format_all_results()
    assert list(frame.columns) == ["file_name", "formula", "charge", "label_percentiles", "spec_id", "rt", "score",
                                   "scaling_factor", "peaks"] and len(frame) == 2
    assert res.determine_max_itensity({"rt": [1, 2], "i": [5, 9], "scores": [0.6, 0.7]})["max I in window (rt)"] == 2


def test_shard_ranges_cover_exactly():
    from pyqms_b200.distributed import shard_range
    for n in (0, 1, 7, 8, 9, 1000003):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(e - b for b, e in parts) - min(e - b for b, e in parts) <= 1


def test_synthetic_generators_are_deterministic():
    from pyqms_b200.synthetic import averagine_formulas, random_tryptic_peptides, synth_run
    a, b = random_tryptic_peptides(50, seed=1), random_tryptic_peptides(50, seed=1)
    assert a == b and len(set(a)) == 50 and all(p[-1] in "KR" and 7 <= len(p) <= 30 for p in a)
    assert averagine_formulas(5, seed=2) == averagine_formulas(5, seed=2)
    mz = [np.array([500.0, 500.5]), np.array([700.0, 700.33, 700.66])]
    ci = [np.array([60000.0, 20000.0]), np.array([50000.0, 30000.0, 9000.0])]
    p1, o1 = synth_run(mz, ci, n_spectra=10, seed=3, mean_peaks=50, min_peaks=10, elution_sigma=2.0)
    p2, o2 = synth_run(mz, ci, n_spectra=10, seed=3, mean_peaks=50, min_peaks=10, elution_sigma=2.0)
    assert np.array_equal(p1, p2) and np.array_equal(o1, o2) and o1[-1] == len(p1)
    for s in range(10):
        seg = p1[o1[s]:o1[s + 1], 0]
        assert np.all(seg[1:] >= seg[:-1])


def test_mzml_round_trip(tmp_path):
    """N3: mzML -> packed spectra; 64/32-bit floats, zlib or plain, MS-level filter, unsorted peaks."""
    from pyqms_b200.mzml import read_mzml, write_mzml
    rng = np.random.default_rng(8)
    spectra = []
    for n in range(7):
        mz = np.sort(rng.uniform(150, 2000, size=int(rng.integers(0, 40))))
        spectra.append(np.column_stack([mz, 10 ** rng.uniform(3, 6, size=len(mz))]))
    spectra[3] = spectra[3][::-1].copy()                     # a file with descending m/z
    levels = [1, 2, 1, 1, 2, 1, 1]
    for precision, compress in ((64, True), (64, False), (32, True)):
        path = tmp_path / ("run_%d_%d.mzML" % (precision, compress))
        write_mzml(path, spectra, rts_minutes=[0.5 * n for n in range(7)], ms_levels=levels, precision=precision, compress=compress)
        run = read_mzml(path, ms_level=1)
        keep = [n for n, level in enumerate(levels) if level == 1]
        assert run.name == path.name and run.spec_ids == [n + 1 for n in keep]
        assert run.rts == [(0.5 * n, "minute") for n in keep]
        assert list(np.diff(run.offsets)) == [len(spectra[n]) for n in keep]
        for k, n in enumerate(keep):
            got = run.peaks[run.offsets[k]:run.offsets[k + 1]]
            want = spectra[n][np.argsort(spectra[n][:, 0], kind="stable")]
            if precision == 64:
                assert np.array_equal(got, want)
            else:
                assert np.allclose(got, want, rtol=1e-6)
            assert np.all(got[1:, 0] >= got[:-1, 0])
        assert len(read_mzml(path, ms_level=None).spec_ids) == 7


def build_c_demo(tmp_path):
    import subprocess
    exe = tmp_path / "c_abi_demo"
    lib_dir = ROOT / "pyqms_b200"
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-I" + str(ROOT / "include"), str(ROOT / "examples" / "c_abi_demo.c"), "-o", str(exe),
           "-L" + str(lib_dir), "-lqms_b200", "-Wl,-rpath," + str(lib_dir), "-lm"]
    done = subprocess.run(cmd, capture_output=True, text=True)
    assert done.returncode == 0, done.stderr
    return exe


def test_c_abi_links_from_plain_c(tmp_path):
    """The boundary is a C ABI: a plain C program (no Python, no torch) compiles against include/qms_b200.h and links
    libqms_b200.so; without a GPU it must fail loudly (there is no CPU fallback)."""
    import subprocess
    import torch
    exe = build_c_demo(tmp_path)
    if not torch.cuda.is_available():
        done = subprocess.run([str(exe)], capture_output=True, text=True)
        assert done.returncode == 2 and "no CPU fallback" in done.stderr


def test_packed_hits_widen_the_compact_columns():
    """What a deferred pass brings back (entry / score / scaling per hit, uint16 rows per hit window, hit offsets per
    spectrum) widens on the host to the columns every consumer reads: spec, win_off, peak -- and mmz / mi by row."""
    from pyqms_b200.isotopologue_library import PackedHits
    rng = np.random.default_rng(11)
    n_spec, n_entries = 40, 25
    lengths = rng.integers(0, 300, size=n_spec)
    lengths[[3, 17]] = 0                                               # empty spectra
    offsets = 1000 + np.concatenate([[0], np.cumsum(lengths)])         # a batch that starts at row 1000 of the array
    peaks = rng.uniform(100.0, 2000.0, size=(int(offsets[-1]), 2))
    entry_windows = rng.integers(1, 12, size=n_entries)
    hits_per_spectrum = np.where(lengths > 0, rng.integers(0, 5, size=n_spec), 0)
    spec = np.repeat(np.arange(n_spec, dtype=np.int32), hits_per_spectrum)
    entry = rng.integers(0, n_entries, size=len(spec)).astype(np.int32)
    win_off = np.concatenate([[0], np.cumsum(entry_windows[entry])]).astype(np.int64)
    first_row = np.repeat(offsets[spec], entry_windows[entry])
    span = np.repeat(lengths[spec], entry_windows[entry])
    local = (rng.uniform(size=win_off[-1]) * span).astype(np.int64)
    missing = rng.uniform(size=win_off[-1]) < 0.2
    peak = np.where(missing, -1, first_row + local)
    out = PackedHits(peaks, offsets, entry_windows)
    out.update({"entry": entry, "score": rng.uniform(size=len(spec)), "scaling": rng.uniform(size=len(spec)),
                "peak16": np.where(missing, 0xFFFF, local).astype(np.uint16),
                "spec_off": np.concatenate([[0], np.cumsum(hits_per_spectrum)]).astype(np.int64), "n_hits": len(spec), "n_windows": int(win_off[-1])})
    assert "spec" not in dict.keys(out) and out.get("nothing", 7) == 7
    assert np.array_equal(out["spec"], spec) and out["spec"].dtype == np.int32
    assert np.array_equal(out["win_off"], win_off) and out["win_off"].dtype == np.int64
    assert np.array_equal(out["peak"], peak) and out["peak"].dtype == np.int64
    expected = np.where(missing, np.nan, peaks[np.maximum(peak, 0), 0])
    assert np.array_equal(out["mmz"], expected, equal_nan=True)
    assert np.array_equal(np.isnan(out["mi"]), missing)
    assert "spec" in dict.keys(out)                                    # widened once, then kept
    # no hits at all
    empty = PackedHits(peaks, offsets, entry_windows)
    empty.update({"entry": entry[:0], "peak16": np.zeros(0, np.uint16), "spec_off": np.zeros(n_spec + 1, np.int64)})
    assert len(empty["spec"]) == 0 and list(empty["win_off"]) == [0] and len(empty["peak"]) == 0 and len(empty["mmz"]) == 0
