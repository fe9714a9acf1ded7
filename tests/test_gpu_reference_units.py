"""The reference's own unit tests (nose generators under /root/reference/tests) re-expressed against the
drop-in class.  Expectations are the reference's literals; each test names the file it restates."""
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pyqms():
    import pyqms_b200
    return pyqms_b200


def test_build_label_percentile_tuples(pyqms):
    """build_label_percentile_tuples_test.py:7-18, increase_element_envelope_test.py:8-13"""
    lib = pyqms.IsotopologueLibrary(molecules=["KLEINERTEST"], charges=[2], verbose=False,
                                    metabolic_labels={"15N": [0, 0.99], "13C": [0.01, 0.98]})
    want = [(("C", "0.010"), ("N", "0.000")), (("C", "0.010"), ("N", "0.990")), (("C", "0.980"), ("N", "0.000")),
            (("C", "0.980"), ("N", "0.990"))]
    assert sorted(lib.labled_percentiles) == want
    lib = pyqms.IsotopologueLibrary(molecules=["KLEINERTEST"], charges=[2], verbose=False, metabolic_labels={"18O": [0, 0.99]})
    assert sorted(lib.labled_percentiles) == [(("O", "0.000"),), (("O", "0.990"),)]


def test_create_combinations(pyqms):
    """create_combinations_test.py:9-27"""
    lib = pyqms.IsotopologueLibrary(molecules=["PAINLESS"], charges=[2], verbose=False)
    assert sorted(lib._create_combinations([("X", 2), ("Y", 2)], all_combos=None, current_combo=None, pos=0)) == [
        [("X", 0), ("Y", 0)], [("X", 0), ("Y", 1)], [("X", 1), ("Y", 0)], [("X", 1), ("Y", 1)]]
    assert sorted(lib._create_combinations([("X", 1), ("Y", 1), ("Z", 1)])) == [[("X", 0), ("Y", 0), ("Z", 0)]]
    assert sorted(lib._create_combinations([("X", 1), ("Y", 2), ("Z", 1)])) == [[("X", 0), ("Y", 0), ("Z", 0)],
                                                                               [("X", 0), ("Y", 1), ("Z", 0)]]


@pytest.mark.parametrize("case", [
    dict(fixed_labels={"R": ["Label:13C(6)15N(4)", ""]}, params={}, out=["KLEINER0TEST", "KLEINER1TEST"]),
    dict(fixed_labels={"K": ["Label:13C(6)15N(2)", "Label:13C(6)15N(4)", ""], "R": ["Label:13C(6)15N(4)", ""]}, params={},
         out=["K0LEINER0TEST", "K0LEINER1TEST", "K1LEINER0TEST", "K1LEINER1TEST", "K2LEINER0TEST", "K2LEINER1TEST"]),
    dict(fixed_labels={"K": ["Label:13C(6)15N(2)", ""], "R": ["Label:13C(6)15N(4)", ""]},
         params={"SILAC_AAS_LOCKED_IN_EXPERIMENT": ["K", "R"]}, out=["K0LEINER0TEST", "K1LEINER1TEST"]),
])
def test_extend_molecules_with_fixed_labels(pyqms, case):
    """extend_molecules_with_fixed_labels_test.py:7-39, extend_kb_with_fixed_labels_test.py:145-157"""
    lib = pyqms.IsotopologueLibrary(molecules=["KLEINERTEST"], charges=[2], params=case["params"],
                                    fixed_labels=case["fixed_labels"], verbose=False)
    assert sorted(lib._extend_molecules_with_fixed_labels(["KLEINERTEST"])) == sorted(case["out"])
    assert sorted(lib.lookup["molecule fixed label variations"]["KLEINERTEST"]) == sorted(case["out"])


def test_extend_kb_with_amino_acids_and_crash(pyqms):
    """extend_kb_with_amino_acids_test.py:12-61 (custom amino acid; unknown amino acid -> SystemExit(1))"""
    lib = pyqms.IsotopologueLibrary(molecules=["UU"], charges=[2], params={"AMINO_ACIDS": {"U": "C(3)H(7)N(1)O(2)Se(1)"}},
                                    verbose=False)
    assert dict(lib.aa_compositions["U"]) == {"C": 3, "H": 7, "N": 1, "O": 2, "Se": 1}
    assert list(lib.keys()) == ["C(6)H(16)N(2)O(5)Se(2)"]
    with pytest.raises(SystemExit) as info:
        pyqms.IsotopologueLibrary(molecules=["KLEINERTEST"], charges=[2], fixed_labels={"X": ["C(-6) 13C(6) N(-4) 15N(4)"]},
                                  verbose=False)
    assert info.value.code == 1


@pytest.mark.parametrize("element,isotope,target", [("N", 15, 0.994), ("O", 18, 0.970), ("N", 15, 0.32)])
def test_recalc_isotopic_distribution(pyqms, element, isotope, target):
    """recalc_isotopic_distribution_test.py:14-73: the target percentile is present and abundances sum to 1"""
    lib = pyqms.IsotopologueLibrary(molecules=["KLEINERTEST"], charges=[2], verbose=False)
    dist = lib._recalc_isotopic_distribution(element=element, target_percentile=target, enriched_isotope=isotope)
    assert str(round(target, 3)) in {str(round(a, 3)) for _, a, _ in dist}
    assert 1 - sum(a for _, a, _ in dist) <= sys.float_info.epsilon


@pytest.mark.parametrize("set1,set2", [
    (dict(molecules=["KLEINERTEST"], fixed_labels={"C": ["C(2) H(3) N(1) O(1)"]}),
     dict(molecules=["TESTKLEINER"], fixed_labels={"C": ["C(2) H(3) 14N(1) O(1)"]})),
    (dict(molecules=["KLEINERTEST"], fixed_labels={"C": ["C(2) H(3) N(1) O(1)"]}, metabolic_labels={"15N": [0.994]}),
     dict(molecules=["TESTKLEINER"], fixed_labels={"C": ["C(2) H(3) 15N(1) O(1)"]}, metabolic_labels={"15N": [0.994]})),
    (dict(molecules=["KLEINERTEST"], fixed_labels={"R": ["C(-6) 13C(6) 15N(-4) 15N(4)"]}, metabolic_labels={"15N": [0.994]}),
     dict(molecules=["TESTKLEINER"], fixed_labels={"R": ["C(-6) 13C(6) N(-4) N(4)"]}, metabolic_labels={"15N": [0.994]})),
])
def test_carbamido_14n_15n(pyqms, set1, set2):
    """carbamido_14N_15N_test.py:16-83: first isotopologue mass agrees between the N / 14N / 15N spellings"""
    lib_1 = pyqms.IsotopologueLibrary(charges=[2], verbose=False, **set1)
    lib_2 = pyqms.IsotopologueLibrary(charges=[2], verbose=False, **set2)
    f1, f2 = list(lib_1.keys())[0], list(lib_2.keys())[0]
    for lp in lib_1[f1]["env"].keys():
        assert lib_1[f1]["env"][lp]["mass"][0] - lib_2[f2]["env"][lp]["mass"][0] < 0.000000001


def test_knowledge_base(pyqms):
    """knowledge_base_tests.py: every element's abundances sum to 1 (to 3 decimals as in the reference)"""
    for element, table in pyqms.knowledge_base.isotopic_distributions.items():
        assert abs(sum(a for _, a in table) - 1.0) < 1e-3, element


def test_slice_list_and_transform_spectrum(pyqms):
    """slice_list_test.py, transform_spectrum_test.py:12-19"""
    lib = pyqms.IsotopologueLibrary(molecules=["PEPTIDE"], charges=[2], verbose=False, params={"INTERNAL_PRECISION": 10000})
    keys, lookup = lib._transform_spectrum([(1000.0009, 10.0)])
    assert keys == {10000009} and lookup == {10000009: [(1000.0009, 10.0)]}
    data = [(1.0, 1), (2.0, 2), (3.0, 3), (4.0, 4), (5.0, 5), (6.0, 6), (7.0, 7)]
    assert lib._slice_list(data, ((4.0, 0), (4.0, 9)), tolerance=1) == [(3.0, 3), (4.0, 4), (5.0, 5)]
    arr = np.array(data, dtype=float)
    assert len(lib._slice_list(arr, ((4.0, 0), (4.0, 0)), tolerance=1.5)) == 3


def test_print_overview_runs(pyqms, capsys):
    lib = pyqms.IsotopologueLibrary(molecules=["PEPTIDE"], charges=[1], verbose=False)
    lib.print_overview("PEPTIDE", charge=1)
    text = capsys.readouterr().out
    assert "C(34)H(53)N(7)O(15)" in text and "799.3599640346" in text  # IL:2057 docstring row 0 (mass column)
    env = lib["C(34)H(53)N(7)O(15)"]["env"][(("N", "0.000"),)]
    assert env["abun"][:4] == [64799, 26251, 7164, 1456]              # IL:2057-2060
