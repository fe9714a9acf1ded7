"""Native host planner (qms_parse_molecules / qms_hill_formulas) and the array-built term tables against the
one-molecule-at-a-time host parser and a direct restatement of the reference's term loop (IL:546-646)."""
import random

import numpy as np
import pytest

import pyqms_b200
from pyqms_b200 import knowledge_base, planner
from pyqms_b200.chemistry import ChemicalComposition
from pyqms_b200.synthetic import averagine_formulas, random_tryptic_peptides
from tests.golden.cases import CASES

SYMBOLS = list(knowledge_base.isotopic_distributions)


def host_composition(molecule):
    factory = ChemicalComposition(aa_compositions=knowledge_base.aa_compositions)
    return factory.composition_of(molecule)


def native_rows(molecules):
    counts, seen, status = planner.parse_molecules(molecules, SYMBOLS, knowledge_base.aa_compositions)
    rows = []
    for i in range(len(molecules)):
        order = [k for k in np.argsort(np.where(seen[i] >= 0, seen[i], 32767), kind="stable") if seen[i, k] >= 0]
        rows.append([(SYMBOLS[k], int(counts[i, k])) for k in order])
    return rows, status, counts


HANDLED = ["+H2O", "+C104H175N29O33S1", "+C(2)H(3)N(1)O(1)", "+C13C6", "+C2 H6 O", "+C(12)H(-2)O(3)H(2)", "+H2OH(-2)O(-1)C",
           "PEPTIDEK", "PEPTIDEK+PO3", "A", "CCTESLVNR", "ACDEFGHIKLMNPQRSTVWY", "+Se2C3", "+NaCl"]
LEFT_TO_HOST = ["PEPTIDEK#Oxidation:3", "PEPTIDEK0", "pEPTIDE", "PEPTIDE+H2O+H2O", "+C(-6)13C(6)", "+Xx2", "+C(2", "+c2", "",
                "PEPTIDEB", "PEP TIDE", "+C6H12O6+", "+(2)"]


def test_native_parser_equals_host_parser_on_handled_syntax():
    rows, status, counts = native_rows(HANDLED)
    assert status.tolist() == [0] * len(HANDLED)
    for molecule, row in zip(HANDLED, rows):
        assert row == list(host_composition(molecule).items()), molecule
    formulas = planner.hill_formulas(counts, SYMBOLS)
    assert formulas == [host_composition(m).hill_notation_unimod() for m in HANDLED]


def test_native_parser_hands_everything_else_to_the_host():
    rows, status, counts = native_rows(LEFT_TO_HOST)
    assert status.tolist() == [1] * len(LEFT_TO_HOST)
    assert not counts.any() and all(row == [] for row in rows)


def test_native_parser_on_random_molecules():
    rng = random.Random(5)
    molecules = averagine_formulas(3000)[::3] + random_tryptic_peptides(1500, seed=3)
    elements = ["C", "H", "N", "O", "S", "P", "Se", "Na", "Cl", "Fe"]
    for _ in range(1500):           # shuffled element order, repeated elements, bracketed and negative counts
        parts = []
        for element in rng.choices(elements, k=rng.randint(1, 7)):
            n = rng.randint(-3, 40)
            parts.append("%s(%d)" % (element, n) if rng.random() < 0.5 or n < 1 else "%s%s" % (element, n if n > 1 else ""))
        molecules.append("+" + "".join(parts))
    rows, status, counts = native_rows(molecules)
    assert not status.any()
    for molecule, row in zip(molecules, rows):
        assert row == list(host_composition(molecule).items()), molecule
    assert planner.hill_formulas(counts, SYMBOLS) == [host_composition(m).hill_notation_unimod() for m in molecules]


def reference_term_loop(lib):
    """The per-formula, per-label-tuple loop the array code replaces (IL:546-646), on lib[formula]["cc"]."""
    levels = lib.params["FIXED_LABEL_ISOTOPE_ENRICHMENT_LEVELS"]
    off, pool, level = [0], [], []
    for formula in lib._formulas:
        composition = lib[formula]["cc"]
        for lp in lib.labled_percentiles:
            merged = {}
            for element, count in composition.items():
                target = element
                if element in levels:
                    bare = element[-1]
                    for lp_element, lp_label in lp:
                        if lp_element == bare and lp_label == lib._fmt(levels[element]):
                            target = bare
                merged[target] = merged.get(target, 0) + count
            for element, n in merged.items():
                if element in lib.metabolically_labeled_elements:
                    continue
                pool.append(lib._pool_ids[(element, lib._default_label[element])])
                level.append(n)
            for element, label in lp:
                n = merged.get(element, 0)
                if n == 0:
                    continue
                pool.append(lib._pool_ids[(element, label)])
                level.append(n)
            off.append(len(pool))
    return off, pool, level


@pytest.mark.parametrize("name", sorted(CASES))
def test_bulk_planning_equals_one_by_one_planning(name):
    kwargs = dict(CASES[name]["lib"])
    native = pyqms_b200.IsotopologueLibrary(verbose=False, _plan_only=True, **kwargs)
    by_hand = pyqms_b200.IsotopologueLibrary(verbose=False, _plan_only=True, _native_planner=False, **kwargs)
    assert native._formulas == by_hand._formulas and native._symbols == by_hand._symbols
    assert native.lookup == by_hand.lookup
    assert list(native._highest_element_count.items()) == list(by_hand._highest_element_count.items())
    assert native._range_for_elements_with_two_isotopes == by_hand._range_for_elements_with_two_isotopes
    assert native._pools == by_hand._pools
    for a, b in zip(native._pool_tables + native._terms, by_hand._pool_tables + by_hand._terms):
        assert a.dtype == b.dtype and np.array_equal(a, b, equal_nan=a.dtype.kind == "f")
    for formula in native._formulas:
        assert list(native[formula]["cc"].items()) == list(by_hand[formula]["cc"].items())
        assert set(native[formula]) == {"env", "cc"} and "cc" in native[formula] and len(native[formula]) == 2
    off, pool, level = reference_term_loop(native)
    assert native._terms[0].tolist() == off and native._terms[1].tolist() == pool and native._terms[2].tolist() == level
    assert native._n_env == len(off) - 1


def test_term_tables_of_a_mixed_library_with_folding():
    """Fixed-label isotopes that fold into a labelled element (Q6), modifications, formulas and peptides together."""
    molecules = random_tryptic_peptides(300, seed=9) + averagine_formulas(300)[::5] + [
        "PEPTIDEK#Oxidation:1", "CCTESLVNR#Carbamidomethyl:1;Carbamidomethyl:2", "+C(-6)13C(6)H(12)O(6)", "ELVISLIVESK#Label:13C(6)15N(2):11",
        "+C(5)15N(2)N(1)H(9)O(2)"]
    kwargs = dict(molecules=molecules, charges=[2], metabolic_labels={"15N": [0, 0.5, 0.994], "13C": [0, 0.994]},
                  fixed_labels={"C": ["C(2)H(3)14N(1)O(1)"]}, verbose=False, _plan_only=True)
    native = pyqms_b200.IsotopologueLibrary(**kwargs)
    by_hand = pyqms_b200.IsotopologueLibrary(_native_planner=False, **kwargs)
    assert native._formulas == by_hand._formulas
    for a, b in zip(native._pool_tables + native._terms, by_hand._pool_tables + by_hand._terms):
        assert np.array_equal(a, b, equal_nan=a.dtype.kind == "f")
    off, pool, level = reference_term_loop(native)
    assert native._terms[0].tolist() == off and native._terms[1].tolist() == pool and native._terms[2].tolist() == level
    assert len(native.labled_percentiles) == 6 and native._n_env == 6 * len(native._formulas)


def test_envelope_terms_do_not_depend_on_the_block_size():
    """planner.envelope_terms works through the formulas in blocks (bounded memory); any block size gives the same tables."""
    kwargs = dict(molecules=random_tryptic_peptides(500, seed=21) + averagine_formulas(200), charges=[2],
                  metabolic_labels={"15N": [0, 0.3, 0.99]}, fixed_labels={"C": ["C(2)H(3)14N(1)O(1)"]}, verbose=False, _plan_only=True)
    lib = pyqms_b200.IsotopologueLibrary(**kwargs)
    levels = lib.params["FIXED_LABEL_ISOTOPE_ENRICHMENT_LEVELS"]
    column = {symbol: k for k, symbol in enumerate(lib._symbols)}
    pool_of_default = np.array([lib._pool_ids.get((s, lib._default_label.get(s)), 0) for s in lib._symbols], dtype=np.int32)
    labelled = np.array([s in lib.metabolically_labeled_elements for s in lib._symbols], dtype=bool)
    fold_target, pool_of_label = [], []
    for lp in lib.labled_percentiles:
        target = list(range(len(lib._symbols)))
        for k, symbol in enumerate(lib._symbols):
            if symbol in levels:
                for element, label in lp:
                    if element == symbol[-1] and label == lib._fmt(levels[symbol]) and symbol[-1] in column:
                        target[k] = column[symbol[-1]]
        fold_target.append(target)
        pool_of_label.append([(column.get(element), lib._pool_ids.get((element, label), 0)) for element, label in lp])
    for rows_per_block in (1, 7, 64, 1 << 16):
        off, pool, level = planner.envelope_terms(lib._cc_counts, lib._cc_seen, lib._symbols, lib.labled_percentiles, fold_target,
                                                  labelled, pool_of_default, pool_of_label, rows_per_block=rows_per_block)
        assert np.array_equal(off, lib._terms[0]) and np.array_equal(pool, lib._terms[1]) and np.array_equal(level, lib._terms[2])


def test_formula_records_behave_like_the_reference_dicts():
    lib = pyqms_b200.IsotopologueLibrary(molecules=["PEPTIDEK", "+C6H12O6"], charges=[1], verbose=False, _plan_only=True)
    formula = lib.lookup["molecule to formula"]["+C6H12O6"]
    record = lib[formula]
    assert "cc" in record and "env" in record and "other" not in record and record.get("other", 5) == 5
    assert sorted(record) == ["cc", "env"] and len(record) == 2 and sorted(record.keys()) == ["cc", "env"]
    assert dict(record["cc"]) == {"C": 6, "H": 12, "O": 6} and list(record["cc"]) == ["C", "H", "O"]
    assert record["cc"] is record["cc"]                                 # built once
    assert [key for key, _ in record.items()] == list(record.keys()) and len(list(record.values())) == 2
    with pytest.raises(KeyError):
        record["nothing"]
    peptide = lib[lib.lookup["molecule to formula"]["PEPTIDEK"]]["cc"]
    assert list(peptide.items()) == list(host_composition("PEPTIDEK").items())


def test_fixed_label_fast_path_equals_the_general_expansion():
    """One label per labelled residue, nothing locked: a str.translate per molecule must give what the site / combination
    loop of IL:1375-1677 gives (expanded molecule set, variation lookup, formulas)."""
    import pyqms_b200
    from pyqms_b200.synthetic import CAM_FIXED_LABEL, random_tryptic_peptides
    peptides = random_tryptic_peptides(3000, seed=3)
    molecules = peptides + [p + "#Oxidation:1" for p in peptides[:40]] + ["+C6H12O6", "PEPTIDEC+PO3", "CCCC", "AAAK"]
    kwargs = dict(molecules=molecules, charges=[2], fixed_labels=dict(CAM_FIXED_LABEL, M=["O(1)"]), verbose=False, _plan_only=True)
    fast = pyqms_b200.IsotopologueLibrary(**kwargs)
    pyqms_b200.IsotopologueLibrary._fixed_label_fast_path = False
    try:
        slow = pyqms_b200.IsotopologueLibrary(**kwargs)
    finally:
        pyqms_b200.IsotopologueLibrary._fixed_label_fast_path = True
    assert list(fast.keys()) == list(slow.keys())
    for name in ("molecule to formula", "formula to molecule", "molecule fixed label variations"):
        assert fast.lookup[name] == slow.lookup[name], name
    assert len(fast.lookup["molecule fixed label variations"]) > 1000


def test_fixed_label_fast_path_over_the_whole_list_at_once():
    """A list of nothing but residue letters is labelled by join / replace / split; the result, the lazily filled
    variation lookup and the compact element columns of the native parser must equal the general expansion."""
    import pyqms_b200
    from pyqms_b200.isotopologue_library import _LazyDict
    from pyqms_b200.synthetic import CAM_FIXED_LABEL, random_tryptic_peptides
    peptides = random_tryptic_peptides(3000, seed=5) + ["CCCC", "AAAK", "C", "MCMK"]
    kwargs = dict(molecules=peptides, charges=[2], fixed_labels=dict(CAM_FIXED_LABEL, M=["O(1)"]), verbose=False, _plan_only=True)
    fast = pyqms_b200.IsotopologueLibrary(**kwargs)
    assert isinstance(fast.lookup["molecule fixed label variations"], _LazyDict)          # the whole-list path ran
    assert len(fast._symbols) <= 6                                                      # C H N O S, not the knowledge base
    pyqms_b200.IsotopologueLibrary._fixed_label_fast_path = False
    try:
        slow = pyqms_b200.IsotopologueLibrary(**kwargs)
    finally:
        pyqms_b200.IsotopologueLibrary._fixed_label_fast_path = True
    assert list(fast.keys()) == list(slow.keys()) and fast._symbols == slow._symbols
    assert np.array_equal(fast._cc_counts, slow._cc_counts) and np.array_equal(fast._cc_seen, slow._cc_seen)
    for name in ("molecule to formula", "formula to molecule", "molecule fixed label variations"):
        assert fast.lookup[name] == slow.lookup[name], name
    assert fast.lookup["molecule fixed label variations"]["MCMK"] == {"M0C0M0K"}
    for a, b in zip(fast._terms + fast._pool_tables, slow._terms + slow._pool_tables):
        assert np.array_equal(a, b, equal_nan=np.asarray(a).dtype.kind == "f")
    # one molecule with anything else (or a line break) sends the list to the per-molecule paths
    label = pyqms_b200.IsotopologueLibrary._label_all_at_once
    assert label(["ACK", "CC"], "ACK", [("C", "C0")]) == ["AC0K", "C0C0"]
    assert label(["ACK", "CC+PO3"], "ACK", [("C", "C0")]) is None and label(["ACK", "C\nC"], "ACK", [("C", "C0")]) is None
    assert label([], "ACK", [("C", "C0")]) is None and label([""], "ACK", [("C", "C0")]) == [""]


def test_compact_symbol_columns_of_the_native_parser():
    symbols = list(knowledge_base.isotopic_distributions)
    residues = dict(knowledge_base.aa_compositions, U="C3H5NOSe")          # a residue that brings its own element
    peptides = random_tryptic_peptides(500, seed=9) + ["UK", "PEPTIDEK", "BK"]     # "B": no such residue, left to the host
    wide = planner.parse_molecules(peptides, symbols, residues)
    counts, seen, status, used = planner.parse_molecules(peptides, symbols, residues, compact=True)
    assert len(used) < 20 and "Se" in used and used == [s for s in symbols if s in used]
    columns = [symbols.index(s) for s in used]
    assert np.array_equal(wide[0][:, columns], counts) and np.array_equal(wide[1][:, columns], seen) and np.array_equal(wide[2], status)
    assert not np.delete(wide[0], columns, axis=1).any() and status[-1] == 1 and status[:-1].sum() == 0
    # formulas: the symbols spelled with the characters of the text (isotope-prefixed ones included), never fewer than used
    mixed = peptides + ["+C6H12O6", "+C(-6)13C(6)", "+Fe2O3Na1"]
    wide = planner.parse_molecules(mixed, symbols, residues)
    counts, seen, status, used = planner.parse_molecules(mixed, symbols, residues, compact=True)
    assert {"Fe", "Na", "Se"} <= set(used) and len(used) < len(symbols)
    columns = [symbols.index(s) for s in used]
    assert np.array_equal(wide[0][:, columns], counts) and np.array_equal(wide[1][:, columns], seen) and np.array_equal(wide[2], status)
    assert not np.delete(wide[0], columns, axis=1).any() and status[-4:].tolist() == [1, 0, 1, 0]       # 13C is not a symbol of the bare knowledge base: host parser


def test_hill_formulas_write_any_count():
    counts = np.array([[1, 0, -6, 2147483647], [12, 1, 0, -2147483648], [0, 0, 0, 0]], dtype=np.int32)
    assert planner.hill_formulas(counts, ["H", "13C", "C", "N"]) == ["C(-6)H(1)N(2147483647)", "H(12)13C(1)N(-2147483648)", ""]
