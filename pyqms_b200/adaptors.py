"""Evidence ingest for the library build (the caller in front of P1).

``parse_evidence`` mirrors ``pyqms.adaptors.parse_evidence`` (adaptors.py:204-506): identification
tables (Ursgal unified ``.csv`` or ``.mzTab`` PSM sections) become

* the fixed-label dictionary in the form ``IsotopologueLibrary(fixed_labels=...)`` takes,
* the evidence lookup ``{formula: {molecule: {"evidences": [...], "trivial_names": [...]}}}`` that the
  library stores under ``lookup["formula to evidences"]`` (IL:253-256) and that
  ``Results._determine_rt_windows_from_evidence`` / ``curate_rt_windows`` turn into RT windows
  (results.py:1283-1300, 1937-2064), and
* the molecule list with the fixed-label modifications stripped.

Host-side string work only; behaviour follows the reference including its corner cases (N-terminal
position 0 addresses the last residue, formula molecules ``+C..`` are scanned letter-wise for fixed-label residues).
"""
import codecs
import csv
import re
from copy import deepcopy

from . import knowledge_base
from .chemistry import ChemicalComposition, default_unimod_table, load_unimod_xml

_TRAILING_POSITION = re.compile(r":(?P<pos>[0-9]*$)")
_NAME_COLUMNS = ("proteinacc_start_stop_pre_post_;", "trivial_name", "Protein ID", "accession")
_LAYOUTS = {  # file kind -> (modifications, retention time [s], sequence) column names
    "csv": ("Modifications", "Retention Time (s)", "Sequence"),
    "mztab": ("modifications", "retention_time", "sequence"),
}


def _composition(unimod_file_list):
    return ChemicalComposition(aa_compositions=knowledge_base.aa_compositions, unimod_file_list=unimod_file_list,
                               add_default_files=False if unimod_file_list else True)


def _unimod_name_of_id(unimod_id, unimod_file_list):
    """First Unimod title of a record id (``UnimodMapper.id2first_name``, adaptors.py:364-366)."""
    for path in unimod_file_list or []:
        ids = _xml_ids(path)
        if unimod_id in ids:
            return ids[unimod_id]
    return default_unimod_table()["ids"].get(unimod_id)


def _xml_ids(path):
    import xml.etree.ElementTree as ET

    ids = {}
    for element in ET.parse(str(path)).getroot().iter():
        if element.tag.endswith("}mod") and element.get("record_id") is not None:
            ids.setdefault(element.get("record_id"), element.get("title"))
    load_unimod_xml(path)
    return ids


def _read_rows(path):
    """Rows of one evidence file as dicts + its layout key (adaptors.py:316-347)."""
    upper = str(path).upper()
    with codecs.open(str(path), mode="r", encoding="utf-8") as handle:
        if upper.endswith("CSV"):
            return list(csv.DictReader(handle)), "csv"
        if upper.endswith("MZTAB"):
            psm_lines = [line for line in handle if line[:3] in ("PSM", "PSH")]
            return list(csv.DictReader(psm_lines, delimiter="\t")), "mztab"
    raise ValueError("The format of {0} is not recognized by the pyQms adaptor function".format(path))


def _molecule_of_row(row, kind, unimod_file_list):
    mods_column, _, seq_column = _LAYOUTS[kind]
    mods = row.get(mods_column, "")
    if mods == "":
        return row[seq_column]
    if kind == "mztab":      # "2-UNIMOD:4,3-UNIMOD:4" -> "Carbamidomethyl:2;Carbamidomethyl:3"
        named = []
        for item in mods.split(","):
            position, accession = item.split("-")
            named.append("{0}:{1}".format(_unimod_name_of_id(accession.split(":")[1], unimod_file_list), position))
        mods = ";".join(named)
    return "{0}#{1}".format(row[seq_column], mods)


def _evidence_of_row(row, kind, score_field):
    _, rt_column, _ = _LAYOUTS[kind]
    evidence = {}
    rt = row.get(rt_column, "")
    if rt != "":
        evidence["RT"] = float(rt) / 60.0      # minutes; both formats store seconds
    score = row.get(score_field, "")
    if score != "":
        evidence["score"], evidence["score_field"] = float(score), score_field
    else:
        evidence["score"], evidence["score_field"] = "None", "None"
    names = [row.get(column, "") for column in _NAME_COLUMNS]
    names = [name for name in names if name != ""]
    if names:
        evidence["trivial_name"] = ";".join(names)
    return evidence, names


def parse_evidence(fixed_labels=None, evidence_files=None, molecules=None, evidence_score_field=None,
                   return_raw_csv_data=False, unimod_file_list=None):
    """Returns ``(formatted_fixed_labels, evidence_lookup, molecule_list)`` (+ the raw rows per file when
    ``return_raw_csv_data``); arguments and result as ``pyqms.adaptors.parse_evidence``."""
    molecules = [] if molecules is None else molecules
    score_field = "PEP" if evidence_score_field is None else evidence_score_field

    # ---- fixed labels: {aa: [{"element_composition": dict | composition, "evidence_mod_name": str}]} ----
    formatted, label_composition, names_of_residue, label_names = None, {}, {}, set()
    if fixed_labels is not None and len(fixed_labels) != 0:
        formatted = {}
        for residue, entries in fixed_labels.items():
            for entry in entries:
                given = entry["element_composition"]
                if isinstance(given, dict):       # plain dicts and composition objects (dict subclasses) alike
                    composition = _composition(unimod_file_list)
                    composition.add_chemical_formula(given)
                else:
                    composition = given
                formatted.setdefault(residue, []).append(composition.hill_notation_unimod())
                label_composition[entry["evidence_mod_name"]] = deepcopy(composition)
                names_of_residue.setdefault(residue, []).append(entry["evidence_mod_name"])
                label_names.add(entry["evidence_mod_name"])
                composition.clear()      # a non-dict composition object is emptied, as in adaptors.py:301

    # ---- evidence rows, grouped by molecule string --------------------------------------------------
    by_molecule, raw_rows, evidence_lookup = {}, {}, None
    for path in evidence_files or []:
        evidence_lookup = {}
        rows, kind = _read_rows(path)
        raw_rows[path] = rows
        for row in rows:
            molecule = _molecule_of_row(row, kind, unimod_file_list)
            evidence, names = _evidence_of_row(row, kind, score_field)
            slot = by_molecule.setdefault(molecule, {"evidences": [], "trivial_names": set()})
            slot["trivial_names"].update(names)
            slot["evidences"].append(evidence)

    # ---- strip fixed-label modifications, derive the complete formula of every molecule ---------------
    factory = _composition(unimod_file_list)
    stripped_molecules = set()
    for tagged in sorted(list(molecules) + list(by_molecule)):
        if tagged in by_molecule:
            by_molecule[tagged]["trivial_names"] = sorted(by_molecule[tagged]["trivial_names"])
        if "#" in tagged:
            molecule, modifications = tagged.split("#")
        else:
            molecule, modifications = tagged, None
        applied_labels = []
        if modifications is not None:
            kept = []
            for item in modifications.split(";"):
                found = _TRAILING_POSITION.search(item)
                if found is None:
                    raise ValueError("modification %r of %r carries no position" % (item, tagged))
                name, position = item[:found.start()], int(found.group("pos"))
                residue = molecule[position - 1]
                if formatted is not None and residue in formatted and name in label_names:
                    applied_labels.append(name)
                else:
                    kept.append(item)
            if kept:
                molecule = "{0}#{1}".format(molecule, ";".join(kept))
        elif formatted is not None:
            for residue in molecule:
                applied_labels.extend(names_of_residue.get(residue, []) if residue in formatted else [])
        if molecule.startswith("+"):
            factory.clear()
            factory.add_chemical_formula(molecule)
        elif "#" in molecule:
            sequence, remaining = molecule.split("#")
            factory.use(sequence=sequence, modifications=remaining)
        else:
            factory.use(sequence=molecule)
        for name in applied_labels:
            factory.add_chemical_formula(label_composition[name])
        formula = factory.hill_notation_unimod()
        stripped_molecules.add(molecule)
        if tagged in by_molecule:
            evidence_lookup.setdefault(formula, {})[tagged] = by_molecule[tagged]
        factory.clear()

    molecule_list = list(stripped_molecules)
    if return_raw_csv_data:
        return formatted, evidence_lookup, molecule_list, raw_rows
    return formatted, evidence_lookup, molecule_list


def calc_amount_function(obj_for_calc_amount):
    """Amounts of one matched isotope chromatogram ``{"rt": [...], "i": [...], "scores": [...]}``: maximum
    intensity with its retention time and score, summed intensity and the trapezoid area under the curve
    (``pyqms.adaptors.calc_amount_function``, adaptors.py:509-569); ``None`` for an empty profile.
    Evaluated by the device kernel behind ``qms_calc_amounts`` -- pass it as ``calc_amount_function`` to
    ``Results.calc_amounts_from_rt_info_file``, or use ``on_device=True`` there to batch all lines in one pass."""
    from collections import namedtuple

    from .isotopologue_library import calc_amounts_of_profiles

    point = namedtuple("point", ["rt", "scaling_factor", "score"])
    profile = [point(*row) for row in zip(obj_for_calc_amount["rt"], obj_for_calc_amount["i"], obj_for_calc_amount["scores"])]
    return calc_amounts_of_profiles([profile])[0]
