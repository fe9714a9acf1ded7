"""Molecule strings -> element counts (host-side string work; no hot-path arithmetic).

The reference delegates this to the PyPI package ``chemical_composition`` (not vendored, unpinned:
requirements.txt:8, call sites IL:31, 314-347, 1006-1013).  This module is an independent
implementation of the subset the library build needs:

* formulas          ``+H2O``, ``C(2)H(3)14N(1)O(1)``, ``C(-6) 13C(6)`` (isotope prefix, signed counts)
* peptides          ``PEPTIDE`` = sum of residues + H2O; residues may carry a fixed-label index (``K0``)
* add-ons           ``PEPTIDE+PO3``
* modifications     ``PEPTIDE#Oxidation:1;Phospho:3`` by Unimod title (position is informational)
* Hill notation     ``C(34)H(53)N(7)O(15)``: C, H, then ``sorted()`` of the remaining symbols
"""
import json
import re
import xml.etree.ElementTree as ET
from pathlib import Path

_TOKEN = re.compile(r"(\d*)([A-Z][a-z]*)(?:\((-?\d+)\)|(\d*))")
_RESIDUE = re.compile(r"([A-Z])(\d*)")
_UNIMOD_NS = "{http://www.unimod.org/xmlns/schema/unimod_2}"
_DEFAULT_TABLE = None
_XML_TABLES = {}


class ChemicalComposition(dict):
    """``{element symbol: count}``; absent elements read as 0 (the library build does ``cc[k] += n``)."""

    def __init__(self, formula=None, sequence=None, aa_compositions=None, isotopic_distributions=None,
                 unimod_file_list=None, add_default_files=True, **_ignored):
        super().__init__()
        self.aa_compositions = {} if aa_compositions is None else aa_compositions
        self.isotopic_distributions = {} if isotopic_distributions is None else isotopic_distributions
        self.unimod_file_list = list(unimod_file_list or [])
        self.add_default_files = add_default_files
        if formula is not None:
            self.add_chemical_formula(formula)
        if sequence is not None:
            self.use(sequence=sequence)

    def __missing__(self, key):
        return 0

    def __reduce__(self):
        return (_rebuild, (dict(self),))

    def add_chemical_formula(self, formula, factor=1):
        items = parse_formula(formula).items() if isinstance(formula, str) else formula.items()
        for symbol, count in items:
            self[symbol] = self[symbol] + factor * count

    def subtract_chemical_formula(self, formula, factor=1):
        self.add_chemical_formula(formula, -factor)

    def use(self, sequence=None, modifications=None, deprecated_format=None):
        """Load a molecule; ``deprecated_format`` is the reference's ``SEQ+ADDON#Mod:pos;Mod:pos`` string."""
        self.clear()
        if deprecated_format is not None:
            sequence, _, modifications = deprecated_format.partition("#")
        sequence = sequence or ""
        sequence, _, addon = sequence.partition("+")
        if sequence:
            for letter, index in _RESIDUE.findall(sequence):
                try:
                    self.add_chemical_formula(self.aa_compositions[letter + index])
                except KeyError:
                    raise KeyError("unknown amino acid %r in %r" % (letter + index, sequence))
            self.add_chemical_formula({"H": 2, "O": 1})
        if addon:
            self.add_chemical_formula(addon)
        for item in filter(None, (modifications or "").split(";")):
            title = item.rsplit(":", 1)[0] if ":" in item else item
            self.add_chemical_formula(self._modification(title))
        for symbol in [s for s, n in self.items() if n == 0]:
            del self[symbol]
        return self

    def composition_of(self, molecule):
        """A fresh ``ChemicalComposition`` of one molecule string (``use`` + ``generate_cc_dict`` without touching
        this object); plain ``+formula`` molecules skip the peptide machinery."""
        out = ChemicalComposition.__new__(ChemicalComposition)
        dict.__init__(out)
        out.aa_compositions = self.aa_compositions
        out.isotopic_distributions = self.isotopic_distributions
        out.unimod_file_list = self.unimod_file_list
        out.add_default_files = self.add_default_files
        if molecule[:1] == "+" and "#" not in molecule:
            for isotope, element, bracketed, bare in _TOKEN.findall(molecule.replace(" ", "")[1:]):
                key = isotope + element
                out[key] = out.get(key, 0) + (int(bracketed) if bracketed != "" else int(bare or 1))
            for symbol in [s for s, n in out.items() if n == 0]:
                del out[symbol]
            return out
        return out.use(deprecated_format=molecule)

    def _modification(self, title):
        for path in self.unimod_file_list:
            table = load_unimod_xml(path)
            if title in table:
                return table[title]
        if self.add_default_files:
            table = default_unimod_table()
            if title in table["mods"]:
                return table["mods"][title]
            if title in table["ids"]:
                return table["mods"][table["ids"][title]]
        raise KeyError("modification %r not found in the Unimod tables" % title)

    def generate_cc_dict(self):
        out = ChemicalComposition(aa_compositions=self.aa_compositions,
                                  isotopic_distributions=self.isotopic_distributions,
                                  unimod_file_list=self.unimod_file_list, add_default_files=self.add_default_files)
        out.update(self)
        return out

    def hill_notation_unimod(self):
        order = [s for s in ("C", "H") if s in self] + sorted(s for s in self if s not in ("C", "H"))
        return "".join("%s(%d)" % (s, self[s]) for s in order if self[s] != 0)

    def hill_notation(self):
        order = [s for s in ("C", "H") if s in self] + sorted(s for s in self if s not in ("C", "H"))
        return "".join("%s%s" % (s, self[s] if self[s] != 1 else "") for s in order if self[s] != 0)


def _rebuild(items):
    cc = ChemicalComposition()
    cc.update(items)
    return cc


def parse_formula(text):
    """``"C(2)H(3)14N(1)O(1)"`` / ``"+H2O"`` / ``"C(-6) 13C(6)"`` -> {symbol: count}."""
    counts = {}
    for isotope, element, bracketed, bare in _TOKEN.findall(text.replace(" ", "").lstrip("+")):
        n = int(bracketed) if bracketed != "" else int(bare or 1)
        counts[isotope + element] = counts.get(isotope + element, 0) + n
    return counts


def default_unimod_table():
    global _DEFAULT_TABLE
    if _DEFAULT_TABLE is None:
        path = Path(__file__).resolve().parent / "kb" / "unimod_table.json"
        _DEFAULT_TABLE = json.loads(path.read_text())
    return _DEFAULT_TABLE


def load_unimod_xml(path):
    """Title -> delta composition of every ``<umod:mod>`` of a Unimod-schema XML file (user files)."""
    key = str(path)
    if key not in _XML_TABLES:
        table = {}
        for mod in ET.parse(key).getroot().iter(_UNIMOD_NS + "mod"):
            delta = mod.find(_UNIMOD_NS + "delta")
            composition = {}
            if delta is not None:
                for el in delta.findall(_UNIMOD_NS + "element"):
                    composition[el.get("symbol")] = composition.get(el.get("symbol"), 0) + int(el.get("number"))
            table[mod.get("title")] = composition
        _XML_TABLES[key] = table
    return _XML_TABLES[key]
