"""Stand-in for the un-vendored PyPI dependency ``chemical_composition``.

TEST INFRASTRUCTURE ONLY.  The reference (`/root/reference/pyqms/isotopologue_library.py:31`)
imports this module solely as its molecule-string -> {element: count} parser; no hot-path
arithmetic lives in it.  This stand-in lets `tests/golden/make_golden.py` import the UNMODIFIED
reference in the build container to generate golden vectors.  It is never imported by the product.
"""
import re
import xml.etree.ElementTree as ET
from pathlib import Path

_TOKEN = re.compile(r"(?P<iso>\d*)(?P<el>[A-Z][a-z]*)(?:\((?P<n>-?\d+)\)|(?P<c>\d*))")
_RESIDUE = re.compile(r"([A-Z])(\d*)")


def _default_unimod():
    """unimod.xml of the pyqms package that imported this stand-in (the reference source tree in the build container, its
    copy under baseline/_ref on the GPU box)."""
    try:
        import pyqms
        found = Path(pyqms.__file__).resolve().parent / "kb" / "ext" / "unimod.xml"
        if found.exists():
            return found
    except Exception:
        pass
    return Path("/root/reference/pyqms/kb/ext/unimod.xml")


_DEFAULT_UNIMOD = _default_unimod()
_XML_CACHE = {}


def _parse_formula(text):
    out = {}
    text = text.replace(" ", "").lstrip("+")
    for m in _TOKEN.finditer(text):
        if not m.group("el"):
            continue
        n = m.group("n")
        if n is None:
            n = m.group("c") or "1"
        key = m.group("iso") + m.group("el")
        out[key] = out.get(key, 0) + int(n)
    return out


def _load_unimod(path):
    path = str(path)
    if path not in _XML_CACHE:
        table = {}
        ns = "{http://www.unimod.org/xmlns/schema/unimod_2}"
        for mod in ET.parse(path).getroot().iter(ns + "mod"):
            delta = mod.find(ns + "delta")
            comp = {}
            for el in delta.findall(ns + "element"):
                comp[el.get("symbol")] = comp.get(el.get("symbol"), 0) + int(el.get("number"))
            table[mod.get("title")] = comp
        _XML_CACHE[path] = table
    return _XML_CACHE[path]


class ChemicalComposition(dict):
    def __init__(self, sequence=None, formula=None, aa_compositions=None, isotopic_distributions=None,
                 unimod_file_list=None, add_default_files=True, **kw):
        super().__init__()
        if aa_compositions is None:      # the real package ships its own residue table; the reference's is the same chemistry
            import pyqms.knowledge_base as _kb
            aa_compositions = _kb.aa_compositions
        self.aa_compositions = aa_compositions
        self.isotopic_distributions = isotopic_distributions if isotopic_distributions is not None else {}
        files = [str(f) for f in (unimod_file_list or [])]
        if add_default_files and str(_DEFAULT_UNIMOD) not in files:
            files.append(str(_DEFAULT_UNIMOD))
        self._unimod_files = files
        if formula is not None:
            self.add_chemical_formula(formula)
        if sequence is not None:
            self.use(sequence=sequence)

    def __missing__(self, key):
        return 0

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)

    def add_chemical_formula(self, formula, factor=1):
        comp = _parse_formula(formula) if isinstance(formula, str) else formula
        for k, v in comp.items():
            self[k] = self[k] + factor * v

    def _lookup_mod(self, name):
        for f in self._unimod_files:
            table = _load_unimod(f)
            if name in table:
                return table[name]
        raise KeyError("unknown unimod modification %r" % name)

    def use(self, sequence=None, modifications=None, deprecated_format=None, **kw):
        self.clear()
        if deprecated_format is not None:
            bits = deprecated_format.split("#")
            sequence = bits[0]
            modifications = bits[1] if len(bits) > 1 else None
        addon = None
        if "+" in sequence:
            sequence, addon = sequence.split("+", 1)
        if sequence:
            for aa, idx in _RESIDUE.findall(sequence):
                comp = self.aa_compositions[aa + idx]
                self.add_chemical_formula(comp if not isinstance(comp, str) else comp)
            self.add_chemical_formula("H2O")
        if addon:
            self.add_chemical_formula(addon)
        if modifications:
            for mod in modifications.split(";"):
                if not mod:
                    continue
                name = mod.rsplit(":", 1)[0]
                self.add_chemical_formula(self._lookup_mod(name))
        for k in [k for k, v in self.items() if v == 0]:
            del self[k]

    def generate_cc_dict(self):
        out = ChemicalComposition(aa_compositions=self.aa_compositions,
                                  isotopic_distributions=self.isotopic_distributions,
                                  unimod_file_list=self._unimod_files, add_default_files=False)
        out.update(self)
        return out

    def hill_notation_unimod(self):
        keys = [k for k in ("C", "H") if k in self] + sorted(k for k in self if k not in ("C", "H"))
        return "".join("%s(%d)" % (k, self[k]) for k in keys if self[k] != 0)
