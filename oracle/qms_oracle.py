"""CPU ORACLE for the two hot paths of pyQms -- TEST INFRASTRUCTURE, NOT PRODUCT.

This file restates, in plain Python, the arithmetic of the reference
`/root/reference/pyqms/isotopologue_library.py` (v0.6.5; abbreviated IL below) for
  P1  isotopologue-library construction   (IL:212-956, 1078-1324, 1679-1792, 2109-2165, 2533-2552)
  P2  match_all / match_isotopologue / score_matches   (IL:1794-2040, 2441-2603)
keeping the reference's operation order wherever a floating-point result or a threshold
decision depends on it, including the behaviours listed as quirks Q1-Q16 in SURVEY.md section 3.4.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  The product (`pyqms_b200/`) never does and has no CPU fallback.

PARITY PINNED: `tests/test_oracle_golden.py` checks this oracle bit-exactly against vectors
generated here by importing the UNMODIFIED reference (`tests/golden/make_golden.py`; library
dumps, match_all results, score KATs) and against the reference's own golden vectors
(`tests/peptide_sequence_test.py:12-37`, `docs/source/quick_start.rst:119-136`,
`tests/data/test_BSA_pyqms_results.pkl`).

The molecule-string parser below replaces the un-vendored PyPI dependency
`chemical_composition` (unpinned, requirements.txt:8); it holds no hot-path arithmetic.
"""
import bisect
import copy
import json
import operator
import re
import sys
import xml.etree.ElementTree as ET
from collections import namedtuple
from pathlib import Path

import numpy as np

EPS = sys.float_info.epsilon
_HERE = Path(__file__).resolve().parent
_KB_DIR = _HERE.parent / "pyqms_b200"

m_key = namedtuple("m_key", ["file_name", "formula", "charge", "label_percentiles"])
match = namedtuple("match", ["spec_id", "rt", "score", "scaling_factor", "peaks"])

# ---------------------------------------------------------------- constants (data, not code)
DEFAULT_PARAMS = {  # pyqms/params.py:29-74 (numeric defaults only; COLORS is plotting)
    "PERCENTILE_FORMAT_STRING": "{0:.3f}",
    "M_SCORE_THRESHOLD": 0.5,
    "ELEMENT_MIN_ABUNDANCE": 1e-3,
    "MIN_REL_PEAK_INTENSITY_FOR_MATCHING": 0.01,
    "REQUIRED_PERCENTILE_PEAK_OVERLAP": 0.50,
    "MINIMUM_NUMBER_OF_MATCHED_ISOTOPOLOGUES": 2,
    "INTENSITY_TRANSFORMATION_FACTOR": 1e5,
    "UPPER_MZ_LIMIT": 2000,
    "LOWER_MZ_LIMIT": 150,
    "MZ_TRANSFORMATION_FACTOR": 10000,
    "REL_MZ_RANGE": 5e-6,
    "REL_I_RANGE": 0.2,
    "INTERNAL_PRECISION": 1000,
    "MAX_MOLECULES_PER_MATCH_BIN": 20,
    "MZ_SCORE_PERCENTILE": 0.4,
    "SILAC_AAS_LOCKED_IN_EXPERIMENT": None,
    "BUILD_RESULT_INDEX": True,
    "MACHINE_OFFSET_IN_PPM": 0.0,
    "FIXED_LABEL_ISOTOPE_ENRICHMENT_LEVELS": {"15N": 0.994, "13C": 0.996, "2H": 0.994},
}


def _load_kb():
    ns = {}
    exec((_KB_DIR / "knowledge_base.py").read_text(), ns)  # data-only module
    return ns["isotopic_distributions"], ns["aa_compositions"], ns["PROTON"]


KB_ISOTOPES, KB_AMINO_ACIDS, PROTON = _load_kb()
_UNIMOD = None
_ISO_EL = re.compile(r"(?P<isotope>[0-9]*)(?P<element>[A-Z][a-z]*)")
_ISO_EL_CNT = re.compile(r"(?P<isotope>[0-9]*)(?P<element>[A-Z][a-z]*)\((?P<count>[-]*[0-9]*)\)")
_TOKEN = re.compile(r"(\d*)([A-Z][a-z]*)(?:\((-?\d+)\)|(\d*))")
_RESIDUE = re.compile(r"([A-Z])(\d*)")


# ---------------------------------------------------------------- molecule parser (host glue)
class Composition(dict):
    """{element: count}; missing keys read 0 (IL:1473 relies on that)."""

    def __missing__(self, key):
        return 0

    def add(self, other, factor=1):
        if isinstance(other, str):
            other = parse_formula(other)
        for k, v in other.items():
            self[k] = self[k] + factor * v
        return self

    def hill(self):
        keys = [k for k in ("C", "H") if k in self] + sorted(k for k in self if k not in ("C", "H"))
        return "".join("%s(%d)" % (k, self[k]) for k in keys if self[k] != 0)


def parse_formula(text):
    out = Composition()
    for iso, el, n, c in _TOKEN.findall(text.replace(" ", "").lstrip("+")):
        out[iso + el] += int(n if n != "" else (c or "1"))
    return out


def unimod_table(extra_files=None, add_default=True):
    global _UNIMOD
    if _UNIMOD is None:
        _UNIMOD = json.loads((_KB_DIR / "kb" / "unimod_table.json").read_text())["mods"]
    tables = []
    tag = "{http://www.unimod.org/xmlns/schema/unimod_2}"
    for f in extra_files or []:
        t = {}
        for mod in ET.parse(str(f)).getroot().iter(tag + "mod"):
            comp = {}
            for el in mod.find(tag + "delta").findall(tag + "element"):
                comp[el.get("symbol")] = comp.get(el.get("symbol"), 0) + int(el.get("number"))
            t[mod.get("title")] = comp
        tables.append(t)
    if add_default:
        tables.append(_UNIMOD)
    return tables


def molecule_to_cc(molecule, aa_compositions, tables):
    seq, _, mods = molecule.partition("#")
    seq, _, addon = seq.partition("+")
    cc = Composition()
    if seq:
        for aa, idx in _RESIDUE.findall(seq):
            cc.add(aa_compositions[aa + idx])
        cc.add("H2O")
    if addon:
        cc.add(addon)
    for mod in filter(None, mods.split(";")):
        name = mod.rsplit(":", 1)[0]
        for t in tables:
            if name in t:
                cc.add(t[name])
                break
        else:
            raise KeyError("unknown modification %r" % name)
    for k in [k for k, v in cc.items() if v == 0]:
        del cc[k]
    return cc


def cartesian_positions(ranges):
    """IL:1028-1076 -- all index combinations, first token slowest."""
    combos = [[]]
    for token, n in ranges:
        combos = [c + [(token, i)] for c in combos for i in range(n)]
    return combos if ranges else []


# ---------------------------------------------------------------- P1: library build
class OracleLibrary(dict):
    """Restatement of IL:212-956; same dict layout as the reference library."""

    def __init__(self, molecules, charges, metabolic_labels=None, fixed_labels=None, params=None,
                 unimod_files=None, add_default_files=True, build=True):
        self.params = copy.deepcopy(DEFAULT_PARAMS)
        self.params.update(params or {})
        P = self.params
        self.fmt = P["PERCENTILE_FORMAT_STRING"].format
        self.zero = self.fmt(0)
        self.metabolic_labels = metabolic_labels or {"15N": [0.0]}          # IL:233
        self.fixed_labels = fixed_labels or {}
        self.charges = charges
        self.lookup = {"molecule to formula": {}, "formula to molecule": {},
                       "molecule fixed label variations": {}}
        # IL:997-1026  knowledge-base cache; pos = enumeration index, then sort by mass (Q1)
        self.aa_compositions = {aa: parse_formula(f) for aa, f in KB_AMINO_ACIDS.items()}
        for aa, f in P.get("AMINO_ACIDS", {}).items():
            self.aa_compositions[aa] = parse_formula(f)
        self.isotopic_distributions = {
            el: {self.zero: sorted((m, a, pos) for pos, (m, a) in enumerate(dist))}
            for el, dist in KB_ISOTOPES.items()}
        # IL:1326-1373  metabolic label pools
        self.metabolically_labeled_elements = set()
        for iso_el, pcts in self.metabolic_labels.items():
            for pct in pcts:
                if pct <= EPS:
                    continue
                mt = _ISO_EL.match(iso_el)
                el = mt.group("element")
                self.isotopic_distributions[el][self.fmt(pct)] = self.recalc_distribution(
                    el, pct, int(round(float(mt.group("isotope")))))
                self.metabolically_labeled_elements.add(el)
        if not self.metabolically_labeled_elements:
            self.metabolically_labeled_elements.add("N")
        # IL:958-995  label-percentile tuples
        self.labled_percentiles = []
        for combo in cartesian_positions([(k, len(v)) for k, v in self.metabolic_labels.items()]):
            d = {}
            for name, idx in combo:
                d[_ISO_EL.match(name).group("element")] = self.fmt(self.metabolic_labels[name][idx])
            self.labled_percentiles.append(tuple(sorted(d.items())))
        if self.fixed_labels:
            self._extend_kb_with_fixed_labels()
            molecules = self.extend_molecules_with_fixed_labels(molecules)
        self._parse_molecules(molecules, unimod_table(unimod_files, add_default_files))
        self.formulas_sorted_by_mz = []
        self.match_sets = {}
        self.match_set_mz_range = [None, None]
        if build:
            self.create_element_trees()
            for formula in self.keys():
                self._build_formula(formula)
            self._build_index()

    # IL:2109-2165
    def recalc_distribution(self, element, target_percentile, enriched_isotope):
        natural = self.isotopic_distributions[element][self.zero]
        others = 0
        for mass, abundance, pos in natural:
            if int(round(mass)) != int(enriched_isotope):
                others += abundance
                continue
            natural_abundance = abundance
        diff = natural_abundance - target_percentile
        new = []
        for mass, abundance, pos in natural:
            share = abundance * diff / others
            if int(round(mass)) == int(enriched_isotope):
                abundance = target_percentile
            else:
                abundance += share
            new.append((mass, abundance, pos))
        return new

    # IL:1375-1499
    def _extend_kb_with_fixed_labels(self, synthetic_abundance=0.994):
        levels = self.params["FIXED_LABEL_ISOTOPE_ENRICHMENT_LEVELS"]
        for aa, definitions in self.fixed_labels.items():
            for idx, unimod_string in enumerate(definitions):
                if aa not in self.aa_compositions:
                    raise SystemExit(1)
                comp = copy.deepcopy(self.aa_compositions[aa])
                for isotope, element, count in _ISO_EL_CNT.findall(unimod_string):
                    if isotope == "":
                        key = element
                    else:
                        top = sorted(self.isotopic_distributions[element][self.zero],
                                     key=operator.itemgetter(1), reverse=True)[0]
                        natural_name = str(round(top[0])).split(".")[0]
                        percentile = 0.0 if isotope == natural_name else synthetic_abundance   # Q6
                        key = isotope + element
                        if key in levels:
                            percentile = levels[key]
                        else:
                            levels[key] = percentile
                    comp[key] += int(count)
                    if isotope != "":
                        self.isotopic_distributions.setdefault(key, {})[self.fmt(percentile)] = \
                            self.recalc_distribution(element, percentile, isotope)
                for k in [k for k, v in comp.items() if v == 0]:
                    del comp[k]
                self.aa_compositions["%s%d" % (aa, idx)] = comp

    # IL:1501-1677
    def extend_molecules_with_fixed_labels(self, molecules):
        out = set()
        locked = self.params["SILAC_AAS_LOCKED_IN_EXPERIMENT"]
        if locked is not None:
            assert len(locked) != 1
            assert len({len(self.fixed_labels[a]) for a in locked}) == 1
        for full in molecules:
            cut = len(full)
            for i, ch in enumerate(full):
                if ch not in self.aa_compositions:
                    cut = i
                    break
            molecule, addon = full[:cut], full[cut:]
            sites = []
            for aa in self.fixed_labels:
                for mt in re.finditer(aa, molecule):
                    sites.append(("%d%s" % (mt.start(), aa), len(self.fixed_labels[aa])))
            if not sites:
                out.add(molecule + addon)
                continue
            for combo in cartesian_positions(sorted(sites, reverse=True)):
                exchange = sorted(((int(tok[:-1]), tok[-1], idx) for tok, idx in combo), reverse=True)
                mod = molecule
                states = set()
                for pos, aa, idx in exchange:
                    mod = "%s%s%d%s" % (mod[:pos], aa, idx, mod[pos + 1:])
                    if locked is not None and aa in locked:
                        states.add(idx)
                keep = True
                if locked is not None:
                    keep = False
                    if len(states) == 1:
                        seen = {int(s) for a, s in re.findall(r"([A-Z]{1})([0-9]+)", mod) if a in locked}
                        keep = seen == states
                if keep:
                    out.add(mod + addon)
                    self.lookup["molecule fixed label variations"].setdefault(full, set()).add(mod + addon)
        return out

    # IL:323-497
    def _parse_molecules(self, molecules, tables):
        P = self.params
        self._highest = {}
        self._two_iso_range = [None, None]
        for molecule in list(set(molecules)):
            if molecule.count("#") > 1:
                raise ValueError("%s contains too many '#'" % molecule)
            cc = molecule_to_cc(molecule, self.aa_compositions, tables)
            formula = cc.hill()
            self.lookup["molecule to formula"][molecule] = formula
            self.lookup["formula to molecule"].setdefault(formula, []).append(molecule)
            if formula not in self:
                self[formula] = {"env": {}, "cc": cc}
            for lp in self.labled_percentiles:
                self[formula]["env"][lp] = {}
            for element in cc:
                if element not in self.isotopic_distributions:          # IL:378-436
                    mt = _ISO_EL.match(element)
                    try:
                        iso = int(round(float(mt.group("isotope"))))
                    except Exception:
                        raise SystemExit(1)
                    template = mt.group("element")
                    pct = P["FIXED_LABEL_ISOTOPE_ENRICHMENT_LEVELS"].get("%d%s" % (iso, template)) or 0.994
                    self.isotopic_distributions[element] = {
                        self.zero: self.recalc_distribution(template, pct, iso)}
            for element, count in cc.items():
                if self._highest.get(element, 0) < count:
                    self._highest[element] = count
                self._highest.setdefault(element, 0)
                for lp in self.isotopic_distributions[element]:
                    if len(self.isotopic_distributions[element][lp]) == 2:
                        r = self._two_iso_range
                        if r[0] is None or count < r[0]:
                            r[0] = count
                        if r[1] is None or count > r[1]:
                            r[1] = count
                        break
        for element in list(self._highest):                               # IL:467-497
            if not element.isalpha():
                template = _ISO_EL.match(element).group("element")
                self._highest.setdefault(template, 0)
                self._highest[template] += self._highest.get(element, 0)
                for lp in self.isotopic_distributions[template]:
                    if len(self.isotopic_distributions[template][lp]) == 2:
                        if self._highest[template] > self._two_iso_range[1]:
                            self._two_iso_range[1] = self._highest[template]

    # IL:1078-1214
    def create_element_trees(self):
        if self._two_iso_range[0] is not None:
            lo, hi = self._two_iso_range
            fact = [1]
            for n in range(1, hi + 2):
                fact.append(fact[-1] * n)
            self._binomials = {n: [fact[n] // (fact[k] * fact[n - k]) for k in range(n + 1)]
                               for n in range(lo, hi + 1)}               # exact ints, IL:1121-1140
        self.element_trees = {}
        for element, count in self._highest.items():
            dists = self.isotopic_distributions[element]
            trees = self.element_trees[element] = {}
            for lp in sorted(dists, reverse=True):
                trees[lp] = {1: {"maxPos": dists[lp][-1][2], "minPos": dists[lp][0][2],
                                 "env": {pos: {"mass": m, "abun": a} for m, a, pos in dists[lp]}}}
            for lp in dists:
                k_iso = len(dists[lp])
                if k_iso == 1:
                    for level in range(2, count + 1):
                        trees[lp][level] = {"env": {0: {"abun": 1, "mass": trees[lp][level - 1]["env"][0]["mass"]
                                                        + dists[lp][0][0]}}, "maxPos": 0, "minPos": 0}
                elif k_iso == 2:
                    self._two_isotope_levels(element, lp, count)
                else:
                    self._multi_isotope_levels(element, lp, count)

    # IL:1216-1324 (Q2, Q4): binomial pmf with exact C(n,k), lazily filled power caches
    def _two_isotope_levels(self, element, lp, count):
        dist = self.isotopic_distributions[element][lp]
        (a_mass, a, _), (b_mass, b, _) = dist[0], dist[1]
        a_pow, b_pow = {1: a}, {1: b}
        a_msum, b_msum = {1: a_mass}, {1: b_mass}
        thr = self.params["ELEMENT_MIN_ABUNDANCE"]
        k_start = 0
        for n in range(self._two_iso_range[0], count + 1):
            level = self.element_trees[element][lp][n] = {"env": {}, "maxPos": None, "minPos": None}
            seen_nonzero = ended = False
            for k in range(k_start, n + 1):
                node = level["env"][k] = {"mass": None, "abun": 0}
                if ended:
                    continue
                m = n - k
                if m not in a_pow:
                    if k != n:
                        a_pow[m] = a * a_pow[m - 1] if (m - 1) in a_pow else a ** m
                    else:
                        a_pow[m] = 1
                if k not in b_pow:
                    b_pow[k] = b * b_pow[k - 1] if k != 0 else 1
                if m not in a_msum:
                    if k != n:
                        a_msum[m] = a_mass + a_msum[m - 1] if (m - 1) in a_msum else a_mass * m
                    else:
                        a_msum[m] = 0
                if k not in b_msum:
                    b_msum[k] = b_mass + b_msum[k - 1] if k != 0 else 0
                node["abun"] = a_pow[m] * b_pow[k] * self._binomials[n][k]
                node["mass"] = a_msum[m] + b_msum[k]
                if node["abun"] <= thr:
                    node["abun"] = 0
                    if seen_nonzero:
                        ended = True
                    else:
                        k_start = k + 1
                else:
                    level["maxPos"] = k
                    if not seen_nonzero:
                        seen_nonzero = True
                        level["minPos"] = k

    # IL:1679-1792 (Q3, Q5): unpruned self-convolution, unweighted mean path mass
    def _multi_isotope_levels(self, element, lp, count):
        dist = self.isotopic_distributions[element][lp]
        trees = self.element_trees[element][lp]
        thr = self.params["ELEMENT_MIN_ABUNDANCE"]
        for n in range(2, count + 2):
            prev = trees[n - 1]["env"]
            env = {}
            for pos in sorted(prev):
                for iso_mass, _, q in dist:
                    node = env.setdefault(pos + q, {"mass": [], "abun": 0})
                    node["abun"] += prev[pos]["abun"] * dist[q][1]
                    node["mass"].append(prev[pos]["mass"] + dist[q][0])
            lo = hi = None
            for pos, node in env.items():
                node["mass"] = float(sum(node["mass"])) / float(len(node["mass"]))
                if node["abun"] > thr:
                    if lo is None:
                        lo = pos
                    hi = pos
            trees[n] = {"env": env, "minPos": lo, "maxPos": hi}

    # IL:518-853
    def _build_formula(self, formula):
        P = self.params
        levels_cfg = P["FIXED_LABEL_ISOTOPE_ENRICHMENT_LEVELS"]
        cc = self[formula]["cc"]
        for lp in list(self[formula]["env"].keys()):
            merged = {}
            for element, count in cc.items():                            # IL:546-588 (Q6 fold)
                target = element
                if element in levels_cfg:
                    bare = element[-1]
                    for el, pct in lp:
                        if el == bare and pct == self.fmt(levels_cfg[element]):
                            target = bare
                merged[target] = merged.get(target, 0) + count
            ranges, stats, zero_pos = [], {}, 0
            for element, level in merged.items():                        # IL:589-617
                if element in self.metabolically_labeled_elements:
                    continue
                label = list(self.element_trees[element].keys())[0]
                node = self.element_trees[element][label][level]
                ranges.append([element, node["maxPos"] - node["minPos"] + 1])
                zero_pos += node["minPos"]
                stats[element] = (label, level)
            for element, pct in lp:                                      # IL:626-646
                level = merged.get(element, 0)
                if level == 0:
                    continue
                node = self.element_trees[element][pct][level]
                ranges.append([element, node["maxPos"] - node["minPos"] + 1])
                stats[element] = (pct, level)
                zero_pos += node["minPos"]
            acc = {}
            for combo in cartesian_positions(ranges):                    # IL:649-702
                pos_sum = sum([p for _, p in combo]) + zero_pos
                slot = acc.setdefault(pos_sum, {"abun": [], "mass": []})
                abun, mass = 1, 0
                for element, p in combo:
                    label, level = stats[element]
                    node = self.element_trees[element][label][level]
                    abun *= node["env"][p + node["minPos"]]["abun"]
                    mass += node["env"][p + node["minPos"]]["mass"]
                slot["abun"].append(abun)
                slot["mass"].append(mass)
            env = self[formula]["env"][lp] = {"isot": [], "mass": [], "abun": [], "relabun": [],
                                              "c_peak_pos": [], "n_c_peaks": 0}
            for z in self.charges:
                env[z] = {"mz": [], "tmzs": [], "atmzs": set()}
            top = 0
            for pos in sorted(acc):                                      # IL:721-725
                total = sum(acc[pos]["abun"])
                if total > top:
                    top = total
            for pos in sorted(acc):                                      # IL:727-811
                total = sum(acc[pos]["abun"])
                if total < EPS:
                    continue
                mass = 0
                for n, m in enumerate(acc[pos]["mass"]):
                    mass += m * acc[pos]["abun"][n] / float(total)
                env["mass"].append(mass)
                env["abun"].append(int(round(total * P["INTENSITY_TRANSFORMATION_FACTOR"])))
                rel = total / float(top)
                env["relabun"].append(rel)
                c_peak = rel >= P["MIN_REL_PEAK_INTENSITY_FOR_MATCHING"]
                if c_peak:
                    env["n_c_peaks"] += 1.0
                env["c_peak_pos"].append(pos if c_peak else None)
                for z in self.charges:
                    mz = (mass + z * PROTON) / float(abs(z))
                    if P["MACHINE_OFFSET_IN_PPM"] != 0:
                        mz = mz + mz * 1e-6 * P["MACHINE_OFFSET_IN_PPM"]
                    env[z]["mz"].append(mz)
                    if c_peak:
                        window = self.mz_to_set(mz)
                        env[z]["tmzs"].append(window)
                        env[z]["atmzs"] |= window
                    else:
                        env[z]["tmzs"].append(None)
            for z in self.charges:                                       # IL:817-844 (Q15)
                lo, hi = env[z]["mz"][0], env[z]["mz"][-1]
                if P["LOWER_MZ_LIMIT"] <= lo and hi <= P["UPPER_MZ_LIMIT"]:
                    self.formulas_sorted_by_mz.append((lo, hi, z, lp, formula))

    # IL:866-939
    def _build_index(self):
        self.formulas_sorted_by_mz.sort()
        step = self.params["MAX_MOLECULES_PER_MATCH_BIN"]
        total = len(self.formulas_sorted_by_mz)
        for number, start in enumerate(range(0, total, step)):
            stop = min(start + step, total)
            entry = self.match_sets[number] = {"ids": [start, stop], "tmzs": set(), "mz_range": [None, None]}
            for index in range(start, stop):
                lo, hi, z, lp, formula = self.formulas_sorted_by_mz[index]
                if step != 1:
                    entry["tmzs"] |= self[formula]["env"][lp][z]["atmzs"]
                    if entry["mz_range"][0] is None or lo < entry["mz_range"][0]:
                        entry["mz_range"][0] = lo
                    if entry["mz_range"][1] is None or hi > entry["mz_range"][1]:
                        entry["mz_range"][1] = hi
                else:
                    entry["tmzs"] = self[formula]["env"][lp][z]["atmzs"]
                    entry["mz_range"] = [lo, hi]
        for entry in self.match_sets.values():
            r = self.match_set_mz_range
            if r[0] is None or entry["mz_range"][0] < r[0]:
                r[0] = entry["mz_range"][0]
            if r[1] is None or entry["mz_range"][1] > r[1]:
                r[1] = entry["mz_range"][1]

    # IL:2533-2552 (Q8)
    def mz_to_set(self, mz):
        err = mz * self.params["REL_MZ_RANGE"]
        lo = (mz - err) * self.params["INTERNAL_PRECISION"]
        hi = (mz + err) * self.params["INTERNAL_PRECISION"]
        return set(range(int(round(lo)), int(round(hi)) + 1))

    # ------------------------------------------------------------ P2: matching
    # IL:2500-2531 (Q12)
    @staticmethod
    def slice_list(source, borders, tolerance=2):
        lower, upper = list(borders[0]), list(borders[1])
        if getattr(source, "tolist", False) is False:
            try:
                lo = bisect.bisect(source, lower) - tolerance
                hi = bisect.bisect(source, upper) + tolerance
            except Exception:
                lo = bisect.bisect(source, tuple(lower)) - tolerance
                hi = bisect.bisect(source, tuple(upper)) + tolerance
            return source[max(lo, 0):min(hi, len(source))]
        return source[np.where((lower[0] - tolerance < source[:, 0]) * (source[:, 0] < upper[0] + tolerance))]

    # IL:2554-2603 (Q9)
    def transform_spectrum(self, peaks, mz_range=None):
        lookup = {}
        target = peaks if mz_range is None else self.slice_list(peaks, [(x, 0.001) for x in mz_range])
        precision = self.params["INTERNAL_PRECISION"]
        if getattr(target, "tolist", False) is False:
            keys = set()
            for mz, intensity in target:
                t = int(round(mz * precision))
                keys.add(t)
                lookup.setdefault(t, []).append((mz, intensity))
        else:
            tmz = np.round(target[:, 0] * precision).astype(int)
            for i, t in enumerate(tmz):
                lookup.setdefault(t, []).append((target[i][0], target[i][1]))
            keys = set(tmz)
        return keys, lookup

    # IL:1794-1883
    def match_all(self, mz_i_list, file_name=None, spec_id=None, spec_rt=None, results=None,
                  mz_score_percentile=None):
        if results is None:
            results = {}
        xi = self.params["MZ_SCORE_PERCENTILE"] if mz_score_percentile is None else mz_score_percentile
        P = self.params
        sliced = self.slice_list(mz_i_list, ((self.match_set_mz_range[0], 0), (self.match_set_mz_range[1], 0)))
        for number in self.match_sets:
            mset = self.match_sets[number]
            keys, lookup = self.transform_spectrum(sliced, mz_range=mset["mz_range"])
            if len(keys & mset["tmzs"]) < P["MINIMUM_NUMBER_OF_MATCHED_ISOTOPOLOGUES"]:
                continue
            for index in range(mset["ids"][0], mset["ids"][1]):
                found = self.match_isotopologue(index, keys, lookup, xi)
                if found is None:
                    continue
                score, scaling, peaks = found
                if score < P["M_SCORE_THRESHOLD"]:
                    continue
                if sum(1 for p in peaks if p[0] is not None) < P["MINIMUM_NUMBER_OF_MATCHED_ISOTOPOLOGUES"]:
                    continue
                lo, hi, z, lp, formula = self.formulas_sorted_by_mz[index]
                results.setdefault(m_key(file_name, formula, z, lp), []).append(
                    match(spec_id, spec_rt, score, scaling, peaks))
        return results

    # IL:1957-1966: match_isotopologue(index=, mz_i_list=): the spectrum is sliced by the entry's own m/z range first
    def match_isotopologue_of_spectrum(self, index, mz_i_list, xi):
        lower, upper = self.formulas_sorted_by_mz[index][:2]
        sliced = self.slice_list(mz_i_list, ((lower, 0), (upper, 0)))
        keys, lookup = self.transform_spectrum(sliced, mz_range=(lower, upper))
        return self.match_isotopologue(index, keys, lookup, xi)

    # IL:1885-2040 (Q10, Q11)
    def match_isotopologue(self, index, keys, lookup, xi):
        lo, hi, z, lp, formula = self.formulas_sorted_by_mz[index]
        env = self[formula]["env"][lp]
        overlap = env[z]["atmzs"] & keys
        P = self.params
        if len(overlap) < P["MINIMUM_NUMBER_OF_MATCHED_ISOTOPOLOGUES"]:
            return None
        if len(overlap) / env["n_c_peaks"] < P["REQUIRED_PERCENTILE_PEAK_OVERLAP"]:
            return None
        rows, cands = {}, {}
        for n, pos in enumerate(env["c_peak_pos"]):
            if pos is None:
                continue
            rows[n] = [None, None, env["relabun"][n], env[z]["mz"][n], env["abun"][n]]
            cands[n] = list(overlap & env[z]["tmzs"][n])
        scored = []
        for combo in cartesian_positions([(n, len(c)) for n, c in sorted(cands.items()) if len(c) > 0]):
            for n, i in combo:
                rows[n][0], rows[n][1] = lookup[cands[n][i]][0]
            score, scaling = self.score_matches(rows.values(), xi)
            scored.append((score, scaling, tuple(tuple(r) for r in rows.values())))
        scored.sort(reverse=True)
        return scored[0] if scored else None

    # IL:2441-2498
    def score_matches(self, matched_peaks, xi):
        P = self.params
        floor = P["MIN_REL_PEAK_INTENSITY_FOR_MATCHING"]
        measured = calculated = ri_sum = 0
        for mmz, mi, ri, cmz, ci in matched_peaks:
            if ri < floor:
                continue
            if mmz is not None:
                calculated += ci * ri
                measured += mi * ri
            ri_sum += ri
        scaling = measured / float(calculated)
        mz_score = i_score = 0
        for mmz, mi, ri, cmz, ci in matched_peaks:
            if ri < floor or mmz is None:
                continue
            si = ci * scaling
            if cmz > EPS:
                err = abs(mmz - cmz) / cmz
                mz_score += (0 if err > P["REL_MZ_RANGE"] else 1 - (err / P["REL_MZ_RANGE"])) * ri
            if si > EPS:
                err = abs(mi - si) / si
                span = 1.0 + P["REL_I_RANGE"] - ri
                i_score += (0 if err > span else 1 - (err / span)) * ri
        mz_score = mz_score / ri_sum
        i_score = i_score / ri_sum
        return mz_score * xi + i_score * (1 - xi), scaling


# ---------------------------------------------------------------- next row N2: amounts
def calc_amount(profile):
    """pyqms/adaptors.py:509-569 (calc_amount_function): max / sum / area under curve of one profile
    {"rt": [...], "i": [...], "scores": [...]}; None for an empty profile."""
    if len(profile["i"]) == 0:
        return None
    top = max(profile["i"])
    at = profile["i"].index(top)
    out = {"max I in window": top, "max I in window (rt)": profile["rt"][at], "max I in window (score)": profile["scores"][at],
           "sum I in window": sum(profile["i"]), "auc in window": 0}
    for pos, y1 in enumerate(profile["i"]):
        if pos == 0:
            continue
        x0, y0, x1 = profile["rt"][pos - 1], profile["i"][pos - 1], profile["rt"][pos]
        xspace = x1 - x0
        triangle = 0.5 * (xspace * abs(y1 - y0))
        out["auc in window"] += xspace * y0
        if y0 < y1:
            out["auc in window"] += triangle
        elif y0 > y1:
            out["auc in window"] -= triangle
    return out


def windowed_profile(matches, start=None, stop=None):
    """pyqms/results.py:1238-1262: the rt window walk of calc_amounts_from_rt_info_file over one key's matches."""
    profile = {"rt": [], "i": [], "scores": [], "spec_ids": []}
    for entry in matches:
        try:
            rt = entry.rt[0] / 60 if entry.rt[1] == "second" else entry.rt[0]
        except TypeError:
            rt = entry.rt
        if start is not None and rt < start:
            continue
        elif stop is not None and rt > stop:
            break
        profile["rt"].append(rt)
        profile["i"].append(entry.scaling_factor)
        profile["scores"].append(entry.score)
        profile["spec_ids"].append(entry.spec_id)
    return profile
