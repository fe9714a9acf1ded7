"""``IsotopologueLibrary`` -- drop-in for ``pyqms.IsotopologueLibrary`` with both hot paths on the GPU.

Host side of the B200-native build.  The constructor signature, the ``params`` dict, the dict layout
``lib[formula] = {"cc", "env": {lp_tuple: {...}}}``, ``match_all(mz_i_list, file_name, spec_id, spec_rt,
results)`` and the ``Results`` / ``match`` structures follow the reference
(pyqms/isotopologue_library.py, "IL" below).  What runs where:

* host (this file): string work only -- molecule parsing, fixed-label expansion (IL:1375-1677), label
  percentile tuples (IL:958-995), the per-envelope (element pool, atom count) descriptors (IL:546-646),
  Python tie-break ranks for the index sort, and lazy dict/list views over arrays exported from the GPU.
* device (libqms_b200.so through ctypes, include/qms_b200.h): enriched isotope tables, element trees,
  envelope convolution, per-charge m/z + ppm windows, index sort + probe tables, spectrum probe,
  candidate assembly, mScore and hit compaction.  There is no CPU fallback.
"""
import contextvars
import copy
import ctypes as C
import operator
import re
import sys

import numpy as np

from . import knowledge_base, planner
from .params import params as _default_params
from ._capi import QMS_E_CAPACITY, Context, QmsHits, QmsIndexInfo, QmsParams, ptr
from ._pinned import PinnedPool
from .chemistry import ChemicalComposition
from .results import Results

_ISOTOPE_ELEMENT = re.compile(r"(?P<isotope>[0-9]*)(?P<element>[A-Z][a-z]*)")
_ISOTOPE_ELEMENT_COUNT = re.compile(r"(?P<isotope>[0-9]*)(?P<element>[A-Z][a-z]*)\((?P<count>[-]*[0-9]*)\)")
_NAN = float("nan")


def _cartesian(ranges):
    """All index combinations of [(token, n), ...], first token slowest (IL:1028-1076)."""
    if not ranges:
        return []
    combos = [[]]
    for token, n in ranges:
        combos = [c + [(token, i)] for c in combos for i in range(n)]
    return combos


class _LazyEnvelopes(dict):
    """``lib[formula]["env"]``: label tuple -> envelope dict, materialised from device exports on access."""

    def __init__(self, library, formula, label_tuples):
        super().__init__((lp, None) for lp in label_tuples)
        self._library = library
        self._formula = formula

    def __getitem__(self, lp):
        value = dict.__getitem__(self, lp)
        if value is None:
            value = self._library._materialise_envelope(self._formula, lp)
            dict.__setitem__(self, lp, value)
        return value

    def get(self, lp, default=None):
        return self[lp] if lp in self else default

    def values(self):
        return [self[lp] for lp in self.keys()]

    def items(self):
        return [(lp, self[lp]) for lp in self.keys()]


_restore_device = contextvars.ContextVar("qms_restore_device", default=None)


def _restore_library(state, formulas):
    library = IsotopologueLibrary.__new__(IsotopologueLibrary)
    dict.__init__(library)
    library.__dict__.update(state)
    dict.update(library, dict.fromkeys(formulas))           # records are made on first access
    library._ctx = Context(_restore_device.get())      # raises without the CUDA library / a GPU: no fallback
    library._views = {}
    library._push_params()
    verbose, library.verbose = library.verbose, False
    try:
        library._build_on_device(distributed=False)
    finally:
        library.verbose = verbose
    return library


class _LazyDict(dict):
    """A dict whose content is produced by ``fill(self)`` on first use: lookups of a large library that the build itself
    never reads (500 000 ``setdefault`` calls cost more than the device part of the build)."""

    def __init__(self, fill):
        super().__init__()
        self._fill = fill

    def _ensure(self):
        fill, self._fill = self._fill, None
        if fill is not None:
            fill(self)

    def __getitem__(self, key):
        self._ensure()
        return dict.__getitem__(self, key)

    def __iter__(self):
        self._ensure()
        return dict.__iter__(self)

    def __len__(self):
        self._ensure()
        return dict.__len__(self)

    def __contains__(self, key):
        self._ensure()
        return dict.__contains__(self, key)

    def __eq__(self, other):
        self._ensure()
        if isinstance(other, _LazyDict):
            other._ensure()
        return dict.__eq__(self, other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __repr__(self):
        self._ensure()
        return dict.__repr__(self)

    def keys(self):
        self._ensure()
        return dict.keys(self)

    def values(self):
        self._ensure()
        return dict.values(self)

    def items(self):
        self._ensure()
        return dict.items(self)

    def get(self, key, default=None):
        self._ensure()
        return dict.get(self, key, default)

    def setdefault(self, key, default=None):
        self._ensure()
        return dict.setdefault(self, key, default)

    def pop(self, *args):
        self._ensure()
        return dict.pop(self, *args)

    def update(self, *args, **kwargs):
        self._ensure()
        return dict.update(self, *args, **kwargs)

    def __setitem__(self, key, value):
        self._ensure()
        return dict.__setitem__(self, key, value)

    def copy(self):
        self._ensure()
        return dict(self)

    def __reduce__(self):
        self._ensure()
        return (dict, (dict(self),))


class _FormulaRecord(dict):
    """``lib[formula]``: ``{"env": label tuple -> envelope, "cc": composition}``, both built on first access (a
    library of 10^6 formulas does not create 2*10^6 dictionaries nobody reads)."""
    _FIELDS = ("env", "cc")
    __slots__ = ("_library", "_row", "_formula")

    def __init__(self, library, row, formula):          # dict.__new__ already made the (empty) mapping
        self._library = library
        self._row = row
        self._formula = formula

    def __missing__(self, key):
        if key == "env":
            value = _LazyEnvelopes(self._library, self._formula, self._library.labled_percentiles)
        elif key == "cc":
            value = self._library._composition_of_formula(self._row)
        else:
            raise KeyError(key)
        dict.__setitem__(self, key, value)
        return value

    def _fill(self):
        for key in self._FIELDS:
            self[key]
        return self

    def __contains__(self, key):
        return key in self._FIELDS or dict.__contains__(self, key)

    def __iter__(self):
        return iter(self._fill().keys())

    def __len__(self):
        return dict.__len__(self._fill())

    def keys(self):
        return dict.keys(self._fill())

    def values(self):
        return dict.values(self._fill())

    def items(self):
        return dict.items(self._fill())

    def get(self, key, default=None):
        return self[key] if key in self else default

    def __reduce__(self):
        return (dict, (dict(self._fill()),))


AMOUNT_FIELDS = ("max I in window", "max I in window (rt)", "max I in window (score)", "sum I in window", "auc in window")
_amount_ctx = None


def calc_amounts_of_profiles(profiles, windows=None, ctx=None):
    """One device pass (``qms_calc_amounts``) over many elution profiles.

    ``profiles[n]`` is a list of ``match`` tuples in acquisition order, ``windows[n] = (start_min, stop_min)`` or
    NaNs for the whole profile; returns one amount dict (``AMOUNT_FIELDS``) or ``None`` per profile, with the
    semantics of results.py:1238-1262 + adaptors.py:509-569 (retention times in minutes, the scan stops at the
    first match beyond the window).  ``ctx`` defaults to a module-wide context on the current device."""
    global _amount_ctx
    if ctx is None:
        if _amount_ctx is None:
            _amount_ctx = Context()
        ctx = _amount_ctx
    seg_off = np.zeros(len(profiles) + 1, dtype=np.int64)
    rts, intensities, scores = [], [], []
    for n, profile in enumerate(profiles):
        for entry in profile:
            try:                                        # results.py:1240-1246
                rt = entry.rt[0] / 60 if entry.rt[1] == "second" else entry.rt[0]
            except TypeError:
                rt = entry.rt
            rts.append(rt)
            intensities.append(entry.scaling_factor)
            scores.append(entry.score)
        seg_off[n + 1] = len(rts)
    rt = np.array(rts, dtype=np.float64)
    inten = np.array(intensities, dtype=np.float64)
    score = np.array(scores, dtype=np.float64)
    start = stop = None
    if windows is not None:
        start = np.array([w[0] for w in windows], dtype=np.float64)
        stop = np.array([w[1] for w in windows], dtype=np.float64)
    out = np.zeros((len(profiles), 5), dtype=np.float64)
    count = np.zeros(len(profiles), dtype=np.int32)
    ctx.call("qms_calc_amounts", len(profiles), ptr(seg_off), ptr(rt), ptr(inten), ptr(score), ptr(start), ptr(stop),
             ptr(out), ptr(count))
    return [dict(zip(AMOUNT_FIELDS, out[n].tolist())) if count[n] else None for n in range(len(profiles))]


class PackedHits(dict):
    """Packed hit arrays of one device pass: spec, entry, score, scaling (per hit), win_off (n_hits + 1), peak (row of the
    matched peak per hit window, -1 = None), n_hits, n_windows -- and ``mmz`` / ``mi``, the matched (m/z, intensity)
    values (NaN = None).  When the pass did not bring the values back from the device they are gathered from the
    spectra by ``peak`` row on first access (the spectra array is kept alive for that; do not overwrite it).

    A pass without the values also leaves the wide index columns on the device and brings their compact forms
    (``qms_hits``: hit_peak16, spec_hit_off): ``spec`` is then expanded from ``spec_off``, ``win_off`` summed from the
    window counts of the hits' entries and ``peak`` widened from ``peak16`` + the first row of the hit's spectrum, each
    on first access."""

    def __init__(self, peaks=None, offsets=None, entry_windows=None):
        super().__init__()
        self._peaks = peaks
        self._offsets = offsets                  # first row of every spectrum (n_spectra + 1), as handed to the matcher
        self._entry_windows = entry_windows      # windows of every index entry
        self._row_base = 0

    def __missing__(self, key):
        have = dict.keys(self)
        if key == "spec" and "spec_off" in have:
            spec_off = dict.__getitem__(self, "spec_off")
            value = np.repeat(np.arange(len(spec_off) - 1, dtype=np.int32), np.diff(spec_off))
        elif key == "win_off" and "entry" in have and self._entry_windows is not None:
            value = np.zeros(len(self["entry"]) + 1, dtype=np.int64)
            np.cumsum(self._entry_windows[self["entry"]], out=value[1:])
        elif key == "peak" and "peak32" in have:      # the device sent the rows as int32 counted from the batch's first row
            narrow = dict.__getitem__(self, "peak32")
            value = narrow.astype(np.int64)
            value[narrow >= 0] += self._row_base
        elif key == "peak" and "peak16" in have:      # ... or as uint16 counted from the first row of the hit's spectrum
            narrow = dict.__getitem__(self, "peak16")
            first_row = np.asarray(self._offsets, dtype=np.int64)[self["spec"]]
            value = np.repeat(first_row, np.diff(self["win_off"]))
            value += narrow
            value[narrow == 0xFFFF] = -1
        elif key in ("mmz", "mi") and self._peaks is not None:
            row = self["peak"]
            value = self._peaks[np.maximum(row, 0), 0 if key == "mmz" else 1]
            value[row < 0] = _NAN
        else:
            raise KeyError(key)
        self[key] = value
        return value

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default


class IsotopologueLibrary(dict):
    """Isotopologue library built and matched on one B200.

    Arguments as in the reference (IL:212-224): ``molecules``, ``charges``, ``metabolic_labels``,
    ``fixed_labels``, ``params``, ``trivial_names``, ``verbose``, ``evidences``, ``unimod_files``,
    ``add_default_files``.  Extra, keyword-only: ``device`` (CUDA ordinal, default ``LOCAL_RANK`` or 0)
    and ``distributed`` (shard the envelope build over the ranks of ``torch.distributed`` and all-gather
    the shards with NCCL; default: on when a process group with more than one rank is initialised).
    """

    def __init__(self, molecules=None, charges=None, metabolic_labels=None, fixed_labels=None, params=None,
                 trivial_names=None, verbose=True, evidences=None, unimod_files=None, add_default_files=True, *,
                 device=None, distributed=None, defer_matching=True, _plan_only=False, _native_planner=True):
        super().__init__()
        assert molecules is not None, "require list of molecules"
        assert charges is not None, "require list of charges"
        self.params = copy.deepcopy(_default_params)
        if params is not None:
            self.params.update(params)
        self._fmt = self.params["PERCENTILE_FORMAT_STRING"].format
        self.zero_labeled_percentile = self._fmt(0)
        if metabolic_labels is None or metabolic_labels == {}:
            metabolic_labels = {"15N": [0.0]}
        self.build_isotoplogues = True
        self.defer_matching = defer_matching
        self._native_planner = _native_planner
        self.verbose = verbose
        self.metabolic_labels = metabolic_labels
        self.fixed_labels = {} if fixed_labels is None else fixed_labels
        self.charges = charges
        self.lookup = {"molecule to formula": {}, "formula to molecule": {}, "molecule to annotations": {},
                       "molecule fixed label variations": {}, "formula to trivial name": {}}
        self.break_points = {}
        self.skipped_molecules = set()
        if trivial_names is not None:
            self.lookup["molecule to trivial name"] = trivial_names
        self.unimod_files = unimod_files
        self.lookup["formula to evidences"] = {} if evidences is None else evidences
        self.regex = {"<isotope><element>(<count>)": _ISOTOPE_ELEMENT_COUNT, "<isotope><element>": _ISOTOPE_ELEMENT}
        self.aa_compositions = {}
        self.isotopic_distributions = {}
        self.labled_percentiles = []
        self.metabolically_labeled_elements = set()
        # _plan_only: host tables only (CPU unit tests of the string/planning layer); nothing can be built or matched
        self._ctx = None if _plan_only else Context(device)   # raises without the CUDA library / a GPU: no fallback
        self._push_params()
        self._views = {}

        self._cache_kb()
        if self.verbose:
            print("> Metabolic labels        >", self.metabolic_labels)
            print("> Fixed labels            >", self.fixed_labels)
            print("> Charges                 >", self.charges)
            print("> Machine ppm offset      > {MACHINE_OFFSET_IN_PPM}".format(**self.params))
        self._extend_isotopic_distributions_with_metabolic_labels()
        self._build_label_percentile_tuples()
        if self.verbose:
            print("> Label percentile tuples >", self.labled_percentiles)
        if len(self.fixed_labels) != 0:
            self._extend_kb_with_fixed_labels()
            molecules = self._extend_molecules_with_fixed_labels(molecules)
            self._push_params()              # enrichment levels may have been added (IL:1469-1471)
        self._register_molecules(molecules, trivial_names, add_default_files)
        self._plan_device_tables()
        if _plan_only:
            return
        self._build_on_device(distributed)
        if self.verbose:
            if self._n_entries == 0:
                print("> No molecules have been added into the library ")
                print("  maybe molecules are out of mz limit range ?")
                print("  Current mz limits in params are: [ {LOWER_MZ_LIMIT} .. {UPPER_MZ_LIMIT}]".format(**self.params))
                raise SystemExit(1)          # IL:941-949
            print("> Created {0} match sets, total mz range [{1:10.5f} .. {2:10.5f}]".format(
                self._n_bins(), *self.match_set_mz_range))

    # ------------------------------------------------------------------ save / load
    _fixed_label_fast_path = True        # tests switch it off to compare with the general expansion
    _hit_pool = None                     # page-locked result arrays, created with the first match
    _hit_caps = (4096, 32768)            # (hits, hit windows) the next result arrays are sized for

    _hit_arena = None                    # distributed.SharedHitArena: result arrays in shared memory instead of the pool
    _compact_hits = True                 # deferred batches fetch the compact hit columns (PackedHits widens them on access)
    _ent_windows = None

    def use_hit_arena(self, arena):
        """Let the hits of batched matching land in ``arena`` (``distributed.SharedHitArena``; None = back to the pool)."""
        self._hit_arena = arena
        if arena is not None:
            arena.entry_windows = self._entry_windows()       # rank 0 widens the compact columns of every rank with it

    def _entry_windows(self):
        """Windows of every index entry = windows of each of its hits."""
        if self._ent_windows is None or len(self._ent_windows) != self._n_entries:
            self._ent_windows = np.diff(self._ent_win_off)
        return self._ent_windows

    # ---- lib[formula] records: made on first access (a placeholder per formula until then) -------------------------
    def __getitem__(self, formula):
        record = dict.__getitem__(self, formula)
        if record is None:
            record = _FormulaRecord(self, self._formula_index[formula], formula)
            dict.__setitem__(self, formula, record)
        return record

    def get(self, formula, default=None):
        return self[formula] if dict.__contains__(self, formula) else default

    def _all_records(self):
        for formula, record in list(dict.items(self)):
            if record is None:
                self[formula]
        return self

    def values(self):
        return dict.values(self._all_records())

    def items(self):
        return dict.items(self._all_records())

    def pop(self, formula, *default):
        if dict.__contains__(self, formula):
            self[formula]
        return dict.pop(self, formula, *default)

    def popitem(self):
        return dict.popitem(self._all_records())

    def setdefault(self, formula, default=None):
        return self[formula] if dict.__contains__(self, formula) else dict.setdefault(self, formula, default)

    def copy(self):
        return dict(self._all_records())

    def __eq__(self, other):
        return dict.__eq__(self._all_records(), other._all_records() if isinstance(other, IsotopologueLibrary) else other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    _DEVICE_STATE = ("_hit_pool", "_hit_arena", "_ctx", "_views", "_entry_static", "_ent_env", "_ent_z", "_ent_lower", "_ent_upper", "_ent_win_off",
                     "_ent_windows", "_win_lo", "_win_hi", "_win_cmz", "_win_ri", "_win_ci")

    def __reduce__(self):
        """Pickling keeps the host plan (knowledge base, lookups, isotope pools, envelope terms) and drops everything
        that lives on or was exported from the device; unpickling creates a context on the current device and runs
        the device build again (milliseconds), skipping all molecule-string work."""
        state = {k: v for k, v in self.__dict__.items() if k not in self._DEVICE_STATE}
        return (_restore_library, (state, list(dict.keys(self))))

    def save(self, path):
        """Write the library plan to ``path`` (pickle); ``IsotopologueLibrary.load`` rebuilds it on a GPU."""
        import pickle
        with open(path, "wb") as handle:
            pickle.dump(self, handle, protocol=pickle.HIGHEST_PROTOCOL)

    @staticmethod
    def load(path, device=None):
        """Library saved with ``save``; ``device`` as in the constructor (default: ``LOCAL_RANK`` or 0)."""
        import pickle
        token = _restore_device.set(device)
        try:
            with open(path, "rb") as handle:
                return pickle.load(handle)
        finally:
            _restore_device.reset(token)

    # ------------------------------------------------------------------ parameters -> device
    def _push_params(self):
        if self._ctx is None:
            return
        P = self.params
        q = QmsParams(
            element_min_abundance=P["ELEMENT_MIN_ABUNDANCE"], min_rel_peak_intensity=P["MIN_REL_PEAK_INTENSITY_FOR_MATCHING"],
            intensity_transformation=P["INTENSITY_TRANSFORMATION_FACTOR"], lower_mz_limit=P["LOWER_MZ_LIMIT"],
            upper_mz_limit=P["UPPER_MZ_LIMIT"], rel_mz_range=P["REL_MZ_RANGE"], rel_i_range=P["REL_I_RANGE"],
            internal_precision=P["INTERNAL_PRECISION"], machine_offset_ppm=P["MACHINE_OFFSET_IN_PPM"],
            m_score_threshold=P["M_SCORE_THRESHOLD"], required_peak_overlap=P["REQUIRED_PERCENTILE_PEAK_OVERLAP"],
            proton=knowledge_base.PROTON, min_matched_isotopologues=int(P["MINIMUM_NUMBER_OF_MATCHED_ISOTOPOLOGUES"]),
            max_molecules_per_bin=int(P["MAX_MOLECULES_PER_MATCH_BIN"]))
        self._ctx.call("qms_set_params", C.byref(q))

    # ------------------------------------------------------------------ knowledge base (IL:997-1026)
    def _cache_kb(self):
        for aa, formula in knowledge_base.aa_compositions.items():
            self.aa_compositions[aa] = ChemicalComposition(formula="+" + formula, unimod_file_list=self.unimod_files)
        for aa, formula in self.params.get("AMINO_ACIDS", {}).items():
            self.aa_compositions[aa] = ChemicalComposition(formula="+" + formula, unimod_file_list=self.unimod_files)
        for element, table in knowledge_base.isotopic_distributions.items():
            rows = sorted((mass, abundance, pos) for pos, (mass, abundance) in enumerate(table))
            self.isotopic_distributions[element] = {self.zero_labeled_percentile: rows}

    def _recalc_isotopic_distribution(self, element=None, target_percentile=None, enriched_isotope=None):
        """Enriched isotope table of ``element`` (IL:2109-2165), computed by ``qms_recalc_distribution``."""
        natural = self.isotopic_distributions[element][self.zero_labeled_percentile]
        if self._ctx is None:               # plan-only: the planner needs the isotope count, not the values
            return [(row[0], _NAN, row[2]) for row in natural]
        mass = np.array([row[0] for row in natural], dtype=np.float64)
        abun = np.array([row[1] for row in natural], dtype=np.float64)
        out = np.empty_like(abun)
        self._ctx.call("qms_recalc_distribution", len(natural), ptr(mass), ptr(abun), int(enriched_isotope),
                       float(target_percentile), ptr(out))
        return [(row[0], float(a), row[2]) for row, a in zip(natural, out)]

    def _extend_isotopic_distributions_with_metabolic_labels(self):
        """One isotope pool per (labelled element, percentile > 0) -- IL:1326-1373."""
        for isotope_element, percentiles in self.metabolic_labels.items():
            for percentile in percentiles:
                if percentile <= sys.float_info.epsilon:
                    continue
                found = _ISOTOPE_ELEMENT.match(isotope_element)
                element = found.group("element")
                self.isotopic_distributions[element][self._fmt(percentile)] = self._recalc_isotopic_distribution(
                    element=element, target_percentile=percentile,
                    enriched_isotope=int(round(float(found.group("isotope")))))
                self.metabolically_labeled_elements.add(element)
        if not self.metabolically_labeled_elements:
            self.metabolically_labeled_elements.add("N")

    def _build_label_percentile_tuples(self):
        """Cartesian product of the label percentile lists (IL:958-995)."""
        for combo in self._create_combinations([(k, len(v)) for k, v in self.metabolic_labels.items()]):
            chosen = {}
            for name, index in combo:
                chosen[_ISOTOPE_ELEMENT.match(name).group("element")] = self._fmt(self.metabolic_labels[name][index])
            self.labled_percentiles.append(tuple(sorted(chosen.items())))

    def _create_combinations(self, input_list, all_combos=None, current_combo=None, pos=0):
        """Position combinations of ``[(token, length), ...]`` (IL:1028-1076)."""
        return _cartesian(list(input_list))

    # ------------------------------------------------------------------ fixed labels (IL:1375-1677)
    def _extend_kb_with_fixed_labels(self, synthetic_abundance=0.994):
        levels = self.params["FIXED_LABEL_ISOTOPE_ENRICHMENT_LEVELS"]
        for amino_acid, definitions in self.fixed_labels.items():
            for index, unimod_string in enumerate(definitions):
                if amino_acid not in self.aa_compositions:
                    print("Error in _extend_kb_with_fixed_labels: unknown amino acid", amino_acid)
                    raise SystemExit(1)
                composition = copy.deepcopy(self.aa_compositions[amino_acid])
                for isotope, element, count in _ISOTOPE_ELEMENT_COUNT.findall(unimod_string):
                    if isotope == "":
                        key = element
                    else:
                        most_abundant = max(self.isotopic_distributions[element][self.zero_labeled_percentile],
                                            key=operator.itemgetter(1))
                        natural_name = str(round(most_abundant[0])).split(".")[0]
                        # a label naming the natural isotope (14N in carbamidomethyl) is "not enriched" (Q6)
                        percentile = 0.0 if isotope == natural_name else synthetic_abundance
                        key = isotope + element
                        if key in levels:
                            percentile = levels[key]
                        else:
                            if self.verbose:
                                print(">Enrichment levels of {0} was not specified in params."
                                      "Falling back to {1} and inserting temporarily into self.params".format(key, percentile))
                            levels[key] = percentile
                    composition[key] += int(count)
                    if isotope != "":
                        self.isotopic_distributions.setdefault(key, {})[self._fmt(percentile)] = \
                            self._recalc_isotopic_distribution(element=element, target_percentile=percentile,
                                                               enriched_isotope=isotope)
                for symbol in [s for s, n in composition.items() if n == 0]:
                    del composition[symbol]
                self.aa_compositions["{0}{1}".format(amino_acid, index)] = composition

    def _extend_molecules_with_fixed_labels(self, molecules):
        locked = self.params["SILAC_AAS_LOCKED_IN_EXPERIMENT"]
        if locked is not None:
            assert len(locked) != 1, 'only one aa in self.params["SILAC_AAS_LOCKED_IN_EXPERIMENT"]  ?'
            assert len({len(self.fixed_labels[aa]) for aa in locked}) == 1, \
                "Length of fixed_labels for aas that ought to be locked into same experiment dont show same number of labels ..."
        expanded = set()
        letters = "".join(re.escape(k) for k in self.aa_compositions if len(k) == 1)
        not_a_residue = re.compile("[^%s]" % letters) if letters else re.compile(".", re.S)
        labelled_residue = re.compile("|".join(re.escape(k) for k in self.fixed_labels)) if self.fixed_labels else None
        # Fast path for the usual case -- every labelled residue has exactly one label (carbamidomethyl cysteine) and
        # nothing is locked: the single combination of IL:1375-1677 puts state 0 behind every labelled residue, which is one
        # str.replace per molecule (0.3 us instead of 7 us; a 500 000-peptide library spent 2 s here).  Molecules that hold
        # anything besides residue letters (modifications, add-ons) take the general path below.
        simple = (self._fixed_label_fast_path and locked is None and labelled_residue is not None and all(len(k) == 1 and len(v) == 1 for k, v in self.fixed_labels.items()))
        if simple:
            residue_letters = "".join(k for k in self.aa_compositions if len(k) == 1)
            states = [(k, k + "0") for k in self.fixed_labels]
            molecule_list = molecules if isinstance(molecules, list) else list(molecules)
            labelled_list = self._label_all_at_once(molecule_list, residue_letters, states)
            if labelled_list is not None:                        # nothing but residue letters anywhere: no Python loop at all
                expanded.update(labelled_list)
                self._note_label_variations(molecule_list, labelled_list)
                return expanded
            variations = self.lookup["molecule fixed label variations"]
            general = []
            for full_molecule in molecule_list:
                if full_molecule.strip(residue_letters):         # something besides residue letters
                    general.append(full_molecule)
                    continue
                labelled = full_molecule
                for residue, with_state in states:
                    labelled = labelled.replace(residue, with_state)
                expanded.add(labelled)
                if len(labelled) != len(full_molecule):
                    variations.setdefault(full_molecule, set()).add(labelled)
            molecules = general
        for full_molecule in molecules:
            stop = not_a_residue.search(full_molecule)      # the sequence ends at the first non-residue character
            cut = stop.start() if stop else len(full_molecule)
            sequence, addon = full_molecule[:cut], full_molecule[cut:]
            if labelled_residue is None or labelled_residue.search(sequence) is None:
                expanded.add(full_molecule)
                continue
            sites = []
            for amino_acid in self.fixed_labels:
                for hit in re.finditer(amino_acid, sequence):
                    sites.append(("{0}{1}".format(hit.start(), amino_acid), len(self.fixed_labels[amino_acid])))
            for combination in self._create_combinations(sorted(sites, reverse=True)):
                exchange = sorted(((int(token[:-1]), token[-1], state) for token, state in combination), reverse=True)
                labelled = sequence
                states = set()
                for position, amino_acid, state in exchange:
                    labelled = "{0}{1}{2}{3}".format(labelled[:position], amino_acid, state, labelled[position + 1:])
                    if locked is not None and amino_acid in locked:
                        states.add(state)
                keep = True
                if locked is not None:
                    keep = False
                    if len(states) == 1:
                        present = {int(state) for aa, state in re.findall(r"([A-Z]{1})([0-9]+)", labelled) if aa in locked}
                        keep = present == states
                if keep:
                    expanded.add(labelled + addon)
                    self.lookup["molecule fixed label variations"].setdefault(full_molecule, set()).add(labelled + addon)
        return expanded

    @staticmethod
    def _label_all_at_once(molecule_list, residue_letters, states):
        """The single-label fast path over the whole list in three C-level passes: join, ``str.replace`` per labelled
        residue, split.  None when a molecule holds anything besides residue letters (or the separator)."""
        if not molecule_list or not residue_letters.isascii():
            return None
        joined = "\n".join(molecule_list)
        if not joined.isascii():
            return None
        rest = joined.encode("ascii").translate(None, residue_letters.encode("ascii"))
        if len(rest) != len(molecule_list) - 1 or rest.count(b"\n") != len(rest):
            return None                                          # other characters, or a separator inside a molecule
        for residue, with_state in states:
            joined = joined.replace(residue, with_state)
        labelled_list = joined.split("\n")
        return labelled_list if len(labelled_list) == len(molecule_list) else None

    def _note_label_variations(self, molecule_list, labelled_list):
        """``lookup["molecule fixed label variations"]`` of the fast path (molecule -> {labelled molecule}), filled when
        somebody reads it."""
        earlier = self.lookup["molecule fixed label variations"]

        def fill(target, earlier=earlier, molecule_list=molecule_list, labelled_list=labelled_list):
            dict.update(target, earlier)
            for molecule, labelled in zip(molecule_list, labelled_list):
                if len(labelled) != len(molecule):
                    dict.setdefault(target, molecule, set()).add(labelled)

        self.lookup["molecule fixed label variations"] = _LazyDict(fill)

    # ------------------------------------------------------------------ molecules -> formulas (IL:312-497)
    def _register_molecules(self, molecules, trivial_names, add_default_files):
        """Molecule strings -> formulas, lookups, per-formula count rows.  Plain formulas and unmodified peptides are
        parsed in bulk by the native planner; the rest goes through ``ChemicalComposition`` one by one."""
        self._highest_element_count = {}
        self._range_for_elements_with_two_isotopes = [None, None]
        factory = ChemicalComposition(aa_compositions=self.aa_compositions,
                                      isotopic_distributions=self.isotopic_distributions,
                                      unimod_file_list=self.unimod_files, add_default_files=add_default_files)
        self._add_default_files = add_default_files
        levels = self.params["FIXED_LABEL_ISOTOPE_ENRICHMENT_LEVELS"]
        known = self.isotopic_distributions
        # the reference iterates set(molecules) (hash order); sorted() makes the envelope order identical in
        # every process, which the sharded multi-GPU build relies on
        ordered = sorted(set(molecules))
        symbols = list(known)
        if self._native_planner:
            counts, seen, status, symbols = planner.parse_molecules(ordered, symbols, self.aa_compositions, compact=True)
        else:                                              # unit tests: every molecule through the host parser
            counts = np.zeros((len(ordered), len(symbols)), dtype=np.int32)
            seen = np.full((len(ordered), len(symbols)), planner.ABSENT, dtype=np.int16)
            status = np.ones(len(ordered), dtype=np.uint8)
        parsed_on_host = {}
        for i in np.flatnonzero(status).tolist():
            molecule = ordered[i]
            if molecule.count("#") > 1:
                raise ValueError(f"{molecule} contains too many '#' {molecule.count('#') + 1} only one allowed")
            composition = parsed_on_host[i] = factory.composition_of(molecule)
            for element in composition:
                if element in known:
                    continue
                # isotope-labelled element introduced by a modification (e.g. 13C of a SILAC/TMT unimod), IL:378-436
                found = _ISOTOPE_ELEMENT.match(element)
                try:
                    enriched = int(round(float(found.group("isotope"))))
                except Exception:
                    print("Failed on", element)
                    print("Maybe element is not in pyqms.knowledge_base.py ?")
                    raise SystemExit(1)
                template = found.group("element")
                target = levels.get("{0}{1}".format(enriched, template)) or 0.994
                if self.verbose:
                    print("> Extending isotopic distribution that are within the modifications for "
                          "element {0}{1}, enrichment levels set to {2}".format(enriched, template, target))
                known[element] = {self.zero_labeled_percentile: self._recalc_isotopic_distribution(
                    element=template, target_percentile=target, enriched_isotope=enriched)}
        # columns for symbols the host parser introduced and for the bare elements fixed-label isotopes fold into, in the
        # order of the knowledge base (the native parser may have come back with the few columns peptides need)
        needed = set(symbols) | {level_symbol[-1] for level_symbol in levels}
        for composition in parsed_on_host.values():
            needed.update(composition)
        wider = [symbol for symbol in known if symbol in needed]
        if wider != symbols:
            place = [wider.index(symbol) for symbol in symbols]
            wide_counts = np.zeros((len(ordered), len(wider)), dtype=np.int32)
            wide_seen = np.full((len(ordered), len(wider)), planner.ABSENT, dtype=np.int16)
            wide_counts[:, place] = counts
            wide_seen[:, place] = seen
            counts, seen, symbols = wide_counts, wide_seen, wider
        column = {symbol: k for k, symbol in enumerate(symbols)}
        for i, composition in parsed_on_host.items():
            for position, (element, count) in enumerate(composition.items()):
                counts[i, column[element]] = count
                seen[i, column[element]] = position
        # keep the columns some molecule uses (+ the fold targets): the knowledge base lists ~60 elements, a library uses ~6
        bare = {level_symbol[-1] for level_symbol in levels}
        keep = [k for k in range(len(symbols)) if symbols[k] in bare or bool((seen[:, k] >= 0).any())]
        symbols = [symbols[k] for k in keep]
        counts, seen = np.ascontiguousarray(counts[:, keep]), np.ascontiguousarray(seen[:, keep])
        formulas = planner.hill_formulas(counts, symbols)
        known_formulas = self.lookup["molecule to formula"]

        def fill_molecule_to_formula(target, earlier=known_formulas, ordered=ordered, formulas=formulas):
            dict.update(target, earlier)
            dict.update(target, zip(ordered, formulas))

        self.lookup["molecule to formula"] = _LazyDict(fill_molecule_to_formula)
        known_pairs = self.lookup["formula to molecule"]

        def fill_formula_to_molecule(target, earlier=known_pairs, ordered=ordered, formulas=formulas):
            dict.update(target, earlier)
            for molecule, formula in zip(ordered, formulas):
                dict.setdefault(target, formula, []).append(molecule)

        self.lookup["formula to molecule"] = _LazyDict(fill_formula_to_molecule)
        if trivial_names is not None:
            for molecule, formula in zip(ordered, formulas):
                name = trivial_names.get(molecule, None)
                if name is not None:
                    self.lookup["formula to trivial name"].setdefault(formula, []).append(name)
                else:
                    print("No trivial name for molecule", molecule)
        # one record per formula, taken from the first molecule (in sorted order) that has it.  The records themselves are made
        # when somebody asks for them (``lib[formula]``): the dict holds a placeholder per formula until then.
        first_molecule = dict(zip(reversed(formulas), range(len(formulas) - 1, -1, -1)))
        if dict.__len__(self):
            for formula in [f for f in first_molecule if dict.__contains__(self, f)]:
                del first_molecule[formula]
        rows = np.fromiter(first_molecule.values(), dtype=np.int64, count=len(first_molecule))
        rows.sort()
        self._symbols = symbols
        self._cc_counts = np.ascontiguousarray(counts[rows]) if len(rows) else np.zeros((0, len(symbols)), dtype=np.int32)
        self._cc_seen = np.ascontiguousarray(seen[rows]) if len(rows) else np.zeros((0, len(symbols)), dtype=np.int16)
        dict.update(self, dict.fromkeys([formulas[i] for i in rows.tolist()]))
        # highest atom count per element (insertion order = first molecule, then position inside it) and the count range
        # of the elements with two-isotope tables (IL:438-466)
        present = seen >= 0
        touched = np.flatnonzero(present.any(axis=0))
        first_row = present[:, touched].argmax(axis=0)
        appearance = sorted(zip(first_row.tolist(), seen[first_row, touched].tolist(), touched.tolist()))
        two_iso = self._range_for_elements_with_two_isotopes
        for _, _, k in appearance:
            element, held = symbols[k], counts[present[:, k], k]
            self._highest_element_count[element] = max(0, int(held.max()))
            if any(len(table) == 2 for table in known[element].values()):
                two_iso[0] = int(held.min()) if two_iso[0] is None else min(two_iso[0], int(held.min()))
                two_iso[1] = int(held.max()) if two_iso[1] is None else max(two_iso[1], int(held.max()))
        # fold the isotope-labelled counts into their template element (IL:467-497)
        for element in list(self._highest_element_count):
            if element.isalpha():
                continue
            template = _ISOTOPE_ELEMENT.match(element).group("element")
            self._highest_element_count.setdefault(template, 0)
            self._highest_element_count[template] += self._highest_element_count.get(element, 0)
            for table in self.isotopic_distributions[template].values():
                if len(table) == 2 and self._highest_element_count[template] > two_iso[1]:
                    two_iso[1] = self._highest_element_count[template]

    def _composition_of_formula(self, k):
        """``lib[formula]["cc"]``: the formula's elements in the order they entered the first molecule's composition."""
        composition = ChemicalComposition(aa_compositions=self.aa_compositions, isotopic_distributions=self.isotopic_distributions,
                                          unimod_file_list=self.unimod_files, add_default_files=self._add_default_files)
        row, seen = self._cc_counts[k], self._cc_seen[k]
        for column in np.argsort(np.where(seen >= 0, seen, 32767), kind="stable")[:int((seen >= 0).sum())].tolist():
            composition[self._symbols[column]] = int(row[column])
        return composition

    # ------------------------------------------------------------------ host tables for the device build
    def _plan_device_tables(self):
        """Pools (one per element x label) and per-envelope (pool, atom count) term lists (IL:546-646)."""
        self._pool_ids = {}             # (element, label) -> pool
        self._pools = []                # [(element, label)]
        iso_off, iso_mass, iso_abun, iso_q, max_level = [0], [], [], [], []
        self._default_label = {}
        for element, count in self._highest_element_count.items():
            assert element in self.isotopic_distributions, \
                "Isotopic distribution and abundance of {0} not in knowledge base".format(element)
            labels = self.isotopic_distributions[element]
            self._default_label[element] = sorted(labels, reverse=True)[0]      # first key of element_trees[element]
            for label, table in labels.items():
                self._pool_ids[(element, label)] = len(self._pools)
                self._pools.append((element, label))
                for _mass, _abundance, q in table:
                    # the reference reads the isotope of a step by its POSITION (IL:1733-1746); for mass-sorted
                    # knowledge-base tables that is the row itself
                    iso_mass.append(table[q][0] if len(table) > 2 else _mass)
                    iso_abun.append(table[q][1] if len(table) > 2 else _abundance)
                    iso_q.append(q)
                iso_off.append(len(iso_q))
                max_level.append(count)
        self._pool_tables = (np.array(iso_off, dtype=np.int32), np.array(iso_mass, dtype=np.float64),
                             np.array(iso_abun, dtype=np.float64), np.array(iso_q, dtype=np.int32),
                             np.array(max_level, dtype=np.int32))
        levels = self.params["FIXED_LABEL_ISOTOPE_ENRICHMENT_LEVELS"]
        symbols = self._symbols
        column = {symbol: k for k, symbol in enumerate(symbols)}
        n_lp = len(self.labled_percentiles)
        self._formulas = list(self.keys())
        self._formula_index = {f: i for i, f in enumerate(self._formulas)}
        self._lp_index = {lp: i for i, lp in enumerate(self.labled_percentiles)}
        pool_of_default = np.array([self._pool_ids.get((s, self._default_label.get(s)), 0) for s in symbols], dtype=np.int32)
        labelled = np.array([s in self.metabolically_labeled_elements for s in symbols], dtype=bool)
        fold_target, pool_of_label = [], []
        for lp in self.labled_percentiles:
            target = list(range(len(symbols)))
            for k, symbol in enumerate(symbols):
                if symbol in levels:
                    bare = symbol[-1]
                    for lp_element, lp_label in lp:
                        if lp_element == bare and lp_label == self._fmt(levels[symbol]) and bare in column:
                            target[k] = column[bare]   # fixed-label isotope folds into the labelled pool (Q6)
            fold_target.append(target)
            pool_of_label.append([(column.get(element), self._pool_ids.get((element, label), 0)) for element, label in lp])
        self._n_env = len(self._formulas) * n_lp
        self._terms = planner.envelope_terms(self._cc_counts, self._cc_seen, symbols, self.labled_percentiles, fold_target,
                                             labelled, pool_of_default, pool_of_label)

    # ------------------------------------------------------------------ device build
    def _build_on_device(self, distributed):
        import time
        self.break_points["Building isotopologues"] = time.time()
        ctx = self._ctx
        iso_off, iso_mass, iso_abun, iso_q, max_level = self._pool_tables
        two_iso_min = self._range_for_elements_with_two_isotopes[0] or 1
        ctx.call("qms_build_element_trees", len(self._pools), ptr(iso_off), ptr(iso_mass), ptr(iso_abun), ptr(iso_q),
                 ptr(max_level), int(two_iso_min))
        term_off, term_pool, term_level = self._terms
        ctx.call("qms_set_envelope_terms", self._n_env, ptr(term_off), ptr(term_pool), ptr(term_level))
        from . import distributed as dist_mod
        world = dist_mod.world_size() if distributed is not False else 1
        if distributed is None:
            distributed = dist_mod.shard_build_by_default(self._n_env, world)      # small libraries are faster replicated
        self._build_stats = {"sharded": bool(distributed and world > 1), "world": world, "envelopes": int(self._n_env)}
        if distributed and world > 1:
            begin, end = dist_mod.shard_range(self._n_env, dist_mod.rank(), world)
            ctx.call("qms_build_envelopes", begin, end)
            t0 = time.perf_counter()
            self._build_stats.update(dist_mod.allgather_envelopes(ctx, self._n_env, world))
            self._build_stats["allgather_wall_s"] = time.perf_counter() - t0
        else:
            ctx.call("qms_build_envelopes", 0, self._n_env)
        # ranks of (label tuple, formula) under Python ordering: tie-break of the index sort (IL:836-867)
        n_lp = len(self.labled_percentiles)
        n_formulas = len(self._formulas)
        formula_rank = np.empty(n_formulas, dtype=np.int64)
        formula_rank[sorted(range(n_formulas), key=self._formulas.__getitem__)] = np.arange(n_formulas, dtype=np.int64)
        lp_rank = np.empty(n_lp, dtype=np.int64)
        lp_rank[sorted(range(n_lp), key=self.labled_percentiles.__getitem__)] = np.arange(n_lp, dtype=np.int64)
        # envelope e = formula * n_lp + label tuple; both key parts are unique, so the pair order is lexicographic
        rank = (lp_rank[None, :] * n_formulas + formula_rank[:, None]).reshape(-1)
        charges = np.array(self.charges, dtype=np.int32)
        ctx.call("qms_finalize_index", len(charges), ptr(charges), ptr(rank))
        info = QmsIndexInfo()
        ctx.call("qms_index_info_get", C.byref(info))
        self._n_entries, self._n_windows = info.n_entries, info.n_windows
        self.match_set_mz_range = [info.mz_min, info.mz_max] if info.n_entries else [None, None]
        # entry table + entry-major windows: needed on the host to label every hit
        n, w = self._n_entries, self._n_windows
        self._ent_env = np.empty(n, dtype=np.int64)
        self._ent_z = np.empty(n, dtype=np.int32)
        self._ent_lower = np.empty(n, dtype=np.float64)
        self._ent_upper = np.empty(n, dtype=np.float64)
        self._ent_win_off = np.zeros(n + 1, dtype=np.int64)
        ctx.call("qms_export_index", ptr(self._ent_env), ptr(self._ent_z), ptr(self._ent_lower), ptr(self._ent_upper),
                 ptr(self._ent_win_off))
        for name in self._WINDOW_VIEWS:                  # exported when somebody reads them (__getattr__): 28 bytes per window
            self.__dict__.pop(name, None)
        self._entry_static = {}
        self._ent_windows = None
        if self.verbose:
            print("\n> Execution time {0:.2f} seconds".format(time.time() - self.break_points["Building isotopologues"]))

    _WINDOW_VIEWS = ("_win_lo", "_win_hi", "_win_cmz", "_win_ri", "_win_ci")

    def __getattr__(self, name):
        """The entry-major window arrays of the index (grid bounds, centre m/z, relative intensity, isotopologue ordinal)
        come to the host on first use: labelling hits needs them, building and matching do not."""
        if name in IsotopologueLibrary._WINDOW_VIEWS and "_ctx" in self.__dict__ and "_n_windows" in self.__dict__:
            w = self._n_windows
            views = {"_win_lo": np.empty(w, dtype=np.int32), "_win_hi": np.empty(w, dtype=np.int32), "_win_cmz": np.empty(w, dtype=np.float64),
                     "_win_ri": np.empty(w, dtype=np.float64), "_win_ci": np.empty(w, dtype=np.int32)}
            self._ctx.call("qms_export_windows", *[ptr(views[k]) for k in IsotopologueLibrary._WINDOW_VIEWS])
            self.__dict__.update(views)
            return views[name]
        raise AttributeError(name)

    # ------------------------------------------------------------------ lazy host views
    def _n_bins(self):
        step = self.params["MAX_MOLECULES_PER_MATCH_BIN"]
        return (self._n_entries + step - 1) // step

    def _export_envelopes(self):
        if "env" not in self._views:
            ctx = self._ctx
            n_env, n_slots = C.c_int64(), C.c_int64()
            ctx.call("qms_envelope_layout", C.byref(n_env), C.byref(n_slots), None)
            slot_off = np.zeros(n_env.value + 1, dtype=np.int64)
            ctx.call("qms_envelope_layout", C.byref(n_env), C.byref(n_slots), ptr(slot_off))
            s = n_slots.value
            view = {"slot_off": slot_off, "mass": np.empty(s), "relabun": np.empty(s), "abun": np.empty(s, dtype=np.int32),
                    "pos": np.empty(s, dtype=np.int32), "c_flag": np.empty(s, dtype=np.uint8),
                    "len": np.empty(n_env.value, dtype=np.int32), "n_c": np.empty(n_env.value, dtype=np.int32), "planes": {}}
            ctx.call("qms_export_envelopes", ptr(view["mass"]), ptr(view["relabun"]), ptr(view["abun"]), ptr(view["pos"]),
                     ptr(view["c_flag"]), ptr(view["len"]), ptr(view["n_c"]))
            self._views["env"] = view
        return self._views["env"]

    def _charge_plane(self, z_index):
        view = self._export_envelopes()
        if z_index not in view["planes"]:
            s = len(view["mass"])
            mz, lo, hi = np.empty(s), np.empty(s, dtype=np.int32), np.empty(s, dtype=np.int32)
            self._ctx.call("qms_export_charge_plane", z_index, ptr(mz), ptr(lo), ptr(hi))
            view["planes"][z_index] = (mz, lo, hi)
        return view["planes"][z_index]

    def _materialise_envelope(self, formula, lp):
        """Reference layout of ``lib[formula]["env"][lp]`` (IL:703-811) from the exported device arrays."""
        view = self._export_envelopes()
        e = self._formula_index[formula] * len(self.labled_percentiles) + self._lp_index[lp]
        a, n = int(view["slot_off"][e]), int(view["len"][e])
        sl = slice(a, a + n)
        c_flag = view["c_flag"][sl].astype(bool)
        env = {"isot": [], "mass": view["mass"][sl].tolist(), "abun": view["abun"][sl].tolist(),
               "relabun": view["relabun"][sl].tolist(),
               "c_peak_pos": [int(p) if c else None for p, c in zip(view["pos"][sl], c_flag)],
               "n_c_peaks": float(view["n_c"][e]) if view["n_c"][e] else 0}
        for z_index, charge in enumerate(self.charges):
            mz, lo, hi = self._charge_plane(z_index)
            tmzs = [set(range(int(l), int(h) + 1)) if c else None for l, h, c in zip(lo[sl], hi[sl], c_flag)]
            env[charge] = {"mz": mz[sl].tolist(), "tmzs": tmzs, "atmzs": set().union(*[t for t in tmzs if t is not None])}
        return env

    def _entry_tuple(self, index):
        n_lp = len(self.labled_percentiles)
        e = int(self._ent_env[index])
        return (float(self._ent_lower[index]), float(self._ent_upper[index]), self.charges[int(self._ent_z[index])],
                self.labled_percentiles[e % n_lp], self._formulas[e // n_lp])

    @property
    def formulas_sorted_by_mz(self):
        """[(lower_mz, upper_mz, charge, label tuple, formula)] in index order (IL:836-867)."""
        if "index" not in self._views:
            self._views["index"] = [self._entry_tuple(i) for i in range(self._n_entries)]
        return self._views["index"]

    @property
    def match_sets(self):
        """Bins of MAX_MOLECULES_PER_MATCH_BIN entries with their window union and m/z range (IL:869-916)."""
        if "bins" not in self._views:
            step = self.params["MAX_MOLECULES_PER_MATCH_BIN"]
            bins = {}
            for number, start in enumerate(range(0, self._n_entries, step)):
                stop = min(start + step, self._n_entries)
                union = set()
                for w in range(int(self._ent_win_off[start]), int(self._ent_win_off[stop])):
                    union.update(range(int(self._win_lo[w]), int(self._win_hi[w]) + 1))
                bins[number] = {"ids": [start, stop], "tmzs": union,
                                "mz_range": [float(self._ent_lower[start:stop].min()), float(self._ent_upper[start:stop].max())]}
            self._views["bins"] = bins
        return self._views["bins"]

    @property
    def element_trees(self):
        """{element: {label: {n: {"env": {pos: {"mass","abun"}}, "minPos", "maxPos"}}}} (IL:1086-1106); only the
        pruned slice [minPos, maxPos] of each level is held on the device and materialised here."""
        if "trees" not in self._views:
            trees = {}
            cap = 1 << 16
            abun, mass = np.empty(cap), np.empty(cap)
            lo, hi, n = C.c_int32(), C.c_int32(), C.c_int32()
            for (element, label), pool in self._pool_ids.items():
                levels = trees.setdefault(element, {}).setdefault(label, {})
                for level in range(1, int(self._pool_tables[4][pool]) + 1):
                    self._ctx.call("qms_export_tree_level", pool, level, C.byref(lo), C.byref(hi), ptr(abun), ptr(mass), cap,
                                   C.byref(n))
                    if n.value == 0:
                        continue
                    levels[level] = {"minPos": lo.value, "maxPos": hi.value,
                                     "env": {lo.value + j: {"mass": float(mass[j]), "abun": float(abun[j])}
                                             for j in range(n.value)}}
            self._views["trees"] = trees
        return self._views["trees"]

    # ------------------------------------------------------------------ matching (P2)
    def _new_results(self):
        return Results(aa_compositions=self.aa_compositions, charges=self.charges, fixed_labels=self.fixed_labels,
                       isotopic_distributions=self.isotopic_distributions, lookup=self.lookup,
                       metabolic_labels=self.metabolic_labels, params=self.params)

    def _run_matcher(self, peaks, offsets, input_kind, xi, only_entry=None, values=True):
        """One ``qms_match_batch`` / ``qms_match_entry`` pass with host buffers.

        The hits land in page-locked arrays of a reusable pool (``_pinned.PinnedPool``): the library copies every
        sub-batch of a large run out while the next one is being matched, and a block returns to the pool when the
        last array viewing it is collected.  The pool arrays are sized from the previous call; a call that needs more
        gets ``QMS_E_CAPACITY`` with the sizes and fetches the hits the device has kept (``qms_fetch_hits``; no second
        matching pass).  ``values=False`` leaves the matched (m/z, intensity) values on the host side: ``PackedHits``
        gathers them from ``peaks`` by ``peak`` row when somebody asks."""
        ctx = self._ctx
        n_spec = len(offsets) - 1
        nh, nw = C.c_int64(), C.c_int64()
        pool = self._hit_pool
        if pool is None:
            pool = self._hit_pool = PinnedPool()
        cap_h, cap_w = self._hit_caps

        # without the values the index columns travel in their compact forms (qms_hits: hit_peak16, spec_hit_off; the
        # window offsets follow from the entries): 20 bytes + 2 per window instead of 32 + 4
        compact = (not values and only_entry is None and self._compact_hits and n_spec > 0
                   and int(np.max(np.diff(offsets))) <= 65534)

        def buffers(h, w):
            out = PackedHits(peaks if not values else None, offsets, self._entry_windows())
            if self._hit_arena is not None and only_entry is None:
                out.update(self._hit_arena.arrays(h, w, "compact" if compact else values, n_spec))    # shared segment of this rank
            elif compact:
                out.update({"entry": pool.array(np.int32, h), "score": pool.array(np.float64, h), "scaling": pool.array(np.float64, h),
                            "peak16": pool.array(np.uint16, w), "spec_off": pool.array(np.int64, n_spec + 1)})
            else:
                out.update({"spec": pool.array(np.int32, h), "entry": pool.array(np.int32, h), "score": pool.array(np.float64, h),
                            "scaling": pool.array(np.float64, h), "win_off": pool.array(np.int64, h + 1)})
                if values:
                    out["mmz"], out["mi"], out["peak"] = pool.array(np.float64, w), pool.array(np.float64, w), pool.array(np.int64, w)
                else:
                    out["peak32"] = pool.array(np.int32, w)       # rows counted from offsets[0]: half the bytes across the bus
            out._row_base = int(offsets[0])
            have = dict.keys(out)
            hits = QmsHits(h, w, *[out[k].ctypes.data if k in have else None
                                   for k in ("spec", "entry", "score", "scaling", "win_off", "mmz", "mi", "peak", "peak32", "nc8", "peak16",
                                             "spec_off")])
            return out, hits

        out, hits = buffers(cap_h, cap_w)
        if only_entry is None:
            rc = ctx.lib.qms_match_batch(ctx.handle, ptr(peaks), ptr(offsets), n_spec, input_kind, float(xi), C.byref(hits),
                                         C.byref(nh), C.byref(nw))
        else:
            rc = ctx.lib.qms_match_entry(ctx.handle, ptr(peaks), len(peaks), input_kind, int(only_entry), float(xi), C.byref(hits),
                                         C.byref(nh), C.byref(nw))
        if rc == QMS_E_CAPACITY:
            del out, hits
            cap_h, cap_w = nh.value + nh.value // 8 + 64, nw.value + nw.value // 8 + 512
            out, hits = buffers(cap_h, cap_w)
            rc = ctx.lib.qms_fetch_hits(ctx.handle, C.byref(hits), C.byref(nh), C.byref(nw))
        ctx.check(rc)
        h, w = nh.value, nw.value
        if only_entry is None:
            # the next call of a run needs about as much: keep the capacities (and with them the pool's size classes)
            # stable, grow them only when a run comes close to them
            self._hit_caps = (cap_h if h < 0.95 * cap_h else h + h // 8 + 64, cap_w if w < 0.95 * cap_w else w + w // 8 + 512)
        have = set(dict.keys(out))
        for k in have & {"spec", "entry", "score", "scaling"}:
            out[k] = out[k][:h]
        if "win_off" in have:
            out["win_off"] = out["win_off"][:h + 1]
            if h == 0:
                out["win_off"][0] = 0
        if "spec_off" in have:
            out["spec_off"] = out["spec_off"][:n_spec + 1]
            if h == 0:
                out["spec_off"][:] = 0
        for k in have & {"peak", "peak32", "peak16", "mmz", "mi"}:
            out[k] = out[k][:w]
        out["n_hits"], out["n_windows"] = h, w
        return out

    @staticmethod
    def _as_peak_array(mz_i_list):
        """(n,2) float64 C-contiguous array + input kind (0 numpy semantics, 1 list semantics; Q12)."""
        is_numpy = getattr(mz_i_list, "tolist", False) is not False
        array = np.ascontiguousarray(mz_i_list, dtype=np.float64)
        if array.size == 0:
            array = array.reshape(0, 2)
        return array, 0 if is_numpy else 1

    def _sorted_or_reordered(self, array, kind):
        """The device wants ascending m/z buckets; an unsorted numpy spectrum is reordered by (bucket, row),
        which keeps 'first row of a bucket' (IL:2592-2600) what it was.  Returns (array, row map or None)."""
        mz = array[:, 0]
        if len(mz) < 2 or bool(np.all(mz[1:] >= mz[:-1])):
            return array, None
        if kind == 1:
            raise ValueError("list input must be sorted by m/z (the reference bisects it, IL:2511)")
        order = np.argsort(np.round(mz * self.params["INTERNAL_PRECISION"]), kind="stable")
        return np.ascontiguousarray(array[order]), order

    def _emit(self, out, results, file_names, spec_ids, spec_rts, python_floats):
        """Packed hits -> ``results.add`` in (spectrum, entry) order (IL:1872-1882).  The matched (m/z,
        intensity) come back from the device: numpy scalars for numpy input, Python floats for list input,
        like the objects the reference hands back (IL:2014-2017, 2594-2600)."""
        n_hits = out["n_hits"]
        if n_hits == 0:
            return results
        # whole-array conversions once per batch; the per-hit loop then only indexes Python lists
        spec, entry, win_off = out["spec"].tolist(), out["entry"].tolist(), out["win_off"].tolist()
        if python_floats:
            mmz, mi, score, scaling = out["mmz"].tolist(), out["mi"].tolist(), out["score"].tolist(), out["scaling"].tolist()
        else:
            mmz, mi, score, scaling = list(out["mmz"]), list(out["mi"]), list(out["score"]), list(out["scaling"])
        statics = self._entry_static
        one_name = not isinstance(file_names, list)
        append = getattr(results, "_append_match", None)     # pyqms_b200.Results: skips per-hit key / index-set work
        if append is not None:
            results._flush()                                  # older deferred batches come first
            new_key, new_match, keys = results._m_key_class, results._match_class, {}
        for h in range(n_hits):
            s, x = spec[h], entry[h]
            static = statics.get(x)
            if static is None:
                lo, hi, charge, lp, formula = self._entry_tuple(x)
                a, b = int(self._ent_win_off[x]), int(self._ent_win_off[x + 1])
                static = statics[x] = (formula, charge, lp, self._win_ri[a:b].tolist(), self._win_cmz[a:b].tolist(),
                                       self._win_ci[a:b].tolist())
            formula, charge, lp, ri, cmz, ci = static
            a = win_off[h]
            peaks = tuple([(None, None, ri[m], cmz[m], ci[m]) if mmz[a + m] != mmz[a + m] else (mmz[a + m], mi[a + m], ri[m], cmz[m], ci[m])
                           for m in range(win_off[h + 1] - a)])
            name = file_names if one_name else file_names[s]
            if append is None:
                results.add((name, formula, charge, lp), (spec_ids[s], spec_rts[s], score[h], scaling[h], peaks))
                continue
            key = keys.get((name, x))
            if key is None:
                key = keys[(name, x)] = new_key(name, formula, charge, lp)
            append(key, new_match(spec_ids[s], spec_rts[s], score[h], scaling[h], peaks))
        return results

    # spectra handed to match_all are queued on the Results object and matched in one device pass when the
    # results are first read (or when the queue is full): a per-spectrum loop then runs at batch speed
    QUEUE_MAX_SPECTRA = 4096
    QUEUE_MAX_PEAKS = 8 << 20

    def _match_queue(self, queue, results):
        """Device pass over the spectra queued by match_all (called by ``Results._flush``), oldest first."""
        start = 0
        while start < len(queue):
            kind = queue[start][1]
            stop = start
            while stop < len(queue) and queue[stop][1] == kind:
                stop += 1
            run = queue[start:stop]
            arrays = [item[0] for item in run]
            offsets = np.zeros(len(run) + 1, dtype=np.int64)
            np.cumsum([len(a) for a in arrays], out=offsets[1:])
            peaks = np.concatenate(arrays) if len(arrays) > 1 else arrays[0]
            out = self._run_matcher(peaks, offsets, kind, results.params["MZ_SCORE_PERCENTILE"])
            self._emit(out, results, [item[2] for item in run], [item[3] for item in run], [item[4] for item in run], kind == 1)
            start = stop

    def match_all(self, mz_i_list=None, file_name=None, spec_id=None, spec_rt=None, results=None):
        """Match every library entry against one spectrum (IL:1794-1883); returns/updates ``results``.

        With ``defer_matching`` (default) and a pyqms_b200 ``Results`` the spectrum is copied into the result
        object's queue and matched with the next device pass (first read of ``results``, a full queue, or
        ``results.flush()``); hits appear in exactly the order of an eager loop."""
        if results is None:
            results = self._new_results()
        if self._n_entries == 0:
            return results
        array, kind = self._as_peak_array(mz_i_list)
        array, _ = self._sorted_or_reordered(array, kind)
        if self.defer_matching and hasattr(results, "enqueue"):
            if array is mz_i_list or array.base is not None or not array.flags["OWNDATA"]:
                array = array.copy()                       # the caller may reuse its buffer
            results.enqueue(self, (array, kind, file_name, spec_id, spec_rt), len(array),
                            self.QUEUE_MAX_SPECTRA, self.QUEUE_MAX_PEAKS)
            return results
        offsets = np.array([0, len(array)], dtype=np.int64)
        out = self._run_matcher(array, offsets, kind, results.params["MZ_SCORE_PERCENTILE"])
        return self._emit(out, results, file_name, [spec_id], [spec_rt], kind == 1)

    def match_many(self, spectra, file_name=None, spec_ids=None, spec_rts=None, results=None, offsets=None, lazy=True):
        """Batched ``match_all``: one device pass over many spectra (additive API).

        ``spectra``: list of (n,2) arrays / lists of (mz, i) tuples, or one packed (N,2) array together with
        ``offsets`` (n_spectra+1 row offsets).  Hits are added to ``results`` in the order a ``match_all`` loop
        over the spectra would add them; with ``lazy`` (default) the ``match`` tuples are only built when
        ``results`` is first read (``Results.defer``), the packed arrays are available at once through
        ``results.packed_batches()``."""
        if results is None:
            results = self._new_results()
        if offsets is None:
            arrays, kinds = [], set()
            for spectrum in spectra:
                array, kind = self._as_peak_array(spectrum)
                array, _ = self._sorted_or_reordered(array, kind)
                kinds.add(kind)
                arrays.append(array)
            if len(kinds) > 1:
                raise ValueError("match_many needs all spectra as numpy arrays or all as lists")
            kind = kinds.pop() if kinds else 0
            peaks = np.concatenate(arrays) if arrays else np.zeros((0, 2))
            offsets = np.zeros(len(arrays) + 1, dtype=np.int64)
            np.cumsum([len(a) for a in arrays], out=offsets[1:])
        else:
            peaks = np.ascontiguousarray(spectra, dtype=np.float64)
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
            kind = 0
        n_spec = len(offsets) - 1
        spec_ids = list(range(n_spec)) if spec_ids is None else spec_ids
        spec_rts = [None] * n_spec if spec_rts is None else spec_rts
        if self._n_entries == 0 or n_spec == 0:
            return results
        if hasattr(results, "flush_queue"):
            results.flush_queue()                          # spectra queued by earlier match_all calls come first
        deferred = lazy and hasattr(results, "defer")
        # a lazy result leaves the matched (m/z, intensity) values where they are, in ``peaks``: they are gathered by
        # peak row if somebody reads them (``peaks`` is kept alive by the result; do not overwrite it meanwhile)
        out = self._run_matcher(peaks, offsets, kind, results.params["MZ_SCORE_PERCENTILE"], values=not deferred)
        if deferred:
            results.defer(out, lambda target: self._emit(out, target, file_name, spec_ids, spec_rts, kind == 1),
                          table=lambda: self._hit_columns(out, file_name, spec_ids, spec_rts))
            return results
        return self._emit(out, results, file_name, spec_ids, spec_rts, kind == 1)

    def _hit_columns(self, out, file_names, spec_ids, spec_rts):
        """Columns of ``Results.format_all_results`` (all but ``peaks``) of one packed batch, by array indexing only."""
        entry, spec = out["entry"], out["spec"]
        n_lp = len(self.labled_percentiles)
        env = self._ent_env[entry]
        formulas = np.array(self._formulas, dtype=object)
        percentiles = np.empty(n_lp, dtype=object)
        percentiles[:] = list(self.labled_percentiles)
        as_object = lambda values: np.array(list(values) + [None], dtype=object)[:-1]   # keeps tuples as cells
        names = as_object(file_names)[spec] if isinstance(file_names, list) else np.full(len(spec), file_names, dtype=object)
        matched = np.zeros(len(spec), dtype=np.int64)
        if out["n_windows"]:
            hit_of_window = np.repeat(np.arange(len(spec)), np.diff(out["win_off"]))
            np.add.at(matched, hit_of_window, out["peak"] >= 0)
        return {"file_name": names, "formula": formulas[env // n_lp], "charge": np.asarray(self.charges)[self._ent_z[entry]],
                "label_percentiles": percentiles[env % n_lp], "spec_id": as_object(spec_ids)[spec],
                "rt": as_object(spec_rts)[spec], "score": out["score"], "scaling_factor": out["scaling"],
                "n_peaks": np.diff(out["win_off"]), "n_matched_peaks": matched}

    def match_packed(self, peaks, offsets, mz_score_percentile=None, input_kind=0):
        """Device pass without Python object creation: returns the packed hit arrays of ``qms_match_batch``
        (dict with spec, entry, score, scaling, win_off, mmz, mi, peak, n_hits, n_windows)."""
        xi = self.params["MZ_SCORE_PERCENTILE"] if mz_score_percentile is None else mz_score_percentile
        return self._run_matcher(np.ascontiguousarray(peaks, dtype=np.float64), np.ascontiguousarray(offsets, dtype=np.int64),
                                 input_kind, xi)

    def match_isotopologue(self, index=None, formula=None, charge=None, label_percentile=None, spec_tmz_set=None,
                           spec_tmz_lookup=None, mz_i_list=None, mz_score_percentile=None):
        """Best match of one library entry (IL:1885-2040): (score, scaling_factor, peaks) or None."""
        if index is None:
            assert formula is not None, "require formula information for match"
            assert charge is not None, "require charge information for match"
            assert label_percentile is not None, "require tuple list"
            e = self._formula_index[formula] * len(self.labled_percentiles) + self._lp_index[label_percentile]
            z_index = list(self.charges).index(charge)
            found = np.flatnonzero((self._ent_env == e) & (self._ent_z == z_index))
            if len(found) == 0:
                raise KeyError("entry is not in the m/z range of the library index")
            index = int(found[0])
        if mz_i_list is not None:
            assert spec_tmz_set is None, "both a transformed spec and mz_i_list were given ... which one to use ?"
            source = mz_i_list
        else:
            source = [spec_tmz_lookup[t][0] for t in sorted(spec_tmz_set)]   # first peak of every bucket (IL:2014)
        array, kind = self._as_peak_array(source)
        if mz_i_list is None:
            kind = 0
        array, row_map = self._sorted_or_reordered(array, kind)
        xi = self.params["MZ_SCORE_PERCENTILE"] if mz_score_percentile is None else mz_score_percentile
        out = self._run_matcher(array, np.array([0, len(array)], dtype=np.int64), kind, xi, only_entry=index)
        if out["n_hits"] == 0:
            return None
        collected = self._emit(out, _Collector(), None, [None], [None], kind == 1 or mz_i_list is None)
        return collected.rows[0]

    def calc_amounts(self, results, rt_windows=None):
        """Amounts of every matched isotope chromatogram in ``results``.

        Per key the matches inside the retention-time window ``rt_windows[key] = (start_min, stop_min)`` (whole
        profile if absent) give ``{"max I in window", "max I in window (rt)", "max I in window (score)",
        "sum I in window", "auc in window"}`` exactly like ``pyqms.adaptors.calc_amount_function``
        (adaptors.py:509-569) driven by ``Results.calc_amounts_from_rt_info_file`` (results.py:1238-1262);
        ``None`` for a key without a match in its window.  Computed by ``qms_calc_amounts`` (one thread per key)."""
        keys = list(results.keys())
        windows = None
        if rt_windows is not None:
            windows = [rt_windows.get(key, (_NAN, _NAN)) for key in keys]
        return dict(zip(keys, calc_amounts_of_profiles([results[key]["data"] for key in keys], windows, ctx=self._ctx)))

    def score_matches(self, matched_peaks, mz_score_percentile):
        """mScore and scaling factor of one list of (mmz, mi, ri, cmz, ci) tuples (IL:2441-2498), on the device."""
        rows = list(matched_peaks)
        cols = [np.array([_NAN if r[j] is None else float(r[j]) for r in rows], dtype=np.float64) for j in range(5)]
        score, scaling = C.c_double(), C.c_double()
        self._ctx.call("qms_score_matches", len(rows), ptr(cols[0]), ptr(cols[1]), ptr(cols[2]), ptr(cols[3]), ptr(cols[4]),
                       float(mz_score_percentile), C.byref(score), C.byref(scaling))
        return score.value, scaling.value

    # ------------------------------------------------------------------ reference helpers kept for API parity
    def _transform_values(self, mz_values):
        mz = np.ascontiguousarray(mz_values, dtype=np.float64)
        lo, hi, tmz = (np.empty(len(mz), dtype=np.int32) for _ in range(3))
        self._ctx.call("qms_transform_mz", len(mz), ptr(mz), ptr(lo), ptr(hi), ptr(tmz))
        return lo, hi, tmz

    def _transform_mz_to_set(self, mz):
        """Integer ppm window of one m/z as a set (IL:2533-2552); edges computed by ``qms_transform_mz``."""
        lo, hi, _ = self._transform_values([mz])
        return set(range(int(lo[0]), int(hi[0]) + 1))

    def _slice_list(self, source_list, borders, tolerance=2):
        """+-2 elements (list) / +-2 Da (numpy) slice around ``borders`` (IL:2500-2531)."""
        import bisect
        lower, upper = list(borders[0]), list(borders[1])
        if getattr(source_list, "tolist", False) is False:
            try:
                start, stop = bisect.bisect(source_list, lower) - tolerance, bisect.bisect(source_list, upper) + tolerance
            except TypeError:
                start = bisect.bisect(source_list, tuple(lower)) - tolerance
                stop = bisect.bisect(source_list, tuple(upper)) + tolerance
            return source_list[max(start, 0):min(stop, len(source_list))]
        keep = (lower[0] - tolerance < source_list[:, 0]) * (source_list[:, 0] < upper[0] + tolerance)
        return source_list[np.where(keep)]

    def _transform_spectrum(self, mz_i_list, mz_range=None):
        """(set of integer m/z buckets, bucket -> [(mz, i), ...]) of a spectrum (IL:2554-2603)."""
        target = mz_i_list if mz_range is None else self._slice_list(mz_i_list, [(x, 0.001) for x in mz_range])
        is_numpy = getattr(target, "tolist", False) is not False
        mz = target[:, 0] if is_numpy else [p[0] for p in target]
        _, _, tmz = self._transform_values(mz)
        lookup = {}
        for row, bucket in enumerate(tmz if is_numpy else tmz.tolist()):
            lookup.setdefault(bucket, []).append((target[row][0], target[row][1]))
        return set(tmz if is_numpy else tmz.tolist()), lookup

    def print_overview(self, formula, charge=None):
        """Isotope pattern table of one formula/molecule (IL:2042-2103)."""
        if formula not in self:
            formula = self.lookup["molecule to formula"][formula]
        if charge is None:
            charge = self.charges[0]
        print("\n\n{0: ^30}\n\n".format("Overview"))
        print("> Chemical formula", formula)
        names = self.lookup["formula to trivial name"].get(formula, None)
        if names is not None:
            print("> Trivial name{0} {1}".format("" if len(names) == 1 else "s", names))
        for lp in sorted(self[formula]["env"]):
            env = self[formula]["env"][lp]
            print("> Label percentile", lp)
            print("> Isotope pattern                                                Abundance\n pos      Mass\t\t\t"
                  "m/z [MH]{0:+1}               transformed    rel.".format(charge))
            for n in range(len(env["abun"])):
                print("{0: >3}\t{1:16.10f}\t{2:16.10f}\t{3:10.0f}\t{4:12.11f}\t{5}".format(
                    n, env["mass"][n], env[charge]["mz"][n], env["abun"][n], env["relabun"][n], env["c_peak_pos"][n]))

    # ------------------------------------------------------------------ introspection for benchmarks/tests
    def device_counters(self):
        return self._ctx.counters()

    def entry_peaks(self):
        """Per index entry: (theoretical c-peak m/z array, integer abundance array) -- workload generators."""
        off = self._ent_win_off
        return ([self._win_cmz[off[i]:off[i + 1]] for i in range(self._n_entries)],
                [self._win_ci[off[i]:off[i + 1]].astype(np.float64) for i in range(self._n_entries)])


class _Collector:
    """Minimal ``results`` stand-in used by match_isotopologue."""

    def __init__(self):
        self.rows = []

    def add(self, key, value):
        self.rows.append((value[2], value[3], value[4]))
