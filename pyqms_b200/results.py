"""Result container of ``match_all`` -- the boundary type of the matching path.

Mirrors the core of the reference's ``pyqms.Results`` (pyqms/results.py:40-57, 81-203): a ``dict``
keyed by ``m_key(file_name, formula, charge, label_percentiles)`` whose values hold the list of
``match(spec_id, rt, score, scaling_factor, peaks)`` tuples plus the running maximum score.
Reporting (csv/xlsx/mzTab writers, RT-window curation, plots; results.py:205-2064) is out of scope
of this build and works on top of this structure unchanged.
"""
from collections import namedtuple

m_key = namedtuple("m_key", ["file_name", "formula", "charge", "label_percentiles"])
match = namedtuple("match", ["spec_id", "rt", "score", "scaling_factor", "peaks"])
combined = namedtuple("combined", ["file_name", "formula", "charge", "label_percentiles", "spec_id", "rt",
                                   "score", "scaling_factor", "peaks"])


class _FlushingIndex(dict):
    """``Results.index`` that brings the owning container up to date before it is read."""

    def __init__(self, owner, content):
        super().__init__(content)
        self._owner = owner

    def __getitem__(self, key):
        if self._owner._queue or self._owner._pending:
            self._owner._flush()
        return dict.__getitem__(self, key)

    def __reduce__(self):
        return (dict, (dict(self),))          # pickles hold a plain dict, like the reference's Results.index


class Results(dict):
    """dict {m_key: {"data": [match, ...], "max_score", "max_score_index", "len_data"}}.

    Batched matching (``IsotopologueLibrary.match_many``) hands over *packed* hit arrays and a thunk that
    turns them into ``match`` tuples; the thunk runs on the first dict access (any read, ``add``, pickling),
    so a whole run can be matched at device speed and the Python objects are only built if somebody looks
    (SURVEY.md section 8f, N1).  ``packed_batches`` exposes the arrays without materialising anything."""

    def __init__(self, lookup=None, params=None, fixed_labels=None, metabolic_labels=None, aa_compositions=None,
                 isotopic_distributions=None, charges=None, verbose=False):
        super().__init__()
        self.params = params
        self.lookup = lookup if lookup is not None else {}
        self.fixed_labels = fixed_labels
        self.metabolic_labels = metabolic_labels
        self.aa_compositions = aa_compositions
        self.charges = charges
        self.isotopic_distribution = isotopic_distributions
        self.verbose = verbose
        self.index = {"files": set(), "charges": set(), "formulas": set(), "label_percentiles": set()}
        self._match_class = match
        self._m_key_class = m_key
        self._combined_class = combined
        self._silac_pairs = None
        self._pending = []          # thunks of packed batches not yet turned into match tuples
        self._packed = []           # every packed batch ever handed over (kept for array consumers)
        self._queue = []            # spectra handed to match_all and not yet matched
        self._queue_peaks = 0
        self._queue_library = None
        self._busy = False
        self.index = _FlushingIndex(self, self.index)

    # ---- deferred materialisation of packed batches -------------------------------------------------
    def defer(self, packed, thunk):
        """Register a packed batch; ``thunk(self)`` must ``add`` its hits in (spectrum, entry) order."""
        self._pending.append(thunk)
        self._packed.append(packed)

    def packed_batches(self):
        """The packed hit arrays of every batched call so far (see ``IsotopologueLibrary.match_packed``)."""
        return list(self._packed)

    def enqueue(self, library, item, n_peaks, max_spectra, max_peaks):
        """Queue one spectrum of a ``match_all`` loop; a full queue is matched at once (bounded memory)."""
        if self._queue_library is not None and self._queue_library is not library:
            self.flush_queue()
        self._queue_library = library
        self._queue.append(item)
        self._queue_peaks += n_peaks
        if len(self._queue) >= max_spectra or self._queue_peaks >= max_peaks:
            self.flush_queue()

    def flush_queue(self):
        """Match the queued spectra now (one device pass) and add their hits (after any older packed batch)."""
        self._flush()

    def flush(self):
        """Bring the container up to date: queued spectra are matched, packed batches materialised."""
        self._flush()

    def _flush(self):
        if self._busy or not (self._pending or self._queue):
            return
        self._busy = True           # add() re-enters _flush while a thunk or the queue is being emitted
        try:
            while self._pending or self._queue:
                if self._pending:   # packed batches are always older than the queue (match_many flushes it first)
                    pending, self._pending = self._pending, []
                    for thunk in pending:
                        thunk(self)
                else:
                    queue, library = self._queue, self._queue_library
                    self._queue, self._queue_peaks = [], 0
                    library._match_queue(queue, self)
        finally:
            self._busy = False

    def __getitem__(self, key):
        self._flush()
        return dict.__getitem__(self, key)

    def __iter__(self):
        self._flush()
        return dict.__iter__(self)

    def __len__(self):
        self._flush()
        return dict.__len__(self)

    def __contains__(self, key):
        self._flush()
        return dict.__contains__(self, key)

    def keys(self):
        self._flush()
        return dict.keys(self)

    def values(self):
        self._flush()
        return dict.values(self)

    def items(self):
        self._flush()
        return dict.items(self)

    def get(self, key, default=None):
        self._flush()
        return dict.get(self, key, default)

    def __reduce_ex__(self, protocol):
        self._flush()
        self._packed = []           # arrays are redundant once materialised; keep pickles reference-sized
        self._queue_library = None  # the device context is not picklable
        return dict.__reduce_ex__(self, protocol)

    def __repr__(self):
        self._flush()
        return dict.__repr__(self)

    def add(self, key, value):
        """Append one match (results.py:144-203); returns the formatted key."""
        self._flush()
        key = self._m_key_class(*key)
        slot = dict.get(self, key)
        if slot is None:
            slot = {"data": [], "max_score": -1, "max_score_index": -1, "len_data": 0}
            dict.__setitem__(self, key, slot)
        entry = self._match_class(*value)
        if slot["max_score"] < entry.score:
            slot["max_score"] = entry.score
            slot["max_score_index"] = len(slot["data"])
        slot["len_data"] += 1
        slot["data"].append(entry)
        index = self.index
        dict.__getitem__(index, "files").add(key.file_name)
        dict.__getitem__(index, "charges").add(key.charge)
        dict.__getitem__(index, "formulas").add(key.formula)
        dict.__getitem__(index, "label_percentiles").add(key.label_percentiles)
        return key

    def _translate_molecules_to_formulas(self, molecules, formulas):
        """Formulas of ``molecules`` (via ``lookup["molecule to formula"]``) added to ``formulas`` (results.py:343-358)."""
        formulas = set(formulas) if formulas is not None else set()
        table = self.lookup.get("molecule to formula", {})
        for molecule in molecules:
            if molecule in table:
                formulas.add(table[molecule])
        return formulas

    def _parse_and_filter(self, molecules=None, charges=None, file_names=None, label_percentiles=None, formulas=None):
        """Yield the keys that pass the given filters; ``None`` = do not filter (results.py:205-248)."""
        if molecules is not None:
            formulas = self._translate_molecules_to_formulas(molecules, formulas)
        for key in self.keys():
            if file_names is not None and key[0] not in file_names:
                continue
            if formulas is not None and key[1] not in formulas:
                continue
            if charges is not None and key[2] not in charges:
                continue
            if label_percentiles is not None and key[3] not in label_percentiles:
                continue
            yield key

    def extract_results(self, molecules=None, charges=None, file_names=None, label_percentiles=None, formulas=None,
                        score_threshold=None):
        """Yield (key, i, match) of every stored match that passes the filters (results.py:250-307)."""
        for key in self._parse_and_filter(molecules=molecules, charges=charges, file_names=file_names,
                                          label_percentiles=label_percentiles, formulas=formulas):
            for i, entry in enumerate(self[key]["data"]):
                if score_threshold is not None and entry.score < score_threshold:
                    continue
                yield key, i, entry

    def format_all_results(self):
        """All matches as a pandas DataFrame with the m_key and match fields as columns (results.py:309-341)."""
        import pandas as pd
        rows = [self._combined_class(*key, *entry) for key in self.keys() for entry in self[key]["data"]]
        return pd.DataFrame(rows)

    def max_score(self, molecules=None, charges=None, file_names=None, label_percentiles=None, formulas=None):
        """[best score, key, index, match] under the filters; [0, None, None, None] if nothing matches
        (results.py:360-398)."""
        best = [0, None, None, None]
        for key, i, entry in self.extract_results(molecules=molecules, charges=charges, file_names=file_names,
                                                  label_percentiles=label_percentiles, formulas=formulas):
            if entry.score > best[0]:
                best = [entry.score, key, i, entry]
        return best

    def determine_max_itensity(self, obj_for_calc_amount):
        """Default amount function: maximum intensity of an elution profile (results.py:400-447)."""
        if len(obj_for_calc_amount["i"]) == 0:
            return None
        top = max(obj_for_calc_amount["i"])
        at = obj_for_calc_amount["i"].index(top)
        return {"max I in window": top, "max I in window (rt)": obj_for_calc_amount["rt"][at],
                "max I in window (score)": obj_for_calc_amount["scores"][at]}
