"""Result container of ``match_all`` -- the boundary type of the matching path.

Mirrors the core of the reference's ``pyqms.Results`` (pyqms/results.py:40-57, 81-203): a ``dict``
keyed by ``m_key(file_name, formula, charge, label_percentiles)`` whose values hold the list of
``match(spec_id, rt, score, scaling_factor, peaks)`` tuples plus the running maximum score.
Reporting (csv/xlsx/mzTab writers, RT-window curation, plots; results.py:205-2064) is out of scope
of this build and works on top of this structure unchanged.
"""
import ast
import codecs
import copy
import csv
import sys
import weakref
from collections import namedtuple

m_key = namedtuple("m_key", ["file_name", "formula", "charge", "label_percentiles"])
match = namedtuple("match", ["spec_id", "rt", "score", "scaling_factor", "peaks"])
combined = namedtuple("combined", ["file_name", "formula", "charge", "label_percentiles", "spec_id", "rt",
                                   "score", "scaling_factor", "peaks"])


class _FlushingIndex(dict):
    """``Results.index`` that brings the owning container up to date before it is read."""

    def __init__(self, owner, content):
        super().__init__(content)
        self._owner = weakref.ref(owner)      # no reference cycle: a dropped Results frees its packed (page-locked) arrays at once

    def __getitem__(self, key):
        owner = self._owner()
        if owner is not None and (owner._queue or owner._pending):
            owner._flush()
        return dict.__getitem__(self, key)

    def __reduce__(self):
        return (dict, (dict(self),))          # pickles hold a plain dict, like the reference's Results.index


_TRANSIENT_STATE = ("_pending", "_packed", "_tables", "_queue", "_queue_peaks", "_queue_library", "_busy")


def _restore_results(cls, state, items):
    obj = cls.__new__(cls)
    dict.update(obj, items)
    obj.__dict__.update(state)
    obj.__dict__.update(_pending=[], _packed=[], _tables=[], _queue=[], _queue_peaks=0, _queue_library=None, _busy=False)
    obj.index = _FlushingIndex(obj, dict(obj.index))
    return obj


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
        self._tables = []           # per packed batch: callable giving its hits as columns
        self._queue = []            # spectra handed to match_all and not yet matched
        self._queue_peaks = 0
        self._queue_library = None
        self._busy = False
        self.index = _FlushingIndex(self, self.index)

    # ---- deferred materialisation of packed batches -------------------------------------------------
    def defer(self, packed, thunk, table=None):
        """Register a packed batch; ``thunk(self)`` must ``add`` its hits in (spectrum, entry) order;
        ``table()`` returns the batch as columns for ``packed_table``."""
        self._pending.append(thunk)
        self._packed.append(packed)
        self._tables.append(table)

    def packed_table(self):
        """The hits of all batched calls (``match_many``) as a pandas DataFrame with the columns of
        ``format_all_results`` except ``peaks`` (plus ``n_peaks`` / ``n_matched_peaks``), built from the packed
        arrays without creating a Python object per hit; rows in (batch, spectrum, index entry) order.
        Hits added one by one (eager ``match_all``, ``add``) are not part of it."""
        import pandas as pd
        frames = [pd.DataFrame(table()) for table in self._tables if table is not None]
        if not frames:
            return pd.DataFrame(columns=list(m_key._fields) + ["spec_id", "rt", "score", "scaling_factor", "n_peaks",
                                                               "n_matched_peaks"])
        return pd.concat(frames, ignore_index=True)

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
        """Pickles and deep copies hold the materialised matches only: the packed arrays are redundant once the match
        tuples exist and the device context is not picklable.  The live object keeps both."""
        self._flush()
        state = {k: v for k, v in self.__dict__.items() if k not in _TRANSIENT_STATE}
        return (_restore_results, (type(self), state, list(dict.items(self))))

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

    def _append_match(self, key, entry):
        """``add`` for emitters that hand over ready ``m_key`` / ``match`` objects in order (no flush, no re-wrapping;
        the index sets only change when a key is new)."""
        slot = dict.get(self, key)
        if slot is None:
            slot = {"data": [], "max_score": -1, "max_score_index": -1, "len_data": 0}
            dict.__setitem__(self, key, slot)
            index = self.index
            dict.__getitem__(index, "files").add(key.file_name)
            dict.__getitem__(index, "charges").add(key.charge)
            dict.__getitem__(index, "formulas").add(key.formula)
            dict.__getitem__(index, "label_percentiles").add(key.label_percentiles)
        if slot["max_score"] < entry.score:
            slot["max_score"] = entry.score
            slot["max_score_index"] = len(slot["data"])
        slot["len_data"] += 1
        slot["data"].append(entry)

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

    # ---- RT windows from identification evidence ----------------------------------------
    def _determine_rt_windows_from_evidence(self, rt_border_tolerance=None):
        """``(rt_border_lookup, molecule_lookup)`` per formula of ``lookup["formula to evidences"]``
        (results.py:1283-1300)."""
        tolerance = 0 if rt_border_tolerance is None else rt_border_tolerance
        borders, molecules = {}, {}
        for formula, evidence_dict in (self.lookup or {}).get("formula to evidences", {}).items():
            if evidence_dict is None:
                continue
            molecules[formula] = evidence_dict.keys()
            borders[formula] = self.curate_rt_windows(evidence_dict, rt_tolerance=tolerance)
        return borders, molecules

    def curate_rt_windows(self, evidence_dict, rt_tolerance):
        """Identification RT span ``[min, max]`` of every molecule that shares a formula, with the borders
        between neighbouring spans settled by ``window_curator`` (results.py:1937-1967)."""
        spans = []
        for molecule, info in evidence_dict.items():
            times = [evidence["RT"] for evidence in info["evidences"]]
            spans.append(([min(times), max(times)], molecule))
        if len(spans) > 1:
            return self.window_curator(spans, rt_tolerance=rt_tolerance)
        return {spans[0][1]: {"rt_window": spans[0][0]}}

    def window_curator(self, window_sort_list, rt_tolerance=None, win_key=None):
        """Walk the RT spans in ascending order; where two neighbours come closer than twice the tolerance,
        either split the gap (or overlap) half-way -- ``upper_window_border`` of the earlier one and
        ``lower_window_border`` of the later one, negative for an overlap -- or flag both as
        ``window_is_unseparable`` when the split point would fall outside one of them or one span encloses
        the other (results.py:1969-2064)."""
        window_sort_list.sort()
        out = {}

        def border(name, which):
            return out[name].get(which, rt_tolerance)

        for position, (span, name) in enumerate(window_sort_list):
            out[name] = {"rt_window": span}
            if position == 0:
                continue
            earlier_span, earlier = window_sort_list[position - 1]
            if not span[0] - rt_tolerance <= earlier_span[-1] + rt_tolerance:
                continue
            half_gap = (span[0] - earlier_span[-1]) / 2.0
            if span[0] - half_gap > span[1] + border(name, "upper_window_border"):
                clash = True
            elif earlier_span[-1] - half_gap < earlier_span[0] - border(earlier, "lower_window_border"):
                clash = True
            elif earlier_span[0] - border(earlier, "lower_window_border") < span[0] - border(name, "lower_window_border"):
                clash = earlier_span[-1] + border(earlier, "upper_window_border") > span[-1] + border(name, "upper_window_border")
            else:
                clash = False
            if clash:
                for member in (name, earlier):
                    out[member].setdefault("window_is_unseparable", []).append(win_key)
                    out[member].pop("lower_window_border", None)
                    out[member].pop("upper_window_border", None)
            else:
                out[name]["lower_window_border"] = half_gap
                out[earlier]["upper_window_border"] = half_gap
        return out

    # ---- quant summary / RT info file (results.py:1138-1281, 1302-1527) -------------------------------------
    def write_rt_info_file(self, output_file=None, list_of_csvdicts=None, trivial_name_lookup=None,
                           rt_border_tolerance=None, update=True, buffer_only=False):
        """One line per (result key, molecule) with the RT window derived from the identification evidence
        (``rt_window`` widened by the curated borders or the tolerance), the evidence list ``score@rt;...`` and
        the trivial names; columns ``RT_INFO_FIELDS``.  Returns the lines; writes ``.csv`` unless ``buffer_only``.
        Mirrors ``pyqms.Results.write_rt_info_file`` (csv only: openpyxl is not a dependency here)."""
        assert output_file is not None, "You need to specify an output file"
        tolerance = 0 if rt_border_tolerance is None else rt_border_tolerance
        if list_of_csvdicts is None:
            list_of_csvdicts = [key._asdict() for key in sorted(self.keys())]
        names_of_formula = {} if trivial_name_lookup is None else trivial_name_lookup
        borders_of_formula, molecules_of_formula = self._determine_rt_windows_from_evidence(rt_border_tolerance=tolerance)
        lines = []
        for line in list_of_csvdicts:
            if not update:
                lines.append(copy.deepcopy(line))
                continue
            formula = line["formula"]
            extra_names = ", ".join(names_of_formula.get(formula, []))
            if line.get("trivial_name(s)", None) is None:
                line["trivial_name(s)"] = extra_names
            elif line["trivial_name(s)"] != "":
                line["trivial_name(s)"] += extra_names
            borders = borders_of_formula.get(formula, {})
            evidence = self.lookup["formula to evidences"].get(formula, None) if len(borders) > 0 else None
            if evidence is not None:
                molecules = molecules_of_formula[formula]
            else:
                molecules = self.lookup.get("formula to molecule", {}).get(formula, [])
            for molecule in molecules:
                line["molecule"] = molecule
                if evidence is not None:
                    listed = []
                    for item in evidence[molecule]["evidences"]:
                        if "RT" not in item:
                            continue
                        text = "{0}@{1}".format(item["score"], item["RT"]) if item["score"] != "None" else "{0}".format(item["RT"])
                        listed.append((item["RT"], text))
                    window = borders[molecule]
                    line["start (min)"] = window["rt_window"][0] - window.get("lower_window_border", tolerance)
                    line["stop (min)"] = window["rt_window"][1] + window.get("upper_window_border", tolerance)
                    line["evidences (min)"] = ";".join(text for _, text in sorted(listed))
                    if len(evidence[molecule]["trivial_names"]) > 0:
                        line["trivial_name(s)"] = ";".join(sorted(set(evidence[molecule]["trivial_names"])))
                lines.append(copy.deepcopy(line))
        if not buffer_only:
            if not str(output_file).endswith(".csv"):
                raise ValueError("Extension: {0} of file {1} not recognized (csv only)".format(
                    str(output_file).split(".")[-1], output_file))
            with codecs.open(str(output_file), mode="w", encoding="utf-8") as handle:
                writer = csv.DictWriter(handle, RT_INFO_FIELDS, extrasaction="ignore",
                                        lineterminator="\n" if sys.platform == "win32" else "\r\n")
                writer.writeheader()
                for line in lines:
                    writer.writerow(line)
        return lines

    def calc_amounts_from_rt_info_file(self, rt_info_file=None, rt_border_tolerance=None, calc_amount_function=None,
                                       evidence_score_field="PEP", buffer_only=False, buffered_csv_dicts=None,
                                       on_device=False):
        """Amounts of every line of a quant summary that carries ``start (min)`` / ``stop (min)``: the matches of
        the line's key inside the window (retention time in minutes; the scan stops at the first match beyond the
        window) go to ``calc_amount_function`` -- default ``determine_max_itensity`` -- and its result is merged
        into the line.  Mirrors ``pyqms.Results.calc_amounts_from_rt_info_file``.

        ``on_device=True`` (additive) computes all five amount fields of all lines in ONE pass of
        ``qms_calc_amounts`` instead of calling a Python function per line."""
        if rt_info_file is not None:
            if not str(rt_info_file).endswith(".csv"):
                raise ValueError("Extension: {0} of file {1} not recognized (csv only)".format(
                    str(rt_info_file).split(".")[-1], rt_info_file))
            with codecs.open(str(rt_info_file), mode="r", encoding="utf-8") as handle:
                given = list(csv.DictReader(handle))
        elif buffered_csv_dicts is not None:
            given = buffered_csv_dicts
        else:
            raise ValueError("No RT info file or buffered csv lines were provided")
        jobs = []       # (line, profile of the key, (start, stop))
        for line in given:
            try:
                line["start (min)"] = float(line.get("start (min)", None))
                line["stop (min)"] = float(line.get("stop (min)", None))
            except (TypeError, ValueError):
                continue
            percentiles = line["label_percentiles"]
            if isinstance(percentiles, str):
                percentiles = ast.literal_eval(percentiles)
            key = self._m_key_class(line["file_name"], line["formula"], int(line["charge"]), percentiles)
            jobs.append((line, self[key]["data"], (line["start (min)"], line["stop (min)"])))
        if on_device:
            from .isotopologue_library import calc_amounts_of_profiles
            amounts = calc_amounts_of_profiles([job[1] for job in jobs], [job[2] for job in jobs])
        else:
            function = self.determine_max_itensity if calc_amount_function is None else calc_amount_function
            amounts = [function(_profile_in_window(profile, *window)) for _, profile, window in jobs]
        for (line, _, _), amount in zip(jobs, amounts):
            if amount is not None:
                line.update(amount)
        if buffer_only is False:
            self.write_rt_info_file(output_file=rt_info_file, list_of_csvdicts=list(given),
                                    rt_border_tolerance=rt_border_tolerance, update=False, buffer_only=buffer_only)
        return list(given)


RT_INFO_FIELDS = ["file_name", "formula", "molecule", "trivial_name(s)", "label_percentiles", "charge", "start (min)",
                  "stop (min)", "max I in window", "max I in window (rt)", "max I in window (score)", "auc in window",
                  "sum I in window", "evidences (min)"]


def _profile_in_window(profile, start, stop):
    """``obj_for_calc_amount`` of one key: the matches with start <= rt <= stop, scanning stops at the first
    rt > stop (results.py:1238-1262)."""
    obj = {"rt": [], "i": [], "scores": [], "spec_ids": []}
    for entry in profile:
        try:
            rt = entry.rt[0] / 60 if entry.rt[1] == "second" else entry.rt[0]
        except TypeError:
            rt = entry.rt
        if rt < start:
            continue
        if rt > stop:
            break
        obj["rt"].append(rt)
        obj["i"].append(entry.scaling_factor)
        obj["scores"].append(entry.score)
        obj["spec_ids"].append(entry.spec_id)
    return obj
