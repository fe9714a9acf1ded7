"""Host planning of the library build in bulk: molecule strings -> count matrix -> envelope term tables.

``parse_molecules`` / ``hill_formulas`` bind the native planner of ``libqms_b200.so`` (``qms_parse_molecules``,
``qms_hill_formulas``: plain formulas and unmodified peptides in one pass over a packed text buffer); molecules it
flags go through ``chemistry.ChemicalComposition`` one by one.  ``envelope_terms`` derives the per-envelope
(pool, atom count) lists of IL:546-646 for all formulas x label tuples with array operations.
"""
import ctypes as C
import re

import numpy as np

from . import _capi
from .chemistry import parse_formula

ABSENT = -1
_RESIDUE_NAME = re.compile(r"^[A-Z][0-9]*$")


def _pack(strings):
    """One byte buffer + offsets (n+1) of ``strings`` laid back to back."""
    blob = "".join(strings).encode("utf-8")
    lengths = np.fromiter(map(len, strings), dtype=np.int64, count=len(strings))
    if int(lengths.sum()) != len(blob):                 # non-ASCII text: byte lengths differ from character counts
        lengths = np.fromiter((len(s.encode("utf-8")) for s in strings), dtype=np.int64, count=len(strings))
    offsets = np.zeros(len(strings) + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    return blob, offsets


def _symbol_table(symbols):
    blob = "".join(symbols).encode("ascii")
    offsets = np.zeros(len(symbols) + 1, dtype=np.int32)
    np.cumsum([len(s) for s in symbols], out=offsets[1:])
    return blob, offsets


_PEPTIDE_BYTES = np.zeros(256, dtype=bool)
_PEPTIDE_BYTES[[ord(c) for c in "ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789"]] = True


def parse_molecules(molecules, symbols, aa_compositions, compact=False):
    """(counts[n,S] int32, first_seen[n,S] int16, status[n] uint8) of ``molecules`` over the element ``symbols``.
    ``first_seen``: rank at which a symbol entered the molecule's composition (ABSENT if not part of it);
    status 1 = left to the host parser (the row is empty).

    ``compact=True`` returns a fourth value, the symbols the columns stand for: only those the text can hold -- spelled with
    characters that occur in it, or part of a residue whose letter occurs (a knowledge base lists ~100 symbols, peptides
    and CHNOS formulas use 5-15: the full-width matrices of a 500 000-peptide library were 300 MB of zeros)."""
    n, n_symbols = len(molecules), len(symbols)
    usable = not (n == 0 or n_symbols == 0 or n_symbols > 127 or not all(s.isascii() for s in symbols))
    names, rows = [], []
    if usable:
        known = set(symbols)
        for name, value in aa_compositions.items():          # "A", "C", and the fixed-label variants "C0", "K1", ...
            if not _RESIDUE_NAME.match(name):
                continue
            items = parse_formula(value) if isinstance(value, str) else dict(value)
            if all(symbol in known for symbol in items):
                names.append(name)
                rows.append(items)
        text, str_off = _pack(molecules)
        if compact:
            # the symbols the text can possibly hold: those spelled with characters that occur in it, those of the residues
            # whose letter occurs, and water (a peptide's H2O)
            held = np.bincount(np.frombuffer(text, dtype=np.uint8), minlength=256) > 0
            wanted = {"H", "O"} | {s for s in symbols if all(held[ord(c)] for c in s)}
            for name, items in zip(names, rows):
                if held[ord(name[0])]:
                    wanted.update(items)
            if len(wanted) < n_symbols:
                symbols = [s for s in symbols if s in wanted]
                n_symbols = len(symbols)
                kept = [k for k, items in enumerate(rows) if all(symbol in wanted for symbol in items)]
                names, rows = [names[k] for k in kept], [rows[k] for k in kept]
    counts = np.zeros((n, n_symbols), dtype=np.int32)
    first_seen = np.full((n, n_symbols), ABSENT, dtype=np.int16)
    status = np.ones(n, dtype=np.uint8)
    if not usable or n_symbols == 0:
        return (counts, first_seen, status, symbols) if compact else (counts, first_seen, status)
    column = {s: i for i, s in enumerate(symbols)}
    residue_counts = np.zeros((len(names), n_symbols), dtype=np.int32)
    residue_order = np.full((len(names), n_symbols), ABSENT, dtype=np.int8)
    for r, items in enumerate(rows):
        for k, (symbol, count) in enumerate(items.items()):
            residue_counts[r, column[symbol]] = count
            residue_order[r, k] = column[symbol]
    residue_blob, residue_off = _symbol_table(names)
    symbols_blob, symbol_off = _symbol_table(symbols)
    rc = _capi.load().qms_parse_molecules(text, _capi.ptr(str_off), n, n_symbols, symbols_blob, _capi.ptr(symbol_off),
                                          len(names), residue_blob, _capi.ptr(residue_off), _capi.ptr(residue_counts),
                                          _capi.ptr(residue_order), _capi.ptr(counts), _capi.ptr(first_seen), _capi.ptr(status))
    if rc != _capi.QMS_OK:
        raise _capi.QmsError(rc, "qms_parse_molecules: bad argument")
    return (counts, first_seen, status, symbols) if compact else (counts, first_seen, status)


def hill_formulas(counts, symbols):
    """Hill-notation key of every count row (list of str), ``C(34)H(53)N(7)O(15)``."""
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    n = len(counts)
    if n == 0:
        return []
    lib = _capi.load()
    symbols_blob, symbol_off = _symbol_table(symbols)
    offsets = np.zeros(n + 1, dtype=np.int64)
    needed = C.c_int64()
    # one pass into a buffer that is large enough for any counts: "(-2147483648)" is 13 characters
    room = n * (len(symbols_blob) + 13 * len(symbols))
    out = C.create_string_buffer(max(1, room))
    rc = lib.qms_hill_formulas(_capi.ptr(counts), n, len(symbols), symbols_blob, _capi.ptr(symbol_off), out, room,
                               _capi.ptr(offsets), C.byref(needed))
    if rc != _capi.QMS_OK:
        raise _capi.QmsError(rc, "qms_hill_formulas failed")
    text = C.string_at(out, needed.value).decode("ascii")
    bounds = offsets.tolist()
    return [text[a:b] for a, b in zip(bounds[:-1], bounds[1:])]


def envelope_terms(counts, first_seen, symbols, label_tuples, fold_target, labelled, pool_of_default, pool_of_label,
                   rows_per_block=1 << 16):
    """Term lists of all envelopes, formula-major / label-tuple-minor: (term_off int64, term_pool int32, term_level int32).

    Per formula and label tuple (IL:546-646): the symbols of the composition in the order they entered it, a
    fixed-label isotope folded into its bare element when the label tuple carries that element at the isotope's
    enrichment (``fold_target[l][s]`` = column the symbol counts towards), metabolically labelled elements
    (``labelled`` mask) left out -- each a term on the element's default pool -- followed by the label tuple's own
    elements, in its order, on their labelled pool (``pool_of_label[l]`` = [(column, pool)]) when the formula holds them."""
    n_formulas, n_symbols = counts.shape
    n_lp = len(label_tuples)
    big = np.int16(32767)
    width = n_symbols + max((len(lp) for lp in label_tuples), default=0)
    offs, pools, levels = [np.zeros(1, dtype=np.int64)], [], []
    total = 0
    step = max(1, rows_per_block // max(1, n_lp))
    for begin in range(0, n_formulas, step):
        c = counts[begin:begin + step]
        seen = first_seen[begin:begin + step]
        rows = len(c)
        pool = np.zeros((rows, n_lp, width), dtype=np.int32)
        level = np.zeros((rows, n_lp, width), dtype=np.int32)
        valid = np.zeros((rows, n_lp, width), dtype=bool)
        present = seen >= 0
        for l in range(n_lp):
            target = fold_target[l]
            merged = np.zeros((rows, n_symbols), dtype=np.int64)
            rank = np.full((rows, n_symbols), big, dtype=np.int16)
            for s in range(n_symbols):
                t = target[s]
                merged[:, t] += c[:, s]
                np.minimum(rank[:, t], np.where(present[:, s], seen[:, s], big), out=rank[:, t])
            in_formula = rank < big
            key = np.where(in_formula & ~labelled[None, :], rank, big)
            order = np.argsort(key, axis=1, kind="stable")
            sorted_key = np.take_along_axis(key, order, axis=1)
            pool[:, l, :n_symbols] = pool_of_default[order]
            level[:, l, :n_symbols] = np.take_along_axis(merged, order, axis=1)
            valid[:, l, :n_symbols] = sorted_key < big
            for j, (column, labelled_pool) in enumerate(pool_of_label[l]):
                if column is None:
                    continue
                held = merged[:, column]
                pool[:, l, n_symbols + j] = labelled_pool
                level[:, l, n_symbols + j] = held
                valid[:, l, n_symbols + j] = (held != 0) & in_formula[:, column]
        per_envelope = valid.sum(axis=2).reshape(-1)
        offs.append(total + np.cumsum(per_envelope, dtype=np.int64))
        total += int(per_envelope.sum())
        pools.append(pool[valid])
        levels.append(level[valid])
    term_off = np.concatenate(offs)
    term_pool = np.concatenate(pools) if pools else np.zeros(0, dtype=np.int32)
    term_level = np.concatenate(levels) if levels else np.zeros(0, dtype=np.int32)
    return term_off, term_pool.astype(np.int32, copy=False), term_level.astype(np.int32, copy=False)
