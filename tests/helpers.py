"""Shared helpers: load golden fixtures, rebuild objects from them, compare result sets."""
import gzip
import json
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"


def load_case(name):
    with gzip.open(GOLDEN / (name + ".json.gz"), "rt") as fh:
        blob = json.load(fh)
    spectra = np.load(GOLDEN / (name + ".npz"))
    return blob, spectra["peaks"], spectra["offsets"]


def lp_tuple(raw):
    return tuple(tuple(p) for p in raw)


def results_to_rows(res):
    """{m_key: [match,...]} or a Results-like {m_key: {"data": [...]}} -> comparable nested lists."""
    out = []
    for key, val in res.items():
        data = val["data"] if isinstance(val, dict) else val
        rows = []
        for m in data:
            rows.append([m[0], None if m[1] is None else float(m[1]), float(m[2]), float(m[3]),
                         [[None if a is None else float(a), None if b is None else float(b), float(c), float(d), int(e)]
                          for a, b, c, d, e in m[4]]])
        out.append([[key[0], key[1], key[2], [list(p) for p in key[3]]], rows])
    return out


def assert_results_equal(got, want, score_tol=0.0, cmz_rel=0.0, ri_rel=0.0):
    """Hit sets, keys, key order and matched (mmz, mi) exact; score/scaling within score_tol;
    cmz / ri within the relative tolerances (0 = bit-exact)."""
    assert [g[0] for g in got] == [w[0] for w in want], "key sets / first-hit order differ"
    for (key, grows), (_, wrows) in zip(got, want):
        assert len(grows) == len(wrows), (key, len(grows), len(wrows))
        for g, w in zip(grows, wrows):
            assert g[0] == w[0] and g[1] == w[1], (key, g[:2], w[:2])
            assert abs(g[2] - w[2]) <= score_tol, (key, g[0], g[2], w[2])
            assert abs(g[3] - w[3]) <= score_tol * max(1.0, abs(w[3])), (key, g[0], g[3], w[3])
            assert len(g[4]) == len(w[4])
            for gp, wp in zip(g[4], w[4]):
                assert gp[0] == wp[0] and gp[1] == wp[1], (key, g[0], gp, wp)      # matched peak set: exact
                assert abs(gp[2] - wp[2]) <= ri_rel * abs(wp[2]), (key, gp, wp)
                assert abs(gp[3] - wp[3]) <= cmz_rel * abs(wp[3]), (key, gp, wp)
                assert gp[4] == wp[4], (key, gp, wp)
