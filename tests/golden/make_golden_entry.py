#!/usr/bin/env python
"""Golden vectors for match_isotopologue(index, mz_i_list=<list>) (IL:1885-2040), from the UNMODIFIED reference.

    python tests/golden/make_golden_entry.py        ->  tests/golden/entry_list.json

The list path slices the spectrum by +-2 ELEMENTS around the entry's own (lower_mz, upper_mz) (IL:1957-1966, Q12):
peaks crowd the first and last window so that the element-wise slice decides what is seen.
"""
import json
import sys
import warnings
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / "_standin"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(HERE.parent.parent))
warnings.filterwarnings("ignore")
try:
    import pyqms  # noqa: E402  the reference (build container only; the tests import LIB from this module)
except ImportError:
    pyqms = None
from pyqms_b200.synthetic import BSA_MOLECULES, CAM_FIXED_LABEL  # noqa: E402

LIB = dict(molecules=BSA_MOLECULES[:6], charges=[2, 3], fixed_labels=CAM_FIXED_LABEL)


def main():
    lib = pyqms.IsotopologueLibrary(verbose=False, **LIB)
    rng = np.random.default_rng(31)
    cases = []
    for index, (lower, upper, z, lp, formula) in enumerate(lib.formulas_sorted_by_mz):
        env = lib[formula]["env"][lp]
        keep = [n for n, p in enumerate(env["c_peak_pos"]) if p is not None]
        cmz = np.array([env[z]["mz"][n] for n in keep])
        ci = np.array([env["abun"][n] for n in keep], dtype=float)
        for variant in range(4):
            mz = list(cmz * (1 + rng.normal(0, 1.5e-6, len(cmz))))
            inten = list(ci * 50.0 * (1 + rng.normal(0, 0.05, len(cmz))))
            # crowd both ends: peaks just below lower_mz / above upper_mz, some inside the first and last window
            for edge, sign in ((lower, -1), (upper, +1)):
                for k in range(int(rng.integers(0, 6))):
                    mz.append(edge + sign * rng.uniform(0.0, 0.02))
                    inten.append(float(10 ** rng.uniform(2, 5)))
                for k in range(int(rng.integers(0, 4))):
                    mz.append(edge * (1 + rng.normal(0, 3e-6)))
                    inten.append(float(10 ** rng.uniform(2, 5)))
            if variant == 3:                                   # most of the pattern missing
                drop = rng.random(len(mz)) < 0.6
                mz = [m for m, d in zip(mz, drop) if not d]
                inten = [i for i, d in zip(inten, drop) if not d]
            for k in range(int(rng.integers(0, 30))):          # background
                mz.append(float(rng.uniform(lower - 30, upper + 30)))
                inten.append(float(10 ** rng.uniform(2, 5)))
            order = np.lexsort((inten, mz))
            spectrum = [(float(mz[i]), float(inten[i])) for i in order]
            found = lib.match_isotopologue(index=index, mz_i_list=spectrum, mz_score_percentile=0.4)
            row = None
            if found is not None:
                score, scaling, peaks = found
                row = [float(score), float(scaling), [[None if a is None else float(a), None if b is None else float(b), float(c),
                                                       float(d), int(e)] for a, b, c, d, e in peaks]]
            cases.append([index, spectrum, row])
    (HERE / "entry_list.json").write_text(json.dumps({"cases": cases}, separators=(",", ":")))
    print(len(cases), "cases,", sum(1 for c in cases if c[2] is not None), "with a match")


if __name__ == "__main__":
    main()
