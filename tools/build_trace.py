#!/usr/bin/env python
"""Wall-clock trace of a large library build (QMS_TRACE=1 prints the stages inside libqms_b200.so).

    QMS_TRACE=1 python tools/build_trace.py [n_formulas] [repeats] [upper_mz_limit]
"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pyqms_b200  # noqa: E402
from pyqms_b200.synthetic import averagine_formulas  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
params = {"UPPER_MZ_LIMIT": float(sys.argv[3]), "LOWER_MZ_LIMIT": 50} if len(sys.argv) > 3 else None
formulas = averagine_formulas(n, seed=20260105)
for repeat in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    t0 = time.perf_counter()
    plan = pyqms_b200.IsotopologueLibrary(molecules=formulas, charges=[1], verbose=False, _plan_only=True, params=params)
    t1 = time.perf_counter()
    lib = pyqms_b200.IsotopologueLibrary(molecules=formulas, charges=[1], verbose=False, params=params)
    t2 = time.perf_counter()
    c = lib.device_counters()
    print("repeat %d: plan only %.3f s, full build %.3f s, envelopes %d, ms_trees %.2f ms_envelopes %.3f ms_index %.2f -> %.0f envelopes/s end to end"
          % (repeat, t1 - t0, t2 - t1, c["envelopes"], c["ms_trees"], c["ms_envelopes"], c["ms_index"], c["envelopes"] / (t2 - t1)), flush=True)
    del lib
