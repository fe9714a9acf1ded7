#!/usr/bin/env python
"""Per-spectrum match_all loop: eager vs deferred queue (what a pyQms user script does, quantify_mzml.py:61-73)."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import pyqms_b200  # noqa: E402
from pyqms_b200.synthetic import BSA_MOLECULES, CAM_FIXED_LABEL, synth_run  # noqa: E402

kw = dict(molecules=BSA_MOLECULES, charges=[1, 2, 3, 4], metabolic_labels={"15N": [0, 0.99]}, fixed_labels=CAM_FIXED_LABEL, verbose=False)
out = {}
for name, defer in (("eager", False), ("deferred", True)):
    lib = pyqms_b200.IsotopologueLibrary(defer_matching=defer, **kw)
    mz, ci = lib.entry_peaks()
    peaks, offsets = synth_run(mz, ci, n_spectra=4000, seed=20260101)
    spectra = [peaks[offsets[s]:offsets[s + 1]] for s in range(4000)]
    res = None
    for s in range(50):
        res = lib.match_all(mz_i_list=spectra[s], file_name="w", spec_id=s, spec_rt=0.0, results=res)
    len(res)
    res = None
    t0 = time.perf_counter()
    for s, spec in enumerate(spectra):
        res = lib.match_all(mz_i_list=spec, file_name="run", spec_id=s, spec_rt=s * 0.2, results=res)
    t_loop = time.perf_counter() - t0
    n_keys = len(res)                       # first read: device pass + materialisation for the deferred path
    t_total = time.perf_counter() - t0
    out[name] = {"spectra": 4000, "loop_s": t_loop, "total_s_incl_first_read": t_total, "spectra_per_s": 4000 / t_total,
                 "us_per_call": 1e6 * t_total / 4000, "keys": n_keys, "hits": sum(v["len_data"] for v in res.values())}
print(json.dumps(out))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "loop_latency.json").write_text(json.dumps(out, indent=1))
