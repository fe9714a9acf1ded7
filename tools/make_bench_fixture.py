#!/usr/bin/env python
"""Write bench_data/proteome500k_shard0_planted.npz (run on the GPU box; lands in gpurun_out/ and is copied from there).

The planted-peak table of shard 0 of the proteome run: the theoretical c-peak m/z and integer abundances of the 2 % of
index entries that elute in the shard (pyqms_b200.synthetic.synth_run, return_planted).  With it bench.py draws the
identical spectra in a process that has no library -- the reference arm (`bench.py --impl reference`) and the
cpu_baseline leg see the same input as the GPU arm.  The table is data of the benchmark workload, not of the product.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import pyqms_b200  # noqa: E402
from pyqms_b200.synthetic import CAM_FIXED_LABEL, random_tryptic_peptides, synth_run  # noqa: E402


def main():
    peptides = random_tryptic_peptides(500000, seed=20260104)
    lib = pyqms_b200.IsotopologueLibrary(molecules=peptides, charges=[1, 2, 3, 4, 5], fixed_labels=CAM_FIXED_LABEL, verbose=False)
    mz, ci = lib.entry_peaks()
    peaks, offsets, table = synth_run(mz, ci, n_spectra=25000, seed=20260104, entry_fraction=0.02, return_planted=True)
    out = ROOT / "gpurun_out" / "proteome500k_shard0_planted.npz"
    out.parent.mkdir(exist_ok=True)
    np.savez_compressed(out, n_entries=np.int64(table["n_entries"]), lens=table["lens"].astype(np.int16), mz=table["mz"],
                        ci=table["ci"].astype(np.int32))
    again, off2 = synth_run(None, None, n_spectra=25000, seed=20260104, entry_fraction=0.02,
                            planted={"n_entries": table["n_entries"], "lens": table["lens"], "mz": table["mz"], "ci": table["ci"]})
    assert np.array_equal(again, peaks) and np.array_equal(off2, offsets)
    print("entries", table["n_entries"], "planted entries", len(table["lens"]), "c-peaks", len(table["mz"]), "run peaks", len(peaks),
          "file bytes", out.stat().st_size)


if __name__ == "__main__":
    main()
