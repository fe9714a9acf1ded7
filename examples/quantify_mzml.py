#!/usr/bin/env python
"""Quantify molecules in an mzML file on the GPU (the flow of the reference's example_scripts/quantify_mzml.py).

    python examples/quantify_mzml.py run.mzML [PEPTIDE ...]

Reads all MS1 spectra into packed arrays, matches them in one device pass and prints, per matched
(formula, charge, label tuple), the best mScore and the maximum-intensity amount of its chromatogram.
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pyqms_b200 as pyqms  # noqa: E402
from pyqms_b200.mzml import read_mzml  # noqa: E402


def main(mzml, molecules):
    lib = pyqms.IsotopologueLibrary(molecules=molecules, charges=[2, 3, 4, 5], metabolic_labels=None, fixed_labels=None,
                                    verbose=True)
    run = read_mzml(mzml, ms_level=1)
    results = lib.match_many(run.peaks, file_name=run.name, spec_ids=run.spec_ids, spec_rts=run.rts, offsets=run.offsets)
    amounts = lib.calc_amounts(results)
    for key in results.keys():
        molecule = results.lookup["formula to molecule"][key.formula][0]
        amount = amounts[key]
        print("{0:<28} z={1} {2}  matches={3:<5} max mScore={4:.3f}  max I={5:.4g} at rt {6:.2f} min".format(
            molecule, key.charge, key.label_percentiles, results[key]["len_data"], results[key]["max_score"],
            amount["max I in window"], amount["max I in window (rt)"]))
    return results


if __name__ == "__main__":
    if len(sys.argv) < 2:
        print(__doc__)
        sys.exit(1)
    main(sys.argv[1], sys.argv[2:] or ["HLVDEPQNLIK", "YICDNQDTISSK", "DLGEEHFK"])
