#!/usr/bin/env python
"""Quantify molecules in an mzML file on the GPU (the flow of the reference's example_scripts/quantify_mzml.py and
parse_ident_file_and_quantify.py + generate_quant_summary_file.py).

    python examples/quantify_mzml.py run.mzML [PEPTIDE ...]
    python examples/quantify_mzml.py run.mzML --evidence ids.csv [--summary quant_summary.csv] [--rt-tolerance 1.0]

Reads all MS1 spectra into packed arrays, matches them in one device pass and prints, per matched
(formula, charge, label tuple), the best mScore and the maximum-intensity amount of its chromatogram.  With an
identification table the molecules come from the evidence, the RT windows from the curated identification times,
and the amounts inside those windows are written to the quant summary.
"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pyqms_b200 as pyqms  # noqa: E402
from pyqms_b200.mzml import read_mzml  # noqa: E402

CAM = {"C": [{"element_composition": {"O": 1, "H": 3, "14N": 1, "C": 2}, "evidence_mod_name": "Carbamidomethyl"}]}


def main(mzml, molecules, evidence=None, summary=None, rt_tolerance=1.0, charges=(2, 3, 4, 5)):
    fixed_labels, evidences = None, None
    if evidence is not None:
        fixed_labels, evidences, from_evidence = pyqms.adaptors.parse_evidence(fixed_labels=CAM, evidence_files=[evidence],
                                                                               molecules=list(molecules))
        molecules = from_evidence
    lib = pyqms.IsotopologueLibrary(molecules=molecules, charges=list(charges), metabolic_labels=None,
                                    fixed_labels=fixed_labels, evidences=evidences, verbose=True)
    run = read_mzml(mzml, ms_level=1)
    results = lib.match_many(run.peaks, file_name=run.name, spec_ids=run.spec_ids, spec_rts=run.rts, offsets=run.offsets)
    amounts = lib.calc_amounts(results)
    for key in results.keys():
        molecule = results.lookup["formula to molecule"][key.formula][0]
        amount = amounts[key]
        print("{0:<28} z={1} {2}  matches={3:<5} max mScore={4:.3f}  max I={5:.4g} at rt {6:.2f} min".format(
            molecule, key.charge, key.label_percentiles, results[key]["len_data"], results[key]["max_score"],
            amount["max I in window"], amount["max I in window (rt)"]))
    if evidence is not None:
        summary = summary or str(Path(mzml).with_suffix(".quant_summary.csv"))
        results.write_rt_info_file(output_file=summary, rt_border_tolerance=rt_tolerance)
        lines = results.calc_amounts_from_rt_info_file(rt_info_file=summary, rt_border_tolerance=rt_tolerance, on_device=True)
        quantified = sum(line.get("max I in window", "") != "" for line in lines)
        print("> {0} of {1} summary lines quantified inside their evidence RT window -> {2}".format(quantified, len(lines), summary))
    return results


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("mzml")
    parser.add_argument("molecules", nargs="*")
    parser.add_argument("--evidence", help="identification table (.csv of Ursgal, or .mzTab)")
    parser.add_argument("--summary", help="quant summary / RT info file to write (.csv)")
    parser.add_argument("--rt-tolerance", type=float, default=1.0, help="RT border tolerance in minutes")
    args = parser.parse_args()
    default = [] if args.evidence else ["HLVDEPQNLIK", "YICDNQDTISSK", "DLGEEHFK"]
    main(args.mzml, args.molecules or default, evidence=args.evidence, summary=args.summary, rt_tolerance=args.rt_tolerance)
