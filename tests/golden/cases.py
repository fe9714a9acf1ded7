"""Parity cases shared by the golden generator (make_golden.py), the oracle tests and the GPU tests.

Each case = constructor kwargs of the library + how its synthetic spectra are drawn.  The run
generator is pyqms_b200.synthetic.synth_run (seeded); the drawn spectra are stored with the
goldens so the fixtures do not depend on the numpy version.
"""
from pyqms_b200.synthetic import BSA_MOLECULES, CAM_FIXED_LABEL, random_tryptic_peptides

QUICK_START_PEAKS = [  # example_scripts/basic_quantification_example.py:45-89 (input data of the docs KAT)
    (404.2492407565097, 2652.905029296875), (405.3003310237508, 4831.56103515625),
    (408.8403673369115, 23153.7109375), (409.17476109421705, 10182.2822265625),
    (409.5098740355617, 4770.97412109375), (411.17196124490727, 3454.364013671875),
    (413.26627826402705, 6861.84912109375), (419.3157903165357, 90201.5625),
    (420.2440507067882, 11098.4716796875), (420.31917273788645, 22288.9140625),
    (420.73825281590496, 8159.7099609375), (421.2406187369968, 3768.656494140625),
    (427.3787652898548, 5680.43212890625), (433.3316647490907, 8430.30859375),
    (434.705984428002, 25924.38671875), (435.2080179219357, 11041.2060546875),
    (443.6708762397708, 4081.282470703125), (443.69049198141124, 5107.13330078125),
    (443.6974813419733, 9135.3125), (443.7112735313511, 2517650.0),
    (443.7282222289076, 5571.26025390625), (443.7379762316008, 5227.4033203125),
    (444.1998579474954, 3021.341796875), (444.21248374593875, 1156173.75),
    (444.71384916266277, 336326.96875), (445.21533524843596, 58547.0703125),
    (445.71700965093, 4182.04345703125), (446.1200302053469, 93216.3359375),
    (447.09963627699824, 3806.537109375), (447.1169242266495, 59846.37109375),
    (447.3464079857604, 13170.9541015625), (448.11566395552086, 9294.5107421875),
    (448.3500303628631, 3213.052490234375), (452.1123280000919, 5092.0869140625),
    (461.1934526664677, 4022.537353515625), (462.1463969367603, 99732.5),
    (463.14561508666384, 24247.015625), (464.1433022096936, 20417.041015625),
    (465.1421080732791, 3222.4052734375), (470.1669593722212, 8621.81640625),
    (475.23989190282134, 3369.073974609375), (493.27465300375036, 2725.885986328125),
    (496.0077303201583, 8604.0830078125),
]

# the docs' known answer for that spectrum (docs/source/quick_start.rst:119-136)
QUICK_START_EXPECT = {
    "key": ("BSA_test", "C(37)H(59)N(9)O(16)", 2, (("N", "0.000"),)),
    "score": 0.9606609710868856, "scaling_factor": 40.75802642055527,
    "peaks": ((443.7112735313511, 2517650.0, 1.0, 443.7112648946701, 62091),
              (444.21248374593875, 1156173.75, 0.4459422196277157, 444.2127374486285, 27689),
              (444.71384916266277, 336326.96875, 0.12958327918547244, 444.7142840859656, 8046),
              (445.21533524843596, 58547.0703125, 0.02805309805863953, 445.21582563050043, 1742)),
}

# tests/peptide_sequence_test.py:12-37 (reference's own golden vectors, charge 2)
PEPTIDE_SEQUENCE_KATS = [
    ("TESTPEPTIDE", "C(50)H(79)N(11)O(24)", [52441, 31423, 11803, 3313, 763, 134, 23, 2, 0, 0, 0],
     (609769, 609775)),
    ("TESTPEPTIDE#Oxidation:1", "C(50)H(79)N(11)O(25)", [52314, 31367, 11894, 3374, 787, 140, 24, 3, 0, 0, 0],
     (617767, 617773)),
    ("TESTPEPTIDE#SILAC TMT:0", "C(52)H(99)13C(10)15N(1)N(12)O(26)",
     [12, 2247, 49894, 30980, 12102, 3539, 849, 156, 28, 3, 0, 0, 0], (726859, 726866)),
]

CASES = {
    "c1_basic": dict(lib=dict(molecules=["DDSPDLPK"], charges=[2]), run=dict(n_spectra=6, mean_peaks=120, seed=11)),
    "bsa_14n15n": dict(
        lib=dict(molecules=BSA_MOLECULES, charges=[1, 2, 3, 4], metabolic_labels={"15N": [0, 0.99]},
                 fixed_labels=CAM_FIXED_LABEL),
        run=dict(n_spectra=30, mean_peaks=200, seed=12, elution_sigma=2.0)),
    "bsa_pct": dict(
        lib=dict(molecules=BSA_MOLECULES[:6], charges=[2, 3], fixed_labels=CAM_FIXED_LABEL,
                 metabolic_labels={"15N": [0.0, 0.1, 0.25, 0.5, 0.75, 0.9, 0.99]}),
        run=dict(n_spectra=24, mean_peaks=200, seed=13, elution_sigma=2.0)),
    "multi_label": dict(
        lib=dict(molecules=BSA_MOLECULES[3:8], charges=[2, 3],
                 metabolic_labels={"15N": [0, 0.37, 0.5], "13C": [0, 0.99]}),
        run=dict(n_spectra=16, mean_peaks=150, seed=14, elution_sigma=2.0)),
    "o18": dict(
        lib=dict(molecules=["KLEINERTEST", "+H2O", "PEPTIDE+PO3", "+C6H12O6"], charges=[1, 2],
                 metabolic_labels={"18O": [0, 0.99]}),
        run=dict(n_spectra=20, mean_peaks=150, seed=15, elution_sigma=3.0)),
    "silac_locked": dict(
        lib=dict(molecules=["EILCEWRRAR", "KLEINERTESTR", "AEFVEVTK"], charges=[2, 3],
                 fixed_labels={"R": ["C(-6) 13C(6) N(-4) 15N(4)", ""], "K": ["C(-6) 13C(6) N(-2) 15N(2)", ""]},
                 params={"SILAC_AAS_LOCKED_IN_EXPERIMENT": ["K", "R"]}),
        run=dict(n_spectra=20, mean_peaks=150, seed=16, elution_sigma=3.0)),
    "sulfur_selenium": dict(
        lib=dict(molecules=["MCSMCSMMCC", "+C100H200S30Se2", "+C20H30Cl4Br2", "+C10H20P2F6Na1"], charges=[1, 2, 3, -1]),
        run=dict(n_spectra=20, mean_peaks=150, seed=17, elution_sigma=3.0)),
    "params_var": dict(
        lib=dict(molecules=BSA_MOLECULES[:10], charges=[1, 2, 3], fixed_labels=CAM_FIXED_LABEL,
                 params={"INTERNAL_PRECISION": 10000, "REL_MZ_RANGE": 1e-5, "MACHINE_OFFSET_IN_PPM": 3.0,
                         "MIN_REL_PEAK_INTENSITY_FOR_MATCHING": 0.05, "M_SCORE_THRESHOLD": 0.3,
                         "MAX_MOLECULES_PER_MATCH_BIN": 5, "MZ_SCORE_PERCENTILE": 0.7,
                         "REQUIRED_PERCENTILE_PEAK_OVERLAP": 0.3, "LOWER_MZ_LIMIT": 300, "UPPER_MZ_LIMIT": 1500}),
        run=dict(n_spectra=30, mean_peaks=150, seed=18, elution_sigma=3.0, ppm_sd=3.0, offset_ppm=3.0)),
    "bin1": dict(
        lib=dict(molecules=BSA_MOLECULES[:8], charges=[2], params={"MAX_MOLECULES_PER_MATCH_BIN": 1,
                                                                  "MINIMUM_NUMBER_OF_MATCHED_ISOTOPOLOGUES": 3}),
        run=dict(n_spectra=20, mean_peaks=150, seed=19, elution_sigma=3.0)),
    "random_peptides": dict(
        lib=dict(molecules=random_tryptic_peptides(150, seed=5), charges=[1, 2, 3, 4, 5], fixed_labels=CAM_FIXED_LABEL),
        run=dict(n_spectra=12, mean_peaks=300, seed=20, elution_sigma=1.0, entry_fraction=0.5)),
    # dense spectra: several candidates per window, duplicate 0.001-Da buckets (Q9-Q12)
    "dense": dict(
        lib=dict(molecules=BSA_MOLECULES[:5], charges=[2, 3], fixed_labels=CAM_FIXED_LABEL),
        run=dict(n_spectra=24, mean_peaks=120, seed=21, elution_sigma=3.0, dense=True)),
    # wide windows at high charge: neighbouring isotope windows overlap, one bucket can sit in two windows (Q10)
    "overlap_windows": dict(
        lib=dict(molecules=BSA_MOLECULES[:6], charges=[4, 5], params={"REL_MZ_RANGE": 5e-4, "REL_I_RANGE": 0.5,
                                                                     "M_SCORE_THRESHOLD": 0.2}),
        run=dict(n_spectra=16, mean_peaks=150, seed=23, elution_sigma=3.0, ppm_sd=100.0)),
    # one ~22 kDa molecule: more than 24 matchable peaks per entry (the matcher's long-entry path)
    "huge_formula": dict(
        lib=dict(molecules=["+C980H1540N270O290S8"], charges=[10, 14, 18], params={"UPPER_MZ_LIMIT": 3000}),
        run=dict(n_spectra=8, mean_peaks=150, seed=24, elution_sigma=3.0), skip_trees=True),
    "big_formula": dict(
        lib=dict(molecules=["+C222H349N61O66S2", "+C444H698N122O133S4", "+C130H210N30O45S1"], charges=[3, 5, 8]),
        run=dict(n_spectra=10, mean_peaks=150, seed=22, elution_sigma=3.0)),
    # averagine at 20 kDa and 50 kDa: the range where the reference's 2-isotope levels use exact big-int binomials
    # (IL:1121-1140, 1216-1324) and the kernels FP64 pow x iterated C(n,k).  heavy: the reference needs 13 s / ~10 min
    # and several GB for these, so the CPU suite does not rebuild them with the oracle (QMS_HEAVY_ORACLE=1 does); the
    # fixtures carry minPos/maxPos of every tree level instead of the full trees.
    "avg_20kda": dict(
        lib=dict(molecules=["+C889H1396N244O266S8"], charges=[11, 14, 18]),
        run=dict(n_spectra=8, mean_peaks=150, seed=25, elution_sigma=3.0), skip_trees=True, tree_bounds=True, heavy=True),
    "avg_50kda": dict(
        lib=dict(molecules=["+C2222H3491N611O665S19"], charges=[26, 32]),
        run=dict(n_spectra=8, mean_peaks=150, seed=26, elution_sigma=3.0), skip_trees=True, tree_bounds=True, heavy=True),
    # BASELINE.json configs[2] at its full shape: all 19 BSA molecules x 100 label percentiles, charges 1-4
    "bsa_pct_full": dict(
        lib=dict(molecules=BSA_MOLECULES, charges=[1, 2, 3, 4], fixed_labels=CAM_FIXED_LABEL,
                 metabolic_labels={"15N": [round(0.01 * i, 2) for i in range(100)]}),
        run=dict(n_spectra=20, mean_peaks=300, seed=27, elution_sigma=1.0, entry_fraction=0.004), skip_trees=True, heavy=True),
    # wide windows, ten candidates in every window of a six-window entry: 10^6 candidate combinations for one pair
    # (IL:2004-2034; the reference enumerates and sorts them all -- minutes and ~1 GB for this one spectrum)
    "many_candidates": dict(
        lib=dict(molecules=["LKPDPNTLCDEFK"], charges=[2], params={"REL_MZ_RANGE": 1e-4, "M_SCORE_THRESHOLD": 0.1}),
        run=dict(n_spectra=1, mean_peaks=40, seed=28, elution_sigma=3.0, entry_fraction=0.0, crowd=10), heavy=True),
}
