"""Default parameters of the library build and the matcher.

Same keys and default values as the reference's ``pyqms.params`` (pyqms/params.py:29-74); the dict is
deep-copied and updated with the user's ``params`` by ``IsotopologueLibrary`` (IL:228-230).  Only the
numeric subset crosses the C ABI (``qms_params`` in include/qms_b200.h).
"""

params = dict(
    PERCENTILE_FORMAT_STRING="{0:.3f}",        # label percentile -> dictionary key
    M_SCORE_THRESHOLD=0.5,                     # matches below are dropped (IL:1860)
    ELEMENT_MIN_ABUNDANCE=1e-3,                # pruning of the per-element isotope patterns
    MIN_REL_PEAK_INTENSITY_FOR_MATCHING=0.01,  # relative abundance that makes a peak matchable
    REQUIRED_PERCENTILE_PEAK_OVERLAP=0.50,     # fraction of matchable peaks that must be found
    MINIMUM_NUMBER_OF_MATCHED_ISOTOPOLOGUES=2,
    INTENSITY_TRANSFORMATION_FACTOR=1e5,       # abundance -> integer "abun"
    UPPER_MZ_LIMIT=2000,
    LOWER_MZ_LIMIT=150,
    MZ_TRANSFORMATION_FACTOR=10000,
    REL_MZ_RANGE=5e-6,                         # m/z tolerance (relative): +-5 ppm
    REL_I_RANGE=0.2,                           # intensity tolerance at relative abundance 1
    INTERNAL_PRECISION=1000,                   # integer m/z grid: 1/1000 Da
    MAX_MOLECULES_PER_MATCH_BIN=20,
    MZ_SCORE_PERCENTILE=0.4,                   # weight of the m/z part of the mScore
    SILAC_AAS_LOCKED_IN_EXPERIMENT=None,
    BUILD_RESULT_INDEX=True,
    MACHINE_OFFSET_IN_PPM=0.0,
    FIXED_LABEL_ISOTOPE_ENRICHMENT_LEVELS={"15N": 0.994, "13C": 0.996, "2H": 0.994},
    COLORS={0.0: (37, 37, 37), 0.1: (99, 99, 99), 0.2: (150, 150, 150), 0.3: (204, 204, 204),
            0.4: (247, 247, 247), 0.5: (203, 27, 29), 0.6: (248, 120, 72), 0.7: (253, 219, 121),
            0.8: (209, 239, 121), 0.9: (129, 202, 78), 1: (27, 137, 62)},
)
