"""Device amounts (qms_calc_amounts) behind the reference's quant-summary flow, against amounts the unmodified
reference computed with pyqms.adaptors.calc_amount_function (tests/golden/make_golden_adaptors.py)."""
from copy import deepcopy

import pytest

from pyqms_b200 import adaptors
from tests.test_cpu_adaptors import GOLDEN, build_results

pytestmark = pytest.mark.gpu
FIELDS = ("max I in window", "max I in window (rt)", "max I in window (score)", "sum I in window", "auc in window")


@pytest.mark.parametrize("tolerance", sorted(GOLDEN["rt_info"]["runs"]))
@pytest.mark.parametrize("mode", ["on_device", "function"])
def test_amounts_match_reference(tolerance, mode):
    want = GOLDEN["rt_info"]["runs"][tolerance]["amounts_full"]
    results = build_results()
    lines = deepcopy(GOLDEN["rt_info"]["runs"][tolerance]["lines"])
    for line in lines:      # JSON turned the percentile tuples into lists
        line["label_percentiles"] = tuple(tuple(pair) for pair in line["label_percentiles"])
    if mode == "on_device":
        got = results.calc_amounts_from_rt_info_file(buffered_csv_dicts=lines, buffer_only=True, on_device=True)
    else:
        got = results.calc_amounts_from_rt_info_file(buffered_csv_dicts=lines, buffer_only=True,
                                                     calc_amount_function=adaptors.calc_amount_function)
    assert len(got) == len(want)
    n_amounts = 0
    for line, ref in zip(got, want):
        assert line["formula"] == ref["formula"] and line["molecule"] == ref["molecule"]
        for field in FIELDS:
            assert (field in line) == (field in ref), (field, line, ref)
            if field in ref:
                n_amounts += 1
                assert line[field] == pytest.approx(ref[field], rel=1e-12, abs=1e-6)
    assert n_amounts >= 5 * (2 if tolerance == "None" else 10)


def test_calc_amount_function_edge_profiles():
    assert adaptors.calc_amount_function({"rt": [], "i": [], "scores": []}) is None
    one = adaptors.calc_amount_function({"rt": [3.0], "i": [7.0], "scores": [0.9]})
    assert one == {"max I in window": 7.0, "max I in window (rt)": 3.0, "max I in window (score)": 0.9,
                   "sum I in window": 7.0, "auc in window": 0.0}
    ramp = adaptors.calc_amount_function({"rt": [0.0, 1.0, 2.0, 4.0], "i": [0.0, 2.0, 2.0, 0.0], "scores": [0.1, 0.2, 0.3, 0.4]})
    assert ramp["max I in window (rt)"] == 1.0 and ramp["max I in window (score)"] == 0.2      # first maximum
    assert ramp["auc in window"] == pytest.approx(1.0 + 2.0 + 2.0) and ramp["sum I in window"] == 4.0
