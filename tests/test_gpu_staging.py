"""Large pageable host buffers go through the library's own staged copies (csrc/ctx.cu: two pinned slots, DMA and the
multi-threaded host copy overlapped).  The data must be what the direct path delivers, also across the 16 MB chunk
boundaries."""
import numpy as np
import pytest

import pyqms_b200
from pyqms_b200.synthetic import BSA_MOLECULES, CAM_FIXED_LABEL, synth_run

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    return pyqms_b200.IsotopologueLibrary(molecules=BSA_MOLECULES, charges=[1, 2, 3, 4], metabolic_labels={"15N": [0, 0.99]},
                                         fixed_labels=CAM_FIXED_LABEL, verbose=False)


def test_staged_download_equals_direct_download(lib):
    n = 10_000_000                                      # 3 x 40 MB of int32 results: three staging chunks each
    mz = np.linspace(150.0, 2000.0, n)
    lo, hi, tmz = lib._transform_values(mz)
    chunk = (16 << 20) // 4                             # elements per staging chunk
    for start in (0, chunk - 500, 2 * chunk - 500, n - 1000):
        part = slice(start, start + 1000)
        lo_s, hi_s, tmz_s = lib._transform_values(mz[part])        # 12 KB: direct path
        assert np.array_equal(lo[part], lo_s) and np.array_equal(hi[part], hi_s) and np.array_equal(tmz[part], tmz_s)
    assert np.array_equal(tmz, np.rint(mz * lib.params["INTERNAL_PRECISION"]).astype(np.int32))
    assert bool(np.all(lo <= tmz)) and bool(np.all(tmz <= hi))


def test_staged_upload_equals_direct_upload(lib):
    peaks, offsets = synth_run(*lib.entry_peaks(), n_spectra=3000, seed=31)
    assert peaks.nbytes > 2 * (16 << 20)                # pageable numpy array: staged in three chunks
    whole = lib.match_packed(peaks, offsets)
    assert whole["n_hits"] > 500
    pieces = []
    for a in range(0, 3000, 250):                       # 250 spectra = ~3.6 MB per call: direct path
        sub = offsets[a:a + 251]
        part = lib.match_packed(np.ascontiguousarray(peaks[sub[0]:sub[-1]]), sub - sub[0])
        part["spec"] = part["spec"] + a
        part["peak"] = np.where(part["peak"] >= 0, part["peak"] + sub[0], -1)
        pieces.append(part)
    for field in ("spec", "entry", "score", "scaling", "peak"):
        assert np.array_equal(np.concatenate([p[field] for p in pieces]), whole[field]), field
    assert np.array_equal(np.concatenate([p["mmz"] for p in pieces]), whole["mmz"], equal_nan=True)
    assert np.array_equal(np.concatenate([p["mi"] for p in pieces]), whole["mi"], equal_nan=True)
