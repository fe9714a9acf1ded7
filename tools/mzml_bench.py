#!/usr/bin/env python
"""Ingest throughput of the mzML readers on this host (no GPU needed): native libqms_mzml.so per thread count against the
ElementTree reader, on a synthetic run written with float and MS-Numpress array encodings.

    python tools/mzml_bench.py [n_spectra=3000] [peaks_per_spectrum=1000]
"""
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pyqms_b200.mzml import read_mzml, write_mzml  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    rng = np.random.default_rng(1)
    spectra = [np.stack([np.sort(rng.uniform(150, 2000, k)), 10 ** rng.uniform(3, 6, k)], axis=1) for _ in range(n)]
    folder = tempfile.mkdtemp()
    print("cores %d" % os.cpu_count())
    for precision, compress, packing in ((64, True, None), (64, False, None), (32, True, None), (64, True, ("linear", "slof")),
                                         (64, False, ("linear", "pic"))):
        path = os.path.join(folder, "run.mzML")
        if packing:                                           # MS-Numpress: counts, not floats, in the intensity column
            for spectrum in spectra:
                spectrum[:, 1] = np.floor(spectrum[:, 1])
        write_mzml(path, spectra, precision=precision, compress=compress, indexed=True,
                   mz_packing=packing and packing[0], intensity_packing=packing and packing[1])
        size = os.path.getsize(path) / 1e6
        t0 = time.perf_counter()
        want = read_mzml(path, native=False)
        slow = time.perf_counter() - t0
        line = "%-12s %-5s %6.1f MB  ElementTree %.2f s" % ("+".join(packing) if packing else "%d-bit" % precision, "zlib" if compress else "plain",
                                                        size, slow)
        for threads in (1, 2, 4, 8, 16):
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                got = read_mzml(path, native=True, threads=threads)
                best = min(best, time.perf_counter() - t0)
            assert np.array_equal(got.peaks, want.peaks)
            line += " | %2d thr %.3f s (%.0f spectra/s)" % (threads, best, n / best)
        print(line, flush=True)


if __name__ == "__main__":
    main()
