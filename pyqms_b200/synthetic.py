"""Deterministic synthetic LC-MS runs and peptide lists.

Workload generators for tests and bench.py; not part of the hot path.  A *run* is returned in the
packed layout the batched matcher consumes: ``peaks`` (n_peaks, 2) float64 rows of (m/z, intensity),
sorted by m/z inside each spectrum, and ``offsets`` (n_spectra + 1) int64 row offsets.
"""
import numpy as np

# the 19 BSA molecules of the reference fixture tests/data/test_BSA_pyqms_results.pkl (lookup
# "formula to molecule"), i.e. the 17 peptides of data/BSA1_omssa_2_1_9.mztab + 2 oxidised forms
BSA_MOLECULES = [
    "HLVDEPQNLIK", "LCVLHEK", "GACLLPK", "EYEATLEECCAK", "CCTESLVNR", "YLYEIAR", "LKPDPNTLCDEFK",
    "LKPDPNTLCDEFK#Oxidation:1", "LKPDPNTLCDEFK#Oxidation:1;Oxidation:2", "EACFAVEGPK", "SHCIAEVEK",
    "LVTDLTK", "LVVSTQTALA", "DDSPDLPK", "ETYGDMADCCEK", "YICDNQDTISSK", "ECCDKPLLEK", "DLGEEHFK",
    "AEFVEVTK",
]
CAM_FIXED_LABEL = {"C": ["C(2)H(3)14N(1)O(1)"]}

# natural amino-acid frequencies (UniProtKB/Swiss-Prot release statistics, per cent)
_AA = "ACDEFGHIKLMNPQRSTVWY"
_AA_FREQ = np.array([8.25, 1.38, 5.46, 6.72, 3.86, 7.07, 2.27, 5.91, 5.80, 9.65, 2.41, 4.06, 4.74, 3.93,
                     5.53, 6.64, 5.36, 6.86, 1.10, 2.92])


def random_tryptic_peptides(n, seed=20260104, min_len=7, max_len=30):
    """n distinct peptides, length U[min_len, max_len], natural AA frequencies, C-terminal K/R."""
    rng = np.random.default_rng(seed)
    freq = _AA_FREQ / _AA_FREQ.sum()
    letters = np.frombuffer(_AA.encode(), dtype="S1")
    out, seen = [], set()
    while len(out) < n:
        m = max(1024, (n - len(out)) * 2)
        lens = rng.integers(min_len, max_len + 1, size=m)
        body = rng.choice(len(_AA), size=int(lens.sum()), p=freq)
        ends = rng.integers(0, 2, size=m)
        pos = 0
        for length, end in zip(lens, ends):
            seq = letters[body[pos:pos + length - 1]].tobytes().decode() + "KR"[end]
            pos += length
            if seq not in seen:
                seen.add(seq)
                out.append(seq)
                if len(out) == n:
                    break
    return out


def averagine_formulas(n, seed=20260105, min_mass=100.0, max_mass=50000.0):
    """n '+C..H..N..O..S..' formula strings, mass log-uniform, averagine +-10 % per-element jitter."""
    rng = np.random.default_rng(seed)
    mass = np.exp(rng.uniform(np.log(min_mass), np.log(max_mass), size=n))
    per = np.array([4.9384, 7.7583, 1.3577, 1.4773, 0.0417]) / 111.1254
    counts = np.rint(mass[:, None] * per[None, :] * rng.uniform(0.9, 1.1, size=(n, 5))).astype(int)
    counts[:, 0] = np.maximum(counts[:, 0], 1)
    counts[:, 1] = np.maximum(counts[:, 1], 1)
    out = []
    for c, h, nn, o, s in counts:
        f = "+C%dH%d" % (c, h)
        if nn:
            f += "N%d" % nn
        if o:
            f += "O%d" % o
        if s:
            f += "S%d" % s
        out.append(f)
    return out


def synth_run(entry_mz, entry_ci, n_spectra, seed, mean_peaks=800, sigma_peaks=0.5, min_peaks=100,
              max_peaks=5000, entry_fraction=1.0, elution_sigma=15.0, mz_lo=150.0, mz_hi=2000.0,
              ppm_sd=1.5, intensity_sd=0.08, dropout=0.10, planted=None, return_planted=False):
    """Synthetic run: uniform background + Gaussian-eluting library entries.

    entry_mz / entry_ci: per library index entry, 1-D arrays of the theoretical c-peak m/z and the
    integer abundances (what a matched peak's ``cmz`` and ``ci`` are).  Every chosen entry elutes
    once (apex uniform over the run, sd ``elution_sigma`` spectra, on for +-3 sd); while on, each
    c-peak is planted at mz*(1+N(0, ppm_sd ppm)) with intensity ci*amount*(1+N(0, intensity_sd)),
    ``dropout`` of them missing.  Returns (peaks (n,2) float64, offsets (n_spectra+1) int64).

    ``planted``: the c-peak table of the chosen entries from an earlier call with ``return_planted=True``
    (dict n_entries, lens, mz, ci); with it the identical run is drawn without a library (entry_mz / entry_ci
    may be None) -- bench.py's reference arm reproduces the GPU arm's spectra from a committed table this way.
    """
    rng = np.random.default_rng(seed)
    n_bg = np.clip(np.rint(rng.lognormal(np.log(mean_peaks), sigma_peaks, size=n_spectra)), min_peaks,
                   max_peaks).astype(np.int64)
    bg_spec = np.repeat(np.arange(n_spectra, dtype=np.int64), n_bg)
    bg_mz = rng.uniform(mz_lo, mz_hi, size=bg_spec.size)
    bg_i = 10.0 ** rng.uniform(3.0, 6.0, size=bg_spec.size)

    n_entries = int(planted["n_entries"]) if planted is not None else len(entry_mz)
    chosen = np.arange(n_entries)
    if entry_fraction < 1.0:
        chosen = np.sort(rng.choice(n_entries, size=max(1, int(round(n_entries * entry_fraction))),
                                    replace=False))
    spec_parts, mz_parts, i_parts = [bg_spec], [bg_mz], [bg_i]
    if len(chosen):
        apex = rng.uniform(0, n_spectra, size=len(chosen))
        amount = 10.0 ** rng.uniform(1.0, 3.0, size=len(chosen))
        half = int(np.ceil(3 * elution_sigma))
        if planted is not None:
            lens = np.asarray(planted["lens"], dtype=np.int64)
            flat_mz = np.asarray(planted["mz"], dtype=np.float64)
            flat_ci = np.asarray(planted["ci"], dtype=np.float64)
            assert len(lens) == len(chosen), "planted table does not belong to this (seed, n_spectra, entry_fraction)"
        else:
            lens = np.array([len(entry_mz[e]) for e in chosen], dtype=np.int64)
            flat_mz = np.concatenate([np.asarray(entry_mz[e], dtype=np.float64) for e in chosen])
            flat_ci = np.concatenate([np.asarray(entry_ci[e], dtype=np.float64) for e in chosen])
        table = {"n_entries": n_entries, "lens": lens, "mz": flat_mz, "ci": flat_ci}
        owner = np.repeat(np.arange(len(chosen)), lens)           # per c-peak: which chosen entry
        shifts = np.arange(-half, half + 1)
        # (c-peak, shift) grid
        spec = np.rint(apex[owner])[:, None].astype(np.int64) + shifts[None, :]
        profile = np.exp(-0.5 * ((spec - apex[owner][:, None]) / elution_sigma) ** 2)
        ok = (spec >= 0) & (spec < n_spectra) & (rng.random(spec.shape) >= dropout)
        mz = flat_mz[:, None] * (1.0 + rng.normal(0.0, ppm_sd * 1e-6, size=spec.shape))
        inten = flat_ci[:, None] * (amount[owner][:, None] * profile) * \
            (1.0 + rng.normal(0.0, intensity_sd, size=spec.shape))
        ok &= inten > 0
        spec_parts.append(spec[ok])
        mz_parts.append(mz[ok])
        i_parts.append(inten[ok])
    spec = np.concatenate(spec_parts)
    mz = np.concatenate(mz_parts)
    inten = np.concatenate(i_parts)
    order = np.lexsort((mz, spec))
    peaks = np.empty((order.size, 2), dtype=np.float64)
    peaks[:, 0] = mz[order]
    peaks[:, 1] = inten[order]
    offsets = np.zeros(n_spectra + 1, dtype=np.int64)
    np.cumsum(np.bincount(spec, minlength=n_spectra), out=offsets[1:])
    if return_planted:
        return peaks, offsets, (table if len(chosen) else {"n_entries": n_entries, "lens": [], "mz": [], "ci": []})
    return peaks, offsets


def split_run(peaks, offsets):
    """List of per-spectrum (n,2) views -- what a per-spectrum ``match_all`` loop is fed."""
    return [peaks[offsets[i]:offsets[i + 1]] for i in range(len(offsets) - 1)]
