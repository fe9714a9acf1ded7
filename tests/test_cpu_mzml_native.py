"""libqms_mzml.so (include/qms_mzml.h): the native mzML reader must return exactly what the ElementTree reader returns
(N3), for every encoding the writer produces and for hand-written files with the variations converters produce."""
import base64
import ctypes
import os
import re
import struct
import zlib
from pathlib import Path

import numpy as np
import pytest

from pyqms_b200 import mzml
from pyqms_b200.mzml import read_mzml, write_mzml

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module", autouse=True)
def built_library():
    if not (ROOT / "pyqms_b200" / "libqms_mzml.so").exists():     # fresh checkout: make builds it with the CUDA library
        import sys
        sys.path.insert(0, str(ROOT))
        import __graft_entry__
        __graft_entry__.build()
        mzml._NATIVE = None


def same_run(a, b):
    return (np.array_equal(a.peaks, b.peaks) and np.array_equal(a.offsets, b.offsets) and a.spec_ids == b.spec_ids and
            a.rts == b.rts and a.ms_levels == b.ms_levels and a.name == b.name)


def test_native_library_exports_the_declared_symbols():
    header = (ROOT / "include" / "qms_mzml.h").read_text()
    declared = sorted(set(re.findall(r"\b(qms_(?:mzml|numpress)_[a-z_]+)\s*\(", re.sub(r"/\*.*?\*/", "", header, flags=re.S))))
    assert len(declared) == 7
    lib = ctypes.CDLL(str(ROOT / "pyqms_b200" / "libqms_mzml.so"))
    for name in declared:
        assert hasattr(lib, name), name


@pytest.mark.parametrize("precision,compress", [(64, True), (64, False), (32, True), (32, False)])
def test_native_reader_equals_elementtree_reader(tmp_path, precision, compress):
    rng = np.random.default_rng(precision + compress)
    spectra, levels = [], []
    for s in range(40):
        k = int(rng.integers(0, 400)) if s != 5 else 0          # an empty spectrum, too
        mz, inten = np.sort(rng.uniform(100, 2000, k)), 10 ** rng.uniform(2, 7, k)
        if s % 6 == 0 and k > 3:                                  # unsorted peaks: both readers reorder stably
            mz[[0, -1]] = mz[[-1, 0]]
            mz[k // 2] = mz[k // 2 - 1]
        spectra.append(np.stack([mz, inten], axis=1))
        levels.append(2 if s % 4 == 3 else 1)
    path = tmp_path / "run.mzML"
    write_mzml(path, spectra, rts_minutes=[0.25 * s for s in range(40)], ms_levels=levels, precision=precision, compress=compress)
    for level in (1, 2, None):
        for sort_peaks in (True, False):
            slow = read_mzml(path, ms_level=level, sort_peaks=sort_peaks, native=False)
            for threads in (1, 3):
                fast = read_mzml(path, ms_level=level, sort_peaks=sort_peaks, native=True, threads=threads)
                assert same_run(slow, fast), (level, sort_peaks, threads)
    assert len(read_mzml(path, ms_level=1).spec_ids) == 30 and same_run(read_mzml(path), read_mzml(path, native=False))


def _array(values, accession, dtype="<f8", compressed=False, wrap=False, group=None):
    raw = np.asarray(values, dtype=dtype).tobytes()
    if compressed:
        raw = zlib.compress(raw)
    text = base64.b64encode(raw).decode()
    if wrap:                                                     # line breaks inside the payload
        text = "\n".join(text[i:i + 60] for i in range(0, len(text), 60)) + "\n"
    width = "MS:1000523" if dtype == "<f8" else "MS:1000521"
    packing = "MS:1000574" if compressed else "MS:1000576"
    params = ('<referenceableParamGroupRef ref="%s"/>' % group) if group else (
        '<cvParam cvRef="MS" accession="%s" name="float"/><cvParam cvRef="MS" accession="%s" name="packing"/>' % (width, packing))
    return ('<binaryDataArray encodedLength="%d">%s<cvParam cvRef="MS" accession="%s" name="array"/><binary>%s</binary>'
            '</binaryDataArray>' % (len(text), params, accession, text))


def _spectrum(index, native_id, level_xml, rt_xml, mz, inten, **kw):
    return ('<spectrum index="%d" id="%s" defaultArrayLength="%d">%s<scanList count="1"><scan>%s</scan></scanList>'
            '<binaryDataArrayList count="2">%s%s</binaryDataArrayList></spectrum>' % (
                index, native_id, len(mz), level_xml, rt_xml, _array(mz, "MS:1000514", **kw), _array(inten, "MS:1000515", **kw)))


def test_converter_style_variations(tmp_path):
    """Param groups, namespace prefixes, comments, '>' inside attribute values, wrapped base64, seconds, missing times."""
    level1 = '<cvParam cvRef="MS" accession="MS:1000511" name="ms level" value="1"/>'
    body = "".join([
        _spectrum(0, "controllerType=0 controllerNumber=1 scan=17", level1,
                  '<cvParam cvRef="MS" accession="MS:1000016" name="scan start time" value="93.5" unitName="second"/>',
                  [400.5, 300.25, 500.125], [10.0, 20.0, 30.0], wrap=True),
        "<!-- a comment with <spectrum> inside -->",
        _spectrum(1, "index=7", '<referenceableParamGroupRef ref="ms1"/>',
                  '<userParam name="filter" value="FTMS + p NSI Full ms [300.00-2000.00] a>b"/>'
                  '<cvParam cvRef="MS" accession="MS:1000016" name="scan start time" value="1.75" unitName="minute"/>',
                  [200.0, 201.0], [5.0, 6.0], compressed=True, group="zlib64"),
        _spectrum(2, "no number here", level1, "", [100.0], [1.0], dtype="<f4"),
        _spectrum(3, "scan=9", '<cvParam cvRef="MS" accession="MS:1000511" name="ms level" value="2"/>', "", [50.0, 60.0], [1.0, 2.0]),
    ])
    text = ('<?xml version="1.0" encoding="utf-8"?>\n<mzML xmlns="http://psi.hupo.org/ms/mzml"><referenceableParamGroupList count="2">'
            '<referenceableParamGroup id="ms1"><cvParam cvRef="MS" accession="MS:1000511" name="ms level" value="1"/>'
            '</referenceableParamGroup><referenceableParamGroup id="zlib64"><cvParam cvRef="MS" accession="MS:1000523" name="64-bit float"/>'
            '<cvParam cvRef="MS" accession="MS:1000574" name="zlib compression"/></referenceableParamGroup></referenceableParamGroupList>'
            '<run id="r"><spectrumList count="4">%s</spectrumList><chromatogramList count="1"><chromatogram index="0" id="TIC" '
            'defaultArrayLength="2"><binaryDataArrayList count="1">%s</binaryDataArrayList></chromatogram></chromatogramList></run></mzML>'
            % (body, _array([1.0, 2.0], "MS:1000595")))
    path = tmp_path / "variations.mzML"
    path.write_text(text)
    run = read_mzml(path, ms_level=1, native=True)
    assert run.spec_ids == [17, 7, 3] and run.ms_levels == [1, 1, 1]
    assert run.rts == [(93.5, "second"), (1.75, "minute"), None]
    assert run.offsets.tolist() == [0, 3, 5, 6]
    assert run.peaks.tolist() == [[300.25, 20.0], [400.5, 10.0], [500.125, 30.0], [200.0, 5.0], [201.0, 6.0], [100.0, 1.0]]
    everything = read_mzml(path, ms_level=None, native=True)
    assert everything.spec_ids == [17, 7, 3, 9] and everything.ms_levels == [1, 1, 1, 2] and len(everything.peaks) == 8
    # the ElementTree reader does not resolve param groups; on the spectra that do not use them both agree
    slow = read_mzml(path, ms_level=2, native=False)
    assert same_run(slow, read_mzml(path, ms_level=2, native=True))


def test_unsupported_encodings_fall_back_and_broken_files_raise(tmp_path):
    level1 = '<cvParam cvRef="MS" accession="MS:1000511" name="ms level" value="1"/>'
    truncated = _spectrum(0, "scan=1", level1, "", [100.0, 200.0], [1.0, 2.0]).replace(
        '<cvParam cvRef="MS" accession="MS:1000514"', '<cvParam cvRef="MS" accession="MS:1003089" name="truncation and zlib compression"/><cvParam cvRef="MS" accession="MS:1000514"')
    path = tmp_path / "truncation.mzML"
    path.write_text("<mzML><run><spectrumList>%s</spectrumList></run></mzML>" % truncated)
    with pytest.raises(ValueError, match="does not decode"):
        read_mzml(path, native=True)
    for native in (None, False):                                 # no reader here decodes it: an error, never made-up numbers
        with pytest.raises(ValueError, match="not decoded"):
            read_mzml(path, native=native)
    broken = _spectrum(0, "scan=1", level1, "", [100.0, 200.0, 300.0], [1.0, 2.0, 3.0], compressed=True)
    payload = re.search(r"<binary>([^<]+)</binary>", broken).group(1)
    path = tmp_path / "broken.mzML"
    path.write_text("<mzML><run><spectrumList>%s</spectrumList></run></mzML>" % broken.replace(payload, payload[:12], 1))
    with pytest.raises(ValueError, match="broken"):
        read_mzml(path, native=True)
    with pytest.raises(OSError):
        read_mzml(tmp_path / "missing.mzML", native=True)
    assert struct.calcsize("d") == 8 and mzml._native() is not None


def test_indexed_files_are_scanned_in_parallel_blocks(tmp_path):
    """indexedmzML: the spectrum offsets of the file's index split the tag scan over the threads; a wrong index is ignored."""
    rng = np.random.default_rng(8)
    spectra = [np.stack([np.sort(rng.uniform(100, 2000, k)), 10 ** rng.uniform(2, 7, k)], axis=1)
               for k in rng.integers(1, 60, size=300)]
    levels = [1 if s % 5 else 2 for s in range(300)]
    path = tmp_path / "indexed.mzML"
    write_mzml(path, spectra, ms_levels=levels, indexed=True)
    plain = tmp_path / "plain.mzML"
    write_mzml(plain, spectra, ms_levels=levels, indexed=False)
    want = read_mzml(plain, ms_level=None, native=False)
    for threads in (1, 2, 5, 16):
        lib = mzml._native()
        handle = ctypes.c_void_p()
        assert lib.qms_mzml_open(str(path).encode(), ctypes.byref(handle)) == 0
        lib.qms_mzml_close(handle)
        got = read_mzml(path, ms_level=None, native=True, threads=threads)
        assert np.array_equal(got.peaks, want.peaks) and np.array_equal(got.offsets, want.offsets)
        assert got.spec_ids == want.spec_ids and got.rts == want.rts and got.ms_levels == want.ms_levels
        assert mzml.LAST_READ["scan_blocks"] > 1 or (os.cpu_count() or 1) < 4     # the index WAS used (half the cores, at least 2)
    assert same_run(read_mzml(path, ms_level=1, native=False)._replace(name="x"), read_mzml(path, ms_level=1, native=True)._replace(name="x"))
    # an index whose offsets do not point at <spectrum> tags (file edited after indexing) is not used
    text = path.read_text()
    shifted = re.sub(r"<offset ([^>]*)>(\d+)</offset>", lambda m: "<offset %s>%d</offset>" % (m.group(1), int(m.group(2)) + 3), text)
    assert shifted != text
    bad = tmp_path / "bad_index.mzML"
    bad.write_text(shifted)
    got = read_mzml(bad, ms_level=None, native=True)
    assert mzml.LAST_READ["scan_blocks"] == 0                                      # untrusted index: sequential scan
    assert np.array_equal(got.peaks, want.peaks) and got.spec_ids == want.spec_ids
    # an index that lists fewer spectra than the file holds (count mismatch inside a block) falls back as well
    fewer = re.sub(r'<offset idRef="[^"]*scan=(?:7|8|9)">\d+</offset>\n', "", text)
    assert fewer != text
    short = tmp_path / "short_index.mzML"
    short.write_text(fewer)
    got = read_mzml(short, ms_level=None, native=True)
    assert mzml.LAST_READ["scan_blocks"] == 0                                      # block counts disagree with the index: sequential scan
    assert np.array_equal(got.peaks, want.peaks) and got.spec_ids == want.spec_ids
    # attribute order and quote character of the <index> element are the writer's business
    quoted = tmp_path / "quoted_index.mzML"
    assert '<index name="spectrum">' in text
    quoted.write_text(text.replace('<index name="spectrum">', "<index  name='spectrum' >"))
    got = read_mzml(quoted, ms_level=None, native=True, threads=4)
    assert mzml.LAST_READ["scan_blocks"] > 1 or (os.cpu_count() or 1) < 4
    assert np.array_equal(got.peaks, want.peaks) and got.spec_ids == want.spec_ids


def test_gzip_compressed_files(tmp_path):
    import gzip
    rng = np.random.default_rng(3)
    spectra = [np.stack([np.sort(rng.uniform(100, 2000, k)), 10 ** rng.uniform(2, 7, k)], axis=1) for k in rng.integers(0, 200, size=80)]
    plain = tmp_path / "run.mzML"
    write_mzml(plain, spectra, indexed=True)
    packed = tmp_path / "run.mzML.gz"
    packed.write_bytes(gzip.compress(plain.read_bytes()))
    want = read_mzml(plain, native=False)
    for native in (True, False):
        got = read_mzml(packed, native=native)
        assert np.array_equal(got.peaks, want.peaks) and np.array_equal(got.offsets, want.offsets)
        assert got.spec_ids == want.spec_ids and got.rts == want.rts and got.name == "run.mzML.gz"
    # several gzip members in one file (bgzip, pigz, cat a.gz b.gz): all of them are read, like gzip.open does
    raw = plain.read_bytes()
    members = tmp_path / "members.mzML.gz"
    members.write_bytes(gzip.compress(raw[:len(raw) // 3]) + gzip.compress(raw[len(raw) // 3:]))
    for native in (True, False):
        got = read_mzml(members, native=native)
        assert np.array_equal(got.peaks, want.peaks) and got.spec_ids == want.spec_ids
    broken = tmp_path / "cut.mzML.gz"
    broken.write_bytes(packed.read_bytes()[:2000])
    with pytest.raises(OSError):
        read_mzml(broken, native=True)


def _check_against(path, want, native, threads=None):
    for level in (None, 1, 2):
        run = read_mzml(path, ms_level=level, native=native, threads=threads)
        keep = [i for i, l in enumerate(want["levels"]) if level is None or l == level]
        peaks = np.concatenate([want["peaks"][want["offsets"][i]:want["offsets"][i + 1]] for i in keep])
        assert np.array_equal(run.peaks, peaks) and run.spec_ids == [int(want["ids"][i]) for i in keep]
        assert np.array_equal(np.diff(run.offsets), np.diff(want["offsets"])[keep])
        assert run.rts == [(float(want["rts"][i]), "minute") for i in keep] and run.ms_levels == [int(want["levels"][i]) for i in keep]


def test_converter_style_document_not_written_by_our_writer(tmp_path):
    """An indexedmzML in the layout msconvert writes (cvList, param groups, pretty printing, unit attributes, precursor
    lists, a chromatogram list and a second index, mixed 32/64-bit and zlib/plain arrays, an empty spectrum), assembled by
    tests/data/make_pwiz_style.py with struct / zlib / base64 only -- both readers, all MS-level filters."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_pwiz_style", ROOT / "tests" / "data" / "make_pwiz_style.py")
    maker = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(maker)
    want = np.load(ROOT / "tests" / "data" / "pwiz_style_small_expected.npz")
    raw, again = maker.document(9)
    assert raw == (ROOT / "tests" / "data" / "pwiz_style_small.mzML").read_bytes()      # the committed file is what the script writes
    for native in (True, False):
        _check_against(ROOT / "tests" / "data" / "pwiz_style_small.mzML", want, native)
    # 450 spectra: enough for the native reader to split the tag scan at the file's own spectrum index
    big, expected = maker.document(450)
    path = tmp_path / "pwiz_style_big.mzML"
    path.write_bytes(big)
    _check_against(path, expected, True, threads=4)
    assert mzml.LAST_READ["scan_blocks"] > 1 or (os.cpu_count() or 1) < 4
    _check_against(path, expected, False)


# ---- MS-Numpress and integer arrays ---------------------------------------------------------------------------------
def _native_numpress(kind, data):
    lib = mzml._native()
    count = ctypes.c_int64()
    code = ("linear", "pic", "slof").index(kind)
    if lib.qms_numpress_decode(code, data, len(data), None, 0, ctypes.byref(count)) != 0:
        raise ValueError(lib.qms_mzml_last_error().decode())
    out = np.empty(count.value)
    assert lib.qms_numpress_decode(code, data, len(data), out.ctypes.data, len(out), ctypes.byref(count)) == 0
    return out


def test_numpress_half_byte_code_and_hand_derived_stream():
    """The half-byte integer code on hand-derived cases, and one linear stream worked out by hand:
    [100, 200, 300.00005, 400.0001] at fixed point 1e5 -> integers 1e7, 2e7, 30000005, 40000010; residuals against the
    linear extrapolation: 5 -> half bytes (7 leading zeros, 5), 0 -> (8); packed 0x75, 0x80 (padded)."""
    from pyqms_b200 import numpress
    cases = {0: [8], 5: [7, 5], 16: [6, 0, 1], -1: [15, 15], -2: [15, 14], -17: [14, 15, 14],
             0x12345678: [0, 8, 7, 6, 5, 4, 3, 2, 1], -0x12345678: [0, 8, 8, 9, 10, 11, 12, 13, 14], 0x0FFFFFFF: [1, 15, 15, 15, 15, 15, 15, 15]}
    for value, expected in cases.items():
        half_bytes = []
        numpress._int_to_half_bytes(value, half_bytes)
        assert half_bytes == expected, value
        reader = numpress._HalfByteReader(numpress._pack_half_bytes(half_bytes), 0)
        assert reader.integer() == value and reader.at_end()
    stream = numpress.encode_linear([100.0, 200.0, 300.00005, 400.0001], 100000.0)
    assert stream == struct.pack(">d", 100000.0) + struct.pack("<II", 10000000, 20000000) + bytes([0x75, 0x80])
    for decode in (numpress.decode_linear, lambda data: _native_numpress("linear", data)):
        assert np.allclose(decode(stream), [100.0, 200.0, 300.00005, 400.0001], rtol=0, atol=1e-9)
        assert len(decode(stream[:8])) == 0 and decode(stream[:12]).tolist() == [100.0] and decode(stream[:16]).tolist() == [100.0, 200.0]
        with pytest.raises(ValueError):
            decode(stream[:10])
        with pytest.raises(ValueError):
            decode(stream[:5])
    assert numpress.decode_slof(struct.pack(">d", 1000.0) + struct.pack("<HH", 0, 6908)).tolist() == [0.0, np.exp(6.908) - 1.0]
    assert numpress.decode_pic(bytes([0x88, 0x80])).tolist() == [0.0, 0.0, 0.0] and numpress.decode_pic(bytes([0x88])).tolist() == [0.0, 0.0]


def test_native_numpress_decoder_equals_the_python_one():
    from pyqms_b200 import numpress
    rng = np.random.default_rng(21)
    for trial in range(60):
        n = int(rng.integers(0, 400))
        mz = np.sort(rng.uniform(100.0, 2500.0, n))
        if trial % 3 == 0:                                        # profile-like: jumps make large and negative residuals
            mz = np.cumsum(rng.choice([0.001, 0.5, 40.0], size=n)) + 100.0
        intensity = np.floor(10.0 ** rng.uniform(0.0, 9.0, n)) * (rng.uniform(size=n) > 0.1)
        fixed = None if trial % 2 else float(rng.choice([1e3, 1e5, 1234567.0]))
        for kind, data in (("linear", numpress.encode_linear(mz, fixed)), ("pic", numpress.encode_pic(intensity)),
                           ("slof", numpress.encode_slof(intensity))):
            slow, fast = numpress.DECODERS[kind](data), _native_numpress(kind, data)
            assert np.array_equal(slow, fast), (trial, kind)
            assert len(slow) == n
        if n:
            assert np.abs(numpress.decode_linear(numpress.encode_linear(mz)) - mz).max() < 5e-6 * max(1.0, 1.0)
            assert np.array_equal(numpress.decode_pic(numpress.encode_pic(intensity)), intensity)
            back = numpress.decode_slof(numpress.encode_slof(intensity))
            assert np.all(np.abs(back - intensity) <= 5e-4 * (intensity + 1.0))
    # a residual that needs all eight half bytes, and values beyond 32 bits once the extrapolation has run
    wild = np.array([1.0, 2.0, 30000.0, 5.0, 4.0e6, 4.0e6 + 1.0])
    data = numpress.encode_linear(wild, 500.0)
    assert np.array_equal(numpress.decode_linear(data), _native_numpress("linear", data)) and np.allclose(numpress.decode_linear(data), wild)


@pytest.mark.parametrize("mz_packing,intensity_packing,compress", [("linear", "slof", True), ("linear", "pic", False), ("linear", "slof", False),
                                                                  (None, "int32", True), ("linear", "int64", False), (None, "pic", True)])
def test_numpress_and_integer_arrays_through_both_readers(tmp_path, mz_packing, intensity_packing, compress):
    rng = np.random.default_rng(5)
    spectra = []
    for n in range(30):
        k = int(rng.integers(0, 300)) if n not in (3, 4, 5) else n - 3          # empty, one and two rows too
        spectra.append(np.column_stack([np.sort(rng.uniform(150.0, 2000.0, k)), np.floor(10.0 ** rng.uniform(2.0, 7.0, k))]))
    path = tmp_path / "packed.mzML"
    write_mzml(path, spectra, mz_packing=mz_packing, intensity_packing=intensity_packing, compress=compress, indexed=True)
    text = path.read_text()
    assert ("MS:100274" in text) == (compress and "linear" == mz_packing or compress and intensity_packing in ("pic", "slof"))
    fast, slow = read_mzml(path, native=True), read_mzml(path, native=False)
    assert same_run(fast, slow) and fast.offsets.tolist() == np.concatenate([[0], np.cumsum([len(s) for s in spectra])]).tolist()
    expected = np.concatenate(spectra)
    assert np.abs(fast.peaks[:, 0] - expected[:, 0]).max() < (1e-6 if mz_packing else 1e-12)
    if intensity_packing == "slof":
        assert np.all(np.abs(fast.peaks[:, 1] - expected[:, 1]) <= 5e-4 * expected[:, 1])
    else:
        assert np.array_equal(fast.peaks[:, 1], expected[:, 1])
    # without the native library the ElementTree reader decodes with numpress.py: same numbers
    saved, mzml._NATIVE = mzml._NATIVE, False
    try:
        assert same_run(read_mzml(path), slow)
    finally:
        mzml._NATIVE = saved
