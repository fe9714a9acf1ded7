"""mzML -> packed spectra (the step before the matching path).

The reference's scripts loop ``pymzml.run.Reader`` and hand every MS1 spectrum's ``centroidedPeaks`` to
``match_all`` (example_scripts/quantify_mzml.py:59-73).  This reader goes from the file straight to the packed
layout the batched matcher consumes -- one (N, 2) float64 array of (m/z, intensity) rows plus row offsets --
without a Python object per peak (plain or gzip-compressed files): binary arrays are base64/zlib-decoded into numpy buffers (mzML 1.1: MS:1000514
m/z array, MS:1000515 intensity array, MS:1000521/523 32/64-bit float, MS:1000519/522 32/64-bit integer, MS:1000574 zlib,
MS:1000576 none, MS:1002312-1002314 / MS:1002746-1002748 MS-Numpress alone / followed by zlib -- ``numpress.py``).

    run = read_mzml("BSA1.mzML", ms_level=1)
    results = lib.match_many(run.peaks, file_name=run.name, spec_ids=run.spec_ids, spec_rts=run.rts, offsets=run.offsets)
"""
import base64
import ctypes as C
import gzip
import os
import re
import zlib
from collections import namedtuple
from pathlib import Path
import xml.etree.ElementTree as ET

import numpy as np

from . import numpress

Run = namedtuple("Run", ["name", "peaks", "offsets", "spec_ids", "rts", "ms_levels"])

_MZ, _INTENSITY = "MS:1000514", "MS:1000515"
_F32, _F64, _ZLIB = "MS:1000521", "MS:1000523", "MS:1000574"
_I32, _I64 = "MS:1000519", "MS:1000522"
_TRUNCATION_CODECS = ("MS:1003088", "MS:1003089", "MS:1003090")
_MS_LEVEL, _SCAN_START = "MS:1000511", "MS:1000016"
_ID_NUMBER = re.compile(r"(?:scan|index|spectrum)=(\d+)")


def _local(tag):
    return tag.rsplit("}", 1)[-1]


def _decode(array_element):
    kind, dtype, compressed, packed, text = None, np.float64, False, None, ""
    for child in array_element:
        name = _local(child.tag)
        if name == "cvParam":
            accession = child.get("accession")
            if accession in (_MZ, _INTENSITY):
                kind = accession
            elif accession == _F32:
                dtype = np.float32
            elif accession == _F64:
                dtype = np.float64
            elif accession == _ZLIB:
                compressed = True
            elif accession in (_I32, _I64):
                packed = np.int32 if accession == _I32 else np.int64
            elif accession in numpress.KIND_OF_ACCESSION:
                packed, then_zlib = numpress.KIND_OF_ACCESSION[accession]
                compressed = compressed or then_zlib
            elif accession in _TRUNCATION_CODECS:
                raise ValueError("binary array encoding %s (%s) is not decoded by this reader" % (accession, child.get("name")))
        elif name == "binary":
            text = child.text or ""
    raw = base64.b64decode(text) if text.strip() else b""
    if compressed and raw:
        raw = zlib.decompress(raw)
    if isinstance(packed, str):
        return kind, _numpress_values(packed, raw)
    if packed is not None:
        dtype = packed
    return kind, np.frombuffer(raw, dtype=np.dtype(dtype).newbyteorder("<")).astype(np.float64)


def _numpress_values(kind, raw):
    """One numpress stream -> float64 values: through libqms_mzml.so when it is built, else the plain Python decoder."""
    lib = _native()
    if lib is None:
        return numpress.DECODERS[kind](raw)
    count = C.c_int64()
    code = ("linear", "pic", "slof").index(kind)
    if lib.qms_numpress_decode(code, raw, len(raw), None, 0, C.byref(count)) != 0:
        raise numpress.NumpressError(lib.qms_mzml_last_error().decode())
    values = np.empty(count.value, dtype=np.float64)
    if lib.qms_numpress_decode(code, raw, len(raw), values.ctypes.data, len(values), C.byref(count)) != 0:
        raise numpress.NumpressError(lib.qms_mzml_last_error().decode())
    return values


def spectrum_id(native_id, index):
    """pymzml-style integer id: the scan number of the native id if there is one, else index + 1."""
    found = _ID_NUMBER.search(native_id or "")
    return int(found.group(1)) if found else index + 1


LAST_READ = {"scan_blocks": 0}      # facts about the last native read: blocks the tag scan was split into (0 = sequential)
_NATIVE = None
_NATIVE_PATH = Path(__file__).resolve().parent / "libqms_mzml.so"


def _native():
    """ctypes handle of libqms_mzml.so (include/qms_mzml.h), or None if it has not been built."""
    global _NATIVE
    if _NATIVE is None:
        if not _NATIVE_PATH.exists():
            _NATIVE = False
        else:
            lib = C.CDLL(str(_NATIVE_PATH))
            lib.qms_mzml_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
            lib.qms_mzml_close.argtypes = [C.c_void_p]
            lib.qms_mzml_close.restype = None
            lib.qms_mzml_last_error.restype = C.c_char_p
            lib.qms_mzml_count.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
            lib.qms_mzml_read.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 6
            lib.qms_mzml_scan_blocks.argtypes = [C.c_void_p]
            lib.qms_numpress_decode.argtypes = [C.c_int32, C.c_char_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
            _NATIVE = lib
    return _NATIVE or None


def _read_native(lib, path, ms_level, sort_peaks, threads):
    """The packed run through libqms_mzml.so; None when the file needs a decoder the native reader lacks."""
    handle = C.c_void_p()
    rc = lib.qms_mzml_open(os.fsencode(str(path)), C.byref(handle))
    if rc != 0:
        raise OSError("%s: %s" % (path, lib.qms_mzml_last_error().decode()))
    try:
        LAST_READ["scan_blocks"] = lib.qms_mzml_scan_blocks(handle)
        n_spectra, n_peaks = C.c_int64(), C.c_int64()
        level = 0 if ms_level is None else int(ms_level)
        rc = lib.qms_mzml_count(handle, level, C.byref(n_spectra), C.byref(n_peaks))
        if rc == -3:
            return None
        if rc != 0:
            raise ValueError(lib.qms_mzml_last_error().decode())
        n = n_spectra.value
        peaks = np.empty((n_peaks.value, 2), dtype=np.float64)
        offsets = np.zeros(n + 1, dtype=np.int64)
        ids, rt = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.float64)
        unit, levels = np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.int32)
        rc = lib.qms_mzml_read(handle, level, 1 if sort_peaks else 0, int(threads or 0), peaks.ctypes.data, offsets.ctypes.data,
                               ids.ctypes.data, rt.ctypes.data, unit.ctypes.data, levels.ctypes.data)
        if rc != 0:
            raise ValueError(lib.qms_mzml_last_error().decode())
    finally:
        lib.qms_mzml_close(handle)
    names = (None, "minute", "second")
    rts = [None if u == 0 else (value, names[u]) for value, u in zip(rt.tolist(), unit.tolist())]
    return Run(os.path.basename(str(path)), peaks, offsets, ids.tolist(), rts, [None if v < 0 else v for v in levels.tolist()])


def read_mzml(path, ms_level=1, sort_peaks=True, native=None, threads=None):
    """Read every spectrum of ``ms_level`` (None = all) into a packed ``Run``.

    ``rts`` holds ``(value, unit)`` tuples like pymzml's ``scan_time`` (unit 'minute' or 'second'), which is
    what ``Results``/``calc_amounts`` understand (results.py:1240-1246).

    ``native``: None (default) uses libqms_mzml.so when it is built -- the file is memory-mapped, indexed by a tag
    scanner and decoded by ``threads`` threads; True insists on the native reader, False on the ElementTree one below (same
    encodings: floats, integers, MS-Numpress, zlib; neither decodes the truncation codecs MS:1003088-1003090)."""
    if native is not False:
        lib = _native()
        if lib is None and native:
            raise ImportError("%s not found: build it with `make -C pyqms_b200/csrc`" % _NATIVE_PATH)
        if lib is not None:
            run = _read_native(lib, path, ms_level, sort_peaks, threads)
            if run is not None:
                return run
            if native:
                raise ValueError("%s uses an array encoding the native reader does not decode" % path)
    mz_parts, int_parts, lengths, ids, rts, levels = [], [], [], [], [], []
    index = -1
    source = gzip.open(str(path), "rb") if str(path).lower().endswith(".gz") else str(path)
    for _, element in ET.iterparse(source, events=("end",)):
        if _local(element.tag) != "spectrum":
            continue
        index += 1
        level, scan_time = None, None
        for param in element.iter():
            if _local(param.tag) != "cvParam":
                continue
            accession = param.get("accession")
            if accession == _MS_LEVEL:
                level = int(param.get("value"))
            elif accession == _SCAN_START and scan_time is None:
                unit = (param.get("unitName") or "minute").lower()
                scan_time = (float(param.get("value")), "second" if unit.startswith("sec") else "minute")
        if ms_level is None or level == ms_level:
            mz = intensity = None
            for array_element in element.iter():
                if _local(array_element.tag) != "binaryDataArray":
                    continue
                kind, values = _decode(array_element)
                if kind == _MZ:
                    mz = values
                elif kind == _INTENSITY:
                    intensity = values
            if mz is None or intensity is None or len(mz) != len(intensity):
                raise ValueError("spectrum %r: m/z and intensity arrays missing or of different length" % element.get("id"))
            if sort_peaks and len(mz) > 1 and np.any(mz[1:] < mz[:-1]):
                order = np.argsort(mz, kind="stable")
                mz, intensity = mz[order], intensity[order]
            mz_parts.append(mz)
            int_parts.append(intensity)
            lengths.append(len(mz))
            ids.append(spectrum_id(element.get("id"), index))
            rts.append(scan_time)
            levels.append(level)
        element.clear()
    offsets = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    peaks = np.empty((int(offsets[-1]), 2), dtype=np.float64)
    if lengths:
        peaks[:, 0] = np.concatenate(mz_parts)
        peaks[:, 1] = np.concatenate(int_parts)
    return Run(os.path.basename(str(path)), peaks, offsets, ids, rts, levels)


def write_mzml(path, spectra, rts_minutes=None, ms_levels=None, precision=64, compress=True, indexed=False, mz_packing=None,
               intensity_packing=None):
    """Minimal mzML 1.1 writer (test fixtures and synthetic runs): ``spectra`` = list of (n, 2) arrays.
    ``indexed``: wrap the document in ``<indexedmzML>`` with the byte offset of every spectrum, as converters do.
    ``mz_packing`` / ``intensity_packing``: None (floats of ``precision`` bits), "int32", "int64", or an MS-Numpress codec
    "linear", "pic", "slof" (with ``compress`` the combined numpress-then-zlib accession is written, as ProteoWizard does)."""
    dtype = np.dtype("<f8" if precision == 64 else "<f4")
    numpress_accession = {("linear", False): numpress.LINEAR, ("pic", False): numpress.PIC, ("slof", False): numpress.SLOF,
                          ("linear", True): numpress.LINEAR_ZLIB, ("pic", True): numpress.PIC_ZLIB, ("slof", True): numpress.SLOF_ZLIB}

    def binary(values, accession, name, packing):
        params = [(_F64 if precision == 64 else _F32, "%d-bit float" % precision)]
        if packing in ("int32", "int64"):
            raw = np.ascontiguousarray(np.rint(values), dtype="<i4" if packing == "int32" else "<i8").tobytes()
            params = [(_I32 if packing == "int32" else _I64, "%s-bit integer" % packing[3:])]
        elif packing in numpress.ENCODERS:
            raw = numpress.ENCODERS[packing](values)
            params.append((numpress_accession[(packing, bool(compress))], "MS-Numpress %s%s" % (packing, " followed by zlib" if compress else "")))
        elif packing is None:
            raw = np.ascontiguousarray(values, dtype=dtype).tobytes()
        else:
            raise ValueError("unknown packing %r" % (packing,))
        if compress:
            raw = zlib.compress(raw)
        if packing not in numpress.ENCODERS:
            params.append((_ZLIB, "zlib compression") if compress else ("MS:1000576", "no compression"))
        text = base64.b64encode(raw).decode()
        return '<binaryDataArray encodedLength="%d">%s<cvParam cvRef="MS" accession="%s" name="%s"/><binary>%s</binary></binaryDataArray>' % (
            len(text), "".join('<cvParam cvRef="MS" accession="%s" name="%s"/>' % p for p in params), accession, name, text)

    with open(path, "w", newline="") as out:
        written, offsets = 0, []

        def emit(text):
            nonlocal written
            out.write(text)
            written += len(text)             # the document is pure ASCII: characters = bytes

        emit('<?xml version="1.0" encoding="utf-8"?>\n')
        if indexed:
            emit('<indexedmzML xmlns="http://psi.hupo.org/ms/mzml">\n')
        emit('<mzML xmlns="http://psi.hupo.org/ms/mzml" version="1.1.0">\n<run id="synthetic"><spectrumList count="%d">\n' % len(spectra))
        for n, spectrum in enumerate(spectra):
            spectrum = np.asarray(spectrum, dtype=np.float64).reshape(-1, 2)
            level = 1 if ms_levels is None else ms_levels[n]
            rt = 0.01 * n if rts_minutes is None else rts_minutes[n]
            offsets.append(written)
            emit('<spectrum index="%d" id="controllerType=0 controllerNumber=1 scan=%d" defaultArrayLength="%d">'
                 '<cvParam cvRef="MS" accession="MS:1000511" name="ms level" value="%d"/>'
                 '<scanList count="1"><scan><cvParam cvRef="MS" accession="MS:1000016" name="scan start time" '
                 'value="%r" unitName="minute"/></scan></scanList><binaryDataArrayList count="2">%s%s'
                 "</binaryDataArrayList></spectrum>\n" % (
                     n, n + 1, len(spectrum), level, rt, binary(spectrum[:, 0], _MZ, "m/z array", mz_packing),
                     binary(spectrum[:, 1], _INTENSITY, "intensity array", intensity_packing)))
        emit("</spectrumList></run></mzML>\n")
        if indexed:
            list_at = written
            emit('<indexList count="1">\n<index name="spectrum">\n')
            for n, at in enumerate(offsets):
                emit('<offset idRef="controllerType=0 controllerNumber=1 scan=%d">%d</offset>\n' % (n + 1, at))
            emit("</index>\n</indexList>\n<indexListOffset>%d</indexListOffset>\n</indexedmzML>\n" % list_at)
