"""MS-Numpress binary-array codecs of mzML (Teleman et al., Mol Cell Proteomics 2014; cvParams MS:1002312-1002314 and,
followed by zlib, MS:1002746-1002748): what pymzml decodes before the reference's scripts see ``centroidedPeaks``
(example_scripts/quantify_mzml.py:59-73), so files written with them are part of the N3 row.

* linear prediction (m/z): 8 bytes fixed point (an IEEE double, most significant byte first), the first two values as
  32-bit little-endian integers ``round(x * fixed_point)``, every further value as the residual against the linear
  extrapolation of the two before it, in the half-byte integer code below.
* positive integer ("pic", ion counts): every value rounded to an integer, in the half-byte code.
* short logged float ("slof", intensities): 8 bytes fixed point, then ``round(ln(x + 1) * fixed_point)`` as 16-bit
  little-endian integers.

Half-byte code of a 32-bit two's-complement integer: one half byte h, then the 8 - (h & 7 or 8 if h == 8) remaining
half bytes, least significant first; h <= 8 counts leading zero half bytes, h > 8 counts (h - 8) leading 0xF half bytes.
Half bytes are packed two per byte, first one high; a stream with an odd count ends in a zero half byte.

The decoders here are plain Python (the ElementTree reader's path and the tests' cross-check of the native reader,
``csrc/mzml_reader.cpp``, which decodes with the same rules at file speed); the encoders exist for the writer and the tests.
"""
import math
import struct

import numpy as np

LINEAR, PIC, SLOF = "MS:1002312", "MS:1002313", "MS:1002314"
LINEAR_ZLIB, PIC_ZLIB, SLOF_ZLIB = "MS:1002746", "MS:1002747", "MS:1002748"
KIND_OF_ACCESSION = {LINEAR: ("linear", False), PIC: ("pic", False), SLOF: ("slof", False),
                     LINEAR_ZLIB: ("linear", True), PIC_ZLIB: ("pic", True), SLOF_ZLIB: ("slof", True)}


class NumpressError(ValueError):
    pass


# ---- half-byte integers -------------------------------------------------------------------------------------------
def _int_to_half_bytes(value, out):
    x = value & 0xFFFFFFFF
    top = x & 0xF0000000
    if top == 0:
        lead = 8
        for i in range(8):
            if x & (0xF0000000 >> (4 * i)):
                lead = i
                break
        out.append(lead)
    elif top == 0xF0000000:
        lead = 7
        for i in range(8):
            mask = 0xF0000000 >> (4 * i)
            if x & mask != mask:
                lead = i
                break
        out.append(lead + 8)
    else:
        lead = 0
        out.append(0)
    for i in range(8 - lead):
        out.append((x >> (4 * i)) & 0xF)


def _pack_half_bytes(half_bytes):
    if len(half_bytes) % 2:
        half_bytes = half_bytes + [0]
    return bytes((half_bytes[i] << 4) | half_bytes[i + 1] for i in range(0, len(half_bytes), 2))


class _HalfByteReader:
    def __init__(self, data, start):
        self.data, self.position, self.low = data, start, False

    def at_end(self):
        """True at the end of the stream, or on the zero padding half byte of its last byte."""
        if self.position >= len(self.data):
            return True
        return self.position == len(self.data) - 1 and self.low and (self.data[self.position] & 0xF) == 0

    def _half_byte(self):
        if self.position >= len(self.data):
            raise NumpressError("numpress stream ends inside an integer")
        byte = self.data[self.position]
        if self.low:
            self.position += 1
            self.low = False
            return byte & 0xF
        self.low = True
        return byte >> 4

    def integer(self):
        head = self._half_byte()
        if head <= 8:
            lead, value = head, 0
        else:
            lead = head - 8
            value = 0
            for i in range(lead):
                value |= 0xF0000000 >> (4 * i)
        for i in range(8 - lead):
            value |= self._half_byte() << (4 * i)
        return value - (1 << 32) if value & 0x80000000 else value


def _fixed_point_of(data):
    if len(data) < 8:
        raise NumpressError("numpress stream shorter than its fixed point")
    return struct.unpack(">d", bytes(data[:8]))[0]


# ---- decoders -----------------------------------------------------------------------------------------------------
def decode_linear(data):
    data = bytes(data)
    if len(data) == 8:
        return np.zeros(0, dtype=np.float64)
    fixed_point = _fixed_point_of(data)
    if len(data) < 12:
        raise NumpressError("linear numpress stream is cut inside its first value")
    first = struct.unpack("<I", data[8:12])[0]
    values = [first]
    if len(data) > 12:
        if len(data) < 16:
            raise NumpressError("linear numpress stream is cut inside its second value")
        values.append(struct.unpack("<I", data[12:16])[0])
        reader = _HalfByteReader(data, 16)
        before, last = values
        while not reader.at_end():
            value = last + (last - before) + reader.integer()
            values.append(value)
            before, last = last, value
    return np.array(values, dtype=np.float64) / fixed_point


def decode_pic(data):
    data = bytes(data)
    reader = _HalfByteReader(data, 0)
    values = []
    while not reader.at_end():
        values.append(reader.integer() & 0xFFFFFFFF)
    return np.array(values, dtype=np.float64)


def decode_slof(data):
    data = bytes(data)
    fixed_point = _fixed_point_of(data)
    if (len(data) - 8) % 2:
        raise NumpressError("slof numpress stream has half a value at its end")
    counts = np.frombuffer(data, dtype="<u2", offset=8).tolist()
    # math.exp is the C library's exp, the one the native reader calls (numpy's vectorised exp may differ in the last bit)
    return np.array([math.exp(count / fixed_point) - 1.0 for count in counts], dtype=np.float64)


DECODERS = {"linear": decode_linear, "pic": decode_pic, "slof": decode_slof}


# ---- encoders (writer, tests) -------------------------------------------------------------------------------------
def optimal_linear_fixed_point(values):
    """Largest fixed point that keeps the first two values and every residual inside 31 bits."""
    values = np.asarray(values, dtype=np.float64)
    if len(values) == 0:
        return 0.0
    if len(values) == 1:
        return math.floor(0xFFFFFFFF / values[0]) if values[0] > 0 else 1.0
    largest = float(max(values[0], values[1]))
    if len(values) > 2:
        residuals = values[2:] - (2.0 * values[1:-1] - values[:-2])
        largest = max(largest, float(np.max(np.ceil(np.abs(residuals) + 1.0))))
    return math.floor(0x7FFFFFFF / largest) if largest > 0 else 1.0


def optimal_slof_fixed_point(values):
    values = np.asarray(values, dtype=np.float64)
    largest = float(np.max(np.log(values + 1.0))) if len(values) else 0.0
    return math.floor(0xFFFF / largest) if largest > 0 else 1.0


def encode_linear(values, fixed_point=None):
    values = np.asarray(values, dtype=np.float64)
    if fixed_point is None:
        fixed_point = optimal_linear_fixed_point(values)
    out = struct.pack(">d", fixed_point)
    if len(values) == 0:
        return out
    scaled = [int(v * fixed_point + 0.5) for v in values.tolist()]
    if min(scaled[:2]) < 0 or max(scaled[:2]) > 0xFFFFFFFF:
        raise NumpressError("linear numpress needs its first two values in [0, 2^32 / fixed point)")
    out += struct.pack("<I", scaled[0])
    if len(scaled) == 1:
        return out
    out += struct.pack("<I", scaled[1])
    half_bytes = []
    for i in range(2, len(scaled)):
        residual = scaled[i] - (2 * scaled[i - 1] - scaled[i - 2])
        if not -(1 << 31) <= residual < (1 << 31):
            raise NumpressError("linear numpress residual does not fit 32 bits (fixed point too large)")
        _int_to_half_bytes(residual, half_bytes)
    return out + _pack_half_bytes(half_bytes)


def encode_pic(values):
    half_bytes = []
    for v in np.asarray(values, dtype=np.float64).tolist():
        count = int(v + 0.5)
        if not 0 <= count <= 0xFFFFFFFF:
            raise NumpressError("pic numpress needs counts in [0, 2^32)")
        _int_to_half_bytes(count, half_bytes)
    return _pack_half_bytes(half_bytes)


def encode_slof(values, fixed_point=None):
    values = np.asarray(values, dtype=np.float64)
    if fixed_point is None:
        fixed_point = optimal_slof_fixed_point(values)
    counts = np.floor(np.log(values + 1.0) * fixed_point + 0.5)
    if len(counts) and (counts.min() < 0 or counts.max() > 0xFFFF):
        raise NumpressError("slof numpress value outside the 16-bit range of this fixed point")
    return struct.pack(">d", fixed_point) + counts.astype("<u2").tobytes()


ENCODERS = {"linear": encode_linear, "pic": encode_pic, "slof": encode_slof}
