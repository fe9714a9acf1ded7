"""ctypes binding of libqms_b200.so (include/qms_b200.h).  No CPU fallback: a missing library or
a missing CUDA device raises."""
import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libqms_b200.so"

QMS_OK, QMS_E_CUDA, QMS_E_ARG, QMS_E_CAPACITY, QMS_E_PARAM, QMS_E_LIMIT, QMS_E_INPUT = 0, -1, -2, -3, -4, -5, -6


class QmsError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libqms_b200 error %d: %s" % (code, message))
        self.code = code


class QmsParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "element_min_abundance", "min_rel_peak_intensity", "intensity_transformation", "lower_mz_limit",
        "upper_mz_limit", "rel_mz_range", "rel_i_range", "internal_precision", "machine_offset_ppm",
        "m_score_threshold", "required_peak_overlap", "proton")] + [
        ("min_matched_isotopologues", C.c_int32), ("max_molecules_per_bin", C.c_int32)]


class QmsCounters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "spectra", "peaks", "live_peaks", "incidences", "candidates", "candidate_windows", "combinations", "hits",
        "hit_windows", "kernel_launches")] + [(n, C.c_double) for n in (
            "ms_probe", "ms_gate", "ms_score", "ms_compact", "ms_h2d", "ms_d2h", "ms_trees", "ms_envelopes", "ms_index")] + [
        ("envelopes", C.c_int64), ("envelope_flops", C.c_int64), ("dense_batches", C.c_int64), ("gated_incidences", C.c_int64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class QmsIndexInfo(C.Structure):
    _fields_ = [("n_entries", C.c_int64), ("n_windows", C.c_int64), ("t_min", C.c_int32), ("t_max", C.c_int32),
                ("max_width", C.c_int32), ("mz_min", C.c_double), ("mz_max", C.c_double)]


class QmsHits(C.Structure):
    _fields_ = [("cap_hits", C.c_int64), ("cap_windows", C.c_int64), ("hit_spec", C.c_void_p), ("hit_entry", C.c_void_p),
                ("hit_score", C.c_void_p), ("hit_scaling", C.c_void_p), ("hit_win_off", C.c_void_p), ("hit_mmz", C.c_void_p),
                ("hit_mi", C.c_void_p), ("hit_peak", C.c_void_p), ("hit_peak32", C.c_void_p), ("hit_nc8", C.c_void_p),
                ("hit_peak16", C.c_void_p), ("spec_hit_off", C.c_void_p)]


EXPORTS = {  # name -> (restype, argtypes); must list every symbol of include/qms_b200.h
    "qms_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "qms_ctx_destroy": (None, [C.c_void_p]),
    "qms_last_error": (C.c_char_p, [C.c_void_p]),
    "qms_set_params": (C.c_int, [C.c_void_p, C.POINTER(QmsParams)]),
    "qms_stream": (C.c_void_p, [C.c_void_p]),
    "qms_synchronize": (C.c_int, [C.c_void_p]),
    "qms_get_counters": (C.c_int, [C.c_void_p, C.POINTER(QmsCounters)]),
    "qms_reset_counters": (C.c_int, [C.c_void_p]),
    "qms_device_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "qms_set_tuning": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32]),
    "qms_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "qms_host_free": (C.c_int, [C.c_void_p]),
    "qms_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "qms_host_unregister": (C.c_int, [C.c_void_p]),
    "qms_recalc_distribution": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p]),
    "qms_build_element_trees": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int32]),
    "qms_export_tree_level": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                        C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "qms_set_envelope_terms": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qms_build_envelopes": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64]),
    "qms_envelope_layout": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p]),
    "qms_envelope_buffers": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_void_p)] * 7),
    "qms_comm_unique_id": (C.c_int, [C.c_void_p, C.c_size_t]),
    "qms_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "qms_allgather_library": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "qms_comm_destroy": (C.c_int, [C.c_void_p]),
    "qms_finalize_index": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "qms_index_info_get": (C.c_int, [C.c_void_p, C.POINTER(QmsIndexInfo)]),
    "qms_export_envelopes": (C.c_int, [C.c_void_p] + [C.c_void_p] * 7),
    "qms_export_charge_plane": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qms_export_index": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5),
    "qms_export_windows": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5),
    "qms_match_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_double,
                                  C.POINTER(QmsHits), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "qms_fetch_hits": (C.c_int, [C.c_void_p, C.POINTER(QmsHits), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "qms_match_entry": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_double,
                                  C.POINTER(QmsHits), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "qms_transform_mz": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qms_calc_amounts": (C.c_int, [C.c_void_p, C.c_int64] + [C.c_void_p] * 8),
    "qms_measure_fp64_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "qms_parse_molecules": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int64, C.c_int32, C.c_char_p, C.c_void_p, C.c_int32,
                                      C.c_char_p] + [C.c_void_p] * 6),
    "qms_hill_formulas": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int64,
                                    C.c_void_p, C.POINTER(C.c_int64)]),
    "qms_score_matches": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """dlopen the in-tree shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        path = Path(os.environ.get("QMS_B200_LIB", LIB_PATH))
        if not path.exists():
            raise ImportError("%s not found: build it with `make -C pyqms_b200/csrc` (or __graft_entry__.build()); "
                              "pyqms_b200 has no CPU fallback" % path)
        lib = C.CDLL(str(path))
        for name, (restype, argtypes) in EXPORTS.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def ptr(array):
    """Host pointer of a C-contiguous numpy array (None -> NULL)."""
    if array is None:
        return None
    assert array.flags["C_CONTIGUOUS"]
    return array.ctypes.data_as(C.c_void_p)


class Context:
    """One ``qms_ctx``: one GPU, one stream, all device buffers of one library."""

    def __init__(self, device=None):
        self.lib = load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        handle = C.c_void_p()
        rc = self.lib.qms_ctx_create(int(device), C.byref(handle))
        if rc != QMS_OK:
            raise QmsError(rc, (self.lib.qms_last_error(None) or b"").decode())
        self.handle = handle
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.qms_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != QMS_OK:
            raise QmsError(rc, (self.lib.qms_last_error(self.handle) or b"").decode())

    def call(self, name, *args):
        self.check(getattr(self.lib, name)(self.handle, *args))

    def counters(self):
        c = QmsCounters()
        self.call("qms_get_counters", C.byref(c))
        return c.as_dict()

    def reset_counters(self):
        self.call("qms_reset_counters")

    def set_tuning(self, key, value):
        self.call("qms_set_tuning", key.encode(), int(value))

    def stream(self):
        return self.lib.qms_stream(self.handle)

    def device_bytes(self):
        n = C.c_int64()
        self.call("qms_device_bytes", C.byref(n))
        return n.value
