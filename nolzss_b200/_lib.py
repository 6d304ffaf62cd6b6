"""ctypes loader for the C ABI in include/nolzss_b200.h (libnolzss_b200.so, built by build.py).

There is no CPU fallback: if the shared library is missing it is built with nvcc; if that fails, or
if no CUDA device is present when a compute entry point is called, the call raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

import numpy as np

from . import build as _build

NLZ_OK, NLZ_ERR_RUNTIME, NLZ_ERR_INVALID, NLZ_ERR_CUDA = 0, 1, 2, 3
MODE_GENERAL, MODE_RC_PREPARED, MODE_DNA_RC = 0, 1, 2
RC_MASK = 1 << 63

_u64 = ctypes.c_uint64
_u64p = ctypes.POINTER(ctypes.c_uint64)
_u64pp = ctypes.POINTER(_u64p)
_vp = ctypes.c_void_p


class Stats(ctypes.Structure):
    _fields_ = [
        ("n_text", _u64), ("n_suffixes", _u64), ("n_factorized", _u64), ("n_factors", _u64),
        ("active_sum", _u64), ("walk_nodes", _u64), ("hard_positions", _u64), ("workspace_bytes", _u64),
        ("key_bits", ctypes.c_uint32), ("sym_bits", ctypes.c_uint32), ("key_syms", ctypes.c_uint32),
        ("doubling_rounds", ctypes.c_uint32), ("tile_sort_rounds", ctypes.c_uint32), ("kernel_launches", ctypes.c_uint32),
        ("host_syncs", ctypes.c_uint32),
        ("ms_total", ctypes.c_float), ("ms_prepare", ctypes.c_float), ("ms_keys", ctypes.c_float),
        ("ms_sort0", ctypes.c_float), ("ms_doubling", ctypes.c_float), ("ms_lcp", ctypes.c_float),
        ("ms_lpnf", ctypes.c_float), ("ms_chain", ctypes.c_float),
        ("rank_records_applied", _u64), ("lcp_marked", _u64), ("n_local_suffixes", _u64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# name -> (restype, argtypes); every symbol declared in include/nolzss_b200.h
SIGNATURES = {
    "nlz_ctx_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_vp)]),
    "nlz_ctx_destroy": (None, [_vp]),
    "nlz_ctx_device": (ctypes.c_int, [_vp]),
    "nlz_last_error": (ctypes.c_char_p, []),
    "nlz_free": (None, [_vp]),
    "nlz_get_stats": (ctypes.c_int, [_vp, ctypes.POINTER(Stats)]),
    "nlz_version": (ctypes.c_char_p, []),
    "nlz_host_register": (ctypes.c_int, [_vp, _u64]),
    "nlz_host_unregister": (ctypes.c_int, [_vp]),
    "nlz_set_profiling": (ctypes.c_int, [_vp, ctypes.c_int]),
    "nlz_kernel_class_count": (ctypes.c_int, []),
    "nlz_set_debug_flags": (ctypes.c_int, [_vp, ctypes.c_int]),
    "nlz_get_kernel_stats": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p),
                                             ctypes.POINTER(ctypes.c_double), _u64p,
                                             ctypes.POINTER(ctypes.c_uint32)]),
    "nlz_factorize_mode": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _u64, _u64, _u64pp, _u64p]),
    "nlz_factorize_mode_into": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _u64, _u64, _vp, _u64, _u64p]),
    "nlz_count_mode": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _u64, _u64, _u64p]),
    "nlz_factorize_device": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _u64, _u64, _vp, _vp, _u64, _u64p]),
    "nlz_factorize_batch": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _vp, _u64, _u64pp, _vp, _u64p]),
    "nlz_dist_create": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _u64, ctypes.c_int, ctypes.POINTER(_vp)]),
    "nlz_dist_destroy": (None, [_vp]),
    "nlz_dist_ipc_handle_bytes": (ctypes.c_int, []),
    "nlz_dist_export": (ctypes.c_int, [_vp, _vp]),
    "nlz_dist_attach": (ctypes.c_int, [_vp, _vp]),
    "nlz_dist_attach_local": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int]),
    "nlz_dist_factorize": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _u64, _u64pp, _u64p]),
    "nlz_dist_factorize_into": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _u64, _vp, _u64, _u64p]),
    "nlz_factorize": (ctypes.c_int, [_vp, _vp, _u64, _u64, _u64pp, _u64p]),
    "nlz_count_factors": (ctypes.c_int, [_vp, _vp, _u64, _u64, _u64p]),
    "nlz_factorize_dna_w_rc": (ctypes.c_int, [_vp, _vp, _u64, _u64pp, _u64p]),
    "nlz_count_factors_dna_w_rc": (ctypes.c_int, [_vp, _vp, _u64, _u64p]),
    "nlz_factorize_multiple_dna_w_rc": (ctypes.c_int, [_vp, _vp, _u64, _u64, _u64pp, _u64p]),
    "nlz_count_factors_multiple_dna_w_rc": (ctypes.c_int, [_vp, _vp, _u64, _u64, _u64p]),
    "nlz_prepare_multiple_dna_sequences_w_rc": (ctypes.c_int, [_vp, _vp, _u64, ctypes.POINTER(_vp), _u64p, _u64p, _u64pp, _u64p]),
    "nlz_prepare_multiple_dna_sequences_no_rc": (ctypes.c_int, [_vp, _vp, _u64, ctypes.POINTER(_vp), _u64p, _u64p, _u64pp, _u64p]),
    "nlz_fasta_parse": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(_vp)]),
    "nlz_fasta_num_sequences": (_u64, [_vp]),
    "nlz_fasta_id": (ctypes.c_char_p, [_vp, _u64]),
    "nlz_fasta_sequence": (_vp, [_vp, _u64, _u64p]),
    "nlz_fasta_free": (None, [_vp]),
    "nlz_identify_sentinel_factors": (ctypes.c_int, [_vp, _u64, _vp, _u64, _u64pp, _u64p]),
    "nlz_write_factor_file": (ctypes.c_int, [ctypes.c_char_p, _vp, _u64, _vp, _u64, _u64, _u64, _u64]),
    "nlz_factorize_file_mode": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_char_p, _u64, _u64pp, _u64p]),
    "nlz_count_file_mode": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_char_p, _u64, _u64p]),
    "nlz_write_factors_binary_file_mode": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, _u64, _u64p]),
    "nlz_parallel_factorize_to_file": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _u64, ctypes.c_char_p, _u64, _u64p]),
    "nlz_factorize_w_reference": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _u64, _vp, _u64, ctypes.c_char_p, _u64pp, _u64p]),
    "nlz_factorize_fasta": (ctypes.c_int, [_vp, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p,
                                            _u64pp, _u64p, _u64pp, _u64p, ctypes.POINTER(_vp)]),
    "nlz_factorize_fasta_per_sequence": (ctypes.c_int, [_vp, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p,
                                                         ctypes.c_int, ctypes.c_int, _u64pp, _u64pp, _u64p,
                                                         ctypes.POINTER(_vp)]),
    "nlz_debug_index": (ctypes.c_int, [_vp, _vp, _u64, _vp, _vp, _vp]),
    "nlz_debug_sort_pairs_u64": (ctypes.c_int, [_vp, _vp, _vp, _u64, ctypes.c_int, ctypes.c_int]),
    "nlz_debug_sort_pairs_u32": (ctypes.c_int, [_vp, _vp, _vp, _u64, ctypes.c_int, ctypes.c_int]),
    "nlz_debug_per_position": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _u64, _vp, _vp, _u64, _u64p]),
}

_lib = None
_lib_lock = threading.Lock()
_ctx = {}


def load():
    """Loads (building if needed) the shared library and binds every declared symbol."""
    global _lib
    with _lib_lock:
        if _lib is None:
            path = _build.LIB
            if not os.path.exists(path):
                _build.build()
            L = ctypes.CDLL(path)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(L, name)   # AttributeError here means header and library disagree
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


def _raise(rc: int):
    msg = load().nlz_last_error().decode("utf-8", "replace")
    if rc == NLZ_ERR_INVALID:
        raise ValueError(msg)
    raise RuntimeError(msg)


def check(rc: int):
    if rc != NLZ_OK:
        _raise(rc)


def context(device: int | None = None) -> int:
    """Process-wide context handle for `device` (default: LOCAL_RANK or 0)."""
    if device is None:
        device = int(os.environ.get("NOLZSS_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    with _lib_lock:
        h = _ctx.get(device)
    if h is None:
        L = load()
        out = _vp()
        check(L.nlz_ctx_create(device, ctypes.byref(out)))
        with _lib_lock:
            h = _ctx.setdefault(device, out.value)
    return h


def _as_buffer(data):
    """bytes-like -> (address, length, keepalive); itemsize-1, 1-D, like bindings.cpp:58-67."""
    mv = memoryview(data)
    if mv.itemsize != 1:
        raise ValueError("buffer must be a bytes-like object with itemsize==1")
    if mv.ndim != 1:
        raise ValueError("buffer must be a 1-dimensional bytes-like object")
    if not mv.c_contiguous:
        mv = memoryview(bytes(mv))
    arr = np.frombuffer(mv, dtype=np.uint8)
    return arr.ctypes.data if arr.size else None, arr.size, arr


def factorize_array(mode: int, data, start_pos: int = 0, device: int | None = None) -> np.ndarray:
    """(z, 3) uint64 array of (start, length, ref-with-RC_MASK)."""
    L = load()
    addr, n, keep = _as_buffer(data)
    out = _u64p()
    cnt = _u64(0)
    check(L.nlz_factorize_mode(context(device), mode, addr, n, start_pos, ctypes.byref(out), ctypes.byref(cnt)))
    z = cnt.value
    if z == 0:
        return np.zeros((0, 3), dtype=np.uint64)
    try:
        arr = np.ctypeslib.as_array(out, shape=(z * 3,)).copy().reshape(z, 3)
    finally:
        L.nlz_free(out)
    return arr


def factorize_batch(records, with_rc: bool, want_factors: bool = True, device: int | None = None):
    """Independent records in one segmented pipeline run (nlz_factorize_batch).

    Returns (triples, counts): record-local (z, 3) uint64 triples concatenated in record order (None when
    want_factors is False) and the per-record factor counts."""
    L = load()
    lens = np.array([len(r) for r in records], dtype=np.uint64)
    offs = np.zeros(len(records), dtype=np.uint64)
    if len(records):
        offs[1:] = np.cumsum(lens)[:-1]
    concat = np.frombuffer(b"".join(bytes(r) for r in records), dtype=np.uint8)
    counts = np.zeros(max(len(records), 1), dtype=np.uint64)
    out = _u64p()
    total = _u64(0)
    check(L.nlz_factorize_batch(context(device), 1 if with_rc else 0, concat.ctypes.data if concat.size else None,
                                offs.ctypes.data, lens.ctypes.data, len(records),
                                ctypes.byref(out) if want_factors else None, counts.ctypes.data, ctypes.byref(total)))
    counts = counts[:len(records)]
    if not want_factors:
        return None, counts
    z = total.value
    if z == 0:
        return np.zeros((0, 3), dtype=np.uint64), counts
    try:
        arr = np.ctypeslib.as_array(out, shape=(z * 3,)).copy().reshape(z, 3)
    finally:
        L.nlz_free(out)
    return arr, counts


def count(mode: int, data, start_pos: int = 0, device: int | None = None) -> int:
    L = load()
    addr, n, keep = _as_buffer(data)
    cnt = _u64(0)
    check(L.nlz_count_mode(context(device), mode, addr, n, start_pos, ctypes.byref(cnt)))
    return cnt.value


def stats(device: int | None = None) -> dict:
    s = Stats()
    check(load().nlz_get_stats(context(device), ctypes.byref(s)))
    return s.as_dict()


def set_profiling(on: bool, device: int | None = None):
    check(load().nlz_set_profiling(context(device), 1 if on else 0))


def kernel_stats(device: int | None = None) -> dict:
    """{class name: {"ms", "bytes", "launches"}} for the last call on this device's context."""
    Lb = load()
    out = {}
    for cls in range(Lb.nlz_kernel_class_count()):
        name = ctypes.c_char_p()
        ms = ctypes.c_double(0)
        by = _u64(0)
        ln = ctypes.c_uint32(0)
        check(Lb.nlz_get_kernel_stats(context(device), cls, ctypes.byref(name), ctypes.byref(ms),
                                      ctypes.byref(by), ctypes.byref(ln)))
        out[name.value.decode()] = {"ms": ms.value, "bytes": by.value, "launches": ln.value}
    return out
