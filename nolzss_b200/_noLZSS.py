"""Drop-in for the reference's pybind11 extension module `noLZSS._noLZSS`
(/root/reference/src/cpp/bindings.cpp:39-1518): the same 42 function names, argument names, defaults,
return shapes and exception types, implemented as a ctypes shim over the C ABI of
include/nolzss_b200.h.  Every factor is computed by the CUDA library; nothing here falls back to a
CPU implementation.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _lib as L

RC_MASK = 1 << 63
_U64 = ctypes.c_uint64
_U64P = ctypes.POINTER(ctypes.c_uint64)


def _version() -> str:
    return L.load().nlz_version().decode()


__version__ = "1.2.0"   # bindings.cpp:1513-1517 (the reference reports its package version here)


# ----------------------------------------------------------------------------- result classes
class Factor:
    """bindings.cpp:44-48 -- start, length, ref (RC_MASK stripped), is_rc."""

    __slots__ = ("start", "length", "_ref")

    def __init__(self, start: int, length: int, ref: int):
        self.start, self.length, self._ref = int(start), int(length), int(ref)

    @property
    def ref(self) -> int:
        return self._ref & ~RC_MASK

    @property
    def is_rc(self) -> bool:
        return bool(self._ref & RC_MASK)

    def __repr__(self):
        return f"Factor(start={self.start}, length={self.length}, ref={self.ref}, is_rc={self.is_rc})"


class FastaFactorizationResult:
    """bindings.cpp:51-53."""

    def __init__(self, factors, sentinel_factor_indices, sequence_ids=None):
        self.factors = factors
        self.sentinel_factor_indices = sentinel_factor_indices
        self.sequence_ids = sequence_ids if sequence_ids is not None else []


class FastaPerSequenceFactorizationResult:
    """bindings.cpp:1208-1213."""

    def __init__(self, per_sequence_factors, sequence_ids):
        self.per_sequence_factors = per_sequence_factors
        self.sequence_ids = sequence_ids


# ----------------------------------------------------------------------------- helpers
def _buf(data, fn: str):
    """bytes-like, itemsize 1, 1-D (bindings.cpp:58-67) -> (address, n, keepalive)."""
    try:
        mv = memoryview(data)
    except TypeError as e:
        raise TypeError(f"{fn}(): incompatible function arguments: expected a bytes-like object") from e
    if mv.itemsize != 1:
        raise ValueError(f"{fn}: buffer must be a bytes-like object with itemsize==1")
    if mv.ndim != 1:
        raise ValueError(f"{fn}: buffer must be a 1-dimensional bytes-like object")
    if not mv.c_contiguous:
        mv = memoryview(bytes(mv))
    arr = np.frombuffer(mv, dtype=np.uint8)
    return (arr.ctypes.data if arr.size else None), arr.size, arr


def _take(ptr, count: int) -> np.ndarray:
    """Copies `count` triples out of a library-owned buffer and frees it."""
    if count == 0 or not ptr:
        if ptr:
            L.load().nlz_free(ptr)
        return np.zeros((0, 3), dtype=np.uint64)
    try:
        return np.ctypeslib.as_array(ptr, shape=(count * 3,)).copy().reshape(count, 3)
    finally:
        L.load().nlz_free(ptr)


def _take_u64(ptr, count: int) -> list:
    if not ptr:
        return []
    try:
        return [int(x) for x in np.ctypeslib.as_array(ptr, shape=(max(count, 1),))[:count]]
    finally:
        L.load().nlz_free(ptr)


def _plain(arr: np.ndarray) -> list:
    return list(map(tuple, arr.tolist()))


def _with_rc_flag(arr: np.ndarray) -> list:
    """(start, length, ref & ~RC_MASK, is_rc) -- bindings.cpp:226."""
    if len(arr) == 0:
        return []
    ref = arr[:, 2]
    flags = (ref >> np.uint64(63)).astype(bool).tolist()
    clean = (ref & np.uint64(RC_MASK - 1)).tolist()
    return list(zip(arr[:, 0].tolist(), arr[:, 1].tolist(), clean, flags))


def _path(p) -> bytes:
    return os.fsencode(os.fspath(p))


def _text(s) -> bytes:
    """std::string argument: str (UTF-8) or bytes."""
    if isinstance(s, str):
        return s.encode("utf-8")
    return bytes(s)


def _mode_flag(sanitize_mode: str) -> int:
    if sanitize_mode == "remove_ambiguous":
        return 0
    if sanitize_mode == "strict":
        return 1
    raise ValueError("Invalid sanitize_mode. Expected 'remove_ambiguous' or 'strict'.")   # bindings.cpp:29-37


def _factorize(mode: int, data, fn: str, start_pos: int = 0) -> np.ndarray:
    addr, n, keep = _buf(data, fn)
    out, cnt = _U64P(), _U64(0)
    L.check(L.load().nlz_factorize_mode(L.context(), mode, addr, n, start_pos, ctypes.byref(out), ctypes.byref(cnt)))
    return _take(out, cnt.value)


def _count(mode: int, data, fn: str, start_pos: int = 0) -> int:
    addr, n, keep = _buf(data, fn)
    cnt = _U64(0)
    L.check(L.load().nlz_count_mode(L.context(), mode, addr, n, start_pos, ctypes.byref(cnt)))
    return cnt.value


def _factorize_file(mode: int, path, start_pos: int = 0) -> np.ndarray:
    out, cnt = _U64P(), _U64(0)
    L.check(L.load().nlz_factorize_file_mode(L.context(), mode, _path(path), start_pos, ctypes.byref(out), ctypes.byref(cnt)))
    return _take(out, cnt.value)


def _count_file(mode: int, path, start_pos: int = 0) -> int:
    cnt = _U64(0)
    L.check(L.load().nlz_count_file_mode(L.context(), mode, _path(path), start_pos, ctypes.byref(cnt)))
    return cnt.value


def _write_file(mode: int, in_path, out_path, start_pos: int = 0) -> int:
    cnt = _U64(0)
    L.check(L.load().nlz_write_factors_binary_file_mode(L.context(), mode, _path(in_path), _path(out_path), start_pos,
                                                        ctypes.byref(cnt)))
    return cnt.value


# ----------------------------------------------------------------------------- general mode (bindings.cpp:56-202)
def factorize(data):
    return _plain(_factorize(L.MODE_GENERAL, data, "factorize"))


def factorize_file(path, reserve_hint=0):
    return _plain(_factorize_file(L.MODE_GENERAL, path))


def count_factors(data):
    return _count(L.MODE_GENERAL, data, "count_factors")


def count_factors_file(path):
    return _count_file(L.MODE_GENERAL, path)


def write_factors_binary_file(in_path, out_path):
    return _write_file(L.MODE_GENERAL, in_path, out_path)


# ----------------------------------------------------------------------------- DNA with reverse complement (:207-357)
def factorize_dna_w_rc(data):
    return _with_rc_flag(_factorize(L.MODE_DNA_RC, data, "factorize_dna_w_rc"))


def factorize_file_dna_w_rc(path, reserve_hint=0):
    return _with_rc_flag(_factorize_file(L.MODE_DNA_RC, path))


def count_factors_dna_w_rc(data):
    return _count(L.MODE_DNA_RC, data, "count_factors_dna_w_rc")


def count_factors_file_dna_w_rc(path):
    return _count_file(L.MODE_DNA_RC, path)


def write_factors_binary_file_dna_w_rc(in_path, out_path):
    return _write_file(L.MODE_DNA_RC, in_path, out_path)


# ----------------------------------------------------------------------------- prepared multi-sequence text (:361-508)
def factorize_multiple_dna_w_rc(data):
    return _with_rc_flag(_factorize(L.MODE_RC_PREPARED, data, "factorize_multiple_dna_w_rc"))


def factorize_file_multiple_dna_w_rc(path, reserve_hint=0):
    return _with_rc_flag(_factorize_file(L.MODE_RC_PREPARED, path))


def count_factors_multiple_dna_w_rc(data):
    return _count(L.MODE_RC_PREPARED, data, "count_factors_multiple_dna_w_rc")


def count_factors_file_multiple_dna_w_rc(path):
    return _count_file(L.MODE_RC_PREPARED, path)


def write_factors_binary_file_multiple_dna_w_rc(in_path, out_path):
    return _write_file(L.MODE_RC_PREPARED, in_path, out_path)


# ----------------------------------------------------------------------------- prepare (:732-797)
def _prepare(fn, sequences):
    seqs = [_text(s) for s in sequences]
    k = len(seqs)
    ptrs = (ctypes.c_char_p * max(k, 1))(*seqs) if k else (ctypes.c_char_p * 1)()
    lens = (ctypes.c_uint64 * max(k, 1))(*[len(s) for s in seqs]) if k else (ctypes.c_uint64 * 1)()
    prepared, plen, olen = ctypes.c_void_p(), _U64(0), _U64(0)
    sent, nsent = _U64P(), _U64(0)
    L.check(fn(ptrs, lens, k, ctypes.byref(prepared), ctypes.byref(plen), ctypes.byref(olen), ctypes.byref(sent),
               ctypes.byref(nsent)))
    try:
        text = ctypes.string_at(prepared.value, plen.value) if prepared.value else b""
    finally:
        if prepared.value:
            L.load().nlz_free(prepared)
    # pybind11 returns std::string as a Python str; the sentinel bytes above 127 only survive as latin-1
    return text.decode("latin-1"), olen.value, _take_u64(sent, nsent.value)


def prepare_multiple_dna_sequences_w_rc(sequences):
    return _prepare(L.load().nlz_prepare_multiple_dna_sequences_w_rc, sequences)


def prepare_multiple_dna_sequences_no_rc(sequences):
    return _prepare(L.load().nlz_prepare_multiple_dna_sequences_no_rc, sequences)


# ----------------------------------------------------------------------------- concatenated FASTA (:511-729)
def _ids_of(handle) -> list:
    lib = L.load()
    try:
        return [lib.nlz_fasta_id(handle, i).decode("utf-8", "replace") for i in range(lib.nlz_fasta_num_sequences(handle))]
    finally:
        lib.nlz_fasta_free(handle)


def _fasta(ref_fasta, fasta_path, with_rc: bool, sanitize_mode: str, out_path=None):
    flag = _mode_flag(sanitize_mode)
    out, cnt = _U64P(), _U64(0)
    sidx, nsidx = _U64P(), _U64(0)
    ids = ctypes.c_void_p()
    want = out_path is None
    L.check(L.load().nlz_factorize_fasta(
        L.context(), None if ref_fasta is None else _path(ref_fasta), _path(fasta_path), 1 if with_rc else 0, flag,
        None if out_path is None else _path(out_path), ctypes.byref(out) if want else None, ctypes.byref(cnt),
        ctypes.byref(sidx) if want else None, ctypes.byref(nsidx) if want else None,
        ctypes.byref(ids) if want else None))
    if not want:
        return cnt.value
    return _with_rc_flag(_take(out, cnt.value)), _take_u64(sidx, nsidx.value), _ids_of(ids)


def factorize_fasta_multiple_dna_w_rc(fasta_path, sanitize_mode="remove_ambiguous"):
    return _fasta(None, fasta_path, True, sanitize_mode)


def factorize_fasta_multiple_dna_no_rc(fasta_path, sanitize_mode="remove_ambiguous"):
    return _fasta(None, fasta_path, False, sanitize_mode)


def factorize_dna_rc_w_ref_fasta_files(reference_fasta_path, target_fasta_path, sanitize_mode="remove_ambiguous"):
    return _fasta(reference_fasta_path, target_fasta_path, True, sanitize_mode)


def write_factors_binary_file_fasta_multiple_dna_w_rc(fasta_path, out_path, sanitize_mode="remove_ambiguous"):
    return _fasta(None, fasta_path, True, sanitize_mode, out_path)


def write_factors_binary_file_fasta_multiple_dna_no_rc(fasta_path, out_path, sanitize_mode="remove_ambiguous"):
    return _fasta(None, fasta_path, False, sanitize_mode, out_path)


def write_factors_dna_w_reference_fasta_files_to_binary(reference_fasta_path, target_fasta_path, out_path,
                                                        sanitize_mode="remove_ambiguous"):
    return _fasta(reference_fasta_path, target_fasta_path, True, sanitize_mode, out_path)


# ----------------------------------------------------------------------------- reference + target text (:800-975)
def _w_reference(dna: bool, reference_seq, target_seq, out_path=None):
    ref, tgt = _text(reference_seq), _text(target_seq)
    out, cnt = _U64P(), _U64(0)
    L.check(L.load().nlz_factorize_w_reference(L.context(), 1 if dna else 0, ref, len(ref), tgt, len(tgt),
                                               None if out_path is None else _path(out_path),
                                               ctypes.byref(out) if out_path is None else None, ctypes.byref(cnt)))
    if out_path is not None:
        return cnt.value
    arr = _take(out, cnt.value)
    return _with_rc_flag(arr) if dna else _plain(arr)


def factorize_dna_w_reference_seq(reference_seq, target_seq):
    return _w_reference(True, reference_seq, target_seq)


def factorize_dna_w_reference_seq_file(reference_seq, target_seq, out_path):
    return _w_reference(True, reference_seq, target_seq, out_path)


def factorize_w_reference(reference_seq, target_seq):
    return _w_reference(False, reference_seq, target_seq)


def factorize_w_reference_file(reference_seq, target_seq, out_path):
    return _w_reference(False, reference_seq, target_seq, out_path)


# ----------------------------------------------------------------------------- "parallel" entry points (:978-1205)
# num_threads is accepted for signature compatibility; the GPU pipeline evaluates every position in
# parallel, so there is no thread count to choose.
def _parallel_text(mode: int, text, output_path, start_pos: int, fn: str) -> int:
    if isinstance(text, str):
        text = text.encode("utf-8")
    addr, n, keep = _buf(text, fn)
    cnt = _U64(0)
    L.check(L.load().nlz_parallel_factorize_to_file(L.context(), mode, addr, n, _path(output_path), start_pos,
                                                    ctypes.byref(cnt)))
    return cnt.value


def parallel_factorize_to_file(text, output_path, num_threads=0, start_pos=0):
    return _parallel_text(L.MODE_GENERAL, text, output_path, start_pos, "parallel_factorize_to_file")


def parallel_factorize_file_to_file(input_path, output_path, num_threads=0, start_pos=0):
    try:
        with open(os.fspath(input_path), "rb") as f:
            data = f.read()
    except OSError:
        raise RuntimeError(f"Cannot open input file: {os.fspath(input_path)}")
    return _parallel_text(L.MODE_GENERAL, data, output_path, start_pos, "parallel_factorize_file_to_file")


def parallel_factorize_dna_w_rc_to_file(text, output_path, num_threads=0):
    return _parallel_text(L.MODE_DNA_RC, text, output_path, 0, "parallel_factorize_dna_w_rc_to_file")


def parallel_factorize_file_dna_w_rc_to_file(input_path, output_path, num_threads=0):
    try:
        with open(os.fspath(input_path), "rb") as f:
            data = f.read()
    except OSError:
        raise RuntimeError(f"Cannot open input file: {os.fspath(input_path)}")
    return _parallel_text(L.MODE_DNA_RC, data, output_path, 0, "parallel_factorize_file_dna_w_rc_to_file")


def parallel_write_factors_binary_file_fasta_multiple_dna_w_rc(fasta_path, out_path, num_threads=0,
                                                               sanitize_mode="remove_ambiguous"):
    return _fasta(None, fasta_path, True, sanitize_mode, out_path)


def parallel_write_factors_binary_file_fasta_multiple_dna_no_rc(fasta_path, out_path, num_threads=0,
                                                                sanitize_mode="remove_ambiguous"):
    return _fasta(None, fasta_path, False, sanitize_mode, out_path)


def parallel_write_factors_dna_w_reference_fasta_files_to_binary(reference_fasta_path, target_fasta_path, out_path,
                                                                 num_threads=0, sanitize_mode="remove_ambiguous"):
    return _fasta(reference_fasta_path, target_fasta_path, True, sanitize_mode, out_path)


# ----------------------------------------------------------------------------- per-sequence FASTA (:1215-1510)
def _per_sequence(fasta_path, with_rc: bool, sanitize_mode: str, out_dir=None, want_factors=False, num_threads=0):
    flag = _mode_flag(sanitize_mode)
    out, counts, total = _U64P(), _U64P(), _U64(0)
    ids = ctypes.c_void_p()
    L.check(L.load().nlz_factorize_fasta_per_sequence(
        L.context(), _path(fasta_path), 1 if with_rc else 0, flag, None if out_dir is None else _path(out_dir),
        1 if want_factors else 0, int(num_threads), ctypes.byref(out) if want_factors else None, ctypes.byref(counts),
        ctypes.byref(total), ctypes.byref(ids)))
    names = _ids_of(ids)
    cnts = _take_u64(counts, len(names))
    arr = _take(out, total.value) if want_factors else None
    return arr, cnts, names, total.value


def factorize_fasta_dna_w_rc_per_sequence(fasta_path, sanitize_mode="remove_ambiguous"):
    arr, cnts, names, _ = _per_sequence(fasta_path, True, sanitize_mode, want_factors=True)
    bounds = np.cumsum([0] + cnts)
    return [_with_rc_flag(arr[bounds[i]:bounds[i + 1]]) for i in range(len(cnts))], names


def factorize_fasta_dna_no_rc_per_sequence(fasta_path, sanitize_mode="remove_ambiguous"):
    arr, cnts, names, _ = _per_sequence(fasta_path, False, sanitize_mode, want_factors=True)
    bounds = np.cumsum([0] + cnts)
    return [_with_rc_flag(arr[bounds[i]:bounds[i + 1]]) for i in range(len(cnts))], names


def count_factors_fasta_dna_w_rc_per_sequence(fasta_path, sanitize_mode="remove_ambiguous"):
    _, cnts, names, total = _per_sequence(fasta_path, True, sanitize_mode)
    return cnts, names, total


def count_factors_fasta_dna_no_rc_per_sequence(fasta_path, sanitize_mode="remove_ambiguous"):
    _, cnts, names, total = _per_sequence(fasta_path, False, sanitize_mode)
    return cnts, names, total


def write_factors_binary_file_fasta_dna_w_rc_per_sequence(fasta_path, out_dir, sanitize_mode="remove_ambiguous"):
    return _per_sequence(fasta_path, True, sanitize_mode, out_dir=out_dir)[3]


def write_factors_binary_file_fasta_dna_no_rc_per_sequence(fasta_path, out_dir, sanitize_mode="remove_ambiguous"):
    return _per_sequence(fasta_path, False, sanitize_mode, out_dir=out_dir)[3]


def parallel_write_factors_binary_file_fasta_dna_w_rc_per_sequence(fasta_path, out_dir, num_threads=0,
                                                                   sanitize_mode="remove_ambiguous"):
    return _per_sequence(fasta_path, True, sanitize_mode, out_dir=out_dir, num_threads=num_threads)[3]


def parallel_write_factors_binary_file_fasta_dna_no_rc_per_sequence(fasta_path, out_dir, num_threads=0,
                                                                    sanitize_mode="remove_ambiguous"):
    return _per_sequence(fasta_path, False, sanitize_mode, out_dir=out_dir, num_threads=num_threads)[3]
