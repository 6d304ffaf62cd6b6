"""TEST INFRASTRUCTURE ONLY: ctypes loader for the CPU oracle (oracle/libnlz_oracle.so).

Import this from tests/, __graft_entry__.smoke() and bench.py's CPU legs only.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnlz_oracle.so")
RC_MASK = 1 << 63


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "nolzss_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libnlz_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        u8p = ctypes.c_char_p
        u64 = ctypes.c_uint64
        pp = ctypes.POINTER(ctypes.POINTER(ctypes.c_uint64))
        for name in ("nlzo_factorize", "nlzo_factorize_multiple_dna_w_rc"):
            f = getattr(L, name)
            f.argtypes = [u8p, u64, u64, pp, ctypes.POINTER(u64)]
            f.restype = ctypes.c_int
        L.nlzo_parallel_factorize_dna_w_rc.argtypes = [u8p, u64, u64, ctypes.c_int, pp, ctypes.POINTER(u64),
                                                       ctypes.POINTER(ctypes.c_int)]
        L.nlzo_parallel_factorize_dna_w_rc.restype = ctypes.c_int
        L.nlzo_free.argtypes = [ctypes.c_void_p]
        L.nlzo_free.restype = None
        L.nlzo_suffix_array_i32.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]
        L.nlzo_suffix_array_i32.restype = ctypes.c_int
        L.nlzo_sa_lcp_bytes.argtypes = [u8p, u64, ctypes.c_void_p, ctypes.c_void_p]
        L.nlzo_sa_lcp_bytes.restype = ctypes.c_int
        L.nlzo_lcp_from_sa_i32.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
        L.nlzo_lcp_from_sa_i32.restype = ctypes.c_int
        L.nlzo_last_timing.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        L.nlzo_last_timing.restype = None
        _lib = L
    return _lib


def _run(fn, data: bytes, start_pos: int) -> np.ndarray:
    data = bytes(data)
    out = ctypes.POINTER(ctypes.c_uint64)()
    cnt = ctypes.c_uint64(0)
    rc = fn(data, len(data), start_pos, ctypes.byref(out), ctypes.byref(cnt))
    if rc == 1:
        raise ValueError("start_pos must be less than the original sequence length")
    if rc != 0:
        raise RuntimeError(f"oracle failed rc={rc}")
    n = cnt.value
    if n == 0:
        return np.zeros((0, 3), dtype=np.uint64)
    arr = np.ctypeslib.as_array(out, shape=(n * 3,)).copy().reshape(n, 3)
    lib().nlzo_free(out)
    return arr


def factorize(data: bytes, start_pos: int = 0) -> np.ndarray:
    """(z,3) uint64 triples, general mode (factorizer_core.hpp:51-119)."""
    return _run(lib().nlzo_factorize, data, start_pos)


def factorize_multiple_dna_w_rc(S: bytes, start_pos: int = 0) -> np.ndarray:
    """(z,3) uint64 triples with RC_MASK in ref (factorizer_core.hpp:177-383)."""
    return _run(lib().nlzo_factorize_multiple_dna_w_rc, S, start_pos)


def parallel_factorize_multiple_dna_w_rc(S: bytes, num_threads: int, start_pos: int = 0):
    """The reference's CPU parallel mode (parallel_factorizer.cpp:849-984): serial index build, chunked chain walk
    on `num_threads` threads, convergence merge.  Returns ((z,3) triples, threads actually used)."""
    used = ctypes.c_int(0)

    def fn(data, n, sp, out, cnt):
        return lib().nlzo_parallel_factorize_dna_w_rc(data, n, sp, int(num_threads), out, cnt, ctypes.byref(used))

    arr = _run(fn, S, start_pos)
    return arr, used.value


def last_timing():
    a, b = ctypes.c_double(0), ctypes.c_double(0)
    lib().nlzo_last_timing(ctypes.byref(a), ctypes.byref(b))
    return a.value, b.value


def suffix_array_i32(s: np.ndarray, K: int) -> np.ndarray:
    """SA of an int32 string whose last symbol is a unique 0."""
    s = np.ascontiguousarray(s, dtype=np.int32)
    sa = np.empty(len(s), dtype=np.int32)
    rc = lib().nlzo_suffix_array_i32(s.ctypes.data, len(s), K, sa.ctypes.data)
    assert rc == 0
    return sa


def sa_lcp_bytes(data: bytes):
    data = bytes(data)
    n1 = len(data) + 1
    sa = np.empty(n1, dtype=np.int32)
    lcp = np.empty(n1 + 1, dtype=np.int32)
    rc = lib().nlzo_sa_lcp_bytes(data, len(data), sa.ctypes.data, lcp.ctypes.data)
    assert rc == 0
    return sa, lcp


def lcp_from_sa_i32(s: np.ndarray, sa: np.ndarray) -> np.ndarray:
    s = np.ascontiguousarray(s, dtype=np.int32)
    sa = np.ascontiguousarray(sa, dtype=np.int32)
    lcp = np.zeros(len(s), dtype=np.int32)
    rc = lib().nlzo_lcp_from_sa_i32(s.ctypes.data, len(s), sa.ctypes.data, lcp.ctypes.data)
    assert rc == 0
    return lcp


def gpu_order_recode(data: bytes):
    """int32 recoding of data·$ under the symbol order the CUDA path sorts by (csrc/sa.cuh):
    bytes that occur once, and the terminator, are sentinel symbols ordered by text position and
    smaller than every repeated byte.  A real smallest terminator 0 is appended for SA-IS.
    Returns (s, K)."""
    x = np.frombuffer(bytes(data), dtype=np.uint8)
    hist = np.bincount(x, minlength=256)
    dense = np.full(256, -1, dtype=np.int64)
    rep = np.nonzero(hist >= 2)[0]
    dense[rep] = np.arange(len(rep))
    codes = dense[x]
    sent_pos = np.nonzero(codes < 0)[0]
    S = len(sent_pos) + 1                      # + virtual terminator at position len(x)
    out = np.empty(len(x) + 2, dtype=np.int32)
    out[: len(x)] = codes + (S + 1)
    out[sent_pos] = np.arange(1, len(sent_pos) + 1)
    out[len(x)] = S
    out[len(x) + 1] = 0
    return out, int(S + 1 + len(rep))


def gpu_order_sa_lcp(data: bytes):
    """(SA, LCP) of data·$ in the CUDA path's symbol order: SA has len+1 entries, LCP len+2
    (LCP[0] = LCP[len+1] = 0)."""
    s, K = gpu_order_recode(data)
    sa_full = suffix_array_i32(s, K)
    assert sa_full[0] == len(s) - 1
    sa = sa_full[1:]
    lcp_full = lcp_from_sa_i32(s, sa_full)
    lcp = np.zeros(len(sa) + 1, dtype=np.int32)
    lcp[1 : len(sa)] = lcp_full[2:]
    return sa, lcp
