"""Low-complexity inputs: time and check against the oracle."""
import os, sys, time
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")]
import numpy as np
import oracle_py as orc
from nolzss_b200 import _lib as L, workloads as wl

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(5)
cases = {"A^n": b"A" * n, "(AC)^n/2": b"AC" * (n // 2), "fib-like": None, "binary": np.frombuffer(b"AC", dtype=np.uint8)[rng.integers(0, 2, n)].tobytes(),
         "A^n/2 C A^n/2": b"A" * (n // 2) + b"C" + b"A" * (n // 2)}
a, b = b"A", b"AB"
while len(b) < n:
    a, b = b, b + a
cases["fib-like"] = b[:n].replace(b"B", b"C")
for name, s in cases.items():
    for mode, mname in ((L.MODE_GENERAL, "general"), (L.MODE_DNA_RC, "rc")):
        t0 = time.perf_counter()
        f = L.factorize_array(mode, s)
        dt = time.perf_counter() - t0
        st = L.stats()
        exp = orc.factorize(s) if mode == L.MODE_GENERAL else orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(s))
        print(f"{name:14s} {mname:8s} n={len(s)} z={len(f)} {'OK' if np.array_equal(f, exp) else 'MISMATCH'} wall={dt*1e3:.1f} ms dev={st['ms_total']:.1f} "
              f"(prep {st['ms_prepare']:.1f} keys {st['ms_keys']:.1f} sort0 {st['ms_sort0']:.1f} doubling {st['ms_doubling']:.1f} lcp {st['ms_lcp']:.1f} lpnf {st['ms_lpnf']:.1f} chain {st['ms_chain']:.1f}) rounds={st['doubling_rounds']} hard={st['hard_positions']}", flush=True)
