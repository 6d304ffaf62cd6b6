"""Small workload for compute-sanitizer: all modes, edge sizes, the fallback paths."""
import os, sys
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
import numpy as np
from nolzss_b200 import _lib as L, workloads as wl

rng = np.random.default_rng(1)
cases = [b"A", b"AC", b"ACGT" * 7, b"A" * 3000, wl.planted_dna(60_000, 3, scale=0.05).tobytes(),
         (b"the quick brown fox jumps over the lazy dog " * 120) + b"!", rng.integers(1, 256, 20_000, dtype=np.uint8).tobytes(),
         wl.uniform_dna(33_333, 9).tobytes()]
tot = 0
for flags in (0, 1, 2):
    L.check(L.load().nlz_set_debug_flags(L.context(), flags))
    for s in cases:
        tot += len(L.factorize_array(L.MODE_GENERAL, s))
        if set(s) <= set(b"ACGT"):
            tot += len(L.factorize_array(L.MODE_DNA_RC, s))
            tot += L.count(L.MODE_RC_PREPARED, wl.prepare_w_rc_single(s))
print("done", tot)
