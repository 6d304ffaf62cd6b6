"""ncu target for the chromosome-scale kernels: one count-only factorization of a 60 Mbp text with the configs[3] recipe
(planted repeats scaled x12: hybrid doubling rounds, deep-nesting positions), RC mode, single GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nolzss_b200 import _lib as L, workloads as wl

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000_000
t = wl.planted_dna(n, 4, scale=n / 5e6).tobytes()
z = L.count(L.MODE_DNA_RC, t)
st = L.stats()
print("factors", z, {k: round(v, 2) for k, v in st.items() if k.startswith("ms_")}, "rounds", st["doubling_rounds"])
