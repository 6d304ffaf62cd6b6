"""Short single-GPU run used as the ncu target: two C2 factorizations (RC mode, count only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nolzss_b200 import _lib as L, workloads as wl

t = wl.c2_text()
for _ in range(2):
    z = L.count(L.MODE_DNA_RC, t)
print("factors", z, L.stats()["ms_total"])
