import os, sys
sys.path.insert(0, "/root/repo")
from nolzss_b200 import _lib as L, workloads as wl
t = wl.c2_text()
L.count(L.MODE_DNA_RC, t)
os.environ["NLZ_TRACE"] = "1"
L.count(L.MODE_DNA_RC, t)
print(L.stats())
