"""Prints the per-stage device times of one pipeline run (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nolzss_b200 import _lib as L, workloads as wl

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
if which == "c1":
    t, mode = wl.c1_text(), L.MODE_GENERAL
elif which == "c2":
    t, mode = wl.c2_text(), L.MODE_DNA_RC
else:
    n = int(which); t, mode = wl.planted_dna(n, 4, scale=max(1.0, n / 5e6)).tobytes(), L.MODE_DNA_RC
L.set_profiling(True)
for it in range(3):
    t0 = time.perf_counter()
    z = L.count(mode, t)
    dt = time.perf_counter() - t0
s = L.stats()
print(f"{which}: wall {dt*1e3:.2f} ms z={z} total={s['ms_total']:.2f} doubling={s['ms_doubling']:.2f} lcp={s['ms_lcp']:.2f} "
      f"lpnf={s['ms_lpnf']:.2f} nodes={s['walk_nodes']} hard={s['hard_positions']} rounds={s['doubling_rounds']} "
      f"tile_rounds={s['tile_sort_rounds']} launches={s['kernel_launches']}")
print("   " + "  ".join(f"{k}={v['ms']:.3f}" for k, v in L.kernel_stats().items() if v['launches']))
