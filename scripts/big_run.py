"""Single-GPU run of a chromosome-sized text (configs[3] at N=1): timing + size-independent checks."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nolzss_b200 import _lib as L, workloads as wl

n = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000_000
mode = L.MODE_DNA_RC if (len(sys.argv) < 3 or sys.argv[2] == "rc") else L.MODE_GENERAL
t0 = time.time()
x = wl.planted_dna(n, 4, scale=10.0)
print(f"generated {n} bases in {time.time()-t0:.1f}s", flush=True)
t = x.tobytes()
for it in range(2):
    t0 = time.perf_counter()
    f = L.factorize_array(mode, t)
    dt = time.perf_counter() - t0
    s = L.stats()
    print(f"iter {it}: wall {dt*1e3:.1f} ms device {s['ms_total']:.1f} ms  z={len(f)}  rounds={s['doubling_rounds']} tile_rounds={s['tile_sort_rounds']} "
          f"hard={s['hard_positions']} ws={s['workspace_bytes']/1e9:.1f} GB  Mbases/s={n/ (s['ms_total']*1e-3)/1e6:.1f}", flush=True)
    print("   stages:", {k: round(v, 1) for k, v in s.items() if k.startswith('ms_')}, flush=True)
# properties: coverage, order, validity of sampled factors
st, ln, rf = f[:, 0].astype(np.int64), f[:, 1].astype(np.int64), f[:, 2]
assert st[0] == 0 and np.all(st[1:] == st[:-1] + ln[:-1]) and st[-1] + ln[-1] == n, "coverage"
rng = np.random.default_rng(0)
idx = rng.integers(0, len(f), 20000)
bad = 0
for k in idx:
    s0, l0, r0 = int(st[k]), int(ln[k]), int(rf[k])
    if r0 >> 63:
        r0 &= (1 << 63) - 1
        ok = r0 + l0 <= s0 and np.array_equal(wl.revcomp(x[r0:r0 + l0]), x[s0:s0 + l0])
    elif r0 == s0:
        ok = l0 == 1
    else:
        ok = r0 + l0 <= s0 and np.array_equal(x[r0:r0 + l0], x[s0:s0 + l0])
    bad += 0 if ok else 1
print("sampled factor validity: bad =", bad, "of", len(idx))
# maximality of sampled forward/RC factors: the next base must not extend the same source match
print("ok" if bad == 0 else "FAILED")
