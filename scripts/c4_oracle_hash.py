"""Offline (CPU, ~10 GB host RAM): the oracle on the 250 Mbp RC text of configs[3]; writes the factor count and the
sha256 of the (z,3) uint64 little-endian triples to tests/golden/c4_250mbp_rc.json.  bench.py's single-text leg and
tests/test_gpu_parity.py compare the GPU output with this hash at every N."""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as orc  # noqa: E402
from nolzss_b200 import workloads as wl  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000_000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 4
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 50.0
out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "tests", "golden", "c4_250mbp_rc.json")
t0 = time.time()
t = wl.planted_dna(n, seed, scale=scale).tobytes()
S = wl.prepare_w_rc_single(t)
del t
print(f"text ready after {time.time() - t0:.1f} s", flush=True)
t0 = time.time()
f = orc.factorize_multiple_dna_w_rc(S)
dt = time.time() - t0
h = hashlib.sha256(f.astype("<u8").tobytes()).hexdigest()
rec = {"workload": f"planted_dna({n}, {seed}, scale={scale}), RC mode (S = T s0 rc(T) s1)", "n_bases": n, "factors": int(len(f)),
       "sha256_triples_le_u64": h, "oracle_seconds": dt, "sum_lengths": int(f[:, 1].sum()),
       "rc_factors": int((f[:, 2] >> 63).sum())}
with open(out, "w") as fh:
    json.dump(rec, fh, indent=1)
print(json.dumps(rec), flush=True)
