"""Offline (CPU): the oracle on every record of configs[2] (10 000 records x 10 kbp, RC mode, seed 3); writes the total
factor count and sha256 hashes of the per-record counts (int64 LE) and of all record-local triples (uint64 LE,
record order) to tests/golden/c3_10000x10k_rc.json.  bench.py's configs[2] leg checks the GPU result against it at
every N."""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import oracle_py as orc  # noqa: E402
from nolzss_b200 import workloads as wl  # noqa: E402

nrec = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
reclen = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
t0 = time.time()
recs = wl.c3_records(nrec, reclen, seed=3)
counts = np.zeros(nrec, dtype=np.int64)
h = hashlib.sha256()
for j, (_, s) in enumerate(recs):
    f = orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(s))
    counts[j] = len(f)
    h.update(f.astype("<u8").tobytes())
rec = {"workload": f"c3_records({nrec}, {reclen}, seed=3), per-record RC mode", "records": nrec, "record_length": reclen,
       "total_factors": int(counts.sum()), "sha256_counts_le_i64": hashlib.sha256(counts.astype("<i8").tobytes()).hexdigest(),
       "sha256_triples_le_u64": h.hexdigest(), "oracle_seconds": time.time() - t0}
with open(os.path.join(ROOT, "tests", "golden", f"c3_{nrec}x{reclen // 1000}k_rc.json"), "w") as fh:
    json.dump(rec, fh, indent=1)
print(json.dumps(rec))
