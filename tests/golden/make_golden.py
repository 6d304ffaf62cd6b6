"""Regenerates tests/golden/golden_factors.json.

The reference implementation itself cannot be built in this environment (its core includes
<sdsl/...> from sdsl-lite v3.0.3, fetched from the network by CMake), so the vectors are produced by
the CPU oracle (oracle/nolzss_oracle.c), which is pinned on the reference's own known-answer tests
and on the literal tree-walk model (tests/test_oracle.py).  Run from the repo root:

    python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import oracle_py as orc  # noqa: E402
import treewalk_model as tm  # noqa: E402
from nolzss_b200 import workloads as wl  # noqa: E402


def triples(a):
    return [[int(x) for x in r] for r in a]


def main():
    rnd = random.Random(20261018)
    cases = []
    # general mode: DNA, protein-like, text, runs
    texts = [b"abracadabra", b"A" * 200, b"AC" * 150, b"mississippi$mississippi",
             (b"the quick brown fox jumps over the lazy dog " * 12)]
    for sig, n in [(4, 300), (2, 400), (4, 2000), (20, 1500), (3, 900)]:
        alpha = b"ACGTDEFHIKLMNPQRSVWY"[:sig]
        texts.append(bytes(rnd.choice(alpha) for _ in range(n)))
    texts.append(wl.planted_dna(6000, 99, scale=0.02).tobytes())
    for t in texts:
        cases.append({"mode": "general", "text": t.decode("latin-1"), "start_pos": 0, "factors": triples(orc.factorize(t))})
    t = texts[7]
    cases.append({"mode": "general", "text": t.decode("latin-1"), "start_pos": 777, "factors": triples(orc.factorize(t, 777))})
    # RC mode on single sequences (factorize_dna_w_rc) and prepared multi-record texts
    dnas = [b"AC", b"ATGCAT", b"ATCGATCG", b"CAAGCACCACCGCGGCGACCGAGGCA", b"A" * 120, b"AT" * 90]
    for n in (250, 1200, 5000):
        dnas.append(wl.planted_dna(n, n, scale=0.02).tobytes())
    for d in dnas:
        S = wl.prepare_w_rc_single(d)
        cases.append({"mode": "dna_rc", "text": d.decode("latin-1"), "start_pos": 0,
                      "factors": triples(orc.factorize_multiple_dna_w_rc(S))})
    recs = [bytes(rnd.choice(b"ACGT") for _ in range(rnd.randint(50, 400))) for _ in range(4)]
    recs[2] = recs[0][10:120] + recs[2]
    S, ol, sent = tm.prepare_multiple_dna_sequences_w_rc(recs)
    for sp in (0, len(recs[0]) + 1):
        cases.append({"mode": "rc_prepared", "text": S.decode("latin-1"), "start_pos": sp,
                      "factors": triples(orc.factorize_multiple_dna_w_rc(S, sp))})
    out = os.path.join(ROOT, "tests", "golden", "golden_factors.json")
    with open(out, "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "cases": cases}, f, separators=(",", ":"))
    print(out, len(cases), "cases", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
