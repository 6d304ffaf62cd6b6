"""Development aid (GPU): sha256 of the factor arrays of a set of texts (general / RC, tandem-heavy, degenerate) with the
stage-3 time of each -- run once per build or per debug switch and diff the hashes (profiles/r2_stage3_rework.md)."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

from nolzss_b200 import _lib as L, workloads as wl

texts = []
texts.append(("c1", wl.c1_text(), L.MODE_GENERAL))
texts.append(("c2", wl.c2_text(), L.MODE_DNA_RC))
texts.append(("planted20M", wl.planted_dna(20_000_000, 4, scale=4.0).tobytes(), L.MODE_DNA_RC))
texts.append(("planted20M_general", wl.planted_dna(20_000_000, 7, scale=4.0).tobytes(), L.MODE_GENERAL))
x = wl.uniform_dna(3_000_000, 3).copy()
x[500_000:1_500_000] = ord("A")
x[2_000_000:2_600_000] = np.frombuffer(b"ACG" * 200_000, dtype=np.uint8)
texts.append(("tandem_rc", x.tobytes(), L.MODE_DNA_RC))
texts.append(("tandem_general", x.tobytes(), L.MODE_GENERAL))
texts.append(("at_rc", b"AT" * 400_000, L.MODE_DNA_RC))
texts.append(("a_rc", b"A" * 300_000, L.MODE_DNA_RC))
texts.append(("acgt_small", b"ACGTTGCA" * 11 + b"GATTACA", L.MODE_DNA_RC))
rng = np.random.default_rng(3)
texts.append(("binary_general", bytes(rng.integers(0, 2, 2_000_000, dtype=np.uint8) + 65), L.MODE_GENERAL))
for name, t, mode in texts:
    f = L.factorize_array(mode, t)
    st = L.stats()
    print(name, len(f), hashlib.sha256(np.ascontiguousarray(f).tobytes()).hexdigest()[:16], f"lpnf={st['ms_lpnf']:.2f}ms hard={st['hard_positions']}", flush=True)
