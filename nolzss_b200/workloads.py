"""Seeded synthetic inputs for the configurations named in BASELINE.json.

Shared by tests/ and bench.py so that parity cases and bench lines use the same bytes.
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b


def revcomp(x: np.ndarray) -> np.ndarray:
    return _COMP[x[::-1]]


def uniform_dna(n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return _ACGT[rng.integers(0, 4, n)]


def planted_dna(n: int, seed: int, scale: float = 1.0, families: int = 20, tandems: int = 40) -> np.ndarray:
    """Uniform background + interspersed repeat families (1 % substitutions, 30 % reverse-complemented
    copies) + tandem arrays.  `scale` stretches family and tandem lengths (C4 uses 50x of C2's)."""
    rng = np.random.default_rng(seed)
    x = _ACGT[rng.integers(0, 4, n)].copy()
    for _ in range(families):
        flen = int(rng.integers(500, 10_000) * scale)
        flen = max(8, min(flen, n // 8))
        copies = int(rng.integers(5, 51))
        unit = _ACGT[rng.integers(0, 4, flen)]
        for _c in range(copies):
            cp = unit.copy()
            nsub = flen // 100
            if nsub:
                pos = rng.integers(0, flen, nsub)
                cp[pos] = _ACGT[rng.integers(0, 4, nsub)]
            if rng.random() < 0.3:
                cp = revcomp(cp)
            at = int(rng.integers(0, n - flen))
            x[at:at + flen] = cp
    for _ in range(tandems):
        period = int(rng.integers(2, 201))
        copies = int(rng.integers(10, 501))
        total = int(min(period * copies * scale, n // 8))
        unit = _ACGT[rng.integers(0, 4, period)]
        arr = np.tile(unit, total // period + 1)[:total]
        at = int(rng.integers(0, n - total))
        x[at:at + total] = arr
    return x


def c1_text() -> bytes:
    """configs[0]: 1 Mbp uniform random ACGT, general mode (noLZSS.factorize)."""
    return uniform_dna(1_000_000, 1).tobytes()


def c2_text(n: int = 5_000_000, seed: int = 2) -> bytes:
    """configs[1]: 5 Mbp bacterial-genome-sized DNA with planted repeats, RC mode."""
    return planted_dna(n, seed).tobytes()


def c3_records(nrec: int = 10_000, reclen: int = 10_000, seed: int = 3):
    """configs[2]: multi-record FASTA, each record uniform with a 500-bp segment copied inside it
    (50 % reverse-complemented)."""
    rng = np.random.default_rng(seed)
    recs = []
    for r in range(nrec):
        x = _ACGT[rng.integers(0, 4, reclen)].copy()
        seg = min(500, reclen // 4)
        if seg > 0:
            a = int(rng.integers(0, reclen - seg))
            b = int(rng.integers(0, reclen - seg))
            piece = x[a:a + seg].copy()
            if rng.random() < 0.5:
                piece = revcomp(piece)
            x[b:b + seg] = piece
        recs.append((f"rec{r:06d}", x.tobytes()))
    return recs


def prepare_w_rc_single(t: bytes) -> bytes:
    """S = T s0 rc(T) s1 for one record (factorizer.cpp:54-172 with k = 1)."""
    x = np.frombuffer(t, dtype=np.uint8)
    return t + b"\x01" + revcomp(x).tobytes() + b"\x02"
