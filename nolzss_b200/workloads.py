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


def planted_dna_big(n: int, seed: int, scale: float = 50.0, families: int = 20, tandems: int = 40, out: np.ndarray | None = None,
                    chunk: int = 1 << 27) -> np.ndarray:
    """Genome-scale variant of `planted_dna` (configs[4]: n = 3.1 * 10^9): the same recipe, but the uniform background
    is drawn in chunks as uint8 (numpy's default int64 draw would need 8 bytes per base) and may be written straight
    into `out` (e.g. a memory-mapped file shared by the ranks of a box).  Not byte-identical to `planted_dna` for the
    same seed (different draw width)."""
    rng = np.random.default_rng(seed)
    x = out if out is not None else np.empty(n, dtype=np.uint8)
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        x[a:b] = _ACGT[rng.integers(0, 4, b - a, dtype=np.uint8)]
    for _ in range(families):
        flen = int(rng.integers(500, 10_000) * scale)
        flen = max(8, min(flen, n // 8))
        copies = int(rng.integers(5, 51))
        unit = _ACGT[rng.integers(0, 4, flen, dtype=np.uint8)]
        for _c in range(copies):
            cp = unit.copy()
            nsub = flen // 100
            if nsub:
                pos = rng.integers(0, flen, nsub)
                cp[pos] = _ACGT[rng.integers(0, 4, nsub, dtype=np.uint8)]
            if rng.random() < 0.3:
                cp = revcomp(cp)
            at = int(rng.integers(0, n - flen))
            x[at:at + flen] = cp
    for _ in range(tandems):
        period = int(rng.integers(2, 201))
        copies = int(rng.integers(10, 501))
        total = int(min(period * copies * scale, n // 8))
        unit = _ACGT[rng.integers(0, 4, period, dtype=np.uint8)]
        arr = np.tile(unit, total // period + 1)[:total]
        at = int(rng.integers(0, n - total))
        x[at:at + total] = arr
    return x


C5_BASES = 3_100_000_000


def c5_text_into(out: np.ndarray, n: int = C5_BASES, seed: int = 5) -> np.ndarray:
    """configs[4]: human-genome-sized synthetic DNA.  Repeat sizes as in configs[3] (families up to 500 kbp, tandem
    arrays up to 5 Mbp: scale 50), their number scaled with the text (12.4 x as many)."""
    k = max(1, round(n / 250_000_000))
    return planted_dna_big(n, seed, scale=50.0, families=20 * k, tandems=40 * k, out=out)


def verify_factors_sample(text: np.ndarray, f: np.ndarray, samples: int = 100_000, seed: int = 0):
    """Size-independent checks of an RC-mode factorization of one record: the factors tile [0, n) and `samples`
    randomly chosen factors are true matches -- forward: T[ref, ref+len) == T[start, start+len) with ref + len <= start;
    reverse complement: revcomp(T[ref, ref+len)) == T[start, start+len) with ref + len <= start (the source ends
    before the factor starts); literals are (p, 1, p).  Returns a dict of counts; raises AssertionError on the first violation."""
    n = len(text)
    start, length, ref = f[:, 0].astype(np.int64), f[:, 1].astype(np.int64), f[:, 2]
    assert start[0] == 0 and np.all(start[1:] == start[:-1] + length[:-1]) and start[-1] + length[-1] == n, "factors do not tile the text"
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, len(f), min(samples, len(f)))
    nf = nr = nl = 0
    rc_mask = np.uint64(1 << 63)
    for k in idx:
        s, ln = int(start[k]), int(length[k])
        is_rc = bool(ref[k] & rc_mask)
        rf = int(ref[k] & ~rc_mask)
        if not is_rc and rf == s:
            assert ln == 1, ("literal", k)
            nl += 1
            continue
        assert rf + ln <= s, ("source overlaps the factor", k, s, ln, rf)
        sub = text[s:s + ln]
        src = text[rf:rf + ln]
        if is_rc:
            assert np.array_equal(sub, revcomp(src)), ("rc mismatch", k)
            nr += 1
        else:
            assert np.array_equal(sub, src), ("forward mismatch", k)
            nf += 1
    return {"factors": int(len(f)), "sampled": int(len(idx)), "forward": nf, "rc": nr, "literals": nl}
