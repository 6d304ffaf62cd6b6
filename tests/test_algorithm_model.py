"""CPU: the Python model of the GPU algorithms (tests/gpu_algorithm_model.py) against the oracle."""
import random

import numpy as np

import gpu_algorithm_model as gm
import oracle_py as orc
import treewalk_model as tm
from kats import GENERAL_KATS, RC_KATS, plain_tuples, rc_tuples
from nolzss_b200 import workloads as wl


def test_model_kats():
    for text, exp in GENERAL_KATS.items():
        assert gm.factorize_model(text, "general") == exp
    for text, exp in RC_KATS.items():
        got = gm.factorize_model(wl.prepare_w_rc_single(text), "rc_prepared")
        assert rc_tuples(got) == exp


def test_model_random_small():
    rnd = random.Random(5)
    for it in range(400):
        sig = rnd.randint(1, 4)
        s = bytes(rnd.choice(b"ACGT"[:sig]) for _ in range(rnd.randint(1, 80)))
        fb = rnd.choice([None, 32, 64])
        ch = rnd.choice([4, 16, 1024])
        assert gm.factorize_model(s, "general", 0, fb, ch) == plain_tuples(orc.factorize(s))
        seqs = [bytes(rnd.choice(b"ACGT"[:sig]) for _ in range(rnd.randint(1, 40))) for _ in range(rnd.randint(1, 3))]
        S, _, _ = tm.prepare_multiple_dna_sequences_w_rc(seqs)
        N = len(S) // 2 - 1
        sp = rnd.randint(0, N - 1) if it % 2 else 0
        assert gm.factorize_model(S, "rc_prepared", sp, fb, ch) == plain_tuples(orc.factorize_multiple_dna_w_rc(S, sp))


def test_model_summary_trees_and_repeats():
    for n, sig, seed in [(6000, 4, 1), (9000, 1, 2), (12000, 2, 3)]:
        rng = np.random.default_rng(seed)
        x = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, sig, n)].copy()
        for _ in range(6):
            L = int(rng.integers(50, 500)); a = int(rng.integers(0, n - L)); b = int(rng.integers(0, n - L))
            x[b:b + L] = x[a:a + L]
        s = x.tobytes()
        assert gm.factorize_model(s, "general") == plain_tuples(orc.factorize(s))
        S = wl.prepare_w_rc_single(s[: n // 2])
        assert gm.factorize_model(S, "rc_prepared") == plain_tuples(orc.factorize_multiple_dna_w_rc(S))


def test_gpu_symbol_order_recode_matches_model():
    rnd = random.Random(9)
    for it in range(100):
        s = bytes(rnd.choice(b"ACGTXY"[: rnd.randint(1, 6)]) for _ in range(rnd.randint(1, 150)))
        if it % 3 == 0:
            s = s + b"\x01" + s[::-1] + b"\x02"
        sa, lcp = orc.gpu_order_sa_lcp(s)
        SA, RANK, _ = gm.suffix_array(s, rnd.choice([None, 32, 64]))
        assert list(sa) == SA and list(lcp) == gm.lcp_array(s, SA, RANK)


def test_model_hybrid_rounds_and_representative_ranks():
    """Model of the hybrid doubling rounds (csrc/big_groups.cuh) and of ranks-as-representatives: tandem-heavy texts
    with tiny tile / outlier capacities so that the split, the pivot partition, the S/B routing, the renaming rule
    and the redo path all run; the suffix array must be the oracle's."""
    rnd = random.Random(12)
    seen = dict(hybrid_rounds=0, fallbacks=0, stream_groups=0, kept=0, renamed=0)
    for it in range(120):
        parts = []
        for _p in range(rnd.randint(1, 4)):
            unit = bytes(rnd.choice(b"ACGT") for _u in range(rnd.randint(1, 5)))
            parts.append(unit * rnd.randint(3, 60))
            parts.append(bytes(rnd.choice(b"ACGT") for _u in range(rnd.randint(0, 12))))
        s = b"".join(parts)
        if it % 4 == 0:
            s = wl.prepare_w_rc_single(s)
        gcap, ocap = rnd.choice([(4, 6), (8, 16), (8, 3), (16, 64), (3, 1)])
        SA, RANK, _, st = gm.suffix_array_hybrid(s, gcap, ocap, rnd.choice([None, 32, 64]), rnd)
        sa, _ = orc.gpu_order_sa_lcp(s)
        assert SA == list(sa), (it, gcap, ocap, s)
        assert all(RANK[SA[k]] == k for k in range(len(SA)))
        for k in seen:
            seen[k] += st[k]
    assert all(v > 0 for v in seen.values()), seen
