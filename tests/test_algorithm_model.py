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


def test_key_derived_lcp_and_marks():
    """sa.cuh key_pair_lcp / LcpSeed: outside tie groups the LCP read off adjacent sort keys IS the LCP; every other slot
    is a tie-group member whose TEXT POSITION is marked; Kasai over the marked positions alone completes the array.
    Short windows (force_w) push most suffixes into tie groups, long ones almost none; texts with sentinels inside the
    window (unique bytes) exercise the offset field."""
    rnd = random.Random(21)
    for it in range(300):
        sig = rnd.randint(1, 4)
        s = bytes(rnd.choice(b"ACGT"[:sig]) for _ in range(rnd.randint(1, 120)))
        if it % 3 == 0:
            s = s + b"\x01" + s[::-1] + b"\x02"
        if it % 7 == 0:
            s = s[: len(s) // 2] + b"N" + s[len(s) // 2:]       # a unique byte mid-text: sentinel class
        sa, lcp = orc.gpu_order_sa_lcp(s)
        so = {}
        SA, RANK, _ = gm.suffix_array(s, rnd.choice([None, 32, 64]), rnd.choice([None, 1, 2, 3, 5]), seed_out=so)
        assert list(sa) == SA
        seeded, need = so["seed"]
        for r in range(len(SA)):
            if seeded[r] is None:
                pass                                             # marks are by text position of the INITIAL occupant ...
            else:
                assert seeded[r] == lcp[r], (s, r)
                assert not need[SA[r]]                           # ... and a slot outside tie groups is final: same suffix
        # every marked position ends in a pending slot and vice versa (tie groups are closed under the doubling rounds)
        assert sorted(SA[r] for r in range(len(SA)) if seeded[r] is None) == [i for i in range(len(SA)) if need[i]]
        assert list(lcp) == gm.lcp_array(s, SA, RANK, Q=rnd.choice([1, 4, 32]), seed=so["seed"])


def test_rolling_window_keys_equal_the_definition():
    """kb_tile_keys (rolling windows, sentinel flags in a bit mask) against kb_key (symbol by symbol), incl. sentinels inside
    the window, windows running past the end of the text, and 32-bit keys."""
    rnd = random.Random(2)
    for it in range(300):
        sig = rnd.randint(1, 6)
        s = bytes(rnd.choice(b"ACGTXY"[:sig]) for _ in range(rnd.randint(1, 200)))
        if it % 2:
            s = s + b"\x01" + s[::-1] + b"\x02"
        if it % 5 == 0:
            s = s[: len(s) // 3] + b"N" + s[len(s) // 3:]
        cls, lay = gm.choose_layout(s, len(s) + 1, rnd.choice([None, 32, 64]), rnd.choice([None, 1, 2, 5]))
        assert gm.build_keys_rolling(s, cls, lay, rnd.choice([1, 3, 8])) == gm.build_keys(s, cls, lay)


def test_piecewise_line_probes():
    """lpnf.cuh line_prev_less / line_next_less (16-byte pieces with masks for the first and last piece) against a scan."""
    rnd = random.Random(8)
    for _ in range(3000):
        n = rnd.randint(1, 32)
        line = [rnd.randint(0, 6) for _ in range(n)]
        d = rnd.randint(0, 7)
        hi = rnd.randint(0, n - 1)
        exp = max([j for j in range(hi + 1) if line[j] < d], default=-1)
        assert gm.Trees.line_prev_less(line, hi, d) == exp
        lo = rnd.randint(0, n - 1)
        last = rnd.randint(lo, n - 1)
        exp = min([j for j in range(lo, last + 1) if line[j] < d], default=-1)
        assert gm.Trees.line_next_less(line, lo, last, d) == exp


def test_layout_symbols_fills_whole_radix_passes():
    # 250 Mbp RC text (n' = 5 * 10^8 + 3, sigma 4) -> 21 symbols + 5 offset bits = 47 bits, 6 passes
    assert gm.layout_symbols(4, 2, 0, 5, 500_000_003) == 21
    assert gm.layout_symbols(4, 2, 0, 5, 6_200_000_003) == 21                 # configs[4]: still 6 passes
    for sigma, b in [(2, 1), (4, 2), (5, 3), (20, 5), (200, 8)]:
        for n1 in [10, 10_000, 10**7, 10**10]:
            W = gm.layout_symbols(sigma, b, 0, 5, n1)
            assert 1 <= W <= min(29, 59 // b)
            assert W == min(29, 59 // b) or max(sigma, 2) ** W >= 64 * n1     # key space covers the text
            assert W == min(29, 59 // b) or (W + 1) * b + 5 > (W * b + 5 + 7) // 8 * 8   # one more symbol = one more pass


def test_tabulated_climb_settles_both_candidates():
    """Round-2 stage 3 (walk_tables: node tables with F-min AND R-max, one climb, RC depth of budget-exhausted positions found
    by the hard kernel) against the search-based model of round 1 (walk) at EVERY position, not only along the chain."""
    rnd = random.Random(33)
    for it in range(250):
        sig = rnd.choice([1, 2, 2, 3, 4])
        n = rnd.randint(1, 90)
        s = bytes(rnd.choice(b"ACGT"[:sig]) for _ in range(n))
        if it % 4 == 0:                                            # tandem array inside random text: long forward-only climbs
            u = bytes(rnd.choice(b"ACGT") for _ in range(rnd.randint(1, 4)))
            s = s[: n // 2] + u * rnd.randint(5, 30) + s[n // 2:]
        for mode in ("general", "rc"):
            if mode == "rc":
                S = wl.prepare_w_rc_single(s)
                N = len(S) // 2 - 1
                rc, nfac = True, N
            else:
                S, N, rc, nfac = s, 0, False, len(s)
            n1 = len(S) + 1
            SA, RANK, _ = gm.suffix_array(S)
            LCP = gm.lcp_array(S, SA, RANK)
            T = gm.Trees(LCP, SA, rc, N)
            ref = gm.walk(T, n1, nfac, RANK, 4, 16)
            for budget in [1, 2, 4, 64, rnd.randint(1, 8)]:
                assert gm.walk_tables(T, n1, nfac, RANK, budget, rnd.choice([1, 4, 16])) == ref, (s, mode, budget)


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
