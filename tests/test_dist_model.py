"""CPU: the boundary exchange of the distributed path (edge staircases -> virtual ranks, tests/dist_model.py,
mirroring run_dist2 in csrc/dist2_host.cuh) against the oracle, for every cut of the suffix array into rank ranges."""
import random

import dist_model as dm
import oracle_py as orc
import treewalk_model as tm
from kats import plain_tuples


def _check(s, mode, G, K, **kw):
    exp = orc.factorize(s) if mode == "general" else orc.factorize_multiple_dna_w_rc(s)
    assert dm.factorize_dist_model(s, mode, G, K, **kw) == plain_tuples(exp), (s, mode, G, K)


def test_dist_model_random_small():
    rnd = random.Random(9)
    for it in range(250):
        sig = rnd.choice([1, 2, 2, 3, 4])
        s = bytes(rnd.choice(b"ACGT"[:sig]) for _ in range(rnd.randint(1, 70)))
        G, K = rnd.randint(2, 5), rnd.randint(1, 6)
        _check(s, "general", G, K)
        seqs = [bytes(rnd.choice(b"ACGT"[:sig]) for _ in range(rnd.randint(1, 30))) for _ in range(rnd.randint(1, 3))]
        S, _, _ = tm.prepare_multiple_dna_sequences_w_rc(seqs)
        _check(S, "rc_prepared", G, K)


def test_dist_model_runs_tandems_and_empty_ranges():
    for s in (b"A" * 90, b"AC" * 40, b"ACG" * 25 + b"T" + b"ACG" * 9, b"AAAAC" * 15):
        for G in (2, 3, 7):
            _check(s, "general", G, 4)
            S, _, _ = tm.prepare_multiple_dna_sequences_w_rc([s])
            _check(S, "rc_prepared", G, 4)
    # ranges that are empty, and ranges that connect entirely (chains through several neighbours)
    s = b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAC"
    S, _, _ = tm.prepare_multiple_dna_sequences_w_rc([s])
    n1 = len(S) + 1
    _check(S, "rc_prepared", 0, 64, bounds=[0, 3, 3, 9, 14, 14, 20, n1])
    _check(s, "general", 0, 64, bounds=[0, 5, 5, 11, 30, len(s) + 1])
