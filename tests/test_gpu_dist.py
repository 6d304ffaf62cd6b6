"""One text across several ranks (csrc/dist2.cuh, csrc/dist2_host.cuh).  The ranks are dealt round-robin to the visible GPUs: on a
multi-GPU box every exchange step (Phi / LCP delivery, rank requests, edge staircases, virtual ranks, barriers) runs
over real peer memory; on a one-GPU box the ranks share cuda:0 and only the peer pointers are local."""
import random

import numpy as np
import pytest

import oracle_py as orc
from nolzss_b200 import _lib as L
from nolzss_b200 import dist as nd
from nolzss_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def _devs(world):
    import torch

    n = max(torch.cuda.device_count(), 1)
    return [g % n for g in range(world)]


def _expected(mode, s):
    if mode == L.MODE_GENERAL:
        return orc.factorize(s)
    if mode == L.MODE_DNA_RC:
        return orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(s))
    return orc.factorize_multiple_dna_w_rc(s)


def _cases():
    rng = random.Random(5)
    cases = [b"ACGT" * 8, b"A" * 300, b"AC" * 200, b"GATTACA" * 30, b"ACGTTGCAACGTTGCA", b"T" + b"A" * 99 + b"C" + b"A" * 99]
    for _ in range(40):
        n = rng.choice([2, 3, 7, 33, 100, 500, 2000])
        sigma = rng.choice([1, 2, 4, 4, 4])
        cases.append(bytes(rng.choice(b"ACGT"[:sigma]) for _ in range(n)))
    cases.append(wl.planted_dna(60_000, 7, scale=0.1).tobytes())
    cases.append(wl.uniform_dna(50_000, 8).tobytes())
    return cases


@pytest.mark.parametrize("world", [1, 2, 3, 4])
def test_dist_small_cases_vs_oracle(world):
    grp = nd.LocalGroup(_devs(world), 200_000, L.MODE_DNA_RC)
    try:
        for s in _cases():
            for mode in (L.MODE_DNA_RC, L.MODE_GENERAL):
                got, _ = grp.factorize(mode, s)
                assert np.array_equal(got, _expected(mode, s)), (world, mode, s[:60], len(s))
    finally:
        grp.close()


def test_dist_general_bytes_and_prepared():
    rng = np.random.default_rng(3)
    text = bytes(rng.integers(97, 123, 30_000, dtype=np.uint8))          # 26-letter alphabet: 8-bit symbols
    words = [b"lorem", b"ipsum", b"dolor", b"sit", b"amet", b"consectetur"]
    prose = b" ".join(words[i] for i in rng.integers(0, len(words), 8000))
    from treewalk_model import prepare_multiple_dna_sequences_w_rc
    S, _, _ = prepare_multiple_dna_sequences_w_rc([wl.uniform_dna(3000, 1).tobytes(), wl.planted_dna(5000, 2, scale=0.02).tobytes(), b"ACGTACGTAC"])
    grp = nd.LocalGroup(_devs(3), 100_000, L.MODE_GENERAL)
    try:
        for s in (text, prose):
            got, _ = grp.factorize(L.MODE_GENERAL, s)
            assert np.array_equal(got, orc.factorize(s))
        got, _ = grp.factorize(L.MODE_RC_PREPARED, S)
        assert np.array_equal(got, orc.factorize_multiple_dna_w_rc(S))
    finally:
        grp.close()


def test_dist_matches_single_gpu_5mbp_rc():
    """configs[1] text on 4 ranks: identical triples to the single-GPU pipeline (itself oracle-checked)."""
    s = wl.c2_text()
    single = L.factorize_array(L.MODE_DNA_RC, s)
    grp = nd.LocalGroup(_devs(4), len(s), L.MODE_DNA_RC)
    try:
        got, stats = grp.factorize(L.MODE_DNA_RC, s)
    finally:
        grp.close()
    assert np.array_equal(got, single)
    assert sum(st["active_sum"] for st in stats) > 0


def test_dist_hybrid_rounds_big_tie_groups():
    """The hybrid doubling rounds (big_groups.cuh) inside the distributed loop: tile capacity lowered to 64 members on
    every rank's context, with and without the forced redo of a round."""
    from test_gpu_parity import _hybrid_cases
    cases = _hybrid_cases()
    for world in (2, 3):
        grp = nd.LocalGroup(_devs(world), 400_000, L.MODE_DNA_RC)
        try:
            for flags in ((64 << 8), (64 << 8) | 8):
                for c in grp.ctxs:
                    L.check(L.load().nlz_set_debug_flags(c, flags))
                for s in cases:
                    modes = (L.MODE_GENERAL, L.MODE_DNA_RC) if set(s) <= set(b"ACGT") else (L.MODE_GENERAL,)
                    for mode in modes:
                        got, _ = grp.factorize(mode, s)
                        assert np.array_equal(got, _expected(mode, s)), (world, hex(flags), mode, len(s))
        finally:
            grp.close()


def test_dist_invalid_nucleotide_is_reported_by_every_rank():
    grp = nd.LocalGroup(_devs(3), 10_000, L.MODE_DNA_RC)
    try:
        s = wl.uniform_dna(5000, 3).tobytes()
        bad = s[:4100] + b"N" + s[4101:]
        with pytest.raises(RuntimeError, match="Invalid nucleotide 'N' found in sequence 0"):
            grp.factorize(L.MODE_DNA_RC, bad)
        got, _ = grp.factorize(L.MODE_DNA_RC, s)          # the group stays usable
        assert np.array_equal(got, _expected(L.MODE_DNA_RC, s))
    finally:
        grp.close()


def test_dist_33_bit_ranks_on_small_texts():
    """Debug flag 0x1000000: every global rank carries an offset of 2^32 + 12345, so the doubling keys (group << 33 |
    rank), the tile sort / group stream / radix rounds and every exchanged record run with ranks beyond 32 bits -- the
    widths of configs[4] (6.2 * 10^9 suffixes) -- on texts the oracle can check."""
    from test_gpu_parity import _hybrid_cases
    cases = _cases()[:30] + _hybrid_cases()[:4] + [wl.planted_dna(300_000, 21, scale=0.3).tobytes()]
    for world in (1, 3):
        grp = nd.LocalGroup(_devs(world), 400_000, L.MODE_DNA_RC)
        try:
            for flags in (0x1000000, 0x1000000 | (64 << 8), 0x1000000 | (64 << 8) | 8, 0x1000000 | 1):
                for c in grp.ctxs:
                    L.check(L.load().nlz_set_debug_flags(c, flags))
                for s in cases:
                    modes = (L.MODE_GENERAL, L.MODE_DNA_RC) if set(s) <= set(b"ACGT") else (L.MODE_GENERAL,)
                    for mode in modes:
                        got, _ = grp.factorize(mode, s)
                        assert np.array_equal(got, _expected(mode, s)), (world, hex(flags), mode, len(s))
        finally:
            grp.close()


def test_dist_device_resident_text_and_caller_buffer():
    """nlz_dist_factorize_into: the text is a device pointer on every rank, rank 0 hands in the output buffer
    (the entry point bench.py times)."""
    for world in (1, 3):
        devs = _devs(world)
        grp = nd.LocalGroup(devs, 400_000, L.MODE_DNA_RC)
        try:
            for s in (wl.planted_dna(200_000, 31, scale=0.2).tobytes(), b"ACGT" * 500, b"A" * 1000):
                for mode in (L.MODE_DNA_RC, L.MODE_GENERAL):
                    got = grp.factorize_device_text(mode, s, devs)
                    assert np.array_equal(got, _expected(mode, s)), (world, mode, len(s))
            # a caller buffer that is too small is reported by rank 0 AFTER the closing barrier: no rank is left waiting,
            # and the group stays usable
            with pytest.raises(RuntimeError, match="output capacity 3 factors is too small"):
                grp.factorize_device_text(L.MODE_DNA_RC, b"ACGT" * 500, devs, capacity=3)
            got = grp.factorize_device_text(L.MODE_DNA_RC, b"ACGT" * 500, devs)
            assert np.array_equal(got, _expected(L.MODE_DNA_RC, b"ACGT" * 500))
        finally:
            grp.close()


def test_dist_partition_balance_on_low_complexity_text():
    """A text whose suffixes crowd into two 12-symbol buckets: 9 Mbp of (AT)^n (its own reverse complement) inside 10 Mbp
    -- 90 % of all suffixes start with ATATATATATAT or TATATATATATA.  A prefix partition cannot balance such a text (a tie
    group is one unit, and deeper splitters do not split (AT)^n either); the run must stay correct, and the imbalance
    is reported (max / mean of the rank-range sizes)."""
    x = wl.uniform_dna(10_000_000, 13).copy()
    x[500_000:9_500_000] = np.frombuffer(b"AT" * 4_500_000, dtype=np.uint8)
    s = x.tobytes()
    single = L.factorize_array(L.MODE_DNA_RC, s)
    grp = nd.LocalGroup(_devs(4), len(s), L.MODE_DNA_RC)
    try:
        got, stats = grp.factorize(L.MODE_DNA_RC, s)
    finally:
        grp.close()
    assert np.array_equal(got, single)
    sizes = [st["n_local_suffixes"] for st in stats]
    assert sum(sizes) == 2 * len(s) + 3
    print(f"rank-range sizes {sizes}: max/mean = {max(sizes) / (sum(sizes) / len(sizes)):.2f}")
    assert max(sizes) / (sum(sizes) / len(sizes)) > 1.5          # the case is unbalanced by construction
