"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle (bit-exact)."""
import ctypes
import random

import numpy as np
import pytest

import oracle_py as orc
import treewalk_model as tm
from kats import ATAT_THIRD, GENERAL_KATS, RC_KATS, plain_tuples, rc_tuples
from nolzss_b200 import _lib as L
from nolzss_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def _sort_case(dtype, m, lo, hi, seed, dup=False):
    rng = np.random.default_rng(seed)
    bits = 32 if dtype == np.uint32 else 64
    keys = rng.integers(0, 2**bits, m, dtype=np.uint64).astype(dtype)
    if dup:
        keys = (keys % dtype(7)).astype(dtype) << dtype(lo)
    vals = np.arange(m, dtype=np.uint32)
    mask = ((1 << (hi - lo)) - 1)
    field = (keys.astype(np.uint64) >> np.uint64(lo)) & np.uint64(mask)
    order = np.argsort(field, kind="stable")
    k2, v2 = keys.copy(), vals.copy()
    fn = L.load().nlz_debug_sort_pairs_u32 if dtype == np.uint32 else L.load().nlz_debug_sort_pairs_u64
    L.check(fn(L.context(), k2.ctypes.data, v2.ctypes.data, m, lo, hi))
    assert np.array_equal(v2, vals[order]), (dtype, m, lo, hi)
    assert np.array_equal(k2, keys[order])


@pytest.mark.parametrize("m", [1, 2, 31, 257, 4096, 4097, 70_001, 1_300_003])
def test_radix_sort_pairs(m):
    _sort_case(np.uint32, m, 0, 32, m)
    _sort_case(np.uint64, m, 0, 64, m + 1)
    _sort_case(np.uint64, m, 32, 53, m + 2)
    _sort_case(np.uint32, m, 4, 13, m + 3)
    _sort_case(np.uint64, m, 3, 11, m + 4, dup=True)


def _index_case(data: bytes):
    n = len(data)
    sa = np.empty(n + 1, dtype=np.uint32)
    isa = np.empty(n + 1, dtype=np.uint32)
    lcp = np.empty(n + 2, dtype=np.uint32)
    arr = np.frombuffer(data, dtype=np.uint8)
    L.check(L.load().nlz_debug_index(L.context(), arr.ctypes.data, n, sa.ctypes.data, isa.ctypes.data, lcp.ctypes.data))
    esa, elcp = orc.gpu_order_sa_lcp(data)
    assert np.array_equal(sa.astype(np.int64), esa.astype(np.int64)), "suffix array differs"
    assert np.array_equal(isa[sa], np.arange(n + 1, dtype=np.uint32)), "ISA is not the inverse of SA"
    assert np.array_equal(lcp.astype(np.int64), elcp.astype(np.int64)), "LCP array differs"


def test_index_small_and_edge():
    rnd = random.Random(1)
    for data in [b"A", b"AC", b"AAAA", b"abracadabra", b"ACGT" * 50, b"A" * 3000, b"AC" * 2500]:
        _index_case(data)
    for it in range(60):
        sig = rnd.randint(1, 6)
        s = bytes(rnd.choice(b"ACGTXY"[:sig]) for _ in range(rnd.randint(1, 400)))
        if it % 3 == 0:
            s = s + b"\x01" + s[::-1] + b"\x02"
        _index_case(s)


def test_index_medium():
    _index_case(wl.uniform_dna(200_000, 4).tobytes())
    _index_case(wl.planted_dna(300_000, 7, scale=0.2).tobytes())
    _index_case(wl.prepare_w_rc_single(wl.planted_dna(150_000, 8, scale=0.2).tobytes()))
    rng = np.random.default_rng(3)
    _index_case(rng.integers(1, 256, 100_000, dtype=np.uint8).tobytes())          # sigma = 255, 64-bit keys
    _index_case(rng.integers(97, 123, 100_000, dtype=np.uint8).tobytes())         # sigma = 26


def test_kats_through_abi():
    for text, exp in GENERAL_KATS.items():
        assert plain_tuples(L.factorize_array(L.MODE_GENERAL, text)) == exp
        assert L.count(L.MODE_GENERAL, text) == len(exp)
    for text, exp in RC_KATS.items():
        assert rc_tuples(L.factorize_array(L.MODE_DNA_RC, text)) == exp
        assert rc_tuples(L.factorize_array(L.MODE_RC_PREPARED, wl.prepare_w_rc_single(text))) == exp
        assert L.count(L.MODE_DNA_RC, text) == len(exp)
    assert rc_tuples(L.factorize_array(L.MODE_DNA_RC, b"ATAT"))[2] == ATAT_THIRD
    assert rc_tuples(L.factorize_array(L.MODE_DNA_RC, b"ATCGATCG"))[3:] == [(3, 3, 0, True), (6, 2, 2, False)]
    assert (7, 2, 0, False) in rc_tuples(L.factorize_array(L.MODE_DNA_RC, b"CAAGCACCACCGCGGCGACCGAGGCA"))


def test_random_small_vs_oracle():
    rnd = random.Random(2)
    for it in range(300):
        sig = rnd.randint(1, 4)
        s = bytes(rnd.choice(b"ACGT"[:sig]) for _ in range(rnd.randint(1, 120)))
        sp = rnd.randint(0, len(s) - 1) if it % 2 else 0
        assert np.array_equal(L.factorize_array(L.MODE_GENERAL, s, sp), orc.factorize(s, sp)), (s, sp)
        seqs = [bytes(rnd.choice(b"ACGT"[:sig]) for _ in range(rnd.randint(1, 60))) for _ in range(rnd.randint(1, 3))]
        S, _, _ = tm.prepare_multiple_dna_sequences_w_rc(seqs)
        N = len(S) // 2 - 1
        sp = rnd.randint(0, N - 1) if it % 2 else 0
        assert np.array_equal(L.factorize_array(L.MODE_RC_PREPARED, S, sp), orc.factorize_multiple_dna_w_rc(S, sp)), (S, sp)
        assert np.array_equal(L.factorize_array(L.MODE_DNA_RC, seqs[0]),
                              orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(seqs[0]))), seqs[0]


def test_general_bytes_vs_oracle():
    rng = np.random.default_rng(12)
    for n, lo, hi in [(5000, 1, 4), (20000, 97, 123), (50000, 1, 256), (3000, 65, 66)]:
        s = rng.integers(lo, hi, n, dtype=np.uint8).tobytes()
        assert np.array_equal(L.factorize_array(L.MODE_GENERAL, s), orc.factorize(s))
    s = (b"the quick brown fox jumps over the lazy dog " * 500) + b"!"
    assert np.array_equal(L.factorize_array(L.MODE_GENERAL, s), orc.factorize(s))


def test_degenerate_runs_and_tandems():
    for s in [b"A" * 1, b"A" * 2, b"A" * 1000, b"A" * 40_000, b"AC" * 20_000, b"ACG" * 7000 + b"T",
              b"A" * 5000 + b"C" + b"A" * 5000]:
        assert np.array_equal(L.factorize_array(L.MODE_GENERAL, s), orc.factorize(s)), s[:20]
        assert np.array_equal(L.factorize_array(L.MODE_DNA_RC, s),
                              orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(s))), s[:20]


def test_c1_config_general_1mbp():
    t = wl.c1_text()
    got = L.factorize_array(L.MODE_GENERAL, t)
    exp = orc.factorize(t)
    assert np.array_equal(got, exp)
    assert L.count(L.MODE_GENERAL, t) == len(exp)
    st = L.stats()
    assert st["n_factors"] == len(exp) and st["kernel_launches"] > 0


def test_planted_repeats_rc_mode():
    t = wl.planted_dna(600_000, 21, scale=0.3).tobytes()
    exp = orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(t))
    assert np.array_equal(L.factorize_array(L.MODE_DNA_RC, t), exp)
    assert np.array_equal(L.factorize_array(L.MODE_RC_PREPARED, wl.prepare_w_rc_single(t)), exp)
    assert np.array_equal(L.factorize_array(L.MODE_GENERAL, t), orc.factorize(t))


def test_c2_config_rc_5mbp():
    t = wl.c2_text()
    got = L.factorize_array(L.MODE_DNA_RC, t)
    exp = orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(t))
    assert np.array_equal(got, exp)


def test_multi_record_prepared_and_start_pos():
    recs = [r for _, r in wl.c3_records(6, 3000, 9)]
    S, ol, sent = tm.prepare_multiple_dna_sequences_w_rc(recs)
    for sp in (0, len(recs[0]) + 1, ol - 5):
        assert np.array_equal(L.factorize_array(L.MODE_RC_PREPARED, S, sp), orc.factorize_multiple_dna_w_rc(S, sp))


def test_errors_through_abi():
    with pytest.raises(RuntimeError, match="Invalid nucleotide"):
        L.factorize_array(L.MODE_DNA_RC, b"ACGTNACGT")
    S = wl.prepare_w_rc_single(b"ACGT")
    with pytest.raises(ValueError, match="start_pos"):
        L.factorize_array(L.MODE_RC_PREPARED, S, 4)
    assert len(L.factorize_array(L.MODE_GENERAL, b"")) == 0
    assert len(L.factorize_array(L.MODE_DNA_RC, b"")) == 0
    assert len(L.factorize_array(L.MODE_RC_PREPARED, b"A\x01")) == 0


def test_device_entry_point_matches_host_entry_point():
    import torch

    t = wl.planted_dna(400_000, 33, scale=0.2).tobytes()
    exp = orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(t))
    d_text = torch.frombuffer(bytearray(t), dtype=torch.uint8).cuda()
    d_out = torch.empty((len(t), 3), dtype=torch.int64, device="cuda")
    cnt = ctypes.c_uint64(0)
    st = torch.cuda.current_stream().cuda_stream
    L.check(L.load().nlz_factorize_device(L.context(0), L.MODE_DNA_RC, d_text.data_ptr(), len(t), 0, st,
                                          d_out.data_ptr(), d_out.shape[0], ctypes.byref(cnt)))
    got = d_out[: cnt.value].cpu().numpy().view(np.uint64)
    assert np.array_equal(got, exp)


def test_group_sort_fallback_paths():
    """The shared-memory group sort has three paths (pivot partition, counting, bitonic network);
    force each of them and check the factors."""
    cases = [(b"the quick brown fox jumps over the lazy dog " * 500) + b"!",
             wl.planted_dna(300_000, 41, scale=0.3).tobytes(), b"AC" * 1500 + b"G" + b"AC" * 700]
    exp = [orc.factorize(s) for s in cases]
    try:
        for flags in (1, 2, 4, 6):
            L.check(L.load().nlz_set_debug_flags(L.context(), flags))
            for s, e in zip(cases, exp):
                assert np.array_equal(L.factorize_array(L.MODE_GENERAL, s), e), flags
    finally:
        L.check(L.load().nlz_set_debug_flags(L.context(), 0))


def _hybrid_cases():
    rnd = random.Random(7)
    out = [b"AC" * 1500 + b"G" + b"AC" * 700, b"A" * 3000, b"ACG" * 900 + b"T" + b"ACG" * 400 + b"GATTACA" * 300,
           wl.planted_dna(200_000, 5, scale=1.0, families=3, tandems=25).tobytes(),
           (b"the quick brown fox jumps over the lazy dog " * 300) + b"!"]
    for _ in range(4):
        parts = []
        for _p in range(rnd.randint(2, 6)):
            unit = bytes(rnd.choice(b"ACGT") for _u in range(rnd.randint(1, 9)))
            parts.append(unit * rnd.randint(50, 2500))
            parts.append(bytes(rnd.choice(b"ACGT") for _u in range(rnd.randint(0, 40))))
        out.append(b"".join(parts))
    return out


def test_hybrid_rounds_big_tie_groups():
    """Doubling rounds with tie groups beyond the tile sort (big_groups.cuh): the debug flags lower the tile
    capacity to 64 / 256 members so that small texts run the split, the group-stream kernel and (flag 8) the
    redo of a round through the radix path; factors against the oracle in general and RC mode."""
    cases = _hybrid_cases()
    exp_g = [orc.factorize(s) for s in cases]
    exp_r = [orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(s)) if set(s) <= set(b"ACGT") else None for s in cases]
    try:
        for flags in ((64 << 8), (64 << 8) | 8, (256 << 8), (1024 << 8) | 8):
            L.check(L.load().nlz_set_debug_flags(L.context(), flags))
            for s, eg, er in zip(cases, exp_g, exp_r):
                assert np.array_equal(L.factorize_array(L.MODE_GENERAL, s), eg), (hex(flags), len(s))
                if er is not None:
                    assert np.array_equal(L.factorize_array(L.MODE_DNA_RC, s), er), (hex(flags), len(s))
        L.check(L.load().nlz_set_debug_flags(L.context(), 64 << 8))
        _index_case(cases[0])
        _index_case(cases[3])
    finally:
        L.check(L.load().nlz_set_debug_flags(L.context(), 0))


def _check_cover_and_valid(x, f, samples=4000, seed=0):
    n = len(x)
    st, ln, rf = f[:, 0].astype(np.int64), f[:, 1].astype(np.int64), f[:, 2]
    assert st[0] == 0 and np.all(st[1:] == st[:-1] + ln[:-1]) and st[-1] + ln[-1] == n
    rng = np.random.default_rng(seed)
    for k in rng.integers(0, len(f), samples):
        s0, l0, r0 = int(st[k]), int(ln[k]), int(rf[k])
        if r0 >> 63:
            r0 &= (1 << 63) - 1
            assert r0 + l0 <= s0 and np.array_equal(wl.revcomp(x[r0:r0 + l0]), x[s0:s0 + l0])
        elif r0 == s0:
            assert l0 == 1
        else:
            assert r0 + l0 <= s0 and np.array_equal(x[r0:r0 + l0], x[s0:s0 + l0])


def test_chromosome_scale_properties_single_gpu():
    """configs[3] at N=1 (scaled repeat recipe, 64-bit sort keys, radix fallback rounds, hard positions):
    too large for the oracle in test time, so the size-independent properties are checked: the factors
    tile the text in order and every sampled factor is a genuine non-overlapping (reverse-complement)
    match; RC and general mode on the same text; count == number of triples."""
    x = wl.planted_dna(60_000_000, 4, scale=10.0)
    t = x.tobytes()
    f = L.factorize_array(L.MODE_DNA_RC, t)
    _check_cover_and_valid(x, f)
    assert L.count(L.MODE_DNA_RC, t) == len(f)
    assert L.stats()["key_bits"] == 64
    g = L.factorize_array(L.MODE_GENERAL, t)
    _check_cover_and_valid(x, g, seed=1)
    assert not np.any(g[:, 2] >> np.uint64(63))
    # RC candidates can only lengthen factors on average: never more factors than 1.02x the forward-only count
    assert len(f) <= len(g) * 1.02


def test_mid_scale_parity_64bit_keys():
    """20 Mbp planted text: 64-bit initial keys + radix-sorted doubling rounds, full parity with the oracle."""
    t = wl.planted_dna(20_000_000, 11, scale=4.0).tobytes()
    assert np.array_equal(L.factorize_array(L.MODE_GENERAL, t), orc.factorize(t))
    assert L.stats()["key_bits"] == 64


def _batch_expected(records, with_rc):
    exp = []
    for s in records:
        if len(s) == 0:
            exp.append(np.zeros((0, 3), dtype=np.uint64))
        elif with_rc:
            exp.append(orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(s)))
        else:
            exp.append(orc.factorize(s))
    return exp


@pytest.mark.parametrize("with_rc", [True, False])
def test_batch_records_vs_oracle_per_record(with_rc):
    """Segmented pipeline (nlz_factorize_batch): every record must factorize exactly as if it were alone."""
    rng = random.Random(11 + with_rc)
    records = [b"A", b"", b"ACGT", b"AAAAAAAAAAAAAAAA", b"ACGTACGTACGTACGT", b"TTTTTTTTAAAAAAAA", b"GATTACA" * 9]
    for _ in range(300):
        n = rng.choice([1, 2, 3, 5, 8, 13, 40, 100, 333])
        sigma = rng.choice([1, 2, 4, 4])
        records.append(bytes(rng.choice(b"ACGT"[:sigma]) for _ in range(n)))
    # identical records and records that are reverse complements of each other must not see one another
    records += [records[10], records[10], wl.revcomp(np.frombuffer(records[10], dtype=np.uint8)).tobytes()]
    for seed in range(4):
        records.append(wl.planted_dna(20_000, 40 + seed, scale=0.05).tobytes())
    records.append(b"AC" * 3000)
    records.append(b"A" * 5000)
    got, counts = L.factorize_batch(records, with_rc)
    exp = _batch_expected(records, with_rc)
    assert counts.tolist() == [len(e) for e in exp]
    at = 0
    for j, e in enumerate(exp):
        assert np.array_equal(got[at:at + len(e)], e), (j, records[j][:40])
        at += len(e)
    assert at == len(got)
    _, counts2 = L.factorize_batch(records, with_rc, want_factors=False)
    assert counts2.tolist() == counts.tolist()


def test_batch_config3_sample():
    """configs[2] recipe (10 kbp records with a planted copy), 200 records, RC mode."""
    recs = [s for _, s in wl.c3_records(200, 10_000, seed=3)]
    got, counts = L.factorize_batch(recs, True)
    at = 0
    for j, s in enumerate(recs):
        e = orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(s))
        assert counts[j] == len(e)
        assert np.array_equal(got[at:at + len(e)], e), j
        at += len(e)


def test_batch_sparse_unordered_offsets():
    """nlz_factorize_batch with records picked from a much larger buffer, in arbitrary order, with gaps and one
    overlap (the staging copy used to span [0, max offset + len) and overran the workspace)."""
    import ctypes

    lib = L.load()
    rng = np.random.default_rng(77)
    big = wl.uniform_dna(3_000_000, 9)
    big[100_000:100_040] = np.frombuffer(b"N" * 40, dtype=np.uint8)     # garbage between the records is never read
    offs = np.array([2_900_000, 5, 1_500_000, 1_500_300, 700_000, 5, 2_999_900], dtype=np.uint64)
    lens = np.array([400, 300, 300, 123, 0, 50, 100], dtype=np.uint64)
    for with_rc in (1, 0):
        counts = np.zeros(len(offs), dtype=np.uint64)
        out, total = L._u64p(), L._u64(0)
        L.check(lib.nlz_factorize_batch(L.context(None), with_rc, big.ctypes.data, offs.ctypes.data, lens.ctypes.data,
                                        len(offs), ctypes.byref(out), counts.ctypes.data, ctypes.byref(total)))
        got = np.ctypeslib.as_array(out, shape=(total.value * 3,)).copy().reshape(-1, 3)
        lib.nlz_free(out)
        at = 0
        for j in range(len(offs)):
            s = big[int(offs[j]):int(offs[j] + lens[j])].tobytes()
            e = (orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(s)) if with_rc else orc.factorize(s)) if s else np.zeros((0, 3), dtype=np.uint64)
            assert counts[j] == len(e), (with_rc, j)
            assert np.array_equal(got[at:at + len(e)], e), (with_rc, j)
            at += len(e)
        assert at == total.value
    # an invalid nucleotide inside a picked record is reported with the record's index among the non-empty ones
    offs2 = np.array([2_900_000, 99_990], dtype=np.uint64)
    lens2 = np.array([100, 30], dtype=np.uint64)
    counts = np.zeros(2, dtype=np.uint64)
    out, total = L._u64p(), L._u64(0)
    rc = lib.nlz_factorize_batch(L.context(None), 1, big.ctypes.data, offs2.ctypes.data, lens2.ctypes.data, 2,
                                 ctypes.byref(out), counts.ctypes.data, ctypes.byref(total))
    assert rc == L.NLZ_ERR_RUNTIME and b"Invalid nucleotide 'N' found in sequence 1" in lib.nlz_last_error()


def test_c4_250mbp_rc_text_matches_oracle_hash():
    """configs[3] at full size on one GPU: the 250 Mbp RC text (26 M deep-nesting positions, hybrid doubling rounds)
    against the sha256 of the oracle's triples (tests/golden/c4_250mbp_rc.json, scripts/c4_oracle_hash.py)."""
    import hashlib
    import json
    import os

    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "c4_250mbp_rc.json")))
    t = wl.planted_dna(gold["n_bases"], 4, scale=50.0).tobytes()
    got = L.factorize_array(L.MODE_DNA_RC, t)
    assert len(got) == gold["factors"]
    assert int(got[:, 1].sum()) == gold["sum_lengths"]
    assert hashlib.sha256(got.astype("<u8").tobytes()).hexdigest() == gold["sha256_triples_le_u64"]
