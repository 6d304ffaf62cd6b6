"""CPU: the oracle (C restatement + literal tree-walk model) against the reference's own vectors."""
import random

import numpy as np

import oracle_py as orc
import treewalk_model as tm
from kats import ATAT_THIRD, GENERAL_KATS, RC_KATS, plain_tuples, rc_tuples
from nolzss_b200 import workloads as wl


def test_general_kats_model_and_oracle():
    for text, exp in GENERAL_KATS.items():
        assert tm.factorize(text) == exp
        assert plain_tuples(orc.factorize(text)) == exp


def test_rc_kats_model_and_oracle():
    for text, exp in RC_KATS.items():
        assert tm.factorize_dna_w_rc(text) == exp
        S = wl.prepare_w_rc_single(text)
        assert rc_tuples(orc.factorize_multiple_dna_w_rc(S)) == exp
    assert tm.factorize_dna_w_rc(b"ATAT")[2] == ATAT_THIRD
    assert rc_tuples(orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(b"ATAT")))[2] == ATAT_THIRD


def test_survey_traps():
    # the code (not docs/RC_ALGORITHM.md) is the authority
    assert tm.factorize_dna_w_rc(b"ATCGATCG")[3:] == [(3, 3, 0, True), (6, 2, 2, False)]
    f = tm.factorize_dna_w_rc(b"CAAGCACCACCGCGGCGACCGAGGCA")
    assert (7, 2, 0, False) in f      # forward-candidate quirk (factorizer_core.hpp:279-287, 322-326)


def test_oracle_matches_literal_model_random():
    rnd = random.Random(11)
    for it in range(600):
        sig = rnd.randint(1, 4)
        s = bytes(rnd.choice(b"ACGT"[:sig]) for _ in range(rnd.randint(1, 60)))
        sp = rnd.randint(0, len(s) - 1)
        assert plain_tuples(orc.factorize(s)) == tm.factorize(s)
        assert plain_tuples(orc.factorize(s, sp)) == tm.factorize(s, sp)
        seqs = [bytes(rnd.choice(b"ACGT"[:sig]) for _ in range(rnd.randint(1, 30))) for _ in range(rnd.randint(1, 3))]
        S, _, _ = tm.prepare_multiple_dna_sequences_w_rc(seqs)
        N = len(S) // 2 - 1
        sp = rnd.randint(0, N - 1) if it % 2 else 0
        assert plain_tuples(orc.factorize_multiple_dna_w_rc(S, sp)) == tm.nolzss_multiple_dna_w_rc(S, sp)


def test_oracle_invariants_on_planted_dna():
    t = wl.planted_dna(60_000, 5, scale=0.05).tobytes()
    f = orc.factorize(t)
    assert int(f[0, 0]) == 0 and int((f[:, 1]).sum()) == len(t)
    assert np.all(f[1:, 0] == f[:-1, 0] + f[:-1, 1])
    x = np.frombuffer(t, dtype=np.uint8)
    for s, l, r in f[:: max(1, len(f) // 300)]:
        s, l, r = int(s), int(l), int(r)
        if r == s:
            assert l == 1
        else:
            assert r + l <= s and np.array_equal(x[r:r + l], x[s:s + l])
    S = wl.prepare_w_rc_single(t)
    g = orc.factorize_multiple_dna_w_rc(S)
    assert int(g[:, 1].sum()) == len(t)
    for s, l, r in g[:: max(1, len(g) // 300)]:
        s, l, r = int(s), int(l), int(r)
        if r & orc.RC_MASK:
            r &= ~orc.RC_MASK
            assert r + l <= s and np.array_equal(wl.revcomp(x[r:r + l]), x[s:s + l])
        elif r != s:
            assert r + l <= s and np.array_equal(x[r:r + l], x[s:s + l])


def test_prepare_functions():
    S, ol, sent = tm.prepare_multiple_dna_sequences_w_rc([b"ATCG", b"ggcc"])
    assert S == b"ATCG\x01GGCC\x02GGCC\x03CGAT\x04" and ol == 10 and sent == [4, 9, 14, 19]
    S, ol, sent = tm.prepare_multiple_dna_sequences_no_rc([b"ATCG", b"GGCC", b"TT"])
    assert S == b"ATCG\x01GGCC\x02TT" and ol == len(S) and sent == [4, 9]
    assert [tm.sentinel_byte(i) for i in (0, 63, 64, 65, 66)] == [1, 64, 66, 68, 69]


def test_parallel_mode_equals_serial():
    """The oracle's restatement of the reference's CPU parallel mode (parallel_factorizer.cpp:849-984: serial index,
    chunked chain walk, convergence merge) must give the serial triples (the reference asserts the same in
    tests/test_parallel_fasta.py:294-330), for any thread count and start position."""
    from nolzss_b200 import workloads as wl
    for seed, n, sp in ((50, 350_000, 0), (51, 420_000, 17), (52, 250_000, 100_001)):
        S = wl.prepare_w_rc_single(wl.planted_dna(n, seed, scale=0.3).tobytes())
        want = orc.factorize_multiple_dna_w_rc(S, sp)
        for threads in (1, 2, 3, 8):
            got, used = orc.parallel_factorize_multiple_dna_w_rc(S, threads, sp)
            assert 1 <= used <= threads and used <= max(1, (n - sp) // 100_000)
            assert np.array_equal(got, want), (seed, threads)
    got, used = orc.parallel_factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(b"ACGTACGTTTGCA"), 8)
    assert used == 1 and np.array_equal(got, orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(b"ACGTACGTTTGCA")))


def test_oracle_64bit_index_build_matches_32bit():
    """-DNLZO_IDX64 (libnlz_oracle64.so): every index widened to 64 bits -- the reference's own width
    (factorizer_core.hpp:211-213) -- gives the same triples as the default 32-bit build."""
    import ctypes
    import os
    import subprocess

    import numpy as np

    from nolzss_b200 import workloads as wl

    here = os.path.dirname(os.path.abspath(orc.__file__))
    subprocess.check_call(["make", "-C", here, "libnlz_oracle64.so"], stdout=subprocess.DEVNULL)
    L64 = ctypes.CDLL(os.path.join(here, "libnlz_oracle64.so"))
    pp = ctypes.POINTER(ctypes.POINTER(ctypes.c_uint64))
    for name in ("nlzo_factorize", "nlzo_factorize_multiple_dna_w_rc"):
        f = getattr(L64, name)
        f.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint64, pp, ctypes.POINTER(ctypes.c_uint64)]
        f.restype = ctypes.c_int
    L64.nlzo_free.argtypes = [ctypes.c_void_p]

    def run64(fn, data, sp=0):
        out = ctypes.POINTER(ctypes.c_uint64)()
        cnt = ctypes.c_uint64(0)
        assert fn(data, len(data), sp, ctypes.byref(out), ctypes.byref(cnt)) == 0
        arr = np.ctypeslib.as_array(out, shape=(cnt.value * 3,)).copy().reshape(-1, 3) if cnt.value else np.zeros((0, 3), np.uint64)
        if cnt.value:
            L64.nlzo_free(out)
        return arr

    texts = [b"abracadabra", b"A" * 500, wl.planted_dna(40_000, 3, scale=0.05).tobytes(), wl.uniform_dna(20_000, 9).tobytes()]
    for t in texts:
        assert np.array_equal(run64(L64.nlzo_factorize, t), orc.factorize(t))
        if set(t) <= set(b"ACGT"):
            S = wl.prepare_w_rc_single(t)
            assert np.array_equal(run64(L64.nlzo_factorize_multiple_dna_w_rc, S), orc.factorize_multiple_dna_w_rc(S))
            assert np.array_equal(run64(L64.nlzo_factorize_multiple_dna_w_rc, S, 100), orc.factorize_multiple_dna_w_rc(S, 100))
