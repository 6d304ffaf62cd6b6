"""Randomised soak: single-GPU, batch and 3-rank distributed paths against the CPU oracle on many generated texts."""
import os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import oracle_py as orc
from nolzss_b200 import _lib as L, dist as nd, workloads as wl

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
import torch
_nd = max(torch.cuda.device_count(), 1)
grp = nd.LocalGroup([g % _nd for g in range(3)], 2_000_000, L.MODE_DNA_RC)
t_end = time.time() + budget
n_cases = 0
def gen():
    kind = rng.choice(["planted", "planted", "uniform", "lowsigma", "tandem", "runs"])
    n = rng.choice([1000, 20_000, 150_000, 600_000])
    seed = rng.randrange(1 << 30)
    if kind == "planted":
        return wl.planted_dna(n, seed, scale=rng.choice([0.02, 0.1, 0.5]), families=rng.randint(1, 20), tandems=rng.randint(0, 40)).tobytes()
    if kind == "uniform":
        return wl.uniform_dna(n, seed).tobytes()
    r = np.random.default_rng(seed)
    if kind == "lowsigma":
        return np.frombuffer(b"ACGT", dtype=np.uint8)[r.integers(0, rng.choice([1, 2, 3]), n)].tobytes()
    if kind == "tandem":
        unit = np.frombuffer(b"ACGT", dtype=np.uint8)[r.integers(0, 4, rng.randint(1, 300))]
        x = np.tile(unit, n // len(unit) + 1)[:n].copy()
        pos = r.integers(0, n, rng.randint(0, 6)); x[pos] = np.frombuffer(b"ACGT", dtype=np.uint8)[r.integers(0, 4, len(pos))]
        return x.tobytes()
    x = np.repeat(np.frombuffer(b"ACGT", dtype=np.uint8)[r.integers(0, 4, n // 50 + 1)], r.integers(1, 100, n // 50 + 1))[:n]
    return x.tobytes()
FLAG_CHOICES = [0, 0, 0, 64 << 8, (64 << 8) | 8, 128 << 8, (256 << 8) | 8, 1024 << 8, 1, 2, 4, (64 << 8) | 1, (512 << 8) | 2]
while time.time() < t_end:
    s = gen()
    # test hooks (nlz_set_debug_flags): tile capacity lowered so that small texts run the hybrid rounds of
    # big_groups.cuh, forced redo of a hybrid round (8), forced tile-sort paths (1, 2, 4)
    flags = rng.choice(FLAG_CHOICES)
    for c in [L.context()] + list(grp.ctxs):
        L.check(L.load().nlz_set_debug_flags(c, flags))
    print(f"case {n_cases}: n={len(s)} flags={flags:#x} head={s[:24]!r}", flush=True)
    exp_rc = orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(s))
    exp_g = orc.factorize(s)
    assert np.array_equal(L.factorize_array(L.MODE_DNA_RC, s), exp_rc), ("single rc", len(s))
    assert np.array_equal(L.factorize_array(L.MODE_GENERAL, s), exp_g), ("single general", len(s))
    got, _ = grp.factorize(L.MODE_DNA_RC, s); assert np.array_equal(got, exp_rc), ("dist rc", len(s))
    got, _ = grp.factorize(L.MODE_GENERAL, s); assert np.array_equal(got, exp_g), ("dist general", len(s))
    # batch: the text cut into random records
    cuts = sorted(rng.sample(range(1, len(s)), min(len(s) - 1, rng.randint(1, 40))))
    recs = [s[a:b] for a, b in zip([0] + cuts, cuts + [len(s)])]
    for with_rc in (True, False):
        got, counts = L.factorize_batch(recs, with_rc)
        at = 0
        for j, rec in enumerate(recs):
            e = orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(rec)) if with_rc else orc.factorize(rec)
            assert counts[j] == len(e) and np.array_equal(got[at:at + len(e)], e), ("batch", with_rc, j, len(rec))
            at += len(e)
    n_cases += 1
grp.close()
print(f"soak ok: {n_cases} texts x (single, distributed x3 ranks, batch) x (rc, general) identical to the oracle")
