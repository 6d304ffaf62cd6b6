"""Development check of the hybrid doubling rounds (big_groups.cuh): small texts with the tile capacity lowered
through the debug flags so that the split / group-stream / fallback paths run, against the CPU oracle."""
import os, sys, random
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")]
import numpy as np
import oracle_py as orc
from nolzss_b200 import _lib as L, dist as nd, workloads as wl


def cases():
    rnd = random.Random(7)
    out = [b"AC" * 1500 + b"G" + b"AC" * 700, b"A" * 3000, b"ACG" * 900 + b"T" + b"ACG" * 400 + b"GATTACA" * 300,
           wl.planted_dna(300_000, 41, scale=0.3).tobytes(), wl.planted_dna(200_000, 5, scale=1.0, families=3, tandems=25).tobytes(),
           (b"the quick brown fox jumps over the lazy dog " * 500) + b"!"]
    for _ in range(6):
        parts = []
        for _p in range(rnd.randint(2, 6)):
            unit = bytes(rnd.choice(b"ACGT") for _u in range(rnd.randint(1, 9)))
            parts.append(unit * rnd.randint(50, 2500))
            parts.append(bytes(rnd.choice(b"ACGT") for _u in range(rnd.randint(0, 40))))
        out.append(b"".join(parts))
    return out


def main():
    cs = cases()
    exp_g = [orc.factorize(s) for s in cs]
    exp_r = [orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(s)) if set(s) <= set(b"ACGT") else None for s in cs]
    bad = 0
    for flags in ((64 << 8), (64 << 8) | 8, (256 << 8), (1024 << 8) | 8, 0):
        L.check(L.load().nlz_set_debug_flags(L.context(), flags))
        for i, s in enumerate(cs):
            got = L.factorize_array(L.MODE_GENERAL, s)
            st = L.stats()
            ok = np.array_equal(got, exp_g[i])
            okr = True
            if exp_r[i] is not None:
                okr = np.array_equal(L.factorize_array(L.MODE_DNA_RC, s), exp_r[i])
            print(f"flags={flags:#x} case {i} n={len(s)} general={'ok' if ok else 'FAIL'} rc={'ok' if okr else 'FAIL'} rounds={st['doubling_rounds']}", flush=True)
            bad += (not ok) + (not okr)
    L.check(L.load().nlz_set_debug_flags(L.context(), 0))
    # distributed (in-process ranks sharing the device)
    for world in (2, 3):
        grp = nd.LocalGroup([0] * world, 400_000, L.MODE_DNA_RC)
        try:
            for flags in ((64 << 8), (64 << 8) | 8):
                for c in grp.ctxs:
                    L.check(L.load().nlz_set_debug_flags(c, flags))
                for i, s in enumerate(cs):
                    got, _ = grp.factorize(L.MODE_GENERAL, s)
                    ok = np.array_equal(got, exp_g[i])
                    okr = True
                    if exp_r[i] is not None:
                        got, _ = grp.factorize(L.MODE_DNA_RC, s)
                        okr = np.array_equal(got, exp_r[i])
                    print(f"world={world} flags={flags:#x} case {i} general={'ok' if ok else 'FAIL'} rc={'ok' if okr else 'FAIL'}", flush=True)
                    bad += (not ok) + (not okr)
        finally:
            grp.close()
    print("HYBRID CHECK", "PASSED" if bad == 0 else f"FAILED ({bad})")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
