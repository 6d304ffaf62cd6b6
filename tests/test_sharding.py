"""CPU: the N>1 path (record sharding + gather) with a world-size-2 gloo group."""
import os
import sys

import torch.multiprocessing as mp

from nolzss_b200.sharding import assign_records

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_assign_records_is_balanced_and_complete():
    lengths = [100, 5, 90, 7, 50, 50, 3, 1]
    shares = assign_records(lengths, 3)
    assert sorted(i for s in shares for i in s) == list(range(len(lengths)))
    loads = [sum(lengths[i] for i in s) for s in shares]
    assert max(loads) - min(loads) <= max(lengths)
    assert assign_records(lengths, 1) == [list(range(len(lengths)))]
    assert assign_records([], 2) == [[], []]


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    import torch.distributed as dist

    import oracle_py as orc
    from nolzss_b200 import workloads as wl
    from nolzss_b200.sharding import factorize_records_distributed

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    recs = wl.c3_records(7, 400, seed=5)
    # stand-in for the GPU call (no GPU in this test): the CPU oracle, so the gathered result is checkable
    one = lambda s: [tuple(int(x) for x in r) for r in orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(s))]
    res = factorize_records_distributed(recs, one)
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_matches_serial():
    sys.path[:0] = [os.path.join(ROOT, "oracle")]
    import oracle_py as orc
    from nolzss_b200 import workloads as wl

    orc.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    recs = wl.c3_records(7, 400, seed=5)
    exp = [(rid, [tuple(int(x) for x in r) for r in orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(s))])
           for rid, s in recs]
    assert got[0] == exp and got[1] == exp


def _oracle_batch(recs, with_rc, want_factors):
    """Stand-in for nlz_factorize_batch on the CPU: the oracle per record (test infrastructure)."""
    import numpy as np

    import oracle_py as orc
    from nolzss_b200 import workloads as wl

    parts = [(orc.factorize_multiple_dna_w_rc(wl.prepare_w_rc_single(s)) if with_rc else orc.factorize(s)) if s
             else np.zeros((0, 3), dtype=np.uint64) for s in recs]
    counts = np.array([len(p) for p in parts], dtype=np.uint64)
    trip = np.concatenate(parts) if parts else np.zeros((0, 3), dtype=np.uint64)
    return (trip if want_factors else None), counts


def _batch_worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    import torch.distributed as dist

    from nolzss_b200 import workloads as wl
    from nolzss_b200.sharding import factorize_batch_distributed
    from test_sharding import _oracle_batch

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    recs = [s for _, s in wl.c3_records(9, 300, seed=6)] + [b"", b"ACGT"]
    counts, trip = factorize_batch_distributed(recs, True, True, batch_fn=_oracle_batch)
    counts2, none = factorize_batch_distributed(recs, True, False, batch_fn=_oracle_batch)
    q.put((rank, counts.tolist(), trip.tolist(), counts2.tolist(), none is None))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_batch_arrays_match_serial():
    """factorize_batch_distributed (configs[2] path): one batch call per rank, counts and triples gathered as arrays."""
    sys.path[:0] = [os.path.join(ROOT, "oracle")]
    import oracle_py as orc
    from nolzss_b200 import workloads as wl

    orc.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 500)
    procs = [ctx.Process(target=_batch_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    recs = [s for _, s in wl.c3_records(9, 300, seed=6)] + [b"", b"ACGT"]
    trip, counts = _oracle_batch(recs, True, True)
    for rank, c, t, c2, none_ok in got:
        assert c == counts.tolist() and c2 == counts.tolist() and none_ok
        assert t == trip.tolist()
