"""Multi-GPU use of the path: independent texts / FASTA records are dealt to ranks, every rank runs
the single-GPU pipeline on its share, results are gathered per record.  No collective on the data
path (first row); torch.distributed is only the plumbing (NCCL on the GPU box,
gloo in the CPU tests)."""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple


def assign_records(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Length-balanced, deterministic deal: longest record first to the least loaded rank."""
    load = [0] * world_size
    shares: List[List[int]] = [[] for _ in range(world_size)]
    for idx in sorted(range(len(lengths)), key=lambda i: (-lengths[i], i)):
        r = min(range(world_size), key=lambda k: (load[k], k))
        shares[r].append(idx)
        load[r] += lengths[idx]
    for s in shares:
        s.sort()
    return shares


def factorize_records_distributed(records: Sequence[Tuple[str, bytes]], factorize_one: Callable[[bytes], object],
                                  group=None):
    """Every rank factorizes its share with `factorize_one`; all ranks return the full, ordered list
    [(id, result)].  `factorize_one` is e.g. nolzss_b200._noLZSS.factorize_dna_w_rc."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    shares = assign_records([len(s) for _, s in records], world)
    mine = [(i, factorize_one(records[i][1])) for i in shares[rank]]
    if world == 1:
        gathered = [mine]
    else:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine, group=group)
    out = [None] * len(records)
    for part in gathered:
        for i, res in part:
            out[i] = (records[i][0], res)
    return out


def factorize_batch_distributed(records: Sequence[bytes], with_rc: bool = True, want_factors: bool = False, group=None,
                                batch_fn=None, device=None):
    """configs[2] across ranks: the records are dealt by `assign_records`, every rank runs ONE `nlz_factorize_batch`
    call over its whole share (the segmented pipeline of csrc: record id = leading sort-key field), and the results
    are gathered as ARRAYS -- per-record factor counts always (one padded int64 all_gather), the record-local triples
    when `want_factors` (one padded all_gather of the flattened shares, stitched back into record order).  No
    collective on the data path: the gather is the only communication.  Replaces the reference's worker pool over an
    atomic record index (parallel_fasta_processor.cpp:360-385).

    `batch_fn(records, with_rc, want_factors)` -> (triples or None, counts) defaults to `_lib.factorize_batch`; the
    CPU tests inject an oracle-backed stand-in.  Returns (counts[int64, k], triples or None) on every rank."""
    import numpy as np
    import torch
    import torch.distributed as dist

    if batch_fn is None:
        from . import _lib as L

        def batch_fn(recs, rc, wf):
            return L.factorize_batch(recs, rc, want_factors=wf, device=device)

    on = dist.is_initialized()
    world = dist.get_world_size(group) if on else 1
    rank = dist.get_rank(group) if on else 0
    k = len(records)
    shares = assign_records([len(s) for s in records], world)
    mine = shares[rank]
    trip, cnts = batch_fn([records[i] for i in mine], with_rc, want_factors)
    cnts = np.asarray(cnts, dtype=np.int64)
    counts = np.zeros(k, dtype=np.int64)
    if world == 1:
        counts[mine] = cnts
        return counts, (np.asarray(trip, dtype=np.uint64) if want_factors else None)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    width = max(len(s) for s in shares)
    pad = torch.zeros(width, dtype=torch.int64)
    pad[: len(mine)] = torch.from_numpy(cnts)
    allc = torch.empty(world * width, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allc, pad.to(dev), group=group)
    allc = allc.cpu().numpy().reshape(world, width)
    for r in range(world):
        counts[shares[r]] = allc[r, : len(shares[r])]
    if not want_factors:
        return counts, None
    per_rank = [int(allc[r, : len(shares[r])].sum()) for r in range(world)]
    zmax = max(per_rank)
    flat = torch.zeros(zmax * 3, dtype=torch.int64)
    if per_rank[rank]:
        flat[: per_rank[rank] * 3] = torch.from_numpy(np.ascontiguousarray(trip, dtype=np.uint64).view(np.int64).reshape(-1))
    allt = torch.empty(world * zmax * 3, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allt, flat.to(dev), group=group)
    allt = allt.cpu().numpy().view(np.uint64).reshape(world, zmax, 3)
    starts = np.zeros(k + 1, dtype=np.int64)
    np.cumsum(counts, out=starts[1:])
    out = np.empty((int(starts[-1]), 3), dtype=np.uint64)
    for r in range(world):
        at = 0
        for i in shares[r]:
            c = int(counts[i])
            out[starts[i]: starts[i] + c] = allt[r, at: at + c]
            at += c
    return counts, out
