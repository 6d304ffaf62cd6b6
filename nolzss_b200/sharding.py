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
