"""One text across the GPUs of a box (configs[3]/[4]; C side: csrc/dist2.cuh + csrc/dist2_host.cuh, nlz_dist_* in the C ABI).

Two ways to form a group of ranks, one rank per GPU:

* ``LocalGroup``   all ranks in this process, one host thread per rank (also used by the tests with several
                   ranks sharing ONE GPU, which exercises every exchange step without a multi-GPU box);
* ``ProcessGroup`` one process per GPU (torchrun): the CUDA IPC handles of the shared segments are
                   all-gathered with ``torch.distributed`` -- that is all torch does here; the data path is
                   peer-memory loads/stores issued by the kernels themselves.
"""
from __future__ import annotations

import ctypes
import threading

import numpy as np

from . import _lib as L

_vp = ctypes.c_void_p


def _triples(out, z):
    if z == 0:
        return np.zeros((0, 3), dtype=np.uint64)
    try:
        return np.ctypeslib.as_array(out, shape=(z * 3,)).copy().reshape(z, 3)
    finally:
        L.load().nlz_free(out)


def _ctx_stats(ctx) -> dict:
    s = L.Stats()
    L.check(L.load().nlz_get_stats(ctx, ctypes.byref(s)))
    return s.as_dict()


class LocalGroup:
    """`world` ranks inside this process; devices[g] is the CUDA device of rank g (repeats allowed)."""

    def __init__(self, devices, max_text_bytes: int, mode: int = L.MODE_DNA_RC):
        lib = L.load()
        self.world = len(devices)
        self.ctxs, self.dists = [], []
        try:
            for g, dev in enumerate(devices):
                c = _vp()
                L.check(lib.nlz_ctx_create(int(dev), ctypes.byref(c)))
                self.ctxs.append(c)
                d = _vp()
                L.check(lib.nlz_dist_create(c, g, self.world, max_text_bytes, mode, ctypes.byref(d)))
                self.dists.append(d)
            arr = (_vp * self.world)(*[d.value for d in self.dists])
            L.check(lib.nlz_dist_attach_local(arr, self.world))
        except Exception:
            self.close()          # a later rank failed (e.g. out of memory): release what the earlier ranks hold
            raise

    def factorize(self, mode: int, data):
        """Runs the collective call on one thread per rank; returns (triples of rank 0, [stats per rank])."""
        lib = L.load()
        addr, n, keep = L._as_buffer(data)
        results = [None] * self.world
        errors = [None] * self.world

        def work(g):
            out = L._u64p()
            cnt = L._u64(0)
            rc = lib.nlz_dist_factorize(self.dists[g], mode, addr, n, ctypes.byref(out) if g == 0 else None, ctypes.byref(cnt))
            if rc != L.NLZ_OK:
                errors[g] = (rc, lib.nlz_last_error().decode("utf-8", "replace"))
                return
            results[g] = (_triples(out, cnt.value) if g == 0 else None, cnt.value)

        threads = [threading.Thread(target=work, args=(g,)) for g in range(self.world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        failed = [(g, e) for g, e in enumerate(errors) if e is not None]
        if failed:
            # a rank that fails leaves the others waiting in a barrier until it times out: report the root cause first
            failed.sort(key=lambda ge: "timed out" in ge[1][1])
            g, e = failed[0]
            msg = "; ".join(f"rank {gg}: {ee[1]}" for gg, ee in failed)
            raise (ValueError if e[0] == L.NLZ_ERR_INVALID else RuntimeError)(msg)
        counts = {r[1] for r in results}
        assert len(counts) == 1, f"ranks disagree on the factor count: {counts}"
        return results[0][0], [_ctx_stats(c) for c in self.ctxs]

    def factorize_device_text(self, mode: int, data, devices, capacity=None):
        """As `factorize`, through nlz_dist_factorize_into: every rank passes a DEVICE pointer to its own copy of the text
        (already resident in its HBM) and rank 0 a caller-provided output buffer.  Returns the triples of rank 0."""
        import torch

        lib = L.load()
        addr, n, keep = L._as_buffer(data)
        host = np.frombuffer(keep, dtype=np.uint8) if n else np.zeros(0, dtype=np.uint8)
        texts = [torch.from_numpy(host.copy()).to(f"cuda:{dev}") for dev in devices]
        cap = n + 16 if capacity is None else int(capacity)
        out = np.zeros((max(cap, 1), 3), dtype=np.uint64)
        counts = [None] * self.world
        errors = [None] * self.world

        def work(g):
            cnt = L._u64(0)
            rc = lib.nlz_dist_factorize_into(self.dists[g], mode, texts[g].data_ptr() if n else None, n,
                                             out.ctypes.data if g == 0 else None, cap if g == 0 else 0, ctypes.byref(cnt))
            if rc != L.NLZ_OK:
                errors[g] = lib.nlz_last_error().decode("utf-8", "replace")
            counts[g] = cnt.value

        threads = [threading.Thread(target=work, args=(g,)) for g in range(self.world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if any(errors):
            raise RuntimeError("; ".join(f"rank {g}: {e}" for g, e in enumerate(errors) if e))
        assert len(set(counts)) == 1
        return out[:counts[0]].copy()

    def close(self):
        lib = L.load()
        for d in self.dists:
            lib.nlz_dist_destroy(d)
        for c in self.ctxs:
            lib.nlz_ctx_destroy(c)
        self.dists, self.ctxs = [], []


class ProcessGroup:
    """This process is one rank of an initialised torch.distributed group (one process per GPU)."""

    def __init__(self, max_text_bytes: int, mode: int = L.MODE_DNA_RC, device: int | None = None):
        import torch
        import torch.distributed as dist

        lib = L.load()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.ctx = L.context(device)
        self.dist = _vp()
        L.check(lib.nlz_dist_create(self.ctx, self.rank, self.world, max_text_bytes, mode, ctypes.byref(self.dist)))
        hb = lib.nlz_dist_ipc_handle_bytes()
        mine = np.zeros(hb, dtype=np.uint8)
        L.check(lib.nlz_dist_export(self.dist, mine.ctypes.data))
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.from_numpy(mine).to(dev)
        gathered = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(gathered, t)
        handles = np.concatenate([g.cpu().numpy() for g in gathered]).astype(np.uint8)
        L.check(lib.nlz_dist_attach(self.dist, handles.ctypes.data))
        dist.barrier()

    def factorize(self, mode: int, data):
        """Collective: every rank passes the same text.  Rank 0 gets the (z, 3) triples, the others None."""
        lib = L.load()
        addr, n, keep = L._as_buffer(data)
        out = L._u64p()
        cnt = L._u64(0)
        L.check(lib.nlz_dist_factorize(self.dist, mode, addr, n, ctypes.byref(out) if self.rank == 0 else None, ctypes.byref(cnt)))
        return (_triples(out, cnt.value) if self.rank == 0 else None), cnt.value

    def stats(self) -> dict:
        return _ctx_stats(self.ctx)

    def close(self):
        if self.dist:
            L.load().nlz_dist_destroy(self.dist)
            self.dist = None
