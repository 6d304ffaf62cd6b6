"""One large text across the GPUs of a box (torchrun, one process per GPU): time + check against one GPU.

torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_run.py [n_bases] [mode]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from nolzss_b200 import _lib as L, dist as nd, workloads as wl

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
mode = {"rc": L.MODE_DNA_RC, "general": L.MODE_GENERAL}[sys.argv[2] if len(sys.argv) > 2 else "rc"]
check = (sys.argv[3] if len(sys.argv) > 3 else "check") == "check"
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t0 = time.perf_counter()
scale = max(1.0, n / 5_000_000)
text = wl.planted_dna(n, 4, scale=scale).tobytes()
if rank == 0: print(f"[dist_run] text of {n} bases generated in {time.perf_counter()-t0:.1f} s; world={world}", flush=True)
grp = nd.ProcessGroup(n, mode, device=local)
for it in range(3):
    dist.barrier()
    t0 = time.perf_counter()
    got, z = grp.factorize(mode, text)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = grp.stats()
    ms = torch.tensor([st["ms_total"]], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    line = (f"rank {rank} it {it}: z={z} wall={dt*1e3:.1f} ms dev={st['ms_total']:.1f} (prep {st['ms_prepare']:.1f} keys {st['ms_keys']:.1f} sort0 {st['ms_sort0']:.1f} "
            f"doubling {st['ms_doubling']:.1f} lcp {st['ms_lcp']:.1f} lpnf {st['ms_lpnf']:.1f} chain {st['ms_chain']:.1f}) rounds={st['doubling_rounds']} "
            f"active_sum={st['active_sum']} records_applied={st['rank_records_applied']} suffixes={st['n_suffixes']} ws={st['workspace_bytes']/2**30:.1f} GiB")
    print(line, flush=True)
    if rank == 0: print(f"[dist_run] it {it}: max-over-ranks device time {ms.item():.1f} ms -> {n/ms.item()/1e3:.1f} Mbases/s on {world} GPUs", flush=True)
L.check(L.load().nlz_set_profiling(grp.ctx, 1))
dist.barrier()
got, z = grp.factorize(mode, text)
Lb = L.load()
import ctypes
ks = {}
for cls in range(Lb.nlz_kernel_class_count()):
    name = ctypes.c_char_p(); ms = ctypes.c_double(0); by = L._u64(0); ln = ctypes.c_uint32(0)
    L.check(Lb.nlz_get_kernel_stats(grp.ctx, cls, ctypes.byref(name), ctypes.byref(ms), ctypes.byref(by), ctypes.byref(ln)))
    if ln.value: ks[name.value.decode()] = (round(ms.value, 1), ln.value)
for r in range(world):
    dist.barrier()
    if r == rank: print(f"rank {rank} kernel classes (ms, launches): {ks}", flush=True)
L.check(L.load().nlz_set_profiling(grp.ctx, 0))
gold_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c4_250mbp_rc.json")
if rank == 0 and n == 250_000_000 and mode == L.MODE_DNA_RC and os.path.exists(gold_path):
    import hashlib, json
    gold = json.load(open(gold_path))
    hsh = hashlib.sha256(got.astype("<u8").tobytes()).hexdigest()
    print(f"[dist_run] oracle sha256 of the 250 Mbp RC text: {'MATCH' if hsh == gold['sha256_triples_le_u64'] and len(got) == gold['factors'] else 'MISMATCH'} ({len(got)} factors)", flush=True)
    assert hsh == gold["sha256_triples_le_u64"]
if check and rank == 0:
    t0 = time.perf_counter()
    single = L.factorize_array(mode, text, device=local)
    st = L.stats(local)
    print(f"[dist_run] single GPU: dev={st['ms_total']:.1f} ms -> {n/st['ms_total']/1e3:.1f} Mbases/s (prep {st['ms_prepare']:.1f} keys {st['ms_keys']:.1f} sort0 {st['ms_sort0']:.1f} "
          f"doubling {st['ms_doubling']:.1f} lcp {st['ms_lcp']:.1f} lpnf {st['ms_lpnf']:.1f} chain {st['ms_chain']:.1f}); identical triples: {np.array_equal(single, got)}", flush=True)
    assert np.array_equal(single, got)
dist.barrier()
grp.close()
dist.destroy_process_group()
