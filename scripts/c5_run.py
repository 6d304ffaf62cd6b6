"""configs[4]: ONE human-genome-sized synthetic text (3.1 Gbp, 6.2 * 10^9 indexed suffixes in RC mode: 33-bit ranks and
S-positions) across the GPUs of a box (torchrun, one process per GPU).

torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 scripts/c5_run.py [n_bases] [runs] [shuffled]

Rank 0 generates the text once into /dev/shm (every rank maps it; a rank only uploads its slice), the group factorizes
it, rank 0 checks that the factors tile the text and verifies 100 000 sampled factors against the text, and one JSON
record goes to gpurun_out/r2_c5.json.  With `shuffled` the per-record-permuted control (here: one record) follows."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from nolzss_b200 import _lib as L, dist as nd, workloads as wl

n = int(sys.argv[1]) if len(sys.argv) > 1 else wl.C5_BASES
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
shuffled = len(sys.argv) > 3 and sys.argv[3] == "shuffled"
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
path = f"/dev/shm/nlz_c5_{n}.bin"
t0 = time.perf_counter()
if rank == 0:
    mm = np.lib.format.open_memmap(path + ".npy", mode="w+", dtype=np.uint8, shape=(n,))
    wl.c5_text_into(mm, n)
    mm.flush()
    del mm
    print(f"[c5] text of {n} bases generated in {time.perf_counter() - t0:.1f} s", flush=True)
dist.barrier()
text = np.load(path + ".npy", mmap_mode="r+")       # writable mapping: the slice this rank uploads gets page-locked
lib = L.load()


def pin_slice(arr):
    per = ((n + world - 1) // world + 255) // 256 * 256
    lo, hi = min(rank * per, n), min((rank + 1) * per, n)
    a0 = (arr.ctypes.data + lo) // 4096 * 4096
    a1 = (arr.ctypes.data + hi + 4095) // 4096 * 4096
    return a0 if lib.nlz_host_register(a0, a1 - a0) == L.NLZ_OK else None


grp = nd.ProcessGroup(n, L.MODE_DNA_RC, device=local)
pin_real = pin_slice(text)
print(f"[c5] rank {rank}: text slice page-locked: {pin_real is not None}", flush=True)
rec = {"workload": f"configs[4]: c5_text_into(n={n}, seed=5): planted repeats (families <= 500 kbp, tandem arrays <= 5 Mbp), RC mode, "
                   f"{2 * n + 3} indexed suffixes, one text across {world} GPUs", "n_bases": n, "n_gpus": world, "runs": []}


def one(tx, label):
    got = None
    for it in range(runs):
        dist.barrier()
        t0 = time.perf_counter()
        got, z = grp.factorize(L.MODE_DNA_RC, tx)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        st = grp.stats()
        ms = torch.tensor([st["ms_total"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        stages = {k: round(st[k], 1) for k in st if k.startswith("ms_")}
        print(f"[c5] {label} rank {rank} it {it}: z={z} wall={wall:.2f} s dev={st['ms_total']:.0f} ms {stages} rounds={st['doubling_rounds']} "
              f"active_sum={st['active_sum']} ws={st['workspace_bytes'] / 2**30:.1f} GiB", flush=True)
        if rank == 0:
            rec["runs"].append({"text": label, "it": it, "factors": int(z), "device_ms_max_over_ranks": ms.item(), "wall_s_rank0": wall,
                                "Mbases_per_s": n / ms.item() / 1e3, "stages_ms_rank0": stages, "doubling_rounds": st["doubling_rounds"],
                                "suffixes": st["n_suffixes"], "workspace_GiB_rank0": st["workspace_bytes"] / 2**30})
    return got


got = one(text, "real")
if rank == 0:
    t0 = time.perf_counter()
    chk = wl.verify_factors_sample(np.asarray(text), got, 100_000)
    rec["check_real"] = chk
    rec["rc_factors_real"] = int((got[:, 2] >> np.uint64(63)).sum())
    rec["max_factor_length_real"] = int(got[:, 1].max())
    print(f"[c5] real text: factors tile the text; {chk} verified in {time.perf_counter() - t0:.1f} s", flush=True)
    # the on-disk factor file of the reference (noLZSSv2: factors, one empty name, footer; factorizer.cpp:597-635)
    real_bin = "/dev/shm/nlz_c5_real.bin"
    t0 = time.perf_counter()
    L.check(lib.nlz_write_factor_file(real_bin.encode(), got.ctypes.data, len(got), b"\0", 1, 1, 0, n))
    rec["real_bin_bytes"] = os.path.getsize(real_bin)
    rec["real_bin_write_s"] = time.perf_counter() - t0
    del got
if shuffled:
    spath = path + ".shuf.npy"
    if rank == 0:
        t0 = time.perf_counter()
        sm = np.lib.format.open_memmap(spath, mode="w+", dtype=np.uint8, shape=(n,))
        rng = np.random.default_rng(6)
        k = 24                                           # per-record permutation (batch_factorize.py:209-270): 24 records
        for r in range(k):
            a, b = r * n // k, (r + 1) * n // k
            sm[a:b] = rng.permutation(np.asarray(text[a:b]))
        sm.flush()
        del sm
        print(f"[c5] shuffled control generated in {time.perf_counter() - t0:.1f} s", flush=True)
    dist.barrier()
    stext = np.load(spath, mmap_mode="r+")
    pin_shuf = pin_slice(stext)
    got = one(stext, "shuffled")
    if rank == 0:
        chk = wl.verify_factors_sample(np.asarray(stext), got, 100_000)
        rec["check_shuffled"] = chk
        from nolzss_b200 import genomics

        shuf_bin = "/dev/shm/nlz_c5_shuf.bin"
        L.check(lib.nlz_write_factor_file(shuf_bin.encode(), got.ctypes.data, len(got), b"\0", 1, 1, 0, n))
        rec["shuf_bin_bytes"] = os.path.getsize(shuf_bin)
        del got
        # the consumer of configs[4]: factor-length threshold from the two factor files
        t0 = time.perf_counter()
        res = genomics.calculate_factor_length_threshold(real_bin, shuf_bin, tau_expected_fp=10.0)
        rec["threshold"] = {k: (int(v) if isinstance(v, (int, np.integer)) else v) for k, v in res.items()
                            if k in ("L_star", "N_real", "N_shuf", "tau_expected_fp")}
        rec["threshold_s"] = time.perf_counter() - t0
        print(f"[c5] shuffled: {chk}; threshold {rec['threshold']} in {rec['threshold_s']:.1f} s", flush=True)
        os.unlink(shuf_bin)
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r2_c5.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec), flush=True)
    for p in (path + ".npy", path + ".npy.shuf.npy", path + ".shuf.npy", "/dev/shm/nlz_c5_real.bin"):
        if os.path.exists(p):
            os.unlink(p)
dist.barrier()
grp.close()
dist.destroy_process_group()
