#!/usr/bin/env python
"""Headline benchmark: Mbases/s factorized end-to-end (SA + LCP + factors) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload at every N: BASELINE.json configs[1] -- 5 Mbp bacterial-genome-sized synthetic DNA with
planted repeats, reverse-complement mode (noLZSS `factorize_dna_w_rc`), one such text per GPU
(seed 2 + rank; weak scaling, no data-path collective: the path shards per text/record).

One JSON line on stdout (rank 0):
  value   : whole-job Mbases/s with the text already resident in HBM (device entry point)
  e2e     : same metric through the C ABI with HOST buffers (pinned H2D of the text and D2H of the
            factor triples inside the timed region)
  roofline: dominant kernel class, algorithmic bytes / CUDA-event time of its launches
  cpu_baseline: the CPU oracle (port of the reference algorithm) on this box's host cores
`--impl reference` times that CPU oracle alone (the reference's SDSL build is not available
offline) and prints the same line shape with "impl": "reference".
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mbases/s factorized end-to-end (SA+LCP+factors)"
UNIT = "Mbases/s"
N_BASES = 5_000_000
WORKLOAD = ("configs[1]: 5 Mbp synthetic DNA with planted repeats (20 interspersed families, 40 tandem "
            "arrays), reverse-complement mode (factorize_dna_w_rc), one text per GPU")


def _text_for_rank(rank: int) -> bytes:
    from nolzss_b200 import workloads as wl

    return wl.c2_text(N_BASES, 2 + rank)


def _traffic(kernel_class: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu
    --set full capture (profiles/r1_traffic.json, written by scripts/summarize_profiles.py); None if not captured."""
    names = {"tile_sort": "k_tile_sort", "gather_rank": "k_gather_rank", "lpnf_rank": "k_lpnf_rank",
             "lcp_kasai": "k_lcp_kasai", "radix_scatter": "k_rs_scatter", "node_tables": "k_node_tables"}
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            t = json.load(f)
        e = t.get(names.get(kernel_class, ""))
        return (e["dram_bytes_per_launch"], e["launches_captured"]) if e else (None, 0)
    except Exception:
        return (None, 0)


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as orc
    from nolzss_b200 import workloads as wl

    # every step factorizes the FULL 5 Mbp text of configs[1] -- the same config as our arm (about 1.3 s of CPU work
    # per step) -- through the reference's own parallel mode (serial index build + chunked chain walk on every host
    # core, convergence merge: src/cpp/parallel_factorizer.cpp:849-984), which is what
    # `parallel_factorize_dna_w_rc_to_file` would run.  The arm is this repo's CPU port of the reference algorithm
    # (cpu_baseline.kind = "port"): the reference's SDSL build is not available offline.
    t = _text_for_rank(0)
    S = wl.prepare_w_rc_single(t)
    cores = os.cpu_count() or 1
    times = []
    z = 0
    used = 1
    index_s = walk_s = 0.0
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        f, used = orc.parallel_factorize_multiple_dna_w_rc(S, cores)
        dt = time.perf_counter() - t0
        z = len(f)
        if it >= args.warmup:
            times.append(dt)
            a, b = orc.last_timing()
            index_s += a
            walk_s += b
    total = sum(times)
    value = len(t) * len(times) / total / 1e6
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_bases_per_gpu": N_BASES, "sample_bases_per_step": len(t), "factors": z,
                   "same_config_as_ours": len(t) == N_BASES},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port",
                         "sample": f"the full {len(t)}-base text of configs[1] per step (oracle port: SA-IS + Kasai "
                                   f"serially, then the per-factor LCP-interval walk on {used} threads with the "
                                   "reference's convergence merge = its parallel mode; the reference's SDSL path "
                                   "cannot be built offline)",
                         "serial_index_fraction": index_s / max(index_s + walk_s, 1e-12),
                         "host_cores_available": cores},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        # NVML from a thread of this process (first sample within a millisecond, one every 5 ms); the nvidia-smi
        # loop below is the fallback -- on an 8-GPU box it needs longer to start than a 50 ms timed region lasts
        self.nvml = None
        try:
            import threading

            import pynvml
            import torch

            pynvml.nvmlInit()
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(self.device).uuid))
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            smax = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            state = {"stop": False, "clk": [], "reasons": 0, "smax": smax, "n": 0}

            def loop():
                while not state["stop"]:
                    try:
                        state["clk"].append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        state["reasons"] |= pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        state["n"] += 1
                    except Exception:
                        pass
                    time.sleep(0.005)

            th = threading.Thread(target=loop, daemon=True)
            th.start()
            self.nvml = (pynvml, state, th)
            return
        except Exception:
            self.nvml = None
        try:
            fd, self.path = tempfile.mkstemp(prefix="nlz_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if getattr(self, "nvml", None):
            pynvml, state, th = self.nvml
            state["stop"] = True
            th.join(timeout=1.0)
            bits = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}
            if state["clk"]:
                out.update(sm_mhz=float(statistics.median(state["clk"])), sm_max_mhz=float(state["smax"]),
                           reasons=sorted(k for k, b in bits.items() if state["reasons"] & b), samples=state["n"],
                           source="nvml, 5 ms period, during the timed regions")
            return out
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            rows = [r.strip().split(",") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
        except Exception:
            return out
        sm, reasons, smax = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 7:
                continue
            try:
                clk, mx, util = float(r[0]), float(r[1]), float(r[6])
            except ValueError:
                continue
            smax = mx
            if util > 0:
                sm.append(clk)
            for name, v in zip(names, r[2:6]):
                if v.strip() == "Active":
                    reasons.add(name)
        allclk = sm or [float(r[0]) for r in rows if len(r) >= 7 and r[0].strip().replace(".", "").isdigit()]
        out.update(sm_mhz=statistics.median(allclk) if allclk else None, sm_max_mhz=smax,
                   reasons=sorted(reasons), samples=len(rows))
        return out


# ------------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch

    from nolzss_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    lib = L.load()
    ctx = L.context(local)
    text = _text_for_rank(rank)
    n = len(text)
    cap = n // 2 + 1024
    d_text = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    d_out = torch.empty((cap, 3), dtype=torch.int64, device="cuda")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    h_text = torch.frombuffer(bytearray(text), dtype=torch.uint8).pin_memory()
    h_out = torch.empty((cap, 3), dtype=torch.int64).pin_memory()
    stream = torch.cuda.current_stream()
    cnt = ctypes.c_uint64(0)

    def step_device():
        L.check(lib.nlz_factorize_device(ctx, L.MODE_DNA_RC, d_text.data_ptr(), n, 0, stream.cuda_stream,
                                         d_out.data_ptr(), cap, ctypes.byref(cnt)))
        return cnt.value

    def step_host():
        L.check(lib.nlz_factorize_mode_into(ctx, L.MODE_DNA_RC, h_text.data_ptr(), n, 0, h_out.data_ptr(), cap,
                                            ctypes.byref(cnt)))
        return cnt.value

    # ---- warm-up
    for _ in range(max(args.warmup, 1)):
        flush.zero_()
        z = step_device()
    for _ in range(max(args.warmup, 1)):
        step_host()
    launches_per_step = L.stats(local)["kernel_launches"]

    # ---- timed: device-resident (value)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        flush.zero_()                       # L2 flush between timed iterations (not timed)
        a.record(stream)
        z = step_device()
        b.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    stage = L.stats(local)

    # ---- timed: end to end through the C ABI with host buffers (e2e)
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        z2 = step_host()
        e2e_s += time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()
    assert z2 == z

    # ---- roofline leg: CUDA events around every launch of every kernel class
    L.set_profiling(True, local)
    ksum = {}
    prof_steps = 3
    for _ in range(prof_steps):
        flush.zero_()
        step_device()
        for name, v in L.kernel_stats(local).items():
            acc = ksum.setdefault(name, {"ms": 0.0, "bytes": 0, "launches": 0})
            for k in acc:
                acc[k] += v[k]
    L.set_profiling(False, local)

    # ---- reduce over ranks (max time)
    times = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max = times.tolist()
    one_text = None
    if not args.no_single_text:
        del d_text, d_out, flush
        torch.cuda.empty_cache()
        one_text = single_text_leg(world, rank, local, dist)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    value = world * n * args.steps / (dev_ms_max * 1e-3) / 1e6
    e2e_value = world * n * args.steps / (e2e_ms_max * 1e-3) / 1e6
    peak, peak_src = _peaks()
    dom = max(ksum, key=lambda k: ksum[k]["ms"])
    d = ksum[dom]
    achieved = d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0.0
    traffic, traffic_launches = _traffic(dom)
    total_kernel_ms = sum(v["ms"] for v in ksum.values())
    classes = {k: {"ms_per_step": v["ms"] / prof_steps, "launches_per_step": v["launches"] // prof_steps,
                   "alg_GB_per_step": v["bytes"] / prof_steps / 1e9,
                   "GBps": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 else None,
                   "frac_of_hbm_peak": (v["bytes"] / (v["ms"] * 1e-3) / 1e9 / peak) if v["ms"] > 0 else None}
               for k, v in ksum.items() if v["launches"]}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_bases_per_gpu": n, "indexed_suffixes_per_gpu": stage["n_suffixes"],
                   "factors": z, "l2": "flushed (512 MiB write) between timed steps",
                   "parallelism": f"{world} independent texts, one per GPU"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n, "d2h_bytes_per_step": int(z) * 24 + 4 * 260,
                "ms_per_step": e2e_ms_max / args.steps},
        "gpu_launches": int(launches_per_step) * args.steps,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "traffic_source": (f"ncu --set full, mean of {traffic_launches} captured launches "
                                        "(profiles/r1_full.md, profiles/r1_traffic.json)") if traffic else None,
                     "kernel_share_of_step": d["ms"] / total_kernel_ms if total_kernel_ms else None,
                     "alg_bytes_per_launch": d["bytes"] / max(d["launches"], 1),
                     "avg_launch_us": 1e3 * d["ms"] / max(d["launches"], 1)},
        "kernel_classes": classes,
        "stages_ms": {k: stage[k] for k in stage if k.startswith("ms_")},
        "pipeline": {k: stage[k] for k in ("key_bits", "sym_bits", "key_syms", "doubling_rounds", "active_sum",
                                           "walk_nodes", "host_syncs", "workspace_bytes")},
        "wall_s_timed_region": t_wall,
    }
    if one_text is not None:
        line["single_text_all_gpus"] = one_text

    if world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_py as orc
        from nolzss_b200 import workloads as wl

        S = wl.prepare_w_rc_single(text)
        t0 = time.perf_counter()
        f = orc.factorize_multiple_dna_w_rc(S)
        dt = time.perf_counter() - t0
        got = h_out[:z].numpy().view(np.uint64)
        # the reference's parallel mode on every host core (serial index build + threaded chain walk)
        cores = os.cpu_count() or 1
        t0 = time.perf_counter()
        fp, used = orc.parallel_factorize_multiple_dna_w_rc(S, cores)
        dtp = time.perf_counter() - t0
        idx_s, walk_s = orc.last_timing()
        line["cpu_baseline"] = {
            "value": n / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "the full 5 Mbp text, once (CPU oracle: SA-IS + Kasai + per-factor LCP-interval walk)",
            "seconds": dt, "host_cores_available": cores,
            "triples_identical_to_gpu": bool(len(f) == z and np.array_equal(f, got)),
            "reference_published": {"value": 0.037, "unit": UNIT,
                                    "note": "~27 s per Mbp for factorize_fasta_multiple_dna_w_rc (RC mode), "
                                            "benchmarks/README.md:293-294 of the reference: published, hardware "
                                            "unknown, not reproduced (the SDSL build is not available offline)"},
            "parallel_mode": {"value": n / dtp / 1e6, "unit": UNIT, "cores": used, "seconds": dtp,
                              "serial_index_seconds": idx_s, "threaded_walk_seconds": walk_s,
                              "triples_identical_to_gpu": bool(len(fp) == z and np.array_equal(fp, got)),
                              "what": "the reference's CPU parallel mode (parallel_factorizer.cpp:849-984): one index "
                                      "built serially, chain walk on all host threads, convergence merge"},
        }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------- configs[3]: one text, all GPUs
C3_BASES = 250_000_000


def single_text_leg(world, rank, local, dist, steps=2):
    """configs[3]: ONE 250 Mbp chromosome-sized text (planted repeats scaled x50), RC mode, across all `world` GPUs
    (nolzss_b200.dist: rank-range-partitioned suffix array over peer memory); at N = 1 the ordinary single-GPU
    pipeline.  Count-only calls from pageable host text; time = CUDA events inside the call, max over ranks."""
    import torch

    from nolzss_b200 import _lib as L
    from nolzss_b200 import dist as nd
    from nolzss_b200 import workloads as wl

    text = wl.planted_dna(C3_BASES, 4, scale=50.0).tobytes()
    out = {"workload": "configs[3]: one 250 Mbp synthetic text with planted tandem/interspersed repeats, RC mode "
                       "(n' = 500 000 003 suffixes), partitioned over all GPUs", "n_bases": C3_BASES, "n_gpus": world}
    times = []
    if world == 1:
        for it in range(steps + 1):
            z = L.count(L.MODE_DNA_RC, text, device=local)
            st = L.stats(local)
            if it:
                times.append(st["ms_total"])
    else:
        grp = nd.ProcessGroup(C3_BASES, L.MODE_DNA_RC, device=local)
        lib = L.load()
        addr, n, keep = L._as_buffer(text)
        for it in range(steps + 1):
            dist.barrier()
            cnt = ctypes.c_uint64(0)
            L.check(lib.nlz_dist_factorize(grp.dist, L.MODE_DNA_RC, addr, n, None, ctypes.byref(cnt)))
            z = cnt.value
            st = grp.stats()
            t = torch.tensor([st["ms_total"]], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if it:
                times.append(t.item())
        grp.close()
    ms = sum(times) / len(times)
    out.update(ms_per_text=ms, value=C3_BASES / ms / 1e3, unit=UNIT, factors=int(z),
               stages_ms_rank0={k: st[k] for k in st if k.startswith("ms_")}, doubling_rounds=st["doubling_rounds"],
               workspace_bytes_rank0=st["workspace_bytes"])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-single-text", action="store_true", help="skip the configs[3] leg (one 250 Mbp text over all GPUs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
