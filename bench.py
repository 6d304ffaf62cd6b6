#!/usr/bin/env python
"""Headline benchmark: Mbases/s factorized end-to-end (SA + LCP + factors) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload at EVERY N (strong scaling): BASELINE.json configs[3] -- ONE 250 Mbp chromosome-sized synthetic text with
planted tandem / interspersed repeats, reverse-complement mode (`factorize_dna_w_rc`; 500 000 003 indexed suffixes).
N = 1: the single-GPU pipeline.  N > 1: the distributed path (csrc/dist2.cuh: RANK partitioned by position, suffix
array by rank range, bucketed bulk exchanges over NVLink peer memory; one process per GPU, torch.distributed/NCCL only
for the rendezvous, the barrier around the timed region and the max-over-ranks reduction).  Rank 0 compares the sha256
of the triples with the CPU oracle's (tests/golden/c4_250mbp_rc.json) at every N.

One JSON line on stdout (rank 0):
  value   : Mbases/s with the text already resident in HBM (device pointers), CUDA events, max over ranks
  e2e     : the same metric through the C ABI with HOST buffers: pinned H2D of the text and D2H of all factor triples
            (245 MB) inside the timed region (wall clock around the call, max over ranks)
  roofline: dominant kernel class of the headline workload (algorithmic bytes / CUDA-event time of its launches)
  cpu_baseline (N = 1): the CPU oracle (port of the reference algorithm) on a bounded sample, this box's host cores
Extra legs (reported next to the headline, not part of `value`):
  configs1  : configs[1], 5 Mbp RC text, one per GPU (replicas) -- with its own kernel classes and python-list-API time
  configs2  : configs[2], 10 000 records x 10 kbp dealt to the ranks, ONE batch call per rank, arrays gathered
  configs4  : (N = 8) configs[4], 3.1 Gbp RC text: 6.2 * 10^9 suffixes, 33-bit ranks, per-stage ms and GB/s
`--impl reference` times the CPU oracle's parallel mode (the reference's SDSL build is not available offline; see
DESIGN.md) on a bounded sample of the same text and prints the same line shape with "impl": "reference".
"""
from __future__ import annotations

import argparse
import ctypes
import hashlib
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
C3_BASES = 250_000_000
C3_SEED, C3_SCALE = 4, 50.0
C1_BASES = 5_000_000
REF_SAMPLE = 5_000_000          # reference arm: bases per step (a prefix of the headline text)
CPU_SAMPLE = 25_000_000         # cpu_baseline of our arm: one run on this prefix
WORKLOAD = ("configs[3]: ONE 250 Mbp synthetic text with planted tandem/interspersed repeats (20 families up to 500 kbp, "
            "40 tandem arrays up to 5 Mbp), reverse-complement mode (factorize_dna_w_rc), strong scaling over the GPUs")


def _config(world: int, factors) -> dict:
    """Identical in both arms (the reference arm runs a bounded sample of this workload; cpu_baseline.sample says which)."""
    return {"workload": WORKLOAD, "n_bases": C3_BASES, "indexed_suffixes": 2 * C3_BASES + 3, "factors": factors,
            "l2": "inputs larger than L2 (250 MB text, 30 GB working set); no flush needed",
            "parallelism": f"one text across {world} GPU(s)"}


def _headline_text():
    from nolzss_b200 import workloads as wl

    return wl.planted_dna(C3_BASES, C3_SEED, scale=C3_SCALE)


def _gold():
    try:
        with open(os.path.join(ROOT, "tests", "golden", "c4_250mbp_rc.json")) as f:
            return json.load(f)
    except Exception:
        return None


def _traffic(kernel_class: str):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture of THIS round
    (profiles/r2_traffic.json, scripts/summarize_profiles.py); None when that kernel was not captured."""
    names = {"tile_sort": "k_tile_sort", "gather_rank": "k_gather_rank", "lpnf_rank": "k_lpnf_rank", "lcp_kasai": "k_lcp_kasai",
             "radix_scatter": "k_rs_scatter", "node_tables": "k_node_tables", "group_stream": "k_group_stream", "lpnf_hard": "k_lpnf_hard"}
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            t = json.load(f)
        e = t.get(names.get(kernel_class, ""))
        return (e["dram_bytes_per_launch"], e["launches_captured"], t.get("_captured_at", ""), e.get("dram_bytes_per_alg_byte")) if e else (None, 0, "", None)
    except Exception:
        return (None, 0, "", None)


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

    # The reference's own CPU implementation of the path on this box's host cores: its parallel mode (serial index
    # build + chunked chain walk on every host core, convergence merge: src/cpp/parallel_factorizer.cpp:849-984), i.e.
    # what `parallel_factorize_dna_w_rc_to_file` runs.  The full 250 Mbp text takes the CPU ~200 s per step (measured
    # offline: tests/golden/c4_250mbp_rc.json, oracle_seconds), so every step factorizes a bounded sample -- the first
    # REF_SAMPLE bases of the SAME text -- as the bench contract allows; the CPU is FASTER per base on the sample than on
    # the whole text (cache-resident index), so the ratio against it is conservative.  The arm is this repo's CPU port
    # of the reference algorithm (cpu_baseline.kind = "port"): the reference's SDSL build is not available offline.
    t = _headline_text()[:REF_SAMPLE].tobytes()
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
    gold = _gold()
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": _config(args.gpus, gold["factors"] if gold else None),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port",
                         "sample": f"the first {len(t)} bases of the 250 Mbp text per step, {z} factors (oracle port: SA-IS + Kasai "
                                   f"serially, then the per-factor LCP-interval walk on {used} threads with the reference's "
                                   "convergence merge = its parallel mode)",
                         "full_text_once": ({"seconds": gold["oracle_seconds"], "value": C3_BASES / gold["oracle_seconds"] / 1e6,
                                             "unit": UNIT, "cores": 1, "where": "authoring container, scripts/c4_oracle_hash.py"}
                                            if gold else None),
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
def _kernel_classes(ksum, steps, peak):
    return {k: {"ms_per_step": v["ms"] / steps, "launches_per_step": v["launches"] // steps,
                "alg_GB_per_step": v["bytes"] / steps / 1e9,
                "GBps": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 else None,
                "frac_of_hbm_peak": (v["bytes"] / (v["ms"] * 1e-3) / 1e9 / peak) if v["ms"] > 0 else None}
            for k, v in ksum.items() if v["launches"]}


def _roofline(ksum, peak, peak_src, exclude=("dist_barrier",)):
    cand = {k: v for k, v in ksum.items() if k not in exclude and v["launches"]}
    dom = max(cand, key=lambda k: cand[k]["ms"])
    d = cand[dom]
    achieved = d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0.0
    traffic, nl, when, per_alg = _traffic(dom)
    total = sum(v["ms"] for v in ksum.values())
    alg_launch = d["bytes"] / max(d["launches"], 1)
    # the capture ran a smaller text of the same recipe: DRAM bytes per ALGORITHMIC byte carry over, absolute bytes do not
    if per_alg:
        traffic = per_alg * alg_launch
    return {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_per_algorithmic_byte": per_alg, "peak_source": peak_src,
            "traffic_source": (f"ncu --set full, {nl} captured launches (profiles/r2_traffic.json, r2_full.md; {when}): DRAM bytes per "
                               "algorithmic byte there x this run's algorithmic bytes per launch; not re-measured in this run") if traffic else None,
            "kernel_share_of_step": d["ms"] / total if total else None,
            "alg_bytes_per_launch": d["bytes"] / max(d["launches"], 1),
            "avg_launch_us": 1e3 * d["ms"] / max(d["launches"], 1)}


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

    def rmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    lib = L.load()
    ctx = L.context(local)
    gold = _gold()
    x = _headline_text()
    n = len(x)
    h_text = torch.from_numpy(x).pin_memory()
    d_text = h_text.cuda()
    cap = (gold["factors"] if gold else n // 16) + 4096
    h_out = torch.empty((cap, 3), dtype=torch.int64).pin_memory() if rank == 0 else None
    d_out = torch.empty((cap, 3), dtype=torch.int64, device="cuda") if world == 1 else None
    stream = torch.cuda.current_stream()
    cnt = ctypes.c_uint64(0)
    grp = None
    if world > 1:
        from nolzss_b200 import dist as nd

        grp = nd.ProcessGroup(n, L.MODE_DNA_RC, device=local)
        gctx = grp.ctx
    else:
        gctx = ctx

    def stats():
        s = L.Stats()
        L.check(lib.nlz_get_stats(gctx, ctypes.byref(s)))
        return s.as_dict()

    def step_device():
        """text resident in HBM; N = 1: triples stay in HBM; N > 1: count-only call (the slices stay where they are)"""
        if world == 1:
            L.check(lib.nlz_factorize_device(ctx, L.MODE_DNA_RC, d_text.data_ptr(), n, 0, stream.cuda_stream,
                                             d_out.data_ptr(), cap, ctypes.byref(cnt)))
        else:
            L.check(lib.nlz_dist_factorize_into(grp.dist, L.MODE_DNA_RC, d_text.data_ptr(), n, None, 0, ctypes.byref(cnt)))
        return cnt.value

    def step_host():
        """pinned host text in, all triples out to pinned host memory (rank 0)"""
        if world == 1:
            L.check(lib.nlz_factorize_mode_into(ctx, L.MODE_DNA_RC, h_text.data_ptr(), n, 0, h_out.data_ptr(), cap, ctypes.byref(cnt)))
        else:
            L.check(lib.nlz_dist_factorize_into(grp.dist, L.MODE_DNA_RC, h_text.data_ptr(), n,
                                                h_out.data_ptr() if rank == 0 else None, cap if rank == 0 else 0, ctypes.byref(cnt)))
        return cnt.value

    # ---- warm-up
    for _ in range(max(args.warmup, 1)):
        z = step_device()
    step_host()
    launches_per_step = stats()["kernel_launches"]

    # ---- timed: device-resident (value): CUDA events inside the call (ms_total), max over ranks per step
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        barrier()
        z = step_device()
        dev_ms += rmax(stats()["ms_total"])
    barrier()
    t_wall = time.perf_counter() - t_wall0
    stage = stats()

    # ---- timed: end to end through the C ABI with host buffers (e2e): wall clock, max over ranks per step
    e2e_ms = 0.0
    for _ in range(args.steps):
        barrier()
        t0 = time.perf_counter()
        z2 = step_host()
        e2e_ms += rmax((time.perf_counter() - t0) * 1e3)
    barrier()
    clocks = sampler.stop()
    assert z2 == z
    parity = None
    if rank == 0 and gold:
        got = h_out[:z].numpy().view(np.uint64)
        parity = {"factors": int(z), "factors_expected": gold["factors"],
                  "sha256_matches_oracle": bool(int(z) == gold["factors"] and
                                                hashlib.sha256(got.astype("<u8").tobytes()).hexdigest() == gold["sha256_triples_le_u64"]),
                  "oracle": "CPU oracle on the full text, offline: tests/golden/c4_250mbp_rc.json (scripts/c4_oracle_hash.py)"}

    # ---- roofline leg: CUDA events around every launch of every kernel class (rank 0's view at N > 1)
    L.check(lib.nlz_set_profiling(gctx, 1))
    ksum = {}
    prof_steps = 2
    for _ in range(prof_steps):
        barrier()
        step_device()
        for cls in range(lib.nlz_kernel_class_count()):
            name = ctypes.c_char_p(); ms = ctypes.c_double(0); by = L._u64(0); ln = ctypes.c_uint32(0)
            L.check(lib.nlz_get_kernel_stats(gctx, cls, ctypes.byref(name), ctypes.byref(ms), ctypes.byref(by), ctypes.byref(ln)))
            acc = ksum.setdefault(name.value.decode(), {"ms": 0.0, "bytes": 0, "launches": 0})
            acc["ms"] += ms.value; acc["bytes"] += by.value; acc["launches"] += ln.value
    L.check(lib.nlz_set_profiling(gctx, 0))
    barrier()

    legs = {}
    del d_text, d_out
    if grp is not None:
        grp.close()
    torch.cuda.empty_cache()
    if not args.no_legs:
        legs["configs1"] = configs1_leg(args, world, rank, local, dist, barrier, rmax)
        legs["configs2"] = configs2_leg(world, rank, local, dist, barrier, rmax)
        if world == 8 and not args.no_c5:
            try:
                legs["configs4"] = configs4_leg(world, rank, local, dist, barrier, rmax)
            except Exception as e:                              # never lose the headline line to the largest leg
                legs["configs4"] = {"error": repr(e)}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    value = n * args.steps / (dev_ms * 1e-3) / 1e6
    e2e_value = n * args.steps / (e2e_ms * 1e-3) / 1e6
    peak, peak_src = _peaks()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32" if world == 1 else "u64", "data": "synthetic",
        "config": _config(world, int(z)),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n, "d2h_bytes_per_step": int(z) * 24 + 4 * 260,
                "ms_per_step": e2e_ms / args.steps,
                "api": "nlz_factorize_mode_into (C ABI, pinned host buffers)" if world == 1 else
                       "nlz_dist_factorize_into (C ABI, pinned host buffers; every rank uploads one slice, rank 0 gathers the triples)"},
        "gpu_launches": int(launches_per_step) * args.steps,
        "parity": parity,
        "roofline": _roofline(ksum, peak, peak_src),
        "kernel_classes": _kernel_classes(ksum, prof_steps, peak),
        "stages_ms": {k: stage[k] for k in stage if k.startswith("ms_")},
        "pipeline": {k: stage[k] for k in ("key_bits", "sym_bits", "key_syms", "doubling_rounds", "active_sum", "walk_nodes",
                                           "hard_positions", "host_syncs", "workspace_bytes", "rank_records_applied")},
        "wall_s_timed_region": t_wall,
    }
    line.update(legs)

    if world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_py as orc
        from nolzss_b200 import workloads as wl

        sample = x[:CPU_SAMPLE].tobytes()
        S = wl.prepare_w_rc_single(sample)
        t0 = time.perf_counter()
        f = orc.factorize_multiple_dna_w_rc(S)
        dt = time.perf_counter() - t0
        got_s = L.factorize_array(L.MODE_DNA_RC, sample, device=local)
        cores = os.cpu_count() or 1
        t0 = time.perf_counter()
        fp, used = orc.parallel_factorize_multiple_dna_w_rc(S, cores)
        dtp = time.perf_counter() - t0
        idx_s, walk_s = orc.last_timing()
        line["cpu_baseline"] = {
            "value": len(sample) / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"the first {len(sample)} bases of the 250 Mbp text, once (CPU oracle: SA-IS + Kasai + per-factor "
                      "LCP-interval walk); the full text needs ~200 s (tests/golden/c4_250mbp_rc.json)",
            "seconds": dt, "host_cores_available": cores,
            "triples_identical_to_gpu": bool(np.array_equal(f, got_s)),
            "full_text_once": ({"seconds": gold["oracle_seconds"], "value": C3_BASES / gold["oracle_seconds"] / 1e6, "unit": UNIT,
                                "cores": 1, "where": "authoring container, scripts/c4_oracle_hash.py"} if gold else None),
            "reference_published": {"value": 0.037, "unit": UNIT,
                                    "note": "~27 s per Mbp for factorize_fasta_multiple_dna_w_rc (RC mode), "
                                            "benchmarks/README.md:293-294 of the reference: published, hardware "
                                            "unknown, not reproduced (the SDSL build is not available offline)"},
            "parallel_mode": {"value": len(sample) / dtp / 1e6, "unit": UNIT, "cores": used, "seconds": dtp,
                              "serial_index_seconds": idx_s, "threaded_walk_seconds": walk_s,
                              "triples_identical_to_gpu": bool(np.array_equal(fp, got_s)),
                              "what": "the reference's CPU parallel mode (parallel_factorizer.cpp:849-984): one index "
                                      "built serially, chain walk on all host threads, convergence merge"},
        }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------- configs[1]: 5 Mbp text per GPU (replicas)
def configs1_leg(args, world, rank, local, dist, barrier, rmax, steps=10):
    """BASELINE.json configs[1]: 5 Mbp bacterial-genome-sized text with planted repeats, RC mode, one text per GPU
    (seed 2 + rank; replicas: no data-path collective).  Device-resident and host-buffer timings, kernel classes of one
    profiled step, and the time of the python LIST API a drop-in user calls (`_noLZSS.factorize_dna_w_rc`)."""
    import torch

    from nolzss_b200 import _lib as L
    from nolzss_b200 import workloads as wl

    lib = L.load()
    ctx = L.context(local)
    text = wl.c2_text(C1_BASES, 2 + rank)
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
        L.check(lib.nlz_factorize_device(ctx, L.MODE_DNA_RC, d_text.data_ptr(), n, 0, stream.cuda_stream, d_out.data_ptr(), cap, ctypes.byref(cnt)))
        return cnt.value

    def step_host():
        L.check(lib.nlz_factorize_mode_into(ctx, L.MODE_DNA_RC, h_text.data_ptr(), n, 0, h_out.data_ptr(), cap, ctypes.byref(cnt)))
        return cnt.value

    for _ in range(3):
        flush.zero_()
        z = step_device()
    step_host()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.zero_()                       # L2 flush between timed iterations (not timed)
        a.record(stream)
        z = step_device()
        b.record(stream)
    barrier()
    dev_ms = rmax(sum(a.elapsed_time(b) for a, b in ev))
    stage = L.stats(local)
    e2e_s = 0.0
    for _ in range(steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step_host()
        e2e_s += time.perf_counter() - t0
    e2e_ms = rmax(e2e_s * 1e3)
    L.set_profiling(True, local)
    ksum = {}
    for _ in range(3):
        flush.zero_()
        step_device()
        for name, v in L.kernel_stats(local).items():
            acc = ksum.setdefault(name, {"ms": 0.0, "bytes": 0, "launches": 0})
            for k in acc:
                acc[k] += v[k]
    L.set_profiling(False, local)
    out = None
    if rank == 0:
        from nolzss_b200 import _noLZSS as ext

        t0 = time.perf_counter()
        lst = ext.factorize_dna_w_rc(text)
        py_s = time.perf_counter() - t0
        peak, peak_src = _peaks()
        out = {"workload": "configs[1]: 5 Mbp synthetic DNA with planted repeats (20 interspersed families, 40 tandem arrays), RC mode, "
                           "one text per GPU (replicas, weak scaling)", "n_bases_per_gpu": n, "n_gpus": world, "factors": int(z),
               "value": world * n * steps / (dev_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": dev_ms / steps,
               "e2e": {"value": world * n * steps / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms / steps,
                       "h2d_bytes_per_step": n, "d2h_bytes_per_step": int(z) * 24 + 4 * 260},
               "python_list_api": {"call": "_noLZSS.factorize_dna_w_rc(bytes) -> list of 4-tuples", "seconds": py_s,
                                   "value": n / py_s / 1e6, "unit": UNIT, "tuples": len(lst)},
               "l2": "flushed (512 MiB write) between timed steps",
               "roofline": _roofline(ksum, peak, peak_src), "kernel_classes": _kernel_classes(ksum, 3, peak),
               "stages_ms": {k: stage[k] for k in stage if k.startswith("ms_")}, "host_syncs": stage["host_syncs"],
               "doubling_rounds": stage["doubling_rounds"]}
    del d_text, d_out, flush
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------- configs[2]: records dealt to the ranks
def configs2_leg(world, rank, local, dist, barrier, rmax, steps=3):
    """BASELINE.json configs[2]: 10 000 records x 10 kbp, per-sequence RC factorization; the records are dealt to the
    ranks by `sharding.assign_records`, every rank runs ONE nlz_factorize_batch call over its share, counts (and, once,
    the triples) are gathered as arrays.  Parity: sha256 of the per-record counts and of all triples against the CPU
    oracle's (tests/golden/c3_10000x10k_rc.json)."""
    import numpy as np

    from nolzss_b200 import _lib as L
    from nolzss_b200 import sharding
    from nolzss_b200 import workloads as wl

    recs = [s for _, s in wl.c3_records(10_000, 10_000, seed=3)]
    nb = sum(len(s) for s in recs)
    group = None
    counts = trip = None
    times, dev = [], []
    for it in range(steps + 1):
        barrier()
        t0 = time.perf_counter()
        counts, tr = sharding.factorize_batch_distributed(recs, True, want_factors=(it == 0), group=group, device=local)
        dt = time.perf_counter() - t0
        if it == 0:
            trip = tr                                         # the triples are gathered once (parity), the timed steps count
        st = L.stats(local)
        if it:
            times.append(rmax(dt * 1e3))
            dev.append(rmax(st["ms_total"]))
    if rank != 0:
        return None
    out = {"workload": "configs[2]: 10 000 records x 10 kbp (a 500-bp segment copied inside each, 50 % reverse-complemented), "
                       "per-sequence RC mode, records dealt to the ranks (length-balanced), one batch call per rank, array gathers",
           "records": len(recs), "n_bases": nb, "n_gpus": world, "total_factors": int(counts.sum()),
           "device_ms_max_over_ranks": sum(dev) / len(dev), "value": nb / (sum(dev) / len(dev)) / 1e3, "unit": UNIT,
           "wall_ms_incl_h2d_and_count_gather": sum(times) / len(times),
           "e2e": {"value": nb / (sum(times) / len(times)) / 1e3, "unit": UNIT, "what": "count path: host records -> per-record counts on every rank"}}
    try:
        with open(os.path.join(ROOT, "tests", "golden", "c3_10000x10k_rc.json")) as f:
            g = json.load(f)
        out["parity"] = {"counts_sha256_matches_oracle": hashlib.sha256(counts.astype("<i8").tobytes()).hexdigest() == g["sha256_counts_le_i64"],
                         "triples_sha256_matches_oracle": hashlib.sha256(np.ascontiguousarray(trip).astype("<u8").tobytes()).hexdigest() == g["sha256_triples_le_u64"],
                         "total_factors_expected": g["total_factors"]}
    except Exception as e:
        out["parity"] = {"error": repr(e)}
    return out


# ------------------------------------------------------------------------------- configs[4]: 3.1 Gbp on 8 GPUs
def configs4_leg(world, rank, local, dist, barrier, rmax, runs=2):
    """BASELINE.json configs[4]: ONE 3.1 Gbp human-genome-sized synthetic text in RC mode -- 6 200 000 003 indexed
    suffixes, 33-bit global ranks and S-positions -- across the 8 GPUs (csrc/dist2.cuh).  The text is generated once
    into /dev/shm and mapped by every rank (a rank uploads only its slice).  Checks: the factors tile the text, 100 000
    sampled factors are true (reverse-complement) matches.  Per-stage ms and algorithmic GB/s from the kernel classes."""
    import numpy as np
    import torch

    from nolzss_b200 import _lib as L
    from nolzss_b200 import dist as nd
    from nolzss_b200 import workloads as wl

    n = wl.C5_BASES
    path = f"/dev/shm/nlz_c5_{n}.npy"
    t0 = time.perf_counter()
    if rank == 0:
        mm = np.lib.format.open_memmap(path, mode="w+", dtype=np.uint8, shape=(n,))
        wl.c5_text_into(mm, n)
        mm.flush()
        del mm
    barrier()
    gen_s = time.perf_counter() - t0
    text = np.load(path, mmap_mode="r+")          # (writable mapping: a read-only one cannot be page-locked on this platform)
    lib = L.load()
    # page-lock the slice of the mapped text this rank uploads (DNA_RC: every rank uploads ONE slice, csrc/dist2_host.cuh):
    # a pageable upload of 388 MB costs ~370 ms of the 1.5 s run, a pinned one ~10 ms
    per = ((n + world - 1) // world + 255) // 256 * 256
    lo, hi = min(rank * per, n), min((rank + 1) * per, n)
    base = text.ctypes.data
    a0 = (base + lo) // 4096 * 4096
    a1 = (base + hi + 4095) // 4096 * 4096
    pinned = lib.nlz_host_register(a0, a1 - a0) == L.NLZ_OK
    grp = nd.ProcessGroup(n, L.MODE_DNA_RC, device=local)
    out = {"workload": f"configs[4]: c5_text_into(n={n}, seed=5): planted repeats (240 families <= 500 kbp, 480 tandem arrays <= 5 Mbp), "
                       f"RC mode, {2 * n + 3} indexed suffixes (33-bit ranks), one text across {world} GPUs",
           "n_bases": n, "n_gpus": world, "text_generation_s": gen_s, "runs": []}
    got = None
    try:
        for it in range(runs):
            prof = it == runs - 1
            L.check(lib.nlz_set_profiling(grp.ctx, 1 if prof else 0))
            barrier()
            t0 = time.perf_counter()
            got, z = grp.factorize(L.MODE_DNA_RC, text)
            wall = time.perf_counter() - t0
            st = grp.stats()
            ms = rmax(st["ms_total"])
            if rank == 0:
                r = {"it": it, "factors": int(z), "device_ms_max_over_ranks": ms, "value": n / ms / 1e3, "unit": UNIT,
                     "wall_s_rank0_incl_h2d_and_d2h_of_all_triples": wall, "text_slice_pinned": pinned,
                     "stages_ms_rank0": {k: st[k] for k in st if k.startswith("ms_")}, "doubling_rounds": st["doubling_rounds"],
                     "workspace_GiB_rank0": st["workspace_bytes"] / 2**30, "profiled": prof}
                if prof:
                    peak, _ = _peaks()
                    ks = {}
                    for cls in range(lib.nlz_kernel_class_count()):
                        name = ctypes.c_char_p(); kms = ctypes.c_double(0); by = L._u64(0); ln = ctypes.c_uint32(0)
                        L.check(lib.nlz_get_kernel_stats(grp.ctx, cls, ctypes.byref(name), ctypes.byref(kms), ctypes.byref(by), ctypes.byref(ln)))
                        if ln.value:
                            ks[name.value.decode()] = {"ms": kms.value, "launches": ln.value, "alg_GB": by.value / 1e9,
                                                       "GBps": by.value / max(kms.value, 1e-9) / 1e6,
                                                       "frac_of_hbm_peak": by.value / max(kms.value, 1e-9) / 1e6 / peak}
                    r["kernel_classes_rank0"] = ks
                out["runs"].append(r)
        L.check(lib.nlz_set_profiling(grp.ctx, 0))
        if rank == 0:
            out["check"] = wl.verify_factors_sample(np.asarray(text), got, 100_000)
            out["rc_factors"] = int((got[:, 2] >> np.uint64(63)).sum())
            out["max_factor_length"] = int(got[:, 1].max())
            out["value"] = max(r["value"] for r in out["runs"] if not r["profiled"]) if runs > 1 else out["runs"][0]["value"]
    finally:
        barrier()
        grp.close()
        if pinned:
            lib.nlz_host_unregister(a0)
        del text
        if rank == 0 and os.path.exists(path):
            os.unlink(path)
    return out if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="headline only (skip the configs[1] / configs[2] / configs[4] legs)")
    ap.add_argument("--no-c5", action="store_true", help="skip the configs[4] leg (3.1 Gbp on 8 GPUs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
