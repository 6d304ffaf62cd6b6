"""Turns ncu outputs brought back under gpurun_out/ into the small text summaries kept in profiles/.

    python scripts/summarize_profiles.py launches <launches.csv> <out.md>
    python scripts/summarize_profiles.py full <report.ncu-rep> <out.md>
"""
import csv
import collections
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]


def launches(src, dst):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "")
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += float(r[vi].replace(",", ""))
    total = sum(v[1] for v in tot.values())
    mine = {k: v for k, v in tot.items() if "nlz::" in k or k.startswith("k_")}
    mtotal = sum(v[1] for v in mine.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)\n\n")
        f.write(f"source: {src}; {sum(v[0] for v in tot.values())} launches, {total/1e6:.3f} ms total, "
                f"{mtotal/1e6:.3f} ms in nlz:: kernels\n\n| kernel | launches | total us | share of nlz:: time |\n|---|---:|---:|---:|\n")
        for k, v in sorted(mine.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]/1e3:.1f} | {100*v[1]/mtotal:.1f}% |\n")
        f.write("\nother kernels in the process (torch fill used as L2 flush etc.):\n\n")
        for k, v in tot.items():
            if k not in mine:
                f.write(f"- `{k}`: {v[0]} launches, {v[1]/1e3:.1f} us\n")


def _raw(src):
    """raw-page CSV of a report: either the .ncu-rep itself or a CSV exported on the GPU box (reports of more than a few
    dozen launches exceed what gpurun brings back)"""
    if src.endswith(".csv"):
        return open(src).read()
    return subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout


def full(src, dst):
    out = _raw(src)
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for r in rows[2:]:
            f.write(f"## `{r[ki].split('(')[0]}`  (ID {r[0]})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")
            f.write("\n")


def traffic(src, dst):
    """mean dram__bytes_read.sum + dram__bytes_write.sum per captured launch of every kernel -> JSON (bench.py's roofline.traffic)"""
    import json
    out = _raw(src)
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = collections.OrderedDict()
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "").split("<")[0].replace("nlz::", "")
        b = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
        a = acc.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += b
    with open(dst, "w") as f:
        d = {k: {"launches_captured": v[0], "dram_bytes_per_launch": v[1] / v[0]} for k, v in acc.items()}
        d["_captured_at"] = sys.argv[4] if len(sys.argv) > 4 else ""
        json.dump(d, f, indent=1)


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
