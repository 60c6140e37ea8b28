"""Summarise ncu outputs into small tracked files under profiles/ (gpurun_out/ is scratch).
  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/out.json
  python tools/ncu_summary.py full gpurun_out/x.ncu-rep profiles/out.json
"""
import csv
import io
import json
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", name)
    return name.strip()[:90]


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 10]
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        unit = r[hdr.index("Metric Unit")]
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        a = agg.setdefault(short(r[kn]), [0, 0.0])
        a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values())
    res = {"source": path, "note": "ncu --metrics gpu__time_duration.sum --clock-control none: per-launch times are "
                                   "cold-cache and serialised; compare SHARES, not absolutes",
           "total_us": tot,
           "kernels": sorted(({"kernel": k, "launches": a[0], "us": a[1], "share": a[1] / tot} for k, a in agg.items()),
                             key=lambda d: -d["us"])}
    json.dump(res, open(out, "w"), indent=1)
    for k in res["kernels"][:12]:
        print(f"{k['share'] * 100:6.2f}%  {k['launches']:5d}  {k['us']:12.1f} us  {k['kernel']}")


def full(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res = {"source": path, "launches": []}
    for r in rows[2:]:
        d = {"kernel": short(r[hdr.index("Kernel Name")])}
        for k in KEYS:
            if k in hdr:
                d[k] = r[hdr.index(k)] + " " + units[hdr.index(k)]
        res["launches"].append(d)
    json.dump(res, open(out, "w"), indent=1)
    for d in res["launches"][:8]:
        print({k: v for k, v in d.items() if k in ("kernel", "gpu__time_duration.sum", "dram__bytes_read.sum",
                                                    "dram__bytes_write.sum", "launch__grid_size",
                                                    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                                                    "sm__inst_executed_pipe_tensor.sum")})


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
