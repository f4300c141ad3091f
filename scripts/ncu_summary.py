#!/usr/bin/env python
"""Summarise ncu outputs (read here, no GPU needed) into the tracked profiles/ directory.

  scripts/ncu_summary.py launches <launches.csv> <out.md>      per-kernel share of the captured window
  scripts/ncu_summary.py full <prof.ncu-rep> <out.md>          key metrics of every launch in a --set full capture
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__cycles_active.avg"]


def short(name):
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*", "", name)


def launches(path, out):
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    agg = collections.OrderedDict()
    for r in rows:
        d = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
        d[0] += 1
        d[1] += float(r["Metric Value"]) / 1e3
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({path})\n\n{len(rows)} launches captured, {tot / 1e3:.2f} ms of device time "
                "(gpu__time_duration.sum, cold-cache, serialised: compare SHARES, not absolutes).\n\n"
                "| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |\n")


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary ({path})\n\n")
        for r in data:
            f.write(f"## {short(r[idx['Kernel Name']])}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
            f.write("\n")


def traffic(out, *paths):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of every kernel in the given --set full
    captures -> JSON that bench.py reads for `roofline.traffic`"""
    import json
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tmult = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    res = {}
    for path in paths:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in data:
            name = re.sub(r"<.*", "", short(r[idx["Kernel Name"]]))
            rd = float(r[idx["dram__bytes_read.sum"]]) * mult[units[idx["dram__bytes_read.sum"]]]
            wr = float(r[idx["dram__bytes_write.sum"]]) * mult[units[idx["dram__bytes_write.sum"]]]
            us = float(r[idx["gpu__time_duration.sum"]]) * tmult[units[idx["gpu__time_duration.sum"]]]
            tp = float(r[idx["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])
            res.setdefault(name, []).append(dict(kernel=short(r[idx["Kernel Name"]]), grid=r[idx["Grid Size"]], dram_read_bytes=rd,
                                                 dram_write_bytes=wr, time_us=us, tensor_pipe_active_pct=tp))
    summary = {k: dict(captured_launches=len(v), dram_bytes_per_launch=sum(x["dram_read_bytes"] + x["dram_write_bytes"] for x in v) / len(v),
                       launches=v, source=[p for p in paths]) for k, v in res.items()}
    json.dump(summary, open(out, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], *sys.argv[3:])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
