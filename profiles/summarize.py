#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into small text summaries that are committed under profiles/.

    python profiles/summarize.py <tag>

Reads  gpurun_out/launches_<tag>.csv   (ncu --metrics gpu__time_duration.sum launch list)
       gpurun_out/prof_<tag>.ncu-rep   (ncu --set full capture, optional)
Writes profiles/launches_<tag>.md      (per-kernel time share of one step, library kernels only)
       profiles/ncu_<tag>.md           (per-launch DRAM bytes / throughput / pipe utilisation)
       profiles/traffic.json           (kernel-name -> measured dram bytes per launch; bench.py reads it)
"""
import csv
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

OURS = ("k1_", "k1m_", "k2_", "k3_", "k4_", "geom_", "mvs_relative", "vis_homography", "pack_")

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % active"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem (CTAs)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1/shared data pipe (LSU wavefronts) %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared data pipe: LSU wavefronts (LDS / STS) % of peak"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared data pipe: tensor-core operand wavefronts % of peak"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "warps stalled: fixed-latency wait / issue"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "warps stalled: long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "warps stalled: short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "warps stalled: math pipe throttle / issue"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def short(name):
    n = name.replace("void ", "").replace("mvsb200::", "")
    return n.split("(")[0]


def launches(tag):
    path = os.path.join(OUT, "launches_%s.csv" % tag)
    if not os.path.exists(path):
        return
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((short(r["Kernel Name"]), us, r["Grid Size"], r["Block Size"]))
    agg = OrderedDict()
    for n, us, g, b in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    total_ours = sum(v[1] for k, v in agg.items() if k.startswith(OURS))
    total = sum(v[1] for v in agg.values())
    with open(os.path.join(PROF, "launches_%s.md" % tag), "w") as f:
        f.write("# ncu launch list `%s` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n" % tag)
        f.write("command: `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e` (first 400 launches)\n\n")
        f.write("library kernels: %.1f us of %.1f us captured (rest = torch memset/copy/packing glue)\n\n" % (total_ours, total))
        f.write("| kernel | launches | total us | mean us | share of library time |\n|---|---|---|---|---|\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if k.startswith(OURS):
                f.write("| `%s` | %d | %.1f | %.1f | %.1f %% |\n" % (k, n, us, us / n, 100 * us / max(total_ours, 1e-9)))
        f.write("\n## per-launch list (library kernels, in launch order)\n\n| # | kernel | grid | block | us |\n|---|---|---|---|---|\n")
        i = 0
        for n, us, g, b in rows:
            if n.startswith(OURS):
                f.write("| %d | `%s` | %s | %s | %.1f |\n" % (i, n, g, b, us))
                i += 1
    print("wrote profiles/launches_%s.md" % tag)


def full(tag):
    rep = os.path.join(OUT, "prof_%s.ncu-rep" % tag)
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    traffic_path = os.path.join(PROF, "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    # the capture holds the kernels of ONE cfg2 step in launch order: name them by layer for bench.py.  A layer is identified
    # by its position AND the kernel it must be (conv6 runs as two launches over the halves of its input channels: their
    # traffic is summed); if the sequence does not match, no per-layer attribution is written rather than a wrong one.
    step_layers = [("k1_cost_volume", "k1m_cost_volume_kernel"), ("conv0", "k2_conv3d_zm_kernel<0, 8>"), ("conv1", "k2_conv3d_zm_kernel<1, 16>"),
                   ("conv2", "k2_conv3d_zm_kernel<0, 16>"), ("conv3", "k2_conv3d_zm_kernel<1, 16>"), ("conv4", "k2_conv3d_zm_kernel<0, 16>"),
                   ("conv5", "k2_conv3d_zm_kernel<1, 16>"), ("conv6", "k2_conv3d_zm_kernel<0, 16>"), ("conv6", "k2_conv3d_zm_kernel<0, 16>"),
                   ("conv7", "k2_conv3d_zm_kernel<2, 16>"), ("conv9", "k2_conv3d_zm_kernel<2, 16>"), ("conv11", "k2_conv3d_zm_kernel<2, 8>"),
                   ("prob", "k2_conv3d_c1_kernel"), ("k3_regress", "k3_depth_regress")]
    mismatch = False
    layers, nth = {}, 0
    with open(os.path.join(PROF, "ncu_%s.md" % tag), "w") as f:
        f.write("# ncu --set full `%s` (--clock-control none, one pass of cfg2 under the profiler; not a bench value)\n\n" % tag)
        for r in rows[2:]:
            name = short(r[hdr.index("Kernel Name")])
            f.write("## `%s`  (launch id %s)\n\n| metric | value |\n|---|---|\n" % (name, r[hdr.index("ID")]))
            rd = wr = None
            for m, label in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write("| %s (`%s`) | %s %s |\n" % (label, m, r[i], units[i]))
                    if m.startswith("dram__bytes"):
                        v = float(r[i].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
                        if "read" in m:
                            rd = v
                        else:
                            wr = v
            if rd is not None and wr is not None:
                f.write("| **dram traffic per launch** | %.1f MB |\n" % ((rd + wr) / 1e6))
                key = name + "#" + r[hdr.index("ID")]
                traffic[key] = rd + wr
                if nth < len(step_layers):
                    lname, kprefix = step_layers[nth]
                    if not name.startswith(kprefix):
                        mismatch = True
                    elif lname in layers:
                        layers[lname]["dram_bytes"] += rd + wr
                        layers[lname]["launches"] += 1
                    else:
                        layers[lname] = {"kernel": name, "dram_bytes": rd + wr, "launches": 1, "tag": tag,
                                         "grid": r[hdr.index("launch__grid_size")] if "launch__grid_size" in hdr else None}
            nth += 1
            f.write("\n")
    if mismatch:
        print("launch sequence does not match one cfg2 step: traffic.json `layers` left unchanged")
    elif len(layers) >= 2:
        traffic["layers"] = layers
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
    print("wrote profiles/ncu_%s.md, profiles/traffic.json" % tag)


if __name__ == "__main__":
    tag = sys.argv[1]
    launches(tag)
    full(tag)
