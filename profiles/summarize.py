#!/usr/bin/env python
"""Turn gpurun_out/prof_<tag>.ncu-rep (+ launches_<tag>.csv) into profiles/<tag>_ncu.md -- the per-kernel summary that is
committed (the .ncu-rep files themselves stay in the git-ignored gpurun_out/).

    python profiles/summarize.py <tag> [bench_json]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1/smem %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2 %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
    ("launch__registers_per_thread", "regs"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank conflicts"),
    ("smsp__inst_executed.sum", "warp inst"),
]


def to_bytes(val, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(val) * mult.get(unit, 1)


def main():
    tag = sys.argv[1]
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu summary `{tag}`", "",
             "`ncu --set full --clock-control none --import-source on` on one bench step (cold-cache, serialised launches: "
             "compare shares, not absolutes).  Source: `gpurun_out/prof_%s.ncu-rep` (not committed); recipe `profiles/run_ncu.sh`." % tag, ""]
    lines.append("| kernel | grid x block | " + " | ".join(n for _, n in METRICS) + " | dram traffic / launch |")
    lines.append("|---|---|" + "---|" * (len(METRICS) + 1))
    seen = set()
    traffic = {}
    warp_inst = {}
    per_step = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("maf::", "")
        per_step[name.split("<")[0]] = per_step.get(name.split("<")[0], 0) + 1
        if name in seen:
            continue
        seen.add(name)
        try:
            warp_inst[name.split("<")[0]] = float(r[hdr.index("smsp__inst_executed.sum")])
        except ValueError:
            pass
        cells = []
        for m, _ in METRICS:
            i = hdr.index(m)
            v = r[i]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            u = units[i]
            cells.append(f"{v} {u}".strip() if u not in ("%", "inst", "register/thread", "") else v)
        rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
        wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
        traffic[name.split("<")[0]] = rd + wr
        lines.append(f"| `{name}` | {r[hdr.index('launch__grid_size')]} x {r[hdr.index('launch__block_size')]} | " + " | ".join(cells) +
                     f" | {(rd + wr) / 1e6:.1f} MB |")
    wl = "cfg2"
    for a in sys.argv[3:]:
        wl = a
    with open(os.path.join(ROOT, "profiles", f"traffic_{wl}.json"), "w") as f:
        json.dump({"tag": tag, "workload": wl, "source": "ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum per launch",
                   "bytes_per_launch": traffic, "warp_inst_per_launch": warp_inst, "launches_per_step": per_step}, f, indent=1)
    lpath = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    if os.path.exists(lpath):
        lr = list(csv.DictReader(l for l in open(lpath) if l.startswith('"')))
        tot = {}
        order = []
        for r in lr:
            n = r["Kernel Name"].split("(")[0].replace("void ", "").replace("maf::", "")
            if n not in tot:
                order.append(n)
            tot.setdefault(n, []).append(int(r["Metric Value"]) / 1000.0)
        total = sum(sum(v) / len(v) for v in tot.values())
        lines += ["", "## launch list (`--metrics gpu__time_duration.sum`, %d launches after warm-up)" % len(lr), "",
                  "| kernel | launches | mean us | share of step |", "|---|---|---|---|"]
        for n in order:
            v = tot[n]
            lines.append(f"| `{n}` | {len(v)} | {sum(v) / len(v):.1f} | {100 * sum(v) / len(v) / total:.1f} % |")
    if len(sys.argv) > 2 and os.path.exists(sys.argv[2]):
        b = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
        st = b["roofline"]["stages"]
        tot = sum(v["ms"] for v in st.values())
        lines += ["", "## CUDA-event stage times of the un-profiled bench run (`%s`)" % os.path.basename(sys.argv[2]), "",
                  "value %.0f %s, %.3f ms/step" % (b["value"], b["unit"], b["ms_per_step"]), "",
                  "| stage | ms | share | algorithmic GB/s |", "|---|---|---|---|"]
        for k, v in st.items():
            lines.append(f"| {k} | {v['ms']:.4f} | {100 * v['ms'] / tot:.1f} % | {v['GBps']} |")
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
