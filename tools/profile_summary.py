#!/usr/bin/env python
"""Turns an .ncu-rep (brought back in gpurun_out/) into the small text summary committed under profiles/.

  python tools/profile_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<name>.md
"""
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor-memory active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    print("# ncu summary of `%s`\n" % rep)
    print("Captured with `ncu --set full --clock-control none --import-source on` under gpurun (one B200); per-launch values,"
          " cold-cache and serialised -- compare shares and utilisation, not absolute times.\n")
    cols = [(hdr.index(m), label, units[hdr.index(m)]) for m, label in METRICS if m in hdr]
    print("| # | kernel | " + " | ".join("%s [%s]" % (l, u) if u else l for _, l, u in cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for i, r in enumerate(data):
        name = r[kn].replace("void ", "").replace("unnamed>::", "").replace("rt::<", "")
        name = name.split("(")[0][:48]
        vals = []
        for c, _, _ in cols:
            v = r[c]
            try:
                v = "%.4g" % float(v)
            except ValueError:
                pass
            vals.append(v)
        print("| %d | `%s` | " % (i, name) + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
