#!/usr/bin/env python
"""profiles/: one NVSmall step of an ncu launch list (gpu__time_duration + dram bytes per launch) as a table + the DRAM traffic
per pair that bench.py reports in its `traffic` fields.

  python tools/step_summary.py gpurun_out/r02_launches_ncu.csv 23 profiles/r02_ncu_step_summary.md profiles/r02_traffic.json
"""
import csv
import json
import sys

STEP_NAMES = ["left conv1 (3->32 5x5 s2, CUDA cores)", "right conv1", "towers conv2 pack (L|R as one batch of 2)", "towers conv2", "towers conv3",
              "towers conv4", "towers conv5", "cost_vol+conv3D_1: pack L", "cost_vol+conv3D_1: conv2d 32->96 (L)", "cost_vol+conv3D_1: pack R",
              "cost_vol+conv3D_1: conv2d 32->96 (R)", "cost_vol+conv3D_1: edge", "cost_vol+conv3D_1: combine (writes conv3D_1 output)",
              "conv3D_2 (depth-stationary kernel)", "conv3D_3ds", "conv3D_4", "conv3D_5", "conv3D_6ds", "conv3D_7", "conv3D_8",
              "deconv3D_1 (+skip, ELU)", "deconv3D_2 (+skip, ELU)", "deconv3D_3 + slice + soft-argmin (one kernel)"]


def main():
    path, per_step, out_md, out_json = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
    rows = [r for r in csv.DictReader(l for l in open(path) if not l.startswith("=="))]
    launches = {}
    for r in rows:
        d = launches.setdefault(int(r["ID"]), {"kernel": r["Kernel Name"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    ids = [i for i in sorted(launches) if "rt::" in launches[i]["kernel"]][-per_step:]      # (torch's own kernels, e.g. the final .mean(), are not part of a step)
    total = sum(launches[i]["gpu__time_duration.sum"] for i in ids)
    lines = ["# ncu launch list of one NVSmall 1025x321 step (round 2)\n",
             "Command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv python tools/onestep.py 3`",
             "(raw: `profiles/r02_launches_ncu.csv`; last step shown, %d launches).  Times under ncu are serialised and cold-cache: the *shares* are what" % per_step,
             "agrees with `bench.py`'s CUDA-event numbers (`layer_ms`), not the absolutes.\n",
             "| # | engine step | kernel | time [us] | share | DRAM read [MB] | DRAM write [MB] |", "|---|---|---|---|---|---|---|"]
    conv_stack = cv1 = 0.0
    for n, i in enumerate(ids):
        L = launches[i]
        k = L["kernel"].replace("void ", "").replace("rt::<unnamed>::", "").split("(")[0]
        t, rd, wr = L["gpu__time_duration.sum"] / 1e3, L["dram__bytes_read.sum"], L["dram__bytes_write.sum"]
        name = STEP_NAMES[n] if per_step == len(STEP_NAMES) else ""
        lines.append("| %d | %s | `%s` | %.1f | %.1f %% | %.1f | %.1f |" % (n, name, k, t, 100 * t * 1e3 / total, rd / 1e6, wr / 1e6))
        if name.startswith(("conv3D_", "deconv3D_")):
            conv_stack += rd + wr
        if name.startswith("cost_vol+conv3D_1"):
            cv1 += rd + wr
    lines.append("| | **total** | | %.1f | | | |" % (total / 1e3))
    open(out_md, "w").write("\n".join(lines) + "\n")
    json.dump({"conv3d_stack_bytes_per_pair": conv_stack, "costvol_conv1_bytes": cv1,
               "source": "%s (ncu dram__bytes_read/write.sum per launch, one NVSmall step, batch 1)" % out_md}, open(out_json, "w"), indent=1)
    print("total %.1f us, conv stack %.1f MB, cost_vol+conv3D_1 %.1f MB" % (total / 1e3, conv_stack / 1e6, cv1 / 1e6))


if __name__ == "__main__":
    main()
