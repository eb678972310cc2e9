#!/usr/bin/env python
"""Dev aid: turn an `ncu --set full` report of the bulk kernels into profiles/ncu_traffic.json, the per-build record bench.py
reads for `roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum per launch) instead of a literal.

    python tools/ncu_traffic.py gpurun_out/r02_prof_bulk2.ncu-rep --frames 17 --width 6000 --height 4000 --channels 3 --masters 3
"""
import argparse
import csv
import json
import os
import subprocess

NAMES = {"k_calibrate_gray_batch": "calibrate_gray", "k_calibrate_gray_ring": "calibrate_gray", "k_warp_accumulate_tma": "warp_accumulate", "k_warp_accumulate_tiled": "warp_accumulate",
         "k_sky_sub_stats": "sky_sub_stats", "k_master_accumulate": "master_accumulate"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3,
        "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--frames", type=int, default=17)
    ap.add_argument("--width", type=int, default=6000)
    ap.add_argument("--height", type=int, default=4000)
    ap.add_argument("--channels", type=int, default=3)
    ap.add_argument("--masters", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json"))
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        return float(r[col[name]].replace(",", "")) * UNIT.get(units[col[name]], 1.0)
    kernels = {}
    for r in rows[2:]:
        fn = r[col["Kernel Name"]].split("(")[0].split("<")[0].split("::")[-1].replace("void ", "").strip()
        name = NAMES.get(fn)
        if not name:
            continue
        rd, wr, ms = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum"), val(r, "gpu__time_duration.sum")
        k = {"kernel": r[col["Kernel Name"]].split("(")[0], "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
             "frames_per_launch": a.frames, "duration_ms_under_ncu": ms,
             "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
             "registers": val(r, "launch__registers_per_thread"), "warp_instructions": val(r, "smsp__inst_executed.sum")}
        if name not in kernels or k["dram_bytes_per_launch"] > kernels[name]["dram_bytes_per_launch"]:
            kernels[name] = k   # several launches of a kernel: keep the full-batch one
    out = {"source": "ncu --set full --clock-control none, report %s, written by tools/ncu_traffic.py" % os.path.basename(a.report),
           "shape": {"width": a.width, "height": a.height, "channels": a.channels, "masters": a.masters, "frames_per_launch": a.frames},
           "kernels": kernels}
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    for n, k in kernels.items():
        print(n, "%.3f GB per launch (%.1f B/px-frame)" % (k["dram_bytes_per_launch"] / 1e9,
                                                            k["dram_bytes_per_launch"] / a.frames / (a.width * a.height)))


if __name__ == "__main__":
    main()
