#!/usr/bin/env python
"""ncu launch list (csv, `--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`) -> profiles/ncu_traffic.json.

  python tools/ncu_traffic.py profiles/r02_ncu_launches.csv "<the ncu command>"

One bench step = the launches from one `expand_bpp_kernel` (first kernel of cic_adaptive_forward) up to the next one; the second
complete step of the capture is used.  Per kernel class: launches, ncu time (cold cache, serialised: shares, not absolutes) and
DRAM bytes (read + write) per step - `bench.py` reads `dram_bytes_per_step` for `roofline.traffic`."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"<.*$", "", name)
    return name.replace("cic::", "")


def main():
    path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = {}
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        d = rows.setdefault(int(r["ID"]), {"name": short(r["Kernel Name"])})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        else:
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            d[r["Metric Name"]] = v * mult
    ids = sorted(rows)
    starts = [i for i in ids if rows[i]["name"] == "expand_bpp_kernel"]
    if len(starts) < 3:
        sys.exit(f"need at least three forward calls in the capture, found {len(starts)}")
    lo, hi = starts[1], starts[2]
    step = [rows[i] for i in ids if lo <= i < hi]
    total_us = sum(d.get("us", 0.0) for d in step)
    out = {"_source": f"{cmd} ({os.path.basename(path)}, launches {lo}..{hi - 1} = one bench step of 256 tiles of 256x256 incl. the evaluation kernels: "
                      f"{len(step)} launches); dram bytes = read + write summed over the launches of the class in that step; ncu times are "
                      f"cold-cache and serialised (shares, not absolutes)"}
    for d in step:
        c = out.setdefault(d["name"], {"dram_bytes_per_step": 0.0, "launches_per_step": 0, "ncu_time_us_per_step": 0.0})
        c["dram_bytes_per_step"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        c["launches_per_step"] += 1
        c["ncu_time_us_per_step"] += d.get("us", 0.0)
    for k, c in out.items():
        if k.startswith("_"):
            continue
        c["dram_bytes_per_launch"] = c["dram_bytes_per_step"] / c["launches_per_step"]
        c["share_of_step_ncu"] = round(c["ncu_time_us_per_step"] / total_us, 4) if total_us else None
        c["ncu_time_us_per_step"] = round(c["ncu_time_us_per_step"], 1)
    json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
    print(f"step = launches {lo}..{hi - 1} ({len(step)} launches, {total_us / 1e3:.2f} ms under ncu)")
    for k, c in sorted(((k, c) for k, c in out.items() if not k.startswith("_")), key=lambda kc: -kc[1]["ncu_time_us_per_step"]):
        print(f"  {k:32s} {c['launches_per_step']:3d} launches {c['ncu_time_us_per_step']:9.1f} us {c['share_of_step_ncu']:.3f}  {c['dram_bytes_per_step'] / 1e6:10.1f} MB")


if __name__ == "__main__":
    main()
