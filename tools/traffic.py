#!/usr/bin/env python
"""DRAM traffic per stage from an ncu launch list: writes profiles/roofline_traffic.json (read by bench.py for `roofline.traffic`).

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_" -c 400 --csv \\
      --log-file gpurun_out/traffic.csv python tools/stage_times.py --variant fov --first 5 --frames 18
  python tools/traffic.py gpurun_out/traffic.csv 5 18

The capture covers bench.py's own frames (frame f = camera f%30, gaze f%9; the default bench run takes its statistics from frames
5..22), the first two rendered frames are warm-up and are dropped; values are means per launch over the remaining frames."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE = {"k_pre": "preprocess", "k_tile_scan": "color", "k_color": "color", "k_scatter": "scatter", "k_lazy_blend": "blend", "k_setup": "setup",
         "k_tile_levels": "setup", "k_tile_infos": "setup"}


def main():
    path, first, frames = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rows = list(csv.reader(open(path)))
    hdr = None
    per = collections.defaultdict(lambda: collections.defaultdict(list))      # kernel -> metric -> values in launch order
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            rec = dict(zip(hdr, r))
            name = rec["Kernel Name"].replace("void ", "").replace("fovgs::", "").split("(")[0]
            v = float(rec["Metric Value"].replace(",", ""))
            unit = rec.get("Metric Unit", "")
            scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "ns": 1e-6}.get(unit, 1.0)
            per[name][rec["Metric Name"]].append(v * scale)
    out = collections.defaultdict(float)
    detail = {}
    for name, m in per.items():
        stage = next((s for k, s in STAGE.items() if name.startswith(k)), None)
        if stage is None:
            continue
        rd, wr, t = m.get("dram__bytes_read.sum", []), m.get("dram__bytes_write.sum", []), m.get("gpu__time_duration.sum", [])
        n = len(rd)
        skip = n - frames if n >= frames else 0                            # leading warm-up launches
        mean = lambda x: sum(x[skip:]) / max(len(x[skip:]), 1)
        detail[name] = {"launches": n - skip, "dram_read": mean(rd), "dram_write": mean(wr), "ncu_ms": mean(t)}
        out[stage] += mean(rd) + mean(wr)
    res = {"note": f"mean DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) over bench.py frames {first}..{first + frames - 1} "
                   "(6 M Gaussians, 1080p, foveated, moving gaze) from one ncu pass; blend = both lazy-blend launches of a frame; "
                   "tools/traffic.py", "kernels": detail}
    res.update({k: int(v) for k, v in out.items()})
    json.dump(res, open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "kernels"}))


if __name__ == "__main__":
    main()
