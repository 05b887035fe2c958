#!/usr/bin/env python3
"""profiles/r2_traffic.json from an ncu launch list of the benched step (gpu__time_duration.sum,
dram__bytes_read.sum, dram__bytes_write.sum per launch): DRAM bytes per frame and launch class.
usage: traffic_from_launches.py launches.csv frames_per_group > profiles/r2_traffic.json"""
import csv
import json
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
group = int(sys.argv[2]) if len(sys.argv) > 2 else 64
hdr = rows[0]
ik, im, iu, iv = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Unit'), hdr.index('Metric Value')
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
CLASS = (("mc_put", "mc_put"), ("mc_obmc", "mc_put"), ("mc_compound", "mc_compound"), ("warp_batch", "warp"),
         ("itx2_task", "itx"), ("intra_", "intra"))
# launches of a class per frame (one frame's submission: recon2.cu group_submit_on)
PER_FRAME = {"mc_put": 2, "mc_compound": 4, "warp": 1, "itx": 6}
out = {}
for r in rows[1:]:
    cls = next((c for k, c in CLASS if k in r[ik]), None)
    if cls is None or not r[im].startswith("dram__bytes"):
        continue
    o = out.setdefault(cls, {"bytes": 0.0, "launches": 0})
    o["bytes"] += float(r[iv].replace(',', '')) * SCALE.get(r[iu], 1)
    if r[im] == "dram__bytes_read.sum":
        o["launches"] += 1
res = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                 "--launch-skip 1944 -c 860 on `python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-verify` "
                 f"(one group submission of {group} frames per step); per-launch values are cold-cache and serialised",
       "frames_in_window": {}}
for cls, o in out.items():
    frames = group if cls == "intra" else o["launches"] / PER_FRAME[cls]
    res["frames_in_window"][cls] = round(frames, 1)
    res[cls] = {"dram_bytes_per_frame": o["bytes"] / frames, "launches": o["launches"]}
res["itx"]["note"] = "inter residuals and the intra residual pre-pass (three size classes each)"
res["intra"]["note"] = "executor + mark / keys / sort / scan / clear of the group; the residual pre-pass is counted under itx"
json.dump(res, sys.stdout, indent=1)
print()
