#!/usr/bin/env python3
"""Where does the CUDA frame differ from the reference driver?  Runs one case of tests/test_reference_driver.py on
the GPU (device-side and recorded levels) and prints the differing 4x4 cells per plane with the intra-class
operation that covers the first one.  usage: tools/diff_frame_vs_driver.py <case name>"""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import _d1pkg
pkg = _d1pkg.load_pkg()
from dav1d_mirror_b200 import frame as F
import refdsp, refframe, test_frame, test_reference_driver as R, test_recorder as TR
ref = refdsp.RefDSP()
name = sys.argv[1]
hf, init = R.make(name)
refs = R.refs_of(hf, name)
want = refframe.run_reference_driver(ref, hf, [p.copy() for p in init], refs)
for rec in (False, True):
    if rec:
        hf.record_levels()
    got = test_frame.run_gpu(hf, refs, init, use_graph=False)
    ops = np.frombuffer(hf.intra.tobytes(), dtype=TR.OP)
    for pl, (a, b) in enumerate(zip(want, got)):
        bad = np.argwhere(a != b)
        if not len(bad):
            continue
        cells = sorted(set((int(y) // 4, int(x) // 4) for y, x in bad))
        print("recorded", rec, "plane", pl, "bad px", len(bad), "cells", len(cells), cells[:12])
        y4, x4 = cells[0]
        for i, o in enumerate(ops):
            if o["plane"] == pl and o["x4"] <= x4 < o["x4"] + o["tw4"] and o["y4"] <= y4 < o["y4"] + o["th4"]:
                print("  op", i, {k: (o[k].tolist() if hasattr(o[k], 'tolist') else o[k]) for k in ("x4", "y4", "tw4", "th4", "mode", "angle_delta", "flags", "aux", "eob", "tx", "tile_x4_start", "tile_y4_start")},
                      "src", (o["aux"] & 0xffff, o["aux"] >> 16) if o["mode"] == 17 else None)
