#!/usr/bin/env python3
"""Write tests/golden/frame_md5.json: MD5 of the frames the reference-driven
oracle (oracle/_ref, i.e. the reference's own C templates) reconstructs for the
fixed synthetic cases of tests/test_frame.py.  Run in the container that has
/root/reference (after `make -C oracle ref`)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refdsp  # noqa: E402
import test_frame as T  # noqa: E402

ref = refdsp.RefDSP()
out = {}
for name, (w, h, bd, seed, kw) in T.CASES.items():
    hf = T.F.HostFrame(w, h, bd, seed, **kw)
    _, _, planes = T.oracle_planes(ref, hf, seed)
    out[name] = T.md5_planes(planes)
    print(name, out[name], "blocks", hf.n_blocks, "intra", hf.n_intra_blocks, "intra ops", hf.n_intra)
with open(T.GOLDEN, "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)

# frames reconstructed by the reference's own driver (dav1d_recon_b_intra, oracle/ref_recon.c)
import refframe  # noqa: E402
import test_reference_driver as R  # noqa: E402

out = {}
for name in R.CASES:
    hf, init = R.make(name)
    out[name] = R.md5_planes(refframe.run_reference_driver(ref, hf, [p.copy() for p in init], R.refs_of(hf, name)))
    print(name, out[name], "blocks", hf.n_block_recs, "intra ops", hf.n_intra)
with open(R.GOLDEN, "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)

# the reference's loop filter over the block records (tests/test_loopfilter.py)
import reflf  # noqa: E402
import test_loopfilter as LF  # noqa: E402

out = {}
for name in LF.CASES:
    hf, src, seed, sharp = LF.make(name)
    planes, _ = reflf.run_reference_lf(ref, hf, [p.copy() for p in src], seed, sharpness=sharp)
    out[name] = LF.md5_planes(planes)
    print(name, out[name])
with open(LF.GOLDEN, "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)

# the reference's CDEF (tests/test_cdef.py)
import test_cdef as CD  # noqa: E402

out = {}
for name in CD.CASES:
    hf, src, seed, damping, ys, us = CD.make(name)
    planes, _ = reflf.run_reference_cdef(ref, hf, [p.copy() for p in src], seed, damping, ys, us)
    out[name] = CD.md5_planes(planes)
    print(name, out[name])
with open(CD.GOLDEN, "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)

# the reference's post-filter chain (tests/test_postfilter_chain.py)
import test_postfilter_chain as PF  # noqa: E402

out = {}
for name in PF.CASES:
    hf, src, seed, par = PF.make(name)
    planes, _ = reflf.run_reference_chain(ref, hf, [p.copy() for p in src], seed, **par)
    out[name] = PF.md5_planes(planes)
    print(name, out[name])
with open(PF.GOLDEN, "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
