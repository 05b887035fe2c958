#!/usr/bin/env python3
"""Randomised parity soak: N synthetic frames with random layouts / bit depths / sizes / block mixes
through the default batched path (alternating single-frame graph and group graph) against the
reference-driven oracle.  usage: tools/soak.py [N] [first_seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _d1pkg  # noqa: E402

pkg = _d1pkg.load_pkg()
from dav1d_mirror_b200 import frame as F  # noqa: E402
import refdsp  # noqa: E402
import refframe  # noqa: E402

_ref = None


def run_one(seed):
    """One random frame: returns (bit_exact, description)."""
    global _ref
    if _ref is None:
        _ref = refdsp.RefDSP()
    ref, L = _ref, pkg.lib()
    r = np.random.default_rng(seed)
    lay = [(1, 1), (1, 0), (0, 0)][r.integers(3)]
    bd = [0xff, 0x3ff, 0xfff][r.integers(3)]
    w = int(r.integers(8, 80)) * 8
    h = int(r.integers(8, 60)) * 8
    kw = dict(ss_hor=lay[0], ss_ver=lay[1], p_intra=float(r.choice([0.0, 0.2, 0.5, 1.0])),
              p_obmc=float(r.choice([0.0, 0.3])), p_ii=float(r.choice([0.0, 0.3])),
              p_ibc=float(r.choice([0.0, 0.3])), p_palette=float(r.choice([0.0, 0.1])),
              p_cfl=float(r.choice([0.0, 0.5])), p_filter_intra=float(r.choice([0.0, 0.2])),
              p_warp=float(r.choice([0.0, 0.1])), mv_range=int(r.choice([16, 128, 400])),
              dense_coefs=int(r.integers(2)), p_tx_split=float(r.choice([0.0, 0.5])))
    if r.integers(8) == 0:
        kw["no_chroma"] = 1
    hf = F.HostFrame(w, h, bd, seed, **kw)
    if seed & 4:
        hf.record_levels()       # the recorder's level pass instead of the device's
    refs = [F.random_planes(hf, seed * 10 + k) for k in range(2)]
    init = F.random_planes(hf, seed * 10 + 5)
    want = refframe.run_oracle(ref, hf, [p.copy() for p in init], refs)
    ctx = F.open_context(0)
    df = F.DeviceFrame(ctx, hf, n_refs=2)
    df.upload_descriptors()
    for k, planes in enumerate(refs):
        df.upload_picture(df.refs[k], planes)
    df.upload_picture(df.dst, init)
    if seed & 1:
        mf = F.MultiFrame(ctx, [df], graph=bool(seed & 2))
        mf.launch()
    else:
        mf = None
        df.build_graph()
        df.launch_graph()
    got = df.download_picture()
    ok = all(np.array_equal(a, b) for a, b in zip(want, got))
    pkg.check_error()
    if mf:
        mf.close()
    df.close()
    L.dav1d_cuda_close(ctx)
    return ok, f"{seed} {w}x{h} {hex(bd)} {kw} intra ops {hf.n_intra}"


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
    bad = 0
    for seed in range(first, first + n):
        ok, what = run_one(seed)
        bad += not ok
        print(("ok  " if ok else "FAIL"), what, flush=True)
    print("soak:", n - bad, "of", n, "frames bit-exact")
    sys.exit(1 if bad else 0)
