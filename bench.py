#!/usr/bin/env python3
"""Headline benchmark: Mpix/s of mc + itx + ipred reconstruction per B200.

Workload (BASELINE.json configs[3]/[4]): full synthetic 4K 10-bit 4:2:0 frame
reconstruction - motion compensation (put / fused compound / warp), inter
residual inverse transforms and intra prediction (+CfL, palette,
filter-intra) - for `--streams` independent streams per GPU, each with its
own reference frames in HBM.  A step = one frame of every stream; the frames of
`--group` streams are submitted together (dav1d_cuda_recon_group_submit).

  python bench.py --gpus N --steps K --warmup W            (our CUDA path)
  python bench.py --impl reference ...                     (reference C templates on host cores)

`value`  : luma Mpix/s with descriptors/coefficients/refs resident in HBM.
`e2e`    : same metric through the C ABI with HOST buffers - every step takes
           the NEXT frame of every stream (a different descriptor set than the
           step before), ships descriptors + coefficients H2D from pinned
           memory, submits and reads the reconstructed frame back D2H, all
           inside the timed region.  The library does no host-side scheduling.
Timing: CUDA events on the launching streams (fork/join through events), max
over ranks.  No collectives on the data path (streams are independent).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L2_MB = 126.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=64, help="independent 4K streams per GPU")
    ap.add_argument("--group", type=int, default=64,
                    help="streams whose frames are submitted together (dav1d_cuda_recon_group_submit: one intra "
                         "executor launch per group) in the device-resident arm")
    ap.add_argument("--e2e-group", type=int, default=16,
                    help="same for the end-to-end arm: smaller groups on their own CUDA streams, so that the "
                         "copies of one group overlap the kernels of the others")
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--bitdepth-max", type=lambda s: int(s, 0), default=0x3ff)
    ap.add_argument("--coefs", default="int16", choices=["int16", "native"],
                    help="coefficient stream of high-bit-depth frames: int16 + escape list (cf_int16) or the "
                         "reference's int32 layout")
    ap.add_argument("--mc-staging", default="tma_put", choices=["tma", "tma_put", "cp_async"],
                    help="how the 32x32-tile MC kernels stage the reference windows (dav1d_cuda_set_mc_tma)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--graph", action="store_true", help="replay the group submission from a captured CUDA graph "
                                                         "(device-resident arm only; e2e always submits directly)")
    a = ap.parse_args()
    a.group = max(1, min(a.group, a.streams, 64))
    a.e2e_group = max(1, min(a.e2e_group, a.streams, 64))
    return a


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.marks = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        """host time stamp: samples between the first and the last mark are 'during the timed region'"""
        self.marks.append(time.time())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def digest(lines):
            sm, mx, reasons = [], [], set()
            for ln in lines:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for n, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            return sm, mx, reasons
        lo, hi = (self.marks[0], self.marks[-1]) if len(self.marks) >= 2 else (0.0, float("inf"))
        inside = [ln for t, ln in self.lines if lo <= t <= hi + 0.02]
        window = "timed region"
        sm, mx, reasons = digest(inside)
        if not sm:       # region shorter than the sampling period: use the warm-up + timed window
            sm, mx, reasons = digest([ln for _, ln in self.lines])
            window = "warm-up + timed region"
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------ reference arm / cpu baseline
def cpu_reference_run(args, steps, warmup, frames_per_step=None):
    """Times the reference's own C templates (oracle/_ref, compiled from /root/reference) replaying
    the same descriptors sequentially, one frame per host thread."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _d1pkg
    _d1pkg.load_pkg()
    from dav1d_mirror_b200 import frame as F
    import refdsp
    import refframe
    ref = refdsp.RefDSP()
    cores = os.cpu_count() or 1
    n = frames_per_step or cores
    # a bounded sample: n frames (distinct seeds cycle over 4 descriptor sets)
    hfs = [F.HostFrame(args.width, args.height, args.bitdepth_max, 1000 + i) for i in range(min(n, 4))]
    keep, ofs = [], (refframe.OracleFrame * n)()
    refs_by_set = [[F.random_planes(hf, 7 + r) for r in range(2)] for hf in hfs]
    for i in range(n):
        hf = hfs[i % len(hfs)]
        dst = F.random_planes(hf, 99 + i)
        keep.append(dst)
        ofs[i] = refframe.make_oracle_frame(hf, dst, refs_by_set[i % len(hfs)], keep)
    fn = ref.lib.oracle_ref_frames_run_mt
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int]
    fn.restype = None
    for _ in range(warmup):
        fn(ofs, n, cores)
    t0 = time.perf_counter()
    for _ in range(steps):
        fn(ofs, n, cores)
    dt = time.perf_counter() - t0
    luma = hfs[0].luma_px
    mpix = n * steps * luma / dt / 1e6
    return {"value": mpix, "unit": "Mpix/s", "cores": min(cores, n), "kind": "reference",
            "sample": f"{n} frames/step x {steps} steps of {args.width}x{args.height} "
                      f"{'10' if args.bitdepth_max == 0x3ff else '12' if args.bitdepth_max > 0x3ff else '8'}-bit 4:2:0, "
                      f"reference C templates (gcc -O3, no asm: no nasm in the image), one frame per thread",
            "seconds": dt, "ms_per_step": dt / steps * 1e3, "frames_per_step": n}


def main_reference(args):
    rank, world, local = dist_env()
    if rank != 0:
        return
    # bounded: a step = one frame per host thread; the requested steps / warm-up are capped so that
    # the run ends within a few minutes, and the line reports what was actually run
    steps, warmup = max(1, min(args.steps, 8)), max(0, min(args.warmup, 1))
    r = cpu_reference_run(args, steps, warmup)
    out = {"metric": "Mpix/s of mc+itx+ipred recon", "impl": "reference", "value": r["value"], "unit": "Mpix/s",
           "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"],
           "requested": {"steps": args.steps, "warmup": args.warmup},
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16" if args.bitdepth_max > 0xff else "u8",
           "data": "synthetic",
           "config": workload_config(args),
           "config_extra": {"frames_per_step": r["frames_per_step"], "host_threads": r["cores"]},
           "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
           "e2e": {"value": r["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def workload_config(args):
    """The same dictionary for both arms (what is measured); arm-specific details go to config_extra."""
    bits = 10 if args.bitdepth_max == 0x3ff else 12 if args.bitdepth_max > 0x3ff else 8
    return {"workload": f"full synthetic {args.width}x{args.height} {bits}-bit 4:2:0 reconstruction "
                        f"(MC put/compound/warp + itx + intra/CfL/palette), independent streams, one frame per "
                        f"stream and step (BASELINE configs[3]/[4])",
            "frame": [args.width, args.height], "bitdepth": bits,
            "mix": "30% intra blocks, 60% blocks with residual, inter: put 50/avg 20/w_avg 10/wedge 10/seg 5/warp 5",
            "parallelism": "independent streams, no collective"}


# ------------------------------------------------------------------ our arm
def pcie_probe(device, h2d_bytes, d2h_bytes, step_ms):
    """Both copy directions at once, pinned host memory, bytes in the proportion of one e2e step of this rank
    (scaled down to at most 512 MB per direction): the time a step's transfers take on this box when nothing
    else happens."""
    import torch
    try:
        scale = min(1.0, 512e6 / max(h2d_bytes, d2h_bytes, 1))
        nh, nd = max(int(h2d_bytes * scale), 1 << 20), max(int(d2h_bytes * scale), 1 << 20)
        dev = torch.device("cuda", device)
        hs, hd = torch.empty(nh, dtype=torch.uint8, pin_memory=True), torch.empty(nd, dtype=torch.uint8, pin_memory=True)
        ds, dd = torch.empty(nh, dtype=torch.uint8, device=dev), torch.empty(nd, dtype=torch.uint8, device=dev)
        s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        best = None
        for it in range(4):
            torch.cuda.synchronize(dev)
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(torch.cuda.current_stream(dev))
            s1.wait_event(e0)
            s2.wait_event(e0)
            with torch.cuda.stream(s1):
                ds.copy_(hs, non_blocking=True)
                e1.record(s1)
            with torch.cuda.stream(s2):
                hd.copy_(dd, non_blocking=True)
                e2.record(s2)
            torch.cuda.synchronize(dev)
            ms = max(e0.elapsed_time(e1), e0.elapsed_time(e2))
            if it and (best is None or ms < best[0]):
                best = (ms, e0.elapsed_time(e1), e0.elapsed_time(e2))
        ms, ms_h2d, ms_d2h = best
        floor_ms = ms / scale
        return {"h2d_gbs": nh / ms_h2d / 1e6, "d2h_gbs": nd / ms_d2h / 1e6, "both_directions": True,
                "step_transfer_floor_ms": floor_ms, "frac_of_floor": floor_ms / step_ms,
                "note": "frac_of_floor = time plain pinned copies of one step's bytes take on this box (both "
                        "directions concurrently) / the measured e2e step"}
    except Exception as e:       # noqa: BLE001 - the probe must never cost the bench line
        return {"error": repr(e)}


def main_ours(args):
    rank, world, local = dist_env()
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    # host threads and the pinned staging buffers of a rank live on the CPUs next to its GPU
    # (NUMA placement matters once several GPUs copy at the same time)
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1 and 64 * i + b < ncpu]
        if cpus and len(cpus) < ncpu:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass
    if world > 1:
        # stdout carries only the JSON line: NCCL prints its version banner to fd 1 when the
        # communicator is created (first collective), so fd 1 points at stderr until that is done
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    import _d1pkg
    pkg = _d1pkg.load_pkg()
    from dav1d_mirror_b200 import frame as F
    L = pkg.lib()
    if not L.dav1d_cuda_available():
        raise RuntimeError("no CUDA device: this benchmark has no CPU fallback")
    L.dav1d_cuda_set_mc_tma({"cp_async": 0, "tma_put": 1, "tma": 2}[args.mc_staging])
    from dav1d_mirror_b200 import dist as D
    S = args.streams
    # weak scaling: world * S independent streams, stream i -> rank i mod world (replicas only)
    my_streams = D.streams_of_rank(world * S, world, rank)
    assert len(my_streams) == S
    # ---- S streams with their own pictures.  Every stream cycles through N_SETS different frames
    # (descriptor sets): a step reconstructs the NEXT frame of every stream, never the same one
    # twice in a row.  The last stream of a rank is 12-bit (BASELINE config 5's spot check; 10- and
    # 12-bit frames share the pixel type and may share a group).
    N_SETS = 4
    bd12 = 0xfff if args.bitdepth_max == 0x3ff else args.bitdepth_max
    sets = {bd: [F.HostFrame(args.width, args.height, bd, 1000 + 17 * rank + i) for i in range(N_SETS)]
            for bd in {args.bitdepth_max, bd12}}
    # The recorder's dependency-level pass (dav1d_cuda_intra_levels: one linear walk over the
    # intra-class descriptors in decode order, host only).  The device-resident arm replays
    # descriptors that carry these levels; the end-to-end arm leaves the levels to the device.
    if args.coefs == "int16":          # the compact stream: int16 storage + escape list (cf_int16), high bit depth only
        for hfs_bd in sets.values():
            for hf in hfs_bd:
                if hf.hbd:
                    hf.pack_coefs()
    t0 = time.perf_counter()
    n_rec = 0
    for hfs_bd in sets.values():
        for hf in hfs_bd:
            hf.record_levels()
            n_rec += 1
    recorder_levels_ms = (time.perf_counter() - t0) * 1e3 / max(n_rec, 1)
    G = args.group
    ctxs, dfs, units = [], [], []
    plane_cache = {}
    main_ctx = F.open_context(local)
    for g0 in range(0, S, G):
        ctx = F.open_context(local)
        ctxs.append(ctx)
        gdfs = []
        for s in range(g0, min(S, g0 + G)):
            bd = bd12 if s == S - 1 else args.bitdepth_max
            hfs = sets[bd][s % N_SETS:] + sets[bd][:s % N_SETS]        # the stream's frame sequence
            df = F.DeviceFrame(ctx, hfs[0], n_refs=2, more_sets=hfs[1:])
            df.upload_descriptors()
            # picture CONTENT is generated once per bit depth and role (host-side set-up time); every
            # stream owns its own reference / destination pictures in HBM
            if bd not in plane_cache:
                plane_cache[bd] = [F.random_planes(hfs[0], 7 + r) for r in range(2)] + [F.random_planes(hfs[0], 99)]
            for r in range(2):
                df.upload_picture(df.refs[r], plane_cache[bd][r])
            df.upload_picture(df.dst, plane_cache[bd][2])
            gdfs.append(df)
            dfs.append(df)
        L.dav1d_cuda_synchronize(ctx)
        units.append([ctx, None, gdfs])
    pkg.check_error()
    luma_px = dfs[0].hf.luma_px
    footprint_mb = S * (4 * args.width * args.height * 1.5 * 2 + max(h.host_bytes() for h in sets[args.bitdepth_max])) / 1e6

    ev_start, ev_stop = L.dav1d_cuda_event_create(), L.dav1d_cuda_event_create()
    ev_done = [L.dav1d_cuda_event_create() for _ in ctxs]
    step_no = [0]
    host_s = [0.0]
    host_cpu = [0.0]

    def run_step(e2e=False, resident_rotate=True):
        """One frame of every stream.  Device-resident arm: the descriptor sets of the step were
        shipped before the timed region (all N_SETS sets of a stream cannot be resident in ONE arena,
        so this arm replays set 0); e2e arm: every stream takes its NEXT frame - the descriptor set
        goes host -> device from pinned memory, the group is submitted, the picture comes back."""
        t0, c0 = time.perf_counter(), time.thread_time()
        k = step_no[0]
        step_no[0] += 1
        for u in units:
            ctx, multi, gdfs = u
            if e2e:
                for df in gdfs:
                    df.use(k)
                    df.upload_descriptors_pinned()
                u[1] = multi = F.MultiFrame(ctx, gdfs)          # batch pointers of the sets in use: host structs only
            multi.launch()
            if e2e:
                for df in gdfs:
                    df.download_pinned()
        host_s[0] += time.perf_counter() - t0
        host_cpu[0] += time.thread_time() - c0

    def timed(nsteps, e2e=False):
        """fork: every group's stream waits for ev_start; join: main stream waits for every done event."""
        L.dav1d_cuda_event_record(main_ctx, ev_start)
        for c in ctxs:
            L.dav1d_cuda_stream_wait_event(c, ev_start)
        for _ in range(nsteps):
            run_step(e2e)
        for c, ev in zip(ctxs, ev_done):
            L.dav1d_cuda_event_record(c, ev)
            L.dav1d_cuda_stream_wait_event(main_ctx, ev)
        L.dav1d_cuda_event_record(main_ctx, ev_stop)
        return L.dav1d_cuda_event_elapsed_ms(ev_start, ev_stop)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- device-resident arm: descriptors, coefficients and references are in HBM when the timed
    # region starts; a step = one group submission per group (a handful of launches: nothing is
    # scheduled, merged or captured on the host)
    for u in units:
        u[1] = F.MultiFrame(u[0], u[2], graph=args.graph)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)          # nvidia-smi needs a moment before its first sample
    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        run_step()
    barrier()
    launches0 = L.dav1d_cuda_launch_count()
    barrier()
    sampler.mark()
    host_s[0] = 0.0
    ms = timed(args.steps)
    barrier()
    sampler.mark()
    launches = L.dav1d_cuda_launch_count() - launches0
    host_resident_ms = host_s[0] * 1e3 / (args.steps * S)
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ms)
    for c in ctxs:
        if L.dav1d_cuda_synchronize(c):
            pkg.check_error()
    pkg.check_error()
    value = world * S * args.steps * luma_px / (ms * 1e-3) / 1e6
    algo_step = sum(df.hf.algo_bytes for df in dfs)
    # the same arm with the dependency levels worked out on the device (nothing but the descriptors
    # themselves comes from the recorder)
    for df in dfs:
        df.set_levels_recorded(False)
    for _ in range(2):
        run_step()
    barrier()
    ms_dev = max_over_ranks(timed(max(2, args.steps // 2)))
    barrier()
    value_device_levels = world * S * max(2, args.steps // 2) * luma_px / (ms_dev * 1e-3) / 1e6
    for df in dfs:
        df.set_levels_recorded(True)
    pkg.check_error()

    # ---- verification of the timed configuration (outside the timed region): one 10-bit and the
    # 12-bit stream of this rank against the oracle (the reference's C templates), bit for bit
    verified = None
    if rank == 0 and not args.no_verify:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import refdsp
        import refframe
        ref = refdsp.RefDSP()
        verified = []
        for s_idx in sorted({0, S - 1}):
            df = dfs[s_idx]
            bd = df.hf.bdmax
            refs = plane_cache[bd][:2]
            df.upload_picture(df.dst, plane_cache[bd][2])
            want = refframe.run_oracle(ref, df.hf, [p.copy() for p in plane_cache[bd][2]], refs)
            mf1 = F.MultiFrame(df.ctx, [df])
            mf1.launch()
            got = df.download_picture()
            ok = all(np.array_equal(a, b) for a, b in zip(want, got))
            verified.append({"stream": s_idx, "bitdepth_max": bd, "bit_exact_vs_oracle": bool(ok)})
            if not ok:
                raise RuntimeError(f"stream {s_idx}: output differs from the oracle")
        pkg.check_error()

    # ---- per-launch-class timing (roofline of the dominant class), measured live with CUDA events
    # in the regime of the timed region: for every class a submission per group that holds only that
    # class's launches (dav1d_cuda_recon_group_submit_phases), all groups in flight at once.
    # Reported per FRAME: class time of one step / frames per step.
    roof = None
    if rank == 0:
        peak, peak_src = peaks()
        cls_bits = {"mc_put": 1, "mc_compound": 2, "warp": 4, "itx": 8, "intra": 16}
        cls_ms = {}
        for name, bit in cls_bits.items():
            cg = [F.MultiFrame(ctx, gdfs, phase_mask=bit) for ctx, _, gdfs in units]

            def step_cls():
                for g in cg:
                    g.launch()
            for _ in range(2):
                step_cls()
            torch.cuda.synchronize()
            L.dav1d_cuda_event_record(main_ctx, ev_start)
            for c in ctxs:
                L.dav1d_cuda_stream_wait_event(c, ev_start)
            reps = 3
            for _ in range(reps):
                step_cls()
            for c, ev in zip(ctxs, ev_done):
                L.dav1d_cuda_event_record(c, ev)
                L.dav1d_cuda_stream_wait_event(main_ctx, ev)
            L.dav1d_cuda_event_record(main_ctx, ev_stop)
            cls_ms[name] = L.dav1d_cuda_event_elapsed_ms(ev_start, ev_stop) / (reps * S)
        dom = max(cls_ms, key=lambda k: cls_ms[k])
        alg = sum(df.hf.algo_class[dom] for df in dfs) / S
        achieved = alg / (cls_ms[dom] * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r2_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(dom, {}).get("dram_bytes_per_frame")
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "traffic_source": "profiles/r2_traffic.json (ncu dram__bytes of this class at the benched shape)"
                if traffic else None,
                "algorithmic_bytes": alg, "ms": cls_ms[dom],
                "per": "frame (launch class of one frame; class time of a step / frames per step)",
                "per_class_ms": cls_ms,
                "per_class_gbs": {k: (sum(df.hf.algo_class[k] for df in dfs) / S / (v * 1e-3) / 1e9 if v > 0 else None)
                                  for k, v in cls_ms.items()},
                "frame_algorithmic_bytes": dfs[0].hf.algo_bytes,
                "frame_algorithmic_bytes_packed_coefs": dfs[0].hf.algo_bytes - dfs[0].hf.dense_coef_bytes +
                (dfs[0].hf.cf.nbytes if dfs[0].hf.cf16 is None else dfs[0].hf.cf16.nbytes),
                "whole_step": {"achieved": algo_step * args.steps / (ms * 1e-3) / 1e9,
                               "frac": algo_step * args.steps / (ms * 1e-3) / 1e9 / peak}}

    # ---- end to end with host buffers (pinned): every step takes the NEXT frame of every stream
    # (a different descriptor set than the step before), ships it H2D, submits, reads the picture back
    e2e = None
    if not args.no_e2e:
        shared_pinned = {}
        for df in dfs:
            df.alloc_pinned(share=shared_pinned)
        # regroup: --e2e-group streams per submission, every group on its own context / CUDA stream
        for u in units:
            if u[1] is not None:
                u[1].close()
        GE = args.e2e_group
        e2e_ctxs = [F.open_context(local) for _ in range(0, S, GE)]
        units[:] = []
        for gi, g0 in enumerate(range(0, S, GE)):
            gdfs = dfs[g0:g0 + GE]
            for df in gdfs:
                df.ctx = e2e_ctxs[gi]
            units.append([e2e_ctxs[gi], None, gdfs])
        ctxs[:] = e2e_ctxs
        while len(ev_done) < len(ctxs):
            ev_done.append(L.dav1d_cuda_event_create())
        for df in dfs:
            df.set_levels_recorded(False)       # nothing is prepared on the host: the device finds the levels
        for _ in range(2):
            run_step(e2e=True)
        barrier()
        e2e_steps = max(3, min(args.steps, 10))
        host_s[0] = host_cpu[0] = 0.0
        ems = max_over_ranks(timed(e2e_steps, e2e=True))
        barrier()
        pkg.check_error()
        h2d = sum(sum(st["bytes"] for st in df._sets) / len(df._sets) for df in dfs)
        d2h = sum(df.pinned_out_bytes for df in dfs)
        e2e = {"value": world * S * e2e_steps * luma_px / (ems * 1e-3) / 1e6, "unit": "Mpix/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
               "ms_per_step": ems / e2e_steps,
               "fresh_frames": f"every stream cycles through {N_SETS} different descriptor sets, one per step",
               "includes": ["descriptor + coefficient upload", "dependency levels + sort of the intra-class operations "
                            "(on the device, inside the submission)", "group submission (launches only: the library does no "
                            "host-side scheduling, table merging or graph capture)", "picture download"],
               "host_ms_per_frame": host_s[0] * 1e3 / (e2e_steps * S), "host_threads": 1,
               "host_cpu_ms_per_frame": host_cpu[0] * 1e3 / (e2e_steps * S),
               "host_note": "host_ms is the wall time of the submitting thread (it blocks when the launch queue is "
                            "full); host_cpu_ms is its CPU time",
               "frames_per_group_submission": GE, "groups_in_flight": len(units)}
        # the bound of this arm: what the PCIe link of THIS box moves when both directions are busy with plain
        # pinned copies in the step's byte proportion (measured, torch streams; not part of any timed region)
        e2e["pcie"] = pcie_probe(local, int(h2d), int(d2h), ems / e2e_steps)
        for u in units:
            for df in u[2]:
                df.use(0)
                df.upload_descriptors_pinned()      # the arena holds the last e2e frame: back to set 0
            L.dav1d_cuda_synchronize(u[0])
            u[1] = F.MultiFrame(u[0], u[2])
        for df in dfs:
            df.set_levels_recorded(True)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        # bounded sample: one frame per host thread, a few steps (about 10-30 s of CPU work in total)
        r = cpu_reference_run(args, steps=6, warmup=1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        out = {"metric": "Mpix/s of mc+itx+ipred recon", "value": value, "unit": "Mpix/s", "n_gpus": world,
               "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "u16" if args.bitdepth_max > 0xff else "u8", "data": "synthetic",
               "config": workload_config(args),
               "config_extra": {
                   "streams_per_gpu": S, "frames_per_group_submission": G,
                   "l2": f"inputs larger than L2: working set {footprint_mb:.0f} MB per GPU vs {L2_MB:.0f} MB L2"
                   if footprint_mb > 2 * L2_MB else f"working set {footprint_mb:.0f} MB; L2 NOT exceeded",
                   "cuda_graph": bool(args.graph),
                   "mc_window_staging": {"tma": "cp.async.bulk.tensor.2d (TMA, two window buffers per warp) in the 32x32-tile kernels",
                                         "tma_put": "cp.async.bulk.tensor.2d (TMA, two window buffers per warp) for single-reference "
                                                    "32x32 tiles; cp.async per lane for compound (measured faster) and 8x8 tiles",
                                         "cp_async": "cp.async per lane"}[args.mc_staging],
                   "submission": f"{len(units)} group submissions per step, each the frames of {G} streams "
                                 f"(dav1d_cuda_recon_group_submit: nothing scheduled on the host)",
                   "twelve_bit_stream": "the last stream of every rank is 12-bit",
                   "coefficient_stream": ("int16 storage + escape list (Dav1dCudaReconBatch.cf_int16; the CPU arm reads "
                                          "the int32 stream the reference keeps)" if args.coefs == "int16" else "int32"),
                   "intra_ops_per_frame": int(dfs[0].hf.n_intra),
                   "intra_dependency_levels": "recorder-side in `value` (dav1d_cuda_intra_levels: a linear pass over "
                                              "the descriptors in decode order, measured below, part of recording like "
                                              "the descriptors themselves); device-side in `value_device_levels` and e2e",
                   "recorder_levels_ms_per_frame_one_core": recorder_levels_ms,
                   "value_device_levels": value_device_levels,
                   "launches_per_frame": launches / max(1, args.steps * S),
                   "host_ms_per_frame_resident": host_resident_ms},
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
               "verified": verified,
               "hbm_frac_of_8TBs": algo_step * args.steps / (ms * 1e-3) / 8e12}
        print(json.dumps(out), flush=True)
    for u in units:
        if u[1] is not None:
            u[1].close()
    for df in dfs:
        df.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
