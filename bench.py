#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (libhpgb.so kernels)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle/_ref)

Workload (config.workload): angle_to_pixel, nside=2^29, nest, lonlat, degrees -- one shard of
BASELINE config 5 per GPU (2^30 uniform-on-sphere points = 16 GiB in + 8 GiB out per GPU, weak
scaling: 8 GPUs = 8.6e9 points).  A step is ONE pass of the kernel over the resident shard.
Inputs (24 GB) are ~200x larger than L2, so nothing survives between steps.

value  : whole-job Gpoints/s, inputs resident in HBM, CUDA events on the launch stream,
         barrier + synchronize on both sides, max over ranks.
e2e    : same metric through the C ABI with HOST buffers (hpgb_host_angle_to_pixel: pinned host
         memory -> chunked H2D / kernel / D2H pipeline -> pinned host memory), timed on the host
         around the blocking call, max over ranks.
roofline: 24 algorithmic bytes per point / the kernel's mean launch time, against the measured
         HBM copy bandwidth of MEASURED_PEAKS.json.
cpu_baseline: the reference's own C (oracle/_ref, all host threads) on a bounded sample, rank 0, N=1.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NSIDE = 2 ** 29
BYTES_PER_POINT = {"angle_to_pixel": 24, "nest_to_ring": 16}
METRIC = "Gpoints/s angle_to_pixel @nside=2^29 nest lonlat degrees (device-resident)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-points", type=int, default=30, help="points per GPU = 2^this")
    ap.add_argument("--log2-e2e-points", type=int, default=27, help="host-buffer batch per GPU = 2^this")
    ap.add_argument("--log2-ref-points", type=int, default=25, help="reference arm: points per step = 2^this")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layout", default="separate", choices=["separate", "carved"],
                    help="operand placement in HBM: three separately allocated tensors (what a caller has), or "
                         "lon | lat | pix carved back to back from one allocation")
    ap.add_argument("--log2-parity-head", type=int, default=24,
                    help="parity sample per rank = the first 2^this elements of the shard + every 1024th after them")
    ap.add_argument("--sustained-launches", type=int, default=150,
                    help="after the timed region: this many back-to-back launches of the headline kernel, reported under "
                         "'sustained' with the clocks seen (0 = skip).  The kernel draws the board's power limit: beyond "
                         "~40 ms of load the SM clock is capped (sw_power_cap) and the rate settles ~15 %% lower.")
    ap.add_argument("--gather", action="store_true",
                    help="also time the OPTIONAL collective: all-gather of a 2^28-pixel result shard per GPU (NCCL)")
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, points):
    """DRAM bytes per launch of the dominant kernel, from the committed `ncu --set full` capture of THIS kernel at
    the metric's size (profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one launch over
    2^30 points, separately allocated operands; the json names its source capture under profiles/), scaled by the point
    count when the bench runs at another size; None if the file is absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            rec = json.load(f)[kernel]
        return rec["dram_bytes_per_launch"] * (points / rec["points_per_launch"])
    except Exception:
        return None


def boundary_induced(ref, nside, a, b, got, n_threads=1):
    """True where the reference returns `got` for some <= 2 ulp perturbation of the input angles: the mismatch is
    a point within 2 ulp of a pixel boundary (BASELINE north_star: such cases are counted and reported)."""
    import numpy as np
    ok = np.zeros(a.size, dtype=bool)
    for da in (-2, -1, 0, 1, 2):
        for db in (-2, -1, 0, 1, 2):
            pa, pb = a + da * np.spacing(a), b + db * np.spacing(b)
            try:
                ok |= ref(nside, pa, pb) == got
            except ValueError:
                pass
    return ok


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.

    The sampler process is started before the warm-up (nvidia-smi needs ~100 ms to produce its
    first row); every row is stamped on arrival and only rows that arrived inside
    [mark_begin(), mark_end()] are summarised (all rows under load if that window is too short
    to contain three samples)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            time.sleep(0.3)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
        rows = [r for t, r in self.rows if len(r) >= 9 and self.t0 is not None and self.t0 <= t <= (self.t1 or 1e30) + 0.03]
        window = "timed region"
        if len(rows) < 3:
            rows = [r for _, r in self.rows if len(r) >= 9]
            window = "warm-up + timed region"

        def num(v):
            try:
                return float(v)
            except ValueError:
                return None

        sm = [num(r[1]) for r in rows if num(r[1]) is not None]
        mx = [num(r[2]) for r in rows if num(r[2]) is not None]
        pw = [num(r[3]) for r in rows if num(r[3]) is not None]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm),
                "window": window}


def sky_points_host(n, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    lon = 360.0 * rng.random(n)
    lat = np.degrees(np.arcsin(2.0 * rng.random(n) - 1.0))
    return lon, lat


def cpu_reference_rate(n, threads, reps=1):
    """Gpoints/s of the reference's own C path (oracle/_ref) -- or the scalar port -- on n points."""
    import oracle
    lon, lat = sky_points_host(n, 12345)
    if oracle.ref is not None:
        kind, cores = "reference", threads
        fn = lambda: oracle.ref.angle_to_pixel(NSIDE, lon, lat, nest=True, lonlat=True, degrees=True, n_threads=threads)
    else:
        kind, cores = "port", 1
        port = oracle.port()
        fn = lambda: port.angle_to_pixel(NSIDE, lon, lat)
    fn()
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return n / best / 1e9, kind, cores, best


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the same path, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n = 1 << args.log2_ref_points
    import oracle
    lon, lat = sky_points_host(n, 12345)
    if oracle.ref is not None:
        kind, cores = "reference", threads
        step = lambda: oracle.ref.angle_to_pixel(NSIDE, lon, lat, nest=True, lonlat=True, degrees=True, n_threads=threads)
    else:
        kind, cores = "port", 1
        port = oracle.port()
        step = lambda: port.angle_to_pixel(NSIDE, lon, lat)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = n * args.steps / dt / 1e9
    sample = "2^%d uniform-on-sphere points per step (bounded sample of the 2^%d-point-per-GPU workload)" % (
        args.log2_ref_points, args.log2_points)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gpoints/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "angle_to_pixel nside=2^29 nest lonlat degrees; reference CPU path "
                               "(unmodified hpgeom C compiled into oracle/_ref), n_threads=%d, %s" % (cores, sample)},
        "cpu_baseline": {"value": value, "unit": "Gpoints/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Gpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from hpgeom_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    n = 1 << args.log2_points
    gen = torch.Generator(device=dev).manual_seed(12345 + rank)
    # HBM layout.  Default: three ordinary, separately allocated tensors -- what a caller of the drop-in has.
    # (--layout carved = one allocation cut in three, the round-1 layout: a few percent faster on some boxes.)
    pool = None
    if args.layout == "carved":
        pool = torch.empty(3 * n, dtype=torch.float64, device=dev)
        lon, lat, pix = pool[:n], pool[n:2 * n], pool[2 * n:].view(torch.int64)
    else:
        lon = torch.empty(n, dtype=torch.float64, device=dev)
        lat = torch.empty(n, dtype=torch.float64, device=dev)
        pix = torch.empty(n, dtype=torch.int64, device=dev)
    torch.rand(n, dtype=torch.float64, device=dev, generator=gen, out=lon).mul_(360.0)
    torch.rand(n, dtype=torch.float64, device=dev, generator=gen, out=lat).mul_(2.0).sub_(1.0).asin_().mul_(180.0 / 3.141592653589793)
    status = torch.empty(2, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)
    stp = ctypes.c_void_p(status.data_ptr())

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_a2p():
        rc = L.hpgb_angle_to_pixel(lon.data_ptr(), lat.data_ptr(), pix.data_ptr(), n, NSIDE, None, 1, 1, 1, stp, sp)
        assert rc == 0, _lib.error_string(rc)

    def timed(step, steps, warmup, sampler=None):
        L.hpgb_status_reset(stp, sp)
        for _ in range(warmup):
            step()
        barrier()
        l0 = L.hpgb_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler is not None:
            sampler.mark_begin()
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        barrier()
        if sampler is not None:
            sampler.mark_end()
        ms = e0.elapsed_time(e1)
        launches = L.hpgb_launch_count() - l0
        t = torch.tensor([ms, float(launches)], dtype=torch.float64, device=dev)
        if world > 1:
            mx = t.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = t.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            return float(mx[0]), int(sm[1])
        return ms, launches

    # ---- headline: angle_to_pixel, device resident ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches = timed(step_a2p, args.steps, max(args.warmup, 3), sampler if rank == 0 else None)
    clocks = sampler.stop() if rank == 0 else None
    assert int(status[0].item()) == -1, "kernel flagged a bad element in valid synthetic input"
    ms_per_step = ms / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e9

    # ---- the same kernel under sustained load (informational; `value` above is the contract's K-step burst) ----
    sustained = None
    if args.sustained_launches > 0:
        s2 = ClockSampler(local)
        if rank == 0:
            s2.start()
        ms_s, _ = timed(step_a2p, args.sustained_launches, 0, s2 if rank == 0 else None)
        c2 = s2.stop() if rank == 0 else None
        v_s = world * n / (ms_s / args.sustained_launches * 1e-3) / 1e9
        sustained = {"value": v_s, "unit": "Gpoints/s", "launches": args.sustained_launches,
                     "ms_per_launch": ms_s / args.sustained_launches, "clocks": c2,
                     "note": "back-to-back launches right after the timed region: the kernel runs into the board power "
                             "limit (sw_power_cap), the SM clock drops and the issue-limited kernel with it"}
        time.sleep(1.0)  # let the clocks recover before the other kernels are timed

    lon_keep, lat_keep = lon[:1 << 20].clone(), lat[:1 << 20].clone()
    # parity sample of THIS rank's shard: the first 2^24 elements + every 1024th after them
    head = min(n, 1 << args.log2_parity_head)
    sample_idx = torch.cat([torch.arange(head, device=dev), torch.arange(head, n, 1024, device=dev)])
    s_lon, s_lat, s_pix = lon[sample_idx].cpu().numpy(), lat[sample_idx].cpu().numpy(), pix[sample_idx].cpu().numpy()
    # ---- second headline of the metric: nest_to_ring on the pixels just produced ----
    ring = torch.empty(n, dtype=torch.int64, device=dev) if args.layout == "separate" else lon.view(torch.int64)
    if args.layout == "separate":
        del lon, lat

    def step_n2r():
        rc = L.hpgb_nest_to_ring(pix.data_ptr(), ring.data_ptr(), n, NSIDE, None, stp, sp)
        assert rc == 0, _lib.error_string(rc)

    ms_n2r, _ = timed(step_n2r, args.steps, max(args.warmup, 3))
    ms_n2r /= args.steps
    n2r_value = world * n / (ms_n2r * 1e-3) / 1e9
    # ---- the other kernels of the path, same protocol (device resident, CUDA events, max over ranks), smaller
    # batches: 2^27 elements at nside 2^29 (configs 2/3), 2^24 at nside 4096 (config 4).  Reported under "also".
    also = {}

    def add_also(name, step, elems, bytes_per_elem, workload):
        ms_k, _ = timed(step, 5, 3)
        ms_k /= 5
        assert int(status[0].item()) == -1, "%s flagged element %d of valid synthetic input (rank %d)" % (
            name, int(status[0].item()), rank)
        also[name] = {"value": world * elems / (ms_k * 1e-3) / 1e9, "unit": "Gelements/s", "ms_per_step": ms_k,
                      "bytes_per_element": bytes_per_elem, "workload": workload}

    def ck(rc):
        assert rc == 0, _lib.error_string(rc)

    m = min(n, 1 << 27)
    ms8 = min(n, 1 << 24)
    f1, f2, f3 = (torch.empty(m, dtype=torch.float64, device=dev) for _ in range(3))
    i1 = torch.empty(m, dtype=torch.int64, device=dev)
    # ring_to_nest writes to a scratch array: it is NOT the inverse of nest_to_ring for every pixel at nside 2^29 --
    # the reference's uncorrected isqrt (hpgeom/healpix_geom.c:60) sends ~1e-8 of the RING pixels of the largest cap
    # rings to garbage, which the kernel reproduces bit for bit; `pix` must stay valid for the kernels below
    add_also("ring_to_nest", lambda: ck(L.hpgb_ring_to_nest(ring.data_ptr(), i1.data_ptr(), m, NSIDE, None, stp, sp)), m, 16,
             "ring_to_nest nside=2^29 on 2^%d pixels per GPU (the nest_to_ring output)" % (m.bit_length() - 1))
    add_also("pixel_to_angle", lambda: ck(L.hpgb_pixel_to_angle(pix.data_ptr(), f1.data_ptr(), f2.data_ptr(), m, NSIDE, None, 1, 1, 1, stp, sp)),
             m, 24, "pixel_to_angle nside=2^29 nest lonlat degrees, 2^%d pixels per GPU" % (m.bit_length() - 1))
    add_also("pixel_to_vector", lambda: ck(L.hpgb_pixel_to_vector(pix.data_ptr(), f1.data_ptr(), f2.data_ptr(), f3.data_ptr(), m, NSIDE, None, 1, stp, sp)),
             m, 32, "pixel_to_vector nside=2^29 nest, 2^%d pixels per GPU" % (m.bit_length() - 1))
    add_also("vector_to_pixel", lambda: ck(L.hpgb_vector_to_pixel(f1.data_ptr(), f2.data_ptr(), f3.data_ptr(), i1.data_ptr(), m, NSIDE, None, 1, stp, sp)),
             m, 32, "vector_to_pixel nside=2^29 nest on the vectors just produced, 2^%d per GPU" % (m.bit_length() - 1))
    assert torch.equal(i1[:1 << 20], pix[:1 << 20]), "vector_to_pixel(pixel_to_vector(p)) != p"
    # BASELINE config 2 as written: nside 2^20, 10^8 random pixels, both schemes (ring-table form of pixel_to_angle:
    # >= 8 pixels per ring)
    m2 = min(m, 100_000_000)
    p20 = pix[:m2] % (12 * (1 << 20) ** 2)
    for nest_flag, tag in ((1, "nest"), (0, "ring")):
        add_also("pixel_to_angle_config2_" + tag,
                 lambda nf=nest_flag: ck(L.hpgb_pixel_to_angle(p20.data_ptr(), f1.data_ptr(), f2.data_ptr(), m2, 1 << 20, None, nf, 1, 1, stp, sp)),
                 m2, 24, "pixel_to_angle nside=2^20 %s lonlat degrees, %d random pixels per GPU (BASELINE config 2)" % (tag, m2))
        add_also("pixel_to_vector_config2_" + tag,
                 lambda nf=nest_flag: ck(L.hpgb_pixel_to_vector(p20.data_ptr(), f1.data_ptr(), f2.data_ptr(), f3.data_ptr(), m2, 1 << 20, None, nf, stp, sp)),
                 m2, 32, "pixel_to_vector nside=2^20 %s, %d random pixels per GPU (BASELINE config 2)" % (tag, m2))
    del p20
    p4 = pix[:ms8] % (12 * 4096 * 4096 - 1)
    # upgrade_pixels refuses the LAST NEST pixel like the reference (hpgeom/hpgeom.py:505); in the RING scheme that
    # is one particular pixel inside the range: keep it out of the RING batch (p4 never holds npix-1 itself)
    import hpgeom_b200
    last_ring = hpgeom_b200.nest_to_ring(4096, torch.tensor([12 * 4096 * 4096 - 1], device=dev))
    p4r = torch.where(p4 == last_ring, p4 - 1, p4)
    reps4 = (ms8 + lon_keep.numel() - 1) // lon_keep.numel()
    lon4, lat4 = lon_keep.repeat(reps4)[:ms8].contiguous(), lat_keep.repeat(reps4)[:ms8].contiguous()
    o8 = i1[:8 * ms8] if 8 * ms8 <= m else torch.empty(8 * ms8, dtype=torch.int64, device=dev)
    w4 = f1[:4 * ms8] if 4 * ms8 <= m else torch.empty(4 * ms8, dtype=torch.float64, device=dev)
    for nest_flag, tag in ((1, "nest"), (0, "ring")):
        add_also("neighbors_" + tag, lambda nf=nest_flag: ck(L.hpgb_neighbors(p4.data_ptr(), o8.data_ptr(), ms8, 4096, None, nf, stp, sp)),
                 ms8, 72, "neighbors nside=4096 %s, 2^%d pixels per GPU" % (tag, ms8.bit_length() - 1))
        add_also("interpolation_weights_" + tag,
                 lambda nf=nest_flag: ck(L.hpgb_interp_weights(lon4.data_ptr(), lat4.data_ptr(), o8.data_ptr(), w4.data_ptr(), ms8, 4096, None, nf, 1, 1, stp, sp)),
                 ms8, 80, "get_interpolation_weights nside=4096 %s lonlat degrees, 2^%d points per GPU" % (tag, ms8.bit_length() - 1))
        add_also("upgrade_pixels_" + tag,
                 lambda nf=nest_flag: ck(L.hpgb_upgrade_pixels((p4 if nf else p4r).data_ptr(), o8.data_ptr(), ms8, 4096, 8192, nf, stp, sp)),
                 ms8, 40, "upgrade_pixels nside=4096 -> 8192 %s, 2^%d input pixels per GPU" % (tag, ms8.bit_length() - 1))
    assert int(status[0].item()) == -1, "a kernel flagged a bad element in valid synthetic input"
    del f1, f2, f3, i1, o8, w4, p4, p4r, lon4, lat4
    # ---- parity accounting, every rank, outside every timed region: this rank's sample of BOTH headline kernels
    # against the reference's own C (oracle/_ref; the port if that build is absent), mismatches classified
    s_ring = ring[sample_idx].cpu().numpy()
    par = [0, 0, 0, 0, 0]  # points, a2p mismatches, of which boundary-induced, n2r points, n2r mismatches
    par_err, against = None, None
    try:
        import numpy as np
        import oracle
        nth = max(1, (os.cpu_count() or 1) // max(world, 1))
        if oracle.ref is not None:
            against = "oracle/_ref (unmodified reference C)"
            ref_a2p = lambda ns, a, b: oracle.ref.angle_to_pixel(ns, a, b, nest=True, lonlat=True, degrees=True, n_threads=nth)
            ref_n2r = lambda ns, p: oracle.ref.nest_to_ring(ns, p, n_threads=nth)
        else:
            against = "oracle port (C restatement)"
            port = oracle.port()
            ref_a2p = lambda ns, a, b: port.angle_to_pixel(ns, a, b)
            ref_n2r = lambda ns, p: port.nest_to_ring(ns, p)
        want = ref_a2p(NSIDE, s_lon, s_lat)
        bad = np.flatnonzero(s_pix != want)
        explained = int(boundary_induced(ref_a2p, NSIDE, s_lon[bad], s_lat[bad], s_pix[bad]).sum()) if bad.size else 0
        # nest_to_ring is judged on the pixels the kernel was actually fed (s_pix), bit for bit
        par = [int(s_pix.size), int(bad.size), explained, int(s_pix.size), int((s_ring != ref_n2r(NSIDE, s_pix)).sum())]
    except Exception as exc:  # the oracle is test infrastructure; its absence must not break the bench
        par_err = repr(exc)
    pt = torch.tensor(par + [1 if par_err else 0], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(pt, op=dist.ReduceOp.SUM)
    pt = [int(v) for v in pt.tolist()]
    parity = {"sample": "per rank: first 2^%d points of the shard + every 1024th after them" % (head.bit_length() - 1),
              "ranks": world, "ranks_failed_to_check": pt[5], "against": against,
              "angle_to_pixel": {"points": pt[0], "mismatches": pt[1], "boundary_induced": pt[2], "unexplained": pt[1] - pt[2]},
              "nest_to_ring": {"points": pt[3], "mismatches": pt[4]}}
    if par_err:
        parity["error_rank0"] = par_err
    # ---- optional collective (never on the data path): all-gather of a result shard over NCCL / NVLink ----
    gather = None
    if args.gather and world > 1:
        from hpgeom_b200 import sharding
        gn = min(n, 1 << 28)
        shard_t = pix[:gn]
        for _ in range(2):
            full = sharding.gather(shard_t)
            del full
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5):
            full = sharding.gather(shard_t)
        e1.record(stream)
        barrier()
        ok = bool(torch.equal(full[rank * gn:(rank + 1) * gn], shard_t))
        tg = torch.tensor([e0.elapsed_time(e1) / 5], dtype=torch.float64, device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        ms_g = float(tg[0])
        gather = {"bytes_per_rank": 8 * gn, "ms": ms_g, "received_GBps_per_gpu": 8 * gn * (world - 1) / (ms_g * 1e-3) / 1e9,
                  "busbw_GBps": 8 * gn * (world - 1) / (ms_g * 1e-3) / 1e9, "own_shard_intact": ok,
                  "nvlink_peer_copy_GBps_measured": 770.0, "api": "sharding.gather -> all_gather_into_tensor (NCCL)"}
        del full
    lon = lat = None
    del ring, lon, lat, pix, pool, lon_keep, lat_keep, sample_idx
    torch.cuda.empty_cache()

    # ---- e2e: the C ABI with HOST (pinned) buffers, copies inside the timed region ----
    ne = 1 << args.log2_e2e_points
    hlon = torch.empty(ne, dtype=torch.float64).pin_memory()
    hlat = torch.empty(ne, dtype=torch.float64).pin_memory()
    hpix = torch.empty(ne, dtype=torch.int64).pin_memory()
    g = torch.Generator().manual_seed(777 + rank)
    hlon.copy_(torch.rand(ne, dtype=torch.float64, generator=g) * 360.0)
    hlat.copy_(torch.rad2deg(torch.asin(torch.rand(ne, dtype=torch.float64, generator=g) * 2.0 - 1.0)))
    fb = ctypes.c_int64(-1)

    def step_e2e():
        rc = L.hpgb_host_angle_to_pixel(hlon.data_ptr(), hlat.data_ptr(), hpix.data_ptr(), ne, NSIDE, None, 1, 1, 1,
                                        local, ctypes.byref(fb))
        assert rc == 0 and fb.value == -1, (rc, fb.value)

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * ne * e2e_steps / float(t[0]) / 1e9
    checksum = int(hpix[:1024].sum().item())  # the step's result is read on the host
    # the same call with PAGEABLE buffers (what a numpy caller of the drop-in passes): staged through pinned
    # buffers by the library's host copy pool.  Reported next to e2e, not instead of it.
    e2e_pageable = None
    if world == 1:
        import numpy as np
        npg = 1 << 26
        plon, plat = hlon[:npg].numpy().copy(), hlat[:npg].numpy().copy()
        ppix = np.empty(npg, dtype=np.int64)
        ppix[:] = 0  # touch the pages

        def step_pg():
            rc = L.hpgb_host_angle_to_pixel(plon.ctypes.data, plat.ctypes.data, ppix.ctypes.data, npg, NSIDE, None, 1, 1, 1,
                                            local, ctypes.byref(fb))
            assert rc == 0 and fb.value == -1, (rc, fb.value)

        step_pg()
        t0 = time.perf_counter()
        for _ in range(3):
            step_pg()
        e2e_pageable = {"value": 3 * npg / (time.perf_counter() - t0) / 1e9, "unit": "Gpoints/s",
                        "points_per_step": npg, "buffers": "pageable numpy arrays (library-side pinned staging, "
                        "host copy pool)", "matches_pinned_result": bool((ppix[:1024] == hpix[:1024].numpy()).all())}

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = BYTES_PER_POINT["angle_to_pixel"] * n / (ms_per_step * 1e-3) / 1e9
        kernel = "k_stream<OpAng2Pix<nest,lonlat,degrees>,3,2,256,2>"
        line = {
            "metric": METRIC, "value": value, "unit": "Gpoints/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "angle_to_pixel nside=2^29 nest lonlat degrees, 2^%d uniform-on-sphere points per GPU "
                                   "(BASELINE config 5 shard), device-resident; inputs 16 B + output 8 B per point "
                                   "(%.1f GB per GPU) >> L2, no flush needed" % (args.log2_points, 24 * n / 1e9),
                       "sharding": "contiguous shards, one process per GPU, no data-path collective",
                       "e2e_batch_points_per_gpu": ne},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic("angle_to_pixel", n), "kernel": kernel, "peak_source": peak_src,
                         "bytes_per_point": 24, "points_per_launch": n,
                         "frac_nominal_8tbs": achieved / 8000.0, "peak_nominal": 8000.0,
                         "traffic_source": "profiles/ncu_traffic.json (ncu --set full of this kernel at 2^30 points)"},
            "e2e": {"value": e2e_value, "unit": "Gpoints/s", "h2d_bytes_per_step": 16 * ne, "d2h_bytes_per_step": 8 * ne,
                    "api": "hpgb_host_angle_to_pixel (C ABI, pinned host buffers)", "steps": e2e_steps,
                    "result_checksum": checksum},
            "e2e_pageable": e2e_pageable,
            "gpu_launches": launches,
            "clocks": clocks,
            "sustained": (dict(sustained, roofline_frac=sustained["value"] / world * 24 / peak) if sustained else None),
            "parity": parity,
            "layout": ("three separately allocated tensors" if args.layout == "separate"
                       else "lon | lat | pix carved back to back from one allocation"),
            "gather": gather,
            "also": {"nest_to_ring": {"value": n2r_value, "unit": "Gpoints/s", "ms_per_step": ms_n2r,
                                      "roofline_frac": 16 * n / (ms_n2r * 1e-3) / 1e9 / peak,
                                      "roofline_frac_nominal_8tbs": 16 * n / (ms_n2r * 1e-3) / 1e9 / 8000.0,
                                      "workload": "nest_to_ring nside=2^29 on the 2^%d pixels per GPU" % args.log2_points}},
        }
        for name, rec in also.items():  # whole-job rate -> per-GPU algorithmic GB/s -> fraction of the measured peak
            rec["roofline_frac"] = rec["value"] / world * rec["bytes_per_element"] / peak
            line["also"][name] = rec
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            probe, kind, cores, _ = cpu_reference_rate(1 << 22, threads)
            # ~10 s of CPU work: a 2^26-point sample (1.5 GB of host arrays) repeated as often as needed
            ns = 1 << 26
            reps = int(min(max(10.0 / (ns / (probe * 1e9)), 1), 200))
            rate, kind, cores, secs = cpu_reference_rate(ns, threads, reps=reps)
            line["cpu_baseline"] = {"value": rate, "unit": "Gpoints/s", "cores": cores, "kind": kind,
                                    "sample": "%d uniform-on-sphere points of the same workload x %d passes "
                                              "(best pass %.2f s), n_threads=%d" % (ns, reps, secs, cores)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
