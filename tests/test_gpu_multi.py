"""Several GPUs in one process, asynchronous (deferred-check) calls, the byte-budgeted host pipeline and the
optional NCCL gather.  The multi-GPU tests skip on a one-GPU box; their host-side logic (shard bounds, gather over
gloo) is covered in the CPU tier (tests/test_sharding.py)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hpg():
    import hpgeom_b200
    return hpgeom_b200


def _ngpu():
    import torch
    return torch.cuda.device_count()


def sky(n, seed):
    rng = np.random.default_rng(seed)
    return 360.0 * rng.random(n), np.degrees(np.arcsin(2.0 * rng.random(n) - 1.0))


def test_second_device_in_one_process(hpg, oracle_port):
    """The > 48 KiB dynamic shared memory opt-in is a per-device function attribute: kernels that were first
    launched on cuda:0 must also launch on cuda:1.  Every kernel family with a big staging
    ring or table: angle_to_pixel (TMA ring), pixel_to_angle / interpolation / boundaries (tables)."""
    import torch
    if _ngpu() < 2:
        pytest.skip("needs two GPUs")
    n = 300_001
    lon, lat = sky(n, 5)
    nside = 2 ** 20
    ref_pix = oracle_port.angle_to_pixel(nside, lon, lat)
    outs = []
    for dev in (0, 1, 0):
        d = torch.device("cuda", dev)
        tl, tb = torch.from_numpy(lon).to(d), torch.from_numpy(lat).to(d)
        pix = hpg.angle_to_pixel(nside, tl, tb)
        assert pix.device == d and np.count_nonzero(pix.cpu().numpy() != ref_pix) <= 1
        a, b = hpg.pixel_to_angle(nside, pix)
        x, y, z = hpg.pixel_to_vector(nside, pix)
        p4, w4 = hpg.get_interpolation_weights(4096, tl, tb)
        ba, bb = hpg.boundaries(4096, pix[:100_000] % (12 * 4096 ** 2), step=2)
        nb = hpg.neighbors(4096, pix[:100_000] % (12 * 4096 ** 2), nest=False)
        ring = hpg.nest_to_ring(nside, pix)
        outs.append([t.cpu() for t in (pix, a, b, x, y, z, p4, w4, ba, bb, nb, ring)])
    for other in outs[1:]:
        for t0, t1 in zip(outs[0], other):
            assert torch.equal(t0, t1)
    # numpy callers on the second GPU
    hpg.set_default_device(1)
    try:
        assert np.array_equal(hpg.angle_to_pixel(nside, lon, lat), outs[0][0].numpy())
        assert np.array_equal(hpg.pixel_to_angle(nside, outs[0][0].numpy())[1], outs[0][2].numpy())
    finally:
        hpg.set_default_device(0)


def test_numpy_call_sharded_over_all_gpus(hpg, oracle_port):
    """set_default_device("all"): one numpy call, contiguous shards over every GPU of the box, one process
    (hpgeom/hpgeom.c:272-279 split); same results, and the FIRST bad element of the whole array still wins."""
    if _ngpu() < 2:
        pytest.skip("needs two GPUs")
    n = 3_000_001
    lon, lat = sky(n, 6)
    nside = 2 ** 29
    one = hpg.angle_to_pixel(nside, lon, lat)
    hpg.set_default_device("all")
    try:
        assert np.array_equal(hpg.angle_to_pixel(nside, lon, lat), one)
        th, ph = hpg.pixel_to_angle(nside, one, lonlat=False)
        nb = hpg.neighbors(nside, one[:1_000_001])
        p4, w4 = hpg.get_interpolation_weights(nside, lon, lat)
        hpg.set_default_device(0)
        assert np.array_equal(hpg.pixel_to_angle(nside, one, lonlat=False)[0], th)
        assert np.array_equal(hpg.neighbors(nside, one[:1_000_001]), nb)
        q4, v4 = hpg.get_interpolation_weights(nside, lon, lat)
        assert np.array_equal(p4, q4) and np.array_equal(w4, v4)
        hpg.set_default_device("all")
        bad = lat.copy()
        bad[n - 5] = 95.0       # in the last shard
        bad[n // 2 + 17] = -93.0  # earlier: must be the one reported
        with pytest.raises(ValueError, match=r"lat = -93 out of range"):
            hpg.angle_to_pixel(nside, lon, bad)
        hpg.set_devices([1])
        assert np.array_equal(hpg.angle_to_pixel(nside, lon, lat), one)
    finally:
        hpg.set_devices(None)
        hpg.set_default_device(0)


def test_deferred_checks_keep_the_stream_asynchronous(hpg, oracle_port):
    """Inside deferred_checks() a chain of device-resident calls queues without any read-back; the ValueError
    the reference would have raised comes out of raise_if_bad(), for the first failing CALL."""
    import torch
    n = 1_000_003
    lon, lat = sky(n, 7)
    nside = 2 ** 29
    tl, tb = torch.from_numpy(lon).cuda(), torch.from_numpy(lat).cuda()
    with hpg.deferred_checks() as pending:
        pix = hpg.angle_to_pixel(nside, tl, tb)
        ring = hpg.nest_to_ring(nside, pix)
        back = hpg.ring_to_nest(nside, ring)
        assert len(pending) == 3
    pending.raise_if_bad()
    assert len(pending) == 0
    assert np.count_nonzero(pix.cpu().numpy() != oracle_port.angle_to_pixel(nside, lon, lat)) <= 1
    assert np.array_equal(ring.cpu().numpy(), oracle_port.nest_to_ring(nside, pix.cpu().numpy()))
    assert torch.equal(back, pix)
    tb2 = tb.clone()
    tb2[12345] = 91.0
    with hpg.deferred_checks() as pending:
        p2 = hpg.angle_to_pixel(nside, tl, tb2)            # bad latitude: flagged, not raised here
        r2 = hpg.nest_to_ring(nside, p2)                   # its -1 pixel fails the next call too
    assert int(p2[12345]) == -1
    with pytest.raises(ValueError, match=r"lat = 91 out of range"):
        pending.raise_if_bad()
    # outside the context the same call raises at once
    with pytest.raises(ValueError, match=r"lat = 91 out of range"):
        hpg.angle_to_pixel(nside, tl, tb2)
    # concurrent streams do not share a status word
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    with hpg.deferred_checks() as pending:
        with torch.cuda.stream(s1):
            a = hpg.angle_to_pixel(nside, tl, tb2)
        with torch.cuda.stream(s2):
            b = hpg.angle_to_pixel(nside, tl, tb)
    torch.cuda.synchronize()
    assert int(a[12345]) == -1 and int(b[12345]) >= 0
    with pytest.raises(ValueError):
        pending.raise_if_bad()


def test_host_pipeline_is_budgeted_in_bytes(hpg, oracle_port):
    """Row-valued outputs: upgrade_pixels 64 -> 8192 writes 128 KiB per input pixel.  The host pipeline sizes its
    slots in bytes (96 MiB), so this runs in many small chunks instead of asking for GBs of pinned memory."""
    rng = np.random.default_rng(8)
    pix = rng.integers(0, 12 * 64 * 64 - 1, 3000)
    for nest in (True, False):
        got = hpg.upgrade_pixels(64, pix, 8192, nest=nest)
        assert got.shape == (3000 * 128 * 128,)
        assert np.array_equal(got, oracle_port.upgrade_pixels(64, pix, 8192, nest=nest))
    a, b = hpg.boundaries(4096, rng.integers(0, 12 * 4096 ** 2, 20_000), step=64)
    assert a.shape == (20_000, 256) and np.isfinite(a).all() and np.isfinite(b).all()
    # reference edge behaviour of upgrade_pixels for degenerate nside
    with pytest.raises(ValueError, match=r"Input pixels/ranges out of range\."):
        hpg.upgrade_pixels(0, np.array([0]), 2)
    with pytest.raises(ValueError, match=r"nside 0 must be positive"):
        hpg.upgrade_pixels(0, np.array([0]), 2, nest=False)
    with pytest.raises(ValueError, match=r"Illegal nside value"):
        hpg.upgrade_pixels(2 ** 30, np.array([0]), 2 ** 30)
    with pytest.raises(ValueError, match=r"must not be greater than 2\*\*29"):
        hpg.upgrade_pixels(2 ** 30, np.array([0]), 2 ** 31, nest=False)


def _gather_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import hpgeom_b200 as hpg
    from hpgeom_b200.sharding import gather, shard_bounds
    ok = True
    for n in (2_000_001, 2_000_000):  # ragged and equal shards
        lon, lat = sky(n, 9)
        lo, hi = shard_bounds(n, world, rank)
        pix = hpg.angle_to_pixel(2 ** 29, torch.from_numpy(lon[lo:hi]).cuda(), torch.from_numpy(lat[lo:hi]).cuda())
        full = gather(pix)
        nb = gather(hpg.neighbors(2 ** 29, pix[:1000]))
        want = hpg.angle_to_pixel(2 ** 29, torch.from_numpy(lon).cuda(), torch.from_numpy(lat).cuda())
        ok = ok and bool(torch.equal(full, want)) and nb.shape == (1000 * world, 8)
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gather_over_nccl(hpg):
    """sharding.gather on real NCCL: one process per GPU, shards converted where they live, all-gathered into one
    preallocated tensor per rank (equal shards: all_gather_into_tensor; ragged: broadcasts into slices)."""
    if _ngpu() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res)
