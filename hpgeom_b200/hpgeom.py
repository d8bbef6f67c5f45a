"""Drop-in Python API for hpgeom's elementwise HEALPix conversion path, backed by sm_100a CUDA.

Same names, positional order, defaults and exceptions as the reference (hpgeom/hpgeom.py and the
C methods it re-exports, hpgeom/hpgeom.c:3898-3929).  Note the *positional* order of the flags is
``lonlat, nest, degrees`` (hpgeom/hpgeom.c:183-184), although the docstrings list ``nest`` first.
``n_threads`` is accepted for compatibility and ignored: the CUDA grid replaces the reference's
pthread split (hpgeom/hpgeom.c:240-316).

Inputs may be numpy arrays / Python scalars (results are numpy, computed on the GPU selected by
``set_default_device``) or device-resident ``torch`` CUDA tensors (results are CUDA tensors on
the same device; the kernels run on the current torch stream).  There is no CPU fallback.
"""
import ctypes
import numbers

import numpy as np

from . import _dispatch as _d
from . import _lib
from ._dispatch import F8, I8, deferred_checks, set_default_device, set_devices  # noqa: F401

__all__ = [
    'thetaphi_to_lonlat',
    'npixel_to_nside',
    'nside_to_pixel_area',
    'nside_to_resolution',
    'nside_to_order',
    'order_to_nside',
    'angle_to_vector',
    'vector_to_angle',
    'reorder',
    'boundaries',
    'max_pixel_radius',
    'angle_to_pixel',
    'pixel_to_angle',
    'lonlat_to_thetaphi',
    'vector_to_pixel',
    'pixel_to_vector',
    'nest_to_ring',
    'ring_to_nest',
    'neighbors',
    'get_interpolation_weights',
    'pixel_ranges_to_pixels',
    'query_circle',
    'query_circle_vec',
    'query_circle_batch',
    'query_polygon',
    'query_polygon_vec',
    'query_ellipse',
    'query_box',
    'upgrade_pixels',
    'upgrade_pixel_ranges',
    'nside_to_npixel',
    'set_default_device',
    'set_devices',
    'deferred_checks',
    'UNSEEN',
]

UNSEEN = -1.6375e+30
max_nside = 1 << 29


def angle_to_pixel(nside, a, b, lonlat=True, nest=True, degrees=True, n_threads=1):
    """Convert angles to pixels (reference: hpgeom/hpgeom.c:170-380).

    a, b are longitude/latitude (``lonlat=True``; degrees or radians per ``degrees``) or
    co-latitude theta / longitude phi in radians (``lonlat=False``).  Raises ValueError for
    out-of-range angles or non-broadcastable inputs, with the reference's messages.
    """
    lonlat, nest, degrees = bool(lonlat), bool(nest), bool(degrees)
    (pix,) = _d.elementwise(
        "angle_to_pixel", nside, [a, b], [F8, F8], [I8], [int(nest), int(lonlat), int(degrees)], nest=nest,
        bcast_msg="nside, a, b arrays could not be broadcast together.",
        explain=lambda ns, v: _lib.explain_angle(ns, nest, v[0], v[1], lonlat, degrees))
    return pix


def pixel_to_angle(nside, pix, lonlat=True, nest=True, degrees=True, n_threads=1):
    """Convert pixels to angles (reference: hpgeom/hpgeom.c:471-676).  Returns ``(a, b)``."""
    lonlat, nest, degrees = bool(lonlat), bool(nest), bool(degrees)
    a, b = _d.elementwise(
        "pixel_to_angle", nside, [pix], [I8], [F8, F8], [int(nest), int(lonlat), int(degrees)], nest=nest,
        bcast_msg="nside, pix arrays could not be broadcast together.",
        explain=lambda ns, v: _lib.explain_pixel(ns, nest, v[0]))
    return a, b


def nest_to_ring(nside, pix, n_threads=1):
    """Convert pixel numbers from nest to ring ordering (reference: hpgeom/hpgeom.c:1462-1653)."""
    (out,) = _d.elementwise(
        "nest_to_ring", nside, [pix], [I8], [I8], [], nest=True,
        bcast_msg="nside, pix arrays could not be broadcast together.",
        explain=lambda ns, v: _lib.explain_pixel(ns, True, v[0]))
    return out


def ring_to_nest(nside, pix, n_threads=1):
    """Convert pixel numbers from ring to nest ordering (reference: hpgeom/hpgeom.c:1732-1927)."""
    (out,) = _d.elementwise(
        "ring_to_nest", nside, [pix], [I8], [I8], [], nest=True,
        bcast_msg="nside, pix arrays could not be broadcast together.",
        explain=lambda ns, v: _lib.explain_pixel(ns, True, v[0]))
    return out


def vector_to_pixel(nside, x, y, z, nest=True, n_threads=1):
    """Convert (un-normalised) vectors to pixels (reference: hpgeom/hpgeom.c:2357-2566)."""
    nest = bool(nest)
    (pix,) = _d.elementwise(
        "vector_to_pixel", nside, [x, y, z], [F8, F8, F8], [I8], [int(nest)], nest=nest,
        bcast_msg="nside, x, y, z arrays could not be broadcast together.",
        explain=lambda ns, v: _lib.explain_nside(ns, nest))
    return pix


def pixel_to_vector(nside, pix, nest=True, n_threads=1):
    """Convert pixels to unit vectors (reference: hpgeom/hpgeom.c:2655-2866).  Returns ``(x, y, z)``."""
    nest = bool(nest)
    x, y, z = _d.elementwise(
        "pixel_to_vector", nside, [pix], [I8], [F8, F8, F8], [int(nest)], nest=nest,
        bcast_msg="nside, pix arrays could not be broadcast together.",
        explain=lambda ns, v: _lib.explain_pixel(ns, nest, v[0]))
    return x, y, z


def neighbors(nside, pix, nest=True, n_threads=1):
    """8 nearest neighbours, order SW, W, NW, N, NE, E, SE, S; -1 where there is none
    (reference: hpgeom/hpgeom.c:2973-3184).  Returns (N, 8), or (8,) for scalar input."""
    nest = bool(nest)

    def dims(ns, arrs):
        if len(arrs[0].shape) > 1:
            raise ValueError("pix array must be at most 1D.")
        if len(ns.shape) > 1:
            raise ValueError("nside array must be at most 1D.")

    (out,) = _d.elementwise(
        "neighbors", nside, [pix], [I8], [I8], [int(nest)], nest=nest, out_cols=[8], max_ndim=dims,
        bcast_msg="nside, pix arrays could not be broadcast together.",
        explain=lambda ns, v: _lib.explain_pixel(ns, nest, v[0]))
    return out


def boundaries(nside, pix, step=1, lonlat=True, nest=True, degrees=True, n_threads=1):
    """Positions along the boundary of each pixel (reference: hpgeom/hpgeom.c:2046-2278, core
    hpgeom/healpix_geom.c:807-888): ``step`` points per edge, starting at the north corner.
    Returns ``(a, b)`` of shape (4*step,) for a scalar pixel, else (N, 4*step)."""
    lonlat, nest, degrees = bool(lonlat), bool(nest), bool(degrees)
    step = int(step)
    if step < 1:
        raise ValueError("step must be positive.")

    def dims(ns, arrs):
        if len(arrs[0].shape) > 1:
            raise ValueError("pix array must be at most 1D.")

    a, b = _d.elementwise(
        "boundaries", nside, [pix], [I8], [F8, F8], [step, int(nest), int(lonlat), int(degrees)], nest=nest,
        out_cols=[4 * step, 4 * step], max_ndim=dims, bcast_msg="nside, pix arrays could not be broadcast together.",
        explain=lambda ns, v: _lib.explain_pixel(ns, nest, v[0]))
    return a, b


def max_pixel_radius(nside, degrees=True, n_threads=1):
    """Maximum angular distance between a pixel centre and its corners, per nside (reference:
    hpgeom/hpgeom.c:3255-3440, core hpgeom/healpix_geom.c:480-502)."""
    be = _d.pick_backend(nside)
    ns = be.coerce(_d._np_view(nside) if be.name == "numpy" else nside, I8)
    shape = be.shape(ns)
    n = int(np.prod(shape, dtype=np.int64))
    out = be.empty(shape, F8)
    if n == 0:
        return out
    nsc = be.expand(ns, shape)
    def on_bad(bad):
        raise ValueError(_lib.explain_nside(int(be.item(nsc, bad)), False))

    rc, bad = be.call("max_pixel_radius", [be.ptr(nsc), be.ptr(out)], [n, int(bool(degrees))], on_bad)
    _lib.check_rc(rc)
    if bad >= 0:
        on_bad(bad)
    return be.scalar(out) if len(shape) == 0 else out


def get_interpolation_weights(nside, a, b, lonlat=True, nest=True, degrees=True, n_threads=1):
    """Bilinear interpolation pixels and weights (reference: hpgeom/hpgeom.c:3537-3777).
    Returns ``(pixels (N, 4) int64, weights (N, 4) float64)``, or (4,) each for scalar input."""
    lonlat, nest, degrees = bool(lonlat), bool(nest), bool(degrees)

    def dims(ns, arrs):
        if len(arrs[0].shape) > 1:
            raise ValueError("a array must be at most 1D.")
        if len(arrs[0].shape) != len(arrs[1].shape):
            raise ValueError("a and b arrays must have same number of dimensions.")

    pix, wgt = _d.elementwise(
        "interp_weights", nside, [a, b], [F8, F8], [I8, F8], [int(nest), int(lonlat), int(degrees)], nest=nest,
        out_cols=[4, 4], max_ndim=dims, bcast_msg="nside, a, b arrays could not be broadcast together.",
        explain=lambda ns, v: _lib.explain_angle(ns, nest, v[0], v[1], lonlat, degrees))
    return pix, wgt


def lonlat_to_thetaphi(lon, lat, degrees=True):
    """Convert longitude/latitude to theta/phi (reference: hpgeom/hpgeom.py:60-88; numpy's
    ``x * (pi/180)`` and Python-style ``%``, NOT the C path's ``(x*pi)/180``).

    Like the reference, lon and lat are NOT broadcast against each other: theta has lat's shape,
    phi has lon's."""
    be = _d.pick_backend(lon, lat)
    if be.name == "numpy":
        lon, lat = _d._np_view(lon), _d._np_view(lat)
    lon_a, lat_a = be.coerce(lon, F8), be.coerce(lat, F8)
    deg = int(bool(degrees))

    def run(lo, la):
        n = int(np.prod(be.shape(lo), dtype=np.int64))
        theta, phi = be.empty(be.shape(lo), F8), be.empty(be.shape(lo), F8)
        if n == 0:
            return theta, phi
        lo, la = be.expand(lo, be.shape(lo)), be.expand(la, be.shape(la))
        def on_bad(bad):
            raise ValueError("Latitude out of range.")

        rc, bad = be.call("lonlat_to_thetaphi", [be.ptr(lo), be.ptr(la), be.ptr(theta), be.ptr(phi)], [n, deg], on_bad)
        _lib.check_rc(rc)
        if bad >= 0:
            on_bad(bad)
        return theta, phi

    if be.shape(lon_a) == be.shape(lat_a):
        theta, phi = run(lon_a, lat_a)
    else:  # evaluate each operand on its own shape, pairing it with zeros
        theta, _ = run(lat_a * 0.0, lat_a)
        _, phi = run(lon_a, lon_a * 0.0)
    if len(be.shape(theta)) == 0:
        theta = be.scalar(theta)
    if len(be.shape(phi)) == 0:
        phi = be.scalar(phi)
    return theta, phi


def thetaphi_to_lonlat(theta, phi, degrees=True):
    """Convert theta/phi (radians) to longitude/latitude (reference: hpgeom/hpgeom.py:91-112).

    ``lat = -(theta - pi/2)``, ``rad2deg(x) = x * (180/pi)``: pure IEEE, so the kernel
    (``hpgb_thetaphi_to_lonlat``) is bit-identical to numpy.  Like the reference, theta and phi are not
    broadcast against each other: lon has phi's shape, lat has theta's."""
    be = _d.pick_backend(theta, phi)
    if be.name == "numpy":
        theta, phi = _d._np_view(theta), _d._np_view(phi)
    th, ph = be.coerce(theta, F8), be.coerce(phi, F8)
    deg = int(bool(degrees))

    def run(t, p_):
        n = int(np.prod(be.shape(t), dtype=np.int64))
        lon, lat = be.empty(be.shape(t), F8), be.empty(be.shape(t), F8)
        if n == 0:
            return lon, lat
        t, p_ = be.expand(t, be.shape(t)), be.expand(p_, be.shape(p_))
        rc, _ = be.call("thetaphi_to_lonlat", [be.ptr(t), be.ptr(p_), be.ptr(lon), be.ptr(lat)], [n, deg])
        _lib.check_rc(rc)
        return lon, lat

    if be.shape(th) == be.shape(ph):
        lon, lat = run(th, ph)
    else:
        _, lat = run(th, th * 0.0)
        lon, _ = run(ph * 0.0, ph)
    if len(be.shape(lon)) == 0:
        lon = be.scalar(lon)
    if len(be.shape(lat)) == 0:
        lat = be.scalar(lat)
    return lon, lat


def npixel_to_nside(npixel):
    """nside for a number of pixels (reference: hpgeom/hpgeom.py:209-229)."""
    if np.any(npixel <= 0):
        raise ValueError("Illegal npixel (must be positive)")
    nside = np.sqrt(np.atleast_1d(npixel) / 12.0)
    if np.any(nside != np.floor(nside)):
        raise ValueError("Illegal npixel (it must be 12*nside*nside)")
    return int(nside[0]) if len(nside) == 1 else nside.astype(np.int64)


def nside_to_pixel_area(nside, degrees=True):
    """Pixel area in square degrees or steradians (reference: hpgeom/hpgeom.py:232-254)."""
    area = 4 * np.pi / nside_to_npixel(nside)
    return np.rad2deg(np.rad2deg(area)) if degrees else area


def nside_to_resolution(nside, units='degrees'):
    """sqrt(pixel area) in the requested units (reference: hpgeom/hpgeom.py:257-292)."""
    res = np.sqrt(nside_to_pixel_area(nside, degrees=False))
    if units == 'radians':
        return res
    scale = {'degrees': None, 'arcminutes': (60.,), 'arcseconds': (60., 60.)}
    if units not in scale:
        raise ValueError("Invalid units.  Must be radians, degrees, arcminutes, or arcseconds.")
    value = np.rad2deg(res)
    for f in scale[units] or ():
        value = value * f
    return value


def nside_to_order(nside):
    """log2(nside) (reference: hpgeom/hpgeom.py:295-318)."""
    _nside = np.atleast_1d(nside)
    if np.any((_nside <= 0) | (_nside > max_nside) | ((_nside & (_nside - 1)) != 0)):
        raise ValueError("nside must be postive power of 2, and less than 2**30")
    if hasattr(nside, '__len__'):
        return np.round(np.log2(nside)).astype(np.int64)
    return int(np.round(np.log2(nside)))


def order_to_nside(order):
    """2**order (reference: hpgeom/hpgeom.py:321-339)."""
    if np.any((order != np.int64(order)) | (order < 0) | (order > 29)):
        raise ValueError("Order must be integer, 0<=order<=29.")
    return 2**order


_A2V_CLASS_PHI = 1 << 62


def angle_to_vector(a, b, lonlat=True, degrees=True):
    """Angles to unit vectors, (N, 3) or (3,) (reference: hpgeom/hpgeom.py:364-393).

    One kernel (``hpgb_angle_to_vector``): numpy's ``lonlat_to_thetaphi`` formulas, then
    ``(sin t cos p, sin t sin p, cos t)`` with sin / cos correctly rounded in practice (what numpy = glibc
    returns for ~99.9 % of arguments); the ValueErrors are the reference's (hpgeom/hpgeom.py:86, :342-361)."""
    be = _d.pick_backend(a, b)
    if be.name == "numpy":
        a, b = _d._np_view(a), _d._np_view(b)
    aa, bb = be.coerce(a, F8), be.coerce(b, F8)
    shape = np.broadcast_shapes(be.shape(aa), be.shape(bb))  # numpy broadcasting of sin(theta) * cos(phi)
    n = int(np.prod(shape, dtype=np.int64))
    vec = be.empty((n, 3), F8)
    if n:
        fa, fb = be.expand(aa, shape), be.expand(bb, shape)
        def on_bad(bad):
            if lonlat:
                raise ValueError("Latitude out of range.")
            raise ValueError("Longitude (phi) out of range." if bad & _A2V_CLASS_PHI else "Co-latitude (theta) out of range.")

        rc, bad = be.call("angle_to_vector", [be.ptr(fa), be.ptr(fb), be.ptr(vec)], [n, int(bool(lonlat)), int(bool(degrees))], on_bad)
        _lib.check_rc(rc)
        if bad >= 0:
            on_bad(bad)
    # np.array([x, y, z]).transpose(): the component axis goes last, the data axes are REVERSED
    out = vec.reshape(tuple(shape) + (3,))
    if len(shape) > 1:
        perm = tuple(range(len(shape) - 1, -1, -1)) + (len(shape),)
        out = out.permute(perm) if _d._is_torch(out) else out.transpose(perm)
    return out


def vector_to_angle(vec, lonlat=True, degrees=True):
    """(N, 3) or (3,) vectors to angles (reference: hpgeom/hpgeom.py:396-422); always returns 1-D arrays.

    One kernel (``hpgb_vector_to_angle``): ``theta = acos(z / |v|)``, ``phi = atan2(y, x) mod 2 pi`` with
    correctly rounded acos / atan2 (numpy's are SVML routines on AVX-512 hosts, within 1 ulp of these)."""
    be = _d.pick_backend(vec)
    if be.name == "numpy":
        vec = np.atleast_1d(_d._np_view(vec))
    v = be.coerce(vec, F8).reshape(-1, 3)
    n = int(be.shape(v)[0])
    a, b = be.empty((n,), F8), be.empty((n,), F8)
    if n:
        v = be.expand(v, be.shape(v))
        rc, _ = be.call("vector_to_angle", [be.ptr(v), be.ptr(a), be.ptr(b)], [n, int(bool(lonlat)), int(bool(degrees))])
        _lib.check_rc(rc)
    return a, b


_REORDER_SIZES = (1, 2, 4, 8, 16)


def reorder(map_in, ring_to_nest=True):
    """Reorder a full-sky map between RING and NEST (reference: hpgeom/hpgeom.py:425-462).

    ONE kernel (``hpgb_reorder_map``): a CTA moves a 64 x 64 block of a base face through shared memory -- one
    contiguous run on the NEST side, one run per ring on the RING side -- so the map is read once and written
    once (the reference converts 24 groups of pixel numbers and scatters with fancy indexing)."""
    is_t = _d._is_torch(map_in)
    npix = map_in.numel() if is_t else map_in.size
    nside = npixel_to_nside(npix)
    nside_to_order(nside)  # ValueError unless nside is a legal NEST resolution
    L = _lib.lib()
    if is_t and map_in.device.type == "cuda":
        torch = _d.torch
        src = map_in.contiguous()
        if src.element_size() not in _REORDER_SIZES:
            raise TypeError("reorder: unsupported element size %d" % src.element_size())
        out = torch.empty_like(src)
        with torch.cuda.device(src.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(src.device).cuda_stream)
            rc = L.hpgb_reorder_map(src.data_ptr(), out.data_ptr(), npix, src.element_size(), int(bool(ring_to_nest)), stream)
        _lib.check_rc(rc)
        return out
    src = np.ascontiguousarray(_d._np_view(map_in))
    if src.dtype.itemsize not in _REORDER_SIZES or src.dtype.hasobject:
        raise TypeError("reorder: unsupported dtype %s" % src.dtype)
    out = np.empty_like(src)
    rc = L.hpgb_host_reorder_map(src.ctypes.data, out.ctypes.data, npix, src.dtype.itemsize, int(bool(ring_to_nest)),
                                 _d.host_device())
    _lib.check_rc(rc)
    return out


def nside_to_npixel(nside):
    """Number of pixels for an nside (reference: hpgeom/hpgeom.py:190-206)."""
    _nside = np.atleast_1d(nside)
    if np.any((_nside < 0) | (_nside > max_nside)):
        raise ValueError("Illegal nside value (must be 0 <= nside <= 2**29)")
    return 12 * nside * nside


def pixel_ranges_to_pixels(pixel_ranges, inclusive=False):
    """Expand an (M, 2) array of [lo, hi) ranges (reference: hpgeom/hpgeom.c:3795-3896).

    Two passes like the reference's two loops, both on the GPU (csrc/k_ranges.cu): the count pass
    validates the rows and prefix-sums their lengths, the expand pass writes the pixel list.
    numpy in -> numpy out through the chunked D2H pipeline; a torch CUDA tensor stays resident.
    """
    be = _d.pick_backend(pixel_ranges)
    inclusive = int(bool(inclusive))
    if be.name == "numpy":
        r = be.coerce(_d._np_view(pixel_ranges), I8)
    else:
        r = be.coerce(pixel_ranges, I8)
    shape = be.shape(r)
    if len(shape) != 2 or shape[1] != 2:
        raise ValueError("pixel_ranges must be 2D, with shape (M, 2).")
    m = int(shape[0])
    if m == 0:
        return be.empty((0,), I8)
    r = be.expand(r, shape)  # C-contiguous rows = 16-byte (lo, hi) pairs
    bad_msg = "pixel_ranges[:, 0] must all be <= pixel_ranges[:, 1]"
    L = _lib.lib()
    if be.name == "numpy":
        total, first_bad = ctypes.c_int64(0), ctypes.c_int64(-1)
        _lib.check_rc(L.hpgb_host_ranges_count(be.ptr(r), m, inclusive, be.device, ctypes.byref(total),
                                               ctypes.byref(first_bad)))
        if first_bad.value >= 0:
            raise ValueError(bad_msg)
        out = be.empty((total.value,), I8)
        _lib.check_rc(L.hpgb_host_ranges_expand(be.ptr(r), m, inclusive, be.ptr(out), total.value, be.device))
        return out
    torch = _d.torch
    with torch.cuda.device(be.device):
        if r.data_ptr() % 16:
            r = r.clone()
        offs = torch.empty(m + 1, dtype=torch.int64, device=be.dev)
        def on_bad(bad):
            raise ValueError(bad_msg)

        rc, first_bad = be.call("ranges_offsets", [be.ptr(r)], [m, inclusive, be.ptr(offs)], on_bad)
        _lib.check_rc(rc)
        if first_bad >= 0:
            on_bad(first_bad)
        total = int(offs[m].item())
        out = torch.empty(total, dtype=torch.int64, device=be.dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream(be.device).cuda_stream)
        _lib.check_rc(L.hpgb_ranges_expand(be.ptr(r), be.ptr(offs), m, 0, total, be.ptr(out), stream))
    return out


def query_circle(nside, a, b, radius, inclusive=False, fact=4, nest=True, lonlat=True, degrees=True,
                 return_pixel_ranges=False, device_resident=False):
    """Pixels whose centres lie within (inclusive: which overlap) a circle
    (reference: hpgeom/hpgeom.c:765-864 -> query_disc, hpgeom/healpix_geom.c:632-747).

    RING scheme: the per-disc scalars are prepared on the host, every ring the disc touches is one
    GPU thread (csrc/k_query.cu).  NEST scheme: the reference's hierarchical search walked breadth
    first, one kernel launch per order over the whole frontier, ranges sorted on the device.  Either
    way the resulting intervals are expanded by the pixel_ranges_to_pixels kernels.  Same argument
    checks, order and messages as the reference.  ``device_resident=True`` (an extension) returns a
    torch CUDA tensor instead of a numpy array.
    """
    nside, a, b, radius = int(nside), float(a), float(b), float(radius)
    inclusive, nest, lonlat, degrees = bool(inclusive), bool(nest), bool(lonlat), bool(degrees)
    fact = int(fact)
    if return_pixel_ranges and not nest:
        raise RuntimeError("Can only use return_pixel_ranges with nest ordering.")
    L = _lib.lib()
    plan = _lib.Disc()
    rc = L.hpgb_query_circle_ring_plan(nside, a, b, radius, int(inclusive), fact, int(lonlat), int(degrees),
                                       ctypes.byref(plan))
    if rc == 6:
        raise ValueError(_lib.explain_angle(1, False, a, b, lonlat, degrees))
    if rc == 7:
        raise ValueError("Radius must be positive.")
    if rc in (1, 2, 3) or (nest and _lib.explain_nside(nside, True)):
        raise ValueError(_lib.explain_nside(nside, nest))
    if inclusive:  # hpgeom/hpgeom_utils.c:73-87
        if fact <= 0:
            raise ValueError("Inclusive factor %d must be positive." % fact)
        if fact * nside > max_nside:
            raise ValueError("Inclusive factor * nside must be <= %d" % max_nside)
        if nest and (fact & (fact - 1)) > 0:
            raise ValueError("Inclusive factor %d must be power of 2 for nest." % fact)
    _lib.check_rc(rc)
    dev = _d.default_device()
    if nest:
        return _query_circle_nest(L, nside, a, b, radius, inclusive, fact, lonlat, degrees, return_pixel_ranges,
                                  device_resident, dev)
    if not device_resident:
        total = ctypes.c_int64(0)
        _lib.check_rc(L.hpgb_host_query_circle_ring_count(ctypes.byref(plan), dev, ctypes.byref(total)))
        out = np.empty(total.value, dtype=np.int64)
        _lib.check_rc(L.hpgb_host_query_circle_ring_fill(ctypes.byref(plan), ctypes.c_void_p(out.ctypes.data),
                                                         total.value, dev))
        return out
    torch = _d.torch
    with torch.cuda.device(dev):
        m = int(plan.m)
        table = torch.empty((m, 2), dtype=torch.int64, device="cuda:%d" % dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        wb = int(L.hpgb_query_circle_ring_work_bytes(ctypes.byref(plan)))
        work = torch.empty(max(wb, 16) // 8, dtype=torch.int64, device="cuda:%d" % dev)
        _lib.check_rc(L.hpgb_query_circle_ring_ranges(ctypes.byref(plan), ctypes.c_void_p(table.data_ptr()),
                                                      ctypes.c_void_p(work.data_ptr()), stream))
    return pixel_ranges_to_pixels(table)


def query_circle_batch(nside, a, b, radius, inclusive=False, fact=4, nest=False, lonlat=True, degrees=True,
                       device_resident=False):
    """Many query_circle calls at once (an extension: the reference has no batched form).

    ``a``, ``b``, ``radius`` are 1-D arrays of D discs (``radius`` may be a scalar); returns
    ``(pixels, offsets)``: the D pixel lists back to back and the D + 1 list boundaries, list d =
    ``pixels[offsets[d]:offsets[d+1]]``, each identical to ``query_circle(nside, a[d], b[d], radius[d],
    nest=nest, ...)``.  One small disc is launch-latency bound on a GPU; a batch runs every (disc, ring)
    pair (RING) or every order of the hierarchical search over all discs (NEST) in one launch
    (csrc/k_query.cu).  Argument errors are reported for the first offending disc with the reference's
    messages.  Note the default here is RING (the cheaper scheme for small discs).
    """
    nside, fact, inclusive, lonlat, degrees = int(nside), int(fact), bool(inclusive), bool(lonlat), bool(degrees)
    if nest:
        return _query_circle_batch_nest(nside, a, b, radius, inclusive, fact, lonlat, degrees, device_resident)
    av = np.ascontiguousarray(np.atleast_1d(_d._np_view(a)), dtype=np.float64)
    bv = np.ascontiguousarray(np.atleast_1d(_d._np_view(b)), dtype=np.float64)
    if av.ndim != 1 or av.shape != bv.shape:
        raise ValueError("a and b must be 1-D arrays of the same length")
    rv = np.ascontiguousarray(np.broadcast_to(np.asarray(_d._np_view(radius), dtype=np.float64), av.shape))
    D = int(av.size)
    L = _lib.lib()
    plans = (_lib.Disc * max(D, 1))()
    bad = ctypes.c_int64(-1)
    rc = L.hpgb_query_circle_ring_plan_many(nside, av.ctypes.data, bv.ctypes.data, rv.ctypes.data, D, int(inclusive), fact,
                                            int(lonlat), int(degrees), plans, ctypes.byref(bad))
    if rc != 0 and bad.value >= 0:
        k = bad.value  # same checks, order and messages as the single call
        query_circle(nside, float(av[k]), float(bv[k]), float(rv[k]), inclusive=inclusive, fact=fact, nest=False,
                     lonlat=lonlat, degrees=degrees)
    _lib.check_rc(rc)
    dev = _d.default_device()
    if not device_resident:
        offsets = np.zeros(D + 1, dtype=np.int64)
        total = ctypes.c_int64(0)
        _lib.check_rc(L.hpgb_host_query_circle_ring_batch(plans, D, dev, offsets.ctypes.data, ctypes.byref(total), None, 0))
        out = np.empty(total.value, dtype=np.int64)
        if total.value:
            _lib.check_rc(L.hpgb_host_query_circle_ring_batch(plans, D, dev, offsets.ctypes.data, ctypes.byref(total),
                                                              out.ctypes.data, total.value))
        return out, offsets
    torch = _d.torch
    ring_start = np.zeros(D + 1, dtype=np.int64)
    nring = int(L.hpgb_query_circle_ring_batch_prepare(plans, D, ring_start.ctypes.data))
    m = 2 * nring + 2 * D
    with torch.cuda.device(dev):
        cuda = "cuda:%d" % dev
        d_plans = torch.from_numpy(np.frombuffer(plans, dtype=np.uint8, count=D * ctypes.sizeof(_lib.Disc)).copy()).to(cuda)
        d_rs = torch.from_numpy(ring_start).to(cuda)
        table = torch.empty((max(m, 1), 2), dtype=torch.int64, device=cuda)
        fct = fact if inclusive else 1
        work = torch.empty(8 * (nring + 1) if fct > 1 else 2, dtype=torch.int64, device=cuda)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if D:
            _lib.check_rc(L.hpgb_query_circle_ring_batch_ranges(d_plans.data_ptr(), d_rs.data_ptr(), D, nring, nside, fct,
                                                                table.data_ptr(), work.data_ptr(), stream))
        table = table[:m]
        pixels = pixel_ranges_to_pixels(table) if m else torch.empty(0, dtype=torch.int64, device=cuda)
        lens = (table[:, 1] - table[:, 0]).clamp_(min=0)
        csum = torch.cat([torch.zeros(1, dtype=torch.int64, device=cuda), torch.cumsum(lens, 0)])
        offsets = csum[2 * d_rs + 2 * torch.arange(D + 1, device=cuda)]
    return pixels, offsets


def _query_circle_batch_nest(nside, a, b, radius, inclusive, fact, lonlat, degrees, device_resident):
    av = np.ascontiguousarray(np.atleast_1d(_d._np_view(a)), dtype=np.float64)
    bv = np.ascontiguousarray(np.atleast_1d(_d._np_view(b)), dtype=np.float64)
    if av.ndim != 1 or av.shape != bv.shape:
        raise ValueError("a and b must be 1-D arrays of the same length")
    rv = np.ascontiguousarray(np.broadcast_to(np.asarray(_d._np_view(radius), dtype=np.float64), av.shape))
    D = int(av.size)
    L = _lib.lib()
    dev = _d.default_device()
    handle, m, total, bad = ctypes.c_void_p(), ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(-1)
    rc = L.hpgb_query_circle_nest_batch_open(nside, av.ctypes.data, bv.ctypes.data, rv.ctypes.data, D, int(inclusive), fact,
                                             int(lonlat), int(degrees), dev, ctypes.byref(handle), ctypes.byref(m),
                                             ctypes.byref(total), ctypes.byref(bad))
    if rc != 0 and bad.value >= 0:
        k = bad.value  # same checks, order and messages as the single call
        query_circle(nside, float(av[k]), float(bv[k]), float(rv[k]), inclusive=inclusive, fact=fact, nest=True,
                     lonlat=lonlat, degrees=degrees)
    _lib.check_rc(rc)
    try:
        offsets = np.zeros(D + 1, dtype=np.int64)
        _lib.check_rc(L.hpgb_query_nest_batch_offsets(handle, offsets.ctypes.data))
        if not device_resident:
            out = np.empty(total.value, dtype=np.int64)
            _lib.check_rc(L.hpgb_query_nest_pixels_host(handle, ctypes.c_void_p(out.ctypes.data), total.value))
            return out, offsets
        torch = _d.torch
        with torch.cuda.device(dev):
            out = torch.empty(total.value, dtype=torch.int64, device="cuda:%d" % dev)
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check_rc(L.hpgb_query_nest_pixels_device(handle, ctypes.c_void_p(out.data_ptr()), total.value, stream))
            torch.cuda.current_stream(dev).synchronize()
            return out, torch.from_numpy(offsets).to("cuda:%d" % dev)
    finally:
        L.hpgb_query_nest_close(handle)


def _query_circle_nest(L, nside, a, b, radius, inclusive, fact, lonlat, degrees, return_pixel_ranges, device_resident, dev):
    handle, m, total = ctypes.c_void_p(), ctypes.c_int64(0), ctypes.c_int64(0)
    _lib.check_rc(L.hpgb_query_circle_nest_open(nside, a, b, radius, int(inclusive), fact, int(lonlat), int(degrees), dev,
                                                ctypes.byref(handle), ctypes.byref(m), ctypes.byref(total)))
    try:
        if return_pixel_ranges:
            table = np.empty((m.value, 2), dtype=np.int64)
            k = ctypes.c_int64(0)
            _lib.check_rc(L.hpgb_query_nest_ranges(handle, ctypes.c_void_p(table.ctypes.data), ctypes.byref(k)))
            table = np.ascontiguousarray(table[:k.value])
            return _d.torch.from_numpy(table).to("cuda:%d" % dev) if device_resident else table
        if not device_resident:
            out = np.empty(total.value, dtype=np.int64)
            _lib.check_rc(L.hpgb_query_nest_pixels_host(handle, ctypes.c_void_p(out.ctypes.data), total.value))
            return out
        torch = _d.torch
        with torch.cuda.device(dev):
            out = torch.empty(total.value, dtype=torch.int64, device="cuda:%d" % dev)
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check_rc(L.hpgb_query_nest_pixels_device(handle, ctypes.c_void_p(out.data_ptr()), total.value, stream))
            torch.cuda.current_stream(dev).synchronize()  # the handle (range table) is freed below
        return out
    finally:
        L.hpgb_query_nest_close(handle)


def query_circle_vec(nside, vec, radius, inclusive=False, fact=4, nest=True):
    """query_circle for a centre given as a unit vector, radius in radians
    (reference: hpgeom/hpgeom.py:115-160)."""
    if len(vec) != 3:
        raise ValueError("vec must be 3 elements.")
    theta, phi = vector_to_angle(vec, lonlat=False)
    return query_circle(nside, theta[0], phi[0], radius, inclusive=inclusive, fact=fact, nest=nest, lonlat=False)


def _upgrade_scalar_checks(nside, nside_upgrade):
    # hpgeom/hpgeom.py:490-499, same order and messages
    if not isinstance(nside, numbers.Integral) or not isinstance(nside_upgrade, numbers.Integral):
        return "Input nside and nside_upgrade must be integers."
    if nside_upgrade < nside:
        return "The value for nside_upgrade must be >= nside."
    if (nside & (nside - 1)) > 0:
        return "Input nside must be a power of two."
    if (nside_upgrade & (nside_upgrade - 1)) > 0:
        return "Output nside_upgrade must be a power of two."
    return None


def upgrade_pixel_ranges(nside, pixel_ranges, nside_upgrade):
    """Upgrade nest pixel ranges to a finer nside (reference: hpgeom/hpgeom.py:465-512)."""
    msg = _upgrade_scalar_checks(nside, nside_upgrade)
    if msg:
        raise ValueError(msg)
    pixel_ranges = np.asarray(_d._np_view(pixel_ranges))
    _pixel_ranges = np.atleast_2d(pixel_ranges).astype(np.int64)
    if _pixel_ranges.ndim != 2 or _pixel_ranges.shape[1] != 2:
        raise ValueError("Input pixel_ranges must be of shape (N, 2).")
    if np.min(pixel_ranges.ravel()) < 0 or np.max(pixel_ranges.ravel()) >= nside_to_npixel(nside):
        raise ValueError("Input pixels/ranges out of range.")
    bit_shift = 2 * int(np.round(np.log2(nside_upgrade / nside)))
    _pixel_ranges <<= bit_shift
    return _pixel_ranges


def upgrade_pixels(nside, pixels, nside_upgrade, nest=True):
    """All pixels of ``nside_upgrade`` covering the input pixels (reference: hpgeom/hpgeom.py:515-556).

    One fused kernel replaces the reference's ring_to_nest -> range build -> shift ->
    pixel_ranges_to_pixels -> nest_to_ring chain; the output order (children in nest order,
    unsorted for ring) and the reference's checks -- including its refusal of the last pixel,
    whose range end ``p + 1`` equals npix (hpgeom/hpgeom.py:505) -- are preserved.
    """
    nest = bool(nest)
    be = _d.pick_backend(pixels)
    if be.name == "numpy":
        pix = np.atleast_1d(_d._np_view(pixels)).astype(np.int64)  # unsafe cast allowed, as the reference
    else:
        pix = pixels.reshape(-1) if pixels.dim() == 0 else pixels
        pix = pix.to(dtype=_d.torch.int64)
    scalar_msg = _upgrade_scalar_checks(nside, nside_upgrade)
    if not nest:
        # the reference calls ring_to_nest first: its nside / pixel errors win
        if scalar_msg is not None:
            ring_to_nest(nside, pix)
    if scalar_msg is not None:
        raise ValueError(scalar_msg)
    if len(be.shape(pix)) != 1:
        raise ValueError("Input pixel_ranges must be of shape (N, 2).")
    n = int(be.shape(pix)[0])
    if n == 0:
        # ring_to_nest of nothing checks nothing; np.min of an empty array raises in the reference
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    ns = int(nside)
    if not nest and (ns <= 0 or ns > max_nside):
        raise ValueError(_lib.explain_nside(ns, True))  # the reference's ring_to_nest call comes first
    nside_to_npixel(ns)  # "Illegal nside value ..." for nside < 0 or > 2**29 (hpgeom/hpgeom.py:505 -> :190-206)
    if ns == 0:
        raise ValueError("Input pixels/ranges out of range.")  # npix = 0: every pixel is out of range
    ratio = (int(nside_upgrade) // ns) ** 2
    out = be.empty((n * ratio,), I8)
    pix = be.expand(pix, be.shape(pix))
    def on_bad(bad):
        # key = (class << 62) | index; class 0 = pixel out of range, class 1 = the last nest pixel
        cls, idx = bad >> 62, bad & ((1 << 62) - 1)
        if not nest and cls == 0:
            raise ValueError(_lib.explain_pixel(int(nside), True, int(be.item(pix, idx))))
        raise ValueError("Input pixels/ranges out of range.")

    rc, bad = be.call("upgrade_pixels", [be.ptr(pix), be.ptr(out)], [n, int(nside), int(nside_upgrade), int(nest)], on_bad)
    if rc in (1, 2, 3):
        raise ValueError(_lib.explain_nside(int(nside), True))
    _lib.check_rc(rc)
    if bad >= 0:
        on_bad(bad)
    return out


def _out_of_scope(name):
    def stub(*args, **kwargs):
        raise NotImplementedError(
            "hpgeom_b200.%s: the polygon / ellipse / box region queries are outside the accelerated path "
            "(SURVEY.md section 2 row 6; README 'Scope').  query_circle, query_circle_vec and query_circle_batch "
            "are available; use the reference package for this call." % name)
    stub.__name__ = name
    stub.__doc__ = "Not implemented (out of scope); raises NotImplementedError.  Reference: hpgeom/hpgeom.c:866-1383."
    return stub


query_polygon = _out_of_scope("query_polygon")
query_polygon_vec = _out_of_scope("query_polygon_vec")
query_ellipse = _out_of_scope("query_ellipse")
query_box = _out_of_scope("query_box")
