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
import numbers

import numpy as np

from . import _dispatch as _d
from . import _lib
from ._dispatch import F8, I8, set_default_device  # noqa: F401

__all__ = [
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
    'upgrade_pixels',
    'upgrade_pixel_ranges',
    'nside_to_npixel',
    'set_default_device',
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
        rc, bad = be.call("lonlat_to_thetaphi", [be.ptr(lo), be.ptr(la), be.ptr(theta), be.ptr(phi)], [n, deg])
        _lib.check_rc(rc)
        if bad >= 0:
            raise ValueError("Latitude out of range.")
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


def nside_to_npixel(nside):
    """Number of pixels for an nside (reference: hpgeom/hpgeom.py:190-206)."""
    _nside = np.atleast_1d(nside)
    if np.any((_nside < 0) | (_nside > max_nside)):
        raise ValueError("Illegal nside value (must be 0 <= nside <= 2**29)")
    return 12 * nside * nside


def pixel_ranges_to_pixels(pixel_ranges, inclusive=False):
    """Expand an (M, 2) array of [lo, hi) ranges (reference: hpgeom/hpgeom.c:3795-3896).

    Host-side helper kept for API completeness (variable-length expansion is not on the
    accelerated path; upgrade_pixels below fuses its equal-length case into one kernel)."""
    r = np.asarray(_d._np_view(pixel_ranges))
    if r.dtype != np.int64:
        r = r.astype(np.int64, casting="safe")
    if r.ndim != 2 or r.shape[1] != 2:
        raise ValueError("pixel_ranges must be 2D, with shape (M, 2).")
    if r.size == 0:
        return np.zeros(0, dtype=np.int64)
    if np.any(r[:, 1] < r[:, 0]):
        raise ValueError("pixel_ranges[:, 0] must all be <= pixel_ranges[:, 1]")
    lens = r[:, 1] - r[:, 0] + int(bool(inclusive))
    starts = np.repeat(r[:, 0] - np.concatenate(([0], np.cumsum(lens)[:-1])), lens)
    return starts + np.arange(int(lens.sum()), dtype=np.int64)


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
    ratio = (int(nside_upgrade) // int(nside)) ** 2
    out = be.empty((n * ratio,), I8)
    if n == 0:
        if not nest:
            ring_to_nest(nside, pix)
        # np.min of an empty array raises in the reference
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    pix = be.expand(pix, be.shape(pix))
    rc, bad = be.call("upgrade_pixels", [be.ptr(pix), be.ptr(out)], [n, int(nside), int(nside_upgrade), int(nest)])
    if rc in (1, 2, 3):
        raise ValueError(_lib.explain_nside(int(nside), True))
    _lib.check_rc(rc)
    if bad >= 0:
        # key = (class << 62) | index; class 0 = pixel out of range, class 1 = the last nest pixel
        cls, idx = bad >> 62, bad & ((1 << 62) - 1)
        if not nest and cls == 0:
            raise ValueError(_lib.explain_pixel(int(nside), True, int(be.item(pix, idx))))
        raise ValueError("Input pixels/ranges out of range.")
    return out
