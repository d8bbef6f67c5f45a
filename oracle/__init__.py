"""oracle -- TEST INFRASTRUCTURE ONLY (see oracle/hpg_oracle.c header).

Two checkers live here:

* ``oracle.port``  : ctypes binding of ``liboracle.so`` (our plain-C restatement, hpg_oracle.c),
  wrapped in functions that mirror the reference's Python signatures for the hot path.
* ``oracle.ref``   : the UNMODIFIED reference C extension compiled from /root/reference into
  ``oracle/_ref/_hpgeom*.so`` by ``oracle/Makefile`` (``None`` when that build is absent).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  hpgeom_b200 never does.
"""
import ctypes
import glob
import importlib.machinery
import importlib.util
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(os.path.join(_HERE, "liboracle.so")) \
            or os.path.getmtime(os.path.join(_HERE, "liboracle.so")) < os.path.getmtime(os.path.join(_HERE, "hpg_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    if not glob.glob(os.path.join(_HERE, "_ref", "_hpgeom*.so")) or force:
        subprocess.call(["make", "-s", "-C", _HERE, "ref"])


def _load_ref():
    cands = glob.glob(os.path.join(_HERE, "_ref", "_hpgeom*.so"))
    if not cands:
        return None
    loader = importlib.machinery.ExtensionFileLoader("_hpgeom", cands[0])
    spec = importlib.util.spec_from_loader("_hpgeom", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


class _Port:
    """Reference-shaped Python front end over liboracle.so (scalar nside, 1-D data)."""

    _I64P = ctypes.POINTER(ctypes.c_int64)
    _F64P = ctypes.POINTER(ctypes.c_double)

    def __init__(self):
        build()
        self.lib = ctypes.CDLL(os.path.join(_HERE, "liboracle.so"))
        for name in ("orc_angle_to_pixel", "orc_pixel_to_angle", "orc_nest_to_ring", "orc_ring_to_nest",
                     "orc_vector_to_pixel", "orc_pixel_to_vector", "orc_neighbors", "orc_interp_weights",
                     "orc_pixel_ranges_to_pixels", "orc_upgrade_pixels"):
            getattr(self.lib, name).restype = ctypes.c_int64
        self.lib.orc_check_nside.restype = ctypes.c_int

    @staticmethod
    def _i(a):
        return np.ascontiguousarray(np.atleast_1d(a), dtype=np.int64)

    @staticmethod
    def _f(a):
        return np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64)

    def _pi(self, a):
        return a.ctypes.data_as(self._I64P)

    def _pf(self, a):
        return a.ctypes.data_as(self._F64P)

    def _nside(self, nside, nest):
        c = self.lib.orc_check_nside(ctypes.c_int64(int(nside)), int(bool(nest)))
        if c:
            raise ValueError("bad nside (code %d)" % c)
        return ctypes.c_int64(int(nside))

    @staticmethod
    def _done(bad, what):
        if bad >= 0:
            raise ValueError("%s out of range at element %d" % (what, bad))

    def angle_to_pixel(self, nside, a, b, nest=True, lonlat=True, degrees=True):
        a, b = self._f(a), self._f(b)
        out = np.empty(a.size, dtype=np.int64)
        bad = self.lib.orc_angle_to_pixel(self._nside(nside, nest), self._pf(a), self._pf(b),
                                          ctypes.c_int64(a.size), int(nest), int(lonlat), int(degrees),
                                          self._pi(out))
        self._done(bad, "angle")
        return out

    def pixel_to_angle(self, nside, pix, nest=True, lonlat=True, degrees=True):
        pix = self._i(pix)
        a = np.empty(pix.size)
        b = np.empty(pix.size)
        bad = self.lib.orc_pixel_to_angle(self._nside(nside, nest), self._pi(pix), ctypes.c_int64(pix.size),
                                          int(nest), int(lonlat), int(degrees), self._pf(a), self._pf(b))
        self._done(bad, "pixel")
        return a, b

    def nest_to_ring(self, nside, pix):
        pix = self._i(pix)
        out = np.empty_like(pix)
        self._done(self.lib.orc_nest_to_ring(self._nside(nside, True), self._pi(pix),
                                             ctypes.c_int64(pix.size), self._pi(out)), "pixel")
        return out

    def ring_to_nest(self, nside, pix):
        pix = self._i(pix)
        out = np.empty_like(pix)
        self._done(self.lib.orc_ring_to_nest(self._nside(nside, True), self._pi(pix),
                                             ctypes.c_int64(pix.size), self._pi(out)), "pixel")
        return out

    def vector_to_pixel(self, nside, x, y, z, nest=True):
        x, y, z = self._f(x), self._f(y), self._f(z)
        out = np.empty(x.size, dtype=np.int64)
        self.lib.orc_vector_to_pixel(self._nside(nside, nest), self._pf(x), self._pf(y), self._pf(z),
                                     ctypes.c_int64(x.size), int(nest), self._pi(out))
        return out

    def pixel_to_vector(self, nside, pix, nest=True):
        pix = self._i(pix)
        x, y, z = np.empty(pix.size), np.empty(pix.size), np.empty(pix.size)
        self._done(self.lib.orc_pixel_to_vector(self._nside(nside, nest), self._pi(pix), ctypes.c_int64(pix.size),
                                                int(nest), self._pf(x), self._pf(y), self._pf(z)), "pixel")
        return x, y, z

    def neighbors(self, nside, pix, nest=True):
        pix = self._i(pix)
        out = np.empty((pix.size, 8), dtype=np.int64)
        self._done(self.lib.orc_neighbors(self._nside(nside, nest), self._pi(pix), ctypes.c_int64(pix.size),
                                          int(nest), self._pi(out)), "pixel")
        return out

    def get_interpolation_weights(self, nside, a, b, nest=True, lonlat=True, degrees=True):
        a, b = self._f(a), self._f(b)
        p = np.empty((a.size, 4), dtype=np.int64)
        w = np.empty((a.size, 4), dtype=np.float64)
        self._done(self.lib.orc_interp_weights(self._nside(nside, nest), self._pf(a), self._pf(b),
                                               ctypes.c_int64(a.size), int(nest), int(lonlat), int(degrees),
                                               self._pi(p), self._pf(w)), "angle")
        return p, w

    def pixel_ranges_to_pixels(self, ranges, inclusive=False):
        r = np.ascontiguousarray(ranges, dtype=np.int64).reshape(-1, 2)
        total = int(np.sum(r[:, 1] - r[:, 0] + int(inclusive))) if r.size else 0
        out = np.empty(max(total, 0), dtype=np.int64)
        k = self.lib.orc_pixel_ranges_to_pixels(self._pi(r), ctypes.c_int64(r.shape[0]), int(inclusive),
                                                self._pi(out))
        if k < 0:
            raise ValueError("pixel_ranges[:, 0] must all be <= pixel_ranges[:, 1]")
        return out[:k]

    def upgrade_pixels(self, nside, pixels, nside_upgrade, nest=True):
        """hpgeom/hpgeom.py:515-556 incl. the checks of upgrade_pixel_ranges (:490-506)."""
        pix = self._i(pixels)
        if nside_upgrade < nside:
            raise ValueError("The value for nside_upgrade must be >= nside.")
        if (nside & (nside - 1)) > 0:
            raise ValueError("Input nside must be a power of two.")
        if (nside_upgrade & (nside_upgrade - 1)) > 0:
            raise ValueError("Output nside_upgrade must be a power of two.")
        p = self.ring_to_nest(nside, pix) if not nest else pix
        # the reference range-checks the (p, p+1) ranges: max(p+1) >= npix raises (hpgeom.py:505)
        if p.size and (p.min() < 0 or p.max() + 1 >= 12 * nside * nside):
            raise ValueError("Input pixels/ranges out of range.")
        ratio = (nside_upgrade // nside) ** 2
        out = np.empty(pix.size * ratio, dtype=np.int64)
        self._done(self.lib.orc_upgrade_pixels(ctypes.c_int64(nside), self._pi(pix), ctypes.c_int64(pix.size),
                                               ctypes.c_int64(nside_upgrade), int(nest), self._pi(out)), "pixel")
        return out

    @staticmethod
    def lonlat_to_thetaphi(lon, lat, degrees=True):
        """hpgeom/hpgeom.py:60-88 (numpy formula: x * (pi/180), Python-style modulo)."""
        lon = np.asarray(lon, dtype=np.float64)
        lat = np.asarray(lat, dtype=np.float64)
        if degrees:
            lon_ = np.deg2rad(lon % 360.)
            lat_ = np.deg2rad(lat)
        else:
            lon_ = lon % (2. * np.pi)
            lat_ = lat
        if np.any((lat_ < -np.pi / 2.) | (lat_ > np.pi / 2.)):
            raise ValueError("Latitude out of range.")
        return np.pi / 2. - lat_, lon_


ref = _load_ref()
_port = None


def port():
    global _port
    if _port is None:
        _port = _Port()
    return _port
