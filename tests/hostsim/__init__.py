"""TEST INFRASTRUCTURE ONLY: ctypes front end of libhostsim.so (host build of the device element
functions, see hostsim.cpp).  Mirrors the reference's Python signatures for 1-D inputs."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_I = ctypes.POINTER(ctypes.c_int64)
_D = ctypes.POINTER(ctypes.c_double)


class HostSim:
    def __init__(self):
        subprocess.check_call(["make", "-s", "-C", _HERE])
        # HPGB_HOSTSIM_LIB: an alternative host build (kernel variants compiled with extra -D flags)
        self.lib = ctypes.CDLL(os.environ.get("HPGB_HOSTSIM_LIB") or os.path.join(_HERE, "libhostsim.so"))
        for n in ("sim_nest_to_ring", "sim_ring_to_nest", "sim_ring_roundtrip_any", "sim_angle_to_pixel",
                  "sim_pixel_to_angle", "sim_pixel_to_vector", "sim_vector_to_pixel", "sim_neighbors", "sim_interp",
                  "sim_lonlat", "sim_nest_to_ring_k", "sim_ring_to_nest_k", "sim_ring_to_nest_fold", "sim_pixel_to_angle_k", "sim_pixel_to_vector_k", "sim_pixel_to_angle_tab", "sim_pixel_to_vector_tab", "sim_interp_k", "sim_query_circle_ring", "sim_query_circle_nest"):
            getattr(self.lib, n).restype = ctypes.c_int64
        self.last_slow = 0

    @staticmethod
    def _i(a):
        return np.ascontiguousarray(np.atleast_1d(a), dtype=np.int64)

    @staticmethod
    def _f(a):
        return np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64)

    @staticmethod
    def _chk(bad):
        if bad >= 0:
            raise ValueError("bad element %d" % bad)

    def _conv(self, fn, nside, pix):
        pix = self._i(pix)
        out = np.empty_like(pix)
        self._chk(getattr(self.lib, fn)(ctypes.c_int64(nside), pix.ctypes.data_as(_I), ctypes.c_int64(pix.size),
                                        out.ctypes.data_as(_I)))
        return out

    def nest_to_ring(self, nside, pix):
        return self._conv("sim_nest_to_ring", nside, pix)

    def ring_to_nest(self, nside, pix):
        return self._conv("sim_ring_to_nest", nside, pix)

    def _conv_k(self, fn, nside, pix, hi):
        pix = self._i(pix)
        out = np.empty_like(pix)
        self._chk(getattr(self.lib, fn)(ctypes.c_int64(nside), pix.ctypes.data_as(_I), ctypes.c_int64(pix.size),
                                        out.ctypes.data_as(_I), int(hi)))
        return out

    def nest_to_ring_k(self, nside, pix, hi=None):
        """hpgb_reorder.h core; hi=None picks the instantiation the kernel launcher would."""
        return self._conv_k("sim_nest_to_ring_k", nside, pix, nside >= 2 ** 16 if hi is None else hi)

    def ring_to_nest_k(self, nside, pix, hi=None):
        return self._conv_k("sim_ring_to_nest_k", nside, pix, nside >= 2 ** 16 if hi is None else hi)

    def ring_to_nest_fold(self, nside, pix, hi=None):
        """the folded-coordinate path of the streaming ring -> nest kernel (hpgb_reorder.h)"""
        return self._conv_k("sim_ring_to_nest_fold", nside, pix, nside >= 2 ** 16 if hi is None else hi)

    def ring_roundtrip_any(self, nside, pix):
        return self._conv("sim_ring_roundtrip_any", nside, pix)

    def angle_to_pixel(self, nside, a, b, nest=True, lonlat=True, degrees=True, exact_only=False):
        a, b = self._f(a), self._f(b)
        out = np.empty(a.size, dtype=np.int64)
        ns = ctypes.c_int64(0)
        self._chk(self.lib.sim_angle_to_pixel(ctypes.c_int64(nside), a.ctypes.data_as(_D), b.ctypes.data_as(_D),
                                              ctypes.c_int64(a.size), int(nest), int(lonlat), int(degrees),
                                              int(exact_only), out.ctypes.data_as(_I), ctypes.byref(ns)))
        self.last_slow = ns.value
        return out

    def pixel_to_angle(self, nside, pix, nest=True, lonlat=True, degrees=True):
        pix = self._i(pix)
        a, b = np.empty(pix.size), np.empty(pix.size)
        self._chk(self.lib.sim_pixel_to_angle(ctypes.c_int64(nside), pix.ctypes.data_as(_I), ctypes.c_int64(pix.size),
                                              int(nest), int(lonlat), int(degrees), a.ctypes.data_as(_D),
                                              b.ctypes.data_as(_D)))
        return a, b

    def pixel_to_angle_k(self, nside, pix, nest=True, lonlat=True, degrees=True):
        """The branch-free power-of-two path of the kernels; self.last_slow = pixels that took a rare-case patch."""
        pix = self._i(pix)
        a, b = np.empty(pix.size), np.empty(pix.size)
        r = self.lib.sim_pixel_to_angle_k(ctypes.c_int64(nside), pix.ctypes.data_as(_I), ctypes.c_int64(pix.size),
                                          int(nest), int(lonlat), int(degrees), a.ctypes.data_as(_D), b.ctypes.data_as(_D))
        if r < 0:
            raise ValueError("bad element %d" % (-1 - r))
        self.last_slow = r
        return a, b

    def pixel_to_vector_k(self, nside, pix, nest=True):
        pix = self._i(pix)
        x, y, z = np.empty(pix.size), np.empty(pix.size), np.empty(pix.size)
        r = self.lib.sim_pixel_to_vector_k(ctypes.c_int64(nside), pix.ctypes.data_as(_I), ctypes.c_int64(pix.size), int(nest),
                                           x.ctypes.data_as(_D), y.ctypes.data_as(_D), z.ctypes.data_as(_D))
        if r < 0:
            raise ValueError("bad element %d" % (-1 - r))
        self.last_slow = r
        return x, y, z

    def pixel_to_angle_tab(self, nside, pix, nest=True, lonlat=True, degrees=True):
        """The ring-table form of the kernels (per-ring colatitude table + per-pixel ring number and phi)."""
        pix = self._i(pix)
        a, b = np.empty(pix.size), np.empty(pix.size)
        r = self.lib.sim_pixel_to_angle_tab(ctypes.c_int64(nside), pix.ctypes.data_as(_I), ctypes.c_int64(pix.size),
                                            int(nest), int(lonlat), int(degrees), a.ctypes.data_as(_D), b.ctypes.data_as(_D))
        if r < 0:
            raise ValueError("bad element %d" % (-1 - r))
        return a, b

    def pixel_to_vector_tab(self, nside, pix, nest=True, sincos=0):
        pix = self._i(pix)
        x, y, z = np.empty(pix.size), np.empty(pix.size), np.empty(pix.size)
        r = self.lib.sim_pixel_to_vector_tab(ctypes.c_int64(nside), pix.ctypes.data_as(_I), ctypes.c_int64(pix.size), int(nest),
                                             int(sincos), x.ctypes.data_as(_D), y.ctypes.data_as(_D), z.ctypes.data_as(_D))
        if r < 0:
            raise ValueError("bad element %d" % (-1 - r))
        return x, y, z

    def pixel_to_vector(self, nside, pix, nest=True):
        pix = self._i(pix)
        x, y, z = np.empty(pix.size), np.empty(pix.size), np.empty(pix.size)
        self._chk(self.lib.sim_pixel_to_vector(ctypes.c_int64(nside), pix.ctypes.data_as(_I), ctypes.c_int64(pix.size),
                                               int(nest), x.ctypes.data_as(_D), y.ctypes.data_as(_D),
                                               z.ctypes.data_as(_D)))
        return x, y, z

    def vector_to_pixel(self, nside, x, y, z, nest=True, exact_only=False):
        x, y, z = self._f(x), self._f(y), self._f(z)
        out = np.empty(x.size, dtype=np.int64)
        self.last_slow = self.lib.sim_vector_to_pixel(ctypes.c_int64(nside), x.ctypes.data_as(_D), y.ctypes.data_as(_D),
                                                      z.ctypes.data_as(_D), ctypes.c_int64(x.size), int(nest),
                                                      int(exact_only), out.ctypes.data_as(_I))
        if self.last_slow < 0:
            raise ValueError("folded-face and direct-face fast paths disagree at element %d" % (-self.last_slow - 4000000000000))
        return out

    def neighbors(self, nside, pix, nest=True):
        pix = self._i(pix)
        out = np.empty((pix.size, 8), dtype=np.int64)
        self._chk(self.lib.sim_neighbors(ctypes.c_int64(nside), pix.ctypes.data_as(_I), ctypes.c_int64(pix.size),
                                         int(nest), out.ctypes.data_as(_I)))
        return out

    def get_interpolation_weights(self, nside, a, b, nest=True, lonlat=True, degrees=True):
        a, b = self._f(a), self._f(b)
        p = np.empty((a.size, 4), dtype=np.int64)
        w = np.empty((a.size, 4), dtype=np.float64)
        self._chk(self.lib.sim_interp(ctypes.c_int64(nside), a.ctypes.data_as(_D), b.ctypes.data_as(_D),
                                      ctypes.c_int64(a.size), int(nest), int(lonlat), int(degrees),
                                      p.ctypes.data_as(_I), w.ctypes.data_as(_D)))
        return p, w

    def get_interpolation_weights_k(self, nside, a, b, nest=True, lonlat=True, degrees=True, exact_only=False):
        """Power-of-two kernel path (fast ring choice + tables); self.last_slow = points that fell back.
        exact_only = 2: the zone-sorted bodies of the kernel (hpgb_interp_belt / hpgb_interp_cap)."""
        a, b = self._f(a), self._f(b)
        p = np.empty((a.size, 4), dtype=np.int64)
        w = np.empty((a.size, 4), dtype=np.float64)
        ns = ctypes.c_int64(0)
        self._chk(self.lib.sim_interp_k(ctypes.c_int64(nside), a.ctypes.data_as(_D), b.ctypes.data_as(_D),
                                        ctypes.c_int64(a.size), int(nest), int(lonlat), int(degrees), int(exact_only),
                                        p.ctypes.data_as(_I), w.ctypes.data_as(_D), ctypes.byref(ns)))
        self.last_slow = ns.value
        return p, w

    def div_const_mismatches(self, a, d):
        """hpgb_div_const(a, d) != a / d, counted over the array a"""
        a = self._f(a)
        self.lib.sim_div_const_check.restype = ctypes.c_int64
        return self.lib.sim_div_const_check(a.ctypes.data_as(_D), ctypes.c_int64(a.size), ctypes.c_double(d))

    def lonlat_to_thetaphi(self, lon, lat, degrees=True):
        lon, lat = self._f(lon), self._f(lat)
        th, ph = np.empty(lon.size), np.empty(lon.size)
        bad = self.lib.sim_lonlat(lon.ctypes.data_as(_D), lat.ctypes.data_as(_D), ctypes.c_int64(lon.size),
                                  int(degrees), th.ctypes.data_as(_D), ph.ctypes.data_as(_D))
        if bad >= 0:
            raise ValueError("Latitude out of range.")
        return th, ph

    def query_circle_ring_ranges(self, nside, theta, phi, radius, fact=0):
        """(m, 2) range table of the RING query_circle kernel path (theta/phi/radius in radians,
        fact=0: non-inclusive)."""
        args = (ctypes.c_int64(nside), ctypes.c_double(theta), ctypes.c_double(phi), ctypes.c_double(radius),
                ctypes.c_int64(fact))
        m = self.lib.sim_query_circle_ring(*args, None, ctypes.c_int64(0))
        if m < 0:
            raise ValueError("plan failed: %d" % -m)
        r = np.zeros((m, 2), dtype=np.int64)
        self.lib.sim_query_circle_ring(*args, r.ctypes.data_as(_I), ctypes.c_int64(m))
        return r

    def query_circle_nest_ranges(self, nside, theta, phi, radius, fact=0):
        """merged (m, 2) range list of the NEST query_circle walk (radians; fact=0: non-inclusive)."""
        args = (ctypes.c_int64(nside), ctypes.c_double(theta), ctypes.c_double(phi), ctypes.c_double(radius),
                ctypes.c_int64(fact))
        m = self.lib.sim_query_circle_nest(*args, None, ctypes.c_int64(0))
        if m < 0:
            raise ValueError("plan failed: %d" % -m)
        r = np.zeros((m, 2), dtype=np.int64)
        self.lib.sim_query_circle_nest(*args, r.ctypes.data_as(_I), ctypes.c_int64(m))
        return r
