"""ctypes binding of libhpgb.so (the C ABI declared in include/hpgb.h).

There is deliberately NO fallback: if the CUDA library is missing or no CUDA device is
visible, every conversion raises.  Nothing under oracle/ is ever imported from here.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# HPGEOM_B200_LIB selects an alternative build of the same library (A/B timing of kernel variants)
LIB_PATH = os.environ.get("HPGEOM_B200_LIB") or os.path.join(_HERE, "libhpgb.so")

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_dbl = ctypes.c_double
c_vp = ctypes.c_void_p

NO_BAD_ELEMENT = 0xFFFFFFFFFFFFFFFF
E_CUDA = 1000

# name -> argtypes (restype is int unless listed in _RESTYPES)
_SIGS = {
    "hpgb_version": [],
    "hpgb_device_count": [],
    "hpgb_error_string": [c_int],
    "hpgb_status_reset": [c_vp, c_vp],
    "hpgb_launch_count": [],
    "hpgb_host_release": [],
    "hpgb_host_set_devices": [ctypes.POINTER(c_int), c_int],
    "hpgb_host_get_devices": [ctypes.POINTER(c_int), c_int],
    "hpgb_pinned_alloc": [ctypes.POINTER(c_vp), c_i64],
    "hpgb_pinned_free": [c_vp],
    "hpgb_host_copy": [c_vp, c_vp, c_i64],
    "hpgb_nest_to_ring": [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp],
    "hpgb_ring_to_nest": [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp],
    "hpgb_host_nest_to_ring": [c_vp, c_vp, c_i64, c_i64, c_vp, c_int, ctypes.POINTER(c_i64)],
    "hpgb_host_ring_to_nest": [c_vp, c_vp, c_i64, c_i64, c_vp, c_int, ctypes.POINTER(c_i64)],
    "hpgb_angle_to_pixel": [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_int, c_int, c_int, c_vp, c_vp],
    "hpgb_host_angle_to_pixel": [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_int, c_int, c_int, c_int,
                                 ctypes.POINTER(c_i64)],
    "hpgb_pixel_to_angle": [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_int, c_int, c_int, c_vp, c_vp],
    "hpgb_host_pixel_to_angle": [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_int, c_int, c_int, c_int,
                                 ctypes.POINTER(c_i64)],
    "hpgb_vector_to_pixel": [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_int, c_vp, c_vp],
    "hpgb_host_vector_to_pixel": [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_int, c_int, ctypes.POINTER(c_i64)],
    "hpgb_pixel_to_vector": [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_int, c_vp, c_vp],
    "hpgb_host_pixel_to_vector": [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_int, c_int, ctypes.POINTER(c_i64)],
    "hpgb_neighbors": [c_vp, c_vp, c_i64, c_i64, c_vp, c_int, c_vp, c_vp],
    "hpgb_host_neighbors": [c_vp, c_vp, c_i64, c_i64, c_vp, c_int, c_int, ctypes.POINTER(c_i64)],
    "hpgb_interp_weights": [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_int, c_int, c_int, c_vp, c_vp],
    "hpgb_host_interp_weights": [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_int, c_int, c_int, c_int,
                                 ctypes.POINTER(c_i64)],
    "hpgb_upgrade_pixels": [c_vp, c_vp, c_i64, c_i64, c_i64, c_int, c_vp, c_vp],
    "hpgb_host_upgrade_pixels": [c_vp, c_vp, c_i64, c_i64, c_i64, c_int, c_int, ctypes.POINTER(c_i64)],
    "hpgb_lonlat_to_thetaphi": [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_vp],
    "hpgb_host_lonlat_to_thetaphi": [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_int, ctypes.POINTER(c_i64)],
    "hpgb_thetaphi_to_lonlat": [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_vp],
    "hpgb_host_thetaphi_to_lonlat": [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_int, ctypes.POINTER(c_i64)],
    "hpgb_angle_to_vector": [c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp],
    "hpgb_host_angle_to_vector": [c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_int, ctypes.POINTER(c_i64)],
    "hpgb_vector_to_angle": [c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp],
    "hpgb_host_vector_to_angle": [c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_int, ctypes.POINTER(c_i64)],
    "hpgb_reorder_map": [c_vp, c_vp, c_i64, c_int, c_int, c_vp],
    "hpgb_host_reorder_map": [c_vp, c_vp, c_i64, c_int, c_int, c_int],
    "hpgb_boundaries": [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp],
    "hpgb_host_boundaries": [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_int,
                             ctypes.POINTER(c_i64)],
    "hpgb_max_pixel_radius": [c_vp, c_vp, c_i64, c_int, c_vp, c_vp],
    "hpgb_host_max_pixel_radius": [c_vp, c_vp, c_i64, c_int, c_int, ctypes.POINTER(c_i64)],
    "hpgb_ranges_offsets": [c_vp, c_i64, c_int, c_vp, c_vp, c_vp],
    "hpgb_ranges_expand": [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp],
    "hpgb_host_ranges_count": [c_vp, c_i64, c_int, c_int, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)],
    "hpgb_host_ranges_expand": [c_vp, c_i64, c_int, c_vp, c_i64, c_int],
    "hpgb_query_circle_ring_plan": [c_i64, c_dbl, c_dbl, c_dbl, c_int, c_i64, c_int, c_int, c_vp],
    "hpgb_query_circle_ring_work_bytes": [c_vp],
    "hpgb_query_circle_ring_ranges": [c_vp, c_vp, c_vp, c_vp],
    "hpgb_host_query_circle_ring_count": [c_vp, c_int, ctypes.POINTER(c_i64)],
    "hpgb_host_query_circle_ring_fill": [c_vp, c_vp, c_i64, c_int],
    "hpgb_query_circle_ring_plan_many": [c_i64, c_vp, c_vp, c_vp, c_i64, c_int, c_i64, c_int, c_int, c_vp,
                                         ctypes.POINTER(c_i64)],
    "hpgb_query_circle_ring_batch_prepare": [c_vp, c_i64, c_vp],
    "hpgb_query_circle_ring_batch_ranges": [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp],
    "hpgb_host_query_circle_ring_batch": [c_vp, c_i64, c_int, c_vp, ctypes.POINTER(c_i64), c_vp, c_i64],
    "hpgb_query_circle_nest_open": [c_i64, c_dbl, c_dbl, c_dbl, c_int, c_i64, c_int, c_int, c_int, ctypes.POINTER(c_vp),
                                    ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)],
    "hpgb_query_circle_nest_batch_open": [c_i64, c_vp, c_vp, c_vp, c_i64, c_int, c_i64, c_int, c_int, c_int,
                                          ctypes.POINTER(c_vp), ctypes.POINTER(c_i64), ctypes.POINTER(c_i64),
                                          ctypes.POINTER(c_i64)],
    "hpgb_query_nest_batch_offsets": [c_vp, c_vp],
    "hpgb_query_nest_ranges": [c_vp, c_vp, ctypes.POINTER(c_i64)],
    "hpgb_query_nest_pixels_host": [c_vp, c_vp, c_i64],
    "hpgb_query_nest_pixels_device": [c_vp, c_vp, c_i64, c_vp],
    "hpgb_query_nest_close": [c_vp],
    "hpgb_explain_nside": [c_i64, c_int, ctypes.c_char_p, c_int],
    "hpgb_explain_angle": [c_i64, c_int, c_dbl, c_dbl, c_int, c_int, ctypes.c_char_p, c_int],
    "hpgb_explain_pixel": [c_i64, c_int, c_i64, ctypes.c_char_p, c_int],
}
_RESTYPES = {"hpgb_error_string": ctypes.c_char_p, "hpgb_launch_count": c_i64,
             "hpgb_query_circle_ring_work_bytes": c_i64, "hpgb_query_circle_ring_batch_prepare": c_i64}

EXPORTS = tuple(_SIGS)


class Disc(ctypes.Structure):
    """hpgb_disc_t of include/hpgb.h"""
    _fields_ = [(n, c_i64) for n in ("nside", "fct", "whole_sphere", "irmin", "irmax", "north_hi", "south_lo", "cpix", "m")] \
        + [(n, c_dbl) for n in ("phi", "z0", "xa", "cosrbig", "cosrsmall")]

_lib = None


def lib():
    """Load libhpgb.so (once).  Raises ImportError with the build hint if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "hpgeom_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C hpgeom_b200/csrc`. There is no CPU fallback." % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        for name, args in _SIGS.items():
            fn = getattr(_lib, name)  # AttributeError here = library older than include/hpgb.h
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, c_int)
    return _lib


def error_string(code):
    return lib().hpgb_error_string(int(code)).decode()


def check_rc(rc, nside=None, nest=True):
    """Map a non-zero HPGB return code to the exception the reference raises."""
    if rc == 0:
        return
    if rc in (1, 2, 3) and nside is not None:
        raise ValueError(explain_nside(nside, nest))
    if rc == 5:
        raise RuntimeError("hpgeom_b200: no CUDA device visible; this package has no CPU fallback")
    raise RuntimeError("hpgeom_b200: %s (code %d)" % (error_string(rc), rc))


def explain_nside(nside, nest):
    buf = ctypes.create_string_buffer(256)
    lib().hpgb_explain_nside(int(nside), int(bool(nest)), buf, 256)
    return buf.value.decode()


def explain_angle(nside, nest, a, b, lonlat, degrees):
    buf = ctypes.create_string_buffer(256)
    lib().hpgb_explain_angle(int(nside), int(bool(nest)), float(a), float(b), int(bool(lonlat)),
                             int(bool(degrees)), buf, 256)
    return buf.value.decode()


def explain_pixel(nside, nest, pix):
    buf = ctypes.create_string_buffer(256)
    lib().hpgb_explain_pixel(int(nside), int(bool(nest)), int(pix), buf, 256)
    return buf.value.decode()
