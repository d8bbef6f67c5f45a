"""healpy-flavoured names over the accelerated path (reference: hpgeom/healpy_compat.py).

Same signatures and defaults as the reference's shims -- RING ordering, theta/phi in radians,
``lonlat=True`` meaning degrees, (8, N) / (4, N) transposed outputs -- each one a call into
``hpgeom_b200.hpgeom``, i.e. into the CUDA kernels.  ``query_disc`` runs the RING query_circle
kernels (healpy's default ordering); ``query_polygon`` is not built and raises
``NotImplementedError``.
"""
import warnings

from .hpgeom import (
    UNSEEN,
    angle_to_pixel,
    angle_to_vector,
    boundaries as _boundaries,
    get_interpolation_weights,
    max_pixel_radius,
    neighbors,
    nest_to_ring,
    npixel_to_nside,
    nside_to_npixel,
    nside_to_order,
    nside_to_pixel_area,
    nside_to_resolution,
    order_to_nside,
    pixel_to_angle,
    pixel_to_vector,
    query_circle_vec,
    ring_to_nest,
    vector_to_angle,
    vector_to_pixel,
)

__all__ = [
    'ang2pix', 'pix2ang', 'query_disc', 'query_polygon', 'ring2nest', 'nest2ring', 'nside2npix', 'npix2nside',
    'nside2pixarea', 'nside2resol', 'nside2order', 'order2nside', 'ang2vec', 'vec2ang', 'pix2vec', 'vec2pix',
    'boundaries', 'get_all_neighbours', 'max_pixrad', 'get_interp_weights', 'UNSEEN',
]


def ang2pix(nside, theta, phi, nest=False, lonlat=False):
    """reference: hpgeom/healpy_compat.py:55-76"""
    return angle_to_pixel(nside, theta, phi, nest=nest, lonlat=lonlat)


def pix2ang(nside, pix, nest=False, lonlat=False):
    """reference: hpgeom/healpy_compat.py:79-100"""
    return pixel_to_angle(nside, pix, nest=nest, lonlat=lonlat)


def query_disc(nside, vec, radius, inclusive=False, fact=4, nest=False, buff=None):
    """reference: hpgeom/healpy_compat.py:103-142"""
    pixels = query_circle_vec(nside, vec, radius, inclusive=inclusive, fact=fact, nest=nest)
    if buff is not None:
        warnings.warn("In hpgeom, setting buff is less performant than simply returning the pixels.")
        buff[0: len(pixels)] = pixels
    return pixels


def query_polygon(nside, vertices, inclusive=False, fact=4, nest=False, buff=None):
    raise NotImplementedError("query_polygon is not built: hpgeom_b200 covers the elementwise conversion path and query_circle")


def nest2ring(nside, pix):
    """reference: hpgeom/healpy_compat.py:185-200"""
    return nest_to_ring(nside, pix)


def ring2nest(nside, pix):
    """reference: hpgeom/healpy_compat.py:203-218"""
    return ring_to_nest(nside, pix)


def nside2npix(nside):
    return nside_to_npixel(nside)


def npix2nside(nside):
    return npixel_to_nside(nside)


def nside2pixarea(nside, degrees=False):
    return nside_to_pixel_area(nside, degrees=degrees)


def nside2resol(nside, arcmin=False):
    return nside_to_resolution(nside, units='arcminutes' if arcmin else 'radians')


def nside2order(nside):
    return nside_to_order(nside)


def order2nside(order):
    return order_to_nside(order)


def ang2vec(theta, phi, lonlat=False):
    """reference: hpgeom/healpy_compat.py:331-349"""
    return angle_to_vector(theta, phi, lonlat=lonlat, degrees=True)


def vec2ang(vectors, lonlat=False):
    """reference: hpgeom/healpy_compat.py:352-374"""
    return vector_to_angle(vectors, lonlat=lonlat, degrees=True)


def boundaries(nside, pix, step=1, nest=False):
    """Pixel boundaries as unit vectors, (3, 4*step) or (N, 3, 4*step) (reference: hpgeom/healpy_compat.py:377-409)."""
    lon, lat = _boundaries(nside, pix, step=step, nest=nest)
    vec = angle_to_vector(lon, lat)  # (4*step, 3) or (4*step, N, 3)
    if lon.ndim == 1:
        return vec.transpose(-1, 0) if hasattr(vec, "is_cuda") else vec.transpose()
    return vec.permute(1, 2, 0) if hasattr(vec, "is_cuda") else vec.transpose([1, 2, 0])


def pix2vec(nside, pix, nest=False):
    return pixel_to_vector(nside, pix, nest=nest)


def vec2pix(nside, x, y, z, nest=False):
    return vector_to_pixel(nside, x, y, z, nest=nest)


def get_all_neighbours(nside, theta, phi=None, nest=False, lonlat=False):
    """8 neighbours as (8,) / (8, N); theta alone = pixel numbers (reference: hpgeom/healpy_compat.py:460-492)."""
    pix = theta if phi is None else angle_to_pixel(nside, theta, phi, nest=nest, lonlat=lonlat, degrees=True)
    return neighbors(nside, pix, nest=nest).transpose(-1, 0) if hasattr(pix, "is_cuda") else neighbors(nside, pix, nest=nest).transpose()


def max_pixrad(nside, degrees=False):
    return max_pixel_radius(nside, degrees=degrees)


def get_interp_weights(nside, theta, phi=None, nest=False, lonlat=False):
    """(4, N) pixels and weights; theta alone = pixel numbers (reference: hpgeom/healpy_compat.py:513-541)."""
    if phi is None:
        _theta, _phi = pixel_to_angle(nside, theta, nest=nest, lonlat=False)
        pixels, weights = get_interpolation_weights(nside, _theta, _phi, nest=nest, lonlat=False)
    else:
        pixels, weights = get_interpolation_weights(nside, theta, phi, nest=nest, lonlat=lonlat)
    return pixels.T, weights.T
