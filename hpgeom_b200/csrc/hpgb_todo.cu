// TEMPORARY: entry points not implemented yet return HPGB_E_BAD_ARGUMENT (removed as kernels land)
#include "../../include/hpgb.h"
#define TODO return HPGB_E_BAD_ARGUMENT
extern "C" {
int hpgb_angle_to_pixel(const double *, const double *, int64_t *, int64_t, int64_t, const int64_t *, int, int, int, hpgb_status_t *, void *) { TODO; }
int hpgb_host_angle_to_pixel(const double *, const double *, int64_t *, int64_t, int64_t, const int64_t *, int, int, int, int, int64_t *) { TODO; }
int hpgb_pixel_to_angle(const int64_t *, double *, double *, int64_t, int64_t, const int64_t *, int, int, int, hpgb_status_t *, void *) { TODO; }
int hpgb_host_pixel_to_angle(const int64_t *, double *, double *, int64_t, int64_t, const int64_t *, int, int, int, int, int64_t *) { TODO; }
int hpgb_vector_to_pixel(const double *, const double *, const double *, int64_t *, int64_t, int64_t, const int64_t *, int, hpgb_status_t *, void *) { TODO; }
int hpgb_pixel_to_vector(const int64_t *, double *, double *, double *, int64_t, int64_t, const int64_t *, int, hpgb_status_t *, void *) { TODO; }
int hpgb_host_vector_to_pixel(const double *, const double *, const double *, int64_t *, int64_t, int64_t, const int64_t *, int, int, int64_t *) { TODO; }
int hpgb_host_pixel_to_vector(const int64_t *, double *, double *, double *, int64_t, int64_t, const int64_t *, int, int, int64_t *) { TODO; }
int hpgb_neighbors(const int64_t *, int64_t *, int64_t, int64_t, const int64_t *, int, hpgb_status_t *, void *) { TODO; }
int hpgb_host_neighbors(const int64_t *, int64_t *, int64_t, int64_t, const int64_t *, int, int, int64_t *) { TODO; }
int hpgb_interp_weights(const double *, const double *, int64_t *, double *, int64_t, int64_t, const int64_t *, int, int, int, hpgb_status_t *, void *) { TODO; }
int hpgb_host_interp_weights(const double *, const double *, int64_t *, double *, int64_t, int64_t, const int64_t *, int, int, int, int, int64_t *) { TODO; }
int hpgb_upgrade_pixels(const int64_t *, int64_t *, int64_t, int64_t, int64_t, int, hpgb_status_t *, void *) { TODO; }
int hpgb_host_upgrade_pixels(const int64_t *, int64_t *, int64_t, int64_t, int64_t, int, int, int64_t *) { TODO; }
int hpgb_lonlat_to_thetaphi(const double *, const double *, double *, double *, int64_t, int, hpgb_status_t *, void *) { TODO; }
int hpgb_host_lonlat_to_thetaphi(const double *, const double *, double *, double *, int64_t, int, int, int64_t *) { TODO; }
}
