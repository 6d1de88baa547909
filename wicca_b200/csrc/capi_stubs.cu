// TEMPORARY: entry points not implemented yet (replaced file by file).
#include "host_common.h"
using namespace wicca;
extern "C" {
int wicca_batch_icons_u8(const uint8_t* const*, const int*, const int*, const int64_t*, int, int, const int*, int, int, double, uint8_t* const*, const int*, int, wicca_timing*) { return fail(WICCA_ESTATE, "not implemented"); }
int wicca_haar_forward_f32(const uint8_t*, int, int, int, int64_t, int, int, double, float*, int, wicca_timing*) { return fail(WICCA_ESTATE, "not implemented"); }
int wicca_haar_inverse_f32(const float*, int, int, int, int, float*, int, wicca_timing*) { return fail(WICCA_ESTATE, "not implemented"); }
int wicca_haar_forward_dev(const uint8_t*, int, int, int, int64_t, int, int, double, float*, float*, int, void*) { return fail(WICCA_ESTATE, "not implemented"); }
int wicca_haar_inverse_dev(const float*, int, int, int, int, float*, float*, int, void*) { return fail(WICCA_ESTATE, "not implemented"); }
int wicca_icon_resize_norm_f32(const uint8_t* const*, const int*, const int*, int, int, int, int, float*, uint8_t*, int, wicca_timing*) { return fail(WICCA_ESTATE, "not implemented"); }
}
