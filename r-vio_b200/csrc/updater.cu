// TEMPORARY skeleton so the library links while the tracker is brought up on the GPU; replaced by the real updater.
#include "common.cuh"
using namespace rvio;
struct rvio_updater { int device; };
#define NOTYET(name) { set_error(name, "updater not built yet"); return RVIO_ERR_STATE; }
extern "C" int rvio_updater_create(const rvio_updater_cfg*, int, rvio_updater**) NOTYET("rvio_updater_create")
extern "C" void rvio_updater_destroy(rvio_updater*) {}
extern "C" int rvio_updater_update(rvio_updater*, const double*, int, const double*, int, const uint8_t*, const int32_t*, const float*, int, double*, double*, rvio_update_info*) NOTYET("rvio_updater_update")
extern "C" int rvio_updater_update_from_tracker(rvio_updater*, rvio_tracker*, const double*, int, const double*, int, double*, double*, rvio_update_info*) NOTYET("x")
extern "C" int rvio_updater_get_debug(rvio_updater*, int, uint8_t*, double*, double*, int32_t*) NOTYET("x")
extern "C" int rvio_updater_get_normal_terms(rvio_updater*, double*, double*, int) NOTYET("x")
extern "C" int rvio_updater_update_begin(rvio_updater*, const double*, int, const double*, int, const uint8_t*, const int32_t*, const float*, int, int, int) NOTYET("x")
extern "C" int rvio_updater_reduce_buffer(rvio_updater*, double**, int*) NOTYET("x")
extern "C" int rvio_updater_update_finish(rvio_updater*, double*, double*, rvio_update_info*) NOTYET("x")
extern "C" void* rvio_updater_stream(rvio_updater*) { return nullptr; }
