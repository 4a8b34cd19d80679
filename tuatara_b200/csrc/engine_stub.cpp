// TEMPORARY: engine entry points until engine.cpp lands.
#include "common.h"
#include "engine.h"
extern "C" {
int tt_engine_create(const char*, const int*, int, const tt_config*, tt_engine**) { tt::set_error("engine: not built yet"); return 1; }
void tt_engine_destroy(tt_engine*) {}
int tt_ocr_pages(tt_engine*, const tt_image*, int, tt_result**) { tt::set_error("engine: not built yet"); return 1; }
void tt_result_free(tt_result*) {}
int tt_craft_forward(tt_engine*, const uint8_t*, int, int, float*) { tt::set_error("engine: not built yet"); return 1; }
int tt_parseq_forward(tt_engine*, const uint8_t*, int, const int32_t*, float*, int32_t*) { tt::set_error("engine: not built yet"); return 1; }
int tt_postprocess_dev(tt_engine*, const float*, int, int, int, int*, void*) { tt::set_error("engine: not built yet"); return 1; }
}
