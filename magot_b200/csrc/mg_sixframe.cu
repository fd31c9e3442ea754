// TEMPORARY stub while K4 is being written (see mg_sixframe.cu.draft)
#include "mg_common.cuh"
void mg_sixframe_free(mg_genome *) {}
extern "C" int mg_sixframe_count(mg_genome *, int64_t, int64_t, int64_t, int64_t *, int64_t *, void *) { mg_set_error("K4 not built yet"); return MG_ESTATE; }
extern "C" int mg_sixframe_emit(mg_genome *, uint8_t *, mg_orf *, void *) { mg_set_error("K4 not built yet"); return MG_ESTATE; }
extern "C" int mg_sixframe_emit_device(mg_genome *, uint8_t *, mg_orf *, void *) { mg_set_error("K4 not built yet"); return MG_ESTATE; }
