// placeholder, replaced below
#include "model.h"
namespace gnnb {
int fused_prepare(gnnb_model *m) { m->fused = nullptr; return GNNB_OK; }
void fused_release(gnnb_model *m) { m->fused = nullptr; }
bool fused_supports(const gnnb_model *, int, int) { return false; }
int fused_run(gnnb_model *, const float *, const int32_t *, const int64_t *, const int64_t *, int,
              float *, cudaStream_t, int *) { set_error("fused path not built"); return GNNB_ERR_INVALID; }
int fused_tile_rows(const gnnb_model *) { return 0; }
}
