// placeholder until the ANIm stages land
#include "common.cuh"
void free_index(MinimizerIndex *) {}
void free_anim_workspace(void *) {}
extern "C" {
const char *b200ani_stage_name(int) { return ""; }
int b200ani_anim(b200ani_ctx *ctx, const int32_t *, int64_t, int, int, b200ani_anim_result *) { FAIL(B200ANI_ERR_ARG, "anim: not built"); }
int64_t b200ani_anim_alignments(b200ani_ctx *ctx, int32_t, int32_t, int, int, int64_t *, int64_t) { FAIL(B200ANI_ERR_ARG, "anim: not built"); }
}
