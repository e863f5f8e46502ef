// attn_tcgen05.cu — K6 placeholder until the tcgen05 attention lands; fails loudly, never falls back.
#include "common.cuh"
int attn_tc_init(nb200_ctx *) { return NB200_OK; }
int launch_attention_tc(nb200_ctx *ctx, const bf16 *, bf16 *, int, int, int) {
    return nb200_fail(ctx, NB200_UNSUPPORTED_SHAPE, "tcgen05 attention kernel not built in this revision (set NB200_ATTN=simt)");
}
