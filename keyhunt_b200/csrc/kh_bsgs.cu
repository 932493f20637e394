// kh_bsgs.cu — BSGS half of the C ABI (placeholder until the kernels land in the next commit).
#include "kh_ctx.cuh"

extern "C" {
int kh_bsgs_build(kh_ctx *c, uint64_t, uint32_t) { return kh_fail(c, KH_ESTATE, "bsgs not built yet"); }
int kh_bsgs_describe(kh_ctx *c, kh_bsgs_desc *) { return kh_fail(c, KH_ESTATE, "bsgs not built yet"); }
int kh_bsgs_export(kh_ctx *c, int, int, void *, uint64_t) { return kh_fail(c, KH_ESTATE, "bsgs not built yet"); }
int kh_bsgs_import(kh_ctx *c, int, int, const void *, uint64_t) { return kh_fail(c, KH_ESTATE, "bsgs not built yet"); }
int kh_bsgs_search(kh_ctx *c, const uint8_t *, const uint8_t *, const uint8_t *, uint8_t *, int *) { return kh_fail(c, KH_ESTATE, "bsgs not built yet"); }
}
