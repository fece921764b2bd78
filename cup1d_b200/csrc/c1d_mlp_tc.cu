// Tensor-core (tcgen05, 3xTF32) emulator path — placeholder until the kernel lands.
#include "c1d_internal.h"

int c1d_tc_prepare(c1d_ctx* ctx) { (void)ctx; return C1D_OK; }

int c1d_tc_launch(c1d_ctx* ctx, const double*, int64_t, double*, int, cudaStream_t) {
  snprintf(ctx->err, sizeof(ctx->err), "tensor-core emulator path not built");
  return C1D_ERR_STATE;
}

void c1d_tc_destroy(c1d_ctx* ctx) { (void)ctx; }
