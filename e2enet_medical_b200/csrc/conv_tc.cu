// tcgen05 / TMA implicit-GEMM kernel for the stride-1 depth-shifted (1,3,3) convolutions.
// (placeholder until the kernel lands: the dispatcher reports it as unsupported)
#include "common.cuh"

int e2e_conv_tc_fwd(const e2e_gemm_t* p, cudaStream_t st) {
  (void)p; (void)st;
  e2e_set_error("conv_tc_fwd: tcgen05 path not built");
  return E2E_ERR_UNSUPPORTED;
}
