// grad_umma.cu — tcgen05 / TMEM gradient contraction (placeholder until the
// kernel lands; AUTO falls back to the CUDA-core path).
#include "gvm_internal.cuh"
bool gvm_grad_umma_supported(const gvm_engine*, const GvmChannel&) { return false; }
int gvm_grad_umma(gvm_engine*, GvmChannel&, int*) {
  gvm_set_error("UMMA gradient path not built");
  return 1;
}
