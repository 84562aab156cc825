// weights_grid.cu — imaging weights and convolutional gridding (placeholder for the
// first GPU bring-up; replaced by the GPU implementation).
#include "gvm_internal.cuh"
extern "C" {
int gvm_weights(int, int, float, int64_t, int64_t, double, double, int, const int64_t*,
                const double* const*, const float*, float* const*, const gvm_taper*) {
  gvm_set_error("gvm_weights: not built yet");
  return 1;
}
int gvm_grid_block(int, int64_t, int64_t, double, double, float, int64_t, const double*,
                   const float*, const float*, const float*, int, int, int, int, double*, float*,
                   float*, int64_t*) {
  gvm_set_error("gvm_grid_block: not built yet");
  return 1;
}
}
