/* Shim for the CUDA-samples helper_cuda.h (absent offline): only the
 * checkCudaErrors macro the reference uses. When the harness runs the
 * reference's HOST-ONLY code (weighting, gridding, CKernel tables) in a
 * container without a GPU it sets gvref_tolerate_cuda_errors so that the
 * incidental cudaMalloc/cudaMemcpy calls in that code do not abort.
 * Test infrastructure only. */
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdio.h>
#include <stdlib.h>
extern int gvref_tolerate_cuda_errors;
template <typename T>
inline void gvref_check(T result, const char* func, const char* file, int line) {
  if (result && !gvref_tolerate_cuda_errors) {
    fprintf(stderr, "CUDA error at %s:%d code=%d \"%s\"\n", file, line,
            (int)result, func);
    exit(EXIT_FAILURE);
  }
}
#define checkCudaErrors(val) gvref_check((val), #val, __FILE__, __LINE__)
#define getLastCudaError(msg) checkCudaErrors(cudaGetLastError())
