// globals.hpp — the state the reference keeps in ~60 mutable globals (src/mfs.cu:4-45,
// `extern`-ed in src/functions.cu:37-71), gathered in one struct, plus the engine handle
// every adapter talks to through the C ABI (include/gvm_b200.h). Host-only C++: this
// layer never includes CUDA headers; all device work goes through gvm_*.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../../include/gvm_b200.h"

namespace gpuvmem {

struct Globals {
  gvm_engine* engine = nullptr;   // replaces vars_gpu / device_* globals
  long M = 0, N = 0;              // src/mfs.cu:4
  int image_count = 2;            // always 2 after MFS::configure (src/mfs.cu:176-180)
  int imagesChanged = 0;
  std::vector<float> penalizators;   // -Z
  int nPenalizators = 0;
  std::vector<float> initial_values; // -z ([0] already multiplied by -eta, src/mfs.cu:169)
  float eta = -1.0f, noise_cut = 10.0f, threshold = 0.0f, nu_0 = -1.0f;
  float noise_jypix = 0.0f, robust_param = 2.0f, random_probability = 1.0f;
  double DELTAX = 0, DELTAY = 0, deltau = 0, deltav = 0, ra = 0, dec = 0, crpix1 = 0, crpix2 = 0;
  double beam_bmaj = 0, beam_bmin = 0, beam_bpa = 0;
  int verbose_flag = 0;
  bool nopositivity = false, apply_noise = false, print_images = false, print_errors = false;
  bool save_model_input = false, radius_mask = false, modify_weights = false;
  int flag_opt = 0;               // src/frprmn.cu:46
  int num_gpus = 1, firstgpu = 0, multigpu = 0;
  int nMeasurementSets = 0;
  int max_number_vis = 0;
  // one process per GPU: this process's place in the job
  int rank = 0, world = 1;
  int dist_kind = 0;              // GVM_DIST_*: what this rank's blocks are (set by MFS::shardAndUpload)
  bool quiet = false;             // ranks > 0 stay silent
};
Globals& G();

// checkCudaErrors convention of the reference (print + exit), applied to the C ABI's
// int status + gvm_last_error().
void gvmCheck(int rc, const char* what, const char* file, int line);
// GVM_PROFILE_HOST=1 in the environment: wall time and call count per C-ABI call site, printed to
// stderr by hostProfileReport() (MFS::unSetDevice) — a host-side profile without external tools.
bool hostProfileOn();
double hostProfileNow();
void hostProfileAdd(const char* what, double seconds);
void hostProfileReport();
#define GVM_CHECK(call)                                                           \
  do {                                                                            \
    if (::gpuvmem::hostProfileOn()) {                                             \
      const double _t0 = ::gpuvmem::hostProfileNow();                             \
      const int _rc = (call);                                                     \
      ::gpuvmem::hostProfileAdd(#call, ::gpuvmem::hostProfileNow() - _t0);        \
      ::gpuvmem::gvmCheck(_rc, #call, __FILE__, __LINE__);                        \
    } else {                                                                      \
      ::gpuvmem::gvmCheck((call), #call, __FILE__, __LINE__);                     \
    }                                                                             \
  } while (0)

// Device buffers of n floats, zero-filled (cudaMalloc + cudaMemset pairs of the reference).
float* devAllocFloats(size_t n);
void devFree(void* p);
void devZero(float* p, size_t n);
void devCopyD2D(float* dst, const float* src, size_t n);
void devUpload(float* dst, const float* src, size_t n);
void devDownload(float* dst, const float* src, size_t n);

}  // namespace gpuvmem
