#include "globals.hpp"

namespace gpuvmem {

Globals& G() {
  static Globals g;
  return g;
}

void gvmCheck(int rc, const char* what, const char* file, int line) {
  if (rc == 0) return;
  std::fprintf(stderr, "gpuvmem_b200 error at %s:%d: %s -> %s\n", file, line, what, gvm_last_error());
  std::exit(-1);
}

float* devAllocFloats(size_t n) {
  void* p = nullptr;
  GVM_CHECK(gvm_dev_alloc(G().engine, n * sizeof(float), &p));
  return static_cast<float*>(p);
}
void devFree(void* p) {
  if (p && G().engine) GVM_CHECK(gvm_dev_free(G().engine, p));
}
void devZero(float* p, size_t n) { GVM_CHECK(gvm_dev_memset(G().engine, p, 0, n * sizeof(float))); }
void devCopyD2D(float* dst, const float* src, size_t n) {
  GVM_CHECK(gvm_dev_copy(G().engine, dst, src, n * sizeof(float), GVM_COPY_D2D));
}
void devUpload(float* dst, const float* src, size_t n) {
  GVM_CHECK(gvm_dev_copy(G().engine, dst, src, n * sizeof(float), GVM_COPY_H2D));
}
void devDownload(float* dst, const float* src, size_t n) {
  GVM_CHECK(gvm_dev_copy(G().engine, dst, src, n * sizeof(float), GVM_COPY_D2H));
}

}  // namespace gpuvmem
