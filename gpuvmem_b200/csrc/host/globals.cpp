#include "globals.hpp"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <map>

namespace gpuvmem {

Globals& G() {
  static Globals g;
  return g;
}

void gvmCheck(int rc, const char* what, const char* file, int line) {
  if (rc == 0) return;
  std::fprintf(stderr, "gpuvmem_b200 error at %s:%d: %s -> %s\n", file, line, what, gvm_last_error());
  // one rank leaving must not strand the others inside a collective
  if (G().engine && G().world > 1) gvm_dist_abort(G().engine);
  std::exit(-1);
}

namespace {
struct Site { double seconds = 0; long calls = 0; };
std::map<std::string, Site>& sites() { static std::map<std::string, Site> m; return m; }
}  // namespace
bool hostProfileOn() {
  static const bool on = [] { const char* v = std::getenv("GVM_PROFILE_HOST"); return v && *v && std::strcmp(v, "0") != 0; }();
  return on;
}
double hostProfileNow() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
void hostProfileAdd(const char* what, double seconds) {
  Site& s = sites()[what];
  s.seconds += seconds;
  s.calls++;
}
void hostProfileReport() {
  if (!hostProfileOn() || sites().empty()) return;
  std::vector<std::pair<std::string, Site>> v(sites().begin(), sites().end());
  std::sort(v.begin(), v.end(), [](const auto& a, const auto& b) { return a.second.seconds > b.second.seconds; });
  std::fprintf(stderr, "host profile (C-ABI call sites, wall time):\n");
  for (const auto& kv : v)
    std::fprintf(stderr, "  %10.3f ms %8ld x %8.1f us  %.90s\n", kv.second.seconds * 1e3, kv.second.calls,
                 kv.second.seconds * 1e6 / kv.second.calls, kv.first.c_str());
  sites().clear();
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
