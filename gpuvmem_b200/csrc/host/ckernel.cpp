// ckernel.cpp — point functions and table construction of the CKernel family.
// fp32 throughout, mirroring the reference's expressions (files cited per function).
#include "ckernel.hpp"

#include <cstdio>

#include "globals.hpp"

namespace gpuvmem {
namespace {

const float kPi = 3.14159265358979323846f;  // CUDART_PI_F (include/MSFITSIO.cuh:52)

// distance() of src/MSFITSIO.cu:47-51
inline float dist2(float x, float y, float cx, float cy) {
  const float dx = x - cx, dy = y - cy;
  return sqrtf(dx * dx + dy * dy);
}

// gaussian2D, src/gaussian2D.cu:17-38: separable super-Gaussian cut at w*sigma
float gaussWindow(float amp, float x, float y, float cx, float cy, float sx, float sy, float w, float alpha) {
  const float rx = dist2(x, 0.0f, cx, 0.0f);
  const float ry = dist2(0.0f, y, 0.0f, cy);
  if (!(rx < w * sx && ry < w * sy)) return 0.0f;
  const float ex = powf(rx / (w * sx), alpha);
  const float ey = powf(ry / (w * sy), alpha);
  return amp * expf(-1.0f * (ex + ey));
}

// sincf + sinc1D, src/sinc2D.cu:3-22
float sincCut(float x, float centre, float sigma, float w) {
  const float r = dist2(x, 0.0f, centre, 0.0f);
  const float t = r / (w * sigma);
  if (!(r < w * sigma)) return 0.0f;
  const float s = (t == 0.0f) ? 1.0f : sinf(kPi * t) / (kPi * t);
  return 1.0f * s;
}

// pswf_11D_func, src/pswf_12D.cu:3-47: rational approximation of the 0-order prolate
// spheroidal function on two sub-intervals (coefficients of :8-13)
float spheroidal(float nu) {
  static const float P[2][5] = {{8.203343e-2, -3.644705e-1, 6.278660e-1, -5.335581e-1, 2.312756e-1},
                                {4.028559e-3, -3.697768e-2, 1.021332e-1, -1.201436e-1, 6.412774e-2}};
  static const float Q[2][3] = {{1.0000000e0, 8.212018e-1, 2.078043e-1}, {1.0000000e0, 9.599102e-1, 2.918724e-1}};
  const float a = fabsf(nu);
  if (a > 1.0f) return 0.0f;
  const int part = (a >= 0.0f && a < 0.75) ? 0 : 1;
  const float edge = part == 0 ? 0.75f : 1.0f;
  const float d = a * a - edge * edge;
  float num = P[part][0], den = Q[part][0];
  for (int k = 1; k < 5; k++) num += P[part][k] * powf(d, k);
  for (int k = 1; k < 3; k++) den += Q[part][k] * powf(d, k);
  return den > 0.0f ? num / den : 0.0f;
}
// pswf_11D, src/pswf_12D.cu:49-62
float spheroidal1D(float amp, float x, float centre, float sigma, float w) {
  const float r = dist2(x, 0.0f, centre, 0.0f);
  const float nu = r / (w * sigma);
  if (nu == 0.0f) return 1.0f;
  const float psi = spheroidal(nu);
  const float nusq = nu * nu;
  return amp * (1.0f - nusq) * psi;
}

}  // namespace

CKernel::~CKernel() {
  if (gpu_kernel && G().engine) devFree(gpu_kernel);
}

void CKernel::copyBaseTo(CKernel* other) const {
  // a clone owns its own (not yet uploaded) device copy and no GCF of its own
  other->gpu_kernel = nullptr;
  other->gpu_stale = true;
  other->gcf = nullptr;
}

template <class F>
void CKernel::fill(F pointFn, float sx, float sy) {
  kernel.assign((size_t)m * n, 0.0f);
  for (int i = 0; i < m; i++) {
    const float y = (i - support_y) * sy;
    for (int j = 0; j < n; j++) {
      const float x = (j - support_x) * sx;
      kernel[(size_t)n * i + j] = pointFn(x, y);
    }
  }
  gpu_stale = true;
}

void CKernel::buildKernel() {
  fill([this](float x, float y) { return kernelAt(amp, x, y, x0, y0, sigma_x, sigma_y); }, sigma_x, sigma_y);
}
void CKernel::buildKernel(float a, float cx, float cy, float sx, float sy) {
  fill([=](float x, float y) { return kernelAt(a, x, y, cx, cy, sx, sy); }, sx, sy);
}
void CKernel::buildGCF() {
  fill([this](float x, float y) { return gcfAt(amp, x, y, x0, y0, sigma_x, sigma_y); }, sigma_x, sigma_y);
}
void CKernel::buildGCF(float a, float cx, float cy, float sx, float sy) {
  fill([=](float x, float y) { return gcfAt(a, x, y, cx, cy, sx, sy); }, sx, sy);
}

void CKernel::setGCF(CKernel* g) {
  if (gcf && gcf != g) delete gcf;
  gcf = g;
}
// ckernel.cuh:82-88: the GCF is a clone of the kernel evaluated on the m x n image grid with
// sigma = pixel size in radians and w = m
void CKernel::initializeGCF(int mm, int nn, float dx, float dy) { initializeGCF(mm, nn, dx, dy, (float)mm); }
void CKernel::initializeGCF(int mm, int nn, float dx, float dy, float ww) {
  CKernel* g = clone();
  g->setmn(mm, nn);
  g->setSigmas(dx, dy);
  g->setW(ww);
  g->buildGCF();
  setGCF(g);
}

float* CKernel::getGPUKernel() {
  if (kernel.empty()) return nullptr;
  if (gpu_stale) {
    if (gpu_kernel) devFree(gpu_kernel);
    gpu_kernel = devAllocFloats(kernel.size());
    devUpload(gpu_kernel, kernel.data(), kernel.size());
    gpu_stale = false;
  }
  return gpu_kernel;
}

void CKernel::printCKernel() const {
  if (m > 16 || n > 16) {
    std::printf("%s: %d x %d table (not printed)\n", name.c_str(), m, n);
    return;
  }
  for (int i = 0; i < m; i++) {
    for (int j = 0; j < n; j++) std::printf("%.6e ", kernel.empty() ? 0.0f : kernel[(size_t)n * i + j]);
    std::printf("\n");
  }
}

// pillBox2D, src/pillBox2D.cu:3-16 with the limits of :133-134
float PillBox2D::kernelAt(float a, float x, float y, float, float, float sx, float sy) const {
  const float lx = (m / 2.0f) * sx, ly = (n / 2.0f) * sy;
  const float bx = (fabs(x) < lx) ? a : 0.0f;
  const float by = (fabs(y) < ly) ? a : 0.0f;
  return bx * by;
}

float Gaussian2D::kernelAt(float a, float x, float y, float cx, float cy, float sx, float sy) const {
  return gaussWindow(a, x, y, cx, cy, sx, sy, w, alpha);
}
// Gaussian2D::GCF, src/gaussian2D.cu:194-205
float Gaussian2D::gcfAt(float a, float x, float y, float cx, float cy, float sx, float sy) const {
  return gaussWindow(a, kPi * x, kPi * y, kPi * cx, kPi * cy, sx, sy, 2.0f * w, alpha);
}

// Sinc2D::buildKernel calls sinc2D(amp, x, y, x0, y0, ...) on a function declared
// sinc2D(amp, x, x0, y, y0, ...) (src/sinc2D.cu:24-31 vs :146-147): the second argument is
// taken as the x centre and x0 as the y coordinate. Kept so that the tables match.
float Sinc2D::kernelAt(float a, float x, float y, float cx, float cy, float sx, float sy) const {
  const float px = sincCut(x, /*centre*/ y, sx, w);
  const float py = sincCut(/*coordinate*/ cx, cy, sy, w);
  return a * px * py;
}
// Sinc2D::GCF, src/sinc2D.cu:169-189
float Sinc2D::gcfAt(float a, float x, float y, float cx, float cy, float sx, float sy) const {
  const float dxs = dist2(x, y, cx, cy) * sx, dys = dist2(x, y, cx, cy) * sy;
  const float bx = (fabs(dxs) < w * sx) ? a : 0.0f;
  const float by = (fabs(dys) < w * sy) ? a : 0.0f;
  return bx * by;
}

// gaussianSinc2D, src/gaussianSinc2D.cu:14-27 (here sinc2D gets its arguments in order)
float GaussianSinc2D::kernelAt(float a, float x, float y, float cx, float cy, float sx, float sy) const {
  const float gpart = gaussWindow(1.0f, x, y, cx, cy, sx, sy, w, alpha);
  const float spart = 1.0f * sincCut(x, cx, sx, w2) * sincCut(y, cy, sy, w2);
  return a * gpart * spart;
}

// pswf_12D, src/pswf_12D.cu:64-76; GCF = 1/pswf (:261-272)
float PSWF_12D::kernelAt(float a, float x, float y, float cx, float cy, float sx, float sy) const {
  const float px = spheroidal1D(1.0f, x, cx, sx, w);
  const float py = spheroidal1D(1.0f, y, cy, sy, w);
  return a * px * py;
}
float PSWF_12D::gcfAt(float a, float x, float y, float cx, float cy, float sx, float sy) const {
  return 1.0f / kernelAt(a, x, y, cx, cy, sx, sy);
}

namespace {
CKernel* makePillBox() { return new PillBox2D; }
CKernel* makeGaussian() { return new Gaussian2D; }
CKernel* makeSinc() { return new Sinc2D; }
CKernel* makeGaussianSinc() { return new GaussianSinc2D; }
CKernel* makePSWF() { return new PSWF_12D; }
const bool kRegistered[] = {
    registerCreationFunction<CKernel, std::string>("PillBox2D", makePillBox),
    registerCreationFunction<CKernel, std::string>("Gaussian2D", makeGaussian),
    registerCreationFunction<CKernel, std::string>("Sinc2D", makeSinc),
    registerCreationFunction<CKernel, std::string>("GaussianSinc2D", makeGaussianSinc),
    registerCreationFunction<CKernel, std::string>("PSWF", makePSWF),
};
}  // namespace

}  // namespace gpuvmem
