// linesearch.cpp — bracketing + Brent minimisation along a search direction.
// Algorithms: Numerical Recipes mnbrak/brent as used by the reference (src/linmin.cu:52-116,
// src/mnbrak.cu:44-98, src/brent.cu:43-125, src/f1dim.cu:49-80). The reference mixes float
// variables with double literals; every expression below spells out the same promotions so
// that the sequence of probe abscissae is identical for identical function values.
#include <cmath>
#include <cstdio>

#include "optimizer.hpp"

namespace gpuvmem {

namespace {
const float kLinminTol = 1.0e-7;   // TOL, src/linmin.cu:35
const double kGold = 1.618034;     // default magnification of successive intervals
const double kGrowLimit = 100.0;   // maximum magnification of a parabolic step
const float kTiny = 1.0e-20;
const int kBrentMaxIter = 500;
const double kCGold = 0.3819660;   // golden-section fraction
const double kZeps = 1.0e-10;

inline float withSignOf(float magnitude, float sign_source) {
  return sign_source >= 0.0 ? std::fabs(magnitude) : -std::fabs(magnitude);
}
}  // namespace

LineSearch::~LineSearch() {
  if (G().engine) {
    devFree(pcom);
    devFree(xicom);
    devFree(xt);
  }
}

void LineSearch::ensure() {
  if (pcom) return;
  const size_t n = (size_t)G().M * G().N * G().image_count;
  pcom = devAllocFloats(n);
  xicom = devAllocFloats(n);
  xt = devAllocFloats(n);
}

float LineSearch::f1dim(float x) {
  Globals& g = G();
  probes++;
  if (probe_override) return probe_override(x);
  // xt = pcom + x*xicom, image 0 floored at -eta*MINPIX unless -x (src/f1dim.cu:59-74)
  if (image->stockMapping(g.nopositivity)) {
    GVM_CHECK(gvm_vec_evaluate_xt(g.engine, xt, pcom, xicom, x, image->getImageCount(), g.nopositivity ? 1 : 0));
  } else {
    imageMap* map = image->getFunctionMapping();
    for (int i = 0; i < image->getImageCount(); i++) map[i].evaluateXt(xt, pcom, xicom, x, i);
  }
  return of->calcFunction(xt);
}

// Given the two starting abscissae, walk downhill until a minimum is bracketed:
// on return f(bx) <= f(ax) and f(bx) <= f(cx).
void LineSearch::mnbrak(Bracket& k) {
  k.fa = f1dim(k.ax);
  k.fb = f1dim(k.bx);
  if (k.fb > k.fa) {  // make a -> b the downhill direction
    std::swap(k.ax, k.bx);
    std::swap(k.fa, k.fb);
  }
  k.cx = (float)(k.bx + kGold * (k.bx - k.ax));
  k.fc = f1dim(k.cx);
  while (k.fb > k.fc) {
    // abscissa of the parabola through a, b, c
    const float r = (k.bx - k.ax) * (k.fb - k.fc);
    const float q = (k.bx - k.cx) * (k.fb - k.fa);
    const float qr = q - r;
    const float guard = std::fabs(qr) > kTiny ? std::fabs(qr) : kTiny;
    const float num = (k.bx - k.cx) * q - (k.bx - k.ax) * r;
    float u = (float)(k.bx - num / (2.0 * withSignOf(guard, qr)));
    const float ulim = (float)(k.bx + kGrowLimit * (k.cx - k.bx));
    float fu;
    if ((k.bx - u) * (u - k.cx) > 0.0) {  // u between b and c
      fu = f1dim(u);
      if (fu < k.fc) {         // minimum between b and c
        k.ax = k.bx; k.fa = k.fb;
        k.bx = u;    k.fb = fu;
        return;
      } else if (fu > k.fb) {  // minimum between a and u
        k.cx = u; k.fc = fu;
        return;
      }
      u = (float)(k.cx + kGold * (k.cx - k.bx));  // parabolic fit was no use
      fu = f1dim(u);
    } else if ((k.cx - u) * (u - ulim) > 0.0) {   // u between c and its allowed limit
      fu = f1dim(u);
      if (fu < k.fc) {
        k.bx = k.cx; k.cx = u;
        u = (float)(k.cx + kGold * (k.cx - k.bx));
        k.fb = k.fc; k.fc = fu;
        fu = f1dim(u);
      }
    } else if ((u - ulim) * (ulim - k.cx) >= 0.0) {  // limit u to its maximum
      u = ulim;
      fu = f1dim(u);
    } else {
      u = (float)(k.cx + kGold * (k.cx - k.bx));
      fu = f1dim(u);
    }
    k.ax = k.bx; k.bx = k.cx; k.cx = u;
    k.fa = k.fb; k.fb = k.fc; k.fc = fu;
  }
}

// Brent's method inside the bracket (ax, bx, cx): parabolic interpolation when it behaves,
// golden section otherwise. Returns f at the minimum, abscissa in *xmin.
float LineSearch::brent(const Bracket& k, float tol, float* xmin) {
  float a = k.ax < k.cx ? k.ax : k.cx;
  float b = k.ax > k.cx ? k.ax : k.cx;
  float x = k.bx, w = k.bx, v = k.bx;
  float fx = f1dim(x), fw = fx, fv = fx;
  float d = 0.0f, e = 0.0f;
  for (int iter = 1; iter <= kBrentMaxIter; iter++) {
    const float xm = (float)(0.5 * (a + b));
    const float tol1 = (float)(tol * std::fabs(x) + kZeps);
    const float tol2 = (float)(2.0 * tol1);
    if (std::fabs(x - xm) <= (tol2 - 0.5 * (b - a))) {
      *xmin = x;
      return fx;
    }
    bool golden = true;
    if (std::fabs(e) > tol1) {  // try a parabolic step
      float r = (x - w) * (fx - fv);
      float q = (x - v) * (fx - fw);
      float p = (x - v) * q - (x - w) * r;
      q = (float)(2.0 * (q - r));
      if (q > 0.0) p = -p;
      q = std::fabs(q);
      const float etemp = e;
      e = d;
      const bool reject = std::fabs(p) >= std::fabs(0.5 * q * etemp) || p <= q * (a - x) || p >= q * (b - x);
      if (!reject) {
        d = p / q;
        const float u = x + d;
        if (u - a < tol2 || b - u < tol2) d = withSignOf(tol1, xm - x);
        golden = false;
      }
    }
    if (golden) {
      e = (x >= xm) ? a - x : b - x;
      d = (float)(kCGold * e);
    }
    const float u = (std::fabs(d) >= tol1) ? x + d : x + withSignOf(tol1, d);
    const float fu = f1dim(u);
    if (fu <= fx) {
      if (u >= x) a = x; else b = x;
      v = w; w = x; x = u;
      fv = fw; fw = fx; fx = fu;
    } else {
      if (u < x) a = u; else b = u;
      if (fu <= fw || w == x) {
        v = w; w = u;
        fv = fw; fw = fu;
      } else if (fu <= fv || v == x || v == w) {
        v = u;
        fv = fu;
      }
    }
  }
  std::printf("Too many iterations in brent\n");
  *xmin = x;
  return fx;
}

float LineSearch::minimize(float* xmin) {
  Bracket k;
  k.ax = 0.0f;
  k.bx = 1.0f;
  k.cx = 0.0f;
  k.fa = k.fb = k.fc = 0.0f;
  mnbrak(k);
  return brent(k, kLinminTol, xmin);
}

void LineSearch::linmin(float* p, float* xi, float* fret) {
  Globals& g = G();
  ensure();
  const size_t n = (size_t)g.M * g.N * g.image_count;
  devCopyD2D(pcom, p, n);
  devCopyD2D(xicom, xi, n);
  float xmin = 0.0f;
  *fret = minimize(&xmin);
  if (g.verbose_flag && !g.quiet) std::printf("Alpha for linear minimization = %f\n\n", xmin);
  // xi *= xmin; p += xi with the positivity projection (src/linmin.cu:92-111)
  if (image->stockMapping(g.nopositivity)) {
    GVM_CHECK(gvm_vec_new_p(g.engine, p, xi, xmin, image->getImageCount(), g.nopositivity ? 1 : 0));
  } else {
    imageMap* map = image->getFunctionMapping();
    for (int i = 0; i < image->getImageCount(); i++) map[i].newP(p, xi, xmin, i);
  }
}

// ----------------------------------------------------------- Image helpers --
// per-image variants (a caller-installed imageMap may mix them freely): each works on ONE
// image through pointer offsets into the shared fused kernels
namespace {
void evalOne(float* xt, float* pcom, float* xicom, float x, int image, bool positive) {
  Globals& g = G();
  const size_t off = (size_t)g.M * g.N * image;
  if (positive && image == 0) {
    GVM_CHECK(gvm_vec_evaluate_xt(g.engine, xt, pcom, xicom, x, 1, 0));
  } else {
    // y = 1*pcom ; y = x*xicom + 1*y
    GVM_CHECK(gvm_vec_evaluate_xt(g.engine, xt + off, pcom + off, xicom + off, x, 1, 1));
  }
}
void newPOne(float* p, float* xi, float xmin, int image, bool positive) {
  Globals& g = G();
  const size_t off = (size_t)g.M * g.N * image;
  if (positive && image == 0) GVM_CHECK(gvm_vec_new_p(g.engine, p, xi, xmin, 1, 0));
  else GVM_CHECK(gvm_vec_new_p(g.engine, p + off, xi + off, xmin, 1, 1));
}
}  // namespace
void defaultNewP(float* p, float* xi, float xmin, int image) { newPOne(p, xi, xmin, image, false); }
void particularNewP(float* p, float* xi, float xmin, int image) { newPOne(p, xi, xmin, image, true); }
void defaultEvaluateXt(float* xt, float* pcom, float* xicom, float x, int image) { evalOne(xt, pcom, xicom, x, image, false); }
void particularEvaluateXt(float* xt, float* pcom, float* xicom, float x, int image) { evalOne(xt, pcom, xicom, x, image, true); }

bool Image::stockMapping(bool nopositivity) const {
  if (!functionMapping) return true;
  for (int i = 0; i < image_count; i++) {
    const bool want_particular = !nopositivity && i == 0;
    if (functionMapping[i].newP != (want_particular ? particularNewP : defaultNewP)) return false;
    if (functionMapping[i].evaluateXt != (want_particular ? particularEvaluateXt : defaultEvaluateXt)) return false;
  }
  return true;
}

}  // namespace gpuvmem
