// optimizer.hpp — Optimizer surface of the reference (include/classes/optimizer.cuh) with
// its two implementations: ConjugateGradient (Fletcher-Reeves/Polak-Ribiere, src/frprmn.cu,
// key "CG-FRPRMN") and LBFGS (src/lbfgs.cu, key "CG-LBFGS"), both on the Numerical-Recipes
// line search (linmin -> mnbrak + brent over f1dim; src/linmin.cu, mnbrak.cu, brent.cu,
// f1dim.cu). All image-sized arithmetic runs in fused gvm_vec_* kernels; work buffers are
// allocated once per optimize() call, not per line-search probe as in the reference.
#pragma once
#include <functional>
#include <vector>

#include "factory.hpp"
#include "image.hpp"
#include "objectivefunction.hpp"

namespace gpuvmem {

class Optimizer {
 public:
  Optimizer() = default;
  Optimizer(int total_iterations, float ftol) : total_iterations(total_iterations), ftol(ftol) {}
  Optimizer(int total_iterations, float ftol, float gtol) : total_iterations(total_iterations), ftol(ftol), gtol(gtol) {}
  virtual ~Optimizer() = default;

  virtual void allocateMemoryGpu() = 0;
  virtual void deallocateMemoryGpu() = 0;
  virtual void optimize() = 0;
  virtual int getK() { return 0; }
  virtual void setK(int) {}

  float getFtol() const { return ftol; }
  float getGtol() const { return gtol; }
  int getCurrentIteration() const { return current_iteration; }
  void setImage(Image* im) { image = im; }
  void setObjectiveFunction(ObjectiveFunction* o) { of = o; }
  void setFlag(int f) { flag = f; }
  void setFTol(float f) { ftol = f; }
  void setGTol(float g) { gtol = g; }
  void setTotalIterations(int n) { total_iterations = n; }
  ObjectiveFunction* getObjectiveFunction() { return of; }
  // why the last optimize() returned: "tolerance", "gradient tolerance", "gg = 0", "iterations"
  const char* getExitReason() const { return exit_reason; }
  // objective value after every outer iteration of the last optimize() ([0] = starting value)
  const std::vector<float>& getHistory() const { return history; }

 protected:
  ObjectiveFunction* of = nullptr;
  Image* image = nullptr;
  int flag = 0;
  int total_iterations = 500;
  int current_iteration = 0;
  float ftol = 1E-12;  // optimizer.cuh:15-16
  float gtol = 1E-12;
  int configured = 1;
  const char* exit_reason = "";
  std::vector<float> history;
};

// One-dimensional minimisation of of(p + x*xi) along xi; owns pcom/xicom/xt.
class LineSearch {
 public:
  LineSearch(ObjectiveFunction* of, Image* image) : of(of), image(image) {}
  ~LineSearch();
  // linmin (src/linmin.cu:52-116): on return p <- p + xmin*xi (projected), xi <- xmin*xi
  void linmin(float* p, float* xi, float* fret);
  float f1dim(float x);  // src/f1dim.cu:49-80
  // the bracketing + Brent part of linmin alone (starting abscissae 0 and 1)
  float minimize(float* xmin);
  // replaces the device evaluation of f1dim (host-logic tests on analytic functions)
  std::function<float(float)> probe_override;
  long probes = 0;

 private:
  struct Bracket { float ax, bx, cx, fa, fb, fc; };
  void mnbrak(Bracket& b);                                     // src/mnbrak.cu:44-98
  float brent(const Bracket& b, float tol, float* xmin);       // src/brent.cu:43-125
  void ensure();
  ObjectiveFunction* of;
  Image* image;
  float *pcom = nullptr, *xicom = nullptr, *xt = nullptr;
};

class ConjugateGradient : public Optimizer {
 public:
  using Optimizer::Optimizer;
  void allocateMemoryGpu() override;
  void deallocateMemoryGpu() override;
  void optimize() override;

 private:
  float *device_g = nullptr, *device_h = nullptr, *xi = nullptr;
  float fret = 0.0f, fp = 0.0f, gg = 0.0f, dgg = 0.0f, gam = 0.0f;
};

class LBFGS : public Optimizer {
 public:
  using Optimizer::Optimizer;
  void allocateMemoryGpu() override;
  void deallocateMemoryGpu() override;
  void optimize() override;
  int getK() override { return K; }
  void setK(int k) override { K = k; }

 private:
  void LBFGS_recursion(float* d_y, float* d_s, float* xi, int par_M, int lbfgs_it, int M, int N);
  float *d_y = nullptr, *d_s = nullptr, *xi = nullptr, *xi_old = nullptr, *p_old = nullptr;
  float *d_q = nullptr, *d_r = nullptr;
  float fret = 0.0f, fp = 0.0f, max_per_it = 0.0f;
  int K = 100;
};

}  // namespace gpuvmem
