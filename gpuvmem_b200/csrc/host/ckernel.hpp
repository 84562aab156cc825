// ckernel.hpp — anti-aliasing convolution kernels and their gridding-correction
// functions (GCF). Class surface of the reference's include/classes/ckernel.cuh:6-551
// and src/{pillBox2D,gaussian2D,gaussianSinc2D,sinc2D,pswf_12D}.cu; factory keys
// "PillBox2D", "Gaussian2D", "GaussianSinc2D", "Sinc2D", "PSWF" (pswf_12D.cu:287).
//
// Design: the reference repeats the table loop in every subclass; here each family only
// supplies its two point functions (kernelAt / gcfAt) and the base class owns the loop,
// the table, its device copy (uploaded through the engine) and the GCF clone. The fp32
// arithmetic of every point function follows the reference expression by expression so
// that the tables are bit-identical (tests/test_host_cpu.py pins them against the
// reference build).
#pragma once
#include <cmath>
#include <string>
#include <vector>

#include "factory.hpp"

namespace gpuvmem {

class Io;

class CKernel {
 public:
  CKernel() { init(7, 7, 1.0f, 1.0f, 1.0f); }
  CKernel(int m, int n) { init(m, n, 1.0f, 1.0f, 1.0f); }
  CKernel(int m, int n, float w) { init(m, n, 1.0f, 1.0f, w); }
  CKernel(int m, int n, float dx, float dy) { init(m, n, dx, dy, 1.0f); }
  CKernel(int m, int n, float dx, float dy, float w) { init(m, n, dx, dy, w); }
  virtual ~CKernel();

  // point functions of the family (host, fp32)
  virtual float kernelAt(float amp, float x, float y, float x0, float y0, float sigma_x, float sigma_y) const = 0;
  virtual float gcfAt(float amp, float x, float y, float x0, float y0, float sigma_x, float sigma_y) const = 0;
  virtual CKernel* clone() const = 0;

  // reference surface
  virtual void buildKernel();
  virtual void buildKernel(float amp, float x0, float y0, float sigma_x, float sigma_y);
  virtual void buildGCF();
  virtual void buildGCF(float amp, float x0, float y0, float sigma_x, float sigma_y);
  virtual float GCF(float amp, float x, float y, float x0, float y0, float sigma_x, float sigma_y, float w) {
    const float keep = this->w;
    this->w = w;
    const float v = gcfAt(amp, x, y, x0, y0, sigma_x, sigma_y);
    this->w = keep;
    return v;
  }
  virtual void initializeGCF() { setGCF(clone()); }
  virtual void initializeGCF(int m, int n, float dx, float dy);
  virtual void initializeGCF(int m, int n, float dx, float dy, float w);
  virtual CKernel* getGCF() { return gcf; }
  virtual void setGCF(CKernel* g);
  virtual float* getGCFGPU() { return gcf ? gcf->getGPUKernel() : nullptr; }
  virtual std::vector<float> getGCFCPU() { return gcf ? gcf->getKernel() : std::vector<float>(); }
  virtual float* getGCFCPUPointer() { return gcf ? gcf->getKernelPointer() : nullptr; }
  virtual float getAlpha() { return 0.0f; }
  virtual void setAlpha(float) {}
  virtual float getW2() { return 0.0f; }
  virtual void setW2(float) {}

  float getAmp() const { return amp; }
  int getm() const { return m; }
  int getn() const { return n; }
  float getSigmaX() const { return sigma_x; }
  float getSigmaY() const { return sigma_y; }
  int getSupportX() const { return support_x; }
  int getSupportY() const { return support_y; }
  int getGPUID() const { return gpu_id; }
  float getW() const { return w; }
  float getX0() const { return x0; }
  float getY0() const { return y0; }
  float getKernelValue(int i, int j) const { return kernel[(size_t)n * i + j]; }
  std::vector<float> getKernel() const { return kernel; }
  float* getKernelPointer() { return kernel.data(); }
  float* getGPUKernel();   // device copy, uploaded lazily through the engine
  std::string getName() const { return name; }
  Io* getImageHandler() { return ioImageHandler; }

  void setName(const std::string& s) { name = s; }
  void setAmp(float a) { amp = a; }
  void setCenter(float cx, float cy) { x0 = cx; y0 = cy; }
  void setmn(int mm, int nn) { m = mm; n = nn; setSupports(); }
  void setSigmas(float dx, float dy) { sigma_x = dx; sigma_y = dy; }
  void setW(float ww) { w = ww; }
  void setIoImageHandler(Io* io) { ioImageHandler = io; }
  void setGPUID(int id) { gpu_id = id; }
  void printCKernel() const;
  void printGCF() const { if (gcf) gcf->printCKernel(); }

 protected:
  int m = 7, n = 7, support_x = 3, support_y = 3, gpu_id = 0;
  float amp = 1.0f, x0 = 0.0f, y0 = 0.0f, sigma_x = 1.0f, sigma_y = 1.0f, w = 1.0f;
  std::vector<float> kernel;
  float* gpu_kernel = nullptr;
  bool gpu_stale = true;
  Io* ioImageHandler = nullptr;
  CKernel* gcf = nullptr;
  std::string name;

  void copyBaseTo(CKernel* other) const;

 private:
  void init(int mm, int nn, float dx, float dy, float ww) {
    amp = 1.0f; m = mm; n = nn; sigma_x = dx; sigma_y = dy; x0 = y0 = 0.0f; w = ww;
    setSupports();
  }
  // both supports derive from m, as the reference does (ckernel.cuh:508-511)
  void setSupports() {
    support_x = (int)std::floor(m / 2.0f);
    support_y = (int)std::floor(m / 2.0f);
  }
  template <class F>
  void fill(F pointFn, float sx, float sy);
};

class PillBox2D : public CKernel {
 public:
  PillBox2D() : CKernel() { setmn(1, 1); name = "Pill Box"; }   // pillBox2D.cu:22-25
  PillBox2D(int m, int n) : CKernel(m, n) { name = "Pill Box"; }
  PillBox2D(int m, int n, float w) : CKernel(m, n, w) { name = "Pill Box"; }
  PillBox2D(int m, int n, float dx, float dy) : CKernel(m, n, dx, dy) { name = "Pill Box"; }
  PillBox2D(int m, int n, float dx, float dy, float w) : CKernel(m, n, dx, dy, w) { name = "Pill Box"; }
  float kernelAt(float amp, float x, float y, float x0, float y0, float sx, float sy) const override;
  float gcfAt(float, float, float, float, float, float, float) const override { return 1.0f; }
  CKernel* clone() const override { auto* k = new PillBox2D(*this); copyBaseTo(k); return k; }
};

class Gaussian2D : public CKernel {
 public:
  Gaussian2D() : CKernel() { name = "Gaussian"; }
  Gaussian2D(int m, int n) : CKernel(m, n) { name = "Gaussian"; }
  Gaussian2D(int m, int n, float w) : CKernel(m, n, w) { name = "Gaussian"; }
  Gaussian2D(int m, int n, float dx, float dy) : CKernel(m, n, dx, dy) { name = "Gaussian"; }
  Gaussian2D(int m, int n, float dx, float dy, float w) : CKernel(m, n, dx, dy, w) { name = "Gaussian"; }
  float getAlpha() override { return alpha; }
  void setAlpha(float a) override { alpha = a; }
  float kernelAt(float amp, float x, float y, float x0, float y0, float sx, float sy) const override;
  float gcfAt(float amp, float x, float y, float x0, float y0, float sx, float sy) const override;
  CKernel* clone() const override { auto* k = new Gaussian2D(*this); copyBaseTo(k); return k; }

 private:
  float alpha = 2.0f;
};

class Sinc2D : public CKernel {
 public:
  Sinc2D() : CKernel() { name = "Sinc"; }
  Sinc2D(int m, int n) : CKernel(m, n) { name = "Sinc"; }
  Sinc2D(int m, int n, float w) : CKernel(m, n, w) { name = "Sinc"; }
  Sinc2D(int m, int n, float dx, float dy) : CKernel(m, n, dx, dy) { name = "Sinc"; }
  Sinc2D(int m, int n, float dx, float dy, float w) : CKernel(m, n, dx, dy, w) { name = "Sinc"; }
  float kernelAt(float amp, float x, float y, float x0, float y0, float sx, float sy) const override;
  float gcfAt(float amp, float x, float y, float x0, float y0, float sx, float sy) const override;
  CKernel* clone() const override { auto* k = new Sinc2D(*this); copyBaseTo(k); return k; }
};

class GaussianSinc2D : public CKernel {
 public:
  GaussianSinc2D() : CKernel() { w = 2.52f; name = "Gaussian Sinc"; }
  GaussianSinc2D(int m, int n) : CKernel(m, n) { w = 2.52f; name = "Gaussian Sinc"; }
  GaussianSinc2D(int m, int n, float w, float w2) : CKernel(m, n, w), w2(w2) { name = "Gaussian Sinc"; }
  GaussianSinc2D(int m, int n, float dx, float dy, float w, float w2) : CKernel(m, n, dx, dy, w), w2(w2) { name = "Gaussian Sinc"; }
  float getAlpha() override { return alpha; }
  void setAlpha(float a) override { alpha = a; }
  float getW2() override { return w2; }
  void setW2(float v) override { w2 = v; }
  float kernelAt(float amp, float x, float y, float x0, float y0, float sx, float sy) const override;
  float gcfAt(float, float, float, float, float, float, float) const override { return 1.0f; }  // gaussianSinc2D.cuh:80-90
  CKernel* clone() const override { auto* k = new GaussianSinc2D(*this); copyBaseTo(k); return k; }

 private:
  float w2 = 1.55f, alpha = 2.0f;
};

class PSWF_12D : public CKernel {
 public:
  PSWF_12D() : CKernel() { w = 6.0f; name = "Prolate Spheroidal Wave Function (PSWF)"; }
  PSWF_12D(int m, int n) : CKernel(m, n) { w = 6.0f; name = "Prolate Spheroidal Wave Function (PSWF)"; }
  PSWF_12D(int m, int n, float w) : CKernel(m, n, w) { name = "Prolate Spheroidal Wave Function (PSWF)"; }
  PSWF_12D(int m, int n, float dx, float dy) : CKernel(m, n, dx, dy) { w = 6.0f; name = "Prolate Spheroidal Wave Function (PSWF)"; }
  PSWF_12D(int m, int n, float dx, float dy, float w) : CKernel(m, n, dx, dy, w) { name = "Prolate Spheroidal Wave Function (PSWF)"; }
  float kernelAt(float amp, float x, float y, float x0, float y0, float sx, float sy) const override;
  float gcfAt(float amp, float x, float y, float x0, float y0, float sx, float sy) const override;
  CKernel* clone() const override { auto* k = new PSWF_12D(*this); copyBaseTo(k); return k; }
};

}  // namespace gpuvmem
