// weightingscheme.hpp — imaging-weight schemes. Surface of the reference's
// include/classes/weightingscheme.cuh:9-81 and src/{natural,uniform,briggs,radial}weightingscheme.cu;
// factory keys "Natural", "Uniform", "Briggs", "Radial". apply() runs on the GPU through
// gvm_weights (bit-exact cell indexing, fp32 sums in the reference's single-thread order).
#pragma once
#include <iostream>
#include <vector>

#include "factory.hpp"
#include "msdata.hpp"
#include "uvtaper.hpp"

namespace gpuvmem {

class WeightingScheme {
 public:
  WeightingScheme() = default;
  explicit WeightingScheme(int threads) : threads(threads) {}
  WeightingScheme(int threads, UVTaper* uvtaper) : threads(threads), uvtaper(uvtaper) {}
  WeightingScheme(int threads, UVTaper* uvtaper, bool modify_weights)
      : threads(threads), uvtaper(uvtaper), modify_weights(modify_weights) {}
  virtual ~WeightingScheme() = default;

  virtual void apply(std::vector<MSDataset>& d) = 0;
  virtual void configure(void* params) = 0;

  bool getModifyWeights() const { return modify_weights; }
  void setModifyWeights(bool m) { modify_weights = m; }
  int getThreads() const { return threads; }
  // kept for source compatibility: the arithmetic runs on the GPU, the count is unused
  void setThreads(int t) { threads = t; }
  UVTaper* getUVTaper() { return uvtaper; }
  void setUVTaper(UVTaper* t) {
    uvtaper = t;
    std::cout << "UVTaper has been set" << std::endl;
    std::cout << "UVTaper Features - bmaj=" << t->getSigma_maj() << ", bmin=" << t->getSigma_min()
              << ", bpa=" << t->getBPA() << std::endl;
  }
  void restoreWeights(std::vector<MSDataset>& d);

 protected:
  int threads = 1;
  UVTaper* uvtaper = nullptr;
  bool modify_weights = false;
  // all (dataset, field, channel, stokes) blocks in the reference's loop order -> one gvm_weights call
  void applyOnGpu(int scheme, float robust, std::vector<MSDataset>& d, const char* label);
};

class NaturalWeightingScheme : public WeightingScheme {
 public:
  using WeightingScheme::WeightingScheme;
  void apply(std::vector<MSDataset>& d) override { applyOnGpu(GVM_W_NATURAL, 0.0f, d, "Natural"); }
  void configure(void*) override {}
};
class UniformWeightingScheme : public WeightingScheme {
 public:
  using WeightingScheme::WeightingScheme;
  void apply(std::vector<MSDataset>& d) override { applyOnGpu(GVM_W_UNIFORM, 0.0f, d, "Uniform"); }
  void configure(void*) override {}
};
class RadialWeightingScheme : public WeightingScheme {
 public:
  using WeightingScheme::WeightingScheme;
  void apply(std::vector<MSDataset>& d) override { applyOnGpu(GVM_W_RADIAL, 0.0f, d, "Radial"); }
  void configure(void*) override {}
};
class BriggsWeightingScheme : public WeightingScheme {
 public:
  using WeightingScheme::WeightingScheme;
  float getRobustParam() const { return robust_param; }
  // briggsweightingscheme.cu:13-21: R must lie in [-2, 2]
  void setRobustParam(float r);
  void configure(void* params) override;
  void apply(std::vector<MSDataset>& d) override { applyOnGpu(GVM_W_BRIGGS, robust_param, d, "Briggs"); }

 private:
  float robust_param = 2.0f;
};

}  // namespace gpuvmem
