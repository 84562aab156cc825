// fi.hpp — terms of the objective function. Surface of the reference's
// include/classes/fi.cuh:13-104 and of its subclasses (src/chi2.cu, entropy.cu, l1norm.cu,
// totalvariation.cu, totalsquaredvariation.cu, laplacian.cu, quadraticpenalization.cu,
// gentropy.cu, gl1norm.cu); factory keys as in those files.
// Each adapter is a thin binding onto the C ABI: gvm_chi2/gvm_dchi2 and
// gvm_prior_value/gvm_prior_grad/gvm_add_to_dphi.
#pragma once
#include <string>
#include <vector>

#include "ckernel.hpp"
#include "factory.hpp"
#include "globals.hpp"

namespace gpuvmem {

class Fi {
 public:
  Fi() = default;
  virtual ~Fi();

  virtual float calcFi(float* p) = 0;
  virtual void calcGi(float* p, float* xi) = 0;
  virtual void restartDGi();
  virtual void addToDphi(float* device_dphi);
  virtual void setPrior(float) {}
  virtual void setPrior(float*) {}
  virtual float getEta() { return 0.0f; }
  virtual void setEta(float) {}
  virtual void setCKernel(CKernel*) {}
  virtual void setFgScale(float) {}
  virtual float getFgScale() { return 1.0f; }
  virtual float calculateSecondDerivate() { return 0.0f; }
  // Single-synchronisation path of ObjectiveFunction::calcFunction (engine extension; plugins that do not
  // override it are evaluated through calcFi as in the reference). enqueueFi launches the term's value
  // asynchronously into result slot `slot` of the engine and returns true (false: not supported);
  // finishFi is what calcFi does once the value is on the host.
  virtual bool enqueueFi(float* p, int slot);
  // Everything enqueueFi bakes into its launches besides the image pointer (kind, parameters, gate; Chi2: the
  // scalars it hands to the engine): ObjectiveFunction replays a captured CUDA graph only while the keys match.
  virtual uint64_t stateKey();
  // the host-side part of enqueueFi, for evaluations that replay a captured graph instead of launching
  virtual void noteEnqueued();
  // Fused gradient step (ObjectiveFunction::calcGradient's fast path): restartDGi + calcGi + addToDphi of this term
  // applied straight to `dphi` (already zeroed) without the per-term device_DS / result buffers. Returns false when
  // the term does not support it (user plugins): the evaluation then takes the reference's loop. `first`: the term
  // is the first of the objective function (Chi2 REPLACES dphi with its gradient, src/chi2.cu:60-70 — that equals
  // accumulating into the zeroed dphi only when nothing was added before).
  virtual bool gradInto(float* p, float* dphi, bool first);
  // image the term's gradient is added to (TVariation: always 0, src/totalvariation.cu:46)
  virtual int dphiImage() const { return imageToAdd; }
  float finishFi(float value);
  // fi.cuh:58-89: penalizatorIndex -1 keeps the current factor; an index past the -Z list
  // disables the term (factor 0); a negative one is a configuration error (print + exit)
  virtual void configure(int penalizatorIndex, int imageIndex, int imageToAdd, bool normalize);

  std::string getName() const { return name; }
  void setName(const std::string& n) { name = n; }
  float get_fivalue() const { return fi_value; }
  bool getNormalize() const { return normalize; }
  float getPenalizationFactor() const { return penalization_factor; }
  void set_fivalue(float fi) { fi_value = fi; }
  void setPenalizationFactor(float p) { penalization_factor = p; }
  void setIteration(int it) { iteration = it; }
  void setNormalize(bool n) { normalize = n; }
  void setS(float* S);
  void setDS(float* DS);

 protected:
  float fi_value = 0.0f;
  float* device_S = nullptr;   // unused by the fused value kernels; kept for setS/setSandDs callers
  float* device_DS = nullptr;  // [M*N] gradient of this term
  float penalization_factor = 1.0f;
  int imageIndex = 0, iteration = 0, mod = 0, order = 0, imageToAdd = 0;
  std::string name = "default";
  bool normalize = false;

  // shared by all priors: the (iter > 0 && lambda) gate of e.g. src/functions.cu:4643 and the
  // flag_opt % 2 == imageIndex gate of the gradients (:4666)
  float priorValue(int kind, float* p, const gvm_prior_params& pp);
  void priorGrad(int kind, float* p, const gvm_prior_params& pp);
  // kind + parameters of a built-in prior (false: not a built-in prior)
  virtual bool priorSpec(int* kind, gvm_prior_params* pp) { (void)kind; (void)pp; return false; }
  bool enqueued_gate_closed = false;   // enqueueFi found the (iter > 0 && lambda) gate closed: the value is 0
};

class Chi2 : public Fi {
 public:
  Chi2() { name = "Chi2"; }
  ~Chi2() override;
  float calcFi(float* p) override;
  void calcGi(float* p, float* xi) override;
  void restartDGi() override;
  void addToDphi(float* device_dphi) override;
  void configure(int penalizatorIndex, int imageIndex, int imageToAdd, bool normalize) override;
  void setCKernel(CKernel* ck) override;
  void setFgScale(float s) override { fg_scale = s; }
  float getFgScale() override { return fg_scale; }
  bool enqueueFi(float* p, int slot) override;
  uint64_t stateKey() override;
  void noteEnqueued() override;
  bool gradInto(float* p, float* dphi, bool first) override;

 private:
  float* result_dchi2 = nullptr;  // [image_count][M*N]
  float fg_scale = 1.0f;
  CKernel* ckernel = nullptr;
};

class Entropy : public Fi {
 public:
  bool priorSpec(int* kind, gvm_prior_params* pp) override;
  Entropy() { name = "Entropy"; }
  explicit Entropy(float prior_value) : prior_value(prior_value) { name = "Entropy"; }
  Entropy(float prior_value, float eta) : prior_value(prior_value), eta(eta) { name = "Entropy"; }
  float getPrior() const { return prior_value; }
  void setPrior(float v) override { prior_value = v; }
  float getEta() override { return eta; }
  void setEta(float e) override { eta = e; }
  float calcFi(float* p) override;
  void calcGi(float* p, float* xi) override;

 private:
  float prior_value = 1.0f, eta = -1.0f;
};

class L1norm : public Fi {
 public:
  bool priorSpec(int* kind, gvm_prior_params* pp) override;
  L1norm() { name = "L1 Norm"; }
  explicit L1norm(float epsilon) : epsilon(epsilon) { name = "L1 Norm"; }
  float getEpsilon() const { return epsilon; }
  void setEpsilon(float e) { epsilon = e; }
  float calcFi(float* p) override;
  void calcGi(float* p, float* xi) override;

 private:
  float epsilon = 1E-12;
};

class TVariation : public Fi {
 public:
  bool priorSpec(int* kind, gvm_prior_params* pp) override;
  TVariation() { name = "Total Variation"; }
  explicit TVariation(float epsilon) : epsilon(epsilon) { name = "Total Variation"; }
  int dphiImage() const override { return 0; }
  float getEpsilon() const { return epsilon; }
  void setEpsilon(float e) { epsilon = e; }
  float calcFi(float* p) override;
  void calcGi(float* p, float* xi) override;
  void addToDphi(float* device_dphi) override;  // always image 0 (src/totalvariation.cu:44-46)

 private:
  float epsilon = 1E-12;
};

class TSqVariation : public Fi {
 public:
  bool priorSpec(int* kind, gvm_prior_params* pp) override;
  TSqVariation() { name = "Total Squared Variation"; }
  float calcFi(float* p) override;
  void calcGi(float* p, float* xi) override;
};

class Laplacian : public Fi {
 public:
  bool priorSpec(int* kind, gvm_prior_params* pp) override;
  Laplacian() { name = "Laplacian"; }
  float calcFi(float* p) override;
  void calcGi(float* p, float* xi) override;
};

class QuadraticP : public Fi {
 public:
  bool priorSpec(int* kind, gvm_prior_params* pp) override;
  QuadraticP() { name = "Quadratic"; }
  float calcFi(float* p) override;
  void calcGi(float* p, float* xi) override;
};

// entropy / L1 against a prior IMAGE (device pointer, M*N floats, owned by the term)
class GEntropy : public Fi {
 public:
  bool priorSpec(int* kind, gvm_prior_params* pp) override;
  GEntropy() { name = "GEntropy"; }
  explicit GEntropy(float* prior) : prior(prior) { name = "GEntropy"; }
  GEntropy(float* prior, float normalization_factor) : prior(prior), normalization_factor(normalization_factor) { name = "GEntropy"; }
  explicit GEntropy(const std::vector<float>& prior_host);
  ~GEntropy() override;
  float getNormalizationFactor() const { return normalization_factor; }
  void setNormalizationFactor(float f) { normalization_factor = f; }
  float getEta() override { return eta; }
  void setEta(float e) override { eta = e; }
  void setPrior(float* p) override;
  void normalizePrior();
  float calcFi(float* p) override;
  void calcGi(float* p, float* xi) override;

 private:
  float* prior = nullptr;
  float normalization_factor = 1.0f, eta = -1.0f;
};

class GL1Norm : public Fi {
 public:
  bool priorSpec(int* kind, gvm_prior_params* pp) override;
  GL1Norm() { name = "G L1-Norm"; }
  explicit GL1Norm(float* prior) : prior(prior) { name = "G L1-Norm"; }
  GL1Norm(float* prior, float epsilon_a, float epsilon_b) : prior(prior), epsilon_a(epsilon_a), epsilon_b(epsilon_b) { name = "G L1-Norm"; }
  explicit GL1Norm(const std::vector<float>& prior_host);
  ~GL1Norm() override;
  float getNormalizationFactor() const { return normalization_factor; }
  void setNormalizationFactor(float f) { normalization_factor = f; }
  void setPrior(float* p) override;
  void setEpsilonA(float e) { epsilon_a = e; }
  void setEpsilonB(float e) { epsilon_b = e; }
  void setEpsilons(float a, float b) { epsilon_a = a; epsilon_b = b; }
  void normalizePrior();
  float calcFi(float* p) override;
  void calcGi(float* p, float* xi) override;
  bool gradInto(float* p, float* dphi, bool first) override;

 private:
  float* prior = nullptr;
  float normalization_factor = 1.0f, epsilon_a = 1E-12, epsilon_b = 1E-12;
};

}  // namespace gpuvmem
