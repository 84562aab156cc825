// fis.cpp — Fi adapters over the C ABI (see fi.hpp).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "fi.hpp"

namespace gpuvmem {

Fi::~Fi() {
  if (G().engine) {
    devFree(device_S);
    devFree(device_DS);
  }
}

void Fi::configure(int penalizatorIndex, int imageIndex_, int imageToAdd_, bool normalize_) {
  Globals& g = G();
  imageIndex = imageIndex_;
  imageToAdd = imageToAdd_;
  normalize = normalize_;
  if (imageIndex > g.image_count - 1 || imageToAdd > g.image_count - 1) {
    std::printf("There is no image for the provided index %s\n", name.c_str());
    std::exit(-1);
  }
  if (penalizatorIndex != -1) {
    if (penalizatorIndex < 0) {
      std::printf("invalid index for penalizator (%s)\n", name.c_str());
      std::exit(-1);
    } else if (penalizatorIndex > g.nPenalizators - 1) {
      penalization_factor = 0.0f;
    } else {
      penalization_factor = g.penalizators[penalizatorIndex];
    }
  }
  if (!device_DS) device_DS = devAllocFloats((size_t)g.M * g.N);
}

void Fi::restartDGi() {
  if (device_DS) devZero(device_DS, (size_t)G().M * G().N);
}
void Fi::addToDphi(float* device_dphi) {
  GVM_CHECK(gvm_add_to_dphi(G().engine, device_dphi, device_DS, imageToAdd));  // linkAddToDPhi
}
void Fi::setS(float* S) { devFree(device_S); device_S = S; }
void Fi::setDS(float* DS) { devFree(device_DS); device_DS = DS; }

float Fi::priorValue(int kind, float* p, const gvm_prior_params& pp) {
  float v = 0.0f;
  if (iteration > 0 && penalization_factor)
    GVM_CHECK(gvm_prior_value(G().engine, kind, p, imageIndex, &pp, &v));
  set_fivalue(v);
  return penalization_factor * v;
}
bool Fi::enqueueFi(float* p, int slot) {
  int kind = 0;
  gvm_prior_params pp;
  if (!priorSpec(&kind, &pp)) return false;
  enqueued_gate_closed = !(iteration > 0 && penalization_factor);
  if (!enqueued_gate_closed) GVM_CHECK(gvm_prior_value_to_slot(G().engine, kind, p, imageIndex, &pp, slot));
  return true;
}
namespace {
uint64_t mix(uint64_t h, const void* data, size_t n) {   // FNV-1a
  const unsigned char* b = static_cast<const unsigned char*>(data);
  for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}
template <typename T> uint64_t mixv(uint64_t h, T v) { return mix(h, &v, sizeof(T)); }
}  // namespace
uint64_t Fi::stateKey() {
  int kind = -1;
  gvm_prior_params pp;
  std::memset(&pp, 0, sizeof(pp));
  priorSpec(&kind, &pp);
  uint64_t h = mixv(1469598103934665603ull, kind);
  h = mixv(h, pp.prior_value); h = mixv(h, pp.eta); h = mixv(h, pp.epsilon); h = mixv(h, pp.epsilon_b);
  h = mixv(h, pp.prior_image_dev); h = mixv(h, imageIndex);
  return mixv(h, (int)(iteration > 0 && penalization_factor));
}
uint64_t Chi2::stateKey() {
  Globals& g = G();
  uint64_t h = mixv(1469598103934665603ull, 0x43686932);
  h = mixv(h, fg_scale); h = mixv(h, g.noise_cut); h = mixv(h, g.threshold); h = mixv(h, g.flag_opt);
  return mixv(h, (int)normalize);
}
bool Fi::gradInto(float* p, float* dphi, bool) {
  int kind = 0;
  gvm_prior_params pp;
  if (!priorSpec(&kind, &pp)) return false;
  if (iteration > 0 && penalization_factor && G().flag_opt % 2 == imageIndex)
    GVM_CHECK(gvm_prior_grad_add(G().engine, kind, p, imageIndex, &pp, penalization_factor, dphi, dphiImage()));
  return true;
}
bool Chi2::gradInto(float* p, float* dphi, bool first) {
  if (!first) return false;   // the overwrite of src/chi2.cu:60-70 would discard what earlier terms added
  Globals& g = G();
  GVM_CHECK(gvm_dchi2(g.engine, p, g.flag_opt, normalize ? 1 : 0, dphi));
  return true;
}
// the swapped-argument quirk (see calcGi): the gradient lands in the prior image, dphi receives nothing
bool GL1Norm::gradInto(float* p, float*, bool) {
  restartDGi();
  calcGi(p, nullptr);
  return true;
}
void Fi::noteEnqueued() { enqueued_gate_closed = !(iteration > 0 && penalization_factor); }
void Chi2::noteEnqueued() {
  Globals& g = G();
  enqueued_gate_closed = false;
  GVM_CHECK(gvm_set_scalars(g.engine, fg_scale, g.noise_cut, g.threshold));
  GVM_CHECK(gvm_set_flag_opt(g.engine, g.flag_opt));
}
float Fi::finishFi(float value) {
  if (enqueued_gate_closed) value = 0.0f;
  set_fivalue(value);
  return penalization_factor * value;
}
void Fi::priorGrad(int kind, float* p, const gvm_prior_params& pp) {
  if (iteration > 0 && penalization_factor && G().flag_opt % 2 == imageIndex)
    GVM_CHECK(gvm_prior_grad(G().engine, kind, p, imageIndex, &pp, penalization_factor, device_DS));
}

// ------------------------------------------------------------------ Chi2 --
Chi2::~Chi2() {
  if (G().engine) devFree(result_dchi2);
}
void Chi2::configure(int penalizatorIndex, int imageIndex_, int, bool normalize_) {
  Globals& g = G();
  imageIndex = imageIndex_;
  normalize = normalize_;
  if (penalizatorIndex != -1) {
    if (penalizatorIndex > g.nPenalizators - 1 || penalizatorIndex < 0) {
      std::printf("invalid index for penalizator (%s)\n", name.c_str());
      std::exit(-1);
    }
    penalization_factor = g.penalizators[penalizatorIndex];
  }
  if (!result_dchi2) result_dchi2 = devAllocFloats((size_t)g.M * g.N * g.image_count);
}
float Chi2::calcFi(float* p) {
  Globals& g = G();
  float v = 0.0f;
  GVM_CHECK(gvm_set_scalars(g.engine, fg_scale, g.noise_cut, g.threshold));
  GVM_CHECK(gvm_set_flag_opt(g.engine, g.flag_opt));
  GVM_CHECK(gvm_chi2(g.engine, p, normalize ? 1 : 0, &v));
  set_fivalue(v);
  return penalization_factor * v;
}
bool Chi2::enqueueFi(float* p, int slot) {
  Globals& g = G();
  enqueued_gate_closed = false;
  GVM_CHECK(gvm_set_scalars(g.engine, fg_scale, g.noise_cut, g.threshold));
  GVM_CHECK(gvm_set_flag_opt(g.engine, g.flag_opt));
  GVM_CHECK(gvm_chi2_to_slot(g.engine, p, normalize ? 1 : 0, slot));
  return true;
}
void Chi2::calcGi(float* p, float*) {
  Globals& g = G();
  GVM_CHECK(gvm_dchi2(g.engine, p, g.flag_opt, normalize ? 1 : 0, result_dchi2));
}
void Chi2::restartDGi() { devZero(result_dchi2, (size_t)G().M * G().N * G().image_count); }
// src/chi2.cu:60-70: with two images the chi2 gradient REPLACES dphi (priors are added after it)
void Chi2::addToDphi(float* device_dphi) {
  Globals& g = G();
  if (g.image_count == 1) GVM_CHECK(gvm_add_to_dphi(g.engine, device_dphi, result_dchi2, 0));
  if (g.image_count > 1) devCopyD2D(device_dphi, result_dchi2, (size_t)g.M * g.N * g.image_count);
}
void Chi2::setCKernel(CKernel* ck) {
  ckernel = ck;
  GVM_CHECK(gvm_set_gcf(G().engine, ck && ck->getGCF() ? ck->getGCFCPUPointer() : nullptr));
}

// ---------------------------------------------------------------- priors --
namespace {
gvm_prior_params params(float prior_value, float eta, float eps_a, float eps_b, const float* prior_image) {
  gvm_prior_params pp;
  pp.prior_value = prior_value; pp.eta = eta; pp.epsilon = eps_a; pp.epsilon_b = eps_b;
  pp.prior_image_dev = prior_image;
  return pp;
}
}  // namespace

bool Entropy::priorSpec(int* kind, gvm_prior_params* pp) { *kind = GVM_PRIOR_ENTROPY; *pp = params(prior_value, eta, 0, 0, nullptr); return true; }
float Entropy::calcFi(float* p) { return priorValue(GVM_PRIOR_ENTROPY, p, params(prior_value, eta, 0, 0, nullptr)); }
void Entropy::calcGi(float* p, float*) { priorGrad(GVM_PRIOR_ENTROPY, p, params(prior_value, eta, 0, 0, nullptr)); }

bool L1norm::priorSpec(int* kind, gvm_prior_params* pp) { *kind = GVM_PRIOR_L1; *pp = params(0, 0, epsilon, 0, nullptr); return true; }
float L1norm::calcFi(float* p) { return priorValue(GVM_PRIOR_L1, p, params(0, 0, epsilon, 0, nullptr)); }
void L1norm::calcGi(float* p, float*) { priorGrad(GVM_PRIOR_L1, p, params(0, 0, epsilon, 0, nullptr)); }

bool TVariation::priorSpec(int* kind, gvm_prior_params* pp) { *kind = GVM_PRIOR_TV; *pp = params(0, 0, epsilon, 0, nullptr); return true; }
float TVariation::calcFi(float* p) { return priorValue(GVM_PRIOR_TV, p, params(0, 0, epsilon, 0, nullptr)); }
void TVariation::calcGi(float* p, float*) { priorGrad(GVM_PRIOR_TV, p, params(0, 0, epsilon, 0, nullptr)); }
void TVariation::addToDphi(float* device_dphi) { GVM_CHECK(gvm_add_to_dphi(G().engine, device_dphi, device_DS, 0)); }

bool TSqVariation::priorSpec(int* kind, gvm_prior_params* pp) { *kind = GVM_PRIOR_TSV; *pp = params(0, 0, 0, 0, nullptr); return true; }
float TSqVariation::calcFi(float* p) { return priorValue(GVM_PRIOR_TSV, p, params(0, 0, 0, 0, nullptr)); }
void TSqVariation::calcGi(float* p, float*) { priorGrad(GVM_PRIOR_TSV, p, params(0, 0, 0, 0, nullptr)); }

bool Laplacian::priorSpec(int* kind, gvm_prior_params* pp) { *kind = GVM_PRIOR_LAPLACIAN; *pp = params(0, 0, 0, 0, nullptr); return true; }
float Laplacian::calcFi(float* p) { return priorValue(GVM_PRIOR_LAPLACIAN, p, params(0, 0, 0, 0, nullptr)); }
void Laplacian::calcGi(float* p, float*) { priorGrad(GVM_PRIOR_LAPLACIAN, p, params(0, 0, 0, 0, nullptr)); }

bool QuadraticP::priorSpec(int* kind, gvm_prior_params* pp) { *kind = GVM_PRIOR_QUADRATIC; *pp = params(0, 0, 0, 0, nullptr); return true; }
float QuadraticP::calcFi(float* p) { return priorValue(GVM_PRIOR_QUADRATIC, p, params(0, 0, 0, 0, nullptr)); }
void QuadraticP::calcGi(float* p, float*) { priorGrad(GVM_PRIOR_QUADRATIC, p, params(0, 0, 0, 0, nullptr)); }

namespace {
float* uploadImage(const std::vector<float>& host) {
  float* d = devAllocFloats(host.size());
  devUpload(d, host.data(), host.size());
  return d;
}
void scaleImage(float* img, float factor) {  // normalizeImage, src/functions.cu:4626
  GVM_CHECK(gvm_vec_scale(G().engine, img, 1.0f / factor, (int64_t)G().M * G().N));
}
}  // namespace

GEntropy::GEntropy(const std::vector<float>& prior_host) : prior(uploadImage(prior_host)) { name = "GEntropy"; }
GEntropy::~GEntropy() { if (G().engine) devFree(prior); }
void GEntropy::setPrior(float* p) { devFree(prior); prior = p; }
void GEntropy::normalizePrior() { scaleImage(prior, normalization_factor); }
bool GEntropy::priorSpec(int* kind, gvm_prior_params* pp) { *kind = GVM_PRIOR_GENTROPY; *pp = params(0, eta, 0, 0, prior); return true; }
float GEntropy::calcFi(float* p) { return priorValue(GVM_PRIOR_GENTROPY, p, params(0, eta, 0, 0, prior)); }
void GEntropy::calcGi(float* p, float*) { priorGrad(GVM_PRIOR_GENTROPY, p, params(0, eta, 0, 0, prior)); }

GL1Norm::GL1Norm(const std::vector<float>& prior_host) : prior(uploadImage(prior_host)) { name = "G L1-Norm"; }
GL1Norm::~GL1Norm() { if (G().engine) devFree(prior); }
void GL1Norm::setPrior(float* p) { devFree(prior); prior = p; }
void GL1Norm::normalizePrior() { scaleImage(prior, normalization_factor); }
bool GL1Norm::priorSpec(int* kind, gvm_prior_params* pp) { *kind = GVM_PRIOR_GL1; *pp = params(0, 0, epsilon_a, epsilon_b, prior); return true; }
float GL1Norm::calcFi(float* p) { return priorValue(GVM_PRIOR_GL1, p, params(0, 0, epsilon_a, epsilon_b, prior)); }
// Reference quirk kept verbatim: GL1Norm::calcGi calls DGL1Norm(p, device_DS, this->prior, ...)
// (src/gl1norm.cu:145-148) against the signature DGL1Norm(I, prior, dgi, ...) (src/functions.cu:4700-4709) — the two
// image arguments are swapped. The kernel therefore reads the device_DS that restartDGi has just zeroed AS THE PRIOR
// and writes the gradient INTO THE PRIOR IMAGE: device_DS stays zero, the term adds nothing to dphi, and every later
// calcFi divides by the overwritten prior. Pinned against the reference build by
// tests/test_parity_reference_gpu.py::test_host_fi_terms_match_reference.
void GL1Norm::calcGi(float* p, float*) {
  if (!(iteration > 0 && penalization_factor && G().flag_opt % 2 == imageIndex)) return;
  const gvm_prior_params pp = params(0, 0, epsilon_a, epsilon_b, device_DS);
  GVM_CHECK(gvm_prior_grad(G().engine, GVM_PRIOR_GL1, p, imageIndex, &pp, penalization_factor, prior));
}

namespace {
Fi* makeChi2() { return new Chi2; }
Fi* makeEntropy() { return new Entropy; }
Fi* makeL1() { return new L1norm; }
Fi* makeTV() { return new TVariation; }
Fi* makeTSV() { return new TSqVariation; }
Fi* makeLaplacian() { return new Laplacian; }
Fi* makeQuadratic() { return new QuadraticP; }
Fi* makeGEntropy() { return new GEntropy; }
Fi* makeGL1() { return new GL1Norm; }
const bool kRegistered[] = {
    registerCreationFunction<Fi, std::string>("Chi2", makeChi2),
    registerCreationFunction<Fi, std::string>("Entropy", makeEntropy),
    registerCreationFunction<Fi, int>(0, makeEntropy),  // src/entropy.cu:78-79
    registerCreationFunction<Fi, std::string>("L1-Norm", makeL1),
    registerCreationFunction<Fi, std::string>("TotalVariation", makeTV),
    registerCreationFunction<Fi, std::string>("TotalSquaredVariation", makeTSV),
    registerCreationFunction<Fi, std::string>("Laplacian", makeLaplacian),
    registerCreationFunction<Fi, std::string>("Quadratic", makeQuadratic),
    registerCreationFunction<Fi, std::string>("GEntropy", makeGEntropy),
    registerCreationFunction<Fi, std::string>("GL1Norm", makeGL1),
};
}  // namespace

}  // namespace gpuvmem
