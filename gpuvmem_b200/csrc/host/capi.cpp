// capi.cpp — C entry points over the host classes (include/gvm_host.h).
#include <cstdio>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "../../../include/gvm_host.h"
#include "fits.hpp"
#include "synthesizer.hpp"

using namespace gpuvmem;

struct gvmh_session {
  Synthesizer* sy = nullptr;
  MFS* mfs = nullptr;
  Optimizer* opt = nullptr;
  ObjectiveFunction* of = nullptr;
  CKernel* ck = nullptr;
  WeightingScheme* scheme = nullptr;
  std::vector<Fi*> terms;
  float* xi = nullptr;        // [2][M][N] gradient of the last gvmh_calc_gradient / eval
  float* image_stage = nullptr;
  bool of_configured = false;
};

namespace {
bool g_quiet_all = false;  // gvmh_set_quiet: library callers that own stdout (bench.py prints ONE JSON line)
std::vector<std::string> splitArgs(const std::string& s) {
  std::vector<std::string> out;
  std::istringstream is(s);
  std::string tok;
  while (is >> tok) out.push_back(tok);
  return out;
}
void defaultOrder(Optimizer* optimizer, Image* image) {  // optimizationOrder, src/main.cu:87-98
  optimizer->setImage(image);
  optimizer->setFlag(0);
  optimizer->optimize();
}
CKernel* makeCKernel(const char* name, int m, int n) {
  const std::string id = name ? name : "PillBox2D";
  if (m <= 0 || n <= 0) return createObject<CKernel, std::string>(id);
  if (id == "PillBox2D") return new PillBox2D(m, n);
  if (id == "Gaussian2D") return new Gaussian2D(m, n);
  if (id == "Sinc2D") return new Sinc2D(m, n);
  if (id == "GaussianSinc2D") return new GaussianSinc2D(m, n);
  if (id == "PSWF") return new PSWF_12D(m, n);
  return createObject<CKernel, std::string>(id);
}
void ensureObjective(gvmh_session* s) {
  if (s->of_configured) return;
  Globals& g = G();
  s->of->configure(g.N, g.M, g.image_count);
  s->of_configured = true;
}
size_t imageFloats() { return (size_t)G().M * G().N * G().image_count; }
}  // namespace

extern "C" {

int gvmh_create(const gvmh_problem* p, const char* args, const char* optimizer, const char* scheme,
                const char* ckernel, int ck_m, int ck_n, const char* fi_spec, int rank, int world,
                const char* nccl_id, gvmh_session** out) {
  if (!out) return 1;   // p == NULL: the datasets and the header come from the -i / -m files named in `args`
  {  // library callers get an error code instead of the reference's print + exit when no B200 is present
    gvm_config probe;
    std::memset(&probe, 0, sizeof(probe));
    probe.M = probe.N = 16; probe.DELTAX = -1e-5; probe.DELTAY = 1e-5; probe.nu_0 = 1e11f; probe.eta = -1.0f;
    probe.fg_scale = 1.0f; probe.noise_cut = 1e30f;
    int dev = 0;
    if (args) {
      const char* gflag = std::strstr(args, "-G ");
      if (gflag) dev = std::atoi(gflag + 3);
    }
    probe.device = dev;
    gvm_engine* e = nullptr;
    if (gvm_create(&probe, &e) != 0) return 2;
    gvm_destroy(e);
  }
  G() = Globals();

  gvmh_session* s = new gvmh_session();
  s->sy = createObject<Synthesizer, std::string>("MFS");
  s->mfs = static_cast<MFS*>(s->sy);
  s->opt = createObject<Optimizer, std::string>(optimizer ? optimizer : "CG-FRPRMN");
  s->ck = makeCKernel(ckernel, ck_m, ck_n);
  s->of = createObject<ObjectiveFunction, std::string>("ObjectiveFunction");
  s->scheme = createObject<WeightingScheme, std::string>(scheme ? scheme : "Natural");
  Io* ioms = createObject<Io, std::string>("IoMS");
  Io* iofits = createObject<Io, std::string>("IoFITS");

  if (p) {
    std::vector<MSDataset> ds(1);
    const double fra = p->has_field_centre ? p->field_ra : p->ra, fdec = p->has_field_centre ? p->field_dec : p->dec;
    fillDataset(&ds[0], p->telescope ? p->telescope : "ALMA", p->antenna_diameter, fra * 3.14159265358979323846 / 180.0,
                fdec * 3.14159265358979323846 / 180.0, p->nchan, p->freqs, p->Z, p->uvw_m, p->Vo, p->w);
    ds[0].name = "memory";
    ds[0].oname = "NULL";
    headerValues h;
    h.M = p->M; h.N = p->N; h.DELTAX = p->DELTAX; h.DELTAY = p->DELTAY; h.ra = p->ra; h.dec = p->dec;
    h.crpix1 = p->crpix1; h.crpix2 = p->crpix2; h.beam_noise = p->beam_noise;
    s->mfs->adoptDatasets(std::move(ds), h);
  }
  s->mfs->setDistributed(rank, world, nccl_id ? std::string(nccl_id, GVM_DIST_ID_BYTES) : std::string());
  if (g_quiet_all) G().quiet = true;

  s->sy->setIoVisibilitiesHandler(ioms);
  s->sy->setIoImageHandler(iofits);
  s->sy->setOrder(&defaultOrder);
  s->sy->setWeightingScheme(s->scheme);
  s->sy->setGriddingKernel(s->ck);
  s->sy->setOptimizator(s->opt);

  std::vector<std::string> toks = splitArgs(args ? args : "");
  std::vector<char*> argv;
  std::string prog = "gpuvmem_b200";
  argv.push_back(&prog[0]);
  for (std::string& t : toks) argv.push_back(&t[0]);
  s->sy->configure((int)argv.size(), argv.data());
  s->opt->setObjectiveFunction(s->of);
  s->sy->setDevice();

  // the Fi terms, main.cu:168-190
  const std::string spec = fi_spec && *fi_spec
                               ? fi_spec
                               : "Chi2:-1:0:0,Entropy:0:0:0,L1-Norm:1:0:0,TotalSquaredVariation:2:0:0,Laplacian:3:0:0";
  std::istringstream is(spec);
  std::string item;
  while (std::getline(is, item, ',')) {
    std::istringstream it(item);
    std::string name, a, b, c, nrm;
    std::getline(it, name, ':');
    std::getline(it, a, ':');
    std::getline(it, b, ':');
    std::getline(it, c, ':');
    std::getline(it, nrm, ':');   // optional 5th field: Fi::configure's `normalize`
    Fi* fi = createObject<Fi, std::string>(name);
    fi->configure(a.empty() ? -1 : std::stoi(a), b.empty() ? 0 : std::stoi(b), c.empty() ? 0 : std::stoi(c),
                  !nrm.empty() && std::stoi(nrm) != 0);
    if (name == "Entropy") fi->setPrior(0.001f);  // main.cu:177
    s->of->addFi(fi);
    s->terms.push_back(fi);
  }
  s->xi = devAllocFloats(imageFloats());
  *out = s;
  return 0;
}

int gvmh_set_quiet(int quiet) {
  g_quiet_all = quiet != 0;
  G().quiet = G().quiet || g_quiet_all;
  return 0;
}

int gvmh_destroy(gvmh_session* s) {
  if (!s) return 0;
  if (G().engine) {
    devFree(s->xi);
    devFree(s->image_stage);
  }
  for (Fi* f : s->terms) delete f;
  delete s->of;
  delete s->opt;
  s->sy->unSetDevice();
  delete s->ck;
  delete s->scheme;
  delete s->sy;
  delete s;
  return 0;
}

int gvmh_run(gvmh_session* s, float* image_out, double* optimize_seconds) {
  // the optimizers configure the objective function on first use (src/frprmn.cu:93-96); if the
  // session already evaluated it directly, keep that allocation
  s->sy->run();
  s->of_configured = true;
  if (image_out) devDownload(image_out, s->sy->getImage()->getImage(), imageFloats());
  if (optimize_seconds) *optimize_seconds = s->mfs->derived().run_seconds;
  return 0;
}
int gvmh_clear_run(gvmh_session* s) { s->sy->clearRun(); return 0; }
int gvmh_set_lbfgs_k(gvmh_session* s, int k) { s->opt->setK(k); return 0; }
int gvmh_write_outputs(gvmh_session* s) {
  s->sy->writeImages();
  s->sy->writeResiduals();
  return 0;
}

int gvmh_use_ckernel_degridding(gvmh_session* s, int on) {
  s->mfs->useCKernelDegridding(on != 0);
  return 0;
}
int gvmh_write_residuals(gvmh_session* s, float* nongridded_chi2) {
  s->sy->writeResiduals();
  if (nongridded_chi2) *nongridded_chi2 = s->mfs->getNonGriddedChi2();
  return 0;
}
int gvmh_get_host_model(gvmh_session* s, int chan, float* Vm, float* Vr) {
  Field& f = s->mfs->getDatasets()[0].fields[0];
  if (chan < 0 || chan >= (int)f.visibilities.size()) return 1;
  const HVis& v = f.visibilities[chan][0];
  if (Vm) std::memcpy(Vm, v.Vm.data(), v.Vm.size() * sizeof(float));
  if (Vr) std::memcpy(Vr, v.Vr.data(), v.Vr.size() * sizeof(float));
  return 0;
}

int gvmh_fi_eval(gvmh_session*, const char* name, const float* I_host, const float* prior_host, float lambda,
                 float prior_value, float eta, float eps_a, float eps_b, int image_index, int iteration, int flag,
                 float* value_out, float* dphi_out, float* prior_after_out) {
  Globals& g = G();
  Fi* f = createObject<Fi, std::string>(name);
  if (!f) return 1;
  const std::string n(name);
  const size_t MN = (size_t)g.M * g.N;
  const int saved_flag = g.flag_opt;
  g.flag_opt = flag;
  f->configure(-1, image_index, image_index, false);
  f->setPenalizationFactor(lambda);
  f->setIteration(iteration);
  float* d_prior = nullptr;
  if (prior_host) {
    d_prior = devAllocFloats(MN);
    devUpload(d_prior, prior_host, MN);
    f->setPrior(d_prior);    // owned by the term from here on
  }
  if (n == "Entropy") { f->setPrior(prior_value); f->setEta(eta); }
  if (n == "GEntropy") f->setEta(eta);
  if (n == "TotalVariation") static_cast<TVariation*>(f)->setEpsilon(eps_a);
  if (n == "L1-Norm") static_cast<L1norm*>(f)->setEpsilon(eps_a);
  if (n == "GL1Norm") static_cast<GL1Norm*>(f)->setEpsilons(eps_a, eps_b);
  float* d_I = devAllocFloats(imageFloats());
  float* d_phi = devAllocFloats(imageFloats());
  devUpload(d_I, I_host, imageFloats());
  f->calcFi(d_I);
  if (value_out) *value_out = f->get_fivalue();
  f->restartDGi();
  f->calcGi(d_I, d_phi);
  f->addToDphi(d_phi);
  if (dphi_out) devDownload(dphi_out, d_phi, imageFloats());
  if (prior_after_out && d_prior) devDownload(prior_after_out, d_prior, MN);
  devFree(d_I);
  devFree(d_phi);
  delete f;
  g.flag_opt = saved_flag;
  return 0;
}

int gvmh_error_image(gvmh_session* s, float* errors_host) {
  Error* est = createObject<Error, std::string>("SecondDerivateError");
  Image* img = s->sy->getImage();
  est->calculateErrorImage(img, nullptr);
  devDownload(errors_host, img->getErrorImage(), imageFloats());
  delete est;
  return 0;
}

int gvmh_set_image(gvmh_session* s, const float* I_host) {
  devUpload(s->sy->getImage()->getImage(), I_host, imageFloats());
  return 0;
}
int gvmh_get_image(gvmh_session* s, float* I_host) {
  devDownload(I_host, s->sy->getImage()->getImage(), imageFloats());
  return 0;
}
int gvmh_set_iteration(gvmh_session* s, int iteration) {
  for (Fi* f : s->of->getFi()) f->setIteration(iteration);
  return 0;
}
int gvmh_set_flag(gvmh_session*, int flag_opt) { G().flag_opt = flag_opt; return 0; }

int gvmh_calc_function(gvmh_session* s, float* value, float* fi_values, int nfi) {
  ensureObjective(s);
  Fi* chi2 = s->of->getFiByName("Chi2");
  if (chi2) chi2->setFgScale(s->sy->getFgScale());
  const float v = s->of->calcFunction(s->sy->getImage()->getImage());
  if (value) *value = v;
  const std::vector<float> fv = s->of->get_fi_values();
  for (int i = 0; i < nfi && i < (int)fv.size(); i++) fi_values[i] = fv[i];
  return 0;
}
int gvmh_calc_gradient(gvmh_session* s, int iteration, float* grad_host) {
  ensureObjective(s);
  s->of->calcGradient(s->sy->getImage()->getImage(), s->xi, iteration);
  if (grad_host) devDownload(grad_host, s->xi, imageFloats());
  return 0;
}
int gvmh_eval_device(gvmh_session* s, int iteration, float* value) {
  ensureObjective(s);
  Fi* chi2 = s->of->getFiByName("Chi2");
  if (chi2) chi2->setFgScale(s->sy->getFgScale());
  float* I = s->sy->getImage()->getImage();
  const float v = s->of->calcFunction(I);
  s->of->calcGradient(I, s->xi, iteration);
  if (value) *value = v;
  return 0;
}
int gvmh_eval_host(gvmh_session* s, const float* I_host, int iteration, float* value, float* grad_host) {
  ensureObjective(s);
  if (!s->image_stage) s->image_stage = devAllocFloats(imageFloats());
  Fi* chi2 = s->of->getFiByName("Chi2");
  if (chi2) chi2->setFgScale(s->sy->getFgScale());
  // Multi-rank: the image crosses PCIe ONCE (rank 0) and reaches the other replicas over NVLink
  // (ncclBroadcast on the engine stream); the gradient, identical on every rank after the all-reduce,
  // goes back to the host from rank 0 only. I_host / grad_host are not touched on the other ranks.
  if (G().rank == 0) devUpload(s->image_stage, I_host, imageFloats());
  if (G().world > 1) GVM_CHECK(gvm_dist_broadcast(G().engine, s->image_stage, (int64_t)imageFloats(), 0));
  const float v = s->of->calcFunction(s->image_stage);
  s->of->calcGradient(s->image_stage, s->xi, iteration);
  if (G().rank == 0) devDownload(grad_host, s->xi, imageFloats());
  if (value) *value = v;
  return 0;
}

gvm_engine* gvmh_engine(gvmh_session*) { return G().engine; }

int gvmh_scalars(gvmh_session* s, double* out) {
  const MFS::Derived& d = s->mfs->derived();
  out[0] = d.fg_scale; out[1] = d.noise_cut; out[2] = d.noise_jypix; out[3] = d.nu_0; out[4] = d.vis_noise;
  out[5] = d.sum_weights; out[6] = d.beam_bmaj_deg; out[7] = d.beam_bmin_deg; out[8] = d.beam_bpa_deg;
  out[9] = d.deltau; out[10] = d.deltav; out[11] = d.xobs_pix; out[12] = d.yobs_pix;
  out[13] = d.total_visibilities; out[14] = s->opt->getCurrentIteration(); out[15] = (double)s->of->getFi().size();
  return 0;
}
int gvmh_stats(gvmh_session* s, double* sec /* [6] */, int64_t* counts /* [2] */) {
  const MFS::Derived& d = s->mfs->derived();
  if (sec) {
    sec[0] = d.setup_seconds; sec[1] = d.weighting_seconds; sec[2] = d.gridding_seconds; sec[3] = d.run_seconds;
    sec[4] = s->of->functionSeconds(); sec[5] = s->of->gradientSeconds();
  }
  if (counts) { counts[0] = s->of->functionEvaluations(); counts[1] = s->of->gradientEvaluations(); }
  return 0;
}
int64_t gvmh_nvis(gvmh_session* s, int chan) {
  Field& f = s->mfs->getDatasets()[0].fields[0];
  if (chan < 0 || chan >= (int)f.visibilities.size()) return -1;
  return (int64_t)f.visibilities[chan][0].size();
}
int gvmh_get_host_vis(gvmh_session* s, int chan, double* uvw_m, float* Vo, float* w) {
  Field& f = s->mfs->getDatasets()[0].fields[0];
  if (chan < 0 || chan >= (int)f.visibilities.size()) return 1;
  const HVis& v = f.visibilities[chan][0];
  if (uvw_m) std::memcpy(uvw_m, v.uvw.data(), v.uvw.size() * sizeof(double));
  if (Vo) std::memcpy(Vo, v.Vo.data(), v.Vo.size() * sizeof(float));
  if (w) std::memcpy(w, v.weight.data(), v.weight.size() * sizeof(float));
  return 0;
}
int gvmh_filter_gridding(gvmh_session* s, const char* ckernel, int ck_m, int ck_n) {
  Filter* f = createObject<Filter, std::string>("Gridding");
  CKernel* ck = ckernel && *ckernel ? makeCKernel(ckernel, ck_m, ck_n) : nullptr;
  static_cast<Gridding*>(f)->setCKernel(ck);
  f->applyCriteria(s->sy->getVisibilities());
  delete f;
  delete ck;
  return 0;
}
const char* gvmh_exit_reason(gvmh_session* s) { return s->opt->getExitReason(); }
int gvmh_history(gvmh_session* s, float* out, int cap) {
  const std::vector<float>& h = s->opt->getHistory();
  const int n = std::min<int>(cap, (int)h.size());
  for (int i = 0; i < n; i++) out[i] = h[i];
  return (int)h.size();
}

// ---------------------------------------------------------------- stateless --
int gvmh_ckernel_table(const char* name, int m, int n, float sx, float sy, float w, float* table, int* support_x,
                       int* support_y) {
  CKernel* ck = makeCKernel(name, m, n);
  ck->setSigmas(sx, sy);
  if (w > 0.0f) ck->setW(w);
  ck->buildKernel();
  std::memcpy(table, ck->getKernelPointer(), sizeof(float) * ck->getm() * ck->getn());
  if (support_x) *support_x = ck->getSupportX();
  if (support_y) *support_y = ck->getSupportY();
  delete ck;
  return 0;
}
int gvmh_ckernel_gcf(const char* name, int m, int n, int M, int N, float dx, float dy, float* gcf) {
  CKernel* ck = makeCKernel(name, m, n);
  ck->initializeGCF(M, N, dx, dy);
  std::memcpy(gcf, ck->getGCFCPUPointer(), sizeof(float) * (size_t)M * N);
  delete ck;
  return 0;
}
int gvmh_factory_has(const char* kind, const char* name) {
  const std::string k = kind, id = name;
  if (k == "Fi") return Singleton<Factory<Fi, std::string>>::Instance().Has(id);
  if (k == "Optimizer") return Singleton<Factory<Optimizer, std::string>>::Instance().Has(id);
  if (k == "CKernel") return Singleton<Factory<CKernel, std::string>>::Instance().Has(id);
  if (k == "WeightingScheme") return Singleton<Factory<WeightingScheme, std::string>>::Instance().Has(id);
  if (k == "Synthesizer") return Singleton<Factory<Synthesizer, std::string>>::Instance().Has(id);
  if (k == "Filter") return Singleton<Factory<Filter, std::string>>::Instance().Has(id);
  if (k == "Error") return Singleton<Factory<Error, std::string>>::Instance().Has(id);
  if (k == "Io") return Singleton<Factory<Io, std::string>>::Instance().Has(id);
  if (k == "ObjectiveFunction") return Singleton<Factory<ObjectiveFunction, std::string>>::Instance().Has(id);
  return 0;
}
int gvmh_fits_read(const char* path, double* header16, float* data_out, int64_t cap) {
  FitsImage img;
  std::string err;
  if (!fitsRead(path, data_out != nullptr, &img, &err)) { std::fprintf(stderr, "gvmh_fits_read: %s\n", err.c_str()); return 1; }
  headerValues h;
  const bool wcs = fitsHeaderValues(img, &h, &err);
  if (header16) {
    const double vals[16] = {(double)img.naxis1, (double)img.naxis2, (double)img.bitpix, wcs ? 1.0 : 0.0, h.DELTAX, h.DELTAY,
                             h.ra, h.dec, h.crpix1, h.crpix2, h.beam_bmaj, h.beam_bmin, h.beam_bpa, (double)h.beam_noise,
                             (double)h.equinox, (double)img.cards.size()};
    for (int i = 0; i < 16; i++) header16[i] = vals[i];
  }
  if (data_out) {
    if ((int64_t)img.data.size() > cap) return 2;
    std::memcpy(data_out, img.data.data(), img.data.size() * sizeof(float));
  }
  return 0;
}
int gvmh_fits_write(const char* path, const float* data, int64_t naxis1, int64_t naxis2, const char* template_path,
                    const char* bunit, int niter, const char* radesys, float equinox, double crval1, double crval2) {
  FitsImage tmpl;
  std::string err;
  if (template_path && *template_path && !fitsRead(template_path, false, &tmpl, &err)) {
    std::fprintf(stderr, "gvmh_fits_write: %s\n", err.c_str());
    return 1;
  }
  if (!fitsWriteFloat(path, data, (long)naxis1, (long)naxis2, tmpl.cards, bunit ? bunit : "", niter,
                      radesys ? radesys : "ICRS", equinox, crval1, crval2, &err)) {
    std::fprintf(stderr, "gvmh_fits_write: %s\n", err.c_str());
    return 1;
  }
  return 0;
}

int gvmh_parse_args(const char* args, char* json_out, size_t cap) {
  G() = Globals();
  std::vector<std::string> toks = splitArgs(args ? args : "");
  std::vector<char*> argv;
  std::string prog = "gpuvmem_b200";
  argv.push_back(&prog[0]);
  for (std::string& t : toks) argv.push_back(&t[0]);
  Vars v;
  const bool ok = getOptions((int)argv.size(), argv.data(), &v);
  Globals& g = G();
  std::snprintf(json_out, cap,
                "{\"ok\": %s, \"input\": \"%s\", \"output\": \"%s\", \"output_image\": \"%s\", \"modin\": \"%s\", "
                "\"path\": \"%s\", \"gpus\": \"%s\", \"ofile\": \"%s\", \"initial_values\": \"%s\", "
                "\"penalization_factors\": \"%s\", \"noise\": %g, \"eta\": %g, \"noise_cut\": %g, \"nu_0\": %g, "
                "\"threshold\": %g, \"randoms\": %g, \"robust_param\": %g, \"it_max\": %d, \"gridding\": %d, "
                "\"blockSizeX\": %d, \"blockSizeY\": %d, \"blockSizeV\": %d, \"verbose\": %d, \"nopositivity\": %d, "
                "\"print_images\": %d, \"modify_weights\": %d}",
                ok ? "true" : "false", v.input.c_str(), v.output.c_str(), v.output_image.c_str(), v.modin.c_str(),
                v.path.c_str(), v.gpus.c_str(), v.ofile.c_str(), v.initial_values.c_str(),
                v.penalization_factors.c_str(), v.noise, v.eta, v.noise_cut, v.nu_0, v.threshold, v.randoms,
                v.robust_param, v.it_max, v.gridding, v.blockSizeX, v.blockSizeY, v.blockSizeV, g.verbose_flag,
                g.nopositivity ? 1 : 0, g.print_images ? 1 : 0, g.modify_weights ? 1 : 0);
  return ok ? 0 : 1;
}
int gvmh_linmin_1d(gvmh_fn1d f, void* user, float* xmin, float* fmin, int* probes) {
  LineSearch ls(nullptr, nullptr);
  ls.probe_override = [=](float x) { return f(x, user); };
  float xm = 0.0f;
  const float fm = ls.minimize(&xm);
  if (xmin) *xmin = xm;
  if (fmin) *fmin = fm;
  if (probes) *probes = (int)ls.probes;
  return 0;
}
int gvmh_shard_plan(int nchan, const int64_t* Z, int world, int rank, int64_t* lo, int64_t* hi) {
  for (int c = 0; c < nchan; c++) {
    size_t a = 0, b = 0;
    if (shardRange(nchan, c, (size_t)Z[c], rank, world, &a, &b)) { lo[c] = (int64_t)a; hi[c] = (int64_t)b; }
    else { lo[c] = hi[c] = 0; }
  }
  return 0;
}
int gvmh_read_gvms(const char* path, double* out) {
  MSDataset ds;
  headerValues h;
  std::string err;
  if (!readGVMS(path, &ds, &h, &err)) {
    std::fprintf(stderr, "%s\n", err.c_str());
    return 1;
  }
  long total = 0;
  for (Field& f : ds.fields)
    for (auto& n : f.numVisibilitiesPerFreq) total += n;
  out[0] = (double)h.M; out[1] = (double)h.N; out[2] = ds.data.total_frequencies; out[3] = (double)total;
  out[4] = ds.data.min_freq; out[5] = ds.data.max_freq; out[6] = ds.data.max_blength; out[7] = ds.data.uvmax_wavelength;
  return 0;
}

}  // extern "C"
