/* oracle/ref_harness.cu — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Builds, together with the UNMODIFIED sources under /root/reference/src (all of
 * them except MSFITSIO.cu and main.cu) and the stub headers in oracle/ref_shim/,
 * the shared library oracle/_ref/libgvref.so.  It gives the parity tests and the
 * bench's reference arm a way to run the reference's own code on synthetic
 * input:
 *   - the reference's CUDA hot path (chi2(), dchi2(), the prior hosts, the
 *     optimizers) exactly as `main.cu` wires it (src/main.cu:147-212), and
 *   - the reference's host-only code (WeightingScheme::apply, do_gridding,
 *     CKernel::buildKernel/GCF), which runs without a GPU.
 *
 * What this file supplies is the link contract SURVEY.md §8c lists: a synthetic
 * stand-in for src/MSFITSIO.cu (casacore/cfitsio are absent offline).  The three
 * unit-conversion helpers restate src/MSFITSIO.cu:36-51; readMS fills the
 * in-memory layout src/MSFITSIO.cu:398-754 produces; FITS writers are no-ops.
 * Nothing here is copied from the reference; the reference sources are compiled
 * where they lie.
 */
#include <omp.h>
#include <cstdarg>
#include <sstream>

#include "framework.cuh"
#include "functions.cuh"
#include "mfs.cuh"
#include "pillBox2D.cuh"
#include "gaussian2D.cuh"
#include "gaussianSinc2D.cuh"
#include "sinc2D.cuh"
#include "pswf_12D.cuh"
#include "imageProcessor.cuh"
#include "totalvariation.cuh"
#include "gl1norm.cuh"
#include "gentropy.cuh"

int num_gpus;  // defined in the reference's main.cu, which is not compiled here
int gvref_tolerate_cuda_errors = 0;  // read by oracle/ref_shim/helper_cuda.h

extern long M, N;
extern double DELTAX, DELTAY, deltau, deltav, ra, dec, crpix1, crpix2;
extern float noise_cut, noise_jypix, eta, nu_0, threshold;
extern float* device_noise_image;
extern std::vector<MSDataset> datasets;
extern int nMeasurementSets, image_count, flag_opt, firstgpu, max_number_vis;
extern double beam_bmaj, beam_bmin, beam_bpa;
extern bool verbose_flag;

/* ---------------------------------------------------------------- MSFITSIO --
 * Restated from src/MSFITSIO.cu:36-51: fp32 wavelength, fp64 division. */
__host__ __device__ float freq_to_wavelength(float freq) {
  return LIGHTSPEED / freq;
}
__host__ __device__ double metres_to_lambda(double uvw_metres, float freq) {
  float lambda = freq_to_wavelength(freq);
  return uvw_metres / lambda;
}
__host__ __device__ float distance(float x, float y, float x0, float y0) {
  float s = (x - x0) * (x - x0) + (y - y0) * (y - y0);
  return sqrtf(s);
}

namespace {
struct SynthChannel {
  std::vector<double3> uvw;
  std::vector<cufftComplex> Vo;
  std::vector<float> w;
};
struct SynthProblem {
  headerValues hdr;
  std::string telescope;
  float dish = 12.0f;
  std::vector<float> freqs;
  std::vector<SynthChannel> chans;
  double max_uv_m = 0.0;
  bool has_field_centre = false;   // pointing / phase centre of the field when it is not the image centre
  double field_ra = 0.0, field_dec = 0.0;   // degrees
} g_prob;

Synthesizer* g_sy = nullptr;
Optimizer* g_opt = nullptr;
ObjectiveFunction* g_of = nullptr;
CKernel* g_ck = nullptr;
Io* g_iofits = nullptr;
Io* g_ioms = nullptr;
WeightingScheme* g_scheme = nullptr;
std::vector<Fi*> g_fis;
bool g_gridding = false;

void fill_dataset(std::vector<MSAntenna>& antennas, std::vector<Field>& fields,
                  MSData* data, int gridding) {
  const int nchan = (int)g_prob.freqs.size();
  data->telescope_name = g_prob.telescope;
  data->nantennas = 1;
  data->nbaselines = 0;
  data->nfields = 1;
  data->nstokes = 1;
  data->corr_type.assign(1, (int)XX);
  data->n_internal_frequencies = 1;
  data->n_internal_frequencies_ids.assign(1, 0);
  data->channels.assign(1, nchan);
  data->total_frequencies = nchan;
  float fmin = g_prob.freqs[0], fmax = g_prob.freqs[0];
  for (float f : g_prob.freqs) { fmin = std::min(fmin, f); fmax = std::max(fmax, f); }
  data->min_freq = fmin;
  data->max_freq = fmax;
  data->ref_freq = 0.5f * (fmin + fmax);
  data->max_blength = (float)g_prob.max_uv_m;
  data->min_blength = 0.0f;
  data->uvmax_wavelength = g_prob.max_uv_m * data->max_freq / LIGHTSPEED;

  /* src/MSFITSIO.cu:510, 535-551: per-telescope beam model */
  float max_wavelength = freq_to_wavelength(data->min_freq);
  antennas.clear();
  antennas.push_back(MSAntenna());
  antennas[0].antenna_id = "A0";
  antennas[0].station = "S0";
  antennas[0].position = {0.0, 0.0, 0.0};
  antennas[0].antenna_diameter = g_prob.dish;
  if (g_prob.telescope == "ALMA") {
    antennas[0].pb_factor = 1.13f;
    antennas[0].primary_beam = AIRYDISK;
  } else if (g_prob.telescope == "EVLA") {
    antennas[0].pb_factor = 1.25f;
    antennas[0].primary_beam = GAUSSIAN;
  } else {
    antennas[0].pb_factor = 3.8317059702075125f / PI; /* cyl_bessel_j_zero(1,1)/pi */
    antennas[0].primary_beam = GAUSSIAN;
  }
  antennas[0].pb_cutoff =
      antennas[0].pb_factor * (max_wavelength / antennas[0].antenna_diameter);

  fields.clear();
  fields.push_back(Field());
  Field& F = fields[0];
  F.id = 0;
  F.ref_ra = F.phs_ra = (g_prob.has_field_centre ? g_prob.field_ra : g_prob.hdr.ra) * (PI_D / 180.0);
  F.ref_dec = F.phs_dec = (g_prob.has_field_centre ? g_prob.field_dec : g_prob.hdr.dec) * (PI_D / 180.0);
  F.nu = g_prob.freqs;
  F.visibilities.resize(nchan, std::vector<HVis>(1, HVis()));
  F.device_visibilities.resize(nchan, std::vector<DVis>(1, DVis()));
  F.numVisibilitiesPerFreqPerStoke.resize(nchan, std::vector<long>(1, 0));
  F.numVisibilitiesPerFreq.resize(nchan, 0);
  F.backup_visibilities.resize(nchan, std::vector<HVis>(1, HVis()));
  if (gridding) {
    F.backup_numVisibilitiesPerFreqPerStoke.resize(nchan, std::vector<long>(1, 0));
    F.backup_numVisibilitiesPerFreq.resize(nchan, 0);
  }
  int mx = 0;
  for (int i = 0; i < nchan; i++) {
    HVis& h = F.visibilities[i][0];
    h.uvw = g_prob.chans[i].uvw;
    h.Vo = g_prob.chans[i].Vo;
    h.weight = g_prob.chans[i].w;
    h.Vm.assign(h.Vo.size(), make_cuFloatComplex(0.0f, 0.0f));
    F.numVisibilitiesPerFreqPerStoke[i][0] = (long)h.Vo.size();
    F.numVisibilitiesPerFreq[i] = (long)h.Vo.size();
    mx = std::max(mx, (int)h.Vo.size());
  }
  F.valid_frequencies = nchan;
  data->max_number_visibilities_in_channel_and_stokes = mx;
}
}  // namespace

__host__ void readMS(const char*, std::vector<MSAntenna>& antennas,
                     std::vector<Field>& fields, MSData* data, bool, bool,
                     float, int gridding) {
  fill_dataset(antennas, fields, data, gridding);
}
__host__ void readMS(const char*, std::string, std::vector<MSAntenna>& antennas,
                     std::vector<Field>& fields, MSData* data, bool, bool,
                     float, int gridding) {
  fill_dataset(antennas, fields, data, gridding);
}
__host__ void MScopy(const char*, const char*) {}
__host__ void writeMS(const char*, const char*, std::vector<Field>, MSData,
                      float, bool, bool, bool) {}
/* Semantics of src/MSFITSIO.cu:1114-1138: the model visibilities come back to the host vectors and are
 * conjugated where the HOST u is positive (the device copy was folded by hermitianSymmetry). */
__host__ void modelToHost(std::vector<Field>& fields, MSData data, int ngpus,
                          int first) {
  for (int f = 0; f < data.nfields; f++)
    for (int i = 0; i < data.total_frequencies; i++) {
      cudaSetDevice((i % ngpus) + first);
      for (int s = 0; s < data.nstokes; s++) {
        long n = fields[f].numVisibilitiesPerFreqPerStoke[i][s];
        if (n <= 0) continue;
        HVis& h = fields[f].visibilities[i][s];
        cudaMemcpy(h.Vm.data(), fields[f].device_visibilities[i][s].Vm,
                   sizeof(cufftComplex) * n, cudaMemcpyDeviceToHost);
        for (long j = 0; j < n; j++)
          if (h.uvw[j].x > 0) h.Vm[j] = cuConjf(h.Vm[j]);
      }
    }
}
__host__ headerValues readOpenedFITSHeader(fitsfile*&, bool) { return g_prob.hdr; }
__host__ headerValues readFITSHeader(const char*) { return g_prob.hdr; }
__host__ fitsfile* openFITS(const char*) { static fitsfile f; return &f; }
__host__ void closeFITS(fitsfile*) {}
__host__ void OCopyFITS(float*, const char*, const char*, const char*, char*,
                        int, int, float, long, long, double, double,
                        std::string, float, bool) {}
__host__ void OCopyFITSCufftComplex(cufftComplex*, const char*, const char*,
                                    const char*, int, float, long, long, int,
                                    bool) {}
extern "C" void fits_report_error(FILE*, int) {}
extern "C" int fits_read_img(fitsfile*, int, long, long, void*, void*, int*, int*) {
  return 0;
}
extern "C" int fits_read_key(fitsfile*, int, const char*, void*, char*, int*) {
  return 0;
}

static void harnessOrder(Optimizer* optimizer, Image* image) {
  optimizer->setImage(image);
  optimizer->setFlag(0);
  optimizer->optimize();
}

static CKernel* make_ckernel(const char* name, int m, int n) {
  std::string s(name ? name : "PillBox2D");
  if (s == "PillBox2D") return new PillBox2D();
  if (s == "Gaussian2D") return new Gaussian2D(m, n);
  if (s == "GaussianSinc2D") return new GaussianSinc2D(m, n);
  if (s == "Sinc2D") return new Sinc2D(m, n);
  if (s == "PSWF" || s == "PSWF_12D") return new PSWF_12D(m, n);
  fprintf(stderr, "gvref: unknown ckernel %s\n", name);
  return nullptr;
}

extern "C" {

/* ------------------------------------------------------- problem definition */
int gvref_problem_begin(long m, long n, double cdelt1, double cdelt2,
                        double ra_deg, double dec_deg, double crp1, double crp2,
                        const char* telescope, float dish, int nchan,
                        const float* freqs) {
  g_prob = SynthProblem();
  g_prob.hdr.M = m;
  g_prob.hdr.N = n;
  g_prob.hdr.DELTAX = cdelt1;
  g_prob.hdr.DELTAY = cdelt2;
  g_prob.hdr.ra = ra_deg;
  g_prob.hdr.dec = dec_deg;
  g_prob.hdr.crpix1 = crp1;
  g_prob.hdr.crpix2 = crp2;
  g_prob.hdr.beam_bmaj = g_prob.hdr.beam_bmin = g_prob.hdr.beam_bpa = 0.0;
  g_prob.hdr.beam_noise = -1.0f;
  g_prob.hdr.radesys = "ICRS";
  g_prob.hdr.equinox = 2000.0f;
  g_prob.hdr.bitpix = -32;
  g_prob.telescope = telescope ? telescope : "ALMA";
  g_prob.dish = dish;
  g_prob.freqs.assign(freqs, freqs + nchan);
  g_prob.chans.assign(nchan, SynthChannel());
  return 0;
}

/* The field's phase/pointing centre (degrees) when it differs from the image centre CRVAL1/2: the
 * reference places it with direccos (src/mfs.cu:660-691). Call after gvref_problem_begin. */
int gvref_set_field_centre(double ra_deg, double dec_deg) {
  g_prob.has_field_centre = true;
  g_prob.field_ra = ra_deg;
  g_prob.field_dec = dec_deg;
  return 0;
}

int gvref_problem_channel(int chan, long Z, const double* uvw_m, const float* Vo,
                          const float* w) {
  if (chan < 0 || chan >= (int)g_prob.chans.size()) return -1;
  SynthChannel& c = g_prob.chans[chan];
  c.uvw.resize(Z);
  c.Vo.resize(Z);
  c.w.assign(w, w + Z);
  for (long k = 0; k < Z; k++) {
    c.uvw[k] = {uvw_m[3 * k], uvw_m[3 * k + 1], uvw_m[3 * k + 2]};
    c.Vo[k] = make_cuFloatComplex(Vo[2 * k], Vo[2 * k + 1]);
    g_prob.max_uv_m = std::max(g_prob.max_uv_m,
                               std::max(fabs(uvw_m[3 * k]), fabs(uvw_m[3 * k + 1])));
  }
  return 0;
}

/* ------------------------------------------------ host-only reference code */
static void cpu_prepare(int gridding) {
  gvref_tolerate_cuda_errors = 1;
  M = g_prob.hdr.M;
  N = g_prob.hdr.N;
  DELTAX = g_prob.hdr.DELTAX;
  DELTAY = g_prob.hdr.DELTAY;
  double dx = RPDEG_D * DELTAX, dy = RPDEG_D * DELTAY; /* src/mfs.cu:493-496 */
  deltau = 1.0 / (M * dx);
  deltav = 1.0 / (N * dy);
  datasets.clear();
  datasets.push_back(MSDataset());
  nMeasurementSets = 1;
  fill_dataset(datasets[0].antennas, datasets[0].fields, &datasets[0].data, gridding);
}

static void copy_out_weights(float** w_out) {
  Field& F = datasets[0].fields[0];
  for (size_t i = 0; i < F.visibilities.size(); i++)
    memcpy(w_out[i], F.visibilities[i][0].weight.data(),
           sizeof(float) * F.visibilities[i][0].weight.size());
}

/* Runs WeightingScheme::apply (src/*weightingscheme.cu) on the stored problem;
 * w_out[chan] receives the Z_chan new weights. */
int gvref_cpu_weights(const char* scheme, float robust, int threads, float** w_out) {
  cpu_prepare(0);
  WeightingScheme* s = createObject<WeightingScheme, std::string>(scheme);
  s->setThreads(threads);
  s->configure(&robust);
  s->apply(datasets);
  copy_out_weights(w_out);
  delete s;
  return 0;
}

/* CKernel table + GCF image (include/classes/ckernel.cuh, src/<kernel>.cu) as
 * MFS::configure builds them (src/mfs.cu:510-516). table: m*n floats, gcf: M*N
 * floats (either may be NULL). */
int gvref_cpu_ckernel(const char* name, int m, int n, float* table, float* gcf,
                      int* support_xy) {
  cpu_prepare(0);
  CKernel* ck = make_ckernel(name, m, n);
  if (!ck) return -1;
  double dx = RPDEG_D * DELTAX, dy = RPDEG_D * DELTAY;
  ck->setSigmas(fabs(deltau), fabs(deltav));
  ck->buildKernel();
  if (table) memcpy(table, ck->getKernelPointer(), sizeof(float) * ck->getm() * ck->getn());
  if (support_xy) { support_xy[0] = ck->getSupportX(); support_xy[1] = ck->getSupportY();
                    support_xy[2] = ck->getm(); support_xy[3] = ck->getn(); }
  if (gcf) {
    ck->initializeGCF(M, N, fabs(dx), fabs(dy));
    memcpy(gcf, ck->getGCFCPUPointer(), sizeof(float) * M * N);
  }
  return 0;
}

/* Optional weighting, then do_gridding (src/functions.cu:1339). Results stay in
 * `datasets`; fetch with gvref_cpu_gridded_count / gvref_cpu_gridded_fetch. */
static double g_cpu_seconds[2] = {0.0, 0.0};   /* WeightingScheme::apply, do_gridding of the last call */
int gvref_cpu_gridding(const char* scheme, float robust, const char* ckname,
                       int m, int n, int threads) {
  cpu_prepare(1);
  g_cpu_seconds[0] = g_cpu_seconds[1] = 0.0;
  if (scheme && scheme[0]) {
    WeightingScheme* s = createObject<WeightingScheme, std::string>(scheme);
    s->setThreads(threads);
    s->configure(&robust);
    double t0 = omp_get_wtime();
    s->apply(datasets);
    g_cpu_seconds[0] = omp_get_wtime() - t0;
    delete s;
  }
  CKernel* ck = make_ckernel(ckname, m, n);
  if (!ck) return -1;
  ck->setSigmas(fabs(deltau), fabs(deltav));
  ck->buildKernel();
  double t0 = omp_get_wtime();
  do_gridding(datasets[0].fields, &datasets[0].data, deltau, deltav, M, N, ck, threads);
  g_cpu_seconds[1] = omp_get_wtime() - t0;
  return 0;
}
/* wall seconds of the two reference host calls inside the last gvref_cpu_gridding */
int gvref_cpu_last_seconds(double* out2) { out2[0] = g_cpu_seconds[0]; out2[1] = g_cpu_seconds[1]; return 0; }
long gvref_cpu_gridded_count(int chan) {
  return datasets[0].fields[0].numVisibilitiesPerFreqPerStoke[chan][0];
}
int gvref_cpu_gridded_fetch(int chan, double* uvw_m, float* Vo, float* w) {
  HVis& h = datasets[0].fields[0].visibilities[chan][0];
  for (size_t k = 0; k < h.Vo.size(); k++) {
    uvw_m[3 * k] = h.uvw[k].x; uvw_m[3 * k + 1] = h.uvw[k].y; uvw_m[3 * k + 2] = h.uvw[k].z;
    Vo[2 * k] = h.Vo[k].x; Vo[2 * k + 1] = h.Vo[k].y;
    w[k] = h.weight[k];
  }
  return 0;
}

/* ------------------------------------------------------ GPU reference path */
/* Wires the object graph as src/main.cu:147-203 does and runs
 * MFS::configure + MFS::setDevice.  `args` is the command line (space
 * separated) the reference binary would get, e.g.
 * "-X 16 -Y 16 -V 256 -z 0.001 -Z 0.01 -t 50 -i synth.ms -o out.ms -m hdr.fits". */
int gvref_init(const char* args, const char* optimizer, const char* scheme,
               const char* ckname, int ck_m, int ck_n, int with_tv) {
  gvref_tolerate_cuda_errors = 0;
  cudaError_t err = cudaGetDeviceCount(&num_gpus);
  if (err != cudaSuccess || num_gpus < 1) return -100;

  std::vector<std::string> toks;
  { std::istringstream ss(args); std::string t; toks.push_back("gpuvmem");
    while (ss >> t) toks.push_back(t); }
  std::vector<char*> argv;
  for (auto& t : toks) argv.push_back(strdup(t.c_str()));
  argv.push_back(nullptr);

  optind = 1; /* getopt state, in case of re-init */
  g_sy = createObject<Synthesizer, std::string>("MFS");
  g_opt = createObject<Optimizer, std::string>(optimizer);
  g_ck = make_ckernel(ckname, ck_m, ck_n);
  g_of = createObject<ObjectiveFunction, std::string>("ObjectiveFunction");
  g_ioms = createObject<Io, std::string>("IoMS");
  g_iofits = createObject<Io, std::string>("IoFITS");
  g_scheme = createObject<WeightingScheme, std::string>(scheme);
  if (!g_sy || !g_opt || !g_ck || !g_of || !g_scheme) return -1;

  g_sy->setIoVisibilitiesHandler(g_ioms);
  g_sy->setIoImageHandler(g_iofits);
  g_sy->setOrder(&harnessOrder);
  g_sy->setWeightingScheme(g_scheme);
  g_sy->setGriddingKernel(g_ck);
  g_sy->setOptimizator(g_opt);
  g_sy->configure((int)argv.size() - 1, argv.data());
  g_opt->setObjectiveFunction(g_of);
  g_gridding = g_sy->getGridding();
  g_sy->setDevice();

  Fi* chi2 = createObject<Fi, std::string>("Chi2");
  Fi* e = createObject<Fi, std::string>("Entropy");
  Fi* l1 = createObject<Fi, std::string>("L1-Norm");
  Fi* tsqv = createObject<Fi, std::string>("TotalSquaredVariation");
  Fi* lap = createObject<Fi, std::string>("Laplacian");
  chi2->configure(-1, 0, 0, false);
  e->configure(0, 0, 0, false);
  e->setPrior(0.001f);
  l1->configure(1, 0, 0, false);
  tsqv->configure(2, 0, 0, false);
  lap->configure(3, 0, 0, false);
  g_fis = {chi2, e, l1, tsqv, lap};
  g_of->addFi(chi2);
  g_of->addFi(e);
  g_of->addFi(l1);
  g_of->addFi(tsqv);
  g_of->addFi(lap);
  if (with_tv) {
    Fi* tv = createObject<Fi, std::string>("TotalVariation");
    tv->configure(4, 0, 0, false);
    g_fis.push_back(tv);
    g_of->addFi(tv);
  }
  g_of->setIo(g_iofits);
  g_of->configure(N, M, image_count);

  /* what MFS::run does before calling the optimizer (src/mfs.cu:978-992) */
  chi2->setFgScale(g_sy->getFgScale());
  if (g_gridding) chi2->setCKernel(g_ck);
  return 0;
}

/* Scalars the reference derived in configure/setDevice; out[0..15]. */
int gvref_scalars(double* out) {
  Field& F = datasets[0].fields[0];
  out[0] = g_sy->getFgScale();
  out[1] = noise_cut;
  out[2] = noise_jypix;
  out[3] = deltau;
  out[4] = deltav;
  out[5] = F.phs_xobs_pix;
  out[6] = F.phs_yobs_pix;
  out[7] = nu_0;
  out[8] = beam_bmaj;
  out[9] = beam_bmin;
  out[10] = beam_bpa;
  out[11] = g_sy->getVisNoise();
  out[12] = datasets[0].antennas[0].pb_cutoff;
  out[13] = datasets[0].antennas[0].pb_factor;
  out[14] = eta;
  out[15] = threshold;
  return 0;
}

int gvref_noise_image(float* out) {
  return (int)cudaMemcpy(out, device_noise_image, sizeof(float) * M * N,
                         cudaMemcpyDeviceToHost);
}

long gvref_nvis(int chan) {
  return datasets[0].fields[0].numVisibilitiesPerFreqPerStoke[chan][0];
}

/* Device-side visibility arrays after hermitianSymmetry / a forward pass. */
int gvref_get_vis(int chan, double* uvw_lambda, float* Vo, float* Vm, float* Vr,
                  float* w) {
  cudaSetDevice(firstgpu + (chan % num_gpus));
  DVis& d = datasets[0].fields[0].device_visibilities[chan][0];
  long n = gvref_nvis(chan);
  if (uvw_lambda) cudaMemcpy(uvw_lambda, d.uvw, sizeof(double3) * n, cudaMemcpyDeviceToHost);
  if (Vo) cudaMemcpy(Vo, d.Vo, sizeof(cufftComplex) * n, cudaMemcpyDeviceToHost);
  if (Vm) cudaMemcpy(Vm, d.Vm, sizeof(cufftComplex) * n, cudaMemcpyDeviceToHost);
  if (Vr) cudaMemcpy(Vr, d.Vr, sizeof(cufftComplex) * n, cudaMemcpyDeviceToHost);
  if (w) cudaMemcpy(w, d.weight, sizeof(float) * n, cudaMemcpyDeviceToHost);
  cudaSetDevice(firstgpu);
  return 0;
}

/* Host weights after weighting(+gridding) in MFS::configure (metres). */
int gvref_get_host_vis(int chan, double* uvw_m, float* Vo, float* w) {
  HVis& h = datasets[0].fields[0].visibilities[chan][0];
  for (size_t k = 0; k < h.Vo.size(); k++) {
    if (uvw_m) { uvw_m[3 * k] = h.uvw[k].x; uvw_m[3 * k + 1] = h.uvw[k].y; uvw_m[3 * k + 2] = h.uvw[k].z; }
    if (Vo) { Vo[2 * k] = h.Vo[k].x; Vo[2 * k + 1] = h.Vo[k].y; }
    if (w) w[k] = h.weight[k];
  }
  return 0;
}

static float* image_dev() { return g_sy->getImage()->getImage(); }

int gvref_set_image(const float* I_host) {
  return (int)cudaMemcpy(image_dev(), I_host, sizeof(float) * M * N * image_count,
                         cudaMemcpyHostToDevice);
}
int gvref_get_image(float* I_host) {
  return (int)cudaMemcpy(I_host, image_dev(), sizeof(float) * M * N * image_count,
                         cudaMemcpyDeviceToHost);
}

/* ObjectiveFunction::calcFunction on the current device image; iteration gates
 * the priors so it is set explicitly. fi_out gets get_fi_values(). */
float gvref_calc_function(int iteration, float* fi_out, int nfi) {
  for (Fi* f : g_of->getFi()) f->setIteration(iteration);
  float v = g_of->calcFunction(image_dev());
  std::vector<float> vals = g_of->get_fi_values();
  for (int i = 0; i < nfi && i < (int)vals.size(); i++) fi_out[i] = vals[i];
  return v;
}

/* ObjectiveFunction::calcGradient; grad_out: image_count*M*N floats (xi). */
int gvref_calc_gradient(int iteration, int flag, float* grad_out) {
  flag_opt = flag;
  float* xi = nullptr;
  cudaMalloc(&xi, sizeof(float) * M * N * image_count);
  g_of->calcGradient(image_dev(), xi, iteration);
  cudaMemcpy(grad_out, xi, sizeof(float) * M * N * image_count, cudaMemcpyDeviceToHost);
  cudaFree(xi);
  return 0;
}

/* n back-to-back (calcFunction + calcGradient) pairs timed with CUDA events on
 * the default stream (the reference synchronises internally). Returns ms/eval. */
float gvref_time_evals(int n, int iteration, int flag) {
  flag_opt = flag;
  float* xi = nullptr;
  cudaMalloc(&xi, sizeof(float) * M * N * image_count);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (Fi* f : g_of->getFi()) f->setIteration(iteration);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < n; i++) {
    g_of->calcFunction(image_dev());
    g_of->calcGradient(image_dev(), xi, iteration);
  }
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0.0f;
  cudaEventElapsedTime(&ms, a, b);
  cudaFree(xi);
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return ms / n;
}

/* Whole reconstruction: MFS::run() (optimizer through harnessOrder). */
int gvref_run(float* image_out, float* wall_ms) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  g_sy->run();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  if (wall_ms) cudaEventElapsedTime(wall_ms, a, b);
  if (image_out) gvref_get_image(image_out);
  return g_opt->getCurrentIteration();
}

/* calculateErrors (src/functions.cu:4966-5040) on the current device image and the residuals
 * of the last chi2(); out: image_count*M*N floats (error_Inu_0, error_alpha). */
int gvref_error_image(float* out) {
  Image* img = g_sy->getImage();
  calculateErrors(img);
  cudaMemcpy(out, img->getErrorImage(), sizeof(float) * M * N * image_count, cudaMemcpyDeviceToHost);
  cudaFree(img->getErrorImage());
  img->setErrorImage(nullptr);
  return 0;
}

/* The reference's degriddingGPU kernel (src/functions.cu:2205-2254; never launched by the
 * reference itself) on caller-supplied arrays: uvw in wavelengths [Z][3], a CENTRED model grid
 * Vg [M][N] complex, a kernel table [km][kn]. Vm_out [Z][2]. */
int gvref_degridding(long Z, const double* uvw_lambda, const float* Vg, const float* table,
                     double du, double dv, int m_img, int n_img, int km, int kn, int sx, int sy,
                     float* Vm_out) {
  double3* d_uvw; cufftComplex *d_vm, *d_vg; float* d_k;
  cudaMalloc(&d_uvw, sizeof(double3) * Z);
  cudaMalloc(&d_vm, sizeof(cufftComplex) * Z);
  cudaMalloc(&d_vg, sizeof(cufftComplex) * m_img * n_img);
  cudaMalloc(&d_k, sizeof(float) * km * kn);
  cudaMemcpy(d_uvw, uvw_lambda, sizeof(double3) * Z, cudaMemcpyHostToDevice);
  cudaMemcpy(d_vg, Vg, sizeof(cufftComplex) * m_img * n_img, cudaMemcpyHostToDevice);
  cudaMemcpy(d_k, table, sizeof(float) * km * kn, cudaMemcpyHostToDevice);
  degriddingGPU<<<(int)((Z + 255) / 256), 256>>>(d_uvw, d_vm, d_vg, d_k, du, dv, (int)Z, m_img, n_img, km, kn, sx, sy);
  cudaError_t err = cudaDeviceSynchronize();
  cudaMemcpy(Vm_out, d_vm, sizeof(cufftComplex) * Z, cudaMemcpyDeviceToHost);
  cudaFree(d_uvw); cudaFree(d_vm); cudaFree(d_vg); cudaFree(d_k);
  return (int)err;
}

/* One Fi of the reference (factory name: "Entropy", "L1-Norm", "TotalVariation", "TotalSquaredVariation",
 * "Laplacian", "Quadratic", "GEntropy", "GL1Norm") evaluated ON ITS OWN on a caller-supplied image
 * [image_count][M][N]: Fi::configure(-1, image_index, image_index, false) (src/main.cu:185-197 pattern),
 * penalization factor = lambda, then calcFi -> get_fivalue() and restartDGi + calcGi + addToDphi into a
 * zeroed dphi [image_count][M][N] (what ObjectiveFunction::calcGradient does per term,
 * include/classes/objectivefunction.cuh). prior_host: M*N floats for GEntropy / GL1Norm, else NULL.
 * Needs gvref_init first (noise image, M, N, image_count).
 * flag: the optimizers' flag_opt (the gradient hosts only act when flag_opt % 2 == imageIndex).
 * prior_after_out (optional, M*N): the term's prior image AFTER calcGi — GL1Norm::calcGi passes its two image
 * arguments to DGL1Norm in swapped order (src/gl1norm.cu:145-148 vs src/functions.cu:4700), so the reference
 * writes that gradient into the prior image and adds zeros to dphi. */
int gvref_prior_eval(const char* name, const float* I_host, const float* prior_host, float lambda,
                     float prior_value, float eta_v, float eps_a, float eps_b, int image_index, int iteration,
                     int flag, float* value_out, float* dphi_out, float* prior_after_out) {
  Fi* f = createObject<Fi, std::string>(name);
  if (!f) return -1;
  flag_opt = flag;
  f->configure(-1, image_index, image_index, false);
  f->setPenalizationFactor(lambda);
  f->setIteration(iteration);
  std::string s(name);
  float* d_prior = nullptr;
  if (prior_host) {
    cudaMalloc(&d_prior, sizeof(float) * M * N);
    cudaMemcpy(d_prior, prior_host, sizeof(float) * M * N, cudaMemcpyHostToDevice);
    f->setPrior(d_prior);   /* owned (and freed) by the Fi from here on */
  }
  if (s == "Entropy") { f->setPrior(prior_value); f->setEta(eta_v); }
  if (s == "GEntropy") f->setEta(eta_v);
  if (s == "TotalVariation") static_cast<TVariation*>(f)->setEpsilon(eps_a);
  if (s == "GL1Norm") static_cast<GL1Norm*>(f)->setEpsilons(eps_a, eps_b);
  float *d_I = nullptr, *d_phi = nullptr;
  const size_t bytes = sizeof(float) * M * N * image_count;
  cudaMalloc(&d_I, bytes);
  cudaMalloc(&d_phi, bytes);
  cudaMemcpy(d_I, I_host, bytes, cudaMemcpyHostToDevice);
  cudaMemset(d_phi, 0, bytes);
  f->calcFi(d_I);
  if (value_out) *value_out = f->get_fivalue();
  f->restartDGi();
  f->calcGi(d_I, d_phi);
  f->addToDphi(d_phi);
  cudaError_t err = cudaDeviceSynchronize();
  if (dphi_out) cudaMemcpy(dphi_out, d_phi, bytes, cudaMemcpyDeviceToHost);
  if (prior_after_out && d_prior) cudaMemcpy(prior_after_out, d_prior, sizeof(float) * M * N, cudaMemcpyDeviceToHost);
  cudaFree(d_I);
  cudaFree(d_phi);
  flag_opt = 0;
  return (int)err;
}

/* MFS::writeResiduals (src/mfs.cu:1115-1155): weights restored (or, in gridded mode, the original samples
 * brought back by getOriginalVisibilitiesBack, src/functions.cu:1844-2010, and the model re-sampled at the
 * original (u,v) by one more chi2()), then modelToHost. Returns through *nongridded_chi2 what that last
 * Chi2::calcFi left in get_fivalue() (the "Non-gridded chi2" the reference prints). Fetch the result with
 * gvref_nvis + gvref_get_host_model. */
int gvref_write_residuals(float* nongridded_chi2) {
  g_sy->writeResiduals();
  if (nongridded_chi2) *nongridded_chi2 = g_fis[0]->get_fivalue();
  return (int)cudaDeviceSynchronize();
}
/* Host arrays after modelToHost: Vm [Z][2], weight [Z], uvw in metres [Z][3], Vo [Z][2]. */
int gvref_get_host_model(int chan, double* uvw_m, float* Vo, float* Vm, float* w) {
  HVis& h = datasets[0].fields[0].visibilities[chan][0];
  for (size_t k = 0; k < h.Vm.size(); k++) {
    if (uvw_m) { uvw_m[3 * k] = h.uvw[k].x; uvw_m[3 * k + 1] = h.uvw[k].y; uvw_m[3 * k + 2] = h.uvw[k].z; }
    if (Vo) { Vo[2 * k] = h.Vo[k].x; Vo[2 * k + 1] = h.Vo[k].y; }
    if (Vm) { Vm[2 * k] = h.Vm[k].x; Vm[2 * k + 1] = h.Vm[k].y; }
    if (w) w[k] = h.weight[k];
  }
  return 0;
}

int gvref_set_lbfgs_k(int k) { g_opt->setK(k); return 0; }
void gvref_set_verbose(int v) { verbose_flag = v != 0; }

}  // extern "C"
