// mfs.cpp — MFS synthesizer: the host flow of src/mfs.cu (configure :79-528, setDevice
// :530-943, clearRun :945-976, run :978-1074, writeImages :1076-1113, writeResiduals
// :1115-1155, unSetDevice :1157-1248) driving the B200 engine through the C ABI.
// One process per GPU; multi-GPU jobs run one MFS per rank over a NCCL communicator.
#include <getopt.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>

#include "synthesizer.hpp"

namespace gpuvmem {

namespace {
const double RPDEG_D = 3.14159265358979323846 / 180.0;  // include/functions.cuh:18
const double RPARCSEC = RPDEG_D / 3600.0;
const double PI_D = 3.14159265358979323846;

double wallSeconds() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
bool usedCorrelation(int c) { return c == LL || c == RR || c == XX || c == YY; }  // src/functions.cu:4378-4381

// src/directioncosines.cu:40-59
void direccos(double ra, double dec, double ra0, double dec0, double* l, double* m) {
  const double dra = ra - ra0;
  *l = std::cos(dec) * std::sin(dra);
  *m = std::sin(dec) * std::cos(dec0) - std::cos(dec) * std::sin(dec0) * std::cos(dra);
}

struct FlagSpec { char key; const char* name; bool takes_value; const char* help; };
const FlagSpec kFlags[] = {
    {'i', "input", true, "Name of the input visibility file/s (separated by a comma)"},
    {'o', "output", true, "Name of the output visibility file/s (separated by a comma)"},
    {'O', "output_image", true, "Name of the output image"},
    {'m', "model_input", true, "File including a complete header for astrometry"},
    {'n', "noise", true, "Noise factor parameter"},
    {'e', "eta", true, "Variable that controls the minimum image value in the entropy prior"},
    {'N', "noise_cut", true, "Noise-cut Parameter"},
    {'F', "ref_frequency", true, "Reference frequency in Hz (if alpha is not zero)"},
    {'T', "threshold", true, "Threshold to calculate the spectral index image above a number of sigmas in I_nu_0"},
    {'p', "path", true, "Path to save images. With last trail / included"},
    {'G', "gpus", true, "Index of the GPU/s you are going to use separated by a comma"},
    {'r', "random_sampling", true, "Percentage of data used when random sampling"},
    {'R', "robust_parameter", true, "Robust weighting parameter when gridding (-2 uniform, 2 natural)"},
    {'f', "output_file", true, "Output file where final objective function values are saved"},
    {'X', "blockSizeX", true, "GPU block X Size for image/Fourier plane (accepted, unused by this engine)"},
    {'Y', "blockSizeY", true, "GPU block Y Size for image/Fourier plane (accepted, unused by this engine)"},
    {'V', "blockSizeV", true, "GPU block V Size for visibilities (accepted, unused by this engine)"},
    {'t', "iterations", true, "Number of iterations for optimization"},
    {'g', "gridding", true, "Use gridded visibilities (any value > 0; gridding runs on the GPU)"},
    {'z', "initial_values", true, "Initial values for image/s"},
    {'Z', "regularization_factors", true, "Regularization factors for each regularization (separated by a comma)"},
    {'U', "user-mask", true, "Use a user created mask instead of using the noise mask"},
    {'K', "grad-mode", true, "Gradient kernel: 0 auto, 1 tensor-core, 2 CUDA-core separable, 3 CUDA-core exact (engine extension)"},
    {'v', "verbose", false, "Shows information through all the execution"},
    {'x', "nopositivity", false, "Runs with no positivity restrictions on the images"},
    {'a', "apply-noise", false, "Applies random gaussian noise to visibilities (not supported)"},
    {'P', "print-images", false, "Prints images per iteration"},
    {'E', "print-errors", false, "Prints final error maps (not supported)"},
    {'s', "save_modelcolumn", false, "Saves the model visibilities"},
    {'M', "use-radius-mask", false, "Use a mask based on a radius instead of the noise estimation (not supported)"},
    {'W', "modify-weights", false, "Modify the WEIGHT column with the computed weights"},
    {'h', "help", false, "Shows this help"},
    {'w', "warranty", false, "Shows warranty details"},
    {'c', "copyright", false, "Shows copyright conditions"},
};
}  // namespace

void print_help() {
  std::printf("gpuvmem_b200 options:\n");
  for (const FlagSpec& f : kFlags)
    std::printf("  -%c, --%-24s %s\n", f.key, f.name, f.help);
}

bool getOptions(int argc, char** argv, Vars* v) {
  Globals& g = G();
  std::string shortopts;
  std::vector<option> longopts;
  for (const FlagSpec& f : kFlags) {
    shortopts += f.key;
    if (f.takes_value) shortopts += ':';
    longopts.push_back({f.name, f.takes_value ? required_argument : no_argument, nullptr, f.key});
  }
  longopts.push_back({nullptr, 0, nullptr, 0});
  bool help = false;
  optind = 0;   // glibc: 0 = full re-initialisation (1 would keep `nextchar` pointing into the previous argv)
  opterr = 0;
  int c;
  while ((c = getopt_long(argc, argv, shortopts.c_str(), longopts.data(), nullptr)) != -1) {
    switch (c) {
      case 'i': v->input = optarg; break;
      case 'o': v->output = optarg; break;
      case 'O': v->output_image = optarg; break;
      case 'm': v->modin = optarg; break;
      case 'n': v->noise = std::stof(optarg); break;
      case 'e': v->eta = std::stof(optarg); break;
      case 'N': v->noise_cut = std::stof(optarg); break;
      case 'F': v->nu_0 = std::stof(optarg); break;
      case 'T': v->threshold = std::stof(optarg); break;
      case 'p': v->path = optarg; break;
      case 'G': v->gpus = optarg; break;
      case 'r': v->randoms = std::stof(optarg); break;
      case 'R': v->robust_param = std::stof(optarg); break;
      case 'f': v->ofile = optarg; break;
      case 'X': v->blockSizeX = std::stoi(optarg); break;
      case 'Y': v->blockSizeY = std::stoi(optarg); break;
      case 'V': v->blockSizeV = std::stoi(optarg); break;
      case 't': v->it_max = std::stoi(optarg); break;
      case 'g': v->gridding = std::stoi(optarg); break;
      case 'z': v->initial_values = optarg; break;
      case 'Z': v->penalization_factors = optarg; break;
      case 'U': v->user_mask = optarg; break;
      case 'K': v->grad_mode = std::stoi(optarg); break;
      case 'v': g.verbose_flag = 1; break;
      case 'x': g.nopositivity = true; break;
      case 'a': g.apply_noise = true; break;
      case 'P': g.print_images = true; break;
      case 'E': g.print_errors = true; break;
      case 's': g.save_model_input = true; break;
      case 'M': g.radius_mask = true; break;
      case 'W': g.modify_weights = true; break;
      case 'h': case 'w': case 'c': help = true; break;
      default: help = true; break;  // unknown flag: the reference prints the help and exits
    }
  }
  if (help) { print_help(); return false; }
  if (v->randoms > 1.0 || v->randoms < 0.0 || v->gridding < 0) { print_help(); return false; }
  if (v->user_mask != "NULL") v->noise_cut = 1.0f;  // src/functions.cu:296-298
  return true;
}

std::vector<std::string> MFS::countAndSeparateStrings(std::string long_str, std::string sep) {
  std::vector<std::string> out;
  size_t start = 0;
  while (start <= long_str.size()) {
    const size_t hit = long_str.find_first_of(sep, start);
    const size_t end = hit == std::string::npos ? long_str.size() : hit;
    out.push_back(long_str.substr(start, end - start));
    if (hit == std::string::npos) break;
    start = hit + 1;
  }
  return out;
}

MFS::~MFS() {}

void MFS::adoptDatasets(std::vector<MSDataset>&& ds, const headerValues& h) {
  datasets = std::move(ds);
  header = h;
  adopted = true;
}

void MFS::setDistributed(int rank, int world, const std::string& id) {
  G().rank = rank;
  G().world = world;
  G().quiet = rank != 0;
  nccl_id = id;
}

void MFS::configure(int argc, char** argv) {
  Globals& g = G();
  t_start = wallSeconds();
  if (!ioImageHandler) ioImageHandler = createObject<Io, std::string>("IoFITS");
  if (!ioVisibilitiesHandler) ioVisibilitiesHandler = createObject<Io, std::string>("IoMS");
  if (!getOptions(argc, argv, &variables)) std::exit(EXIT_SUCCESS);

  msinput = variables.input;
  msoutput = variables.output;
  modinput = variables.modin;
  out_image = variables.output_image;
  ioImageHandler->setInput(modinput);
  ioImageHandler->setOutput(out_image);
  ioImageHandler->setPath(variables.path);
  optimizer->setTotalIterations(variables.it_max);
  setVisNoise(variables.noise);
  g.noise_cut = variables.noise_cut;
  g.random_probability = variables.randoms;
  g.eta = variables.eta;
  setGriddingThreads(variables.gridding);
  g.nu_0 = variables.nu_0;
  g.robust_param = variables.robust_param;
  g.threshold = variables.threshold * 5.0;
  ioImageHandler->setPrintImages(g.print_images);
  if (g.apply_noise || variables.randoms < 1.0f) {
    std::printf("ERROR: -a and -r < 1 are outside this engine's scope (DESIGN.md §7)\n");
    std::exit(-1);
  }

  if (!adopted) {
    if (msinput == "NULL") {
      std::printf("Datasets files were not provided\n");
      print_help();
      std::exit(-1);
    }
    if (msoutput == "NULL") {
      std::printf("Output/s was/were not provided\n");
      print_help();
      std::exit(-1);
    }
    const std::vector<std::string> ins = countAndSeparateStrings(msinput, ",");
    const std::vector<std::string> outs = countAndSeparateStrings(msoutput, ",");
    if (ins.size() != outs.size()) {
      std::printf("Number of input datasets should be equal to the number of output datasets\n");
      std::exit(-1);
    }
    datasets.assign(ins.size(), MSDataset());
    for (size_t i = 0; i < ins.size(); i++) {
      datasets[i].name = ins[i];
      datasets[i].oname = outs[i];
    }
  }
  g.nMeasurementSets = (int)datasets.size();
  if (!visibilities) visibilities = new Visibilities();      // src/mfs.cu:489-491
  visibilities->setMSDataset(datasets);
  visibilities->setNDatasets(g.nMeasurementSets);
  if (g.verbose_flag && !g.quiet) std::printf("Number of input datasets %d\n", g.nMeasurementSets);

  if (variables.initial_values == "NULL") {
    std::printf("Initial values for image/s were not provided\n");
    print_help();
    std::exit(-1);
  }
  const std::vector<std::string> init = countAndSeparateStrings(variables.initial_values, ",");
  g.image_count = (int)init.size();
  g.initial_values.clear();
  for (int i = 0; i < g.image_count; i++)
    g.initial_values.push_back(i == 0 ? std::stof(init[i]) * -1.0f * g.eta : std::stof(init[i]));
  g.imagesChanged = 0;
  if (g.image_count == 1) {  // src/mfs.cu:176-180: a second (alpha) image is always carried
    g.initial_values.push_back(0.0f);
    g.image_count++;
    g.imagesChanged = 1;
  }

  if (!adopted) header = ioImageHandler->readHeader(modinput);
  g.M = header.M;
  g.N = header.N;
  g.DELTAX = header.DELTAX;
  g.DELTAY = header.DELTAY;
  g.ra = header.ra;
  g.dec = header.dec;
  g.crpix1 = header.crpix1;
  g.crpix2 = header.crpix2;
  ioImageHandler->setMN(g.M, g.N);
  ioImageHandler->setRADec(g.ra, g.dec);
  ioImageHandler->setPixelGrid(g.DELTAX, g.DELTAY, g.crpix1, g.crpix2);
  ioImageHandler->setFrame(header.radesys);
  ioImageHandler->setEquinox(header.equinox);
  if (header.beam_noise > 0.0f) setVisNoise(header.beam_noise);
  if (ckernel) ckernel->setIoImageHandler(ioImageHandler);

  if (!adopted)
    for (MSDataset& ds : datasets)
      ioVisibilitiesHandler->read(ds.name, ds.antennas, ds.fields, &ds.data);

  float max_freq = 0, min_freq = 0, max_blength = 0;
  double max_uvmax = 0;
  for (size_t d = 0; d < datasets.size(); d++) {
    const MSData& md = datasets[d].data;
    if (d == 0) { max_freq = md.max_freq; min_freq = md.min_freq; }
    max_freq = std::max(max_freq, md.max_freq);
    min_freq = std::min(min_freq, md.min_freq);
    max_blength = std::max(max_blength, md.max_blength);
    max_uvmax = std::max(max_uvmax, md.uvmax_wavelength);
    if (!g.quiet)
      std::printf("Dataset %zu: %s - Antenna diameter: %.3f metres\n", d, datasets[d].name.c_str(),
                  datasets[d].antennas[0].antenna_diameter);
  }
  max_uvmax += 1E-5;
  if (!g.quiet) {
    const float resolution_arcsec = (freq_to_wavelength(max_freq) / max_blength) / RPARCSEC;
    std::printf("The maximum u,v in wavelength units is: %e\n", max_uvmax);
    std::printf("The maximum theoretical resolution of this/these dataset/s is ~%f arcsec\n", resolution_arcsec);
  }
  if (g.nu_0 < 0.0) {
    if (!g.quiet)
      std::printf("WARNING: Reference frequency not provided. It will be calculated as the middle of the frequency range.\n");
    g.nu_0 = 0.5f * (max_freq + min_freq);
  }
  if (!g.quiet) {
    std::printf("Reference frequency: %e Hz\n", g.nu_0);
    const double deltau_theo = 2.0 * max_uvmax / (g.M - 1);
    std::printf("The pixel size has to be less or equal to %lf arcsec\n", 1.0 / (g.M * deltau_theo) / RPARCSEC);
    std::printf("Actual pixel size is %lf arcsec\n", std::fabs(g.DELTAX) * 3600.0);
  }

  // -G: with one process per GPU the list names the devices of the job; this rank drives entry
  // `rank` of it (LOCAL_RANK order). The reference's single-process multi-GPU mode does not exist.
  const std::vector<std::string> gpu_list = countAndSeparateStrings(variables.gpus, ",");
  g.firstgpu = 0;
  if (!gpu_list.empty() && !gpu_list[0].empty()) {
    const size_t pick = (size_t)g.rank < gpu_list.size() ? (size_t)g.rank : 0;
    g.firstgpu = std::stoi(gpu_list[pick]);
    if (g.world == 1 && gpu_list.size() > 1 && !g.quiet)
      std::printf("NOTE: %zu GPUs listed but this is a single process; launch one process per GPU "
                  "(RANK/WORLD_SIZE/LOCAL_RANK) for multi-GPU. Using GPU %d.\n", gpu_list.size(), g.firstgpu);
  }
  g.num_gpus = g.world;
  g.multigpu = g.world > 1 ? g.world : 0;
  if (ckernel) ckernel->setGPUID(g.firstgpu);

  g.penalizators.clear();
  if (variables.penalization_factors != "NULL") {
    for (const std::string& s : countAndSeparateStrings(variables.penalization_factors, ","))
      g.penalizators.push_back(std::stof(s));
  } else if (!g.quiet) {
    std::printf("No regularization factors provided\n");
  }
  g.nPenalizators = (int)g.penalizators.size();

  const double deltax = RPDEG_D * g.DELTAX, deltay = RPDEG_D * g.DELTAY;  // radians
  g.deltau = 1.0 / (g.M * deltax);
  g.deltav = 1.0 / (g.N * deltay);
  der.deltau = g.deltau;
  der.deltav = g.deltav;

  // The engine (per-GPU scratch, cuFFT plan, NCCL communicator) is created here rather than in setDevice: with
  // several ranks the weighting and the gridding below are already distributed and need the communicator.
  createEngine();

  if (!scheme) scheme = createObject<WeightingScheme, std::string>("Natural");
  if (gridding) scheme->setThreads(griddingThreads);
  scheme->configure(&g.robust_param);
  scheme->setModifyWeights(g.modify_weights);
  double t0 = wallSeconds();
  if (gridding) {
    // weighting and gridding share one device work arena (this thread): size it once for the largest block, so that
    // the gridding does not free and re-allocate what the weighting allocated (counted as weighting time)
    int64_t zmax = 0;
    for (auto& ds : datasets)
      for (auto& f : ds.fields)
        for (auto& chan : f.visibilities)
          for (auto& v : chan) zmax = std::max<int64_t>(zmax, (int64_t)v.size());
    GVM_CHECK(gvm_grid_reserve(g.firstgpu, g.M, g.N, zmax, g.world > 1 ? g.world : 1));
  }
  scheme->apply(datasets);
  der.weighting_seconds = wallSeconds() - t0;

  if (gridding) {
    if (!ckernel) ckernel = new PillBox2D();
    if (!g.quiet) std::cout << "Doing gridding" << std::endl;
    ckernel->setSigmas(std::fabs(g.deltau), std::fabs(g.deltav));
    ckernel->buildKernel();
    ckernel->initializeGCF(g.M, g.N, std::fabs(deltax), std::fabs(deltay));
    if (!g.quiet)
      std::printf("Using an antialiasing kernel %s of size (%d, %d) and support (%d, %d)\n",
                  ckernel->getName().c_str(), ckernel->getm(), ckernel->getn(), ckernel->getSupportX(),
                  ckernel->getSupportY());
    t0 = wallSeconds();
    doGridding();
    der.gridding_seconds = wallSeconds() - t0;
  }
}

// do_gridding (src/functions.cu:1339-1653) for every dataset/field/channel/stokes through
// gvm_grid_block; the originals are kept for the residual write-back.
void MFS::doGridding() {
  Globals& g = G();
  // The originals move aside (no copy: they can be gigabytes); the gridded samples go into shells
  // that carry the same metadata (do_gridding replaces the block, src/functions.cu:1577-1612).
  ungridded = std::move(datasets);
  datasets.clear();
  for (MSDataset& src : ungridded) {
    datasets.emplace_back();
    MSDataset& ds = datasets.back();
    ds.name = src.name; ds.oname = src.oname; ds.antennas = src.antennas; ds.data = src.data;
    int max = 0;
    for (Field& sf : src.fields) {
      ds.fields.emplace_back();
      Field& f = ds.fields.back();
      f.id = sf.id; f.valid_frequencies = sf.valid_frequencies;
      f.ref_ra = sf.ref_ra; f.ref_dec = sf.ref_dec; f.phs_ra = sf.phs_ra; f.phs_dec = sf.phs_dec;
      f.ref_xobs_pix = sf.ref_xobs_pix; f.ref_yobs_pix = sf.ref_yobs_pix;
      f.phs_xobs_pix = sf.phs_xobs_pix; f.phs_yobs_pix = sf.phs_yobs_pix;
      f.nu = sf.nu;
      f.numVisibilitiesPerFreqPerStoke = sf.numVisibilitiesPerFreqPerStoke;
      f.numVisibilitiesPerFreq = sf.numVisibilitiesPerFreq;
      f.visibilities.resize(sf.visibilities.size());
      for (size_t i = 0; i < sf.visibilities.size(); i++) {
        long per_freq = 0;
        f.visibilities[i].resize(sf.visibilities[i].size());
        for (size_t s = 0; s < sf.visibilities[i].size(); s++) {
          const HVis& in = sf.visibilities[i][s];
          HVis& v = f.visibilities[i][s];
          int64_t nout = 0;
          if (g.engine && g.world > 1)   // every rank grids a slice of the block; all end with the whole result
            GVM_CHECK(gvm_grid_block_dist(g.engine, f.nu[i], (int64_t)in.size(), in.uvw.data(), in.Vo.data(),
                                          in.weight.data(), ckernel->getKernelPointer(), ckernel->getm(), ckernel->getn(),
                                          ckernel->getSupportX(), ckernel->getSupportY(), &nout));
          else
            GVM_CHECK(gvm_grid_block(g.firstgpu, g.M, g.N, g.deltau, g.deltav, f.nu[i], (int64_t)in.size(),
                                     in.uvw.data(), in.Vo.data(), in.weight.data(), ckernel->getKernelPointer(),
                                     ckernel->getm(), ckernel->getn(), ckernel->getSupportX(),
                                     ckernel->getSupportY(), nullptr, nullptr, nullptr, &nout));
          v.uvw.resize(3 * nout);
          v.Vo.resize(2 * nout);
          v.weight.resize(nout);
          GVM_CHECK(gvm_grid_fetch(v.uvw.data(), v.Vo.data(), v.weight.data()));
          v.Vm.assign(2 * nout, 0.0f);
          v.Vr.assign(2 * nout, 0.0f);
          f.numVisibilitiesPerFreqPerStoke[i][s] = (long)nout;
          per_freq += (long)nout;
          max = std::max<long>(max, (long)nout);
        }
        f.numVisibilitiesPerFreq[i] = per_freq;
      }
    }
    ds.data.max_number_visibilities_in_channel_and_stokes = max;
  }
  datasets_are_gridded = true;
  // the gridding work buffers stay allocated (a later block or run reuses them; cudaFree of tens of GB costs
  // more than the gridding itself) and go back in MFS::unSetDevice
}

bool shardRange(int max_nfreq, int chan, size_t Z, int rank, int world, size_t* lo, size_t* hi) {
  *lo = 0;
  *hi = Z;
  if (world <= 1 || max_nfreq < 0) return true;   // max_nfreq < 0: replicated (every rank holds everything)
  if (max_nfreq >= world) return chan % world == rank;  // whole channels, the reference's rule
  *lo = Z * (size_t)rank / world;                         // contiguous visibility chunks
  *hi = Z * (size_t)(rank + 1) / world;
  return true;
}

// Which part of every (dataset, field, channel, stokes) block lives on this rank, then the
// uploads. Whole channels go to rank (i % world) as in the reference (src/mfs.cu:568,
// src/functions.cu:4341) when there are at least `world` channels; otherwise every block is cut
// into `world` contiguous visibility chunks (chi2 and its gradient are sums over visibilities).
void MFS::shardAndUpload() {
  Globals& g = G();
  int max_nfreq = 1;
  for (MSDataset& ds : datasets) max_nfreq = std::max(max_nfreq, ds.data.total_frequencies);
  g.dist_kind = g.world <= 1 ? GVM_DIST_NONE : (max_nfreq >= g.world ? GVM_DIST_BLOCKS : GVM_DIST_CHUNKS);
  // Gridded samples (-g) are few (at most one per uv cell) and their objective is image-sized work — FFTs and image
  // kernels that do not shard: every rank keeps all of them and evaluates the objective itself, without the
  // all-reduce of the [2][M][N] gradient (which alone costs more than the evaluation). The weighting and the
  // gridding before it ARE distributed (gvm_weights_dist, gvm_grid_block_dist).
  const bool replicated = g.world > 1 && gridding && datasets_are_gridded;
  if (g.world > 1) GVM_CHECK(gvm_dist_set_replicated(g.engine, replicated ? 1 : 0));
  if (replicated) { g.dist_kind = GVM_DIST_NONE; max_nfreq = -1; }
  for (MSDataset& ds : datasets)
    for (Field& f : ds.fields) {
      f.engine_slot.assign(f.visibilities.size(), std::vector<int>(ds.data.nstokes, -1));
      for (size_t i = 0; i < f.visibilities.size(); i++)
        for (int s = 0; s < ds.data.nstokes; s++) {
          if (!usedCorrelation(ds.data.corr_type[s])) continue;
          HVis& v = f.visibilities[i][s];
          size_t lo = 0, hi = 0;
          if (!shardRange(max_nfreq, (int)i, v.size(), g.rank, g.world, &lo, &hi)) continue;
          gvm_channel_desc cd;
          cd.freq = f.nu[i];
          cd.antenna_diameter = ds.antennas[0].antenna_diameter;
          cd.pb_factor = ds.antennas[0].pb_factor;
          cd.pb_cutoff = ds.antennas[0].pb_cutoff;
          cd.primary_beam = ds.antennas[0].primary_beam;
          cd.ref_xobs_pix = f.ref_xobs_pix; cd.ref_yobs_pix = f.ref_yobs_pix;
          cd.phs_xobs_pix = f.phs_xobs_pix; cd.phs_yobs_pix = f.phs_yobs_pix;
          int slot = -1;
          GVM_CHECK(gvm_add_channel(g.engine, &cd, (int64_t)(hi - lo), v.uvw.data() + 3 * lo, v.Vo.data() + 2 * lo,
                                    v.weight.data() + lo, &slot));
          // -normalize divides by the size of the whole block, not of this rank's slice (src/functions.cu:4439-4441)
          if (hi - lo != v.size()) GVM_CHECK(gvm_set_block_nvis(g.engine, slot, (int64_t)v.size()));
          f.engine_slot[i][s] = slot;
        }
    }
}

// gvm_create + gvm_dist_init, once (device-side part of MFS::setDevice, src/mfs.cu:530-560)
void MFS::createEngine() {
  Globals& g = G();
  if (g.engine) return;
  gvm_config cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.M = g.M; cfg.N = g.N; cfg.DELTAX = g.DELTAX; cfg.DELTAY = g.DELTAY;
  cfg.nu_0 = g.nu_0; cfg.eta = g.eta; cfg.minpix = g.initial_values[0];
  cfg.noise_cut = 1e30f; cfg.threshold = g.threshold; cfg.fg_scale = 1.0f;
  cfg.device = g.firstgpu; cfg.grad_mode = variables.grad_mode;
  cfg.keep_vm = 1;  // -o is mandatory in the reference: the model visibilities are always written back
  if (gvm_create(&cfg, &g.engine) != 0) {
    std::printf("ERROR: %s\n", gvm_last_error());
    std::exit(-1);
  }
  if (g.world > 1) GVM_CHECK(gvm_dist_init(g.engine, g.rank, g.world, nccl_id.data(), nccl_id.size()));
}

void MFS::setDevice() {
  Globals& g = G();
  const double deltax = RPDEG_D * g.DELTAX, deltay = RPDEG_D * g.DELTAY;
  g.deltau = 1.0 / (g.M * deltax);
  g.deltav = 1.0 / (g.N * deltay);

  // calculateNoiseAndBeam (src/functions.cu:1700-1813): weighted second moments of the uv
  // coverage in fp64, sum of weights as a running fp32 sum per block (reduceCPU, :354-367)
  double s_uu = 0.0, s_vv = 0.0, s_uv = 0.0;
  sum_weights = 0.0f;
  total_visibilities = 0;
  for (MSDataset& ds : datasets)
    for (Field& f : ds.fields)
      for (size_t i = 0; i < f.visibilities.size(); i++)
        for (int s = 0; s < ds.data.nstokes; s++) {
          if (!usedCorrelation(ds.data.corr_type[s])) continue;
          const HVis& v = f.visibilities[i][s];
          const size_t Z = v.size();
          if (Z == 0) continue;
          double uu = 0.0, vv = 0.0, uv = 0.0;
          for (size_t k = 0; k < Z; k++) {
            const double ul = metres_to_lambda(v.uvw[3 * k], f.nu[i]);
            const double vl = metres_to_lambda(v.uvw[3 * k + 1], f.nu[i]);
            uu += ul * ul * v.weight[k];
            vv += vl * vl * v.weight[k];
            uv += ul * vl * v.weight[k];
          }
          s_uu += uu; s_vv += vv; s_uv += uv;
          float block = v.weight[0];
          for (size_t k = 1; k < Z; k++) block = block + v.weight[k];
          sum_weights += block;
          total_visibilities += (int)Z;
        }
  if (!(sum_weights > 0.0f)) {
    std::printf("Error: The sum of the visibility weights cannot be zero\n");
    std::exit(-1);
  }
  s_uu /= sum_weights; s_vv /= sum_weights; s_uv /= sum_weights;
  const float variance = 1.0f / sum_weights;
  {  // calc_beamSize (:1685-1698)
    const double diff = s_uu - s_vv, sum = s_uu + s_vv;
    const double root = std::sqrt(diff * diff + 4.0 * (s_uv * s_uv));
    g.beam_bmaj = 1.0 / std::sqrt(2.0) / PI_D / std::sqrt(sum - root) / RPDEG_D;
    g.beam_bmin = 1.0 / std::sqrt(2.0) / PI_D / std::sqrt(sum + root) / RPDEG_D;
    g.beam_bpa = -0.5 * std::atan2(2.0 * s_uv, diff) / RPDEG_D;
  }
  if (vis_noise <= 0.0) vis_noise = 0.5f * sqrtf(variance);
  der.beam_bmaj_deg = g.beam_bmaj; der.beam_bmin_deg = g.beam_bmin; der.beam_bpa_deg = g.beam_bpa;
  der.sum_weights = sum_weights; der.vis_noise = vis_noise; der.total_visibilities = total_visibilities;
  if (visibilities) visibilities->setTotalVisibilities(total_visibilities);   // src/mfs.cu:553

  g.max_number_vis = 0;
  for (MSDataset& ds : datasets)
    g.max_number_vis = std::max(g.max_number_vis, ds.data.max_number_visibilities_in_channel_and_stokes);
  if (g.max_number_vis == 0) {
    std::printf("Max number of visibilities cannot be zero for image synthesis\n");
    std::exit(-1);
  }
  if (!g.quiet)
    std::printf("Estimated beam size: %e x %e (arcsec) / %lf (degrees)\n", g.beam_bmaj * 3600.0,
                g.beam_bmin * 3600.0, g.beam_bpa);
  g.beam_bmaj = g.beam_bmaj / std::fabs(g.DELTAX);  // to pixels
  g.beam_bmin = g.beam_bmin / std::fabs(g.DELTAX);
  g.noise_jypix = vis_noise / (PI_D * g.beam_bmaj * g.beam_bmin / (4.0 * logf(2.0)));
  der.noise_jypix = g.noise_jypix;

  // phase / pointing centres in pixels (src/mfs.cu:660-691; both are set from the PHASE centre)
  const double raimage = g.ra * RPDEG_D, decimage = g.dec * RPDEG_D;
  for (MSDataset& ds : datasets)
    for (Field& f : ds.fields) {
      double lphs, mphs;
      direccos(f.phs_ra, f.phs_dec, raimage, decimage, &lphs, &mphs);
      const double lpix = lphs / deltax, mpix = mphs / deltay;
      f.ref_xobs_pix = f.phs_xobs_pix = lpix + (g.crpix1 - 1.0f);
      f.ref_yobs_pix = f.phs_yobs_pix = mpix + (g.crpix2 - 1.0f);
      der.xobs_pix = f.phs_xobs_pix; der.yobs_pix = f.phs_yobs_pix;
      if (f.ref_xobs_pix < 0 || f.ref_xobs_pix >= g.M || f.ref_yobs_pix < 0 || f.ref_yobs_pix >= g.N) {
        std::printf("Dataset: %s\nPointing reference center (%f,%f) is outside the range of the image\n",
                    ds.name.c_str(), f.ref_xobs_pix, f.ref_yobs_pix);
        std::exit(0);  // goToError()
      }
    }

  // the engine exists since configure(); the visibility blocks go up now (varsPerGPU + device_visibilities)
  createEngine();
  shardAndUpload();

  // starting image (src/mfs.cu:742-750) and the Image object with its update rules (:798-811)
  const size_t MN = (size_t)g.M * g.N;
  host_I.assign(MN * g.image_count, 0.0f);
  for (int k = 0; k < g.image_count; k++) std::fill_n(host_I.begin() + MN * k, MN, g.initial_values[k]);
  device_Image = devAllocFloats(MN * g.image_count);
  devUpload(device_Image, host_I.data(), MN * g.image_count);
  image = new Image(device_Image, g.image_count);
  functionPtr = new imageMap[g.image_count];
  for (int i = 0; i < g.image_count; i++) {
    const bool positive = !g.nopositivity && i == 0;
    functionPtr[i].evaluateXt = positive ? particularEvaluateXt : defaultEvaluateXt;
    functionPtr[i].newP = positive ? particularNewP : defaultNewP;
  }
  image->setFunctionMapping(functionPtr);

  // noise image -> fg_scale, noise_cut (src/mfs.cu:850-916)
  if (gvm_num_channels(g.engine) == 0) {
    std::printf("ERROR: rank %d holds no visibility block\n", g.rank);
    gvm_dist_abort(g.engine);   // the peers must not wait in their first collective for a rank that is gone
    std::exit(-1);
  }
  float noise_min = 0.0f;
  // one attenuation^2 per (dataset, field), whatever was sharded to this rank (src/mfs.cu:870-888)
  std::vector<gvm_channel_desc> beam_fields;
  for (MSDataset& ds : datasets)
    for (Field& f : ds.fields) {
      gvm_channel_desc cd;
      std::memset(&cd, 0, sizeof(cd));
      cd.freq = g.nu_0;
      cd.antenna_diameter = ds.antennas[0].antenna_diameter;
      cd.pb_factor = ds.antennas[0].pb_factor;
      cd.pb_cutoff = ds.antennas[0].pb_cutoff;
      cd.primary_beam = ds.antennas[0].primary_beam;
      cd.ref_xobs_pix = f.ref_xobs_pix; cd.ref_yobs_pix = f.ref_yobs_pix;
      cd.phs_xobs_pix = f.phs_xobs_pix; cd.phs_yobs_pix = f.phs_yobs_pix;
      beam_fields.push_back(cd);
    }
  GVM_CHECK(gvm_build_noise_image_fields(g.engine, g.noise_jypix, (int)beam_fields.size(), beam_fields.data(), &noise_min));
  fg_scale = noise_min;
  g.noise_cut = g.noise_cut * noise_min;
  GVM_CHECK(gvm_set_scalars(g.engine, fg_scale, g.noise_cut, g.threshold));
  // -U / -M: the mask plane REPLACES the noise image after fg_scale and noise_cut were derived from it
  // (src/mfs.cu:922-931). -U: user plane, noise_cut was forced to 1 (x min noise) by getOptions;
  // -M: distance_image (src/functions.cu:2360-2380) of the last field: 1 everywhere, 0 within 4.5e-5
  // arcsec of the pointing-centre pixel.
  if (variables.user_mask != "NULL") {
    ioImageHandler->setMN(g.M, g.N);
    const std::vector<float> u_mask = ioImageHandler->read_data_float_FITS(variables.user_mask);
    GVM_CHECK(gvm_set_noise_image(g.engine, u_mask.data(), 0));
  } else if (g.radius_mask) {
    std::vector<float> dist_img(MN, 1.0f);
    const Field& lf = datasets.back().fields.back();
    const int x0 = (int)lf.ref_xobs_pix, y0 = (int)lf.ref_yobs_pix;
    for (long i = 0; i < g.N; i++)
      for (long j = 0; j < g.M; j++) {
        const float x = (float)((j - x0) * g.DELTAX * 3600.0), y = (float)((i - y0) * g.DELTAY * 3600.0);
        if (std::sqrt(x * x + y * y) < 4.5e-05f) dist_img[g.N * i + j] = 0.0f;
      }
    GVM_CHECK(gvm_set_noise_image(g.engine, dist_img.data(), 0));
  }
  if (gridding && ckernel) GVM_CHECK(gvm_set_gcf(g.engine, ckernel->getGCFCPUPointer()));
  der.fg_scale = fg_scale; der.noise_cut = g.noise_cut; der.nu_0 = g.nu_0;
  if (g.verbose_flag && !g.quiet) {
    std::printf("fg_scale = %e\n", fg_scale);
    std::printf("noise (Jy/pix) = %e\n", g.noise_jypix);
  }
  der.setup_seconds = wallSeconds() - t_start;
}

void MFS::useCKernelDegridding(bool on) {
  Globals& g = G();
  if (!on || !ckernel) {
    GVM_CHECK(gvm_set_degrid_kernel(g.engine, nullptr, 0, 0, 0, 0));
    if (!gridding) GVM_CHECK(gvm_set_gcf(g.engine, nullptr));
    return;
  }
  const double deltax = RPDEG_D * g.DELTAX, deltay = RPDEG_D * g.DELTAY;
  ckernel->setSigmas(std::fabs(g.deltau), std::fabs(g.deltav));
  ckernel->buildKernel();
  ckernel->initializeGCF(g.M, g.N, std::fabs(deltax), std::fabs(deltay));
  GVM_CHECK(gvm_set_degrid_kernel(g.engine, ckernel->getKernelPointer(), ckernel->getm(), ckernel->getn(),
                                  ckernel->getSupportX(), ckernel->getSupportY()));
  GVM_CHECK(gvm_set_gcf(g.engine, ckernel->getGCFCPUPointer()));
}

void MFS::clearRun() {
  Globals& g = G();
  devUpload(device_Image, host_I.data(), (size_t)g.M * g.N * g.image_count);
}

void MFS::run() {
  Globals& g = G();
  ObjectiveFunction* of = optimizer->getObjectiveFunction();
  of->setIo(ioImageHandler);
  Fi* chi2 = of->getFiByName("Chi2");
  if (chi2 && chi2->getNormalize()) fg_scale = 1.0f;
  if (chi2) chi2->setFgScale(fg_scale);
  if (gridding && chi2) chi2->setCKernel(ckernel);

  if (!g.quiet) std::printf("\n\nStarting optimizer\n");
  const double t0 = wallSeconds();
  if (!Order) {
    if (g.imagesChanged) {
      optimizer->setImage(image);
      optimizer->optimize();
    } else if (g.image_count == 2) {
      optimizer->setImage(image);
      for (int flag = 0; flag < 4; flag++) {
        optimizer->setFlag(flag);
        optimizer->optimize();
      }
    }
  } else {
    Order(optimizer, image);
  }
  GVM_CHECK(gvm_synchronize(g.engine));
  der.run_seconds = wallSeconds() - t0;

  const float chi2_final = chi2 ? chi2->get_fivalue() : 0.0f;
  Fi* entropy = of->getFiByName("Entropy");
  const float final_S = entropy ? entropy->get_fivalue() : 0.0f;
  const float lambda_S = entropy ? entropy->getPenalizationFactor() : 0.0f;
  auto report = [&](std::FILE* out) {
    std::fprintf(out, "Iterations: %d\n", optimizer->getCurrentIteration());
    std::fprintf(out, "chi2: %f\n", 2.0f * chi2_final);
    std::fprintf(out, "0.5*chi2: %f\n", chi2_final);
    std::fprintf(out, "Total visibilities: %d\n", total_visibilities);
    std::fprintf(out, "Reduced-chi2 (Num visibilities): %f\n", chi2_final / total_visibilities);
    std::fprintf(out, "Reduced-chi2 (Weights sum): %f\n", chi2_final / sum_weights);
    std::fprintf(out, "S: %f\n", final_S);
    std::fprintf(out, "Normalized S: %f\n", final_S / (g.M * g.N));
    std::fprintf(out, "lambda*S: %f\n", lambda_S * final_S);
    std::fprintf(out, "Wall time: %lf\n", der.run_seconds);
  };
  if (!g.quiet) {
    std::printf("Minimization ended successfully\n\n");
    report(stdout);
    std::printf("Objective evaluations: %ld, gradient evaluations: %ld\n\n", of->functionEvaluations(),
                of->gradientEvaluations());
  }
  if (variables.ofile != "NULL" && g.rank == 0) {
    std::FILE* out = std::fopen(variables.ofile.c_str(), "w");
    if (!out) {
      std::printf("Error opening output file!\n");
      std::exit(0);
    }
    report(out);
    std::fclose(out);
  }
}

void MFS::writeImages() {
  Globals& g = G();
  if (g.rank == 0) {
    std::printf("Saving final image to disk\n");
    if (IoOrderEnd) {
      IoOrderEnd(image->getImage(), ioImageHandler);
    } else {
      ioImageHandler->printImage(image->getImage(), out_image, "JY/PIXEL", optimizer->getCurrentIteration(), 0,
                                 fg_scale, g.M, g.N, true);
      if (g.print_images)
        ioImageHandler->printNotNormalizedImage(image->getImage(), "alpha.fits", "", optimizer->getCurrentIteration(), 1, true);
    }
  }
  // -E: error images (src/mfs.cu:1090-1113). Every rank takes part: the sums run over its shard.
  if (g.print_errors) {
    if (!error) error = createObject<Error, std::string>("SecondDerivateError");
    if (!g.quiet) std::printf("Calculating Error Images\n");
    error->calculateErrorImage(image, visibilities);
    if (g.rank != 0) return;
    if (IoOrderError) {
      IoOrderError(image->getErrorImage(), ioImageHandler);
    } else if (g.print_images) {
      ioImageHandler->printNotNormalizedImage(image->getErrorImage(), "error_Inu_0.fits", "JY/PIXEL",
                                              optimizer->getCurrentIteration(), 0, true);
      ioImageHandler->printNotNormalizedImage(image->getErrorImage(), "error_alpha_0.fits", "",
                                              optimizer->getCurrentIteration(), 1, true);
    }
  }
}

// calculateErrors (src/functions.cu:4966-5040) uses the residuals Vr the last chi2() left on the
// device; the error image is allocated here and owned by the Image, as in the reference.
void SecondDerivateError::calculateErrorImage(Image* I, Visibilities*) {
  Globals& g = G();
  if (!I->getErrorImage()) I->setErrorImage(devAllocFloats((size_t)I->getImageCount() * g.M * g.N));
  GVM_CHECK(gvm_error_maps(g.engine, I->getImage(), g.dist_kind, I->getErrorImage()));
}

// Residual / model write-back (src/mfs.cu:1115-1155). Ungridded runs: the scheme's backup weights come
// back (WeightingScheme::restoreWeights). Gridded runs: the ORIGINAL samples come back as
// getOriginalVisibilitiesBack leaves them (src/functions.cu:1844-2010) — uvw and Vo of the input, and the
// weights do_gridding backed up, i.e. the weights AFTER the weighting scheme (:1398-1409 overwrite the
// scheme's own backup) — and one more Chi2::calcFi re-samples the model at the original (u,v) with the
// bilinear degridder; the Chi2 term still carries the CKernel, so apply_GCF stays in that evaluation
// (src/mfs.cu:989-992, src/functions.cu:4358). Then modelToHost (src/MSFITSIO.cu:1114-1138): Vm comes back and
// is CONJUGATED where the host u > 0 (the device samples were folded by hermitianSymmetry, the host ones were
// not), and the residual that goes to the file is host Vo - that Vm (writeMS, :1217).
void MFS::writeResiduals() {
  Globals& g = G();
  if (!g.quiet) std::printf("Transferring residuals to host memory\n");
  Fi* chi2 = optimizer->getObjectiveFunction()->getFiByName("Chi2");
  nongridded_chi2 = 0.0f;
  if (gridding && !ungridded.empty()) {
    // the originals come back (no copy); they take the pointing / phase-centre pixels setDevice derived for the
    // gridded shells (the reference keeps ONE Field object per field, so its centres survive the swap)
    for (size_t d = 0; d < ungridded.size() && d < datasets.size(); d++)
      for (size_t f = 0; f < ungridded[d].fields.size() && f < datasets[d].fields.size(); f++) {
        Field& dst = ungridded[d].fields[f];
        const Field& src = datasets[d].fields[f];
        dst.ref_xobs_pix = src.ref_xobs_pix; dst.ref_yobs_pix = src.ref_yobs_pix;
        dst.phs_xobs_pix = src.phs_xobs_pix; dst.phs_yobs_pix = src.phs_yobs_pix;
      }
    datasets = std::move(ungridded);   // a second call finds them in place
    ungridded.clear();
    datasets_are_gridded = false;      // the originals shard over the ranks again (chunks / channels)
    GVM_CHECK(gvm_clear_channels(g.engine));
    shardAndUpload();
    if (chi2) {
      nongridded_chi2 = chi2->calcFi(image->getImage());
      if (!g.quiet) std::printf("Non-gridded chi2 after de-gridding using bilinear interpolation %f\n", nongridded_chi2);
    }
  } else if (!gridding) {
    scheme->restoreWeights(datasets);
  }
  for (MSDataset& ds : datasets)
    for (Field& f : ds.fields)
      for (size_t i = 0; i < f.visibilities.size(); i++)
        for (size_t s = 0; s < f.visibilities[i].size(); s++) {
          const int slot = f.engine_slot.empty() ? -1 : f.engine_slot[i][s];
          if (slot < 0) continue;
          HVis& v = f.visibilities[i][s];
          const size_t Z = (size_t)gvm_channel_nvis(g.engine, slot);
          size_t lo = 0;
          if (g.world > 1 && Z != v.size()) lo = v.size() * (size_t)g.rank / g.world;
          v.Vm.assign(2 * v.size(), 0.0f);
          v.Vr.assign(2 * v.size(), 0.0f);
          GVM_CHECK(gvm_get_vis(g.engine, slot, nullptr, nullptr, nullptr, v.Vm.data() + 2 * lo, nullptr, nullptr));
          for (size_t k = lo; k < lo + Z; k++) {
            if (v.uvw[3 * k] > 0.0) v.Vm[2 * k + 1] = -v.Vm[2 * k + 1];
            v.Vr[2 * k] = v.Vo[2 * k] - v.Vm[2 * k];
            v.Vr[2 * k + 1] = v.Vo[2 * k + 1] - v.Vm[2 * k + 1];
          }
        }
  if (msoutput != "NULL") {
    const std::vector<std::string> outs = countAndSeparateStrings(msoutput, ",");
    for (size_t d = 0; d < datasets.size() && d < outs.size(); d++) {
      const std::string file = g.world > 1 ? outs[d] + ".rank" + std::to_string(g.rank) : outs[d];
      ioVisibilitiesHandler->writeModelVisibilities(file, datasets[d].fields, datasets[d].data);
    }
    if (!g.quiet) std::printf("Residuals and model visibilities saved.\n");
  }
}

void MFS::unSetDevice() {
  Globals& g = G();
  hostProfileReport();
  if (!g.quiet) std::printf("Freeing device memory\n");
  gvm_grid_release();
  if (g.engine) {
    devFree(device_Image);
    device_Image = nullptr;
    if (image && image->getErrorImage()) { devFree(image->getErrorImage()); image->setErrorImage(nullptr); }
    GVM_CHECK(gvm_destroy(g.engine));
    g.engine = nullptr;
  }
  delete image;
  image = nullptr;
  delete[] functionPtr;
  functionPtr = nullptr;
}


// ---- Filter "Gridding" (src/gridding.cu) ------------------------------------------------------
void gridDatasetsInPlace(std::vector<MSDataset>& datasets, CKernel* ckernel) {
  Globals& g = G();
  for (MSDataset& ds : datasets) {
    int max = 0;
    for (Field& f : ds.fields)
      for (size_t i = 0; i < f.visibilities.size(); i++) {
        long per_freq = 0;
        for (size_t s = 0; s < f.visibilities[i].size(); s++) {
          HVis& v = f.visibilities[i][s];
          int64_t nout = 0;
          GVM_CHECK(gvm_grid_block(g.firstgpu, g.M, g.N, g.deltau, g.deltav, f.nu[i], (int64_t)v.size(), v.uvw.data(),
                                   v.Vo.data(), v.weight.data(), ckernel->getKernelPointer(), ckernel->getm(),
                                   ckernel->getn(), ckernel->getSupportX(), ckernel->getSupportY(), nullptr, nullptr,
                                   nullptr, &nout));
          v.uvw.resize(3 * nout);
          v.Vo.resize(2 * nout);
          v.weight.resize(nout);
          GVM_CHECK(gvm_grid_fetch(v.uvw.data(), v.Vo.data(), v.weight.data()));
          v.Vm.assign(2 * nout, 0.0f);
          v.Vr.assign(2 * nout, 0.0f);
          if (i < f.numVisibilitiesPerFreqPerStoke.size() && s < f.numVisibilitiesPerFreqPerStoke[i].size())
            f.numVisibilitiesPerFreqPerStoke[i][s] = (long)nout;
          per_freq += (long)nout;
          max = std::max<long>(max, (long)nout);
        }
        if (i < f.numVisibilitiesPerFreq.size()) f.numVisibilitiesPerFreq[i] = per_freq;
      }
    ds.data.max_number_visibilities_in_channel_and_stokes = max;
  }
  GVM_CHECK(gvm_grid_release());
}

void Gridding::setThreadsChecked(int t) {
  if (t != 1 && t >= 1) threads = t;
  else if (t != 1) std::printf("Number of threads set to 1\n");
}

void Gridding::applyCriteria(Visibilities* v) {
  Globals& g = G();
  if (g.deltau == 0.0 || g.deltav == 0.0) {
    std::printf("ERROR: Gridding::applyCriteria needs the uv cell size (MFS::configure first)\n");
    std::exit(-1);
  }
  PillBox2D fallback;
  CKernel* ck = ckernel ? ckernel : &fallback;
  ck->setSigmas(std::fabs(g.deltau), std::fabs(g.deltav));
  ck->buildKernel();
  gridDatasetsInPlace(v->getMSDataset(), ck);
}

namespace {
Synthesizer* makeMFS() { return new MFS; }
const bool kRegistered = registerCreationFunction<Synthesizer, std::string>("MFS", makeMFS);
Error* makeSecondDerivateError() { return new SecondDerivateError; }
const bool kRegisteredError = registerCreationFunction<Error, std::string>("SecondDerivateError", makeSecondDerivateError);
Filter* makeGridding() { return new Gridding; }
const bool kRegisteredGridding = registerCreationFunction<Filter, std::string>("Gridding", makeGridding);
}  // namespace

}  // namespace gpuvmem
