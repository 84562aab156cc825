// synthesizer.hpp — Synthesizer surface (reference include/classes/synthesizer.cuh) and its
// one implementation, MFS (src/mfs.cu, include/mfs.cuh; factory key "MFS"): flag parsing,
// weights, optional gridding, device upload, noise/mask image, optimizer run, outputs.
#pragma once
#include <string>
#include <vector>

#include "ckernel.hpp"
#include "error.hpp"
#include "factory.hpp"
#include "image.hpp"
#include "io.hpp"
#include "optimizer.hpp"
#include "visibilities.hpp"
#include "weightingscheme.hpp"

namespace gpuvmem {

// the parsed command line (reference: Vars, include/functions.cuh; getOptions src/functions.cu:181-301)
struct Vars {
  std::string input = "NULL", output = "NULL", output_image = "mod_out.fits", modin = "mod_in_0.fits";
  std::string path = "mem/", gpus = "0", ofile = "NULL", initial_values = "NULL";
  std::string penalization_factors = "NULL", user_mask = "NULL";
  float noise = -1.0f, eta = -1.0f, noise_cut = 10.0f, nu_0 = -1.0f, threshold = 0.0f, randoms = 1.0f;
  float robust_param = 2.0f;
  int blockSizeX = -1, blockSizeY = -1, blockSizeV = -1, it_max = 500, gridding = 0;
  // engine extension (not in the reference): gradient kernel selection, GVM_GRAD_*
  int grad_mode = 0;
};
// Parses the reference's flags (-i -o -O -m -n -e -N -F -T -p -G -r -R -f -X -Y -V -t -g -z -Z -U
// and the boolean -v -x -a -P -E -s -M -W -h -w -c, long names as in the reference). Returns false
// (after printing the help) where the reference exits.
bool getOptions(int argc, char** argv, Vars* out);
void print_help();

// The part [lo, hi) of a block of Z visibilities in channel `chan` that rank `rank` of `world`
// holds; false when the block lives elsewhere. Whole channels go to rank chan % world (the
// reference's rule, src/functions.cu:4341) when the job has at least `world` channels; otherwise
// every block is cut into `world` contiguous chunks.
bool shardRange(int max_nfreq, int chan, size_t Z, int rank, int world, size_t* lo, size_t* hi);

class Synthesizer {
 public:
  virtual ~Synthesizer() = default;
  virtual void run() = 0;
  virtual void setOutPut(char*) {}
  virtual void setDevice() = 0;
  virtual void unSetDevice() = 0;
  virtual std::vector<std::string> countAndSeparateStrings(std::string long_str, std::string sep) = 0;
  virtual void configure(int argc, char** argv) = 0;
  virtual void writeImages() = 0;
  virtual void clearRun() = 0;
  virtual void writeResiduals() = 0;

  void setError(Error* e) { error = e; }
  void setVisibilities(Visibilities* v) { visibilities = v; }
  Visibilities* getVisibilities() { return visibilities; }
  void setOptimizator(Optimizer* min) { optimizer = min; }
  void setIoImageHandler(Io* h) { ioImageHandler = h; }
  void setIoVisibilitiesHandler(Io* h) { ioVisibilitiesHandler = h; }
  void setWeightingScheme(WeightingScheme* s) { scheme = s; }
  void setOrder(void (*func)(Optimizer* o, Image* I)) { Order = func; }
  Image* getImage() { return image; }
  void setImage(Image* i) { image = i; }
  void setIoOrderEnd(void (*func)(float* I, Io* io)) { IoOrderEnd = func; }
  void setIoOrderError(void (*func)(float* I, Io* io)) { IoOrderError = func; }
  void setIoOrderIterations(void (*func)(float* I, Io* io)) { IoOrderIterations = func; }
  Optimizer* getOptimizator() { return optimizer; }
  void setGriddingKernel(CKernel* ck) { ckernel = ck; }
  bool getGridding() const { return gridding; }
  void setGridding(bool gr) { gridding = gr; }
  int getGriddingThreads() const { return griddingThreads; }
  void setGriddingThreads(int t) {
    if (t > 0) { griddingThreads = t; gridding = true; }
    else if (t < 0) std::cout << "Gridding threads cannot be less than 0" << std::endl;
  }
  float getVisNoise() const { return vis_noise; }
  void setVisNoise(float n) { vis_noise = n; }
  float getFgScale() const { return fg_scale; }
  void setFgScale(float s) { fg_scale = s; }

 protected:
  Image* image = nullptr;
  Optimizer* optimizer = nullptr;
  CKernel* ckernel = nullptr;
  Error* error = nullptr;
  Visibilities* visibilities = nullptr;
  Io* ioImageHandler = nullptr;
  Io* ioVisibilitiesHandler = nullptr;
  WeightingScheme* scheme = nullptr;
  void (*Order)(Optimizer* o, Image* I) = nullptr;
  void (*IoOrderEnd)(float* I, Io* io) = nullptr;
  void (*IoOrderError)(float* I, Io* io) = nullptr;
  void (*IoOrderIterations)(float* I, Io* io) = nullptr;
  bool gridding = false;
  int griddingThreads = 0;
  float vis_noise = -1.0f, fg_scale = 1.0f;
};

class MFS : public Synthesizer {
 public:
  ~MFS() override;
  void configure(int argc, char** argv) override;
  void setDevice() override;
  void run() override;
  void clearRun() override;
  void writeImages() override;
  void writeResiduals() override;
  void createEngine();
  void unSetDevice() override;
  std::vector<std::string> countAndSeparateStrings(std::string long_str, std::string sep) override;

  // library callers: hand over datasets that are already in memory instead of -i/-m files
  void adoptDatasets(std::vector<MSDataset>&& ds, const headerValues& header);
  // one process per GPU: this process's rank / world size and the NCCL id from rank 0
  void setDistributed(int rank, int world, const std::string& nccl_id);
  std::vector<MSDataset>& getDatasets() { return datasets; }
  float getNonGriddedChi2() const { return nongridded_chi2; }
  // Forward-model option: the gridding kernel handed to setGriddingKernel is used as a
  // convolutional DEGRIDDING kernel on the ungridded samples (degriddingGPU, src/functions.cu:2205-2254) with its
  // gridding-correction image in front of the FFT, instead of the bilinear vis_mod. Call after setDevice.
  void useCKernelDegridding(bool on);
  // scalars derived in configure/setDevice (parity tests compare them with the reference's)
  struct Derived {
    double beam_bmaj_deg = 0, beam_bmin_deg = 0, beam_bpa_deg = 0, deltau = 0, deltav = 0;
    float sum_weights = 0, vis_noise = 0, noise_jypix = 0, fg_scale = 0, noise_cut = 0, nu_0 = 0;
    float xobs_pix = 0, yobs_pix = 0;
    int total_visibilities = 0;
    double setup_seconds = 0, run_seconds = 0, gridding_seconds = 0, weighting_seconds = 0;
  };
  const Derived& derived() const { return der; }
  const Vars& vars() const { return variables; }

 private:
  void shardAndUpload();
  void doGridding();
  Vars variables;
  std::vector<MSDataset> datasets;
  bool datasets_are_gridded = false;  // `datasets` currently holds the output of doGridding
  float nongridded_chi2 = 0.0f;       // the "Non-gridded chi2" of the last writeResiduals (gridded runs)
  std::vector<MSDataset> ungridded;   // originals kept when -g replaces them (residual write-back)
  headerValues header;
  bool adopted = false;
  std::string nccl_id;
  std::vector<float> host_I;
  float* device_Image = nullptr;
  imageMap* functionPtr = nullptr;
  std::string msinput, msoutput, modinput, out_image;
  float sum_weights = 0.0f;
  int total_visibilities = 0;
  Derived der;
  double t_start = 0;
};

}  // namespace gpuvmem
