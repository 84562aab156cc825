#include "weightingscheme.hpp"

#include "globals.hpp"

namespace gpuvmem {

void WeightingScheme::restoreWeights(std::vector<MSDataset>& d) {
  for (auto& ds : d)
    for (auto& f : ds.fields)
      for (size_t i = 0; i < f.visibilities.size(); i++)
        for (size_t s = 0; s < f.visibilities[i].size(); s++)
          if (i < f.backup_visibilities.size() && s < f.backup_visibilities[i].size())
            f.visibilities[i][s].weight = f.backup_visibilities[i][s].weight;
}

void WeightingScheme::applyOnGpu(int scheme, float robust, std::vector<MSDataset>& d, const char* label) {
  Globals& g = G();
  if (!g.quiet) std::cout << "Running " << label << " weighting scheme on GPU " << g.firstgpu << std::endl;
  std::vector<int64_t> Z;
  std::vector<const double*> uvw;
  std::vector<float> freqs;
  std::vector<float*> w;
  for (auto& ds : d)
    for (auto& f : ds.fields) {
      f.backup_visibilities.resize(f.visibilities.size());
      for (size_t i = 0; i < f.visibilities.size(); i++) {
        f.backup_visibilities[i].resize(f.visibilities[i].size());
        for (size_t s = 0; s < f.visibilities[i].size(); s++) {
          HVis& v = f.visibilities[i][s];
          Z.push_back((int64_t)v.size());
          uvw.push_back(v.uvw.data());
          freqs.push_back(f.nu[i]);
          w.push_back(v.weight.data());
          // backup_visibilities: the weights before the scheme (one copy, taken now), or the new ones with -W
          if (!modify_weights) f.backup_visibilities[i][s].weight = v.weight;
        }
      }
    }
  gvm_taper taper;
  if (uvtaper) taper = uvtaper->abi();
  // several ranks: every rank weighs a slice of every block, the cell sums travel down the ranks in sample order
  // (bit-identical result, gvm_weights_dist); the engine and its communicator exist since MFS::configure
  if (g.engine && g.world > 1)
    GVM_CHECK(gvm_weights_dist(g.engine, scheme, robust, (int)Z.size(), Z.data(), uvw.data(), freqs.data(), w.data(),
                               uvtaper ? &taper : nullptr));
  else
    GVM_CHECK(gvm_weights(g.firstgpu, scheme, robust, g.M, g.N, g.deltau, g.deltav, (int)Z.size(), Z.data(),
                          uvw.data(), freqs.data(), w.data(), uvtaper ? &taper : nullptr));
  if (modify_weights)
    for (auto& ds : d)
      for (auto& f : ds.fields)
        for (size_t i = 0; i < f.visibilities.size(); i++)
          for (size_t s = 0; s < f.visibilities[i].size(); s++)
            f.backup_visibilities[i][s].weight = f.visibilities[i][s].weight;
}

void BriggsWeightingScheme::setRobustParam(float r) {
  if (r >= -2.0f && r <= 2.0f) {
    robust_param = r;
  } else {
    std::cout << "Error. Robust parameter must have values between -2.0 and 2.0" << std::endl;
    std::exit(-1);
  }
}
void BriggsWeightingScheme::configure(void* params) {
  setRobustParam(*static_cast<float*>(params));
  if (!G().quiet) std::cout << "Using robust " << robust_param << " for Briggs weighting" << std::endl;
}

namespace {
WeightingScheme* makeNatural() { return new NaturalWeightingScheme; }
WeightingScheme* makeUniform() { return new UniformWeightingScheme; }
WeightingScheme* makeBriggs() { return new BriggsWeightingScheme; }
WeightingScheme* makeRadial() { return new RadialWeightingScheme; }
const bool kRegistered[] = {
    registerCreationFunction<WeightingScheme, std::string>("Natural", makeNatural),
    registerCreationFunction<WeightingScheme, std::string>("Uniform", makeUniform),
    registerCreationFunction<WeightingScheme, std::string>("Briggs", makeBriggs),
    registerCreationFunction<WeightingScheme, std::string>("Radial", makeRadial),
};
}  // namespace

}  // namespace gpuvmem
