// visibilities.hpp — the Visibilities holder and the Filter plugin family of the reference
// (include/classes/visibilities.cuh:7-36, include/classes/filter.cuh:4-8, include/gridding.cuh:4-15,
// src/gridding.cu; factory key "Gridding"). The reference's Visibilities COPIES the dataset vector in
// setMSDataset/getMSDataset (gigabytes at the BASELINE sizes); here it refers to the synthesizer's
// datasets, so filters act on the data the run uses.
#pragma once
#include <vector>

#include "ckernel.hpp"
#include "factory.hpp"
#include "msdata.hpp"
#include "weightingscheme.hpp"

namespace gpuvmem {

class Visibilities {
 public:
  void setMSDataset(std::vector<MSDataset>& d) { datasets = &d; }
  void setTotalVisibilities(int t) { total_visibilities = t; }
  void setNDatasets(int t) { ndatasets = t; }
  void setMaxNumberVis(int t) { max_number_vis = t; }
  std::vector<MSDataset>& getMSDataset() { return *datasets; }
  int getTotalVisibilities() const { return total_visibilities; }
  int getMaxNumberVis() const { return max_number_vis; }
  int getNDatasets() const { return ndatasets; }
  void applyWeightingScheme(WeightingScheme* scheme) { scheme->apply(*datasets); }

 private:
  std::vector<MSDataset>* datasets = nullptr;
  int ndatasets = 0, total_visibilities = 0, max_number_vis = 0;
};

class Filter {
 public:
  virtual ~Filter() = default;
  virtual void applyCriteria(Visibilities* v) = 0;
  virtual void configure(void* params) = 0;
};

// do_gridding over every dataset (src/gridding.cu:13-28). The reference passes a NULL kernel there (it
// would dereference it); here the kernel is the one given with setCKernel, a PillBox2D otherwise — the
// reference's default gridding kernel (src/main.cu:152). The thread count is kept for the surface only:
// gridding runs on the GPU.
class Gridding : public Filter {
 public:
  Gridding() = default;
  explicit Gridding(int threads_) { setThreadsChecked(threads_); }
  void applyCriteria(Visibilities* v) override;
  void configure(void* params) override { setThreadsChecked(*static_cast<int*>(params)); }
  void setThreads(int t) { threads = t; }
  int getThreads() const { return threads; }
  void setCKernel(CKernel* ck) { ckernel = ck; }

 private:
  void setThreadsChecked(int t);
  int threads = 1;
  CKernel* ckernel = nullptr;
};

// do_gridding for every (field, channel, stokes) block of `datasets`, in place (the gridded samples
// replace the block). Shared by MFS::doGridding and the Gridding filter.
void gridDatasetsInPlace(std::vector<MSDataset>& datasets, CKernel* ckernel);

}  // namespace gpuvmem
