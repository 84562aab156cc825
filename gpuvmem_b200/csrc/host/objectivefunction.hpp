// objectivefunction.hpp — Phi(I) = sum_i lambda_i f_i(I) and its gradient. Surface of the
// reference's include/classes/objectivefunction.cuh:6-112 (factory key "ObjectiveFunction").
#pragma once
#include <string>
#include <vector>

#include "factory.hpp"
#include "fi.hpp"
#include "io.hpp"

namespace gpuvmem {

class ObjectiveFunction {
 public:
  ObjectiveFunction() = default;
  ~ObjectiveFunction();
  // terms with a zero factor are dropped (objectivefunction.cuh:9-14)
  void addFi(Fi* fi);
  float calcFunction(float* p);
  // restartDPhi, then per term setIteration/calcGi/addToDphi, then dphi -> xi (:28-45)
  void calcGradient(float* p, float* xi, int iter);
  void restartDPhi();
  void copyDphiToXi(float* xi);
  std::vector<Fi*> getFi() { return fis; }
  Fi* getFiByName(const std::string& fi_name);
  void setN(long n) { N = n; }
  void setM(long m) { M = m; }
  void setImageCount(int I) { image_count = I; }
  void setIo(Io* i) { io = i; }
  void setIoOrderIterations(void (*func)(float* I, Io* io)) { IoOrderIterations = func; }
  void configure(long N, long M, int I);
  std::vector<float> get_fi_values() { return fi_values; }
  // evaluation counters (bench / statistics block)
  long functionEvaluations() const { return n_function; }
  long gradientEvaluations() const { return n_gradient; }
  // host wall time spent inside calcFunction / calcGradient (the calls return after the GPU work)
  double functionSeconds() const { return t_function; }
  double gradientSeconds() const { return t_gradient; }
  // one host synchronisation per calcFunction (Fi::enqueueFi) instead of one per term; on by default,
  // GVM_SINGLE_SYNC=0 in the environment or setSingleSync(false) selects the reference's per-term loop
  void setSingleSync(bool on) { single_sync = on; }
  // one CUDA graph per objective evaluation: the second evaluation with the same image pointer and term state is
  // captured, later ones replay it with a single launch (GVM_GRAPHS=0 or setGraphs(false): plain launches)
  void setGraphs(bool on) { graphs = on; }
  // gradient terms written straight into xi (Fi::gradInto); GVM_FUSED_GRADIENT=0 or setFusedGradient(false): the
  // reference's restartDGi / calcGi / addToDphi loop
  void setFusedGradient(bool on) { fused_gradient = on; }
  long graphReplays() const { return n_replays; }

 private:
  std::vector<Fi*> fis;
  std::vector<float> fi_values;
  Io* io = nullptr;
  float* dphi = nullptr;
  long N = 0, M = 0;
  void (*IoOrderIterations)(float* I, Io* io) = nullptr;
  int image_count = 1;
  long n_function = 0, n_gradient = 0;
  double t_function = 0.0, t_gradient = 0.0;
  bool single_sync = defaultSingleSync();
  static bool defaultSingleSync();
  // captured objective evaluations, keyed by (image pointer, term states, engine epoch)
  struct GraphEntry { uint64_t key = 0; void* exec = nullptr; int seen = 0; long last_use = 0; };
  std::vector<GraphEntry> graph_cache;
  bool graphs = defaultGraphs();
  long n_replays = 0;
  static bool defaultGraphs();
  bool fused_gradient = defaultFusedGradient();
  static bool defaultFusedGradient();
  void dropGraphs();
};

}  // namespace gpuvmem
