#include "objectivefunction.hpp"

#include <chrono>
#include <cstdint>
#include <cstdlib>

namespace gpuvmem {

namespace {
double nowS() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

bool ObjectiveFunction::defaultSingleSync() {
  const char* v = std::getenv("GVM_SINGLE_SYNC");
  return !(v && *v == '0');
}

bool ObjectiveFunction::defaultFusedGradient() {
  const char* v = std::getenv("GVM_FUSED_GRADIENT");
  return !(v && *v == '0');
}

bool ObjectiveFunction::defaultGraphs() {
  const char* v = std::getenv("GVM_GRAPHS");
  return !(v && *v == '0');
}

void ObjectiveFunction::dropGraphs() {
  for (GraphEntry& g : graph_cache)
    if (g.exec && G().engine) gvm_graph_destroy(G().engine, g.exec);
  graph_cache.clear();
}

ObjectiveFunction::~ObjectiveFunction() {
  dropGraphs();
  if (G().engine) devFree(dphi);
}

void ObjectiveFunction::addFi(Fi* fi) {
  if (fi->getPenalizationFactor()) {
    fis.push_back(fi);
    fi_values.push_back(0.0f);
  }
}

float ObjectiveFunction::calcFunction(float* p) {
  const double t0 = nowS();
  float value = 0.0f;
  size_t k = 0;
  // Fast path: every term launches its value asynchronously into a result slot of the engine and ONE
  // stream synchronisation brings them all back (the reference synchronises several times per term).
  // Terms that do not support it (user plugins) switch the whole evaluation to the reference's loop.
  bool fast = single_sync && fis.size() <= (size_t)GVM_OBJ_SLOTS;
  // Graph path: an evaluation is identified by the image pointer, every term's state and the engine's epoch.
  // First sight: plain launches (buffers that are created on first use appear). Second sight: the same launches
  // are captured into a CUDA graph. From then on: one cudaGraphLaunch + one synchronisation per evaluation —
  // the line search probes the same work buffer hundreds of times (src/f1dim.cu:49-80).
  GraphEntry* entry = nullptr;
  if (fast && graphs && G().world <= 1) {
    uint64_t key = 1469598103934665603ull;
    auto fold = [&key](uint64_t v) { key = (key ^ v) * 1099511628211ull; };
    fold((uint64_t)(uintptr_t)p);
    fold((uint64_t)gvm_state_epoch(G().engine));
    for (Fi* fi : fis) fold(fi->stateKey());
    for (GraphEntry& g : graph_cache)
      if (g.key == key) entry = &g;
    if (!entry) {
      if (graph_cache.size() >= 8) {   // evict the least recently used
        size_t lru = 0;
        for (size_t i = 1; i < graph_cache.size(); i++)
          if (graph_cache[i].last_use < graph_cache[lru].last_use) lru = i;
        if (graph_cache[lru].exec) gvm_graph_destroy(G().engine, graph_cache[lru].exec);
        graph_cache.erase(graph_cache.begin() + lru);
      }
      graph_cache.push_back(GraphEntry());
      entry = &graph_cache.back();
      entry->key = key;
    }
    entry->seen++;
    entry->last_use = n_function;
  }
  if (entry && entry->exec) {
    for (Fi* fi : fis) fi->noteEnqueued();
    GVM_CHECK(gvm_graph_launch(G().engine, entry->exec));
    n_replays++;
  } else {
    const bool capture = entry && entry->seen == 2;
    if (capture && gvm_graph_begin(G().engine) != 0) entry = nullptr;
    size_t enq = 0;
    for (; fast && enq < fis.size(); enq++) fast = fis[enq]->enqueueFi(p, (int)enq);
    if (capture && entry) {
      if (fast) GVM_CHECK(gvm_fetch_slots_enqueue(G().engine, (int)fis.size()));
      void* exec = nullptr;
      if (gvm_graph_end(G().engine, &exec) == 0 && fast) {
        entry->exec = exec;
        GVM_CHECK(gvm_graph_launch(G().engine, exec));   // the capture executed nothing
      } else {
        // not capturable (a term allocated or synchronised): stay on plain launches for this key
        if (exec) gvm_graph_destroy(G().engine, exec);
        entry->seen = 3;
        if (fast) {
          enq = 0;
          for (; fast && enq < fis.size(); enq++) fast = fis[enq]->enqueueFi(p, (int)enq);
          if (fast) GVM_CHECK(gvm_fetch_slots_enqueue(G().engine, (int)fis.size()));
        }
      }
    } else if (fast) {
      GVM_CHECK(gvm_fetch_slots_enqueue(G().engine, (int)fis.size()));
    }
  }
  if (fast) {
    double vals[GVM_OBJ_SLOTS];
    GVM_CHECK(gvm_fetch_slots_wait(G().engine, (int)fis.size(), vals));
    for (Fi* fi : fis) {
      const float term = fi->finishFi((float)vals[k]);
      fi_values[k++] = fi->get_fivalue();
      value += term;
    }
  } else {
    for (Fi* fi : fis) {
      const float term = fi->calcFi(p);
      fi_values[k++] = fi->get_fivalue();
      value += term;
    }
  }
  n_function++;
  t_function += nowS() - t0;
  return value;
}

void ObjectiveFunction::calcGradient(float* p, float* xi, int iter) {
  const double t0 = nowS();
  if (io && io->getPrintImages()) {
    if (IoOrderIterations) {
      IoOrderIterations(p, io);
    } else {
      io->printImageIteration(p, "I_nu_0", "JY/PIXEL", iter, 0, true);
      io->printImageIteration(p, "alpha", "JY/PIXEL", iter, 1, true);
    }
  }
  // Fast path: every term writes straight into xi (zeroed once) — the chi2 gradient accumulates over the channels
  // there, every prior adds lambda * dS in one fused pass. The reference's loop moves 16 image planes more per
  // evaluation (zeroing and filling a device_DS per term, result_dchi2 -> dphi, dphi -> xi); the values are the
  // same, bit for bit. Terms that do not support it (user plugins, Chi2 not first) keep the reference's loop.
  bool fused = fused_gradient && p != xi;
  if (fused) {
    for (Fi* fi : fis) fi->setIteration(iter);
    devZero(xi, (size_t)M * N * image_count);
    size_t k = 0;
    for (; fused && k < fis.size(); k++) fused = fis[k]->gradInto(p, xi, k == 0);
  }
  if (!fused) {
    restartDPhi();
    for (Fi* fi : fis) {
      fi->setIteration(iter);
      fi->calcGi(p, xi);
      fi->addToDphi(dphi);
    }
    copyDphiToXi(xi);
  }
  GVM_CHECK(gvm_synchronize(G().engine));
  n_gradient++;
  t_gradient += nowS() - t0;
}

void ObjectiveFunction::restartDPhi() {
  for (Fi* fi : fis) fi->restartDGi();
  devZero(dphi, (size_t)M * N * image_count);
}

void ObjectiveFunction::copyDphiToXi(float* xi) { devCopyD2D(xi, dphi, (size_t)M * N * image_count); }

Fi* ObjectiveFunction::getFiByName(const std::string& fi_name) {
  for (Fi* fi : fis)
    if (fi->getName() == fi_name) return fi;
  return nullptr;
}

void ObjectiveFunction::configure(long N_, long M_, int I) {
  setN(N_);
  setM(M_);
  setImageCount(I);
  if (dphi) devFree(dphi);
  dphi = devAllocFloats((size_t)M * N * I);
}

namespace {
ObjectiveFunction* makeObjectiveFunction() { return new ObjectiveFunction; }
const bool kRegistered =
    registerCreationFunction<ObjectiveFunction, std::string>("ObjectiveFunction", makeObjectiveFunction);
}  // namespace

}  // namespace gpuvmem
