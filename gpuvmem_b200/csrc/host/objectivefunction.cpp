#include "objectivefunction.hpp"

#include <chrono>
#include <cstdlib>

namespace gpuvmem {

namespace {
double nowS() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

bool ObjectiveFunction::defaultSingleSync() {
  const char* v = std::getenv("GVM_SINGLE_SYNC");
  return !(v && *v == '0');
}

ObjectiveFunction::~ObjectiveFunction() {
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
  size_t enq = 0;
  for (; fast && enq < fis.size(); enq++) fast = fis[enq]->enqueueFi(p, (int)enq);
  if (fast) {
    double vals[GVM_OBJ_SLOTS];
    GVM_CHECK(gvm_fetch_slots(G().engine, (int)fis.size(), vals));
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
  restartDPhi();
  for (Fi* fi : fis) {
    fi->setIteration(iter);
    fi->calcGi(p, xi);
    fi->addToDphi(dphi);
  }
  copyDphiToXi(xi);
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
