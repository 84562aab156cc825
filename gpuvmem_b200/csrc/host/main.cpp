// main.cpp — the gpuvmem command line on the B200 engine. Mirrors the object graph of the
// reference's src/main.cu:100-229: factories by name, MFS::configure(argc, argv), setDevice,
// the five Fi terms with main.cu's penalizer indices, run, writeImages, writeResiduals.
// Input is the GVMS container (gpuvmem_b200/synth.py) instead of a Measurement Set + FITS.
// Multi-GPU: start one process per GPU with RANK / WORLD_SIZE / LOCAL_RANK in the environment
// (e.g. torchrun --no-python, or mpirun exporting them); rank 0 publishes the NCCL id through
// a per-launch file (GVM_RENDEZVOUS, default /tmp/gvm_nccl_<uid>_<MASTER_PORT or 29500>_<launch nonce>.id)
// that rank 0 removes as soon as the communicator exists.
#include <fcntl.h>
#include <unistd.h>

#include <cctype>
#include <cerrno>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "synthesizer.hpp"

using namespace gpuvmem;

namespace {
void optimizationOrder(Optimizer* optimizer, Image* image) {
  optimizer->setImage(image);
  optimizer->setFlag(0);
  optimizer->optimize();
}
int envInt(const char* name, int fallback) {
  const char* v = std::getenv(name);
  return v && *v ? std::atoi(v) : fallback;
}
// The NCCL id travels through a file that belongs to THIS launch: its name carries a per-launch nonce
// (GVM_RUN_ID, else TORCHELASTIC_RUN_ID, else the launcher's pid — every rank of one launch has the same
// parent) and the uid, it is created exclusively with mode 0600 (no symlink games in a world-writable /tmp),
// and rank 0 removes it once the communicator exists (rendezvousDone), so a later job never reads a stale id.
std::string rendezvousPath() {
  if (const char* named = std::getenv("GVM_RENDEZVOUS")) return named;
  std::string nonce;
  if (const char* v = std::getenv("GVM_RUN_ID")) nonce = v;
  else if (const char* t = std::getenv("TORCHELASTIC_RUN_ID")) nonce = t;
  else nonce = "ppid" + std::to_string((long)getppid());
  for (char& ch : nonce)
    if (!std::isalnum((unsigned char)ch) && ch != '-' && ch != '_') ch = '_';
  return "/tmp/gvm_nccl_" + std::to_string((long)getuid()) + "_" + std::to_string(envInt("MASTER_PORT", 29500)) + "_" + nonce + ".id";
}
std::string rendezvous(int rank, int world) {
  if (world <= 1) return std::string();
  const std::string path = rendezvousPath();
  std::string id(GVM_DIST_ID_BYTES, '\0');
  if (rank == 0) {
    if (gvm_dist_unique_id(&id[0], id.size()) != 0) {
      std::printf("ERROR: %s\n", gvm_last_error());
      std::exit(-1);
    }
    const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    unlink(tmp.c_str());
    const int fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_NOFOLLOW, 0600);
    if (fd < 0) {
      std::printf("ERROR: cannot create the rendezvous file %s (%s)\n", tmp.c_str(), std::strerror(errno));
      std::exit(-1);
    }
    const ssize_t put = write(fd, id.data(), id.size());
    if (close(fd) != 0 || put != (ssize_t)id.size() || std::rename(tmp.c_str(), path.c_str()) != 0) {
      std::printf("ERROR: cannot publish the NCCL id in %s (%s)\n", path.c_str(), std::strerror(errno));
      unlink(tmp.c_str());
      std::exit(-1);
    }
  } else {
    for (int tries = 0; tries < 600; tries++) {
      std::FILE* fp = std::fopen(path.c_str(), "rb");
      if (fp) {
        const size_t got = std::fread(&id[0], 1, id.size(), fp);
        std::fclose(fp);
        if (got == id.size()) return id;
      }
      usleep(100000);
    }
    std::printf("ERROR: rank %d timed out waiting for %s\n", rank, path.c_str());
    std::exit(-1);
  }
  return id;
}
// ncclCommInitRank is collective: when it has returned on rank 0 every rank has read the id
void rendezvousDone(int rank, int world) {
  if (world > 1 && rank == 0) unlink(rendezvousPath().c_str());
}
}  // namespace

int main(int argc, char** argv) {
  const int rank = envInt("RANK", 0), world = envInt("WORLD_SIZE", 1);
  // engine-side choices that the reference makes by editing main.cu: optimizer, weighting
  // scheme and gridding kernel, selectable here through the environment
  const char* opt_name = std::getenv("GVM_OPTIMIZER") ? std::getenv("GVM_OPTIMIZER") : "CG-FRPRMN";
  const char* scheme_name = std::getenv("GVM_WEIGHTING") ? std::getenv("GVM_WEIGHTING") : "Natural";
  const char* ck_name = std::getenv("GVM_CKERNEL") ? std::getenv("GVM_CKERNEL") : "PillBox2D";

  Synthesizer* sy = createObject<Synthesizer, std::string>("MFS");
  Optimizer* cg = createObject<Optimizer, std::string>(opt_name);
  if (std::getenv("GVM_LBFGS_K")) cg->setK(envInt("GVM_LBFGS_K", 100));
  CKernel* sc = createObject<CKernel, std::string>(ck_name);
  if (std::getenv("GVM_CKERNEL_SIZE")) sc->setmn(envInt("GVM_CKERNEL_SIZE", 7), envInt("GVM_CKERNEL_SIZE", 7));
  ObjectiveFunction* of = createObject<ObjectiveFunction, std::string>("ObjectiveFunction");
  Io* ioms = createObject<Io, std::string>("IoMS");
  Io* iofits = createObject<Io, std::string>("IoFITS");
  WeightingScheme* scheme = createObject<WeightingScheme, std::string>(scheme_name);

  static_cast<MFS*>(sy)->setDistributed(rank, world, rendezvous(rank, world));
  sy->setIoVisibilitiesHandler(ioms);
  sy->setIoImageHandler(iofits);
  sy->setOrder(&optimizationOrder);
  sy->setWeightingScheme(scheme);
  sy->setGriddingKernel(sc);
  sy->setOptimizator(cg);
  sy->configure(argc, argv);
  cg->setObjectiveFunction(of);
  sy->setDevice();
  rendezvousDone(rank, world);

  Fi* chi2 = createObject<Fi, std::string>("Chi2");
  Fi* e = createObject<Fi, std::string>("Entropy");
  Fi* l1 = createObject<Fi, std::string>("L1-Norm");
  Fi* tsqv = createObject<Fi, std::string>("TotalSquaredVariation");
  Fi* lap = createObject<Fi, std::string>("Laplacian");
  chi2->configure(-1, 0, 0, false);  // (penalizatorIndex, imageIndex, imageToAddDphi, normalize)
  e->configure(0, 0, 0, false);
  e->setPrior(0.001f);
  l1->configure(1, 0, 0, false);
  tsqv->configure(2, 0, 0, false);
  lap->configure(3, 0, 0, false);
  of->addFi(chi2);
  of->addFi(e);
  of->addFi(l1);
  of->addFi(tsqv);
  of->addFi(lap);

  sy->run();
  sy->writeImages();
  sy->writeResiduals();
  sy->unSetDevice();
  return 0;
}
