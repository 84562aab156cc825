// dist_nccl.cu — multi-GPU plumbing of the engine: one process per GPU, one NCCL
// communicator over NVLink/NVSwitch, sum all-reduces issued on the engine stream.
//
// What is reduced: the chi2 scalar of every objective
// evaluation (fp64, 1 element) and the image-sized gradient [2][M][N] fp32 of every
// gradient evaluation. The reference instead accumulates per-GPU gradients into GPU 0
// through peer-to-peer loads under an OpenMP critical section
// (src/functions.cu:4534-4549) and sums chi2 on the host (:4437-4446).
//
// NCCL is bound at run time with dlopen("libnccl.so.2") so that a process that already
// carries a NCCL (torch's bundled one) shares it, and single-GPU users need no NCCL.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>

#include "gvm_internal.cuh"

namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { gvm_set_error("gvm_dist: cannot load libnccl.so.2 (%s)", dlerror()); return 1; }
#define GVM_SYM(field, name)                                                        \
  *(void**)(&g_nccl.field) = dlsym(h, name);                                        \
  if (!g_nccl.field) { gvm_set_error("gvm_dist: libnccl lacks %s", name); return 1; }
  GVM_SYM(GetUniqueId, "ncclGetUniqueId")
  GVM_SYM(CommInitRank, "ncclCommInitRank")
  GVM_SYM(AllReduce, "ncclAllReduce")
  GVM_SYM(Broadcast, "ncclBroadcast")
  GVM_SYM(Send, "ncclSend")
  GVM_SYM(Recv, "ncclRecv")
  GVM_SYM(GroupStart, "ncclGroupStart")
  GVM_SYM(GroupEnd, "ncclGroupEnd")
  GVM_SYM(CommDestroy, "ncclCommDestroy")
  GVM_SYM(CommAbort, "ncclCommAbort")
  GVM_SYM(GetErrorString, "ncclGetErrorString")
  GVM_SYM(GetVersion, "ncclGetVersion")
#undef GVM_SYM
  g_nccl.handle = h;
  return 0;
}

#define GVM_NCCL(call)                                                              \
  do {                                                                              \
    ncclResult_t _r = (call);                                                       \
    if (_r != ncclSuccess) {                                                        \
      gvm_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(_r)); \
      return 1;                                                                     \
    }                                                                               \
  } while (0)
}  // namespace

#define GVM_DIST_ALIVE(e) \
  if (!(e)->nccl_comm) { gvm_set_error("gvm_dist: the communicator was aborted after a failure on this rank"); return 1; }

int gvm_dist_allreduce_f32(gvm_engine* e, float* buf, size_t n) {
  if (e->world <= 1 || e->replicated) return 0;
  GVM_DIST_ALIVE(e)
  GVM_NCCL(g_nccl.AllReduce(buf, buf, n, ncclFloat, ncclSum, (ncclComm_t)e->nccl_comm, e->stream));
  e->collectives++;
  return 0;
}
int gvm_dist_allreduce_f64(gvm_engine* e, double* buf, size_t n) {
  if (e->world <= 1 || e->replicated) return 0;
  GVM_DIST_ALIVE(e)
  GVM_NCCL(g_nccl.AllReduce(buf, buf, n, ncclDouble, ncclSum, (ncclComm_t)e->nccl_comm, e->stream));
  e->collectives++;
  return 0;
}
int gvm_dist_broadcast_f32(gvm_engine* e, float* buf, size_t n, int root) {
  if (e->world <= 1) return 0;
  GVM_DIST_ALIVE(e)
  GVM_NCCL(g_nccl.Broadcast(buf, buf, n, ncclFloat, root, (ncclComm_t)e->nccl_comm, e->stream));
  e->collectives++;
  return 0;
}
// point-to-point and byte-wise collectives of the distributed preprocessing (weights_grid.cu)
int gvm_dist_send(gvm_engine* e, const void* buf, size_t bytes, int peer) {
  GVM_DIST_ALIVE(e)
  GVM_NCCL(g_nccl.Send(buf, bytes, ncclInt8, peer, (ncclComm_t)e->nccl_comm, e->stream));
  return 0;
}
int gvm_dist_recv(gvm_engine* e, void* buf, size_t bytes, int peer) {
  GVM_DIST_ALIVE(e)
  GVM_NCCL(g_nccl.Recv(buf, bytes, ncclInt8, peer, (ncclComm_t)e->nccl_comm, e->stream));
  return 0;
}
int gvm_dist_group_begin(gvm_engine* e) {
  GVM_DIST_ALIVE(e)
  GVM_NCCL(g_nccl.GroupStart());
  return 0;
}
int gvm_dist_group_end(gvm_engine* e) {
  GVM_DIST_ALIVE(e)
  GVM_NCCL(g_nccl.GroupEnd());
  e->collectives++;
  return 0;
}
int gvm_dist_broadcast_bytes(gvm_engine* e, void* buf, size_t bytes, int root) {
  if (e->world <= 1) return 0;
  GVM_DIST_ALIVE(e)
  GVM_NCCL(g_nccl.Broadcast(buf, buf, bytes, ncclInt8, root, (ncclComm_t)e->nccl_comm, e->stream));
  e->collectives++;
  return 0;
}
// element-wise max of unsigned words: with disjoint ownership (every word non-zero on at most one rank) this is an
// exact bit-for-bit merge, which a floating-point sum is not (-0.0 + 0.0 = +0.0)
int gvm_dist_allreduce_u32_max(gvm_engine* e, uint32_t* buf, size_t n) {
  if (e->world <= 1) return 0;
  GVM_DIST_ALIVE(e)
  GVM_NCCL(g_nccl.AllReduce(buf, buf, n, ncclUint32, ncclMax, (ncclComm_t)e->nccl_comm, e->stream));
  e->collectives++;
  return 0;
}
void gvm_dist_abort_comm(gvm_engine* e) {
  if (e->world <= 1 || !e->nccl_comm) return;
  if (g_nccl.CommAbort) g_nccl.CommAbort((ncclComm_t)e->nccl_comm);
  e->nccl_comm = nullptr;   // every later collective of this engine fails instead of touching a dead communicator
  e->dist_aborted = true;
}
void gvm_dist_release(gvm_engine* e) {
  if (e->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)e->nccl_comm);
  e->nccl_comm = nullptr;
  e->world = 1;
  e->rank = 0;
}

extern "C" {

int gvm_dist_unique_id(char* id_out, size_t bytes) {
  if (!id_out || bytes < sizeof(ncclUniqueId)) { gvm_set_error("gvm_dist_unique_id: need %zu bytes", sizeof(ncclUniqueId)); return 1; }
  if (load_nccl()) return 1;
  ncclUniqueId id;
  GVM_NCCL(g_nccl.GetUniqueId(&id));
  std::memcpy(id_out, &id, sizeof(id));
  return 0;
}

int gvm_dist_init(gvm_engine* e, int rank, int world, const char* id, size_t bytes) {
  if (!e || world < 1 || rank < 0 || rank >= world) { gvm_set_error("gvm_dist_init: bad rank/world %d/%d", rank, world); return 1; }
  if (world == 1) { e->rank = 0; e->world = 1; return 0; }
  if (!id || bytes < sizeof(ncclUniqueId)) { gvm_set_error("gvm_dist_init: need the %zu-byte id of gvm_dist_unique_id", sizeof(ncclUniqueId)); return 1; }
  if (load_nccl()) return 1;
  GVM_CUDA(cudaSetDevice(e->cfg.device));
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  ncclComm_t comm = nullptr;
  GVM_NCCL(g_nccl.CommInitRank(&comm, world, uid, rank));
  e->nccl_comm = comm;
  e->rank = rank;
  e->world = world;
  // NCCL connects lazily: the first use of every transport (ring / tree of the collectives, every point-to-point
  // pair) costs hundreds of milliseconds. Pay that here, once, with messages large enough to bring up all channels,
  // instead of inside the first weighting / gridding / gradient of the run (GVM_NCCL_WARMUP=0 skips it).
  const char* wu = getenv("GVM_NCCL_WARMUP");
  if (!(wu && *wu == '0')) {
    const size_t n = (size_t)8 << 20;   // 32 MB of floats per message
    float* buf = nullptr;
    GVM_CUDA(cudaMalloc(&buf, (size_t)(world + 1) * n * sizeof(float)));
    GVM_CUDA(cudaMemsetAsync(buf, 0, (size_t)(world + 1) * n * sizeof(float), e->stream));
    int rc = gvm_dist_allreduce_f32(e, buf, n) || gvm_dist_allreduce_u32_max(e, reinterpret_cast<uint32_t*>(buf), n) ||
             gvm_dist_broadcast_bytes(e, buf, n * sizeof(float), 0) || gvm_dist_group_begin(e);
    for (int peer = 0; peer < world && !rc; peer++)
      rc = gvm_dist_send(e, buf, n * sizeof(float), peer) || gvm_dist_recv(e, buf + (size_t)(peer + 1) * n, n * sizeof(float), peer);
    rc = rc || gvm_dist_group_end(e);
    if (!rc && cudaStreamSynchronize(e->stream) != cudaSuccess) { gvm_set_error("gvm_dist_init: warm-up failed"); rc = 1; }
    cudaFree(buf);
    e->collectives = 0;
    if (rc) return 1;
  }
  return 0;
}

int gvm_dist_rank(gvm_engine* e) { return e->rank; }
int gvm_dist_world(gvm_engine* e) { return e->world; }
int64_t gvm_dist_collectives(gvm_engine* e) { return e->collectives; }

int gvm_dist_set_replicated(gvm_engine* e, int on) {
  e->replicated = on != 0;
  e->epoch++;
  return 0;
}

int gvm_dist_abort(gvm_engine* e) {
  gvm_dist_abort_comm(e);
  return 0;
}

int gvm_dist_broadcast(gvm_engine* e, float* buf_dev, int64_t n, int root) {
  if (root < 0 || root >= e->world) { gvm_set_error("gvm_dist_broadcast: root %d of %d", root, e->world); return 1; }
  return gvm_dist_broadcast_f32(e, buf_dev, (size_t)n, root);
}

/* In-place sum all-reduce of n floats on the engine stream (optimizer-level use). */
int gvm_dist_allreduce(gvm_engine* e, float* buf_dev, int64_t n) {
  return gvm_dist_allreduce_f32(e, buf_dev, (size_t)n);
}

}  // extern "C"
