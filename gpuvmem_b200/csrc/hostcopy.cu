// hostcopy.cu — large copies between PAGEABLE host memory and the device.
//
// cudaMemcpy from pageable memory runs at ~4-5 GB/s on this platform (the driver stages through one
// small pinned buffer on the calling thread), which made the once-per-run uploads of the raw samples
// (36 B/sample, twice: weighting and gridding) the largest part of the preprocessing time. Here the
// copy is a pipeline over a ring of pinned 32 MB buffers: worker threads fill a buffer from the
// caller's memory in parallel while the DMA engine drains the previous ones at PCIe speed.
// Small copies (< 8 MB) go through plain cudaMemcpy.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "gvm_internal.cuh"

namespace {
constexpr size_t kChunk = (size_t)32 << 20;
constexpr int kRing = 4;
constexpr size_t kSmall = (size_t)8 << 20;

struct Ring {
  void* buf[kRing] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t done[kRing];
  bool ok = false;
  int device = -1;
  ~Ring() {
    if (!ok) return;
    for (int i = 0; i < kRing; i++) { cudaFreeHost(buf[i]); cudaEventDestroy(done[i]); }
  }
  bool ensure() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (ok && dev == device) return true;
    if (ok) { for (int i = 0; i < kRing; i++) { cudaFreeHost(buf[i]); cudaEventDestroy(done[i]); } ok = false; }
    for (int i = 0; i < kRing; i++) {
      if (cudaMallocHost(&buf[i], kChunk) != cudaSuccess) { cudaGetLastError(); return false; }
      cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming);
    }
    device = dev;
    ok = true;
    return true;
  }
};
thread_local Ring g_ring;

int copy_threads() {
  static const int n = [] {
    const char* s = getenv("GVM_COPY_THREADS");
    int v = s ? atoi(s) : 0;
    if (v <= 0) {
      // measured on a 16-core B200 host, 1.8 GB from pageable memory: 6 threads 73 ms, 10 threads 62 ms, 14 threads 57 ms
      const unsigned hc = std::thread::hardware_concurrency();
      v = hc >= 16 ? 10 : (hc >= 8 ? 4 : 2);
      // one process per GPU: share the cores with the other ranks of this node (torchrun exports LOCAL_WORLD_SIZE)
      const char* lws = getenv("LOCAL_WORLD_SIZE");
      const int ranks = lws ? atoi(lws) : 1;
      if (ranks > 1 && hc > 0) v = std::max(2, std::min(v, (int)hc / ranks));
    }
    return v > 16 ? 16 : v;
  }();
  return n;
}

// dst <- src for n bytes, split over the copy threads (the calling thread takes the first part)
void parallel_memcpy(void* dst, const void* src, size_t n) {
  const int T = copy_threads();
  if (T <= 1 || n < ((size_t)4 << 20)) { std::memcpy(dst, src, n); return; }
  const size_t part = ((n / T) + 4095) & ~(size_t)4095;
  std::vector<std::thread> th;
  for (int t = 1; t < T; t++) {
    const size_t off = (size_t)t * part;
    if (off >= n) break;
    const size_t len = std::min(part, n - off);
    th.emplace_back([=] { std::memcpy((char*)dst + off, (const char*)src + off, len); });
  }
  std::memcpy(dst, src, std::min(part, n));
  for (auto& x : th) x.join();
}
}  // namespace

// Host (pageable or pinned) -> device. Synchronous: returns when the data is on the device.
int gvm_fast_h2d(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return 0;
  cudaPointerAttributes attr;
  const bool pinned = cudaPointerGetAttributes(&attr, src_host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  if (bytes < kSmall || pinned || !g_ring.ensure()) {
    GVM_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, stream));
    GVM_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }
  Ring& r = g_ring;
  size_t off = 0;
  for (long c = 0; off < bytes; c++, off += kChunk) {
    const int slot = (int)(c % kRing);
    const size_t n = std::min(kChunk, bytes - off);
    if (c >= kRing) GVM_CUDA(cudaEventSynchronize(r.done[slot]));     // the DMA that used this buffer has finished
    parallel_memcpy(r.buf[slot], (const char*)src_host + off, n);
    GVM_CUDA(cudaMemcpyAsync((char*)dst_dev + off, r.buf[slot], n, cudaMemcpyHostToDevice, stream));
    GVM_CUDA(cudaEventRecord(r.done[slot], stream));
  }
  GVM_CUDA(cudaStreamSynchronize(stream));
  return 0;
}

// Device -> host (pageable or pinned). Synchronous.
int gvm_fast_d2h(void* dst_host, const void* src_dev, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return 0;
  cudaPointerAttributes attr;
  const bool pinned = cudaPointerGetAttributes(&attr, dst_host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  if (bytes < kSmall || pinned || !g_ring.ensure()) {
    GVM_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, stream));
    GVM_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }
  Ring& r = g_ring;
  const long nchunks = (long)((bytes + kChunk - 1) / kChunk);
  // DMA runs kRing - 1 chunks ahead of the copy out of the pinned ring
  for (long c = 0; c < nchunks + (kRing - 1); c++) {
    if (c < nchunks) {
      const int slot = (int)(c % kRing);
      const size_t off = (size_t)c * kChunk, n = std::min(kChunk, bytes - off);
      GVM_CUDA(cudaMemcpyAsync(r.buf[slot], (const char*)src_dev + off, n, cudaMemcpyDeviceToHost, stream));
      GVM_CUDA(cudaEventRecord(r.done[slot], stream));
    }
    const long d = c - (kRing - 1);
    if (d >= 0) {
      const int slot = (int)(d % kRing);
      const size_t off = (size_t)d * kChunk, n = std::min(kChunk, bytes - off);
      GVM_CUDA(cudaEventSynchronize(r.done[slot]));
      parallel_memcpy((char*)dst_host + off, r.buf[slot], n);
    }
  }
  return 0;
}

// The pinned ring costs ~0.1 s to create (cudaMallocHost); gvm_create does it once so that the first large upload of
// this thread (weighting, gridding, gvm_add_channel) does not pay for it.
void gvm_hostcopy_warm() { g_ring.ensure(); }
