// sort.cu — hand-written device primitives of the upload / preprocessing paths: a stable LSD radix sort of
// (uint32 key, uint32 value) pairs and an exclusive prefix sum. No library calls (north_star: the hot path's
// kernels are written for sm_100a; cuFFT is the one library exception).
//
// Radix sort, 8 bits per pass, three launches per pass:
//   k_sort_hist     every block counts the digits of its chunk (kChunk keys)      -> hist[digit][block]
//   scan            exclusive prefix sum over hist in (digit, block) order         -> global start of every (digit, block)
//   k_sort_scatter  every block ranks its keys STABLY and writes them to their final place
// Stability inside a block: warp w owns the contiguous sub-chunk [w * kPerWarp, (w+1) * kPerWarp) and walks it
// in rounds of 32 consecutive keys; inside a round __match_any_sync groups the lanes with equal digits and the
// rank is the number of lower lanes in the group; a per-warp running count per digit carries the rank across
// rounds, and a prefix over the warps of the block (one thread per digit) carries it across warps. Equal keys
// therefore keep their input order — the gridding path needs that (ascending sample index inside a tile is
// the reference's summation order, src/functions.cu:1418-1508) and the forward degridder gets a reproducible
// order for its chi2 sum.
#include "gvm_internal.cuh"

namespace {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kRounds = 16;                          // rounds of 32 keys per warp
constexpr int kPerWarp = 32 * kRounds;               // 512 keys per warp
constexpr int kChunk = kPerWarp * kSortWarps;        // 4096 keys per block

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

__global__ void __launch_bounds__(kSortThreads) k_sort_hist(const uint32_t* __restrict__ keys, size_t n, int shift,
                                                            uint32_t* __restrict__ hist, unsigned nblocks) {
  __shared__ uint32_t s_hist[256];
  s_hist[threadIdx.x] = 0u;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * kChunk;
  for (int r = 0; r < kChunk / kSortThreads; r++) {
    const size_t k = base + (size_t)r * kSortThreads + threadIdx.x;
    if (k < n) atomicAdd(&s_hist[(keys[k] >> shift) & 0xFFu], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = s_hist[threadIdx.x];
}

__global__ void __launch_bounds__(kSortThreads) k_sort_scatter(const uint32_t* __restrict__ keys_in,
                                                               const uint32_t* __restrict__ vals_in,
                                                               uint32_t* __restrict__ keys_out,
                                                               uint32_t* __restrict__ vals_out, size_t n, int shift,
                                                               const uint32_t* __restrict__ hist_scanned,
                                                               unsigned nblocks) {
  __shared__ uint32_t s_warp[kSortWarps][256];   // pass A: digit counts of each warp; then start of the warp inside the block
  __shared__ uint32_t s_run[kSortWarps][256];    // pass B: keys of this digit the warp has already placed
  __shared__ uint32_t s_base[256];               // global position of the block's first key of every digit
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) {
    (&s_warp[0][0])[i] = 0u;
    (&s_run[0][0])[i] = 0u;
  }
  s_base[threadIdx.x] = hist_scanned[(size_t)threadIdx.x * nblocks + blockIdx.x];
  __syncthreads();
  const size_t wbase = (size_t)blockIdx.x * kChunk + (size_t)warp * kPerWarp;
  uint32_t key[kRounds];
  // pass A: per-warp digit counts
#pragma unroll
  for (int r = 0; r < kRounds; r++) {
    const size_t k = wbase + (size_t)r * 32 + lane;
    const bool in = k < n;
    key[r] = in ? keys_in[k] : 0xFFFFFFFFu;
    const unsigned active = __ballot_sync(0xffffffffu, in);
    if (in) {
      const uint32_t d = (key[r] >> shift) & 0xFFu;
      const unsigned peers = __match_any_sync(active, d);
      if ((peers & lanemask_lt()) == 0u) s_warp[warp][d] += (uint32_t)__popc(peers);
    }
    __syncwarp();
  }
  __syncthreads();
  {  // one thread per digit: exclusive prefix over the warps
    const int d = threadIdx.x;
    uint32_t run = 0u;
#pragma unroll
    for (int w = 0; w < kSortWarps; w++) {
      const uint32_t c = s_warp[w][d];
      s_warp[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
  // pass B: final positions
#pragma unroll
  for (int r = 0; r < kRounds; r++) {
    const size_t k = wbase + (size_t)r * 32 + lane;
    const bool in = k < n;
    const unsigned active = __ballot_sync(0xffffffffu, in);
    if (in) {
      const uint32_t d = (key[r] >> shift) & 0xFFu;
      const unsigned peers = __match_any_sync(active, d);
      const uint32_t rank = (uint32_t)__popc(peers & lanemask_lt());
      const uint32_t pos = s_base[d] + s_warp[warp][d] + s_run[warp][d] + rank;
      keys_out[pos] = key[r];
      vals_out[pos] = vals_in[k];
      __syncwarp(active);
      if (rank == 0u) s_run[warp][d] += (uint32_t)__popc(peers);
    }
    __syncwarp();
  }
}

// ---- exclusive prefix sum (uint32), two-level, recursive over the block totals
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanChunk = kScanThreads * kScanItems;   // 2048 per block

__global__ void __launch_bounds__(kScanThreads) k_scan_block(uint32_t* __restrict__ data, size_t n,
                                                             uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t s_warp[kScanThreads / 32];
  const size_t base = (size_t)blockIdx.x * kScanChunk + (size_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t sum = 0u;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    v[i] = base + i < n ? data[base + i] : 0u;
    sum += v[i];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < kScanThreads / 32 ? s_warp[lane] : 0u;
#pragma unroll
    for (int o = 1; o < kScanThreads / 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    if (lane < kScanThreads / 32) s_warp[lane] = w;   // inclusive over warps
  }
  __syncthreads();
  uint32_t excl = incl - sum + (warp > 0 ? s_warp[warp - 1] : 0u);
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    if (base + i < n) data[base + i] = excl;
    excl += v[i];
  }
  if (threadIdx.x == kScanThreads - 1 && block_sums) block_sums[blockIdx.x] = excl;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_add(uint32_t* __restrict__ data, size_t n,
                                                           const uint32_t* __restrict__ block_offsets) {
  const uint32_t off = block_offsets[blockIdx.x];
  const size_t base = (size_t)blockIdx.x * kScanChunk;
  for (int i = threadIdx.x; i < kScanChunk; i += kScanThreads)
    if (base + i < n) data[base + i] += off;
}

size_t scan_temp_words(size_t n) {
  size_t words = 0;
  while (n > (size_t)kScanChunk) {
    n = (n + kScanChunk - 1) / kScanChunk;
    words += n;
  }
  return words + 1;
}

}  // namespace

size_t gvm_scan_temp_bytes(size_t n) { return scan_temp_words(n) * sizeof(uint32_t); }

// In-place exclusive prefix sum of n uint32 values (totals must stay below 2^32).
int gvm_exclusive_scan_u32(uint32_t* data, size_t n, void* temp, cudaStream_t stream) {
  if (n == 0) return 0;
  const size_t blocks = (n + kScanChunk - 1) / kScanChunk;
  uint32_t* sums = static_cast<uint32_t*>(temp);
  k_scan_block<<<(unsigned)blocks, kScanThreads, 0, stream>>>(data, n, blocks > 1 ? sums : nullptr);
  GVM_CUDA(cudaGetLastError());
  if (blocks > 1) {
    if (gvm_exclusive_scan_u32(sums, blocks, sums + blocks, stream)) return 1;
    k_scan_add<<<(unsigned)blocks, kScanThreads, 0, stream>>>(data, n, sums);
    GVM_CUDA(cudaGetLastError());
  }
  return 0;
}

size_t gvm_sort_temp_bytes(size_t n) {
  const size_t nblocks = (n + kChunk - 1) / kChunk;
  const size_t hist = 256 * (nblocks ? nblocks : 1);
  return (2 * (n ? n : 1) + hist) * sizeof(uint32_t) + gvm_scan_temp_bytes(hist) + 64;
}

// Stable sort of n (key, value) pairs by the low `key_bits` bits of the key; the result is left in
// keys / vals. temp: gvm_sort_temp_bytes(n) bytes of device memory.
int gvm_sort_pairs_u32(uint32_t* keys, uint32_t* vals, size_t n, int key_bits, void* temp, cudaStream_t stream) {
  if (n == 0) return 0;
  if (n >= ((size_t)1 << 32)) { gvm_set_error("gvm_sort_pairs_u32: %zu pairs exceed 32-bit positions", n); return 1; }
  const unsigned nblocks = (unsigned)((n + kChunk - 1) / kChunk);
  uint32_t* k2 = static_cast<uint32_t*>(temp);
  uint32_t* v2 = k2 + n;
  uint32_t* hist = v2 + n;
  void* scan_tmp = hist + 256 * (size_t)nblocks;
  uint32_t *ki = keys, *vi = vals, *ko = k2, *vo = v2;
  const int passes = (key_bits + 7) / 8;
  for (int p = 0; p < passes; p++) {
    k_sort_hist<<<nblocks, kSortThreads, 0, stream>>>(ki, n, 8 * p, hist, nblocks);
    GVM_CUDA(cudaGetLastError());
    if (gvm_exclusive_scan_u32(hist, 256 * (size_t)nblocks, scan_tmp, stream)) return 1;
    k_sort_scatter<<<nblocks, kSortThreads, 0, stream>>>(ki, vi, ko, vo, n, 8 * p, hist, nblocks);
    GVM_CUDA(cudaGetLastError());
    uint32_t* t = ki; ki = ko; ko = t;
    t = vi; vi = vo; vo = t;
  }
  if (ki != keys) {   // odd number of passes: bring the result home
    GVM_CUDA(cudaMemcpyAsync(keys, ki, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
    GVM_CUDA(cudaMemcpyAsync(vals, vi, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
  }
  return 0;
}

// Test entry (include/gvm_b200.h): the sort on host arrays, in place.
extern "C" int gvm_sort_pairs_host(int device, uint32_t* keys, uint32_t* vals, int64_t n, int key_bits) {
  if (n < 0 || key_bits < 1 || key_bits > 32 || !keys || !vals) { gvm_set_error("gvm_sort_pairs_host: bad argument"); return 1; }
  if (n == 0) return 0;
  GVM_CUDA(cudaSetDevice(device));
  uint32_t *dk = nullptr, *dv = nullptr;
  void* tmp = nullptr;
  int rc = 1;
  do {
    if (cudaMalloc(&dk, n * sizeof(uint32_t)) != cudaSuccess || cudaMalloc(&dv, n * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(&tmp, gvm_sort_temp_bytes((size_t)n)) != cudaSuccess) { gvm_set_error("gvm_sort_pairs_host: out of device memory"); break; }
    if (cudaMemcpy(dk, keys, n * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) break;
    if (cudaMemcpy(dv, vals, n * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) break;
    if (gvm_sort_pairs_u32(dk, dv, (size_t)n, key_bits, tmp, nullptr)) break;
    if (cudaMemcpy(keys, dk, n * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) break;
    if (cudaMemcpy(vals, dv, n * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) break;
    rc = 0;
  } while (0);
  if (rc && cudaPeekAtLastError() != cudaSuccess) gvm_set_error("gvm_sort_pairs_host: %s", cudaGetErrorString(cudaGetLastError()));
  cudaFree(dk); cudaFree(dv); cudaFree(tmp);
  return rc;
}
