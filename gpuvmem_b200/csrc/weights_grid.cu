// weights_grid.cu — imaging weights (natural / uniform / Briggs / radial) and convolutional
// gridding on the GPU, bit-exact with the reference's host code run with ONE thread.
//
// Reference: WeightingScheme::apply (src/{natural,uniform,briggs,radial}weightingscheme.cu)
// and do_gridding (src/functions.cu:1339-1653). Both accumulate fp32 sums sample by sample
// (under `omp critical` / `omp atomic`), so the result depends on the order; the only
// deterministic reference order is the single-thread one (ascending sample index). It is
// reproduced exactly: samples are STABLY radix-sorted by grid cell (sort.cu, hand-written) and every cell is summed
// sequentially in ascending sample order by one thread; all fp32 arithmetic uses explicit
// round-to-nearest intrinsics so that nvcc cannot contract what gcc does not (the reference's
// host code is compiled for baseline x86-64: separate multiply and add). Cell indices use the
// reference's mixed fp32/fp64 arithmetic verbatim.
// The two order-dependent SCALARS of Briggs (sum of weights, sum of squared grid weights over
// the half plane; src/briggsweightingscheme.cu:46-110) are sequential fp32 sums over host /
// downloaded data in the reference's loop order. UVTaper (include/classes/uvtaper.cuh:100) is
// evaluated on the host: it calls libm's exp/cosf/sinf, which no device routine matches bit for bit.
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "gvm_internal.cuh"

namespace {

#define WG_CUDA(call)                                                                \
  do {                                                                               \
    cudaError_t _e = (call);                                                         \
    if (_e != cudaSuccess) {                                                         \
      gvm_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

constexpr uint32_t kNoCell = 0xFFFFFFFFu;

// src/uniformweightingscheme.cu:36-49 / src/briggsweightingscheme.cu:75-88
__global__ void __launch_bounds__(256) k_weight_cells(const double* __restrict__ uvw_m, long Z, float freq,
                                                      double adu, double adv, long M, long N,
                                                      uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const long z = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (z >= Z) return;
  double u = gvm_metres_to_lambda(uvw_m[3 * z], freq);
  double v = gvm_metres_to_lambda(uvw_m[3 * z + 1], freq);
  if (u < 0.0) { u *= -1.0; v *= -1.0; }
  const double gx = u / adu, gy = v / adv;
  const int x = (int)(gx + (double)(int)(N / 2) + 0.5);
  const int y = (int)(gy + (double)(int)(M / 2) + 0.5);
  keys[z] = (x >= 0 && y >= 0 && x < N && y < M) ? (uint32_t)(N * y + x) : kNoCell;
  vals[z] = (uint32_t)z;
}

// One thread per segment head of the sorted (cell, sample) list: sequential fp32 sum of the
// cell's samples in ascending sample order, starting from what the grid already holds.
__global__ void __launch_bounds__(256) k_cell_accumulate(const uint32_t* __restrict__ keys,
                                                         const uint32_t* __restrict__ vals, long n,
                                                         const float* __restrict__ w, float* __restrict__ grid) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint32_t c = keys[t];
  if (c == kNoCell || (t > 0 && keys[t - 1] == c)) return;
  float g = grid[c];
  for (long s = t; s < n && keys[s] == c; s++) g = __fadd_rn(g, w[vals[s]]);
  grid[c] = g;
}

// uniform: w /= g (src/uniformweightingscheme.cu:82-86); Briggs: w /= (1.0 + g * f2) in double
// (src/briggsweightingscheme.cu:172-173); off-grid samples get weight 0.
__global__ void __launch_bounds__(256) k_weight_apply(const uint32_t* __restrict__ keys,
                                                      const uint32_t* __restrict__ vals, long n,
                                                      const float* __restrict__ grid, int briggs, float f2,
                                                      float* __restrict__ w) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint32_t c = keys[t], z = vals[t];
  if (c == kNoCell) { w[z] = 0.0f; return; }
  const float g = grid[c];
  if (briggs) w[z] = (float)((double)w[z] / (1.0 + (double)__fmul_rn(g, f2)));
  else w[z] = __fdiv_rn(w[z], g);
}

__global__ void __launch_bounds__(256) k_clear_cells(const uint32_t* __restrict__ keys, long n,
                                                     float* __restrict__ grid) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t < n && keys[t] != kNoCell) grid[keys[t]] = 0.0f;
}

// src/radialweightingscheme.cu: w *= distance((float)u, (float)v, 0, 0), no Hermitian fold
__global__ void __launch_bounds__(256) k_radial(const double* __restrict__ uvw_m, long Z, float freq,
                                                float* __restrict__ w) {
  const long z = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (z >= Z) return;
  const float u = (float)gvm_metres_to_lambda(uvw_m[3 * z], freq);
  const float v = (float)gvm_metres_to_lambda(uvw_m[3 * z + 1], freq);
  const float d = sqrtf(__fadd_rn(__fmul_rn(u, u), __fmul_rn(v, v)));
  w[z] = __fmul_rn(w[z], d);
}

// ------------------------------------------------------------------ gridding
// Centre cell of every Hermitian-doubled sample on the grid EXTENDED by the kernel support
// (src/functions.cu:1432-1461); samples whose taps cannot reach the grid are dropped.
__global__ void __launch_bounds__(256) k_grid_centres(const double* __restrict__ uvw_m, long Z, float freq,
                                                      double deltau, double deltav, long M, long N, int sx,
                                                      int sy, uint32_t* __restrict__ keys,
                                                      uint32_t* __restrict__ vals) {
  const long z = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (z >= 2 * Z) return;
  const long vi = (z < Z) ? z : z - Z;
  double u = uvw_m[3 * vi], v = uvw_m[3 * vi + 1];
  if (z >= Z) { u *= -1.0; v *= -1.0; }
  u = gvm_metres_to_lambda(u, freq);
  v = gvm_metres_to_lambda(v, freq);
  const double gx = u / deltau, gy = v / deltav;
  const double j_fp = gx + floor(N / 2.0) + 0.5, k_fp = gy + floor(M / 2.0) + 0.5;
  const int j = (int)j_fp, k = (int)k_fp;
  const long EW = N + 2L * sx;
  const bool ok = (j >= -sx && j < N + sx && k >= -sy && k < M + sy);
  keys[z] = ok ? (uint32_t)((long)(k + sy) * EW + (j + sx)) : kNoCell;
  vals[z] = (uint32_t)z;
}

__device__ __forceinline__ long lower_bound_u32(const uint32_t* __restrict__ a, long n, uint32_t key) {
  long lo = 0, hi = n;
  while (lo < hi) {
    const long mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

constexpr int kMaxTaps = 17 * 17;
constexpr int kTapBits = 9;               // tap index (< 512) in the low bits of the per-list meta word

// One thread per output cell: merge the (<= taps) sorted sample lists of the centre cells whose
// kernel footprint covers this cell, in ascending sample order, and accumulate exactly like the
// reference's sequential loop (src/functions.cu:1466-1505), then normalise (:1537-1558).
// occupancy of the extended grid: count[key]++ for every (sample, twin) centre; an exclusive scan
// turns it into start[key] = first position of that centre cell in the sorted list
__global__ void __launch_bounds__(256) k_cell_count(const uint32_t* __restrict__ keys, long n,
                                                    int* __restrict__ count) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n && keys[i] != kNoCell) atomicAdd(count + keys[i], 1);
}

// Per-thread merge state lives in SHARED memory, one column per thread (conflict-free): the head
// sample index of every non-empty list (hz), its position (head) and remaining-count | tap index
// (meta). With thread-local arrays the same state spilled to DRAM: 322 GB of traffic for ~1 GB of
// algorithmic bytes (ncu, profiles/r1c_k_grid_accumulate_*): the kernel was HBM-bound on its own spills.
template <int kThreads>
__global__ void __launch_bounds__(kThreads) k_grid_accumulate(
    const int* __restrict__ start, const uint32_t* __restrict__ vals, long n, long Z,
    const float2* __restrict__ Vo, const float* __restrict__ w, const float* __restrict__ kernel, int ck_m,
    int ck_n, int sx, int sy, long M, long N, int taps, float* __restrict__ out_w, float2* __restrict__ out_V) {
  extern __shared__ uint32_t s_state[];
  uint32_t* const s_hz = s_state + threadIdx.x;                          // [taps][kThreads]
  uint32_t* const s_head = s_state + (size_t)taps * kThreads + threadIdx.x;
  uint32_t* const s_meta = s_state + 2 * (size_t)taps * kThreads + threadIdx.x;
  const long cell = blockIdx.x * (long)kThreads + threadIdx.x;
  if (cell >= M * N) return;
  const int gk = (int)(cell / N), gj = (int)(cell % N);
  const long EW = N + 2L * sx;
  int nl = 0;
  for (int m = -sy; m <= sy; m++) {
    // the centres of one footprint row are consecutive keys: one look-up tells whether the row is empty
    // (most of the uv plane is, so most cells leave after 2 * (2 sy + 1) loads)
    const long row0 = (long)(gk - m + sy) * EW + gj;   // nn = +sx ... -sx  <->  ej = gj ... gj + 2 sx
    if (start[row0 + 2L * sx + 1] == start[row0]) continue;
    for (int nn = -sx; nn <= sx; nn++) {
      const int ki = m + sy, kj = nn + sx;
      if (ki < 0 || ki >= ck_m || kj < 0 || kj >= ck_n) continue;
      // centre (k, j) with k + m == gk, j + nn == gj
      const long ek = (long)(gk - m + sy), ej = (long)(gj - nn + sx);
      const uint32_t key = (uint32_t)(ek * EW + ej);
      const int b = start[key], e = start[key + 1];
      if (e > b) {
        s_hz[nl * kThreads] = vals[b];
        s_head[nl * kThreads] = (uint32_t)b;
        s_meta[nl * kThreads] = ((uint32_t)(e - b) << kTapBits) | (uint32_t)(ck_n * ki + kj);
        nl++;
      }
    }
  }
  float gw = 0.f, gw2 = 0.f, gvr = 0.f, gvi = 0.f;
  while (true) {
    uint32_t best = kNoCell; int bl = -1;
    for (int l = 0; l < nl; l++) {
      const uint32_t z = s_hz[l * kThreads];
      if (z < best) { best = z; bl = l; }
    }
    if (bl < 0) break;
    const uint32_t meta = s_meta[bl * kThreads] - (1u << kTapBits);   // one sample fewer in this list
    const uint32_t nh = s_head[bl * kThreads] + 1u;
    s_meta[bl * kThreads] = meta;
    s_head[bl * kThreads] = nh;
    s_hz[bl * kThreads] = (meta >> kTapBits) ? vals[nh] : kNoCell;
    const long vi = (best < (uint32_t)Z) ? (long)best : (long)best - Z;
    const float wt = w[vi];
    float2 vo = Vo[vi];
    if (best >= (uint32_t)Z) vo.y *= -1.0f;
    const float ck = __ldg(&kernel[meta & ((1u << kTapBits) - 1u)]);
    const float ck2 = __fmul_rn(ck, ck);
    gw = __fadd_rn(gw, __fmul_rn(wt, ck));
    gw2 = __fadd_rn(gw2, __fmul_rn(wt, ck2));
    gvr = __fadd_rn(gvr, __fmul_rn(__fmul_rn(wt, vo.x), ck));
    gvi = __fadd_rn(gvi, __fmul_rn(__fmul_rn(wt, vo.y), ck));
  }
  float weight = 0.f, orr = 0.f, oi = 0.f;
  if (gw2 != 0.0f && gw != 0.0f) {
    weight = __fdiv_rn(__fmul_rn(gw, gw), gw2);
    orr = __fdiv_rn(gvr, gw);
    oi = __fdiv_rn(gvi, gw);
  }
  out_w[cell] = weight;
  out_V[cell] = make_float2(orr, oi);
}

// ---------------------------------------------------------------------------------------------
// Tile-sequential gridding (default path). The uv grid is cut into kTile x kTile output tiles; every
// Hermitian-doubled sample is listed under each tile its kernel footprint touches (1, 2 or 4 tiles), the
// (tile, sample) pairs are stably radix-sorted by tile, and ONE WARP per tile replays its samples in
// ascending sample order — the reference's loop order (src/functions.cu:1466-1505) — with the lanes
// spread over the kernel taps and the four accumulators of the tile's cells in shared memory. A tap of
// one sample touches a cell at most once, so lanes never collide inside a sample; __syncwarp() orders
// consecutive samples. Per cell the sequence of fp32 operations is exactly the reference's, so the result
// is bit-identical to the single-thread CPU code (and to k_grid_accumulate, kept as a cross-check).
// The per-pair record (centre, weight, visibility) is gathered into sorted order first, so the replay
// streams 16 B per pair with one coalesced load per 32 samples.
constexpr int kTile = 16;

__device__ __forceinline__ bool grid_centre(const double* __restrict__ uvw_m, long z, long Z, float freq,
                                            double deltau, double deltav, long M, long N, int sx, int sy, int* j,
                                            int* k) {
  const long vi = (z < Z) ? z : z - Z;
  double u = uvw_m[3 * vi], v = uvw_m[3 * vi + 1];
  if (z >= Z) { u *= -1.0; v *= -1.0; }
  u = gvm_metres_to_lambda(u, freq);
  v = gvm_metres_to_lambda(v, freq);
  const double gx = u / deltau, gy = v / deltav;
  const double j_fp = gx + floor(N / 2.0) + 0.5, k_fp = gy + floor(M / 2.0) + 0.5;
  *j = (int)j_fp;
  *k = (int)k_fp;
  return *j >= -sx && *j < N + sx && *k >= -sy && *k < M + sy;   // some tap reaches the grid
}

// pass 1: centre of every doubled sample (packed, offset by the support) and the number of tiles it touches
__global__ void __launch_bounds__(256) k_tile_count(const double* __restrict__ uvw_m, long Z, float freq,
                                                    double deltau, double deltav, long M, long N, int sx, int sy,
                                                    uint32_t* __restrict__ cpos, int* __restrict__ cnt) {
  const long z = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (z >= 2 * Z) return;
  int j, k;
  if (!grid_centre(uvw_m, z, Z, freq, deltau, deltav, M, N, sx, sy, &j, &k)) { cpos[z] = kNoCell; cnt[z] = 0; return; }
  cpos[z] = ((uint32_t)(k + sy) << 16) | (uint32_t)(j + sx);
  const int tx0 = max(j - sx, 0) / kTile, tx1 = min(j + sx, (int)N - 1) / kTile;
  const int ty0 = max(k - sy, 0) / kTile, ty1 = min(k + sy, (int)M - 1) / kTile;
  cnt[z] = (tx1 - tx0 + 1) * (ty1 - ty0 + 1);
}

// pass 2: the (tile, sample) pairs at the offsets of the exclusive scan
__global__ void __launch_bounds__(256) k_tile_emit(const uint32_t* __restrict__ cpos, const int* __restrict__ off,
                                                   long n2, long M, long N, int sx, int sy, int ntx,
                                                   uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const long z = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (z >= n2) return;
  const uint32_t cp = cpos[z];
  if (cp == kNoCell) return;
  const int j = (int)(cp & 0xFFFFu) - sx, k = (int)(cp >> 16) - sy;
  const int tx0 = max(j - sx, 0) / kTile, tx1 = min(j + sx, (int)N - 1) / kTile;
  const int ty0 = max(k - sy, 0) / kTile, ty1 = min(k + sy, (int)M - 1) / kTile;
  int o = off[z];
  for (int ty = ty0; ty <= ty1; ty++)
    for (int tx = tx0; tx <= tx1; tx++) {
      keys[o] = (uint32_t)(ty * ntx + tx);
      vals[o] = (uint32_t)z;
      o++;
    }
}

// tiles sorted by decreasing sample count: the long replays start first (longest-processing-time order)
__global__ void __launch_bounds__(256) k_tile_order_keys(const int* __restrict__ tstart, const int* __restrict__ tend,
                                                         long ntiles, uint32_t* __restrict__ keys,
                                                         uint32_t* __restrict__ vals) {
  const long t = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  keys[t] = 0x7FFFFFFFu - (uint32_t)(tend[t] - tstart[t]);
  vals[t] = (uint32_t)t;
}

// Tile-sequential replay with one THREAD PER CELL (256 threads per 16 x 16 tile): every thread walks the tile's records
// (sorted by tile, ascending sample index inside a tile) and adds the tap that lands on its cell, accumulators in
// registers. Per cell the fp32 operation sequence is exactly the reference's loop (src/functions.cu:1418-1508: ascending
// sample index, the same products), so the result is bit-identical to it. Round 1 ran one WARP per tile with the lanes
// over the taps and the accumulators in shared memory (bit-identical as well): a dense tile (44 k samples against a mean
// of 755 at C5) was serialised on that warp and lanes idled on taps outside the tile; 38 -> 28 ms at C5 / 4.
__global__ void __launch_bounds__(kTile * kTile) k_grid_tiles_cells(
    const uint32_t* __restrict__ order, const int* __restrict__ tstart, const int* __restrict__ tend,
    const float4* __restrict__ rec, const float* __restrict__ kernel, int ck_m, int ck_n, int sx, int sy, long M, long N,
    int ntx, float* __restrict__ out_w, float2* __restrict__ out_V) {
  __shared__ float4 s_rec[kTile * kTile];
  __shared__ float s_ck[kMaxTaps];
  const int tid = threadIdx.x;
  const int tile = (int)order[blockIdx.x];
  const int b = tstart[tile], e = tend[tile];
  const int ty = tile / ntx, tx = tile - ty * ntx;
  const int k0 = ty * kTile, j0 = tx * kTile;
  const int ck_ = tid / kTile, cj = tid - ck_ * kTile;
  const bool on_grid = (long)(j0 + cj) < N && (long)(k0 + ck_) < M;
  const long cell = (long)(k0 + ck_) * N + (j0 + cj);
  if (b >= e) {                       // empty tile: the grids are not cleared beforehand
    if (on_grid) { out_w[cell] = 0.f; out_V[cell] = make_float2(0.f, 0.f); }
    return;
  }
  // the taps that exist: rows ki < min(2 sy + 1, ck_m), columns kj < min(2 sx + 1, ck_n) of the table (<= kMaxTaps)
  const unsigned tw = (unsigned)min(2 * sx + 1, ck_n), th = (unsigned)min(2 * sy + 1, ck_m);
  for (int t = tid; t < (int)(tw * th); t += kTile * kTile) s_ck[t] = kernel[ck_n * (t / (int)tw) + t % (int)tw];
  // tap (ki, kj) of a record centred at (lk, lj) lands on this cell when ki = ck_ - lk + sy, kj = cj - lj + sx
  const int cjb = cj + j0 + 2 * sx, ckb = ck_ + k0 + 2 * sy;
  float ax = 0.f, ay = 0.f, az = 0.f, aw = 0.f;
  for (int base = b; base < e; base += kTile * kTile) {
    const int n = min(kTile * kTile, e - base);
    __syncthreads();
    if (tid < n) s_rec[tid] = __ldg(&rec[base + tid]);
    __syncthreads();
    if (!on_grid) continue;
#pragma unroll 4
    for (int i = 0; i < n; i++) {
      const float4 r = s_rec[i];
      const uint32_t cp = __float_as_uint(r.x);
      const unsigned kj = (unsigned)(cjb - (int)(cp & 0xFFFFu)), ki = (unsigned)(ckb - (int)(cp >> 16));
      if (kj < tw && ki < th) {
        const float ckv = s_ck[tw * ki + kj];
        const float wt = r.y;
        ax = __fadd_rn(ax, __fmul_rn(wt, ckv));
        ay = __fadd_rn(ay, __fmul_rn(wt, __fmul_rn(ckv, ckv)));
        az = __fadd_rn(az, __fmul_rn(__fmul_rn(wt, r.z), ckv));
        aw = __fadd_rn(aw, __fmul_rn(__fmul_rn(wt, r.w), ckv));
      }
    }
  }
  if (!on_grid) return;
  // normalise (src/functions.cu:1537-1558) and write the cell
  float weight = 0.f, orr = 0.f, oi = 0.f;
  if (ay != 0.0f && ax != 0.0f) {
    weight = __fdiv_rn(__fmul_rn(ax, ax), ay);
    orr = __fdiv_rn(az, ax);
    oi = __fdiv_rn(aw, ax);
  }
  out_w[cell] = weight;
  out_V[cell] = make_float2(orr, oi);
}

// ---------------------------------------------------------------------------------------------
// Distributed preprocessing (gvm_weights_dist / gvm_grid_block_dist): kernels of the exchange.
// destination of every (tile, sample) pair: the tile's owner rank, originals before Hermitian twins
__device__ __forceinline__ int tile_owner(uint32_t tile, int ntx, int world) {
  const int ty = (int)(tile / (uint32_t)ntx), tx = (int)(tile - (uint32_t)ty * (uint32_t)ntx);
  return (tx + ty) % world;   // neighbouring tiles (the dense centre of the uv plane) go to different ranks
}
__global__ void __launch_bounds__(256) k_pair_dest(const uint32_t* __restrict__ tile, const uint32_t* __restrict__ zl,
                                                   long npairs, long nloc, int ntx, int world,
                                                   uint32_t* __restrict__ dkey, uint32_t* __restrict__ didx) {
  const long i = blockIdx.x * 256L + threadIdx.x;
  if (i >= npairs) return;
  const int half = zl[i] >= (uint32_t)nloc;
  dkey[i] = (uint32_t)(half * world + tile_owner(tile[i], ntx, world));
  didx[i] = (uint32_t)i;
}
// first pair of every destination in the destination-sorted list (entries of absent destinations stay at npairs)
__global__ void __launch_bounds__(256) k_dest_starts(const uint32_t* __restrict__ dkey, long npairs,
                                                     uint32_t* __restrict__ start) {
  const long i = blockIdx.x * 256L + threadIdx.x;
  if (i >= npairs) return;
  if (i == 0 || dkey[i - 1] != dkey[i]) start[dkey[i]] = (uint32_t)i;
}
// send buffers in destination order: tile id and the 16-byte record (centre, w, Vo) of every pair
__global__ void __launch_bounds__(256) k_pair_pack(const uint32_t* __restrict__ didx, const uint32_t* __restrict__ tile,
                                                   const uint32_t* __restrict__ zl, long npairs, long nloc,
                                                   const uint32_t* __restrict__ cpos, const float2* __restrict__ Vo,
                                                   const float* __restrict__ w, uint32_t* __restrict__ skey,
                                                   float4* __restrict__ srec) {
  const long j = blockIdx.x * 256L + threadIdx.x;
  if (j >= npairs) return;
  const uint32_t i = didx[j], z = zl[i];
  const long vi = z < (uint32_t)nloc ? (long)z : (long)z - nloc;
  float2 vo = Vo[vi];
  if (z >= (uint32_t)nloc) vo.y *= -1.0f;
  skey[j] = tile[i];
  srec[j] = make_float4(__uint_as_float(cpos[z]), w[vi], vo.x, vo.y);
}
__global__ void __launch_bounds__(256) k_iota(uint32_t* __restrict__ v, long n) {
  const long i = blockIdx.x * 256L + threadIdx.x;
  if (i < n) v[i] = (uint32_t)i;
}
// after the stable sort of the RECEIVED pairs by tile: records in replay order + first / last pair of every tile
__global__ void __launch_bounds__(256) k_recv_gather(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ idx,
                                                     long npairs, const float4* __restrict__ rrec,
                                                     int* __restrict__ tstart, int* __restrict__ tend,
                                                     float4* __restrict__ rec) {
  const long i = blockIdx.x * 256L + threadIdx.x;
  if (i >= npairs) return;
  const uint32_t t = keys[i];
  if (i == 0 || keys[i - 1] != t) tstart[t] = (int)i;
  if (i == npairs - 1 || keys[i + 1] != t) tend[t] = (int)(i + 1);
  rec[i] = rrec[idx[i]];
}

__global__ void __launch_bounds__(256) k_grid_flags(const float* __restrict__ wgt, long MN,
                                                    int* __restrict__ flags) {
  const long c = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (c < MN) flags[c] = wgt[c] > 0.0f ? 1 : 0;
}

// Row-major compaction (src/functions.cu:1591-1612) with the cell-centre coordinates in metres
// (:1524-1532): u = (j - floor(N/2)) * deltau * lambda.
__global__ void __launch_bounds__(256) k_grid_compact(const float* __restrict__ wgt,
                                                      const float2* __restrict__ V,
                                                      const int* __restrict__ pos, long M, long N,
                                                      double deltau, double deltav, float lambda,
                                                      double* __restrict__ uvw_out, float2* __restrict__ Vo_out,
                                                      float* __restrict__ w_out) {
  const long c = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (c >= M * N) return;
  const float weight = wgt[c];
  if (!(weight > 0.0f)) return;
  const long gk = c / N, gj = c % N;
  const int o = pos[c];
  const double ul = __dmul_rn((double)gj - floor(N / 2.0), deltau);
  const double vl = __dmul_rn((double)gk - floor(M / 2.0), deltav);
  uvw_out[3 * o] = __dmul_rn(ul, (double)lambda);
  uvw_out[3 * o + 1] = __dmul_rn(vl, (double)lambda);
  uvw_out[3 * o + 2] = 0.0;
  Vo_out[o] = V[c];
  w_out[o] = weight;
}

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  ~DevBuf() { cudaFree(p); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { cudaFree(p); p = o.p; bytes = o.bytes; o.p = nullptr; o.bytes = 0; }
    return *this;
  }
  int ensure(size_t n) {
    if (n <= bytes) return 0;
    cudaFree(p); p = nullptr; bytes = 0;
    if (cudaMalloc(&p, n) != cudaSuccess) { gvm_set_error("weights_grid: cudaMalloc(%zu) failed", n); return 1; }
    bytes = n;
    return 0;
  }
  template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

// One cudaMalloc instead of a dozen: cudaMalloc / cudaFree cost 10-45 ms EACH on this platform (more than most of
// the kernels here), so the work buffers of a call are carved out of a per-thread arena that only ever grows and is
// kept between calls (gvm_grid_release returns it).
struct Arena {
  DevBuf buf;
  size_t used = 0;
  static size_t round(size_t n) { return (n + 511) & ~(size_t)511; }
  int reserve(size_t total) { used = 0; return buf.ensure(total + 512); }
  template <class T> T* take(size_t count) {
    T* p = reinterpret_cast<T*>(static_cast<char*>(buf.p) + used);
    used += round(count * sizeof(T));
    return used <= buf.bytes ? p : nullptr;
  }
};
thread_local Arena g_arena_a, g_arena_b, g_arena_c, g_arena_r;
// arena A of one tile-replay gridding call: the slice (zz samples), the grids, per-tile and per-sample tables
size_t grid_arena_a_bytes(size_t zz, size_t MN, long ntiles, int world, size_t ck_elems) {
  const size_t scan_n2 = gvm_scan_temp_bytes(2 * zz), scan_mn = gvm_scan_temp_bytes(MN), sort_tiles = gvm_sort_temp_bytes((size_t)ntiles);
  size_t tmpA = scan_n2 > scan_mn ? scan_n2 : scan_mn;
  tmpA = tmpA > sort_tiles ? tmpA : sort_tiles;
  return zz * 24 + zz * 8 + zz * 4 + ck_elems * 4 + MN * 4 + MN * 8 + MN * 4 + MN * 4 + 4 * (size_t)ntiles * 4 + 3 * (2 * zz * 4) +
         ((size_t)2 * world + 1) * 4 + (size_t)world * 2 * world * 4 + tmpA + 24 * 512;
}
// a carved buffer with DevBuf's accessors
struct Raw {
  void* p;
  template <class T> T* as() { return static_cast<T*>(p); }
};

// result of the last gvm_grid_block of this thread (device resident until fetched)
struct GridResult {
  DevBuf uvw, Vo, w;           // merge path (GVM_GRID_MERGE=1)
  double* uvw_p = nullptr;     // where the compacted samples are: the buffers above or the result arena
  float2* Vo_p = nullptr;
  float* w_p = nullptr;
  long count = 0;
};
thread_local GridResult g_grid_result;
// work buffers of gvm_grid_block, kept between calls (a cudaMalloc/cudaFree pair of several GB per
// block costs more than the kernels); gvm_grid_release() returns them
struct GridWork {
  DevBuf uvw, Vo, w, ck, k0, v0, tmp, gw, gV, flags, pos, start;   // merge path (the tile replay uses the arenas)
};
thread_local GridWork g_grid_work;

// stable in-place sort of (key, value) pairs by the low end_bit bits of the key (sort.cu)
int sort_pairs(DevBuf& tmp, uint32_t* keys, uint32_t* vals, long n, int end_bit, cudaStream_t stream = nullptr) {
  if (n <= 0) return 0;
  if (tmp.ensure(gvm_sort_temp_bytes((size_t)n))) return 1;
  return gvm_sort_pairs_u32(keys, vals, (size_t)n, end_bit, tmp.p, stream);
}
// out = exclusive prefix sum of in (n 32-bit counts; out may alias in)
int exclusive_scan(DevBuf& tmp, const void* in, void* out, long n, cudaStream_t stream = nullptr) {
  if (n <= 0) return 0;
  if (tmp.ensure(gvm_scan_temp_bytes((size_t)n))) return 1;
  if (out != in) WG_CUDA(cudaMemcpyAsync(out, in, (size_t)n * 4, cudaMemcpyDeviceToDevice, stream));
  return gvm_exclusive_scan_u32(static_cast<uint32_t*>(out), (size_t)n, tmp.p, stream);
}
__global__ void __launch_bounds__(256) k_max_int(const int* __restrict__ v, long n, int* __restrict__ out) {
  int m = 0;
  for (long i = blockIdx.x * 256L + threadIdx.x; i < n; i += (long)gridDim.x * 256) m = max(m, v[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// GVM_GRID_TIMING=1: wall time of the phases of gvm_grid_block / gvm_weights on stderr (device-synchronised)
struct PhaseTimer {
  bool on;
  double t0;
  static double now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
  }
  PhaseTimer() {
    const char* v = getenv("GVM_GRID_TIMING");
    on = v && *v == '1';
    t0 = on ? now() : 0.0;
  }
  void mark(const char* what) {
    if (!on) return;
    cudaDeviceSynchronize();
    const double t = now();
    fprintf(stderr, "[gvm timing] %-28s %9.3f ms\n", what, 1e3 * (t - t0));
    t0 = t;
  }
};

// Replay of the sorted pair records tile by tile (rec, tstart, tend -> gw, gV): tiles in decreasing order of their
// sample count, one block of 256 threads (one per cell) per tile. ordk / ord: ntiles words each; sort_tmp: gvm_sort_temp_bytes(ntiles).
int tile_replay_raw(uint32_t* ordk, uint32_t* ord, void* sort_tmp, const int* tstart, const int* tend, const float4* rec,
                    const float* ck, float* gw, float2* gV, long ntiles, int ntx, int ck_m, int ck_n, int sx, int sy,
                    long M, long N, cudaStream_t stream) {
  k_tile_order_keys<<<(int)((ntiles + 255) / 256), 256, 0, stream>>>(tstart, tend, ntiles, ordk, ord);
  WG_CUDA(cudaGetLastError());
  if (gvm_sort_pairs_u32(ordk, ord, (size_t)ntiles, 31, sort_tmp, stream)) return 1;
  k_grid_tiles_cells<<<(unsigned)ntiles, kTile * kTile, 0, stream>>>(ord, tstart, tend, rec, ck, ck_m, ck_n, sx, sy, M, N, ntx, gw, gV);
  WG_CUDA(cudaGetLastError());
  return 0;
}
// std::accumulate(weights, 0.0f) per block, then summed over the blocks (src/briggsweightingscheme.cu:46-57)
float briggs_sum_of_weights(int nblocks, const int64_t* Z, float* const* w) {
  float sum_w = 0.0f;
  for (int b = 0; b < nblocks; b++) {
    float acc = 0.0f;
    const float* wb = w[b];
    for (long z = 0; z < (long)Z[b]; z++) acc += wb[z];
    sum_w += acc;
  }
  return sum_w;
}


// Briggs' second order-dependent scalar: sum over m, n in [N/2, N) of grid^2 as ONE sequential fp32 sum
// (src/briggsweightingscheme.cu:96-106). Adding the +0.0 of an empty cell leaves a non-negative fp32 sum unchanged, so
// only the cells whose square is non-zero matter, in the same order: they are compacted on the device in row-major
// half-plane order (flags -> exclusive scan -> scatter) and the host adds the few that remain — instead of copying the
// whole grid back (268 MB at 8192^2) and walking 33 M cells.
__global__ void __launch_bounds__(256) k_half_flags(const float* __restrict__ grid, long M, long N, long nh,
                                                    uint32_t* __restrict__ flags) {
  const long c = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (c >= M * nh) return;
  const long m = c / nh, n = N / 2 + c % nh;
  const float v = grid[N * m + n];
  flags[c] = (v * v != 0.0f) ? 1u : 0u;
}
__global__ void __launch_bounds__(256) k_half_compact(const float* __restrict__ grid, long M, long N, long nh,
                                                      const uint32_t* __restrict__ pos, float* __restrict__ out) {
  const long c = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (c >= M * nh) return;
  const long m = c / nh, n = N / 2 + c % nh;
  const float v = grid[N * m + n];
  if (v * v != 0.0f) out[pos[c]] = v;
}
inline size_t half_plane_scratch_bytes(long M, long N) {
  const size_t nhalf = (size_t)M * (size_t)(N - N / 2);
  return 2 * ((nhalf * 4 + 511) & ~(size_t)511) + gvm_scan_temp_bytes(nhalf) + 512;
}
// sum_g2 += the half plane's squares, in the reference's order; scratch: half_plane_scratch_bytes(M, N) device bytes
int half_plane_sum_sq(const float* d_grid, long M, long N, void* scratch, cudaStream_t st, float* sum_g2) {
  const long nh = N - N / 2;
  const size_t nhalf = (size_t)M * (size_t)nh;
  if (nhalf == 0) return 0;
  const size_t plane = (nhalf * 4 + 511) & ~(size_t)511;
  uint32_t* pos = static_cast<uint32_t*>(scratch);
  float* vals = reinterpret_cast<float*>(static_cast<char*>(scratch) + plane);
  void* scan_tmp = static_cast<char*>(scratch) + 2 * plane;
  const int blocks = (int)((nhalf + 255) / 256);
  k_half_flags<<<blocks, 256, 0, st>>>(d_grid, M, N, nh, pos);
  uint32_t last_flag = 0, last_pos = 0;
  WG_CUDA(cudaMemcpyAsync(&last_flag, pos + (nhalf - 1), 4, cudaMemcpyDeviceToHost, st));
  if (gvm_exclusive_scan_u32(pos, nhalf, scan_tmp, st)) return 1;
  WG_CUDA(cudaMemcpyAsync(&last_pos, pos + (nhalf - 1), 4, cudaMemcpyDeviceToHost, st));
  k_half_compact<<<blocks, 256, 0, st>>>(d_grid, M, N, nh, pos, vals);
  WG_CUDA(cudaGetLastError());
  WG_CUDA(cudaStreamSynchronize(st));
  const size_t count = (size_t)last_pos + last_flag;
  if (count == 0) return 0;
  std::vector<float> h(count);
  if (gvm_fast_d2h(h.data(), vals, count * 4, st)) return 1;
  float acc = *sum_g2;
  for (size_t i = 0; i < count; i++) acc += h[i] * h[i];
  *sum_g2 = acc;
  return 0;
}

// UVTaper::getValue (include/classes/uvtaper.cuh:100-118), host libm, folded coordinates
void apply_taper_host(const gvm_taper* t, int scheme, long Z, const double* uvw_m, float freq, float* w) {
  const float cb = cosf(t->bpa), sb = sinf(t->bpa), s2 = sinf(2.0f * t->bpa);
  const float a = (cb * cb) / (2.0f * t->sigma_maj * t->sigma_maj) + (sb * sb) / (2.0f * t->sigma_min * t->sigma_min);
  const float b = s2 / (2.0f * t->sigma_maj * t->sigma_maj) - s2 / (2.0f * t->sigma_min * t->sigma_min);
  const float c = (sb * sb) / (2.0f * t->sigma_maj * t->sigma_maj) + (cb * cb) / (2.0f * t->sigma_min * t->sigma_min);
  for (long z = 0; z < Z; z++) {
    double u = gvm_metres_to_lambda(uvw_m[3 * z], freq), v = gvm_metres_to_lambda(uvw_m[3 * z + 1], freq);
    if (scheme != GVM_W_RADIAL && u < 0.0) { u *= -1.0; v *= -1.0; }
    const double x = u - t->u_0, y = v - t->v_0;
    w[z] *= (float)(t->amplitude * exp(-a * x * x - b * x * y - c * y * y));
  }
}

}  // namespace

extern "C" {

static int grid_block_core(gvm_engine* e, int device, cudaStream_t st, long M, long N, double deltau, double deltav, float freq,
                           int64_t Z, const double* uvw_m, const float* Vo, const float* w, const float* ckernel, int ck_m,
                           int ck_n, int sx, int sy, int64_t* nout);

int gvm_weights(int device, int scheme, float robust, int64_t M, int64_t N, double deltau, double deltav,
                int nblocks, const int64_t* Z, const double* const* uvw_m, const float* freqs,
                float* const* w, const gvm_taper* taper) {
  if (scheme < GVM_W_NATURAL || scheme > GVM_W_RADIAL) { gvm_set_error("gvm_weights: unknown scheme %d", scheme); return 1; }
  if (scheme == GVM_W_BRIGGS && (robust < -2.0f || robust > 2.0f)) {
    gvm_set_error("gvm_weights: Briggs robust must be in [-2, 2] (src/briggsweightingscheme.cu:13-21)");
    return 1;
  }
  if (M * N >= (int64_t)kNoCell) { gvm_set_error("gvm_weights: grid too large"); return 1; }
  const bool use_taper = taper && taper->enabled;
  if (scheme != GVM_W_NATURAL) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
      gvm_set_error("gvm_weights: no CUDA device %d (no CPU fallback)", device);
      return 1;
    }
    WG_CUDA(cudaSetDevice(device));
  }
  if (scheme == GVM_W_NATURAL || scheme == GVM_W_RADIAL) {
    DevBuf d_uvw, d_w;
    for (int b = 0; b < nblocks; b++) {
      const long z = (long)Z[b];
      if (scheme == GVM_W_RADIAL && z > 0) {
        if (d_uvw.ensure((size_t)z * 24) || d_w.ensure((size_t)z * 4)) return 1;
        if (gvm_fast_h2d(d_uvw.p, uvw_m[b], (size_t)z * 24, 0)) return 1;
        if (gvm_fast_h2d(d_w.p, w[b], (size_t)z * 4, 0)) return 1;
        k_radial<<<(int)((z + 255) / 256), 256>>>(d_uvw.as<double>(), z, freqs[b], d_w.as<float>());
        WG_CUDA(cudaGetLastError());
        if (gvm_fast_d2h(w[b], d_w.p, (size_t)z * 4, 0)) return 1;
      }
      if (use_taper) apply_taper_host(taper, scheme, z, uvw_m[b], freqs[b], w[b]);
    }
    return 0;
  }

  const size_t MN = (size_t)(M * N);
  const double adu = fabs(deltau), adv = fabs(deltav);
  int end_bit = 1;
  while (end_bit < 32 && (1ull << end_bit) <= MN) end_bit++;
  end_bit = 32;  // the off-grid sentinel is all ones: sort on all 32 bits
  long zmax = 1;
  for (int b = 0; b < nblocks; b++) zmax = Z[b] > zmax ? (long)Z[b] : zmax;
  if (zmax >= (long)0x7FFFFFFF) { gvm_set_error("gvm_weights: block too large"); return 1; }
  PhaseTimer pt;
  Arena& A = g_arena_a;   // one allocation for everything (kept between calls)
  const size_t half_bytes = scheme == GVM_W_BRIGGS ? half_plane_scratch_bytes(M, N) : 0;
  if (A.reserve(MN * 4 + (size_t)zmax * (24 + 4 + 4 + 4) + gvm_sort_temp_bytes((size_t)zmax) + half_bytes + 8 * 512)) return 1;
  pt.mark("weights: arena");
  Raw d_grid{A.take<float>(MN)}, d_uvw{A.take<double>((size_t)zmax * 3)}, d_w{A.take<float>((size_t)zmax)},
      d_k0{A.take<uint32_t>((size_t)zmax)}, d_v0{A.take<uint32_t>((size_t)zmax)};
  void* d_half = A.take<char>(half_bytes);
  void* d_tmp = A.take<char>(gvm_sort_temp_bytes((size_t)zmax));
  if (!d_tmp) { gvm_set_error("gvm_weights: arena too small"); return 1; }
  WG_CUDA(cudaMemset(d_grid.p, 0, MN * 4));

  auto load_and_sort = [&](int b) -> int {
    const long z = (long)Z[b];
    if (gvm_fast_h2d(d_uvw.p, uvw_m[b], (size_t)z * 24, 0)) return 1;
    if (gvm_fast_h2d(d_w.p, w[b], (size_t)z * 4, 0)) return 1;
    k_weight_cells<<<(int)((z + 255) / 256), 256>>>(d_uvw.as<double>(), z, freqs[b], adu, adv, M, N,
                                                    d_k0.as<uint32_t>(), d_v0.as<uint32_t>());
    WG_CUDA(cudaGetLastError());
    return z > 0 ? gvm_sort_pairs_u32(d_k0.as<uint32_t>(), d_v0.as<uint32_t>(), (size_t)z, end_bit, d_tmp, nullptr) : 0;
  };

  float f_squared = 0.0f;
  if (scheme == GVM_W_BRIGGS) {
    float sum_w = 0.0f, sum_g2 = 0.0f;
    // the reference's sequential fp32 sum of all weights (src/briggsweightingscheme.cu:46-57) cannot be split
    // without changing its rounding; it runs on a host thread next to the GPU's first pass
    std::thread sum_thread([&] { sum_w = briggs_sum_of_weights(nblocks, Z, w); });
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{sum_thread};
    for (int b = 0; b < nblocks; b++) {
      const long z = (long)Z[b];
      if (z > 0) {
        if (load_and_sort(b)) return 1;
        k_cell_accumulate<<<(int)((z + 255) / 256), 256>>>(d_k0.as<uint32_t>(), d_v0.as<uint32_t>(), z,
                                                           d_w.as<float>(), d_grid.as<float>());
        WG_CUDA(cudaGetLastError());
      }
      // the first-pass grid is never cleared between blocks and the half-plane sum of squares is
      // taken after each one (src/briggsweightingscheme.cu:59-106)
      pt.mark("weights: upload + cells + sort + cell sums");
      if (half_plane_sum_sq(d_grid.as<float>(), M, N, d_half, 0, &sum_g2)) return 1;
      pt.mark("weights: half-plane sum of squares");
    }
    sum_thread.join();
    pt.mark("weights: wait for the host sum of weights");
    const float avg = sum_g2 / sum_w;
    f_squared = (5.0f * powf(10.0f, -robust)) * (5.0f * powf(10.0f, -robust)) / avg;
    WG_CUDA(cudaMemset(d_grid.p, 0, MN * 4));
  }
  // with a single block the first Briggs pass has left this block's sorted cells and weights on the device
  const bool sorted_resident = scheme == GVM_W_BRIGGS && nblocks == 1;
  for (int b = 0; b < nblocks; b++) {
    const long z = (long)Z[b];
    if (z > 0) {
      if (!sorted_resident && load_and_sort(b)) return 1;
      const int blocks = (int)((z + 255) / 256);
      k_cell_accumulate<<<blocks, 256>>>(d_k0.as<uint32_t>(), d_v0.as<uint32_t>(), z, d_w.as<float>(),
                                         d_grid.as<float>());
      k_weight_apply<<<blocks, 256>>>(d_k0.as<uint32_t>(), d_v0.as<uint32_t>(), z, d_grid.as<float>(),
                                      scheme == GVM_W_BRIGGS, f_squared, d_w.as<float>());
      k_clear_cells<<<blocks, 256>>>(d_k0.as<uint32_t>(), z, d_grid.as<float>());
      WG_CUDA(cudaGetLastError());
      if (gvm_fast_d2h(w[b], d_w.p, (size_t)z * 4, 0)) return 1;
      pt.mark("weights: second pass + weights to host");
    }
    if (use_taper) apply_taper_host(taper, scheme, z, uvw_m[b], freqs[b], w[b]);
  }
  return 0;
}

int gvm_grid_block(int device, int64_t M, int64_t N, double deltau, double deltav, float freq, int64_t Z,
                   const double* uvw_m, const float* Vo, const float* w, const float* ckernel, int ck_m,
                   int ck_n, int support_x, int support_y, double* uvw_out, float* Vo_out, float* w_out,
                   int64_t* nout) {
  if (!nout || Z < 0 || ck_m < 1 || ck_n < 1 || support_x < 0 || support_y < 0) {
    gvm_set_error("gvm_grid_block: bad argument");
    return 1;
  }
  if ((2 * support_x + 1) * (2 * support_y + 1) > kMaxTaps) {
    gvm_set_error("gvm_grid_block: kernel support %d x %d exceeds %d taps", support_x, support_y, kMaxTaps);
    return 1;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    gvm_set_error("gvm_grid_block: no CUDA device %d (no CPU fallback)", device);
    return 1;
  }
  WG_CUDA(cudaSetDevice(device));
  const size_t MN = (size_t)(M * N);
  const long n2 = 2 * (long)Z;
  const size_t ext = (size_t)(M + 2 * support_y) * (size_t)(N + 2 * support_x);
  if (ext >= (size_t)kNoCell || n2 >= (long)0x7FFFFFFF) { gvm_set_error("gvm_grid_block: problem too large"); return 1; }
  *nout = 0;
  {
    // default: the tile-sequential replay (the same code as the multi-rank path, with one rank); the per-cell k-way
    // merge below is the fallback for grids beyond the 16-bit centre packing and the cross-check (GVM_GRID_MERGE=1);
    // both give the reference's summation order, i.e. bit-identical results
    const char* force_merge = getenv("GVM_GRID_MERGE");
    const bool fits = M + 2L * support_y < 65536 && N + 2L * support_x < 65536 && n2 <= 400000000L;
    if (fits && !(force_merge && *force_merge == '1')) {
      if (grid_block_core(nullptr, device, nullptr, M, N, deltau, deltav, freq, Z, uvw_m, Vo, w, ckernel, ck_m, ck_n, support_x,
                          support_y, nout))
        return 1;
      if (uvw_out || Vo_out || w_out) return gvm_grid_fetch(uvw_out, Vo_out, w_out);
      return 0;
    }
  }
  PhaseTimer pt;
  GridWork& wk = g_grid_work;
  DevBuf &d_uvw = wk.uvw, &d_Vo = wk.Vo, &d_w = wk.w, &d_ck = wk.ck, &d_k0 = wk.k0, &d_v0 = wk.v0,
         &d_tmp = wk.tmp, &d_gw = wk.gw, &d_gV = wk.gV, &d_flags = wk.flags, &d_pos = wk.pos, &d_start = wk.start;
  GridResult& res = g_grid_result;   // compacted output stays on the device until it is fetched
  DevBuf &d_uo = res.uvw, &d_Vout = res.Vo, &d_wo = res.w;
  res.count = 0;
  const size_t zz = (size_t)(Z > 0 ? Z : 1);
  if (d_uvw.ensure(zz * 24) || d_Vo.ensure(zz * 8) || d_w.ensure(zz * 4) || d_ck.ensure((size_t)ck_m * ck_n * 4) ||
      d_k0.ensure(2 * zz * 4) || d_v0.ensure(2 * zz * 4) || d_gw.ensure(MN * 4) || d_gV.ensure(MN * 8) || d_flags.ensure(MN * 4) || d_pos.ensure(MN * 4))
    return 1;
  if (Z > 0) {
  pt.mark("grid: device buffers");
    if (gvm_fast_h2d(d_uvw.p, uvw_m, zz * 24, 0) || gvm_fast_h2d(d_Vo.p, Vo, zz * 8, 0) || gvm_fast_h2d(d_w.p, w, zz * 4, 0)) return 1;
  }
  pt.mark("grid: upload");
  WG_CUDA(cudaMemcpy(d_ck.p, ckernel, (size_t)ck_m * ck_n * 4, cudaMemcpyHostToDevice));
  {
    if (n2 > 0) {
      k_grid_centres<<<(int)((n2 + 255) / 256), 256>>>(d_uvw.as<double>(), (long)Z, freq, deltau, deltav, M, N,
                                                       support_x, support_y, d_k0.as<uint32_t>(),
                                                       d_v0.as<uint32_t>());
      WG_CUDA(cudaGetLastError());
      if (sort_pairs(d_tmp, d_k0.as<uint32_t>(), d_v0.as<uint32_t>(), n2, 32)) return 1;
    }
    // start table over the extended grid (ext + 1 entries): histogram of the centre cells + exclusive scan
    if (d_start.ensure((ext + 2) * 4)) return 1;
    WG_CUDA(cudaMemset(d_start.p, 0, (ext + 2) * 4));
    if (n2 > 0) {
      k_cell_count<<<(int)((n2 + 255) / 256), 256>>>(d_k0.as<uint32_t>(), n2, d_start.as<int>());
      WG_CUDA(cudaGetLastError());
    }
    {
      // the merge kernel packs a list's remaining count into 32 - kTapBits bits
      int* d_maxc = d_flags.as<int>();     // free until k_grid_flags
      WG_CUDA(cudaMemset(d_maxc, 0, 4));
      k_max_int<<<1024, 256>>>(d_start.as<int>(), (long)(ext + 1), d_maxc);
      int maxc = 0;
      WG_CUDA(cudaMemcpy(&maxc, d_maxc, 4, cudaMemcpyDeviceToHost));
      if (maxc >= (1 << (32 - kTapBits))) {
        gvm_set_error("gvm_grid_block: %d samples fall into one uv cell (limit %d)", maxc, (1 << (32 - kTapBits)) - 1);
        return 1;
      }
    }
    if (exclusive_scan(d_tmp, d_start.p, d_start.p, (long)(ext + 1))) return 1;
    {
      // threads per CTA from the tap count: 12 bytes of shared state per (thread, tap)
      const int taps = (2 * support_x + 1) * (2 * support_y + 1);
  #define GVM_GRID_ACC(T)                                                                                      \
    do {                                                                                                       \
      const size_t smem = (size_t)taps * (T) * 3 * sizeof(uint32_t);                                           \
      WG_CUDA(cudaFuncSetAttribute(k_grid_accumulate<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      k_grid_accumulate<T><<<(int)((MN + (T) - 1) / (T)), (T), smem>>>(                                        \
          d_start.as<int>(), d_v0.as<uint32_t>(), n2, (long)Z, d_Vo.as<float2>(), d_w.as<float>(),             \
          d_ck.as<float>(), ck_m, ck_n, support_x, support_y, M, N, taps, d_gw.as<float>(), d_gV.as<float2>()); \
    } while (0)
      if (taps <= 49) GVM_GRID_ACC(128);
      else if (taps <= 121) GVM_GRID_ACC(64);
      else GVM_GRID_ACC(32);
  #undef GVM_GRID_ACC
    }
    WG_CUDA(cudaGetLastError());
  }
  k_grid_flags<<<(int)((MN + 255) / 256), 256>>>(d_gw.as<float>(), (long)MN, d_flags.as<int>());
  pt.mark("grid: merge path / flags");
  if (exclusive_scan(d_tmp, d_flags.p, d_pos.p, (long)MN)) return 1;
  int last_pos = 0, last_flag = 0;
  WG_CUDA(cudaMemcpy(&last_pos, d_pos.as<int>() + (MN - 1), 4, cudaMemcpyDeviceToHost));
  WG_CUDA(cudaMemcpy(&last_flag, d_flags.as<int>() + (MN - 1), 4, cudaMemcpyDeviceToHost));
  const long count = (long)last_pos + last_flag;
  if (count > 0) {
    if (d_uo.ensure((size_t)count * 24) || d_Vout.ensure((size_t)count * 8) || d_wo.ensure((size_t)count * 4)) return 1;
    k_grid_compact<<<(int)((MN + 255) / 256), 256>>>(d_gw.as<float>(), d_gV.as<float2>(), d_pos.as<int>(), M, N,
                                                     deltau, deltav, gvm_freq_to_wavelength(freq),
                                                     d_uo.as<double>(), d_Vout.as<float2>(), d_wo.as<float>());
    WG_CUDA(cudaGetLastError());
    WG_CUDA(cudaDeviceSynchronize());
  }
  pt.mark("grid: compact");
  res.uvw_p = d_uo.as<double>(); res.Vo_p = d_Vout.as<float2>(); res.w_p = d_wo.as<float>();
  res.count = count;
  *nout = count;
  // outputs may be omitted: the caller sizes its arrays from *nout and calls gvm_grid_fetch
  if (uvw_out || Vo_out || w_out) return gvm_grid_fetch(uvw_out, Vo_out, w_out);
  return 0;
}

int gvm_grid_reserve(int device, int64_t M, int64_t N, int64_t Zmax, int world) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    gvm_set_error("gvm_grid_reserve: no CUDA device %d (no CPU fallback)", device);
    return 1;
  }
  if (M < 1 || N < 1 || Zmax < 0 || world < 1) { gvm_set_error("gvm_grid_reserve: bad argument"); return 1; }
  WG_CUDA(cudaSetDevice(device));
  const size_t MN = (size_t)(M * N);
  const size_t zz = (size_t)(Zmax / world + 1);
  const long ntiles = (long)((N + kTile - 1) / kTile) * (long)((M + kTile - 1) / kTile);
  const size_t grid = grid_arena_a_bytes(zz, MN, ntiles, world, (size_t)kMaxTaps);
  const size_t weights = MN * 4 + zz * (24 + 4 + 4 + 4) + (size_t)Zmax * 4 + gvm_sort_temp_bytes(zz) + half_plane_scratch_bytes(M, N) + 8 * 512;
  return g_arena_a.reserve(grid > weights ? grid : weights);
}

int gvm_grid_release(void) {
  g_grid_work = GridWork();
  g_arena_a = Arena(); g_arena_b = Arena(); g_arena_c = Arena(); g_arena_r = Arena();
  g_grid_result = GridResult();
  return 0;
}

int gvm_grid_fetch(double* uvw_out, float* Vo_out, float* w_out) {
  GridResult& res = g_grid_result;
  const size_t count = (size_t)res.count;
  if (count > 0) {
    if (uvw_out && gvm_fast_d2h(uvw_out, res.uvw_p, count * 24, 0)) return 1;
    if (Vo_out && gvm_fast_d2h(Vo_out, res.Vo_p, count * 8, 0)) return 1;
    if (w_out && gvm_fast_d2h(w_out, res.w_p, count * 4, 0)) return 1;
  }
  return 0;
}

// =============================================================================================
// Multi-rank preprocessing. Every rank passes the SAME full host arrays; rank r uploads and processes only the
// contiguous slice [Z r / W, Z (r + 1) / W) of every block, so uploads, cell indexing and sorting divide by W.
// The results are bit-identical to the single-rank (and therefore to the reference's one-thread) results:
//  * weights: the per-cell fp32 sums are sequential in ascending sample index; with contiguous slices that is
//    "rank 0's samples, then rank 1's, ..." — the grid travels down the ranks (ncclSend / ncclRecv), every rank
//    continues the sums of its predecessor (k_cell_accumulate starts from what the grid holds), and the last rank
//    broadcasts the finished grid;
//  * gridding: a (tile, sample) pair is owned by the rank that owns the tile; pairs are exchanged all-to-all in
//    (originals | twins) x (source rank) order, which IS ascending doubled-sample order, so the stable sort by
//    tile on the owner leaves every tile's samples in the reference's loop order (src/functions.cu:1418-1508);
//    the tile results are merged with an unsigned max (every cell is non-zero on at most one rank: exact).
static inline void slice_of(int64_t Z, int rank, int world, int64_t* lo, int64_t* hi) {
  *lo = Z * rank / world;
  *hi = Z * (rank + 1) / world;
}

int gvm_weights_dist(gvm_engine* e, int scheme, float robust, int nblocks, const int64_t* Z,
                     const double* const* uvw_m, const float* freqs, float* const* w, const gvm_taper* taper) {
  const gvm_config& g = e->cfg;
  const double deltax = GVM_RPDEG_D * g.DELTAX, deltay = GVM_RPDEG_D * g.DELTAY;
  const double deltau = 1.0 / (g.M * deltax), deltav = 1.0 / (g.N * deltay);
  if (e->world <= 1 || scheme == GVM_W_NATURAL)
    return gvm_weights(g.device, scheme, robust, g.M, g.N, deltau, deltav, nblocks, Z, uvw_m, freqs, w, taper);
  if (scheme < GVM_W_NATURAL || scheme > GVM_W_RADIAL) { gvm_set_error("gvm_weights_dist: unknown scheme %d", scheme); return 1; }
  if (scheme == GVM_W_BRIGGS && (robust < -2.0f || robust > 2.0f)) {
    gvm_set_error("gvm_weights_dist: Briggs robust must be in [-2, 2] (src/briggsweightingscheme.cu:13-21)");
    return 1;
  }
  const long M = g.M, N = g.N;
  if (M * N >= (int64_t)kNoCell) { gvm_set_error("gvm_weights_dist: grid too large"); return 1; }
  WG_CUDA(cudaSetDevice(g.device));
  cudaStream_t st = e->stream;
  const int rank = e->rank, world = e->world;
  const bool use_taper = taper && taper->enabled;
  const size_t MN = (size_t)(M * N);
  const double adu = fabs(deltau), adv = fabs(deltav);
  PhaseTimer pt;
  long zmax = 1, smax = 1;
  for (int b = 0; b < nblocks; b++) {
    zmax = Z[b] > zmax ? (long)Z[b] : zmax;
    const long per = (long)(Z[b] / world) + 1;
    smax = per > smax ? per : smax;
  }
  if (zmax >= (long)0x7FFFFFFF) { gvm_set_error("gvm_weights_dist: block too large"); return 1; }
  Arena& A = g_arena_a;   // one allocation for everything (kept between calls)
  const size_t half_bytes = scheme == GVM_W_BRIGGS ? half_plane_scratch_bytes(M, N) : 0;
  if (A.reserve(MN * 4 + (size_t)smax * (24 + 4 + 4 + 4) + (size_t)zmax * 4 + gvm_sort_temp_bytes((size_t)smax) + half_bytes + 8 * 512)) return 1;
  Raw d_grid{A.take<float>(MN)}, d_uvw{A.take<double>((size_t)smax * 3)}, d_w{A.take<float>((size_t)smax)},
      d_k0{A.take<uint32_t>((size_t)smax)}, d_v0{A.take<uint32_t>((size_t)smax)}, d_wfull{A.take<float>((size_t)zmax)};
  void* d_half = A.take<char>(half_bytes);
  void* d_tmp = A.take<char>(gvm_sort_temp_bytes((size_t)smax));
  if (!d_tmp) { gvm_set_error("gvm_weights_dist: arena too small"); return 1; }

  // this rank's slice of block b on the device (+ its cells, sorted) ; returns the slice
  auto load_slice = [&](int b, int64_t* lo, int64_t* hi, bool cells) -> int {
    slice_of(Z[b], rank, world, lo, hi);
    const long n = (long)(*hi - *lo);
    if (n <= 0) return 0;
    if (gvm_fast_h2d(d_uvw.p, uvw_m[b] + 3 * *lo, (size_t)n * 24, st)) return 1;
    if (gvm_fast_h2d(d_w.p, w[b] + *lo, (size_t)n * 4, st)) return 1;
    if (!cells) return 0;
    k_weight_cells<<<(int)((n + 255) / 256), 256, 0, st>>>(d_uvw.as<double>(), n, freqs[b], adu, adv, M, N,
                                                           d_k0.as<uint32_t>(), d_v0.as<uint32_t>());
    WG_CUDA(cudaGetLastError());
    return gvm_sort_pairs_u32(d_k0.as<uint32_t>(), d_v0.as<uint32_t>(), (size_t)n, 32, d_tmp, st);
  };
  // the grid of cell sums goes down the ranks: receive what the lower ranks summed, continue with this rank's
  // samples, hand it on; the last rank holds the complete sums and broadcasts them
  auto ring_accumulate = [&](long n) -> int {
    if (rank > 0 && gvm_dist_recv(e, d_grid.p, MN * 4, rank - 1)) return 1;
    if (n > 0) {
      k_cell_accumulate<<<(int)((n + 255) / 256), 256, 0, st>>>(d_k0.as<uint32_t>(), d_v0.as<uint32_t>(), n,
                                                                d_w.as<float>(), d_grid.as<float>());
      WG_CUDA(cudaGetLastError());
    }
    if (rank < world - 1 && gvm_dist_send(e, d_grid.p, MN * 4, rank + 1)) return 1;
    return gvm_dist_broadcast_bytes(e, d_grid.p, MN * 4, world - 1);
  };
  // every rank's slice of the new weights -> the whole block on every rank -> host
  auto gather_weights = [&](int b, int64_t lo, int64_t hi) -> int {
    if (hi > lo) WG_CUDA(cudaMemcpyAsync(d_wfull.as<float>() + lo, d_w.p, (size_t)(hi - lo) * 4, cudaMemcpyDeviceToDevice, st));
    if (gvm_dist_group_begin(e)) return 1;
    for (int r = 0; r < world; r++) {
      int64_t rlo, rhi;
      slice_of(Z[b], r, world, &rlo, &rhi);
      if (rhi > rlo && gvm_dist_broadcast_bytes(e, d_wfull.as<float>() + rlo, (size_t)(rhi - rlo) * 4, r)) return 1;
    }
    if (gvm_dist_group_end(e)) return 1;
    return Z[b] > 0 ? gvm_fast_d2h(w[b], d_wfull.p, (size_t)Z[b] * 4, st) : 0;
  };

  if (scheme == GVM_W_RADIAL) {
    for (int b = 0; b < nblocks; b++) {
      int64_t lo, hi;
      if (load_slice(b, &lo, &hi, false)) return 1;
      if (hi > lo) {
        k_radial<<<(int)((hi - lo + 255) / 256), 256, 0, st>>>(d_uvw.as<double>(), (long)(hi - lo), freqs[b], d_w.as<float>());
        WG_CUDA(cudaGetLastError());
      }
      if (gather_weights(b, lo, hi)) return 1;
      if (use_taper) apply_taper_host(taper, scheme, (long)Z[b], uvw_m[b], freqs[b], w[b]);
    }
    return 0;
  }

  float f_squared = 0.0f;
  bool grid_resident = false;   // single block: the first Briggs pass leaves the finished sums (and the sorted slice) in place
  if (scheme == GVM_W_BRIGGS) {
    float sum_w = 0.0f, sum_g2 = 0.0f;
    std::thread sum_thread([&] { sum_w = briggs_sum_of_weights(nblocks, Z, w); });
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{sum_thread};
    WG_CUDA(cudaMemsetAsync(d_grid.p, 0, MN * 4, st));
    for (int b = 0; b < nblocks; b++) {
      int64_t lo, hi;
      if (load_slice(b, &lo, &hi, true)) return 1;
      pt.mark("dist weights: slice upload + cells + sort");
      if (ring_accumulate((long)(hi - lo))) return 1;
      pt.mark("dist weights: ring accumulate + broadcast");
      // the first-pass grid is never cleared between blocks and the half-plane sum of squares is taken after
      // each one (src/briggsweightingscheme.cu:59-106): every rank repeats that sequential host sum
      if (half_plane_sum_sq(d_grid.as<float>(), M, N, d_half, st, &sum_g2)) return 1;
    }
    pt.mark("dist weights: half-plane sum of squares (non-zero cells, compacted)");
    sum_thread.join();
    pt.mark("dist weights: wait for the host sum of weights");
    const float avg = sum_g2 / sum_w;
    f_squared = (5.0f * powf(10.0f, -robust)) * (5.0f * powf(10.0f, -robust)) / avg;
    grid_resident = nblocks == 1;
    if (!grid_resident) WG_CUDA(cudaMemsetAsync(d_grid.p, 0, MN * 4, st));
  } else {
    WG_CUDA(cudaMemsetAsync(d_grid.p, 0, MN * 4, st));
  }
  for (int b = 0; b < nblocks; b++) {
    int64_t lo, hi;
    slice_of(Z[b], rank, world, &lo, &hi);
    if (!grid_resident) {
      if (load_slice(b, &lo, &hi, true)) return 1;
      if (ring_accumulate((long)(hi - lo))) return 1;
    }
    const long n = (long)(hi - lo);
    if (n > 0) {
      k_weight_apply<<<(int)((n + 255) / 256), 256, 0, st>>>(d_k0.as<uint32_t>(), d_v0.as<uint32_t>(), n, d_grid.as<float>(),
                                                             scheme == GVM_W_BRIGGS, f_squared, d_w.as<float>());
      WG_CUDA(cudaGetLastError());
    }
    if (gather_weights(b, lo, hi)) return 1;
    pt.mark("dist weights: apply + all-gather + weights to host");
    if (b + 1 < nblocks) WG_CUDA(cudaMemsetAsync(d_grid.p, 0, MN * 4, st));   // per-block sums in the second pass
    if (use_taper) apply_taper_host(taper, scheme, (long)Z[b], uvw_m[b], freqs[b], w[b]);
  }
  WG_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// The tile-replay gridding of one block on `world` ranks (e == nullptr: one rank, default stream, no collectives).
static int grid_block_core(gvm_engine* e, int device, cudaStream_t st, long M, long N, double deltau, double deltav, float freq,
                           int64_t Z, const double* uvw_m, const float* Vo, const float* w, const float* ckernel, int ck_m,
                           int ck_n, int sx, int sy, int64_t* nout) {
  const int rank = e ? e->rank : 0, world = e ? e->world : 1;
  WG_CUDA(cudaSetDevice(device));
  const size_t MN = (size_t)(M * N);
  *nout = 0;
  PhaseTimer pt;
  GridResult& res = g_grid_result;
  res.count = 0;
  int64_t lo, hi;
  slice_of(Z, rank, world, &lo, &hi);
  const long nloc = (long)(hi - lo), n2 = 2 * nloc;
  const size_t zz = (size_t)(nloc > 0 ? nloc : 1);
  const int ntx = (int)((N + kTile - 1) / kTile), nty = (int)((M + kTile - 1) / kTile);
  const long ntiles = (long)ntx * nty;
  // ---- arena A: the slice, the grids, per-tile and per-sample tables (three arenas per call instead of ~20 cudaMallocs)
  const size_t scan_n2 = gvm_scan_temp_bytes(2 * zz), scan_mn = gvm_scan_temp_bytes(MN), sort_tiles = gvm_sort_temp_bytes((size_t)ntiles);
  size_t tmpA = scan_n2 > scan_mn ? scan_n2 : scan_mn;
  tmpA = tmpA > sort_tiles ? tmpA : sort_tiles;
  Arena& A = g_arena_a;
  if (A.reserve(grid_arena_a_bytes(zz, MN, ntiles, world, (size_t)ck_m * ck_n))) return 1;
  double* d_uvw = A.take<double>(zz * 3);
  float2* d_Vo = A.take<float2>(zz);
  float* d_w = A.take<float>(zz);
  float* d_ck = A.take<float>((size_t)ck_m * ck_n);
  float* d_gw = A.take<float>(MN);
  float2* d_gV = A.take<float2>(MN);
  int* d_flags = A.take<int>(MN);
  int* d_pos = A.take<int>(MN);
  int* d_tstart = A.take<int>((size_t)ntiles);
  int* d_tend = A.take<int>((size_t)ntiles);
  uint32_t* d_ordk = A.take<uint32_t>((size_t)ntiles);
  uint32_t* d_ord = A.take<uint32_t>((size_t)ntiles);
  uint32_t* d_cpos = A.take<uint32_t>(2 * zz);
  int* d_cnt = A.take<int>(2 * zz);
  int* d_off = A.take<int>(2 * zz);
  uint32_t* d_start = A.take<uint32_t>((size_t)2 * world + 1);
  uint32_t* d_counts = A.take<uint32_t>((size_t)world * 2 * world);
  void* d_tmpA = A.take<char>(tmpA);
  if (!d_tmpA) { gvm_set_error("gvm_grid_block_dist: arena A too small"); return 1; }
  pt.mark("dist grid: arena A");
  if (nloc > 0)
    if (gvm_fast_h2d(d_uvw, uvw_m + 3 * lo, zz * 24, st) || gvm_fast_h2d(d_Vo, Vo + 2 * lo, zz * 8, st) ||
        gvm_fast_h2d(d_w, w + lo, zz * 4, st))
      return 1;
  WG_CUDA(cudaMemcpyAsync(d_ck, ckernel, (size_t)ck_m * ck_n * 4, cudaMemcpyHostToDevice, st));
  pt.mark("dist grid: upload of the slice");
  // ---- local (tile, sample) pairs of the slice, in ascending local doubled index (originals, then twins)
  long npairs = 0;
  if (n2 > 0) {
    const int blocks = (int)((n2 + 255) / 256);
    k_tile_count<<<blocks, 256, 0, st>>>(d_uvw, nloc, freq, deltau, deltav, M, N, sx, sy, d_cpos, d_cnt);
    WG_CUDA(cudaGetLastError());
    WG_CUDA(cudaMemcpyAsync(d_off, d_cnt, (size_t)n2 * 4, cudaMemcpyDeviceToDevice, st));
    if (gvm_exclusive_scan_u32(reinterpret_cast<uint32_t*>(d_off), (size_t)n2, d_tmpA, st)) return 1;
    int last_off = 0, last_cnt = 0;
    WG_CUDA(cudaMemcpyAsync(&last_off, d_off + (n2 - 1), 4, cudaMemcpyDeviceToHost, st));
    WG_CUDA(cudaMemcpyAsync(&last_cnt, d_cnt + (n2 - 1), 4, cudaMemcpyDeviceToHost, st));
    WG_CUDA(cudaStreamSynchronize(st));
    npairs = (long)last_off + last_cnt;
    if (npairs >= (long)0x7FFFFFFF) { gvm_set_error("gvm_grid_block_dist: too many (tile, sample) pairs on one rank"); return 1; }
  }
  // ---- arena B: the pairs of this rank (tile id, sample, destination) and the send buffers
  const size_t np = (size_t)(npairs > 0 ? npairs : 1);
  Arena& B = g_arena_b;
  if (B.reserve(np * (4 * 5 + 16) + gvm_sort_temp_bytes(np) + 16 * 512)) return 1;
  uint32_t* d_pk = B.take<uint32_t>(np);     // tile id of the pair
  uint32_t* d_pz = B.take<uint32_t>(np);     // local doubled sample index
  uint32_t* d_dk = B.take<uint32_t>(np);     // destination key
  uint32_t* d_dv = B.take<uint32_t>(np);     // pair index, sorted by destination
  uint32_t* d_sk = B.take<uint32_t>(np);     // send: tile ids in destination order
  float4* d_srec = B.take<float4>(np);       // send: records in destination order
  void* d_tmpB = B.take<char>(gvm_sort_temp_bytes(np));
  if (!d_tmpB) { gvm_set_error("gvm_grid_block_dist: arena B too small"); return 1; }
  std::vector<uint32_t> dstart((size_t)2 * world + 1, (uint32_t)npairs);
  if (npairs > 0) {
    const int blocks = (int)((n2 + 255) / 256), pb = (int)((npairs + 255) / 256);
    k_tile_emit<<<blocks, 256, 0, st>>>(d_cpos, d_off, n2, M, N, sx, sy, ntx, d_pk, d_pz);
    k_pair_dest<<<pb, 256, 0, st>>>(d_pk, d_pz, npairs, nloc, ntx, world, d_dk, d_dv);
    WG_CUDA(cudaGetLastError());
    int bits = 1;
    while ((1 << bits) < 2 * world) bits++;
    if (gvm_sort_pairs_u32(d_dk, d_dv, (size_t)npairs, bits, d_tmpB, st)) return 1;
    std::vector<uint32_t> init((size_t)2 * world + 1, (uint32_t)npairs);
    WG_CUDA(cudaMemcpyAsync(d_start, init.data(), init.size() * 4, cudaMemcpyHostToDevice, st));
    k_dest_starts<<<pb, 256, 0, st>>>(d_dk, npairs, d_start);
    k_pair_pack<<<pb, 256, 0, st>>>(d_dv, d_pk, d_pz, npairs, nloc, d_cpos, d_Vo, d_w, d_sk, d_srec);
    WG_CUDA(cudaGetLastError());
    WG_CUDA(cudaMemcpyAsync(dstart.data(), d_start, dstart.size() * 4, cudaMemcpyDeviceToHost, st));
    WG_CUDA(cudaStreamSynchronize(st));
    for (int d = 2 * world - 1; d >= 0; d--)     // absent destinations: empty range in front of the next present one
      if (dstart[d] > dstart[d + 1]) dstart[d] = dstart[d + 1];
  }
  pt.mark("dist grid: local pairs + destination sort");
  // ---- counts of every (source, destination): counts[s][d], d = half * W + owner
  std::vector<uint32_t> counts((size_t)world * 2 * world, 0u);
  {
    std::vector<uint32_t> mine((size_t)2 * world);
    for (int d = 0; d < 2 * world; d++) mine[d] = dstart[d + 1] - dstart[d];
    if (world > 1) {
      WG_CUDA(cudaMemcpyAsync(d_counts + (size_t)rank * 2 * world, mine.data(), mine.size() * 4, cudaMemcpyHostToDevice, st));
      if (gvm_dist_group_begin(e)) return 1;
      for (int r = 0; r < world; r++)
        if (gvm_dist_broadcast_bytes(e, d_counts + (size_t)r * 2 * world, (size_t)2 * world * 4, r)) return 1;
      if (gvm_dist_group_end(e)) return 1;
      WG_CUDA(cudaMemcpyAsync(counts.data(), d_counts, counts.size() * 4, cudaMemcpyDeviceToHost, st));
      WG_CUDA(cudaStreamSynchronize(st));
    } else {
      counts = mine;
    }
  }
  // receive layout on this rank: (originals | twins) x (source rank) — ascending doubled-sample order
  std::vector<size_t> roff((size_t)2 * world + 1, 0);
  for (int h = 0; h < 2; h++)
    for (int s_ = 0; s_ < world; s_++)
      roff[(size_t)h * world + s_ + 1] = roff[(size_t)h * world + s_] + counts[(size_t)s_ * 2 * world + (size_t)h * world + rank];
  const size_t nrecv = roff[(size_t)2 * world];
  if (nrecv >= (size_t)0x7FFFFFFF) { gvm_set_error("gvm_grid_block_dist: too many pairs for one owner"); return 1; }
  // ---- arena C: what this rank receives as a tile owner
  const size_t nr = nrecv > 0 ? nrecv : 1;
  Arena& Cc = g_arena_c;
  if (Cc.reserve(nr * (4 + 16 + 16 + 4) + gvm_sort_temp_bytes(nr) + 8 * 512)) return 1;
  uint32_t* d_rk = Cc.take<uint32_t>(nr);
  float4* d_rrec = Cc.take<float4>(nr);
  float4* d_rec = Cc.take<float4>(nr);
  uint32_t* d_ridx = Cc.take<uint32_t>(nr);
  void* d_tmpC = Cc.take<char>(gvm_sort_temp_bytes(nr));
  if (!d_tmpC) { gvm_set_error("gvm_grid_block_dist: arena C too small"); return 1; }
  if (world > 1 && gvm_dist_group_begin(e)) return 1;
  for (int peer = 0; peer < world; peer++)
    for (int h = 0; h < 2; h++) {
      const int d = h * world + peer;                                   // what this rank sends to `peer`
      const size_t ns = (size_t)(dstart[d + 1] - dstart[d]);
      const size_t nrv = roff[(size_t)h * world + peer + 1] - roff[(size_t)h * world + peer];   // what it receives from `peer`
      if (peer == rank) {   // this rank's own tiles: a device copy
        if (ns > 0) {
          WG_CUDA(cudaMemcpyAsync(d_rk + roff[(size_t)h * world + peer], d_sk + dstart[d], ns * 4, cudaMemcpyDeviceToDevice, st));
          WG_CUDA(cudaMemcpyAsync(d_rrec + roff[(size_t)h * world + peer], d_srec + dstart[d], ns * 16, cudaMemcpyDeviceToDevice, st));
        }
        continue;
      }
      if (ns > 0) {
        if (gvm_dist_send(e, d_sk + dstart[d], ns * 4, peer)) return 1;
        if (gvm_dist_send(e, d_srec + dstart[d], ns * 16, peer)) return 1;
      }
      if (nrv > 0) {
        if (gvm_dist_recv(e, d_rk + roff[(size_t)h * world + peer], nrv * 4, peer)) return 1;
        if (gvm_dist_recv(e, d_rrec + roff[(size_t)h * world + peer], nrv * 16, peer)) return 1;
      }
    }
  if (world > 1 && gvm_dist_group_end(e)) return 1;
  pt.mark("dist grid: all-to-all of the pairs");
  // ---- owner side: stable sort by tile, records in replay order, replay
  WG_CUDA(cudaMemsetAsync(d_tstart, 0, (size_t)ntiles * 4, st));
  WG_CUDA(cudaMemsetAsync(d_tend, 0, (size_t)ntiles * 4, st));
  if (nrecv > 0) {
    const int rb = (int)((nrecv + 255) / 256);
    k_iota<<<rb, 256, 0, st>>>(d_ridx, (long)nrecv);
    int bits = 1;
    while ((1L << bits) < ntiles) bits++;
    if (gvm_sort_pairs_u32(d_rk, d_ridx, nrecv, bits, d_tmpC, st)) return 1;
    k_recv_gather<<<rb, 256, 0, st>>>(d_rk, d_ridx, (long)nrecv, d_rrec, d_tstart, d_tend, d_rec);
    WG_CUDA(cudaGetLastError());
  }
  if (tile_replay_raw(d_ordk, d_ord, d_tmpA, d_tstart, d_tend, d_rec, d_ck, d_gw, d_gV, ntiles, ntx, ck_m, ck_n, sx, sy, M, N, st))
    return 1;
  pt.mark("dist grid: owner sort + tile replay");
  // every cell was computed by its owner and is exactly zero elsewhere: merge the bit patterns
  if (world > 1) {
    if (gvm_dist_allreduce_u32_max(e, reinterpret_cast<uint32_t*>(d_gw), MN)) return 1;
    if (gvm_dist_allreduce_u32_max(e, reinterpret_cast<uint32_t*>(d_gV), 2 * MN)) return 1;
  }
  k_grid_flags<<<(int)((MN + 255) / 256), 256, 0, st>>>(d_gw, (long)MN, d_flags);
  WG_CUDA(cudaMemcpyAsync(d_pos, d_flags, MN * 4, cudaMemcpyDeviceToDevice, st));
  if (gvm_exclusive_scan_u32(reinterpret_cast<uint32_t*>(d_pos), MN, d_tmpA, st)) return 1;
  int last_pos = 0, last_flag = 0;
  WG_CUDA(cudaMemcpyAsync(&last_pos, d_pos + (MN - 1), 4, cudaMemcpyDeviceToHost, st));
  WG_CUDA(cudaMemcpyAsync(&last_flag, d_flags + (MN - 1), 4, cudaMemcpyDeviceToHost, st));
  WG_CUDA(cudaStreamSynchronize(st));
  const long count = (long)last_pos + last_flag;
  if (count > 0) {
    Arena& R = g_arena_r;   // the compacted result stays there until gvm_grid_fetch
    if (R.reserve((size_t)count * 36 + 4 * 512)) return 1;
    res.uvw_p = R.take<double>((size_t)count * 3);
    res.Vo_p = R.take<float2>((size_t)count);
    res.w_p = R.take<float>((size_t)count);
    k_grid_compact<<<(int)((MN + 255) / 256), 256, 0, st>>>(d_gw, d_gV, d_pos, M, N, deltau, deltav,
                                                            gvm_freq_to_wavelength(freq), res.uvw_p, res.Vo_p, res.w_p);
    WG_CUDA(cudaGetLastError());
    WG_CUDA(cudaStreamSynchronize(st));
  }
  pt.mark("dist grid: merge + compact");
  res.count = count;
  *nout = count;
  return 0;
}

int gvm_grid_block_dist(gvm_engine* e, float freq, int64_t Z, const double* uvw_m, const float* Vo, const float* w,
                        const float* ckernel, int ck_m, int ck_n, int support_x, int support_y, int64_t* nout) {
  const gvm_config& g = e->cfg;
  const double deltax = GVM_RPDEG_D * g.DELTAX, deltay = GVM_RPDEG_D * g.DELTAY;
  const double deltau = 1.0 / (g.M * deltax), deltav = 1.0 / (g.N * deltay);
  const long M = g.M, N = g.N;
  const bool fits = M + 2L * support_y < 65536 && N + 2L * support_x < 65536;   // 16-bit centre packing of the tile replay
  if (e->world <= 1 || !fits)
    return gvm_grid_block(g.device, M, N, deltau, deltav, freq, Z, uvw_m, Vo, w, ckernel, ck_m, ck_n, support_x, support_y,
                          nullptr, nullptr, nullptr, nout);
  if (!nout || Z < 0 || ck_m < 1 || ck_n < 1 || support_x < 0 || support_y < 0) { gvm_set_error("gvm_grid_block_dist: bad argument"); return 1; }
  if ((2 * support_x + 1) * (2 * support_y + 1) > kMaxTaps) {
    gvm_set_error("gvm_grid_block_dist: kernel support %d x %d exceeds %d taps", support_x, support_y, kMaxTaps);
    return 1;
  }
  return grid_block_core(e, g.device, e->stream, M, N, deltau, deltav, freq, Z, uvw_m, Vo, w, ckernel, ck_m, ck_n, support_x,
                         support_y, nout);
}

}  // extern "C"
