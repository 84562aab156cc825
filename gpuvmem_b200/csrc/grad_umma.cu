// grad_umma.cu — the chi2 gradient as a tensor-core contraction (tcgen05 / TMEM, CTA pairs).
//
// Reference: DChi2 (src/functions.cu:3698-3791) computes, per unmasked pixel (i,j),
//   d[i,j] = sum_k w_k (Vr_k.re cos 2 pi phi + Vr_k.im sin 2 pi phi),
//   phi    = x_j u_k + y_i v_k + (z_ij - 1) w_k,
// one sincospif per (pixel, visibility) pair. With the separable part of the w-term
//  this is the real GEMM
//   d[i,j] = sum_k  Qr_k(i) Br_k(j) + Qi_k(i) Bi_k(j),                K' = 2 Z
//   Q_k(i) = exp(+2 pi i (v_k y_i + w_k gB(y_i)))                     ("A", M = i)
//   B_k(j) = w_k |Vr_k| exp(i (arg Vr_k - 2 pi (u_k x_j + w_k gA(x_j))))  ("B", N = j)
// Neither operand can be materialised (B alone is N x 2Z), so K-blocks of both are
// GENERATED into shared memory by CUDA-core warps — fixed-point phases (exact integer
// wrap), MUFU sin/cos, error-compensated split of every fp32 value x into hi = rn16(x)
// and lo = x - hi — and consumed by tcgen05.mma with fp32 accumulation in TMEM:
//   mixed split (default):  Ah*Bh as kind::f16, and the two corrections Al*Bh + Ah*Bl as
//     ONE kind::f8f6f4 contraction over 8-bit-float copies (see split_mixed) on the same
//     accumulator — 2 fp16-MMA times per useful product;
//   fp16x3 (GVM_UMMA_SPLIT=fp16x3, round 1):  Ah*Bh + Ah*Bl + Al*Bh, all kind::f16.
// The dropped Al*Bl is 2^-24 relative. The exponent range is handled by an exact
// power-of-two scale of the amplitudes taken from max_k w|Vr| (reduced in the forward pass).
//
// Tile plan. DChi2 returns early for masked pixels (noise >= noise_cut, :3723-3726), so
// only the unmasked part of the image is covered: 256-row bands starting at the first
// unmasked row, each band cut into tiles of two UMMA column blocks whose width (a
// multiple of 16, <= 256) is fitted to the band's unmasked column extent. A CTA PAIR
// (two SMs of one TPC, cta_group::2) owns one 256 x (2 x nbw) tile for one K slice:
// two UMMA 256 x nbw x 16 accumulators in TMEM. Per visibility each SM generates
// 128 A rows + 2 x nbw/2 B rows for 256 x 2nbw / 2 outputs — half the generation work and
// half the shared-memory operand reads per output of a single-CTA 128 x 256 kernel, which
// is what keeps the tensor pipe busy when every operand byte is computed, not loaded
// (per SM and visibility at nbw = 256: tensor 96 clk, generation ~50 clk of issue slots,
// shared memory 48 + 24 + 12 wavefronts; with the mixed split the tensor time is 64 clk).
//
// Warp roles per CTA (576 threads): 0-11 operand generators (one row per thread: A,
// B block 0, B block 1), 12-15 epilogue (tcgen05.ld -> st/red.global into the split-K
// scratch slice of this tile), 16 record producer (bulk asynchronous copies of the
// visibility arrays, per-tile phase bases -> 16-byte records broadcast to the rows), 17 MMA
// issuer (leader CTA only) + TMEM owner.
// Cross-CTA protocol: every generator warp of BOTH CTAs arrives (cluster scope) on the
// LEADER's operand-full barrier; the leader's tcgen05.commit multicasts the stage-free
// and accumulator-full arrivals to both CTAs; both epilogues arrive on the leader's
// accumulator-empty barrier. The accumulator is drained every `chunk` visibilities:
// TMEM accumulation is fp32 with round-toward-zero per instruction (measured,
// scripts/diag/umma_arith.cu), a shrink that grows linearly with the chunk length
// (6e-9 relative per visibility); the epilogue gives the expected loss back, which leaves
// the gradient of a coherent sky at 1-4e-6 of the fp64 truth at chunk = 4096.
#include <climits>
#include <algorithm>
#include <cstdlib>
#include <string>
#include <type_traits>

#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "gvm_internal.cuh"

namespace {

constexpr int KV = 32;                      // visibilities per operand stage (64 fp16 = one 128 B row)
constexpr int NSTAGE = 2;
constexpr int NVS = 4;
constexpr int BLK_BYTES = 128 * 128;        // one 128-row operand block, K-major SWIZZLE_128B
constexpr int STAGE_BYTES = 6 * BLK_BYTES;  // A_hi | A_lo | B0_hi | B0_lo | B1_hi | B1_lo
constexpr int TILE_I = 256, TILE_J = 512;   // output tile of a CTA pair
constexpr int REC_BYTES = NVS * KV * 16;
constexpr int OFF_RECA = NSTAGE * STAGE_BYTES;
constexpr int OFF_RECB = OFF_RECA + REC_BYTES;
constexpr int OFF_BAR = OFF_RECB + REC_BYTES;
constexpr int VBATCH = 256;                 // visibilities per bulk-copied batch of the five input streams
constexpr int VRING = 2;
constexpr int VB_BYTES = VBATCH * 28;       // du64 | dv64 (8 B) | wz | amp | gam (4 B)
constexpr int NBAR = 2 * NSTAGE + 2 * NVS + 2 + VRING;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int OFF_VIN = (OFF_TMEM + 16 + 127) & ~127;
constexpr int SMEM_BYTES = OFF_VIN + VRING * VB_BYTES + 1024;
// kSplit generator warps share a row: 12 * kSplit generator warps, each thread fills 32 / kSplit visibilities of a stage
__host__ __device__ constexpr int gen_warps(int split) { return 12 * split; }
__host__ __device__ constexpr int nthreads(int split) { return (gen_warps(split) + 6) * 32; }
constexpr uint32_t TMEM_COLS = 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on a barrier that lives in another CTA of the cluster (address from mapa)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// The suspend-time hint lets a waiting warp sleep until the phase completes (or the hint expires) instead of polling:
// the four epilogue warps wait for 64 stages at a time and their polls took 12 % of the issue slots of the generators
// (ncu source view, profiles/r2d_k_grad_umma_mixed_c2_ncu_full_summary.txt).
// Operand hand-over to the tensor core of the pair: the generic-proxy stores were made visible to the async proxy by
// fence.proxy.async; the arrive itself needs no cluster-scope release (which costs MEMBAR.ALL.GPU + ERRBAR per warp and
// stage) — the consumer is tcgen05.mma reading shared memory through the async proxy, not a generic load.
__device__ __forceinline__ void mbar_arrive_operands(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
  }
}
// wait with cluster-scope acquire: the arrivals come from both CTAs of the pair (accumulator hand-over)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
  }
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// arrive (once the MMAs issued so far have retired) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_pair_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// same accumulator, E4M3 operands (K = 32 per instruction): the correction products of the mixed split
__device__ __forceinline__ void tc_mma_pair_f8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// 16 TMEM lanes x 64 columns: thread t holds, for j = 0..7, lanes t/4 (v[4j], v[4j+1]) and t/4 + 8 (v[4j+2], v[4j+3]),
// columns 8j + 2 (t % 4) + {0, 1} — four neighbouring threads cover one 32-byte sector of a row
__device__ __forceinline__ void tc_ld_16x256b_x8(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
        "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
        "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
        "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                             uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ void st_global_v2(float* p, float a, float b) {
  asm volatile("st.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_global_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"):
// rows of 128 B (64 fp16 along K), 8-row groups 1024 B apart (SBO), 16-byte chunk c of
// row r stored at chunk (c ^ (r & 7)). Bit layout: start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version [46,48), layout type [61,64) (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (built per tile): kind::f16, D fp32 (bit 4), A/B fp16 K-major,
// N>>3 at [17,23), M>>4 at [24,29) with M = 256 for cta_group::2.

// fractional turn (top 23 bits of a 0.32 fixed-point phase) -> angle - pi, radians; one SHF + one FFMA
__device__ __forceinline__ float phase_to_angle(uint32_t ph) {
  const float f = __uint_as_float(__funnelshift_r(ph, 0x7Fu, 9));   // 0x3F800000 | (ph >> 9) = 1 + frac
  return fmaf(f, 6.283185307179586f, -9.42477796076938f);
}
__device__ __forceinline__ void split2(float c, float s, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(c, s);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(c - hf.x, s - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// Mixed split (kMixed): x = hi + lo with hi = rn16(x) as before, but the two correction products are formed from 8-bit
// float operands at twice the fp16 tensor rate. The amplitude w|Vr| 2^e (< 2^14.5) rides on the A rows, the B rows are
// unit phasors, and the second block of every operand holds 128 one-byte K elements per row:
//   A rows: [ e5m2(lo) x 64 | e5m2(hi) x 64 ]      B rows: [ e4m3(hi) x 64 | e5m2(lo) x 64 ]
// so bytes 0-63 contract to Al Bh (E5M2 x E4M3) and bytes 64-127 to Ah Bl (E5M2 x E5M2). Everything that carries the
// amplitude, or is 2^-12 of something, is E5M2: its five exponent bits take |Ah| < 2^14.5, |Bl| <= 2^-12 and |Al| from 8
// down to 2^-16 as they are — no scaling instruction in the generators, and visibilities 2^15 times weaker than the
// strongest one (which sets the fp16 scale) still get their correction (with E4M3 for Al a population 2^12 below the
// maximum lost it: 1.9e-4 on that population in the numpy model). Each correction is ~2^-12 of its term and is kept to
// 2^-3 / 2^-4: 1.4e-5 rms of the term (scripts/diag/mixed_split_model.py; the third fp16 product kept it at
// 1e-7 — both are at or below the 1.2e-5 of the phases).
template <bool kIsA>
__device__ __forceinline__ void split_mixed(float c, float s, uint32_t& hi, uint32_t& first8, uint32_t& second8) {
  const __half2 h = __floats2half2_rn(c, s);
  const float2 hf = __half22float2(h);
  const float2 lo = make_float2(c - hf.x, s - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  if (kIsA) {
    first8 = __nv_cvt_float2_to_fp8x2(lo, __NV_SATFINITE, __NV_E5M2);
    second8 = __nv_cvt_halfraw2_to_fp8x2(static_cast<__half2_raw>(h), __NV_SATFINITE, __NV_E5M2);
  } else {
    first8 = __nv_cvt_halfraw2_to_fp8x2(static_cast<__half2_raw>(h), __NV_SATFINITE, __NV_E4M3);
    second8 = __nv_cvt_float2_to_fp8x2(lo, __NV_SATFINITE, __NV_E5M2);
  }
}

// ---------------------------------------------------------------------------
// Per evaluation, per visibility: amplitude and phase of w_k Vr_k, scaled by an exact
// power of two so that the largest amplitude sits near 2^14 (fp16 max is 65504).
__global__ void __launch_bounds__(256) k_grad_coeff(const float2* __restrict__ Vr,
                                                    const float* __restrict__ w, long Z,
                                                    const float* __restrict__ max_in, float im_sign,
                                                    float* __restrict__ inv_scale_out,
                                                    float* __restrict__ amp,
                                                    uint32_t* __restrict__ gam) {
  const float mx = *max_in;   // max_k w * max(|Vr.re|, |Vr.im|)  (forward pass)
  int ex = 0;
  if (mx > 0.f) (void)frexpf(mx, &ex);          // mx = m * 2^ex, m in [0.5, 1)
  const float scale = ldexpf(1.0f, 14 - ex);    // sqrt(2) * mx * scale < 2^14.5
  const long k = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (k == 0) *inv_scale_out = ldexpf(1.0f, ex - 14);
  if (k >= Z) return;
  const float wk = w[k];
  const float2 vr = Vr[k];
  const float re = wk * vr.x, im = im_sign * wk * vr.y;   // im_sign = -1: error maps
  amp[k] = sqrtf(re * re + im * im) * scale;
  const float t = atan2f(im, re) * 0.15915494309189535f;       // turns in [-0.5, 0.5]
  gam[k] = (uint32_t)(int32_t)__float2int_rn(t * 2147483648.0f) << 1;
}

// ---------------------------------------------------------------------------
template <bool kUseW, bool kMixed, int kSplit>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(nthreads(kSplit), 1) k_grad_umma(
    const uint64_t* __restrict__ du64, const uint64_t* __restrict__ dv64,
    const float* __restrict__ wz, const float* __restrict__ amp, const uint32_t* __restrict__ gam,
    const float* __restrict__ gA, const float* __restrict__ gB, const int4* __restrict__ tile_list,
    int ntiles, long Z, int N, int x0, int y0, long klen, int chunk_stages,
    float* __restrict__ scratch) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar0 = sbase + OFF_BAR;
  auto BAR_OP_FULL = [&](int s) { return bar0 + 8u * s; };
  auto BAR_OP_EMPTY = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
  auto BAR_VIS_FULL = [&](int s) { return bar0 + 8u * (2 * NSTAGE + s); };
  auto BAR_VIS_EMPTY = [&](int s) { return bar0 + 8u * (2 * NSTAGE + NVS + s); };
  const uint32_t BAR_ACC_FULL = bar0 + 8u * (2 * NSTAGE + 2 * NVS);
  const uint32_t BAR_ACC_EMPTY = BAR_ACC_FULL + 8u;
  auto BAR_VIN = [&](int s) { return BAR_ACC_EMPTY + 8u + 8u * s; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + OFF_TMEM);

  constexpr int GEN_WARPS = gen_warps(kSplit);
  constexpr int W_EPI = GEN_WARPS, W_PROD = GEN_WARPS + 4, W_MMA = GEN_WARPS + 5;   // W_EPI % 4 == 0: TMEM lane quadrants
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader (issues the MMAs)
  const int tile = blockIdx.x >> 1;
  const int4 tl = tile_list[tile];                  // (i0, j0, column-block width, -)
  const int i0 = tl.x + 128 * (int)rank;            // this CTA's 128 rows of the 256-row tile
  const int nbw = tl.z;                             // UMMA N: multiple of 16, <= 256
  const int hb = nbw >> 1;                          // B rows each CTA supplies per column block
  const int jb = tl.y + hb * (int)rank;             // first column of this CTA's share of block 0
  const int span = nbw + hb;                        // this CTA's B columns lie in [jb, jb + span)
  const int ks = blockIdx.y;
  const long kbeg = ks * klen;
  const long kend = (kbeg + klen < Z) ? kbeg + klen : Z;
  const int nst = (int)((kend - kbeg + KV - 1) / KV);

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; s++) {
      mbar_init(BAR_OP_FULL(s), 2 * GEN_WARPS);   // generator warps of both CTAs (used in the leader)
      mbar_init(BAR_OP_EMPTY(s), 1);
    }
    for (int s = 0; s < NVS; s++) {
      mbar_init(BAR_VIS_FULL(s), 32);
      mbar_init(BAR_VIS_EMPTY(s), GEN_WARPS);
    }
    mbar_init(BAR_ACC_FULL, 1);
    mbar_init(BAR_ACC_EMPTY, 8);                  // 4 epilogue warps x 2 CTAs (used in the leader)
    for (int s = 0; s < VRING; s++) mbar_init(BAR_VIN(s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     sbase + OFF_TMEM), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // barriers of both CTAs initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < GEN_WARPS) {
    // ================================================= operand generators (one row per thread)
    const int grp = (warp >> 2) % 3;         // 0: A rows, 1: B block 0, 2: B block 1
    const int part = warp / 12;              // which 32 / kSplit visibilities of a stage this thread fills
    const int r = tid & 127;
    const uint32_t rowoff = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
    const uint32_t swz = (uint32_t)(r & 7);
    // rows are addressed as (record centre) + dr: A centre i0+64; B centre jb+span/2
    const int dr = (grp == 0) ? r - 64 : (r + nbw * (grp - 1) - (span >> 1));
    const bool active = (grp == 0) || (r < hb);   // narrow column blocks leave B rows idle
    float g2 = 0.f;                          // 2 pi * g(row): w-term slope per wavelength of w
    if (kUseW) {
      if (grp == 0) g2 = 6.283185307179586f * gB[max(0, min(i0 + r, N - 1))];
      else g2 = 6.283185307179586f * gA[max(0, min(jb + nbw * (grp - 1) + r, N - 1))];
    }
    const uint32_t rec0 = sbase + (grp == 0 ? OFF_RECA : OFF_RECB);
    const uint32_t blk_hi = (uint32_t)(2 * grp) * BLK_BYTES + rowoff;
    const uint32_t full_remote0 = mapa_rank(BAR_OP_FULL(0), 0);   // the leader's barriers
    for (int it = 0; it < nst; it++) {
      const int s = it % NSTAGE, vs = it % NVS;
      mbar_wait(BAR_VIS_FULL(vs), (uint32_t)((it / NVS) & 1));
      mbar_wait(BAR_OP_EMPTY(s), (uint32_t)(((it / NSTAGE) & 1) ^ 1));
      const uint32_t hi_row = sbase + (uint32_t)s * STAGE_BYTES + blk_hi, lo_row = hi_row + BLK_BYTES;
      const uint32_t recs = rec0 + (uint32_t)vs * (KV * 16);
      if (!kMixed) {
#pragma unroll 2
        for (int kq = part * (KV / 4 / kSplit); kq < (active ? (part + 1) * (KV / 4 / kSplit) : 0); kq++) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {
            const uint4 rec = ld_shared_v4(recs + (uint32_t)(kq * 4 + kk) * 16);
            float a = phase_to_angle(rec.x + (uint32_t)dr * rec.y);
            if (kUseW) a = fmaf(__uint_as_float(rec.w), g2, a);
            const float am = __uint_as_float(rec.z);          // w|Vr| 2^e on the A rows, 1 on the B rows
            split2(am * __cosf(a), am * __sinf(a), hi[kk], lo[kk]);
          }
          const uint32_t off = ((uint32_t)kq ^ swz) << 4;
          st_shared_v4(hi_row + off, hi[0], hi[1], hi[2], hi[3]);
          st_shared_v4(lo_row + off, lo[0], lo[1], lo[2], lo[3]);
        }
      } else {
        // eight visibilities per pass: two 16-byte chunks of the fp16 row, one chunk in each half of the byte row
        auto pass = [&](auto is_a) {
          constexpr bool kIsA = decltype(is_a)::value;
#pragma unroll
          for (int k8 = part * (KV / 8 / kSplit); k8 < (active ? (part + 1) * (KV / 8 / kSplit) : 0); k8++) {
            uint32_t hi[8], f8[8], s8[8];
#pragma unroll
            for (int kk = 0; kk < 8; kk++) {
              const uint4 rec = ld_shared_v4(recs + (uint32_t)(k8 * 8 + kk) * 16);
              float a = phase_to_angle(rec.x + (uint32_t)dr * rec.y);
              if (kUseW) a = fmaf(__uint_as_float(rec.w), g2, a);
              float c = __cosf(a), sn = __sinf(a);
              if (kIsA) { const float am = __uint_as_float(rec.z); c *= am; sn *= am; }
              split_mixed<kIsA>(c, sn, hi[kk], f8[kk], s8[kk]);
            }
            st_shared_v4(hi_row + (((uint32_t)(2 * k8) ^ swz) << 4), hi[0], hi[1], hi[2], hi[3]);
            st_shared_v4(hi_row + (((uint32_t)(2 * k8 + 1) ^ swz) << 4), hi[4], hi[5], hi[6], hi[7]);
            st_shared_v4(lo_row + (((uint32_t)k8 ^ swz) << 4), f8[0] | (f8[1] << 16), f8[2] | (f8[3] << 16),
                         f8[4] | (f8[5] << 16), f8[6] | (f8[7] << 16));
            st_shared_v4(lo_row + (((uint32_t)(4 + k8) ^ swz) << 4), s8[0] | (s8[1] << 16), s8[2] | (s8[3] << 16),
                         s8[4] | (s8[5] << 16), s8[6] | (s8[7] << 16));
          }
        };
        if (grp == 0) pass(std::true_type{}); else pass(std::false_type{});
      }
      fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_operands(full_remote0 + 8u * s);
        mbar_arrive(BAR_VIS_EMPTY(vs));
      }
    }
  } else if (warp < W_PROD) {
    // ================================================= epilogue: TMEM -> split-K scratch slice
    const int q = warp - W_EPI;
    const int nchunks = (nst + chunk_stages - 1) / chunk_stages;
    const uint32_t acc_empty_remote = mapa_rank(BAR_ACC_EMPTY, 0);
    // 16x256b TMEM loads: four neighbouring threads hold one 32-byte sector of an output row, so every st / red
    // instruction of the warp touches 8 full sectors (the 32x32b shape gave 32 half-used ones: 1.3 % of the step).
    // The adds are fire-and-forget fp32 reductions in L2 and each address is only ever touched by one thread of one
    // CTA, so the result does not depend on scheduling. They are the cost of draining (20 ms of a C2 step, element-
    // bound in L2; the TMEM reads are free).
    float* tile_base = scratch + ((size_t)ks * ntiles + tile) * TILE_I * TILE_J;   // compact scratch: [K slice][tile][256 rows][2 x 256 columns]
    const int n64 = (nbw + 63) >> 6;
    for (int c = 0; c < nchunks; c++) {
      mbar_wait(BAR_ACC_FULL, (uint32_t)(c & 1));
      tc_fence_after();
      // Expected truncation loss of this chunk, given back: every tcgen05.mma adds its K-sum to the accumulator with
      // round-toward-zero (scripts/diag/umma_arith.cu), i.e. loses on average half an ulp of |acc| = 0.5 * 2^-23 * E[1/m]
      // = 4.2e-8 of |acc| per instruction (mantissa m in [1, 2)); for an accumulator that grows from 0 over n
      // instructions that is a shrink by n * 2.1e-8. Measured without the correction: 1.17e-5 at n = 512, 1.56e-5 at
      // n = 768, linear in the chunk length; the factor that minimises the error against the fp64 oracle is 1.9e-8
      // (scripts/diag/noise_like_gradient.py, coherent sky) to 2.3e-8 (C2) per instruction. With it the gradient of a
      // coherent sky is at 3-6e-6 of the oracle at chunk = 4096 instead of 2.3e-5 (what remains is the pixel-to-pixel
      // spread of the loss); a random-walk accumulator (noise-like residuals) keeps 0.58 of the shrink as rms error
      // after the correction, 1.15 without.
      const int stages_here = min(chunk_stages, nst - c * chunk_stages);
      const float unshrink = 1.0f + 2.1e-8f * (float)(stages_here * (kMixed ? 8 : 12));
#pragma unroll 1
      for (int b = 0; b < 4 * n64; b++) {
        const int hh = b & 1, cbn = b >> 1;          // lane half of the quadrant, 64-column block
        const int nb = cbn >= n64, cb = cbn - nb * n64;
        const int col = nb * 256 + cb * 64;
        uint32_t v[32];
        tc_ld_16x256b_x8(tmem_base + ((uint32_t)(32 * q + 16 * hh) << 16) + (uint32_t)col, v);
        tc_wait_ld();
        float* pa = tile_base + (size_t)(128 * rank + 32 * q + 16 * hh + (lane >> 2)) * TILE_J + col + 2 * (lane & 3);
        float* pb = pa + 8 * TILE_J;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          if (cb * 64 + 8 * j < nbw) {   // nbw % 16 == 0
            const float a0 = unshrink * __uint_as_float(v[4 * j]), a1 = unshrink * __uint_as_float(v[4 * j + 1]),
                        b0 = unshrink * __uint_as_float(v[4 * j + 2]), b1 = unshrink * __uint_as_float(v[4 * j + 3]);
            if (c == 0) { st_global_v2(pa + 8 * j, a0, a1); st_global_v2(pb + 8 * j, b0, b1); }
            else { red_global_v2(pa + 8 * j, a0, a1); red_global_v2(pb + 8 * j, b0, b1); }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(acc_empty_remote);
    }
  } else if (warp == W_PROD) {
    // ================================================= record producer (lane = visibility)
    // The five per-visibility input streams are STAGED BY BULK ASYNCHRONOUS COPIES (cp.async.bulk + mbarrier, TMA
    // without a tensor map): batches of VBATCH visibilities in a 2-deep shared-memory ring, the next batch in flight
    // while this warp turns the current one into records. Batches start on multiples of KV = 32 visibilities (128 /
    // 256-byte aligned); a ragged tail is rounded up to 4 visibilities (the arrays carry 8 elements of slack).
    const int ic = i0 + 64, jc = jb + (span >> 1);
    const long nvis = kend - kbeg;
    const int nbatch = (int)((nvis + VBATCH - 1) / VBATCH);
    auto post_batch = [&](int b) {
      const long k0 = kbeg + (long)b * VBATCH;
      long n = kend - k0 < VBATCH ? kend - k0 : VBATCH;
      n = (n + 3) & ~3L;
      const uint32_t dst = sbase + OFF_VIN + (uint32_t)(b % VRING) * VB_BYTES, bar = BAR_VIN(b % VRING);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(n * (kUseW ? 28 : 24))) : "memory");
      auto cp = [&](uint32_t off, const void* src, uint32_t bytes) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst + off), "l"(src), "r"(bytes), "r"(bar) : "memory");
      };
      cp(0, du64 + k0, (uint32_t)n * 8);
      cp(VBATCH * 8, dv64 + k0, (uint32_t)n * 8);
      if (kUseW) cp(VBATCH * 16, wz + k0, (uint32_t)n * 4);
      cp(VBATCH * 20, amp + k0, (uint32_t)n * 4);
      cp(VBATCH * 24, gam + k0, (uint32_t)n * 4);
    };
    if (lane == 0 && nbatch > 0) post_batch(0);
    for (int it = 0; it < nst; it++) {
      const int vs = it % NVS;
      const int b = it / (VBATCH / KV), within = (it % (VBATCH / KV)) * KV + lane;
      if (it % (VBATCH / KV) == 0) {
        // the other ring slot was consumed by the previous batch (program order of this warp): refill it, then wait for ours
        __syncwarp();
        if (lane == 0 && b + 1 < nbatch) post_batch(b + 1);
        mbar_wait(BAR_VIN(b % VRING), (uint32_t)((b / VRING) & 1));
      }
      mbar_wait(BAR_VIS_EMPTY(vs), (uint32_t)(((it / NVS) & 1) ^ 1));
      const long k = kbeg + (long)it * KV + lane;
      uint4 ra = make_uint4(0u, 0u, 0u, 0u), rb = make_uint4(0u, 0u, 0x3F800000u, 0u);   // padding: amplitude 0
      if (k < kend) {
        const uint8_t* vin = sgen + OFF_VIN + (b % VRING) * VB_BYTES;
        const uint64_t du = reinterpret_cast<const uint64_t*>(vin)[within];
        const uint64_t dv = reinterpret_cast<const uint64_t*>(vin + VBATCH * 8)[within];
        const float wzk = kUseW ? reinterpret_cast<const float*>(vin + VBATCH * 16)[within] : 0.f;
        // A: +phase of v_k y_i ; B: arg(Vr_k) - phase of u_k x_j (and -w for the w-term).
        // The per-row increment is ROUNDED to 32 bits: |dr| <= 192 rows => <= 2.3e-8 turns.
        ra.x = (uint32_t)((dv * (uint64_t)(int64_t)(ic - y0)) >> 32);
        ra.y = (uint32_t)((dv + 0x80000000ull) >> 32);
        ra.z = reinterpret_cast<const uint32_t*>(vin + VBATCH * 20)[within];
        ra.w = __float_as_uint(wzk);
        const uint32_t pu = (uint32_t)((du * (uint64_t)(int64_t)(jc - x0)) >> 32);
        rb.x = reinterpret_cast<const uint32_t*>(vin + VBATCH * 24)[within] - pu;
        rb.y = 0u - (uint32_t)((du + 0x80000000ull) >> 32);
        rb.w = __float_as_uint(-wzk);
      }
      *reinterpret_cast<uint4*>(sgen + OFF_RECA + (vs * KV + lane) * 16) = ra;
      *reinterpret_cast<uint4*>(sgen + OFF_RECB + (vs * KV + lane) * 16) = rb;
      mbar_arrive(BAR_VIS_FULL(vs));
    }
  } else if (rank == 0 && lane == 0) {
    // ================================================= MMA issuer (one thread of the leader CTA)
    const uint32_t idesc = (1u << 4) | ((uint32_t)(nbw >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // formats 0: F16 / E4M3
    const uint32_t idesc_albh = idesc | (1u << 7);                 // A E5M2 (bits 7-9 = 1), B E4M3
    const uint32_t idesc_ahbl = idesc | (1u << 7) | (1u << 10);    // A E5M2, B E5M2 (bits 10-12 = 1)
    for (int it = 0; it < nst; it++) {
      const int s = it % NSTAGE;
      const int cpos = it % chunk_stages;
      if (cpos == 0) {
        mbar_wait_cluster(BAR_ACC_EMPTY, (uint32_t)(((it / chunk_stages) & 1) ^ 1));
        tc_fence_after();
      }
      mbar_wait(BAR_OP_FULL(s), (uint32_t)((it / NSTAGE) & 1));
      tc_fence_after();
      const uint32_t st0 = sbase + (uint32_t)s * STAGE_BYTES;
      const uint32_t a_hi = st0, a_lo = st0 + BLK_BYTES;
#pragma unroll
      for (int nb = 0; nb < 2; nb++) {
        const uint32_t b_hi = st0 + (uint32_t)(2 + 2 * nb) * BLK_BYTES, b_lo = b_hi + BLK_BYTES;
        const uint32_t d = tmem_base + (uint32_t)(nb * 256);
#pragma unroll
        for (int kstep = 0; kstep < 4; kstep++) {   // 4 x (K = 16 fp16 = 32 B) inside the 128 B row
          const uint32_t ko = (uint32_t)kstep * 32;
          tc_mma_pair_f16(d, umma_desc_sw128(a_hi + ko), umma_desc_sw128(b_hi + ko), idesc,
                          (cpos > 0 || kstep > 0) ? 1u : 0u);
          if (!kMixed) {
            tc_mma_pair_f16(d, umma_desc_sw128(a_hi + ko), umma_desc_sw128(b_lo + ko), idesc, 1u);
            tc_mma_pair_f16(d, umma_desc_sw128(a_lo + ko), umma_desc_sw128(b_hi + ko), idesc, 1u);
          }
        }
        if (kMixed) {
#pragma unroll
          for (int kstep = 0; kstep < 4; kstep++) {   // 4 x (K = 32 bytes): Al Bh (E5M2 x E4M3) over bytes 0-63, Ah Bl (E5M2 x E5M2) over 64-127
            const uint32_t ko = (uint32_t)kstep * 32;
            tc_mma_pair_f8(d, umma_desc_sw128(a_lo + ko), umma_desc_sw128(b_lo + ko), kstep < 2 ? idesc_albh : idesc_ahbl, 1u);
          }
        }
      }
      tc_commit_pair(BAR_OP_EMPTY(s));                 // stage reusable (both CTAs) once retired
      if (cpos == chunk_stages - 1 || it == nst - 1) tc_commit_pair(BAR_ACC_FULL);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's TMEM / smem / barriers stay alive until both are done
  if (warp == W_MMA) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                 : "memory");
  }
}


// ---------------------------------------------------------------------------
// First / last unmasked column of every image row (one block per row).
__global__ void __launch_bounds__(128) k_row_extent(const float* __restrict__ noise, float noise_cut,
                                                    int N, int2* __restrict__ row_ext) {
  __shared__ int s_min[4], s_max[4];
  const int i = blockIdx.x;
  int lo = INT_MAX, hi = -1;
  for (int j = threadIdx.x; j < N; j += blockDim.x)
    if (noise[(size_t)i * N + j] < noise_cut) {
      lo = min(lo, j);
      hi = max(hi, j);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = lo; s_max[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0)
    row_ext[i] = make_int2(min(min(s_min[0], s_min[1]), min(s_min[2], s_min[3])),
                           max(max(s_max[0], s_max[1]), max(s_max[2], s_max[3])));
}

// Sum the K slices of the compact tile scratch in a fixed order, undo the fp16 scale, then the
// shared finishing math (scale, chain rule, += result).
__global__ void __launch_bounds__(256) k_grad_finish_tiled(
    const float* __restrict__ scratch, int ksplit, int ntiles, const float* __restrict__ inv_scale,
    const float* __restrict__ noise, float noise_cut, const int4* __restrict__ band_tab, int imin,
    GvmFinishParams p) {
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long MN = p.M * p.N;
  if (idx >= MN) return;
  if (noise[idx] >= noise_cut) {  // DChi2 returns early; device_dchi2 was memset to 0
    if (p.dchi2_out) p.dchi2_out[idx] = 0.0f;
    return;
  }
  const int i = (int)(idx / p.N), j = (int)(idx % p.N);
  const int b = (i - imin) / TILE_I;
  const int4 bt = band_tab[b];                       // (jmin, first tile, tiles, tile width)
  const int lj = j - bt.x;
  const int t = lj / bt.w, within = lj - t * bt.w;
  const int nbw = bt.w >> 1;
  const int nb = within >= nbw, c = within - nb * nbw;
  const size_t off = ((size_t)(bt.y + t) * TILE_I + (i - imin - b * TILE_I)) * TILE_J + nb * 256 + c;
  const size_t stride = (size_t)ntiles * TILE_I * TILE_J;
  float d = 0.0f;
  for (int s = 0; s < ksplit; s++) d += scratch[(size_t)s * stride + off];
  d *= *inv_scale;
  gvm_finish_pixel(p, d, idx, i, j);
}

}  // namespace

// max |w_k| * max |g| over the image, in turns: the argument handed to MUFU grows by
// this much beyond [-pi, pi); the approximation degrades slowly, keep it small.
static double wterm_turns(const gvm_engine* e, const GvmChannel& c) {
  const double dx = e->cfg.DELTAX * GVM_RPDEG_D, dy = e->cfg.DELTAY * GVM_RPDEG_D;
  const int N = (int)e->cfg.N;
  const int x0 = (int)c.d.phs_xobs_pix, y0 = (int)c.d.phs_yobs_pix;
  double worst = 0.0;
  const int ends[2] = {0, N - 1};
  for (int a = 0; a < 2; a++) {
    const double x = (ends[a] - x0) * dx, y = (ends[a] - y0) * dy;
    if (x * x >= 1.0 || y * y >= 1.0) return 1e30;
    const double ga = x * x / (1.0 + sqrt(1.0 - x * x)), gb = y * y / (1.0 + sqrt(1.0 - y * y));
    if (ga > worst) worst = ga;
    if (gb > worst) worst = gb;
  }
  return worst * (double)c.max_abs_wz;
}

bool gvm_grad_umma_supported(const gvm_engine* e, const GvmChannel& c) {
  return wterm_turns(e, c) <= 4.0;
}

// (Re)build the tile plan from the noise mask: bands of 256 rows from the first unmasked row,
// tiles of two column blocks fitted to each band's unmasked extent.
static int build_plan(gvm_engine* e) {
  const int N = (int)e->cfg.N;
  if (!e->row_ext) GVM_CUDA(cudaMalloc(&e->row_ext, (size_t)N * sizeof(int2)));
  k_row_extent<<<N, 128, 0, e->stream>>>(e->noise, e->cfg.noise_cut, N, e->row_ext);
  GVM_LAUNCH(e);
  std::vector<int2> ext(N);
  GVM_CUDA(cudaMemcpyAsync(ext.data(), e->row_ext, (size_t)N * sizeof(int2), cudaMemcpyDeviceToHost, e->stream));
  GVM_CUDA(cudaStreamSynchronize(e->stream));
  int imin = -1, imax = -1;
  for (int i = 0; i < N; i++)
    if (ext[i].y >= ext[i].x) { if (imin < 0) imin = i; imax = i; }
  std::vector<int4> tiles, bands;
  long pixels = 0;
  if (imin >= 0) {
    for (int i0 = imin; i0 <= imax; i0 += TILE_I) {
      int jmin = INT_MAX, jmax = -1;
      for (int i = i0; i < i0 + TILE_I && i < N; i++)
        if (ext[i].y >= ext[i].x) { jmin = std::min(jmin, ext[i].x); jmax = std::max(jmax, ext[i].y); }
      if (jmax < jmin) { bands.push_back(make_int4(0, (int)tiles.size(), 0, 32)); continue; }
      const int w = jmax - jmin + 1;
      const int nt = (w + TILE_J - 1) / TILE_J;
      int tw = (w + nt - 1) / nt;
      tw = ((tw + 31) / 32) * 32;                     // two column blocks, each a multiple of 16
      bands.push_back(make_int4(jmin, (int)tiles.size(), nt, tw));
      for (int t = 0; t < nt; t++) tiles.push_back(make_int4(i0, jmin + t * tw, tw / 2, 0));
      pixels += (long)nt * TILE_I * tw;
    }
  }
  cudaFree(e->tile_list); cudaFree(e->band_tab);
  e->tile_list = nullptr; e->band_tab = nullptr;
  if (!tiles.empty()) {
    GVM_CUDA(cudaMalloc(&e->tile_list, tiles.size() * sizeof(int4)));
    GVM_CUDA(cudaMalloc(&e->band_tab, bands.size() * sizeof(int4)));
    GVM_CUDA(cudaMemcpy(e->tile_list, tiles.data(), tiles.size() * sizeof(int4), cudaMemcpyHostToDevice));
    GVM_CUDA(cudaMemcpy(e->band_tab, bands.data(), bands.size() * sizeof(int4), cudaMemcpyHostToDevice));
  }
  e->plan_ntiles = (int)tiles.size();
  e->plan_nbands = (int)bands.size();
  e->plan_imin = imin < 0 ? 0 : imin;
  e->plan_pixels = pixels;
  e->plan_dirty = false;
  return 0;
}

int gvm_grad_umma(gvm_engine* e, GvmChannel& c, const float* I_dev, int flag_opt, int normalize,
                  float* result_dev) {
  const int N = (int)e->cfg.N;
  if (!e->err_variant && !gvm_grad_umma_supported(e, c)) {
    gvm_set_error("gvm_grad_umma: the w-term exceeds 4 turns across the image; use the SIMT kernels");
    return 1;
  }
  // GVM_UMMA_SPLIT=fp16x3 selects the three-product fp16 split (round 1); default: fp16 + two E4M3 correction products.
  // GVM_UMMA_GENSPLIT=1|2: generator warps per operand row (12 or 24 generator warps per CTA).
  bool mixed = true;
  if (const char* s = getenv("GVM_UMMA_SPLIT")) mixed = std::string(s) != "fp16x3";
  int gensplit = 1;
  if (const char* s = getenv("GVM_UMMA_GENSPLIT")) gensplit = atoi(s) == 2 ? 2 : 1;
  if (e->plan_dirty)
    if (build_plan(e)) return 1;
  const int ntiles = e->plan_ntiles;
  if (ntiles == 0) return 0;      // every pixel is masked: the gradient is exactly zero
  if (!c.amp) {
    const size_t z = (size_t)(c.Z > 0 ? c.Z : 1);
    GVM_CUDA(cudaMalloc(&c.amp, (z + 8) * sizeof(float)));
    GVM_CUDA(cudaMalloc(&c.gam, (z + 8) * sizeof(uint32_t)));
    GVM_CUDA(cudaMemsetAsync(c.amp + z, 0, 8 * sizeof(float), e->stream));
    GVM_CUDA(cudaMemsetAsync(c.gam + z, 0, 8 * sizeof(uint32_t), e->stream));
  }
  float* inv_scale = e->red_max + e->red_slots + c.slot;
  k_grad_coeff<<<(int)((c.Z + 255) / 256), 256, 0, e->stream>>>(c.Vr, c.w, c.Z, e->red_max + c.slot,
                                                                 e->err_variant ? -1.0f : 1.0f, inv_scale, c.amp, c.gam);
  GVM_LAUNCH(e);
  const bool use_w = c.max_abs_wz > 0.f && !e->err_variant;
  if (use_w)
    if (gvm_build_pixtab(e, c)) return 1;

  // visibilities per TMEM accumulation chunk (bounds what is left of the round-toward-zero bias after the epilogue's
  // correction; every drain costs 512 KB of L2 reductions per CTA pair: 2048 -> 4096 is 3.8 % of a C2 step)
  long chunk = 4096;
  if (const char* s = getenv("GVM_UMMA_CHUNK")) chunk = atol(s);
  chunk = (chunk / KV) * KV;
  if (chunk < KV) chunk = KV;

  // split K so that tiles * ksplit fills whole waves of CTA pairs; slices >= 2048 samples
  const size_t tile_floats = (size_t)TILE_I * TILE_J;
  const int pairs_per_wave = e->sm_count / 2;
  long max_ks = c.Z / 2048;
  if (max_ks < 1) max_ks = 1;
  while (max_ks > 1 && (size_t)max_ks * ntiles * tile_floats * sizeof(float) > ((size_t)2 << 30)) max_ks--;
  if (max_ks > 1024) max_ks = 1024;
  int best = 1;
  double best_eff = -1.0;
  for (long ks = 1; ks <= max_ks; ks++) {
    const long ctas = (long)ntiles * ks;
    const long waves = (ctas + pairs_per_wave - 1) / pairs_per_wave;
    double eff = (double)ctas / (double)(waves * pairs_per_wave);
    if (waves < 2 && ks < max_ks) eff *= 0.5 + 0.25 * waves;   // prefer >= 2 waves when possible
    if (eff > best_eff + 1e-9) { best_eff = eff; best = (int)ks; }
    if (ctas >= 8L * pairs_per_wave && eff > 0.97) break;
  }
  long klen = (c.Z + best - 1) / best;
  klen = ((klen + KV - 1) / KV) * KV;
  int ksplit = (int)((c.Z + klen - 1) / klen);
  if (ksplit < 1) ksplit = 1;
  if (gvm_ensure_grad_scratch(e, (size_t)ksplit * ntiles * tile_floats)) return 1;
  const int x0 = (int)c.d.phs_xobs_pix, y0 = (int)c.d.phs_yobs_pix;
  dim3 grid(2 * ntiles, ksplit);
  gvm_ev_begin(e);
  int launch_rc = 0;
  auto launch = [&](auto kern, int threads, unsigned bit) {
    // per engine (the attribute belongs to its device): opt in to the dynamic shared memory of this instantiation
    if (!(e->umma_attr_set & bit)) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) { launch_rc = 1; return; }
      e->umma_attr_set |= bit;
    }
    kern<<<grid, threads, SMEM_BYTES, e->stream>>>(c.du64, c.dv64, c.wz, c.amp, c.gam, e->pixtab, e->pixtab + N, e->tile_list,
                                                   ntiles, c.Z, N, x0, y0, klen, (int)(chunk / KV), e->grad_scratch);
  };
  auto pick = [&](auto usew, auto mix) {
    constexpr bool W = decltype(usew)::value, M = decltype(mix)::value;
    constexpr unsigned bit = 1u << (4 * W + 2 * M);
    if (gensplit == 2) launch(k_grad_umma<W, M, 2>, nthreads(2), bit << 1); else launch(k_grad_umma<W, M, 1>, nthreads(1), bit);
  };
  if (use_w) { if (mixed) pick(std::true_type{}, std::true_type{}); else pick(std::true_type{}, std::false_type{}); }
  else { if (mixed) pick(std::false_type{}, std::true_type{}); else pick(std::false_type{}, std::false_type{}); }
  if (launch_rc) { gvm_set_error("gvm_grad_umma: cudaFuncSetAttribute failed"); return 1; }
  gvm_ev_end(e);
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  const long MN = e->cfg.M * e->cfg.N;
  k_grad_finish_tiled<<<(int)((MN + 255) / 256), 256, 0, e->stream>>>(
      e->grad_scratch, ksplit, ntiles, inv_scale, e->noise, e->cfg.noise_cut, e->band_tab, e->plan_imin,
      gvm_finish_params(e, c, I_dev, flag_opt, normalize, result_dev));
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  return 0;
}
