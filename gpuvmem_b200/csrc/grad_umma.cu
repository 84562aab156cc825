// grad_umma.cu — the chi2 gradient as a tensor-core contraction (tcgen05 / TMEM).
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
// wrap), MUFU sin/cos, error-compensated split of every fp32 value x into fp16
// hi = rn(x), lo = rn(x - hi) — and consumed by tcgen05.mma (kind::f16, fp32 accumulate
// in TMEM) as three products  Ah*Bh + Ah*Bl + Al*Bh  (the dropped Al*Bl is 2^-24
// relative). fp16x3 carries the same 11+11 mantissa bits as a 3xTF32 split at twice
// the tensor rate and half the shared-memory bytes; the exponent range is handled by
// an exact power-of-two scale of B taken from max_k w|Vr| (reduced in the forward pass).
//
// Kernel layout (one CTA per SM, 448 threads, cta_group::1, UMMA 128 x 256 x 16):
//   warps 0-3   generate A rows (1 row / thread)        \  st.shared into the canonical
//   warps 4-7   generate B rows (2 rows / thread)       /  K-major SWIZZLE_128B layout
//   warps 8-11  epilogue: tcgen05.ld TMEM -> registers -> += split-K scratch slice
//   warp  12    record producer: coalesced loads of the visibility arrays, per-tile
//               phase bases -> 16-byte records in shared memory (broadcast to the rows)
//   warp  13    one thread issues tcgen05.mma / tcgen05.commit; owns TMEM alloc/dealloc
// Pipelines: records (4 stages), operands (2 stages x 96 KB), accumulator (1 x 256
// TMEM columns, drained every `chunk` visibilities so that no fp32 TMEM accumulation
// runs longer than 2*chunk*3 products).
#include <cstdlib>

#include <cuda_fp16.h>

#include "gvm_internal.cuh"

namespace {

constexpr int TI = 128;                    // UMMA M  (image rows)
constexpr int TJ = 256;                    // UMMA N  (image columns)
constexpr int KV = 32;                     // visibilities per operand stage = 64 fp16 = 128 B rows
constexpr int NSTAGE = 2;
constexpr int NVS = 4;
constexpr int A_BYTES = TI * 128;
constexpr int B_BYTES = TJ * 128;
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // A_hi | A_lo | B_hi | B_lo
constexpr int REC_BYTES = NVS * KV * 16;
constexpr int OFF_RECA = NSTAGE * STAGE_BYTES;
constexpr int OFF_RECB = OFF_RECA + REC_BYTES;
constexpr int OFF_BAR = OFF_RECB + REC_BYTES;
constexpr int NBAR = 2 * NSTAGE + 2 * NVS + 2;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16 + 1024;           // + alignment slack
constexpr int NTHREADS = 448;
constexpr int GEN_THREADS = 256;
constexpr uint32_t TMEM_COLS = 256;

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
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
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 (fp16 inputs, fp32 accumulate)
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
        "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
        "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
        "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
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
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "r"(addr));
  return r;
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"):
// rows of 128 B (64 fp16 along K), 8-row groups 1024 B apart (SBO), 16-byte chunk c of
// row r stored at chunk (c ^ (r & 7)). Bit layout: start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version [46,48), layout type [61,64) (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// Instruction descriptor, kind::f16: D fp32 (bit 4), A/B fp16 (0), both K-major,
// N>>3 at [17,23), M>>4 at [24,29).
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(TJ >> 3) << 17) | ((uint32_t)(TI >> 4) << 24);

// fractional turn (top 23 bits of a 0.32 fixed-point phase) -> angle - pi, in radians
__device__ __forceinline__ float phase_to_angle(uint32_t ph) {
  const float f = __uint_as_float(0x3F800000u | (ph >> 9));   // 1 + frac  in [1, 2)
  return fmaf(f, 6.283185307179586f, -9.42477796076938f);     // 2 pi frac - pi
}
// error-compensated fp16 split of (c, s): hi = rn(x), lo = rn(x - hi); packed (c | s << 16)
__device__ __forceinline__ void split2(float c, float s, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(c, s);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(c - hf.x, s - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ---------------------------------------------------------------------------
// Per evaluation, per visibility: amplitude and phase of w_k Vr_k, scaled by an exact
// power of two so that the largest amplitude sits near 2^14 (fp16 max is 65504).
__global__ void __launch_bounds__(256) k_grad_coeff(const float2* __restrict__ Vr,
                                                    const float* __restrict__ w, long Z,
                                                    const float* __restrict__ max_in,
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
  const float re = wk * vr.x, im = wk * vr.y;
  amp[k] = sqrtf(re * re + im * im) * scale;
  const float t = atan2f(im, re) * 0.15915494309189535f;       // turns in [-0.5, 0.5]
  gam[k] = (uint32_t)(int32_t)__float2int_rn(t * 2147483648.0f) << 1;
}

// ---------------------------------------------------------------------------
template <bool kUseW>
__global__ void __launch_bounds__(NTHREADS, 1) k_grad_umma(
    const uint64_t* __restrict__ du64, const uint64_t* __restrict__ dv64,
    const float* __restrict__ wz, const float* __restrict__ amp, const uint32_t* __restrict__ gam,
    const float* __restrict__ gA, const float* __restrict__ gB, long Z, int N, int x0, int y0,
    long klen, int chunk_stages, float* __restrict__ scratch) {
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
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + OFF_TMEM);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_j = (N + TJ - 1) / TJ;
  const int tj = blockIdx.x % tiles_j, ti = blockIdx.x / tiles_j;
  const int i0 = ti * TI, j0 = tj * TJ;
  const int ks = blockIdx.y;
  const long kbeg = ks * klen;
  const long kend = (kbeg + klen < Z) ? kbeg + klen : Z;
  const int nst = (int)((kend - kbeg + KV - 1) / KV);

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; s++) {
      mbar_init(BAR_OP_FULL(s), GEN_THREADS);
      mbar_init(BAR_OP_EMPTY(s), 1);
    }
    for (int s = 0; s < NVS; s++) {
      mbar_init(BAR_VIS_FULL(s), 32);
      mbar_init(BAR_VIS_EMPTY(s), GEN_THREADS);
    }
    mbar_init(BAR_ACC_FULL, 1);
    mbar_init(BAR_ACC_EMPTY, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 13) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     sbase + OFF_TMEM),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ================================================= operand generators
    const bool isA = warp < 4;
    const int r = tid & 127;                 // row within the 128-row group
    const int dr = r - 64;
    const uint32_t rowoff = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
    const uint32_t swz = (uint32_t)(r & 7);
    float g2a = 0.f, g2b = 0.f;              // 2 pi * g(row): w-term slope per wavelength of w
    if (kUseW) {
      if (isA) {
        const int gi = min(i0 + r, N - 1);
        g2a = 6.283185307179586f * gB[gi];
      } else {
        const int gj1 = min(j0 + r, N - 1), gj2 = min(j0 + r + 128, N - 1);
        g2a = 6.283185307179586f * gA[gj1];
        g2b = 6.283185307179586f * gA[gj2];
      }
    }
    const uint32_t rec0 = sbase + (isA ? OFF_RECA : OFF_RECB);
    for (int it = 0; it < nst; it++) {
      const int s = it % NSTAGE, vs = it % NVS;
      mbar_wait(BAR_VIS_FULL(vs), (uint32_t)((it / NVS) & 1));
      mbar_wait(BAR_OP_EMPTY(s), (uint32_t)(((it / NSTAGE) & 1) ^ 1));
      const uint32_t st0 = sbase + (uint32_t)s * STAGE_BYTES;
      const uint32_t recs = rec0 + (uint32_t)vs * (KV * 16);
      if (isA) {
        const uint32_t hi_row = st0 + rowoff, lo_row = st0 + A_BYTES + rowoff;
#pragma unroll 2
        for (int kq = 0; kq < KV / 4; kq++) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {
            const uint4 rec = ld_shared_v4(recs + (uint32_t)(kq * 4 + kk) * 16);
            const uint32_t ph = rec.x + (uint32_t)dr * rec.y;
            float a = phase_to_angle(ph);
            if (kUseW) a = fmaf(__uint_as_float(rec.z), g2a, a);
            split2(__cosf(a), __sinf(a), hi[kk], lo[kk]);
          }
          const uint32_t off = ((uint32_t)kq ^ swz) << 4;
          st_shared_v4(hi_row + off, hi[0], hi[1], hi[2], hi[3]);
          st_shared_v4(lo_row + off, lo[0], lo[1], lo[2], lo[3]);
        }
      } else {
        const uint32_t hi_row = st0 + 2 * A_BYTES + rowoff;
        const uint32_t lo_row = hi_row + B_BYTES;
#pragma unroll 2
        for (int kq = 0; kq < KV / 4; kq++) {
          uint32_t hi1[4], lo1[4], hi2[4], lo2[4];
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {
            const uint4 rec = ld_shared_v4(recs + (uint32_t)(kq * 4 + kk) * 16);
            const uint32_t ph1 = rec.x + (uint32_t)dr * rec.y;
            const uint32_t ph2 = ph1 + (rec.y << 7);
            const float am = __uint_as_float(rec.z);
            float a1 = phase_to_angle(ph1), a2 = phase_to_angle(ph2);
            if (kUseW) {
              const float wn = __uint_as_float(rec.w);
              a1 = fmaf(wn, g2a, a1);
              a2 = fmaf(wn, g2b, a2);
            }
            split2(am * __cosf(a1), am * __sinf(a1), hi1[kk], lo1[kk]);
            split2(am * __cosf(a2), am * __sinf(a2), hi2[kk], lo2[kk]);
          }
          const uint32_t off = ((uint32_t)kq ^ swz) << 4;
          st_shared_v4(hi_row + off, hi1[0], hi1[1], hi1[2], hi1[3]);
          st_shared_v4(lo_row + off, lo1[0], lo1[1], lo1[2], lo1[3]);
          st_shared_v4(hi_row + 16 * 1024 + off, hi2[0], hi2[1], hi2[2], hi2[3]);   // row + 128
          st_shared_v4(lo_row + 16 * 1024 + off, lo2[0], lo2[1], lo2[2], lo2[3]);
        }
      }
      fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(BAR_OP_FULL(s));
      mbar_arrive(BAR_VIS_EMPTY(vs));
    }
  } else if (warp < 12) {
    // ================================================= epilogue: TMEM -> scratch slice (+=)
    const int q = warp - 8;
    const int nchunks = (nst + chunk_stages - 1) / chunk_stages;
    const int gi = i0 + 32 * q + lane;
    float* orow = scratch + (size_t)ks * N * N + (size_t)gi * N + j0;
    for (int c = 0; c < nchunks; c++) {
      mbar_wait(BAR_ACC_FULL, (uint32_t)(c & 1));
      tc_fence_after();
#pragma unroll 1
      for (int cb = 0; cb < TJ / 32; cb++) {
        uint32_t v[32];
        tc_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(cb * 32), v);
        tc_wait_ld();
        if (gi < N) {
#pragma unroll
          for (int g = 0; g < 8; g++) {
            const int j = j0 + cb * 32 + g * 4;
            if (j < N) {   // N % 4 == 0 (checked on the host)
              float4* p = reinterpret_cast<float4*>(orow + cb * 32 + g * 4);
              float4 o = make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]),
                                     __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
              if (c > 0) {
                const float4 prev = *p;
                o.x += prev.x; o.y += prev.y; o.z += prev.z; o.w += prev.w;
              }
              *p = o;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(BAR_ACC_EMPTY);
    }
  } else if (warp == 12) {
    // ================================================= record producer (lane = visibility)
    const int ic = i0 + 64, jc = j0 + 64;   // rows are addressed as centre + dr, |dr| <= 64 (+128)
    for (int it = 0; it < nst; it++) {
      const int vs = it % NVS;
      mbar_wait(BAR_VIS_EMPTY(vs), (uint32_t)(((it / NVS) & 1) ^ 1));
      const long k = kbeg + (long)it * KV + lane;
      uint4 ra = make_uint4(0u, 0u, 0u, 0u), rb = make_uint4(0u, 0u, 0u, 0u);
      if (k < kend) {
        const uint64_t du = __ldg(&du64[k]), dv = __ldg(&dv64[k]);
        const float wzk = kUseW ? __ldg(&wz[k]) : 0.f;
        // A: +phase of v_k y_i ; B: arg(Vr_k) - phase of u_k x_j (and -w for the w-term)
        ra.x = (uint32_t)((dv * (uint64_t)(int64_t)(ic - y0)) >> 32);
        ra.y = (uint32_t)(dv >> 32);
        ra.z = __float_as_uint(wzk);
        const uint32_t pu = (uint32_t)((du * (uint64_t)(int64_t)(jc - x0)) >> 32);
        rb.x = __ldg(&gam[k]) - pu;
        rb.y = 0u - (uint32_t)(du >> 32);
        rb.z = __float_as_uint(__ldg(&amp[k]));
        rb.w = __float_as_uint(-wzk);
      }
      *reinterpret_cast<uint4*>(sgen + OFF_RECA + (vs * KV + lane) * 16) = ra;
      *reinterpret_cast<uint4*>(sgen + OFF_RECB + (vs * KV + lane) * 16) = rb;
      mbar_arrive(BAR_VIS_FULL(vs));
    }
  } else if (lane == 0) {
    // ================================================= MMA issuer (one thread)
    for (int it = 0; it < nst; it++) {
      const int s = it % NSTAGE;
      const int cpos = it % chunk_stages;
      if (cpos == 0) {
        mbar_wait(BAR_ACC_EMPTY, (uint32_t)(((it / chunk_stages) & 1) ^ 1));
        tc_fence_after();
      }
      mbar_wait(BAR_OP_FULL(s), (uint32_t)((it / NSTAGE) & 1));
      tc_fence_after();
      const uint32_t st0 = sbase + (uint32_t)s * STAGE_BYTES;
      const uint32_t a_hi = st0, a_lo = st0 + A_BYTES, b_hi = st0 + 2 * A_BYTES,
                     b_lo = st0 + 2 * A_BYTES + B_BYTES;
#pragma unroll
      for (int kstep = 0; kstep < 4; kstep++) {   // 4 x K=16 fp16 = 32 B steps inside the 128 B row
        const uint32_t ko = (uint32_t)kstep * 32;
        tc_mma_f16(tmem_base, umma_desc_sw128(a_hi + ko), umma_desc_sw128(b_hi + ko), kIdesc,
                   (cpos > 0 || kstep > 0) ? 1u : 0u);
        tc_mma_f16(tmem_base, umma_desc_sw128(a_hi + ko), umma_desc_sw128(b_lo + ko), kIdesc, 1u);
        tc_mma_f16(tmem_base, umma_desc_sw128(a_lo + ko), umma_desc_sw128(b_hi + ko), kIdesc, 1u);
      }
      tc_commit(BAR_OP_EMPTY(s));                  // smem stage reusable once these MMAs retire
      if (cpos == chunk_stages - 1 || it == nst - 1) tc_commit(BAR_ACC_FULL);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(TMEM_COLS)
                 : "memory");
  }
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
  if (e->cfg.N % 4 != 0) return false;
  if (wterm_turns(e, c) > 4.0) return false;
  return true;
}

int gvm_grad_umma(gvm_engine* e, GvmChannel& c, int* ksplit_out) {
  const int N = (int)e->cfg.N;
  if (!gvm_grad_umma_supported(e, c)) {
    gvm_set_error("gvm_grad_umma: unsupported problem (N %% 4 != 0 or w-term beyond 4 turns)");
    return 1;
  }
  static bool attr_set = false;
  if (!attr_set) {
    GVM_CUDA(cudaFuncSetAttribute(k_grad_umma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    GVM_CUDA(cudaFuncSetAttribute(k_grad_umma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  if (!c.amp) {
    const size_t z = (size_t)(c.Z > 0 ? c.Z : 1);
    GVM_CUDA(cudaMalloc(&c.amp, z * sizeof(float)));
    GVM_CUDA(cudaMalloc(&c.gam, z * sizeof(uint32_t)));
  }
  float* inv_scale = e->red_max + e->red_slots + c.slot;
  k_grad_coeff<<<(int)((c.Z + 255) / 256), 256, 0, e->stream>>>(c.Vr, c.w, c.Z, e->red_max + c.slot,
                                                                 inv_scale, c.amp, c.gam);
  GVM_LAUNCH(e);
  const bool use_w = c.max_abs_wz > 0.f;
  if (use_w)
    if (gvm_build_pixtab(e, c)) return 1;

  // visibilities per TMEM accumulation chunk (fp32 accumulation length control)
  long chunk = 8192;
  if (const char* s = getenv("GVM_UMMA_CHUNK")) chunk = atol(s);
  if (chunk < KV) chunk = KV;
  chunk = (chunk / KV) * KV;

  if (const char* s = getenv("GVM_UMMA_KERNEL"))
    if (atoi(s) == 2) return gvm_grad_umma2_launch(e, c, use_w, chunk, ksplit_out);

  const int tiles = ((N + TJ - 1) / TJ) * ((N + TI - 1) / TI);
  // split K so that tiles*ksplit fills whole waves of one-CTA-per-SM, slices >= 2048 samples
  long max_ks = c.Z / 2048;
  if (max_ks < 1) max_ks = 1;
  while (max_ks > 1 && (size_t)max_ks * N * N * sizeof(float) > ((size_t)2 << 30)) max_ks--;
  if (max_ks > 4096) max_ks = 4096;
  int best = 1;
  double best_eff = -1.0;
  for (long ks = 1; ks <= max_ks; ks++) {
    const long ctas = (long)tiles * ks;
    const long waves = (ctas + e->sm_count - 1) / e->sm_count;
    double eff = (double)ctas / (double)(waves * e->sm_count);
    if (waves < 2 && ks < max_ks) eff *= 0.5 + 0.25 * waves;   // prefer >= 2 waves when possible
    if (eff > best_eff + 1e-9) { best_eff = eff; best = (int)ks; }
    if (ctas >= 8L * e->sm_count && eff > 0.97) break;
  }
  long klen = (c.Z + best - 1) / best;
  klen = ((klen + KV - 1) / KV) * KV;
  int ksplit = (int)((c.Z + klen - 1) / klen);
  if (ksplit < 1) ksplit = 1;
  if (gvm_ensure_grad_scratch(e, (size_t)ksplit * N * N)) return 1;
  const int x0 = (int)c.d.phs_xobs_pix, y0 = (int)c.d.phs_yobs_pix;
  dim3 grid(tiles, ksplit);
  gvm_ev_begin(e);
  if (use_w)
    k_grad_umma<true><<<grid, NTHREADS, SMEM_BYTES, e->stream>>>(
        c.du64, c.dv64, c.wz, c.amp, c.gam, e->pixtab, e->pixtab + N, c.Z, N, x0, y0, klen,
        (int)(chunk / KV), e->grad_scratch);
  else
    k_grad_umma<false><<<grid, NTHREADS, SMEM_BYTES, e->stream>>>(
        c.du64, c.dv64, c.wz, c.amp, c.gam, e->pixtab, e->pixtab + N, c.Z, N, x0, y0, klen,
        (int)(chunk / KV), e->grad_scratch);
  gvm_ev_end(e);
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  *ksplit_out = ksplit;
  return 0;
}
