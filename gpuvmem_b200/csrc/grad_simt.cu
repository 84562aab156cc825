// grad_simt.cu — chi2 gradient on CUDA cores, plus the pieces every gradient
// path shares (pixel w-term tables, split-K scratch, the finishing pass).
//
// Reference: DChi2 (src/functions.cu:3698-3791 / 3793-3888) evaluates, per
// unmasked pixel, d = sum_k w_k (Vr.re cos(2 pi phi) + Vr.im sin(2 pi phi)),
// phi = x u + y v + (z-1) w, with one sincospif per (pixel, visibility) pair, then
// DChi2_total_I_nu_0 / DChi2_total_alpha (:4000 / :3968) apply the MFS chain rule.
//
// Here the sum is recast as a contraction:
//   d[i,j] = Re sum_k P_k(j) Q_k(i),
//   P_k(j) = c_k exp(2 pi i (u_k x_j + w_k gA(x_j))),  c_k = w_k conj(Vr_k)
//   Q_k(i) =     exp(2 pi i (v_k y_i + w_k gB(y_i))),
// with gA(x) = sqrt(1-x^2)-1, gB(y) = sqrt(1-y^2)-1.  The only non-separable
// piece of (z-1) is the cross term ~ -x^2 y^2 / 4; the engine bounds
// max|w| * max|cross| and uses k_grad_exact (per-pair phase) when it matters.
// Phases u_k x_j are exact integer arithmetic on 0.64 fixed-point turns.
//   k_grad_sep    64x64 output tile per CTA, 4x4 outputs per thread, K chunks of
//                 32 visibilities staged as P/Q phasors in shared memory: 2 FMA
//                 per pair instead of a sincos per pair. FP32-pipe bound.
//   k_grad_exact  same tiling, phase per pair (fallback / cross-check).
// Split-K partial sums go to scratch[ks][M*N] (each CTA owns its tile: no
// atomics, deterministic); k_grad_finish adds them in a fixed order.
#include "gvm_internal.cuh"

namespace {

constexpr int TJ = 64, TI = 64, KC = 32;

__device__ __forceinline__ float turns_from_fixed(uint64_t d64, int off) {
  // top 32 bits of d64 * off (mod 2^64) = fractional turn of off * delta
  const uint64_t prod = d64 * (uint64_t)(int64_t)off;
  const int32_t ph = (int32_t)(uint32_t)(prod >> 32);
  return (float)ph * 2.3283064365386963e-10f;  // 2^-32 -> [-0.5, 0.5)
}

template <bool kUseW>
__global__ void __launch_bounds__(256) k_grad_sep(
    const uint64_t* __restrict__ du64, const uint64_t* __restrict__ dv64,
    const float* __restrict__ wz, const float2* __restrict__ Vr, const float* __restrict__ w,
    const float* __restrict__ gA, const float* __restrict__ gB, long Z, int N, int x0, int y0,
    long klen, float im_sign, float* __restrict__ scratch) {
  __shared__ __align__(16) float2 sP[KC][TJ];
  __shared__ __align__(16) float2 sQ[KC][TI];
  const int tiles_j = (N + TJ - 1) / TJ;
  const int tj = blockIdx.x % tiles_j, ti = blockIdx.x / tiles_j;
  const int ks = blockIdx.y;
  const long kbeg = ks * klen;
  const long kend = (kbeg + klen < Z) ? kbeg + klen : Z;
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int gl = t & 63;   // generated row within the tile
  const int gk = t >> 6;   // 0..3
  const int jg = tj * TJ + gl, ig = ti * TI + gl;
  const float gAj = (kUseW && jg < N) ? gA[jg] : 0.f;
  const float gBi = (kUseW && ig < N) ? gB[ig] : 0.f;

  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = 0.f;

  for (long k0 = kbeg; k0 < kend; k0 += KC) {
#pragma unroll
    for (int m = 0; m < KC / 4; m++) {
      const int kl = gk + 4 * m;
      const long k = k0 + kl;
      float2 p = make_float2(0.f, 0.f), q = make_float2(0.f, 0.f);
      if (k < kend) {
        const float wk = __ldg(&w[k]);
        const float2 vr = __ldg(&Vr[k]);
        const float cr = wk * vr.x, ci = -im_sign * wk * vr.y;   // im_sign = -1: w Vr (error maps)
        float tu = turns_from_fixed(__ldg(&du64[k]), jg - x0);
        float tv = turns_from_fixed(__ldg(&dv64[k]), ig - y0);
        if (kUseW) {
          const float wzk = __ldg(&wz[k]);
          tu = fmaf(wzk, gAj, tu);
          tv = fmaf(wzk, gBi, tv);
        }
        float s, c;
        sincospif(2.0f * tu, &s, &c);
        p = make_float2(cr * c - ci * s, -(cr * s + ci * c));  // (Re P, -Im P)
        sincospif(2.0f * tv, &s, &c);
        q = make_float2(c, s);
      }
      sP[kl][gl] = p;
      sQ[kl][gl] = q;
    }
    __syncthreads();
#pragma unroll 8
    for (int kl = 0; kl < KC; kl++) {
      const float4 p01 = *reinterpret_cast<const float4*>(&sP[kl][tx * 4]);
      const float4 p23 = *reinterpret_cast<const float4*>(&sP[kl][tx * 4 + 2]);
      const float4 q01 = *reinterpret_cast<const float4*>(&sQ[kl][ty * 4]);
      const float4 q23 = *reinterpret_cast<const float4*>(&sQ[kl][ty * 4 + 2]);
      const float pr[4] = {p01.x, p01.z, p23.x, p23.z};
      const float pi[4] = {p01.y, p01.w, p23.y, p23.w};
      const float qr[4] = {q01.x, q01.z, q23.x, q23.z};
      const float qi[4] = {q01.y, q01.w, q23.y, q23.w};
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
          acc[a][b] = fmaf(pr[b], qr[a], acc[a][b]);
          acc[a][b] = fmaf(pi[b], qi[a], acc[a][b]);
        }
    }
    __syncthreads();
  }
  float* out = scratch + (size_t)ks * N * N;
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const int i = ti * TI + ty * 4 + a;
    if (i >= N) continue;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int j = tj * TJ + tx * 4 + b;
      if (j < N) out[(size_t)i * N + j] = acc[a][b];
    }
  }
}

// Per-pair phase including the full (z-1) w term: the reference formula, with the
// u x + v y part still exact fixed-point.
__global__ void __launch_bounds__(256) k_grad_exact(
    const uint64_t* __restrict__ du64, const uint64_t* __restrict__ dv64,
    const float* __restrict__ wz, const float2* __restrict__ Vr, const float* __restrict__ w,
    long Z, int N, int x0, int y0, double dx_rad, double dy_rad, long klen,
    float* __restrict__ scratch) {
  __shared__ uint32_t sU[KC][TJ];
  __shared__ uint32_t sV[KC][TI];
  __shared__ float2 sC[KC];
  __shared__ float sW[KC];
  const int tiles_j = (N + TJ - 1) / TJ;
  const int tj = blockIdx.x % tiles_j, ti = blockIdx.x / tiles_j;
  const int ks = blockIdx.y;
  const long kbeg = ks * klen;
  const long kend = (kbeg + klen < Z) ? kbeg + klen : Z;
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int gl = t & 63, gk = t >> 6;
  const int jg = tj * TJ + gl, ig = ti * TI + gl;

  float acc[4][4], zm1[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) {
      acc[a][b] = 0.f;
      const double x = (tj * TJ + tx * 4 + b - x0) * dx_rad;
      const double y = (ti * TI + ty * 4 + a - y0) * dy_rad;
      const double r2 = x * x + y * y;
      zm1[a][b] = (float)(-r2 / (1.0 + sqrt(1.0 - r2)));
    }

  for (long k0 = kbeg; k0 < kend; k0 += KC) {
#pragma unroll
    for (int m = 0; m < KC / 4; m++) {
      const int kl = gk + 4 * m;
      const long k = k0 + kl;
      uint32_t pu = 0, pv = 0;
      if (k < kend) {
        pu = (uint32_t)((__ldg(&du64[k]) * (uint64_t)(int64_t)(jg - x0)) >> 32);
        pv = (uint32_t)((__ldg(&dv64[k]) * (uint64_t)(int64_t)(ig - y0)) >> 32);
      }
      sU[kl][gl] = pu;
      sV[kl][gl] = pv;
    }
    if (t < KC) {
      const long k = k0 + t;
      float2 c = make_float2(0.f, 0.f);
      float wzk = 0.f;
      if (k < kend) {
        const float wk = __ldg(&w[k]);
        const float2 vr = __ldg(&Vr[k]);
        c = make_float2(wk * vr.x, wk * vr.y);
        wzk = __ldg(&wz[k]);
      }
      sC[t] = c;
      sW[t] = wzk;
    }
    __syncthreads();
    for (int kl = 0; kl < KC; kl++) {
      const float2 c = sC[kl];
      const float wzk = sW[kl];
#pragma unroll
      for (int a = 0; a < 4; a++) {
        const uint32_t pv = sV[kl][ty * 4 + a];
#pragma unroll
        for (int b = 0; b < 4; b++) {
          const uint32_t ph = sU[kl][tx * 4 + b] + pv;
          float tt = (float)(int32_t)ph * 2.3283064365386963e-10f;
          tt = fmaf(wzk, zm1[a][b], tt);
          float s, cs;
          sincospif(2.0f * tt, &s, &cs);
          acc[a][b] += c.x * cs + c.y * s;
        }
      }
    }
    __syncthreads();
  }
  float* out = scratch + (size_t)ks * N * N;
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const int i = ti * TI + ty * 4 + a;
    if (i >= N) continue;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int j = tj * TJ + tx * 4 + b;
      if (j < N) out[(size_t)i * N + j] = acc[a][b];
    }
  }
}

// gA(x_j) = sqrt(1-x^2)-1 and gB(y_i), in the cancellation-free form, as floats.
__global__ void k_pixtab(float* __restrict__ tab, int N, int x0, int y0, double dx_rad,
                         double dy_rad) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const double x = (n - x0) * dx_rad, y = (n - y0) * dy_rad;
  tab[n] = (float)(-(x * x) / (1.0 + sqrt(1.0 - x * x)));
  tab[N + n] = (float)(-(y * y) / (1.0 + sqrt(1.0 - y * y)));
}

// Sum the split-K slices in a fixed order, then gvm_finish_pixel (scale, chain rule, +=).
__global__ void __launch_bounds__(256) k_grad_finish(const float* __restrict__ scratch, int ksplit,
                                                     const float* __restrict__ noise,
                                                     float noise_cut, GvmFinishParams p) {
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long MN = p.M * p.N;
  if (idx >= MN) return;
  if (noise[idx] >= noise_cut) {  // DChi2 returns early; device_dchi2 was memset to 0
    if (p.dchi2_out) p.dchi2_out[idx] = 0.0f;
    return;
  }
  float d = 0.0f;
  for (int s = 0; s < ksplit; s++) d += scratch[(size_t)s * MN + idx];
  gvm_finish_pixel(p, d, idx, (int)(idx / p.N), (int)(idx % p.N));
}

// Four pixels per thread with 128-bit loads (cached beam plane, not the raw variant, M N a multiple of 4): the mask is
// read first and fully masked quads — most of a large field — cost nothing else. Per pixel the arithmetic is
// gvm_finish_pixel's.
__global__ void __launch_bounds__(256) k_grad_finish4(const float* __restrict__ scratch, int ksplit,
                                                      const float* __restrict__ noise, float noise_cut, GvmFinishParams p) {
  const long MN = p.M * p.N, nq = MN / 4;
  // two quads per thread, both mask loads in flight before anything depends on them
  const long q0 = 2 * (blockIdx.x * (long)blockDim.x + threadIdx.x);
  if (q0 >= nq) return;
  const bool two = q0 + 1 < nq;
  const float4 nz0 = __ldg(reinterpret_cast<const float4*>(noise) + q0);
  const float4 nz1 = two ? __ldg(reinterpret_cast<const float4*>(noise) + q0 + 1) : make_float4(noise_cut, noise_cut, noise_cut, noise_cut);
  const bool odd = p.flag_opt % 2 != 0;
  const float nudiv = p.freq / p.nu_0;
  const float lognu = logf(nudiv);
  auto one = [&](float dv, float atten, float g, float i0, float alpha, float res, bool masked) -> float {
    if (masked) return res;
    float scale_factor = p.fg_scale * atten;
    if (p.gcf) scale_factor = scale_factor * g;
    dv *= scale_factor;
    if (p.normalize) dv /= p.Z;
    const float dchi2 = -dv;
    const float dI = powf(nudiv, alpha);
    if (!odd) return res + dchi2 * dI;
    const float dalpha = i0 * dI * p.fg_scale * lognu;
    return i0 > p.threshold ? res + dchi2 * dalpha : res;
  };
  auto quad = [&](long q, const float4& nz) {
    const bool m0 = nz.x >= noise_cut, m1 = nz.y >= noise_cut, m2 = nz.z >= noise_cut, m3 = nz.w >= noise_cut;
    if (m0 && m1 && m2 && m3) return;
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < ksplit; s++) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(scratch + (size_t)s * MN) + q);
      d.x += t.x; d.y += t.y; d.z += t.z; d.w += t.w;
    }
    const float4 at = __ldg(reinterpret_cast<const float4*>(p.atten) + q);
    float4 gc = make_float4(1.f, 1.f, 1.f, 1.f);
    if (p.gcf) gc = __ldg(reinterpret_cast<const float4*>(p.gcf) + q);
    const float4 I0 = __ldg(reinterpret_cast<const float4*>(p.I) + q);
    const float4 al = __ldg(reinterpret_cast<const float4*>(p.I + MN) + q);
    float4* rp = reinterpret_cast<float4*>(p.result + (odd ? MN : 0)) + q;
    float4 r = *rp;
    r.x = one(d.x, at.x, gc.x, I0.x, al.x, r.x, m0);
    r.y = one(d.y, at.y, gc.y, I0.y, al.y, r.y, m1);
    r.z = one(d.z, at.z, gc.z, I0.z, al.z, r.z, m2);
    r.w = one(d.w, at.w, gc.w, I0.w, al.w, r.w, m3);
    *rp = r;
  };
  quad(q0, nz0);
  if (two) quad(q0 + 1, nz1);
}

}  // namespace

int gvm_ensure_grad_scratch(gvm_engine* e, size_t floats) {
  if (floats <= e->grad_scratch_floats) return 0;
  if (e->grad_scratch) cudaFree(e->grad_scratch);
  e->grad_scratch = nullptr;
  e->grad_scratch_floats = 0;
  GVM_CUDA(cudaMalloc(&e->grad_scratch, floats * sizeof(float)));
  e->grad_scratch_floats = floats;
  return 0;
}

int gvm_build_pixtab(gvm_engine* e, const GvmChannel& c) {
  const int N = (int)e->cfg.N;
  k_pixtab<<<(N + 255) / 256, 256, 0, e->stream>>>(e->pixtab, N, (int)c.d.phs_xobs_pix,
                                                   (int)c.d.phs_yobs_pix,
                                                   e->cfg.DELTAX * GVM_RPDEG_D,
                                                   e->cfg.DELTAY * GVM_RPDEG_D);
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  return 0;
}

// max_k |w_k| * max_pixels |(z-1) - gA(x) - gB(y)| in turns: the phase error of
// dropping the non-separable cross term. Evaluated at the four image corners,
// where it is largest.
double gvm_wterm_cross_bound(const gvm_engine* e, const GvmChannel& c) {
  const double dx = e->cfg.DELTAX * GVM_RPDEG_D, dy = e->cfg.DELTAY * GVM_RPDEG_D;
  const int x0 = (int)c.d.phs_xobs_pix, y0 = (int)c.d.phs_yobs_pix;
  const int N = (int)e->cfg.N;
  double worst = 0.0;
  const int js[2] = {0, N - 1}, is[2] = {0, N - 1};
  for (int a = 0; a < 2; a++)
    for (int b = 0; b < 2; b++) {
      const long double x = (long double)((js[a] - x0) * dx), y = (long double)((is[b] - y0) * dy);
      const long double r2 = x * x + y * y;
      if (r2 >= 1.0L) return 1e30;
      const long double zm1 = -r2 / (1.0L + sqrtl(1.0L - r2));
      const long double ga = -(x * x) / (1.0L + sqrtl(1.0L - x * x));
      const long double gb = -(y * y) / (1.0L + sqrtl(1.0L - y * y));
      const double cross = (double)fabsl(zm1 - ga - gb);
      if (cross > worst) worst = cross;
    }
  return worst * (double)c.max_abs_wz;
}

static int pick_ksplit(const gvm_engine* e, int tiles, long Z, int kc) {
  // enough CTAs for ~4 waves on all SMs, but never chunks shorter than 8*kc
  long want = ((long)e->sm_count * 8 + tiles - 1) / tiles;
  long maxs = Z / (8L * kc);
  if (maxs < 1) maxs = 1;
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  if (want > 65535) want = 65535;
  return (int)want;
}

int gvm_grad_simt(gvm_engine* e, GvmChannel& c, bool exact, int* ksplit_out) {
  const int N = (int)e->cfg.N;
  const int tiles = ((N + TJ - 1) / TJ) * ((N + TI - 1) / TI);
  int ksplit = pick_ksplit(e, tiles, c.Z, KC);
  // keep the scratch below ~2 GiB
  while (ksplit > 1 && (size_t)ksplit * N * N * sizeof(float) > ((size_t)2 << 30)) ksplit--;
  long klen = (c.Z + ksplit - 1) / ksplit;
  klen = ((klen + KC - 1) / KC) * KC;
  ksplit = (int)((c.Z + klen - 1) / klen);
  if (ksplit < 1) ksplit = 1;
  if (gvm_ensure_grad_scratch(e, (size_t)ksplit * N * N)) return 1;
  const int x0 = (int)c.d.phs_xobs_pix, y0 = (int)c.d.phs_yobs_pix;
  dim3 grid(tiles, ksplit);
  const bool use_w = c.max_abs_wz > 0.f && !e->err_variant;
  if (e->err_variant) exact = false;   // no w-term: the separable form is the formula itself
  const float im_sign = e->err_variant ? -1.0f : 1.0f;
  if (!exact && use_w)
    if (gvm_build_pixtab(e, c)) return 1;
  gvm_ev_begin(e);
  if (exact) {
    k_grad_exact<<<grid, 256, 0, e->stream>>>(c.du64, c.dv64, c.wz, c.Vr, c.w, c.Z, N, x0, y0,
                                              e->cfg.DELTAX * GVM_RPDEG_D,
                                              e->cfg.DELTAY * GVM_RPDEG_D, klen, e->grad_scratch);
  } else if (use_w) {
    k_grad_sep<true><<<grid, 256, 0, e->stream>>>(c.du64, c.dv64, c.wz, c.Vr, c.w, e->pixtab,
                                                  e->pixtab + N, c.Z, N, x0, y0, klen,
                                                  im_sign, e->grad_scratch);
  } else {
    k_grad_sep<false><<<grid, 256, 0, e->stream>>>(c.du64, c.dv64, c.wz, c.Vr, c.w, e->pixtab,
                                                   e->pixtab + N, c.Z, N, x0, y0, klen,
                                                   im_sign, e->grad_scratch);
  }
  gvm_ev_end(e);
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  *ksplit_out = ksplit;
  return 0;
}

GvmFinishParams gvm_finish_params(gvm_engine* e, const GvmChannel& c, const float* I_dev, int flag_opt,
                                  int normalize, float* result_dev) {
  const gvm_config& g = e->cfg;
  GvmFinishParams p;
  p.atten = gvm_channel_atten(e, const_cast<GvmChannel&>(c)); p.gcf = e->gcf; p.I = I_dev; p.result = result_dev;
  p.dchi2_out = e->err_variant ? e->dchi2 : nullptr;   // the per-channel plane is only read by the error maps
  p.N = g.N; p.M = g.M; p.Z = (long)c.Znorm;
  p.fg_scale = g.fg_scale; p.D = c.d.antenna_diameter; p.pb_factor = c.d.pb_factor;
  p.pb_cutoff = c.d.pb_cutoff; p.freq = c.d.freq; p.xobs = c.d.ref_xobs_pix; p.yobs = c.d.ref_yobs_pix;
  p.nu_0 = g.nu_0; p.threshold = g.threshold; p.DELTAX = g.DELTAX; p.DELTAY = g.DELTAY;
  p.primary_beam = c.d.primary_beam; p.flag_opt = flag_opt; p.normalize = normalize;
  p.raw = e->err_variant;
  return p;
}

int gvm_grad_finish(gvm_engine* e, GvmChannel& c, const float* I_dev, int ksplit, int flag_opt,
                    int normalize, float* result_dev) {
  const long MN = e->cfg.M * e->cfg.N;
  const GvmFinishParams fp = gvm_finish_params(e, c, I_dev, flag_opt, normalize, result_dev);
  const bool vec4 = !fp.raw && fp.atten && MN % 4 == 0 && (((uintptr_t)I_dev | (uintptr_t)result_dev) & 15) == 0;
  if (vec4)
    k_grad_finish4<<<(int)((MN / 8 + 255) / 256), 256, 0, e->stream>>>(e->grad_scratch, ksplit, e->noise, e->cfg.noise_cut, fp);
  else
    k_grad_finish<<<(int)((MN + 255) / 256), 256, 0, e->stream>>>(e->grad_scratch, ksplit, e->noise, e->cfg.noise_cut, fp);
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  return 0;
}
