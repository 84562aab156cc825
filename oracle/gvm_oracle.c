/* oracle/gvm_oracle.c — CPU restatement of gpuvmem's objective/gradient hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product (gpuvmem_b200/, include/) may
 * link, import or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do, and only as the checker.
 *
 * Each function cites the reference lines it follows (paths relative to the
 * reference tree). Discrete decisions (cell indices, masks, clips, weights, the
 * gridding accumulators) use exactly the reference's mixed fp32/fp64 types so
 * they can be compared bit-for-bit; the continuous sums (FFT, chi2, the DFT
 * gradient) are evaluated in fp64 so that they measure how far BOTH the
 * reference's fp32 CUDA kernels and the new engine are from the exact value.
 *
 * Pinning: the reference ships no golden vectors. The host-side
 * pieces restated here (weighting schemes, do_gridding, CKernel tables) are
 * checked bit-for-bit against the reference's own host code compiled unmodified
 * into oracle/_ref/libgvref.so (tests/test_oracle_vs_reference_cpu.py, runs
 * without a GPU). The CUDA-only pieces (forward model, chi2, DChi2, priors) are
 * checked against the reference's own kernels from the same library on the GPU
 * box (tests/test_parity_reference_gpu.py) and against fixtures produced by
 * that library (tests/golden/).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GVO_PI_F 3.14159265358979323846f /* CUDART_PI_F rounds to the same float */
#define GVO_PI_D 3.14159265358979323846
#define GVO_RPDEG_D (GVO_PI_D / 180.0)   /* include/functions.cuh:18 */
#define GVO_LIGHTSPEED 2.99792458E8f      /* include/MSFITSIO.cuh:54 */
static const float GVO_RZ = 1.2196698912665045; /* include/functions.cuh:23 */

/* src/MSFITSIO.cu:36-45 */
float gvo_freq_to_wavelength(float freq) { return GVO_LIGHTSPEED / freq; }
double gvo_metres_to_lambda(double m, float freq) {
  float lambda = gvo_freq_to_wavelength(freq);
  return m / lambda;
}
/* src/MSFITSIO.cu:47-51 */
static float gvo_distance(float x, float y, float x0, float y0) {
  float sumsqr = (x - x0) * (x - x0) + (y - y0) * (y - y0);
  return sqrtf(sumsqr);
}

int gvo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void gvo_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------ upload --
 * hermitianSymmetry (src/functions.cu:2256-2273) + the static part of vis_mod
 * (src/functions.cu:2569-2586, 2607). cell[2k] = i1 (u axis), cell[2k+1] = j1,
 * -1/-1 and weight 0 when outside the grid. frac = (du, dv) in fp64. */
void gvo_prep(long Z, const double* uvw_m, const float* Vo_in, const float* w_in, float freq,
              double deltau, double deltav, long N, double* uvw_l, int32_t* cell, double* frac,
              float* Vo, float* w) {
  for (long k = 0; k < Z; k++) {
    double um = uvw_m[3 * k], vm = uvw_m[3 * k + 1], wm = uvw_m[3 * k + 2];
    float vr = Vo_in[2 * k], vi = Vo_in[2 * k + 1];
    if (um > 0.0) {
      um *= -1.0;
      vm *= -1.0;
      vi *= -1.0f;
    }
    double u = gvo_metres_to_lambda(um, freq);
    double v = gvo_metres_to_lambda(vm, freq);
    double wl = gvo_metres_to_lambda(wm, freq);
    uvw_l[3 * k] = u; uvw_l[3 * k + 1] = v; uvw_l[3 * k + 2] = wl;
    Vo[2 * k] = vr; Vo[2 * k + 1] = vi;
    double uv_x = u / deltau, uv_y = v / deltav;
    if (uv_x < 0.0) uv_x += N;
    if (uv_y < 0.0) uv_y += N;
    int i1 = (int)floor(uv_x), j1 = (int)floor(uv_y);
    if (i1 >= 0 && i1 < N && j1 >= 0 && j1 < N) {
      cell[2 * k] = i1; cell[2 * k + 1] = j1;
      frac[2 * k] = uv_x - i1; frac[2 * k + 1] = uv_y - j1;
      w[k] = w_in[k];
    } else {
      cell[2 * k] = -1; cell[2 * k + 1] = -1;
      frac[2 * k] = 0.0; frac[2 * k + 1] = 0.0;
      w[k] = 0.0f;
    }
  }
}

/* attenuation(): src/functions.cu:2304-2333, AiryDiskBeam :2275, GaussianBeam :2293 */
float gvo_attenuation(int i, int j, float D, float pb_factor, float pb_cutoff, float freq,
                      float xobs, float yobs, double DELTAX, double DELTAY, int primary_beam) {
  int x0 = (int)xobs, y0 = (int)yobs;
  float x = (float)((j - x0) * DELTAX * GVO_RPDEG_D);
  float y = (float)((i - y0) * DELTAY * GVO_RPDEG_D);
  float arc = gvo_distance(x, y, 0.0f, 0.0f);
  float lambda = gvo_freq_to_wavelength(freq);
  float atten;
  if (primary_beam == 0) {
    atten = 1.0f;
    if (arc != 0.0f) {
      float arg = GVO_PI_F * arc * D / lambda * (GVO_RZ / pb_factor);
      float b = (float)j1((double)arg);
      atten = 4.0f * (b / arg) * (b / arg);
    }
  } else {
    float fwhm = pb_factor * lambda / D;
    float c = 4.0f * logf(2.0f);
    float r = arc / fwhm;
    atten = expf(-c * r * r);
  }
  return (arc <= pb_cutoff) ? atten : 0.0f;
}

/* clip2IWNoise: src/functions.cu:2694-2719 (mutates I) */
void gvo_clip(float* I, const float* noise, long M, long N, float noise_cut, float minpix,
              float eta, float threshold, int schedule) {
  for (long idx = 0; idx < M * N; idx++) {
    if (noise[idx] > noise_cut) {
      I[idx] = (eta > 0.0f) ? 0.0f : (float)(-1.0 * eta * minpix);
      I[M * N + idx] = 0.0f;
    } else if (I[idx] < threshold && schedule > 0) {
      I[M * N + idx] = 0.0f;
    }
  }
}

/* in-place radix-2 FFT, sign = +1 is cuFFT's CUFFT_INVERSE (unnormalised) */
static void fft1d(double* re, double* im, long n, long stride, int sign) {
  for (long i = 1, j = 0; i < n; i++) {
    long bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) {
      double t = re[i * stride]; re[i * stride] = re[j * stride]; re[j * stride] = t;
      t = im[i * stride]; im[i * stride] = im[j * stride]; im[j * stride] = t;
    }
  }
  for (long len = 2; len <= n; len <<= 1) {
    double ang = sign * 2.0 * GVO_PI_D / (double)len;
    for (long i = 0; i < n; i += len)
      for (long k = 0; k < len / 2; k++) {
        double wr = cos(ang * k), wi = sin(ang * k);
        long a = (i + k) * stride, b = (i + k + len / 2) * stride;
        double xr = re[b] * wr - im[b] * wi, xi = re[b] * wi + im[b] * wr;
        re[b] = re[a] - xr; im[b] = im[a] - xi;
        re[a] += xr; im[a] += xi;
      }
  }
}

/* The grid the interpolator reads: calculateInu (src/functions.cu:3939-3966),
 * apply_beam2I (:2424-2444), apply_GCF (:2468), cuFFT inverse (:2165),
 * phase_rotate (:2483-2518). N must be a power of two. Vre/Vim: N*N doubles. */
int gvo_model_grid(const float* I, const float* gcf, long N, float nu, float nu_0, float minpix,
                   float eta, float fg_scale, float D, float pb_factor, float pb_cutoff,
                   float xobs, float yobs, double xphs, double yphs, double DELTAX, double DELTAY,
                   int primary_beam, double* Vre, double* Vim) {
  if (N & (N - 1)) return 1;
  const long MN = N * N;
#pragma omp parallel for
  for (long idx = 0; idx < MN; idx++) {
    int i = (int)(idx / N), j = (int)(idx % N);
    float nudiv = nu / nu_0;
    float v = I[idx] * powf(nudiv, I[MN + idx]);
    float fl = -1.0f * eta * minpix;
    if (v < fl) v = fl;
    float at = gvo_attenuation(i, j, D, pb_factor, pb_cutoff, nu, xobs, yobs, DELTAX, DELTAY,
                               primary_beam);
    v = v * at * fg_scale;
    if (gcf) v = v * gcf[idx];
    Vre[idx] = (double)v;
    Vim[idx] = 0.0;
  }
#pragma omp parallel for
  for (long r = 0; r < N; r++) fft1d(Vre + r * N, Vim + r * N, N, 1, +1);
#pragma omp parallel for
  for (long c = 0; c < N; c++) fft1d(Vre + c, Vim + c, N, N, +1);
#pragma omp parallel for
  for (long idx = 0; idx < MN; idx++) {
    int i = (int)(idx / N), j = (int)(idx % N);
    double upix = xphs / (double)N, vpix = yphs / (double)N;
    float u = (j < N / 2) ? (float)(upix * j) : (float)(upix * (j - N));
    float v = (i < N / 2) ? (float)(vpix * i) : (float)(vpix * (i - N));
    float phase = -2.0f * (u + v);
    double c = cos(GVO_PI_D * (double)phase), s = sin(GVO_PI_D * (double)phase);
    double a = Vre[idx], b = Vim[idx];
    Vre[idx] = a * c - b * s;
    Vim[idx] = a * s + b * c;
  }
  return 0;
}

/* vis_mod (src/functions.cu:2588-2606) + residual (:2663) + chi2Vector (:2867) +
 * the sum; returns sum_k w |Vr|^2 (NOT halved). Vm, Vr: 2Z floats (may be NULL). */
double gvo_degrid_chi2(long Z, long N, const double* Vre, const double* Vim, const int32_t* cell,
                       const double* frac, const float* Vo, const float* w, float* Vm, float* Vr) {
  double sum = 0.0;
#pragma omp parallel for reduction(+ : sum)
  for (long k = 0; k < Z; k++) {
    double mr = 0.0, mi = 0.0;
    int i1 = cell[2 * k], j1 = cell[2 * k + 1];
    if (i1 >= 0) {
      int i2 = (i1 + 1) % N, j2 = (j1 + 1) % N;
      double du = frac[2 * k], dv = frac[2 * k + 1];
      double w11 = (1.0 - du) * (1.0 - dv), w12 = (1.0 - du) * dv, w21 = du * (1.0 - dv), w22 = du * dv;
      mr = w11 * Vre[N * j1 + i1] + w12 * Vre[N * j2 + i1] + w21 * Vre[N * j1 + i2] + w22 * Vre[N * j2 + i2];
      mi = w11 * Vim[N * j1 + i1] + w12 * Vim[N * j2 + i1] + w21 * Vim[N * j1 + i2] + w22 * Vim[N * j2 + i2];
    }
    double rr = (double)Vo[2 * k] - mr, ri = (double)Vo[2 * k + 1] - mi;
    if (Vm) { Vm[2 * k] = (float)mr; Vm[2 * k + 1] = (float)mi; }
    if (Vr) { Vr[2 * k] = (float)rr; Vr[2 * k + 1] = (float)ri; }
    sum += (double)w[k] * (rr * rr + ri * ri);
  }
  return sum;
}

/* degriddingGPU (src/functions.cu:2205-2254; defined but never launched by the reference):
 * convolutional-kernel degridding of a CENTRED model grid Vg [M][N] (interleaved re, im) at
 * (u, v) in wavelengths. j = int(u/deltau + int(floorf(N/2)) + 0.5), k likewise with M (:2222-2223);
 * taps outside the grid are skipped (:2231-2232); cells with shifted_j < N/2 are read through
 * their Hermitian twin [M - k][N - j], conjugated (:2238-2243). The twin index reaches M (or N)
 * when shifted_k (shifted_j) is 0, which the reference would read out of bounds: such taps are
 * skipped here and the parity tests keep samples away from row/column 0. fp32 accumulation in
 * the kernel's tap order. */
void gvo_degrid_conv(long Z, const double* uvw_l, const float* Vg, const float* table, double deltau,
                     double deltav, int M, int N, int kn, int sx, int sy, float* Vm) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < Z; i++) {
    int j = (int)(uvw_l[3 * i] / deltau + (int)floorf((float)(N / 2)) + 0.5);
    int k = (int)(uvw_l[3 * i + 1] / deltav + (int)floorf((float)(M / 2)) + 0.5);
    float re = 0.0f, im = 0.0f;
    for (int m = -sy; m <= sy; m++)
      for (int n = -sx; n <= sx; n++) {
        int sj = j + n, sk = k + m;
        if (sk < 0 || sk >= M || sj < 0 || sj >= N) continue;
        float kv = table[kn * (m + sy) + (n + sx)];
        if (sj >= N / 2) {
          re += kv * Vg[2 * ((long)N * sk + sj)];
          im += kv * Vg[2 * ((long)N * sk + sj) + 1];
        } else {
          int hj = N - sj, hk = M - sk;
          if (hj >= N || hk >= M) continue;
          re += kv * Vg[2 * ((long)N * hk + hj)];
          im -= kv * Vg[2 * ((long)N * hk + hj) + 1];
        }
      }
    Vm[2 * i] = re;
    Vm[2 * i + 1] = im;
  }
}

/* DChi2 (src/functions.cu:3698-3791 / 3793-3888) in fp64 at the pixels listed in
 * pix[npix] (flat indices N*i + j). Vr, w as floats (what the forward pass left).
 * out[p] = dChi2 value (incl. -1, fg_scale, atten, gcf, /Z); masked pixels -> 0.
 * fp32_phase != 0 reproduces the reference's float conversion of the phase
 * argument before sincospif (:3766), to separate that effect from the rest. */
void gvo_dchi2(long npix, const long* pix, long N, long Z, const double* uvw_l, const float* Vr,
               const float* w, const float* noise, const float* gcf, float noise_cut,
               float fg_scale, float D, float pb_factor, float pb_cutoff, float freq, float ref_xobs,
               float ref_yobs, float phs_xobs, float phs_yobs, double DELTAX, double DELTAY,
               int primary_beam, int normalize, int fp32_phase, double* out) {
#pragma omp parallel for schedule(dynamic, 4)
  for (long p = 0; p < npix; p++) {
    long idx = pix[p];
    int i = (int)(idx / N), j = (int)(idx % N);
    if (noise[idx] >= noise_cut) { out[p] = 0.0; continue; }
    int x0 = (int)phs_xobs, y0 = (int)phs_yobs;
    double x = (j - x0) * DELTAX * GVO_RPDEG_D;
    double y = (i - y0) * DELTAY * GVO_RPDEG_D;
    double z = sqrt(1.0 - x * x - y * y);
    double zm1 = z - 1.0;
    float atten = gvo_attenuation(i, j, D, pb_factor, pb_cutoff, freq, ref_xobs, ref_yobs, DELTAX,
                                  DELTAY, primary_beam);
    double scale = (double)fg_scale * (double)atten;
    if (gcf) scale *= (double)gcf[idx];
    double d = 0.0;
    for (long k = 0; k < Z; k++) {
      double phase = 2.0 * (x * uvw_l[3 * k] + y * uvw_l[3 * k + 1] + zm1 * uvw_l[3 * k + 2]);
      if (fp32_phase) phase = (double)(float)phase;
      phase -= 2.0 * floor(phase * 0.5); /* sin/cos(pi*phase) have period 2 */
      double c = cos(GVO_PI_D * phase), s = sin(GVO_PI_D * phase);
      d += (double)w[k] * ((double)Vr[2 * k] * c + (double)Vr[2 * k + 1] * s);
    }
    d *= scale;
    if (normalize) d /= (double)Z;
    out[p] = -d;
  }
}

/* One block's contribution to the error maps BEFORE noise_reduction, at the pixels pix[npix]:
 * I_nu_0_Noise (src/functions.cu:4076-4111) -> acc0[p] += atten^2 * sum_w * (nu/nu0)^(2 alpha)
 * alpha_Noise  (src/functions.cu:4113-4177) -> acc1[p] += ln^2(nu/nu0) * atten * I_nu * s  if s > 0,
 *   s = sum_k w_k (atten I_nu + Vr.re cos(2 pi phi) - Vr.im sin(2 pi phi)), phi = x u + y v (no w-term).
 * Continuous sums in fp64; masked pixels (noise >= noise_cut) are SET to 0 as both kernels do.
 * fp32_xy != 0 reproduces the reference's `float x, y` and float products Ukv, Vkv (:4133-4160).
 * calculateErrors (:4966-5040) calls this per (field, channel, stokes) and then noise_reduction
 * (:4179-4193): e = e > 0 ? 1/sqrt(e) : 0 — see gvo_error_reduce. */
void gvo_error_accumulate(long npix, const long* pix, long N, long Z, const double* uvw_l, const float* Vr,
                          const float* w, const float* noise, const float* I, float noise_cut, float D,
                          float pb_factor, float pb_cutoff, float freq, float nu_0, float ref_xobs,
                          float ref_yobs, double DELTAX, double DELTAY, int primary_beam, int fp32_xy,
                          double* acc0, double* acc1) {
  const long MN = N * N;
  double sum_w = 0.0;
  for (long k = 0; k < Z; k++) sum_w += (double)w[k];
#pragma omp parallel for schedule(dynamic, 4)
  for (long p = 0; p < npix; p++) {
    long idx = pix[p];
    int i = (int)(idx / N), j = (int)(idx % N);
    if (!(noise[idx] < noise_cut)) { acc0[p] = 0.0; acc1[p] = 0.0; continue; }
    int x0 = (int)ref_xobs, y0 = (int)ref_yobs;
    double x = (j - x0) * DELTAX * GVO_RPDEG_D;
    double y = (i - y0) * DELTAY * GVO_RPDEG_D;
    if (fp32_xy) { x = (double)(float)x; y = (double)(float)y; }
    float atten = gvo_attenuation(i, j, D, pb_factor, pb_cutoff, freq, ref_xobs, ref_yobs, DELTAX,
                                  DELTAY, primary_beam);
    float nudiv = freq / nu_0;
    double alpha = (double)I[MN + idx], I0 = (double)I[idx];
    double I_nu = I0 * pow((double)nudiv, alpha);
    double log_nu = log((double)nudiv);
    acc0[p] += (double)atten * (double)atten * sum_w * pow((double)nudiv, 2.0 * alpha);
    double d = 0.0;
    for (long k = 0; k < Z; k++) {
      double ukv = x * uvw_l[3 * k], vkv = y * uvw_l[3 * k + 1];
      if (fp32_xy) { ukv = (double)(float)ukv; vkv = (double)(float)vkv; }
      double phase = 2.0 * (ukv + vkv);
      if (fp32_xy) phase = (double)(float)phase;
      phase -= 2.0 * floor(phase * 0.5);
      double c = cos(GVO_PI_D * phase), sn = sin(GVO_PI_D * phase);
      d += (double)w[k] * ((double)Vr[2 * k] * c - (double)Vr[2 * k + 1] * sn);
    }
    double s = (double)atten * I_nu * sum_w + d;
    if (s > 0.0) acc1[p] += log_nu * log_nu * (double)atten * I_nu * s;
  }
}
void gvo_error_reduce(long n, double* acc) {
  for (long p = 0; p < n; p++) acc[p] = acc[p] > 0.0 ? 1.0 / sqrt(acc[p]) : 0.0;
}

/* DChi2_total_I_nu_0 (:4000) / DChi2_total_alpha (:3968): multiplier applied to
 * dchi2 at pixel idx for image (flag_opt % 2). */
double gvo_chain(const float* I, long MN, long idx, float nu, float nu_0, float fg_scale,
                 float threshold, int flag_opt) {
  float nudiv = nu / nu_0;
  float dI = powf(nudiv, I[MN + idx]);
  if (flag_opt % 2 == 0) return (double)dI;
  float dalpha = I[idx] * dI * fg_scale * logf(nudiv);
  return (I[idx] > threshold) ? (double)dalpha : 0.0;
}

/* ------------------------------------------------------------------ priors --
 * kinds as in include/gvm_b200.h. Per-pixel float formulas (same operation order
 * as the reference device functions, src/functions.cu:2882-3545), fp64 sum. */
static float approx_abs(float v, float e) { return sqrtf(v * v + e); }

static float prior_value_at(int kind, const float* I, const float* noise, const float* P, long N,
                            int i, int j, float noise_cut, float G, float eta, float eps,
                            float eps_b) {
  long idx = N * i + j;
  if (!(noise[idx] < noise_cut)) return 0.0f;
  float c = I[idx];
  switch (kind) {
    case 0: return c * logf((c / G) + (eta + 1.0f));
    case 6: return c * logf((c / P[idx]) + (eta + 1.0f));
    case 1: return approx_abs(c, eps);
    case 7: return approx_abs(c, eps) / (approx_abs(P[idx], eps) + eps_b);
    case 2:
      if (i < N - 1 && j < N - 1) {
        float r = I[idx + 1], d = I[idx + N];
        float a = (r - c) * (r - c), b = (d - c) * (d - c);
        return sqrtf(a + b + eps);
      }
      return c;
    case 3:
      if (i < N - 1 && j < N - 1) {
        float r = I[idx + 1], d = I[idx + N];
        float dx = c - r, dy = c - d;
        return dx * dx + dy * dy;
      }
      return c;
    case 4:
      if ((i > 0 && i < N - 1) && (j > 0 && j < N - 1)) {
        float l = I[idx - 1], r = I[idx + 1], d = I[idx + N], u = I[idx - N];
        float Dx = l - 2.0f * c + r, Dy = u - 2.0f * c + d;
        return 0.5f * (Dx + Dy) * (Dx + Dy);
      }
      return c;
    case 5:
      if ((i > 0 && i < N - 1) && (j > 0 && j < N - 1)) {
        float l = I[idx - 1], r = I[idx + 1], d = I[idx + N], u = I[idx - N];
        float qp = (c - l) * (c - l) + (c - r) * (c - r) + (c - u) * (c - u) + (c - d) * (c - d);
        return qp / 2.0f;
      }
      return c;
  }
  return 0.0f;
}

static float prior_grad_at(int kind, const float* I, const float* noise, const float* P, long N,
                           int i, int j, float noise_cut, float G, float eta, float eps,
                           float eps_b, float lambda) {
  long idx = N * i + j;
  float g = 0.0f;
  if (noise[idx] < noise_cut) {
    float c = I[idx];
    int inner1 = (i > 0 && i < N - 1) && (j > 0 && j < N - 1);
    switch (kind) {
      case 0: g = logf((c / G) + (eta + 1.0f)) + 1.0f / (1.0f + (((eta + 1.0f) * G) / c)); break;
      case 6: { float Gp = P[idx]; g = logf((c / Gp) + (eta + 1.0f)) + 1.0f / (1.0f + (((eta + 1.0f) * Gp) / c)); break; }
      case 1: g = c / approx_abs(c, eps); break;
      case 7: g = c / (approx_abs(c, eps) * (approx_abs(P[idx], eps) + eps_b)); break;
      case 2:
        if (inner1) {
          float d = I[idx + N], u = I[idx - N], r = I[idx + 1], l = I[idx - 1];
          float dl = I[idx + N - 1], ru = I[idx - N + 1];
          float n0 = 2.0f * c - r - d, n1 = c - l, n2 = c - u;
          float a0 = (c - r) * (c - r) + (c - d) * (c - d) + eps;
          float a1 = (l - c) * (l - c) + (l - dl) * (l - dl) + eps;
          float a2 = (u - ru) * (u - ru) + (u - c) * (u - c) + eps;
          g = n0 / sqrtf(a0) + n1 / sqrtf(a1) + n2 / sqrtf(a2);
        } else g = c;
        break;
      case 3:
        if (inner1) {
          float d = I[idx + N], u = I[idx - N], r = I[idx + 1], l = I[idx - 1];
          g = 8.0f * c - 2.0f * (u + l + d + r);
        } else g = c;
        break;
      case 4:
        if ((i > 1 && i < N - 2) && (j > 1 && j < N - 2)) {
          float d = I[idx + N], u = I[idx - N], r = I[idx + 1], l = I[idx - 1];
          float dl = I[idx + N - 1], dr = I[idx + N + 1], lu = I[idx - N - 1], ru = I[idx - N + 1];
          float d2 = I[idx + 2 * N], u2 = I[idx - 2 * N], l2 = I[idx - 2], r2 = I[idx + 2];
          g = 20.0f * c - 8.0f * (d - r - u - l) + 2.0f * (dl + dr + lu + ru) + d2 + r2 + u2 + l2;
        } else g = 0.0f;
        break;
      case 5:
        if (inner1) {
          float d = I[idx + N], u = I[idx - N], r = I[idx + 1], l = I[idx - 1];
          g = 2.0f * (4.0f * c - d + u + r + l);
        } else g = c;
        break;
    }
  }
  return g * lambda;
}

double gvo_prior_value(int kind, const float* I, const float* noise, const float* prior_image,
                       long N, float noise_cut, float G, float eta, float eps, float eps_b) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s)
  for (long idx = 0; idx < N * N; idx++)
    s += (double)prior_value_at(kind, I, noise, prior_image, N, (int)(idx / N), (int)(idx % N),
                                noise_cut, G, eta, eps, eps_b);
  return s;
}
void gvo_prior_grad(int kind, const float* I, const float* noise, const float* prior_image, long N,
                    float noise_cut, float G, float eta, float eps, float eps_b, float lambda,
                    float* out) {
#pragma omp parallel for
  for (long idx = 0; idx < N * N; idx++)
    out[idx] = prior_grad_at(kind, I, noise, prior_image, N, (int)(idx / N), (int)(idx % N),
                             noise_cut, G, eta, eps, eps_b, lambda);
}

/* --------------------------------------------------------- noise image etc --
 * total_attenuation/weight_image/noise_image (src/functions.cu:2349-2422) for one
 * field, as MFS::setDevice sequences them (src/mfs.cu:850-916). Returns min. */
float gvo_noise_image(long N, float D, float pb_factor, float pb_cutoff, float nu_0, float xobs,
                      float yobs, double DELTAX, double DELTAY, int primary_beam,
                      float noise_jypix, float* noise) {
  long MN = N * N;
  float* wt = (float*)malloc(sizeof(float) * MN);
  float mx = 0.0f;
  for (long idx = 0; idx < MN; idx++) {
    float at = gvo_attenuation((int)(idx / N), (int)(idx % N), D, pb_factor, pb_cutoff, nu_0, xobs,
                               yobs, DELTAX, DELTAY, primary_beam);
    wt[idx] = at * at;
    if (wt[idx] > mx) mx = wt[idx];
  }
  float mn = INFINITY;
  for (long idx = 0; idx < MN; idx++) {
    float nsq = noise_jypix * noise_jypix;
    float nw = (wt[idx] / mx) / nsq;
    noise[idx] = sqrtf(1.0f / nw);
    if (noise[idx] < mn) mn = noise[idx];
  }
  free(wt);
  return mn;
}

/* calculateNoiseAndBeam (src/functions.cu:1700-1840) + calc_sBeam (:1655) +
 * calc_beamSize (:1685) + reduceCPU (:353) for the blocks given; out = {sum_weights,
 * vis_noise, bmaj_deg, bmin_deg, bpa_deg}. The s_uu sums are OpenMP reductions in
 * the reference (order-free), so they are compared with a tolerance. */
void gvo_noise_and_beam(int nblocks, const long* Z, const double* const* uvw_m,
                        const float* const* w, const float* freqs, double* out) {
  double s_uu = 0.0, s_vv = 0.0, s_uv = 0.0;
  float sum_weights = 0.0f;
  for (int b = 0; b < nblocks; b++) {
    if (Z[b] <= 0) continue;
    double luu = 0.0, lvv = 0.0, luv = 0.0;
    for (long k = 0; k < Z[b]; k++) {
      double u = gvo_metres_to_lambda(uvw_m[b][3 * k], freqs[b]);
      double v = gvo_metres_to_lambda(uvw_m[b][3 * k + 1], freqs[b]);
      luu += u * u * w[b][k]; lvv += v * v * w[b][k]; luv += u * v * w[b][k];
    }
    s_uu += luu; s_vv += lvv; s_uv += luv;
    /* reduceCPU<float>: plain running float sum (the compensation term is unused) */
    float sum = w[b][0];
    for (long k = 1; k < Z[b]; k++) sum = sum + w[b][k];
    sum_weights += sum;
  }
  s_uu /= sum_weights; s_vv /= sum_weights; s_uv /= sum_weights;
  float variance = 1.0f / sum_weights;
  double uv2 = s_uv * s_uv, dmv = s_uu - s_vv, dpv = s_uu + s_vv;
  double sq = sqrt(dmv * dmv + 4.0 * uv2);
  double bx = 1.0 / sqrt(2.0) / GVO_PI_D / sqrt(dpv - sq);
  double by = 1.0 / sqrt(2.0) / GVO_PI_D / sqrt(dpv + sq);
  double bz = -0.5 * atan2(2.0 * s_uv, dmv);
  out[0] = sum_weights;
  out[1] = 0.5f * sqrtf(variance);
  out[2] = bx / GVO_RPDEG_D; out[3] = by / GVO_RPDEG_D; out[4] = bz / GVO_RPDEG_D;
}

/* ------------------------------------------------------------- weighting ---
 * Cell index shared by Uniform/Briggs: src/uniformweightingscheme.cu:33-49,
 * src/briggsweightingscheme.cu:72-88. Returns N*y+x or -1. */
static long weight_cell(double um, double vm, float freq, double deltau, double deltav, long M,
                        long N) {
  double u = gvo_metres_to_lambda(um, freq), v = gvo_metres_to_lambda(vm, freq);
  if (u < 0.0) { u *= -1.0; v *= -1.0; }
  double gx = u / fabs(deltau), gy = v / fabs(deltav);
  int x = gx + (int)(floor(N / 2)) + 0.5;
  int y = gy + (int)(floor(M / 2)) + 0.5;
  if (x >= 0 && y >= 0 && x < N && y < M) return N * y + x;
  return -1;
}
void gvo_weight_cells(long Z, const double* uvw_m, float freq, double deltau, double deltav, long M,
                      long N, int64_t* cells) {
  for (long z = 0; z < Z; z++)
    cells[z] = weight_cell(uvw_m[3 * z], uvw_m[3 * z + 1], freq, deltau, deltav, M, N);
}

/* UVTaper::getValue, include/classes/uvtaper.cuh:100-118 */
static float taper_value(const float* t, double u, double v) {
  /* t = {sigma_maj, sigma_min, bpa, amplitude}, centre 0 */
  double x = u, y = v;
  float cb = cosf(t[2]), sb = sinf(t[2]), s2 = sinf(2.0f * t[2]);
  float a = (cb * cb) / (2.0f * t[0] * t[0]) + (sb * sb) / (2.0f * t[1] * t[1]);
  float b = s2 / (2.0f * t[0] * t[0]) - s2 / (2.0f * t[1] * t[1]);
  float c = (sb * sb) / (2.0f * t[0] * t[0]) + (cb * cb) / (2.0f * t[1] * t[1]);
  return (float)(t[3] * exp(-a * x * x - b * x * y - c * y * y));
}

/* scheme: 0 natural, 1 uniform, 2 briggs, 3 radial (single-thread order). Blocks
 * are given in the reference's loop order (field, channel, stokes). taper may be
 * NULL. Weights updated in place. */
void gvo_weights(int scheme, float robust, long M, long N, double deltau, double deltav,
                 int nblocks, const long* Z, const double* const* uvw_m, const float* freqs,
                 float* const* w, const float* taper) {
  float* g = (float*)calloc((size_t)(M * N), sizeof(float));
  if (scheme == 0 || scheme == 3) {
    for (int b = 0; b < nblocks; b++)
      for (long z = 0; z < Z[b]; z++) {
        double u = gvo_metres_to_lambda(uvw_m[b][3 * z], freqs[b]);
        double v = gvo_metres_to_lambda(uvw_m[b][3 * z + 1], freqs[b]);
        if (scheme == 3) {
          /* src/radialweightingscheme.cu: no Hermitian fold, distance in float */
          w[b][z] *= gvo_distance((float)u, (float)v, 0.0f, 0.0f);
        } else if (u < 0.0) { u *= -1.0; v *= -1.0; }
        if (taper) w[b][z] *= taper_value(taper, u, v);
      }
    free(g);
    return;
  }
  float f_squared = 0.0f;
  if (scheme == 2) {
    /* src/briggsweightingscheme.cu:46-110 */
    float sum_w = 0.0f, sum_g2 = 0.0f;
    for (int b = 0; b < nblocks; b++) {
      float acc = 0.0f;
      for (long z = 0; z < Z[b]; z++) acc += w[b][z]; /* std::accumulate(..., 0.0f) */
      sum_w += acc;
    }
    for (int b = 0; b < nblocks; b++) {
      for (long z = 0; z < Z[b]; z++) {
        long c = weight_cell(uvw_m[b][3 * z], uvw_m[b][3 * z + 1], freqs[b], deltau, deltav, M, N);
        if (c >= 0) g[c] += w[b][z];
      }
      for (long m = 0; m < M; m++)
        for (long n = N / 2; n < N; n++) sum_g2 += g[N * m + n] * g[N * m + n];
    }
    float avg = sum_g2 / sum_w;
    f_squared = (5.0f * powf(10.0f, -robust)) * (5.0f * powf(10.0f, -robust)) / avg;
    memset(g, 0, sizeof(float) * (size_t)(M * N));
  }
  for (int b = 0; b < nblocks; b++) {
    int64_t* cells = (int64_t*)malloc(sizeof(int64_t) * (size_t)(Z[b] > 0 ? Z[b] : 1));
    for (long z = 0; z < Z[b]; z++) {
      cells[z] = weight_cell(uvw_m[b][3 * z], uvw_m[b][3 * z + 1], freqs[b], deltau, deltav, M, N);
      if (cells[z] >= 0) g[cells[z]] += w[b][z];
    }
    for (long z = 0; z < Z[b]; z++) {
      if (cells[z] >= 0) {
        if (scheme == 1) w[b][z] /= g[cells[z]];
        else w[b][z] /= (1.0 + g[cells[z]] * f_squared);
      } else {
        w[b][z] = 0.0f;
      }
      if (taper) {
        double u = gvo_metres_to_lambda(uvw_m[b][3 * z], freqs[b]);
        double v = gvo_metres_to_lambda(uvw_m[b][3 * z + 1], freqs[b]);
        if (u < 0.0) { u *= -1.0; v *= -1.0; }
        w[b][z] *= taper_value(taper, u, v);
      }
    }
    for (long z = 0; z < Z[b]; z++)
      if (cells[z] >= 0) g[cells[z]] = 0.0f; /* std::fill_n(g, M*N, 0) */
    free(cells);
  }
  free(g);
}

/* --------------------------------------------------------------- gridding --
 * do_gridding for one block (src/functions.cu:1418-1612), single-thread order.
 * kernel: ck_m x ck_n table. Outputs sized M*N; returns the number of cells. */
long gvo_gridding(long M, long N, double deltau, double deltav, float freq, long Z,
                  const double* uvw_m, const float* Vo, const float* w, const float* kernel,
                  int ck_m, int ck_n, int support_x, int support_y, double* uvw_out,
                  float* Vo_out, float* w_out) {
  size_t MN = (size_t)(M * N);
  float* gw = (float*)calloc(MN, sizeof(float));
  float* gw2 = (float*)calloc(MN, sizeof(float));
  float* gvr = (float*)calloc(MN, sizeof(float));
  float* gvi = (float*)calloc(MN, sizeof(float));
  double center_j = floor(N / 2.0), center_k = floor(M / 2.0);
  float lambda = gvo_freq_to_wavelength(freq);
  for (long z = 0; z < 2 * Z; z++) {
    long vi = (z < Z) ? z : z - Z;
    double u = uvw_m[3 * vi], v = uvw_m[3 * vi + 1];
    float wt = w[vi], vr = Vo[2 * vi], vim = Vo[2 * vi + 1];
    if (z >= Z) { u *= -1.0; v *= -1.0; vim *= -1.0f; }
    u = gvo_metres_to_lambda(u, freq);
    v = gvo_metres_to_lambda(v, freq);
    double gx = u / deltau, gy = v / deltav;
    double j_fp = gx + center_j + 0.5, k_fp = gy + center_k + 0.5;
    int j = (int)j_fp, k = (int)k_fp;
    for (int m = -support_y; m <= support_y; m++)
      for (int n = -support_x; n <= support_x; n++) {
        int sj = j + n, sk = k + m, kj = n + support_x, ki = m + support_y;
        if (sk >= 0 && sk < M && sj >= 0 && sj < N && ki >= 0 && ki < ck_m && kj >= 0 && kj < ck_n) {
          float ck = kernel[ck_n * ki + kj];
          float ck2 = ck * ck;
          long gi = N * sk + sj;
          gw[gi] += wt * ck;
          gw2[gi] += wt * ck2;
          gvr[gi] += wt * vr * ck;
          gvi[gi] += wt * vim * ck;
        }
      }
  }
  long nout = 0;
  for (long gk = 0; gk < M; gk++)
    for (long gj = 0; gj < N; gj++) {
      long gi = N * gk + gj;
      float ws = gw[gi], aux = gw2[gi];
      float weight = 0.0f, orr = 0.0f, oi = 0.0f;
      if (aux != 0.0f && ws != 0.0f) {
        weight = ws * ws / aux;
        orr = gvr[gi] / ws;
        oi = gvi[gi] / ws;
      }
      if (weight > 0.0f) {
        double ul = (gj - center_j) * deltau, vl = (gk - center_k) * deltav;
        uvw_out[3 * nout] = ul * lambda;
        uvw_out[3 * nout + 1] = vl * lambda;
        uvw_out[3 * nout + 2] = 0.0;
        Vo_out[2 * nout] = orr;
        Vo_out[2 * nout + 1] = oi;
        w_out[nout] = weight;
        nout++;
      }
    }
  free(gw); free(gw2); free(gvr); free(gvi);
  return nout;
}

/* ---------------------------------------------------------------- CKernels --
 * kind: 0 PillBox2D, 1 Gaussian2D, 2 GaussianSinc2D, 3 Sinc2D, 4 PSWF_12D.
 * Restates src/pillBox2D.cu:3-16,149-165; gaussian2D.cu:2-32; sinc2D.cu:2-31;
 * gaussianSinc2D.cu:2-24; pswf_12D.cu:2-76 and the buildKernel loops; support as
 * include/classes/ckernel.cuh:508-511 (both from m). */
static float ck_gaussian2D(float amp, float x, float y, float x0, float y0, float sx, float sy,
                           float w, float alpha) {
  float rx = gvo_distance(x, 0.0f, x0, 0.0f), ry = gvo_distance(0.0f, y, 0.0, y0);
  if (rx < w * sx && ry < w * sy) {
    float fx = rx / (w * sx), fy = ry / (w * sy);
    float vx = powf(fx, alpha), vy = powf(fy, alpha);
    return amp * expf(-1.0f * (vx + vy));
  }
  return 0.0f;
}
static float ck_sincf(float x) { return (x == 0.0f) ? 1.0f : sinf(GVO_PI_F * x) / (GVO_PI_F * x); }
static float ck_sinc1D(float amp, float x, float x0, float sigma, float w) {
  float radius = gvo_distance(x, 0.0f, x0, 0.0f);
  float val = radius / (w * sigma);
  return (radius < w * sigma) ? amp * ck_sincf(val) : 0.0f;
}
static float ck_sinc2D(float amp, float x, float x0, float y, float y0, float sx, float sy, float w) {
  float a = ck_sinc1D(1.0f, x, x0, sx, w), b = ck_sinc1D(1.0f, y, y0, sy, w);
  return amp * a * b;
}
static float ck_pswf_func(float nu) {
  const float mat_p[2][5] = {{8.203343e-2, -3.644705e-1, 6.278660e-1, -5.335581e-1, 2.312756e-1},
                             {4.028559e-3, -3.697768e-2, 1.021332e-1, -1.201436e-1, 6.412774e-2}};
  const float mat_q[2][3] = {{1.0000000e0, 8.212018e-1, 2.078043e-1},
                             {1.0000000e0, 9.599102e-1, 2.918724e-1}};
  float n_nu = fabsf(nu), res = 0.0f;
  if (n_nu > 1.0f) return 0.0f;
  int idx; float nu_end;
  if (n_nu >= 0.0f && n_nu < 0.75) { idx = 0; nu_end = 0.75f; } else { idx = 1; nu_end = 1.0f; }
  float dnusq = n_nu * n_nu - nu_end * nu_end;
  float top = mat_p[idx][0], bottom = mat_q[idx][0];
  for (int i = 1; i < 5; i++) top += mat_p[idx][i] * powf(dnusq, i);
  for (int i = 1; i < 3; i++) bottom += mat_q[idx][i] * powf(dnusq, i);
  if (bottom > 0.0f) res = top / bottom;
  return res;
}
static float ck_pswf_11D(float amp, float x, float x0, float sigma, float w) {
  float radius = gvo_distance(x, 0.0f, x0, 0.0f);
  float nu = radius / (w * sigma);
  if (nu == 0.0f) return 1.0f;
  float p = ck_pswf_func(nu);
  return amp * (1.0f - nu * nu) * p;
}
static float ck_eval(int kind, float x, float y, float sx, float sy, float w, int m, int n,
                     int gcf) {
  switch (kind) {
    case 0: {
      if (gcf) return 1.0f;
      float lx = (m / 2.0f) * sx, ly = (n / 2.0f) * sy;
      float a = (fabs(x) < lx) ? 1.0f : 0.0f, b = (fabs(y) < ly) ? 1.0f : 0.0f;
      return a * b;
    }
    case 1:
      if (gcf) return ck_gaussian2D(1.0f, GVO_PI_F * x, GVO_PI_F * y, GVO_PI_F * 0.0f, GVO_PI_F * 0.0f,
                                    sx, sy, 2.0f * w, 2.0f);
      return ck_gaussian2D(1.0f, x, y, 0.0f, 0.0f, sx, sy, w, 2.0f);
    case 2: {
      if (gcf) return 1.0f;
      float G = ck_gaussian2D(1.0f, x, y, 0.0f, 0.0f, sx, sy, w, 2.0f);
      float S = ck_sinc2D(1.0f, x, 0.0f, y, 0.0f, sx, sy, 1.55f);
      return 1.0f * G * S;
    }
    case 3:
      if (gcf) {
        float dxx = gvo_distance(x, y, 0.0f, 0.0f) * sx, dyy = gvo_distance(x, y, 0.0f, 0.0f) * sy;
        float a = (fabs(dxx) < w * sx) ? 1.0f : 0.0f, b = (fabs(dyy) < w * sy) ? 1.0f : 0.0f;
        return a * b;
      }
      /* Reference quirk (src/sinc2D.cu:146-147, 161-163): buildKernel passes
       * (amp, x, y, x0, y0, ...) to sinc2D(amp, x, x0, y, y0, ...), i.e. x0 := y and
       * y := x0 = 0, so the table is sinc(|x - y|) along the diagonal. Kept verbatim. */
      return ck_sinc2D(1.0f, x, y, 0.0f, 0.0f, sx, sy, w);
    case 4: {
      float a = ck_pswf_11D(1.0f, x, 0.0f, sx, w), b = ck_pswf_11D(1.0f, y, 0.0f, sy, w);
      float v = 1.0f * a * b;
      return gcf ? 1.0f / v : v;
    }
  }
  return 0.0f;
}
/* default w per kind: Gaussian 1, GaussianSinc 2.52, Sinc 1, PSWF 6 */
float gvo_ckernel_default_w(int kind) {
  switch (kind) { case 1: return 1.0f; case 2: return 2.52f; case 3: return 1.0f; case 4: return 6.0f; }
  return 1.0f;
}
/* table[m*n]; gcf != 0 builds the GCF image variant the way initializeGCF does
 * (clone, setmn(M,N), setSigmas(dx,dy), setW(M), buildGCF; ckernel.cuh:82-88). */
void gvo_ckernel(int kind, int m, int n, float sx, float sy, float w, int gcf, float* table) {
  int support_x = (int)floorf(m / 2.0f), support_y = (int)floorf(m / 2.0f);
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      float y = (i - support_y) * sy, x = (j - support_x) * sx;
      table[n * i + j] = ck_eval(kind, x, y, sx, sy, w, m, n, gcf);
    }
}

/* ------------------------------------------------------------ line search --
 * Restates linmin's 1-D search (src/linmin.cu:78-84: ax = 0, xx = 1, mnbrak then
 * brent with TOL 1e-7) with mnbrak (src/mnbrak.cu:44-98) and brent
 * (src/brent.cu:43-125) and the macros of include/nrutil.h:27-29,55. The reference is
 * C++: fabs() on a float is the float overload, literals without suffix are double. */
typedef float (*gvo_fn1d)(float, void*);
static float lm_sign(float a, float b) { return b >= 0.0 ? fabsf(a) : -fabsf(a); }
static float lm_fmax(float a, float b) { return a > b ? a : b; }
int gvo_linmin_1d(gvo_fn1d func, void* user, float* xmin_out, float* fmin_out, int* probes_out) {
  int probes = 0;
#define F(x) (probes++, func((x), user))
  float ax = 0.0f, bx = 1.0f, cx, fa, fb, fc;
  { /* mnbrak */
    const double GOLD = 1.618034, GLIMIT = 100.0;
    const float TINY = (float)1.0e-20;
    float ulim, u, r, q, fu, dum;
    fa = F(ax);
    fb = F(bx);
    if (fb > fa) { dum = ax; ax = bx; bx = dum; dum = fb; fb = fa; fa = dum; }
    cx = (float)(bx + GOLD * (bx - ax));
    fc = F(cx);
    while (fb > fc) {
      r = (bx - ax) * (fb - fc);
      q = (bx - cx) * (fb - fa);
      u = (float)(bx - ((bx - cx) * q - (bx - ax) * r) / (2.0 * lm_sign(lm_fmax(fabsf(q - r), TINY), q - r)));
      ulim = (float)(bx + GLIMIT * (cx - bx));
      if ((bx - u) * (u - cx) > 0.0) {
        fu = F(u);
        if (fu < fc) { ax = bx; bx = u; fa = fb; fb = fu; break; }
        else if (fu > fb) { cx = u; fc = fu; break; }
        u = (float)(cx + GOLD * (cx - bx));
        fu = F(u);
      } else if ((cx - u) * (u - ulim) > 0.0) {
        fu = F(u);
        if (fu < fc) {
          bx = cx; cx = u; u = (float)(cx + GOLD * (cx - bx));
          fb = fc; fc = fu; fu = F(u);
        }
      } else if ((u - ulim) * (ulim - cx) >= 0.0) {
        u = ulim;
        fu = F(u);
      } else {
        u = (float)(cx + GOLD * (cx - bx));
        fu = F(u);
      }
      ax = bx; bx = cx; cx = u;
      fa = fb; fb = fc; fc = fu;
    }
  }
  { /* brent */
    const double CGOLD = 0.3819660, ZEPS = 1.0e-10;
    const float tol = (float)1.0e-7;
    float a, b, d = 0.0f, etemp, fu, fv, fw, fx, p, q, r, tol1, tol2, u, v, w, x, xm;
    float e = 0.0f;
    int iter;
    a = (ax < cx ? ax : cx);
    b = (ax > cx ? ax : cx);
    x = w = v = bx;
    fw = fv = fx = F(x);
    for (iter = 1; iter <= 500; iter++) {
      xm = (float)(0.5 * (a + b));
      tol2 = (float)(2.0 * (tol1 = (float)(tol * fabsf(x) + ZEPS)));
      if (fabsf(x - xm) <= (tol2 - 0.5 * (b - a))) break;
      if (fabsf(e) > tol1) {
        r = (x - w) * (fx - fv);
        q = (x - v) * (fx - fw);
        p = (x - v) * q - (x - w) * r;
        q = (float)(2.0 * (q - r));
        if (q > 0.0) p = -p;
        q = fabsf(q);
        etemp = e;
        e = d;
        if (fabsf(p) >= fabs(0.5 * q * etemp) || p <= q * (a - x) || p >= q * (b - x))
          d = (float)(CGOLD * (e = (x >= xm ? a - x : b - x)));
        else {
          d = p / q;
          u = x + d;
          if (u - a < tol2 || b - u < tol2) d = lm_sign(tol1, xm - x);
        }
      } else {
        d = (float)(CGOLD * (e = (x >= xm ? a - x : b - x)));
      }
      u = (fabsf(d) >= tol1 ? x + d : x + lm_sign(tol1, d));
      fu = F(u);
      if (fu <= fx) {
        if (u >= x) a = x; else b = x;
        v = w; w = x; x = u;
        fv = fw; fw = fx; fx = fu;
      } else {
        if (u < x) a = u; else b = u;
        if (fu <= fw || w == x) { v = w; w = u; fv = fw; fw = fu; }
        else if (fu <= fv || v == x || v == w) { v = u; fv = fu; }
      }
    }
    *xmin_out = x;
    *fmin_out = fx;
  }
#undef F
  if (probes_out) *probes_out = probes;
  return 0;
}
