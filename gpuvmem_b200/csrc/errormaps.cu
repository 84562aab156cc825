// errormaps.cu — error images of I_nu0 and alpha.
//
// Reference: calculateErrors (src/functions.cu:4966-5040), called by
// SecondDerivateError::calculateErrorImage (src/secondderivateerror.cu:6-10) from
// MFS::writeImages under -E (src/mfs.cu:1090-1113). Per (field, channel, stokes) block:
//   I_nu_0_Noise (:4076-4111)  err[0] += atten^2 * sum_k w_k * (nu/nu0)^(2 alpha)
//   alpha_Noise  (:4113-4177)  s = sum_k w_k (atten I_nu + Vr.re cos 2 pi phi - Vr.im sin 2 pi phi),
//                              phi = x u_k + y v_k (no w-term);  err[1] += ln^2(nu/nu0) atten I_nu s  if s > 0
// masked pixels are SET to 0 by every block; then noise_reduction (:4179-4193): 1/sqrt where > 0.
//
// alpha_Noise is the same O(M N Z) direct DFT as DChi2 with w_k Vr_k in place of
// w_k conj(Vr_k) and without the w-term, so it runs on the same contraction kernels
// (tcgen05 k_grad_umma, or the FFT for gridded samples): s = atten I_nu sum_k w_k + d.
#include "gvm_internal.cuh"

namespace {

__global__ void __launch_bounds__(256) k_wsum_partial(const float* __restrict__ w, long Z,
                                                      double* __restrict__ partials) {
  __shared__ double s_part[8];
  double acc = 0.0;
  for (long k = blockIdx.x * 256L + threadIdx.x; k < Z; k += (long)gridDim.x * 256L) acc += (double)w[k];
  acc = gvm_warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; i++) t += s_part[i];
    partials[blockIdx.x] = t;
  }
}
__global__ void k_wsum_finish(const double* __restrict__ partials, int n, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < n; i++) t += partials[i];   // fixed order: deterministic
    out[0] = t;
  }
}

struct ErrParams {
  const float* atten;   // cached attenuation plane or null
  long N, M;
  float noise_cut, freq, nu_0, D, pb_factor, pb_cutoff, xobs, yobs;
  double DELTAX, DELTAY;
  int primary_beam;
};

// I_nu_0_Noise + the per-pixel tail of alpha_Noise in one pass over the image.
__global__ void __launch_bounds__(256) k_err_accumulate(float* __restrict__ err, const float* __restrict__ I,
                                                        const float* __restrict__ noise,
                                                        const float* __restrict__ d_raw,
                                                        const double* __restrict__ wsum, ErrParams p) {
  const long idx = blockIdx.x * 256L + threadIdx.x;
  const long MN = p.M * p.N;
  if (idx >= MN) return;
  if (!(noise[idx] < p.noise_cut)) {
    err[idx] = 0.0f;
    err[MN + idx] = 0.0f;
    return;
  }
  const int i = (int)(idx / p.N), j = (int)(idx % p.N);
  const float atten = p.atten ? p.atten[idx]
                              : gvm_attenuation(i, j, p.D, p.pb_factor, p.pb_cutoff, p.freq, p.xobs, p.yobs,
                                                p.DELTAX, p.DELTAY, p.primary_beam);
  const float sum_weights = (float)wsum[0];
  const float nudiv = p.freq / p.nu_0;
  const float I0 = I[idx], alpha = I[MN + idx];
  err[idx] += atten * atten * sum_weights * powf(nudiv, 2.0f * alpha);
  const float I_nu = I0 * powf(nudiv, alpha);
  const float log_nu = logf(nudiv);
  const float sum_noise = fmaf(atten * I_nu, sum_weights, d_raw[idx]);
  if (sum_noise > 0.0f) err[MN + idx] += log_nu * log_nu * atten * I_nu * sum_noise;
}

__global__ void __launch_bounds__(256) k_err_reduction(float* __restrict__ err, long n) {
  const long idx = blockIdx.x * 256L + threadIdx.x;
  if (idx >= n) return;
  const float v = err[idx];
  err[idx] = v > 0.0f ? 1.0f / sqrtf(v) : 0.0f;
}

}  // namespace

static int error_maps_impl(gvm_engine* e, const float* I_dev, int dist_mode, float* errors_dev);
extern "C" int gvm_error_maps(gvm_engine* e, const float* I_dev, int dist_mode, float* errors_dev) {
  const int rc = error_maps_impl(e, I_dev, dist_mode, errors_dev);
  if (rc && e->world > 1) gvm_dist_abort_comm(e);   // peers must not wait for this rank's collectives
  return rc;
}
static int error_maps_impl(gvm_engine* e, const float* I_dev, int dist_mode, float* errors_dev) {
  GVM_CUDA(cudaSetDevice(e->cfg.device));
  const gvm_config& g = e->cfg;
  const long MN = g.M * g.N;
  const int pix_blocks = (int)((MN + 255) / 256);
  if (e->world <= 1 || e->replicated) dist_mode = GVM_DIST_NONE;
  GVM_CUDA(cudaMemsetAsync(errors_dev, 0, 2 * (size_t)MN * sizeof(float), e->stream));
  double* wsum = e->red_out + 1;
  e->ev_used = 0;
  for (size_t s = 0; s < e->chans.size(); s++) {
    GvmChannel& c = e->chans[s];
    // CHUNKS: every rank holds a slice of the same block and takes part in its two all-reduces, so "empty" is
    // decided from the size of the WHOLE block (identical on all ranks); a rank whose slice is empty contributes zeros
    const bool chunks = dist_mode == GVM_DIST_CHUNKS;
    if ((chunks ? c.Znorm : c.Z) <= 0) continue;
    if (c.Z <= 0) {
      GVM_CUDA(cudaMemsetAsync(e->dchi2, 0, (size_t)MN * sizeof(float), e->stream));
      GVM_CUDA(cudaMemsetAsync(wsum, 0, sizeof(double), e->stream));
      if (gvm_dist_allreduce_f32(e, e->dchi2, (size_t)MN)) return 1;
      if (gvm_dist_allreduce_f64(e, wsum, 1)) return 1;
    }
    if (c.Z > 0 && c.slot < 0) { gvm_set_error("gvm_error_maps: call gvm_chi2 first (Vr comes from the forward pass)"); return 1; }
    if (c.Z > 0) {
      long want = (c.Z + 256L * 16 - 1) / (256L * 16);
      const int blocks = (int)(want < 1 ? 1 : (want > e->red_blocks ? e->red_blocks : want));
      k_wsum_partial<<<blocks, 256, 0, e->stream>>>(c.w, c.Z, e->red_partials);
      GVM_LAUNCH(e);
      k_wsum_finish<<<1, 32, 0, e->stream>>>(e->red_partials, blocks, wsum);
      GVM_LAUNCH(e);
      // raw DFT sum d[i,j] -> e->dchi2 through the gradient machinery
      e->err_variant = 1;
      int rc = 0;
      const int mode = gvm_pick_grad_mode(e, c);
      e->last_grad_mode = mode;
      if (mode == GVM_GRAD_GRIDFFT) {
        rc = gvm_grad_gridfft(e, c) || gvm_grad_finish(e, c, I_dev, 1, 0, 0, nullptr);
      } else if (mode == GVM_GRAD_SIMT || mode == GVM_GRAD_SIMT_EXACT) {
        int ksplit = 1;
        rc = gvm_grad_simt(e, c, false, &ksplit) || gvm_grad_finish(e, c, I_dev, ksplit, 0, 0, nullptr);
      } else {
        rc = gvm_grad_umma(e, c, I_dev, 0, 0, nullptr);
      }
      e->err_variant = 0;
      if (rc) return 1;
      if (dist_mode == GVM_DIST_CHUNKS) {   // every rank holds a slice of THIS block: finish the sums first
        if (gvm_dist_allreduce_f32(e, e->dchi2, (size_t)MN)) return 1;
        if (gvm_dist_allreduce_f64(e, wsum, 1)) return 1;
      }
    }
    ErrParams p;
    p.atten = gvm_channel_atten(e, c);
    p.N = g.N; p.M = g.M; p.noise_cut = g.noise_cut; p.freq = c.d.freq; p.nu_0 = g.nu_0;
    p.D = c.d.antenna_diameter; p.pb_factor = c.d.pb_factor; p.pb_cutoff = c.d.pb_cutoff;
    p.xobs = c.d.ref_xobs_pix; p.yobs = c.d.ref_yobs_pix; p.DELTAX = g.DELTAX; p.DELTAY = g.DELTAY;
    p.primary_beam = c.d.primary_beam;
    k_err_accumulate<<<pix_blocks, 256, 0, e->stream>>>(errors_dev, I_dev, e->noise, e->dchi2, wsum, p);
    GVM_LAUNCH(e);
    GVM_CUDA(cudaGetLastError());
  }
  if (dist_mode == GVM_DIST_BLOCKS)       // ranks hold disjoint blocks (channels i % world)
    if (gvm_dist_allreduce_f32(e, errors_dev, 2 * (size_t)MN)) return 1;
  k_err_reduction<<<(int)((2 * MN + 255) / 256), 256, 0, e->stream>>>(errors_dev, 2 * MN);
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  return 0;
}
