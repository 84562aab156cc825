// grad_gridfft.cu — chi2 gradient of GRIDDED samples as one inverse FFT.
//
// DChi2 (src/functions.cu:3698-3791) evaluates, per unmasked pixel,
//   d[i,j] = sum_k w_k (Vr_k.re cos 2 pi phi + Vr_k.im sin 2 pi phi),  phi = x_j u_k + y_i v_k + (z-1) w_k
// with x_j = (j - x0) dx, y_i = (i - y0) dy. After do_gridding (src/functions.cu:1339-1653) every
// sample sits on the centre of a uv cell, u_k = m_k du with du = 1/(N dx), and w_k = 0, so
//   phi = ((j - x0) m_k + (i - y0) n_k) / N    (mod 1)
// and the sum is EXACTLY an N x N inverse DFT of the cell-indexed coefficients
//   C[n, m] = sum_{k in cell} w_k conj(Vr_k) exp(-2 pi i (x0 m + y0 n) / N),   d = Re IFFT(C).
// The reference still loops over the sparse list per pixel (O(N^2 Z)); here it is one scatter
// (HBM-bound, 20 B read per sample) onto the Hermitian half plane, one cuFFT complex-to-real transform
// (the library call the spec allows for dense 2-D FFTs; half the bytes of C2C, real output) and the
// common finishing pass. k_prep_channel proves applicability per block
// (GvmChannel::offgrid == 0) from the same fixed-point phase steps the other kernels use.
#include "gvm_internal.cuh"

namespace {

// One thread per sample: 20 B in (du64, dv64 top words, Vr, w), atomics out. The result is REAL, so only the
// Hermitian part of C matters: d = Re IFFT(C) = IFFT((C + C^H)/2) with C^H[n][m] = conj(C[-n][-m]); the
// kernel accumulates H = (C + C^H)/2 on the half plane m <= N/2 only ([N][N/2+1], what cuFFT C2R reads):
// a sample a at (n, m) adds a/2 to H[n][m] when m <= N/2 and conj(a)/2 to H[-n][-m] when -m mod N <= N/2.
// Folded samples have m < N/2, so normally one 8-byte update per sample (two on the columns 0 and N/2).
__global__ void __launch_bounds__(256) k_gridfft_scatter(
    const uint64_t* __restrict__ du64, const uint64_t* __restrict__ dv64, const float2* __restrict__ Vr,
    const float* __restrict__ w, long Z, int N, int x0, int y0, float im_sign, float2* __restrict__ H) {
  const long k = blockIdx.x * 256L + threadIdx.x;
  if (k >= Z) return;
  const float wk = w[k];
  if (wk == 0.0f) return;
  // cell index = phase step per pixel in units of 1/N turn (exact: the step is m/N by construction)
  const double scale = (double)N * 5.421010862427522e-20;  // N * 2^-64
  int m = (int)rint((double)du64[k] * scale);
  int n = (int)rint((double)dv64[k] * scale);
  if (m >= N) m -= N;
  if (n >= N) n -= N;
  // exp(-2 pi i (x0 m + y0 n) / N): reduce the integer product mod N first, so the angle is exact
  const long r = ((long)x0 * m + (long)y0 * n) % N;
  float s, c;
  sincospif(-2.0f * (float)r / (float)N, &s, &c);
  const float2 v = Vr[k];
  const float ar = wk * v.x, ai = -im_sign * wk * v.y;  // w conj(Vr); im_sign = -1 (error maps): w Vr
  const float hr = 0.5f * (ar * c - ai * s), hi = 0.5f * (ar * s + ai * c);
  const int NH = N / 2 + 1;
  if (m <= N / 2) {
    float* cell = reinterpret_cast<float*>(H + ((size_t)n * NH + m));
    atomicAdd(cell, hr);
    atomicAdd(cell + 1, hi);
  }
  const int mm = m ? N - m : 0, nn = n ? N - n : 0;
  if (mm <= N / 2) {
    float* cell = reinterpret_cast<float*>(H + ((size_t)nn * NH + mm));
    atomicAdd(cell, hr);
    atomicAdd(cell + 1, -hi);
  }
}

}  // namespace

int gvm_grad_gridfft(gvm_engine* e, GvmChannel& c) {
  const int N = (int)e->cfg.N;
  const long MN = (long)N * N;
  if (N & 1) { gvm_set_error("gvm_grad_gridfft: odd image size"); return 1; }
  if (gvm_ensure_grad_scratch(e, (size_t)MN)) return 1;
  if (!e->have_plan_c2r) {
    if (cufftPlan2d(&e->plan_c2r, N, N, CUFFT_C2R) != CUFFT_SUCCESS) {
      gvm_set_error("gvm_grad_gridfft: cufftPlan2d(C2R) failed");
      return 1;
    }
    e->have_plan_c2r = true;
    cufftSetStream(e->plan_c2r, e->stream);
  }
  // I_nu / V are free between evaluations: Vr (per sample) is all the gradient needs of the forward pass
  const size_t half_bytes = (size_t)N * (N / 2 + 1) * sizeof(float2);
  GVM_CUDA(cudaMemsetAsync(e->I_nu, 0, half_bytes, e->stream));
  gvm_ev_begin(e);
  k_gridfft_scatter<<<(int)((c.Z + 255) / 256), 256, 0, e->stream>>>(
      c.du64, c.dv64, c.Vr, c.w, c.Z, N, (int)c.d.phs_xobs_pix, (int)c.d.phs_yobs_pix,
      e->err_variant ? -1.0f : 1.0f, e->I_nu);
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  // complex-to-real: the inverse transform of the Hermitian half plane lands in the split-K scratch as the
  // real image d (no separate real-part pass)
  if (cufftExecC2R(e->plan_c2r, reinterpret_cast<cufftComplex*>(e->I_nu),
                   reinterpret_cast<cufftReal*>(e->grad_scratch)) != CUFFT_SUCCESS) {
    gvm_set_error("gvm_grad_gridfft: cufftExecC2R failed");
    return 1;
  }
  GVM_LAUNCH(e);
  gvm_ev_end(e);
  GVM_CUDA(cudaGetLastError());
  return 0;
}
