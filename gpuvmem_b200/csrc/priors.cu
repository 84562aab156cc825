// priors.cu — regularisation terms (values and gradients), the noise image, and
// the image-sized vector kernels the optimizers use.
//
// Reference: one kernel writing a per-pixel term into a
// scratch image + deviceReduce (cudaMalloc/D2H/free per call) for every value,
// one kernel per gradient, one kernel per image for every optimizer update
// (src/functions.cu:2878-3545 device, :4633-4964 hosts, :2721-2865, :3556-3687).
// Here every value is ONE launch: the term is computed and reduced in the same
// pass (warp shuffles -> block partial -> last block finishes in fp64 in a fixed
// order), nothing is allocated, and the per-image loops are folded into the grid.
// All of these are HBM-bound streaming passes over 4*M*N bytes per operand.
#include <cmath>
#include "gvm_internal.cuh"

namespace {

constexpr int kT = 256;

// ---- per-pixel terms, formulas and edge rules copied in meaning (not text) from
// calculateS/DS (:3036/:3057), calculateL1norm/DNormL1 (:2882/:2915),
// calculateTV/DTV (:3238/:3283), calculateTSV/DTSV (:3353/:3397),
// calculateL/DL (:3443/:3490, incl. the "-8(d - r - u - l)" sign slip),
// calculateQP/DQ (:3146/:3189, incl. "4c - d + u + r + l"),
// calculateGL1norm/DGNormL1 (:2954/:2990), SGVector/DSG (:3114/:3130).
struct PriorArgs {
  const float* I;      // image `index` base: I + N*M*index
  const float* noise;
  const float* prior_image;
  float noise_cut, G, eta, eps, eps_b, lambda;
  int N;
};

__device__ __forceinline__ float approx_abs(float v, float eps) { return sqrtf(v * v + eps); }

// the four kinds that only look at the pixel itself (and the prior image): shared by the scalar and the 4-wide kernels
template <int KIND> struct IsPointwise {
  static constexpr bool value = KIND == GVM_PRIOR_ENTROPY || KIND == GVM_PRIOR_GENTROPY || KIND == GVM_PRIOR_L1 || KIND == GVM_PRIOR_GL1;
};
template <int KIND>
__device__ __forceinline__ float pw_value(const PriorArgs& a, float c, float pr) {
  if (KIND == GVM_PRIOR_ENTROPY) return c * logf((c / a.G) + (a.eta + 1.0f));
  if (KIND == GVM_PRIOR_GENTROPY) return c * logf((c / pr) + (a.eta + 1.0f));
  if (KIND == GVM_PRIOR_L1) return approx_abs(c, a.eps);
  return approx_abs(c, a.eps) / (approx_abs(pr, a.eps) + a.eps_b);   // GL1
}
template <int KIND>
__device__ __forceinline__ float pw_grad(const PriorArgs& a, float c, float pr) {
  if (KIND == GVM_PRIOR_ENTROPY) return logf((c / a.G) + (a.eta + 1.0f)) + 1.0f / (1.0f + (((a.eta + 1.0f) * a.G) / c));
  if (KIND == GVM_PRIOR_GENTROPY) return logf((c / pr) + (a.eta + 1.0f)) + 1.0f / (1.0f + (((a.eta + 1.0f) * pr) / c));
  if (KIND == GVM_PRIOR_L1) return c / approx_abs(c, a.eps);
  return c / (approx_abs(c, a.eps) * (approx_abs(pr, a.eps) + a.eps_b));   // GL1
}

template <int KIND>
__device__ __forceinline__ float prior_value_at(const PriorArgs& a, int i, int j) {
  const int N = a.N;
  const long idx = (long)N * i + j;
  const float* I = a.I;
  if (!(a.noise[idx] < a.noise_cut)) return 0.0f;
  const float c = I[idx];
  if (IsPointwise<KIND>::value)
    return pw_value<KIND>(a, c, (KIND == GVM_PRIOR_GENTROPY || KIND == GVM_PRIOR_GL1) ? a.prior_image[idx] : 0.f);
  if (KIND == GVM_PRIOR_TV) {
    if (i < N - 1 && j < N - 1) {
      const float r = I[idx + 1], d = I[idx + N];
      const float dxy0 = (r - c) * (r - c), dxy1 = (d - c) * (d - c);
      return sqrtf(dxy0 + dxy1 + a.eps);
    }
    return c;
  }
  if (KIND == GVM_PRIOR_TSV) {
    if (i < N - 1 && j < N - 1) {
      const float r = I[idx + 1], d = I[idx + N];
      const float dx = c - r, dy = c - d;
      return dx * dx + dy * dy;
    }
    return c;
  }
  if (KIND == GVM_PRIOR_LAPLACIAN) {
    if ((i > 0 && i < N - 1) && (j > 0 && j < N - 1)) {
      const float l = I[idx - 1], r = I[idx + 1], d = I[idx + N], u = I[idx - N];
      const float Dx = l - 2.0f * c + r, Dy = u - 2.0f * c + d;
      return 0.5f * (Dx + Dy) * (Dx + Dy);
    }
    return c;
  }
  if (KIND == GVM_PRIOR_QUADRATIC) {
    if ((i > 0 && i < N - 1) && (j > 0 && j < N - 1)) {
      const float l = I[idx - 1], r = I[idx + 1], d = I[idx + N], u = I[idx - N];
      float qp = (c - l) * (c - l) + (c - r) * (c - r) + (c - u) * (c - u) + (c - d) * (c - d);
      return qp / 2.0f;
    }
    return c;
  }
  return 0.0f;
}

template <int KIND>
__device__ __forceinline__ float prior_grad_at(const PriorArgs& a, int i, int j) {
  const int N = a.N;
  const long idx = (long)N * i + j;
  const float* I = a.I;
  float g = 0.0f;
  if (a.noise[idx] < a.noise_cut) {
    const float c = I[idx];
    if (IsPointwise<KIND>::value) {
      g = pw_grad<KIND>(a, c, (KIND == GVM_PRIOR_GENTROPY || KIND == GVM_PRIOR_GL1) ? a.prior_image[idx] : 0.f);
    } else if (KIND == GVM_PRIOR_TV) {
      if ((i > 0 && i < N - 1) && (j > 0 && j < N - 1)) {
        const float d = I[idx + N], u = I[idx - N], r = I[idx + 1], l = I[idx - 1];
        const float dl = I[idx + N - 1], ru = I[idx - N + 1];
        const float num0 = 2.0f * c - r - d, num1 = c - l, num2 = c - u;
        const float a0 = (c - r) * (c - r) + (c - d) * (c - d) + a.eps;
        const float a1 = (l - c) * (l - c) + (l - dl) * (l - dl) + a.eps;
        const float a2 = (u - ru) * (u - ru) + (u - c) * (u - c) + a.eps;
        g = num0 / sqrtf(a0) + num1 / sqrtf(a1) + num2 / sqrtf(a2);
      } else {
        g = c;
      }
    } else if (KIND == GVM_PRIOR_TSV) {
      if ((i > 0 && i < N - 1) && (j > 0 && j < N - 1)) {
        const float d = I[idx + N], u = I[idx - N], r = I[idx + 1], l = I[idx - 1];
        g = 8.0f * c - 2.0f * (u + l + d + r);
      } else {
        g = c;
      }
    } else if (KIND == GVM_PRIOR_LAPLACIAN) {
      if ((i > 1 && i < N - 2) && (j > 1 && j < N - 2)) {
        const float d = I[idx + N], u = I[idx - N], r = I[idx + 1], l = I[idx - 1];
        const float dl = I[idx + N - 1], dr = I[idx + N + 1], lu = I[idx - N - 1], ru = I[idx - N + 1];
        const float d2 = I[idx + 2 * (long)N], u2 = I[idx - 2 * (long)N], l2 = I[idx - 2], r2 = I[idx + 2];
        g = 20.0f * c - 8.0f * (d - r - u - l) + 2.0f * (dl + dr + lu + ru) + d2 + r2 + u2 + l2;
      } else {
        g = 0.0f;
      }
    } else if (KIND == GVM_PRIOR_QUADRATIC) {
      if ((i > 0 && i < N - 1) && (j > 0 && j < N - 1)) {
        const float d = I[idx + N], u = I[idx - N], r = I[idx + 1], l = I[idx - 1];
        g = 2.0f * (4.0f * c - d + u + r + l);
      } else {
        g = c;
      }
    }
  }
  return g * a.lambda;
}

// Block-level finish shared by every reducing kernel in this file: up to two
// sums and one max per block; the last block to arrive adds the partials in
// index order in fp64.
__device__ __forceinline__ void block_reduce_finish(float s0, float s1, float mx,
                                                    double* __restrict__ partials,
                                                    float* __restrict__ pmax,
                                                    unsigned int* __restrict__ counter,
                                                    double* __restrict__ out) {
  __shared__ float sh0[kT / 32], sh1[kT / 32], shm[kT / 32];
  __shared__ bool last;
  s0 = gvm_warp_sum(s0);
  s1 = gvm_warp_sum(s1);
  mx = gvm_warp_max(mx);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh0[warp] = s0; sh1[warp] = s1; shm[warp] = mx; }
  __syncthreads();
  if (warp == 0) {
    float a = lane < kT / 32 ? sh0[lane] : 0.f, b = lane < kT / 32 ? sh1[lane] : 0.f;
    float m = lane < kT / 32 ? shm[lane] : 0.f;
    a = gvm_warp_sum(a); b = gvm_warp_sum(b); m = gvm_warp_max(m);
    if (lane == 0) {
      partials[2 * blockIdx.x] = (double)a;
      partials[2 * blockIdx.x + 1] = (double)b;
      pmax[blockIdx.x] = m;
      __threadfence();
      last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (last && warp == 0) {
    __threadfence();
    double a = 0.0, b = 0.0;
    float m = 0.f;
    for (unsigned int k = lane; k < gridDim.x; k += 32) {
      a += partials[2 * k];
      b += partials[2 * k + 1];
      m = fmaxf(m, pmax[k]);
    }
    a = gvm_warp_sum_d(a); b = gvm_warp_sum_d(b); m = gvm_warp_max(m);
    if (lane == 0) { out[0] = a; out[1] = b; out[2] = (double)m; *counter = 0u; }
  }
}

template <int KIND>
__global__ void __launch_bounds__(kT) k_prior_value(PriorArgs a, double* partials, float* pmax,
                                                    unsigned int* counter, double* out) {
  // rows over the blocks, columns over the threads: (i, j) without a 64-bit division per pixel, four independent
  // coalesced loads in flight per thread
  float s = 0.f;
  for (int i = blockIdx.x; i < a.N; i += gridDim.x) {
#pragma unroll 4
    for (int j = threadIdx.x; j < a.N; j += kT) s += prior_value_at<KIND>(a, i, j);
  }
  block_reduce_finish(s, 0.f, 0.f, partials, pmax, counter, out);
}

// kAdd: dgi += lambda * g in the same pass (DS written by the gradient kernel and then added by AddToDPhi in the
// reference, src/functions.cu:3890 — the same fp32 value is added, without the round trip through device_DS).
// 2-D launch (x: column, y: row): no 64-bit division per pixel.
template <int KIND, bool kAdd>
__global__ void __launch_bounds__(kT) k_prior_grad(PriorArgs a, float* __restrict__ dgi) {
  const int j = blockIdx.x * kT + threadIdx.x, i = blockIdx.y;
  if (j >= a.N) return;
  const long idx = (long)a.N * i + j;
  const float g = prior_grad_at<KIND>(a, i, j);
  dgi[idx] = kAdd ? __fadd_rn(dgi[idx], g) : g;   // never contracted with the lambda product: same value as DS + AddToDPhi
}

__global__ void __launch_bounds__(kT) k_add_to_dphi(float* __restrict__ dphi,
                                                    const float* __restrict__ dgi, long MN) {
  const long idx = blockIdx.x * (long)kT + threadIdx.x;
  if (idx < MN) dphi[idx] += dgi[idx];
}

// ------------------------------------------------------------- noise image --
// total_attenuation (:2349) -> weight_image (:2401) accumulated per field.
__global__ void __launch_bounds__(kT) k_weight_accum(float* __restrict__ weight, long N, long M,
                                                     float D, float pbf, float pbc, float nu0,
                                                     float xobs, float yobs, double DELTAX,
                                                     double DELTAY, int pb) {
  const long idx = blockIdx.x * (long)kT + threadIdx.x;
  if (idx >= M * N) return;
  const float at = gvm_attenuation((int)(idx / N), (int)(idx % N), D, pbf, pbc, nu0, xobs, yobs,
                                   DELTAX, DELTAY, pb);
  weight[idx] += at * at;
}
__global__ void __launch_bounds__(kT) k_max_reduce(const float* __restrict__ v, long n,
                                                   double* partials, float* pmax,
                                                   unsigned int* counter, double* out) {
  float m = 0.f;
  for (long idx = blockIdx.x * (long)kT + threadIdx.x; idx < n; idx += (long)gridDim.x * kT)
    m = fmaxf(m, v[idx]);
  block_reduce_finish(0.f, 0.f, m, partials, pmax, counter, out);
}
// noise_image (:2409-2422); the min is taken as max of -noise... noise > 0 so we
// reduce 1/noise with max and invert on the host instead.
__global__ void __launch_bounds__(kT) k_noise_image(float* __restrict__ noise,
                                                    const float* __restrict__ weight, long n,
                                                    float max_weight, float noise_jypix) {
  const long idx = blockIdx.x * (long)kT + threadIdx.x;
  if (idx >= n) return;
  const float nsq = noise_jypix * noise_jypix;
  const float nw = (weight[idx] / max_weight) / nsq;
  noise[idx] = sqrtf(1.0f / nw);
}
__global__ void __launch_bounds__(kT) k_min_reduce(const float* __restrict__ v, long n,
                                                   float* __restrict__ out_bits) {
  float m = CUDART_INF_F;
  for (long idx = blockIdx.x * (long)kT + threadIdx.x; idx < n; idx += (long)gridDim.x * kT)
    m = fminf(m, v[idx]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
  // positive floats (and +inf) order like their bit patterns
  if ((threadIdx.x & 31) == 0) atomicMin(reinterpret_cast<int*>(out_bits), __float_as_int(m));
}

// ---------------------------------------------------------------- vector ops --
__global__ void __launch_bounds__(kT) k_evaluate_xt(float* __restrict__ xt,
                                                    const float* __restrict__ pcom,
                                                    const float* __restrict__ xicom, float x,
                                                    long MN, int image_count, float floor0,
                                                    int nopositivity) {
  const long idx = blockIdx.x * (long)kT + threadIdx.x;
  if (idx >= MN * image_count) return;
  const float v = pcom[idx] + x * xicom[idx];
  if (idx < MN && !nopositivity) xt[idx] = (v > floor0) ? v : floor0;
  else xt[idx] = v;
}
__global__ void __launch_bounds__(kT) k_new_p(float* __restrict__ p, float* __restrict__ xi,
                                              float xmin, long MN, int image_count, float floor0,
                                              int nopositivity) {
  const long idx = blockIdx.x * (long)kT + threadIdx.x;
  if (idx >= MN * image_count) return;
  float x = xi[idx] * xmin;
  const float pv = p[idx];
  if (idx < MN && !nopositivity) {
    if (pv + x > floor0) { p[idx] = pv + x; }
    else { p[idx] = floor0; x = 0.0f; }
  } else {
    p[idx] = pv + x;
  }
  xi[idx] = x;
}
__global__ void __launch_bounds__(kT) k_dot(const float* __restrict__ a, const float* __restrict__ b,
                                            long n, double* partials, float* pmax,
                                            unsigned int* counter, double* out) {
  float s = 0.f;
  for (long idx = blockIdx.x * (long)kT + threadIdx.x; idx < n; idx += (long)gridDim.x * kT)
    s += a[idx] * b[idx];
  block_reduce_finish(s, 0.f, 0.f, partials, pmax, counter, out);
}
__global__ void __launch_bounds__(kT) k_gg_dgg(const float* __restrict__ xi,
                                               const float* __restrict__ g, long n,
                                               double* partials, float* pmax,
                                               unsigned int* counter, double* out) {
  float s0 = 0.f, s1 = 0.f;
  for (long idx = blockIdx.x * (long)kT + threadIdx.x; idx < n; idx += (long)gridDim.x * kT) {
    const float gv = g[idx], xv = xi[idx];
    s0 += gv * gv;
    s1 += (xv + gv) * xv;
  }
  block_reduce_finish(s0, s1, 0.f, partials, pmax, counter, out);
}
__global__ void __launch_bounds__(kT) k_grad_condition(const float* __restrict__ xi,
                                                       const float* __restrict__ p, float den,
                                                       long n, double* partials, float* pmax,
                                                       unsigned int* counter, double* out) {
  float m = 0.f;
  for (long idx = blockIdx.x * (long)kT + threadIdx.x; idx < n; idx += (long)gridDim.x * kT)
    m = fmaxf(m, fabsf(xi[idx]) * fmaxf(fabsf(p[idx]), 1.0f) / den);
  block_reduce_finish(0.f, 0.f, m, partials, pmax, counter, out);
}
__global__ void __launch_bounds__(kT) k_new_xi(float* __restrict__ g, float* __restrict__ xi,
                                               float* __restrict__ h, float gam, long n) {
  const long idx = blockIdx.x * (long)kT + threadIdx.x;
  if (idx >= n) return;
  const float gv = -xi[idx];
  g[idx] = gv;
  const float hv = (gam == 0.0f) ? gv : gv + gam * h[idx];
  h[idx] = hv;
  xi[idx] = hv;
}
__global__ void __launch_bounds__(kT) k_axpby(float a, const float* __restrict__ x, float b,
                                              float* __restrict__ y, long n) {
  const long idx = blockIdx.x * (long)kT + threadIdx.x;
  if (idx >= n) return;
  y[idx] = (b == 0.0f) ? a * x[idx] : a * x[idx] + b * y[idx];
}

__global__ void __launch_bounds__(kT) k_absmax(const float* __restrict__ v, long n, double* partials,
                                               float* pmax, unsigned int* counter, double* out) {
  float m = 0.f;
  for (long idx = blockIdx.x * (long)kT + threadIdx.x; idx < n; idx += (long)gridDim.x * kT)
    m = fmaxf(m, fabsf(v[idx]));
  block_reduce_finish(0.f, 0.f, m, partials, pmax, counter, out);
}
__global__ void __launch_bounds__(kT) k_scale(float* v, float s, long n) {
  const long idx = blockIdx.x * (long)kT + threadIdx.x;
  if (idx < n) v[idx] *= s;
}
// r = y + s (in place on y's copy): y_out = a + b, s_out = c - d (calculateSandY, src/functions.cu:3636-3653)
__global__ void __launch_bounds__(kT) k_lbfgs_sy(float* __restrict__ y_out, float* __restrict__ s_out,
                                                 const float* __restrict__ xi, const float* __restrict__ xi_old,
                                                 const float* __restrict__ p, const float* __restrict__ p_old,
                                                 long n) {
  const long idx = blockIdx.x * (long)kT + threadIdx.x;
  if (idx >= n) return;
  y_out[idx] = xi[idx] - (-1.0f * xi_old[idx]);
  s_out[idx] = p[idx] - p_old[idx];
}

struct RedBuf {
  double* partials; float* pmax; unsigned int* counter; double* out;
};
RedBuf aux_red(gvm_engine* e) {
  // the last reduction slot is reserved for image-sized reductions
  const int slot = e->red_slots - 1;
  RedBuf r;
  r.partials = e->red_partials + (size_t)slot * e->red_blocks - e->red_blocks;  // 2 doubles per block: use two slots
  r.pmax = reinterpret_cast<float*>(e->red_partials + (size_t)e->red_slots * e->red_blocks) +
           (size_t)slot * e->red_blocks;
  r.counter = e->red_counter + slot;
  r.out = e->red_out + 4;
  return r;
}
int red_grid(gvm_engine* e, long n) {
  long b = (n + (long)kT * 4 - 1) / ((long)kT * 4);
  if (b < 1) b = 1;
  if (b > e->red_blocks) b = e->red_blocks;
  return (int)b;
}
int fetch_red(gvm_engine* e, double* v3) {
  GVM_CUDA(cudaMemcpyAsync(e->h_red + 4, e->red_out + 4, 3 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  GVM_CUDA(cudaStreamSynchronize(e->stream));
  v3[0] = e->h_red[4]; v3[1] = e->h_red[5]; v3[2] = e->h_red[6];
  return 0;
}

// 4-wide versions for the pointwise kinds (N a multiple of 4, planes 16-byte aligned): 128-bit loads of the mask, the
// image, the prior image and (kAdd) dphi; same per-pixel arithmetic (pw_value / pw_grad)
template <int KIND>
__global__ void __launch_bounds__(kT) k_prior_value4(PriorArgs a, double* partials, float* pmax, unsigned int* counter,
                                                     double* out) {
  const long nq = (long)a.N * a.N / 4;
  constexpr bool kPrior = KIND == GVM_PRIOR_GENTROPY || KIND == GVM_PRIOR_GL1;
  float s = 0.f;
#pragma unroll 2
  for (long q = blockIdx.x * (long)kT + threadIdx.x; q < nq; q += (long)gridDim.x * kT) {
    const float4 nz = __ldg(reinterpret_cast<const float4*>(a.noise) + q);
    const float4 c = __ldg(reinterpret_cast<const float4*>(a.I) + q);
    float4 pr = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kPrior) pr = __ldg(reinterpret_cast<const float4*>(a.prior_image) + q);
    // same order as the scalar kernel's per-thread sum would be irrelevant: the partial sums differ anyway by layout
    if (nz.x < a.noise_cut) s += pw_value<KIND>(a, c.x, pr.x);
    if (nz.y < a.noise_cut) s += pw_value<KIND>(a, c.y, pr.y);
    if (nz.z < a.noise_cut) s += pw_value<KIND>(a, c.z, pr.z);
    if (nz.w < a.noise_cut) s += pw_value<KIND>(a, c.w, pr.w);
  }
  block_reduce_finish(s, 0.f, 0.f, partials, pmax, counter, out);
}
template <int KIND, bool kAdd>
__global__ void __launch_bounds__(kT) k_prior_grad4(PriorArgs a, float* __restrict__ dgi) {
  const long q = blockIdx.x * (long)kT + threadIdx.x;
  if (4 * q >= (long)a.N * a.N) return;
  constexpr bool kPrior = KIND == GVM_PRIOR_GENTROPY || KIND == GVM_PRIOR_GL1;
  const float4 nz = __ldg(reinterpret_cast<const float4*>(a.noise) + q);
  const bool u0 = nz.x < a.noise_cut, u1 = nz.y < a.noise_cut, u2 = nz.z < a.noise_cut, u3 = nz.w < a.noise_cut;
  float4* dp = reinterpret_cast<float4*>(dgi) + q;
  if (!(u0 || u1 || u2 || u3)) {          // masked quad: the gradient is +0 * lambda; adding it changes nothing but -0
    if (!kAdd) *dp = make_float4(0.0f * a.lambda, 0.0f * a.lambda, 0.0f * a.lambda, 0.0f * a.lambda);
    else {
      float4 r = *dp;
      const float z = 0.0f * a.lambda;
      r.x = __fadd_rn(r.x, z); r.y = __fadd_rn(r.y, z); r.z = __fadd_rn(r.z, z); r.w = __fadd_rn(r.w, z);
      *dp = r;
    }
    return;
  }
  const float4 c = __ldg(reinterpret_cast<const float4*>(a.I) + q);
  float4 pr = make_float4(0.f, 0.f, 0.f, 0.f);
  if (kPrior) pr = __ldg(reinterpret_cast<const float4*>(a.prior_image) + q);
  float4 g;
  g.x = (u0 ? pw_grad<KIND>(a, c.x, pr.x) : 0.0f) * a.lambda;
  g.y = (u1 ? pw_grad<KIND>(a, c.y, pr.y) : 0.0f) * a.lambda;
  g.z = (u2 ? pw_grad<KIND>(a, c.z, pr.z) : 0.0f) * a.lambda;
  g.w = (u3 ? pw_grad<KIND>(a, c.w, pr.w) : 0.0f) * a.lambda;
  if (kAdd) {
    const float4 r = *dp;
    g.x = __fadd_rn(r.x, g.x); g.y = __fadd_rn(r.y, g.y); g.z = __fadd_rn(r.z, g.z); g.w = __fadd_rn(r.w, g.w);
  }
  *dp = g;
}
inline bool prior_vec4_ok(const PriorArgs& a, const void* out) {
  return a.N % 4 == 0 && (((uintptr_t)a.I | (uintptr_t)a.noise | (uintptr_t)a.prior_image | (uintptr_t)out) & 15) == 0;
}

template <int KIND>
int launch_value(gvm_engine* e, const PriorArgs& a, double* out = nullptr) {
  RedBuf r = aux_red(e);
  if (out) r.out = out;
  if (IsPointwise<KIND>::value && prior_vec4_ok(a, nullptr)) {
    k_prior_value4<KIND><<<red_grid(e, (long)a.N * a.N / 4), kT, 0, e->stream>>>(a, r.partials, r.pmax, r.counter, r.out);
    return 0;
  }
  const int blocks = a.N < e->red_blocks ? a.N : e->red_blocks;
  k_prior_value<KIND><<<blocks, kT, 0, e->stream>>>(a, r.partials, r.pmax, r.counter, r.out);
  return 0;
}
template <int KIND>
int launch_grad(gvm_engine* e, const PriorArgs& a, float* dgi, bool add) {
  if (IsPointwise<KIND>::value && prior_vec4_ok(a, dgi)) {
    const unsigned blocks = (unsigned)(((long)a.N * a.N / 4 + kT - 1) / kT);
    if (add) k_prior_grad4<KIND, true><<<blocks, kT, 0, e->stream>>>(a, dgi);
    else k_prior_grad4<KIND, false><<<blocks, kT, 0, e->stream>>>(a, dgi);
    return 0;
  }
  const dim3 grid((unsigned)((a.N + kT - 1) / kT), (unsigned)a.N);
  if (add) k_prior_grad<KIND, true><<<grid, kT, 0, e->stream>>>(a, dgi);
  else k_prior_grad<KIND, false><<<grid, kT, 0, e->stream>>>(a, dgi);
  return 0;
}

int make_args(gvm_engine* e, int kind, const float* I_dev, int image_index,
              const gvm_prior_params* p, float lambda, PriorArgs* a) {
  if (image_index < 0 || image_index > 1) { gvm_set_error("prior: image index %d out of range", image_index); return 1; }
  const long MN = (long)e->cfg.M * e->cfg.N;
  a->I = I_dev + MN * image_index;
  a->noise = e->noise;
  a->prior_image = p ? p->prior_image_dev : nullptr;
  a->noise_cut = e->cfg.noise_cut;
  a->G = p ? p->prior_value : 1.0f;
  a->eta = p ? p->eta : e->cfg.eta;
  a->eps = p ? p->epsilon : 0.0f;
  a->eps_b = p ? p->epsilon_b : 0.0f;
  a->lambda = lambda;
  a->N = (int)e->cfg.N;
  if ((kind == GVM_PRIOR_GENTROPY || kind == GVM_PRIOR_GL1) && !a->prior_image) {
    gvm_set_error("prior kind %d needs prior_image_dev", kind);
    return 1;
  }
  return 0;
}

}  // namespace

// the forked branch joins the main stream (before the slots are copied back / the capture ends)
int gvm_join_branch(gvm_engine* e) {
  if (e->join_pending) {
    GVM_CUDA(cudaStreamWaitEvent(e->stream, e->ev_join, 0));
    e->join_pending = false;
  }
  return 0;
}

extern "C" {

static int prior_value_launch(gvm_engine* e, int kind, const float* I_dev, int image_index,
                              const gvm_prior_params* p, double* out) {
  GVM_CUDA(cudaSetDevice(e->cfg.device));
  PriorArgs a;
  if (make_args(e, kind, I_dev, image_index, p, 1.0f, &a)) return 1;
  switch (kind) {
    case GVM_PRIOR_ENTROPY: launch_value<GVM_PRIOR_ENTROPY>(e, a, out); break;
    case GVM_PRIOR_L1: launch_value<GVM_PRIOR_L1>(e, a, out); break;
    case GVM_PRIOR_TV: launch_value<GVM_PRIOR_TV>(e, a, out); break;
    case GVM_PRIOR_TSV: launch_value<GVM_PRIOR_TSV>(e, a, out); break;
    case GVM_PRIOR_LAPLACIAN: launch_value<GVM_PRIOR_LAPLACIAN>(e, a, out); break;
    case GVM_PRIOR_QUADRATIC: launch_value<GVM_PRIOR_QUADRATIC>(e, a, out); break;
    case GVM_PRIOR_GENTROPY: launch_value<GVM_PRIOR_GENTROPY>(e, a, out); break;
    case GVM_PRIOR_GL1: launch_value<GVM_PRIOR_GL1>(e, a, out); break;
    default: gvm_set_error("gvm_prior_value: unknown kind %d", kind); return 1;
  }
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  return 0;
}

int gvm_prior_value(gvm_engine* e, int kind, const float* I_dev, int image_index,
                    const gvm_prior_params* p, float* value_out) {
  if (prior_value_launch(e, kind, I_dev, image_index, p, nullptr)) return 1;
  double v[3];
  if (fetch_red(e, v)) return 1;
  *value_out = (float)v[0];
  return 0;
}

// ---- one host synchronisation for all terms of an objective evaluation ----
int gvm_prior_value_to_slot(gvm_engine* e, int kind, const float* I_dev, int image_index,
                            const gvm_prior_params* p, int slot) {
  if (slot < 0 || slot >= GVM_OBJ_SLOTS) { gvm_set_error("gvm_prior_value_to_slot: slot %d out of range", slot); return 1; }
  if (e->capturing && e->fork_valid) {
    // captured evaluation: a branch that starts behind the image preparation of chi2 and joins before the slots are read
    GVM_CUDA(cudaStreamWaitEvent(e->stream2, e->ev_fork, 0));
    cudaStream_t main_stream = e->stream;
    e->stream = e->stream2;
    const int rc = prior_value_launch(e, kind, I_dev, image_index, p, e->obj_slots + 3 * slot);
    e->stream = main_stream;
    if (rc) return rc;
    GVM_CUDA(cudaEventRecord(e->ev_join, e->stream2));
    e->join_pending = true;
    return 0;
  }
  return prior_value_launch(e, kind, I_dev, image_index, p, e->obj_slots + 3 * slot);
}

int gvm_chi2_to_slot(gvm_engine* e, float* I_dev, int normalize, int slot) {
  if (slot < 0 || slot >= GVM_OBJ_SLOTS) { gvm_set_error("gvm_chi2_to_slot: slot %d out of range", slot); return 1; }
  return gvm_chi2_async(e, I_dev, normalize, e->obj_slots + 3 * slot);
}
int gvm_fetch_slots_enqueue(gvm_engine* e, int n) {
  if (n < 0 || n > GVM_OBJ_SLOTS) { gvm_set_error("gvm_fetch_slots: %d slots requested", n); return 1; }
  if (gvm_join_branch(e)) return 1;
  GVM_CUDA(cudaMemcpyAsync(e->h_slots, e->obj_slots, 3 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  return 0;
}
int gvm_fetch_slots_wait(gvm_engine* e, int n, double* values_out) {
  if (n < 0 || n > GVM_OBJ_SLOTS) { gvm_set_error("gvm_fetch_slots: %d slots requested", n); return 1; }
  GVM_CUDA(cudaStreamSynchronize(e->stream));
  for (int i = 0; i < n; i++) values_out[i] = e->h_slots[3 * i];
  return 0;
}
int gvm_fetch_slots(gvm_engine* e, int n, double* values_out) {
  return gvm_fetch_slots_enqueue(e, n) || gvm_fetch_slots_wait(e, n, values_out);
}

static int prior_grad_launch(gvm_engine* e, int kind, const float* I_dev, int image_index,
                             const gvm_prior_params* p, float lambda, float* out_dev, bool add) {
  GVM_CUDA(cudaSetDevice(e->cfg.device));
  PriorArgs a;
  if (make_args(e, kind, I_dev, image_index, p, lambda, &a)) return 1;
  switch (kind) {
    case GVM_PRIOR_ENTROPY: launch_grad<GVM_PRIOR_ENTROPY>(e, a, out_dev, add); break;
    case GVM_PRIOR_L1: launch_grad<GVM_PRIOR_L1>(e, a, out_dev, add); break;
    case GVM_PRIOR_TV: launch_grad<GVM_PRIOR_TV>(e, a, out_dev, add); break;
    case GVM_PRIOR_TSV: launch_grad<GVM_PRIOR_TSV>(e, a, out_dev, add); break;
    case GVM_PRIOR_LAPLACIAN: launch_grad<GVM_PRIOR_LAPLACIAN>(e, a, out_dev, add); break;
    case GVM_PRIOR_QUADRATIC: launch_grad<GVM_PRIOR_QUADRATIC>(e, a, out_dev, add); break;
    case GVM_PRIOR_GENTROPY: launch_grad<GVM_PRIOR_GENTROPY>(e, a, out_dev, add); break;
    case GVM_PRIOR_GL1: launch_grad<GVM_PRIOR_GL1>(e, a, out_dev, add); break;
    default: gvm_set_error("gvm_prior_grad: unknown kind %d", kind); return 1;
  }
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  return 0;
}

int gvm_prior_grad(gvm_engine* e, int kind, const float* I_dev, int image_index,
                   const gvm_prior_params* p, float lambda, float* dgi_dev) {
  return prior_grad_launch(e, kind, I_dev, image_index, p, lambda, dgi_dev, false);
}

int gvm_prior_grad_add(gvm_engine* e, int kind, const float* I_dev, int image_index, const gvm_prior_params* p,
                       float lambda, float* dphi_dev, int image_to_add) {
  if (image_to_add < 0 || image_to_add > 1) { gvm_set_error("gvm_prior_grad_add: image %d out of range", image_to_add); return 1; }
  const long MN = (long)e->cfg.M * e->cfg.N;
  if (dphi_dev + MN * image_to_add == I_dev + MN * image_index) { gvm_set_error("gvm_prior_grad_add: dphi aliases the image"); return 1; }
  return prior_grad_launch(e, kind, I_dev, image_index, p, lambda, dphi_dev + MN * image_to_add, true);
}

int gvm_add_to_dphi(gvm_engine* e, float* dphi_dev, const float* dgi_dev, int index) {
  const long MN = (long)e->cfg.M * e->cfg.N;
  k_add_to_dphi<<<(int)((MN + kT - 1) / kT), kT, 0, e->stream>>>(dphi_dev + MN * index, dgi_dev, MN);
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  return 0;
}

int gvm_build_noise_image(gvm_engine* e, float noise_jypix, float* fg_scale_out) {
  if (e->chans.empty()) { gvm_set_error("gvm_build_noise_image: add the visibility blocks first"); return 1; }
  // no field list given: one attenuation pattern per distinct (pointing centre, beam model) of the uploaded blocks
  std::vector<gvm_channel_desc> fields;
  for (auto& c : e->chans) {
    bool seen = false;
    for (auto& f : fields)
      if (f.ref_xobs_pix == c.d.ref_xobs_pix && f.ref_yobs_pix == c.d.ref_yobs_pix &&
          f.antenna_diameter == c.d.antenna_diameter && f.pb_factor == c.d.pb_factor &&
          f.pb_cutoff == c.d.pb_cutoff && f.primary_beam == c.d.primary_beam) seen = true;
    if (!seen) fields.push_back(c.d);
  }
  return gvm_build_noise_image_fields(e, noise_jypix, (int)fields.size(), fields.data(), fg_scale_out);
}

int gvm_build_noise_image_fields(gvm_engine* e, float noise_jypix, int nfields, const gvm_channel_desc* fields_in,
                                 float* fg_scale_out) {
  GVM_CUDA(cudaSetDevice(e->cfg.device));
  const gvm_config& g = e->cfg;
  const long MN = g.M * g.N;
  const int blocks = (int)((MN + kT - 1) / kT);
  if (nfields < 1 || !fields_in) { gvm_set_error("gvm_build_noise_image_fields: no fields"); return 1; }
  const std::vector<gvm_channel_desc> fields(fields_in, fields_in + nfields);
  float* weight = nullptr;
  GVM_CUDA(cudaMalloc(&weight, MN * sizeof(float)));
  GVM_CUDA(cudaMemsetAsync(weight, 0, MN * sizeof(float), e->stream));
  for (auto& f : fields) {
    k_weight_accum<<<blocks, kT, 0, e->stream>>>(weight, g.N, g.M, f.antenna_diameter, f.pb_factor,
                                                 f.pb_cutoff, g.nu_0, f.ref_xobs_pix, f.ref_yobs_pix,
                                                 g.DELTAX, g.DELTAY, f.primary_beam);
    GVM_LAUNCH(e);
  }
  RedBuf r = aux_red(e);
  k_max_reduce<<<red_grid(e, MN), kT, 0, e->stream>>>(weight, MN, r.partials, r.pmax, r.counter, r.out);
  GVM_LAUNCH(e);
  double v[3];
  if (fetch_red(e, v)) { cudaFree(weight); return 1; }
  const float max_weight = (float)v[2];
  e->plan_dirty = true;
  e->epoch++;
  k_noise_image<<<blocks, kT, 0, e->stream>>>(e->noise, weight, MN, max_weight, noise_jypix);
  GVM_LAUNCH(e);
  float* d_min = nullptr;
  GVM_CUDA(cudaMalloc(&d_min, sizeof(float)));
  const float inf = INFINITY;
  GVM_CUDA(cudaMemcpyAsync(d_min, &inf, sizeof(float), cudaMemcpyHostToDevice, e->stream));
  k_min_reduce<<<red_grid(e, MN), kT, 0, e->stream>>>(e->noise, MN, d_min);
  GVM_LAUNCH(e);
  float mn = 0.f;
  GVM_CUDA(cudaMemcpyAsync(&mn, d_min, sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  GVM_CUDA(cudaStreamSynchronize(e->stream));
  cudaFree(weight);
  cudaFree(d_min);
  if (fg_scale_out) *fg_scale_out = mn;
  return 0;
}

int gvm_vec_evaluate_xt(gvm_engine* e, float* xt, const float* pcom, const float* xicom, float x,
                        int image_count, int nopositivity) {
  const long MN = (long)e->cfg.M * e->cfg.N;
  const long n = MN * image_count;
  k_evaluate_xt<<<(int)((n + kT - 1) / kT), kT, 0, e->stream>>>(
      xt, pcom, xicom, x, MN, image_count, -1.0f * e->cfg.eta * e->cfg.minpix, nopositivity);
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  return 0;
}
int gvm_vec_new_p(gvm_engine* e, float* p, float* xi, float xmin, int image_count, int nopositivity) {
  const long MN = (long)e->cfg.M * e->cfg.N;
  const long n = MN * image_count;
  k_new_p<<<(int)((n + kT - 1) / kT), kT, 0, e->stream>>>(p, xi, xmin, MN, image_count,
                                                         -1.0f * e->cfg.eta * e->cfg.minpix, nopositivity);
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  return 0;
}
int gvm_vec_dot(gvm_engine* e, const float* a, const float* b, int64_t n, float* out) {
  RedBuf r = aux_red(e);
  k_dot<<<red_grid(e, n), kT, 0, e->stream>>>(a, b, n, r.partials, r.pmax, r.counter, r.out);
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  double v[3];
  if (fetch_red(e, v)) return 1;
  *out = (float)v[0];
  return 0;
}
int gvm_vec_gg_dgg(gvm_engine* e, const float* xi, const float* g, int image_count, float* gg, float* dgg) {
  const long n = (long)e->cfg.M * e->cfg.N * image_count;
  RedBuf r = aux_red(e);
  k_gg_dgg<<<red_grid(e, n), kT, 0, e->stream>>>(xi, g, n, r.partials, r.pmax, r.counter, r.out);
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  double v[3];
  if (fetch_red(e, v)) return 1;
  *gg = (float)v[0];
  *dgg = (float)v[1];
  return 0;
}
int gvm_vec_grad_condition(gvm_engine* e, const float* xi, const float* p, float den, int image_count, float* gmax) {
  const long n = (long)e->cfg.M * e->cfg.N * image_count;
  RedBuf r = aux_red(e);
  k_grad_condition<<<red_grid(e, n), kT, 0, e->stream>>>(xi, p, den, n, r.partials, r.pmax, r.counter, r.out);
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  double v[3];
  if (fetch_red(e, v)) return 1;
  *gmax = (float)v[2];
  return 0;
}
int gvm_vec_new_xi(gvm_engine* e, float* g, float* xi, float* h, float gam, int image_count) {
  const long n = (long)e->cfg.M * e->cfg.N * image_count;
  k_new_xi<<<(int)((n + kT - 1) / kT), kT, 0, e->stream>>>(g, xi, h, gam, n);
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  return 0;
}
int gvm_vec_axpby(gvm_engine* e, float a, const float* x, float b, float* y, int64_t n) {
  k_axpby<<<(int)((n + kT - 1) / kT), kT, 0, e->stream>>>(a, x, b, y, n);
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  return 0;
}

/* normArray + deviceMaxReduce (src/lbfgs.cu:151-160): max |v| over n floats. */
int gvm_vec_absmax(gvm_engine* e, const float* v, int64_t n, float* out) {
  RedBuf r = aux_red(e);
  k_absmax<<<red_grid(e, n), kT, 0, e->stream>>>(v, n, r.partials, r.pmax, r.counter, r.out);
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  double v3[3];
  if (fetch_red(e, v3)) return 1;
  *out = (float)v3[2];
  return 0;
}
/* searchDirection_LBFGS (src/functions.cu:3564): v *= s. */
int gvm_vec_scale(gvm_engine* e, float* v, float s, int64_t n) {
  k_scale<<<(int)((n + kT - 1) / kT), kT, 0, e->stream>>>(v, s, n);
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  return 0;
}
/* calculateSandY (src/functions.cu:3636): y = xi - (-xi_old), s = p - p_old over n floats. */
int gvm_vec_lbfgs_sy(gvm_engine* e, float* y_out, float* s_out, const float* xi, const float* xi_old,
                     const float* p, const float* p_old, int64_t n) {
  k_lbfgs_sy<<<(int)((n + kT - 1) / kT), kT, 0, e->stream>>>(y_out, s_out, xi, xi_old, p, p_old, n);
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
