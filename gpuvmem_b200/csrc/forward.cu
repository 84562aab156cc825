// forward.cu — image -> visibility forward model and the chi2 reduction.
//
// Reference path: clip2IWNoise, calculateInu, apply_beam2I,
// apply_GCF, cufftExecC2C(INVERSE), phase_rotate, vis_mod, residual, chi2Vector,
// deviceReduce (src/functions.cu:4323-4454).  Here:
//   k_prep_channel   once per upload: hermitianSymmetry + metres->lambda + the
//                    static part of vis_mod (cell index, fractions, OOB weights)
//                    + fixed-point phase increments for the gradient
//   k_image_prep     clip (first channel only) + I_nu + beam + GCF -> complex grid
//   cuFFT            dense 2-D inverse C2C (the one library call the spec allows)
//   k_phase_rotate   post-FFT modulation
//   k_degrid_tiled   the degridder: samples are sorted by uv TILE at upload (32 x 32 cells; stable hand-written
//                    radix sort, sort.cu), a block stages the tile of the model grid (+ halo) in shared memory
//                    once and streams its samples' SoA arrays through a double-buffered shared-memory ring
//                    filled by 1-D bulk asynchronous copies (cp.async.bulk + mbarrier, TMA without a tensor
//                    map); taps are gathered from shared memory: bilinear vis_mod or the CKernel sum
//   k_degrid_chi2    untiled fallback (grids that are not a multiple of the tile, the half-plane model of
//                    gridded data whose samples are already in cell order): 4-tap gather from global memory
//   both             + residual + w|Vr|^2, warp-shuffle block partials, last block finishes the sum in fp64
//                    in a fixed order
// All streaming kernels are HBM-bound: 24 B read + 8 B written per visibility (16 + 8 with a CKernel),
// grids sized in multiples of the SM count.
#include "gvm_internal.cuh"
#include "gvm_ptx.cuh"

namespace {

constexpr int kVisThreads = 256;
constexpr int kVisPerThread = 4;

// ---------------------------------------------------------------------------
// Upload-time preprocessing. One thread per visibility.
// hermitianSymmetry: src/functions.cu:2256-2273 (w is NOT negated there).
// vis_mod static part: src/functions.cu:2569-2586, 2607.
// Position in the device arrays: p. Sample it holds: k = perm ? perm[p] : p (tile-sorted upload). uvw_l stays
// in the caller's order (read-back, fallback conv degridding); everything the kernels stream is in position order.
__global__ void __launch_bounds__(256) k_prep_channel(
    const double* __restrict__ uvw_m, const float2* __restrict__ Vo_in,
    const float* __restrict__ w_in, const uint32_t* __restrict__ perm, float freq, double deltau, double deltav, double dx_turn,
    double dy_turn, long N, long Z, double* __restrict__ uvw_l, uint32_t* __restrict__ cell,
    float2* __restrict__ frac, uint32_t* __restrict__ ccell, float2* __restrict__ Vo, float* __restrict__ w,
    uint64_t* __restrict__ du64, uint64_t* __restrict__ dv64, float* __restrict__ wz,
    float* __restrict__ max_abs_wz, unsigned long long* __restrict__ offgrid) {
  const long p = blockIdx.x * (long)blockDim.x + threadIdx.x;
  float my_wz = 0.f;
  bool off = false;
  if (p < Z) {
    const long k = perm ? (long)perm[p] : p;
    double um = uvw_m[3 * k], vm = uvw_m[3 * k + 1], wm = uvw_m[3 * k + 2];
    float2 vo = Vo_in[k];
    if (um > 0.0) {
      um *= -1.0;
      vm *= -1.0;
      vo.y *= -1.0f;
    }
    double u = gvm_metres_to_lambda(um, freq);
    double v = gvm_metres_to_lambda(vm, freq);
    double wl = gvm_metres_to_lambda(wm, freq);
    uvw_l[3 * k] = u;
    uvw_l[3 * k + 1] = v;
    uvw_l[3 * k + 2] = wl;
    Vo[p] = vo;

    double uv_x = u / deltau;
    double uv_y = v / deltav;
    // centre cell of the convolutional degridder (degriddingGPU, src/functions.cu:2222-2223), CENTRED grid
    // coordinates; kept for on-grid samples only (0xFFFFFFFF otherwise: the kernel falls back to uvw_l)
    const int half = (int)(N / 2);
    const int jc = (int)(uv_x + (double)half + 0.5);
    const int kc = (int)(uv_y + (double)half + 0.5);
    if (uv_x < 0.0) uv_x += N;
    if (uv_y < 0.0) uv_y += N;
    const int i1 = __double2int_rd(uv_x);
    const int j1 = __double2int_rd(uv_y);
    const double du = uv_x - i1;
    const double dv = uv_y - j1;
    float wk = w_in[k];
    if (i1 >= 0 && i1 < N && j1 >= 0 && j1 < N) {
      cell[p] = (uint32_t)i1 | ((uint32_t)j1 << 16);
      frac[p] = make_float2((float)du, (float)dv);
      ccell[p] = (jc >= 0 && jc < 65536 && kc >= 0 && kc < 65536) ? ((uint32_t)jc | ((uint32_t)kc << 16)) : GVM_CELL_INVALID;
    } else {
      cell[p] = GVM_CELL_INVALID;
      frac[p] = make_float2(0.f, 0.f);
      ccell[p] = GVM_CELL_INVALID;
      wk = 0.0f;  // vis_mod: weight[i] = 0 for samples that fall off the grid
    }
    w[p] = wk;

    // phase increment per pixel step, as a 0.64 fixed-point fraction of a turn
    double tu = u * dx_turn;
    double tv = v * dy_turn;
    tu -= floor(tu);
    tv -= floor(tv);
    du64[p] = __double2ull_rd(tu * 18446744073709551616.0);
    dv64[p] = __double2ull_rd(tv * 18446744073709551616.0);
    wz[p] = (float)wl;
    my_wz = fabsf((float)wl);
    // Is the sample the centre of a uv cell with w = 0 (output of do_gridding)? Then its phase
    // step per pixel is an integer multiple of 1/N turns and the DFT gradient is an FFT.
    const double gu = tu * (double)N, gv = tv * (double)N;
    off = fabs(gu - rint(gu)) > 1e-6 || fabs(gv - rint(gv)) > 1e-6 || wl != 0.0;
  }
  const unsigned any_off = __ballot_sync(0xffffffffu, off);
  if ((threadIdx.x & 31) == 0 && any_off) atomicAdd(offgrid, (unsigned long long)__popc(any_off));
  my_wz = gvm_warp_max(my_wz);
  if ((threadIdx.x & 31) == 0 && my_wz > 0.f)
    atomicMax(reinterpret_cast<int*>(max_abs_wz), __float_as_int(my_wz));  // non-negative floats order as ints
}

// ---------------------------------------------------------------------------
// Tile-sorted upload. Key of a sample: the 32 x 32-cell tile of its vis_mod cell (same fp64 arithmetic as
// k_prep_channel), or `ntiles` for samples that fall off the grid (they form the last bucket).
constexpr int kTile = 32;            // cells per tile edge
constexpr int kChunkV = 384;         // samples per work item of the tiled degridder
constexpr int kStages = 4;           // shared-memory stages: the streams of the next kStages - 1 items are in flight
constexpr int kAlignV = 16;          // an item's streams are bulk-copied from a 16-sample boundary
constexpr int kStageCap = kChunkV + kAlignV;   // samples a stage can hold

__global__ void __launch_bounds__(256) k_tile_keys(const double* __restrict__ uvw_m, float freq, double deltau,
                                                   double deltav, long N, long Z, int ntx, uint32_t ntiles,
                                                   uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                   uint32_t* __restrict__ counts) {
  const long k = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (k >= Z) return;
  double um = uvw_m[3 * k], vm = uvw_m[3 * k + 1];
  if (um > 0.0) { um *= -1.0; vm *= -1.0; }
  double uv_x = gvm_metres_to_lambda(um, freq) / deltau;
  double uv_y = gvm_metres_to_lambda(vm, freq) / deltav;
  if (uv_x < 0.0) uv_x += N;
  if (uv_y < 0.0) uv_y += N;
  const int i1 = __double2int_rd(uv_x), j1 = __double2int_rd(uv_y);
  uint32_t key = ntiles;
  if (i1 >= 0 && i1 < N && j1 >= 0 && j1 < N) key = (uint32_t)(j1 / kTile) * (uint32_t)ntx + (uint32_t)(i1 / kTile);
  keys[k] = key;
  vals[k] = (uint32_t)k;
  atomicAdd(&counts[key], 1u);
}
// counts -> number of work items (chunks of kChunkV samples) per bucket
__global__ void __launch_bounds__(256) k_item_counts(const uint32_t* __restrict__ counts, uint32_t nbuckets,
                                                     uint32_t* __restrict__ nitems) {
  const uint32_t t = blockIdx.x * 256u + threadIdx.x;
  if (t < nbuckets) nitems[t] = (counts[t] + kChunkV - 1) / kChunkV;
}
// (bucket, first sample, samples) of every work item, in bucket order
__global__ void __launch_bounds__(256) k_make_items(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ start,
                                                    const uint32_t* __restrict__ item_start, uint32_t nbuckets,
                                                    uint4* __restrict__ items) {
  const uint32_t t = blockIdx.x * 256u + threadIdx.x;
  if (t >= nbuckets) return;
  const uint32_t c = counts[t], s0 = start[t], i0 = item_start[t];
  for (uint32_t k = 0, i = 0; k < c; k += kChunkV, i++)
    items[i0 + i] = make_uint4(t, s0 + k, min(c - k, (uint32_t)kChunkV), 0u);
}
// out[perm[p]] = in[p]: back to the caller's sample order (gvm_get_vis)
template <typename T>
__global__ void __launch_bounds__(256) k_unpermute(const T* __restrict__ in, const uint32_t* __restrict__ perm, long Z,
                                                   T* __restrict__ out) {
  const long p = blockIdx.x * 256L + threadIdx.x;
  if (p < Z) out[perm[p]] = in[p];
}

// ---------------------------------------------------------------------------
// clip2IWNoise (src/functions.cu:2694-2719) fused with calculateInu (:3939-3966),
// apply_beam2I (:2424-2444) and apply_GCF (:2468-2476). One thread per pixel.
// kReal: the pre-FFT image is written as a real plane (half-plane forward model, see k_degrid_chi2).
__global__ void __launch_bounds__(256) k_atten_image(float* __restrict__ out, long N, long M, float D,
                                                     float pb_factor, float pb_cutoff, float nu, float xobs,
                                                     float yobs, double DELTAX, double DELTAY, int primary_beam) {
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  out[idx] = gvm_attenuation((int)(idx / N), (int)(idx % N), D, pb_factor, pb_cutoff, nu, xobs, yobs, DELTAX,
                             DELTAY, primary_beam);
}

struct PrepArgs {
  float* I;
  const float* noise;
  const float* gcf;
  const float* atten_plane;
  void* out;
  long N, M;
  float noise_cut, minpix, eta, threshold, nu, nu_0, fg_scale, D, pb_factor, pb_cutoff, xobs, yobs;
  double DELTAX, DELTAY;
  int schedule, primary_beam;
};
// one pixel of clip2IWNoise + calculateInu + apply_beam2I + apply_GCF; I0 / alpha are updated by the clip
template <bool kClip>
__device__ __forceinline__ float prep_pixel(const PrepArgs& a, long idx, float& I0, float& alpha, float noise,
                                            float atten, float gcf, bool& touched) {
  if (kClip) {
    if (noise > a.noise_cut) {
      I0 = (a.eta > 0.0f) ? 0.0f : -1.0f * a.eta * a.minpix;
      alpha = 0.0f;
      touched = true;
    } else if (I0 < a.threshold && a.schedule > 0) {
      alpha = 0.0f;
      touched = true;
    }
  }
  const float nudiv = a.nu / a.nu_0;
  float v = I0 * powf(nudiv, alpha);
  const float floor_v = -1.0f * a.eta * a.minpix;
  if (v < floor_v) v = floor_v;
  v = v * atten * a.fg_scale;
  if (a.gcf != nullptr) v = v * gcf;
  return v;
}
__device__ __forceinline__ float prep_atten(const PrepArgs& a, long idx) {
  return gvm_attenuation((int)(idx / a.N), (int)(idx % a.N), a.D, a.pb_factor, a.pb_cutoff, a.nu, a.xobs, a.yobs, a.DELTAX,
                         a.DELTAY, a.primary_beam);
}

// scalar version: any image size / alignment
template <bool kClip, bool kReal>
__global__ void __launch_bounds__(256) k_image_prep(PrepArgs a) {
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long MN = a.M * a.N;
  if (idx >= MN) return;
  float I0 = a.I[idx], alpha = a.I[MN + idx];
  bool touched = false;
  const float atten = a.atten_plane ? a.atten_plane[idx] : prep_atten(a, idx);
  const float v = prep_pixel<kClip>(a, idx, I0, alpha, kClip ? a.noise[idx] : 0.f, atten, a.gcf ? a.gcf[idx] : 1.f, touched);
  if (kClip && touched) { a.I[idx] = I0; a.I[MN + idx] = alpha; }
  if (kReal) reinterpret_cast<float*>(a.out)[idx] = v;
  else reinterpret_cast<float2*>(a.out)[idx] = make_float2(v, 0.0f);
}
// four pixels per thread, 128-bit loads and stores (M N a multiple of 4, planes 16-byte aligned)
template <bool kClip, bool kReal>
__global__ void __launch_bounds__(256) k_image_prep4(PrepArgs a) {
  const long q = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long MN = a.M * a.N, idx = 4 * q;
  if (idx >= MN) return;
  float4 I0 = reinterpret_cast<const float4*>(a.I)[q];
  float4 al = reinterpret_cast<const float4*>(a.I + MN)[q];
  float4 nz = make_float4(0.f, 0.f, 0.f, 0.f), at, gc = make_float4(1.f, 1.f, 1.f, 1.f);
  if (kClip) nz = __ldg(reinterpret_cast<const float4*>(a.noise) + q);
  if (a.atten_plane) at = __ldg(reinterpret_cast<const float4*>(a.atten_plane) + q);
  else at = make_float4(prep_atten(a, idx), prep_atten(a, idx + 1), prep_atten(a, idx + 2), prep_atten(a, idx + 3));
  if (a.gcf) gc = __ldg(reinterpret_cast<const float4*>(a.gcf) + q);
  bool touched = false;
  float4 v;
  v.x = prep_pixel<kClip>(a, idx, I0.x, al.x, nz.x, at.x, gc.x, touched);
  v.y = prep_pixel<kClip>(a, idx + 1, I0.y, al.y, nz.y, at.y, gc.y, touched);
  v.z = prep_pixel<kClip>(a, idx + 2, I0.z, al.z, nz.z, at.z, gc.z, touched);
  v.w = prep_pixel<kClip>(a, idx + 3, I0.w, al.w, nz.w, at.w, gc.w, touched);
  if (kClip && touched) {
    reinterpret_cast<float4*>(a.I)[q] = I0;
    reinterpret_cast<float4*>(a.I + MN)[q] = al;
  }
  if (kReal) {
    reinterpret_cast<float4*>(a.out)[q] = v;
  } else {
    float4* o = reinterpret_cast<float4*>(a.out) + 2 * q;
    o[0] = make_float4(v.x, 0.0f, v.y, 0.0f);
    o[1] = make_float4(v.z, 0.0f, v.w, 0.0f);
  }
}

// phase_rotate: src/functions.cu:2483-2518.
__global__ void __launch_bounds__(256) k_phase_rotate(float2* __restrict__ data, long M, long N,
                                                     double xphs, double yphs) {
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  const int i = (int)(idx / N), j = (int)(idx % N);
  const double upix = xphs / (double)M;
  const double vpix = yphs / (double)N;
  float u, v;
  if (j < M / 2) u = upix * j; else u = upix * (j - M);
  if (i < N / 2) v = vpix * i; else v = vpix * (i - N);
  const float phase = -2.0f * (u + v);
  float s, c;
  sincospif(phase, &s, &c);
  const float2 d = data[idx];
  data[idx] = make_float2(d.x * c - d.y * s, d.x * s + d.y * c);  // cuCmulf
}

// ---------------------------------------------------------------------------
// vis_mod (dynamic part, src/functions.cu:2588-2606) + residual (:2663) +
// chi2Vector (:2867) + first reduction level. Each thread handles kVisPerThread
// consecutive-by-stride samples so that every load is coalesced.
//
// kConv: the model visibility is a convolutional-kernel degridding of the same grid instead of
// the bilinear vis_mod — the forward-model option the reference sketches in degriddingGPU
// (src/functions.cu:2205-2254, never launched): nearest cell j = int(u/deltau + N/2 + 0.5),
// k likewise (:2222-2223), Vm = sum over taps [-sy, sy] x [-sx, sx] of kernel[kn*(m+sy) + (n+sx)] *
// V_g[cell + tap], taps outside the centred grid skipped (:2231-2232). degriddingGPU reads a
// CENTRED grid and folds its left half through the Hermitian twin (:2235-2243); the engine's
// grid is the full plane with DC at [0,0] (cuFFT output, no shift), so the same cell is read
// directly at ((j - N/2) mod N, (k - M/2) mod N) — identical for the Hermitian grid of a real image.
struct GvmConvDegrid {
  const double* uvw_l;   // [Z][3] wavelengths (after the fold), in the CALLER's sample order
  const uint32_t* perm;  // position -> caller's sample index (tile-sorted upload) or null
  const float* table;    // [km][kn] kernel, device
  double deltau, deltav;
  int km, kn, sx, sy;
  double upix, vpix;     // half-plane mode: xphs / M, yphs / N of phase_rotate
};

enum { kGridFull = 0, kGridConv = 1, kGridHalf = 2 };

// kGridHalf — half-plane forward model for blocks with few samples per pixel (gridded data, 4 Z <= M N):
// the pre-FFT image is REAL (calculateInu writes imag = 0, src/functions.cu:3963), so its unnormalised
// inverse DFT is the conjugate of a real-to-complex forward transform, known on the half plane
// Vh[N][N/2+1]: V[r][c] = conj(Vh[r][c]) for c <= N/2, V[r][c] = Vh[(N-r)%N][N-c] beyond. cuFFT R2C moves
// half the bytes of the C2C transform, and phase_rotate (an image-sized pass) is applied to the four taps
// of every sample instead — the same float arithmetic per tap (src/functions.cu:2491-2517), 4 Z
// sincospif instead of M N.
__device__ __forceinline__ float2 gvm_fetch_half(const float2* __restrict__ Vh, int row, int col, int N,
                                                 double upix, double vpix) {
  const int NH = N / 2 + 1;
  float2 g;
  if (col <= N / 2) {
    g = __ldg(&Vh[(long)row * NH + col]);
    g.y = -g.y;
  } else {
    g = __ldg(&Vh[(long)(row ? N - row : 0) * NH + (N - col)]);
  }
  float u, v;
  if (col < N / 2) u = upix * col; else u = upix * (col - N);
  if (row < N / 2) v = vpix * row; else v = vpix * (row - N);
  const float phase = -2.0f * (u + v);
  float s, c;
  sincospif(phase, &s, &c);
  return make_float2(g.x * c - g.y * s, g.x * s + g.y * c);  // cuCmulf, as k_phase_rotate
}

template <bool kKeepVm, int kMode>
__global__ void __launch_bounds__(kVisThreads) k_degrid_chi2(
    const float2* __restrict__ V, const uint32_t* __restrict__ cell,
    const float2* __restrict__ frac, const float2* __restrict__ Vo, const float* __restrict__ w,
    float2* __restrict__ Vr, float2* __restrict__ Vm, long Z, int N,
    double* __restrict__ partials, float* __restrict__ partial_max,
    unsigned int* __restrict__ counter, double* __restrict__ out_sum, float* __restrict__ out_max,
    GvmConvDegrid cv) {
  __shared__ float s_sum[kVisThreads / 32];
  __shared__ float s_max[kVisThreads / 32];
  __shared__ bool s_last;
  constexpr bool kConv = kMode == kGridConv;
  __shared__ float s_tab[kConv ? GVM_MAX_CKERNEL : 1];
  if (kConv) {
    for (int t = threadIdx.x; t < cv.km * cv.kn; t += blockDim.x) s_tab[t] = cv.table[t];
    __syncthreads();
  }
  float acc = 0.f, mx = 0.f;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long k = blockIdx.x * (long)blockDim.x + threadIdx.x; k < Z; k += stride) {
    const uint32_t c = __ldg(&cell[k]);
    const float2 f = __ldg(&frac[k]);
    const float2 vo = __ldg(&Vo[k]);
    const float wk = __ldg(&w[k]);
    float2 vm = make_float2(0.f, 0.f);
    if (kConv) {
      const int half = N / 2;
      const long ko = cv.perm ? (long)cv.perm[k] : k;
      const int jc = (int)(cv.uvw_l[3 * ko] / cv.deltau + (double)half + 0.5);
      const int kc = (int)(cv.uvw_l[3 * ko + 1] / cv.deltav + (double)half + 0.5);
      for (int m = -cv.sy; m <= cv.sy; m++) {
        const int sk = kc + m;
        if (sk < 0 || sk >= N) continue;
        const int row = sk >= half ? sk - half : sk + N - half;      // (sk - N/2) mod N
        for (int n = -cv.sx; n <= cv.sx; n++) {
          const int sj = jc + n;
          if (sj < 0 || sj >= N) continue;
          const int col = sj >= half ? sj - half : sj + N - half;
          const float kv = s_tab[cv.kn * (m + cv.sy) + (n + cv.sx)];
          const float2 g = __ldg(&V[(long)N * row + col]);
          vm.x += kv * g.x;
          vm.y += kv * g.y;
        }
      }
    } else if (c != GVM_CELL_INVALID) {
      const int i1 = (int)(c & 0xFFFFu), j1 = (int)(c >> 16);
      const int i2 = (i1 + 1 == N) ? 0 : i1 + 1;
      const int j2 = (j1 + 1 == N) ? 0 : j1 + 1;
      float2 v11, v12, v21, v22;
      if (kMode == kGridHalf) {
        v11 = gvm_fetch_half(V, j1, i1, N, cv.upix, cv.vpix);
        v12 = gvm_fetch_half(V, j2, i1, N, cv.upix, cv.vpix);
        v21 = gvm_fetch_half(V, j1, i2, N, cv.upix, cv.vpix);
        v22 = gvm_fetch_half(V, j2, i2, N, cv.upix, cv.vpix);
      } else {
        v11 = __ldg(&V[(long)N * j1 + i1]);
        v12 = __ldg(&V[(long)N * j2 + i1]);
        v21 = __ldg(&V[(long)N * j1 + i2]);
        v22 = __ldg(&V[(long)N * j2 + i2]);
      }
      const float du = f.x, dv = f.y;
      const float w11 = (1.0f - du) * (1.0f - dv);
      const float w12 = (1.0f - du) * dv;
      const float w21 = du * (1.0f - dv);
      const float w22 = du * dv;
      vm.x = w11 * v11.x + w12 * v12.x + w21 * v21.x + w22 * v22.x;
      vm.y = w11 * v11.y + w12 * v12.y + w21 * v21.y + w22 * v22.y;
    }
    const float2 vr = make_float2(vo.x - vm.x, vo.y - vm.y);
    Vr[k] = vr;
    if (kKeepVm) Vm[k] = vm;
    acc += wk * (vr.x * vr.x + vr.y * vr.y);
    mx = fmaxf(mx, wk * fmaxf(fabsf(vr.x), fabsf(vr.y)));
  }
  acc = gvm_warp_sum(acc);
  mx = gvm_warp_max(mx);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_sum[warp] = acc; s_max[warp] = mx; }
  __syncthreads();
  if (warp == 0) {
    float a = (lane < kVisThreads / 32) ? s_sum[lane] : 0.f;
    float m = (lane < kVisThreads / 32) ? s_max[lane] : 0.f;
    a = gvm_warp_sum(a);
    m = gvm_warp_max(m);
    if (lane == 0) {
      partials[blockIdx.x] = (double)a;
      partial_max[blockIdx.x] = m;
      __threadfence();
      const unsigned int done = atomicAdd(counter, 1u);
      s_last = (done == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (s_last && warp == 0) {
    // fixed-order fp64 finish: deterministic for a given grid size
    __threadfence();
    double t = 0.0;
    float m = 0.f;
    for (unsigned int b = lane; b < gridDim.x; b += 32) {
      t += partials[b];
      m = fmaxf(m, partial_max[b]);
    }
    t = gvm_warp_sum_d(t);
    m = gvm_warp_max(m);
    if (lane == 0) {
      *out_sum = t;
      *out_max = m;
      *counter = 0u;
    }
  }
}

// ---------------------------------------------------------------------------
// The tiled degridder. Work item = (bucket, first position, count <= kChunkV) over the tile-sorted arrays; a block
// owns a contiguous run of items, so consecutive items of one tile reuse the staged grid tile. Per item one thread
// posts the bulk copies of the NEXT item's streams into the other stage (expect_tx on that stage's mbarrier) and the
// block then consumes the current stage; copies start at the 16-sample boundary below the item so that every address
// and size is a multiple of 16 bytes (the few foreign samples in front are skipped). Per-sample arithmetic is the
// same as k_degrid_chi2's, tap for tap.
struct TiledArgs {
  const float2* V;
  const uint32_t* cell;     // bilinear: i1 | j1 << 16
  const uint32_t* ccell;    // CKernel: jc | kc << 16, centred grid coordinates
  const float2* frac;
  const float2* Vo;
  const float* w;
  float2* Vr;
  float2* Vm;
  const uint4* items;
  const uint32_t* block_first;   // [blocks + 1] first item of every block (cost-balanced at upload)
  int N, ntx;
  uint32_t invalid_bucket;
};

template <bool kKeepVm, bool kConv>
__global__ void __launch_bounds__(256) k_degrid_tiled(TiledArgs a, double* __restrict__ partials,
                                                     float* __restrict__ partial_max, unsigned int* __restrict__ counter,
                                                     double* __restrict__ out_sum, float* __restrict__ out_max,
                                                     GvmConvDegrid cv) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ float s_sum[8], s_max[8];
  __shared__ bool s_last;
  __shared__ __align__(8) uint64_t s_bar[kStages];
  // stage layout (per stage): idx u32[cap] | frac float2[cap] (bilinear only) | Vo float2[cap] | w float[cap]
  constexpr int kStageBytes = kStageCap * (kConv ? 16 : 24);
  constexpr int kTileRegs = kConv ? 1 : ((kTile + 1) * (kTile + 1) + 255) / 256;   // bilinear tile elements per thread
  const int lo_x = kConv ? cv.sx : 0, lo_y = kConv ? cv.sy : 0;
  const int tw = kTile + (kConv ? 2 * cv.sx + 1 : 1), th = kTile + (kConv ? 2 * cv.sy + 1 : 1);
  float2* s_tile = reinterpret_cast<float2*>(smem + kStages * kStageBytes);
  float* s_tab = reinterpret_cast<float*>(s_tile + (size_t)tw * th);
  auto stage_idx = [&](int st) { return reinterpret_cast<uint32_t*>(smem + st * kStageBytes); };
  auto stage_frac = [&](int st) { return reinterpret_cast<float2*>(smem + st * kStageBytes + kStageCap * 4); };
  auto stage_vo = [&](int st) { return reinterpret_cast<float2*>(smem + st * kStageBytes + kStageCap * (kConv ? 4 : 12)); };
  auto stage_w = [&](int st) { return reinterpret_cast<float*>(smem + st * kStageBytes + kStageCap * (kConv ? 12 : 20)); };
  const int tid = threadIdx.x;
  const int N = a.N;
  if (tid == 0) {
    for (int st = 0; st < kStages; st++) gvmptx::mbar_init(gvmptx::smem_addr(&s_bar[st]), 1);
    gvmptx::fence_barrier_init();
  }
  if (kConv)
    for (int t = tid; t < cv.km * cv.kn; t += 256) s_tab[t] = cv.table[t];
  __syncthreads();

  const int it0 = (int)a.block_first[blockIdx.x], it1 = (int)a.block_first[blockIdx.x + 1];
  auto post = [&](int it) {   // one thread: bulk copies of item `it` into stage (it - it0) % kStages
    const uint4 w = a.items[it];
    const int st = (it - it0) % kStages;
    const uint32_t a0 = w.y & ~(uint32_t)(kAlignV - 1);
    const uint32_t n = (w.y + w.z - a0 + kAlignV - 1) & ~(uint32_t)(kAlignV - 1);
    const uint32_t bar = gvmptx::smem_addr(&s_bar[st]);
    gvmptx::mbar_expect_tx(bar, n * (kConv ? 16u : 24u));
    gvmptx::bulk_g2s(gvmptx::smem_addr(stage_idx(st)), (kConv ? a.ccell : a.cell) + a0, n * 4u, bar);
    if (!kConv) gvmptx::bulk_g2s(gvmptx::smem_addr(stage_frac(st)), a.frac + a0, n * 8u, bar);
    gvmptx::bulk_g2s(gvmptx::smem_addr(stage_vo(st)), a.Vo + a0, n * 8u, bar);
    gvmptx::bulk_g2s(gvmptx::smem_addr(stage_w(st)), a.w + a0, n * 4u, bar);
  };
  if (tid == 0)
    for (int it = it0; it < it1 && it < it0 + kStages - 1; it++) post(it);

  // the grid tile (+ halo) of bucket b; rows and columns wrap (the grid is periodic: DC at [0,0])
  auto tile_elem = [&](uint32_t b, int t) -> float2 {
    const int bx = (int)(b % (uint32_t)a.ntx) * kTile, by = (int)(b / (uint32_t)a.ntx) * kTile;
    const int lr = t / tw, lc = t - lr * tw;
    int row = by - lo_y + lr, col = bx - lo_x + lc;
    row += row < 0 ? N : 0; row -= row >= N ? N : 0;
    col += col < 0 ? N : 0; col -= col >= N ? N : 0;
    return __ldg(&a.V[(long)N * row + col]);
  };

  float acc = 0.f, mx = 0.f;
  uint32_t staged_bucket = 0xFFFFFFFFu;
  int x0 = 0, y0 = 0;
  uint4 wi = it0 < it1 ? a.items[it0] : make_uint4(0u, 0u, 0u, 0u);
  for (int it = it0; it < it1; it++) {
    const int st = (it - it0) % kStages;
    // the stage consumed in the previous iteration is free (barrier at its end): refill it kStages - 1 items ahead
    if (tid == 0 && it + kStages - 1 < it1) post(it + kStages - 1);
    const uint4 wn = it + 1 < it1 ? a.items[it + 1] : wi;
    if (wi.x != staged_bucket && wi.x != a.invalid_bucket) {   // first item of the block (or after the off-grid bucket)
      for (int t = tid; t < tw * th; t += 256) s_tile[t] = tile_elem(wi.x, t);
      staged_bucket = wi.x;
      x0 = (int)(wi.x % (uint32_t)a.ntx) * kTile;
      y0 = (int)(wi.x / (uint32_t)a.ntx) * kTile;
      __syncthreads();
    }
    // bilinear: the NEXT item's tile travels in registers while this item is computed (its latency hides behind the
    // work below); it is written to shared memory after the barrier that ends this item
    const bool next_tile = !kConv && wn.x != staged_bucket && wn.x != a.invalid_bucket && it + 1 < it1;
    float2 tr[kTileRegs];
    if (next_tile) {
#pragma unroll
      for (int r = 0; r < kTileRegs; r++) {
        const int t = tid + 256 * r;
        tr[r] = t < tw * th ? tile_elem(wn.x, t) : make_float2(0.f, 0.f);
      }
    }
    gvmptx::mbar_wait(gvmptx::smem_addr(&s_bar[st]), (uint32_t)(((it - it0) / kStages) & 1));
    const uint32_t skip = wi.y & (uint32_t)(kAlignV - 1);
    const uint32_t* s_idx = stage_idx(st) + skip;
    const float2* s_fr = stage_frac(st) + skip;
    const float2* s_vo = stage_vo(st) + skip;
    const float* s_w = stage_w(st) + skip;
    for (uint32_t t = tid; t < wi.z; t += 256) {
      const uint32_t c = s_idx[t];
      const float2 vo = s_vo[t];
      const float wk = s_w[t];
      float2 vm = make_float2(0.f, 0.f);
      if (kConv) {
        const int half = N / 2;
        int jc, kc;
        if (c != GVM_CELL_INVALID) {
          jc = (int)(c & 0xFFFFu); kc = (int)(c >> 16);
        } else {   // off-grid sample: centre from the fp64 coordinates, taps from global memory below
          const long ko = cv.perm ? (long)cv.perm[wi.y + t] : (long)(wi.y + t);
          jc = (int)(cv.uvw_l[3 * ko] / cv.deltau + (double)half + 0.5);
          kc = (int)(cv.uvw_l[3 * ko + 1] / cv.deltav + (double)half + 0.5);
        }
        const bool in_tile = wi.x != a.invalid_bucket && c != GVM_CELL_INVALID;
        // centre in DC-origin coordinates relative to the staged tile
        const int colc = jc - half;                                   // >= 0 for folded samples
        const int rowc = kc >= half ? kc - half : kc + N - half;
        int lr0 = rowc - y0; lr0 += lr0 < 0 ? N : 0;                  // rowc in {j1, j1 + 1 (mod N)}
        const int lc0 = colc - x0 + lo_x, lrb = lr0 + lo_y;
        for (int m = -cv.sy; m <= cv.sy; m++) {
          const int sk = kc + m;
          if (sk < 0 || sk >= N) continue;
          for (int n = -cv.sx; n <= cv.sx; n++) {
            const int sj = jc + n;
            if (sj < 0 || sj >= N) continue;
            const float kv = s_tab[cv.kn * (m + cv.sy) + (n + cv.sx)];
            float2 g;
            if (in_tile) {
              g = s_tile[(lrb + m) * tw + (lc0 + n)];
            } else {
              const int row = sk >= half ? sk - half : sk + N - half;
              const int col = sj >= half ? sj - half : sj + N - half;
              g = __ldg(&a.V[(long)N * row + col]);
            }
            vm.x += kv * g.x;
            vm.y += kv * g.y;
          }
        }
      } else if (c != GVM_CELL_INVALID) {
        const int li = (int)(c & 0xFFFFu) - x0, lj = (int)(c >> 16) - y0;
        const float2 f = s_fr[t];
        const float2 v11 = s_tile[lj * tw + li], v12 = s_tile[(lj + 1) * tw + li];
        const float2 v21 = s_tile[lj * tw + li + 1], v22 = s_tile[(lj + 1) * tw + li + 1];
        const float du = f.x, dv = f.y;
        const float w11 = (1.0f - du) * (1.0f - dv);
        const float w12 = (1.0f - du) * dv;
        const float w21 = du * (1.0f - dv);
        const float w22 = du * dv;
        vm.x = w11 * v11.x + w12 * v12.x + w21 * v21.x + w22 * v22.x;
        vm.y = w11 * v11.y + w12 * v12.y + w21 * v21.y + w22 * v22.y;
      }
      const float2 vr = make_float2(vo.x - vm.x, vo.y - vm.y);
      a.Vr[wi.y + t] = vr;
      if (kKeepVm) a.Vm[wi.y + t] = vm;
      acc += wk * (vr.x * vr.x + vr.y * vr.y);
      mx = fmaxf(mx, wk * fmaxf(fabsf(vr.x), fabsf(vr.y)));
    }
    __syncthreads();   // stage `st` and (if the next item changes bucket) the tile are free again
    if (next_tile) {
#pragma unroll
      for (int r = 0; r < kTileRegs; r++) {
        const int t = tid + 256 * r;
        if (t < tw * th) s_tile[t] = tr[r];
      }
      staged_bucket = wn.x;
      x0 = (int)(wn.x % (uint32_t)a.ntx) * kTile;
      y0 = (int)(wn.x / (uint32_t)a.ntx) * kTile;
      __syncthreads();
    }
    wi = wn;
  }
  acc = gvm_warp_sum(acc);
  mx = gvm_warp_max(mx);
  const int warp = tid >> 5, lane = tid & 31;
  if (lane == 0) { s_sum[warp] = acc; s_max[warp] = mx; }
  __syncthreads();
  if (warp == 0) {
    float sa = lane < 8 ? s_sum[lane] : 0.f;
    float sm = lane < 8 ? s_max[lane] : 0.f;
    sa = gvm_warp_sum(sa);
    sm = gvm_warp_max(sm);
    if (lane == 0) {
      partials[blockIdx.x] = (double)sa;
      partial_max[blockIdx.x] = sm;
      __threadfence();
      s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (s_last && warp == 0) {   // fixed-order fp64 finish: deterministic for a given upload
    __threadfence();
    double t = 0.0;
    float m = 0.f;
    for (unsigned int b = lane; b < gridDim.x; b += 32) {
      t += partials[b];
      m = fmaxf(m, partial_max[b]);
    }
    t = gvm_warp_sum_d(t);
    m = gvm_warp_max(m);
    if (lane == 0) {
      *out_sum = t;
      *out_max = m;
      *counter = 0u;
    }
  }
}

// Combines the per-block sums the way chi2() does on the host
// (src/functions.cu:4439-4453): float accumulation, optional /Z, 0.5f * total.
__global__ void k_chi2_combine(const double* __restrict__ sums, const long* __restrict__ Zs,
                               int nslots, int normalize, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float reduced = 0.0f;
    for (int s = 0; s < nslots; s++) {
      if (Zs[s] <= 0) continue;
      float r = (float)sums[s];
      if (normalize) r /= (float)Zs[s];
      reduced += r;
    }
    out[0] = (double)(0.5f * reduced);
  }
}

}  // namespace

// Tile-sort plan of one block: the permutation (position -> caller's index), bucket sizes and the work items of
// k_degrid_tiled. Nothing is sorted (perm stays null, every kernel sees the caller's order) when the grid is not
// a multiple of the tile edge, for empty blocks, or with GVM_FORWARD_UNTILED=1.
static int build_tile_plan(gvm_engine* e, GvmChannel& c, const double* uvw_m_dev, double deltau, double deltav) {
  const long N = e->cfg.N;
  static const bool disabled = [] { const char* s = getenv("GVM_FORWARD_UNTILED"); return s && *s == '1'; }();
  if (disabled || N % kTile != 0 || c.Z <= 0 || c.Z >= ((int64_t)1 << 31)) return 0;
  const int ntx = (int)(N / kTile);
  const uint32_t ntiles = (uint32_t)ntx * (uint32_t)ntx, nb = ntiles + 1;   // + the off-grid bucket
  const size_t Z = (size_t)c.Z;
  uint32_t *keys = nullptr, *counts = nullptr;   // counts | start | nitems | item_start, nb + 1 words each
  void* tmp = nullptr;
  int rc = 1;
  do {
    if (cudaMalloc(&keys, Z * sizeof(uint32_t)) != cudaSuccess || cudaMalloc(&c.perm, Z * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(&counts, 4 * (size_t)(nb + 1) * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(&tmp, gvm_sort_temp_bytes(Z) + gvm_scan_temp_bytes(nb + 1)) != cudaSuccess) {
      gvm_set_error("tile plan: out of device memory for %zu samples", Z);
      break;
    }
    uint32_t *start = counts + (nb + 1), *nitems = start + (nb + 1), *item_start = nitems + (nb + 1);
    if (cudaMemsetAsync(counts, 0, 4 * (size_t)(nb + 1) * sizeof(uint32_t), e->stream) != cudaSuccess) break;
    k_tile_keys<<<(unsigned)((Z + 255) / 256), 256, 0, e->stream>>>(uvw_m_dev, c.d.freq, deltau, deltav, N, c.Z, ntx, ntiles,
                                                                    keys, c.perm, counts);
    GVM_LAUNCH(e);
    int bits = 1;
    while (((uint32_t)1 << bits) <= ntiles) bits++;
    if (gvm_sort_pairs_u32(keys, c.perm, Z, bits, tmp, e->stream)) break;
    if (cudaMemcpyAsync(start, counts, (nb + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, e->stream) != cudaSuccess) break;
    if (gvm_exclusive_scan_u32(start, nb + 1, tmp, e->stream)) break;
    k_item_counts<<<(nb + 255) / 256, 256, 0, e->stream>>>(counts, nb, nitems);
    if (cudaMemcpyAsync(item_start, nitems, (nb + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, e->stream) != cudaSuccess) break;
    if (gvm_exclusive_scan_u32(item_start, nb + 1, tmp, e->stream)) break;
    uint32_t total = 0;
    if (cudaMemcpyAsync(&total, item_start + nb, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream) != cudaSuccess) break;
    if (cudaStreamSynchronize(e->stream) != cudaSuccess) break;
    if (cudaMalloc(&c.items, (size_t)(total ? total : 1) * sizeof(uint4)) != cudaSuccess) break;
    k_make_items<<<(nb + 255) / 256, 256, 0, e->stream>>>(counts, start, item_start, nb, c.items);
    GVM_LAUNCH(e);
    // Contiguous runs of items per block, balanced by COST, not by count: an item moves 40 B per sample and, when
    // it opens a new bucket, the 9 KB grid tile — sparse outer tiles are all tile, dense central ones all samples.
    std::vector<uint4> h_items(total);
    if (total && cudaMemcpyAsync(h_items.data(), c.items, (size_t)total * sizeof(uint4), cudaMemcpyDeviceToHost, e->stream) != cudaSuccess) break;
    if (cudaStreamSynchronize(e->stream) != cudaSuccess) break;
    int blocks = e->sm_count * 4;   // 4 stages x 9.4 KB + 8.5 KB tile = 47 KB: four blocks per SM
    if (blocks > (int)total) blocks = (int)total;
    if (blocks > e->red_blocks) blocks = e->red_blocks;
    if (blocks < 1) blocks = 1;
    // bytes-equivalent cost: streams, a fixed per-item overhead (barriers, mbarrier wait), the tile when the bucket changes
    const double tile_cost = (double)(kTile + 1) * (kTile + 1) * sizeof(float2) + 4096.0;
    auto cost = [&](uint32_t i) {
      return 40.0 * h_items[i].z + 4096.0 + ((i == 0 || h_items[i].x != h_items[i - 1].x) ? tile_cost : 0.0);
    };
    double all = 0.0;
    for (uint32_t i = 0; i < total; i++) all += cost(i);
    std::vector<uint32_t> first((size_t)blocks + 1, total);
    first[0] = 0;
    double run = 0.0;
    int b = 1;
    for (uint32_t i = 0; i < total && b < blocks; i++) {
      run += cost(i);
      while (b < blocks && run >= all * b / blocks) first[b++] = i + 1;
    }
    if (cudaMalloc(&c.block_first, ((size_t)blocks + 1) * sizeof(uint32_t)) != cudaSuccess) break;
    if (cudaMemcpy(c.block_first, first.data(), ((size_t)blocks + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) break;
    c.tiled_blocks = blocks;
    c.nitems = (int)total;
    c.ntx = ntx;
    c.invalid_bucket = ntiles;
    rc = 0;
  } while (0);
  if (rc && cudaPeekAtLastError() != cudaSuccess) gvm_set_error("tile plan: %s", cudaGetErrorString(cudaGetLastError()));
  cudaFree(keys); cudaFree(counts); cudaFree(tmp);
  if (rc) {
    cudaFree(c.perm); cudaFree(c.items); cudaFree(c.block_first);
    c.perm = nullptr; c.items = nullptr; c.block_first = nullptr; c.nitems = 0;
  }
  return rc;
}

int gvm_launch_prep_channel(gvm_engine* e, GvmChannel& c, const double* uvw_m_dev,
                            const float2* Vo_dev, const float* w_dev) {
  const gvm_config& g = e->cfg;
  const double deltax = GVM_RPDEG_D * g.DELTAX, deltay = GVM_RPDEG_D * g.DELTAY;  // src/mfs.cu:493-496
  const double deltau = 1.0 / (g.M * deltax), deltav = 1.0 / (g.N * deltay);
  if (build_tile_plan(e, c, uvw_m_dev, deltau, deltav)) return 1;
  float* d_max = nullptr;
  GVM_CUDA(cudaMalloc(&d_max, 16));
  GVM_CUDA(cudaMemsetAsync(d_max, 0, 16, e->stream));
  unsigned long long* d_off = reinterpret_cast<unsigned long long*>(d_max) + 1;
  const int blocks = (int)((c.Z + 255) / 256);
  k_prep_channel<<<blocks, 256, 0, e->stream>>>(uvw_m_dev, Vo_dev, w_dev, c.perm, c.d.freq, deltau, deltav,
                                                 g.DELTAX * GVM_RPDEG_D, g.DELTAY * GVM_RPDEG_D,
                                                 g.N, c.Z, c.uvw_l, c.cell, c.frac, c.ccell, c.Vo, c.w,
                                                 c.du64, c.dv64, c.wz, d_max, d_off);
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  unsigned long long h_off = 0;
  GVM_CUDA(cudaMemcpyAsync(&h_off, d_off, sizeof(h_off), cudaMemcpyDeviceToHost, e->stream));
  GVM_CUDA(cudaMemcpyAsync(&c.max_abs_wz, d_max, sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  GVM_CUDA(cudaStreamSynchronize(e->stream));
  c.offgrid = (long)h_off;
  cudaFree(d_max);
  return 0;
}

// dst_dev[caller's index] = src_dev[position] for the per-sample arrays a caller reads back (gvm_get_vis)
int gvm_unpermute(gvm_engine* e, const GvmChannel& c, const void* src_dev, void* dst_dev, int elem_bytes) {
  const unsigned blocks = (unsigned)((c.Z + 255) / 256);
  if (elem_bytes == 4)
    k_unpermute<uint32_t><<<blocks, 256, 0, e->stream>>>(static_cast<const uint32_t*>(src_dev), c.perm, c.Z, static_cast<uint32_t*>(dst_dev));
  else if (elem_bytes == 8)
    k_unpermute<uint2><<<blocks, 256, 0, e->stream>>>(static_cast<const uint2*>(src_dev), c.perm, c.Z, static_cast<uint2*>(dst_dev));
  else { gvm_set_error("gvm_unpermute: element size %d", elem_bytes); return 1; }
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  return 0;
}

const float* gvm_channel_atten(gvm_engine* e, GvmChannel& c) {
  if (c.atten) return c.atten;
  static const size_t budget = [] {
    const char* s = getenv("GVM_ATTEN_CACHE_MB");
    return (size_t)(s ? atol(s) : 8192) << 20;
  }();
  const gvm_config& g = e->cfg;
  const size_t bytes = (size_t)g.M * g.N * sizeof(float);
  if (e->atten_cache_bytes + bytes > budget) return nullptr;
  if (cudaMalloc(&c.atten, bytes) != cudaSuccess) { c.atten = nullptr; cudaGetLastError(); return nullptr; }
  e->atten_cache_bytes += bytes;
  const long MN = g.M * g.N;
  k_atten_image<<<(int)((MN + 255) / 256), 256, 0, e->stream>>>(
      c.atten, g.N, g.M, c.d.antenna_diameter, c.d.pb_factor, c.d.pb_cutoff, c.d.freq, c.d.ref_xobs_pix,
      c.d.ref_yobs_pix, g.DELTAX, g.DELTAY, c.d.primary_beam);
  GVM_LAUNCH(e);
  return c.atten;
}

int gvm_forward_channel(gvm_engine* e, GvmChannel& c, float* I_dev, bool first, int flag_opt,
                        int slot) {
  const gvm_config& g = e->cfg;
  const long MN = g.M * g.N;
  const int pix_blocks = (int)((MN + 255) / 256);
  // half-plane forward model (R2C + per-tap phase rotation) when the block has few samples per pixel
  bool half = e->forward_mode == GVM_FORWARD_HALF || (e->forward_mode == GVM_FORWARD_AUTO && 4 * (long)c.Z <= MN);
  if (e->degrid_table || (g.N & 1)) half = false;
  if (half && !e->have_plan_r2c) {
    if (cufftPlan2d(&e->plan_r2c, (int)g.N, (int)g.M, CUFFT_R2C) != CUFFT_SUCCESS) {
      gvm_set_error("cufftPlan2d(R2C) failed");
      return 1;
    }
    e->have_plan_r2c = true;
    cufftSetStream(e->plan_r2c, e->stream);
  }
  e->last_forward_half = half ? 1 : 0;
  const float* atten_plane = gvm_channel_atten(e, c);
  PrepArgs pa;
  pa.I = I_dev; pa.noise = e->noise; pa.gcf = e->gcf; pa.atten_plane = atten_plane; pa.out = e->I_nu;
  pa.N = g.N; pa.M = g.M; pa.noise_cut = g.noise_cut; pa.minpix = g.minpix; pa.eta = g.eta; pa.threshold = g.threshold;
  pa.nu = c.d.freq; pa.nu_0 = g.nu_0; pa.fg_scale = g.fg_scale; pa.D = c.d.antenna_diameter; pa.pb_factor = c.d.pb_factor;
  pa.pb_cutoff = c.d.pb_cutoff; pa.xobs = c.d.ref_xobs_pix; pa.yobs = c.d.ref_yobs_pix; pa.DELTAX = g.DELTAX;
  pa.DELTAY = g.DELTAY; pa.schedule = flag_opt; pa.primary_beam = c.d.primary_beam;
  const bool vec4 = MN % 4 == 0 && ((uintptr_t)I_dev & 15) == 0;   // engine-owned planes come from cudaMalloc
  const int prep_blocks = vec4 ? (int)((MN / 4 + 255) / 256) : pix_blocks;
#define GVM_PREP(CLIP, REAL)                                                          \
  do {                                                                                \
    if (vec4) k_image_prep4<CLIP, REAL><<<prep_blocks, 256, 0, e->stream>>>(pa);     \
    else k_image_prep<CLIP, REAL><<<prep_blocks, 256, 0, e->stream>>>(pa);           \
  } while (0)
  if (first) { if (half) GVM_PREP(true, true); else GVM_PREP(true, false); }
  else       { if (half) GVM_PREP(false, true); else GVM_PREP(false, false); }
#undef GVM_PREP
  GVM_LAUNCH(e);
  if (first && e->capturing && e->ev_fork) {   // fork point of the captured evaluation (see gvm_prior_value_to_slot)
    GVM_CUDA(cudaEventRecord(e->ev_fork, e->stream));
    e->fork_valid = true;
  }
  if (half) {
    if (cufftExecR2C(e->plan_r2c, reinterpret_cast<cufftReal*>(e->I_nu),
                     reinterpret_cast<cufftComplex*>(e->V)) != CUFFT_SUCCESS) {
      gvm_set_error("cufftExecR2C failed");
      return 1;
    }
    GVM_LAUNCH(e);
  } else {
    if (cufftExecC2C(e->plan, reinterpret_cast<cufftComplex*>(e->I_nu),
                     reinterpret_cast<cufftComplex*>(e->V), CUFFT_INVERSE) != CUFFT_SUCCESS) {
      gvm_set_error("cufftExecC2C failed");
      return 1;
    }
    GVM_LAUNCH(e);
    k_phase_rotate<<<pix_blocks, 256, 0, e->stream>>>(e->V, g.M, g.N, (double)c.d.phs_xobs_pix,
                                                       (double)c.d.phs_yobs_pix);
    GVM_LAUNCH(e);
  }
  double* partials = e->red_partials + (size_t)slot * e->red_blocks;
  float* pmax = reinterpret_cast<float*>(e->red_partials + (size_t)e->red_slots * e->red_blocks) +
                (size_t)slot * e->red_blocks;
  if (c.Z > 0) {
    GvmConvDegrid cv = {};
    if (e->degrid_table) {
      const double deltax = GVM_RPDEG_D * g.DELTAX, deltay = GVM_RPDEG_D * g.DELTAY;
      cv.uvw_l = c.uvw_l; cv.perm = c.perm; cv.table = e->degrid_table;
      cv.deltau = 1.0 / (g.M * deltax); cv.deltav = 1.0 / (g.N * deltay);
      cv.km = e->degrid_m; cv.kn = e->degrid_n; cv.sx = e->degrid_sx; cv.sy = e->degrid_sy;
    }
    cv.upix = (double)c.d.phs_xobs_pix / (double)g.M;
    cv.vpix = (double)c.d.phs_yobs_pix / (double)g.N;
    const bool conv = e->degrid_table != nullptr;
    if (c.items && !half && (!conv || g.N <= 32768)) {
      // tiled degridder: grid tile + streams in shared memory, one contiguous run of work items per block
      const int tw = kTile + (conv ? 2 * cv.sx + 1 : 1), th = kTile + (conv ? 2 * cv.sy + 1 : 1);
      const size_t smem = (size_t)kStages * kStageCap * (conv ? 16 : 24) + (size_t)tw * th * sizeof(float2) +
                          (conv ? (size_t)cv.km * cv.kn * sizeof(float) : 0);
      const int blocks = c.tiled_blocks;
      TiledArgs a;
      a.V = e->V; a.cell = c.cell; a.ccell = c.ccell; a.frac = c.frac; a.Vo = c.Vo; a.w = c.w; a.Vr = c.Vr;
      a.Vm = g.keep_vm ? c.Vm : nullptr;
      a.items = c.items; a.block_first = c.block_first;
      a.N = (int)g.N; a.ntx = c.ntx; a.invalid_bucket = c.invalid_bucket;
#define GVM_TILED(KEEP, CONV)                                                                                       \
  do {                                                                                                              \
    GVM_CUDA(cudaFuncSetAttribute(k_degrid_tiled<KEEP, CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_degrid_tiled<KEEP, CONV><<<blocks, 256, smem, e->stream>>>(a, partials, pmax, e->red_counter + slot,          \
                                                                e->red_sum + slot, e->red_max + slot, cv);         \
  } while (0)
      if (conv) { if (g.keep_vm) GVM_TILED(true, true); else GVM_TILED(false, true); }
      else      { if (g.keep_vm) GVM_TILED(true, false); else GVM_TILED(false, false); }
#undef GVM_TILED
    } else {
      long want = (c.Z + (long)kVisThreads * kVisPerThread - 1) / ((long)kVisThreads * kVisPerThread);
      const int blocks = (int)(want < 1 ? 1 : (want > e->red_blocks ? e->red_blocks : want));
#define GVM_DEGRID(KEEP, MODE)                                                                     \
  k_degrid_chi2<KEEP, MODE><<<blocks, kVisThreads, 0, e->stream>>>(                                \
      e->V, c.cell, c.frac, c.Vo, c.w, c.Vr, KEEP ? c.Vm : nullptr, c.Z, (int)g.N, partials, pmax, \
      e->red_counter + slot, e->red_sum + slot, e->red_max + slot, cv)
      if (conv)      { if (g.keep_vm) GVM_DEGRID(true, kGridConv); else GVM_DEGRID(false, kGridConv); }
      else if (half) { if (g.keep_vm) GVM_DEGRID(true, kGridHalf); else GVM_DEGRID(false, kGridHalf); }
      else           { if (g.keep_vm) GVM_DEGRID(true, kGridFull); else GVM_DEGRID(false, kGridFull); }
#undef GVM_DEGRID
    }
    GVM_LAUNCH(e);
  }
  c.slot = slot;
  GVM_CUDA(cudaGetLastError());
  return 0;
}

int gvm_reduce_finish(gvm_engine* e, int nslots, int normalize, double* out_dev) {
  k_chi2_combine<<<1, 32, 0, e->stream>>>(e->red_sum, e->red_Z, nslots, normalize, out_dev);
  GVM_LAUNCH(e);
  GVM_CUDA(cudaGetLastError());
  return 0;
}
