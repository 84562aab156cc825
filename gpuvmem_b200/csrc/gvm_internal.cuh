// gvm_internal.cuh — engine state and device helpers shared by the kernels.
// B200 (sm_100a) only. See DESIGN.md for the data layout.
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <math_constants.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/gvm_b200.h"

#define GVM_PI_F CUDART_PI_F
#define GVM_PI_D CUDART_PI
#define GVM_RPDEG_D (CUDART_PI / 180.0)          // include/functions.cuh:18
#define GVM_LIGHTSPEED 2.99792458E8f             // include/MSFITSIO.cuh:54
#define GVM_RZ 1.2196698912665045f               // include/functions.cuh:23 (stored as float)
#define GVM_CELL_INVALID 0xFFFFFFFFu
#define GVM_MAX_CKERNEL 1024                     // floats of a convolution-kernel table kept in shared memory

void gvm_set_error(const char* fmt, ...);

#define GVM_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t _e = (call);                                                        \
    if (_e != cudaSuccess) {                                                        \
      gvm_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

struct GvmChannel {
  gvm_channel_desc d;
  int64_t Z = 0;               // samples this rank holds
  int64_t Znorm = 0;           // samples of the WHOLE block (numVisibilitiesPerFreqPerStoke): the `normalize` divisor;
                               // differs from Z when the block is cut into visibility chunks over the ranks
  // SoA, device
  double* uvw_l = nullptr;     // [Z][3] wavelengths after the Hermitian fold (kept for readback / exact checks)
  // tile-sorted upload (forward.cu): every per-sample array below except uvw_l is stored in TILE order; position p
  // holds the caller's sample perm[p]. Null: the caller's order (grid not a multiple of the tile, empty block).
  uint32_t* perm = nullptr;
  uint4* items = nullptr;      // work items of k_degrid_tiled: (bucket, first position, count, -)
  uint32_t* block_first = nullptr;   // [tiled_blocks + 1] first item of every block of the tiled degridder
  int nitems = 0, ntx = 0, tiled_blocks = 0;
  uint32_t invalid_bucket = 0;
  uint32_t* cell = nullptr;    // i1 | j1 << 16, GVM_CELL_INVALID when outside the grid
  uint32_t* ccell = nullptr;   // CKernel degridding: centre cell jc | kc << 16 in centred grid coordinates
  float2* frac = nullptr;      // (du, dv) bilinear fractions
  float2* Vo = nullptr;
  float* w = nullptr;
  float2* Vr = nullptr;
  float2* Vm = nullptr;        // only when cfg.keep_vm
  // gradient inputs (static per block): phase increments per pixel step as
  // 0.64 fixed-point fractions of a turn, and the w coordinate in wavelengths
  uint64_t* du64 = nullptr;    // frac(u_lambda * DELTAX * pi/180) * 2^64
  uint64_t* dv64 = nullptr;    // frac(v_lambda * DELTAY * pi/180) * 2^64
  float* wz = nullptr;
  // per-evaluation gradient coefficients of the tensor-core path (grad_umma.cu)
  float* amp = nullptr;        // w_k |Vr_k| * 2^e
  uint32_t* gam = nullptr;     // arg(Vr_k) as a 0.32 fixed-point turn
  float max_abs_wz = 0.f;      // max |w| (wavelengths)
  long offgrid = -1;           // samples that are NOT the centre of a uv cell with w = 0 (0: gridded data)
  int slot = -1;               // reduction slot of the last forward pass
  // attenuation(i, j) of this block's beam (src/functions.cu:2304-2333), cached on first use: the Airy
  // beam costs a j1f per pixel and the reference re-evaluates it in every kernel of every evaluation
  float* atten = nullptr;      // [MN] or null (cache budget exhausted: evaluated on the fly)
};

struct gvm_engine {
  gvm_config cfg;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cufftHandle plan = 0;
  bool have_plan = false;
  cufftHandle plan_r2c = 0;       // half-plane forward model (forward.cu), created on first use
  bool have_plan_r2c = false;
  cufftHandle plan_c2r = 0;       // gridded gradient / error maps (grad_gridfft.cu), created on first use
  bool have_plan_c2r = false;
  int forward_mode = 0;           // GVM_FORWARD_*
  int last_forward_half = 0;
  int sm_count = 148;
  // image-sized scratch
  float2* I_nu = nullptr;   // [MN] complex
  float2* V = nullptr;      // [MN] complex
  float* noise = nullptr;   // [MN]
  float* gcf = nullptr;     // [MN] or null
  // optional convolutional degridding in the forward model (gvm_set_degrid_kernel); null = bilinear vis_mod
  float* degrid_table = nullptr;
  int degrid_m = 0, degrid_n = 0, degrid_sx = 0, degrid_sy = 0;
  float* dchi2 = nullptr;   // [MN] per-channel gradient before the chain rule
  float* grad_scratch = nullptr;  // [ksplit][MN] partial sums
  size_t grad_scratch_floats = 0;
  size_t atten_cache_bytes = 0;   // attenuation planes held by the channels (budget: GVM_ATTEN_CACHE_MB, default 8192)
  float* pixtab = nullptr;  // [2][N] gA(x_j), gB(y_i) tables (w-term), rebuilt per channel
  // reductions
  double* red_partials = nullptr;  // per-block partials
  unsigned int* red_counter = nullptr;
  double* red_sum = nullptr;       // [slot] sum_k w|Vr|^2 of the last forward pass
  float* red_max = nullptr;        // [slot] max_k w*max(|Vr.re|,|Vr.im|)  (fp16 scaling of the UMMA path)
  double* red_out = nullptr;       // [0] = 0.5*chi2 of the last gvm_chi2
  double* h_red = nullptr;         // pinned host mirror of red_out
  double* obj_slots = nullptr;     // [GVM_OBJ_SLOTS][3] device results of gvm_*_to_slot (one sync per objective)
  double* h_slots = nullptr;       // pinned host mirror
  long* red_Z = nullptr;           // [slot] visibilities per block (device)
  int red_blocks = 0;
  int red_slots = 0;
  int flag_opt = 0;                // optimizer schedule flag (src/frprmn.cu:46), read by the clip
  // staging
  float* I_stage = nullptr;     // [2][MN] device image for gvm_eval_host
  float* grad_stage = nullptr;  // [2][MN]
  std::vector<GvmChannel> chans;
  int64_t launches = 0;
  int64_t epoch = 0;               // bumped by every call that invalidates captured graphs (gvm_state_epoch)
  bool capturing = false;          // between gvm_graph_begin and gvm_graph_end
  // inside a capture the prior values run on a forked branch of the graph: they only need the image as the first
  // channel's preparation pass left it (clip2IWNoise), not the FFT / degridding chain behind it
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool fork_valid = false, join_pending = false;
  // telemetry
  std::vector<cudaEvent_t> ev;
  int ev_used = 0;
  int last_grad_mode = 0;
  unsigned int* tile_counter = nullptr;
  // tensor-core gradient: tile plan over the unmasked part of the image (grad_umma.cu)
  unsigned umma_attr_set = 0;      // dynamic shared-memory opt-in of the k_grad_umma instantiations (bit mask) done on this engine's device
  bool plan_dirty = true;          // noise image / noise_cut changed since the last plan
  int2* row_ext = nullptr;         // [N] (first, last) unmasked column of every row (device)
  int4* tile_list = nullptr;       // [ntiles] (i0, j0, block width, 0)   (device)
  int4* band_tab = nullptr;        // [nbands] (jmin, first tile, tiles, tile width) (device)
  int plan_ntiles = 0, plan_nbands = 0, plan_imin = 0;
  long plan_pixels = 0;            // output pixels the plan computes (algorithmic flops = 4 * this * Z)
  // multi-GPU (dist_nccl.cu): one process per GPU, NCCL communicator over NVLink
  void* nccl_comm = nullptr;
  bool dist_aborted = false;
  bool replicated = false;         // every rank holds ALL blocks (gridded data): sums are complete locally, no all-reduce
  int rank = 0, world = 1;
  float* dist_grad = nullptr;      // [2][MN] this rank's gradient contribution before the all-reduce
  int64_t collectives = 0;
  // error maps (errormaps.cu): while set, the gradient contraction runs on w_k Vr_k instead of
  // w_k conj(Vr_k), without the w-term, and the finishing pass stores the raw sum in `dchi2`
  // (alpha_Noise, src/functions.cu:4113-4177, is the same direct DFT as DChi2)
  int err_variant = 0;
  // caller-visible device memory (gvm_dev_alloc/free): a caching pool — cudaMalloc/cudaFree cost
  // 15-45 ms each on a loaded context, and the optimizers allocate work buffers per optimize() call
  std::map<void*, size_t> pool_live;          // pointer -> bytes of every block handed out
  std::multimap<size_t, void*> pool_free;     // bytes -> cached free blocks
  std::vector<void*> pool_slabs;              // what cudaMalloc returned (blocks are carved out of these)
};

#define GVM_LAUNCH(e) ((e)->launches++)

// ---------------------------------------------------------------- device helpers
// Restates src/MSFITSIO.cu:36-45 (fp32 wavelength, fp64 division).
__host__ __device__ inline float gvm_freq_to_wavelength(float freq) { return GVM_LIGHTSPEED / freq; }
__host__ __device__ inline double gvm_metres_to_lambda(double m, float freq) {
  float lambda = gvm_freq_to_wavelength(freq);
  return m / lambda;
}

// attenuation(): src/functions.cu:2304-2333 with AiryDiskBeam (:2275) / GaussianBeam (:2293).
__device__ inline float gvm_attenuation(int i, int j, float D, float pb_factor, float pb_cutoff,
                                        float freq, float xobs, float yobs, double DELTAX,
                                        double DELTAY, int primary_beam) {
  int x0 = (int)xobs;
  int y0 = (int)yobs;
  float x = (float)((j - x0) * DELTAX * GVM_RPDEG_D);
  float y = (float)((i - y0) * DELTAY * GVM_RPDEG_D);
  // distance(): src/MSFITSIO.cu:47-51; nvcc contracts it as fma(x, x, y*y) (checked in the reference's SASS)
  float arc = sqrtf(fmaf(x, x, __fmul_rn(y, y)));
  float lambda = gvm_freq_to_wavelength(freq);
  float atten;
  if (primary_beam == GVM_BEAM_AIRYDISK) {
    atten = 1.0f;
    if (arc != 0.0f) {
      float arg = GVM_PI_F * arc * D / lambda * (GVM_RZ / pb_factor);
      float b = j1f(arg);
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

__device__ inline float gvm_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ inline double gvm_warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ inline float gvm_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// What DChi2 does after its visibility loop (src/functions.cu:3779-3790) followed by the chain
// rule of DChi2_total_I_nu_0 (:4000-4024, flag_opt even) / DChi2_total_alpha (:3968-3998, odd),
// accumulated into result. `d` is the raw sum over visibilities for an UNMASKED pixel.
struct GvmFinishParams {
  const float* atten;   // cached attenuation plane of the block, or null
  const float* gcf;
  const float* I;
  float* result;
  float* dchi2_out;
  long N, M, Z;
  float fg_scale, D, pb_factor, pb_cutoff, freq, xobs, yobs, nu_0, threshold;
  double DELTAX, DELTAY;
  int primary_beam, flag_opt, normalize;
  int raw;   // error maps: dchi2_out = raw sum over visibilities, nothing else
};
__device__ inline void gvm_finish_pixel(const GvmFinishParams& p, float d, long idx, int i, int j) {
  const long MN = p.M * p.N;
  if (p.raw) { p.dchi2_out[idx] = d; return; }
  const float atten = p.atten ? p.atten[idx]
                              : gvm_attenuation(i, j, p.D, p.pb_factor, p.pb_cutoff, p.freq, p.xobs, p.yobs,
                                                p.DELTAX, p.DELTAY, p.primary_beam);
  float scale_factor = p.fg_scale * atten;
  if (p.gcf) scale_factor = scale_factor * p.gcf[idx];
  d *= scale_factor;
  if (p.normalize) d /= p.Z;
  const float dchi2 = -d;
  if (p.dchi2_out) p.dchi2_out[idx] = dchi2;
  const float I0 = p.I[idx];
  const float alpha = p.I[MN + idx];
  const float nudiv = p.freq / p.nu_0;
  const float dI = powf(nudiv, alpha);
  if (p.flag_opt % 2 == 0) {
    p.result[idx] += dchi2 * dI;
  } else {
    const float dalpha = I0 * dI * p.fg_scale * logf(nudiv);
    if (I0 > p.threshold) p.result[MN + idx] += dchi2 * dalpha;
  }
}
GvmFinishParams gvm_finish_params(gvm_engine* e, const GvmChannel& c, const float* I_dev, int flag_opt,
                                  int normalize, float* result_dev);

// ------------------------------------------------------------ kernel launchers
// forward.cu
// the block's cached attenuation plane (built on first use; null when over budget)
const float* gvm_channel_atten(gvm_engine* e, GvmChannel& c);
int gvm_launch_prep_channel(gvm_engine* e, GvmChannel& c, const double* uvw_m_dev,
                            const float2* Vo_dev, const float* w_dev);
int gvm_forward_channel(gvm_engine* e, GvmChannel& c, float* I_dev, bool first, int flag_opt,
                        int slot);
int gvm_reduce_finish(gvm_engine* e, int nslots, int normalize, double* out_dev);
int gvm_unpermute(gvm_engine* e, const GvmChannel& c, const void* src_dev, void* dst_dev, int elem_bytes);
// sort.cu: hand-written stable radix sort of (u32 key, u32 value) pairs and exclusive scan
size_t gvm_sort_temp_bytes(size_t n);
size_t gvm_scan_temp_bytes(size_t n);
int gvm_sort_pairs_u32(uint32_t* keys, uint32_t* vals, size_t n, int key_bits, void* temp, cudaStream_t stream);
int gvm_exclusive_scan_u32(uint32_t* data, size_t n, void* temp, cudaStream_t stream);
// grad_simt.cu
int gvm_grad_simt(gvm_engine* e, GvmChannel& c, bool exact, int* ksplit_out);
// grad_umma.cu
// grad_umma.cu: tensor-core gradient incl. its own finishing pass (accumulates into result_dev)
int gvm_grad_umma(gvm_engine* e, GvmChannel& c, const float* I_dev, int flag_opt, int normalize,
                  float* result_dev);
bool gvm_grad_umma_supported(const gvm_engine* e, const GvmChannel& c);
// grad_gridfft.cu: gradient of gridded samples (cell centres, w = 0) as ONE inverse FFT
int gvm_grad_gridfft(gvm_engine* e, GvmChannel& c);
// shared by gradient paths
int gvm_grad_finish(gvm_engine* e, GvmChannel& c, const float* I_dev, int ksplit, int flag_opt,
                    int normalize, float* result_dev);
int gvm_ensure_grad_scratch(gvm_engine* e, size_t floats);
int gvm_build_pixtab(gvm_engine* e, const GvmChannel& c);
double gvm_wterm_cross_bound(const gvm_engine* e, const GvmChannel& c);
// dist_nccl.cu: in-place sum all-reduce on the engine stream (no-op when world == 1)
int gvm_dist_allreduce_f32(gvm_engine* e, float* buf, size_t n);
int gvm_dist_allreduce_f64(gvm_engine* e, double* buf, size_t n);
int gvm_join_branch(gvm_engine* e);   // priors.cu: the forked prior-value branch of a captured evaluation rejoins e->stream
int gvm_dist_broadcast_f32(gvm_engine* e, float* buf, size_t n, int root);
int gvm_dist_send(gvm_engine* e, const void* buf, size_t bytes, int peer);
int gvm_dist_recv(gvm_engine* e, void* buf, size_t bytes, int peer);
int gvm_dist_group_begin(gvm_engine* e);
int gvm_dist_group_end(gvm_engine* e);
int gvm_dist_broadcast_bytes(gvm_engine* e, void* buf, size_t bytes, int root);
int gvm_dist_allreduce_u32_max(gvm_engine* e, uint32_t* buf, size_t n);
// a rank that fails locally before a collective tears the communicator down so that its peers error out of
// their pending collective instead of waiting for it forever
void gvm_dist_abort_comm(gvm_engine* e);
void gvm_dist_release(gvm_engine* e);
// errormaps.cu needs the mode rule of gvm_dchi2
int gvm_pick_grad_mode(gvm_engine* e, GvmChannel& c);
// hostcopy.cu: pipelined copies between pageable host memory and the device (synchronous)
int gvm_fast_h2d(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t stream);
int gvm_fast_d2h(void* dst_host, const void* src_dev, size_t bytes, cudaStream_t stream);
void gvm_hostcopy_warm();
void gvm_ev_begin(gvm_engine* e);
void gvm_ev_end(gvm_engine* e);
