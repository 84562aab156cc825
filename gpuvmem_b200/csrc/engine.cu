// engine.cu — the C-ABI (include/gvm_b200.h): lifecycle, uploads, and the host
// drivers of the hot path. The kernels live in forward.cu, grad_simt.cu, grad_umma.cu,
// grad_gridfft.cu, errormaps.cu, priors.cu (priors + optimizer vector ops) and weights_grid.cu.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <iterator>

#include "gvm_internal.cuh"

static thread_local char g_err[1024] = "";

void gvm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void gvm_ev_begin(gvm_engine* e) {
  if (e->ev_used + 2 > (int)e->ev.size()) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    e->ev.push_back(a);
    e->ev.push_back(b);
  }
  cudaEventRecord(e->ev[e->ev_used], e->stream);
}
void gvm_ev_end(gvm_engine* e) {
  cudaEventRecord(e->ev[e->ev_used + 1], e->stream);
  e->ev_used += 2;
}

extern "C" {

const char* gvm_last_error(void) { return g_err; }
int gvm_version(void) { return 100; }

// (Re)allocate the per-block reduction state for `slots` blocks, keeping what the existing blocks left
// there (Z, the sums and maxima of the last forward pass). Grows on demand in gvm_add_channel: large
// mosaics / many-channel datasets have thousands of (field, channel, stokes) blocks.
static int ensure_red_slots(gvm_engine* e, int slots) {
  if (slots <= e->red_slots) return 0;
  int n = e->red_slots > 0 ? e->red_slots : 1024;
  while (n < slots) n *= 2;
  const size_t rp = (size_t)n * e->red_blocks;
  double* partials = nullptr; unsigned int* counter = nullptr; double* sum = nullptr; long* Zs = nullptr; float* mx = nullptr;
  GVM_CUDA(cudaStreamSynchronize(e->stream));
  GVM_CUDA(cudaMalloc(&partials, rp * (sizeof(double) + sizeof(float))));
  GVM_CUDA(cudaMalloc(&counter, n * sizeof(unsigned int)));
  GVM_CUDA(cudaMemset(counter, 0, n * sizeof(unsigned int)));
  GVM_CUDA(cudaMalloc(&sum, n * sizeof(double)));
  GVM_CUDA(cudaMemset(sum, 0, n * sizeof(double)));
  GVM_CUDA(cudaMalloc(&Zs, n * sizeof(long)));
  GVM_CUDA(cudaMemset(Zs, 0, n * sizeof(long)));
  GVM_CUDA(cudaMalloc(&mx, 3 * (size_t)n * sizeof(float)));
  GVM_CUDA(cudaMemset(mx, 0, 3 * (size_t)n * sizeof(float)));
  if (e->red_slots > 0) {
    const int o = e->red_slots;
    GVM_CUDA(cudaMemcpy(sum, e->red_sum, o * sizeof(double), cudaMemcpyDeviceToDevice));
    GVM_CUDA(cudaMemcpy(Zs, e->red_Z, o * sizeof(long), cudaMemcpyDeviceToDevice));
    for (int k = 0; k < 3; k++)   // [max | inv_scale | spare] planes of `slots` floats each
      GVM_CUDA(cudaMemcpy(mx + (size_t)k * n, e->red_max + (size_t)k * o, o * sizeof(float), cudaMemcpyDeviceToDevice));
  }
  cudaFree(e->red_partials); cudaFree(e->red_counter); cudaFree(e->red_sum); cudaFree(e->red_Z); cudaFree(e->red_max);
  e->red_partials = partials; e->red_counter = counter; e->red_sum = sum; e->red_Z = Zs; e->red_max = mx;
  e->red_slots = n;
  return 0;
}

static int create_impl(gvm_engine* e, const gvm_config* cfg, const cudaDeviceProp& prop);

int gvm_create(const gvm_config* cfg, gvm_engine** out) {
  if (!cfg || !out) { gvm_set_error("gvm_create: null argument"); return 1; }
  if (cfg->M != cfg->N || cfg->N <= 0) {
    gvm_set_error("gvm_create: M == N > 0 required (got %ld x %ld); the reference kernels assume it",
                  (long)cfg->M, (long)cfg->N);
    return 1;
  }
  if (cfg->N > 65536) { gvm_set_error("gvm_create: N > 65536 unsupported"); return 1; }
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev < 1) {
    gvm_set_error("gvm_create: no CUDA device (%s); this engine has no CPU fallback",
                  cudaGetErrorString(err));
    return 1;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { gvm_set_error("gvm_create: bad device %d", cfg->device); return 1; }
  GVM_CUDA(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  GVM_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) {
    gvm_set_error("gvm_create: built for sm_100a only, device is sm_%d%d", prop.major, prop.minor);
    return 1;
  }
  gvm_engine* e = new gvm_engine();
  e->cfg = *cfg;
  if (create_impl(e, cfg, prop)) {   // nothing leaks on a failed allocation: gvm_destroy frees what exists
    gvm_destroy(e);
    return 1;
  }
  *out = e;
  return 0;
}

static int create_impl(gvm_engine* e, const gvm_config* cfg, const cudaDeviceProp& prop) {
  e->sm_count = prop.multiProcessorCount;
  // a BLOCKING stream: it orders itself against the legacy default stream, which is what the
  // reference's host code (and torch, by default) launches on — safe drop-in semantics
  GVM_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamDefault));
  e->own_stream = true;
  GVM_CUDA(cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking));
  GVM_CUDA(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
  GVM_CUDA(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
  const size_t MN = (size_t)cfg->M * cfg->N;
  GVM_CUDA(cudaMalloc(&e->I_nu, MN * sizeof(float2)));
  GVM_CUDA(cudaMalloc(&e->V, MN * sizeof(float2)));
  GVM_CUDA(cudaMalloc(&e->noise, MN * sizeof(float)));
  GVM_CUDA(cudaMemset(e->noise, 0, MN * sizeof(float)));
  GVM_CUDA(cudaMalloc(&e->dchi2, MN * sizeof(float)));
  GVM_CUDA(cudaMalloc(&e->pixtab, 2 * (size_t)cfg->N * sizeof(float)));
  GVM_CUDA(cudaMemset(e->pixtab, 0, 2 * (size_t)cfg->N * sizeof(float)));
  GVM_CUDA(cudaMalloc(&e->I_stage, 2 * MN * sizeof(float)));
  GVM_CUDA(cudaMalloc(&e->grad_stage, 2 * MN * sizeof(float)));
  e->red_blocks = e->sm_count * 8;
  e->red_slots = 0;
  if (ensure_red_slots(e, 1024)) return 1;
  GVM_CUDA(cudaMalloc(&e->red_out, 8 * sizeof(double)));
  GVM_CUDA(cudaMemset(e->red_out, 0, 8 * sizeof(double)));
  GVM_CUDA(cudaMallocHost(&e->h_red, 8 * sizeof(double)));
  GVM_CUDA(cudaMalloc(&e->obj_slots, 3 * GVM_OBJ_SLOTS * sizeof(double)));
  GVM_CUDA(cudaMemset(e->obj_slots, 0, 3 * GVM_OBJ_SLOTS * sizeof(double)));
  GVM_CUDA(cudaMallocHost(&e->h_slots, 3 * GVM_OBJ_SLOTS * sizeof(double)));
  GVM_CUDA(cudaMalloc(&e->tile_counter, 16 * sizeof(unsigned int)));
  GVM_CUDA(cudaMemset(e->tile_counter, 0, 16 * sizeof(unsigned int)));
  // cufftPlan2d(N, M, C2C): src/functions.cu:2149
  if (cufftPlan2d(&e->plan, (int)cfg->N, (int)cfg->M, CUFFT_C2C) != CUFFT_SUCCESS) {
    gvm_set_error("gvm_create: cufftPlan2d failed");
    return 1;
  }
  e->have_plan = true;
  cufftSetStream(e->plan, e->stream);
  gvm_hostcopy_warm();
  return 0;
}

static void free_channel(GvmChannel& c) {
  cudaFree(c.uvw_l); cudaFree(c.cell); cudaFree(c.ccell); cudaFree(c.frac); cudaFree(c.Vo); cudaFree(c.w);
  cudaFree(c.perm); cudaFree(c.items); cudaFree(c.block_first);
  c.perm = nullptr; c.items = nullptr; c.block_first = nullptr;
  cudaFree(c.Vr); cudaFree(c.Vm); cudaFree(c.du64); cudaFree(c.dv64); cudaFree(c.wz);
  cudaFree(c.amp); cudaFree(c.gam); cudaFree(c.atten);
  c.atten = nullptr;
}

int gvm_destroy(gvm_engine* e) {
  if (!e) return 0;
  cudaSetDevice(e->cfg.device);
  cudaDeviceSynchronize();
  for (auto& c : e->chans) free_channel(c);
  if (e->have_plan) cufftDestroy(e->plan);
  if (e->have_plan_r2c) cufftDestroy(e->plan_r2c);
  if (e->have_plan_c2r) cufftDestroy(e->plan_c2r);
  cudaFree(e->I_nu); cudaFree(e->V); cudaFree(e->noise); cudaFree(e->gcf); cudaFree(e->dchi2);
  cudaFree(e->degrid_table);
  cudaFree(e->grad_scratch); cudaFree(e->pixtab); cudaFree(e->I_stage); cudaFree(e->grad_stage);
  cudaFree(e->red_partials); cudaFree(e->red_counter); cudaFree(e->red_sum); cudaFree(e->red_Z);
  cudaFree(e->red_max); cudaFree(e->red_out); cudaFreeHost(e->h_red); cudaFree(e->obj_slots); cudaFreeHost(e->h_slots); cudaFree(e->tile_counter);
  cudaFree(e->row_ext); cudaFree(e->tile_list); cudaFree(e->band_tab);
  gvm_dist_release(e);
  cudaFree(e->dist_grad);
  for (void* slab : e->pool_slabs) cudaFree(slab);   // blocks (live or cached) are carved out of the slabs
  for (auto ev : e->ev) cudaEventDestroy(ev);
  if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
  if (e->stream2) cudaStreamDestroy(e->stream2);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  delete e;
  return 0;
}

int gvm_set_stream(gvm_engine* e, void* s) {
  e->epoch++;
  if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
  e->own_stream = false;
  e->stream = (cudaStream_t)s;
  if (e->have_plan) cufftSetStream(e->plan, e->stream);
  if (e->have_plan_r2c) cufftSetStream(e->plan_r2c, e->stream);
  if (e->have_plan_c2r) cufftSetStream(e->plan_c2r, e->stream);
  return 0;
}
void* gvm_get_stream(gvm_engine* e) { return (void*)e->stream; }
int gvm_synchronize(gvm_engine* e) {
  GVM_CUDA(cudaStreamSynchronize(e->stream));
  return 0;
}

int gvm_set_scalars(gvm_engine* e, float fg_scale, float noise_cut, float threshold) {
  // the gradient's tile plan depends on the mask (noise < noise_cut) only
  if (e->cfg.noise_cut != noise_cut) e->plan_dirty = true;
  e->cfg.fg_scale = fg_scale;
  e->cfg.noise_cut = noise_cut;
  e->cfg.threshold = threshold;
  return 0;
}
int gvm_set_grad_mode(gvm_engine* e, int m) { e->cfg.grad_mode = m; return 0; }
int gvm_set_flag_opt(gvm_engine* e, int f) { e->flag_opt = f; return 0; }

int gvm_set_noise_image(gvm_engine* e, const float* noise, int src_is_device) {
  const size_t MN = (size_t)e->cfg.M * e->cfg.N;
  e->plan_dirty = true;
  e->epoch++;
  GVM_CUDA(cudaMemcpyAsync(e->noise, noise, MN * sizeof(float),
                           src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, e->stream));
  GVM_CUDA(cudaStreamSynchronize(e->stream));
  return 0;
}
int gvm_get_noise_image(gvm_engine* e, float* out) {
  const size_t MN = (size_t)e->cfg.M * e->cfg.N;
  GVM_CUDA(cudaMemcpyAsync(out, e->noise, MN * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  GVM_CUDA(cudaStreamSynchronize(e->stream));
  return 0;
}
int gvm_set_gcf(gvm_engine* e, const float* gcf_host) {
  const size_t MN = (size_t)e->cfg.M * e->cfg.N;
  e->epoch++;
  if (!gcf_host) { cudaFree(e->gcf); e->gcf = nullptr; return 0; }
  if (!e->gcf) GVM_CUDA(cudaMalloc(&e->gcf, MN * sizeof(float)));
  GVM_CUDA(cudaMemcpyAsync(e->gcf, gcf_host, MN * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  GVM_CUDA(cudaStreamSynchronize(e->stream));
  return 0;
}

int gvm_set_degrid_kernel(gvm_engine* e, const float* table_host, int m, int n, int support_x, int support_y) {
  GVM_CUDA(cudaSetDevice(e->cfg.device));
  e->epoch++;
  GVM_CUDA(cudaStreamSynchronize(e->stream));
  cudaFree(e->degrid_table);
  e->degrid_table = nullptr;
  if (!table_host) return 0;                      // back to the bilinear vis_mod
  if (m < 1 || n < 1 || m * n > GVM_MAX_CKERNEL || support_x < 0 || support_y < 0 ||
      2 * support_y + 1 > m || 2 * support_x + 1 > n) {
    gvm_set_error("gvm_set_degrid_kernel: table %dx%d with supports (%d, %d) is not usable (max %d entries)", m, n,
                  support_x, support_y, GVM_MAX_CKERNEL);
    return 1;
  }
  GVM_CUDA(cudaMalloc(&e->degrid_table, (size_t)m * n * sizeof(float)));
  GVM_CUDA(cudaMemcpy(e->degrid_table, table_host, (size_t)m * n * sizeof(float), cudaMemcpyHostToDevice));
  e->degrid_m = m; e->degrid_n = n; e->degrid_sx = support_x; e->degrid_sy = support_y;
  return 0;
}

int gvm_set_forward_mode(gvm_engine* e, int mode) {
  if (mode < GVM_FORWARD_AUTO || mode > GVM_FORWARD_HALF) { gvm_set_error("gvm_set_forward_mode: unknown mode %d", mode); return 1; }
  e->forward_mode = mode;
  e->epoch++;
  return 0;
}
int gvm_last_forward_mode(gvm_engine* e) { return e->last_forward_half ? GVM_FORWARD_HALF : GVM_FORWARD_FULL; }

int gvm_get_model_grid(gvm_engine* e, float* V_host) {
  if (e->last_forward_half) { gvm_set_error("gvm_get_model_grid: the last forward pass used the half-plane model (gvm_set_forward_mode)"); return 1; }
  const size_t MN = (size_t)e->cfg.M * e->cfg.N;
  GVM_CUDA(cudaSetDevice(e->cfg.device));
  GVM_CUDA(cudaMemcpyAsync(V_host, e->V, MN * sizeof(float2), cudaMemcpyDeviceToHost, e->stream));
  GVM_CUDA(cudaStreamSynchronize(e->stream));
  return 0;
}

static int add_channel_impl(gvm_engine* e, GvmChannel& c, int64_t Z, const double* uvw_m, const float* Vo, const float* w) {
  const size_t z = (size_t)(Z > 0 ? Z : 1);
  // the streamed arrays carry 32 samples of slack: the tiled degridder's bulk copies start and end on 16-sample
  // boundaries and may read (never use) up to that far past the last sample
  const size_t zs = z + 32;
  GVM_CUDA(cudaMalloc(&c.uvw_l, z * 3 * sizeof(double)));
  GVM_CUDA(cudaMalloc(&c.cell, zs * sizeof(uint32_t)));
  GVM_CUDA(cudaMalloc(&c.ccell, zs * sizeof(uint32_t)));
  GVM_CUDA(cudaMalloc(&c.frac, zs * sizeof(float2)));
  GVM_CUDA(cudaMalloc(&c.Vo, zs * sizeof(float2)));
  GVM_CUDA(cudaMalloc(&c.w, zs * sizeof(float)));
  GVM_CUDA(cudaMemsetAsync(c.cell + (zs - 32 - (Z > 0 ? 0 : 1)), 0xFF, (32 + (Z > 0 ? 0 : 1)) * sizeof(uint32_t), e->stream));
  GVM_CUDA(cudaMemsetAsync(c.ccell + (zs - 32 - (Z > 0 ? 0 : 1)), 0xFF, (32 + (Z > 0 ? 0 : 1)) * sizeof(uint32_t), e->stream));
  GVM_CUDA(cudaMemsetAsync(c.frac + (zs - 32 - (Z > 0 ? 0 : 1)), 0, (32 + (Z > 0 ? 0 : 1)) * sizeof(float2), e->stream));
  GVM_CUDA(cudaMemsetAsync(c.Vo + (zs - 32 - (Z > 0 ? 0 : 1)), 0, (32 + (Z > 0 ? 0 : 1)) * sizeof(float2), e->stream));
  GVM_CUDA(cudaMemsetAsync(c.w + (zs - 32 - (Z > 0 ? 0 : 1)), 0, (32 + (Z > 0 ? 0 : 1)) * sizeof(float), e->stream));
  GVM_CUDA(cudaMalloc(&c.Vr, z * sizeof(float2)));
  GVM_CUDA(cudaMemset(c.Vr, 0, z * sizeof(float2)));
  if (e->cfg.keep_vm) {
    GVM_CUDA(cudaMalloc(&c.Vm, z * sizeof(float2)));
    GVM_CUDA(cudaMemset(c.Vm, 0, z * sizeof(float2)));
  }
  // + 8 elements: the gradient kernel's bulk copies round a ragged tail up to 4 visibilities
  GVM_CUDA(cudaMalloc(&c.du64, (z + 8) * sizeof(uint64_t)));
  GVM_CUDA(cudaMalloc(&c.dv64, (z + 8) * sizeof(uint64_t)));
  GVM_CUDA(cudaMalloc(&c.wz, (z + 8) * sizeof(float)));
  GVM_CUDA(cudaMemsetAsync(c.du64 + z, 0, 8 * sizeof(uint64_t), e->stream));
  GVM_CUDA(cudaMemsetAsync(c.dv64 + z, 0, 8 * sizeof(uint64_t), e->stream));
  GVM_CUDA(cudaMemsetAsync(c.wz + z, 0, 8 * sizeof(float), e->stream));
  if (Z > 0) {
    double* d_uvw = nullptr; float2* d_vo = nullptr; float* d_w = nullptr;
    int rc = cudaMalloc(&d_uvw, z * 3 * sizeof(double)) != cudaSuccess || cudaMalloc(&d_vo, z * sizeof(float2)) != cudaSuccess ||
             cudaMalloc(&d_w, z * sizeof(float)) != cudaSuccess;
    if (rc) gvm_set_error("gvm_add_channel: out of device memory staging %ld visibilities", (long)Z);
    rc = rc || gvm_fast_h2d(d_uvw, uvw_m, z * 3 * sizeof(double), e->stream) || gvm_fast_h2d(d_vo, Vo, z * sizeof(float2), e->stream) ||
         gvm_fast_h2d(d_w, w, z * sizeof(float), e->stream) || gvm_launch_prep_channel(e, c, d_uvw, d_vo, d_w);
    cudaFree(d_uvw); cudaFree(d_vo); cudaFree(d_w);
    if (rc) return 1;
  }
  return 0;
}

int gvm_add_channel(gvm_engine* e, const gvm_channel_desc* desc, int64_t Z, const double* uvw_m,
                    const float* Vo, const float* w, int* chan_out) {
  if (!e || !desc || Z < 0) { gvm_set_error("gvm_add_channel: bad argument"); return 1; }
  GVM_CUDA(cudaSetDevice(e->cfg.device));
  e->epoch++;
  // the last two reduction slots belong to the image-sized reductions (priors.cu aux_red)
  if (ensure_red_slots(e, (int)e->chans.size() + 3)) return 1;
  GvmChannel c;
  c.d = *desc;
  c.Z = Z;
  c.Znorm = Z;
  if (add_channel_impl(e, c, Z, uvw_m, Vo, w)) {
    free_channel(c);     // a failed upload leaves nothing behind
    return 1;
  }
  const int slot = (int)e->chans.size();
  long zl = (long)Z;
  GVM_CUDA(cudaMemcpy(e->red_Z + slot, &zl, sizeof(long), cudaMemcpyHostToDevice));
  GVM_CUDA(cudaMemset(e->red_sum + slot, 0, sizeof(double)));   // a rank-empty block contributes 0, not a stale sum
  e->chans.push_back(c);
  if (chan_out) *chan_out = slot;
  return 0;
}

int gvm_set_block_nvis(gvm_engine* e, int chan, int64_t Z_block) {
  if (chan < 0 || chan >= (int)e->chans.size() || Z_block < e->chans[chan].Z) {
    gvm_set_error("gvm_set_block_nvis: bad channel %d or block size %ld", chan, (long)Z_block);
    return 1;
  }
  GVM_CUDA(cudaSetDevice(e->cfg.device));
  e->chans[chan].Znorm = Z_block;
  long zl = (long)Z_block;
  GVM_CUDA(cudaMemcpy(e->red_Z + chan, &zl, sizeof(long), cudaMemcpyHostToDevice));
  return 0;
}

int gvm_clear_channels(gvm_engine* e) {
  GVM_CUDA(cudaSetDevice(e->cfg.device));
  e->epoch++;
  GVM_CUDA(cudaStreamSynchronize(e->stream));
  for (auto& c : e->chans) free_channel(c);
  e->chans.clear();
  e->atten_cache_bytes = 0;
  return 0;
}

int gvm_num_channels(gvm_engine* e) { return (int)e->chans.size(); }
int64_t gvm_channel_nvis(gvm_engine* e, int chan) {
  if (chan < 0 || chan >= (int)e->chans.size()) return -1;
  return e->chans[chan].Z;
}

int gvm_get_vis(gvm_engine* e, int chan, double* uvw_lambda, int32_t* cell, float* Vo, float* Vm,
                float* Vr, float* w) {
  if (chan < 0 || chan >= (int)e->chans.size()) { gvm_set_error("gvm_get_vis: bad channel"); return 1; }
  GvmChannel& c = e->chans[chan];
  const size_t Z = (size_t)c.Z;
  GVM_CUDA(cudaSetDevice(e->cfg.device));
  GVM_CUDA(cudaStreamSynchronize(e->stream));
  if (Z == 0) return 0;
  if (Vm && !c.Vm) { gvm_set_error("gvm_get_vis: Vm not kept (cfg.keep_vm = 0)"); return 1; }
  // the device arrays are in tile order (c.perm); the caller gets its own sample order back
  void* tmp = nullptr;
  if (c.perm) GVM_CUDA(cudaMalloc(&tmp, Z * sizeof(float2)));
  auto fetch = [&](void* dst_host, const void* src_dev, int elem) -> int {
    if (c.perm) {
      if (gvm_unpermute(e, c, src_dev, tmp, elem)) return 1;
      src_dev = tmp;
    }
    return gvm_fast_d2h(dst_host, src_dev, Z * (size_t)elem, e->stream);
  };
  int rc = 0;
  if (uvw_lambda) rc = rc || gvm_fast_d2h(uvw_lambda, c.uvw_l, Z * 3 * sizeof(double), e->stream);
  if (Vo) rc = rc || fetch(Vo, c.Vo, sizeof(float2));
  if (Vr) rc = rc || fetch(Vr, c.Vr, sizeof(float2));
  if (Vm) rc = rc || fetch(Vm, c.Vm, sizeof(float2));
  if (w) rc = rc || fetch(w, c.w, sizeof(float));
  if (cell && !rc) {
    std::vector<uint32_t> packed(Z);
    rc = fetch(packed.data(), c.cell, sizeof(uint32_t));
    for (size_t k = 0; k < Z && !rc; k++) {
      if (packed[k] == GVM_CELL_INVALID) { cell[2 * k] = -1; cell[2 * k + 1] = -1; }
      else { cell[2 * k] = (int32_t)(packed[k] & 0xFFFFu); cell[2 * k + 1] = (int32_t)(packed[k] >> 16); }
    }
  }
  cudaFree(tmp);
  return rc;
}

// ------------------------------------------------------------------ hot path
static int chi2_async_impl(gvm_engine* e, float* I_dev, int normalize, double* chi2_dev);
// A failure on ONE rank before its collective must not leave the peers blocked in theirs: the failing
// rank aborts the communicator (peers return an error from the pending all-reduce).
int gvm_chi2_async(gvm_engine* e, float* I_dev, int normalize, double* chi2_dev) {
  const int rc = chi2_async_impl(e, I_dev, normalize, chi2_dev);
  if (rc && e->world > 1) gvm_dist_abort_comm(e);
  return rc;
}
static int chi2_async_impl(gvm_engine* e, float* I_dev, int normalize, double* chi2_dev) {
  GVM_CUDA(cudaSetDevice(e->cfg.device));
  if (e->chans.empty()) { gvm_set_error("gvm_chi2: no visibility blocks uploaded"); return 1; }
  bool first = true;
  for (size_t s = 0; s < e->chans.size(); s++) {
    GvmChannel& c = e->chans[s];
    // the reference skips empty blocks (src/functions.cu:4382) but still clips once per call
    if (c.Z <= 0 && !(first && s + 1 == e->chans.size())) continue;
    if (gvm_forward_channel(e, c, I_dev, first, e->flag_opt, (int)s)) return 1;
    first = false;
  }
  double* out = chi2_dev ? chi2_dev : e->red_out;
  if (gvm_reduce_finish(e, (int)e->chans.size(), normalize, out)) return 1;
  // multi-GPU: chi2 is a plain sum over the visibility shards
  return gvm_dist_allreduce_f64(e, out, 1);
}

int gvm_chi2(gvm_engine* e, float* I_dev, int normalize, float* chi2_out) {
  if (gvm_chi2_async(e, I_dev, normalize, e->red_out)) return 1;
  GVM_CUDA(cudaMemcpyAsync(e->h_red, e->red_out, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  GVM_CUDA(cudaStreamSynchronize(e->stream));
  if (chi2_out) *chi2_out = (float)e->h_red[0];
  return 0;
}

}  // extern "C"
int gvm_pick_grad_mode(gvm_engine* e, GvmChannel& c) {
  int mode = e->cfg.grad_mode;
  if (mode == GVM_GRAD_SIMT_EXACT || mode == GVM_GRAD_SIMT) return mode;
  if (mode == GVM_GRAD_GRIDFFT || (mode == GVM_GRAD_AUTO && c.offgrid == 0)) {
    if (c.offgrid == 0) return GVM_GRAD_GRIDFFT;
    mode = GVM_GRAD_AUTO;  // not applicable to these samples
  }
  const bool sep_ok = gvm_wterm_cross_bound(e, c) <= 2e-6;  // turns; DESIGN.md §3.4
  if (mode == GVM_GRAD_UMMA) return GVM_GRAD_UMMA;
  if (!sep_ok) return GVM_GRAD_SIMT_EXACT;
  return gvm_grad_umma_supported(e, c) ? GVM_GRAD_UMMA : GVM_GRAD_SIMT;
}
extern "C" {

static __global__ void k_add_inplace(float* __restrict__ dst, const float* __restrict__ src, long n) {
  const long idx = blockIdx.x * 256L + threadIdx.x;
  if (idx < n) dst[idx] += src[idx];
}

static int dchi2_impl(gvm_engine* e, const float* I_dev, int flag_opt, int normalize, float* result_dchi2_dev);
int gvm_dchi2(gvm_engine* e, const float* I_dev, int flag_opt, int normalize,
              float* result_dchi2_dev) {
  const int rc = dchi2_impl(e, I_dev, flag_opt, normalize, result_dchi2_dev);
  if (rc && e->world > 1) gvm_dist_abort_comm(e);
  return rc;
}
static int dchi2_impl(gvm_engine* e, const float* I_dev, int flag_opt, int normalize, float* result_dchi2_dev) {
  GVM_CUDA(cudaSetDevice(e->cfg.device));
  e->flag_opt = flag_opt;
  e->ev_used = 0;
  float* const caller_result = result_dchi2_dev;
  const size_t MN2 = 2 * (size_t)e->cfg.M * e->cfg.N;
  const bool reduce = e->world > 1 && !e->replicated;
  if (reduce) {
    // this rank's shard accumulates into a private buffer; ONE all-reduce of the image-sized
    // gradient replaces the reference's serialised peer-to-peer accumulate (src/functions.cu:4534-4549)
    if (!e->dist_grad) GVM_CUDA(cudaMalloc(&e->dist_grad, MN2 * sizeof(float)));
    GVM_CUDA(cudaMemsetAsync(e->dist_grad, 0, MN2 * sizeof(float), e->stream));
    result_dchi2_dev = e->dist_grad;
  }
  for (size_t s = 0; s < e->chans.size(); s++) {
    GvmChannel& c = e->chans[s];
    if (c.Z <= 0) continue;
    if (c.slot < 0) { gvm_set_error("gvm_dchi2: call gvm_chi2 first (Vr comes from the forward pass)"); return 1; }
    const int mode = gvm_pick_grad_mode(e, c);
    e->last_grad_mode = mode;
    if (mode == GVM_GRAD_UMMA) {
      if (gvm_grad_umma(e, c, I_dev, flag_opt, normalize, result_dchi2_dev)) return 1;
    } else if (mode == GVM_GRAD_GRIDFFT) {
      if (gvm_grad_gridfft(e, c)) return 1;
      if (gvm_grad_finish(e, c, I_dev, 1, flag_opt, normalize, result_dchi2_dev)) return 1;
    } else {
      int ksplit = 1;
      if (gvm_grad_simt(e, c, mode == GVM_GRAD_SIMT_EXACT, &ksplit)) return 1;
      if (gvm_grad_finish(e, c, I_dev, ksplit, flag_opt, normalize, result_dchi2_dev)) return 1;
    }
  }
  if (reduce) {
    if (gvm_dist_allreduce_f32(e, e->dist_grad, MN2)) return 1;
    k_add_inplace<<<(int)((MN2 + 255) / 256), 256, 0, e->stream>>>(caller_result, e->dist_grad, (long)MN2);
    GVM_LAUNCH(e);
    GVM_CUDA(cudaGetLastError());
  }
  return 0;
}

int gvm_eval_host(gvm_engine* e, const float* I_host, int flag_opt, int normalize, float* chi2_out,
                  float* grad_host) {
  const size_t MN = (size_t)e->cfg.M * e->cfg.N;
  GVM_CUDA(cudaSetDevice(e->cfg.device));
  e->flag_opt = flag_opt;
  GVM_CUDA(cudaMemcpyAsync(e->I_stage, I_host, 2 * MN * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  if (gvm_chi2_async(e, e->I_stage, normalize, e->red_out)) return 1;
  GVM_CUDA(cudaMemsetAsync(e->grad_stage, 0, 2 * MN * sizeof(float), e->stream));  // Chi2::restartDGi
  if (gvm_dchi2(e, e->I_stage, flag_opt, normalize, e->grad_stage)) return 1;
  GVM_CUDA(cudaMemcpyAsync(grad_host, e->grad_stage, 2 * MN * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  GVM_CUDA(cudaMemcpyAsync(e->h_red, e->red_out, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  GVM_CUDA(cudaStreamSynchronize(e->stream));
  if (chi2_out) *chi2_out = (float)e->h_red[0];
  return 0;
}

// ------------------------------------------------------------ device memory
int gvm_dev_alloc(gvm_engine* e, size_t bytes, void** out) {
  GVM_CUDA(cudaSetDevice(e->cfg.device));
  const size_t want = ((bytes ? bytes : 4) + 255) & ~(size_t)255;
  void* p = nullptr;
  // exact-size reuse (the callers' sizes repeat), most recently freed block first
  auto hit = e->pool_free.upper_bound(want);
  if (hit != e->pool_free.begin() && std::prev(hit)->first == want) {
    --hit;
    p = hit->second;
    e->pool_free.erase(hit);
  } else {
    // a miss costs a cudaMalloc (6-9 ms each on this platform, 13 of them per optimize() call): take a slab of
    // eight blocks of this size at once and keep the other seven for the callers that follow (the optimizers and
    // the Fi terms allocate runs of image-sized buffers of two sizes)
    size_t count = want <= ((size_t)256 << 20) ? 8 : 1;
    if (cudaMalloc(&p, count * want) != cudaSuccess) {
      cudaGetLastError();
      count = 1;
      GVM_CUDA(cudaMalloc(&p, want));
    }
    e->pool_slabs.push_back(p);
    for (size_t k = 1; k < count; k++) e->pool_free.emplace(want, static_cast<char*>(p) + k * want);
  }
  e->pool_live[p] = want;
  GVM_CUDA(cudaMemsetAsync(p, 0, want, e->stream));
  *out = p;
  return 0;
}
int gvm_dev_free(gvm_engine* e, void* p) {
  if (!p) return 0;
  auto it = e->pool_live.find(p);
  if (it == e->pool_live.end()) { gvm_set_error("gvm_dev_free: %p was not allocated by gvm_dev_alloc", p); return 1; }
  // stream-ordered reuse: every user of the block works on the engine stream, so a later
  // gvm_dev_alloc may hand it out again without a device synchronisation
  e->pool_free.emplace(it->second, p);
  e->pool_live.erase(it);
  return 0;
}
int gvm_dev_memset(gvm_engine* e, void* p, int value, size_t bytes) {
  GVM_CUDA(cudaMemsetAsync(p, value, bytes, e->stream));
  return 0;
}
int gvm_dev_copy(gvm_engine* e, void* dst, const void* src, size_t bytes, int kind) {
  GVM_CUDA(cudaSetDevice(e->cfg.device));
  if (kind == GVM_COPY_H2D) return gvm_fast_h2d(dst, src, bytes, e->stream);   // synchronous, pipelined for large pageable buffers
  if (kind == GVM_COPY_D2H) return gvm_fast_d2h(dst, src, bytes, e->stream);
  GVM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, e->stream));
  return 0;
}

// --------------------------------------------------------------- CUDA graphs
int64_t gvm_state_epoch(gvm_engine* e) { return e->epoch; }
int gvm_graph_begin(gvm_engine* e) {
  if (e->world > 1) { gvm_set_error("gvm_graph_begin: multi-rank engines are not captured (NCCL collectives inside)"); return 1; }
  if (e->capturing) { gvm_set_error("gvm_graph_begin: already capturing"); return 1; }
  GVM_CUDA(cudaSetDevice(e->cfg.device));
  GVM_CUDA(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
  e->capturing = true;
  e->fork_valid = false;
  e->join_pending = false;
  return 0;
}
int gvm_graph_end(gvm_engine* e, void** graph_exec_out) {
  if (!e->capturing) { gvm_set_error("gvm_graph_end: no capture in progress"); return 1; }
  gvm_join_branch(e);          // a forked prior-value branch must be back on the origin stream
  e->capturing = false;
  e->fork_valid = false;
  cudaGraph_t graph = nullptr;
  cudaError_t err = cudaStreamEndCapture(e->stream, &graph);
  if (err != cudaSuccess || !graph) {
    gvm_set_error("gvm_graph_end: capture failed (%s): a captured call allocated or synchronised", cudaGetErrorString(err));
    cudaGetLastError();
    return 1;
  }
  cudaGraphExec_t exec = nullptr;
  err = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (err != cudaSuccess) { gvm_set_error("gvm_graph_end: cudaGraphInstantiate -> %s", cudaGetErrorString(err)); return 1; }
  *graph_exec_out = exec;
  return 0;
}
int gvm_graph_launch(gvm_engine* e, void* graph_exec) {
  GVM_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_exec, e->stream));
  GVM_LAUNCH(e);
  return 0;
}
int gvm_graph_destroy(gvm_engine*, void* graph_exec) {
  if (graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)graph_exec);
  return 0;
}

int64_t gvm_launch_count(gvm_engine* e) { return e->launches; }
int gvm_last_grad_mode(gvm_engine* e) { return e->last_grad_mode; }
int gvm_grad_plan(gvm_engine* e, int* ntiles, int64_t* pixels) {
  if (ntiles) *ntiles = e->plan_ntiles;
  if (pixels) *pixels = e->plan_pixels;
  return 0;
}
int gvm_last_grad_kernel_ms(gvm_engine* e, float* ms, int* launches) {
  float total = 0.f;
  GVM_CUDA(cudaStreamSynchronize(e->stream));
  for (int i = 0; i + 1 < e->ev_used; i += 2) {
    float t = 0.f;
    GVM_CUDA(cudaEventElapsedTime(&t, e->ev[i], e->ev[i + 1]));
    total += t;
  }
  if (ms) *ms = total;
  if (launches) *launches = e->ev_used / 2;
  return 0;
}

}  // extern "C"
