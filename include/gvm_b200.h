/* gvm_b200.h — C-ABI of the B200-native objective/gradient engine.
 *
 * gpuvmem has no FFI: its plugin surface is C++ abstract classes compiled into
 * one executable. This header is the boundary a maintainer
 * would bind those classes to; every entry point names the reference interface
 * it replaces (paths relative to the reference tree). The C++ adapters that
 * reproduce the reference's class surface on top of it live in
 * gpuvmem_b200/csrc/host/ (see INTEGRATION.md).
 *
 * Conventions kept from the reference: images are row-major fp32
 * I[image][i*N + j] with i = row (y), j = column (x), image 0 = I_nu0 (in units
 * of fg_scale), image 1 = alpha; M = NAXIS1, N = NAXIS2 and M == N is required
 * (the reference assumes it, src/functions.cu:2574-2588). All `*_dev` pointers
 * are device pointers on the engine's GPU; everything else is host memory.
 * Every function returns 0 on success, non-zero on error (gvm_last_error()).
 * There is no CPU fallback: without a CUDA device gvm_create fails.
 */
#ifndef GVM_B200_H
#define GVM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gvm_engine gvm_engine;

enum { GVM_BEAM_AIRYDISK = 0, GVM_BEAM_GAUSSIAN = 1 }; /* include/MSFITSIO.cuh:56 */

/* Which gradient kernel gvm_dchi2 runs. AUTO picks UMMA when the separable
 * w-term bound holds, else SIMT_EXACT. */
enum {
  GVM_GRAD_AUTO = 0,
  GVM_GRAD_UMMA = 1,       /* tcgen05 / TMEM, error-compensated split: fp16 + two 8-bit-float correction products */
  GVM_GRAD_SIMT = 2,       /* separable outer-product on CUDA cores, fp32 */
  GVM_GRAD_SIMT_EXACT = 3, /* per-pair phase incl. the full w-term (reference formula) */
  GVM_GRAD_GRIDFFT = 4     /* samples on uv-cell centres with w = 0 (gridded mode, -g): the DFT
                              is evaluated exactly as one inverse FFT; AUTO picks it when it applies */
};

/* Globals of the reference that the hot path reads (src/functions.cu:37-71,
 * defined src/mfs.cu:4-45). */
typedef struct gvm_config {
  int64_t M, N;          /* image size; M == N */
  double DELTAX, DELTAY; /* CDELT1, CDELT2 in degrees (DELTAX < 0 for RA) */
  float nu_0;            /* reference frequency (Hz), -F */
  float eta;             /* -e, default -1 */
  float minpix;          /* initial_values[0] = -eta * (-z value)  (src/mfs.cu:169) */
  float noise_cut;       /* ALREADY scaled by min(noise) as src/mfs.cu:916 does */
  float threshold;       /* -T * 5 (src/mfs.cu:107) */
  float fg_scale;        /* min(noise image) (src/mfs.cu:912) or 1 if normalize */
  int device;            /* CUDA device ordinal ("firstgpu") */
  int grad_mode;         /* GVM_GRAD_* */
  int keep_vm;           /* also store model visibilities Vm (residual write-back) */
} gvm_config;

/* One (field, channel, stokes) block of visibilities, i.e. one entry of
 * fields[f].visibilities[i][s] (include/MSFITSIO.cuh:82-121) plus the antenna
 * beam model of datasets[d].antennas[0] (include/MSFITSIO.cuh:123-131). */
typedef struct gvm_channel_desc {
  float freq;            /* fields[f].nu[i] (float, as the reference stores it) */
  float antenna_diameter, pb_factor, pb_cutoff;
  int primary_beam;      /* GVM_BEAM_* */
  float ref_xobs_pix, ref_yobs_pix; /* pointing centre (attenuation) */
  float phs_xobs_pix, phs_yobs_pix; /* phase centre (phase_rotate, DChi2) */
} gvm_channel_desc;

const char* gvm_last_error(void);
int gvm_version(void);

/* ------------------------------------------------------------- lifecycle -- */
/* Replaces the device-side part of MFS::setDevice (src/mfs.cu:530-916): per-GPU
 * scratch (varsPerGPU, include/framework.cuh:49-55), cuFFT plan (initFFT,
 * src/functions.cu:2142). */
int gvm_create(const gvm_config* cfg, gvm_engine** out);
int gvm_destroy(gvm_engine* e);
/* Use the caller's CUDA stream (cudaStream_t as void*) for all engine work. */
int gvm_set_stream(gvm_engine* e, void* cuda_stream);
void* gvm_get_stream(gvm_engine* e);
int gvm_synchronize(gvm_engine* e);
/* Update scalars that the reference mutates between runs (fg_scale:
 * Chi2::setFgScale src/chi2.cu:74; noise_cut; threshold). */
int gvm_set_scalars(gvm_engine* e, float fg_scale, float noise_cut, float threshold);
int gvm_set_grad_mode(gvm_engine* e, int grad_mode);
/* The optimizers' schedule flag `flag_opt` (src/frprmn.cu:46, set from
 * Optimizer::setFlag): read by the clip in gvm_chi2; gvm_dchi2 also sets it. */
int gvm_set_flag_opt(gvm_engine* e, int flag_opt);

/* device_noise_image (src/mfs.cu:28; built src/mfs.cu:850-916). M*N floats. */
int gvm_set_noise_image(gvm_engine* e, const float* noise, int src_is_device);
/* Builds the noise image on the GPU the way MFS::setDevice does
 * (total_attenuation -> weight_image -> noise_image, src/functions.cu:2382-2422,
 * src/mfs.cu:850-916) from the channels already added; returns min(noise) =
 * fg_scale. Does NOT rescale noise_cut (the caller does, as src/mfs.cu:916). */
int gvm_build_noise_image(gvm_engine* e, float noise_jypix, float* fg_scale_out);
/* Same with the fields named explicitly: the reference adds attenuation^2 ONCE PER (dataset, field), in
 * dataset-then-field order and without de-duplication (src/mfs.cu:870-888) — and the list must not depend on
 * which blocks were sharded to this rank, or the replicas of a multi-rank job would mask different pixels.
 * Only the beam / pointing members of each descriptor are read (freq is ignored: the pattern is taken at nu_0).
 * gvm_build_noise_image derives the list from the uploaded blocks (one entry per distinct pointing + beam). */
int gvm_build_noise_image_fields(gvm_engine* e, float noise_jypix, int nfields,
                                 const gvm_channel_desc* fields, float* fg_scale_out);
int gvm_get_noise_image(gvm_engine* e, float* noise_host);
/* Gridding-correction image (CKernel::getGCFGPU, include/classes/ckernel.cuh:57);
 * NULL disables it (ip->getCKernel() == NULL, src/functions.cu:4358). */
int gvm_set_gcf(gvm_engine* e, const float* gcf_host);

/* Forward-model option: convolutional-kernel degridding instead of the bilinear
 * vis_mod. This is what the reference sketches in degriddingGPU
 * (src/functions.cu:2205-2254; defined, never launched — its call site is
 * commented out at :2127): nearest uv cell by the gridding index rule
 * (:2222-2223), then the sum over the (2*support_y+1) x (2*support_x+1) taps of
 * the CKernel table (CKernel::getGPUKernel / getm / getn / getSupportX/Y,
 * include/classes/ckernel.cuh) times the model grid; with gvm_set_gcf holding the
 * kernel's GCF image (apply_GCF, :2468) the pair is a gridding-corrected degridder.
 * table_host [m][n]; NULL restores the reference's bilinear interpolation. The
 * gradient stays the exact DFT of the residuals (DChi2), as in the reference. */
int gvm_set_degrid_kernel(gvm_engine* e, const float* table_host, int m, int n,
                          int support_x, int support_y);
/* How gvm_chi2 transforms the image. FULL is the reference's pipeline: complex
 * pre-FFT image, cuFFT C2C inverse (src/functions.cu:2165), phase_rotate over the
 * whole grid (:2483). HALF uses that the pre-FFT image is real: cuFFT R2C onto the
 * half plane (conjugated on read) and the phase rotation applied to the four
 * bilinear taps of every sample — half the FFT bytes and no image-sized rotation
 * pass; same float arithmetic per tap. AUTO picks HALF for blocks with
 * 4 Z <= M N (gridded data). Not combined with gvm_set_degrid_kernel. */
enum { GVM_FORWARD_AUTO = 0, GVM_FORWARD_FULL = 1, GVM_FORWARD_HALF = 2 };
int gvm_set_forward_mode(gvm_engine* e, int mode);
int gvm_last_forward_mode(gvm_engine* e);   /* GVM_FORWARD_FULL or GVM_FORWARD_HALF */
/* The model grid of the LAST channel evaluated by gvm_chi2 (after phase_rotate):
 * [M][N] complex as interleaved floats, DC at [0,0] — device_V of varsPerGPU
 * (include/framework.cuh:49-55). For tests and diagnostics. */
int gvm_get_model_grid(gvm_engine* e, float* V_host);

/* Upload one visibility block: raw uvw in METRES as [Z][3] doubles, Vo as [Z][2]
 * floats, weights [Z]. Performs on the GPU what MFS::setDevice + the
 * hermitianSymmetry kernel do (src/mfs.cu:555-617, src/functions.cu:2256-2273):
 * u>0 -> (u,v) negated, Vo conjugated (w untouched), metres -> lambda with the
 * fp32 wavelength (src/MSFITSIO.cu:36-45); and precomputes the vis_mod cell
 * index / fractions (src/functions.cu:2569-2586), zeroing the weight of
 * out-of-grid samples as vis_mod does (:2607). Returns the channel slot. */
int gvm_add_channel(gvm_engine* e, const gvm_channel_desc* desc, int64_t Z,
                    const double* uvw_m, const float* Vo, const float* w,
                    int* chan_out);
/* When `chan` holds only a slice of a (field, channel, stokes) block (visibility-chunk sharding over ranks),
 * the size of the WHOLE block: chi2() and DChi2 normalise by numVisibilitiesPerFreqPerStoke of the block
 * (src/functions.cu:4439-4441, :3785), not by the slice. Default: the uploaded Z. */
int gvm_set_block_nvis(gvm_engine* e, int chan, int64_t Z_block);
/* Drop every uploaded block (MFS::writeResiduals re-uploads the ungridded samples after a
 * gridded run, src/mfs.cu:1118-1139). */
int gvm_clear_channels(gvm_engine* e);
int gvm_num_channels(gvm_engine* e);
int64_t gvm_channel_nvis(gvm_engine* e, int chan);
/* Device-side state of a block after upload / a forward pass; any pointer may be
 * NULL. uvw_lambda [Z][3] doubles, cell [Z][2] int32 (i1, j1), Vo/Vm/Vr [Z][2]. */
int gvm_get_vis(gvm_engine* e, int chan, double* uvw_lambda, int32_t* cell,
                float* Vo, float* Vm, float* Vr, float* w);

/* -------------------------------------------------------------- hot path -- */
/* chi2() (src/functions.cu:4323-4454) incl. ip->clipWNoise (clip2IWNoise, :2694 —
 * MUTATES I_dev like the reference), calculateInu (:3939), apply_beam2I (:2424),
 * apply_GCF (:2468), cuFFT inverse C2C (:2165), phase_rotate (:2483), vis_mod
 * (:2557), residual (:2663), chi2Vector (:2867), deviceReduce (:605).
 * I_dev: [2][M][N]. Leaves Vr per block for gvm_dchi2. chi2_out = 0.5*sum. */
int gvm_chi2(gvm_engine* e, float* I_dev, int normalize, float* chi2_out);
/* Same, but the result stays on the device (no host sync); for graphs. */
int gvm_chi2_async(gvm_engine* e, float* I_dev, int normalize, double* chi2_dev);

/* dchi2() (src/functions.cu:4456-4558): per block DChi2 (:3698 / :3793 with GCF)
 * then DChi2_total_I_nu_0 (:4000, flag_opt even) or DChi2_total_alpha (:3968,
 * flag_opt odd) ACCUMULATED (+=) into result_dchi2_dev [2][M][N], as the
 * reference does (Chi2::restartDGi zeroes it first, src/chi2.cu:55). */
int gvm_dchi2(gvm_engine* e, const float* I_dev, int flag_opt, int normalize,
              float* result_dchi2_dev);

/* End-to-end convenience used by bench.py's e2e leg and the smoke test: host
 * image in (pinned or pageable), H2D, gvm_chi2, zero + gvm_dchi2, D2H of the
 * gradient [2][M][N] and chi2. */
int gvm_eval_host(gvm_engine* e, const float* I_host, int flag_opt, int normalize,
                  float* chi2_out, float* grad_host);

/* calculateErrors() (src/functions.cu:4966-5040; Error "SecondDerivateError",
 * src/secondderivateerror.cu:6-10, run by MFS::writeImages under -E,
 * src/mfs.cu:1090-1113): per block I_nu_0_Noise (:4076) and alpha_Noise (:4113,
 * a direct DFT over the residuals Vr of the last gvm_chi2 — evaluated on the same
 * contraction kernels as gvm_dchi2), then noise_reduction (:4179).
 * errors_dev: [2][M][N], overwritten. dist_mode says what the blocks of a
 * multi-rank engine are: slices of the same blocks on every rank (sums are
 * completed per block, every rank gets the full result) or disjoint blocks (one
 * all-reduce of the accumulated maps). Ignored when world == 1. */
enum { GVM_DIST_NONE = 0, GVM_DIST_CHUNKS = 1, GVM_DIST_BLOCKS = 2 };
int gvm_error_maps(gvm_engine* e, const float* I_dev, int dist_mode, float* errors_dev);

/* --------------------------------------------------------------- priors --- */
enum {
  GVM_PRIOR_ENTROPY = 0,   /* SEntropy/DEntropy  src/functions.cu:4722/4744 */
  GVM_PRIOR_L1 = 1,        /* L1Norm/DL1Norm     :4633/4655 */
  GVM_PRIOR_TV = 2,        /* totalvariation/DTVariation :4886/4907 */
  GVM_PRIOR_TSV = 3,       /* TotalSquaredVariation/DTSVariation :4927/4947 */
  GVM_PRIOR_LAPLACIAN = 4, /* laplacian/DLaplacian :4808/4828 */
  GVM_PRIOR_QUADRATIC = 5, /* quadraticP/DQuadraticP :4847/4867 */
  GVM_PRIOR_GENTROPY = 6,  /* SGEntropy/DGEntropy :4765/4787 (prior image) */
  GVM_PRIOR_GL1 = 7        /* GL1NormK/DGL1Norm :4675/4700 (prior image) */
};
typedef struct gvm_prior_params {
  float prior_value;  /* Entropy G (Fi::setPrior(float)) */
  float eta;          /* Entropy eta */
  float epsilon;      /* L1 (1e-12, src/l1norm.cu:10) / TV epsilon / GL1 epsilon_a */
  float epsilon_b;    /* GL1 */
  const float* prior_image_dev; /* GEntropy / GL1Norm prior image, M*N, device */
} gvm_prior_params;

/* The value host functions: sum over the image of the per-pixel term, masked by
 * noise < noise_cut. The reference gates on (iter > 0 && lambda != 0)
 * (e.g. :4643); the gate is the caller's (Fi adapter) business here. */
int gvm_prior_value(gvm_engine* e, int kind, const float* I_dev, int image_index,
                    const gvm_prior_params* p, float* value_out);
/* One host synchronisation for ALL terms of an objective evaluation (the reference synchronises, copies
 * and frees inside every deviceReduce, i.e. several times per Fi): each term launches into a result slot
 * of the engine, gvm_fetch_slots copies the first n slots back with a single stream synchronisation.
 * gvm_chi2_to_slot = gvm_chi2 (same side effects), gvm_prior_value_to_slot = gvm_prior_value. */
#define GVM_OBJ_SLOTS 16
int gvm_chi2_to_slot(gvm_engine* e, float* I_dev, int normalize, int slot);
int gvm_prior_value_to_slot(gvm_engine* e, int kind, const float* I_dev, int image_index,
                            const gvm_prior_params* p, int slot);
int gvm_fetch_slots(gvm_engine* e, int n, double* values_out);
/* The two halves of gvm_fetch_slots: enqueue the device->host copy of the first n slots on the engine stream
 * (capturable), and wait for it + read the values. */
int gvm_fetch_slots_enqueue(gvm_engine* e, int n);
int gvm_fetch_slots_wait(gvm_engine* e, int n, double* values_out);

/* One CUDA graph per objective evaluation (the reference launches ~10 kernels and synchronises several times per
 * term in every line-search probe, src/f1dim.cu:49-80): everything the engine enqueues on its stream between
 * gvm_graph_begin and gvm_graph_end — gvm_chi2_to_slot, gvm_prior_value_to_slot, gvm_fetch_slots_enqueue — is
 * captured instead of executed and instantiated; gvm_graph_launch replays it with ONE launch. The captured calls
 * must not allocate or synchronise: run the same sequence once un-captured first (beam planes, FFT plans are
 * created on first use). A graph bakes in every argument (image pointer, scalars, flag_opt) and the engine's
 * buffers: gvm_state_epoch changes whenever a call invalidates them (blocks added or cleared, noise image, GCF,
 * degridding kernel, forward mode); compare it before replaying. Single-rank engines only. */
int gvm_graph_begin(gvm_engine* e);
int gvm_graph_end(gvm_engine* e, void** graph_exec_out);
int gvm_graph_launch(gvm_engine* e, void* graph_exec);
int gvm_graph_destroy(gvm_engine* e, void* graph_exec);
int64_t gvm_state_epoch(gvm_engine* e);
/* The gradient host functions: dgi_dev[M*N] = lambda * d(term)/dI, written (not
 * accumulated), like DS/DL1NormK/... write device_DS. */
int gvm_prior_grad(gvm_engine* e, int kind, const float* I_dev, int image_index,
                   const gvm_prior_params* p, float lambda, float* dgi_dev);
/* The two steps above in one pass: dphi[image_to_add] += lambda * d(term)/dI without the round trip through a
 * per-term device_DS buffer (the gradient kernel + AddToDPhi pair of every prior, src/functions.cu:3890) — the same
 * fp32 value is added, so the result is bit-identical. dphi_dev: [2][M][N]. */
int gvm_prior_grad_add(gvm_engine* e, int kind, const float* I_dev, int image_index, const gvm_prior_params* p,
                       float lambda, float* dphi_dev, int image_to_add);
/* linkAddToDPhi / AddToDPhi (src/functions.cu:4560/3890): dphi[index] += dgi. */
int gvm_add_to_dphi(gvm_engine* e, float* dphi_dev, const float* dgi_dev, int index);

/* ------------------------------------------------- optimizer vector ops ---
 * Image-sized kernels the optimizers launch (src/functions.cu:2721-2865,
 * 3556-3687), one fused launch each, reductions finished on the device. */
/* evaluateXt / evaluateXtNoPositivity (:2832/:2853) for all images:
 * xt = pcom + x*xicom, image 0 floored at -eta*minpix unless nopositivity. */
int gvm_vec_evaluate_xt(gvm_engine* e, float* xt, const float* pcom,
                        const float* xicom, float x, int image_count, int nopositivity);
/* newP / newPNoPositivity (:2779/:2801): xi *= xmin; p += xi with projection
 * (xi zeroed where clipped). */
int gvm_vec_new_p(gvm_engine* e, float* p, float* xi, float xmin, int image_count,
                  int nopositivity);
/* sum(a*b) over n floats, fp32 pairwise -> fp64 final. */
int gvm_vec_dot(gvm_engine* e, const float* a, const float* b, int64_t n, float* out);
/* getGGandDGG + two deviceReduce (src/frprmn.cu:157-170): gg = sum g*g,
 * dgg = sum (xi+g)*xi over all images. */
int gvm_vec_gg_dgg(gvm_engine* e, const float* xi, const float* g, int image_count,
                   float* gg, float* dgg);
/* CGGradCondition + deviceMaxReduce (src/frprmn.cu:139-148):
 * max |xi|*max(|p|,1)/den. */
int gvm_vec_grad_condition(gvm_engine* e, const float* xi, const float* p, float den,
                           int image_count, float* gmax);
/* searchDirection (:3655): g = -xi; xi = h = g.  newXi (:3678): g = -xi;
 * xi = h = g + gam*h. (gam = 0 and first = 1 give searchDirection.) */
int gvm_vec_new_xi(gvm_engine* e, float* g, float* xi, float* h, float gam,
                   int image_count);
/* y = a*x + b*y over n floats (L-BFGS pieces: updateQ :3613, getR :3625). */
int gvm_vec_axpby(gvm_engine* e, float a, const float* x, float b, float* y, int64_t n);
/* normArray + deviceMaxReduce (src/lbfgs.cu:151-160; kernel :3590): max |v|. */
int gvm_vec_absmax(gvm_engine* e, const float* v, int64_t n, float* out);
/* searchDirection_LBFGS (:3564): v *= s. */
int gvm_vec_scale(gvm_engine* e, float* v, float s, int64_t n);
/* calculateSandY (:3636): y_out = xi - (-1*xi_old), s_out = p - p_old over n floats. */
int gvm_vec_lbfgs_sy(gvm_engine* e, float* y_out, float* s_out, const float* xi,
                     const float* xi_old, const float* p, const float* p_old, int64_t n);

/* ------------------------------------------------------- device memory -----
 * The engine owns device memory for its callers (the reference's classes call
 * cudaMalloc/cudaMemset/cudaMemcpy directly: include/classes/fi.cuh:84-88,
 * src/frprmn.cpp allocateMemoryGpu, objectivefunction.cuh:80-85). Allocations are
 * zero-filled like the reference's malloc+memset pairs. Copies are ordered on
 * the engine stream; H2D/D2H return after completion. */
enum { GVM_COPY_H2D = 0, GVM_COPY_D2H = 1, GVM_COPY_D2D = 2 };
int gvm_dev_alloc(gvm_engine* e, size_t bytes, void** out);
int gvm_dev_free(gvm_engine* e, void* p);
int gvm_dev_memset(gvm_engine* e, void* p, int value, size_t bytes);
int gvm_dev_copy(gvm_engine* e, void* dst, const void* src, size_t bytes, int kind);

/* ------------------------------------------------------------ multi-GPU ----
 * One process per GPU. Every rank holds a shard of the visibility blocks
 * (whole channels, as the reference's `i % num_gpus` rule, src/functions.cu:4341,
 * or contiguous visibility chunks of a channel) and a replica of the image.
 * After gvm_dist_init, gvm_chi2 all-reduces the chi2 scalar and gvm_dchi2
 * all-reduces the [2][M][N] gradient (NCCL, sum) before adding it to
 * result_dchi2_dev — replacing the reference's serialised peer-to-peer
 * accumulate (:4534-4549). Rank 0 creates the id, the launcher distributes it. */
#define GVM_DIST_ID_BYTES 128
int gvm_dist_unique_id(char* id_out, size_t bytes);
int gvm_dist_init(gvm_engine* e, int rank, int world, const char* id, size_t bytes);
int gvm_dist_rank(gvm_engine* e);
int gvm_dist_world(gvm_engine* e);
/* Replicated mode: every rank holds ALL visibility blocks and computes the complete chi2 / gradient itself, so
 * gvm_chi2, gvm_dchi2 and gvm_error_maps skip their all-reduces. For gridded data (-g): the evaluation is then
 * image-sized work (FFTs, image kernels) that does not shard, while the all-reduce of the [2][M][N] gradient would
 * cost more than the few samples it saves; only the preprocessing (gvm_weights_dist / gvm_grid_block_dist) is
 * distributed. Identical inputs give identical results on every rank. */
int gvm_dist_set_replicated(gvm_engine* e, int on);
int gvm_dist_allreduce(gvm_engine* e, float* buf_dev, int64_t n);
/* Broadcast n floats from `root` on the engine stream (the image replica of a multi-rank job: one rank
 * uploads it, the others receive it over NVLink instead of each pulling it over PCIe). */
int gvm_dist_broadcast(gvm_engine* e, float* buf_dev, int64_t n, int root);
int64_t gvm_dist_collectives(gvm_engine* e);
/* A rank that must bail out (a local failure) calls this first: the communicator is aborted, so the peers'
 * pending collectives return an error instead of waiting forever. The engine does it itself when gvm_chi2,
 * gvm_dchi2 or gvm_error_maps fail on a multi-rank engine. */
int gvm_dist_abort(gvm_engine* e);

/* ------------------------------------------------- weights and gridding ---
 * WeightingScheme::apply (src/{natural,uniform,briggs,radial}weightingscheme.cu)
 * for ONE dataset of `nblocks` (field,channel,stokes) blocks in the reference's
 * loop order. uvw_m[b] -> [Z[b]][3] doubles in metres, w[b] updated in place.
 * All host pointers; the arithmetic runs on the GPU. */
enum { GVM_W_NATURAL = 0, GVM_W_UNIFORM = 1, GVM_W_BRIGGS = 2, GVM_W_RADIAL = 3 };
typedef struct gvm_taper {           /* UVTaper::getValue, include/classes/uvtaper.cuh:100 */
  int enabled;
  float sigma_maj, sigma_min, bpa, amplitude;
  double u_0, v_0;
} gvm_taper;
int gvm_weights(int device, int scheme, float robust, int64_t M, int64_t N,
                double deltau, double deltav, int nblocks, const int64_t* Z,
                const double* const* uvw_m, const float* freqs, float* const* w,
                const gvm_taper* taper);

/* do_gridding (src/functions.cu:1339-1653) for one block: convolutional
 * gridding of the Hermitian-doubled samples with an m x n CKernel table, then
 * w_eff = gw^2/gw2, V = gV/gw, row-major compaction of cells with w > 0.
 * Outputs are written to caller-provided host arrays sized M*N (upper bound);
 * *nout = number of gridded samples. uvw_out in METRES (cell centres), w = 0. */
int gvm_grid_block(int device, int64_t M, int64_t N, double deltau, double deltav,
                   float freq, int64_t Z, const double* uvw_m, const float* Vo,
                   const float* w, const float* ckernel, int ck_m, int ck_n,
                   int support_x, int support_y, double* uvw_out, float* Vo_out,
                   float* w_out, int64_t* nout);
/* Two-phase use for large grids: call gvm_grid_block with the three output pointers NULL to get
 * *nout, size the host arrays, then fetch the gridded samples of that last call (same thread). */
int gvm_grid_fetch(double* uvw_out, float* Vo_out, float* w_out);
/* Multi-rank versions on an engine that went through gvm_dist_init (grid size and cell size are the engine's).
 * EVERY rank passes the same full host arrays; rank r uploads and processes only the samples
 * [Z r / W, Z (r + 1) / W) of every block, and every rank ends with the complete result: the new weights of all
 * samples in w[b] / the gridded samples through gvm_grid_fetch. Bit-identical to gvm_weights / gvm_grid_block
 * (DESIGN.md §6): the per-cell weight sums travel down the ranks in sample order, the (tile, sample) pairs of the
 * gridding are exchanged all-to-all with the tile owners in ascending sample order. With world == 1 they are
 * gvm_weights / gvm_grid_block. */
int gvm_weights_dist(gvm_engine* e, int scheme, float robust, int nblocks, const int64_t* Z,
                     const double* const* uvw_m, const float* freqs, float* const* w, const gvm_taper* taper);
int gvm_grid_block_dist(gvm_engine* e, float freq, int64_t Z, const double* uvw_m, const float* Vo, const float* w,
                        const float* ckernel, int ck_m, int ck_n, int support_x, int support_y, int64_t* nout);
/* gvm_grid_block keeps its device work buffers between calls (this thread); this returns them. */
int gvm_grid_release(void);
/* Optional: size that work arena once, before the first gvm_weights* / gvm_grid_block* call of this thread, for blocks
 * of up to Zmax samples on `world` ranks — otherwise weighting allocates it for its own (smaller) need and gridding
 * frees and re-allocates it (a cudaMalloc of several GB costs 0.1-0.5 s here). The reference allocates its gridding
 * buffers per call on the host (src/functions.cu:1339-1416). */
int gvm_grid_reserve(int device, int64_t M, int64_t N, int64_t Zmax, int world);

/* The engine's own stable radix sort (csrc/sort.cu: the tile-sorted upload of gvm_add_channel and the gridding
 * path order their samples with it instead of a library sort) on host arrays, in place: n (key, value) pairs by
 * the low key_bits bits of the key, equal keys keep their input order. For tests. */
int gvm_sort_pairs_host(int device, uint32_t* keys, uint32_t* vals, int64_t n, int key_bits);

/* ------------------------------------------------------------- telemetry -- */
/* Number of kernels (ours + cuFFT) launched by the engine since creation. */
int64_t gvm_launch_count(gvm_engine* e);
/* Device time (ms, CUDA events on the engine stream) of the dominant gradient
 * kernel over the last gvm_dchi2 call, and how many launches it comprised. */
int gvm_last_grad_kernel_ms(gvm_engine* e, float* ms, int* launches);
/* Tile plan of the tensor-core gradient after the last gvm_dchi2: number of 256-row tiles that
 * cover the unmasked pixels (DChi2 skips masked pixels, src/functions.cu:3723-3726) and the
 * number of output pixels they contain (algorithmic flops of one launch = 4 * pixels * Z). */
int gvm_grad_plan(gvm_engine* e, int* ntiles, int64_t* pixels);
/* Which kernel the last gvm_dchi2 used (GVM_GRAD_*). */
int gvm_last_grad_mode(gvm_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* GVM_B200_H */
