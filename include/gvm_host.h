/* gvm_host.h — C entry points of the C++ host layer (gpuvmem_b200/libgvmhost.so).
 *
 * The host layer restates gpuvmem's plugin surface in C++ (headers under gpuvmem_b200/csrc/host:
 * Synthesizer "MFS", Optimizer "CG-FRPRMN"/"CG-LBFGS", ObjectiveFunction, the Fi terms,
 * CKernel and WeightingScheme families, the string-keyed factories and the reference's
 * command line) on top of the engine's C ABI (gvm_b200.h). A C++ program uses the classes
 * directly (gpuvmem_b200/csrc/host/main.cpp is the reference's src/main.cu:100-229 on this
 * layer); these functions expose the same flows to non-C++ callers (tests, bench.py).
 * Every function returns 0 on success; errors that the reference turns into print+exit
 * do the same here.
 */
#ifndef GVM_HOST_H
#define GVM_HOST_H

#include <stddef.h>
#include <stdint.h>

#include "gvm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gvmh_session gvmh_session;

/* What readMS + readFITSHeader hand to MFS::configure (src/MSFITSIO.cu:398-754,
 * include/MSFITSIO.cuh:140-150): one dataset, one field, one correlation (XX). */
typedef struct gvmh_problem {
  int64_t M, N;
  double DELTAX, DELTAY;   /* CDELT1, CDELT2 (deg) */
  double ra, dec;          /* CRVAL1, CRVAL2 (deg) = phase centre */
  double crpix1, crpix2;
  const char* telescope;   /* "ALMA", "EVLA", other */
  float antenna_diameter;
  float beam_noise;        /* NOISE header keyword, <= 0: estimate */
  int nchan;
  const float* freqs;      /* [nchan] */
  const int64_t* Z;        /* [nchan] */
  const double* const* uvw_m; /* [nchan] -> [Z][3] metres */
  const float* const* Vo;     /* [nchan] -> [Z][2] */
  const float* const* w;      /* [nchan] -> [Z] */
  int has_field_centre;       /* 0: the field's phase/pointing centre is (ra, dec) */
  double field_ra, field_dec; /* deg: Field::phs_ra/phs_dec = ref_ra/ref_dec when it is not the image centre */
} gvmh_problem;

/* `problem` may be NULL: the visibilities and the image header are then read from the files named by
 * -i / -m in `args` (GVMS container / FITS model image), as the command-line program does. */
/* Everything src/main.cu:147-212 does up to (not including) sy->run():
 * factories -> MFS::configure(args) -> setDevice -> Fi terms -> ObjectiveFunction.
 *   args      the reference's command line as one string, e.g.
 *             "-z 0.001 -Z 0.01,0.0,1e-4 -t 50 -R 0.0 -g 0 -v"
 *   optimizer "CG-FRPRMN" | "CG-LBFGS"; scheme "Natural"|"Uniform"|"Briggs"|"Radial";
 *   ckernel   "PillBox2D"|"Gaussian2D"|"Sinc2D"|"GaussianSinc2D"|"PSWF" with size ck_m x ck_n
 *   fi_spec   comma list of name:penalizatorIndex:imageIndex:imageToAdd[:normalize], NULL = main.cu's
 *             "Chi2:-1:0:0,Entropy:0:0:0,L1-Norm:1:0:0,TotalSquaredVariation:2:0:0,Laplacian:3:0:0"
 *   rank/world/nccl_id  one process per GPU; nccl_id from gvm_dist_unique_id on rank 0 */
int gvmh_create(const gvmh_problem* p, const char* args, const char* optimizer, const char* scheme,
                const char* ckernel, int ck_m, int ck_n, const char* fi_spec, int rank, int world,
                const char* nccl_id, gvmh_session** out);
int gvmh_destroy(gvmh_session* s);
/* 1: the host layer prints nothing to stdout (the reference's progress text), for callers that
 * own stdout. Applies to sessions created afterwards and to the current one. */
int gvmh_set_quiet(int quiet);

/* Synthesizer::run (+ the default optimisation order of main.cu: flag 0 only). image_out
 * [2][M][N] host, optional. */
int gvmh_run(gvmh_session* s, float* image_out, double* optimize_seconds);
int gvmh_clear_run(gvmh_session* s);
int gvmh_set_lbfgs_k(gvmh_session* s, int k);
int gvmh_write_outputs(gvmh_session* s);   /* writeImages + writeResiduals */
/* Forward-model option: use the session's CKernel (gvmh_create's ckernel / ck_m / ck_n) as a convolutional DEGRIDDING
 * kernel on the ungridded samples — what the reference sketches in degriddingGPU (src/functions.cu:2205-2254) — with
 * its gridding-correction image (apply_GCF, :2468) in front of the FFT; the gradient stays the exact DFT. on = 0
 * restores the bilinear vis_mod. This is BASELINE config 4 read literally ("PSWF_12D degridding kernel"). */
int gvmh_use_ckernel_degridding(gvmh_session* s, int on);
/* MFS::writeResiduals alone (src/mfs.cu:1115-1155): weights restored / original samples brought back
 * (getOriginalVisibilitiesBack, src/functions.cu:1844-2010) and re-evaluated, then modelToHost
 * (src/MSFITSIO.cu:1114-1138). nongridded_chi2 (optional): the "Non-gridded chi2" of a gridded run, else 0.
 * Afterwards gvmh_nvis / gvmh_get_host_vis describe the samples that go to the output file and
 * gvmh_get_host_model returns their model Vm [Z][2] (conjugated back where u > 0) and residual Vo - Vm. */
int gvmh_write_residuals(gvmh_session* s, float* nongridded_chi2);
int gvmh_get_host_model(gvmh_session* s, int chan, float* Vm, float* Vr);
/* Error "SecondDerivateError" (src/secondderivateerror.cu:6-10 -> calculateErrors,
 * src/functions.cu:4966-5040) on the session's image and the residuals of the last objective
 * evaluation; errors_host [2][M][N]: sigma(I_nu0), sigma(alpha). */
int gvmh_error_image(gvmh_session* s, float* errors_host);
/* ONE Fi term of the factory ("Entropy", "L1-Norm", "TotalVariation", "TotalSquaredVariation", "Laplacian",
 * "Quadratic", "GEntropy", "GL1Norm") on its own, the way ObjectiveFunction drives it (include/classes/
 * objectivefunction.cuh; Fi::configure(-1, image_index, image_index, false), include/classes/fi.cuh:62-95):
 * calcFi -> value_out = get_fivalue(); restartDGi + calcGi + addToDphi into a zeroed dphi_out [2][M][N].
 * I_host [2][M][N]; prior_host M*N or NULL; flag = flag_opt; prior_after_out (optional, M*N): the term's prior image
 * after calcGi (see GL1Norm::calcGi). The session provides the engine, mask and image size. */
int gvmh_fi_eval(gvmh_session* s, const char* name, const float* I_host, const float* prior_host, float lambda,
                 float prior_value, float eta, float eps_a, float eps_b, int image_index, int iteration, int flag,
                 float* value_out, float* dphi_out, float* prior_after_out);
/* Filter "Gridding" (src/gridding.cu:13-28): do_gridding over the session's Visibilities, in place, with
 * the named CKernel (NULL/"": PillBox2D, the reference's default). The host-side samples change
 * (gvmh_get_host_vis); the engine keeps what was uploaded. */
int gvmh_filter_gridding(gvmh_session* s, const char* ckernel, int ck_m, int ck_n);

/* ObjectiveFunction::calcFunction / calcGradient on the session's device image. */
int gvmh_set_image(gvmh_session* s, const float* I_host);
int gvmh_get_image(gvmh_session* s, float* I_host);
int gvmh_set_iteration(gvmh_session* s, int iteration);  /* Fi::setIteration on every term */
int gvmh_set_flag(gvmh_session* s, int flag_opt);
int gvmh_calc_function(gvmh_session* s, float* value, float* fi_values, int nfi);
int gvmh_calc_gradient(gvmh_session* s, int iteration, float* grad_host /* [2][M][N] or NULL */);
/* One objective + gradient evaluation, device resident (bench `value`) ... */
int gvmh_eval_device(gvmh_session* s, int iteration, float* value);
/* ... and end to end: pinned/pageable host image in, gradient + value out (bench `e2e`). With several ranks
 * the image is uploaded by rank 0 and broadcast over NVLink, and only rank 0 receives the gradient
 * (I_host / grad_host may be NULL elsewhere); the value is returned on every rank. */
int gvmh_eval_host(gvmh_session* s, const float* I_host, int iteration, float* value, float* grad_host);

gvm_engine* gvmh_engine(gvmh_session* s);
/* out[16]: fg_scale, noise_cut, noise_jypix, nu_0, vis_noise, sum_weights, bmaj_deg, bmin_deg,
 * bpa_deg, deltau, deltav, xobs_pix, yobs_pix, total_visibilities, iterations_done, n_fi */
int gvmh_scalars(gvmh_session* s, double* out);
/* seconds6: setup, weighting, gridding, optimize (last run), time inside calcFunction, time inside
 * calcGradient (both cumulative); counts2: function evaluations, gradient evaluations */
int gvmh_stats(gvmh_session* s, double* seconds6, int64_t* counts2);
/* the visibilities as the hot path sees them after weighting (+ gridding): per channel */
int64_t gvmh_nvis(gvmh_session* s, int chan);
int gvmh_get_host_vis(gvmh_session* s, int chan, double* uvw_m, float* Vo, float* w);
const char* gvmh_exit_reason(gvmh_session* s);
int gvmh_history(gvmh_session* s, float* out, int cap);

/* --- stateless pieces (no GPU needed) ---------------------------------------------- */
/* CKernel::buildKernel table (m x n) with sigmas (sx, sy); w <= 0 keeps the family default. */
int gvmh_ckernel_table(const char* name, int m, int n, float sx, float sy, float w, float* table,
                       int* support_x, int* support_y);
/* CKernel::initializeGCF(M, N, dx, dy) image (M x N). */
int gvmh_ckernel_gcf(const char* name, int m, int n, int M, int N, float dx, float dy, float* gcf);
/* 1 if `name` is registered in the factory of `kind` ("Fi", "Optimizer", "CKernel",
 * "WeightingScheme", "Synthesizer", "Io", "ObjectiveFunction", "Error"). */
int gvmh_factory_has(const char* kind, const char* name);
/* The dependency-free FITS image reader/writer of the Io handlers (csrc/host/fits.hpp), for tests and
 * Python callers. header16: naxis1, naxis2, bitpix, has_wcs, CDELT1, CDELT2, CRVAL1, CRVAL2, CRPIX1, CRPIX2,
 * BMAJ, BMIN, BPA, NOISE (-1: absent), EQUINOX, number of header cards. data_out may be NULL. */
int gvmh_fits_read(const char* path, double* header16, float* data_out, int64_t cap);
/* OCopyFITS (src/MSFITSIO.cu:93-165): header of template_path (may be NULL) copied, BUNIT/NITER/NAXISn/
 * RADESYS/EQUINOX/CRVALn replaced, BITPIX -32. */
int gvmh_fits_write(const char* path, const float* data, int64_t naxis1, int64_t naxis2, const char* template_path,
                    const char* bunit, int niter, const char* radesys, float equinox, double crval1, double crval2);
/* getOptions on a command-line string; writes a JSON object of the parsed Vars. */
int gvmh_parse_args(const char* args, char* json_out, size_t cap);
/* linmin's bracketing + Brent search on a caller-supplied 1-D function (host logic test). */
typedef float (*gvmh_fn1d)(float x, void* user);
int gvmh_linmin_1d(gvmh_fn1d f, void* user, float* xmin, float* fmin, int* probes);
/* Multi-GPU sharding rule of MFS::setDevice: the part [lo[c], hi[c]) of channel c's Z[c]
 * visibilities that `rank` of `world` holds (lo == hi: none). */
int gvmh_shard_plan(int nchan, const int64_t* Z, int world, int rank, int64_t* lo, int64_t* hi);
/* readGVMS summary: out[8] = M, N, nchan, total visibilities, min_freq, max_freq, max_blength, uvmax_wavelength */
int gvmh_read_gvms(const char* path, double* out);

#ifdef __cplusplus
}
#endif
#endif /* GVM_HOST_H */
