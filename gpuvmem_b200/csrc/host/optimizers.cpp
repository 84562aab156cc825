// optimizers.cpp — ConjugateGradient (src/frprmn.cu:85-195) and LBFGS (src/lbfgs.cu:99-343)
// outer loops on top of LineSearch and the fused gvm_vec_* kernels.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <iostream>

#include "optimizer.hpp"

namespace gpuvmem {

namespace {
const double kEps = 1.0e-10;  // EPS of src/frprmn.cu:48 / src/lbfgs.cu:49
double nowSeconds() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
bool chatty() { return G().verbose_flag && !G().quiet; }
void say(const char* msg) {
  if (!G().quiet) std::cout << msg << std::endl;
}
}  // namespace

// ------------------------------------------------------ ConjugateGradient --
void ConjugateGradient::allocateMemoryGpu() {
  const size_t n = (size_t)G().M * G().N * image->getImageCount();
  device_g = devAllocFloats(n);
  device_h = devAllocFloats(n);
  xi = devAllocFloats(n);
}
void ConjugateGradient::deallocateMemoryGpu() {
  devFree(device_g); devFree(device_h); devFree(xi);
  device_g = device_h = xi = nullptr;
}

void ConjugateGradient::optimize() {
  Globals& g = G();
  if (!g.quiet) std::printf("\n\nStarting Fletcher Reeves Polak Ribiere method (Conj. Grad.)\n\n");
  g.flag_opt = flag;
  allocateMemoryGpu();
  if (configured) {
    of->configure(g.N, g.M, image->getImageCount());
    configured = 0;
  }
  LineSearch search(of, image);
  const int images = image->getImageCount();
  history.clear();
  exit_reason = "iterations";
  current_iteration = 0;

  auto finish = [&](const char* why, const char* msg) {
    exit_reason = why;
    say(msg);
    of->calcFunction(image->getImage());
    deallocateMemoryGpu();
  };

  fp = of->calcFunction(image->getImage());
  history.push_back(fp);
  if (chatty()) std::printf("Starting function value = %.4f\n", fp);
  of->calcGradient(image->getImage(), xi, 0);
  // g = -xi ; xi = h = g
  GVM_CHECK(gvm_vec_new_xi(g.engine, device_g, xi, device_h, 0.0f, images));

  for (int i = 1; i <= total_iterations; i++) {
    const double t0 = nowSeconds();
    current_iteration = i;
    if (chatty()) std::printf("\n\n********** Iteration %d **********\n\n", i);
    search.linmin(image->getImage(), xi, &fret);
    if (2.0f * fabsf(fret - fp) <= ftol * (fabsf(fret) + fabsf(fp) + kEps))
      return finish("tolerance", "Exit due to tolerance");

    fp = of->calcFunction(image->getImage());
    history.push_back(fp);
    if (chatty()) std::printf("Function value = %.4f\n", fp);
    of->calcGradient(image->getImage(), xi, i);

    const float den = std::max(fp, 1.0f);
    float gmax = 0.0f;
    GVM_CHECK(gvm_vec_grad_condition(g.engine, xi, image->getImage(), den, images, &gmax));
    if (gmax < gtol) return finish("gradient tolerance", "Exit due to gradient tolerance");

    // gg = g.g ; dgg = (xi + g).xi  (Polak-Ribiere)
    GVM_CHECK(gvm_vec_gg_dgg(g.engine, xi, device_g, images, &gg, &dgg));
    if (gg == 0.0f) return finish("gg = 0", "Exit due to gg = 0");
    gam = std::max(0.0f, dgg / gg);
    // g = -xi ; xi = h = g + gam*h
    GVM_CHECK(gvm_vec_new_xi(g.engine, device_g, xi, device_h, gam, images));
    if (chatty()) std::printf("Time: %.4f seconds\n", nowSeconds() - t0);
  }
  finish("iterations", "Too many iterations in frprmn");
}

// ------------------------------------------------------------------ LBFGS --
void LBFGS::allocateMemoryGpu() {
  const size_t n = (size_t)G().M * G().N * image->getImageCount();
  d_y = devAllocFloats(n * K);
  d_s = devAllocFloats(n * K);
  p_old = devAllocFloats(n);
  xi = devAllocFloats(n);
  xi_old = devAllocFloats(n);
  d_q = devAllocFloats(n);
  d_r = devAllocFloats(n);
}
void LBFGS::deallocateMemoryGpu() {
  devFree(d_y); devFree(d_s); devFree(xi); devFree(xi_old); devFree(p_old); devFree(d_q); devFree(d_r);
  d_y = d_s = xi = xi_old = p_old = d_q = d_r = nullptr;
}

void LBFGS::optimize() {
  Globals& g = G();
  say("\n\nStarting Lbfgs\n");
  g.flag_opt = flag;
  allocateMemoryGpu();
  if (configured) {
    of->configure(g.N, g.M, image->getImageCount());
    configured = 0;
  }
  LineSearch search(of, image);
  const int images = image->getImageCount();
  const size_t MN = (size_t)g.M * g.N;
  const size_t n = MN * images;
  history.clear();
  exit_reason = "iterations";
  current_iteration = 0;

  auto finish = [&](const char* why, const char* msg) {
    exit_reason = why;
    say(msg);
    of->calcFunction(image->getImage());
    deallocateMemoryGpu();
  };

  fp = of->calcFunction(image->getImage());
  history.push_back(fp);
  if (chatty()) std::printf("Starting function value = %.4f\n", fp);
  of->calcGradient(image->getImage(), xi, 0);
  GVM_CHECK(gvm_vec_scale(g.engine, xi, -1.0f, (int64_t)n));  // searchDirection_LBFGS

  for (int i = 1; i <= total_iterations; i++) {
    const double t0 = nowSeconds();
    current_iteration = i;
    max_per_it = 0.0f;
    if (chatty()) std::printf("\n\n********** Iteration %d **********\n\n", i);
    devCopyD2D(p_old, image->getImage(), n);
    devCopyD2D(xi_old, xi, n);

    search.linmin(image->getImage(), xi, &fret);
    if ((fp - fret) / std::max({fabsf(fret), fabsf(fp), 1.0f}) <= ftol)
      return finish("tolerance", "Exit due to tolerance");

    float step_max = 0.0f;  // normArray + deviceMaxReduce over the scaled step xi
    GVM_CHECK(gvm_vec_absmax(g.engine, xi, (int64_t)n, &step_max));
    max_per_it = std::max(max_per_it, step_max);
    if (max_per_it <= gtol) return finish("gradient tolerance", "Exit due to gnorm ~ 0");

    fp = of->calcFunction(image->getImage());
    history.push_back(fp);
    if (chatty()) std::printf("Function value = %.4f\n", fp);
    of->calcGradient(image->getImage(), xi, i);

    // history slot of this iteration; the offsets keep the reference's indexing
    // (MN*image*k + MN*image, src/functions.cu:3636-3653) — see LBFGS_recursion
    const int slot = (current_iteration - 1) % K;
    for (int im = 0; im < images; im++) {
      const size_t hist = MN * im * slot + MN * im, cur = MN * im;
      GVM_CHECK(gvm_vec_lbfgs_sy(g.engine, d_y + hist, d_s + hist, xi + cur, xi_old + cur,
                                 image->getImage() + cur, p_old + cur, (int64_t)MN));
    }
    LBFGS_recursion(d_y, d_s, xi, std::min(K, current_iteration), slot, (int)g.M, (int)g.N);
    if (chatty()) std::printf("Time: %.4f seconds\n", nowSeconds() - t0);
  }
  finish("iterations", "Too many iterations in LBFGS");
}

// Two-loop recursion, per image. Two different address rules are in play in the reference and
// both are kept: the dot products and calculateSandY address slot k of image
// `im` at MN*im*k + MN*im (getDot_LBFGS_ff, src/functions.cu:3575-3588), the updates at
// MN*im + MN*k (updateQ, :3613-3623). For image 1 they coincide; for image 0 the dots always
// see slot 0 while the updates walk through the slots.
void LBFGS::LBFGS_recursion(float* d_y, float* d_s, float* xi, int par_M, int lbfgs_it, int M, int N) {
  Globals& g = G();
  const int images = image->getImageCount();
  const size_t MN = (size_t)M * N;
  auto dotAt = [&](int im, int k) { return MN * im * k + MN * im; };
  auto updAt = [&](int im, int k) { return MN * im + MN * k; };
  auto dot = [&](const float* a, const float* b) {
    float v = 0.0f;
    GVM_CHECK(gvm_vec_dot(g.engine, a, b, (int64_t)MN, &v));
    return v;
  };
  std::vector<std::vector<float>> alpha(images, std::vector<float>(par_M, 0.0f));

  devZero(d_r, MN * images);
  devCopyD2D(d_q, xi, MN * images);

  for (int im = 0; im < images; im++) {
    for (int k = par_M - 1; k >= 0; k--) {
      const float rho_den = dot(d_y + dotAt(im, k), d_s + dotAt(im, k));
      const float rho = rho_den != 0.0f ? (float)(1.0 / rho_den) : 0.0f;
      alpha[im][k] = rho * dot(d_s + dotAt(im, k), d_q + dotAt(im, 0));
      // q -= alpha_k * y_k
      GVM_CHECK(gvm_vec_axpby(g.engine, -alpha[im][k], d_y + updAt(im, k), 1.0f, d_q + MN * im, (int64_t)MN));
    }
  }

  // initial Hessian scale (s.y)/(y.y) of the newest pair, summed over the images
  float sy_yy = 0.0f;
  for (int im = 0; im < images; im++) {
    const float sy = dot(d_y + dotAt(im, lbfgs_it), d_s + dotAt(im, lbfgs_it));
    const float yy = dot(d_y + dotAt(im, lbfgs_it), d_y + dotAt(im, lbfgs_it));
    if (yy != 0.0f) sy_yy += sy / yy;
  }

  for (int im = 0; im < images; im++) {
    // r = q * sy_yy
    GVM_CHECK(gvm_vec_axpby(g.engine, sy_yy, d_q + MN * im, 0.0f, d_r + MN * im, (int64_t)MN));
    for (int k = 0; k < par_M; k++) {
      const float rho_den = dot(d_y + dotAt(im, k), d_s + dotAt(im, k));
      const float rho = rho_den != 0.0f ? 1.0f / rho_den : 0.0f;
      const float beta = rho * dot(d_y + dotAt(im, k), d_r + dotAt(im, 0));
      // r += s_k * (alpha_k - beta)
      GVM_CHECK(gvm_vec_axpby(g.engine, alpha[im][k] - beta, d_s + updAt(im, k), 1.0f, d_r + MN * im, (int64_t)MN));
    }
  }
  GVM_CHECK(gvm_vec_scale(g.engine, d_r, -1.0f, (int64_t)(MN * images)));
  devCopyD2D(xi, d_r, MN * images);
}

namespace {
Optimizer* makeCG() { return new ConjugateGradient; }
Optimizer* makeLBFGS() { return new LBFGS; }
const bool kRegistered[] = {
    registerCreationFunction<Optimizer, std::string>("CG-FRPRMN", makeCG),
    registerCreationFunction<Optimizer, std::string>("CG-LBFGS", makeLBFGS),
};
}  // namespace

}  // namespace gpuvmem
