// uvtaper.hpp — elliptical Gaussian taper in the uv plane. Surface of the reference's
// include/classes/uvtaper.cuh (constructors, FWHM setters, getValue :100-117). The
// per-visibility evaluation itself runs on the GPU inside gvm_weights (gvm_taper).
#pragma once
#include <cmath>

#include "../../../include/gvm_b200.h"

namespace gpuvmem {

class UVTaper {
 public:
  UVTaper() : UVTaper(1.0f, 1.0f, 0.0f) {}
  explicit UVTaper(float sigma) : UVTaper(sigma, sigma, 0.0f) {}
  UVTaper(float sigma_maj, float sigma_min, float bpa, float amplitude = 1.0f, double u_0 = 0.0, double v_0 = 0.0)
      : sigma_maj(sigma_maj), sigma_min(sigma_min), bpa(bpa), amplitude(amplitude), u_0(u_0), v_0(v_0) {}

  float getSigma_maj() const { return sigma_maj; }
  float getSigma_min() const { return sigma_min; }
  float getBPA() const { return bpa; }
  float getAmplitude() const { return amplitude; }
  void setSigma_maj(float s) { sigma_maj = s; }
  void setSigma_min(float s) { sigma_min = s; }
  void setSigmas(float maj, float min) { sigma_maj = maj; sigma_min = min; }
  void setAmplitude(float a) { amplitude = a; }
  void setBPA(float b) { bpa = b; }
  void setCenter(double u, double v) { u_0 = u; v_0 = v; }
  void setGaussianParameters(float maj, float min, float b) { sigma_maj = maj; sigma_min = min; bpa = b; }
  static float fwhmToSigma(float fwhm) { return fwhm / (2.0f * sqrtf(2.0f * logf(2.0f))); }
  void setFWHM(float fwhm_maj, float fwhm_min) { sigma_maj = fwhmToSigma(fwhm_maj); sigma_min = fwhmToSigma(fwhm_min); }
  void setFWHM_arcsec(float maj_arcsec, float min_arcsec) {
    const float pi = 3.14159265358979323846f;
    const float maj_rad = (maj_arcsec / 3600.0f) * pi / 180.0, min_rad = (min_arcsec / 3600.0f) * pi / 180.0;
    sigma_maj = fwhmToSigma(1.0f / maj_rad);
    sigma_min = fwhmToSigma(1.0f / min_rad);
  }
  void setFWHM_arcsec(float arcsec) { setFWHM_arcsec(arcsec, arcsec); }

  // host evaluation, same expression order as uvtaper.cuh:100-117 (used by tests)
  float getValue(double u, double v) const {
    const double x = u - u_0, y = v - v_0;
    const float c = cosf(bpa), s = sinf(bpa), s2 = sinf(2.0f * bpa);
    const float a = (c * c) / (2.0f * sigma_maj * sigma_maj) + (s * s) / (2.0f * sigma_min * sigma_min);
    const float b = s2 / (2.0f * sigma_maj * sigma_maj) - s2 / (2.0f * sigma_min * sigma_min);
    const float cc = (s * s) / (2.0f * sigma_maj * sigma_maj) + (c * c) / (2.0f * sigma_min * sigma_min);
    return amplitude * exp(-a * x * x - b * x * y - cc * y * y);
  }
  gvm_taper abi() const {
    gvm_taper t;
    t.enabled = 1; t.sigma_maj = sigma_maj; t.sigma_min = sigma_min; t.bpa = bpa; t.amplitude = amplitude;
    t.u_0 = u_0; t.v_0 = v_0;
    return t;
  }

 private:
  float sigma_maj, sigma_min, bpa, amplitude;
  double u_0, v_0;
};

}  // namespace gpuvmem
