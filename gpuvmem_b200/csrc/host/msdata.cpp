#include "msdata.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace gpuvmem {

void beamModel(const std::string& telescope, float antenna_diameter, float min_freq, MSAntenna* a) {
  const float max_wavelength = freq_to_wavelength(min_freq);
  // boost::math::cyl_bessel_j_zero(1.0f, 1) / PI of the reference (src/MSFITSIO.cu:522-523)
  const float first_j1_zero = 3.8317059702075125f;
  const float pb_default = first_j1_zero / 3.14159265358979323846f;
  if (telescope == "ALMA") { a->pb_factor = 1.13f; a->primary_beam = AIRYDISK; }
  else if (telescope == "EVLA") { a->pb_factor = 1.25f; a->primary_beam = GAUSSIAN; }
  else { a->pb_factor = pb_default; a->primary_beam = GAUSSIAN; }
  a->antenna_diameter = antenna_diameter;
  a->pb_cutoff = a->pb_factor * (max_wavelength / antenna_diameter);
}

void finishDataset(MSDataset* ds, float antenna_diameter) {
  MSData& d = ds->data;
  d.nfields = (int)ds->fields.size();
  float fmin = 0, fmax = 0;
  bool first = true;
  double bmax = 0, bmin = 1e300, uvmax_m = 0;
  int maxvis = 0;
  for (Field& f : ds->fields) {
    d.total_frequencies = (int)f.nu.size();
    f.numVisibilitiesPerFreq.assign(f.nu.size(), 0);
    f.numVisibilitiesPerFreqPerStoke.assign(f.nu.size(), std::vector<long>(d.nstokes, 0));
    f.valid_frequencies = 0;
    for (size_t i = 0; i < f.nu.size(); i++) {
      if (first) { fmin = fmax = f.nu[i]; first = false; }
      fmin = std::min(fmin, f.nu[i]);
      fmax = std::max(fmax, f.nu[i]);
      for (int s = 0; s < d.nstokes; s++) {
        const HVis& v = f.visibilities[i][s];
        const long Z = (long)v.size();
        f.numVisibilitiesPerFreqPerStoke[i][s] = Z;
        f.numVisibilitiesPerFreq[i] += Z;
        maxvis = std::max<long>(maxvis, Z);
        for (long k = 0; k < Z; k++) {
          const double u = v.uvw[3 * k], vv = v.uvw[3 * k + 1];
          const double b = std::sqrt(u * u + vv * vv);
          bmax = std::max(bmax, b);
          bmin = std::min(bmin, b);
          uvmax_m = std::max(uvmax_m, std::max(std::fabs(u), std::fabs(vv)));
        }
      }
      if (f.numVisibilitiesPerFreq[i] > 0) f.valid_frequencies++;
    }
  }
  d.min_freq = fmin;
  d.max_freq = fmax;
  d.ref_freq = 0.5f * (fmin + fmax);
  d.max_blength = (float)bmax;
  d.min_blength = (float)(bmin == 1e300 ? 0.0 : bmin);
  d.uvmax_wavelength = uvmax_m * d.max_freq / LIGHTSPEED;  // src/MSFITSIO.cu:508
  d.max_number_visibilities_in_channel_and_stokes = maxvis;
  d.max_number_vis = maxvis;
  ds->antennas.assign(1, MSAntenna());
  beamModel(d.telescope_name, antenna_diameter, d.min_freq, &ds->antennas[0]);
  d.nantennas = 1;
}

void fillDataset(MSDataset* ds, const std::string& telescope, float antenna_diameter, double ra_rad,
                 double dec_rad, int nchan, const float* nu, const int64_t* Z, const double* const* uvw_m,
                 const float* const* Vo, const float* const* w) {
  ds->fields.assign(1, Field());
  ds->data = MSData();
  ds->data.telescope_name = telescope;
  ds->data.nstokes = 1;
  ds->data.corr_type.assign(1, XX);
  Field& f = ds->fields[0];
  f.ref_ra = f.phs_ra = ra_rad;
  f.ref_dec = f.phs_dec = dec_rad;
  f.nu.assign(nu, nu + nchan);
  f.visibilities.assign(nchan, std::vector<HVis>(1));
  for (int i = 0; i < nchan; i++) {
    HVis& v = f.visibilities[i][0];
    const size_t z = (size_t)Z[i];
    v.uvw.assign(uvw_m[i], uvw_m[i] + 3 * z);
    v.Vo.assign(Vo[i], Vo[i] + 2 * z);
    v.weight.assign(w[i], w[i] + z);
    v.Vm.assign(2 * z, 0.0f);
    v.Vr.assign(2 * z, 0.0f);
  }
  finishDataset(ds, antenna_diameter);
}

namespace {
template <class T>
bool rd(std::FILE* fp, T* out, size_t count = 1) {
  return std::fread(out, sizeof(T), count, fp) == count;
}
}  // namespace

bool readGVMS(const std::string& path, MSDataset* ds, headerValues* h, std::string* err) {
  std::FILE* fp = std::fopen(path.c_str(), "rb");
  if (!fp) { if (err) *err = "cannot open " + path; return false; }
  auto fail = [&](const char* why) { if (err) *err = path + ": " + why; std::fclose(fp); return false; };
  char magic[8];
  if (!rd(fp, magic, 8) || std::memcmp(magic, "GVMS0001", 8) != 0) return fail("not a GVMS0001 container");
  int64_t M, N;
  double hd[6];
  float noise, dish;
  char tel[32];
  int32_t nf, nc, ns;
  if (!rd(fp, &M) || !rd(fp, &N) || !rd(fp, hd, 6) || !rd(fp, &noise) || !rd(fp, &dish) || !rd(fp, tel, 32) ||
      !rd(fp, &nf) || !rd(fp, &nc) || !rd(fp, &ns))
    return fail("truncated header");
  if (nf < 1 || nc < 1 || ns < 1 || ns > 16) return fail("bad counts");
  tel[31] = 0;
  h->M = M; h->N = N; h->DELTAX = hd[0]; h->DELTAY = hd[1]; h->ra = hd[2]; h->dec = hd[3];
  h->crpix1 = hd[4]; h->crpix2 = hd[5]; h->beam_noise = noise;
  ds->data = MSData();
  ds->data.telescope_name = tel;
  ds->data.nstokes = ns;
  std::vector<int32_t> corr(ns);
  if (!rd(fp, corr.data(), ns)) return fail("truncated correlation types");
  ds->data.corr_type.assign(corr.begin(), corr.end());
  ds->fields.assign(nf, Field());
  for (int f = 0; f < nf; f++) {
    Field& F = ds->fields[f];
    F.id = f;
    double dir[4];
    if (!rd(fp, dir, 4)) return fail("truncated field");
    F.ref_ra = dir[0]; F.ref_dec = dir[1]; F.phs_ra = dir[2]; F.phs_dec = dir[3];
    F.nu.resize(nc);
    if (!rd(fp, F.nu.data(), nc)) return fail("truncated frequencies");
    F.visibilities.assign(nc, std::vector<HVis>(ns));
    for (int i = 0; i < nc; i++)
      for (int s = 0; s < ns; s++) {
        int64_t Z;
        if (!rd(fp, &Z) || Z < 0) return fail("truncated block");
        HVis& v = F.visibilities[i][s];
        v.uvw.resize(3 * Z); v.Vo.resize(2 * Z); v.weight.resize(Z);
        if (Z && (!rd(fp, v.uvw.data(), 3 * Z) || !rd(fp, v.Vo.data(), 2 * Z) || !rd(fp, v.weight.data(), Z)))
          return fail("truncated visibilities");
        v.Vm.assign(2 * Z, 0.0f);
        v.Vr.assign(2 * Z, 0.0f);
      }
  }
  std::fclose(fp);
  ds->name = path;
  finishDataset(ds, dish);
  return true;
}

}  // namespace gpuvmem
