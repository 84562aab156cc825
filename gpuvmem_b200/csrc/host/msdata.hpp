// msdata.hpp — the in-memory visibility model the hot path consumes, laid out like the
// reference's include/MSFITSIO.cuh:58-150 (MSData / HVis / Field / MSAntenna / MSDataset /
// headerValues) with flat storage: uvw is a [Z][3] double array (the reference's
// std::vector<double3> has the same bytes), Vo a [Z][2] float array.
// casacore/cfitsio ingestion (src/MSFITSIO.cu) is replaced by the GVMS container that
// gpuvmem_b200/synth.py writes (readGVMS below): same fields, no third-party reader.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace gpuvmem {

enum { AIRYDISK = 0, GAUSSIAN = 1 };  // include/MSFITSIO.cuh:56
// correlation codes, include/functions.cuh:25-59 (casacore numbering)
enum { I_ST = 1, Q_ST, U_ST, V_ST, RR, RL, LR, LL, XX, XY, YX, YY };

const float LIGHTSPEED = 2.99792458E8f;  // include/MSFITSIO.cuh:54

// src/MSFITSIO.cu:36-45: fp32 wavelength, fp64 division
inline float freq_to_wavelength(float freq) { return LIGHTSPEED / freq; }
inline double metres_to_lambda(double uvw_metres, float freq) {
  const float lambda = freq_to_wavelength(freq);
  return uvw_metres / lambda;
}

struct MSData {
  int total_frequencies = 0, nfields = 0, nsamples = 0, nstokes = 0, nantennas = 0, nbaselines = 0;
  float ref_freq = 0, min_freq = 0, max_freq = 0, max_blength = 0, min_blength = 0;
  double uvmax_wavelength = 0;
  std::string telescope_name;
  std::vector<int> corr_type;
  int max_number_visibilities_in_channel_and_stokes = 0;
  int max_number_vis = 0;
};

struct HVis {
  std::vector<double> uvw;    // [Z][3] metres
  std::vector<float> weight;  // [Z]
  std::vector<float> Vo;      // [Z][2]
  std::vector<float> Vm;      // [Z][2]
  std::vector<float> Vr;      // [Z][2]
  size_t size() const { return weight.size(); }
};

struct Field {
  int id = 0, valid_frequencies = 0;
  double ref_ra = 0, ref_dec = 0, phs_ra = 0, phs_dec = 0;
  float ref_xobs_pix = 0, ref_yobs_pix = 0, phs_xobs_pix = 0, phs_yobs_pix = 0;
  std::vector<float> nu;  // channel frequencies, float as in the reference
  std::vector<std::vector<long>> numVisibilitiesPerFreqPerStoke;
  std::vector<long> numVisibilitiesPerFreq;
  std::vector<std::vector<HVis>> visibilities;         // [chan][stokes]
  std::vector<std::vector<HVis>> backup_visibilities;  // weights before the scheme
  std::vector<std::vector<int>> engine_slot;           // [chan][stokes] -> gvm slot on THIS rank (-1: elsewhere)
};

struct MSAntenna {
  std::string antenna_id, station;
  float antenna_diameter = 0, pb_factor = 0, pb_cutoff = 0;
  int primary_beam = AIRYDISK;
};

struct MSDataset {
  std::string name, oname;
  std::vector<Field> fields;
  std::vector<MSAntenna> antennas;
  MSData data;
};

struct headerValues {
  double DELTAX = 0, DELTAY = 0, ra = 0, dec = 0, crpix1 = 0, crpix2 = 0;
  long M = 0, N = 0;
  double beam_bmaj = 0, beam_bmin = 0, beam_bpa = 0;
  float beam_noise = -1.0f;
  std::string radesys = "ICRS";
  float equinox = 2000.0f;
  int bitpix = -32;
};

// GVMS container (little-endian), written by gpuvmem_b200/synth.py:write_gvms:
//   char magic[8] = "GVMS0001"; int64 M, N; double DELTAX, DELTAY, ra, dec (deg), crpix1, crpix2;
//   float beam_noise; float antenna_diameter; char telescope[32];
//   int32 nfields, nchan, nstokes; int32 corr_type[nstokes];
//   per field: double ref_ra, ref_dec, phs_ra, phs_dec (radians); float nu[nchan];
//     per chan, per stokes: int64 Z; double uvw[Z][3]; float Vo[Z][2]; float weight[Z]
// Fills the dataset the way readMS does (src/MSFITSIO.cu:398-754), including the
// per-telescope beam model (:510-551) and the MSData summary values.
bool readGVMS(const std::string& path, MSDataset* ds, headerValues* header, std::string* err);
// The same from memory-resident arrays (tests, bench, Python callers): one field, one stokes (XX).
void fillDataset(MSDataset* ds, const std::string& telescope, float antenna_diameter, double ra_rad,
                 double dec_rad, int nchan, const float* nu, const int64_t* Z, const double* const* uvw_m,
                 const float* const* Vo, const float* const* w);
// telescope -> (pb_factor, pb_cutoff, primary_beam), src/MSFITSIO.cu:510-551
void beamModel(const std::string& telescope, float antenna_diameter, float min_freq, MSAntenna* out);
void finishDataset(MSDataset* ds, float antenna_diameter);  // MSData summary + antenna beam model

}  // namespace gpuvmem
