// fits.hpp — a dependency-free reader/writer for the FITS images the hot path's callers exchange
// (cfitsio is not available offline). Covers what the reference does with cfitsio in
// src/MSFITSIO.cu:65-360: read the astrometry header of the -m model image (readFITSHeader,
// :262-325: CDELT1/2, CRVAL1/2, CRPIX1/2, NAXIS1/2, BMAJ/BMIN/BPA, NOISE, RADESYS, EQUINOX, BITPIX),
// read an image plane as floats (open_fits / read_data_float_FITS), and write a result image with
// the model's header copied and BUNIT/NITER/NAXISn/RADESYS/EQUINOX/CRVALn updated (OCopyFITS, :93-165).
// Primary HDU only, NAXIS >= 2 (degenerate extra axes allowed), BITPIX 8/16/32/-32/-64 with
// BSCALE/BZERO on input; -32 on output.
#pragma once
#include <string>
#include <vector>

#include "msdata.hpp"

namespace gpuvmem {

struct FitsImage {
  std::vector<std::string> cards;   // 80-character header cards of the primary HDU, END excluded
  long naxis1 = 0, naxis2 = 0;
  int bitpix = -32;
  std::vector<float> data;          // [naxis2][naxis1], first plane
};

bool isFitsFile(const std::string& path);
// header only when want_data is false
bool fitsRead(const std::string& path, bool want_data, FitsImage* out, std::string* err);
// value of a header card as text (quotes and comment stripped); false when absent
bool fitsCard(const std::vector<std::string>& cards, const std::string& key, std::string* value);
// the reference's headerValues from the cards (readFITSHeader, src/MSFITSIO.cu:262-325)
bool fitsHeaderValues(const FitsImage& img, headerValues* h, std::string* err);
// `template_cards`: header to copy (may be empty); the keys OCopyFITS updates are replaced
bool fitsWriteFloat(const std::string& path, const float* data, long naxis1, long naxis2,
                    const std::vector<std::string>& template_cards, const std::string& bunit, int niter,
                    const std::string& radesys, float equinox, double crval1, double crval2, std::string* err);

}  // namespace gpuvmem
