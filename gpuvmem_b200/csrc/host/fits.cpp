// fits.cpp — see fits.hpp. FITS standard 4.0: 2880-byte blocks, 80-character cards, big-endian data.
#include "fits.hpp"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace gpuvmem {

namespace {
constexpr size_t kBlock = 2880, kCard = 80;

std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(' '), b = s.find_last_not_of(' ');
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
std::string keyOf(const std::string& card) { return trim(card.substr(0, 8)); }

std::string makeCard(const std::string& key, const std::string& value, const std::string& comment) {
  char buf[200];
  std::snprintf(buf, sizeof(buf), "%-8s= %20s%s%s", key.c_str(), value.c_str(), comment.empty() ? "" : " / ",
                comment.c_str());
  std::string c(buf);
  c.resize(kCard, ' ');
  return c;
}
std::string stringCard(const std::string& key, const std::string& value, const std::string& comment) {
  std::string v = "'" + value;
  while (v.size() < 9) v += ' ';   // at least 8 characters between the quotes
  v += "'";
  char buf[200];
  std::snprintf(buf, sizeof(buf), "%-8s= %-20s%s%s", key.c_str(), v.c_str(), comment.empty() ? "" : " / ",
                comment.c_str());
  std::string c(buf);
  c.resize(kCard, ' ');
  return c;
}
std::string numCard(const std::string& key, double v, const std::string& comment) {
  char num[40];
  std::snprintf(num, sizeof(num), "%.17G", v);   // 17 significant digits: doubles round-trip
  if (!std::strpbrk(num, ".EN")) std::strcat(num, ".");   // keep it a floating-point literal
  return makeCard(key, num, comment);
}
std::string intCard(const std::string& key, long v, const std::string& comment) {
  return makeCard(key, std::to_string(v), comment);
}

template <class T>
T bigEndian(const unsigned char* p) {
  unsigned char b[sizeof(T)];
  for (size_t i = 0; i < sizeof(T); i++) b[i] = p[sizeof(T) - 1 - i];
  T v;
  std::memcpy(&v, b, sizeof(T));
  return v;
}
}  // namespace

bool isFitsFile(const std::string& path) {
  std::FILE* fp = std::fopen(path.c_str(), "rb");
  if (!fp) return false;
  char head[10] = {0};
  const size_t got = std::fread(head, 1, 9, fp);
  std::fclose(fp);
  return got == 9 && std::strncmp(head, "SIMPLE  =", 9) == 0;
}

bool fitsCard(const std::vector<std::string>& cards, const std::string& key, std::string* value) {
  for (const std::string& c : cards) {
    if (keyOf(c) != key || c.size() < 10 || c[8] != '=') continue;
    std::string v = c.substr(10);
    const size_t q = v.find('\'');
    if (q != std::string::npos && trim(v.substr(0, q)).empty()) {   // string value: up to the closing quote
      size_t e = q + 1;
      std::string out;
      while (e < v.size()) {
        if (v[e] == '\'') {
          if (e + 1 < v.size() && v[e + 1] == '\'') { out += '\''; e += 2; continue; }
          break;
        }
        out += v[e++];
      }
      size_t last = out.find_last_not_of(' ');
      *value = last == std::string::npos ? std::string() : out.substr(0, last + 1);
    } else {
      const size_t slash = v.find('/');
      *value = trim(slash == std::string::npos ? v : v.substr(0, slash));
    }
    return true;
  }
  return false;
}

bool fitsRead(const std::string& path, bool want_data, FitsImage* out, std::string* err) {
  std::FILE* fp = std::fopen(path.c_str(), "rb");
  if (!fp) { *err = "cannot open " + path; return false; }
  out->cards.clear();
  bool ended = false;
  std::vector<char> block(kBlock);
  while (!ended) {
    if (std::fread(block.data(), 1, kBlock, fp) != kBlock) { std::fclose(fp); *err = path + ": truncated FITS header"; return false; }
    for (size_t c = 0; c < kBlock / kCard; c++) {
      std::string card(block.data() + c * kCard, kCard);
      if (keyOf(card) == "END") { ended = true; break; }
      out->cards.push_back(card);
    }
  }
  std::string v;
  if (out->cards.empty() || keyOf(out->cards[0]) != "SIMPLE") { std::fclose(fp); *err = path + ": not a FITS file"; return false; }
  long naxis = 0;
  if (fitsCard(out->cards, "NAXIS", &v)) naxis = std::atol(v.c_str());
  if (fitsCard(out->cards, "BITPIX", &v)) out->bitpix = std::atoi(v.c_str());
  if (naxis < 2 || !fitsCard(out->cards, "NAXIS1", &v)) { std::fclose(fp); *err = path + ": primary HDU holds no image (NAXIS < 2)"; return false; }
  out->naxis1 = std::atol(v.c_str());
  fitsCard(out->cards, "NAXIS2", &v);
  out->naxis2 = std::atol(v.c_str());
  if (out->naxis1 <= 0 || out->naxis2 <= 0) { std::fclose(fp); *err = path + ": empty image"; return false; }
  if (!want_data) { std::fclose(fp); return true; }
  double bscale = 1.0, bzero = 0.0;
  if (fitsCard(out->cards, "BSCALE", &v)) bscale = std::atof(v.c_str());
  if (fitsCard(out->cards, "BZERO", &v)) bzero = std::atof(v.c_str());
  const int bytes = std::abs(out->bitpix) / 8;
  if (bytes != 1 && bytes != 2 && bytes != 4 && bytes != 8) { std::fclose(fp); *err = path + ": unsupported BITPIX"; return false; }
  const size_t n = (size_t)out->naxis1 * out->naxis2;
  std::vector<unsigned char> raw(n * bytes);
  if (std::fread(raw.data(), 1, raw.size(), fp) != raw.size()) { std::fclose(fp); *err = path + ": truncated FITS data"; return false; }
  std::fclose(fp);
  out->data.resize(n);
  for (size_t i = 0; i < n; i++) {
    const unsigned char* p = raw.data() + i * bytes;
    double x;
    switch (out->bitpix) {
      case 8: x = (double)p[0]; break;
      case 16: x = (double)bigEndian<int16_t>(p); break;
      case 32: x = (double)bigEndian<int32_t>(p); break;
      case -32: x = (double)bigEndian<float>(p); break;
      default: x = bigEndian<double>(p); break;   // -64
    }
    out->data[i] = (float)(bscale * x + bzero);
  }
  return true;
}

bool fitsHeaderValues(const FitsImage& img, headerValues* h, std::string* err) {
  std::string v;
  auto num = [&](const char* key, double* dst) {
    if (!fitsCard(img.cards, key, &v)) return false;
    for (char& ch : v) if (ch == 'D' || ch == 'd') ch = 'E';   // Fortran-style exponents
    *dst = std::atof(v.c_str());
    return true;
  };
  if (!num("CDELT1", &h->DELTAX) || !num("CDELT2", &h->DELTAY) || !num("CRVAL1", &h->ra) || !num("CRVAL2", &h->dec) ||
      !num("CRPIX1", &h->crpix1) || !num("CRPIX2", &h->crpix2)) {
    *err = "the FITS header lacks CDELT1/2, CRVAL1/2 or CRPIX1/2";
    return false;
  }
  h->M = img.naxis1;
  h->N = img.naxis2;
  h->bitpix = img.bitpix;
  num("BMAJ", &h->beam_bmaj);
  num("BMIN", &h->beam_bmin);
  num("BPA", &h->beam_bpa);
  double noise = -1.0;
  h->beam_noise = num("NOISE", &noise) ? (float)noise : -1.0f;      // src/MSFITSIO.cu:289-299
  h->radesys = fitsCard(img.cards, "RADESYS", &v) ? v : "ICRS";
  double eq = 2000.0;
  h->equinox = num("EQUINOX", &eq) ? (float)eq : 2000.0f;
  return true;
}

bool fitsWriteFloat(const std::string& path, const float* data, long naxis1, long naxis2,
                    const std::vector<std::string>& template_cards, const std::string& bunit, int niter,
                    const std::string& radesys, float equinox, double crval1, double crval2, std::string* err) {
  static const char* kStructural[] = {"SIMPLE", "BITPIX", "NAXIS", "NAXIS1", "NAXIS2", "NAXIS3", "NAXIS4", "EXTEND",
                                      "BSCALE", "BZERO", "BLANK", "BUNIT", "NITER", "RADESYS", "EQUINOX", "CRVAL1",
                                      "CRVAL2", "END", "DATAMAX", "DATAMIN"};
  std::vector<std::string> cards;
  cards.push_back(makeCard("SIMPLE", "T", "conforms to FITS standard"));
  cards.push_back(intCard("BITPIX", -32, "IEEE single precision"));
  cards.push_back(intCard("NAXIS", 2, ""));
  cards.push_back(intCard("NAXIS1", naxis1, ""));
  cards.push_back(intCard("NAXIS2", naxis2, ""));
  for (const std::string& c : template_cards) {
    const std::string k = keyOf(c);
    bool skip = false;
    for (const char* s : kStructural) skip = skip || k == s;
    // axes beyond the second (degenerate frequency / stokes axes of the template) are dropped with their WCS
    if (k.size() == 6 && (k.compare(0, 5, "CTYPE") == 0 || k.compare(0, 5, "CRVAL") == 0 || k.compare(0, 5, "CDELT") == 0 ||
                          k.compare(0, 5, "CRPIX") == 0 || k.compare(0, 5, "CUNIT") == 0 || k.compare(0, 5, "CROTA") == 0) &&
        k[5] > '2')
      skip = true;
    if (!skip) { std::string cc = c; cc.resize(kCard, ' '); cards.push_back(cc); }
  }
  cards.push_back(stringCard("BUNIT", bunit, "Unit of measurement"));
  cards.push_back(intCard("NITER", niter, "Number of iteration in gpuvmem software"));
  cards.push_back(stringCard("RADESYS", radesys, "Changed by gpuvmem"));
  cards.push_back(numCard("EQUINOX", equinox, "Changed by gpuvmem"));
  cards.push_back(numCard("CRVAL1", crval1, "Changed by gpuvmem"));
  cards.push_back(numCard("CRVAL2", crval2, "Changed by gpuvmem"));
  std::string end = "END";
  end.resize(kCard, ' ');
  cards.push_back(end);
  std::FILE* fp = std::fopen(path.c_str(), "wb");
  if (!fp) { *err = "cannot write " + path; return false; }
  size_t written = 0;
  for (const std::string& c : cards) written += std::fwrite(c.data(), 1, kCard, fp);
  const std::string blank(kCard, ' ');
  while (written % kBlock) written += std::fwrite(blank.data(), 1, kCard, fp);
  const size_t n = (size_t)naxis1 * naxis2;
  std::vector<unsigned char> raw(n * 4);
  for (size_t i = 0; i < n; i++) {
    unsigned char b[4];
    std::memcpy(b, &data[i], 4);
    raw[4 * i] = b[3]; raw[4 * i + 1] = b[2]; raw[4 * i + 2] = b[1]; raw[4 * i + 3] = b[0];
  }
  std::fwrite(raw.data(), 1, raw.size(), fp);
  const size_t pad = (kBlock - raw.size() % kBlock) % kBlock;
  const std::vector<unsigned char> zeros(pad, 0);
  if (pad) std::fwrite(zeros.data(), 1, pad, fp);
  std::fclose(fp);
  return true;
}

}  // namespace gpuvmem
