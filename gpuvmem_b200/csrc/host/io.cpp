// io.cpp — the Io handler registered under both of the reference's keys, "IoMS" and "IoFITS":
// visibilities come from the GVMS container (casacore is not available offline); images and
// image headers are FITS (dependency-free reader/writer in fits.cpp) when the file is FITS /
// the output name ends in ".fits", raw little-endian fp32 + a JSON sidecar otherwise.
#include "io.hpp"

#include <cstdio>
#include <cstdlib>

#include "fits.hpp"
#include "globals.hpp"

namespace gpuvmem {

class IoGVMS : public Io {
 public:
  void read(const std::string& path, std::vector<MSAntenna>& antennas, std::vector<Field>& fields,
            MSData* data) override {
    MSDataset ds;
    headerValues h;
    std::string err;
    if (!readGVMS(path, &ds, &h, &err)) {
      std::printf("ERROR: %s\n", err.c_str());
      std::exit(-1);
    }
    antennas = std::move(ds.antennas);
    fields = std::move(ds.fields);
    *data = ds.data;
  }
  headerValues readHeader(const std::string& path) override {
    headerValues h;
    std::string err;
    if (isFitsFile(path)) {   // the -m model image: astrometry header (readFITSHeader, src/MSFITSIO.cu:262-325)
      FitsImage img;
      if (!fitsRead(path, false, &img, &err) || !fitsHeaderValues(img, &h, &err)) {
        std::printf("ERROR: %s\n", err.c_str());
        std::exit(-1);
      }
      template_cards = img.cards;   // copied into every image written later (OCopyFITS copies the header)
      return h;
    }
    MSDataset ds;
    // the container carries the FITS-header values in its first 100 bytes; reading it whole
    // keeps one code path (headers are read once per run)
    if (!readGVMS(path, &ds, &h, &err)) {
      std::printf("ERROR: %s\n", err.c_str());
      std::exit(-1);
    }
    return h;
  }
  void printImage(float* I_dev, const std::string& name, const std::string& units, int iteration, int index,
                  float fg_scale, long M_, long N_, bool) override {
    const size_t MN = (size_t)M_ * N_;
    std::vector<float> host(MN);
    devDownload(host.data(), I_dev + MN * index, MN);
    for (float& v : host) v *= fg_scale;  // OCopyFITS, src/MSFITSIO.cu:152-156
    const std::string file = name.find('/') == std::string::npos && name != output ? path + name : name;
    if (file.size() > 5 && file.compare(file.size() - 5, 5, ".fits") == 0) {   // OCopyFITS, src/MSFITSIO.cu:93-165
      std::string err;
      std::vector<std::string> cards = template_cards;
      if (cards.empty()) {   // no -m FITS template (library callers, GVMS headers): a minimal celestial WCS
        auto card = [&](const char* fmt, auto value) {
          char buf[96];
          std::snprintf(buf, sizeof(buf), fmt, value);
          std::string c(buf);
          c.resize(80, ' ');
          cards.push_back(c);
        };
        card("CTYPE1  = %-20s", "'RA---SIN'");
        card("CDELT1  = %20.13E", cdelt1);
        card("CRPIX1  = %20.6f", crpix1);
        card("CUNIT1  = %-20s", "'deg     '");
        card("CTYPE2  = %-20s", "'DEC--SIN'");
        card("CDELT2  = %20.13E", cdelt2);
        card("CRPIX2  = %20.6f", crpix2);
        card("CUNIT2  = %-20s", "'deg     '");
      }
      if (!fitsWriteFloat(file, host.data(), M_, N_, cards, units, iteration, frame, equinox, ra, dec, &err))
        std::printf("ERROR: %s\n", err.c_str());
      return;
    }
    std::FILE* fp = std::fopen(file.c_str(), "wb");
    if (!fp) {
      std::printf("ERROR: cannot write %s\n", file.c_str());
      return;
    }
    std::fwrite(host.data(), sizeof(float), MN, fp);
    std::fclose(fp);
    std::FILE* js = std::fopen((file + ".json").c_str(), "w");
    if (js) {
      std::fprintf(js,
                   "{\"dtype\": \"float32\", \"shape\": [%ld, %ld], \"bunit\": \"%s\", \"niter\": %d, "
                   "\"crval1\": %.12g, \"crval2\": %.12g, \"radesys\": \"%s\", \"equinox\": %g, \"scale\": %.9g}\n",
                   M_, N_, units.c_str(), iteration, ra, dec, frame.c_str(), equinox, fg_scale);
      std::fclose(js);
    }
  }
  std::vector<float> read_data_float_FITS(const std::string& file) override {
    if (isFitsFile(file)) {
      FitsImage f;
      std::string err;
      if (!fitsRead(file, true, &f, &err) || f.naxis1 != M || f.naxis2 != N) {
        std::printf("ERROR: %s\n", err.empty() ? (file + ": image size differs from the model header").c_str() : err.c_str());
        std::exit(-1);
      }
      return f.data;
    }
    std::vector<float> img((size_t)M * N);
    std::FILE* fp = std::fopen(file.c_str(), "rb");
    if (!fp || std::fread(img.data(), sizeof(float), img.size(), fp) != img.size()) {
      std::printf("ERROR: cannot read %ld x %ld fp32 values from %s\n", M, N, file.c_str());
      std::exit(-1);
    }
    std::fclose(fp);
    return img;
  }
  std::vector<std::string> template_cards;

  void writeModelVisibilities(const std::string& out, std::vector<Field>& fields, MSData& data) override {
    // residuals + model per block: int64 Z; float Vm[Z][2]; float Vr[Z][2]; float weight[Z]
    std::FILE* fp = std::fopen(out.c_str(), "wb");
    if (!fp) {
      std::printf("ERROR: cannot write %s\n", out.c_str());
      return;
    }
    std::fwrite("GVMR0001", 1, 8, fp);
    const int32_t dims[3] = {(int32_t)fields.size(), data.total_frequencies, data.nstokes};
    std::fwrite(dims, sizeof(int32_t), 3, fp);
    for (Field& f : fields)
      for (auto& chan : f.visibilities)
        for (HVis& v : chan) {
          const int64_t Z = (int64_t)v.size();
          std::fwrite(&Z, sizeof(Z), 1, fp);
          std::fwrite(v.Vm.data(), sizeof(float), 2 * Z, fp);
          std::fwrite(v.Vr.data(), sizeof(float), 2 * Z, fp);
          std::fwrite(v.weight.data(), sizeof(float), Z, fp);
        }
    std::fclose(fp);
  }
};

namespace {
Io* makeIo() { return new IoGVMS; }
const bool kRegistered[] = {
    registerCreationFunction<Io, std::string>("IoMS", makeIo),
    registerCreationFunction<Io, std::string>("IoFITS", makeIo),
};
}  // namespace

}  // namespace gpuvmem
