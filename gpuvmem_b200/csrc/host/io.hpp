// io.hpp — the Io surface the hot path touches (reference include/classes/io.cuh):
// reading visibilities + the image header, writing result images. MS/FITS are out of scope
// (casacore/cfitsio absent); the registered handlers "IoMS"/"IoFITS" read the GVMS container
// and write raw little-endian fp32 images with a one-line JSON sidecar.
#pragma once
#include <string>
#include <vector>

#include "factory.hpp"
#include "msdata.hpp"

namespace gpuvmem {

class Io {
 public:
  virtual ~Io() = default;
  // visibilities
  virtual void read(const std::string& path, std::vector<MSAntenna>& antennas, std::vector<Field>& fields,
                    MSData* data) = 0;
  virtual headerValues readHeader(const std::string& path) = 0;
  // images: device pointer to [image_count][M][N]; index selects one image
  virtual void printImage(float* I_dev, const std::string& name, const std::string& units, int iteration,
                          int index, float fg_scale, long M, long N, bool normalize) = 0;
  virtual void printNotNormalizedImage(float* I_dev, const std::string& name, const std::string& units,
                                       int iteration, int index, bool) {
    printImage(I_dev, name, units, iteration, index, 1.0f, M, N, false);
  }
  virtual void printImageIteration(float* I_dev, const std::string& name, const std::string& units, int iteration,
                                   int index, bool) {
    printImage(I_dev, name + "_" + std::to_string(iteration), units, iteration, index, 1.0f, M, N, false);
  }
  virtual void writeModelVisibilities(const std::string& path, std::vector<Field>& fields, MSData& data) = 0;
  // an image-sized float plane from a file (reference: IoFITS::read_data_float_FITS, src/iofits.cu:98-129;
  // used for the -U user mask): raw little-endian fp32, M*N values
  virtual std::vector<float> read_data_float_FITS(const std::string& file) = 0;

  void setInput(const std::string& s) { input = s; }
  void setOutput(const std::string& s) { output = s; }
  void setPath(const std::string& s) { path = s; }
  void setMN(long m, long n) { M = m; N = n; }
  void setRADec(double r, double d) { ra = r; dec = d; }
  // pixel scale (degrees) and reference pixel, for image headers written without a template
  void setPixelGrid(double dx, double dy, double cp1, double cp2) { cdelt1 = dx; cdelt2 = dy; crpix1 = cp1; crpix2 = cp2; }
  void setFrame(const std::string& s) { frame = s; }
  void setEquinox(float e) { equinox = e; }
  void setPrintImages(bool p) { print_images = p; }
  bool getPrintImages() const { return print_images; }
  void setGridding(int g) { gridding = g; }
  void setRandomProbability(float r) { random_probability = r; }
  void setApplyNoiseInput(bool b) { apply_noise = b; }
  void setStoreModelVisInput(bool b) { store_model = b; }
  const std::string& getOutput() const { return output; }
  const std::string& getPath() const { return path; }

 protected:
  std::string input, output, path = "mem/", frame = "ICRS";
  long M = 0, N = 0;
  double ra = 0, dec = 0, cdelt1 = 0, cdelt2 = 0, crpix1 = 0, crpix2 = 0;
  float equinox = 2000.0f, random_probability = 1.0f;
  bool print_images = false, apply_noise = false, store_model = false;
  int gridding = 0;
};

}  // namespace gpuvmem
