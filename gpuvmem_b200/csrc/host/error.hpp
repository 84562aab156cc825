// error.hpp — error-image estimators (reference include/classes/error.cuh:7-10 and
// include/secondderivateerror.cuh:6-10; factory key "SecondDerivateError",
// src/secondderivateerror.cu:12-18). MFS::writeImages runs one under -E (src/mfs.cu:1090-1113).
#pragma once
#include "factory.hpp"
#include "image.hpp"

namespace gpuvmem {

class Visibilities;   // visibilities.hpp; the reference passes its (unused) Visibilities object through

class Error {
 public:
  virtual ~Error() = default;
  // fills I->getErrorImage(): [image_count][M][N] on the device, allocated here like the reference does
  virtual void calculateErrorImage(Image* I, Visibilities* v) = 0;
};

// calculateErrors (src/functions.cu:4966-5040) -> gvm_error_maps
class SecondDerivateError : public Error {
 public:
  void calculateErrorImage(Image* I, Visibilities* v) override;
};

}  // namespace gpuvmem
