// image.hpp — the optimisation variable: [image_count][M][N] fp32 on the GPU plus the
// per-image update rules. Surface of the reference's include/classes/image.cuh:4-38.
#pragma once

namespace gpuvmem {

// per-image projection rules (reference: function pointers chosen in src/mfs.cu:798-811;
// image 0 is kept >= -eta*MINPIX unless -x, the others are unconstrained)
typedef struct functionMap {
  void (*newP)(float*, float*, float, int);
  void (*evaluateXt)(float*, float*, float*, float, int);
} imageMap;

// src/functions.cu:4566-4596
void defaultNewP(float* p, float* xi, float xmin, int image);
void particularNewP(float* p, float* xi, float xmin, int image);
void defaultEvaluateXt(float* xt, float* pcom, float* xicom, float x, int image);
void particularEvaluateXt(float* xt, float* pcom, float* xicom, float x, int image);

class Image {
 public:
  Image(float* image, int image_count) : image_count(image_count), image(image) {}
  int getImageCount() const { return image_count; }
  float* getImage() { return image; }
  float* getErrorImage() { return error_image; }
  imageMap* getFunctionMapping() { return functionMapping; }
  void setImageCount(int i) { image_count = i; }
  void setErrorImage(float* f) { error_image = f; }
  void setImage(float* i) { image = i; }
  void setFunctionMapping(imageMap* f) { functionMapping = f; }
  // true when the mapping is the stock one (image 0 projected, others free / all free):
  // the optimizers then use ONE fused kernel for all images instead of a launch per image
  bool stockMapping(bool nopositivity) const;

 private:
  int image_count;
  float* image;
  float* error_image = nullptr;
  imageMap* functionMapping = nullptr;
};

}  // namespace gpuvmem
