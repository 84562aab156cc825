// Diagnostic: accuracy of the MUFU sin/cos path used by the tensor-core gradient kernel.
#include <cstdio>
#include <cmath>
#include <cstdint>
#include <vector>
__device__ __forceinline__ float phase_to_angle(uint32_t ph) {
  const float f = __uint_as_float(0x3F800000u | (ph >> 9));
  return fmaf(f, 6.283185307179586f, -9.42477796076938f);
}
__global__ void k(const uint32_t* ph, float* c, float* s, float* c2, float* s2, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = phase_to_angle(ph[i]);
  c[i] = __cosf(a); s[i] = __sinf(a);
  float t = (float)(int32_t)ph[i] * 2.3283064365386963e-10f;  // turns in [-0.5,0.5)
  sincospif(2.0f * t, &s2[i], &c2[i]);
}
int main() {
  const int n = 1 << 22;
  std::vector<uint32_t> h(n);
  uint64_t x = 88172645463325252ull;
  for (int i = 0; i < n; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; h[i] = (uint32_t)(x >> 16); }
  uint32_t* d; float *c, *s, *c2, *s2;
  cudaMalloc(&d, n * 4); cudaMalloc(&c, n * 4); cudaMalloc(&s, n * 4); cudaMalloc(&c2, n * 4); cudaMalloc(&s2, n * 4);
  cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
  k<<<(n + 255) / 256, 256>>>(d, c, s, c2, s2, n);
  std::vector<float> hc(n), hs(n), hc2(n), hs2(n);
  cudaMemcpy(hc.data(), c, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hs.data(), s, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hc2.data(), c2, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hs2.data(), s2, n * 4, cudaMemcpyDeviceToHost);
  double mx = 0, ss = 0, mx2 = 0, ss2 = 0;
  for (int i = 0; i < n; i++) {
    double turn = (double)h[i] / 4294967296.0;            // exact fraction of a turn
    double ec = -cos(2 * M_PI * turn), es = -sin(2 * M_PI * turn);   // kernel computes angle - pi
    double e = fmax(fabs(hc[i] - ec), fabs(hs[i] - es));
    mx = fmax(mx, e); ss += (hc[i] - ec) * (hc[i] - ec) + (hs[i] - es) * (hs[i] - es);
    double tc = cos(2 * M_PI * turn), ts = sin(2 * M_PI * turn);
    double e2 = fmax(fabs(hc2[i] - tc), fabs(hs2[i] - ts));
    mx2 = fmax(mx2, e2); ss2 += (hc2[i] - tc) * (hc2[i] - tc) + (hs2[i] - ts) * (hs2[i] - ts);
  }
  printf("MUFU path : max abs err %.3e  rms %.3e\n", mx, sqrt(ss / (2.0 * n)));
  printf("sincospif : max abs err %.3e  rms %.3e\n", mx2, sqrt(ss2 / (2.0 * n)));
  return 0;
}
