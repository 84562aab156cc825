// Diagnostic: arithmetic of tcgen05.mma kind::f16 with fp32 accumulation on B200 — how many bits
// below the largest addend survive inside one K=16 instruction, and how the accumulator rounds.
// One CTA; A is 128 x 64 fp16 (4 K-steps of 16), B is 64 x 64 fp16; D = sum over 4 sequential MMAs.
#include <cstdio>
#include <cmath>
#include <cstdint>
#include <vector>
#include <cuda_fp16.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__global__ void __launch_bounds__(128, 1) k(const __half* A, const __half* B, float* D, int nsteps) {
  extern __shared__ uint8_t raw[];
  const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* s = raw + (sbase - smem_u32(raw));
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  // A tile at s[0..16K), B tile at s[16K..24K): row r, 16B chunk c -> r/8*1024 + r%8*128 + ((c ^ r%8) * 16)
  for (int idx = tid; idx < 128 * 8; idx += 128) {
    int r = idx / 8, c = idx % 8;
    *reinterpret_cast<uint4*>(s + (r / 8) * 1024 + (r % 8) * 128 + ((c ^ (r % 8)) * 16)) = *reinterpret_cast<const uint4*>(A + r * 64 + c * 8);
  }
  for (int idx = tid; idx < 64 * 8; idx += 128) {
    int r = idx / 8, c = idx % 8;
    *reinterpret_cast<uint4*>(s + 16384 + (r / 8) * 1024 + (r % 8) * 128 + ((c ^ (r % 8)) * 16)) = *reinterpret_cast<const uint4*>(B + r * 64 + c * 8);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    for (int st = 0; st < nsteps; st++) {
      uint32_t acc = st > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                   "l"(desc_sw128(sbase + st * 32)), "l"(desc_sw128(sbase + 16384 + st * 32)), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int cb = 0; cb < 2; cb++) {
    uint32_t v[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + cb * 32) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 32; c++) D[tid * 64 + cb * 32 + c] = __uint_as_float(v[c]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}
static std::vector<__half> hA(128 * 64), hB(64 * 64);
static std::vector<float> hD(128 * 64);
static __half *dA, *dB; static float* dD;
static void run(int nsteps) {
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  k<<<1, 128, 16384 + 8192 + 1024>>>(dA, dB, dD, nsteps);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
}
static void clear() { for (auto& x : hA) x = __float2half(0.f); for (auto& x : hB) x = __float2half(0.f); }
int main() {
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, hD.size() * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 8192 + 1024);
  // sanity: random small integers, K = 64
  clear();
  srand(1);
  for (auto& x : hA) x = __float2half((float)(rand() % 7 - 3));
  for (auto& x : hB) x = __float2half((float)(rand() % 7 - 3));
  run(4);
  double maxerr = 0;
  for (int m = 0; m < 128; m++) for (int n = 0; n < 64; n++) {
    double ref = 0; for (int kk = 0; kk < 64; kk++) ref += (double)__half2float(hA[m * 64 + kk]) * __half2float(hB[n * 64 + kk]);
    maxerr = fmax(maxerr, fabs(ref - hD[m * 64 + n]));
  }
  printf("sanity (integer GEMM 128x64x64): max err %.1f\n", maxerr);
  // T1: inside ONE instruction: +2^24 - 2^24 + 14 * 2^(24-n): which n survive?
  clear();
  for (int m = 0; m < 128; m++) { hA[m * 64 + 0] = __float2half(4096.f); hA[m * 64 + 1] = __float2half(-4096.f); }
  for (int n = 0; n < 64; n++) {
    hB[n * 64 + 0] = __float2half(4096.f); hB[n * 64 + 1] = __float2half(4096.f);
    int sh = n < 48 ? n : 48;
    float a = ldexpf(1.f, 12 - (sh + 1) / 2), b = ldexpf(1.f, 12 - sh / 2);
    for (int kk = 2; kk < 16; kk++) { for (int m = 0; m < 128; m++) hA[m * 64 + kk] = __float2half(1.f); hB[n * 64 + kk] = __float2half(a * b / 1.f > 65000.f ? 0.f : 0.f); }
    // put the small product as A=2^x (per row same) * B: A must be per-row constant, so carry all of it in B with A = 2^-8
  }
  // simpler encoding: A[m][k>=2] = 2^-10, B[n][k>=2] = 2^(34-n) clipped to fp16 range (n >= 19 -> <= 2^15)
  for (int m = 0; m < 128; m++) for (int kk = 2; kk < 16; kk++) hA[m * 64 + kk] = __float2half(ldexpf(1.f, -10));
  for (int n = 0; n < 64; n++) for (int kk = 2; kk < 16; kk++) {
    int e = 34 - n; hB[n * 64 + kk] = (e <= 15 && e >= -24) ? __float2half(ldexpf(1.f, e)) : __float2half(0.f);
  }
  run(1);
  printf("T1 one instruction: 2^24 - 2^24 + 14*2^(24-n); got/expected per n (expected exact 14*2^(24-n)):\n");
  for (int n = 19; n < 58; n++) { int e = 34 - n; if (e < -24) break; printf("  n=%2d small=2^%d: got %.6g expected %.6g\n", n, 24 - n, hD[n], 14.0 * ldexp(1.0, 24 - n)); }
  // T2: accumulator holds 2^24 (step 0), step 1 adds 14 * 2^(24-n) (n as above), step 2 adds -2^24
  clear();
  for (int m = 0; m < 128; m++) { hA[m * 64 + 0] = __float2half(4096.f); hA[m * 64 + 32] = __float2half(-4096.f); }
  for (int n = 0; n < 64; n++) { hB[n * 64 + 0] = __float2half(4096.f); hB[n * 64 + 32] = __float2half(4096.f); }
  for (int m = 0; m < 128; m++) for (int kk = 18; kk < 32; kk++) hA[m * 64 + kk] = __float2half(ldexpf(1.f, -10));
  for (int n = 0; n < 64; n++) for (int kk = 18; kk < 32; kk++) { int e = 34 - n; hB[n * 64 + kk] = (e <= 15 && e >= -24) ? __float2half(ldexpf(1.f, e)) : __float2half(0.f); }
  run(3);
  printf("T2 three instructions: acc=2^24; += 14*2^(24-n); += -2^24:\n");
  for (int n = 19; n < 40; n++) printf("  n=%2d small=2^%d: got %.6g expected %.6g\n", n, 24 - n, hD[n], 14.0 * ldexp(1.0, 24 - n));
  // T3: rounding of the accumulator: acc = 2^24 (ulp 2); add x in {1, 3, 5, -1, -3, 0.5, 1.5}; then subtract 2^24
  clear();
  const float xs[8] = {1.f, 3.f, 5.f, -1.f, -3.f, 0.5f, 1.5f, 2.5f};
  for (int m = 0; m < 128; m++) { hA[m * 64 + 0] = __float2half(4096.f); hA[m * 64 + 16] = __float2half(1.f); hA[m * 64 + 32] = __float2half(-4096.f); }
  for (int n = 0; n < 64; n++) { hB[n * 64 + 0] = __float2half(4096.f); hB[n * 64 + 16] = __float2half(xs[n % 8]); hB[n * 64 + 32] = __float2half(4096.f); }
  run(3);
  printf("T3 rounding: (2^24 + x) - 2^24 for x:\n");
  for (int n = 0; n < 8; n++) printf("  x=%5.1f -> %.1f\n", xs[n], hD[n]);
  // T4: same but negative accumulator
  for (int m = 0; m < 128; m++) { hA[m * 64 + 0] = __float2half(-4096.f); hA[m * 64 + 32] = __float2half(4096.f); }
  run(3);
  printf("T4 rounding: (-2^24 + x) + 2^24 for x:\n");
  for (int n = 0; n < 8; n++) printf("  x=%5.1f -> %.1f\n", xs[n], hD[n]);
  return 0;
}
