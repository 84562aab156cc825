// Diagnostic for the fp16 + fp8-correction split of the gradient contraction:
//   A  : can tcgen05.mma kind::f16 and kind::f8f6f4 (E4M3) accumulate into the SAME fp32 TMEM accumulator?
//        D = A16 B16^T (4 x K=16) + A8 B8^T (4 x K=32), checked against the host.
//   B  : tensor-pipe time per 32-visibility stage on all SMs, sustained (power-capped) for
//        mode 0: 12 x f16 (the fp16x3 split)   mode 1: 4 x f16 + 4 x f8 (the mixed split)
//        mode 2: 4 x f16 only                  mode 3: 4 x f8 only
//   C  : issue rate of the conversions the operand generators would need
//        (cvt.rn.f16x2.f32, cvt.rn.satfinite.e4m3x2.f32, cvt.rn.satfinite.e4m3x2.f16x2).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_fp8mix umma_fp8mix.cu
#include <cstdio>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
// copy `rows` rows of 128 bytes into the K-major SWIZZLE_128B layout
__device__ void fill_sw128(uint8_t* dst, const uint8_t* src, int rows, int tid, int nthr) {
  for (int idx = tid; idx < rows * 8; idx += nthr) {
    const int r = idx / 8, c = idx % 8;
    *reinterpret_cast<uint4*>(dst + (r / 8) * 1024 + (r % 8) * 128 + ((c ^ (r % 8)) * 16)) = *reinterpret_cast<const uint4*>(src + r * 128 + c * 16);
  }
}

// ------------------------------------------------------------------ part A
__global__ void __launch_bounds__(128, 1) k_mix(const uint8_t* A16, const uint8_t* B16, const uint8_t* A8, const uint8_t* B8, float* D, int use16, int use8) {
  extern __shared__ uint8_t raw[];
  const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* s = raw + (sbase - smem_u32(raw));
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  fill_sw128(s, A16, 128, tid, 128);            // 16 KB
  fill_sw128(s + 16384, A8, 128, tid, 128);     // 16 KB
  fill_sw128(s + 32768, B16, 64, tid, 128);     // 8 KB
  fill_sw128(s + 40960, B8, 64, tid, 128);      // 8 KB
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
    const uint32_t idesc = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);   // formats 0: F16 / E4M3
    uint32_t acc = 0;
    if (use16) for (int st = 0; st < 4; st++) { mma_f16(tmem, desc_sw128(sbase + st * 32), desc_sw128(sbase + 32768 + st * 32), idesc, acc); acc = 1; }
    if (use8) for (int st = 0; st < 4; st++) { mma_f8(tmem, desc_sw128(sbase + 16384 + st * 32), desc_sw128(sbase + 40960 + st * 32), idesc, acc); acc = 1; }
    commit(smem_u32(&bar));
  }
  wait_bar(smem_u32(&bar), 0);
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

// ------------------------------------------------------------------ part B
// smem: Ah16 | Al16 | A8 (16 KB each) | Bh16 | Bl16 | B8 (32 KB each, N = 256)
template <int mode>
__global__ void __launch_bounds__(128, 1) k_rate(const uint8_t* src, int iters, long long* cycles) {
  extern __shared__ uint8_t raw[];
  const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* s = raw + (sbase - smem_u32(raw));
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  fill_sw128(s, src, 3 * 128 + 3 * 256, tid, 128);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t ah = sbase, al = sbase + 16384, a8 = sbase + 32768;
    const uint32_t bh = sbase + 49152, bl = bh + 32768, b8 = bl + 32768;
    const long long t0 = clock64();
    uint32_t par = 0;
    for (int it = 0; it < iters; it++) {
      if (mode <= 3) {
        for (int nb = 0; nb < 2; nb++) {
          const uint32_t d = tmem + nb * 256;
          uint32_t acc = 0;
          if (mode == 0 || mode == 1 || mode == 2)
#pragma unroll
            for (int st = 0; st < 4; st++) {
              mma_f16(d, desc_sw128(ah + st * 32), desc_sw128(bh + st * 32), idesc, acc); acc = 1;
              if (mode == 0) {
                mma_f16(d, desc_sw128(ah + st * 32), desc_sw128(bl + st * 32), idesc, 1);
                mma_f16(d, desc_sw128(al + st * 32), desc_sw128(bh + st * 32), idesc, 1);
              }
            }
          if (mode == 1 || mode == 3)
#pragma unroll
            for (int st = 0; st < 4; st++) { mma_f8(d, desc_sw128(a8 + st * 32), desc_sw128(b8 + st * 32), idesc, acc); acc = 1; }
        }
      } else if (mode == 4) {          // mixed, one kind switch per stage: f16 for both accumulators, then f8 for both
#pragma unroll
        for (int nb = 0; nb < 2; nb++)
#pragma unroll
          for (int st = 0; st < 4; st++) mma_f16(tmem + nb * 256, desc_sw128(ah + st * 32), desc_sw128(bh + st * 32), idesc, st > 0);
#pragma unroll
        for (int nb = 0; nb < 2; nb++)
#pragma unroll
          for (int st = 0; st < 4; st++) mma_f8(tmem + nb * 256, desc_sw128(a8 + st * 32), desc_sw128(b8 + st * 32), idesc, 1);
      } else if (mode == 5) {          // f16 only, the A operand shared by consecutive instructions (accumulators alternate)
#pragma unroll
        for (int st = 0; st < 4; st++)
#pragma unroll
          for (int nb = 0; nb < 2; nb++) mma_f16(tmem + nb * 256, desc_sw128(ah + st * 32), desc_sw128(bh + st * 32), idesc, st > 0);
      } else if (mode == 6) {          // mixed, A shared by consecutive instructions
#pragma unroll
        for (int st = 0; st < 4; st++)
#pragma unroll
          for (int nb = 0; nb < 2; nb++) mma_f16(tmem + nb * 256, desc_sw128(ah + st * 32), desc_sw128(bh + st * 32), idesc, st > 0);
#pragma unroll
        for (int st = 0; st < 4; st++)
#pragma unroll
          for (int nb = 0; nb < 2; nb++) mma_f8(tmem + nb * 256, desc_sw128(a8 + st * 32), desc_sw128(b8 + st * 32), idesc, 1);
      } else if (mode == 7) {          // f16 x 8 on ONE accumulator with distinct operands (Ah Bh / Al Bl alternate)
#pragma unroll
        for (int st = 0; st < 4; st++) {
          mma_f16(tmem, desc_sw128(ah + st * 32), desc_sw128(bh + st * 32), idesc, st > 0);
          mma_f16(tmem, desc_sw128(al + st * 32), desc_sw128(bl + st * 32), idesc, 1);
        }
      }
      if ((it & 63) == 63) {            // keep the issue queue bounded like a real pipeline does
        commit(smem_u32(&bar));
        wait_bar(smem_u32(&bar), par);
        par ^= 1;
      }
    }
    commit(smem_u32(&bar));
    wait_bar(smem_u32(&bar), par);
    cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ------------------------------------------------------------------ part C
template <int OP>
__global__ void __launch_bounds__(1024, 1) k_cvt(float* out, int iters, long long* cycles) {
  float x[8];
  uint32_t acc[8];
  for (int i = 0; i < 8; i++) { x[i] = 0.37f + 0.01f * i + 1e-4f * threadIdx.x; acc[i] = 0; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (OP == 0) {          // cvt.rn.f16x2.f32
        uint32_t r;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[i]), "f"(x[(i + 1) & 7]));
        acc[i] ^= r;
      } else if (OP == 1) {   // cvt.rn.satfinite.e4m3x2.f32
        uint16_t r;
        asm volatile("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(x[i]), "f"(x[(i + 1) & 7]));
        acc[i] ^= r;
      } else if (OP == 2) {   // cvt.rn.satfinite.e4m3x2.f16x2
        uint16_t r;
        asm volatile("cvt.rn.satfinite.e4m3x2.f16x2 %0, %1;" : "=h"(r) : "r"(__float_as_uint(x[i])));
        acc[i] ^= r;
      } else {                // reference: one FFMA
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(1.0001f), "f"(1e-6f));
      }
    }
    x[it & 7] += 1e-3f;
  }
  const long long t1 = clock64();
  uint32_t a = 0; float f = 0;
  for (int i = 0; i < 8; i++) { a ^= acc[i]; f += x[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = f + __uint_as_float(a & 0x3fffff);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static uint8_t to_e4m3(float x) { return (uint8_t)__nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E4M3); }
static float from_e4m3(uint8_t b) { __half_raw h = __nv_cvt_fp8_to_halfraw(b, __NV_E4M3); return __half2float(__half(h)); }

int main(int argc, char** argv) {
  const double seconds = argc > 1 ? atof(argv[1]) : 2.0;
  srand(7);
  auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  // ---------------- part A
  {
    std::vector<__half> A16(128 * 64), B16(64 * 64);
    std::vector<uint8_t> A8(128 * 128), B8(64 * 128);
    for (auto& v : A16) v = __float2half(rnd());
    for (auto& v : B16) v = __float2half(rnd() * 100.f);
    for (auto& v : A8) v = to_e4m3(rnd() * 0.05f);
    for (auto& v : B8) v = to_e4m3(rnd() * 60.f);
    uint8_t *dA16, *dB16, *dA8, *dB8; float* dD;
    CK(cudaMalloc(&dA16, A16.size() * 2)); CK(cudaMalloc(&dB16, B16.size() * 2)); CK(cudaMalloc(&dA8, A8.size())); CK(cudaMalloc(&dB8, B8.size()));
    CK(cudaMalloc(&dD, 128 * 64 * 4));
    CK(cudaMemcpy(dA16, A16.data(), A16.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB16, B16.data(), B16.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dA8, A8.data(), A8.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB8, B8.data(), B8.size(), cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024));
    for (int variant = 0; variant < 3; variant++) {
      const int use16 = variant != 1, use8 = variant != 0;
      k_mix<<<1, 128, 49152 + 1024>>>(dA16, dB16, dA8, dB8, dD, use16, use8);
      CK(cudaDeviceSynchronize());
      std::vector<float> D(128 * 64);
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      double num = 0, den = 0;
      for (int m = 0; m < 128; m++) for (int n = 0; n < 64; n++) {
        double ref = 0;
        if (use16) for (int k = 0; k < 64; k++) ref += (double)__half2float(A16[m * 64 + k]) * __half2float(B16[n * 64 + k]);
        if (use8) for (int k = 0; k < 128; k++) ref += (double)from_e4m3(A8[m * 128 + k]) * from_e4m3(B8[n * 128 + k]);
        num += (ref - D[m * 64 + n]) * (ref - D[m * 64 + n]); den += ref * ref;
      }
      printf("A: %s%s into one accumulator: rel-L2 vs host = %.3e\n", use16 ? "f16 " : "", use8 ? "e4m3 " : "", sqrt(num / den));
    }
  }
  // ---------------- part B
  {
    const size_t bytes = (3 * 128 + 3 * 256) * 128;
    std::vector<uint8_t> src(bytes);
    // rows: Ah16 (128) Al16 (128) A8 (128) Bh16 (256) Bl16 (256) B8 (256); 128 B each
    auto fill16 = [&](size_t row0, int rows, float scale) { __half* p = reinterpret_cast<__half*>(src.data() + row0 * 128); for (int i = 0; i < rows * 64; i++) p[i] = __float2half(rnd() * scale); };
    auto fill8 = [&](size_t row0, int rows, float scale) { uint8_t* p = src.data() + row0 * 128; for (int i = 0; i < rows * 128; i++) p[i] = to_e4m3(rnd() * scale); };
    fill16(0, 128, 1.f); fill16(128, 128, 2.4e-4f); fill8(256, 128, 0.06f);
    fill16(384, 256, 16000.f); fill16(640, 256, 8.f); fill8(896, 256, 60.f);
    uint8_t* dsrc; long long* dcyc;
    int dev = 0, sms = 0; CK(cudaGetDevice(&dev)); CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaMalloc(&dsrc, bytes)); CK(cudaMalloc(&dcyc, sms * 8));
    CK(cudaMemcpy(dsrc, src.data(), bytes, cudaMemcpyHostToDevice));
    const int smem = (int)bytes + 1024;
    auto launch = [&](int mode, int iters) {
#define L(m) case m: CK(cudaFuncSetAttribute(k_rate<m>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); k_rate<m><<<sms, 128, smem>>>(dsrc, iters, dcyc); break;
      switch (mode) { L(0) L(1) L(2) L(3) L(4) L(5) L(6) L(7) }
#undef L
    };
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const char* names[8] = {"12 x f16 (fp16x3)", "4 x f16 + 4 x f8 (mixed)", "4 x f16", "4 x f8", "mixed, 1 switch/stage", "4 x f16, A shared", "mixed, A shared", "8 x f16 one accumulator"};
    for (int mode = 0; mode < 8; mode++) {
      const int iters = 20000;
      launch(mode, iters);   // warm-up
      CK(cudaDeviceSynchronize());
      float ms = 0; int launches = 0; double total_ms = 0;
      std::vector<long long> cyc(sms);
      while (total_ms < seconds * 1000) {
        CK(cudaEventRecord(e0));
        launch(mode, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        total_ms += ms; launches++;
      }
      CK(cudaMemcpy(cyc.data(), dcyc, sms * 8, cudaMemcpyDeviceToHost));
      long long mx = 0; for (auto c : cyc) mx = c > mx ? c : mx;
      // one iteration = one 32-visibility stage of a 128 x 512 tile per SM (two N = 256 accumulators)
      const double ns_per_stage = ms * 1e6 / iters, clk_per_stage = (double)mx / iters;
      const double useful = 2.0 * 128 * 512 * 64 * (double)iters * sms / (ms * 1e-3) / 1e12;   // K' = 64 real MACs per stage
      printf("B: mode %d %-26s last launch %.2f ms, %.1f ns/stage, %.1f clk/stage (%.2f GHz), useful %.1f TFLOP/s (all %d SMs, after %.1f s)\n",
             mode, names[mode], ms, ns_per_stage, clk_per_stage, clk_per_stage / ns_per_stage, useful, sms, total_ms / 1000);
    }
  }
  // ---------------- part C
  {
    float* dout; long long* dcyc;
    CK(cudaMalloc(&dout, 148 * 1024 * 4)); CK(cudaMalloc(&dcyc, 148 * 8));
    const int iters = 20000;
    const char* names[4] = {"cvt.rn.f16x2.f32", "cvt.rn.satfinite.e4m3x2.f32", "cvt.rn.satfinite.e4m3x2.f16x2", "fma.rn.f32"};
    for (int op = 0; op < 4; op++) {
      if (op == 0) k_cvt<0><<<148, 1024>>>(dout, iters, dcyc);
      if (op == 1) k_cvt<1><<<148, 1024>>>(dout, iters, dcyc);
      if (op == 2) k_cvt<2><<<148, 1024>>>(dout, iters, dcyc);
      if (op == 3) k_cvt<3><<<148, 1024>>>(dout, iters, dcyc);
      CK(cudaDeviceSynchronize());
      long long c; CK(cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost));
      printf("C: %-32s %.2f thread-ops/clk/SM (loop carries 1 FADD + XORs per 8 ops)\n", names[op], 1024.0 * 8 * iters / (double)c);
    }
  }
  return 0;
}
