// Micro-benchmarks of the warp-level tensor path (mma.sync -> HMMA) and ldmatrix on sm_100a, for the banded-Toeplitz
// FIR of psk_main (fp16 hi/lo split operands, fp32 accumulation).  Prints warp-instructions per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tensor tensor.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define ITERS 2048

__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

template <int KIND, int NACC>
__global__ void k_mma(float* out, uint32_t seed) {
  float d[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
  uint32_t a[4] = {seed, seed + 1, seed + 2, seed + 3}, b[2] = {seed ^ 5, seed ^ 9};
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (KIND == 0) mma_f16(d[i], a, b);
      else if (KIND == 1) mma_bf16(d[i], a, b);
      else mma_tf32(d[i], a, b);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ldmatrix.x4 alone: rows at a 160-byte pitch (80 fp16 samples), the A fragment of a 16-window m-tile
template <int PITCH>
__global__ void k_ldsm(float* out) {
  extern __shared__ __align__(16) unsigned char sm[];
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + (uint32_t)((lane & 15) * PITCH + (lane >> 4) * 16 + (warp & 3) * 2560);
  uint32_t acc = 0;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      uint32_t r[4];
      ldsm4(r, base + (uint32_t)(((it & 7) * 8 + k) * 32));
      acc ^= r[0] + r[1] + r[2] + r[3];
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
}

// the planned FIR step: 2 ldmatrix.x4 (hi, lo sample fragments) + 6 HMMA (hh, hl, lh  x  2 n-tiles), taps in registers;
// FMA2 > 0 adds that many independent FFMA2 per step (pipe overlap with the slow-pole / slicer work)
template <int FMA2, int NM>
__global__ void k_fir(float* out, uint32_t seed) {
  extern __shared__ __align__(16) unsigned char sm[];
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + (uint32_t)((lane & 15) * 160 + (lane >> 4) * 16 + (warp & 7) * 2560);
  float d[NM][2][2][4];
#pragma unroll
  for (int m = 0; m < NM; ++m)
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int j = 0; j < 4; ++j) d[m][i][h][j] = 0.f;
  uint32_t bh[2][2] = {{seed, seed + 1}, {seed + 2, seed + 3}}, bl[2][2] = {{seed ^ 5, seed ^ 9}, {seed ^ 6, seed ^ 10}};
  unsigned long long f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = 0x3f8000003f800000ull + i;
  const unsigned long long w = 0x3f7ff0003f7ff000ull;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int m = 0; m < NM; ++m) {
        uint32_t ah[4], al[4];
        ldsm4(ah, base + (uint32_t)(m * 20480 + ((it & 3) * 4 + k) * 32));
        ldsm4(al, base + (uint32_t)(m * 20480 + 40960 + ((it & 3) * 4 + k) * 32));
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          mma_f16(d[m][n][0], ah, bh[n]);
          mma_f16(d[m][n][1], ah, bl[n]);
          mma_f16(d[m][n][1], al, bh[n]);
        }
      }
#pragma unroll
      for (int q = 0; q < FMA2; ++q) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(f[q & 7]) : "l"(w));
    }
  }
  float s = 0;
#pragma unroll
  for (int m = 0; m < NM; ++m)
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h) s += d[m][i][h][0] + d[m][i][h][1] + d[m][i][h][2] + d[m][i][h][3];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += (float)(f[i] & 0xff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// fp32 -> (hi, lo) fp16 split of sample pairs: the staging conversion
__global__ void k_split(float* out, const float* in) {
  float2 x = make_float2(in[threadIdx.x & 255], in[(threadIdx.x + 1) & 255]);
  uint32_t acc = 0;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const __half2 hi = __float22half2_rn(x);
      const float2 hf = __half22float2(hi);
      const float2 r = make_float2((x.x - hf.x) * 2048.f, (x.y - hf.y) * 2048.f);
      const __half2 lo = __float22half2_rn(r);
      acc += *reinterpret_cast<const uint32_t*>(&hi) ^ *reinterpret_cast<const uint32_t*>(&lo);
      x.x += 0.001f; x.y -= 0.001f;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
}

// rounding of the fp32 accumulation inside HMMA: C = 1, one product of 2^-24 * q, q = 1, 2, 3 -> RN gives 1, 1 (tie to even), 1 + 2^-23
__global__ void k_round(float* out) {
  const int lane = threadIdx.x & 31;
  for (int q = 1; q <= 3; ++q) {
    uint32_t a[4] = {0, 0, 0, 0}, b[2] = {0, 0};
    // A[row g][k = 2*(lane%4)]: put 2^-12 at k = 0 of every row; B[k = 0][n]: q * 2^-12
    const __half2 av = __floats2half2_rn(lane % 4 == 0 ? 0.000244140625f : 0.f, 0.f);
    const __half2 bv = __floats2half2_rn(lane % 4 == 0 ? q * 0.000244140625f : 0.f, 0.f);
    a[0] = *reinterpret_cast<const uint32_t*>(&av); a[1] = a[0];
    b[0] = *reinterpret_cast<const uint32_t*>(&bv);
    float d[4] = {1.f, 1.f, 1.f, 1.f};
    mma_f16(d, a, b);
    float dn[4] = {-1.f, -1.f, -1.f, -1.f};
    a[0] ^= (lane % 4 == 0) ? 0x8000u : 0u; a[1] = a[0];
    mma_f16(dn, a, b);
    if (threadIdx.x == 0) { out[2 * (q - 1)] = (d[0] - 1.f) * 8388608.f; out[2 * (q - 1) + 1] = (dn[0] + 1.f) * 8388608.f; }
  }
  // 16 products of 2^-25 each summed to 2^-21 = 4 ulp of 1: are the products summed before the add to C?
  {
    const __half2 av = __floats2half2_rn(0.000244140625f, 0.000244140625f);
    const __half2 bv = __floats2half2_rn(0.0001220703125f, 0.0001220703125f);
    uint32_t a[4], b[2];
    a[0] = a[1] = a[2] = a[3] = *reinterpret_cast<const uint32_t*>(&av);
    b[0] = b[1] = *reinterpret_cast<const uint32_t*>(&bv);
    float d[4] = {1.f, 1.f, 1.f, 1.f};
    mma_f16(d, a, b);
    if (threadIdx.x == 0) out[6] = (d[0] - 1.f) * 8388608.f;
  }
}

template <typename F> static float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount, clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  float *out, *in;
  cudaMalloc(&out, (size_t)sms * 8 * 1024 * 4);
  cudaMalloc(&in, 4096 * 4);
  cudaMemset(in, 0, 4096 * 4);
  printf("%s sms=%d clock=%d kHz\n", p.name, sms, clk_khz);
  auto rep = [&](const char* name, float ms, double warp_instr_per_thread_iter, int blocks, int threads, double macs) {
    const double wi = (double)blocks * (threads / 32) * ITERS * warp_instr_per_thread_iter;
    const double per = wi / (ms * 1e-3) / sms / (clk_khz * 1e3);
    printf("%-34s %8.3f ms  %6.3f warp-instr/clk/SM  %7.0f MAC/clk/SM\n", name, ms, per, per * macs);
  };
  for (int threads : {128, 256, 512}) {
    const int blocks = sms * 2;
    char nm[64];
    snprintf(nm, 64, "HMMA f16 16816 x8acc  %3d thr", threads);
    rep(nm, timeit([&] { k_mma<0, 8><<<blocks, threads>>>(out, 1); }), 8, blocks, threads, 2048);
    snprintf(nm, 64, "HMMA bf16 16816 x8acc %3d thr", threads);
    rep(nm, timeit([&] { k_mma<1, 8><<<blocks, threads>>>(out, 1); }), 8, blocks, threads, 2048);
    snprintf(nm, 64, "HMMA tf32 1688 x8acc  %3d thr", threads);
    rep(nm, timeit([&] { k_mma<2, 8><<<blocks, threads>>>(out, 1); }), 8, blocks, threads, 1024);
  }
  rep("HMMA f16 x2acc 512 thr (latency)", timeit([&] { k_mma<0, 2><<<sms * 2, 512>>>(out, 1); }), 2, sms * 2, 512, 2048);
  {
    cudaFuncSetAttribute(k_ldsm<160>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    cudaFuncSetAttribute(k_ldsm<144>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    rep("ldmatrix.x4 pitch 160", timeit([&] { k_ldsm<160><<<sms * 2, 512, 48 * 1024>>>(out); }), 8, sms * 2, 512, 0);
    rep("ldmatrix.x4 pitch 144", timeit([&] { k_ldsm<144><<<sms * 2, 512, 48 * 1024>>>(out); }), 8, sms * 2, 512, 0);
  }
  {
    auto run = [&](auto kern, const char* name, int threads, int nm) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      rep(name, timeit([&] { kern<<<sms * 2, threads, 96 * 1024>>>(out, 1); }), 4.0 * 6 * nm, sms * 2, threads, 2048);
    };
    run(k_fir<0, 1>, "FIR step 2 ldsm+6 HMMA 256 thr", 256, 1);
    run(k_fir<0, 1>, "FIR step 2 ldsm+6 HMMA 512 thr", 512, 1);
    run(k_fir<0, 2>, "FIR step x2 m-tiles 256 thr", 256, 2);
    run(k_fir<6, 1>, "FIR step + 6 FFMA2 256 thr", 256, 1);
    run(k_fir<12, 1>, "FIR step + 12 FFMA2 256 thr", 256, 1);
    run(k_fir<24, 1>, "FIR step + 24 FFMA2 256 thr", 256, 1);
  }
  rep("fp32 -> fp16 hi/lo split (pairs)", timeit([&] { k_split<<<sms * 4, 512>>>(out, in); }), 8, sms * 4, 512, 0);
  k_round<<<1, 32>>>(out);
  float r[7];
  cudaMemcpy(r, out, sizeof(r), cudaMemcpyDeviceToHost);
  printf("HMMA accumulate rounding, (d - c) in ulps of 1.0 for exact sums of 0.5, 1.0, 1.5 ulp: +c: %.1f %.1f %.1f   -c: %.1f %.1f %.1f   16 x 0.25 ulp: %.1f\n",
         r[0], r[2], r[4], r[1], r[3], r[5], r[6]);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
