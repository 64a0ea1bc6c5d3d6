// Micro-benchmarks of the sm_100a CUDA-core pipes the demodulators lean on: FFMA (3-reg), FFMA2 (fma.rn.f32x2),
// DFMA, F2F.F64.F32, LDS.128.  Prints thread-ops per clock per SM.  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#define ITERS 4096
#define NACC 16

__global__ void k_ffma(float* out, float a, float b) {
  float acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 0.001f + i;
  float x = a + threadIdx.x, y = b;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], x, y);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float* out, float a, float b) {
  unsigned long long acc[NACC];
  float x = a + threadIdx.x, y = b;
  unsigned long long x2, y2;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x2) : "f"(x), "f"(x + 1.f));
  asm("mov.b64 %0, {%1, %2};" : "=l"(y2) : "f"(y), "f"(y + 1.f));
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    float lo = threadIdx.x * 0.001f + i, hi = lo + 0.5f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(lo), "f"(hi));
  }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(x2), "l"(y2));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dfma(float* out, double a, double b) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 0.001 + i;
  double x = a + threadIdx.x, y = b;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}

// F2F.F64.F32 + DFMA pairs: x float -> double, acc = fma(xd, w, acc)
__global__ void k_cvt_dfma(float* out, const float* in, double w) {
  double acc[NACC];
  float xs[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc[i] = i; xs[i] = in[threadIdx.x + i]; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      float xf;
      asm volatile("add.f32 %0, %1, 0f3F800000;" : "=f"(xf) : "f"(xs[i]));   // keep the conversion inside the loop
      xs[i] = xf;
      acc[i] = fma((double)xf, w, acc[i]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}

// mixed: 1 LDS.128 per 8 FFMA2 (the FIR inner-loop mix)
__global__ void k_mix(float* out, float a) {
  __shared__ float4 sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(a + i, a, a * i, 1.f);
  __syncthreads();
  unsigned long long acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0ull;
  unsigned long long y2;
  asm("mov.b64 %0, {%1, %2};" : "=l"(y2) : "f"(a), "f"(a + 1.f));
  int idx = threadIdx.x & 31;
  for (int it = 0; it < ITERS; ++it) {
    const float4 t0 = sm[(idx + it) & 1023];
    const float4 t1 = sm[(idx + it + 512) & 1023];
    unsigned long long p0, p1, p2, p3;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(t0.x), "f"(t0.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(t0.z), "f"(t0.w));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(t1.x), "f"(t1.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(t1.z), "f"(t1.w));
#pragma unroll
    for (int i = 0; i < NACC; i += 4) {
      asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i]) : "l"(p0), "l"(y2));
      asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i + 1]) : "l"(p1), "l"(y2));
      asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i + 2]) : "l"(p2), "l"(y2));
      asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i + 3]) : "l"(p3), "l"(y2));
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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
  const int threads = 512, blocks = sms * 4;
  float *out, *in;
  cudaMalloc(&out, (size_t)blocks * threads * 4);
  cudaMalloc(&in, 4096 * 4);
  cudaMemset(in, 0, 4096 * 4);
  printf("%s sms=%d clock=%d kHz\n", p.name, sms, clk_khz);
  const double total = (double)blocks * threads * ITERS * NACC;
  auto rep = [&](const char* name, float ms, double ops_per_inst) {
    const double per_clk_sm = total * ops_per_inst / (ms * 1e-3) / sms / (clk_khz * 1e3);
    printf("%-12s %8.3f ms  %7.1f thread-ops/clk/SM (at nominal %d MHz)  %.2f Tops/s\n", name, ms, per_clk_sm, clk_khz / 1000,
           total * ops_per_inst / (ms * 1e-3) / 1e12);
  };
  rep("FFMA", timeit([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 1);
  rep("FFMA2(x2)", timeit([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 2);
  rep("DFMA", timeit([&] { k_dfma<<<blocks, threads>>>(out, 1.0001, 0.5); }), 1);
  rep("CVT+DFMA", timeit([&] { k_cvt_dfma<<<blocks, threads>>>(out, in, 0.5); }), 1);
  rep("LDS+FFMA2", timeit([&] { k_mix<<<blocks, threads>>>(out, 1.0f); }), 2);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
