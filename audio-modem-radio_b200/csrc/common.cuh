// Shared declarations for libfbdsp.so (sm_100a).  Host-side handle + device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fbdsp.h"

#define FB_THREADS 256

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

// Per-recording geometry, computed on the host (psk_plan.cpp logic lives in psk_v2.cu) and read by every kernel.
struct RecPlan {
  uint64_t off;        // first sample of the recording inside `samples` (elements)
  uint64_t n;          // samples
  uint64_t word_off;   // first 32-bit word of this recording's decided bit stream in the workspace
  uint64_t out_off;    // first byte of the output slot
  uint64_t out_cap;    // bytes available in the output slot
  int32_t nsym, ndsym; // symbols / differential symbols
  int32_t dl32, dr32;  // main-kernel dsym range [dl32, dr32); empty when the edge kernel takes the whole recording
  int32_t status;      // FB_ST_*
  int32_t pad;
};

// One window evaluated by the edge kernel (float64, the reference recurrences step by step).
struct EdgeJob {
  int32_t rec, k_lo, k_hi, pad;
  int64_t wa, wb;      // band-pass window (sample indices of the odd-extended record)
  int64_t fa, fb;      // range on which the band-pass output f is needed
  int64_t la, lb;      // low-pass window (sample indices of the odd-extended mixed record)
  uint64_t scratch_off;// doubles
};

struct fb_handle {
  // Every work-submitting entry point holds this for the whole call, so two host threads that share a handle (the
  // reference calls decode_from_buffer from the Qt GUI thread and from a QThread, filebeep_advanced_v2.py:324,1112) are
  // serialised instead of interleaving launches and workspace growth.  Recursive: fb_psk_demod_batch re-enters itself.
  std::recursive_mutex mu;
  int device = 0, sm_count = 148;
  cudaStream_t stream = nullptr, stream2 = nullptr, stream_copy = nullptr;   // work, edge kernels (high priority), host->device staging
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;
  cudaEvent_t ev_copy[16] = {};   // one per in-flight host->device group (fb_psk_demod_batch pipelines copies and kernels)
  bool profiling = false, k_recorded = false;
  uint64_t launches = 0;
  std::string err;
  // device workspace (grown on demand, never shrunk)
  DevBuf in, out, out_len, sync_idx, status, bits, plans, tile_first, tiles, jobs, scratch, taps, slow_w, sync_raw;
  DevBuf fec_in, fec_out, fec_meta, misc, redo, mma_trace, fftws, psk_tabs;
  std::vector<void*> mma_cache;   // psk_mma.cu: tensor-core tables per design (host + device copies), built once
  // host copy of the last PSK plan (fb_psk_last_bits)
  std::vector<RecPlan> last_plans;
  // fp32 DPSK kernel: the design-dependent tables of the last call (psk_tabs on the device, the constant-bank part on the host);
  // a call with the same design and taps re-uses them instead of rebuilding and uploading
  std::vector<unsigned char> psk_tab_key;
  std::vector<unsigned char> psk_tab_host;
  int last_bps = 0;
};

#define FB_LOCK(h) std::lock_guard<std::recursive_mutex> fb_lock__((h)->mu)

#define FB_CUDA(h, call)                                                                       \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      (h)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                          \
      return (e__ == cudaErrorMemoryAllocation) ? FB_ENOMEM : FB_ECUDA;                        \
    }                                                                                          \
  } while (0)

int fb_ensure(fb_handle* h, DevBuf& b, size_t bytes);
void fb_fsk_release(fb_handle* h);   // fsk_v2.cu
void fb_resample_release(fb_handle* h);   // resample.cu

// bits back end (backend.cu): first-occurrence magic search + shifted byte packing
int fb_bits_backend(fb_handle* h, int n_rec, const RecPlan* d_plans, const std::vector<RecPlan>& plans, int bps,
                    const uint32_t* d_bits, uint8_t* d_out, uint64_t* d_out_len, int64_t* d_sync, int32_t* d_status);

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cfma(float2 a, float2 b, float2 c) {  // a*b + c
  return make_float2(fmaf(a.x, b.x, fmaf(-a.y, b.y, c.x)), fmaf(a.x, b.y, fmaf(a.y, b.x, c.y)));
}
// Dibit / bit decision on d = y[k+1] conj(y[k]) rho.  Returns the 2-bit code (b0<<1|b1) for DQPSK
// (modem.py:219-241: 00 | 01 | 11 | 10 by 90-degree sectors centred on 0, pi/2, pi, -pi/2) or the single bit for DBPSK
// (modem.py:105: Re(d) < 0 -> 1).  d == 0: np.angle is atan2(+-0, re): 0 for re = +0 -> 00, +-pi for re = -0 -> 11 (products
// of symbols below 1e-162 underflow to signed zeros in the reference, too).
template <typename R>
__device__ __forceinline__ uint32_t psk_decide(R dr, R di, int bps) {
  const R a = dr + di, b = dr - di;
  // a > 0: 00 | 01;  else b < 0: 11;  else origin -> 00 / 11 by the sign of the real zero, otherwise 10
  const uint32_t q = (a > R(0)) ? ((b > R(0)) ? 0u : 1u) : ((b < R(0)) ? 3u : ((a == R(0) && b == R(0)) ? (signbit(dr) ? 3u : 0u) : 2u));
  const uint32_t p = dr < R(0) ? 1u : 0u;
  return bps == 1 ? p : q;
}
template <typename T> __device__ __forceinline__ float load_sample(const void* base, uint64_t i);
template <> __device__ __forceinline__ float load_sample<float>(const void* base, uint64_t i) {
  return __ldg(reinterpret_cast<const float*>(base) + i);
}
template <> __device__ __forceinline__ float load_sample<double>(const void* base, uint64_t i) {
  return (float)__ldg(reinterpret_cast<const double*>(base) + i);
}
template <> __device__ __forceinline__ float load_sample<int16_t>(const void* base, uint64_t i) {
  return (float)__ldg(reinterpret_cast<const int16_t*>(base) + i) * (1.0f / 32768.0f);
}
template <typename T> __device__ __forceinline__ double load_sample_d(const void* base, uint64_t i);
template <> __device__ __forceinline__ double load_sample_d<float>(const void* base, uint64_t i) {
  return (double)__ldg(reinterpret_cast<const float*>(base) + i);
}
template <> __device__ __forceinline__ double load_sample_d<double>(const void* base, uint64_t i) {
  return __ldg(reinterpret_cast<const double*>(base) + i);
}
template <> __device__ __forceinline__ double load_sample_d<int16_t>(const void* base, uint64_t i) {
  return (double)__ldg(reinterpret_cast<const int16_t*>(base) + i) * (1.0 / 32768.0);
}
#endif
