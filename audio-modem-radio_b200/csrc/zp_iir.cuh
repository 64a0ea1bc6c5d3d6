// Zero-phase IIR (scipy.signal.filtfilt: odd extension by padlen, lfilter_zi start-up, DF2T forward then backward) in
// float64, chunk-parallel with coalesced HBM traffic.  Used by the v1 Goertzel FSK pre-filter and the
// v2 FSK tone filters (modem.py:307-308).
//
// The extended record e in [0, Next), Next = N + 2 pad, ext[e] = x_ext(e - pad), is cut into chunks of L positions; one
// thread owns one chunk and runs the recurrence serially: exact state (zi * first value) at the true ends, zero state
// plus >= w warm-up positions (pole decay to 1e-12) at interior cuts.  A CTA = 64 consecutive chunks advancing in
// lock-step through blocks of 16 positions:
//   forward   the 64 x 16 input values of a block are loaded cooperatively (each warp load touches two 64-byte runs
//             instead of 32 scattered sectors) into shared memory; outputs go to a TRANSPOSED scratch
//             y[(e % L) * nch + e / L], so the 64 threads store consecutive doubles
//   backward  reads that scratch coalesced (own chunk, warm-up from the next chunk's head), stages the 64 x 16 outputs
//             in shared memory and stores them cooperatively in natural order
#pragma once
#include "common.cuh"
#include <stdlib.h>
#include <algorithm>

#define ZP_THREADS 64
#define ZP_EB 16
#ifndef ZP_MINB
#define ZP_MINB 8              // resident CTAs per SM the register budget is held to
#endif

template <int ORD>
struct ZpFilt {
  double b[ORD + 1], a[ORD + 1], zi[ORD];
  int32_t w, pad;                 // warm-up positions at interior cuts; filtfilt padlen
};

struct ZpRec {
  uint64_t off;                   // first sample of the recording in `samples` (elements)
  int64_t N;                      // samples (0: skip)
  uint64_t y_off;                 // first double of this recording's transposed forward scratch (nch * L doubles)
  uint64_t out_off;               // first element of this recording in the output buffer
  int64_t nch;                    // chunks = ceil((N + 2 pad) / L)
};

template <typename TIn>
__device__ __forceinline__ double zp_x_ext(const void* samples, uint64_t off, int64_t N, int64_t n) {
  // scipy.signal._arraytools.odd_ext
  if (n < 0) return 2.0 * load_sample_d<TIn>(samples, off) - load_sample_d<TIn>(samples, off + (uint64_t)(-n));
  if (n > N - 1)
    return 2.0 * load_sample_d<TIn>(samples, off + (uint64_t)(N - 1)) - load_sample_d<TIn>(samples, off + (uint64_t)(2 * (N - 1) - n));
  return load_sample_d<TIn>(samples, off + (uint64_t)n);
}

template <int ORD>
__device__ __forceinline__ double zp_step(const ZpFilt<ORD>& t, double (&z)[ORD], double xv) {
  const double y = t.b[0] * xv + z[0];
#pragma unroll
  for (int k = 0; k < ORD - 1; ++k) z[k] = t.b[k + 1] * xv + z[k + 1] - t.a[k + 1] * y;
  z[ORD - 1] = t.b[ORD] * xv - t.a[ORD] * y;
  return y;
}

// grid (ceil(max nch / 64), recordings); W16 = warm-up rounded up to a multiple of 16 (<= L).
// Blocks whose 64 x 16 positions all lie strictly inside the record take a check-free path (plain strided pointers);
// only the first / last CTA of a recording pays for the odd extension and the end-state logic.
template <typename TIn, int ORD>
__global__ void __launch_bounds__(ZP_THREADS, ZP_MINB) zp_fwd_kernel(const void* samples, const ZpRec* recs, ZpFilt<ORD> t, int L, int W16, double* ybase) {
  __shared__ double xs[ZP_THREADS][ZP_EB + 1];
  const ZpRec rc = recs[blockIdx.y];
  const int64_t N = rc.N, Next = N + 2 * t.pad, nch = rc.nch;
  const int64_t cb = (int64_t)blockIdx.x * ZP_THREADS;            // first chunk of this CTA
  if (N <= 0 || cb >= nch) return;
  const int tid = threadIdx.x;
  const int64_t ch = cb + tid, c0 = ch * L;
  double* yp = ybase + rc.y_off + ch;                             // + (position in chunk) * nch
  const TIn* xin = reinterpret_cast<const TIn*>(samples) + rc.off;
  double z[ORD];
#pragma unroll
  for (int i = 0; i < ORD; ++i) z[i] = 0.0;
  const int sg = tid >> 4, su = tid & 15;
  auto is_inner = [&](int kb) {                                     // all 64 x 16 positions of block kb are plain samples
    const int64_t e_lo = cb * L + kb, e_hi = (cb + ZP_THREADS - 1) * L + kb + ZP_EB - 1;
    return kb < L && (e_lo - t.pad >= 0) && (e_hi - t.pad <= N - 1) && (e_lo > 0);
  };
  // raw values: converting here would make the warp wait for the loads before the current block's math
  auto load_inner = [&](int kb, TIn (&v)[ZP_EB]) {                  // this thread's share of the cooperative load
    const TIn* p = xin + (cb * L + kb + (int64_t)sg * L + su - t.pad);
    const int64_t stride = 4 * (int64_t)L;
#pragma unroll
    for (int i = 0; i < ZP_EB; ++i) { v[i] = __ldg(p); p += stride; }
  };
  TIn pre[ZP_EB];                                                   // next block's inputs, in flight during this block's math
  bool pre_ok = is_inner(-W16);
  if (pre_ok) load_inner(-W16, pre);
  for (int kb = -W16; kb < L; kb += ZP_EB) {
    const bool inner = pre_ok;
    if (inner) {
#pragma unroll
      for (int i = 0; i < ZP_EB; ++i) xs[sg + 4 * i][su] = sizeof(TIn) == 2 ? (double)pre[i] * (1.0 / 32768.0) : (double)pre[i];
    } else {
      for (int idx = tid; idx < ZP_THREADS * ZP_EB; idx += ZP_THREADS) {
        const int seg = idx >> 4, u = idx & 15;
        const int64_t e = (cb + seg) * L + kb + u;
        xs[seg][u] = (e >= 0 && e < Next) ? zp_x_ext<TIn>(samples, rc.off, N, e - t.pad) : 0.0;
      }
    }
    __syncthreads();
    pre_ok = is_inner(kb + ZP_EB);
    if (pre_ok) load_inner(kb + ZP_EB, pre);
    if (inner) {
      double* q = yp + (int64_t)kb * nch;
#pragma unroll
      for (int u = 0; u < ZP_EB; ++u) {
        const double yv = zp_step<ORD>(t, z, xs[tid][u]);
        if (kb >= 0) *q = yv;
        q += nch;
      }
    } else if (ch < nch) {
#pragma unroll
      for (int u = 0; u < ZP_EB; ++u) {
        const int64_t e = c0 + kb + u;
        if (e >= 0 && e < Next) {
          const double xv = xs[tid][u];
          if (e == 0) {
#pragma unroll
            for (int i = 0; i < ORD; ++i) z[i] = t.zi[i] * xv;    // lfilter_zi start-up at the true left end
          }
          const double yv = zp_step<ORD>(t, z, xv);
          if (kb + u >= 0) yp[(int64_t)(kb + u) * nch] = yv;
        }
      }
    }
    __syncthreads();
  }
}

template <typename TOut, int ORD>
__global__ void __launch_bounds__(ZP_THREADS, ZP_MINB) zp_bwd_kernel(const double* ybase, const ZpRec* recs, ZpFilt<ORD> t, int L, int W16, TOut* out) {
  __shared__ double os[ZP_THREADS][ZP_EB + 1];
  const ZpRec rc = recs[blockIdx.y];
  const int64_t N = rc.N, Next = N + 2 * t.pad, nch = rc.nch;
  const int64_t cb = (int64_t)blockIdx.x * ZP_THREADS;
  if (N <= 0 || cb >= nch) return;
  const int tid = threadIdx.x;
  const int64_t ch = cb + tid, c0 = ch * L;
  const double* yp = ybase + rc.y_off + ch;
  TOut* op = out + rc.out_off;
  double z[ORD];
#pragma unroll
  for (int i = 0; i < ORD; ++i) z[i] = 0.0;
  const int sg = tid >> 4, su = tid & 15;
  // every position of every chunk of this CTA (warm-up included) strictly below the right end, outputs inside [0, N)
  const bool inner = ((cb + ZP_THREADS) * L + W16 < Next - 1) && (cb * L - t.pad >= 0) && ((cb + ZP_THREADS) * L - 1 - t.pad <= N - 1);
  auto load_y = [&](int kb, double (&v)[ZP_EB]) {                   // own chunk, or (warm-up) the head of the next chunk
    const double* q = kb >= L ? yp + 1 + (int64_t)(kb - L) * nch : yp + (int64_t)kb * nch;
#pragma unroll
    for (int u = 0; u < ZP_EB; ++u) { v[u] = *q; q += nch; }
  };
  double pre[ZP_EB];
  if (inner) load_y(L + W16 - ZP_EB, pre);
  for (int kb = L + W16 - ZP_EB; kb >= 0; kb -= ZP_EB) {          // positions c0 + kb + 15 down to c0 + kb
    double yb[ZP_EB];
    if (inner) {
#pragma unroll
      for (int u = 0; u < ZP_EB; ++u) yb[u] = pre[u];
      if (kb >= ZP_EB) load_y(kb - ZP_EB, pre);                     // next block's values in flight during this block's math
#pragma unroll
      for (int u = ZP_EB - 1; u >= 0; --u) os[tid][u] = zp_step<ORD>(t, z, yb[u]);
    } else {
#pragma unroll
      for (int u = 0; u < ZP_EB; ++u) {
        const int k = kb + u;
        const int64_t e = c0 + k;
        const int64_t te = ch + (k >= L ? 1 : 0);
        const int ie = k >= L ? k - L : k;
        yb[u] = (ch < nch && e < Next) ? ybase[rc.y_off + (int64_t)ie * nch + te] : 0.0;
      }
      if (ch < nch) {
#pragma unroll
        for (int u = ZP_EB - 1; u >= 0; --u) {
          const int64_t e = c0 + kb + u;
          if (e < Next) {
            if (e == Next - 1) {
#pragma unroll
              for (int i = 0; i < ORD; ++i) z[i] = t.zi[i] * yb[u];  // lfilter_zi start-up at the true right end
            }
            os[tid][u] = zp_step<ORD>(t, z, yb[u]);
          }
        }
      }
    }
    __syncthreads();
    if (kb < L) {
      if (inner) {
        TOut* o = op + (cb * L + kb + (int64_t)sg * L + su - t.pad);
        const int64_t stride = 4 * (int64_t)L;
#pragma unroll
        for (int i = 0; i < ZP_EB; ++i) { *o = (TOut)os[sg + 4 * i][su]; o += stride; }
      } else {
        for (int idx = tid; idx < ZP_THREADS * ZP_EB; idx += ZP_THREADS) {
          const int seg = idx >> 4, u = idx & 15;
          const int64_t e = (cb + seg) * L + kb + u, n = e - t.pad;
          if (cb + seg < nch && n >= 0 && n < N) op[n] = (TOut)os[seg][u];
        }
      }
    }
    __syncthreads();
  }
}

// chunk length for a filter with warm-up w: >= 1024, >= w, multiple of 16
static inline int zp_chunk_len(int w) {
  int lmin = 1024;
  if (const char* e = getenv("FB_ZP_L")) lmin = std::max(256, atoi(e) / 16 * 16);   // tuning knob
  return std::max(lmin, (w + 15) / 16 * 16);
}
