// DPSK (v2) receive chain on sm_100a: replaces bpsk_demodulate / qpsk_demodulate (modem.py:68-135, 189-266).
//
// Three kernels write one decided bit stream per recording (MSB-first, 32-bit words stored big-endian
// so memory order == stream order); backend.cu then does the magic search and byte packing.
//
//   psk_main_kernel  interior symbols.  One CTA = one tile of T differential symbols of one recording.
//                    Samples are staged into shared memory de-interleaved by polyphase row
//                    X[j][c] = x[n0 + c*sps + j]; the composite zero-phase kernel is evaluated at symbol
//                    instants only, as a register-tiled polyphase FIR (S symbols x 4 taps per step,
//                    LDS.128 windows) plus a decimated first-order recursion per slow pole pair
//                    (block features -> decaying scan -> residue combine).  See fbdsp/design.py.
//   psk_edge_kernel  the first / last few hundred symbols of every recording (and whole short
//                    recordings): float64 step-by-step evaluation of the reference recurrences
//                    (scipy filtfilt: odd extension, lfilter_zi start-up, DF2T forward + backward;
//                    LO mix; second filtfilt) on a window.  One thread per window.
//
// Algorithmic HBM bytes per recording: N * sizeof(sample) read + raw bytes written (<= 0.6 %).
#include "common.cuh"
#include "psk_shared.cuh"

#include <algorithm>
#include <math.h>
#include <stdlib.h>

#define PM_CH 8              // columns (= symbols) per thread: FIR register tile and slow-pole scan chunk
#ifndef PM_THREADS
#define PM_THREADS 256
#endif
#ifndef PM_MINB
#define PM_MINB 2
#endif
#define PM_MAXTAB 3712       // float2 entries of the per-launch constant table (taps + slow-pole row weights)
#define SLOW_TBL 48          // per pole pair: 32 lane powers, 5 warp-scan multipliers, 9 warp powers (float2 each)

__global__ void __launch_bounds__(256) psk_tiles_kernel(const RecPlan* plans, const uint32_t* tile_first, int n_rec, uint32_t n_tiles,
                                                         int T, PskTile* tiles) {
  const uint32_t tile = blockIdx.x * blockDim.x + threadIdx.x;
  if (tile >= n_tiles) return;
  int lo = 0, hi = n_rec;                             // largest r with tile_first[r] <= tile
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (__ldg(&tile_first[mid]) <= tile) lo = mid; else hi = mid; }
  const RecPlan pl = plans[lo];
  PskTile t;
  t.off = pl.off; t.n = pl.n; t.word_off = pl.word_off;
  t.d0 = pl.dl32 + (int)(tile - __ldg(&tile_first[lo])) * T;
  t.d1 = min(t.d0 + T, pl.dr32);
  tiles[tile] = t;
}

// Per-launch arguments live in the kernel parameter space (constant bank 0): the FIR taps and the slow-pole row
// weights are read as FFMA2 uniform-register operands (LDCU.64 c[0x0][..]), so the inner loop has no tap loads
// through the LSU and nothing is shared between handles / streams.
struct PskMainArgs {
  const void* samples;
  const PskTile* tiles;         // one descriptor per CTA
  const uint32_t* redo;         // non-null: evaluate only tiles redo[1 .. redo[0]] (the ones psk_mma.cu could not take), CTAs stride over the list
  uint32_t n_tiles, pf_dist;    // pf_dist: the tile this many CTAs ahead gets its samples prefetched into L2
  // uniform batch (equal-length recordings back to back, e.g. the parts of one file): the descriptor is arithmetic on
  // these parameters -- no dependent global load at the top of the CTA.  uni_tpr == 0: read tiles[tile].
  uint32_t uni_tpr;             // tiles per recording
  int32_t uni_dl, uni_dr;       // dl32, dr32 of every recording
  uint64_t uni_off0, uni_n, uni_wstride;
  const float4* slow_pw4;       // [pairs][wpad]  {p_a^k, p_b^k}, zero for k > wlen: weights of the tile-boundary state sums (global)
  int wpad;                     // entries per pole pair in slow_pw4 (>= wlen + 1 + 1024, so over-reads hit zeros)
  const float2* slow_tbl;       // [nslow][SLOW_TBL]  powers of m = lam^PM_CH used by the column scan (global)
  uint32_t* bits;
  int n_rec;
  int sps, n0, bps, nt, ntp, dl, dh, nslow, wlen, pad_bp;
  int T;                        // tile size in differential symbols (multiple of 32)
  int P;                        // shared-memory row pitch in floats (multiple of 4)
  int padl;                     // columns staged left of (d0 - dh): makes the centre column 16-byte aligned
  int wc_off;                   // tab[wc_off + (i*sps + j)*2 + {0,1}] = {p_i^(sps-j), p_i^j}
  float2 rho;
  float2 lam[FB_MAX_SLOW];      // p^sps
  // R F + R' conj(F) as a real 2x2 map of (Re F, Im F): {a11, a12, a21, a22}; af = forward, ab = backward residues
  float4 af[FB_MAX_SLOW], ab[FB_MAX_SLOW];
  double slow_p[2 * FB_MAX_SLOW];
  float2 tab[PM_MAXTAB];        // [sps][ntp] reversed taps, then the slow-pole row weights
};

// 4 consecutive samples starting at element index i (i and the base pointer aligned to 4 elements)
template <typename T> __device__ __forceinline__ float4 load4(const void* base, uint64_t i);
template <> __device__ __forceinline__ float4 load4<float>(const void* base, uint64_t i) {
  return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + i));
}
template <> __device__ __forceinline__ float4 load4<double>(const void* base, uint64_t i) {
  const double2* p = reinterpret_cast<const double2*>(reinterpret_cast<const double*>(base) + i);
  const double2 a = __ldg(p), b = __ldg(p + 1);
  return make_float4((float)a.x, (float)a.y, (float)b.x, (float)b.y);
}
template <> __device__ __forceinline__ float4 load4<int16_t>(const void* base, uint64_t i) {
  const short4 v = __ldg(reinterpret_cast<const short4*>(reinterpret_cast<const int16_t*>(base) + i));
  const float k = 1.0f / 32768.0f;
  return make_float4((float)v.x * k, (float)v.y * k, (float)v.z * k, (float)v.w * k);
}

// two consecutive samples starting at an EVEN element index (8- / 4- / 16-byte aligned), as floats
template <typename T> __device__ __forceinline__ float2 load_pair(const T* p);
template <> __device__ __forceinline__ float2 load_pair<float>(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
template <> __device__ __forceinline__ float2 load_pair<double>(const double* p) {
  const double2 v = __ldg(reinterpret_cast<const double2*>(p));
  return make_float2((float)v.x, (float)v.y);
}
template <> __device__ __forceinline__ float2 load_pair<int16_t>(const int16_t* p) {
  const short2 v = __ldg(reinterpret_cast<const short2*>(p));
  return make_float2((float)v.x * (1.0f / 32768.0f), (float)v.y * (1.0f / 32768.0f));
}

// packed fp32 pairs (FFMA2 on sm_100a): a complex accumulator is one 64-bit register pair
__device__ __forceinline__ float2 bfma(float x, float2 w, float2 acc) {       // acc + x * w   (x broadcast)
  return __ffma2_rn(make_float2(x, x), w, acc);
}
__device__ __forceinline__ float2 cfma2(float2 a, float2 b, float2 c) {        // a*b + c  (complex)
  return bfma(b.y, make_float2(-a.y, a.x), bfma(b.x, a, c));
}
__device__ __forceinline__ float2 map22(float4 m, float fr, float fi, float2 acc) {   // acc + [m.x m.y; m.z m.w] (fr, fi)
  return bfma(fi, make_float2(m.y, m.w), bfma(fr, make_float2(m.x, m.z), acc));
}

// NT > 0: taps per polyphase row known at compile time (14 / 16 / 18: the window of NT + 7 columns lives in registers
// and every tap index is an immediate offset into the parameter bank); NT == 0: runtime nt (rows padded to ntp % 4 == 0).
// staged column c lives at X[j][pm_swz(c)]: bit 2 is flipped in odd 32-column blocks, so the 32-byte-strided LDS.128
// window loads of a quarter warp (8 symbols per thread) hit 8 distinct 16-byte bank groups instead of 4
__device__ __forceinline__ int pm_swz(int c) { return c ^ (((c >> 5) & 1) << 2); }

// SPS > 0: samples per symbol (and, with PP > 0, the shared-memory row pitch) known at compile time (even): the staging loop is fully unrolled with all
// of a thread's loads in flight before the first store; SPS == 0: runtime sps.
#ifdef PM_TRACE
// Phase timeline of the main kernel (experiments only; build with NVCC_EXTRA=-DPM_TRACE into a separate .so):
// thread 0 of the first PM_TRACE_N CTAs records %globaltimer at the phase boundaries and its SM id.
#define PM_TRACE_N 16384
__device__ unsigned long long pm_trace_t[PM_TRACE_N][12];
__device__ unsigned int pm_trace_sm[PM_TRACE_N];
__device__ __forceinline__ void pm_mark(int k) {
  if (threadIdx.x == 0 && blockIdx.x < PM_TRACE_N) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    pm_trace_t[blockIdx.x][k] = t;
    if (k == 0) { unsigned int sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); pm_trace_sm[blockIdx.x] = sm; }
  }
}
extern "C" int fb_debug_pm_trace(unsigned long long* t, unsigned int* sm) {
  if (cudaMemcpyFromSymbol(t, pm_trace_t, sizeof(pm_trace_t)) != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(sm, pm_trace_sm, sizeof(pm_trace_sm)) != cudaSuccess) return -1;
  return PM_TRACE_N;
}
#else
__device__ __forceinline__ void pm_mark(int) {}
#endif

template <typename TIn, int NT, int SPS, int PP, int NSL, bool REDO>
__global__ void __launch_bounds__(PM_THREADS, PM_MINB) psk_main_kernel(const __grid_constant__ PskMainArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float2 s_bnd[2][2][4][2];   // boundary-state partial sums [pair][dir][slice][pole of the pair]
  __shared__ float2 s_tot[8][4];         // warp totals of the column scan [warp][seq]
  __shared__ float2 s_y0[9];             // y of each warp's first symbol (differential across warp edges)
  __shared__ double finit_sh[2 * FB_MAX_SLOW];
  __shared__ float2 s_tbl[FB_MAX_SLOW * SLOW_TBL];   // scan multipliers (a.slow_tbl): read at LDS latency inside the dependent scan chain
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x, nwarp = nthr >> 5;
  const int sps = SPS ? SPS : a.sps;
  const int dh = NT ? NT / 2 : a.dh, dl = NT ? NT / 2 - 1 : a.dl;   // the host rejects designs with dh != nt / 2 or dl != nt / 2 - 1
  const int nslow = NSL ? NSL : a.nslow;           // NSL > 0: number of slow poles known at compile time (pair loops unroll)
  constexpr int PADL = NT ? (4 - (NT / 2) % 4) % 4 : 0;
  pm_mark(0);
  for (int i = threadIdx.x; i < nslow * SLOW_TBL; i += blockDim.x) s_tbl[i] = __ldg(&a.slow_tbl[i]);   // visible after the staging barrier
  // ---- this CTA's tile(s): blockIdx.x, or -- in redo mode -- a stride over the redo list ---------------------------------------
  // (REDO is a template flag so that the plain one-tile-per-CTA build keeps its register allocation)
  const uint32_t n_work = REDO ? __ldg(&a.redo[0]) : a.n_tiles;      // redo: [0] count, [1] scheduler counter of psk_mma.cu, [2..] tiles
  uint32_t work = blockIdx.x;
  if (REDO && work >= n_work) return;
  do {
  const uint32_t tile = REDO ? __ldg(&a.redo[2 + work]) : work;
  auto get_tile = [&](uint32_t t) {
    PskTile q;
    if (a.uni_tpr) {
      const uint32_t r = t / a.uni_tpr, i = t - r * a.uni_tpr;
      q.off = a.uni_off0 + (uint64_t)r * a.uni_n; q.n = a.uni_n; q.word_off = (uint64_t)r * a.uni_wstride;
      q.d0 = a.uni_dl + (int)i * a.T; q.d1 = min(q.d0 + a.T, a.uni_dr);
    } else {
      q = a.tiles[t];
    }
    return q;
  };
  const PskTile pl = get_tile(tile);
  const int d0 = pl.d0, d1 = pl.d1;
  const int ns = d1 - d0 + 1;                       // symbols d0 .. d1
  const int64_t N = (int64_t)pl.n;
  float* X = smem;                                  // [sps][P]   X[j][c] = x[n0 + (ca + c) sps + j]
  const int P = PP ? PP : a.P;                      // PP > 0: row pitch known at compile time (immediate store / load offsets)
  const int ca = d0 - dh - PADL;                  // global column (== symbol index) of staged column 0
  const int cc = dh + PADL;                       // staged column of symbol d0
  const int64_t n_d0 = (int64_t)a.n0 + (int64_t)d0 * sps;              // sample index of symbol d0
  const int64_t n_e1 = (int64_t)a.n0 + (int64_t)(d1 + 1) * sps;        // first sample after the tile's last column
  // ---- stage samples: thread <-> column (sps consecutive samples), conflict-free row stores ------------
  // A warp reads 32*sps consecutive samples; each 128-byte line is fetched from L2 once and re-hit in L1.
  {
    const int64_t n_a = (int64_t)a.n0 + (int64_t)ca * sps;          // sample index of staged element 0
    const int ncols = min(P, cc + ns + dl + 1);
    bool done = false;
    if constexpr (SPS > 0 && (SPS & 1) == 0) {
      // pair loads need an even element index: start one sample early when the column start is odd (the parity is
      // the same for every column).  Loaded element k of column c is X[k - s][c]; k - s == -1 is X[sps-1][c-1].
      const int s = (int)((pl.off + (uint64_t)n_a) & 1);
      if (n_a - s >= 0 && n_a + (int64_t)(ncols + 1) * SPS <= N) {      // whole staged range inside the recording
        const TIn* base = reinterpret_cast<const TIn*>(a.samples) + pl.off + n_a - s;
        if constexpr (SPS >= 40) {
          // long columns (1200 sym/s: 80 samples): one column per thread at a time, ten pair loads in flight per step
          constexpr int CH = 10;
          static_assert((SPS / 2) % CH == 0, "column length must be a multiple of 20 samples");
          for (int c = tid; c < ncols + s; c += nthr) {
            const TIn* src = base + (int64_t)c * SPS;
            float* dc = X + pm_swz(c);
#pragma unroll 1
            for (int i0 = 0; i0 < SPS / 2; i0 += CH) {
              float2 v[CH];
#pragma unroll
              for (int i = 0; i < CH; ++i) v[i] = load_pair<TIn>(src + 2 * (i0 + i));
              if (s) {
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                  const int e = 2 * (i0 + i);               // loaded elements e, e + 1 -> rows e - 1, e (e - 1 == -1: the previous column's last row)
                  if (e == 0) { if (c > 0) X[(SPS - 1) * P + pm_swz(c - 1)] = v[i].x; }
                  else if (c < ncols) dc[(e - 1) * P] = v[i].x;
                  if (c < ncols) dc[e * P] = v[i].y;
                }
              } else {
#pragma unroll
                for (int i = 0; i < CH; ++i) { dc[2 * (i0 + i) * P] = v[i].x; dc[(2 * (i0 + i) + 1) * P] = v[i].y; }
              }
            }
          }
        } else {
        constexpr int KC = 4;
        for (int c0 = tid; c0 < ncols + s; c0 += KC * nthr) {
          constexpr int HS = SPS > 1 ? SPS / 2 : 1;
          float2 v[KC][HS];
#pragma unroll
          for (int k = 0; k < KC; ++k) {
            const int c = c0 + k * nthr;
            if (c < ncols + s) {
              const TIn* src = base + (int64_t)c * SPS;
#pragma unroll
              for (int i = 0; i < SPS / 2; ++i) v[k][i] = load_pair<TIn>(src + 2 * i);
            }
          }
#pragma unroll
          for (int k = 0; k < KC; ++k) {
            const int c = c0 + k * nthr;
            if (c < ncols + s) {
              float* dc = X + pm_swz(c);
              if (s) {
                if (c > 0) X[(SPS - 1) * P + pm_swz(c - 1)] = v[k][0].x;
                if (c < ncols) {
                  dc[0] = v[k][0].y;
#pragma unroll
                  for (int i = 1; i < SPS / 2; ++i) { dc[(2 * i - 1) * P] = v[k][i].x; dc[2 * i * P] = v[k][i].y; }
                }
              } else {
#pragma unroll
                for (int i = 0; i < SPS / 2; ++i) { dc[2 * i * P] = v[k][i].x; dc[(2 * i + 1) * P] = v[k][i].y; }
              }
            }
          }
        }
        }
        done = true;
      }
    }
    if (!done) {
      for (int c = tid; c < ncols; c += nthr) {
        const int64_t n = n_a + (int64_t)c * sps;
        float* dst = X + pm_swz(c);
        if (n >= 0 && n + sps <= N) {
          const uint64_t g = pl.off + (uint64_t)n;
          for (int j = 0; j < sps; ++j) dst[j * P] = load_sample<TIn>(a.samples, g + j);
        } else {
          for (int j = 0; j < sps; ++j)
            dst[j * P] = (n + j >= 0 && n + j < N) ? load_sample<TIn>(a.samples, pl.off + (uint64_t)(n + j)) : 0.f;
        }
      }
    }
  }
  pm_mark(1);
  // ---- L2 prefetch for the CTA that will run about one tile-time from now (same SM slot, pf_dist tiles ahead): its
  // staging and boundary loads then hit L2 instead of paying the DRAM latency
  if (tile + a.pf_dist < a.n_tiles) {
    const PskTile nx = get_tile(tile + a.pf_dist);
    const int64_t p0 = max((int64_t)0, (int64_t)a.n0 + (int64_t)(nx.d0 - dh - PADL) * sps - a.wlen);
    const int64_t p1 = min((int64_t)nx.n, (int64_t)a.n0 + (int64_t)(nx.d1 + 2 + dl) * sps + a.wlen);
    const char* base = reinterpret_cast<const char*>(a.samples) + (nx.off + (uint64_t)p0) * sizeof(TIn);
    const int64_t nbytes = (p1 - p0) * (int64_t)sizeof(TIn);
    for (int64_t b = (int64_t)tid * 128; b < nbytes; b += (int64_t)nthr * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(base + b));
  }
  // ---- exact start state of the forward slow recursion at column 0 (left record edge) ----------
  const bool near_left = (n_d0 - a.wlen) <= (int64_t)a.n0;          // the boundary sum would reach column 0
  if (near_left && tid < nslow) {
    // Fst[0] = sum_{n < n0} p^(n0-n) xL[n],  xL = scipy's odd extension (pad_bp samples) then the constant
    // xL[-pad_bp] for ever (that is what the lfilter_zi start-up of filtfilt's forward pass stands for).
    const double pr = a.slow_p[2 * tid], pi = a.slow_p[2 * tid + 1];
    const double x0 = load_sample_d<TIn>(a.samples, pl.off);
    double cr = 2.0 * x0 - load_sample_d<TIn>(a.samples, pl.off + a.pad_bp);     // xL[-pad]
    // s = c0 / (1 - p): state (sum_{k>=0} p^k c0) just before n = -pad
    const double den = (1.0 - pr) * (1.0 - pr) + pi * pi;
    double sr = cr * (1.0 - pr) / den, si = cr * pi / den;
    for (int n = -a.pad_bp; n < a.n0; ++n) {       // s <- p*s + xL[n]
      double xv = (n < 0) ? 2.0 * x0 - load_sample_d<TIn>(a.samples, pl.off + (uint64_t)(-n))
                          : load_sample_d<TIn>(a.samples, pl.off + (uint64_t)n);
      double tr = pr * sr - pi * si + xv, ti = pr * si + pi * sr;
      sr = tr; si = ti;
    }
    finit_sh[2 * tid] = pr * sr - pi * si;         // Fst = p * s
    finit_sh[2 * tid + 1] = pr * si + pi * sr;
  }

  // ---- tile-boundary states of the slow recursions (all pole pairs): Fst[d0] and Bfull[d1+1] as direct sums over the
  // previous / next `wlen` samples against the power table {p_a^k, p_b^k} (zero beyond wlen).  2 directions x 4 slices
  // of 128 samples per step, one warp per (direction, slice) job.
  for (int pair = 0; pair < nslow; pair += 2) {
    const float4* pw = a.slow_pw4 + (size_t)(pair >> 1) * a.wpad;
    for (int job = warp; job < 8; job += nwarp) {
      const bool fwd = job < 4;
      const int slice = job & 3;
      // forward: sample n_d0 - k has weight p^k, k = 1 .. cntf;  backward: sample n_e1 + k has weight p^k, k = 0 .. cntb-1
      const int cntf = (int)min((int64_t)a.wlen, n_d0 - (near_left ? (int64_t)a.n0 : (int64_t)0));
      const int cntb = (int)max((int64_t)0, min((int64_t)a.wlen + 1, N - n_e1));
      const int klo = fwd ? 1 : 0, khi = fwd ? cntf : cntb - 1;
      float2 bs0 = make_float2(0.f, 0.f), bs1 = make_float2(0.f, 0.f);
      const TIn* xb = reinterpret_cast<const TIn*>(a.samples) + pl.off + (fwd ? n_d0 : n_e1);
      // three 512-sample steps per trip with all their loads in flight together (wlen is ~1200-1800 samples)
      for (int k00 = klo + 4 * (slice * 32 + lane); k00 <= khi; k00 += 3 * 512) {
        float xv[3][4];
        float4 w[3][4];
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int kk = k00 + 512 * q + u;
            const int k = min(kk, khi);                           // clamped address; the duplicate is zeroed below
            xv[q][u] = load_sample<TIn>(xb, (uint64_t)(int64_t)(fwd ? -k : k));       // 32-bit offset from the tile edge
            w[q][u] = __ldg(&pw[min(kk, a.wpad - 1)]);
          }
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float x = (k00 + 512 * q + u <= khi) ? xv[q][u] : 0.f;
            bs0 = bfma(x, make_float2(w[q][u].x, w[q][u].y), bs0);
            bs1 = bfma(x, make_float2(w[q][u].z, w[q][u].w), bs1);
          }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        bs0.x += __shfl_xor_sync(0xffffffffu, bs0.x, off); bs0.y += __shfl_xor_sync(0xffffffffu, bs0.y, off);
        bs1.x += __shfl_xor_sync(0xffffffffu, bs1.x, off); bs1.y += __shfl_xor_sync(0xffffffffu, bs1.y, off);
      }
      if (lane == 0) { s_bnd[pair >> 1][fwd ? 0 : 1][slice][0] = bs0; s_bnd[pair >> 1][fwd ? 0 : 1][slice][1] = bs1; }
    }
  }
  pm_mark(2);
  __syncthreads();                                  // staged samples, finit and the boundary sums are visible
  pm_mark(3);

  // this thread's PM_CH symbols e0 .. e0+7 (global d0 + e); y accumulates the slow part, then the FIR
  const int e0 = tid * PM_CH;
  const bool active = e0 < ns;
  float2 y[PM_CH];
#pragma unroll
  for (int i = 0; i < PM_CH; ++i) y[i] = make_float2(0.f, 0.f);

  // ---- slow pole pairs ------------------------------------------------------------------------------
  // y_slow[c] = sum_i  R+_i Fst_i[c] + R+'_i conj(Fst_i[c]) + R-_i Bst_i[c] + R-'_i conj(Bst_i[c])
  //   Fst[c]   = sum_{n < n_c} p^(n_c - n) x[n]      Fst[c+1] = lam Fst[c] + sum_j p^(sps-j) X[j][c]
  //   Bfull[c] = sum_{n >= n_c} p^(n - n_c) x[n]     Bfull[c] = sum_j p^j X[j][c] + lam Bfull[c+1];  Bst = Bfull - x[n_c]
  // The states at the tile boundaries (Fst[d0], Bfull[d1+1]) are direct sums over the previous / next `wlen`
  // samples against the power table p^k (read straight from global memory: no halo staging, no inter-CTA
  // dependency); inside the tile the recursion runs on per-column features z (FFMA2, weights from the parameter
  // bank), a thread-local fold, a decaying warp-shuffle scan (up for the forward states, down for the backward
  // ones) and a serial carry over the <= 8 warps.
  for (int pair = 0; pair < nslow; pair += 2) {
    const bool two = pair + 1 < nslow;
    const int i1 = two ? pair + 1 : pair;
    const float2 lam0 = a.lam[pair], lam1 = a.lam[i1];
    // (2) per-column features of this thread's columns: z[0],z[1] forward (pole 0,1), z[2],z[3] backward
    float2 z[4][PM_CH];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int i = 0; i < PM_CH; ++i) z[k][i] = make_float2(0.f, 0.f);
    if (active) {
      const float* rf = X;
      const int fo0 = pm_swz(cc + e0), fo1 = pm_swz(cc + e0 + 4);
      const float2* w0p = a.tab + a.wc_off + (size_t)pair * sps * 2;
      const float2* w1p = a.tab + a.wc_off + (size_t)i1 * sps * 2;
      for (int j = 0; j < sps; ++j) {
        float xs[PM_CH];
        if (NT) {
          const float4 u0 = *reinterpret_cast<const float4*>(rf + fo0), u1 = *reinterpret_cast<const float4*>(rf + fo1);
          xs[0] = u0.x; xs[1] = u0.y; xs[2] = u0.z; xs[3] = u0.w; xs[4] = u1.x; xs[5] = u1.y; xs[6] = u1.z; xs[7] = u1.w;
        } else {
#pragma unroll
          for (int i = 0; i < PM_CH; ++i) xs[i] = rf[pm_swz(cc + e0 + i)];
        }
        const float2 wf0 = w0p[2 * j], wb0 = w0p[2 * j + 1], wf1 = w1p[2 * j], wb1 = w1p[2 * j + 1];
#pragma unroll
        for (int i = 0; i < PM_CH; ++i) {
          z[0][i] = bfma(xs[i], wf0, z[0][i]);
          z[1][i] = bfma(xs[i], wf1, z[1][i]);
          z[2][i] = bfma(xs[i], wb0, z[2][i]);
          z[3][i] = bfma(xs[i], wb1, z[3][i]);
        }
        rf += P;
      }
      // columns past the tile contribute nothing; the backward boundary state enters at the last column:
      // Bfull[ns-1] = zb[ns-1] + lam Bfull[ns]
      const float2 bb0 = make_float2(s_bnd[pair >> 1][1][0][0].x + s_bnd[pair >> 1][1][1][0].x + s_bnd[pair >> 1][1][2][0].x + s_bnd[pair >> 1][1][3][0].x,
                                     s_bnd[pair >> 1][1][0][0].y + s_bnd[pair >> 1][1][1][0].y + s_bnd[pair >> 1][1][2][0].y + s_bnd[pair >> 1][1][3][0].y);
      const float2 bb1 = make_float2(s_bnd[pair >> 1][1][0][1].x + s_bnd[pair >> 1][1][1][1].x + s_bnd[pair >> 1][1][2][1].x + s_bnd[pair >> 1][1][3][1].x,
                                     s_bnd[pair >> 1][1][0][1].y + s_bnd[pair >> 1][1][1][1].y + s_bnd[pair >> 1][1][2][1].y + s_bnd[pair >> 1][1][3][1].y);
      const float2 inj0 = cfma2(lam0, bb0, make_float2(0.f, 0.f)), inj1 = cfma2(lam1, bb1, make_float2(0.f, 0.f));
#pragma unroll
      for (int i = 0; i < PM_CH; ++i) {
        if (e0 + i >= ns) {
#pragma unroll
          for (int k = 0; k < 4; ++k) z[k][i] = make_float2(0.f, 0.f);
        }
        if (e0 + i == ns - 1) {
          z[2][i].x += inj0.x; z[2][i].y += inj0.y;
          z[3][i].x += inj1.x; z[3][i].y += inj1.y;
        }
      }
    }
    pm_mark(7);
    // (3) thread totals -> warp scan (shuffle up: forward, shuffle down: backward) -> serial carry over warps
    const float2* tb0 = s_tbl + pair * SLOW_TBL;
    const float2* tb1 = s_tbl + i1 * SLOW_TBL;
    float2 v[4];
    {
      float2 s0 = make_float2(0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
#pragma unroll
      for (int i = 0; i < PM_CH; ++i) {
        s0 = cfma2(lam0, s0, z[0][i]);
        s1 = cfma2(lam1, s1, z[1][i]);
        s2 = cfma2(lam0, s2, z[2][PM_CH - 1 - i]);
        s3 = cfma2(lam1, s3, z[3][PM_CH - 1 - i]);
      }
      v[0] = s0; v[1] = s1; v[2] = s2; v[3] = s3;
    }
#pragma unroll
    for (int st = 0; st < 5; ++st) {
      const float2 m0 = tb0[32 + st], m1 = tb1[32 + st];   // m^(2^st), m = lam^PM_CH
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 mm = (k & 1) ? m1 : m0;
        if (k < 2) {
          const float ox = __shfl_up_sync(0xffffffffu, v[k].x, 1 << st), oy = __shfl_up_sync(0xffffffffu, v[k].y, 1 << st);
          if (lane >= (1 << st)) v[k] = cfma2(mm, make_float2(ox, oy), v[k]);
        } else {
          const float ox = __shfl_down_sync(0xffffffffu, v[k].x, 1 << st), oy = __shfl_down_sync(0xffffffffu, v[k].y, 1 << st);
          if (lane + (1 << st) < 32) v[k] = cfma2(mm, make_float2(ox, oy), v[k]);
        }
      }
    }
    pm_mark(8);
    if (lane == 31) { s_tot[warp][0] = v[0]; s_tot[warp][1] = v[1]; }
    if (lane == 0) { s_tot[warp][2] = v[2]; s_tot[warp][3] = v[3]; }
    __syncthreads();
    // state entering this warp: lanes 0-3 (k = lane) fold the totals of the warps before (forward) / after (backward) it,
    // every warp for itself -- no serial section behind a second CTA barrier
    float2 car[4];
    {
      const int k = lane & 3;
      const float2* tb = (k & 1) ? tb1 : tb0;
      const float2 M = tb[38];                        // m^32
      float2 c = make_float2(0.f, 0.f);
      if (lane < 4) {
        if (k < 2) {                                  // from the left; warp 0: Fst[d0]
          c = make_float2(s_bnd[pair >> 1][0][0][k].x + s_bnd[pair >> 1][0][1][k].x + s_bnd[pair >> 1][0][2][k].x + s_bnd[pair >> 1][0][3][k].x,
                          s_bnd[pair >> 1][0][0][k].y + s_bnd[pair >> 1][0][1][k].y + s_bnd[pair >> 1][0][2][k].y + s_bnd[pair >> 1][0][3][k].y);
          if (near_left) {                            // + p^(n_d0 - n0) Fst[0]
            const int i = (k == 0) ? pair : i1;
            const float4 pw4 = __ldg(&a.slow_pw4[(size_t)(pair >> 1) * a.wpad + (int)(n_d0 - a.n0)]);
            const float2 pw = (k == 0) ? make_float2(pw4.x, pw4.y) : make_float2(pw4.z, pw4.w);
            c = cfma2(pw, make_float2((float)finit_sh[2 * i], (float)finit_sh[2 * i + 1]), c);
          }
          for (int w = 0; w < warp; ++w) c = cfma2(M, c, s_tot[w][k]);
        } else {                                      // from the right (boundary already injected)
          for (int w = nwarp - 1; w > warp; --w) c = cfma2(M, c, s_tot[w][k]);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) car[q] = make_float2(__shfl_sync(0xffffffffu, c.x, q), __shfl_sync(0xffffffffu, c.y, q));
    }
    pm_mark(9);
    // (4) states entering this thread's chunk, then the per-column recursion and the residue maps
    {
      float2 sc[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2* tb = (k & 1) ? tb1 : tb0;
        float ex, ey;
        if (k < 2) { ex = __shfl_up_sync(0xffffffffu, v[k].x, 1); ey = __shfl_up_sync(0xffffffffu, v[k].y, 1); }
        else { ex = __shfl_down_sync(0xffffffffu, v[k].x, 1); ey = __shfl_down_sync(0xffffffffu, v[k].y, 1); }
        const bool has = (k < 2) ? lane > 0 : lane < 31;
        const float2 excl = has ? make_float2(ex, ey) : make_float2(0.f, 0.f);
        sc[k] = cfma2(tb[(k < 2) ? lane : 31 - lane], car[k], excl);
      }
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 af0 = a.af[pair], af1 = two ? a.af[i1] : z4, ab0 = a.ab[pair], ab1 = two ? a.ab[i1] : z4;
#pragma unroll
      for (int i = 0; i < PM_CH; ++i) {               // forward: Fst[d0 + e0 + i] is the state BEFORE column i
        y[i] = map22(af0, sc[0].x, sc[0].y, y[i]);
        y[i] = map22(af1, sc[1].x, sc[1].y, y[i]);
        sc[0] = cfma2(lam0, sc[0], z[0][i]);
        sc[1] = cfma2(lam1, sc[1], z[1][i]);
      }
      float x0c[PM_CH];                               // x[n_c] of the thread's columns (row 0)
#pragma unroll
      for (int i = 0; i < PM_CH; ++i) x0c[i] = 0.f;
      if (active) {
        if (NT) {                                     // cc + e0 is a multiple of 4: two aligned 16-byte loads
          const float4 u0 = *reinterpret_cast<const float4*>(X + pm_swz(cc + e0)), u1 = *reinterpret_cast<const float4*>(X + pm_swz(cc + e0 + 4));
          x0c[0] = u0.x; x0c[1] = u0.y; x0c[2] = u0.z; x0c[3] = u0.w; x0c[4] = u1.x; x0c[5] = u1.y; x0c[6] = u1.z; x0c[7] = u1.w;
        } else {
#pragma unroll
          for (int i = 0; i < PM_CH; ++i) x0c[i] = X[pm_swz(cc + e0 + i)];
        }
      }
#pragma unroll
      for (int i = PM_CH - 1; i >= 0; --i) {          // backward: Bfull[col] includes the column; Bst = Bfull - x[n_col]
        sc[2] = cfma2(lam0, sc[2], z[2][i]);
        sc[3] = cfma2(lam1, sc[3], z[3][i]);
        y[i] = map22(ab0, sc[2].x - x0c[i], sc[2].y, y[i]);
        y[i] = map22(ab1, sc[3].x - x0c[i], sc[3].y, y[i]);
      }
    }
    if (pair + 2 < nslow) __syncthreads();          // s_tot / s_car are rewritten by the next pair
  }

  pm_mark(4);
  // ---- fast part: register-tiled polyphase FIR at symbol instants --------------------------------
  //   y[s] += tapsR[j][t'] * X[j][e0 + PADL + s + t']   -- one FFMA2 per (complex tap, real sample), taps as uniform operands
  if (active) {
    const float* row = X;
    if (NT) {
      constexpr int NW = (PADL + NT + PM_CH - 1 + 3) / 4;      // float4 loads covering the window
      const float2* tp = a.tab;
      int wo[NW];
#pragma unroll
      for (int q = 0; q < NW; ++q) wo[q] = pm_swz(e0 + 4 * q);
      for (int j = 0; j < sps; ++j) {
        float win[4 * NW];
#pragma unroll
        for (int q = 0; q < NW; ++q) {
          const float4 u = *reinterpret_cast<const float4*>(row + wo[q]);
          win[4 * q] = u.x; win[4 * q + 1] = u.y; win[4 * q + 2] = u.z; win[4 * q + 3] = u.w;
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const float2 w = tp[t];
#pragma unroll
          for (int s = 0; s < PM_CH; ++s) y[s] = bfma(win[PADL + s + t], w, y[s]);
        }
        row += P;
        tp += NT;
      }
    } else {
      const float2* tp = a.tab;
      for (int j = 0; j < sps; ++j) {
        float win[PM_CH + 4];
        {
          const float4 u0 = *reinterpret_cast<const float4*>(row + pm_swz(e0)), u1 = *reinterpret_cast<const float4*>(row + pm_swz(e0 + 4));
          win[0] = u0.x; win[1] = u0.y; win[2] = u0.z; win[3] = u0.w; win[4] = u1.x; win[5] = u1.y; win[6] = u1.z; win[7] = u1.w;
        }
        for (int t4 = 0; t4 < a.ntp; t4 += 4) {
          const float4 u = *reinterpret_cast<const float4*>(row + pm_swz(e0 + 8 + t4));
          win[8] = u.x; win[9] = u.y; win[10] = u.z; win[11] = u.w;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 w = tp[t4 + t];
#pragma unroll
            for (int s = 0; s < PM_CH; ++s) y[s] = bfma(win[s + t], w, y[s]);
          }
#pragma unroll
          for (int s = 0; s < PM_CH; ++s) win[s] = win[s + 4];
        }
        row += P;
        tp += a.ntp;
      }
    }
  }

  pm_mark(5);
  // ---- differential decisions: PM_CH per thread, 32-bit big-endian words assembled across 2 (DQPSK) / 4 (DBPSK) lanes --
  {
    // warp w only needs the first symbol of warp w + 1: a producer / consumer pair on named barrier w + 1 (64 threads:
    // the arriving warp and the waiting one) instead of a CTA-wide barrier
    if (lane == 0) s_y0[warp] = y[0];
    __syncwarp();
    if (warp > 0) asm volatile("bar.arrive %0, 64;" ::"r"(warp) : "memory");
    float nx = __shfl_down_sync(0xffffffffu, y[0].x, 1), nyv = __shfl_down_sync(0xffffffffu, y[0].y, 1);
    if (warp + 1 < nwarp) {
      asm volatile("bar.sync %0, 64;" ::"r"(warp + 1) : "memory");
      if (lane == 31) { nx = s_y0[warp + 1].x; nyv = s_y0[warp + 1].y; }
    }
    uint32_t part = 0;
    if (a.bps == 2) {                                 // slicer specialised outside the unrolled loop (uniform branch)
#pragma unroll
      for (int i = 0; i < PM_CH; ++i) {
        const float2 prev = y[i];
        const float2 cur = (i + 1 < PM_CH) ? y[(i + 1) % PM_CH] : make_float2(nx, nyv);
        // d = cur * conj(prev) * rho
        const float tr = fmaf(cur.x, prev.x, cur.y * prev.y), ti = fmaf(cur.y, prev.x, -cur.x * prev.y);
        const float dr = fmaf(tr, a.rho.x, -ti * a.rho.y), di = fmaf(tr, a.rho.y, ti * a.rho.x);
        part = (part << 2) | psk_decide<float>(dr, di, 2);
      }
    } else {
#pragma unroll
      for (int i = 0; i < PM_CH; ++i) {
        const float2 prev = y[i];
        const float2 cur = (i + 1 < PM_CH) ? y[(i + 1) % PM_CH] : make_float2(nx, nyv);
        const float tr = fmaf(cur.x, prev.x, cur.y * prev.y), ti = fmaf(cur.y, prev.x, -cur.x * prev.y);
        const float dr = fmaf(tr, a.rho.x, -ti * a.rho.y);
        part = (part << 1) | (dr < 0.f ? 1u : 0u);
      }
    }
    const int nd = d1 - d0;                           // multiple of 32
    if (a.bps == 2) {                                 // 16 bits per thread, 2 threads per word
      const uint32_t other = __shfl_down_sync(0xffffffffu, part, 1);
      if ((lane & 1) == 0 && e0 < nd)
        a.bits[pl.word_off + (uint64_t)((d0 + e0) >> 4)] = __byte_perm((part << 16) | other, 0, 0x0123);
    } else {                                          // 8 bits per thread, 4 threads per word
      uint32_t wv = part << (24 - 8 * (lane & 3));
      wv |= __shfl_xor_sync(0xffffffffu, wv, 1);
      wv |= __shfl_xor_sync(0xffffffffu, wv, 2);
      if ((lane & 3) == 0 && e0 < nd) a.bits[pl.word_off + (uint64_t)((d0 + e0) >> 5)] = __byte_perm(wv, 0, 0x0123);
    }
  }
  pm_mark(6);
  if (REDO) __syncthreads();                            // the next tile of the list reuses the shared arrays
  } while (REDO && (work += gridDim.x) < n_work);
}

// =====================================================================================================
// Edge kernel: the reference recurrences, step by step, in float64, on a window.
// =====================================================================================================
struct PskEdgeArgs {
  const void* samples;
  const RecPlan* plans;
  const EdgeJob* jobs;
  double* scratch;
  uint32_t* bits;
  int n_jobs;
  fb_psk_design d;
};

template <typename TIn>
__device__ __forceinline__ double x_ext(const void* samples, uint64_t off, int64_t N, int64_t n) {
  // scipy.signal._arraytools.odd_ext: 2*x[0] - x[-n] on the left, 2*x[N-1] - x[2(N-1)-n] on the right
  if (n < 0) return 2.0 * load_sample_d<TIn>(samples, off) - load_sample_d<TIn>(samples, off + (uint64_t)(-n));
  if (n > N - 1)
    return 2.0 * load_sample_d<TIn>(samples, off + (uint64_t)(N - 1)) -
           load_sample_d<TIn>(samples, off + (uint64_t)(2 * (N - 1) - n));
  return load_sample_d<TIn>(samples, off + (uint64_t)n);
}

template <typename TIn>
__global__ void __launch_bounds__(32) psk_edge_kernel(const PskEdgeArgs a) {
  const int jid = blockIdx.x * blockDim.x + threadIdx.x;
  if (jid >= a.n_jobs) return;
  const EdgeJob jb = a.jobs[jid];
  const RecPlan pl = a.plans[jb.rec];
  const fb_psk_design& d = a.d;
  const int64_t N = (int64_t)pl.n;
  double* A = a.scratch + jb.scratch_off;                    // band-pass window, in place
  const int64_t Lb = jb.wb - jb.wa + 1;
  double* U = A + Lb;                                        // low-pass window, complex interleaved, in place
  const int64_t Ll = jb.lb - jb.la + 1;

  // All four passes run the recurrence on blocks of EB values fetched up front, so the (independent) loads of a
  // block are in flight together instead of one dependent DRAM/L2 round trip per step.
  constexpr int EB = 16;
  // ---- band-pass forward (scipy lfilter, direct form II transposed; a[0] == 1) ------------------
  {
    double z[8];
    const double x0 = x_ext<TIn>(a.samples, pl.off, N, jb.wa);
    const bool exact = (jb.wa == -(int64_t)d.pad_bp);
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = exact ? d.bp_zi[i] * x0 : 0.0;
    for (int64_t i0 = 0; i0 < Lb; i0 += EB) {
      double xb[EB];
#pragma unroll
      for (int u = 0; u < EB; ++u) xb[u] = (i0 + u < Lb) ? x_ext<TIn>(a.samples, pl.off, N, jb.wa + i0 + u) : 0.0;
#pragma unroll
      for (int u = 0; u < EB; ++u) {
        const double xv = xb[u];
        const double y = d.bp_b[0] * xv + z[0];
#pragma unroll
        for (int k = 0; k < 7; ++k) z[k] = d.bp_b[k + 1] * xv + z[k + 1] - d.bp_a[k + 1] * y;
        z[7] = d.bp_b[8] * xv - d.bp_a[8] * y;
        if (i0 + u < Lb) A[i0 + u] = y;
      }
    }
  }
  // ---- band-pass backward, in place ----------------------------------------------------------------
  {
    double z[8];
    const bool exact = (jb.wb == N - 1 + (int64_t)d.pad_bp);
    const double y0 = A[Lb - 1];
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = exact ? d.bp_zi[i] * y0 : 0.0;
    for (int64_t i0 = Lb - 1; i0 >= 0; i0 -= EB) {
      double xb[EB];
#pragma unroll
      for (int u = 0; u < EB; ++u) xb[u] = (i0 - u >= 0) ? A[i0 - u] : 0.0;
#pragma unroll
      for (int u = 0; u < EB; ++u) {
        if (i0 - u >= 0) {
          const double xv = xb[u];
          const double y = d.bp_b[0] * xv + z[0];
#pragma unroll
          for (int k = 0; k < 7; ++k) z[k] = d.bp_b[k + 1] * xv + z[k + 1] - d.bp_a[k + 1] * y;
          z[7] = d.bp_b[8] * xv - d.bp_a[8] * y;
          A[i0 - u] = y;
        }
      }
    }
  }
  // ---- mix with the continuous LO: u[n] = f[n] exp(-j 2 pi fc n / fs)  (modem.py:80-83, 200-201) --
  auto mixed = [&](int64_t n, double& ur, double& ui) {
    const double f = A[n - jb.wa];
    double ph = d.cycles_per_sample * (double)n;
    ph -= floor(ph);
    double s, c;
    sincospi(2.0 * ph, &s, &c);
    ur = f * c; ui = -f * s;
  };
  auto mixed_ext = [&](int64_t m, double& ur, double& ui) {   // odd extension of the mixed record
    if (m < 0) {
      double ar, ai, br, bi;
      mixed(0, ar, ai); mixed(-m, br, bi);
      ur = 2.0 * ar - br; ui = 2.0 * ai - bi;
    } else if (m > N - 1) {
      double ar, ai, br, bi;
      mixed(N - 1, ar, ai); mixed(2 * (N - 1) - m, br, bi);
      ur = 2.0 * ar - br; ui = 2.0 * ai - bi;
    } else {
      mixed(m, ur, ui);
    }
  };
  // ---- low-pass forward on the complex record ----------------------------------------------------------
  {
    constexpr int LB = 8;
    double zr[4], zi[4];
    double x0r, x0i;
    mixed_ext(jb.la, x0r, x0i);
    const bool exact = (jb.la == -(int64_t)d.pad_lp);
#pragma unroll
    for (int i = 0; i < 4; ++i) { zr[i] = exact ? d.lp_zi[i] * x0r : 0.0; zi[i] = exact ? d.lp_zi[i] * x0i : 0.0; }
    for (int64_t i0 = 0; i0 < Ll; i0 += LB) {
      double xr[LB], xi[LB];
#pragma unroll
      for (int u = 0; u < LB; ++u) {
        xr[u] = 0.0; xi[u] = 0.0;
        if (i0 + u < Ll) mixed_ext(jb.la + i0 + u, xr[u], xi[u]);
      }
#pragma unroll
      for (int u = 0; u < LB; ++u) {
        const double yr = d.lp_b[0] * xr[u] + zr[0], yi = d.lp_b[0] * xi[u] + zi[0];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          zr[k] = d.lp_b[k + 1] * xr[u] + zr[k + 1] - d.lp_a[k + 1] * yr;
          zi[k] = d.lp_b[k + 1] * xi[u] + zi[k + 1] - d.lp_a[k + 1] * yi;
        }
        zr[3] = d.lp_b[4] * xr[u] - d.lp_a[4] * yr;
        zi[3] = d.lp_b[4] * xi[u] - d.lp_a[4] * yi;
        if (i0 + u < Ll) { U[2 * (i0 + u)] = yr; U[2 * (i0 + u) + 1] = yi; }
      }
    }
  }
  // ---- low-pass backward, in place ---------------------------------------------------------------------
  {
    constexpr int LB = 8;
    double zr[4], zi[4];
    const bool exact = (jb.lb == N - 1 + (int64_t)d.pad_lp);
    const double y0r = U[2 * (Ll - 1)], y0i = U[2 * (Ll - 1) + 1];
#pragma unroll
    for (int i = 0; i < 4; ++i) { zr[i] = exact ? d.lp_zi[i] * y0r : 0.0; zi[i] = exact ? d.lp_zi[i] * y0i : 0.0; }
    for (int64_t i0 = Ll - 1; i0 >= 0; i0 -= LB) {
      double xr[LB], xi[LB];
#pragma unroll
      for (int u = 0; u < LB; ++u) {
        xr[u] = (i0 - u >= 0) ? U[2 * (i0 - u)] : 0.0;
        xi[u] = (i0 - u >= 0) ? U[2 * (i0 - u) + 1] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < LB; ++u) {
        if (i0 - u >= 0) {
          const double yr = d.lp_b[0] * xr[u] + zr[0], yi = d.lp_b[0] * xi[u] + zi[0];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            zr[k] = d.lp_b[k + 1] * xr[u] + zr[k + 1] - d.lp_a[k + 1] * yr;
            zi[k] = d.lp_b[k + 1] * xi[u] + zi[k + 1] - d.lp_a[k + 1] * yi;
          }
          zr[3] = d.lp_b[4] * xr[u] - d.lp_a[4] * yr;
          zi[3] = d.lp_b[4] * xi[u] - d.lp_a[4] * yi;
          U[2 * (i0 - u)] = yr; U[2 * (i0 - u) + 1] = yi;
        }
      }
    }
  }
  // ---- symbols, differential decisions, packed words ------------------------------------------------------
  {
    const int bps = d.bits_per_sym;
    const int dper = 32 / bps;
    uint32_t* wout = a.bits + pl.word_off + (uint64_t)(jb.k_lo / dper);
    uint32_t word = 0;
    int filled = 0;
    double pr = 0.0, pi = 0.0;
    for (int k = jb.k_lo; k <= jb.k_hi; ++k) {
      const int64_t nk = (int64_t)d.n0 + (int64_t)k * d.sps;
      const double sr = U[2 * (nk - jb.la)], si = U[2 * (nk - jb.la) + 1];
      if (k > jb.k_lo) {
        // s[k] conj(s[k-1]) -- both already carry the LO phase, as in the reference (modem.py:100, 214).  Evaluated the way
        // numpy's complex multiply is on every FMA-capable host (one rounded product, then a fused multiply-add; b = conj(s[k-1])
        // = (pr, -pi)): the rounding only matters where the products are denormal or underflow to signed zeros.
        const double dr = __fma_rn(sr, pr, __dmul_rn(si, pi)), di = __fma_rn(sr, -pi, __dmul_rn(si, pr));
        word = (word << bps) | psk_decide<double>(dr, di, bps);
        if (++filled == dper) {
          *wout++ = __byte_perm(word, 0, 0x0123);
          word = 0; filled = 0;
        }
      }
      pr = sr; pi = si;
    }
    if (filled) {
      word <<= (dper - filled) * bps;
      *wout = __byte_perm(word, 0, 0x0123);
    }
  }
}

// =====================================================================================================
// Host side: plan + launches
// =====================================================================================================
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t fdiv64(int64_t a, int64_t b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

extern "C" uint64_t fb_psk_out_bound(const fb_psk_design* d, uint64_t n_samples) {
  if (!d || d->sps < 1) return 0;
  const int64_t N = (int64_t)n_samples;
  const int64_t nsym = N > d->n0 ? cdiv64(N - d->n0, d->sps) : 0;
  const int64_t nbits = std::max<int64_t>(nsym - 1, 0) * d->bits_per_sym;
  return (uint64_t)(nbits / 8);
}

static void make_job(const fb_psk_design& d, int rec, int64_t N, int k_lo, int k_hi, uint64_t& scratch_doubles,
                     std::vector<EdgeJob>& jobs) {
  EdgeJob j{};
  j.rec = rec; j.k_lo = k_lo; j.k_hi = k_hi;
  const int64_t n_lo = (int64_t)d.n0 + (int64_t)k_lo * d.sps, n_hi = (int64_t)d.n0 + (int64_t)k_hi * d.sps;
  j.la = std::max<int64_t>(-(int64_t)d.pad_lp, n_lo - d.w_lp);
  j.lb = std::min<int64_t>(N - 1 + d.pad_lp, n_hi + d.w_lp);
  j.fa = std::max<int64_t>(0, j.la);
  j.fb = std::min<int64_t>(N - 1, j.lb);
  if (j.la < 0) j.fb = std::max<int64_t>(j.fb, std::min<int64_t>(N - 1, d.pad_lp));
  if (j.lb > N - 1) j.fa = std::min<int64_t>(j.fa, std::max<int64_t>(0, N - 1 - d.pad_lp));
  j.wa = std::max<int64_t>(-(int64_t)d.pad_bp, j.fa - d.w_bp);
  j.wb = std::min<int64_t>(N - 1 + d.pad_bp, j.fb + d.w_bp);
  j.scratch_off = scratch_doubles;
  scratch_doubles += (uint64_t)(j.wb - j.wa + 1) + 2ull * (uint64_t)(j.lb - j.la + 1);
  jobs.push_back(j);
}

template <typename TIn>
static int launch_psk(fb_handle* h, PskMainArgs& ma, uint32_t n_tiles, int nthreads, size_t smem, const PskEdgeArgs& ea,
                      bool use_mma, const fb_psk_design& d, const float* taps, int dtype, uint64_t total_samples) {
  // edge windows on the second stream, interior tiles on the first: they write disjoint words
  FB_CUDA(h, cudaEventRecord(h->ev_fork, h->stream));
  FB_CUDA(h, cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
  if (ea.n_jobs > 0 && !getenv("FB_PSK_NO_EDGE")) {      // (the knob is for timing experiments only: the record edges stay undecided)
    // The edge kernel runs beside the interior kernel, whose CTAs need (nearly) all of an SM's shared memory: ask for the
    // same L1 / shared-memory split, otherwise an SM that hosts an edge CTA must drain before it can be re-configured and the
    // interior CTA placed there starts ~2 ms late (the length of the edge kernel).
    FB_CUDA(h, cudaFuncSetAttribute(psk_edge_kernel<TIn>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    psk_edge_kernel<TIn><<<(ea.n_jobs + 31) / 32, 32, 0, h->stream2>>>(ea);
    h->launches++;
  }
  uint32_t grid = n_tiles;
  if (n_tiles > 0 && use_mma) {
    // interior tiles on the tensor pipe (psk_mma.cu); the few tiles it hands back (samples outside the fp16 split's range)
    // are then evaluated by the fp32 kernel below, CTAs striding over the redo list
    if (h->profiling) FB_CUDA(h, cudaEventRecord(h->ev_k0, h->stream));
    int rc = fb_psk_mma_launch(h, d, taps, ma.samples, total_samples, dtype, ma.tiles, n_tiles, ma.bits, (uint32_t*)h->redo.p,
                               (ea.n_jobs > 0 && !getenv("FB_PSK_NO_EDGE")) ? (int)((ea.n_jobs + 31) / 32) : 0);
    if (rc) return rc;
    if (h->profiling) { FB_CUDA(h, cudaEventRecord(h->ev_k1, h->stream)); h->k_recorded = true; }
    ma.redo = (const uint32_t*)h->redo.p;
    ma.uni_tpr = 0;
    grid = std::min<uint32_t>(n_tiles, (uint32_t)(2 * h->sm_count));
  }
  if (n_tiles > 0) {
    if (h->profiling && !use_mma) FB_CUDA(h, cudaEventRecord(h->ev_k0, h->stream));
    const int ntv = ma.nt == ma.ntp ? ma.nt : 0;
#define FB_LAUNCH_MAIN(NTV, SPSV, PPV, NSLV, REDOV)                                                                                              \
    do {                                                                                                                                       \
      FB_CUDA(h, cudaFuncSetAttribute(psk_main_kernel<TIn, NTV, SPSV, PPV, NSLV, REDOV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      psk_main_kernel<TIn, NTV, SPSV, PPV, NSLV, REDOV><<<grid, nthreads, smem, h->stream>>>(ma);                                                  \
    } while (0)
    if (ma.redo) {
      if (ntv == 16 && ma.sps == 10) FB_LAUNCH_MAIN(16, 10, 0, 0, true);           // the class psk_mma.cu serves
      else return FB_EINVAL;
    }
    else if (ntv == 16 && ma.sps == 10 && ma.P == 2048 && ma.nslow == 2) FB_LAUNCH_MAIN(16, 10, 2048, 2, false);   // 9600 sym/s at 96 kHz, full-size tiles, one pole pair
    else if (ntv == 16 && ma.sps == 10 && ma.P == 2048) FB_LAUNCH_MAIN(16, 10, 2048, 0, false);
    else if (ntv == 16 && ma.sps == 10) FB_LAUNCH_MAIN(16, 10, 0, 0, false);
    else if (ntv == 14 && ma.sps == 80 && ma.P == 320 && ma.nslow == 1) FB_LAUNCH_MAIN(14, 80, 320, 1, false);     // 1200 sym/s at 96 kHz (the reference's argument defaults)
    else if (ntv == 14 && ma.sps == 20 && ma.P == 1280 && ma.nslow == 1) FB_LAUNCH_MAIN(14, 20, 1280, 1, false);   // 4800 sym/s at 96 kHz (DBPSK-4800)
    else if (ntv == 14) FB_LAUNCH_MAIN(14, 0, 0, 0, false);
    else if (ntv == 16) FB_LAUNCH_MAIN(16, 0, 0, 0, false);
    else if (ntv == 18) FB_LAUNCH_MAIN(18, 0, 0, 0, false);
    else FB_LAUNCH_MAIN(0, 0, 0, 0, false);
#undef FB_LAUNCH_MAIN
    if (h->profiling && !use_mma) { FB_CUDA(h, cudaEventRecord(h->ev_k1, h->stream)); h->k_recorded = true; }
    h->launches++;
  }
  FB_CUDA(h, cudaEventRecord(h->ev_join, h->stream2));
  FB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
  FB_CUDA(h, cudaGetLastError());
  return FB_OK;
}

extern "C" int fb_psk_demod_batch(fb_handle* h, const fb_psk_design* dp, const float* taps, const float* slow_w,
                                  int n_rec, const void* samples, const uint64_t* offsets, int dtype, int flags,
                                  uint8_t* out, const uint64_t* out_offsets, uint64_t* out_len, int64_t* sync_idx,
                                  int32_t* status) {
  if (!h || !dp || n_rec < 0 || !offsets || !out_offsets) return FB_EINVAL;
  FB_LOCK(h);
  if (dtype != FB_F32 && dtype != FB_F64 && dtype != FB_S16) return FB_EINVAL;
  const fb_psk_design& d = *dp;
  if (d.sps < 1 || (d.bits_per_sym != 1 && d.bits_per_sym != 2)) return FB_EINVAL;
  if (!d.emulate_only && (!taps || d.nt < 1 || d.nslow < 0 || d.nslow > FB_MAX_SLOW || (d.nslow && !slow_w))) return FB_EINVAL;
  if ((flags & FB_SAMPLES_ON_DEVICE) && ((uintptr_t)samples & 15)) return FB_EINVAL;   // 16-byte aligned batch buffer
  FB_CUDA(h, cudaSetDevice(h->device));
  if (n_rec == 0) return FB_OK;
  const size_t esz = dtype == FB_F32 ? 4 : dtype == FB_F64 ? 8 : 2;
  const int bps = d.bits_per_sym, dper = 32 / bps;

  // ---- host samples, large batch: pipeline the host->device copy with the kernels --------------------------------------
  // The batch is cut into up to 16 groups of whole recordings; group g+1 is copied on the copy stream while group g is
  // demodulated (device-resident recursion on the work stream), so only the first group's copy is exposed.
  size_t pipe_min_bytes = (size_t)256 << 20;
  if (const char* e = getenv("FB_PSK_PIPE_MB")) pipe_min_bytes = (size_t)std::max(0, atoi(e)) << 20;   // tests: force the pipelined path on small batches
  if (!(flags & FB_SAMPLES_ON_DEVICE) && n_rec >= 8 && (size_t)offsets[n_rec] * esz >= pipe_min_bytes) {
    const uint64_t total = offsets[n_rec], total_out_b = out_offsets[n_rec];
    int rc;
    if ((rc = fb_ensure(h, h->in, (size_t)total * esz + 16))) return rc;
    const bool host_out = !(flags & FB_OUT_ON_DEVICE);
    uint8_t* d_out = out; uint64_t* d_out_len = out_len; int64_t* d_sync = sync_idx; int32_t* d_status = status;
    if (host_out) {
      if ((rc = fb_ensure(h, h->out, (size_t)total_out_b + 16))) return rc;
      if ((rc = fb_ensure(h, h->out_len, (size_t)n_rec * 8))) return rc;
      if ((rc = fb_ensure(h, h->sync_idx, (size_t)n_rec * 8))) return rc;
      if ((rc = fb_ensure(h, h->status, (size_t)n_rec * 4))) return rc;
      d_out = (uint8_t*)h->out.p; d_out_len = (uint64_t*)h->out_len.p; d_sync = (int64_t*)h->sync_idx.p; d_status = (int32_t*)h->status.p;
    }
    const int n_groups = std::min(16, n_rec / 4);
    std::vector<int> first(n_groups + 1, n_rec);
    first[0] = 0;
    for (int g = 1, r = 0; g < n_groups; ++g) {                       // equal shares of the samples
      const uint64_t want = total / n_groups * g;
      while (r < n_rec && offsets[r] < want) ++r;
      first[g] = std::max(r, first[g - 1]);
    }
    FB_CUDA(h, cudaEventRecord(h->ev_fork, h->stream));               // the copy stream starts after earlier work on this handle
    FB_CUDA(h, cudaStreamWaitEvent(h->stream_copy, h->ev_fork, 0));
    for (int g = 0; g < n_groups; ++g) {
      const int r0 = first[g], r1 = first[g + 1];
      if (r1 <= r0) continue;
      const uint64_t e0 = offsets[r0], e1 = offsets[r1];
      FB_CUDA(h, cudaMemcpyAsync((char*)h->in.p + e0 * esz, (const char*)samples + e0 * esz, (size_t)(e1 - e0) * esz,
                                 cudaMemcpyHostToDevice, h->stream_copy));
      FB_CUDA(h, cudaEventRecord(h->ev_copy[g], h->stream_copy));
    }
    for (int g = 0; g < n_groups; ++g) {
      const int r0 = first[g], r1 = first[g + 1];
      if (r1 <= r0) continue;
      FB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_copy[g], 0));
      rc = fb_psk_demod_batch(h, dp, taps, slow_w, r1 - r0, h->in.p, offsets + r0, dtype, FB_SAMPLES_ON_DEVICE | FB_OUT_ON_DEVICE | FB_ASYNC,
                              d_out, out_offsets + r0, d_out_len + r0, d_sync + r0, d_status + r0);
      if (rc) return rc;
    }
    if (host_out) {
      if (total_out_b) FB_CUDA(h, cudaMemcpyAsync(out, d_out, (size_t)total_out_b, cudaMemcpyDeviceToHost, h->stream));
      FB_CUDA(h, cudaMemcpyAsync(out_len, d_out_len, (size_t)n_rec * 8, cudaMemcpyDeviceToHost, h->stream));
      FB_CUDA(h, cudaMemcpyAsync(sync_idx, d_sync, (size_t)n_rec * 8, cudaMemcpyDeviceToHost, h->stream));
      FB_CUDA(h, cudaMemcpyAsync(status, d_status, (size_t)n_rec * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    if (!(flags & FB_ASYNC) || host_out) FB_CUDA(h, cudaStreamSynchronize(h->stream));
    return FB_OK;
  }

  // ---- tile geometry of the main kernel -------------------------------------------------------------
  int T = 0, P = 0, nthreads = PM_THREADS, ntp = d.nt, padl = 0;
  const int wlen = d.wcols * d.sps;
  size_t smem = 0;
  bool emulate_only = d.emulate_only != 0;
  const bool use_mma = !emulate_only && taps && fb_psk_mma_usable(h, d, taps);
  if (!emulate_only) {
    if ((d.nt & 1) || d.dh != d.nt / 2 || d.dl != d.nt / 2 - 1) return FB_EINVAL;
    const bool spec = d.nt == 14 || d.nt == 16 || d.nt == 18;      // compile-time tap count: window in registers
    ntp = spec ? d.nt : (d.nt + 3) / 4 * 4;
    padl = spec ? (4 - d.dh % 4) % 4 : 0;
    // the FIR windows are whole float4s: they over-read a few columns right of the last tap
    const int win = spec ? (padl + d.nt + PM_CH - 1 + 3) / 4 * 4 : PM_CH + ntp;
    if ((size_t)d.sps * ntp + 2 * (size_t)std::max(1, d.nslow) * d.sps > PM_MAXTAB) emulate_only = true;   // constant table full
    size_t budget = 100 * 1024;                                    // two CTAs per SM
    int tmax = PM_THREADS * PM_CH - 32;
    if (const char* e = getenv("FB_PSK_T")) { tmax = std::max(32, std::min(tmax, atoi(e) / 32 * 32)); }   // tuning knob
    if (use_mma) tmax = fb_psk_mma_tile_syms();                    // both interior kernels share one tile table
    if (const char* e = getenv("FB_PSK_SMEM_KB")) { budget = (size_t)std::max(8, atoi(e)) * 1024; }
    for (T = tmax; T >= 32; T -= 32) {
      P = (T + 1 + PM_CH - 1) / PM_CH * PM_CH + win;
      smem = (size_t)d.sps * P * 4;
      if (smem <= budget) break;
    }
    if (T < 32) emulate_only = true;
    nthreads = ((T + 1 + PM_CH - 1) / PM_CH + 31) / 32 * 32;
    if (const char* e = getenv("FB_PSK_EXTRA_SMEM_KB")) smem += (size_t)atoi(e) * 1024;   // occupancy experiments
  }

  // ---- per-recording plan ---------------------------------------------------------------------------------
  std::vector<RecPlan> plans(n_rec);
  std::vector<uint32_t> tile_first(n_rec + 1, 0);
  std::vector<EdgeJob> jobs;
  uint64_t words = 0, scratch_doubles = 0;
  uint32_t n_tiles = 0;
  // Whole-record float64 evaluation (`emulate_only` parameter sets, e.g. PSK31's 62 Hz band): one window per recording,
  // 3 doubles of scratch per sample.  Bounded by half of the free device memory (2^26 samples at most) instead of a fixed
  // 4 M samples, so that a 3-minute recording of such a set returns the reference's bytes as well.
  const uint64_t EMU_MAX = 1ull << 26;
  uint64_t emu_budget = ~0ull;             // doubles; queried on the first whole-record job only (cudaMemGetInfo is not free)
  auto emu_budget_doubles = [&]() -> uint64_t {
    if (emu_budget == ~0ull) {
      size_t fr = 0, tot = 0;
      if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) emu_budget = (uint64_t)fr / 2 / 8; else { cudaGetLastError(); emu_budget = 0; }
    }
    return emu_budget;
  };
  for (int r = 0; r < n_rec; ++r) {
    RecPlan& p = plans[r];
    p.off = offsets[r];
    p.n = offsets[r + 1] - offsets[r];
    p.out_off = out_offsets[r];
    p.out_cap = out_offsets[r + 1] - out_offsets[r];
    const int64_t N = (int64_t)p.n;
    p.nsym = (int32_t)(N > d.n0 ? cdiv64(N - d.n0, d.sps) : 0);
    p.ndsym = std::max(p.nsym - 1, 0);
    p.dl32 = p.dr32 = 0;
    p.status = FB_ST_OK;
    p.word_off = words;
    tile_first[r] = n_tiles;
    if (N <= d.pad_bp) { p.status = FB_ST_TOO_SHORT; p.nsym = p.ndsym = 0; continue; }
    if (p.nsym < 2) { p.status = FB_ST_EMPTY; p.ndsym = 0; continue; }
    words += ((uint64_t)p.ndsym * bps + 31) / 32 + 2;
    bool whole = true;
    if (!emulate_only) {
      const int64_t kmin = std::max<int64_t>(0, cdiv64((int64_t)d.zone_left - d.n0, d.sps));
      const int64_t kmax = fdiv64(N - 1 - d.zone_right - d.n0, d.sps);
      const int64_t dl32 = cdiv64(kmin, 32) * 32, dr32 = kmax >= 0 ? kmax / 32 * 32 : 0;
      if (dr32 > dl32 && dr32 <= p.ndsym) {
        p.dl32 = (int32_t)dl32; p.dr32 = (int32_t)dr32;
        n_tiles += (uint32_t)cdiv64(dr32 - dl32, T);
        make_job(d, r, N, 0, p.dl32, scratch_doubles, jobs);
        make_job(d, r, N, p.dr32, p.nsym - 1, scratch_doubles, jobs);
        whole = false;
      }
    }
    if (whole) {
      if (p.n > EMU_MAX || scratch_doubles + 3 * p.n + 4096 > emu_budget_doubles() + h->scratch.cap / 8) { p.status = FB_ST_UNSUPPORTED; p.ndsym = 0; continue; }
      make_job(d, r, N, 0, p.nsym - 1, scratch_doubles, jobs);
    }
  }
  tile_first[n_rec] = n_tiles;
  (void)dper;

  // ---- device buffers -------------------------------------------------------------------------------------------
  const uint64_t total_samples = offsets[n_rec], total_out = out_offsets[n_rec];
  const void* d_samples = samples;
  if (!(flags & FB_SAMPLES_ON_DEVICE)) {
    int rc = fb_ensure(h, h->in, (size_t)total_samples * esz + 16);
    if (rc) return rc;
    FB_CUDA(h, cudaMemcpyAsync(h->in.p, samples, (size_t)total_samples * esz, cudaMemcpyHostToDevice, h->stream));
    d_samples = h->in.p;
  }
  uint8_t* d_out = out; uint64_t* d_out_len = out_len; int64_t* d_sync = sync_idx; int32_t* d_status = status;
  if (!(flags & FB_OUT_ON_DEVICE)) {
    int rc;
    if ((rc = fb_ensure(h, h->out, (size_t)total_out + 16))) return rc;
    if ((rc = fb_ensure(h, h->out_len, (size_t)n_rec * 8))) return rc;
    if ((rc = fb_ensure(h, h->sync_idx, (size_t)n_rec * 8))) return rc;
    if ((rc = fb_ensure(h, h->status, (size_t)n_rec * 4))) return rc;
    d_out = (uint8_t*)h->out.p; d_out_len = (uint64_t*)h->out_len.p; d_sync = (int64_t*)h->sync_idx.p; d_status = (int32_t*)h->status.p;
  }
  int rc;
  if ((rc = fb_ensure(h, h->bits, (size_t)(words + 4) * 4))) return rc;
  if ((rc = fb_ensure(h, h->plans, (size_t)n_rec * sizeof(RecPlan)))) return rc;
  if ((rc = fb_ensure(h, h->tile_first, (size_t)(n_rec + 1) * 4))) return rc;
  if ((rc = fb_ensure(h, h->jobs, std::max<size_t>(1, jobs.size()) * sizeof(EdgeJob)))) return rc;
  if ((rc = fb_ensure(h, h->scratch, (size_t)(scratch_doubles + 2) * 8))) return rc;
  FB_CUDA(h, cudaMemcpyAsync(h->plans.p, plans.data(), (size_t)n_rec * sizeof(RecPlan), cudaMemcpyHostToDevice, h->stream));
  FB_CUDA(h, cudaMemcpyAsync(h->tile_first.p, tile_first.data(), (size_t)(n_rec + 1) * 4, cudaMemcpyHostToDevice, h->stream));
  if (!jobs.empty())
    FB_CUDA(h, cudaMemcpyAsync(h->jobs.p, jobs.data(), jobs.size() * sizeof(EdgeJob), cudaMemcpyHostToDevice, h->stream));

  PskMainArgs ma{};
  if (!emulate_only) {
    // constant table: reversed taps tab[j*ntp + t'] = taps[j][nt-1-t'] (zero padded to ntp), then the slow-pole row
    // weights {p^(sps-j), p^j}.  Global tables (float64 powers of the poles, rounded once):
    //   pwv  [nslow][wlen+1]    p^k                                 tile-boundary state sums
    //   tbl  [nslow][SLOW_TBL]  m^l, m^(2^st), M^w                  scan multipliers, lam = p^sps, m = lam^PM_CH, M = m^32
    ma.wc_off = d.sps * ntp;
    const int npairs = std::max(1, (d.nslow + 1) / 2), ns_ = std::max(1, d.nslow);
    const int wpad = (wlen + 1 + 1024 + 3) / 4 * 4;
    const size_t o_tb = ((size_t)npairs * wpad * 4 * 4 + 255) / 256 * 256, tab_bytes = o_tb + (size_t)ns_ * SLOW_TBL * 2 * 4;
    // the tables depend on the design and the taps only: built and uploaded once per parameter set, not per call
    std::vector<unsigned char> key(sizeof(d) + (size_t)d.sps * d.nt * 8 + 8);
    memcpy(key.data(), &d, sizeof(d));
    memcpy(key.data() + sizeof(d), taps, (size_t)d.sps * d.nt * 8);
    memcpy(key.data() + sizeof(d) + (size_t)d.sps * d.nt * 8, &ntp, 4);
    memcpy(key.data() + sizeof(d) + (size_t)d.sps * d.nt * 8 + 4, &wlen, 4);
    const bool tab_hit = h->psk_tabs.p && h->psk_tabs.cap >= tab_bytes && key == h->psk_tab_key && h->psk_tab_host.size() == sizeof(ma.tab);
    if (tab_hit) memcpy(ma.tab, h->psk_tab_host.data(), sizeof(ma.tab));
    else {
    for (int j = 0; j < d.sps; ++j)
      for (int t = 0; t < d.nt; ++t)
        ma.tab[j * ntp + t] = make_float2(taps[((size_t)j * d.nt + (d.nt - 1 - t)) * 2], taps[((size_t)j * d.nt + (d.nt - 1 - t)) * 2 + 1]);
    std::vector<float> pwv((size_t)npairs * wpad * 4, 0.f), tbl((size_t)ns_ * SLOW_TBL * 2, 0.f);
    for (int i = 0; i < d.nslow; ++i) {
      const double pr = d.slow_p[2 * i], pi = d.slow_p[2 * i + 1];
      std::vector<double> pk((size_t)(std::max(wlen, d.sps) + 1) * 2);
      pk[0] = 1.0; pk[1] = 0.0;
      for (int k = 1; k <= std::max(wlen, d.sps); ++k) {
        pk[2 * k] = pk[2 * k - 2] * pr - pk[2 * k - 1] * pi;
        pk[2 * k + 1] = pk[2 * k - 2] * pi + pk[2 * k - 1] * pr;
      }
      for (int k = 0; k <= wlen; ++k) {
        pwv[((size_t)(i >> 1) * wpad + k) * 4 + 2 * (i & 1)] = (float)pk[2 * k];
        pwv[((size_t)(i >> 1) * wpad + k) * 4 + 2 * (i & 1) + 1] = (float)pk[2 * k + 1];
      }
      for (int j = 0; j < d.sps; ++j) {
        ma.tab[ma.wc_off + (i * d.sps + j) * 2] = make_float2((float)pk[2 * (d.sps - j)], (float)pk[2 * (d.sps - j) + 1]);
        ma.tab[ma.wc_off + (i * d.sps + j) * 2 + 1] = make_float2((float)pk[2 * j], (float)pk[2 * j + 1]);
      }
      auto cpowd = [](double br, double bi, int n, double& rr, double& ri) {
        rr = 1.0; ri = 0.0;
        for (int k = 0; k < n; ++k) { const double t = rr * br - ri * bi; ri = rr * bi + ri * br; rr = t; }
      };
      const double lr = pk[2 * d.sps], li = pk[2 * d.sps + 1];
      double cr, ci, mr, mi, tr_, ti_;
      cpowd(lr, li, PM_CH, cr, ci);                   // m = lam^PM_CH
      cpowd(cr, ci, 32, mr, mi);                      // M = m^32
      float* tb = &tbl[(size_t)i * SLOW_TBL * 2];
      for (int l = 0; l < 32; ++l) { cpowd(cr, ci, l, tr_, ti_); tb[2 * l] = (float)tr_; tb[2 * l + 1] = (float)ti_; }
      for (int st = 0; st < 5; ++st) { cpowd(cr, ci, 1 << st, tr_, ti_); tb[2 * (32 + st)] = (float)tr_; tb[2 * (32 + st) + 1] = (float)ti_; }
      for (int w = 0; w < 9; ++w) { cpowd(mr, mi, w, tr_, ti_); tb[2 * (37 + w)] = (float)tr_; tb[2 * (37 + w) + 1] = (float)ti_; }
    }
    h->psk_tab_key.clear();
    if ((rc = fb_ensure(h, h->psk_tabs, tab_bytes))) return rc;
    FB_CUDA(h, cudaMemcpyAsync((char*)h->psk_tabs.p, pwv.data(), pwv.size() * 4, cudaMemcpyHostToDevice, h->stream));
    FB_CUDA(h, cudaMemcpyAsync((char*)h->psk_tabs.p + o_tb, tbl.data(), tbl.size() * 4, cudaMemcpyHostToDevice, h->stream));
    // the std::vector staging above is pageable: cudaMemcpyAsync has copied it out before returning
    h->psk_tab_host.assign(reinterpret_cast<const unsigned char*>(ma.tab), reinterpret_cast<const unsigned char*>(ma.tab) + sizeof(ma.tab));
    h->psk_tab_key = key;
    }
    char* tabs = (char*)h->psk_tabs.p;
    if ((rc = fb_ensure(h, h->tiles, (size_t)std::max<uint32_t>(1, n_tiles) * sizeof(PskTile)))) return rc;
    if (use_mma && (rc = fb_ensure(h, h->redo, ((size_t)n_tiles + 4) * 4))) return rc;
    if (n_tiles > 0) {
      psk_tiles_kernel<<<(n_tiles + 255) / 256, 256, 0, h->stream>>>((const RecPlan*)h->plans.p, (const uint32_t*)h->tile_first.p, n_rec,
                                                                      n_tiles, T, (PskTile*)h->tiles.p);
      h->launches++;
    }
    ma.samples = d_samples; ma.tiles = (const PskTile*)h->tiles.p; ma.n_tiles = n_tiles;
    {
      // uniform batch?  every recording decodable by the main kernel, same length, back to back, same tile range
      bool uni = n_rec > 0 && n_tiles > 0 && !getenv("FB_PSK_NO_UNIFORM");
      const RecPlan& p0 = plans[0];
      const uint64_t wst = n_rec > 1 ? plans[1].word_off - p0.word_off : 0;
      for (int r = 0; uni && r < n_rec; ++r) {
        const RecPlan& q = plans[r];
        uni = q.status == FB_ST_OK && q.n == p0.n && q.off == p0.off + (uint64_t)r * p0.n && q.dl32 == p0.dl32 && q.dr32 == p0.dr32 &&
              q.dr32 > q.dl32 && q.word_off == (uint64_t)r * wst && tile_first[r] == (uint32_t)r * tile_first[std::min(1, n_rec - 1)] &&
              p0.word_off == 0;
      }
      if (uni && n_rec > 1 && tile_first[1] == 0) uni = false;
      ma.uni_tpr = 0;
      if (uni) {
        ma.uni_tpr = n_rec > 1 ? tile_first[1] : n_tiles;
        ma.uni_dl = p0.dl32; ma.uni_dr = p0.dr32; ma.uni_off0 = p0.off; ma.uni_n = p0.n; ma.uni_wstride = wst;
        if ((uint64_t)ma.uni_tpr * (uint64_t)n_rec != n_tiles) ma.uni_tpr = 0;
      }
    }
    ma.pf_dist = (uint32_t)(2 * h->sm_count);        // resident CTAs: two per SM
    ma.pf_dist = 0x7fffffffu;                        // measured: the prefetch costs more than it hides (2 CTAs/SM already overlap)
    if (const char* e = getenv("FB_PSK_PF")) { if (atoi(e) > 0) ma.pf_dist = (uint32_t)atoi(e); }   // tuning knob
    ma.slow_pw4 = (const float4*)tabs; ma.wpad = wpad; ma.slow_tbl = (const float2*)(tabs + o_tb); ma.bits = (uint32_t*)h->bits.p;
    ma.n_rec = n_rec; ma.sps = d.sps; ma.n0 = d.n0; ma.bps = bps; ma.nt = d.nt; ma.ntp = ntp; ma.dl = d.dl; ma.dh = d.dh;
    ma.nslow = d.nslow; ma.wlen = wlen; ma.pad_bp = d.pad_bp; ma.T = T; ma.P = P; ma.padl = padl;
    ma.rho = make_float2(d.rho[0], d.rho[1]);
    for (int i = 0; i < FB_MAX_SLOW; ++i) {
      ma.lam[i] = make_float2(d.slow_lam[2 * i], d.slow_lam[2 * i + 1]);
      {
        const float rpx = d.slow_rp[2 * i], rpy = d.slow_rp[2 * i + 1], cx = d.slow_rpc[2 * i], cy = d.slow_rpc[2 * i + 1];
        ma.af[i] = make_float4(rpx + cx, cy - rpy, rpy + cy, rpx - cx);
        const float rmx = d.slow_rm[2 * i], rmy = d.slow_rm[2 * i + 1], mx = d.slow_rmc[2 * i], my = d.slow_rmc[2 * i + 1];
        ma.ab[i] = make_float4(rmx + mx, my - rmy, rmy + my, rmx - mx);
      }
      ma.slow_p[2 * i] = d.slow_p[2 * i]; ma.slow_p[2 * i + 1] = d.slow_p[2 * i + 1];
    }
  }
  PskEdgeArgs ea{};
  ea.samples = d_samples; ea.plans = (const RecPlan*)h->plans.p; ea.jobs = (const EdgeJob*)h->jobs.p;
  ea.scratch = (double*)h->scratch.p; ea.bits = (uint32_t*)h->bits.p; ea.n_jobs = (int)jobs.size(); ea.d = d;

  const bool mma_now = use_mma && !emulate_only && T == fb_psk_mma_tile_syms();
  if (dtype == FB_F32) rc = launch_psk<float>(h, ma, n_tiles, nthreads, smem, ea, mma_now, d, taps, dtype, total_samples);
  else if (dtype == FB_F64) rc = launch_psk<double>(h, ma, n_tiles, nthreads, smem, ea, mma_now, d, taps, dtype, total_samples);
  else rc = launch_psk<int16_t>(h, ma, n_tiles, nthreads, smem, ea, mma_now, d, taps, dtype, total_samples);
  if (rc) return rc;

  rc = fb_bits_backend(h, n_rec, (const RecPlan*)h->plans.p, plans, bps, (const uint32_t*)h->bits.p, d_out, d_out_len, d_sync, d_status);
  if (rc) return rc;

  if (!(flags & FB_OUT_ON_DEVICE)) {
    if (total_out) FB_CUDA(h, cudaMemcpyAsync(out, d_out, (size_t)total_out, cudaMemcpyDeviceToHost, h->stream));
    FB_CUDA(h, cudaMemcpyAsync(out_len, d_out_len, (size_t)n_rec * 8, cudaMemcpyDeviceToHost, h->stream));
    FB_CUDA(h, cudaMemcpyAsync(sync_idx, d_sync, (size_t)n_rec * 8, cudaMemcpyDeviceToHost, h->stream));
    FB_CUDA(h, cudaMemcpyAsync(status, d_status, (size_t)n_rec * 4, cudaMemcpyDeviceToHost, h->stream));
  }
  h->last_plans = plans;
  h->last_bps = bps;
  if (!(flags & FB_ASYNC) || !(flags & FB_OUT_ON_DEVICE)) FB_CUDA(h, cudaStreamSynchronize(h->stream));
  return FB_OK;
}

extern "C" int fb_psk_last_bits(fb_handle* h, int rec, uint8_t* bits_out, uint64_t cap_bytes, uint64_t* n_bits) {
  if (!h || !n_bits) return FB_EINVAL;
  FB_LOCK(h);
  if (rec < 0 || rec >= (int)h->last_plans.size()) return FB_EINVAL;
  const RecPlan& p = h->last_plans[rec];
  *n_bits = (uint64_t)p.ndsym * h->last_bps;
  const uint64_t nbytes = (*n_bits + 7) / 8;
  if (!bits_out) return FB_OK;
  if (cap_bytes < nbytes) return FB_EINVAL;
  FB_CUDA(h, cudaSetDevice(h->device));
  FB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (nbytes) FB_CUDA(h, cudaMemcpy(bits_out, (const uint8_t*)h->bits.p + p.word_off * 4, (size_t)nbytes, cudaMemcpyDeviceToHost));
  return FB_OK;
}
