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

#include <algorithm>
#include <math.h>

#define MAIN_S 4             // symbols per thread in the FIR phase
#define SLOW_CH 5            // columns per thread in the slow-pole phase (odd: conflict-free LDS)
#define SLOW_TBL 48          // per pole pair: 32 lane powers, 5 warp-scan multipliers, 9 warp powers (float2 each)

struct PskMainArgs {
  const void* samples;
  const RecPlan* plans;
  const uint32_t* tile_first;   // n_rec + 1 prefix of main tiles
  const float2* taps_r;         // [sps][nt] reversed tap order
  const float4* slow_wc;        // [nslow][sps]  {p^(sps-j), p^j}: forward / backward feature weights of row j
  const float2* slow_pw;        // [nslow][wlen + 1]  p^k: weights of the tile-boundary state sums
  const float2* slow_tbl;       // [nslow][SLOW_TBL]  powers of lam = p^sps used by the column scan
  uint32_t* bits;
  int n_rec;
  int sps, n0, bps, nt, dl, dh, nslow, wlen, pad_bp;
  int T;                        // tile size in differential symbols (multiple of 32)
  int P;                        // shared-memory row pitch in floats (multiple of 4)
  int rg;                       // row groups (threads cooperating on one symbol chunk)
  int right;                    // columns staged right of the last tile symbol
  float2 rho;
  float2 lam[FB_MAX_SLOW];
  // R F + R' conj(F) as a real 2x2 map of (Re F, Im F): {a11, a12, a21, a22}; af = forward, ab = backward residues
  float4 af[FB_MAX_SLOW], ab[FB_MAX_SLOW];
  double slow_p[2 * FB_MAX_SLOW];
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

// Register-tiled polyphase FIR over the rows j = g, g+rg, ... of one symbol chunk.
//   acc[s] += taps_r[j][t'] * X[j][base + s + t']      s < S, t' < nt (nt even: one LDS.128 = two complex taps)
// NTH > 0: tap-pair count known at compile time (window fully in registers); NTH == 0: runtime loop.
template <int S, int NTH>
__device__ __forceinline__ void fir_rows(const float* X, const float2* taps, int P, int nt, int sps, int g, int rg,
                                         int base, float (&accr)[S], float (&acci)[S]) {
  for (int j = g; j < sps; j += rg) {
    const float4* row = reinterpret_cast<const float4*>(X + (size_t)j * P + base);
    const float4* tp = reinterpret_cast<const float4*>(taps + (size_t)j * nt);
    if (NTH > 0) {
      constexpr int NW = (2 * NTH + S - 1 + 3) / 4;           // float4 loads covering nt + S - 1 columns
      float win[4 * NW];
#pragma unroll
      for (int q = 0; q < NW; ++q) {
        const float4 v = row[q];
        win[4 * q] = v.x; win[4 * q + 1] = v.y; win[4 * q + 2] = v.z; win[4 * q + 3] = v.w;
      }
#pragma unroll
      for (int th = 0; th < NTH; ++th) {
        const float4 t01 = tp[th];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          accr[s] = fmaf(t01.x, win[2 * th + s], accr[s]);     acci[s] = fmaf(t01.y, win[2 * th + s], acci[s]);
          accr[s] = fmaf(t01.z, win[2 * th + s + 1], accr[s]); acci[s] = fmaf(t01.w, win[2 * th + s + 1], acci[s]);
        }
      }
    } else {
      float win[S + 4];
#pragma unroll
      for (int q = 0; q < S / 4; ++q) {
        const float4 v = row[q];
        win[4 * q] = v.x; win[4 * q + 1] = v.y; win[4 * q + 2] = v.z; win[4 * q + 3] = v.w;
      }
      const int nth = nt / 2;
      for (int th = 0; th < nth; th += 2) {                    // two tap pairs per step (second may be absent)
        const float4 v = row[th / 2 + S / 4];
        win[S] = v.x; win[S + 1] = v.y; win[S + 2] = v.z; win[S + 3] = v.w;
        const float4 t01 = tp[th];
        const float4 t23 = (th + 1 < nth) ? tp[th + 1] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int s = 0; s < S; ++s) {
          accr[s] = fmaf(t01.x, win[s], accr[s]);     acci[s] = fmaf(t01.y, win[s], acci[s]);
          accr[s] = fmaf(t01.z, win[s + 1], accr[s]); acci[s] = fmaf(t01.w, win[s + 1], acci[s]);
          accr[s] = fmaf(t23.x, win[s + 2], accr[s]); acci[s] = fmaf(t23.y, win[s + 2], acci[s]);
          accr[s] = fmaf(t23.z, win[s + 3], accr[s]); acci[s] = fmaf(t23.w, win[s + 3], acci[s]);
        }
#pragma unroll
        for (int s = 0; s < S; ++s) win[s] = win[s + 4];
      }
    }
  }
}

template <typename TIn>
__global__ void __launch_bounds__(FB_THREADS, 3) psk_main_kernel(const PskMainArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sps = a.sps;
  // ---- which recording / tile -----------------------------------------------------------------
  // largest r with tile_first[r] <= tile: 32-way splitter search (two dependent loads for up to 1024 recordings)
  const uint32_t tile = blockIdx.x;
  int lo = 0, hi = a.n_rec;
  while (hi - lo > 1) {
    const int step = (hi - lo + 31) >> 5;
    const int probe = lo + (lane + 1) * step;
    const bool le = probe < hi && __ldg(&a.tile_first[probe]) <= tile;
    const int cnt = __popc(__ballot_sync(0xffffffffu, le));       // probes are monotone: the first cnt satisfy it
    const int nlo = lo + cnt * step;
    hi = min(hi, nlo + step);
    lo = nlo;
  }
  const RecPlan pl = a.plans[lo];
  const int d0 = pl.dl32 + (int)(tile - a.tile_first[lo]) * a.T;
  const int d1 = min(d0 + a.T, pl.dr32);
  const int ns = d1 - d0 + 1;                       // symbols d0 .. d1
  const int64_t N = (int64_t)pl.n;
  // ---- shared memory carve-up -----------------------------------------------------------------
  float* X = smem;                                  // [sps][P]
  float2* taps = reinterpret_cast<float2*>(X + (size_t)sps * a.P);      // [sps][nt]
  float4* wc = reinterpret_cast<float4*>(taps + (size_t)sps * a.nt);     // [nslow][sps]
  float2* Y = reinterpret_cast<float2*>(wc + (size_t)max(1, a.nslow) * sps);   // [T + 4] slow contribution, then y'
  float2* red = Y + (a.T + 4);                      // [8 warps][4] partials, [8][4] warp carries, [4] boundary states x2
  __shared__ double finit_sh[2 * FB_MAX_SLOW];

  const int ca = d0 - a.dh;                         // first staged column (global column == symbol index)
  const int ncols = (d1 + a.right) - ca + 1;
  const int64_t n_d0 = (int64_t)a.n0 + (int64_t)d0 * sps;              // sample index of symbol d0
  const int64_t n_e1 = (int64_t)a.n0 + (int64_t)(d1 + 1) * sps;        // first sample after the tile's last column
  // ---- stage samples: thread <-> column (sps consecutive samples), conflict-free row stores ------------
  // A warp reads 32*sps consecutive samples; each 128-byte line is fetched from L2 once and re-hit in L1.
  {
    const int64_t n_a = (int64_t)a.n0 + (int64_t)ca * sps;          // sample index of staged element 0
    for (int c = tid; c < ncols; c += FB_THREADS) {
      const int64_t n = n_a + (int64_t)c * sps;
      float* dst = X + c;
      if (n >= 0 && n + sps <= N) {
        const uint64_t g = pl.off + (uint64_t)n;
        if (sizeof(TIn) == 4 && (sps & 1) == 0 && (g & 1) == 0) {  // 8-byte aligned pairs
          const float2* src = reinterpret_cast<const float2*>(reinterpret_cast<const float*>(a.samples) + g);
          const int P2 = 2 * a.P;
          float* d2 = dst;
#pragma unroll 5
          for (int j = 0; j < sps; j += 2) {
            const float2 v = __ldg(src + (j >> 1));
            d2[0] = v.x;
            d2[a.P] = v.y;
            d2 += P2;
          }
        } else {
          for (int j = 0; j < sps; ++j) dst[j * a.P] = load_sample<TIn>(a.samples, g + j);
        }
      } else {
        for (int j = 0; j < sps; ++j)
          dst[j * a.P] = (n + j >= 0 && n + j < N) ? load_sample<TIn>(a.samples, pl.off + (uint64_t)(n + j)) : 0.f;
      }
    }
    for (int i = tid; i < sps * a.nt; i += FB_THREADS) taps[i] = a.taps_r[i];
    for (int i = tid; i < a.nslow * sps; i += FB_THREADS) wc[i] = a.slow_wc[i];
    for (int i = tid; i < a.T + 4; i += FB_THREADS) Y[i] = make_float2(0.f, 0.f);
  }
  // ---- exact start state of the forward slow recursion at column 0 (left record edge) ----------
  const bool near_left = (n_d0 - a.wlen) <= (int64_t)a.n0;          // the boundary sum would reach column 0
  if (near_left && tid < a.nslow) {
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
    // Fst = p * s
    finit_sh[2 * tid] = pr * sr - pi * si;
    finit_sh[2 * tid + 1] = pr * si + pi * sr;
  }

  // ---- slow pole pairs ------------------------------------------------------------------------------
  // y_slow[c] = sum_i  R+_i Fst_i[c] + R+'_i conj(Fst_i[c]) + R-_i Bst_i[c] + R-'_i conj(Bst_i[c])
  //   Fst[c]   = sum_{n < n_c} p^(n_c - n) x[n]      Fst[c+1] = lam Fst[c] + sum_j p^(sps-j) X[j][c]
  //   Bfull[c] = sum_{n >= n_c} p^(n - n_c) x[n]     Bfull[c] = sum_j p^j X[j][c] + lam Bfull[c+1];  Bst = Bfull - x[n_c]
  // The states at the tile boundaries (Fst[d0], Bfull[d1+1]) are direct sums over the previous / next `wlen`
  // samples against the power table p^k (read straight from global memory: no halo staging); inside the tile
  // the recursion runs on per-column features with a decaying scan (registers + warp shuffles).
  for (int pair = 0; pair < a.nslow; pair += 2) {
    const bool two = pair + 1 < a.nslow;
    const int i1 = two ? pair + 1 : pair;
    const float2 lam0 = a.lam[pair], lam1 = a.lam[i1];
    // (1) boundary sums: index 0,1 = forward pole 0,1; 2,3 = backward pole 0,1.  Warps 0-1 take the forward sum,
    //     warps 2-3 the backward one (4 consecutive samples per thread and step); the other warps go straight on.
    {
      float2 bs0 = make_float2(0.f, 0.f), bs1 = make_float2(0.f, 0.f);
      if (warp < 4) {
        const float2* pw0 = a.slow_pw + (size_t)pair * (a.wlen + 1);
        const float2* pw1 = a.slow_pw + (size_t)i1 * (a.wlen + 1);
        const bool fwd = warp < 2;
        const int t64 = tid & 63;
        int cnt; int64_t nbeg; int kbeg, kstep;                 // sample n = nbeg + m has weight p^(kbeg + kstep*m)
        if (fwd) {
          const int64_t flo = near_left ? (int64_t)a.n0 : n_d0 - a.wlen;       // n in [flo, n_d0), weight p^(n_d0 - n)
          cnt = (int)(n_d0 - flo); nbeg = flo; kbeg = cnt; kstep = -1;
        } else {
          cnt = (int)max((int64_t)0, min((int64_t)a.wlen, N - n_e1));           // n in [n_e1, n_e1 + cnt), weight p^(n - n_e1)
          nbeg = n_e1; kbeg = 0; kstep = 1;
        }
        for (int m0 = 4 * t64; m0 < cnt; m0 += 256) {
          float xv[4];
          float2 w0[4], w1[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool ok = m0 + u < cnt;
            const int k = kbeg + kstep * (m0 + u);
            xv[u] = ok ? load_sample<TIn>(a.samples, pl.off + (uint64_t)(nbeg + m0 + u)) : 0.f;
            w0[u] = ok ? __ldg(&pw0[k]) : make_float2(0.f, 0.f);
            w1[u] = ok ? __ldg(&pw1[k]) : make_float2(0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            bs0.x = fmaf(w0[u].x, xv[u], bs0.x); bs0.y = fmaf(w0[u].y, xv[u], bs0.y);
            bs1.x = fmaf(w1[u].x, xv[u], bs1.x); bs1.y = fmaf(w1[u].y, xv[u], bs1.y);
          }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          bs0.x += __shfl_xor_sync(0xffffffffu, bs0.x, off); bs0.y += __shfl_xor_sync(0xffffffffu, bs0.y, off);
          bs1.x += __shfl_xor_sync(0xffffffffu, bs1.x, off); bs1.y += __shfl_xor_sync(0xffffffffu, bs1.y, off);
        }
        if (lane == 0) { red[64 + warp * 2] = bs0; red[64 + warp * 2 + 1] = bs1; }   // [w0 f0,f1][w1 f0,f1][w2 b0,b1][w3 b0,b1]
      }
    }
    if (pair == 0) __syncthreads();                   // staged samples (and finit) visible
    // (2) per-column features of this thread's SLOW_CH columns (scan index e: forward column d0 + e, backward d1 - e)
    const int e0 = tid * SLOW_CH;
    float2 z[4][SLOW_CH];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int i = 0; i < SLOW_CH; ++i) z[k][i] = make_float2(0.f, 0.f);
    if (e0 < ns) {
      // columns d0 + e0 + i (forward) and d1 - e0 - i (backward): contiguous, so one base address each and
      // immediate offsets; columns past the tile are staged halo / zero and are discarded just below
      const float4* w0p = wc + (size_t)pair * sps;
      const float4* w1p = wc + (size_t)i1 * sps;
      const float* rf = X + (d0 + e0 - ca);
      const float* rb = X + (d1 - e0 - ca);
      for (int j = 0; j < sps; ++j) {
        const float4 w0 = w0p[j], w1 = w1p[j];
#pragma unroll
        for (int i = 0; i < SLOW_CH; ++i) {
          const float xf = rf[i], xb = rb[-i];
          z[0][i].x = fmaf(w0.x, xf, z[0][i].x); z[0][i].y = fmaf(w0.y, xf, z[0][i].y);
          z[1][i].x = fmaf(w1.x, xf, z[1][i].x); z[1][i].y = fmaf(w1.y, xf, z[1][i].y);
          z[2][i].x = fmaf(w0.z, xb, z[2][i].x); z[2][i].y = fmaf(w0.w, xb, z[2][i].y);
          z[3][i].x = fmaf(w1.z, xb, z[3][i].x); z[3][i].y = fmaf(w1.w, xb, z[3][i].y);
        }
        rf += a.P; rb += a.P;
      }
#pragma unroll
      for (int i = 0; i < SLOW_CH; ++i)
        if (e0 + i >= ns) {
#pragma unroll
          for (int k = 0; k < 4; ++k) z[k][i] = make_float2(0.f, 0.f);
        }
    }
    // (3) decaying scan: thread totals -> warp shuffle scan -> warp carries (warp 0) -> per-column states
    const float2* tb0 = a.slow_tbl + (size_t)pair * SLOW_TBL;
    const float2* tb1 = a.slow_tbl + (size_t)i1 * SLOW_TBL;
    float2 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 lam = (k & 1) ? lam1 : lam0;
      float2 s = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < SLOW_CH; ++i) s = cfma(lam, s, z[k][i]);
      v[k] = s;
    }
#pragma unroll
    for (int st = 0; st < 5; ++st) {
      const float2 m0 = __ldg(&tb0[32 + st]), m1 = __ldg(&tb1[32 + st]);   // (lam^CH)^(2^st)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float ox = __shfl_up_sync(0xffffffffu, v[k].x, 1 << st), oy = __shfl_up_sync(0xffffffffu, v[k].y, 1 << st);
        if (lane >= (1 << st)) v[k] = cfma((k & 1) ? m1 : m0, make_float2(ox, oy), v[k]);
      }
    }
    if (lane == 31) {
#pragma unroll
      for (int k = 0; k < 4; ++k) red[warp * 4 + k] = v[k];
    }
    __syncthreads();
    if (warp == 0) {                                  // lane = 4 w + k: carry entering warp w for sequence k
      const int k = lane & 3, w = lane >> 2;
      const float2* tb = (k & 1) ? tb1 : tb0;
      float2 t = red[lane];                           // total of warp w (zero start)
#pragma unroll
      for (int st = 0; st < 3; ++st) {                // inclusive scan over w with multiplier M^(2^st), M = lam^(32 CH)
        const float2 mm = __ldg(&tb[37 + (1 << st)]); // M^1, M^2, M^4
        const float ox = __shfl_up_sync(0xffffffffu, t.x, 4 << st), oy = __shfl_up_sync(0xffffffffu, t.y, 4 << st);
        if (w >= (1 << st)) t = cfma(mm, make_float2(ox, oy), t);
      }
      const float ex = __shfl_up_sync(0xffffffffu, t.x, 4), ey = __shfl_up_sync(0xffffffffu, t.y, 4);
      float2 cin = (w > 0) ? make_float2(ex, ey) : make_float2(0.f, 0.f);
      // boundary state of sequence k (0,1 = forward pole 0,1; 2,3 = backward): two warp partials (+ the exact left start)
      const int q = 64 + (k >> 1) * 4 + (k & 1);
      float2 bnd = make_float2(red[q].x + red[q + 2].x, red[q].y + red[q + 2].y);
      if (k < 2 && near_left) {                       // + p^(n_d0 - n0) Fst[0]
        const int i = (k == 0) ? pair : i1;
        const float2 pw = __ldg(&a.slow_pw[(size_t)i * (a.wlen + 1) + (int)(n_d0 - a.n0)]);
        bnd = cfma(pw, make_float2((float)finit_sh[2 * i], (float)finit_sh[2 * i + 1]), bnd);
      }
      cin = cfma(__ldg(&tb[37 + w]), bnd, cin);       // + M^w * boundary state
      red[32 + lane] = cin;
    }
    __syncthreads();
    float2 sc[4];                                     // state entering this thread's first column
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float ex = __shfl_up_sync(0xffffffffu, v[k].x, 1), ey = __shfl_up_sync(0xffffffffu, v[k].y, 1);
      const float2 excl = lane > 0 ? make_float2(ex, ey) : make_float2(0.f, 0.f);
      sc[k] = cfma(__ldg(&((k & 1) ? tb1 : tb0)[lane]), red[32 + warp * 4 + k], excl);
    }
    if (e0 < ns) {
      const float4 af0 = a.af[pair], af1 = a.af[i1], ab0 = a.ab[pair], ab1 = a.ab[i1];
#pragma unroll
      for (int i = 0; i < SLOW_CH; ++i) {
        const int e = e0 + i;
        if (e < ns) {                                 // forward: Fst[d0 + e] = S[e] (state before column e)
          float ar = fmaf(af0.x, sc[0].x, af0.y * sc[0].y), ai = fmaf(af0.z, sc[0].x, af0.w * sc[0].y);
          if (two) {
            ar = fmaf(af1.x, sc[1].x, fmaf(af1.y, sc[1].y, ar)); ai = fmaf(af1.z, sc[1].x, fmaf(af1.w, sc[1].y, ai));
          }
          atomicAdd(&Y[e].x, ar); atomicAdd(&Y[e].y, ai);
        }
        sc[0] = cfma(lam0, sc[0], z[0][i]);
        sc[1] = cfma(lam1, sc[1], z[1][i]);
        sc[2] = cfma(lam0, sc[2], z[2][i]);
        sc[3] = cfma(lam1, sc[3], z[3][i]);
        if (e < ns) {                                 // backward: Bfull[d1 - e] = S[e + 1]; Bst = Bfull - x[n_col]
          const int col = d1 - e;
          const float x0 = X[col - ca];
          float br = sc[2].x - x0;
          float ar = fmaf(ab0.x, br, ab0.y * sc[2].y), ai = fmaf(ab0.z, br, ab0.w * sc[2].y);
          if (two) {
            br = sc[3].x - x0;
            ar = fmaf(ab1.x, br, fmaf(ab1.y, sc[3].y, ar)); ai = fmaf(ab1.z, br, fmaf(ab1.w, sc[3].y, ai));
          }
          atomicAdd(&Y[col - d0].x, ar); atomicAdd(&Y[col - d0].y, ai);
        }
      }
    }
    __syncthreads();
  }
  if (a.nslow == 0) __syncthreads();

  // ---- fast part: register-tiled polyphase FIR at symbol instants --------------------------------
  {
    const int nchunks = (ns + MAIN_S - 1) / MAIN_S;
    const int chunk = tid % nchunks, g = tid / nchunks;
    if (g < a.rg) {
      float accr[MAIN_S], acci[MAIN_S];
#pragma unroll
      for (int s = 0; s < MAIN_S; ++s) { accr[s] = 0.f; acci[s] = 0.f; }
      // y[s] += tapsR[j][t'] * X[j][base + s + t'],  base = chunk*S  (staging starts at column d0 - dh)
      const int base = chunk * MAIN_S;
      switch (a.nt / 2) {
        case 8: fir_rows<MAIN_S, 8>(X, taps, a.P, a.nt, sps, g, a.rg, base, accr, acci); break;
        case 9: fir_rows<MAIN_S, 9>(X, taps, a.P, a.nt, sps, g, a.rg, base, accr, acci); break;
        case 10: fir_rows<MAIN_S, 10>(X, taps, a.P, a.nt, sps, g, a.rg, base, accr, acci); break;
        default: fir_rows<MAIN_S, 0>(X, taps, a.P, a.nt, sps, g, a.rg, base, accr, acci); break;
      }
#pragma unroll
      for (int s = 0; s < MAIN_S; ++s) {
        const int e = chunk * MAIN_S + s;
        if (e < ns) {
          if (a.rg == 1) {
            float2 y = Y[e];
            Y[e] = make_float2(y.x + accr[s], y.y + acci[s]);
          } else {
            atomicAdd(&Y[e].x, accr[s]);
            atomicAdd(&Y[e].y, acci[s]);
          }
        }
      }
    }
  }
  __syncthreads();

  // ---- differential decisions, 32 bits per thread, big-endian words -------------------------------
  {
    const int dper = 32 / a.bps;                      // dsyms per word
    const int nwords = (d1 - d0) / dper;
    for (int wd = tid; wd < nwords; wd += FB_THREADS) {
      uint32_t word = 0;
      float2 prev = Y[wd * dper];
      for (int k = 0; k < dper; ++k) {
        const float2 cur = Y[wd * dper + k + 1];
        // d = cur * conj(prev) * rho
        const float tr = fmaf(cur.x, prev.x, cur.y * prev.y), ti = fmaf(cur.y, prev.x, -cur.x * prev.y);
        const float dr = fmaf(tr, a.rho.x, -ti * a.rho.y), di = fmaf(tr, a.rho.y, ti * a.rho.x);
        word = (word << a.bps) | psk_decide<float>(dr, di, a.bps);
        prev = cur;
      }
      a.bits[pl.word_off + (uint64_t)(d0 / dper) + wd] = __byte_perm(word, 0, 0x0123);
    }
  }
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
        // s[k] conj(s[k-1]) -- both already carry the LO phase, as in the reference (modem.py:100, 214)
        const double dr = sr * pr + si * pi, di = si * pr - sr * pi;
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
static int launch_psk(fb_handle* h, const PskMainArgs& ma, uint32_t n_tiles, size_t smem, const PskEdgeArgs& ea) {
  // edge windows on the second stream, interior tiles on the first: they write disjoint words
  FB_CUDA(h, cudaEventRecord(h->ev_fork, h->stream));
  FB_CUDA(h, cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
  if (ea.n_jobs > 0) {
    psk_edge_kernel<TIn><<<(ea.n_jobs + 31) / 32, 32, 0, h->stream2>>>(ea);
    h->launches++;
  }
  if (n_tiles > 0) {
    FB_CUDA(h, cudaFuncSetAttribute(psk_main_kernel<TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (h->profiling) FB_CUDA(h, cudaEventRecord(h->ev_k0, h->stream));
    psk_main_kernel<TIn><<<n_tiles, FB_THREADS, smem, h->stream>>>(ma);
    if (h->profiling) { FB_CUDA(h, cudaEventRecord(h->ev_k1, h->stream)); h->k_recorded = true; }
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
  if (dtype != FB_F32 && dtype != FB_F64 && dtype != FB_S16) return FB_EINVAL;
  const fb_psk_design& d = *dp;
  if (d.sps < 1 || (d.bits_per_sym != 1 && d.bits_per_sym != 2)) return FB_EINVAL;
  if (!d.emulate_only && (!taps || d.nt < 1 || d.nslow < 0 || d.nslow > FB_MAX_SLOW || (d.nslow && !slow_w))) return FB_EINVAL;
  if ((flags & FB_SAMPLES_ON_DEVICE) && ((uintptr_t)samples & 15)) return FB_EINVAL;   // 16-byte aligned batch buffer
  FB_CUDA(h, cudaSetDevice(h->device));
  if (n_rec == 0) return FB_OK;
  const size_t esz = dtype == FB_F32 ? 4 : dtype == FB_F64 ? 8 : 2;
  const int bps = d.bits_per_sym, dper = 32 / bps;

  // ---- tile geometry of the main kernel -------------------------------------------------------------
  int T = 0, P = 0, rg = 1, right = 0;
  const int wlen = d.wcols * d.sps;
  size_t smem = 0;
  if (!d.emulate_only) {
    if ((d.nt & 1) || d.dh < SLOW_CH - 1) return FB_EINVAL;
    // the FIR windows are whole float4s: they over-read a few columns right of the last tap
    right = d.dl + MAIN_S + 12;
    const size_t budget = 72 * 1024;
    for (T = std::min(FB_THREADS * MAIN_S - 32, (FB_THREADS * SLOW_CH - 1) / 32 * 32); T >= 32; T -= 32) {
      const int ncols = T + 1 + d.dh + right;
      P = (ncols + 3) / 4 * 4 + 4;
      smem = (size_t)d.sps * P * 4 + (size_t)d.sps * d.nt * 8 + (size_t)std::max(1, d.nslow) * d.sps * 16 +
             (size_t)(T + 4) * 8 + 72 * 8;
      if (smem <= budget) break;
    }
    if (T < 32) return FB_EUNSUPPORTED;   // design.py marks such parameter sets emulate_only
    const int nchunks = (T + 1 + MAIN_S - 1) / MAIN_S;
    rg = std::max(1, std::min(d.sps, FB_THREADS / nchunks));
  }

  // ---- per-recording plan ---------------------------------------------------------------------------------
  std::vector<RecPlan> plans(n_rec);
  std::vector<uint32_t> tile_first(n_rec + 1, 0);
  std::vector<EdgeJob> jobs;
  uint64_t words = 0, scratch_doubles = 0;
  uint32_t n_tiles = 0;
  const uint64_t EMU_MAX = 1ull << 22;     // samples a single window may span when the whole record is emulated
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
    if (!d.emulate_only) {
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
      if (p.n > EMU_MAX) { p.status = FB_ST_UNSUPPORTED; p.ndsym = 0; continue; }
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
  if (!d.emulate_only) {
    // reversed tap table: taps_r[j][t'] = taps[j][nt-1-t']
    std::vector<float> tr((size_t)d.sps * d.nt * 2, 0.f);
    for (int j = 0; j < d.sps; ++j)
      for (int t = 0; t < d.nt; ++t) {
        tr[((size_t)j * d.nt + t) * 2] = taps[((size_t)j * d.nt + (d.nt - 1 - t)) * 2];
        tr[((size_t)j * d.nt + t) * 2 + 1] = taps[((size_t)j * d.nt + (d.nt - 1 - t)) * 2 + 1];
      }
    // slow-pole tables (float64 powers of the poles, rounded once):
    //   wcv  [nslow][sps]       {p^(sps-j), p^j}                   per-row feature weights
    //   pwv  [nslow][wlen+1]    p^k                                 tile-boundary state sums
    //   tbl  [nslow][SLOW_TBL]  (lam^CH)^l, (lam^CH)^(2^st), M^w    scan multipliers, lam = p^sps, M = lam^(32 CH)
    const int ns_ = std::max(1, d.nslow);
    std::vector<float> wcv((size_t)ns_ * d.sps * 4, 0.f), pwv((size_t)ns_ * (wlen + 1) * 2, 0.f), tbl((size_t)ns_ * SLOW_TBL * 2, 0.f);
    for (int i = 0; i < d.nslow; ++i) {
      const double pr = d.slow_p[2 * i], pi = d.slow_p[2 * i + 1];
      std::vector<double> pk((size_t)(std::max(wlen, d.sps) + 1) * 2);
      pk[0] = 1.0; pk[1] = 0.0;
      for (int k = 1; k <= std::max(wlen, d.sps); ++k) {
        pk[2 * k] = pk[2 * k - 2] * pr - pk[2 * k - 1] * pi;
        pk[2 * k + 1] = pk[2 * k - 2] * pi + pk[2 * k - 1] * pr;
      }
      for (int k = 0; k <= wlen; ++k) {
        pwv[((size_t)i * (wlen + 1) + k) * 2] = (float)pk[2 * k];
        pwv[((size_t)i * (wlen + 1) + k) * 2 + 1] = (float)pk[2 * k + 1];
      }
      for (int j = 0; j < d.sps; ++j) {
        float* o = &wcv[((size_t)i * d.sps + j) * 4];
        o[0] = (float)pk[2 * (d.sps - j)]; o[1] = (float)pk[2 * (d.sps - j) + 1];
        o[2] = (float)pk[2 * j]; o[3] = (float)pk[2 * j + 1];
      }
      auto cpowd = [](double br, double bi, int n, double& rr, double& ri) {
        rr = 1.0; ri = 0.0;
        for (int k = 0; k < n; ++k) { const double t = rr * br - ri * bi; ri = rr * bi + ri * br; rr = t; }
      };
      const double lr = pk[2 * d.sps], li = pk[2 * d.sps + 1];
      double cr, ci, mr, mi, tr_, ti_;
      cpowd(lr, li, SLOW_CH, cr, ci);                 // lam^CH
      cpowd(cr, ci, 32, mr, mi);                      // M = lam^(32 CH)
      float* tb = &tbl[(size_t)i * SLOW_TBL * 2];
      for (int l = 0; l < 32; ++l) { cpowd(cr, ci, l, tr_, ti_); tb[2 * l] = (float)tr_; tb[2 * l + 1] = (float)ti_; }
      for (int st = 0; st < 5; ++st) { cpowd(cr, ci, 1 << st, tr_, ti_); tb[2 * (32 + st)] = (float)tr_; tb[2 * (32 + st) + 1] = (float)ti_; }
      for (int w = 0; w < 9; ++w) { cpowd(mr, mi, w, tr_, ti_); tb[2 * (37 + w)] = (float)tr_; tb[2 * (37 + w) + 1] = (float)ti_; }
    }
    const size_t o_wc = (tr.size() * 4 + 255) / 256 * 256, o_pw = o_wc + (wcv.size() * 4 + 255) / 256 * 256,
                 o_tb = o_pw + (pwv.size() * 4 + 255) / 256 * 256, tab_bytes = o_tb + tbl.size() * 4;
    if ((rc = fb_ensure(h, h->taps, tab_bytes))) return rc;
    char* tabs = (char*)h->taps.p;
    FB_CUDA(h, cudaMemcpyAsync(tabs, tr.data(), tr.size() * 4, cudaMemcpyHostToDevice, h->stream));
    FB_CUDA(h, cudaMemcpyAsync(tabs + o_wc, wcv.data(), wcv.size() * 4, cudaMemcpyHostToDevice, h->stream));
    FB_CUDA(h, cudaMemcpyAsync(tabs + o_pw, pwv.data(), pwv.size() * 4, cudaMemcpyHostToDevice, h->stream));
    FB_CUDA(h, cudaMemcpyAsync(tabs + o_tb, tbl.data(), tbl.size() * 4, cudaMemcpyHostToDevice, h->stream));
    // the std::vector staging above is pageable: cudaMemcpyAsync has copied it out before returning
    ma.samples = d_samples; ma.plans = (const RecPlan*)h->plans.p; ma.tile_first = (const uint32_t*)h->tile_first.p;
    ma.taps_r = (const float2*)tabs; ma.slow_wc = (const float4*)(tabs + o_wc); ma.slow_pw = (const float2*)(tabs + o_pw);
    ma.slow_tbl = (const float2*)(tabs + o_tb); ma.bits = (uint32_t*)h->bits.p;
    ma.n_rec = n_rec; ma.sps = d.sps; ma.n0 = d.n0; ma.bps = bps; ma.nt = d.nt; ma.dl = d.dl; ma.dh = d.dh;
    ma.nslow = d.nslow; ma.wlen = wlen; ma.pad_bp = d.pad_bp; ma.T = T; ma.P = P; ma.rg = rg; ma.right = right;
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

  if (dtype == FB_F32) rc = launch_psk<float>(h, ma, n_tiles, smem, ea);
  else if (dtype == FB_F64) rc = launch_psk<double>(h, ma, n_tiles, smem, ea);
  else rc = launch_psk<int16_t>(h, ma, n_tiles, smem, ea);
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
  if (!h || rec < 0 || rec >= (int)h->last_plans.size() || !n_bits) return FB_EINVAL;
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
