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
#define MAIN_Q 4             // taps per register-tile step

struct PskMainArgs {
  const void* samples;
  const RecPlan* plans;
  const uint32_t* tile_first;   // n_rec + 1 prefix of main tiles
  const float2* taps_r;         // [sps][ntp] reversed tap order, zero padded to ntp (multiple of MAIN_Q)
  const float2* slow_w;         // [nslow][sps + 1]  p^m
  uint32_t* bits;
  int n_rec;
  int sps, n0, bps, nt, ntp, dl, dh, nslow, wcols, pad_bp;
  int T;                        // tile size in differential symbols (multiple of 32)
  int P;                        // shared-memory row pitch in floats (multiple of 4)
  int lead4;                    // columns staged left of the first tile symbol
  int rg;                       // row groups (threads cooperating on one symbol chunk)
  int right;                    // columns staged right of the last tile symbol (covers slow warm-up and padded taps)
  int zcap;                     // scan scratch capacity (complex)
  float2 rho;
  float2 lam[FB_MAX_SLOW], rp[FB_MAX_SLOW], rpc[FB_MAX_SLOW], rm[FB_MAX_SLOW], rmc[FB_MAX_SLOW];
  double slow_p[2 * FB_MAX_SLOW];
};

__device__ __forceinline__ float2 cpow_int(float2 z, int n) {   // z^n, n >= 0
  float2 r = make_float2(1.f, 0.f);
  while (n > 0) {
    if (n & 1) r = cmul(r, z);
    z = cmul(z, z);
    n >>= 1;
  }
  return r;
}

// In-place inclusive decaying scan over Z[0..L): S[i] = Z[i] + lam * S[i-1], S[-1] = init.
// Called by all FB_THREADS threads; wtot is 8 complex of shared scratch.
__device__ void decaying_scan(float2* Z, int L, float2 lam, float2 init, float2* wtot) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = (L + FB_THREADS - 1) / FB_THREADS;
  const int lo = min(L, tid * chunk), hi = min(L, lo + chunk);
  float2 s = make_float2(0.f, 0.f);
  for (int i = lo; i < hi; ++i) s = cfma(lam, s, Z[i]);
  // each thread stands for `chunk` elements: shifting by one thread decays by lam^chunk
  float2 m = cpow_int(lam, chunk);
  const float2 m1 = m;
  float2 v = s;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    float ox = __shfl_up_sync(0xffffffffu, v.x, off), oy = __shfl_up_sync(0xffffffffu, v.y, off);
    if (lane >= off) v = cfma(m, make_float2(ox, oy), v);
    m = cmul(m, m);
  }
  // m is now lam^(32*chunk)
  if (lane == 31) wtot[warp] = v;
  __syncthreads();
  float2 carry = init;                                      // state entering this warp's first element
  for (int w = 0; w < warp; ++w) carry = cfma(m, carry, wtot[w]);
  float ex = __shfl_up_sync(0xffffffffu, v.x, 1), ey = __shfl_up_sync(0xffffffffu, v.y, 1);
  float2 in = (lane > 0) ? make_float2(ex, ey) : make_float2(0.f, 0.f);
  in = cfma(cpow_int(m1, lane), carry, in);
  s = in;
  for (int i = lo; i < hi; ++i) {
    s = cfma(lam, s, Z[i]);
    Z[i] = s;
  }
  __syncthreads();
}

template <typename TIn>
__global__ void __launch_bounds__(FB_THREADS) psk_main_kernel(const PskMainArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int sps = a.sps;
  // ---- which recording / tile -----------------------------------------------------------------
  int lo = 0, hi = a.n_rec;
  const uint32_t tile = blockIdx.x;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (a.tile_first[mid] <= tile) lo = mid; else hi = mid;
  }
  const RecPlan pl = a.plans[lo];
  const int d0 = pl.dl32 + (int)(tile - a.tile_first[lo]) * a.T;
  const int d1 = min(d0 + a.T, pl.dr32);
  const int ns = d1 - d0 + 1;                       // symbols d0 .. d1
  const int64_t N = (int64_t)pl.n;
  // ---- shared memory carve-up -----------------------------------------------------------------
  float* X = smem;                                  // [sps][P]
  float2* taps = reinterpret_cast<float2*>(X + (size_t)sps * a.P);      // [sps][ntp]
  float2* Y = taps + (size_t)sps * a.ntp;           // [T + 4]   slow contribution, then y'
  float2* Z = Y + (a.T + 4);                        // [zcap]    scan scratch
  float2* wtot = Z + a.zcap;                        // [8]
  __shared__ double finit_sh[2 * FB_MAX_SLOW];

  const int ca = d0 - a.lead4;                      // first staged column (global column == symbol index)
  const int ncols = (d1 + a.right) - ca + 1;
  // ---- stage samples: coalesced global reads, de-interleaved stores ----------------------------
  {
    const int64_t n_a = (int64_t)a.n0 + (int64_t)ca * sps;
    const int total = ncols * sps;
    int c = tid / sps, j = tid - c * sps;
    const int dc = FB_THREADS / sps, dj = FB_THREADS - dc * sps;
    for (int idx = tid; idx < total; idx += FB_THREADS) {
      const int64_t n = n_a + idx;
      float v = 0.f;
      if (n >= 0 && n < N) v = load_sample<TIn>(a.samples, pl.off + (uint64_t)n);
      X[j * a.P + c] = v;
      c += dc; j += dj;
      if (j >= sps) { j -= sps; ++c; }
    }
    for (int i = tid; i < sps * a.ntp; i += FB_THREADS) taps[i] = a.taps_r[i];
    for (int i = tid; i < a.T + 4; i += FB_THREADS) Y[i] = make_float2(0.f, 0.f);
  }
  // ---- exact start state of the forward slow recursion at column 0 (left record edge) ----------
  const int fa = max(0, d0 - a.wcols);              // first feature column of the forward recursion
  if (fa == 0 && tid < a.nslow) {
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
  __syncthreads();

  // ---- slow pole pairs: block features -> decaying scan -> residue combine ----------------------
  for (int i = 0; i < a.nslow; ++i) {
    const float2* w = a.slow_w + (size_t)i * (sps + 1);
    const float2 lam = a.lam[i];
    // forward: Fst[c+1] = lam Fst[c] + sum_j p^(sps-j) X[j][c]
    {
      const int L = d1 - fa;                        // features at columns fa .. d1-1
      for (int e = tid; e < L; e += FB_THREADS) {
        const float* col = X + (fa + e - ca);
        float zr = 0.f, zi = 0.f;
        for (int j = 0; j < sps; ++j) {
          const float xv = col[j * a.P];
          const float2 ww = __ldg(&w[sps - j]);
          zr = fmaf(ww.x, xv, zr); zi = fmaf(ww.y, xv, zi);
        }
        Z[e] = make_float2(zr, zi);
      }
      __syncthreads();
      float2 init = make_float2(0.f, 0.f);
      if (fa == 0) init = make_float2((float)finit_sh[2 * i], (float)finit_sh[2 * i + 1]);
      decaying_scan(Z, L, lam, init, wtot);
      // Fst[c] for c in [d0, d1]: c == fa -> init, else Z[c - fa - 1]
      for (int e = tid; e < ns; e += FB_THREADS) {
        const int c = d0 + e;
        const float2 f = (c == fa) ? init : Z[c - fa - 1];
        float2 acc = Y[e];
        acc = cfma(a.rp[i], f, acc);
        acc = cfma(a.rpc[i], make_float2(f.x, -f.y), acc);
        Y[e] = acc;
      }
      __syncthreads();
    }
    // backward: Bfull[c] = sum_j p^j X[j][c] + lam Bfull[c+1];  Bst[c] = Bfull[c] - X[0][c]
    {
      const int fb = d1 + a.wcols;                  // last feature column
      const int L = fb - d0 + 1;
      for (int e = tid; e < L; e += FB_THREADS) {   // e-th element is column fb - e
        const float* col = X + (fb - e - ca);
        float zr = 0.f, zi = 0.f;
        for (int j = 0; j < sps; ++j) {
          const float xv = col[j * a.P];
          const float2 ww = __ldg(&w[j]);
          zr = fmaf(ww.x, xv, zr); zi = fmaf(ww.y, xv, zi);
        }
        Z[e] = make_float2(zr, zi);
      }
      __syncthreads();
      decaying_scan(Z, L, lam, make_float2(0.f, 0.f), wtot);
      for (int e = tid; e < ns; e += FB_THREADS) {
        const int c = d0 + e;
        float2 b = Z[fb - c];
        b.x -= X[c - ca];                           // row 0
        float2 acc = Y[e];
        acc = cfma(a.rm[i], b, acc);
        acc = cfma(a.rmc[i], make_float2(b.x, -b.y), acc);
        Y[e] = acc;
      }
      __syncthreads();
    }
  }

  // ---- fast part: register-tiled polyphase FIR at symbol instants --------------------------------
  {
    const int nchunks = (ns + MAIN_S - 1) / MAIN_S;
    const int chunk = tid % nchunks, g = tid / nchunks;
    if (g < a.rg) {
      float accr[MAIN_S], acci[MAIN_S];
#pragma unroll
      for (int s = 0; s < MAIN_S; ++s) { accr[s] = 0.f; acci[s] = 0.f; }
      // y[s] += tapsR[j][t'] * X[j][base + s + t'],  base = (d0 - dh - ca) + chunk*S  (multiple of 4)
      const int base = (d0 - a.dh - ca) + chunk * MAIN_S;
      for (int j = g; j < sps; j += a.rg) {
        const float4* row = reinterpret_cast<const float4*>(X + (size_t)j * a.P + base);
        const float4* tp = reinterpret_cast<const float4*>(taps + (size_t)j * a.ntp);
        float4 w0 = row[0];
        for (int tq = 0; tq < a.ntp / MAIN_Q; ++tq) {
          const float4 w1 = row[tq + 1];
          const float4 t01 = tp[2 * tq], t23 = tp[2 * tq + 1];     // taps t', t'+1 | t'+2, t'+3 (re,im pairs)
          const float win[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int s = 0; s < MAIN_S; ++s) {
            accr[s] = fmaf(t01.x, win[s], accr[s]);     acci[s] = fmaf(t01.y, win[s], acci[s]);
            accr[s] = fmaf(t01.z, win[s + 1], accr[s]); acci[s] = fmaf(t01.w, win[s + 1], acci[s]);
            accr[s] = fmaf(t23.x, win[s + 2], accr[s]); acci[s] = fmaf(t23.y, win[s + 2], acci[s]);
            accr[s] = fmaf(t23.z, win[s + 3], accr[s]); acci[s] = fmaf(t23.w, win[s + 3], acci[s]);
          }
          w0 = w1;
        }
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

  // ---- band-pass forward (scipy lfilter, direct form II transposed; a[0] == 1) ------------------
  {
    double z[8];
    const double x0 = x_ext<TIn>(a.samples, pl.off, N, jb.wa);
    const bool exact = (jb.wa == -(int64_t)d.pad_bp);
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = exact ? d.bp_zi[i] * x0 : 0.0;
    for (int64_t i = 0; i < Lb; ++i) {
      const double xv = x_ext<TIn>(a.samples, pl.off, N, jb.wa + i);
      const double y = d.bp_b[0] * xv + z[0];
#pragma unroll
      for (int k = 0; k < 7; ++k) z[k] = d.bp_b[k + 1] * xv + z[k + 1] - d.bp_a[k + 1] * y;
      z[7] = d.bp_b[8] * xv - d.bp_a[8] * y;
      A[i] = y;
    }
  }
  // ---- band-pass backward, in place ----------------------------------------------------------------
  {
    double z[8];
    const bool exact = (jb.wb == N - 1 + (int64_t)d.pad_bp);
    const double y0 = A[Lb - 1];
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = exact ? d.bp_zi[i] * y0 : 0.0;
    for (int64_t i = Lb - 1; i >= 0; --i) {
      const double xv = A[i];
      const double y = d.bp_b[0] * xv + z[0];
#pragma unroll
      for (int k = 0; k < 7; ++k) z[k] = d.bp_b[k + 1] * xv + z[k + 1] - d.bp_a[k + 1] * y;
      z[7] = d.bp_b[8] * xv - d.bp_a[8] * y;
      A[i] = y;
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
    double zr[4], zi[4];
    double x0r, x0i;
    mixed_ext(jb.la, x0r, x0i);
    const bool exact = (jb.la == -(int64_t)d.pad_lp);
#pragma unroll
    for (int i = 0; i < 4; ++i) { zr[i] = exact ? d.lp_zi[i] * x0r : 0.0; zi[i] = exact ? d.lp_zi[i] * x0i : 0.0; }
    for (int64_t i = 0; i < Ll; ++i) {
      double xr, xi;
      mixed_ext(jb.la + i, xr, xi);
      const double yr = d.lp_b[0] * xr + zr[0], yi = d.lp_b[0] * xi + zi[0];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        zr[k] = d.lp_b[k + 1] * xr + zr[k + 1] - d.lp_a[k + 1] * yr;
        zi[k] = d.lp_b[k + 1] * xi + zi[k + 1] - d.lp_a[k + 1] * yi;
      }
      zr[3] = d.lp_b[4] * xr - d.lp_a[4] * yr;
      zi[3] = d.lp_b[4] * xi - d.lp_a[4] * yi;
      U[2 * i] = yr; U[2 * i + 1] = yi;
    }
  }
  // ---- low-pass backward, in place ---------------------------------------------------------------------
  {
    double zr[4], zi[4];
    const bool exact = (jb.lb == N - 1 + (int64_t)d.pad_lp);
    const double y0r = U[2 * (Ll - 1)], y0i = U[2 * (Ll - 1) + 1];
#pragma unroll
    for (int i = 0; i < 4; ++i) { zr[i] = exact ? d.lp_zi[i] * y0r : 0.0; zi[i] = exact ? d.lp_zi[i] * y0i : 0.0; }
    for (int64_t i = Ll - 1; i >= 0; --i) {
      const double xr = U[2 * i], xi = U[2 * i + 1];
      const double yr = d.lp_b[0] * xr + zr[0], yi = d.lp_b[0] * xi + zi[0];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        zr[k] = d.lp_b[k + 1] * xr + zr[k + 1] - d.lp_a[k + 1] * yr;
        zi[k] = d.lp_b[k + 1] * xi + zi[k + 1] - d.lp_a[k + 1] * yi;
      }
      zr[3] = d.lp_b[4] * xr - d.lp_a[4] * yr;
      zi[3] = d.lp_b[4] * xi - d.lp_a[4] * yi;
      U[2 * i] = yr; U[2 * i + 1] = yi;
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
    psk_main_kernel<TIn><<<n_tiles, FB_THREADS, smem, h->stream>>>(ma);
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
  FB_CUDA(h, cudaSetDevice(h->device));
  if (n_rec == 0) return FB_OK;
  const size_t esz = dtype == FB_F32 ? 4 : dtype == FB_F64 ? 8 : 2;
  const int bps = d.bits_per_sym, dper = 32 / bps;

  // ---- tile geometry of the main kernel -------------------------------------------------------------
  int T = 0, P = 0, lead4 = 0, rg = 1, ntp = 0, zcap = 0, right = 0;
  size_t smem = 0;
  if (!d.emulate_only) {
    ntp = (d.nt + MAIN_Q - 1) / MAIN_Q * MAIN_Q;
    const int lead = std::max(d.dh, d.wcols);
    lead4 = d.dh + (lead - d.dh + 3) / 4 * 4;
    // taps are read t' = 0 .. ntp-1 from column base (+ one prefetched float4 window); the zero-padded taps and the
    // last partial symbol chunk reach further right than dl: those columns must hold staged values, not garbage
    right = std::max(d.dl + (ntp - d.nt) + MAIN_S + 2 * MAIN_Q, d.wcols + 1);
    const size_t budget = 72 * 1024;
    for (T = (FB_THREADS * MAIN_S - 1) / 32 * 32; T >= 32; T -= 32) {
      const int ncols = T + 1 + lead4 + right;
      P = (ncols + 3) / 4 * 4 + 4;
      zcap = T + d.wcols + 8;
      smem = (size_t)d.sps * P * 4 + (size_t)d.sps * ntp * 8 + (size_t)(T + 4) * 8 + (size_t)zcap * 8 + 8 * 8;
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
    // reversed, zero-padded tap table: taps_r[j][t'] = taps[j][nt-1-t'], t' < nt
    std::vector<float> tr((size_t)d.sps * ntp * 2, 0.f);
    for (int j = 0; j < d.sps; ++j)
      for (int t = 0; t < d.nt; ++t) {
        tr[((size_t)j * ntp + t) * 2] = taps[((size_t)j * d.nt + (d.nt - 1 - t)) * 2];
        tr[((size_t)j * ntp + t) * 2 + 1] = taps[((size_t)j * d.nt + (d.nt - 1 - t)) * 2 + 1];
      }
    if ((rc = fb_ensure(h, h->taps, tr.size() * 4))) return rc;
    FB_CUDA(h, cudaMemcpyAsync(h->taps.p, tr.data(), tr.size() * 4, cudaMemcpyHostToDevice, h->stream));
    const size_t swb = (size_t)std::max(1, d.nslow) * (d.sps + 1) * 8;
    if ((rc = fb_ensure(h, h->slow_w, swb))) return rc;
    if (d.nslow) FB_CUDA(h, cudaMemcpyAsync(h->slow_w.p, slow_w, (size_t)d.nslow * (d.sps + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    // the std::vector staging above is pageable: cudaMemcpyAsync has copied it out before returning
    ma.samples = d_samples; ma.plans = (const RecPlan*)h->plans.p; ma.tile_first = (const uint32_t*)h->tile_first.p;
    ma.taps_r = (const float2*)h->taps.p; ma.slow_w = (const float2*)h->slow_w.p; ma.bits = (uint32_t*)h->bits.p;
    ma.n_rec = n_rec; ma.sps = d.sps; ma.n0 = d.n0; ma.bps = bps; ma.nt = d.nt; ma.ntp = ntp; ma.dl = d.dl; ma.dh = d.dh;
    ma.nslow = d.nslow; ma.wcols = d.wcols; ma.pad_bp = d.pad_bp; ma.T = T; ma.P = P; ma.lead4 = lead4; ma.rg = rg; ma.zcap = zcap; ma.right = right;
    ma.rho = make_float2(d.rho[0], d.rho[1]);
    for (int i = 0; i < FB_MAX_SLOW; ++i) {
      ma.lam[i] = make_float2(d.slow_lam[2 * i], d.slow_lam[2 * i + 1]);
      ma.rp[i] = make_float2(d.slow_rp[2 * i], d.slow_rp[2 * i + 1]);
      ma.rpc[i] = make_float2(d.slow_rpc[2 * i], d.slow_rpc[2 * i + 1]);
      ma.rm[i] = make_float2(d.slow_rm[2 * i], d.slow_rm[2 * i + 1]);
      ma.rmc[i] = make_float2(d.slow_rmc[2 * i], d.slow_rmc[2 * i + 1]);
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
