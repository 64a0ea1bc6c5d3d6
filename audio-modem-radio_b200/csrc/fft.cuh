// Hand-written float64 FFT for the two whole-record transforms of the path (no library on it):
//   * scipy.signal.hilbert inside modem.py:306-312 (fsk_demodulate's envelope): rfft-like forward, one-sided mask, inverse;
//   * scipy.signal.resample inside decoder.py:385-387 (WAV files that are not at 96 kHz): forward, crop / pad, inverse.
// Both are defined by a length-N DFT of the WHOLE record (N is whatever the file holds: 17 280 000 = 2^10 3^3 5^4 for three
// minutes at 96 kHz), so the transform has to take any N:
//   * N = 2^a 3^b 5^c 7^d: out-of-place Stockham autosort passes (decimation in frequency), radix 32 / 16 / 8 / 4 / 2 / 27 / 9 / 3 / 25 / 5 / 7,
//     one pass = one streaming read + one streaming write of the N complex128 points, twiddles by sincospi on exact
//     integer fractions (no table, no accumulated rotation: relative error ~ log2(N) ulp);
//   * any other N: Bluestein's chirp-z over a power-of-two length M >= 2N - 1 with the same pass kernels, chirp phases from
//     n^2 mod 2N in 64-bit integers.
// Real input / output of even length goes through a complex transform of HALF the length (z[j] = x[2j] + i x[2j+1] and the
// usual untangling pass); odd lengths take the full-length complex transform (pack / Hermitian expand kernels).
#pragma once
#include "common.cuh"

#include <vector>

namespace fbfft {

__host__ __device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__host__ __device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// exp(sign * 2 pi i * num / den), num < den exact integers
__host__ __device__ __forceinline__ double2 unit(int64_t num, int64_t den, int sign) {
  double s, c;
  sincospi(2.0 * ((double)num / (double)den), &s, &c);
  return make_double2(c, sign < 0 ? -s : s);
}

// In-place DFT of P points held in registers, v[i] -> sum_i v[i] W_P^(ik);  w: roots table of a multiple R of P (W_R^e),
// ws = R / P.  P = 2 and 4 need no multiplications.
template <int P>
__host__ __device__ __forceinline__ void dft_small(double2 (&v)[P], const double2* __restrict__ w, int ws, int sign) {
  if (P == 1) return;
  if (P == 2) {
    const double2 t = csub(v[0], v[1]);
    v[0] = cadd(v[0], v[1]); v[1] = t;
  } else if (P == 4) {
    const double2 e0 = cadd(v[0], v[2]), e1 = csub(v[0], v[2]), o0 = cadd(v[1], v[3]), o1 = csub(v[1], v[3]);
    const double2 jo1 = sign < 0 ? make_double2(o1.y, -o1.x) : make_double2(-o1.y, o1.x);      // (sign i) * o1
    v[0] = cadd(e0, o0); v[2] = csub(e0, o0); v[1] = cadd(e1, jo1); v[3] = csub(e1, jo1);
  } else {
    double2 o[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      double2 acc = v[0];
#pragma unroll
      for (int k = 1; k < P; ++k) {
        const int e = (j * k) % P;
        if (e == 0) acc = cadd(acc, v[k]); else acc = cadd(acc, cmul(v[k], w[e * ws]));
      }
      o[j] = acc;
    }
#pragma unroll
    for (int j = 0; j < P; ++j) v[j] = o[j];
  }
}

// One butterfly of a radix-R Stockham DIF pass, R = R1 * R2.  n: current transform length, s: stride (product of the radices
// already applied), m = n / R;  butterfly t = p * s + q, p < m, q < s:
//   a_i = x[q + s (p + i m)];   y[q + s (R p + k)] = (sum_i a_i W_R^(ik)) * exp(sign 2 pi i p k / n)
// The R-point DFT is done in registers as R2 DFTs of R1 points, the twiddles W_R^(k1 i2), and R1 DFTs of R2 points
// (i = i1 R2 + i2, k = k1 + R1 k2).  w: the R roots W_R^e (shared memory on the device: every lane reads the same entry).
template <int R1, int R2>
__host__ __device__ __forceinline__ void butterfly(const double2* __restrict__ x, double2* __restrict__ y, int64_t n, int64_t s, int64_t t, int sign,
                                                   const double2* __restrict__ w) {
  constexpr int R = R1 * R2;
  const int64_t m = n / R, p = t / s, q = t - p * s;
  double2 a[R];
#pragma unroll
  for (int i = 0; i < R; ++i) a[i] = x[q + s * (p + i * m)];
#pragma unroll
  for (int i2 = 0; i2 < R2; ++i2) {
    double2 v[R1];
#pragma unroll
    for (int i1 = 0; i1 < R1; ++i1) v[i1] = a[i1 * R2 + i2];
    dft_small<R1>(v, w, R2, sign);
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) a[k1 * R2 + i2] = (k1 * i2) % R == 0 ? v[k1] : cmul(v[k1], w[(k1 * i2) % R]);
  }
  // pass twiddles exp(sign 2 pi i p k / n), k = k1 + R1 k2: one sincospi per k1 and one for the step exp(sign 2 pi i p R1 / n),
  // the powers along k2 by multiplication (at most R2 - 1 <= 8 steps: a few ulp; a sincospi per output made the composite
  // passes compute-bound)
  double2* yo = y + q + s * (R * p);
  const double2 stepw = R2 > 1 ? unit(p * R1, n, sign) : make_double2(1.0, 0.0);
#pragma unroll
  for (int k1 = 0; k1 < R1; ++k1) {
    double2 v[R2];
#pragma unroll
    for (int i2 = 0; i2 < R2; ++i2) v[i2] = a[k1 * R2 + i2];
    dft_small<R2>(v, w, R1, sign);
    double2 tw = k1 == 0 ? make_double2(1.0, 0.0) : unit(p * k1, n, sign);
#pragma unroll
    for (int k2 = 0; k2 < R2; ++k2) {
      const int k = k1 + R1 * k2;
      yo[k * s] = k == 0 ? v[k2] : cmul(v[k2], tw);
      if (k2 + 1 < R2) tw = cmul(tw, stepw);
    }
  }
}

// the R-th roots of unity a pass needs (R <= FFT_MAX_RADIX)
constexpr int FFT_MAX_RADIX = 32;
__host__ __device__ __forceinline__ void roots(double2* w, int R, int k, int sign) { w[k] = unit(k, R, sign); }

// Bluestein chirp w(i) = exp(sign * pi * i * i^2 / n), the phase from i^2 mod 2n in exact integers (i < 2^31)
__host__ __device__ __forceinline__ double2 chirp(int64_t i, int64_t n, int sign) {
  const uint64_t r = ((uint64_t)i * (uint64_t)i) % (uint64_t)(2 * n);
  double s, c;
  sincospi((double)r / (double)n, &s, &c);
  return make_double2(c, sign < 0 ? -s : s);
}

// radices of a {2,3,5,7}-smooth length (composite radices 32 / 16 / 27 / 9 / 25 first: a pass is one read and one write of all the
// points, so fewer passes; the radix-R butterfly is a direct R x R product); empty when n has another prime factor
inline std::vector<int> smooth_radices(int64_t n) {
  std::vector<int> r;
  if (n < 1) return r;
  int twos = 0, threes = 0, fives = 0;
  while (n % 2 == 0) { n /= 2; ++twos; }
  while (n % 3 == 0) { n /= 3; ++threes; }
  while (n % 5 == 0) { n /= 5; ++fives; }
  while (twos >= 5) { r.push_back(32); twos -= 5; }
  if (twos == 4) r.push_back(16);
  if (twos == 3) r.push_back(8);
  if (twos == 2) r.push_back(4);
  if (twos == 1) r.push_back(2);
  while (threes >= 3) { r.push_back(27); threes -= 3; }
  if (threes == 2) r.push_back(9);
  if (threes == 1) r.push_back(3);
  while (fives >= 2) { r.push_back(25); fives -= 2; }
  if (fives) r.push_back(5);
  while (n % 7 == 0) { n /= 7; r.push_back(7); }
  if (n != 1) r.clear();
  return r;
}

inline int64_t next_pow2(int64_t v) { int64_t m = 1; while (m < v) m <<= 1; return m; }

}  // namespace fbfft

// fft.cu -- all on h->stream, workspace on the handle (h->fftws), results written before the call returns control to the stream
// X[0 .. n/2] = sum_j x[j] exp(-2 pi i j k / n)                         (numpy.fft.rfft)
int fb_fft_d2z(fb_handle* h, const double* d_x, double2* d_X, int64_t n);
// y[j] = X[0] + 2 Re sum_{0<k<n/2} X[k] exp(+2 pi i j k / n) (+ Re X[n/2] (-1)^j)    (numpy.fft.irfft(X, n) * n: unnormalised)
int fb_fft_z2d(fb_handle* h, const double2* d_X, double* d_y, int64_t n);
// debug / test hook: full complex transform of host data on the device, sign -1 forward / +1 inverse (unnormalised)
extern "C" int fb_debug_fft_c2c(fb_handle* h, const double* host_in_ri, double* host_out_ri, int64_t n, int sign);
