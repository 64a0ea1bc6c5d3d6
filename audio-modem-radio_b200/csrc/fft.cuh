// Hand-written float64 FFT for the two whole-record transforms of the path (no library on it):
//   * scipy.signal.hilbert inside modem.py:306-312 (fsk_demodulate's envelope): rfft-like forward, one-sided mask, inverse;
//   * scipy.signal.resample inside decoder.py:385-387 (WAV files that are not at 96 kHz): forward, crop / pad, inverse.
// Both are defined by a length-N DFT of the WHOLE record (N is whatever the file holds: 17 280 000 = 2^10 3^3 5^4 for three
// minutes at 96 kHz), so the transform has to take any N:
//   * N = 2^a 3^b 5^c 7^d: out-of-place Stockham autosort passes (decimation in frequency), radix 8 / 4 / 2 / 3 / 5 / 7,
//     one pass = one streaming read + one streaming write of the N complex128 points, twiddles by sincospi on exact
//     integer fractions (no table, no accumulated rotation: relative error ~ log2(N) ulp);
//   * any other N: Bluestein's chirp-z over a power-of-two length M >= 2N - 1 with the same pass kernels, chirp phases from
//     n^2 mod 2N in 64-bit integers.
// Real input / output goes through the complex transform (pack / Hermitian expand kernels): these transforms run once per
// recording and are HBM-bound; the factor two is not where their time goes.
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

// One butterfly of a radix-R Stockham DIF pass.  n: current transform length, s: stride (product of the radices already
// applied), m = n / R;  butterfly t = p * s + q, p < m, q < s:
//   a_k = x[q + s (p + k m)];   y[q + s (R p + j)] = (sum_k a_k W_R^(jk)) * exp(sign 2 pi i p j / n)
template <int R>
__host__ __device__ __forceinline__ void butterfly(const double2* __restrict__ x, double2* __restrict__ y, int64_t n, int64_t s, int64_t t, int sign) {
  const int64_t m = n / R, p = t / s, q = t - p * s;
  double2 a[R], b[R];
#pragma unroll
  for (int k = 0; k < R; ++k) a[k] = x[q + s * (p + k * m)];
  if (R == 2) {
    b[0] = cadd(a[0], a[1]); b[1] = csub(a[0], a[1]);
  } else if (R == 4) {
    const double2 e0 = cadd(a[0], a[2]), e1 = csub(a[0], a[2]), o0 = cadd(a[1], a[3]), o1 = csub(a[1], a[3]);
    const double2 jo1 = sign < 0 ? make_double2(o1.y, -o1.x) : make_double2(-o1.y, o1.x);      // (sign i) * o1
    b[0] = cadd(e0, o0); b[2] = csub(e0, o0); b[1] = cadd(e1, jo1); b[3] = csub(e1, jo1);
  } else {
    double2 w[R];                                        // W_R^k = exp(sign 2 pi i k / R)
#pragma unroll
    for (int k = 0; k < R; ++k) w[k] = unit(k, R, sign);
#pragma unroll
    for (int j = 0; j < R; ++j) {
      double2 acc = a[0];
#pragma unroll
      for (int k = 1; k < R; ++k) acc = cadd(acc, cmul(a[k], w[(j * k) % R]));
      b[j] = acc;
    }
  }
  y[q + s * (R * p)] = b[0];
#pragma unroll
  for (int j = 1; j < R; ++j) y[q + s * (R * p + j)] = cmul(b[j], unit(p * j, n, sign));
}

// Bluestein chirp w(i) = exp(sign * pi * i * i^2 / n), the phase from i^2 mod 2n in exact integers (i < 2^31)
__host__ __device__ __forceinline__ double2 chirp(int64_t i, int64_t n, int sign) {
  const uint64_t r = ((uint64_t)i * (uint64_t)i) % (uint64_t)(2 * n);
  double s, c;
  sincospi((double)r / (double)n, &s, &c);
  return make_double2(c, sign < 0 ? -s : s);
}

// radices of a {2,3,5,7}-smooth length (8s and 4s first: fewer passes); empty when n has another prime factor
inline std::vector<int> smooth_radices(int64_t n) {
  std::vector<int> r;
  if (n < 1) return r;
  int twos = 0;
  while (n % 2 == 0) { n /= 2; ++twos; }
  while (twos >= 3) { r.push_back(8); twos -= 3; }
  if (twos == 2) r.push_back(4);
  if (twos == 1) r.push_back(2);
  for (int p : {3, 5, 7}) while (n % p == 0) { n /= p; r.push_back(p); }
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
