// Hand-written float64 FFT (see fft.cuh): Stockham pass kernels, Bluestein for lengths that are not {2,3,5,7}-smooth, and the
// real-input / real-output wrappers the Hilbert envelope (fsk_v2.cu) and the resampler (resample.cu) call.
#include "fft.cuh"

namespace {

using fbfft::butterfly;

template <int R1, int R2>
__global__ void __launch_bounds__(FB_THREADS) fft_pass_kernel(const double2* __restrict__ x, double2* __restrict__ y, int64_t n, int64_t s,
                                                              int64_t count, int sign) {
  constexpr int R = R1 * R2;
  __shared__ double2 w[R];
  if (threadIdx.x < R) fbfft::roots(w, R, (int)threadIdx.x, sign);
  __syncthreads();
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x)
    butterfly<R1, R2>(x, y, n, s, t, sign, w);
}

__global__ void __launch_bounds__(FB_THREADS) pack_real_kernel(const double* __restrict__ x, double2* __restrict__ c, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) c[i] = make_double2(x[i], 0.0);
}
__global__ void __launch_bounds__(FB_THREADS) copy_c_kernel(const double2* __restrict__ c, double2* __restrict__ X, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) X[i] = c[i];
}
// full spectrum of a real sequence from its first n/2 + 1 bins; the imaginary parts of bin 0 and (n even) bin n/2 are ignored,
// as numpy.fft.irfft / cuFFT Z2D do
__global__ void __launch_bounds__(FB_THREADS) hermitian_kernel(const double2* __restrict__ X, double2* __restrict__ c, int64_t n) {
  const int64_t nh = n / 2;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k <= nh; k += (int64_t)gridDim.x * blockDim.x) {
    double2 v = X[k];
    if (k == 0 || 2 * k == n) v.y = 0.0;
    c[k] = v;
    if (k != 0 && 2 * k != n) c[n - k] = make_double2(v.x, -v.y);
  }
}
__global__ void __launch_bounds__(FB_THREADS) take_real_kernel(const double2* __restrict__ c, double* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = c[i].x;
}

// ---- even n: the real transform through a complex one of half the length (z[j] = x[2j] + i x[2j+1]) --------------------------
__global__ void __launch_bounds__(FB_THREADS) pack_pairs_kernel(const double* __restrict__ x, double2* __restrict__ z, int64_t h) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < h; j += (int64_t)gridDim.x * blockDim.x) z[j] = make_double2(x[2 * j], x[2 * j + 1]);
}
__global__ void __launch_bounds__(FB_THREADS) unpack_pairs_kernel(const double2* __restrict__ z, double* __restrict__ y, int64_t h) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < h; j += (int64_t)gridDim.x * blockDim.x) { const double2 v = z[j]; y[2 * j] = v.x; y[2 * j + 1] = v.y; }
}
// X[k] = (A + B) / 2 - (i / 2) w (A - B),  A = Z[k mod h], B = conj(Z[(h - k) mod h]), w = exp(-2 pi i k / n),  k = 0 .. h
__global__ void __launch_bounds__(FB_THREADS) r2c_post_kernel(const double2* __restrict__ Z, double2* __restrict__ X, int64_t h) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k <= h; k += (int64_t)gridDim.x * blockDim.x) {
    const double2 A = Z[k == h ? 0 : k], Bc = Z[(k == 0 || k == h) ? 0 : h - k];
    const double2 B = make_double2(Bc.x, -Bc.y);
    const double2 sm = fbfft::cadd(A, B), df = fbfft::csub(A, B);
    const double2 w = k == h ? make_double2(-1.0, 0.0) : fbfft::unit(k, 2 * h, -1);
    const double2 t = fbfft::cmul(w, df);                            // -(i/2) t = (t.y, -t.x) / 2
    X[k] = make_double2(0.5 * (sm.x + t.y), 0.5 * (sm.y - t.x));
  }
}
// Z'[k] = (A + B) + i w (A - B),  A = X[k], B = conj(X[h - k]), w = exp(+2 pi i k / n),  k = 0 .. h-1;  IDFT_h(Z')[j] = y[2j] + i y[2j+1]
// (the imaginary parts of X[0] and X[h] are ignored, as numpy.fft.irfft does)
__global__ void __launch_bounds__(FB_THREADS) c2r_pre_kernel(const double2* __restrict__ X, double2* __restrict__ Zp, int64_t h) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < h; k += (int64_t)gridDim.x * blockDim.x) {
    double2 A = X[k], Bc = X[h - k];
    if (k == 0) { A.y = 0.0; Bc.y = 0.0; }
    const double2 B = make_double2(Bc.x, -Bc.y);
    const double2 sm = fbfft::cadd(A, B), df = fbfft::csub(A, B);
    const double2 t = fbfft::cmul(fbfft::unit(k, 2 * h, +1), df);    // i t = (-t.y, t.x)
    Zp[k] = make_double2(sm.x - t.y, sm.y + t.x);
  }
}

// ---- Bluestein:  X[k] = w(k) * sum_j (x[j] w(j)) conj(w(k - j)),   w(i) = exp(sign pi i i^2 / n) -------------------------
__global__ void __launch_bounds__(FB_THREADS) blu_in_kernel(const double2* __restrict__ x, double2* __restrict__ a, int64_t n, int64_t M, int sign) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x)
    a[i] = i < n ? fbfft::cmul(x[i], fbfft::chirp(i, n, sign)) : make_double2(0.0, 0.0);
}
__global__ void __launch_bounds__(FB_THREADS) blu_chirp_kernel(double2* __restrict__ b, int64_t n, int64_t M, int sign) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t d = i < n ? i : (M - i < n ? M - i : -1);          // |lag| of slot i, or outside the kernel's support
    double2 v = make_double2(0.0, 0.0);
    if (d >= 0) { v = fbfft::chirp(d, n, sign); v.y = -v.y; }
    b[i] = v;
  }
}
__global__ void __launch_bounds__(FB_THREADS) blu_mul_kernel(double2* __restrict__ a, const double2* __restrict__ b, int64_t M) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) a[i] = fbfft::cmul(a[i], b[i]);
}
__global__ void __launch_bounds__(FB_THREADS) blu_out_kernel(const double2* __restrict__ conv, double2* __restrict__ X, int64_t n, int64_t M, int sign) {
  const double inv = 1.0 / (double)M;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const double2 v = fbfft::cmul(conv[k], fbfft::chirp(k, n, sign));
    X[k] = make_double2(v.x * inv, v.y * inv);
  }
}

inline int grid_for(int64_t count) { return (int)std::max<int64_t>(1, std::min<int64_t>(148 * 8, (count + FB_THREADS - 1) / FB_THREADS)); }

// Stockham passes over a {2,3,5,7}-smooth length: ping-pong between a and b, *res = the buffer that holds the result
int c2c_smooth(fb_handle* h, double2* a, double2* b, int64_t n, int sign, const std::vector<int>& radices, double2** res) {
  int64_t cur = n, s = 1;
  double2 *x = a, *y = b;
  for (int r : radices) {
    const int64_t count = n / r;
    const int g = grid_for(count);
    switch (r) {
#define FB_FFT_PASS(R1, R2) case R1 * R2: fft_pass_kernel<R1, R2><<<g, FB_THREADS, 0, h->stream>>>(x, y, cur, s, count, sign); break
      FB_FFT_PASS(4, 8); FB_FFT_PASS(3, 9); FB_FFT_PASS(4, 4); FB_FFT_PASS(4, 2); FB_FFT_PASS(4, 1); FB_FFT_PASS(2, 1); FB_FFT_PASS(3, 3); FB_FFT_PASS(3, 1); FB_FFT_PASS(5, 5); FB_FFT_PASS(5, 1);
      FB_FFT_PASS(7, 1);
#undef FB_FFT_PASS
      default: return FB_EINVAL;
    }
    h->launches++;
    cur /= r; s *= r;
    std::swap(x, y);
  }
  FB_CUDA(h, cudaGetLastError());
  *res = x;
  return FB_OK;
}

// Complex transform of the n points in ws[0 .. n); workspace ws holds fb_fft_ws_points(n) points.  *res points into ws.
int c2c(fb_handle* h, double2* ws, int64_t n, int sign, double2** res) {
  if (n == 1) { *res = ws; return FB_OK; }
  const std::vector<int> rad = fbfft::smooth_radices(n);
  if (!rad.empty()) return c2c_smooth(h, ws, ws + n, n, sign, rad, res);
  // Bluestein: [x (n)] [a (M)] [b (M)] [t (M)]
  const int64_t M = fbfft::next_pow2(2 * n - 1);
  const std::vector<int> rm = fbfft::smooth_radices(M);
  double2 *x = ws, *a = ws + n, *b = a + M, *t = b + M, *fa, *fb, *cv;
  const int g = grid_for(M);
  blu_in_kernel<<<g, FB_THREADS, 0, h->stream>>>(x, a, n, M, sign);
  h->launches++;
  int rc = c2c_smooth(h, a, t, M, -1, rm, &fa);                       // fa is a or t; the other one is free
  if (rc) return rc;
  double2* spare = fa == a ? t : a;
  blu_chirp_kernel<<<g, FB_THREADS, 0, h->stream>>>(b, n, M, sign);
  h->launches++;
  if ((rc = c2c_smooth(h, b, spare, M, -1, rm, &fb))) return rc;      // fb is b or spare
  blu_mul_kernel<<<g, FB_THREADS, 0, h->stream>>>(fa, fb, M);
  h->launches++;
  double2* spare2 = fb == b ? spare : b;
  if ((rc = c2c_smooth(h, fa, spare2, M, +1, rm, &cv))) return rc;
  double2* out = cv == fa ? spare2 : fa;                               // any M-point buffer that is not cv
  blu_out_kernel<<<grid_for(n), FB_THREADS, 0, h->stream>>>(cv, out, n, M, sign);
  h->launches++;
  FB_CUDA(h, cudaGetLastError());
  *res = out;
  return FB_OK;
}

int64_t ws_points(int64_t n) {
  if (!fbfft::smooth_radices(n).empty() || n == 1) return 2 * n;
  return n + 3 * fbfft::next_pow2(2 * n - 1);
}

}  // namespace

int fb_fft_d2z(fb_handle* h, const double* d_x, double2* d_X, int64_t n) {
  if (n < 1) return FB_EINVAL;
  const bool half = n % 2 == 0;
  const int64_t nc = half ? n / 2 : n;
  int rc = fb_ensure(h, h->fftws, (size_t)ws_points(nc) * sizeof(double2));
  if (rc) return rc;
  double2* ws = (double2*)h->fftws.p;
  if (half) pack_pairs_kernel<<<grid_for(nc), FB_THREADS, 0, h->stream>>>(d_x, ws, nc);
  else pack_real_kernel<<<grid_for(n), FB_THREADS, 0, h->stream>>>(d_x, ws, n);
  double2* res;
  if ((rc = c2c(h, ws, nc, -1, &res))) return rc;
  if (half) r2c_post_kernel<<<grid_for(nc + 1), FB_THREADS, 0, h->stream>>>(res, d_X, nc);
  else copy_c_kernel<<<grid_for(n / 2 + 1), FB_THREADS, 0, h->stream>>>(res, d_X, n / 2 + 1);
  h->launches += 2;
  FB_CUDA(h, cudaGetLastError());
  return FB_OK;
}

int fb_fft_z2d(fb_handle* h, const double2* d_X, double* d_y, int64_t n) {
  if (n < 1) return FB_EINVAL;
  const bool half = n % 2 == 0;
  const int64_t nc = half ? n / 2 : n;
  int rc = fb_ensure(h, h->fftws, (size_t)ws_points(nc) * sizeof(double2));
  if (rc) return rc;
  double2* ws = (double2*)h->fftws.p;
  if (half) c2r_pre_kernel<<<grid_for(nc), FB_THREADS, 0, h->stream>>>(d_X, ws, nc);
  else hermitian_kernel<<<grid_for(n / 2 + 1), FB_THREADS, 0, h->stream>>>(d_X, ws, n);
  double2* res;
  if ((rc = c2c(h, ws, nc, +1, &res))) return rc;
  if (half) unpack_pairs_kernel<<<grid_for(nc), FB_THREADS, 0, h->stream>>>(res, d_y, nc);
  else take_real_kernel<<<grid_for(n), FB_THREADS, 0, h->stream>>>(res, d_y, n);
  h->launches += 2;
  FB_CUDA(h, cudaGetLastError());
  return FB_OK;
}

extern "C" int fb_debug_fft_c2c(fb_handle* h, const double* host_in_ri, double* host_out_ri, int64_t n, int sign) {
  if (!h || !host_in_ri || !host_out_ri || n < 1 || (sign != 1 && sign != -1)) return FB_EINVAL;
  FB_LOCK(h);
  FB_CUDA(h, cudaSetDevice(h->device));
  int rc = fb_ensure(h, h->fftws, (size_t)ws_points(n) * sizeof(double2));
  if (rc) return rc;
  double2* ws = (double2*)h->fftws.p;
  FB_CUDA(h, cudaMemcpyAsync(ws, host_in_ri, (size_t)n * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
  double2* res;
  if ((rc = c2c(h, ws, n, sign, &res))) return rc;
  FB_CUDA(h, cudaMemcpyAsync(host_out_ri, res, (size_t)n * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
  FB_CUDA(h, cudaStreamSynchronize(h->stream));
  return FB_OK;
}

// ---- host execution of the same butterflies (test hook, no GPU needed): tests/test_fft_host.py checks the index math, the
// radix schedule and the Bluestein chirps against numpy on the CPU ----------------------------------------------------------
namespace {
template <int R1, int R2>
void host_pass(const double2* x, double2* y, int64_t n_total, int64_t cur, int64_t s, int sign) {
  constexpr int R = R1 * R2;
  double2 w[R];
  for (int k = 0; k < R; ++k) fbfft::roots(w, R, k, sign);
  for (int64_t t = 0; t < n_total / R; ++t) butterfly<R1, R2>(x, y, cur, s, t, sign, w);
}
std::vector<double2> host_smooth(std::vector<double2> a, int sign) {
  const int64_t n = (int64_t)a.size();
  const std::vector<int> rad = fbfft::smooth_radices(n);
  std::vector<double2> b(a.size());
  double2 *x = a.data(), *y = b.data();
  int64_t cur = n, s = 1;
  for (int r : rad) {
    switch (r) {
#define FB_HOST_PASS(R1, R2) case R1 * R2: host_pass<R1, R2>(x, y, n, cur, s, sign); break
      FB_HOST_PASS(4, 8); FB_HOST_PASS(3, 9); FB_HOST_PASS(4, 4); FB_HOST_PASS(4, 2); FB_HOST_PASS(4, 1); FB_HOST_PASS(2, 1); FB_HOST_PASS(3, 3); FB_HOST_PASS(3, 1); FB_HOST_PASS(5, 5); FB_HOST_PASS(5, 1);
      FB_HOST_PASS(7, 1);
#undef FB_HOST_PASS
      default: return {};
    }
    cur /= r; s *= r;
    std::swap(x, y);
  }
  return std::vector<double2>(x, x + n);
}
}  // namespace

extern "C" int fb_debug_fft_host(const double* in_ri, double* out_ri, int64_t n, int sign) {
  if (!in_ri || !out_ri || n < 1 || (sign != 1 && sign != -1)) return FB_EINVAL;
  const double2* in = reinterpret_cast<const double2*>(in_ri);
  std::vector<double2> res;
  if (n == 1) res.assign(in, in + 1);
  else if (!fbfft::smooth_radices(n).empty()) res = host_smooth(std::vector<double2>(in, in + n), sign);
  else {
    const int64_t M = fbfft::next_pow2(2 * n - 1);
    std::vector<double2> a(M), b(M);
    for (int64_t i = 0; i < M; ++i) {
      a[i] = i < n ? fbfft::cmul(in[i], fbfft::chirp(i, n, sign)) : make_double2(0.0, 0.0);
      const int64_t d = i < n ? i : (M - i < n ? M - i : -1);
      double2 v = make_double2(0.0, 0.0);
      if (d >= 0) { v = fbfft::chirp(d, n, sign); v.y = -v.y; }
      b[i] = v;
    }
    std::vector<double2> fa = host_smooth(a, -1), fb = host_smooth(b, -1);
    for (int64_t i = 0; i < M; ++i) fa[i] = fbfft::cmul(fa[i], fb[i]);
    const std::vector<double2> cv = host_smooth(fa, +1);
    res.resize(n);
    for (int64_t k = 0; k < n; ++k) { const double2 v = fbfft::cmul(cv[k], fbfft::chirp(k, n, sign)); res[k] = make_double2(v.x / (double)M, v.y / (double)M); }
  }
  if ((int64_t)res.size() != n) return FB_EINVAL;
  memcpy(out_ri, res.data(), (size_t)n * sizeof(double2));
  return FB_OK;
}
