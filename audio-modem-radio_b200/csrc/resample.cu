// WAV ingest on the device: channel 0 of interleaved PCM16 / float frames -> float64, and the FFT resampler
// decode_wav_file applies when the file is not at 96 kHz (decoder.py:381-387: soundfile.read -> data[:, 0] ->
// scipy.signal.resample(data, int(round(len(data) * 96000 / sr)))).
//
// scipy.signal.resample for a real record (scipy 1.18, `domain='time'`):  X = rfft(x);  keep the first m2 = m/2 + 1 bins,
// m = min(num, n);  when m is even and num != n the unpaired bin m/2 is doubled (down-sampling) or halved (up-sampling);
// x_r = irfft(X / (n / num), num).  Here: the hand-written float64 FFT of fft.cu (Stockham passes for 2-3-5-7-smooth lengths,
// Bluestein otherwise) -> resample_bins_kernel (crop / zero-pad, unpaired bin, both scale factors folded into one multiply)
// -> the inverse transform of fft.cu.
#include "common.cuh"
#include "fft.cuh"

// channel 0 of `n` interleaved frames, any supported sample type, to float64 (PCM16 scaled by 1/32768 like soundfile)
template <typename TIn>
__global__ void __launch_bounds__(FB_THREADS) ingest_kernel(const void* in, uint64_t n, int n_channels, double* out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = load_sample_d<TIn>(in, i * (uint64_t)n_channels);
}

// Y[k] = X[k] * scale for k < m2 (bin m/2 adjusted), 0 above; scale = (num / n) / num: irfft's 1/num and scipy's 1/s_fac
__global__ void __launch_bounds__(FB_THREADS) resample_bins_kernel(const double2* X, double2* Y, int64_t n, int64_t num) {
  const int64_t m = min(n, num), m2 = m / 2 + 1, ny = num / 2 + 1;
  const double scale = 1.0 / (double)n;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < ny; k += (int64_t)gridDim.x * blockDim.x) {
    double2 v = make_double2(0.0, 0.0);
    if (k < m2) {
      v = X[k];
      double f = scale;
      if ((m & 1) == 0 && num != n && k == m / 2) f *= (num < n) ? 2.0 : 0.5;
      v.x *= f; v.y *= f;
    }
    Y[k] = v;
  }
}

void fb_resample_release(fb_handle*) {}   // (nothing cached per handle any more: the FFT workspace is a handle buffer)

extern "C" int fb_ingest_resample(fb_handle* h, const void* in, uint64_t n_frames, int n_channels, int dtype, uint64_t n_out,
                                  double* out, int flags) {
  if (!h || !in || !out || n_channels < 1 || n_frames < 1 || n_out < 1) return FB_EINVAL;
  FB_LOCK(h);
  if (dtype != FB_F32 && dtype != FB_F64 && dtype != FB_S16) return FB_EINVAL;
  if (n_frames > ((uint64_t)1 << 31) - 64 || n_out > ((uint64_t)1 << 31) - 64) return FB_EUNSUPPORTED;
  FB_CUDA(h, cudaSetDevice(h->device));
  const size_t esz = dtype == FB_F32 ? 4 : dtype == FB_F64 ? 8 : 2;
  const int64_t n = (int64_t)n_frames, num = (int64_t)n_out;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  // workspace: [x float64][X][Y][y float64] (+ the raw frames when they come from the host)
  const size_t o_X = al((size_t)n * 8), o_Y = o_X + al(((size_t)n / 2 + 2) * 16), o_y = o_Y + al(((size_t)num / 2 + 2) * 16),
               o_raw = o_y + al((size_t)num * 8), raw_bytes = (size_t)n * n_channels * esz;
  const bool host_in = !(flags & FB_SAMPLES_ON_DEVICE), host_out = !(flags & FB_OUT_ON_DEVICE);
  int rc;
  if ((rc = fb_ensure(h, h->misc, o_raw + (host_in ? raw_bytes + 16 : 16)))) return rc;
  char* ws = (char*)h->misc.p;
  const void* d_in = in;
  if (host_in) {
    FB_CUDA(h, cudaMemcpyAsync(ws + o_raw, in, raw_bytes, cudaMemcpyHostToDevice, h->stream));
    d_in = ws + o_raw;
  }
  double* x = (double*)ws;
  double* y = host_out ? (double*)(ws + o_y) : out;
  const int g = (int)std::min<int64_t>(148 * 8, (n + FB_THREADS - 1) / FB_THREADS);
  double* ingest_dst = (num == n) ? y : x;               // same rate: ingest straight into the result
  if (dtype == FB_F32) ingest_kernel<float><<<g, FB_THREADS, 0, h->stream>>>(d_in, (uint64_t)n, n_channels, ingest_dst);
  else if (dtype == FB_F64) ingest_kernel<double><<<g, FB_THREADS, 0, h->stream>>>(d_in, (uint64_t)n, n_channels, ingest_dst);
  else ingest_kernel<int16_t><<<g, FB_THREADS, 0, h->stream>>>(d_in, (uint64_t)n, n_channels, ingest_dst);
  h->launches++;
  if (num != n) {
    double2* X = (double2*)(ws + o_X);
    double2* Y = (double2*)(ws + o_Y);
    if ((rc = fb_fft_d2z(h, x, X, n))) return rc;
    const int g2 = (int)std::min<int64_t>(148 * 8, (num / 2 + 1 + FB_THREADS - 1) / FB_THREADS);
    resample_bins_kernel<<<g2, FB_THREADS, 0, h->stream>>>(X, Y, n, num);
    h->launches++;
    if ((rc = fb_fft_z2d(h, Y, y, num))) return rc;
  }
  FB_CUDA(h, cudaGetLastError());
  if (host_out) FB_CUDA(h, cudaMemcpyAsync(out, y, (size_t)num * 8, cudaMemcpyDeviceToHost, h->stream));
  if (!(flags & FB_ASYNC) || host_out) FB_CUDA(h, cudaStreamSynchronize(h->stream));
  return FB_OK;
}
