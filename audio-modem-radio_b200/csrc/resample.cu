// WAV ingest on the device: channel 0 of interleaved PCM16 / float frames -> float64, and the FFT resampler
// decode_wav_file applies when the file is not at 96 kHz (decoder.py:381-387: soundfile.read -> data[:, 0] ->
// scipy.signal.resample(data, int(round(len(data) * 96000 / sr)))).
//
// scipy.signal.resample for a real record (scipy 1.18, `domain='time'`):  X = rfft(x);  keep the first m2 = m/2 + 1 bins,
// m = min(num, n);  when m is even and num != n the unpaired bin m/2 is doubled (down-sampling) or halved (up-sampling);
// x_r = irfft(X / (n / num), num).  Here: cuFFT D2Z (the one library-shaped op, as the reference uses pocketfft) ->
// resample_bins_kernel (crop / zero-pad, unpaired bin, both scale factors folded into one multiply) -> cuFFT Z2D.
#include "common.cuh"

#include <cufft.h>
#include <map>
#include <utility>

// channel 0 of `n` interleaved frames, any supported sample type, to float64 (PCM16 scaled by 1/32768 like soundfile)
template <typename TIn>
__global__ void __launch_bounds__(FB_THREADS) ingest_kernel(const void* in, uint64_t n, int n_channels, double* out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = load_sample_d<TIn>(in, i * (uint64_t)n_channels);
}

// Y[k] = X[k] * scale for k < m2 (bin m/2 adjusted), 0 above; scale = (num / n) / num: irfft's 1/num and scipy's 1/s_fac
__global__ void __launch_bounds__(FB_THREADS) resample_bins_kernel(const cufftDoubleComplex* X, cufftDoubleComplex* Y, int64_t n, int64_t num) {
  const int64_t m = min(n, num), m2 = m / 2 + 1, ny = num / 2 + 1;
  const double scale = 1.0 / (double)n;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < ny; k += (int64_t)gridDim.x * blockDim.x) {
    cufftDoubleComplex v = make_cuDoubleComplex(0.0, 0.0);
    if (k < m2) {
      v = X[k];
      double f = scale;
      if ((m & 1) == 0 && num != n && k == m / 2) f *= (num < n) ? 2.0 : 0.5;
      v.x *= f; v.y *= f;
    }
    Y[k] = v;
  }
}

struct ResamplePlans {                                   // (length, 0 = D2Z | 1 = Z2D) -> plan, with the tick of its last use
  std::map<std::pair<int64_t, int>, cufftHandle> plans;
  std::map<std::pair<int64_t, int>, uint64_t> used;
  uint64_t tick = 0;
};
static std::map<fb_handle*, ResamplePlans> g_rs_plans;

void fb_resample_release(fb_handle* h) {
  auto it = g_rs_plans.find(h);
  if (it == g_rs_plans.end()) return;
  for (auto& p : it->second.plans) cufftDestroy(p.second);
  g_rs_plans.erase(it);
}

static int rs_plan(fb_handle* h, int64_t len, int inverse, cufftHandle* out) {
  ResamplePlans& rp = g_rs_plans[h];
  auto key = std::make_pair(len, inverse);
  auto it = rp.plans.find(key);
  if (it == rp.plans.end()) {
    if (rp.plans.size() >= 16) {                           // evict the least recently used plan
      auto lru = rp.used.begin();
      for (auto u = rp.used.begin(); u != rp.used.end(); ++u) if (u->second < lru->second) lru = u;
      FB_CUDA(h, cudaStreamSynchronize(h->stream));        // transforms queued on it must have finished
      cufftDestroy(rp.plans[lru->first]);
      rp.plans.erase(lru->first);
      rp.used.erase(lru);
    }
    cufftHandle pl;
    if (cufftPlan1d(&pl, (int)len, inverse ? CUFFT_Z2D : CUFFT_D2Z, 1) != CUFFT_SUCCESS) { h->err = "cufftPlan1d failed"; return FB_ECUDA; }
    cufftSetStream(pl, h->stream);
    it = rp.plans.emplace(key, pl).first;
  }
  rp.used[key] = ++rp.tick;
  *out = it->second;
  return FB_OK;
}

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
    cufftHandle fwd, inv;
    if ((rc = rs_plan(h, n, 0, &fwd)) || (rc = rs_plan(h, num, 1, &inv))) return rc;
    cufftDoubleComplex* X = (cufftDoubleComplex*)(ws + o_X);
    cufftDoubleComplex* Y = (cufftDoubleComplex*)(ws + o_Y);
    if (cufftExecD2Z(fwd, x, X) != CUFFT_SUCCESS) { h->err = "cufftExecD2Z failed"; return FB_ECUDA; }
    const int g2 = (int)std::min<int64_t>(148 * 8, (num / 2 + 1 + FB_THREADS - 1) / FB_THREADS);
    resample_bins_kernel<<<g2, FB_THREADS, 0, h->stream>>>(X, Y, n, num);
    h->launches++;
    if (cufftExecZ2D(inv, Y, y) != CUFFT_SUCCESS) { h->err = "cufftExecZ2D failed"; return FB_ECUDA; }
  }
  FB_CUDA(h, cudaGetLastError());
  if (host_out) FB_CUDA(h, cudaMemcpyAsync(out, y, (size_t)num * 8, cudaMemcpyDeviceToHost, h->stream));
  if (!(flags & FB_ASYNC) || host_out) FB_CUDA(h, cudaStreamSynchronize(h->stream));
  return FB_OK;
}
