// v2 FSK receive chain on sm_100a: replaces fsk_demodulate (modem.py:298-341) for valid tone sets.
//
// Per tone (modem.py:306-309):  f = filtfilt(butter(3, [f-b, f+b]), x);  env = |hilbert(f)|
//   zp_fwd_kernel / zp_bwd_kernel (zp_iir.cuh)   scipy.signal.filtfilt in float64: odd extension by padlen, lfilter_zi
//                                     start-up, DF2T forward then backward.  Chunk-parallel with coalesced traffic; all
//                                     recordings of a group in one launch per pass.
//   fb_fft_d2z -> hilbert_spec -> fb_fft_z2d  scipy.signal.hilbert *is* one length-N FFT, a one-sided mask and one length-N
//                                     inverse FFT over the whole recording (circular): the hand-written float64 FFT of
//                                     fft.cu (Stockham passes; Bluestein for lengths with other prime factors).  Only the
//                                     imaginary part needs the inverse (the real part is the input), so it is a real transform.
// Decision (modem.py:315-323): bits[n] = env_mark[n] > env_space[n]; per bit a majority vote over the centre half
// (window truncated at the record end; spb < 4 -> empty window -> no bits).  fsk_vote_kernel packs the decided bits
// into the same big-endian word stream the DPSK kernels write; backend.cu does the magic search and byte packing.
#include "common.cuh"
#include "zp_iir.cuh"

#include "fft.cuh"
#include <algorithm>
#include <map>

#define FSK_ORD 6            // butter(3, band) -> 6th order, 7 coefficients

// scipy.signal.hilbert: x_a = ifft(fft(x) h), h = [1, 2, ..., 2, 1 (N even), 0, ...].  Its real part is x itself and its
// imaginary part is H = irfft(-j X[k]) over 0 < k < N/2 (DC and Nyquist bins dropped): one real-to-complex and one
// complex-to-real transform instead of a full complex inverse.  In place on the half spectrum.
__global__ void __launch_bounds__(FB_THREADS) hilbert_spec_kernel(double2* X, int64_t N) {
  const int64_t nh = N / 2 + 1;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nh; k += (int64_t)gridDim.x * blockDim.x) {
    const double2 v = X[k];
    const bool keep = k > 0 && 2 * k < N;
    X[k] = keep ? make_double2(v.y, -v.x) : make_double2(0.0, 0.0);      // -j (a + j b) = b - j a
  }
}

// |x_a|^2 N^2 = (N f)^2 + H^2 (the inverse transform is unnormalised).  First tone: keep it; second tone:
// cmp[n] = env_mark > env_space  (squares compare like the envelopes; the common N^2 scale drops out)
__global__ void __launch_bounds__(FB_THREADS) env_kernel(const double* f, const double* H, int64_t N, double* env2, uint8_t* cmp, int second) {
  const double dn = (double)N;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    const double re = dn * f[n], im = H[n];
    const double e = re * re + im * im;
    if (!second) env2[n] = e;
    else cmp[n] = env2[n] > e ? 1 : 0;
  }
}

// majority vote per bit (modem.py:320-323), 32 bits per thread, big-endian words
__global__ void __launch_bounds__(FB_THREADS) fsk_vote_kernel(const uint8_t* cmp, int64_t N, int spb, int64_t nbits, uint32_t* words) {
  const int64_t nwords = (nbits + 31) / 32;
  const int q = spb / 4;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (int64_t)gridDim.x * blockDim.x) {
    uint32_t word = 0;
    for (int k = 0; k < 32; ++k) {
      const int64_t i = w * 32 + k;
      uint32_t bit = 0;
      if (i < nbits) {
        const int64_t c = spb / 2 + i * spb;
        const int64_t lo = c - q, hi = min(c + q, N);
        int ones = 0;
        for (int64_t n = lo; n < hi; ++n) ones += cmp[n];
        bit = (2 * ones > (int)(hi - lo)) ? 1u : 0u;       // np.mean(chunk) > 0.5
      }
      word = (word << 1) | bit;
    }
    words[w] = __byte_perm(word, 0, 0x0123);
  }
}

// ------------------------------------------------------------------------------------------------ host side
void fb_fsk_release(fb_handle*) {}   // (nothing cached per handle any more: the FFT workspace is a handle buffer)

// Hilbert envelope compare + vote for one recording whose two tone-filtered copies f0, f1 (float64) are ready
static int fsk_one(fb_handle* h, const fb_fsk_design& d, int64_t N, int64_t nbits, uint32_t* d_words,
                   double* f0, double* f1, double2* X, double2* Z, double* env, uint8_t* cmp) {
  for (int tone = 0; tone < 2; ++tone) {
    double* ft = tone ? f1 : f0;
    int rc = fb_fft_d2z(h, ft, X, N);
    if (rc) return rc;
    const int g = (int)std::min<int64_t>(148 * 8, (N + FB_THREADS - 1) / FB_THREADS);
    hilbert_spec_kernel<<<g, FB_THREADS, 0, h->stream>>>(X, N);
    if ((rc = fb_fft_z2d(h, X, reinterpret_cast<double*>(Z), N))) return rc;
    env_kernel<<<g, FB_THREADS, 0, h->stream>>>(ft, reinterpret_cast<const double*>(Z), N, env, cmp, tone);
    h->launches += 2;
  }
  if (nbits > 0) {
    const int g = (int)std::min<int64_t>(148 * 8, ((nbits + 31) / 32 + FB_THREADS - 1) / FB_THREADS);
    fsk_vote_kernel<<<g, FB_THREADS, 0, h->stream>>>(cmp, N, d.spb, nbits, d_words);
    h->launches++;
  }
  FB_CUDA(h, cudaGetLastError());
  return FB_OK;
}

extern "C" uint64_t fb_fsk_out_bound(const fb_fsk_design* d, uint64_t n_samples) {
  if (!d || d->spb < 1 || d->spb / 4 == 0) return 0;
  const int64_t N = (int64_t)n_samples, c0 = d->spb / 2;
  const int64_t nbits = N > c0 ? (N - c0 + d->spb - 1) / d->spb : 0;
  return (uint64_t)(nbits / 8);
}

extern "C" int fb_fsk_demod_batch(fb_handle* h, const fb_fsk_design* dp, int n_rec, const void* samples, const uint64_t* offsets,
                                  int dtype, int flags, uint8_t* out, const uint64_t* out_offsets, uint64_t* out_len,
                                  int64_t* sync_idx, int32_t* status) {
  if (!h || !dp || n_rec < 0 || !offsets || !out_offsets) return FB_EINVAL;
  FB_LOCK(h);
  if (dtype != FB_F32 && dtype != FB_F64 && dtype != FB_S16) return FB_EINVAL;
  const fb_fsk_design& d = *dp;
  if (d.spb < 1 || d.pad < 1) return FB_EINVAL;
  FB_CUDA(h, cudaSetDevice(h->device));
  if (n_rec == 0) return FB_OK;
  const size_t esz = dtype == FB_F32 ? 4 : dtype == FB_F64 ? 8 : 2;
  std::vector<RecPlan> plans(n_rec);
  uint64_t words = 0;
  int64_t maxN = 0;
  for (int r = 0; r < n_rec; ++r) {
    RecPlan& p = plans[r];
    p.off = offsets[r]; p.n = offsets[r + 1] - offsets[r];
    p.out_off = out_offsets[r]; p.out_cap = out_offsets[r + 1] - out_offsets[r];
    p.word_off = words; p.dl32 = p.dr32 = 0; p.nsym = 0; p.ndsym = 0; p.status = FB_ST_OK; p.pad = 0;
    const int64_t N = (int64_t)p.n;
    if (N <= d.pad) { p.status = FB_ST_TOO_SHORT; continue; }
    if (N > ((int64_t)1 << 31) - 64) { p.status = FB_ST_UNSUPPORTED; continue; }
    const int64_t c0 = d.spb / 2;
    int64_t nbits = (d.spb / 4 > 0 && N > c0) ? (N - c0 + d.spb - 1) / d.spb : 0;   // len(range(spb//2, N, spb)); q == 0 -> none
    p.nsym = p.ndsym = (int32_t)nbits;                                              // backend reads ndsym * bps bits
    words += ((uint64_t)nbits + 31) / 32 + 2;
    maxN = std::max(maxN, N);
  }
  const uint64_t total_samples = offsets[n_rec], total_out = out_offsets[n_rec];
  const void* d_samples = samples;
  int rc;
  if (!(flags & FB_SAMPLES_ON_DEVICE)) {
    if ((rc = fb_ensure(h, h->in, (size_t)total_samples * esz + 16))) return rc;
    FB_CUDA(h, cudaMemcpyAsync(h->in.p, samples, (size_t)total_samples * esz, cudaMemcpyHostToDevice, h->stream));
    d_samples = h->in.p;
  }
  uint8_t* d_out = out; uint64_t* d_out_len = out_len; int64_t* d_sync = sync_idx; int32_t* d_status = status;
  if (!(flags & FB_OUT_ON_DEVICE)) {
    if ((rc = fb_ensure(h, h->out, (size_t)total_out + 16))) return rc;
    if ((rc = fb_ensure(h, h->out_len, (size_t)n_rec * 8))) return rc;
    if ((rc = fb_ensure(h, h->sync_idx, (size_t)n_rec * 8))) return rc;
    if ((rc = fb_ensure(h, h->status, (size_t)n_rec * 4))) return rc;
    d_out = (uint8_t*)h->out.p; d_out_len = (uint64_t*)h->out_len.p; d_sync = (int64_t*)h->sync_idx.p; d_status = (int32_t*)h->status.p;
  }
  if ((rc = fb_ensure(h, h->bits, (size_t)(words + 4) * 4))) return rc;
  if ((rc = fb_ensure(h, h->plans, (size_t)n_rec * sizeof(RecPlan)))) return rc;
  FB_CUDA(h, cudaMemcpyAsync(h->plans.p, plans.data(), (size_t)n_rec * sizeof(RecPlan), cudaMemcpyHostToDevice, h->stream));
  // scratch: [X | Z | env | cmp for the longest recording][f0 | f1 | transposed forward scratch for one GROUP of recordings]
  // [ZpRec table].  The two tone filters run for a whole group per launch (grid.y = recording); the Hilbert step then
  // goes recording by recording (one library FFT each).
  const size_t nN = (size_t)maxN + 16;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };            // 16-byte aligned complex buffers
  const size_t o_Z = al((nN / 2 + 2) * 16), o_env = al(o_Z + nN * 16), o_cmp = al(o_env + nN * 8), o_grp = al(o_cmp + nN + 64);
  const int Lc = zp_chunk_len(std::max(d.w[0], d.w[1]));
  std::vector<ZpRec> zr(n_rec);
  std::vector<uint64_t> f_off(n_rec, 0);
  std::vector<std::pair<int, int>> groups;
  const uint64_t group_doubles = (uint64_t)1 << 28;                    // 2 GB of transposed scratch per group
  uint64_t max_y = 0, max_f = 0;
  for (int r0 = 0; r0 < n_rec;) {
    uint64_t used = 0, fused = 0;
    int r1 = r0;
    while (r1 < n_rec && r1 - r0 < 65535) {
      ZpRec& q = zr[r1];
      q.off = plans[r1].off;
      q.N = plans[r1].status == FB_ST_OK ? (int64_t)plans[r1].n : 0;
      q.nch = q.N > 0 ? (q.N + 2 * d.pad + Lc - 1) / Lc : 0;
      const uint64_t need = (uint64_t)q.nch * Lc;
      if (r1 > r0 && used + need > group_doubles) break;
      q.y_off = used; used += need;
      q.out_off = fused; f_off[r1] = fused; fused += ((uint64_t)q.N + 31) / 32 * 32;
      ++r1;
    }
    max_y = std::max(max_y, used); max_f = std::max(max_f, fused);
    groups.emplace_back(r0, r1);
    r0 = r1;
  }
  const size_t o_f1 = o_grp + al((size_t)max_f * 8), o_y = o_f1 + al((size_t)max_f * 8), o_t = o_y + al((size_t)max_y * 8);
  if (maxN > 0) {
    if ((rc = fb_ensure(h, h->scratch, o_t + (size_t)n_rec * sizeof(ZpRec) + 64))) return rc;
  }
  char* sc = (char*)h->scratch.p;
  if (maxN > 0) {
    ZpRec* d_zr = (ZpRec*)(sc + o_t);
    FB_CUDA(h, cudaMemcpyAsync(d_zr, zr.data(), (size_t)n_rec * sizeof(ZpRec), cudaMemcpyHostToDevice, h->stream));
    double* fbuf[2] = {(double*)(sc + o_grp), (double*)(sc + o_f1)};
    double* ytr = (double*)(sc + o_y);
    for (auto& gr : groups) {
      int64_t gmax = 0;
      for (int r = gr.first; r < gr.second; ++r) gmax = std::max<int64_t>(gmax, zr[r].nch);
      if (gmax == 0) continue;
      const dim3 grid((unsigned)((gmax + ZP_THREADS - 1) / ZP_THREADS), gr.second - gr.first);
      for (int tone = 0; tone < 2; ++tone) {
        ZpFilt<FSK_ORD> t;
        for (int i = 0; i <= FSK_ORD; ++i) { t.b[i] = d.b[tone][i]; t.a[i] = d.a[tone][i]; }
        for (int i = 0; i < FSK_ORD; ++i) t.zi[i] = d.zi[tone][i];
        t.w = d.w[tone]; t.pad = d.pad;
        const int W16 = (std::min(d.w[tone], Lc) + 15) / 16 * 16;
        if (dtype == FB_F32) zp_fwd_kernel<float, FSK_ORD><<<grid, ZP_THREADS, 0, h->stream>>>(d_samples, d_zr + gr.first, t, Lc, W16, ytr);
        else if (dtype == FB_F64) zp_fwd_kernel<double, FSK_ORD><<<grid, ZP_THREADS, 0, h->stream>>>(d_samples, d_zr + gr.first, t, Lc, W16, ytr);
        else zp_fwd_kernel<int16_t, FSK_ORD><<<grid, ZP_THREADS, 0, h->stream>>>(d_samples, d_zr + gr.first, t, Lc, W16, ytr);
        zp_bwd_kernel<double, FSK_ORD><<<grid, ZP_THREADS, 0, h->stream>>>(ytr, d_zr + gr.first, t, Lc, W16, fbuf[tone]);
        h->launches += 2;
      }
      for (int r = gr.first; r < gr.second; ++r) {
        const RecPlan& p = plans[r];
        if (p.status != FB_ST_OK) continue;
        const int64_t N = (int64_t)p.n;
        uint32_t* d_words = (uint32_t*)h->bits.p + p.word_off;
        rc = fsk_one(h, d, N, p.ndsym, d_words, fbuf[0] + f_off[r], fbuf[1] + f_off[r],
                     (double2*)sc, (double2*)(sc + o_Z), (double*)(sc + o_env), (uint8_t*)(sc + o_cmp));
        if (rc) return rc;
      }
    }
  }
  rc = fb_bits_backend(h, n_rec, (const RecPlan*)h->plans.p, plans, 1, (const uint32_t*)h->bits.p, d_out, d_out_len, d_sync, d_status);
  if (rc) return rc;
  if (!(flags & FB_OUT_ON_DEVICE)) {
    if (total_out) FB_CUDA(h, cudaMemcpyAsync(out, d_out, (size_t)total_out, cudaMemcpyDeviceToHost, h->stream));
    FB_CUDA(h, cudaMemcpyAsync(out_len, d_out_len, (size_t)n_rec * 8, cudaMemcpyDeviceToHost, h->stream));
    FB_CUDA(h, cudaMemcpyAsync(sync_idx, d_sync, (size_t)n_rec * 8, cudaMemcpyDeviceToHost, h->stream));
    FB_CUDA(h, cudaMemcpyAsync(status, d_status, (size_t)n_rec * 4, cudaMemcpyDeviceToHost, h->stream));
  }
  h->last_plans = plans;
  h->last_bps = 1;
  if (!(flags & FB_ASYNC) || !(flags & FB_OUT_ON_DEVICE)) FB_CUDA(h, cudaStreamSynchronize(h->stream));
  return FB_OK;
}
