// "v1" streaming demodulators (bytecode-only in the reference, parity unpinned):
//   I/Q integrate-and-dump BPSK / QPSK / 8PSK (B.4-B.6), per-symbol DFT OFDM demap (B.7), Goertzel tone-pair FSK
//   (B.2, B.3).  All are one pattern: per symbol a small bank of correlations  F_m = sum_j x[j] W[m][j]  over the
//   symbol's own samples, then a hard decision -- 2..2*nf FMA per sample, one read of every sample: HBM-bound.
//
//   v1_corr_kernel   CTA = tile of 256 symbols.  sps <= 32: thread per symbol (each warp covers 32*sps contiguous
//                    samples; lines are fetched once and re-hit in L1); larger sps: warp per symbol, lanes stride the
//                    samples (coalesced) and the correlations are reduced with warp shuffles.  Decisions go to
//                    shared memory, are packed 32 bits per thread and stored big-endian straight into the output
//                    slot (PSK / OFDM / FSK-HS have no sync search: bytes are the bit stream truncated to x8).
//   uart_deframe_kernel   B.2's start/8 data LSB-first/stop deframer: inherently sequential, one thread per recording.
//   Goertzel: power = s1^2 + s2^2 - coeff s1 s2 == |sum x[k] e^{-jwk}|^2, evaluated as that correlation in float64.
#include "common.cuh"

#include <algorithm>

enum { V1_BPSK = 0, V1_QPSK = 1, V1_PSK8 = 2, V1_OFDM = 3, V1_FSK = 4 };
#define V1_TILE 256

struct V1Args {
  const void* samples;           // prefiltered float32 buffer for FSK, caller samples otherwise
  const RecPlan* plans;          // nsym = symbols, word_off = bit-stream words (FSK+UART) , out_off/out_cap
  const uint32_t* tile_first;
  const double2* table;          // [nf][len] complex weights (float64)
  uint8_t* out;
  uint32_t* bits;                // workspace words (uart mode) or nullptr
  int n_rec, mode, sps, off0, len, nf, bpsym, to_workspace;
};

__device__ __forceinline__ uint32_t quadrant_code(double re, double im) {      // B.5
  if (re >= 0.0 && im >= 0.0) return 0u;
  if (re < 0.0 && im >= 0.0) return 1u;
  if (re < 0.0 && im < 0.0) return 3u;
  return 2u;
}

template <typename TIn>
__global__ void __launch_bounds__(V1_TILE) v1_corr_kernel(const V1Args a) {
  extern __shared__ __align__(16) unsigned char v1_smem[];
  double2* W = reinterpret_cast<double2*>(v1_smem);                       // [nf][len]
  uint32_t* codes = reinterpret_cast<uint32_t*>(W + (size_t)a.nf * a.len);   // [V1_TILE]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int lo = 0, hi = a.n_rec;
  const uint32_t tile = blockIdx.x;
  while (hi - lo > 1) {
    const int step = (hi - lo + 31) >> 5;
    const int probe = lo + (lane + 1) * step;
    const bool le = probe < hi && __ldg(&a.tile_first[probe]) <= tile;
    const int cnt = __popc(__ballot_sync(0xffffffffu, le));
    const int nlo = lo + cnt * step;
    hi = min(hi, nlo + step);
    lo = nlo;
  }
  const RecPlan pl = a.plans[lo];
  const int k0 = (int)(tile - a.tile_first[lo]) * V1_TILE;                 // first symbol of the tile
  const int ns = min(V1_TILE, pl.nsym - k0);
  for (int i = tid; i < a.nf * a.len; i += V1_TILE) W[i] = a.table[i];
  codes[tid] = 0;
  __syncthreads();

  auto decide = [&](const double* fr, const double* fi) -> uint32_t {
    switch (a.mode) {
      case V1_BPSK: return fr[0] > 0.0 ? 0u : 1u;                          // B.4: '0' if I > 0 else '1'
      case V1_QPSK: return quadrant_code(fr[0], fi[0]);
      case V1_PSK8: {                                                      // B.6
        double phi = atan2(fi[0], fr[0]);
        if (phi < 0.0) phi += 2.0 * 3.141592653589793;
        uint32_t code = 0;
#pragma unroll
        for (int t = 1; t <= 13; t += 2) code += (phi >= t * (3.141592653589793 / 8.0)) ? 1u : 0u;
        return code;
      }
      case V1_OFDM: {                                                      // B.7: bins in order, 2 bits each
        uint32_t code = 0;
        for (int m = 0; m < a.nf; ++m) code = (code << 2) | quadrant_code(fr[m], fi[m]);
        return code;
      }
      default: {                                                           // V1_FSK: bit = p_mark > p_space
        const double pm = fr[0] * fr[0] + fi[0] * fi[0], ps = fr[1] * fr[1] + fi[1] * fi[1];
        return pm > ps ? 1u : 0u;
      }
    }
  };

  constexpr int MAXF = 8;
  if (a.sps <= 32) {
    if (tid < ns) {                                                        // thread per symbol
      const uint64_t base = pl.off + (uint64_t)(k0 + tid) * a.sps + a.off0;
      double fr[MAXF], fi[MAXF];
#pragma unroll
      for (int m = 0; m < MAXF; ++m) { fr[m] = 0.0; fi[m] = 0.0; }
      for (int j = 0; j < a.len; ++j) {
        const double x = load_sample_d<TIn>(a.samples, base + j);
#pragma unroll
        for (int m = 0; m < MAXF; ++m)
          if (m < a.nf) { const double2 w = W[m * a.len + j]; fr[m] = fma(x, w.x, fr[m]); fi[m] = fma(x, w.y, fi[m]); }
      }
      codes[tid] = decide(fr, fi);
    }
  } else {
    for (int s = warp; s < ns; s += V1_TILE / 32) {                        // warp per symbol
      const uint64_t base = pl.off + (uint64_t)(k0 + s) * a.sps + a.off0;
      double fr[MAXF], fi[MAXF];
#pragma unroll
      for (int m = 0; m < MAXF; ++m) { fr[m] = 0.0; fi[m] = 0.0; }
      for (int j = lane; j < a.len; j += 32) {
        const double x = load_sample_d<TIn>(a.samples, base + j);
#pragma unroll
        for (int m = 0; m < MAXF; ++m)
          if (m < a.nf) { const double2 w = W[m * a.len + j]; fr[m] = fma(x, w.x, fr[m]); fi[m] = fma(x, w.y, fi[m]); }
      }
#pragma unroll
      for (int m = 0; m < MAXF; ++m)
        if (m < a.nf) {
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            fr[m] += __shfl_xor_sync(0xffffffffu, fr[m], off);
            fi[m] += __shfl_xor_sync(0xffffffffu, fi[m], off);
          }
        }
      if (lane == 0) codes[s] = decide(fr, fi);
    }
  }
  __syncthreads();
  // ---- pack: V1_TILE * bpsym bits = 8 * bpsym whole words per tile ---------------------------------------
  const int nwords = (ns * a.bpsym + 31) / 32;
  for (int w = tid; w < nwords; w += V1_TILE) {
    uint32_t word = 0;
    for (int b = 0; b < 32; ++b) {
      const int bit = w * 32 + b;
      const int sym = bit / a.bpsym, pos = bit - sym * a.bpsym;
      const uint32_t v = (sym < ns) ? ((codes[sym] >> (a.bpsym - 1 - pos)) & 1u) : 0u;
      word = (word << 1) | v;
    }
    const uint64_t widx = (uint64_t)k0 * a.bpsym / 32 + w;
    if (a.to_workspace) {
      a.bits[pl.word_off + widx] = __byte_perm(word, 0, 0x0123);
    } else {
      const uint64_t nbytes = min((uint64_t)pl.nsym * a.bpsym / 8, pl.out_cap);   // truncate to a multiple of 8 bits
      uint8_t* o = a.out + pl.out_off;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (widx * 4 + k < nbytes) o[widx * 4 + k] = (uint8_t)(word >> (24 - 8 * k));
    }
  }
}

// B.2 UART deframer (pyc src 310-324): one thread per recording, bits from the workspace word stream
__global__ void __launch_bounds__(32) uart_deframe_kernel(const RecPlan* plans, int n_rec, const uint32_t* bits, uint8_t* out,
                                                           uint64_t* out_len) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rec) return;
  const RecPlan pl = plans[r];
  const uint32_t* w = bits + pl.word_off;
  const int64_t n = pl.nsym;
  uint8_t* o = out + pl.out_off;
  uint64_t cnt = 0;
  auto bit = [&](int64_t i) -> uint32_t { return (__byte_perm(w[i >> 5], 0, 0x0123) >> (31 - (i & 31))) & 1u; };
  int64_t i = 0;
  while (i + 10 <= n) {
    if (bit(i) != 0) { ++i; continue; }
    if (bit(i + 9) != 1) { ++i; continue; }
    uint32_t b = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) b |= bit(i + 1 + k) << k;                  // LSB first
    if (cnt < pl.out_cap) o[cnt] = (uint8_t)b;
    ++cnt;
    i += 10;
  }
  out_len[r] = min(cnt, pl.out_cap);
}

__global__ void __launch_bounds__(FB_THREADS) v1_finish_kernel(const RecPlan* plans, int n_rec, int bpsym, uint64_t* out_len, int32_t* status,
                                                                int write_len) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rec) return;
  if (write_len) out_len[r] = min((uint64_t)plans[r].nsym * bpsym / 8, plans[r].out_cap);
  status[r] = plans[r].status;
}

// ------------------------------------------------------------------------------------------------ pre-filter (B.1)
// scipy filtfilt (order <= 8, float64), chunk-parallel exactly like fsk_v2.cu; output cast to float32 (B.2: the v1 code
// re-casts the filtered record to float32 before the Goertzel loop).
#define PF_ORD 8
#define PF_CHUNK 2048
struct PfFilt { double b[PF_ORD + 1], a[PF_ORD + 1], zi[PF_ORD]; int32_t w, pad; };

template <typename TIn>
__device__ __forceinline__ double pf_x_ext(const void* samples, uint64_t off, int64_t N, int64_t n) {
  if (n < 0) return 2.0 * load_sample_d<TIn>(samples, off) - load_sample_d<TIn>(samples, off + (uint64_t)(-n));
  if (n > N - 1)
    return 2.0 * load_sample_d<TIn>(samples, off + (uint64_t)(N - 1)) - load_sample_d<TIn>(samples, off + (uint64_t)(2 * (N - 1) - n));
  return load_sample_d<TIn>(samples, off + (uint64_t)n);
}

template <typename TIn>
__global__ void __launch_bounds__(64) pf_fwd_kernel(const void* samples, uint64_t off, int64_t N, PfFilt t, double* yfwd) {
  const int64_t Next = N + 2 * t.pad;
  const int64_t c0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * PF_CHUNK;
  if (c0 >= Next) return;
  const int64_t c1 = min(Next, c0 + PF_CHUNK), s = max((int64_t)0, c0 - t.w);
  double z[PF_ORD];
  const double x0 = pf_x_ext<TIn>(samples, off, N, s - t.pad);
#pragma unroll
  for (int i = 0; i < PF_ORD; ++i) z[i] = (s == 0) ? t.zi[i] * x0 : 0.0;
  constexpr int EB = 16;
  for (int64_t e0 = s; e0 < c1; e0 += EB) {
    double xb[EB];
#pragma unroll
    for (int u = 0; u < EB; ++u) xb[u] = (e0 + u < c1) ? pf_x_ext<TIn>(samples, off, N, e0 + u - t.pad) : 0.0;
#pragma unroll
    for (int u = 0; u < EB; ++u) {
      const double xv = xb[u], y = t.b[0] * xv + z[0];
#pragma unroll
      for (int k = 0; k < PF_ORD - 1; ++k) z[k] = t.b[k + 1] * xv + z[k + 1] - t.a[k + 1] * y;
      z[PF_ORD - 1] = t.b[PF_ORD] * xv - t.a[PF_ORD] * y;
      if (e0 + u >= c0 && e0 + u < c1) yfwd[e0 + u] = y;
    }
  }
}

__global__ void __launch_bounds__(64) pf_bwd_kernel(const double* yfwd, int64_t N, PfFilt t, float* f32) {
  const int64_t Next = N + 2 * t.pad;
  const int64_t c0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * PF_CHUNK;
  if (c0 >= Next) return;
  const int64_t c1 = min(Next, c0 + PF_CHUNK), s = min(Next - 1, c1 - 1 + t.w);
  double z[PF_ORD];
  const double y0 = yfwd[s];
#pragma unroll
  for (int i = 0; i < PF_ORD; ++i) z[i] = (s == Next - 1) ? t.zi[i] * y0 : 0.0;
  constexpr int EB = 16;
  for (int64_t e0 = s; e0 >= c0; e0 -= EB) {
    double xb[EB];
#pragma unroll
    for (int u = 0; u < EB; ++u) xb[u] = (e0 - u >= c0) ? yfwd[e0 - u] : 0.0;
#pragma unroll
    for (int u = 0; u < EB; ++u) {
      const int64_t e = e0 - u;
      if (e >= c0) {
        const double xv = xb[u], y = t.b[0] * xv + z[0];
#pragma unroll
        for (int k = 0; k < PF_ORD - 1; ++k) z[k] = t.b[k + 1] * xv + z[k + 1] - t.a[k + 1] * y;
        z[PF_ORD - 1] = t.b[PF_ORD] * xv - t.a[PF_ORD] * y;
        const int64_t n = e - t.pad;
        if (e < c1 && n >= 0 && n < N) f32[n] = (float)y;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
extern "C" uint64_t fb_v1_out_bound(const fb_v1_params* p, uint64_t n_samples) {
  if (!p || p->sps < 1) return 0;
  const uint64_t nsym = n_samples / (uint64_t)p->sps;
  if (p->uart) return nsym / 10 + 1;
  return nsym * (uint64_t)p->bits_per_sym / 8;
}

extern "C" int fb_v1_demod_batch(fb_handle* h, const fb_v1_params* pp, const double* table, int n_rec, const void* samples,
                                 const uint64_t* offsets, int dtype, int flags, uint8_t* out, const uint64_t* out_offsets,
                                 uint64_t* out_len, int32_t* status) {
  if (!h || !pp || !table || n_rec < 0 || !offsets || !out_offsets) return FB_EINVAL;
  if (dtype != FB_F32 && dtype != FB_F64 && dtype != FB_S16) return FB_EINVAL;
  const fb_v1_params& p = *pp;
  if (p.sps < 1 || p.len < 1 || p.off0 < 0 || p.off0 + p.len > p.sps || p.nf < 1 || p.nf > 8 || p.bits_per_sym < 1 ||
      p.bits_per_sym > 16 || p.mode < V1_BPSK || p.mode > V1_FSK)
    return FB_EINVAL;
  FB_CUDA(h, cudaSetDevice(h->device));
  if (n_rec == 0) return FB_OK;
  const size_t esz = dtype == FB_F32 ? 4 : dtype == FB_F64 ? 8 : 2;
  std::vector<RecPlan> plans(n_rec);
  std::vector<uint32_t> tile_first(n_rec + 1, 0);
  uint32_t n_tiles = 0;
  uint64_t words = 0;
  int64_t maxN = 0;
  for (int r = 0; r < n_rec; ++r) {
    RecPlan& q = plans[r];
    q.off = offsets[r]; q.n = offsets[r + 1] - offsets[r];
    q.out_off = out_offsets[r]; q.out_cap = out_offsets[r + 1] - out_offsets[r];
    q.word_off = words; q.dl32 = q.dr32 = 0; q.status = FB_ST_OK; q.pad = 0;
    if (q.out_off & 3) return FB_EINVAL;                                   // words are stored into the slots
    if (p.prefilter && (int64_t)q.n <= p.bp_pad) { q.status = FB_ST_TOO_SHORT; q.nsym = q.ndsym = 0; tile_first[r] = n_tiles; continue; }
    q.nsym = q.ndsym = (int32_t)std::min<uint64_t>(q.n / (uint64_t)p.sps, 0x7fffffff);
    tile_first[r] = n_tiles;
    n_tiles += (uint32_t)((q.nsym + V1_TILE - 1) / V1_TILE);
    words += ((uint64_t)q.nsym * p.bits_per_sym + 31) / 32 + 2;
    maxN = std::max<int64_t>(maxN, (int64_t)q.n);
  }
  tile_first[n_rec] = n_tiles;
  const uint64_t total_samples = offsets[n_rec], total_out = out_offsets[n_rec];
  const void* d_samples = samples;
  int rc;
  if (!(flags & FB_SAMPLES_ON_DEVICE)) {
    if ((rc = fb_ensure(h, h->in, (size_t)total_samples * esz + 16))) return rc;
    FB_CUDA(h, cudaMemcpyAsync(h->in.p, samples, (size_t)total_samples * esz, cudaMemcpyHostToDevice, h->stream));
    d_samples = h->in.p;
  }
  uint8_t* d_out = out; uint64_t* d_out_len = out_len; int32_t* d_status = status;
  if (!(flags & FB_OUT_ON_DEVICE)) {
    if ((rc = fb_ensure(h, h->out, (size_t)total_out + 16))) return rc;
    if ((rc = fb_ensure(h, h->out_len, (size_t)n_rec * 8))) return rc;
    if ((rc = fb_ensure(h, h->status, (size_t)n_rec * 4))) return rc;
    d_out = (uint8_t*)h->out.p; d_out_len = (uint64_t*)h->out_len.p; d_status = (int32_t*)h->status.p;
  }
  if ((rc = fb_ensure(h, h->plans, (size_t)n_rec * sizeof(RecPlan)))) return rc;
  if ((rc = fb_ensure(h, h->tile_first, (size_t)(n_rec + 1) * 4))) return rc;
  if ((rc = fb_ensure(h, h->taps, (size_t)p.nf * p.len * 16))) return rc;
  if (p.uart && (rc = fb_ensure(h, h->bits, (size_t)(words + 4) * 4))) return rc;
  FB_CUDA(h, cudaMemcpyAsync(h->plans.p, plans.data(), (size_t)n_rec * sizeof(RecPlan), cudaMemcpyHostToDevice, h->stream));
  FB_CUDA(h, cudaMemcpyAsync(h->tile_first.p, tile_first.data(), (size_t)(n_rec + 1) * 4, cudaMemcpyHostToDevice, h->stream));
  FB_CUDA(h, cudaMemcpyAsync(h->taps.p, table, (size_t)p.nf * p.len * 16, cudaMemcpyHostToDevice, h->stream));

  int kdtype = dtype;
  if (p.prefilter) {
    // filtered copy of the whole batch as float32 (B.2 casts the filtered record back to float32): scratch = [f32 batch][yfwd]
    const size_t o_y = ((size_t)total_samples * 4 + 255) / 256 * 256;
    if ((rc = fb_ensure(h, h->scratch, o_y + ((size_t)maxN + 2 * p.bp_pad + 16) * 8))) return rc;
    float* f32 = (float*)h->scratch.p;
    double* yfwd = (double*)((char*)h->scratch.p + o_y);
    PfFilt t;
    for (int i = 0; i <= PF_ORD; ++i) { t.b[i] = p.bp_b[i]; t.a[i] = p.bp_a[i]; }
    for (int i = 0; i < PF_ORD; ++i) t.zi[i] = p.bp_zi[i];
    t.w = p.bp_w; t.pad = p.bp_pad;
    for (int r = 0; r < n_rec; ++r) {
      if (plans[r].status != FB_ST_OK) continue;
      const int64_t N = (int64_t)plans[r].n, Next = N + 2 * p.bp_pad;
      const int nthreads = (int)((Next + PF_CHUNK - 1) / PF_CHUNK), nblocks = (nthreads + 63) / 64;
      if (dtype == FB_F32) pf_fwd_kernel<float><<<nblocks, 64, 0, h->stream>>>(d_samples, plans[r].off, N, t, yfwd);
      else if (dtype == FB_F64) pf_fwd_kernel<double><<<nblocks, 64, 0, h->stream>>>(d_samples, plans[r].off, N, t, yfwd);
      else pf_fwd_kernel<int16_t><<<nblocks, 64, 0, h->stream>>>(d_samples, plans[r].off, N, t, yfwd);
      pf_bwd_kernel<<<nblocks, 64, 0, h->stream>>>(yfwd, N, t, f32 + plans[r].off);
      h->launches += 2;
    }
    d_samples = f32;
    kdtype = FB_F32;
  }
  V1Args a{};
  a.samples = d_samples; a.plans = (const RecPlan*)h->plans.p; a.tile_first = (const uint32_t*)h->tile_first.p;
  a.table = (const double2*)h->taps.p; a.out = d_out; a.bits = p.uart ? (uint32_t*)h->bits.p : nullptr;
  a.n_rec = n_rec; a.mode = p.mode; a.sps = p.sps; a.off0 = p.off0; a.len = p.len; a.nf = p.nf; a.bpsym = p.bits_per_sym;
  a.to_workspace = p.uart ? 1 : 0;
  const size_t smem = (size_t)p.nf * p.len * 16 + V1_TILE * 4;
  if (smem > 200 * 1024) return FB_EUNSUPPORTED;
  if (n_tiles > 0) {
    if (h->profiling) FB_CUDA(h, cudaEventRecord(h->ev_k0, h->stream));
    if (kdtype == FB_F32) {
      FB_CUDA(h, cudaFuncSetAttribute(v1_corr_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      v1_corr_kernel<float><<<n_tiles, V1_TILE, smem, h->stream>>>(a);
    } else if (kdtype == FB_F64) {
      FB_CUDA(h, cudaFuncSetAttribute(v1_corr_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      v1_corr_kernel<double><<<n_tiles, V1_TILE, smem, h->stream>>>(a);
    } else {
      FB_CUDA(h, cudaFuncSetAttribute(v1_corr_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      v1_corr_kernel<int16_t><<<n_tiles, V1_TILE, smem, h->stream>>>(a);
    }
    if (h->profiling) { FB_CUDA(h, cudaEventRecord(h->ev_k1, h->stream)); h->k_recorded = true; }
    h->launches++;
  }
  if (p.uart) {
    uart_deframe_kernel<<<(n_rec + 31) / 32, 32, 0, h->stream>>>((const RecPlan*)h->plans.p, n_rec, (const uint32_t*)h->bits.p, d_out, d_out_len);
    h->launches++;
  }
  v1_finish_kernel<<<(n_rec + FB_THREADS - 1) / FB_THREADS, FB_THREADS, 0, h->stream>>>((const RecPlan*)h->plans.p, n_rec, p.bits_per_sym,
                                                                                      d_out_len, d_status, p.uart ? 0 : 1);
  h->launches++;
  FB_CUDA(h, cudaGetLastError());
  if (!(flags & FB_OUT_ON_DEVICE)) {
    if (total_out) FB_CUDA(h, cudaMemcpyAsync(out, d_out, (size_t)total_out, cudaMemcpyDeviceToHost, h->stream));
    FB_CUDA(h, cudaMemcpyAsync(out_len, d_out_len, (size_t)n_rec * 8, cudaMemcpyDeviceToHost, h->stream));
    FB_CUDA(h, cudaMemcpyAsync(status, d_status, (size_t)n_rec * 4, cudaMemcpyDeviceToHost, h->stream));
  }
  h->last_plans = plans;
  h->last_bps = p.bits_per_sym;
  if (!(flags & FB_ASYNC) || !(flags & FB_OUT_ON_DEVICE)) FB_CUDA(h, cudaStreamSynchronize(h->stream));
  return FB_OK;
}
