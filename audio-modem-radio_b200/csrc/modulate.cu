// batch modulators (TX side) -- bpsk_modulate (modem.py:28-65), qpsk_modulate (modem.py:138-186) and
// fsk_modulate (modem.py:270-295) for many payloads in one call, waveforms written straight into HBM (or back to the
// host).  Two kernels:
//   mod_phase_kernel   the reference accumulates the symbol phase with sequential float64 adds (modem.py:44-48,170-174)
//                      and, for CPFSK, a float modulo per bit (modem.py:292-293).  Floating-point addition does not
//                      re-associate, so the walk stays sequential: one thread per payload, phases to a float64
//                      workspace.  (1.7 M dependent adds for a 3-minute DQPSK part; all payloads walk in parallel.)
//   mod_wave_kernel    sample (k, j) = float32( sin(base[j] + phase[k]) * env[j] )  for PSK  (float64 sin, as numpy)
//                                    = float32( sin(w_k t[j] + phase[k]) ) * 0.9f       for CPFSK
//                      one thread per sample, coalesced float32 stores; base / env / t are sps-entry tables built on
//                      the host with the reference's own numpy expressions.
// Output equals the reference's float32 samples except where CUDA's and numpy's float64 sin differ in the last place
// AND that flips the float32 rounding (tests/test_gpu_modulate.py: <= 1 float32 ulp, on < 1e-6 of the samples).
#include "common.cuh"

struct ModRec {
  uint64_t data_off, n_bytes;     // payload bytes in `data`
  uint64_t sym_off, n_sym;        // symbols (bits for CPFSK) in the phase workspace
  uint64_t out_off;               // first output sample
};

__device__ __forceinline__ uint32_t mod_bit(const uint8_t* d, uint64_t i) { return (d[i >> 3] >> (7 - (int)(i & 7))) & 1u; }

// Python's `phase %= two_pi` on non-negative floats is C fmod: the exact remainder x - n y, always representable.  With
// n = floor(x / y) found by a multiply and corrected by at most one, fma(-n, y, x) rounds an exactly representable value,
// i.e. not at all -- the same result as fmod() at a fraction of its latency on the sequential walk.
__device__ __forceinline__ double mod_fmod_pos(double x, double y, double inv_y) {
  double n = floor(x * inv_y);
  double r = fma(-n, y, x);
  if (r < 0.0) { n -= 1.0; r = fma(-n, y, x); }
  else if (r >= y) { n += 1.0; r = fma(-n, y, x); }
  return r;
}

__global__ void __launch_bounds__(32) mod_phase_kernel(const fb_mod_params p, const ModRec* recs, int n_rec, const uint8_t* data,
                                                        double* phases) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rec) return;
  const ModRec rc = recs[r];
  const uint8_t* d = data + rc.data_off;
  double* ph = phases + rc.sym_off;
  // one payload byte per trip (its load does not depend on the phase chain, so it runs ahead of it)
  if (p.kind == FB_MOD_DBPSK) {
    // preamble [1, 0] * 40, then the payload bits MSB first; 1 -> += pi (modem.py:33-48)
    double cur = 0.0;
    for (int k = 0; k < 80; ++k) { if (~k & 1) cur += p.inc[1]; ph[k] = cur; }
    ph += 80;
    for (uint64_t i = 0; i < rc.n_bytes; ++i) {
      const uint32_t byte = d[i];
#pragma unroll
      for (int k = 0; k < 8; ++k) { if ((byte >> (7 - k)) & 1u) cur += p.inc[1]; ph[k] = cur; }
      ph += 8;
    }
  } else if (p.kind == FB_MOD_DQPSK) {
    // preamble [0,0]*30 + [1,1]*10, then dibits MSB first; phase change by 2*b0 + b1 (modem.py:150-174)
    double cur = 0.0;
    for (int k = 0; k < 40; ++k) { cur += p.inc[k < 30 ? 0 : 3]; ph[k] = cur; }
    ph += 40;
    const double i0 = p.inc[0], i1 = p.inc[1], i2 = p.inc[2], i3 = p.inc[3];
    for (uint64_t i = 0; i < rc.n_bytes; ++i) {
      const uint32_t byte = d[i];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t code = (byte >> (6 - 2 * k)) & 3u;
        cur += (code & 2u) ? ((code & 1u) ? i3 : i2) : ((code & 1u) ? i1 : i0);
        ph[k] = cur;
      }
      ph += 4;
    }
  } else {
    // CPFSK: preamble AA AA AA AA; the bit is generated with the phase carried in, then
    // phase += 2 pi f (spb / fs); phase %= 2 pi  (modem.py:283-293)
    const double two_pi = 6.283185307179586, inv = 1.0 / 6.283185307179586;
    const double inc0 = p.inc[0], inc1 = p.inc[1];
    double cur = 0.0;
    for (uint64_t i = 0; i < rc.n_bytes + 4; ++i) {
      const uint32_t byte = i < 4 ? 0xAAu : d[i - 4];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        ph[k] = cur;
        cur = mod_fmod_pos(__dadd_rn(cur, ((byte >> (7 - k)) & 1u) ? inc1 : inc0), two_pi, inv);
      }
      ph += 8;
    }
  }
}

#define MOD_THREADS 256
__global__ void __launch_bounds__(MOD_THREADS) mod_wave_kernel(const fb_mod_params p, const ModRec* recs, const double* tab /*[2][sps]*/,
                                                                const uint8_t* data, const double* phases, float* out) {
  extern __shared__ double s_tab[];                  // base | env
  const ModRec rc = recs[blockIdx.y];
  const int sps = p.sps;
  for (int i = threadIdx.x; i < 2 * sps; i += MOD_THREADS) s_tab[i] = tab[i];
  __syncthreads();
  const uint64_t n = rc.n_sym * (uint64_t)sps;
  const double* ph = phases + rc.sym_off;
  float* o = out + rc.out_off;
  const uint8_t* d = data + rc.data_off;
  for (uint64_t i = (uint64_t)blockIdx.x * MOD_THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * MOD_THREADS) {
    const uint64_t k = i / (uint32_t)sps;
    const int j = (int)(i - k * (uint32_t)sps);
    if (p.kind == FB_MOD_CPFSK) {
      const uint32_t bit = k < 32 ? (uint32_t)(~k & 1u) : mod_bit(d, k - 32);
      // product and sum rounded separately, as numpy does (no FMA contraction); float32 array * 0.9 (modem.py:295)
      o[i] = __fmul_rn((float)sin(__dadd_rn(__dmul_rn(p.wfreq[bit], s_tab[j]), ph[k])), p.gain);
    } else {
      o[i] = (float)__dmul_rn(sin(__dadd_rn(s_tab[j], ph[k])), s_tab[sps + j]);
    }
  }
}

extern "C" uint64_t fb_mod_out_samples(const fb_mod_params* p, uint64_t n_bytes) {
  if (!p || p->sps < 0) return 0;
  const uint64_t nsym = p->kind == FB_MOD_DBPSK ? 80 + 8 * n_bytes : p->kind == FB_MOD_DQPSK ? 40 + 4 * n_bytes : 8 * (4 + n_bytes);
  return nsym * (uint64_t)p->sps;
}

extern "C" int fb_modulate_batch(fb_handle* h, const fb_mod_params* pp, const double* base, const double* env, int n_rec,
                                 const uint8_t* data, const uint64_t* data_offsets, float* out, const uint64_t* out_offsets, int flags) {
  if (!h || !pp || !base || n_rec < 0 || !data_offsets || !out_offsets) return FB_EINVAL;
  FB_LOCK(h);
  const fb_mod_params& p = *pp;
  if (p.kind < FB_MOD_DBPSK || p.kind > FB_MOD_CPFSK || p.sps < 0 || p.sps > 1 << 20) return FB_EINVAL;
  if (p.kind != FB_MOD_CPFSK && !env) return FB_EINVAL;
  FB_CUDA(h, cudaSetDevice(h->device));
  if (n_rec == 0 || p.sps == 0) return FB_OK;
  std::vector<ModRec> recs(n_rec);
  uint64_t syms = 0, max_n = 0;
  for (int r = 0; r < n_rec; ++r) {
    ModRec& q = recs[r];
    q.data_off = data_offsets[r]; q.n_bytes = data_offsets[r + 1] - data_offsets[r];
    q.n_sym = fb_mod_out_samples(pp, q.n_bytes) / (uint64_t)p.sps;
    q.sym_off = syms; syms += q.n_sym;
    q.out_off = out_offsets[r];
    if (out_offsets[r + 1] - out_offsets[r] < q.n_sym * (uint64_t)p.sps) return FB_EINVAL;
    max_n = std::max<uint64_t>(max_n, q.n_sym * (uint64_t)p.sps);
  }
  const uint64_t total_data = data_offsets[n_rec], total_out = out_offsets[n_rec];
  int rc;
  const uint8_t* d_data = data;
  if (!(flags & FB_SAMPLES_ON_DEVICE)) {
    if ((rc = fb_ensure(h, h->fec_in, (size_t)total_data + 16))) return rc;
    if (total_data) FB_CUDA(h, cudaMemcpyAsync(h->fec_in.p, data, (size_t)total_data, cudaMemcpyHostToDevice, h->stream));
    d_data = (const uint8_t*)h->fec_in.p;
  }
  float* d_out = out;
  if (!(flags & FB_OUT_ON_DEVICE)) {
    if ((rc = fb_ensure(h, h->in, (size_t)total_out * 4 + 16))) return rc;
    d_out = (float*)h->in.p;
  }
  // workspace: [phases f64][tables 2*sps f64][ModRec]
  const size_t o_tab = ((size_t)syms * 8 + 255) / 256 * 256, o_rec = o_tab + ((size_t)2 * p.sps * 8 + 255) / 256 * 256;
  if ((rc = fb_ensure(h, h->scratch, o_rec + (size_t)n_rec * sizeof(ModRec) + 16))) return rc;
  double* d_ph = (double*)h->scratch.p;
  double* d_tab = (double*)((char*)h->scratch.p + o_tab);
  ModRec* d_recs = (ModRec*)((char*)h->scratch.p + o_rec);
  std::vector<double> tab((size_t)2 * p.sps, 1.0);
  for (int j = 0; j < p.sps; ++j) { tab[j] = base[j]; if (env) tab[p.sps + j] = env[j]; }
  FB_CUDA(h, cudaMemcpyAsync(d_tab, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, h->stream));
  FB_CUDA(h, cudaMemcpyAsync(d_recs, recs.data(), (size_t)n_rec * sizeof(ModRec), cudaMemcpyHostToDevice, h->stream));
  mod_phase_kernel<<<(n_rec + 31) / 32, 32, 0, h->stream>>>(p, d_recs, n_rec, d_data, d_ph);
  const size_t smem = (size_t)2 * p.sps * 8;
  if (smem > 48 * 1024) FB_CUDA(h, cudaFuncSetAttribute(mod_wave_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int r0 = 0; r0 < n_rec; r0 += 65535) {
    const int nr = std::min(65535, n_rec - r0);
    const unsigned gx = (unsigned)std::min<uint64_t>((max_n + MOD_THREADS - 1) / MOD_THREADS, (uint64_t)h->sm_count * 32);
    mod_wave_kernel<<<dim3(std::max(1u, gx), nr), MOD_THREADS, smem, h->stream>>>(p, d_recs + r0, d_tab, d_data, d_ph, d_out);
    h->launches++;
  }
  h->launches++;
  FB_CUDA(h, cudaGetLastError());
  if (!(flags & FB_OUT_ON_DEVICE) && total_out)
    FB_CUDA(h, cudaMemcpyAsync(out, d_out, (size_t)total_out * 4, cudaMemcpyDeviceToHost, h->stream));
  if (!(flags & FB_ASYNC) || !(flags & FB_OUT_ON_DEVICE)) FB_CUDA(h, cudaStreamSynchronize(h->stream));
  return FB_OK;
}
