// Bit-stream back end shared by every v2 demodulator (modem.py:111-135, 244-266, 326-341):
//   sync_search_kernel  first occurrence of the 16-bit pattern 0100011001000010 ("FB") in the decided
//                       bit stream  ==  bit_str.find(magic): a min-reduction over matching bit offsets
//   pack_bytes_kernel   bytes from the sync offset (or from bit 0 when absent), MSB first:
//                       an unaligned funnel-shift gather of the packed words
// The decided bits live in the workspace as big-endian 32-bit words (memory order == stream order).
#include "common.cuh"

#define FB_MAGIC16 0x4642u   // 0100 0110 0100 0010

__global__ void __launch_bounds__(FB_THREADS) sync_init_kernel(unsigned long long* sync_raw, int n_rec) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_rec) sync_raw[i] = ~0ull;
}

// Words [w_lo, w_hi) of every recording.  Two launches: the head (where the preamble puts the magic of any real
// recording), then the rest -- whose CTAs return at once for recordings the head already settled.
__global__ void __launch_bounds__(FB_THREADS) sync_search_kernel(const RecPlan* plans, int bps, const uint32_t* bits,
                                                                  unsigned long long* sync_raw, uint64_t w_lo, uint64_t w_hi) {
  const RecPlan pl = plans[blockIdx.y];
  const uint64_t nbits = (uint64_t)pl.ndsym * bps;
  if (nbits < 16) return;
  if (w_lo > 0 && sync_raw[blockIdx.y] != ~0ull) return;          // found in the head: nothing further out can be first
  const uint64_t nwords = min((nbits + 31) / 32, w_hi);
  const uint32_t* w = bits + pl.word_off;
  unsigned long long best = ~0ull;
  for (uint64_t i = w_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x) {
    if (sync_raw[blockIdx.y] < i * 32) break;        // an earlier match is already known (monotone, benign race)
    const uint32_t hi = __byte_perm(w[i], 0, 0x0123);
    const uint32_t lo = (i + 1 < (nbits + 31) / 32) ? __byte_perm(w[i + 1], 0, 0x0123) : 0u;
    const uint64_t win = ((uint64_t)hi << 32) | lo;
#pragma unroll 8
    for (int o = 0; o < 32; ++o) {
      if ((uint32_t)((win >> (48 - o)) & 0xFFFFu) == FB_MAGIC16) {
        const uint64_t pos = i * 32 + o;
        if (pos + 16 <= nbits && pos < best) best = pos;
      }
    }
  }
  if (best != ~0ull) atomicMin(&sync_raw[blockIdx.y], best);
}

__global__ void __launch_bounds__(FB_THREADS) pack_bytes_kernel(const RecPlan* plans, int bps, const uint32_t* bits,
                                                                 const unsigned long long* sync_raw, uint8_t* out,
                                                                 uint64_t* out_len, int64_t* sync_idx, int32_t* status) {
  const int r = blockIdx.y;
  const RecPlan pl = plans[r];
  const uint64_t nbits = (uint64_t)pl.ndsym * bps;
  const unsigned long long s = sync_raw[r];
  const uint64_t start = (s == ~0ull) ? 0 : s;
  uint64_t nbytes = (nbits - start) / 8;             // floor: "for i in range(0, len(valid) - 7, 8)"
  if (nbytes > pl.out_cap) nbytes = pl.out_cap;      // never write outside the caller's slot
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    out_len[r] = nbytes;
    sync_idx[r] = (s == ~0ull) ? -1 : (int64_t)s;
    status[r] = pl.status;
  }
  const uint32_t* w = bits + pl.word_off;
  uint8_t* o = out + pl.out_off;
  const uint64_t ngroups = (nbytes + 3) / 4;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t p = start + 32 * g;
    const uint64_t wi = p >> 5;
    const uint32_t sh = (uint32_t)(p & 31);
    const uint32_t hi = __byte_perm(w[wi], 0, 0x0123);
    const uint32_t lo = __byte_perm(w[wi + 1], 0, 0x0123);     // workspace keeps 2 spare words per recording
    const uint32_t v = __funnelshift_l(lo, hi, sh);
    const uint64_t b0 = 4 * g;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (b0 + k < nbytes) o[b0 + k] = (uint8_t)(v >> (24 - 8 * k));
  }
}

int fb_bits_backend(fb_handle* h, int n_rec, const RecPlan* d_plans, const std::vector<RecPlan>& plans, int bps,
                    const uint32_t* d_bits, uint8_t* d_out, uint64_t* d_out_len, int64_t* d_sync, int32_t* d_status) {
  int rc = fb_ensure(h, h->sync_raw, (size_t)n_rec * 8);
  if (rc) return rc;
  uint64_t max_words = 1;
  for (const RecPlan& p : plans) max_words = std::max<uint64_t>(max_words, ((uint64_t)p.ndsym * bps + 31) / 32);
  unsigned long long* sr = (unsigned long long*)h->sync_raw.p;
  sync_init_kernel<<<(n_rec + FB_THREADS - 1) / FB_THREADS, FB_THREADS, 0, h->stream>>>(sr, n_rec);
  // recordings on grid.y (<= 65535 per launch), words / byte groups strided on grid.x
  const int gx_search = (int)std::min<uint64_t>(64, (max_words + FB_THREADS - 1) / FB_THREADS);
  const int gx_pack = (int)std::min<uint64_t>(64, (max_words + FB_THREADS - 1) / FB_THREADS);
  for (int r0 = 0; r0 < n_rec; r0 += 65535) {
    const int nr = std::min(65535, n_rec - r0);
    const uint64_t head = 4096;                                   // words: the first 131 072 bits
    sync_search_kernel<<<dim3((unsigned)std::min<uint64_t>(gx_search, head / FB_THREADS), nr), FB_THREADS, 0, h->stream>>>(d_plans + r0, bps, d_bits,
                                                                                                                          sr + r0, 0, head);
    if (max_words > head) {
      sync_search_kernel<<<dim3(gx_search, nr), FB_THREADS, 0, h->stream>>>(d_plans + r0, bps, d_bits, sr + r0, head, ~0ull);
      h->launches += 1;
    }
    pack_bytes_kernel<<<dim3(gx_pack, nr), FB_THREADS, 0, h->stream>>>(d_plans + r0, bps, d_bits, sr + r0, d_out,
                                                                        d_out_len + r0, d_sync + r0, d_status + r0);
    h->launches += 2;
  }
  h->launches += 1;
  FB_CUDA(h, cudaGetLastError());
  return FB_OK;
}
