// Block-parallel decode of the reference's fec.py codes (neither is a real Reed-Solomon / Viterbi decoder;
// these kernels compute exactly what the reference computes):
//   rs_decode_kernel       ReedSolomonFEC.decode (fec.py:34-69): (b1, b2, parity) triples -> (b1, b2) or (b1, 0x3F)
//                          on a parity mismatch, raw tail bytes, trailing CRC32 (LE) checked against the decoded block
//   viterbi_decode_kernel  ViterbiDecoder.decode (fec.py:126-155): MSB-first bits, drop the last 12, keep the
//                          even-index bits, repack MSB-first with the final partial byte right-aligned
//   crc32_kernel           zlib.crc32 of arbitrary byte ranges (frame payloads, decoder.py:194)
// One CTA per block; blocks are independent (CSR offsets).  HBM-bound byte work: n read + ~2n/3 (n/2) written.
#include "common.cuh"
#include "crc32.cuh"

struct FecBlock {
  uint64_t in_off, in_len, out_off, out_cap;
};

__global__ void __launch_bounds__(FB_THREADS) rs_decode_kernel(const FecBlock* blocks, const uint8_t* in, uint8_t* out,
                                                                uint64_t* out_len, int32_t* crc_ok) {
  __shared__ uint32_t tab[1024];
  __shared__ uint32_t scratch[33];
  const FecBlock b = blocks[blockIdx.x];
  const uint8_t* src = in + b.in_off;
  uint8_t* dst = out + b.out_off;
  const uint64_t n = b.in_len;
  if (n < 4) {                                               // fec.py:36-37: returned unchanged
    for (uint64_t i = threadIdx.x; i < n && i < b.out_cap; i += blockDim.x) dst[i] = src[i];
    if (threadIdx.x == 0) { out_len[blockIdx.x] = min(n, b.out_cap); crc_ok[blockIdx.x] = 1; }
    return;
  }
  crc_tables_init(tab);
  const uint64_t m = n - 4, ntr = m / 3, tail = m - 3 * ntr;
  const uint64_t olen = 2 * ntr + tail;
  if (olen > b.out_cap) {                                    // caller's slot too small: report, write nothing
    if (threadIdx.x == 0) { out_len[blockIdx.x] = 0; crc_ok[blockIdx.x] = -1; }
    return;
  }
  // 4 triples (12 bytes in, 8 bytes out) per thread and step
  for (uint64_t g = threadIdx.x; g * 4 < ntr; g += blockDim.x) {
    const uint64_t t0 = g * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t t = t0 + k;
      if (t < ntr) {
        const uint8_t b1 = src[3 * t], b2 = src[3 * t + 1], p = src[3 * t + 2];
        dst[2 * t] = b1;
        dst[2 * t + 1] = ((b1 ^ b2) == p) ? b2 : (uint8_t)0x3F;        // fec.py:53-57
      }
    }
  }
  if (threadIdx.x < tail) dst[2 * ntr + threadIdx.x] = src[3 * ntr + threadIdx.x];   // fec.py:60-62
  __syncthreads();                                                      // dst is complete and visible to the whole CTA
  const uint32_t crc = block_crc32<false>(tab, scratch, dst, olen);     // fec.py:65; coherent loads: dst was written by this kernel
  if (threadIdx.x == 0) {
    const uint32_t want = (uint32_t)src[m] | ((uint32_t)src[m + 1] << 8) | ((uint32_t)src[m + 2] << 16) | ((uint32_t)src[m + 3] << 24);
    out_len[blockIdx.x] = olen;
    crc_ok[blockIdx.x] = (crc == want) ? 1 : 0;
  }
}

__device__ __forceinline__ uint32_t even_bits16(uint32_t w) {   // bits 15,13,...,1 of a 16-bit word -> 8 bits, MSB first
  uint32_t x = (w & 0xAAAAu) >> 1;
  x = (x | (x >> 1)) & 0x3333u;
  x = (x | (x >> 2)) & 0x0F0Fu;
  x = (x | (x >> 4)) & 0x00FFu;
  return x;
}

__global__ void __launch_bounds__(FB_THREADS) viterbi_decode_kernel(const FecBlock* blocks, const uint8_t* in, uint8_t* out,
                                                                     uint64_t* out_len) {
  const FecBlock b = blocks[blockIdx.x];
  const uint8_t* src = in + b.in_off;
  uint8_t* dst = out + b.out_off;
  uint64_t nbits = b.in_len * 8;
  if (nbits >= 12) nbits -= 12;                              // fec.py:138-139
  const uint64_t used = (nbits + 1) / 2;                     // even-index bits, fec.py:144-146
  const uint64_t full = used / 8, rem = used - 8 * full;
  const uint64_t olen = full + (rem ? 1 : 0);
  if (olen > b.out_cap) {
    if (threadIdx.x == 0) out_len[blockIdx.x] = 0;
    return;
  }
  for (uint64_t o = threadIdx.x; o < full; o += blockDim.x)
    dst[o] = (uint8_t)even_bits16(((uint32_t)src[2 * o] << 8) | src[2 * o + 1]);
  if (threadIdx.x == 0) {
    if (rem) {                                               // fec.py:148-153: missing bits are skipped, not padded
      const uint32_t hi = src[2 * full];
      const uint32_t lo = (2 * full + 1 < b.in_len) ? src[2 * full + 1] : 0u;
      dst[full] = (uint8_t)(even_bits16((hi << 8) | lo) >> (8 - rem));
    }
    out_len[blockIdx.x] = olen;
  }
}

__global__ void __launch_bounds__(FB_THREADS) crc32_kernel(const FecBlock* blocks, const uint8_t* data, uint32_t* crc) {
  __shared__ uint32_t tab[1024];
  __shared__ uint32_t scratch[33];
  crc_tables_init(tab);
  const FecBlock b = blocks[blockIdx.x];
  const uint32_t c = block_crc32(tab, scratch, data + b.in_off, b.in_len);
  if (threadIdx.x == 0) crc[blockIdx.x] = c;
}

// ------------------------------------------------------------------------------------------------ host side
extern "C" uint64_t fb_rs_out_bound(uint64_t n) { return n < 4 ? n : 2 * ((n - 4) / 3) + (n - 4) % 3; }
extern "C" uint64_t fb_viterbi_out_bound(uint64_t n) {
  uint64_t nbits = n * 8;
  if (nbits >= 12) nbits -= 12;
  return ((nbits + 1) / 2 + 7) / 8;
}

enum { FEC_RS = 0, FEC_VIT = 1, FEC_CRC = 2 };

static int fec_run(fb_handle* h, int op, int n_blk, std::vector<FecBlock>& blocks, uint64_t total_in, uint64_t total_out, const uint8_t* in,
                   uint8_t* out, uint64_t* out_len, int32_t* aux, int flags);

static int fec_batch(fb_handle* h, int op, int n_blk, const uint8_t* in, const uint64_t* in_offsets, uint8_t* out,
                     const uint64_t* out_offsets, uint64_t* out_len, int32_t* aux, int flags) {
  if (!h || n_blk < 0 || !in_offsets || (op != FEC_CRC && !out_offsets)) return FB_EINVAL;
  FB_LOCK(h);
  FB_CUDA(h, cudaSetDevice(h->device));
  if (n_blk == 0) return FB_OK;
  std::vector<FecBlock> blocks(n_blk);
  for (int i = 0; i < n_blk; ++i) {
    blocks[i].in_off = in_offsets[i];
    blocks[i].in_len = in_offsets[i + 1] - in_offsets[i];
    blocks[i].out_off = out_offsets ? out_offsets[i] : 0;
    blocks[i].out_cap = out_offsets ? out_offsets[i + 1] - out_offsets[i] : 0;
  }
  return fec_run(h, op, n_blk, blocks, in_offsets[n_blk], out_offsets ? out_offsets[n_blk] : 0, in, out, out_len, aux, flags);
}

static int fec_run(fb_handle* h, int op, int n_blk, std::vector<FecBlock>& blocks, uint64_t total_in, uint64_t total_out, const uint8_t* in,
                   uint8_t* out, uint64_t* out_len, int32_t* aux, int flags) {
  int rc;
  if ((rc = fb_ensure(h, h->fec_meta, (size_t)n_blk * sizeof(FecBlock)))) return rc;
  FB_CUDA(h, cudaMemcpyAsync(h->fec_meta.p, blocks.data(), (size_t)n_blk * sizeof(FecBlock), cudaMemcpyHostToDevice, h->stream));
  const uint8_t* d_in = in;
  if (!(flags & FB_SAMPLES_ON_DEVICE)) {
    if ((rc = fb_ensure(h, h->fec_in, (size_t)total_in + 16))) return rc;
    if (total_in) FB_CUDA(h, cudaMemcpyAsync(h->fec_in.p, in, (size_t)total_in, cudaMemcpyHostToDevice, h->stream));
    d_in = (const uint8_t*)h->fec_in.p;
  }
  uint8_t* d_out = out; uint64_t* d_len = out_len; int32_t* d_aux = aux;
  const size_t aux_sz = (size_t)n_blk * 4, len_sz = (size_t)n_blk * 8;
  if (!(flags & FB_OUT_ON_DEVICE)) {
    // one workspace: [out bytes][pad][out_len][aux]
    const size_t o_len = ((size_t)total_out + 15) / 16 * 16, o_aux = o_len + len_sz;
    if ((rc = fb_ensure(h, h->fec_out, o_aux + aux_sz + 16))) return rc;
    d_out = (uint8_t*)h->fec_out.p; d_len = (uint64_t*)((char*)h->fec_out.p + o_len); d_aux = (int32_t*)((char*)h->fec_out.p + o_aux);
  }
  const FecBlock* d_blocks = (const FecBlock*)h->fec_meta.p;
  if (op == FEC_RS) rs_decode_kernel<<<n_blk, FB_THREADS, 0, h->stream>>>(d_blocks, d_in, d_out, d_len, d_aux);
  else if (op == FEC_VIT) viterbi_decode_kernel<<<n_blk, FB_THREADS, 0, h->stream>>>(d_blocks, d_in, d_out, d_len);
  else crc32_kernel<<<n_blk, FB_THREADS, 0, h->stream>>>(d_blocks, d_in, (uint32_t*)d_aux);
  h->launches++;
  FB_CUDA(h, cudaGetLastError());
  if (!(flags & FB_OUT_ON_DEVICE)) {
    if (op != FEC_CRC) {
      if (total_out) FB_CUDA(h, cudaMemcpyAsync(out, d_out, (size_t)total_out, cudaMemcpyDeviceToHost, h->stream));
      FB_CUDA(h, cudaMemcpyAsync(out_len, d_len, len_sz, cudaMemcpyDeviceToHost, h->stream));
    }
    if (op != FEC_VIT && aux) FB_CUDA(h, cudaMemcpyAsync(aux, d_aux, aux_sz, cudaMemcpyDeviceToHost, h->stream));
  }
  if (!(flags & FB_ASYNC) || !(flags & FB_OUT_ON_DEVICE)) FB_CUDA(h, cudaStreamSynchronize(h->stream));
  return FB_OK;
}

extern "C" int fb_rs_decode_batch(fb_handle* h, int n_blk, const uint8_t* in, const uint64_t* in_offsets, uint8_t* out,
                                  const uint64_t* out_offsets, uint64_t* out_len, int32_t* crc_ok, int flags) {
  return fec_batch(h, FEC_RS, n_blk, in, in_offsets, out, out_offsets, out_len, crc_ok, flags);
}

// Blocks given as (start, length) spans inside `in` -- e.g. the payloads of the frames fb_parse_frames_batch found in the raw
// streams still resident on the device -- instead of CSR offsets: no gather copy between the parser and the decoder.
extern "C" int fb_rs_decode_spans(fb_handle* h, int n_blk, const uint8_t* in, const uint64_t* in_start, const uint64_t* in_len,
                                  uint8_t* out, const uint64_t* out_offsets, uint64_t* out_len, int32_t* crc_ok, int flags) {
  if (!h || n_blk < 0 || !in_start || !in_len || !out_offsets || !(flags & FB_SAMPLES_ON_DEVICE)) return FB_EINVAL;
  FB_LOCK(h);
  FB_CUDA(h, cudaSetDevice(h->device));
  if (n_blk == 0) return FB_OK;
  std::vector<FecBlock> blocks(n_blk);
  for (int i = 0; i < n_blk; ++i) {
    blocks[i].in_off = in_start[i];
    blocks[i].in_len = in_len[i];
    blocks[i].out_off = out_offsets[i];
    blocks[i].out_cap = out_offsets[i + 1] - out_offsets[i];
  }
  return fec_run(h, FEC_RS, n_blk, blocks, 0, out_offsets[n_blk], in, out, out_len, crc_ok, flags);
}

extern "C" int fb_viterbi_decode_batch(fb_handle* h, int n_blk, const uint8_t* in, const uint64_t* in_offsets, uint8_t* out,
                                       const uint64_t* out_offsets, uint64_t* out_len, int flags) {
  return fec_batch(h, FEC_VIT, n_blk, in, in_offsets, out, out_offsets, out_len, nullptr, flags);
}

extern "C" int fb_crc32_batch(fb_handle* h, int n_blk, const uint8_t* data, const uint64_t* offsets, uint32_t* crc, int flags) {
  return fec_batch(h, FEC_CRC, n_blk, data, offsets, nullptr, nullptr, nullptr, (int32_t*)crc, flags);
}
