// FBPC frame parser on the device: replaces decoder.parse_fbp_stream_enhanced (decoder.py:142-208) for the
// demodulated byte stream of every recording of a batch, without a host bounce.
//   One CTA per recording: (1) overlapping scan for b"FBPC" (decoder.py:155-159), candidates kept in offset
//   order; (2) per candidate the reference's header checks in the reference's order (decoder.py:166-189);
//   (3) CRC32 of the payload by the whole CTA (decoder.py:194) -- accepted frames are appended to the
//   recording's slot of the frame table.  Byte work, ~2 passes over <= 0.6 % of the input bytes.
#include "common.cuh"
#include "crc32.cuh"

#define MAX_CAND 64

struct ParseRec {
  uint64_t raw_off;      // first byte of this recording's raw stream
  uint64_t raw_cap;      // slot size (upper bound of the length)
};

__global__ void __launch_bounds__(FB_THREADS) parse_frames_kernel(const ParseRec* recs, const uint8_t* raw,
                                                                   const uint64_t* raw_len, int max_frames,
                                                                   fb_frame* frames, int32_t* n_frames,
                                                                   uint64_t* payload_bytes) {
  __shared__ uint32_t tab[1024];
  __shared__ uint32_t scratch[33];
  __shared__ unsigned long long cand[MAX_CAND];
  __shared__ int n_cand;
  const int r = blockIdx.x;
  const ParseRec pr = recs[r];
  const uint64_t len = min(raw_len[r], pr.raw_cap);
  const uint8_t* p = raw + pr.raw_off;
  if (threadIdx.x == 0) n_cand = 0;
  crc_tables_init(tab);
  // ---- (1) all offsets i with p[i..i+4) == "FBPC" (overlapping; the pattern cannot overlap itself) ----------
  // 16 bytes per thread and step (one LDG.128 once the pointer is aligned) plus the 3-byte look-ahead from the next
  // chunk; four steps unrolled so several loads are in flight per thread (the byte-per-lane loop was latency-bound).
  {
    const uint64_t head = min(len, (uint64_t)((16 - ((uintptr_t)p & 15)) & 15));       // bytes before the first 16-byte boundary
    auto hit = [&](uint64_t i) {
      const int slot = atomicAdd(&n_cand, 1);
      if (slot < MAX_CAND) cand[slot] = i;
    };
    for (uint64_t i = threadIdx.x; i < head && i + 4 <= len; i += blockDim.x)
      if (p[i] == 'F' && p[i + 1] == 'B' && p[i + 2] == 'P' && p[i + 3] == 'C') hit(i);
    const uint64_t nchunk = (len - head) / 16;                                          // whole aligned chunks
    const uint4* q = reinterpret_cast<const uint4*>(p + head);
#pragma unroll 4
    for (uint64_t c = threadIdx.x; c < nchunk; c += blockDim.x) {
      const uint4 v = __ldg(q + c);
      const uint64_t base = head + c * 16;
      uint32_t nxt = 0;                                                                 // the 3 bytes after this chunk
      if (c + 1 < nchunk) nxt = __ldg(reinterpret_cast<const uint32_t*>(q + c + 1));
      else {
        for (int k = 0; k < 3; ++k) if (base + 16 + k < len) nxt |= (uint32_t)p[base + 16 + k] << (8 * k);
      }
      const uint32_t w[5] = {v.x, v.y, v.z, v.w, nxt};
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int wi = k >> 2, sh = (k & 3) * 8;
        const uint32_t four = sh ? (w[wi] >> sh) | (w[wi + 1] << (32 - sh)) : w[wi];
        if (four == 0x43504246u && base + k + 4 <= len) hit(base + k);                  // "FBPC" little-endian
      }
    }
    for (uint64_t i = head + nchunk * 16 + threadIdx.x; i + 4 <= len; i += blockDim.x)  // ragged tail (< 16 bytes)
      if (p[i] == 'F' && p[i + 1] == 'B' && p[i + 2] == 'P' && p[i + 3] == 'C') hit(i);
  }
  __syncthreads();
  const int nc = min(n_cand, MAX_CAND);
  if (threadIdx.x == 0) {                                   // ascending offsets, as raw.find() enumerates them
    for (int a = 1; a < nc; ++a) {
      const unsigned long long v = cand[a];
      int b = a - 1;
      while (b >= 0 && cand[b] > v) { cand[b + 1] = cand[b]; --b; }
      cand[b + 1] = v;
    }
  }
  __syncthreads();
  int nf = 0;
  uint64_t pbytes = 0;
  // the reference's header checks in the reference's order, then the payload CRC32 by the whole CTA (uniform control flow)
  auto try_candidate = [&](uint64_t start) {
    if (start + 30 > len) return;                           // decoder.py:166
    const uint32_t name_len = p[start + 4];
    if (name_len == 0) return;                              // decoder.py:170
    const uint64_t meta = start + 5 + name_len;
    if (meta + 24 > len) return;                            // decoder.py:179
    uint32_t f[6];
#pragma unroll
    for (int k = 0; k < 6; ++k)
      f[k] = (uint32_t)p[meta + 4 * k] | ((uint32_t)p[meta + 4 * k + 1] << 8) | ((uint32_t)p[meta + 4 * k + 2] << 16) |
             ((uint32_t)p[meta + 4 * k + 3] << 24);
    const uint32_t dlen = f[4], pcrc = f[5];
    if (dlen > 50000000u || dlen == 0) return;              // decoder.py:184
    const uint64_t pay = meta + 24;
    if (pay + dlen > len) return;                           // decoder.py:187
    const uint32_t crc = block_crc32(tab, scratch, p + pay, dlen);
    if (crc != pcrc) return;                                // decoder.py:195
    if (threadIdx.x == 0 && nf < max_frames) {
      fb_frame& o = frames[(size_t)r * max_frames + nf];
      o.offset = start; o.name_off = start + 5; o.payload_off = pay;
      o.name_len = name_len; o.part = f[0]; o.total = f[1]; o.file_size = f[2]; o.file_crc = f[3];
      o.data_len = dlen; o.payload_crc = pcrc;
    }
    ++nf;
    pbytes += dlen;
  };
  if (n_cand <= MAX_CAND) {
    for (int ci = 0; ci < nc; ++ci) try_candidate(cand[ci]);
  } else {
    // More "FBPC" hits than the table holds (a payload that contains the magic many times, e.g. this project's own
    // sources): walk the stream in order, MAX_CAND * 4 bytes at a time -- a window of that size cannot hold more than
    // MAX_CAND non-overlapping 4-byte patterns -- so no candidate is ever dropped (the reference parser has no limit).
    __shared__ int win_cnt;
    const uint64_t WIN = (uint64_t)MAX_CAND * 4;
    for (uint64_t w0 = 0; w0 + 4 <= len; w0 += WIN) {
      __syncthreads();
      if (threadIdx.x == 0) win_cnt = 0;
      __syncthreads();
      for (uint64_t i = w0 + threadIdx.x; i < w0 + WIN && i + 4 <= len; i += blockDim.x)
        if (p[i] == 'F' && p[i + 1] == 'B' && p[i + 2] == 'P' && p[i + 3] == 'C') cand[atomicAdd(&win_cnt, 1)] = i;
      __syncthreads();
      const int wc = win_cnt;
      if (threadIdx.x == 0) {
        for (int a = 1; a < wc; ++a) {
          const unsigned long long v = cand[a];
          int b = a - 1;
          while (b >= 0 && cand[b] > v) { cand[b + 1] = cand[b]; --b; }
          cand[b + 1] = v;
        }
      }
      __syncthreads();
      for (int ci = 0; ci < wc; ++ci) try_candidate(cand[ci]);
    }
  }
  if (threadIdx.x == 0) {
    n_frames[r] = nf;                                       // may exceed max_frames: only the first max_frames are stored
    payload_bytes[r] = pbytes;
  }
}

extern "C" int fb_parse_frames_batch(fb_handle* h, int n_rec, const uint8_t* raw, const uint64_t* raw_offsets,
                                     const uint64_t* raw_len, int max_frames, fb_frame* frames, int32_t* n_frames,
                                     uint64_t* payload_bytes, int flags) {
  if (!h || n_rec < 0 || !raw_offsets || !raw_len || max_frames < 1 || !frames || !n_frames || !payload_bytes) return FB_EINVAL;
  FB_LOCK(h);
  FB_CUDA(h, cudaSetDevice(h->device));
  if (n_rec == 0) return FB_OK;
  std::vector<ParseRec> recs(n_rec);
  for (int i = 0; i < n_rec; ++i) { recs[i].raw_off = raw_offsets[i]; recs[i].raw_cap = raw_offsets[i + 1] - raw_offsets[i]; }
  const uint64_t total = raw_offsets[n_rec];
  int rc;
  // workspace: [ParseRec table][raw_len][frames][n_frames][payload_bytes]  (+ raw bytes when they come from the host)
  const size_t o_len = (size_t)n_rec * sizeof(ParseRec), o_fr = o_len + (size_t)n_rec * 8,
               o_nf = o_fr + (size_t)n_rec * max_frames * sizeof(fb_frame), o_pb = (o_nf + (size_t)n_rec * 4 + 7) / 8 * 8,
               o_raw = o_pb + (size_t)n_rec * 8;
  const bool host_in = !(flags & FB_SAMPLES_ON_DEVICE), host_out = !(flags & FB_OUT_ON_DEVICE);
  if ((rc = fb_ensure(h, h->misc, o_raw + (host_in ? (size_t)total + 16 : 16)))) return rc;
  char* ws = (char*)h->misc.p;
  FB_CUDA(h, cudaMemcpyAsync(ws, recs.data(), o_len, cudaMemcpyHostToDevice, h->stream));
  const uint8_t* d_raw = raw;
  const uint64_t* d_len = raw_len;
  if (host_in) {
    if (total) FB_CUDA(h, cudaMemcpyAsync(ws + o_raw, raw, (size_t)total, cudaMemcpyHostToDevice, h->stream));
    FB_CUDA(h, cudaMemcpyAsync(ws + o_len, raw_len, (size_t)n_rec * 8, cudaMemcpyHostToDevice, h->stream));
    d_raw = (const uint8_t*)(ws + o_raw);
    d_len = (const uint64_t*)(ws + o_len);
  }
  fb_frame* d_fr = host_out ? (fb_frame*)(ws + o_fr) : frames;
  int32_t* d_nf = host_out ? (int32_t*)(ws + o_nf) : n_frames;
  uint64_t* d_pb = host_out ? (uint64_t*)(ws + o_pb) : payload_bytes;
  parse_frames_kernel<<<n_rec, FB_THREADS, 0, h->stream>>>((const ParseRec*)ws, d_raw, d_len, max_frames, d_fr, d_nf, d_pb);
  h->launches++;
  FB_CUDA(h, cudaGetLastError());
  if (host_out) {
    FB_CUDA(h, cudaMemcpyAsync(frames, d_fr, (size_t)n_rec * max_frames * sizeof(fb_frame), cudaMemcpyDeviceToHost, h->stream));
    FB_CUDA(h, cudaMemcpyAsync(n_frames, d_nf, (size_t)n_rec * 4, cudaMemcpyDeviceToHost, h->stream));
    FB_CUDA(h, cudaMemcpyAsync(payload_bytes, d_pb, (size_t)n_rec * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  if (!(flags & FB_ASYNC) || host_out) FB_CUDA(h, cudaStreamSynchronize(h->stream));
  return FB_OK;
}
