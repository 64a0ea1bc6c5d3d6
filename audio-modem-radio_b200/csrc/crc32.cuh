// CRC-32 (zlib / binascii.crc32: reflected 0xEDB88320, init and final xor 0xFFFFFFFF) for a whole CTA.
// Every thread reduces one contiguous segment with a slice-by-4 table walk (state 0 = pure polynomial
// remainder); segment remainders are aligned with x^(8 * bytes_after) mod P (square-and-multiply in GF(2))
// and xor-reduced.  Used by the frame parser (decoder.py:194) and the RS block check (fec.py:65).
#pragma once
#include <stdint.h>

#define CRC_POLY 0xEDB88320u

// a(x) * b(x) mod P(x), reflected bit order (bit 31 = x^0)
__device__ __forceinline__ uint32_t gf2_mulmod(uint32_t a, uint32_t b) {
  uint32_t r = 0;
#pragma unroll 4
  for (int i = 0; i < 32; ++i) {
    r ^= (b & 0x80000000u) ? a : 0u;
    a = (a >> 1) ^ ((a & 1u) ? CRC_POLY : 0u);
    b <<= 1;
  }
  return r;
}

// x^(8 * nbytes) mod P
__device__ __forceinline__ uint32_t gf2_xpow8n(uint64_t nbytes) {
  uint32_t r = 0x80000000u;            // x^0
  uint32_t sq = 0x00800000u;           // x^8
  while (nbytes) {
    if (nbytes & 1) r = gf2_mulmod(r, sq);
    sq = gf2_mulmod(sq, sq);
    nbytes >>= 1;
  }
  return r;
}

// tab: 4 * 256 entries of shared memory; call once per CTA before crc use, followed by __syncthreads()
__device__ __forceinline__ void crc_tables_init(uint32_t* tab) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    uint32_t c = (uint32_t)i;
    for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1u) ? CRC_POLY : 0u);
    tab[i] = c;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    uint32_t c = tab[i];
    for (int t = 1; t < 4; ++t) {
      c = tab[c & 0xFF] ^ (c >> 8);
      tab[t * 256 + i] = c;
    }
  }
  __syncthreads();
}

// RO = true: the bytes are read-only for the whole kernel (true inputs) -> ld.global.nc; RO = false: the bytes were
// written earlier in the SAME kernel (rs_decode_kernel's decoded block): the non-coherent path is undefined for those,
// so they are read through L2 (ld.global.cg) after the writer's __syncthreads().
template <bool RO> __device__ __forceinline__ uint32_t crc_ld32(const uint32_t* p) { return RO ? __ldg(p) : __ldcg(p); }
template <bool RO> __device__ __forceinline__ uint4 crc_ld128(const uint4* p) { return RO ? __ldg(p) : __ldcg(p); }
template <bool RO> __device__ __forceinline__ uint8_t crc_ld8(const uint8_t* p) { return RO ? __ldg(p) : __ldcg(p); }

// remainder of data[0..len) starting from state 0
template <bool RO = true>
__device__ __forceinline__ uint32_t crc_raw_segment(const uint32_t* tab, const uint8_t* p, uint64_t len) {
  uint32_t c = 0;
  while (len && ((uintptr_t)p & 3)) { c = tab[(c ^ crc_ld8<RO>(p++)) & 0xFF] ^ (c >> 8); --len; }
  const uint32_t* p4 = reinterpret_cast<const uint32_t*>(p);
#define CRC_WORD(wv)                                                                                              \
  do {                                                                                                            \
    c ^= (wv);                                                                                                    \
    c = tab[768 + (c & 0xFF)] ^ tab[512 + ((c >> 8) & 0xFF)] ^ tab[256 + ((c >> 16) & 0xFF)] ^ tab[c >> 24];      \
  } while (0)
  while (len >= 4 && ((uintptr_t)p4 & 15)) { CRC_WORD(crc_ld32<RO>(p4)); ++p4; len -= 4; }
  // 32 bytes per trip: both 16-byte loads are issued before the (serial) table walk over their eight words
  for (; len >= 32; len -= 32) {
    const uint4 a = crc_ld128<RO>(reinterpret_cast<const uint4*>(p4)), b = crc_ld128<RO>(reinterpret_cast<const uint4*>(p4) + 1);
    CRC_WORD(a.x); CRC_WORD(a.y); CRC_WORD(a.z); CRC_WORD(a.w);
    CRC_WORD(b.x); CRC_WORD(b.y); CRC_WORD(b.z); CRC_WORD(b.w);
    p4 += 8;
  }
  for (; len >= 4; len -= 4) { CRC_WORD(crc_ld32<RO>(p4)); ++p4; }
#undef CRC_WORD
  p = reinterpret_cast<const uint8_t*>(p4);
  while (len--) c = tab[(c ^ crc_ld8<RO>(p++)) & 0xFF] ^ (c >> 8);
  return c;
}

// zlib.crc32(data[0..len)) computed by the whole CTA; the result is valid in every thread.
// scratch: 33 uint32 of shared memory.  All threads of the CTA must call it.
template <bool RO = true>
__device__ __forceinline__ uint32_t block_crc32(const uint32_t* tab, uint32_t* scratch, const uint8_t* data, uint64_t len) {
  const int nt = blockDim.x, t = threadIdx.x;
  uint64_t seg = (len + nt - 1) / nt;
  seg = (seg + 15) & ~(uint64_t)15;                        // 16-byte segments keep the word loop aligned
  const uint64_t lo = min(len, (uint64_t)t * seg), hi = min(len, lo + seg);
  uint32_t c = 0;
  if (hi > lo) {
    c = crc_raw_segment<RO>(tab, data + lo, hi - lo);
    if (len - hi) c = gf2_mulmod(c, gf2_xpow8n(len - hi));
  }
  if (t == 0) c ^= gf2_mulmod(0xFFFFFFFFu, gf2_xpow8n(len));   // init state 0xFFFFFFFF carried through len bytes
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) c ^= __shfl_xor_sync(0xffffffffu, c, off);
  __syncthreads();
  if ((t & 31) == 0) scratch[t >> 5] = c;
  __syncthreads();
  if (t == 0) {
    uint32_t r = 0;
    for (int w = 0; w < (nt + 31) / 32; ++w) r ^= scratch[w];
    scratch[32] = r ^ 0xFFFFFFFFu;
  }
  __syncthreads();
  return scratch[32];
}
