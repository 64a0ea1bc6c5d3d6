// DPSK (v2) interior symbols on the tensor pipe: the composite zero-phase kernel of fbdsp/design.py evaluated as ONE
// banded-Toeplitz contraction per group of 8 symbols (see fbdsp/mma_tables.py for the algebra; modem.py:73-105, 194-241
// for what it reproduces).
//
//   D[group r][(symbol s, re/im)] = sum_kk  x[w0'(r) + kk] * B[kk][(s, re/im)]        mma.sync m16n8k16, fp16 x fp16 -> fp32
//
// Rows are consecutive groups (16 per m-tile), so the A operand is a Hankel matrix of the sample stream: row r starts
// 8*sps samples after row r-1.  The samples are staged ONCE, contiguously (fp16 hi / lo pieces, 22 significant bits), and
// ldmatrix reads the overlapping rows straight out of that buffer -- there is no im2col copy.  B (the taps, hi / lo) lives
// in registers for the whole kernel.  A second 8-column MMA on the same rows gives the group features of the slow poles;
// the slow-pole memory then costs one scan element per GROUP (80 samples) instead of 4 FMA per sample.
//
// Persistent CTAs, warp-specialised: 4 loader warps (aligned 16-byte global loads -> scale -> hi/lo split -> shared
// memory, two stages, mbarrier full/empty) and 8 MMA warps (ldmatrix + HMMA, group scan, slicer).  A CTA walks a contiguous
// range of tiles so the forward slow-pole state is carried from tile to tile; the backward state comes from one look-ahead
// m-tile of features.  Tiles whose samples do not fit the fp16 split (|x| >= 4, or quieter than 2^-18) are appended to a
// redo list and evaluated by the fp32 kernel (psk_v2.cu) afterwards.
#include "common.cuh"
#include "psk_shared.cuh"

#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

namespace {

// ---- schedule of the headline class: sps 10, 16 taps per polyphase row (dl 7, dh 8), two slow poles --------------------
struct Sched10 {
  static constexpr int SPS = 10, G = 8, SEG = SPS * G, PITCH = SEG + 8;   // halves per staged segment: +8 keeps ldmatrix conflict-free
  static constexpr int KS = 15;                       // k-steps of 16 samples: 7 + 70 + 160 = 237 <= 240
  static constexpr int HH0A = 0, HH0B = 12, HH1A = 2, HH1B = 14;          // non-zero 16x8 blocks of B per n-tile [first, last]
  static constexpr int LO0A = 2, LO0B = 10, LO1A = 4, LO1B = 12;          // blocks that get the lo products
  static constexpr int FTA = 5, FTB = 10;                                 // k-steps that cover the group's own samples
  static constexpr int LOA = 2, LOB = 12;                                 // k-steps whose lo samples are needed at all
  static constexpr int N_HH0 = HH0B - HH0A + 1, N_HH1 = HH1B - HH1A + 1, N_LO0 = LO0B - LO0A + 1, N_LO1 = LO1B - LO1A + 1,
                       N_FT = FTB - FTA + 1;
  static constexpr int O_HH1 = N_HH0, O_LO0 = O_HH1 + N_HH1, O_LO1 = O_LO0 + N_LO0, O_FTH = O_LO1 + N_LO1, O_FTL = O_FTH + N_FT,
                       NFRAG = O_FTL + N_FT;          // 56
};

constexpr int MMA_WARPS = 8, LOAD_WARPS = 4, MMA_THREADS = 32 * MMA_WARPS, LOAD_THREADS = 32 * LOAD_WARPS;
constexpr int ROWS = 256;                  // group rows per tile: 240 main (15 m-tiles) + 16 look-ahead (features only)
constexpr int MAIN_ROWS = 240;
constexpr int ADV_ROWS = 236;              // rows a tile advances by: 1888 symbols = 59 words of 32 symbols
constexpr int TILE_SYMS = ADV_ROWS * 8;    // 1888 differential symbols decided per tile
constexpr int SEGS = ROWS + 2;             // staged segments: the last row's window reaches 2 segments further
constexpr float SX = 16384.f;              // 2^14

struct MmaArgs {
  const void* samples;
  const PskTile* tiles;
  uint32_t n_tiles;
  const uint2* frags;           // [8 shifts][NFRAG][32 lanes]
  const float4* maps;           // [2 poles][8 symbols][2 dirs]  real 2x2 maps in accumulator units
  const float4* slow_pw4;       // {p_a^k, p_b^k}: direct-sum weights of the forward state at a range start
  int wpad, wlen, n0, pad_bp, bps;
  float2 lam[2];                // p^(8 sps) per pole
  float2 lam_pow[2][6];         // lam^(2^st), st = 0..4, and lam^32
  double slow_p[4];
  float state_scale;            // Sx * Sf: accumulator units of the slow-pole states
  float2 rho;
  uint32_t* bits;
  uint32_t* redo_count;         // redo_list[atomicAdd(redo_count)] = tile
  uint32_t* redo_list;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("{ .reg .b64 t; mbarrier.arrive.shared::cta.b64 t, [%0]; }" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
      ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void hmma(float (&d)[4], const uint32_t (&a)[4], const uint2 b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
__device__ __forceinline__ void bar_mma() { asm volatile("bar.sync 1, %0;" ::"n"(MMA_THREADS) : "memory"); }

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x)); }
__device__ __forceinline__ float2 cfmaf(float2 a, float2 b, float2 c) {   // a*b + c
  return make_float2(fmaf(a.x, b.x, fmaf(-a.y, b.y, c.x)), fmaf(a.x, b.y, fmaf(a.y, b.x, c.y)));
}
__device__ __forceinline__ float2 mapf(float4 m, float2 f, float2 acc) {  // acc + [m.x m.y; m.z m.w] (f.x, f.y)
  return make_float2(fmaf(m.x, f.x, fmaf(m.y, f.y, acc.x)), fmaf(m.z, f.x, fmaf(m.w, f.y, acc.y)));
}

// 8 consecutive samples starting at element e (a multiple of 8: 16-byte aligned for every storage type), as floats * 2^14
template <typename T> __device__ __forceinline__ void load8(const void* base, uint64_t e, float (&v)[8]);
template <> __device__ __forceinline__ void load8<float>(const void* base, uint64_t e, float (&v)[8]) {
  const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + e);
  const float4 a = __ldg(p), b = __ldg(p + 1);
  v[0] = a.x * SX; v[1] = a.y * SX; v[2] = a.z * SX; v[3] = a.w * SX; v[4] = b.x * SX; v[5] = b.y * SX; v[6] = b.z * SX; v[7] = b.w * SX;
}
template <> __device__ __forceinline__ void load8<int16_t>(const void* base, uint64_t e, float (&v)[8]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const int16_t*>(base) + e));
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {                  // value / 32768 * 2^14 = value / 2: exact
    v[2 * i] = (float)(int16_t)(w[i] & 0xFFFFu) * 0.5f;
    v[2 * i + 1] = (float)(int16_t)(w[i] >> 16) * 0.5f;
  }
}
template <> __device__ __forceinline__ void load8<double>(const void* base, uint64_t e, float (&v)[8]) {
  const double2* p = reinterpret_cast<const double2*>(reinterpret_cast<const double*>(base) + e);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double2 a = __ldg(p + i);
    v[2 * i] = (float)a.x * SX; v[2 * i + 1] = (float)a.y * SX;   // the fp32 kernel rounds to float first, too
  }
}
template <typename T> __device__ __forceinline__ float load1s(const void* base, uint64_t e) { return load_sample<T>(base, e) * SX; }

// hi = v rounded to 11 significant bits (exact in fp16 for |v| < 65520), lo = fp16(v - hi)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const float ha = __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xFFFFE000u);
  const float hb = __uint_as_float((__float_as_uint(b) + 0x1000u) & 0xFFFFE000u);
  const __half2 h = __floats2half2_rn(ha, hb), l = __floats2half2_rn(a - ha, b - hb);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <typename S> struct Smem {
  static constexpr int ARR = SEGS * S::PITCH * 2;          // bytes of one fp16 array of a stage
  static constexpr int STAGE = 2 * ARR;                    // hi then lo
  static constexpr int O_Z = 2 * STAGE;                    // float2 [ROWS][4]: group features, then states (in place)
  static constexpr int O_U = O_Z + ROWS * 32;              // float2 [MAIN_ROWS * 8 + 8]: symbols of the tile
  static constexpr int O_MAX = O_U + (MAIN_ROWS * 8 + 8) * 8;   // float [2][LOAD_THREADS]
  static constexpr int O_MISC = O_MAX + 2 * LOAD_THREADS * 4;   // warp totals, carry, flags, mbarriers
  static constexpr int TOTAL = O_MISC + 512;
};

struct Misc {
  float2 tot[MMA_WARPS][4];      // warp totals of the group scan
  float2 carry[2][2];            // [tile parity][pole]: forward states entering row 0 (written by the previous tile at its row ADV_ROWS)
  int ok;                        // the tile's samples fit the fp16 split
  int pad;
  uint64_t full[2], empty[2];
};

template <typename TIn, typename S>
__global__ void __launch_bounds__(MMA_THREADS + LOAD_THREADS, 1) psk_mma_kernel(const __grid_constant__ MmaArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  using L = Smem<S>;
  Misc* misc = reinterpret_cast<Misc*>(smem + L::O_MISC);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&misc->full[0], LOAD_THREADS); mbar_init(&misc->full[1], LOAD_THREADS);
    mbar_init(&misc->empty[0], MMA_WARPS); mbar_init(&misc->empty[1], MMA_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // contiguous tile range of this CTA
  const uint32_t t_lo = (uint32_t)(((uint64_t)a.n_tiles * blockIdx.x) / gridDim.x), t_hi = (uint32_t)(((uint64_t)a.n_tiles * (blockIdx.x + 1)) / gridDim.x);

  if (warp >= MMA_WARPS) {
    // ================================================== loader warps ==================================================
    const int lt = tid - MMA_THREADS;
    for (uint32_t t = t_lo, it = 0; t < t_hi; ++t, ++it) {
      const int st = it & 1;
      mbar_wait(&misc->empty[st], ((it >> 1) & 1) ^ 1);
      const PskTile pl = a.tiles[t];
      const int64_t N = (int64_t)pl.n;
      // window origin of row 0, moved down to a multiple of 8 elements of the sample buffer
      const int64_t w0 = (int64_t)a.n0 + (int64_t)pl.d0 * S::SPS - S::SPS * S::G;     // H+ = 8 sps
      const int sh = (int)((pl.off + (uint64_t)w0) & 7);
      const int64_t w0a = w0 - sh;
      unsigned char* sb = smem + st * L::STAGE;
      float mx = 0.f;
      for (int u = lt; u < SEGS * (S::SEG / 8); u += LOAD_THREADS) {
        const int seg = u / (S::SEG / 8), within = (u - seg * (S::SEG / 8)) * 8;
        const int64_t n = w0a + (int64_t)u * 8;
        float v[8];
        if (n >= 0 && n + 8 <= N) {
          load8<TIn>(a.samples, pl.off + (uint64_t)n, v);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = (n + i >= 0 && n + i < N) ? load1s<TIn>(a.samples, pl.off + (uint64_t)(n + i)) : 0.f;
        }
        uint4 hi, lo;
        split2(v[0], v[1], hi.x, lo.x); split2(v[2], v[3], hi.y, lo.y); split2(v[4], v[5], hi.z, lo.z); split2(v[6], v[7], hi.w, lo.w);
#pragma unroll
        for (int i = 0; i < 8; ++i) mx = fmaxf(mx, fabsf(v[i]));
        const int o = (seg * S::PITCH + within) * 2;
        *reinterpret_cast<uint4*>(sb + o) = hi;
        *reinterpret_cast<uint4*>(sb + L::ARR + o) = lo;
      }
      reinterpret_cast<float*>(smem + L::O_MAX)[st * LOAD_THREADS + lt] = mx;
      mbar_arrive(&misc->full[st]);
    }
    return;
  }

  // ==================================================== MMA warps =====================================================
  uint2 bf[S::NFRAG];                       // B fragments (taps hi / lo, feature weights hi / lo) of the current shift
  int cur_sh = -1;
  float2* Zs = reinterpret_cast<float2*>(smem + L::O_Z);
  float2* Us = reinterpret_cast<float2*>(smem + L::O_U);
  const int g = lane >> 2, q = lane & 3;
  uint64_t prev_off = ~0ull;
  int prev_d0 = 0;

  for (uint32_t t = t_lo, it = 0; t < t_hi; ++t, ++it) {
    const int st = it & 1;
    const PskTile pl = a.tiles[t];
    const int64_t N = (int64_t)pl.n;
    const int64_t w0 = (int64_t)a.n0 + (int64_t)pl.d0 * S::SPS - S::SPS * S::G;
    const int sh = (int)((pl.off + (uint64_t)w0) & 7);
    if (sh != cur_sh) {
      const uint2* src = a.frags + ((size_t)sh * S::NFRAG) * 32 + lane;
#pragma unroll
      for (int i = 0; i < S::NFRAG; ++i) bf[i] = __ldg(src + i * 32);
      cur_sh = sh;
    }
    const bool chained = (pl.off == prev_off) && (pl.d0 == prev_d0 + TILE_SYMS);   // forward state carried from the previous tile
    prev_off = pl.off; prev_d0 = pl.d0;

    // ---- forward slow-pole state at the tile's first group when it cannot be carried: direct sum over the previous wlen
    // samples (exact start-up state of scipy's filtfilt when the record start is within reach; psk_v2.cu has the algebra)
    if (!chained && warp == 0) {
      const int64_t n_d0 = (int64_t)a.n0 + (int64_t)pl.d0 * S::SPS;
      const bool near_left = (n_d0 - a.wlen) <= (int64_t)a.n0;
      const int cnt = (int)min((int64_t)a.wlen, n_d0 - (near_left ? (int64_t)a.n0 : (int64_t)0));
      float2 s0 = make_float2(0.f, 0.f), s1 = s0;
      for (int k = 1 + lane; k <= cnt; k += 32) {
        const float x = load_sample<TIn>(a.samples, pl.off + (uint64_t)(n_d0 - k));
        const float4 w = __ldg(&a.slow_pw4[k]);
        s0.x = fmaf(x, w.x, s0.x); s0.y = fmaf(x, w.y, s0.y); s1.x = fmaf(x, w.z, s1.x); s1.y = fmaf(x, w.w, s1.y);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        s0.x += __shfl_xor_sync(0xffffffffu, s0.x, off); s0.y += __shfl_xor_sync(0xffffffffu, s0.y, off);
        s1.x += __shfl_xor_sync(0xffffffffu, s1.x, off); s1.y += __shfl_xor_sync(0xffffffffu, s1.y, off);
      }
      if (near_left && lane < 2) {
        // F[0] = sum_{n < n0} p^(n0-n) xL[n], xL = scipy's odd extension (pad_bp samples), then the constant xL[-pad_bp] for ever
        const double pr = a.slow_p[2 * lane], pi = a.slow_p[2 * lane + 1];
        const double x0 = load_sample_d<TIn>(a.samples, pl.off);
        const double cr = 2.0 * x0 - load_sample_d<TIn>(a.samples, pl.off + a.pad_bp);
        const double den = (1.0 - pr) * (1.0 - pr) + pi * pi;
        double sr = cr * (1.0 - pr) / den, si = cr * pi / den;
        for (int n = -a.pad_bp; n < a.n0; ++n) {
          const double xv = (n < 0) ? 2.0 * x0 - load_sample_d<TIn>(a.samples, pl.off + (uint64_t)(-n)) : load_sample_d<TIn>(a.samples, pl.off + (uint64_t)n);
          const double tr = pr * sr - pi * si + xv, ti = pr * si + pi * sr;
          sr = tr; si = ti;
        }
        const double fr = pr * sr - pi * si, fi = pr * si + pi * sr;        // Fst[0] = p * s
        const float4 pw4 = __ldg(&a.slow_pw4[(int)(n_d0 - a.n0)]);
        const float2 pw = lane == 0 ? make_float2(pw4.x, pw4.y) : make_float2(pw4.z, pw4.w);
        const float2 add = cmulf(pw, make_float2((float)fr, (float)fi));
        if (lane == 0) { s0.x += add.x; s0.y += add.y; } else { s1.x += add.x; s1.y += add.y; }
      }
      const float2 mine = lane == 0 ? s0 : s1;
      if (lane < 2) misc->carry[it & 1][lane] = make_float2(mine.x * a.state_scale, mine.y * a.state_scale);
    }

    mbar_wait(&misc->full[st], (it >> 1) & 1);
    if (warp == 0) {                                  // range check of the tile's samples (scaled by 2^14)
      const float* mxs = reinterpret_cast<const float*>(smem + L::O_MAX) + st * LOAD_THREADS;
      float m = fmaxf(fmaxf(mxs[lane], mxs[lane + 32]), fmaxf(mxs[lane + 64], mxs[lane + 96]));
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
      if (lane == 0) misc->ok = (m < 65000.f && m >= 0.0625f) ? 1 : 0;
    }

    // ---- MMA phase: two m-tiles per warp (the very last m-tile of the CTA tile is the look-ahead: features only) ----------
    const uint32_t sbase = smem_u32(smem + st * L::STAGE);
    float u[2][2][4];                                 // [m-tile][n-tile][d0..d3]
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const int mt = 2 * warp + mi;
      const bool main_tile = mt < MAIN_ROWS / 16;
      const uint32_t abase = sbase + (uint32_t)(((16 * mt + (lane & 15)) * S::PITCH + (lane >> 4) * 8) * 2);
      float ah[2][4], al[2][4], fh[4], fl[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { ah[0][i] = ah[1][i] = al[0][i] = al[1][i] = fh[i] = fl[i] = 0.f; }
      if (main_tile) {
#pragma unroll
        for (int k = 0; k < S::KS; ++k) {
          const uint32_t ko = (uint32_t)(((k / 5) * S::PITCH + (k % 5) * 16) * 2);
          uint32_t xa[4], xl[4];
          ldsm4(xa, abase + ko);
          const bool need_lo = k >= S::LOA && k <= S::LOB;
          if (need_lo) ldsm4(xl, abase + L::ARR + ko);
          if (k >= S::HH0A && k <= S::HH0B) hmma(ah[0], xa, bf[k - S::HH0A]);
          if (k >= S::HH1A && k <= S::HH1B) hmma(ah[1], xa, bf[S::O_HH1 + k - S::HH1A]);
          if (k >= S::LO0A && k <= S::LO0B) { hmma(al[0], xa, bf[S::O_LO0 + k - S::LO0A]); hmma(al[0], xl, bf[k - S::HH0A]); }
          if (k >= S::LO1A && k <= S::LO1B) { hmma(al[1], xa, bf[S::O_LO1 + k - S::LO1A]); hmma(al[1], xl, bf[S::O_HH1 + k - S::HH1A]); }
          if (k >= S::FTA && k <= S::FTB) {
            hmma(fh, xa, bf[S::O_FTH + k - S::FTA]);
            hmma(fl, xa, bf[S::O_FTL + k - S::FTA]);
            hmma(fl, xl, bf[S::O_FTH + k - S::FTA]);
          }
        }
      } else {
#pragma unroll
        for (int k = S::FTA; k <= S::FTB; ++k) {
          const uint32_t ko = (uint32_t)(((k / 5) * S::PITCH + (k % 5) * 16) * 2);
          uint32_t xa[4], xl[4];
          ldsm4(xa, abase + ko);
          ldsm4(xl, abase + L::ARR + ko);
          hmma(fh, xa, bf[S::O_FTH + k - S::FTA]);
          hmma(fl, xa, bf[S::O_FTL + k - S::FTA]);
          hmma(fl, xl, bf[S::O_FTH + k - S::FTA]);
        }
      }
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) u[mi][n][i] = ah[n][i] + al[n][i];
      // group features: thread (g, q) holds feature q (Zf0, Zb0, Zf1, Zb1) of rows g and g + 8
      Zs[(16 * mt + g) * 4 + q] = make_float2(fh[0] + fl[0], fh[1] + fl[1]);
      Zs[(16 * mt + g + 8) * 4 + q] = make_float2(fh[2] + fl[2], fh[3] + fl[3]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&misc->empty[st]);      // this warp no longer reads the sample stage
    bar_mma();

    // ---- group scan: thread r <-> row r.  F[r+1] = lam F[r] + Zf[r] (F[0] = carry);  Bk[r] = Zb[r+1] + lam Bk[r+1] (Bk[255] = 0) ----
    {
      const int r = tid;
      const float4 z01 = *reinterpret_cast<const float4*>(&Zs[r * 4]), z23 = *reinterpret_cast<const float4*>(&Zs[r * 4 + 2]);
      // inclusive scans: forward over Zf (value at r = state entering row r+1), backward over Zb (value at r = Zb[r] + lam * ...)
      float2 v[4] = {make_float2(z01.x, z01.y), make_float2(z23.x, z23.y), make_float2(z01.z, z01.w), make_float2(z23.z, z23.w)};   // Zf0, Zf1, Zb0, Zb1
#pragma unroll
      for (int s = 0; s < 5; ++s) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 mm = a.lam_pow[k & 1][s];
          if (k < 2) {
            const float ox = __shfl_up_sync(0xffffffffu, v[k].x, 1 << s), oy = __shfl_up_sync(0xffffffffu, v[k].y, 1 << s);
            if (lane >= (1 << s)) v[k] = cfmaf(mm, make_float2(ox, oy), v[k]);
          } else {
            const float ox = __shfl_down_sync(0xffffffffu, v[k].x, 1 << s), oy = __shfl_down_sync(0xffffffffu, v[k].y, 1 << s);
            if (lane + (1 << s) < 32) v[k] = cfmaf(mm, make_float2(ox, oy), v[k]);
          }
        }
      }
      if (lane == 31) { misc->tot[warp][0] = v[0]; misc->tot[warp][1] = v[1]; }
      if (lane == 0) { misc->tot[warp][2] = v[2]; misc->tot[warp][3] = v[3]; }
      bar_mma();
      // state entering this warp's 32 rows from the left (k < 2) / right (k >= 2)
      float2 c = make_float2(0.f, 0.f);
      if (lane < 4) {
        const int k = lane;
        const float2 M = a.lam_pow[k & 1][5];         // lam^32
        if (k < 2) {
          c = misc->carry[it & 1][k];
          for (int w = 0; w < warp; ++w) c = cfmaf(M, c, misc->tot[w][k]);
        } else {
          for (int w = MMA_WARPS - 1; w > warp; --w) c = cfmaf(M, c, misc->tot[w][k]);
        }
      }
      float2 car[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) car[k] = make_float2(__shfl_sync(0xffffffffu, c.x, k), __shfl_sync(0xffffffffu, c.y, k));
      // F[r] (state entering row r) = lam^lane * car + exclusive prefix;  Bk[r] (sources after row r) = lam^(31-lane) * car + exclusive suffix
      float2 outv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float ex, ey;
        if (k < 2) { ex = __shfl_up_sync(0xffffffffu, v[k].x, 1); ey = __shfl_up_sync(0xffffffffu, v[k].y, 1); }
        else { ex = __shfl_down_sync(0xffffffffu, v[k].x, 1); ey = __shfl_down_sync(0xffffffffu, v[k].y, 1); }
        const bool has = (k < 2) ? lane > 0 : lane < 31;
        const float2 excl = has ? make_float2(ex, ey) : make_float2(0.f, 0.f);
        // lam^n for n = lane (forward) or 31 - lane (backward), from the binary powers
        const int n = (k < 2) ? lane : 31 - lane;
        float2 pw = make_float2(1.f, 0.f);
#pragma unroll
        for (int s = 0; s < 5; ++s) if (n & (1 << s)) pw = cmulf(pw, a.lam_pow[k & 1][s]);
        outv[k] = cfmaf(pw, car[k], excl);
      }
      // the forward state a chained next tile starts from: the state entering row ADV_ROWS
      if (r == ADV_ROWS) { misc->carry[(it + 1) & 1][0] = outv[0]; misc->carry[(it + 1) & 1][1] = outv[1]; }
      // in place: row r now holds F0, B0, F1, B1  (the order the epilogue maps expect)
      *reinterpret_cast<float4*>(&Zs[r * 4]) = make_float4(outv[0].x, outv[0].y, outv[2].x, outv[2].y);
      *reinterpret_cast<float4*>(&Zs[r * 4 + 2]) = make_float4(outv[1].x, outv[1].y, outv[3].x, outv[3].y);
    }
    bar_mma();

    // ---- symbols: in-group part (accumulators) + out-of-group sources through the states; to shared memory by symbol index ----
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const int mt = 2 * warp + mi;
      if (mt < MAIN_ROWS / 16) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int row = 16 * mt + g + 8 * h;
          const float4 s01 = *reinterpret_cast<const float4*>(&Zs[row * 4]), s23 = *reinterpret_cast<const float4*>(&Zs[row * 4 + 2]);
          const float2 F0 = make_float2(s01.x, s01.y), B0 = make_float2(s01.z, s01.w), F1 = make_float2(s23.x, s23.y), B1 = make_float2(s23.z, s23.w);
#pragma unroll
          for (int n = 0; n < 2; ++n) {
            const int s = 4 * n + q;
            float2 y = make_float2(u[mi][n][2 * h], u[mi][n][2 * h + 1]);
            y = mapf(__ldg(&a.maps[(0 * 8 + s) * 2 + 0]), F0, y);
            y = mapf(__ldg(&a.maps[(0 * 8 + s) * 2 + 1]), B0, y);
            y = mapf(__ldg(&a.maps[(1 * 8 + s) * 2 + 0]), F1, y);
            y = mapf(__ldg(&a.maps[(1 * 8 + s) * 2 + 1]), B1, y);
            Us[row * 8 + s] = y;
          }
        }
      }
    }
    bar_mma();

    // ---- differential decisions: thread e <-> symbols 8e .. 8e+7 of the tile (236 threads), words of 32 bits as in psk_v2.cu ----
    {
      const int e0 = tid * 8;
      const int nd = pl.d1 - pl.d0;                   // multiple of 32, <= TILE_SYMS
      uint32_t part = 0;
      if (e0 < TILE_SYMS) {
        float2 y[9];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 p = *reinterpret_cast<const float4*>(&Us[e0 + 2 * i]);
          y[2 * i] = make_float2(p.x, p.y); y[2 * i + 1] = make_float2(p.z, p.w);
        }
        y[8] = Us[e0 + 8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 prev = y[i], cur = y[i + 1];
          // the products reach 2^60 in accumulator units: scale one factor down (a positive factor does not move the angle)
          const float cx = cur.x * 5.9604645e-8f, cy = cur.y * 5.9604645e-8f;
          const float tr = fmaf(cx, prev.x, cy * prev.y), ti = fmaf(cy, prev.x, -cx * prev.y);
          if (a.bps == 2) {
            const float dr = fmaf(tr, a.rho.x, -ti * a.rho.y), di = fmaf(tr, a.rho.y, ti * a.rho.x);
            part = (part << 2) | psk_decide<float>(dr, di, 2);
          } else {
            const float dr = fmaf(tr, a.rho.x, -ti * a.rho.y);
            part = (part << 1) | (dr < 0.f ? 1u : 0u);
          }
        }
      }
      if (a.bps == 2) {
        const uint32_t other = __shfl_down_sync(0xffffffffu, part, 1);
        if ((lane & 1) == 0 && e0 < nd && misc->ok) a.bits[pl.word_off + (uint64_t)((pl.d0 + e0) >> 4)] = __byte_perm((part << 16) | other, 0, 0x0123);
      } else {
        uint32_t wv = part << (24 - 8 * (lane & 3));
        wv |= __shfl_xor_sync(0xffffffffu, wv, 1);
        wv |= __shfl_xor_sync(0xffffffffu, wv, 2);
        if ((lane & 3) == 0 && e0 < nd && misc->ok) a.bits[pl.word_off + (uint64_t)((pl.d0 + e0) >> 5)] = __byte_perm(wv, 0, 0x0123);
      }
      if (tid == 0 && !misc->ok) a.redo_list[atomicAdd(a.redo_count, 1u)] = t;
    }
    // the next tile's barriers order every reuse of Zs / Us / misc against the reads above
  }
}

}  // namespace

// =====================================================================================================================
// Host side: tables (built once per design and cached on the handle) and the launch
// =====================================================================================================================
struct MmaTables {
  fb_psk_design d;
  std::vector<float> taps;
  bool usable = false;
  std::vector<uint2> frags;        // [8][NFRAG][32]
  std::vector<float4> maps;        // [2][8][2]
  std::vector<float4> pw4;         // {p0^k, p1^k}
  float2 lam[2], lam_pow[2][6];
  float state_scale = 0.f;
  int wlen = 0;
  void* d_blob = nullptr;          // device copy: frags | maps | pw4
  size_t o_maps = 0, o_pw = 0;
};

static inline uint16_t h16(double v) { const __half h = __double2half(v); uint16_t r; memcpy(&r, &h, 2); return r; }
static inline double h16d(uint16_t b) { __half h; memcpy(&h, &b, 2); return (double)__half2float(h); }

struct cd { double r, i; };
static inline cd cmul_(cd a, cd b) { return {a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r}; }
static cd cpow_(cd b, int n) { cd r{1.0, 0.0}; cd x = b; while (n > 0) { if (n & 1) r = cmul_(r, x); x = cmul_(x, x); n >>= 1; } return r; }

// Builds the band matrices of fbdsp/mma_tables.py for the Sched10 class.  Returns false when the design does not fit the
// compile-time schedule (other sps / tap counts / pole counts, or non-zero blocks outside it): the fp32 kernel serves those.
static bool build_mma_tables(const fb_psk_design& d, const float* taps, MmaTables& T, double* dbg_bfir /* [8][240][16] or null */) {
  using S = Sched10;
  if (d.sps != S::SPS || d.nt != 16 || d.dl != 7 || d.dh != 8 || d.nslow != 2 || d.emulate_only) return false;
  const int sps = d.sps, gs = S::G * sps, Hp = S::G * sps, KP = 16 * S::KS;
  const int hpos = d.dh * sps, hneg = d.dl * sps + sps - 1;
  if (Hp != hpos || (ROWS - 1 - ADV_ROWS) * gs < d.wcols * sps) return false;      // look-ahead must cover the slow-pole memory
  cd p[2], rp[2], rpc[2], rm[2], rmc[2];
  for (int i = 0; i < 2; ++i) {
    p[i] = {d.slow_p[2 * i], d.slow_p[2 * i + 1]};
    rp[i] = {d.slow_rp[2 * i], d.slow_rp[2 * i + 1]}; rpc[i] = {d.slow_rpc[2 * i], d.slow_rpc[2 * i + 1]};
    rm[i] = {d.slow_rm[2 * i], d.slow_rm[2 * i + 1]}; rmc[i] = {d.slow_rmc[2 * i], d.slow_rmc[2 * i + 1]};
  }
  auto cfast = [&](int q) -> cd {
    const int j = ((-q) % sps + sps) % sps, t = d.dl + (q + j) / sps;
    if (t < 0 || t >= d.nt) return {0.0, 0.0};
    return {(double)taps[((size_t)j * d.nt + t) * 2], (double)taps[((size_t)j * d.nt + t) * 2 + 1]};
  };
  auto cslow = [&](int q) -> cd {
    cd v{0.0, 0.0};
    for (int i = 0; i < 2; ++i) {
      const cd pc{p[i].r, -p[i].i};
      if (q > 0) { const cd a1 = cmul_(rp[i], cpow_(p[i], q)), a2 = cmul_(rpc[i], cpow_(pc, q)); v.r += a1.r + a2.r; v.i += a1.i + a2.i; }
      else if (q < 0) { const cd a1 = cmul_(rm[i], cpow_(p[i], -q)), a2 = cmul_(rmc[i], cpow_(pc, -q)); v.r += a1.r + a2.r; v.i += a1.i + a2.i; }
    }
    return v;
  };
  std::vector<double> bfir((size_t)8 * KP * 16, 0.0), bft((size_t)8 * KP * 8, 0.0);
  double mfir = 0.0;
  for (int sh = 0; sh < 8; ++sh)
    for (int kk = 0; kk < KP; ++kk) {
      const int jj = kk - Hp - sh;
      for (int s = 0; s < S::G; ++s) {
        const int q = Hp + sps * s + sh - kk;
        cd v{0.0, 0.0};
        if (q >= -hneg && q <= hpos) v = cfast(q);
        if (jj >= 0 && jj < gs && q != 0) { const cd c = cslow(q); v.r += c.r; v.i += c.i; }
        bfir[((size_t)sh * KP + kk) * 16 + 2 * s] = v.r; bfir[((size_t)sh * KP + kk) * 16 + 2 * s + 1] = v.i;
        mfir = std::max(mfir, std::max(fabs(v.r), fabs(v.i)));
      }
      if (jj >= 0 && jj < gs)
        for (int i = 0; i < 2; ++i) {
          const cd zf = cpow_(p[i], gs - jj), zb = cpow_(p[i], jj);
          double* o = &bft[((size_t)sh * KP + kk) * 8 + 4 * i];
          o[0] = zf.r; o[1] = zf.i; o[2] = zb.r; o[3] = zb.i;
        }
    }
  if (!(mfir > 0.0)) return false;
  if (dbg_bfir) memcpy(dbg_bfir, bfir.data(), bfir.size() * sizeof(double));
  const double st = exp2(floor(log2(16384.0 / mfir))), sf = 8192.0;
  // blocks outside the schedule must be zero (hh) / below the lo threshold
  const double lo_tol = 2.0e-5 * mfir;
  auto in = [](int k, int a, int b) { return k >= a && k <= b; };
  for (int sh = 0; sh < 8; ++sh)
    for (int k = 0; k < S::KS; ++k) {
      for (int n = 0; n < 2; ++n) {
        double bm = 0.0;
        for (int r = 0; r < 16; ++r) for (int c = 0; c < 8; ++c) bm = std::max(bm, fabs(bfir[((size_t)sh * KP + 16 * k + r) * 16 + 8 * n + c]));
        const bool hh = n == 0 ? in(k, S::HH0A, S::HH0B) : in(k, S::HH1A, S::HH1B), lo = n == 0 ? in(k, S::LO0A, S::LO0B) : in(k, S::LO1A, S::LO1B);
        if ((bm != 0.0 && !hh) || (bm >= lo_tol && !lo)) return false;
      }
      double fm = 0.0;
      for (int r = 0; r < 16; ++r) for (int c = 0; c < 8; ++c) fm = std::max(fm, fabs(bft[((size_t)sh * KP + 16 * k + r) * 8 + c]));
      if (fm != 0.0 && !in(k, S::FTA, S::FTB)) return false;
    }
  // fragments: register j of lane l holds B[16k + 2 (l % 4) + 8 j + {0, 1}][8 n + l / 4], the lower k in the low half
  T.frags.assign((size_t)8 * S::NFRAG * 32, make_uint2(0, 0));
  auto put = [&](int sh, int frag, const double* mat, int ncols, int k, int n, double scale, bool lo_piece) {
    for (int l = 0; l < 32; ++l) {
      uint32_t reg[2];
      for (int j = 0; j < 2; ++j) {
        uint16_t hw[2];
        for (int e = 0; e < 2; ++e) {
          const double v = mat[((size_t)sh * KP + 16 * k + 2 * (l % 4) + 8 * j + e) * ncols + 8 * n + l / 4] * scale;
          const uint16_t hi = h16(v);
          hw[e] = lo_piece ? h16(v - h16d(hi)) : hi;
        }
        reg[j] = (uint32_t)hw[0] | ((uint32_t)hw[1] << 16);
      }
      T.frags[((size_t)sh * S::NFRAG + frag) * 32 + l] = make_uint2(reg[0], reg[1]);
    }
  };
  for (int sh = 0; sh < 8; ++sh) {
    for (int k = S::HH0A; k <= S::HH0B; ++k) put(sh, k - S::HH0A, bfir.data(), 16, k, 0, st, false);
    for (int k = S::HH1A; k <= S::HH1B; ++k) put(sh, S::O_HH1 + k - S::HH1A, bfir.data(), 16, k, 1, st, false);
    for (int k = S::LO0A; k <= S::LO0B; ++k) put(sh, S::O_LO0 + k - S::LO0A, bfir.data(), 16, k, 0, st, true);
    for (int k = S::LO1A; k <= S::LO1B; ++k) put(sh, S::O_LO1 + k - S::LO1A, bfir.data(), 16, k, 1, st, true);
    for (int k = S::FTA; k <= S::FTB; ++k) { put(sh, S::O_FTH + k - S::FTA, bft.data(), 8, k, 0, sf, false); put(sh, S::O_FTL + k - S::FTA, bft.data(), 8, k, 0, sf, true); }
  }
  // maps: a F + b conj(F) as a real 2x2 map of (Re F, Im F), accumulator units: (Sx St y) from (Sx Sf state)
  T.maps.assign(2 * 8 * 2, make_float4(0, 0, 0, 0));
  for (int i = 0; i < 2; ++i) {
    const cd pc{p[i].r, -p[i].i};
    for (int s = 0; s < S::G; ++s) {
      const cd af = cmul_(rp[i], cpow_(p[i], sps * s)), bfw = cmul_(rpc[i], cpow_(pc, sps * s));
      const cd ab = cmul_(rm[i], cpow_(p[i], gs - sps * s)), bbw = cmul_(rmc[i], cpow_(pc, gs - sps * s));
      const double k = st / sf;
      T.maps[(i * 8 + s) * 2 + 0] = make_float4((float)((af.r + bfw.r) * k), (float)((bfw.i - af.i) * k), (float)((af.i + bfw.i) * k), (float)((af.r - bfw.r) * k));
      T.maps[(i * 8 + s) * 2 + 1] = make_float4((float)((ab.r + bbw.r) * k), (float)((bbw.i - ab.i) * k), (float)((ab.i + bbw.i) * k), (float)((ab.r - bbw.r) * k));
    }
    const cd lam = cpow_(p[i], gs);
    T.lam[i] = make_float2((float)lam.r, (float)lam.i);
    cd x = lam;
    for (int s = 0; s < 6; ++s) { T.lam_pow[i][s] = make_float2((float)x.r, (float)x.i); x = cmul_(x, x); }
  }
  T.wlen = d.wcols * sps;
  const int wpad = T.wlen + 2 + std::max(T.wlen, 4096);
  T.pw4.assign(wpad, make_float4(0, 0, 0, 0));
  {
    cd a{1.0, 0.0}, b{1.0, 0.0};
    for (int k = 0; k < wpad; ++k) {
      T.pw4[k] = make_float4((float)a.r, (float)a.i, (float)b.r, (float)b.i);
      a = cmul_(a, p[0]); b = cmul_(b, p[1]);
    }
  }
  T.state_scale = (float)(16384.0 * sf);
  T.d = d;
  T.taps.assign(taps, taps + (size_t)d.sps * d.nt * 2);
  T.usable = true;
  return true;
}

// Debug / test hook (host only, no GPU needed): the float64 FIR band [8 shifts][240][16] the fragments are cut from, or -1.
extern "C" int fb_debug_mma_band(const fb_psk_design* d, const float* taps, double* bfir_out) {
  MmaTables T;
  return build_mma_tables(*d, taps, T, bfir_out) ? Sched10::KS : -1;
}

void fb_psk_mma_release(fb_handle* h) {
  for (void* v : h->mma_cache) {
    MmaTables* T = (MmaTables*)v;
    if (T->d_blob) cudaFree(T->d_blob);
    delete T;
  }
  h->mma_cache.clear();
}

static MmaTables* get_tables(fb_handle* h, const fb_psk_design& d, const float* taps, int* rc) {
  *rc = FB_OK;
  for (void* v : h->mma_cache) {
    MmaTables* T = (MmaTables*)v;
    if (memcmp(&T->d, &d, sizeof(d)) == 0 && memcmp(T->taps.data(), taps, T->taps.size() * 4) == 0) return T;
  }
  MmaTables* T = new MmaTables();
  T->d = d;
  T->taps.assign(taps, taps + (size_t)std::max(0, d.sps) * std::max(0, d.nt) * 2);
  if (build_mma_tables(d, taps, *T, nullptr)) {
    const size_t b_fr = T->frags.size() * sizeof(uint2), b_mp = T->maps.size() * sizeof(float4), b_pw = T->pw4.size() * sizeof(float4);
    T->o_maps = (b_fr + 255) / 256 * 256; T->o_pw = (T->o_maps + b_mp + 255) / 256 * 256;
    if (cudaMalloc(&T->d_blob, T->o_pw + b_pw) != cudaSuccess) { cudaGetLastError(); delete T; *rc = FB_ENOMEM; return nullptr; }
    cudaMemcpyAsync((char*)T->d_blob, T->frags.data(), b_fr, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync((char*)T->d_blob + T->o_maps, T->maps.data(), b_mp, cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync((char*)T->d_blob + T->o_pw, T->pw4.data(), b_pw, cudaMemcpyHostToDevice, h->stream);
    cudaStreamSynchronize(h->stream);      // the host vectors stay alive, but keep the first call simple
  }
  if (h->mma_cache.size() >= 32) fb_psk_mma_release(h);
  h->mma_cache.push_back(T);
  return T;
}

bool fb_psk_mma_usable(fb_handle* h, const fb_psk_design& d, const float* taps) {
  if (getenv("FB_PSK_NO_MMA")) return false;
  int rc;
  MmaTables* T = get_tables(h, d, taps, &rc);
  return T && T->usable;
}

int fb_psk_mma_tile_syms() { return TILE_SYMS; }

// Launches the tensor-core kernel over `n_tiles` tile descriptors (tile size TILE_SYMS).  redo: device buffer of
// 1 + n_tiles uint32 (count, then the tiles the fp32 kernel must evaluate); the count is zeroed here.
int fb_psk_mma_launch(fb_handle* h, const fb_psk_design& d, const float* taps, const void* d_samples, int dtype, const PskTile* d_tiles,
                      uint32_t n_tiles, uint32_t* d_bits, uint32_t* d_redo) {
  int rc;
  MmaTables* T = get_tables(h, d, taps, &rc);
  if (!T || !T->usable) return rc ? rc : FB_EUNSUPPORTED;
  using S = Sched10;
  MmaArgs a{};
  a.samples = d_samples; a.tiles = d_tiles; a.n_tiles = n_tiles;
  a.frags = (const uint2*)T->d_blob; a.maps = (const float4*)((char*)T->d_blob + T->o_maps); a.slow_pw4 = (const float4*)((char*)T->d_blob + T->o_pw);
  a.wpad = (int)T->pw4.size(); a.wlen = T->wlen; a.n0 = d.n0; a.pad_bp = d.pad_bp; a.bps = d.bits_per_sym;
  for (int i = 0; i < 2; ++i) {
    a.lam[i] = T->lam[i];
    for (int s = 0; s < 6; ++s) a.lam_pow[i][s] = T->lam_pow[i][s];
    a.slow_p[2 * i] = d.slow_p[2 * i]; a.slow_p[2 * i + 1] = d.slow_p[2 * i + 1];
  }
  a.state_scale = T->state_scale;
  a.rho = make_float2(d.rho[0], d.rho[1]);
  a.bits = d_bits; a.redo_count = d_redo; a.redo_list = d_redo + 1;
  FB_CUDA(h, cudaMemsetAsync(d_redo, 0, 4, h->stream));
  const int smem = Smem<S>::TOTAL;
  const int grid = (int)std::min<uint32_t>(n_tiles, (uint32_t)h->sm_count);
#define FB_MMA_LAUNCH(TIN)                                                                                              \
  do {                                                                                                                  \
    FB_CUDA(h, cudaFuncSetAttribute(psk_mma_kernel<TIN, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));        \
    psk_mma_kernel<TIN, S><<<grid, MMA_THREADS + LOAD_THREADS, smem, h->stream>>>(a);                                   \
  } while (0)
  if (dtype == FB_F32) FB_MMA_LAUNCH(float);
  else if (dtype == FB_F64) FB_MMA_LAUNCH(double);
  else FB_MMA_LAUNCH(int16_t);
#undef FB_MMA_LAUNCH
  h->launches++;
  FB_CUDA(h, cudaGetLastError());
  return FB_OK;
}
