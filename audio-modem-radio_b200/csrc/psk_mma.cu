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

#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>
#include <utility>

namespace {

// ---- schedule of the headline class: sps 10, 16 taps per polyphase row (dl 7, dh 8), two slow poles --------------------
struct Sched10 {
  static constexpr int SPS = 10, G = 8, SEG = SPS * G;
  // A staged segment (8 sps samples) is SEG / 8 slots of 32 bytes plus 16 bytes of padding (336 bytes: the 8 rows of an
  // ldmatrix phase then fall into 8 different 16-byte bank groups).  A slot arrives as 8 raw fp32 samples (TMA) and is
  // converted in place to [8 x fp16 hi | 8 x fp16 lo].
  static constexpr int SEGB = SEG * 4 + 16;
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
constexpr int CHUNK = 16;                  // tiles per scheduling unit (the forward slow-pole state is carried inside a chunk)
constexpr int TR_TILES = 48;               // tiles per CTA the phase trace covers

struct MmaArgs {
  const void* samples;
  const PskTile* tiles;
  uint32_t n_tiles;
  const uint2* frags;           // [8 shifts][NFRAG][32 lanes]
  const float4* maps;           // [2 poles][8 symbols][2 dirs]  real 2x2 maps in accumulator units
  const float4* slow_pw4;       // {p_a^k, p_b^k}: direct-sum weights of the forward state at a range start
  int wpad, wlen, n0, pad_bp, bps;
  float2 lam[2];                // p^(8 sps) per pole
  float2 lam_pow[2][7];         // lam^(2^st), st = 0..6
  double slow_p[4];
  float state_scale;            // Sx * Sf: accumulator units of the slow-pole states
  float2 rho;
  uint32_t* bits;
  uint32_t* redo_count;         // redo_list[atomicAdd(redo_count)] = tile
  uint32_t* redo_list;
  uint32_t* chunk_ctr;          // dynamic scheduling: CTAs claim chunks of CHUNK consecutive tiles
  long long* trace;             // FB_MMA_TRACE: [CTA][TR_TILES][16] clock64 stamps of the phase boundaries (experiments only)
  float4 maps_c[2 * 8 * 2];     // `maps` rearranged for packed FMAs {m.x, m.z, m.y, m.w}: constant-bank operands of the post warps
  uint32_t tm_rows;             // rows of the tensor map over the sample buffer (0: no TMA, every tile through the guarded loads)
  int dbg;                      // timing experiments only (FB_MMA_DBG): 1 = loaders skip their work, 2 = MMA warps skip theirs, 4 = no epilogue
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
// Hand-off barriers between the roles: hardware named barriers (bar.arrive by the producers, bar.sync by the consumers), so that
// a waiting warp costs no issue slots and no shared-memory traffic (polling mbarriers took 47 % of the executed instructions).
// Each barrier is used once per tile and stage, producers + consumers threads; ids 4 .. 9 (0: __syncthreads, 2: post, 3: loaders).
constexpr int HB_FULL = 4, HB_UFULL = 6, HB_UEMPTY = 8;
__device__ __forceinline__ void hb_arrive(int id) { __syncwarp(); asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(384) : "memory"); }
__device__ __forceinline__ void hb_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(384) : "memory"); }
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void hmma(float (&d)[4], const uint32_t (&a)[4], const uint2 b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
__device__ __forceinline__ void tr_mark(const MmaArgs& a, uint32_t it, int slot, bool who) {
  if (a.trace && who && it < TR_TILES) a.trace[((size_t)blockIdx.x * TR_TILES + it) * 16 + slot] = clock64();
}

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x)); }
__device__ __forceinline__ float2 cfmaf(float2 a, float2 b, float2 c) {   // a*b + c
  return make_float2(fmaf(a.x, b.x, fmaf(-a.y, b.y, c.x)), fmaf(a.x, b.y, fmaf(a.y, b.x, c.y)));
}

// One staging unit = 8 consecutive samples starting at element e (a multiple of 8: 16-byte aligned for every storage type).
// The raw 16-byte loads of several units are issued back to back (memory-level parallelism), the conversion to floats * 2^14
// happens when the unit is consumed.
template <typename T> struct Raw8;
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const void* base, uint64_t e) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + e);
    a = __ldg(p); b = __ldg(p + 1);
  }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    const float2 k = make_float2(SX, SX);
    const float2 p0 = __fmul2_rn(make_float2(a.x, a.y), k), p1 = __fmul2_rn(make_float2(a.z, a.w), k);
    const float2 p2 = __fmul2_rn(make_float2(b.x, b.y), k), p3 = __fmul2_rn(make_float2(b.z, b.w), k);
    v[0] = p0.x; v[1] = p0.y; v[2] = p1.x; v[3] = p1.y; v[4] = p2.x; v[5] = p2.y; v[6] = p3.x; v[7] = p3.y;
  }
};
template <> struct Raw8<int16_t> {
  uint4 a;
  __device__ __forceinline__ void load(const void* base, uint64_t e) { a = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const int16_t*>(base) + e)); }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {                // value / 32768 * 2^14 = value / 2: exact
      v[2 * i] = (float)(int16_t)(w[i] & 0xFFFFu) * 0.5f;
      v[2 * i + 1] = (float)(int16_t)(w[i] >> 16) * 0.5f;
    }
  }
};
template <> struct Raw8<double> {
  double2 a[4];
  __device__ __forceinline__ void load(const void* base, uint64_t e) {
    const double2* p = reinterpret_cast<const double2*>(reinterpret_cast<const double*>(base) + e);
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __ldg(p + i);
  }
  __device__ __forceinline__ void get(float (&v)[8]) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = (float)a[i].x * SX; v[2 * i + 1] = (float)a[i].y * SX; }   // the fp32 kernel rounds to float first, too
  }
};
template <typename T> __device__ __forceinline__ float load1s(const void* base, uint64_t e) { return load_sample<T>(base, e) * SX; }

// hi = v truncated to 11 significant bits (exact in fp16 for |v| < 65520), lo = fp16(v - hi)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const float ha = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
  const float hb = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
  const float2 d = __fadd2_rn(make_float2(a, b), make_float2(-ha, -hb));
  const __half2 h = __floats2half2_rn(ha, hb), l = __floats2half2_rn(d.x, d.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

constexpr int POST_WARPS = 4, POST_THREADS = 32 * POST_WARPS;
constexpr int NQ = 4, QROWS = 64, QBOX = 66;   // a stage lands as NQ TMA boxes of QBOX segments, QROWS apart (SEGS = 3 * 64 + 66)
static_assert((NQ - 1) * QROWS + QBOX == SEGS, "landing boxes must cover the stage exactly");
constexpr int ALL_THREADS = LOAD_THREADS + MMA_THREADS + POST_THREADS;          // 512: warps 0-3 convert, 4-11 MMA, 12-15 post (post warp 0 also fetches)
// Handoff buffers MMA warps -> post warps, laid out for the READER (post thread p owns symbols 16p .. 16p+15 = rows 2p, 2p+1):
//   U[j][p] (stride US): symbol 16p + j;   Z[f][row] (stride ZS): feature f of a row  -- reads are conflict-free, writes 2-way
constexpr int US = 121, NU = 16 * US, ZS = 264, NZ = 4 * ZS;

template <typename S> struct Smem {
  static constexpr int STAGE = (SEGS * S::SEGB + 127) / 128 * 128;   // bytes of one sample stage (TMA destinations are 128-byte aligned)
  static constexpr int O_Z = 2 * STAGE;                    // float2 [2][ROWS][4]: group features per tile parity
  static constexpr int O_U = O_Z + 2 * NZ * 8;             // float2 [2][NU]: in-group symbols (accumulator units) per tile parity
  static constexpr int O_MAX = O_U + 2 * NU * 8;           // float [2][LOAD_THREADS]
  static constexpr int O_MISC = O_MAX + 2 * LOAD_THREADS * 4;   // (only 2 x LOAD_WARPS floats of the slot are in use)
  static constexpr int TOTAL = O_MISC + 1792 + 128;        // + slack to align the base
};

struct Misc {
  float2 lpow[2][32];            // lam^(2n), n = 0..31, per pole: two rows per post thread
  float2 tot[POST_WARPS][4];     // warp totals of the group scan
  float2 carry[2][2];            // [tile parity][pole]: forward states entering row 0 (written by the previous tile at its row ADV_ROWS)
  float2 yex[POST_WARPS + 1];    // first symbol of every post warp (differential across warp edges)
  int ok[2];                     // [tile parity]: the tile's samples fit the fp16 split
  uint32_t tile_s[2], tile_u[2]; // tile index travelling with the sample stage / with the handoff buffers (~0u: no more work)
  PskTile pl_s[2], pl_u[2];      // ... and its descriptor (one global load per tile, by the loaders)
  int inb_s[2];                  // the stage is filled by TMA (raw fp32, tile inside its recording); 0: guarded loads by the converters
  uint64_t raw[2][NQ];           // sample stages: bytes of a landing sub-block (TMA, complete_tx)
};

static_assert(sizeof(Misc) <= 1792, "Misc outgrew its shared-memory slot");
__device__ __forceinline__ void bar_post() { asm volatile("bar.sync 2, %0;" ::"n"(POST_THREADS) : "memory"); }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
// acc + f.x * (m.x, m.y) + f.y * (m.z, m.w): one real 2x2 map of a complex state as two packed FMAs
__device__ __forceinline__ float2 map2(float4 m, float2 f, float2 acc) {
  return ffma2(make_float2(f.y, f.y), make_float2(m.z, m.w), ffma2(make_float2(f.x, f.x), make_float2(m.x, m.y), acc));
}
template <typename F, int... I> __device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) {
  (f(std::integral_constant<int, I>{}), ...);
}
template <int N, typename F> __device__ __forceinline__ void static_for(F&& f) { static_for_impl(f, std::make_integer_sequence<int, N>{}); }
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// 120 registers per thread at launch (61 440 of the SM's 65 536: the float64 edge kernel's 32-thread CTAs still fit beside a
// resident CTA of this kernel); the roles then re-balance per scheduler (one loader, two MMA, one post warp each): 48 + 2 x 168 + 96.
template <typename TIn, typename S>
__global__ void __maxnreg__(120) psk_mma_kernel(const __grid_constant__ MmaArgs a, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  using L = Smem<S>;
  Misc* misc = reinterpret_cast<Misc*>(smem + L::O_MISC);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      for (int q = 0; q < NQ; ++q) mbar_init(&misc->raw[i][q], 1);

    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x < 128) {         // lam^(2n) by binary powers
    const int i = ((int)threadIdx.x - 64) >> 5, n = 2 * (threadIdx.x & 31);
    float2 pw = make_float2(1.f, 0.f);
#pragma unroll
    for (int sft = 0; sft < 6; ++sft) if (n & (1 << sft)) pw = cmulf(pw, a.lam_pow[i][sft]);
    misc->lpow[i][n >> 1] = pw;
  }
  __syncthreads();
  // Work distribution: chunks of CHUNK consecutive tiles claimed from a global counter by the loaders; the tile index then
  // travels down the pipeline with the data (tile_s with the sample stage, tile_u with the handoff buffers).  A CTA that
  // starts late (e.g. its SM hosted a CTA of the float64 edge kernel) simply claims fewer chunks.

  if (wid < LOAD_WARPS) {
    setmaxnreg_dec<48>();
    // ================================================== loader warps ==================================================
    // A tile's samples arrive by TMA (requested by the first thread of the post warps, see there), four boxes, each
    // completing its own mbarrier.  120 converter threads rewrite the landed slots in place, [8 x fp32] -> [8 x fp16 hi |
    // 8 x fp16 lo] (x 2^14; 22 significant bits): a thread owns slot c of segments rg, rg + 12, ... (mapping below)  Tiles at the
    // ends of a recording and other sample types come through guarded loads instead.
    const int lt = (int)threadIdx.x;
    constexpr int UPS = S::SEG / 8, LROWS = 12, NIT = (SEGS + LROWS - 1) / LROWS;
    const bool lactive = lt < UPS * LROWS;
    // Slot of this thread inside a row group of 12 segments.  A slot is 32 bytes read and written as two 16-byte halves, so
    // the 8 lanes of a quarter warp must hit 8 different 16-byte bank groups: bank group = (21 seg + 2 c) mod 8, i.e. four
    // slots of an even segment give {0,2,4,6} + const and four of an odd one the other parity.  Threads 0..95: segment pairs
    // x slot blocks c 0-3 / 4-7; threads 96..119: the slots c 8, 9 of segments {0,1,4,5}, {2,3,6,7}, {8,9,10,11} (the last
    // group keeps a 2-way conflict).  (The plain mapping 10 rg + c made every one of these accesses a 2-way conflict.)
    int rg, c;
    if (lt < 96) { const int w16 = lt & 15, l8 = w16 & 7; rg = 2 * (lt >> 4) + (l8 >> 2); c = 4 * (w16 >> 3) + (l8 & 3); }
    else { const int j = lt - 96, g3 = j >> 3, l8 = j & 7, k = l8 >> 1; rg = g3 == 2 ? 8 + k : 2 * g3 + (k & 1) + 4 * (k >> 1); c = 8 + (l8 & 1); }
    for (uint32_t it = 0;; ++it) {
      const int st = it & 1;
      const uint32_t par = (it >> 1) & 1;
      tr_mark(a, it, 10, lt == 0);
      mbar_wait(&misc->raw[st][0], par);
      tr_mark(a, it, 11, lt == 0);
      if (misc->tile_s[st] == ~0u) { hb_arrive(HB_FULL + st); break; }
      // ---- convert the current tile in place -----------------------------------------------------------------------------
      unsigned char* sb = smem + st * L::STAGE + rg * S::SEGB + c * 32;
      float mx = 0.f;
      if (a.dbg & 1) {
        mx = 1.f;
#pragma unroll
        for (int q = 1; q < NQ; ++q) mbar_wait(&misc->raw[st][q], par);
      } else if (misc->inb_s[st]) {
        static_for<NIT>([&](auto I) {
          constexpr int item = decltype(I)::value;
          constexpr int q_now = (LROWS * item + LROWS - 1) / QROWS < NQ - 1 ? (LROWS * item + LROWS - 1) / QROWS : NQ - 1;
          constexpr int q_before = item == 0 ? 0 : ((LROWS * item - 1) / QROWS < NQ - 1 ? (LROWS * item - 1) / QROWS : NQ - 1);
          if (q_now != q_before) mbar_wait(&misc->raw[st][q_now], par);
          if (lactive && !(item == NIT - 1 && rg + LROWS * item >= SEGS)) {
            unsigned char* q8 = sb + item * (LROWS * S::SEGB);
            const float4 r0 = *reinterpret_cast<const float4*>(q8), r1 = *reinterpret_cast<const float4*>(q8 + 16);
            const float2 k = make_float2(SX, SX);
            const float2 p0 = __fmul2_rn(make_float2(r0.x, r0.y), k), p1 = __fmul2_rn(make_float2(r0.z, r0.w), k);
            const float2 p2 = __fmul2_rn(make_float2(r1.x, r1.y), k), p3 = __fmul2_rn(make_float2(r1.z, r1.w), k);
            uint4 hi, lo;
            split2(p0.x, p0.y, hi.x, lo.x); split2(p1.x, p1.y, hi.y, lo.y); split2(p2.x, p2.y, hi.z, lo.z); split2(p3.x, p3.y, hi.w, lo.w);
            mx = fmaxf(fmaxf(fmaxf(fabsf(p0.x), fabsf(p0.y)), fmaxf(fabsf(p1.x), fabsf(p1.y))), fmaxf(mx, fmaxf(fmaxf(fabsf(p2.x), fabsf(p2.y)), fmaxf(fabsf(p3.x), fabsf(p3.y)))));
            *reinterpret_cast<uint4*>(q8) = hi;
            *reinterpret_cast<uint4*>(q8 + 16) = lo;
          }
        });
      } else {
#pragma unroll
        for (int q = 1; q < NQ; ++q) mbar_wait(&misc->raw[st][q], par);     // (plain arrivals: keeps the phases in step)
        const PskTile pl = misc->pl_s[st];
        const int64_t c_N = (int64_t)pl.n;
        const int64_t w0 = (int64_t)a.n0 + (int64_t)pl.d0 * S::SPS - S::SPS * S::G;
        const int64_t c_w0a = w0 - (int64_t)((pl.off + (uint64_t)w0) & 7);
        for (int item = 0; item < NIT && lactive; ++item) {       // tiles at the ends of a recording, and other sample types: guarded loads
          if (rg + LROWS * item >= SEGS) break;
          const int64_t n = c_w0a + (int64_t)(rg * S::SEG + c * 8) + (int64_t)item * (LROWS * S::SEG);
          float v[8];
          if (n >= 0 && n + 8 <= c_N) {
            Raw8<TIn> r;
            r.load(a.samples, pl.off + (uint64_t)n);
            r.get(v);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (n + j >= 0 && n + j < c_N) ? load1s<TIn>(a.samples, pl.off + (uint64_t)(n + j)) : 0.f;
          }
          uint4 hi, lo;
          split2(v[0], v[1], hi.x, lo.x); split2(v[2], v[3], hi.y, lo.y); split2(v[4], v[5], hi.z, lo.z); split2(v[6], v[7], hi.w, lo.w);
#pragma unroll
          for (int j = 0; j < 8; ++j) mx = fmaxf(mx, fabsf(v[j]));
          *reinterpret_cast<uint4*>(sb + item * (LROWS * S::SEGB)) = hi;
          *reinterpret_cast<uint4*>(sb + item * (LROWS * S::SEGB) + 16) = lo;
        }
      }
      {
        const uint32_t wm = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));   // non-negative floats order like their bit patterns
        if (lane == 0) reinterpret_cast<uint32_t*>(smem + L::O_MAX)[st * LOAD_WARPS + wid] = wm;
      }
      tr_mark(a, it, 12, lt == 0);
      hb_arrive(HB_FULL + st);
    }
    return;
  }

  if (wid < LOAD_WARPS + MMA_WARPS) {
    setmaxnreg_inc<168>();
    // ==================================================== MMA warps =====================================================
    const int tid = (int)threadIdx.x - LOAD_THREADS, warp = tid >> 5;
    uint2 bf[S::NFRAG];                       // B fragments (taps hi / lo, feature weights hi / lo) of the current shift
    int cur_sh = -1;
    const int g = lane >> 2, q = lane & 3;
    for (uint32_t it = 0;; ++it) {
      const int st = it & 1;
      tr_mark(a, it, 0, tid == 0);
      hb_sync(HB_FULL + st);
      tr_mark(a, it, 1, tid == 0);
      const uint32_t t = misc->tile_s[st];
      if (t == ~0u) {                                    // end marker: forward it to the post warps
        if (it >= 2) hb_sync(HB_UEMPTY + st);
        if (tid == 0) misc->tile_u[st] = ~0u;
        __syncwarp();
        hb_arrive(HB_UFULL + st);
        break;
      }
      const PskTile pl = misc->pl_s[st];
      const int64_t w0 = (int64_t)a.n0 + (int64_t)pl.d0 * S::SPS - S::SPS * S::G;
      const int sh = (int)((pl.off + (uint64_t)w0) & 7);
      if (sh != cur_sh) {
        const uint2* src = a.frags + ((size_t)sh * S::NFRAG) * 32 + lane;
#pragma unroll
        for (int i = 0; i < S::NFRAG; ++i) bf[i] = __ldg(src + i * 32);
        cur_sh = sh;
      }
      if (a.dbg & 2) { if (it >= 2) hb_sync(HB_UEMPTY + st); if (tid == 0) misc->tile_u[st] = t; __syncwarp(); hb_arrive(HB_UFULL + st); continue; }
      int okv = 1;
      if (tid == 0) {                                   // range check of the tile's samples (scaled by 2^14): one maximum per loader warp
        const float4 mxs = *reinterpret_cast<const float4*>(smem + L::O_MAX + st * LOAD_WARPS * 4);
        const float m = fmaxf(fmaxf(mxs.x, mxs.y), fmaxf(mxs.z, mxs.w));
        okv = (m < 65000.f && m >= 0.0625f) ? 1 : 0;
      }
      tr_mark(a, it, 13, tid == 0);
      // the symbol / feature buffers of this parity must have been drained by the post warps (two tiles ago)
      if (it >= 2) hb_sync(HB_UEMPTY + st);
      tr_mark(a, it, 14, tid == 0);
      if (tid == 0) { misc->ok[st] = okv; misc->tile_u[st] = t; misc->pl_u[st] = pl; }
      __syncwarp();                                     // (the waits above leave the lanes diverged; ldmatrix / mma are warp-collective)
      float2* Zs = reinterpret_cast<float2*>(smem + L::O_Z) + st * NZ;
      float2* Us = reinterpret_cast<float2*>(smem + L::O_U) + st * NU;

      // ---- two m-tiles per warp (the very last m-tile of the CTA tile is the look-ahead: features only) -----------------
      const uint32_t sbase = smem_u32(smem + st * L::STAGE);
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int mt = 2 * warp + mi;
        const bool main_tile = mt < MAIN_ROWS / 16;
        const uint32_t abase = sbase + (uint32_t)((16 * mt + (lane & 15)) * S::SEGB + (lane >> 4) * 32);
        float ah[2][4], al[2][4], fh[4], fl[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { ah[0][i] = ah[1][i] = al[0][i] = al[1][i] = fh[i] = fl[i] = 0.f; }
        if (main_tile) {
#pragma unroll
          for (int k = 0; k < S::KS; ++k) {
            const uint32_t ko = (uint32_t)((k / 5) * S::SEGB + (k % 5) * 64);
            uint32_t xa[4], xl[4];
            ldsm4(xa, abase + ko);
            const bool need_lo = k >= S::LOA && k <= S::LOB;
            if (need_lo) ldsm4(xl, abase + 16 + ko);
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
          // in-group symbols by symbol index: thread (g, q) holds symbols q and 4 + q of rows g and g + 8
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int n = 0; n < 2; ++n)
            {
              const int row = 16 * mt + g + 8 * h;          // symbol 8 row + 4 n + q = 16 (row / 2) + (8 (row & 1) + 4 n + q)
              Us[(8 * (row & 1) + 4 * n + q) * US + (row >> 1)] = make_float2(ah[n][2 * h] + al[n][2 * h], ah[n][2 * h + 1] + al[n][2 * h + 1]);
            }
        } else {
#pragma unroll
          for (int k = S::FTA; k <= S::FTB; ++k) {
            const uint32_t ko = (uint32_t)((k / 5) * S::SEGB + (k % 5) * 64);
            uint32_t xa[4], xl[4];
            ldsm4(xa, abase + ko);
            ldsm4(xl, abase + 16 + ko);
            hmma(fh, xa, bf[S::O_FTH + k - S::FTA]);
            hmma(fl, xa, bf[S::O_FTL + k - S::FTA]);
            hmma(fl, xl, bf[S::O_FTH + k - S::FTA]);
          }
        }
        // group features: thread (g, q) holds feature q (Zf0, Zb0, Zf1, Zb1) of rows g and g + 8
        Zs[q * ZS + 16 * mt + g] = make_float2(fh[0] + fl[0], fh[1] + fl[1]);
        Zs[q * ZS + 16 * mt + g + 8] = make_float2(fh[2] + fl[2], fh[3] + fl[3]);
        if (mi == 0) tr_mark(a, it, 15, tid == 0);
      }
      __syncwarp();
      tr_mark(a, it, 2, tid == 0);
      hb_arrive(HB_UFULL + st);                          // stage free for the loaders, symbols ready for the post warps
    }
    return;
  }

  // ===================================================== post warps ======================================================
  // thread p <-> rows 2p, 2p + 1 of the tile = symbols 16p .. 16p + 15: group scan of the slow-pole states (registers), the
  // out-of-group sources added to the symbols, differential decisions -- one 32-bit word of DQPSK bits per thread
  setmaxnreg_dec<96>();
  {
    const int p = (int)threadIdx.x - LOAD_THREADS - MMA_THREADS, warp = p >> 5;
    const int trp = (a.dbg >> 4) * 32;               // experiments: which post warp stamps the trace (FB_MMA_DBG = 16 w)
    // ---- fetch (first post thread): claims tiles (chunks of CHUNK consecutive tiles from a global counter) and brings a tile's
    // samples in with four TMA tile copies through a 2-D tensor map of the sample buffer whose rows are the 80-sample
    // segments and whose box is 84 samples wide -- the 336-byte shared-memory pitch that keeps ldmatrix conflict-free comes
    // for free, no load instruction is issued, nothing passes through registers or the LSU, the whole tile is in flight at
    // once.  A stage is free the moment its tile's hand-off to the post warps completes (every MMA warp has read it), which
    // is exactly when this thread wakes up for that tile: tile it + 2 is requested at the top of iteration it.
    const uint32_t n_chunks = (a.n_tiles + CHUNK - 1) / CHUNK;
    uint32_t f_t = 0, f_end = 0;
    bool f_done = false, f_have = false;
    PskTile f_pl{};
    auto fetch = [&](int fs) {                           // thread p == 0 only; stage fs is free
      if (f_t == f_end) {
        const uint32_t ch = atomicAdd(a.chunk_ctr, 1u);
        if (ch >= n_chunks) {                            // no more work: pass the end marker down the pipeline
          misc->tile_s[fs] = ~0u;
#pragma unroll
          for (int q = 0; q < NQ; ++q) mbar_arrive(&misc->raw[fs][q]);
          f_done = true;
          return;
        }
        f_t = ch * CHUNK; f_end = min(a.n_tiles, f_t + CHUNK); f_have = false;
      }
      const uint32_t tc = f_t++;
      const PskTile pl = f_have ? f_pl : a.tiles[tc];
      // window origin of row 0, moved down to a multiple of 8 elements of the sample buffer
      const int64_t w0 = (int64_t)a.n0 + (int64_t)pl.d0 * S::SPS - S::SPS * S::G;     // H+ = 8 sps
      const int64_t w0a = w0 - (int64_t)((pl.off + (uint64_t)w0) & 7);
      const uint64_t e0 = pl.off + (uint64_t)w0a;       // first element of the stage in the sample buffer
      const uint32_t e8 = (uint32_t)(e0 >> 3);          // (e0 is a multiple of 8 and below 2^35: 32-bit arithmetic)
      const uint32_t c1 = e8 / (S::SEG / 8), c0 = (e8 - c1 * (S::SEG / 8)) * 8;
      const bool inb = sizeof(TIn) == 4 && w0a >= 0 && w0a + (int64_t)SEGS * S::SEG <= (int64_t)pl.n && (e0 >> 35) == 0 && (uint64_t)c1 + SEGS <= a.tm_rows && !(a.dbg & 1);
      // the stage was last written (converters) and read (ldmatrix) through the generic proxy; (before this thread's own
      // stores below, so the fence has nothing of ours to wait for)
      if (inb) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      misc->tile_s[fs] = tc; misc->pl_s[fs] = pl; misc->inb_s[fs] = inb ? 1 : 0;
      if (inb) {
        const uint32_t dst = smem_u32(smem + fs * L::STAGE);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const uint32_t mb = smem_u32(&misc->raw[fs][q]);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "n"(QBOX * S::SEGB) : "memory");
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                       ::"r"(dst + q * (QROWS * S::SEGB)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(c0), "r"(c1 + q * QROWS), "r"(mb) : "memory");
        }
      } else {
#pragma unroll
        for (int q = 0; q < NQ; ++q) mbar_arrive(&misc->raw[fs][q]);
      }
      f_have = f_t < f_end;
      if (f_have) f_pl = a.tiles[f_t];                   // consumed by the next call: the load's latency stays off the critical path
    };
    if (p == 0) { fetch(0); if (!f_done) fetch(1); }
    uint64_t prev_off = ~0ull;
    bool prev_ok = false;                              // the previous tile's samples fitted the fp16 split: its carried state is usable
    int prev_d0 = 0;
    for (uint32_t it = 0;; ++it) {
      const int st = it & 1;
      tr_mark(a, it, 3, p == trp);
      hb_sync(HB_UFULL + st);
      __syncwarp();                                     // the wait leaves the lanes diverged; the scan below is all warp shuffles
      tr_mark(a, it, 4, p == trp);
      const uint32_t t = misc->tile_u[st];
      if (t == ~0u) break;
      if (p == 0 && !f_done) fetch(st);                  // stage st is free: request tile it + 2
      const PskTile pl = misc->pl_u[st];
      // forward state carried from the previous tile -- unless that tile went to the redo list: a stretch louder than the
      // fp16 range turns its products, and with them the carried state, into inf / NaN for the rest of the chunk
      const bool chained = (pl.off == prev_off) && (pl.d0 == prev_d0 + TILE_SYMS) && prev_ok;
      prev_off = pl.off; prev_d0 = pl.d0;
      prev_ok = misc->ok[st] != 0;
      // ---- forward slow-pole state at the tile's first group when it cannot be carried: direct sum over the previous wlen
      // samples (exact start-up state of scipy's filtfilt when the record start is within reach; psk_v2.cu has the algebra)
      if (!chained) {                                   // uniform over the post warps
        const int64_t n_d0 = (int64_t)a.n0 + (int64_t)pl.d0 * S::SPS;
        const bool near_left = (n_d0 - a.wlen) <= (int64_t)a.n0;
        const int cnt = (int)min((int64_t)a.wlen, n_d0 - (near_left ? (int64_t)a.n0 : (int64_t)0));
        float2 s0 = make_float2(0.f, 0.f), s1 = s0;
        for (int k0 = 1 + p; k0 <= cnt; k0 += 8 * POST_THREADS) {       // 8 independent loads in flight per thread
          float xs[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { const int k = k0 + j * POST_THREADS; xs[j] = k <= cnt ? load_sample<TIn>(a.samples, pl.off + (uint64_t)(n_d0 - k)) : 0.f; }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 w = __ldg(&a.slow_pw4[min(k0 + j * POST_THREADS, a.wpad - 1)]);
            s0.x = fmaf(xs[j], w.x, s0.x); s0.y = fmaf(xs[j], w.y, s0.y); s1.x = fmaf(xs[j], w.z, s1.x); s1.y = fmaf(xs[j], w.w, s1.y);
          }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          s0.x += __shfl_xor_sync(0xffffffffu, s0.x, off); s0.y += __shfl_xor_sync(0xffffffffu, s0.y, off);
          s1.x += __shfl_xor_sync(0xffffffffu, s1.x, off); s1.y += __shfl_xor_sync(0xffffffffu, s1.y, off);
        }
        if (lane == 0) { misc->tot[warp][0] = s0; misc->tot[warp][1] = s1; }
        bar_post();
        if (warp == 0 && lane < 2) {
          float2 sum = make_float2(0.f, 0.f);
          for (int w = 0; w < POST_WARPS; ++w) { sum.x += misc->tot[w][lane].x; sum.y += misc->tot[w][lane].y; }
          if (near_left) {
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
            sum.x += add.x; sum.y += add.y;
          }
          misc->carry[it & 1][lane] = make_float2(sum.x * a.state_scale, sum.y * a.state_scale);
        }
        bar_post();                                       // tot is reused by the scan below
      }
      const float2* Zs = reinterpret_cast<const float2*>(smem + L::O_Z) + st * NZ;
      const float2* Us = reinterpret_cast<const float2*>(smem + L::O_U) + st * NU;
      const int okv = misc->ok[st];

      // ---- group scan, two rows per thread.  F[r+1] = lam F[r] + Zf[r] (F[0] = carry);  Bk[r] = Zb[r+1] + lam Bk[r+1] (Bk[255] = 0) ----
      // features Zf0, Zb0, Zf1, Zb1 (f = 0..3) of rows 2p (a) and 2p + 1 (b): one 16-byte load per feature
      const float4 q0 = *reinterpret_cast<const float4*>(&Zs[0 * ZS + 2 * p]), q1 = *reinterpret_cast<const float4*>(&Zs[1 * ZS + 2 * p]);
      const float4 q2 = *reinterpret_cast<const float4*>(&Zs[2 * ZS + 2 * p]), q3 = *reinterpret_cast<const float4*>(&Zs[3 * ZS + 2 * p]);
      const float2 zfa[2] = {make_float2(q0.x, q0.y), make_float2(q2.x, q2.y)}, zba[2] = {make_float2(q1.x, q1.y), make_float2(q3.x, q3.y)};
      const float2 zfb[2] = {make_float2(q0.z, q0.w), make_float2(q2.z, q2.w)}, zbb[2] = {make_float2(q1.z, q1.w), make_float2(q3.z, q3.w)};
      float2 v[4];                                      // thread totals: forward (pole 0, 1): lam a + b;  backward: a + lam b
#pragma unroll
      for (int i = 0; i < 2; ++i) { v[i] = cfmaf(a.lam_pow[i][0], zfa[i], zfb[i]); v[2 + i] = cfmaf(a.lam_pow[i][0], zbb[i], zba[i]); }
#pragma unroll
      for (int sft = 0; sft < 5; ++sft) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 mm = a.lam_pow[k & 1][sft + 1];  // (lam^2)^(2^sft)
          if (k < 2) {
            const float ox = __shfl_up_sync(0xffffffffu, v[k].x, 1 << sft), oy = __shfl_up_sync(0xffffffffu, v[k].y, 1 << sft);
            if (lane >= (1 << sft)) v[k] = cfmaf(mm, make_float2(ox, oy), v[k]);
          } else {
            const float ox = __shfl_down_sync(0xffffffffu, v[k].x, 1 << sft), oy = __shfl_down_sync(0xffffffffu, v[k].y, 1 << sft);
            if (lane + (1 << sft) < 32) v[k] = cfmaf(mm, make_float2(ox, oy), v[k]);
          }
        }
      }
      if (lane == 31) { misc->tot[warp][0] = v[0]; misc->tot[warp][1] = v[1]; }
      if (lane == 0) { misc->tot[warp][2] = v[2]; misc->tot[warp][3] = v[3]; }
      tr_mark(a, it, 6, p == trp);
      bar_post();
      tr_mark(a, it, 7, p == trp);
      float2 c = make_float2(0.f, 0.f);                 // state entering this warp's 64 rows from the left (k < 2) / right (k >= 2)
      if (lane < 4) {
        const int k = lane;
        const float2 M = a.lam_pow[k & 1][6];           // lam^64
        if (k < 2) {
          c = misc->carry[it & 1][k];
          for (int w = 0; w < warp; ++w) c = cfmaf(M, c, misc->tot[w][k]);
        } else {
          for (int w = POST_WARPS - 1; w > warp; --w) c = cfmaf(M, c, misc->tot[w][k]);
        }
      }
      float2 Fa[2], Fb[2], Ba[2], Bb[2];                // states of rows a = 2p, b = 2p + 1, per pole
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 car = make_float2(__shfl_sync(0xffffffffu, c.x, k), __shfl_sync(0xffffffffu, c.y, k));
        float ex, ey;
        if (k < 2) { ex = __shfl_up_sync(0xffffffffu, v[k].x, 1); ey = __shfl_up_sync(0xffffffffu, v[k].y, 1); }
        else { ex = __shfl_down_sync(0xffffffffu, v[k].x, 1); ey = __shfl_down_sync(0xffffffffu, v[k].y, 1); }
        const bool has = (k < 2) ? lane > 0 : lane < 31;
        const float2 excl = has ? make_float2(ex, ey) : make_float2(0.f, 0.f);
        const float2 e = cfmaf(misc->lpow[k & 1][(k < 2) ? lane : 31 - lane], car, excl);
        if (k < 2) { Fa[k] = e; Fb[k] = cfmaf(a.lam_pow[k][0], e, zfa[k]); }          // F[2p], F[2p+1] = lam F[2p] + Zf[2p]
        else { Bb[k - 2] = e; Ba[k - 2] = cfmaf(a.lam_pow[k - 2][0], e, zbb[k - 2]); }  // Bk[2p+1], Bk[2p] = Zb[2p+1] + lam Bk[2p+1]
      }
      // the forward state a chained next tile starts from: the state entering row ADV_ROWS (an even row)
      if (2 * p == ADV_ROWS) { misc->carry[(it + 1) & 1][0] = Fa[0]; misc->carry[(it + 1) & 1][1] = Fa[1]; }

      // ---- symbols of the two rows: in-group part + out-of-group sources through the states --------------------------------
      float2 y[17];
      if (p < MAIN_ROWS / 2) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float2 F0 = h ? Fb[0] : Fa[0], F1 = h ? Fb[1] : Fa[1], B0 = h ? Bb[0] : Ba[0], B1 = h ? Bb[1] : Ba[1];
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            {
              float2 yy = Us[(8 * h + s) * US + p];
              yy = map2(a.maps_c[(0 * 8 + s) * 2 + 0], F0, yy);
              yy = map2(a.maps_c[(0 * 8 + s) * 2 + 1], B0, yy);
              yy = map2(a.maps_c[(1 * 8 + s) * 2 + 0], F1, yy);
              yy = map2(a.maps_c[(1 * 8 + s) * 2 + 1], B1, yy);
              y[8 * h + s] = yy;
            }
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) y[i] = make_float2(0.f, 0.f);
      }
      __syncwarp();
      tr_mark(a, it, 8, p == trp);
      hb_arrive(HB_UEMPTY + st);                            // this warp no longer reads the handoff buffers
      // first symbol of the next thread (across warps through shared memory)
      if (lane == 0) misc->yex[warp] = y[0];
      y[16] = make_float2(__shfl_down_sync(0xffffffffu, y[0].x, 1), __shfl_down_sync(0xffffffffu, y[0].y, 1));
      bar_post();
      tr_mark(a, it, 9, p == trp);
      if (lane == 31 && warp + 1 < POST_WARPS) y[16] = misc->yex[warp + 1];

      // ---- differential decisions: symbols 16p .. 16p + 15 against their successors ------------------------------------
      {
        const int e0 = p * 16;
        const int nd = pl.d1 - pl.d0;                   // multiple of 32, <= TILE_SYMS
        // d = y[k+1] conj(y[k]) rho in accumulator units (|y| < 2^35, so the products stay far below the fp32 range).
        // DQPSK: the dibit of psk_decide() is (sign(dr + di), sign(dr - di)) -- two funnel shifts, no branches, the 16 symbols
        // independent of each other.  (+ 0.f turns -0 into +0: all-zero input decides 00 like the reference.  The only other
        // difference from the literal rule is a sum that is EXACTLY zero next to a non-zero one, a symbol of margin 0.)
        uint32_t part = 0;
        if (a.bps == 2) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 prev = y[i], cur = y[i + 1];
            const float tr = fmaf(cur.x, prev.x, cur.y * prev.y), ti = fmaf(cur.y, prev.x, -cur.x * prev.y);
            const float dr = fmaf(tr, a.rho.x, -ti * a.rho.y), di = fmaf(tr, a.rho.y, ti * a.rho.x);
            const float sa = (dr + di) + 0.f, sb = (dr - di) + 0.f;
            part = __funnelshift_l(__float_as_uint(sa), part, 1);
            part = __funnelshift_l(__float_as_uint(sb), part, 1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 prev = y[i], cur = y[i + 1];
            const float tr = fmaf(cur.x, prev.x, cur.y * prev.y), ti = fmaf(cur.y, prev.x, -cur.x * prev.y);
            const float dr = fmaf(tr, a.rho.x, -ti * a.rho.y);
            part = (part << 1) | (dr < 0.f ? 1u : 0u);
          }
        }
        if (a.bps == 2) {
          if (e0 < nd && e0 < TILE_SYMS && okv) a.bits[pl.word_off + (uint64_t)((pl.d0 + e0) >> 4)] = __byte_perm(part, 0, 0x0123);
        } else {
          const uint32_t other = __shfl_down_sync(0xffffffffu, part, 1);
          if ((lane & 1) == 0 && e0 < nd && e0 < TILE_SYMS && okv) a.bits[pl.word_off + (uint64_t)((pl.d0 + e0) >> 5)] = __byte_perm((part << 16) | other, 0, 0x0123);
        }
        if (p == 0 && !okv) a.redo_list[atomicAdd(a.redo_count, 1u)] = t;
      }
      tr_mark(a, it, 5, p == trp);
      bar_post();                                        // tot / yex / carry are rewritten by the next tile
    }
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
  float2 lam[2], lam_pow[2][7];
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
    for (int s = 0; s < 7; ++s) { T.lam_pow[i][s] = make_float2((float)x.r, (float)x.i); x = cmul_(x, x); }
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

// Experiments only: clock64 stamps of the last traced launch, [sm_count][TR_TILES][16]; returns TR_TILES (0 when not traced).
extern "C" int fb_debug_mma_trace(fb_handle* h, long long* out, int n_ctas) {
  if (!h || !h->mma_trace.p || !out) return 0;
  cudaStreamSynchronize(h->stream);
  cudaMemcpy(out, h->mma_trace.p, (size_t)std::min(n_ctas, h->sm_count) * TR_TILES * 16 * 8, cudaMemcpyDeviceToHost);
  return TR_TILES;
}

// Launches the tensor-core kernel over `n_tiles` tile descriptors (tile size TILE_SYMS).  redo: device buffer of
// 1 + n_tiles uint32 (count, then the tiles the fp32 kernel must evaluate); the count is zeroed here.
// 2-D tensor map over the float32 sample buffer: row r = elements [80 r, 80 r + 164) (rows overlap: the stride is one segment),
// box = 84 x QBOX elements, so that a box lands as QBOX segments at the 336-byte pitch of the staged layout.  Returns the row count.
typedef CUresult (*fb_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static uint32_t make_sample_tmap(const void* d_samples, uint64_t total_samples, CUtensorMap* tm) {
  using S = Sched10;
  // resolved once (function-local static: thread-safe initialisation; handles of several host threads share it)
  static const fb_tmap_encode_fn enc = []() -> fb_tmap_encode_fn {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) return (fb_tmap_encode_fn)fn;
    cudaGetLastError();
    return nullptr;
  }();
  memset(tm, 0, sizeof(*tm));
  const uint64_t inner = (uint64_t)S::SEG + S::SEGB / 4;          // 164: a box (84 wide) may start anywhere inside its first segment
  if (!enc || getenv("FB_PSK_NO_TMA") || ((uintptr_t)d_samples & 15) || total_samples < inner + (uint64_t)S::SEG * SEGS) return 0;
  const uint64_t rows = (total_samples - inner) / S::SEG + 1;
  if (rows > 0xffffffffull) return 0;
  const cuuint64_t gdim[2] = {inner, rows}, gstr[1] = {(cuuint64_t)S::SEG * 4};
  const cuuint32_t box[2] = {(cuuint32_t)(S::SEGB / 4), (cuuint32_t)QBOX}, estr[2] = {1, 1};
  if (enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(d_samples), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 0;
  return (uint32_t)rows;
}

int fb_psk_mma_launch(fb_handle* h, const fb_psk_design& d, const float* taps, const void* d_samples, uint64_t total_samples, int dtype,
                      const PskTile* d_tiles, uint32_t n_tiles, uint32_t* d_bits, uint32_t* d_redo, int edge_ctas) {
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
    for (int s = 0; s < 7; ++s) a.lam_pow[i][s] = T->lam_pow[i][s];
    a.slow_p[2 * i] = d.slow_p[2 * i]; a.slow_p[2 * i + 1] = d.slow_p[2 * i + 1];
  }
  for (int i = 0; i < 32; ++i) a.maps_c[i] = make_float4(T->maps[i].x, T->maps[i].z, T->maps[i].y, T->maps[i].w);
  a.state_scale = T->state_scale;
  a.rho = make_float2(d.rho[0], d.rho[1]);
  CUtensorMap tm;
  a.tm_rows = dtype == FB_F32 ? make_sample_tmap(d_samples, total_samples, &tm) : (memset(&tm, 0, sizeof(tm)), 0u);
  a.bits = d_bits; a.redo_count = d_redo; a.chunk_ctr = d_redo + 1; a.redo_list = d_redo + 2;
  if (const char* e = getenv("FB_MMA_DBG")) a.dbg = atoi(e);
  if (getenv("FB_MMA_TRACE")) {
    int rc2 = fb_ensure(h, h->mma_trace, (size_t)h->sm_count * TR_TILES * 16 * 8);
    if (rc2) return rc2;
    FB_CUDA(h, cudaMemsetAsync(h->mma_trace.p, 0, (size_t)h->sm_count * TR_TILES * 16 * 8, h->stream));
    a.trace = (long long*)h->mma_trace.p;
  }
  FB_CUDA(h, cudaMemsetAsync(d_redo, 0, 8, h->stream));
  const int smem = Smem<S>::TOTAL;
  // One persistent CTA per SM -- minus the SMs left to the float64 edge kernel that runs beside this one (32-thread CTAs, 8 to
  // an SM).  A CTA of this kernel takes 61 440 of an SM's 65 536 registers, so an edge CTA (4 096) fits beside it only when
  // nothing else holds a register on that SM; in a process whose NCCL communicator has NVLS enabled that is no longer true
  // (measured on 2 GPUs: 13.0 ms per step with no SM left free, 9.7 with one, 5.8 with two; 6.0 without NCCL at all).
  int reserve = edge_ctas > 0 ? std::min(std::max(1, (edge_ctas + 7) / 8), h->sm_count / 8) : 0;
  if (const char* e = getenv("FB_MMA_GRID_MINUS")) reserve = std::max(0, atoi(e));             // tuning knob
  const int grid = (int)std::min<uint32_t>(n_tiles, (uint32_t)std::max(1, h->sm_count - reserve));
#define FB_MMA_LAUNCH(TIN)                                                                                              \
  do {                                                                                                                  \
    FB_CUDA(h, cudaFuncSetAttribute(psk_mma_kernel<TIN, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));        \
    psk_mma_kernel<TIN, S><<<grid, ALL_THREADS, smem, h->stream>>>(a, tm);                                 \
  } while (0)
  if (dtype == FB_F32) FB_MMA_LAUNCH(float);
  else if (dtype == FB_F64) FB_MMA_LAUNCH(double);
  else FB_MMA_LAUNCH(int16_t);
#undef FB_MMA_LAUNCH
  h->launches++;
  FB_CUDA(h, cudaGetLastError());
  return FB_OK;
}
