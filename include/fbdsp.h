/* libfbdsp.so -- C ABI of the B200-native FileBeep batch demodulation engine.
 *
 * The reference (szumanski/Audio-Modem-Radio) has no FFI: its receive hot path is a set of
 * Python module-level functions imported by name (decoder.py:12-14).  This header is the
 * boundary a binding for that path talks to; each entry point cites the reference interface
 * it replaces.  INTEGRATION.md shows the ctypes stub a maintainer adds on the reference side.
 *
 * Conventions: plain pointers and sizes, no torch / C++ types; caller owns every buffer;
 * return 0 on success or a negative FB_E* code (fb_strerror); per-recording results carry
 * their own FB_ST_* status so one bad recording never fails a batch.  One handle owns one
 * CUDA stream and its device workspace; calls on one handle from several host threads are
 * serialised by the library (a per-handle lock held for the whole call), different handles
 * are independent and run concurrently.  There is NO CPU fallback: every entry point that does work
 * fails with FB_ECUDA when no sm_100 device is usable.
 */
#ifndef FBDSP_H
#define FBDSP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB_ABI_VERSION 1

/* sample storage of the input recordings */
enum { FB_F32 = 0, FB_F64 = 1, FB_S16 = 2 };   /* S16: PCM16, value/32768 as soundfile does (decoder.py:381) */

/* flags for *_batch calls */
enum {
  FB_SAMPLES_ON_DEVICE = 1,   /* `samples` is a device pointer (else host; pinned host copies fastest) */
  FB_OUT_ON_DEVICE     = 2,   /* out / out_len / sync_idx / status are device pointers               */
  FB_ASYNC             = 4    /* do not synchronise the handle's stream before returning              */
};

/* library return codes */
enum { FB_OK = 0, FB_EINVAL = -1, FB_ECUDA = -2, FB_ENOMEM = -3, FB_EUNSUPPORTED = -4 };

/* per-recording status */
enum {
  FB_ST_OK = 0,
  FB_ST_EMPTY = 1,       /* fewer than 2 symbols: the reference returns b'' (modem.py:95-96, 211)          */
  FB_ST_TOO_SHORT = 2,   /* N <= filtfilt padlen: the reference raises ValueError (scipy filtfilt)       */
  FB_ST_UNSUPPORTED = 3  /* parameter set only servable by whole-record float64 evaluation and the record does not fit (more
                            than 2^26 samples, or 24 bytes of scratch per sample exceed half of the free device memory) */
};

typedef struct fb_handle fb_handle;

/* Host-designed description of one DPSK parameter set (scheme, baud, carrier, fs).  Filter design
 * (scipy.signal.butter, float64) stays in the host wrapper so coefficient parity with the reference
 * is exact (modem.py:76,87,197,203); see audio-modem-radio_b200/fbdsp/design.py for every field. */
#define FB_MAX_SLOW 4
typedef struct fb_psk_design {
  int32_t sps, n0, bits_per_sym;     /* int(fs/baud); first symbol instant; 1 = DBPSK slicer, 2 = DQPSK slicer */
  int32_t nt, dl, dh;                /* polyphase taps per row; tap span in symbols: q in [-(dl*sps+sps-1), dh*sps] */
  int32_t nslow, wcols;              /* slow conjugate pole pairs; warm-up length of their recursion (symbols) */
  int32_t zone_left, zone_right;     /* samples near each end that belong to the edge kernel */
  int32_t w_bp, w_lp;                /* decay lengths (samples) of the band-pass / low-pass recursions */
  int32_t emulate_only, pad_bp, pad_lp, reserved;
  double  cycles_per_sample;         /* carrier / fs */
  double  bp_b[9], bp_a[9], bp_zi[8];/* butter(4, band) and lfilter_zi: what filtfilt runs (modem.py:76-77,197-198) */
  double  lp_b[5], lp_a[5], lp_zi[4];/* butter(4, baud/nyq)                               (modem.py:87-88,203-204) */
  float   rho[2];                    /* exp(-j 2 pi carrier sps / fs): the LO's rotation between two symbols */
  double  slow_p[2 * FB_MAX_SLOW];
  float   slow_lam[2 * FB_MAX_SLOW];
  float   slow_rp[2 * FB_MAX_SLOW], slow_rpc[2 * FB_MAX_SLOW];
  float   slow_rm[2 * FB_MAX_SLOW], slow_rmc[2 * FB_MAX_SLOW];
} fb_psk_design;

int         fb_abi_version(void);
int         fb_device_count(void);                 /* number of usable CUDA devices (0 if none)            */
const char* fb_strerror(int code);
const char* fb_last_error(fb_handle* h);           /* text of the last CUDA failure on this handle         */

fb_handle*  fb_create(int device);                 /* NULL when the device cannot be opened                */
void        fb_destroy(fb_handle* h);
void*       fb_stream(fb_handle* h);               /* the handle's cudaStream_t, for event timing / interop */
int         fb_sync(fb_handle* h);
uint64_t    fb_kernel_launches(fb_handle* h);      /* kernels launched by this handle so far               */

/* Per-kernel device timing for roofline accounting (bench.py): when enabled, CUDA events are recorded on the
 * launching stream right before and after the dominant kernel of each *_batch call; fb_kernel_ms returns the
 * duration of the most recent one (blocks until it has finished), or a negative value if none was recorded. */
int         fb_set_profiling(fb_handle* h, int on);
float       fb_kernel_ms(fb_handle* h);

/* Upper bound of the raw byte count one recording of n_samples can produce (size your `out` slots). */
uint64_t    fb_psk_out_bound(const fb_psk_design* d, uint64_t n_samples);

/* Batch DPSK demodulation: replaces bpsk_demodulate (modem.py:68-135), qpsk_demodulate
 * (modem.py:189-266) and their aliases psk8_demodulate / ofdm_demodulate_simple / psk31_demodulate
 * (modem.py:348,375,397) for n_rec independent recordings that share one parameter set.
 *   taps      host, complex float [sps][nt]      (design.py PskDesign.taps)
 *   slow_w    host, complex float [nslow][sps+1] (p^j)
 *   samples   recordings back to back; recording r is samples[offsets[r] .. offsets[r+1])  (element units)
 *   offsets, out_offsets   HOST arrays of n_rec+1 entries; out slot r is out[out_offsets[r] .. out_offsets[r+1])
 *   out_len[r]   bytes written for recording r (what the reference returns as `bytes`)
 *   sync_idx[r]  bit index where "0100011001000010" was first found (modem.py:118,248), -1 if absent
 *   status[r]    FB_ST_*
 */
int fb_psk_demod_batch(fb_handle* h, const fb_psk_design* d, const float* taps, const float* slow_w,
                       int n_rec, const void* samples, const uint64_t* offsets, int dtype, int flags,
                       uint8_t* out, const uint64_t* out_offsets,
                       uint64_t* out_len, int64_t* sync_idx, int32_t* status);

/* Debug / test hook (host only, no device needed): the float64 band matrix [8 window shifts][240][16] of the tensor-core
 * DPSK kernel (csrc/psk_mma.cu; algebra in audio-modem-radio_b200/fbdsp/mma_tables.py) for design d, so that the C++ table
 * builder can be checked against the Python statement of the same algebra.  Returns the number of 16-sample k-steps, or
 * -1 when the design is not served by that kernel (the fp32 kernel takes it). */
int fb_debug_mma_band(const fb_psk_design* d, const float* taps, double* bfir_out);

/* Debug / test hook: decided bit stream of the last fb_psk_demod_batch call on this handle
 * (MSB-first packed, recording r at byte offset bit_offsets[r], n_bits[r] valid bits).  Host buffers. */
int fb_psk_last_bits(fb_handle* h, int rec, uint8_t* bits_out, uint64_t cap_bytes, uint64_t* n_bits);

/* ---- v2 FSK: replaces fsk_demodulate (modem.py:298-341) and its aliases fsk_high_speed_demodulate / ft8_demodulate
 * (modem.py:355,391) for tone sets whose Butterworth design is valid (every product default raises ValueError in the
 * host wrapper, exactly as scipy does in the reference).  tone 0 = mark, tone 1 = space.                           */
typedef struct fb_fsk_design {
  int32_t spb, pad;                  /* int(fs/baud); filtfilt padlen = 3*max(len(a),len(b)) = 21                    */
  int32_t w[2];                      /* warm-up samples per tone (pole decay to 1e-12)                               */
  double  b[2][7], a[2][7], zi[2][6];/* butter(3, [(f-baud)/nyq, (f+baud)/nyq], 'band') and lfilter_zi (modem.py:307) */
} fb_fsk_design;
uint64_t fb_fsk_out_bound(const fb_fsk_design* d, uint64_t n_samples);
/* Same calling convention as fb_psk_demod_batch. */
int fb_fsk_demod_batch(fb_handle* h, const fb_fsk_design* d, int n_rec, const void* samples, const uint64_t* offsets,
                       int dtype, int flags, uint8_t* out, const uint64_t* out_offsets,
                       uint64_t* out_len, int64_t* sync_idx, int32_t* status);

/* ---- "v1" streaming demodulators (present in the reference only as __pycache__/modem.cpython-39.pyc;
 * these are the north_star's named kernels -- Goertzel tone bank, I/Q integrate-and-dump, 8PSK slicer, per-symbol DFT OFDM
 * demap.  Parity is against oracle/modem_v1.py, which no reference test pins).  Per symbol k the kernel evaluates
 * F_m = sum_{j<len} x[k*sps + off0 + j] * table[m][j] (complex float64 weights) and applies the mode's hard decision.   */
enum { FB_V1_BPSK = 0, FB_V1_QPSK = 1, FB_V1_PSK8 = 2, FB_V1_OFDM = 3, FB_V1_FSK = 4 };
typedef struct fb_v1_params {
  int32_t mode, sps, off0, len;      /* sps = int(round(fs/baud)); OFDM: off0 = sps/4 (cyclic prefix), len = sps - off0   */
  int32_t nf, bits_per_sym;          /* correlators per symbol (1; OFDM: bins; FSK: 2) and decided bits per symbol        */
  int32_t uart, prefilter;           /* FSK: UART deframe (B.2) / Butterworth-4 zero-phase band-pass before Goertzel (B.1) */
  int32_t bp_w, bp_pad;              /* pre-filter warm-up (pole decay to 1e-12) and filtfilt padlen                       */
  double  bp_b[9], bp_a[9], bp_zi[8];
} fb_v1_params;
uint64_t fb_v1_out_bound(const fb_v1_params* p, uint64_t n_samples);
/* table: HOST, nf*len complex float64 (re, im interleaved).  out slots must be 4-byte aligned.  No sync search in v1:
 * out_len[r] = floor(bits/8) (or the UART-deframed byte count).                                                         */
int fb_v1_demod_batch(fb_handle* h, const fb_v1_params* p, const double* table, int n_rec, const void* samples,
                      const uint64_t* offsets, int dtype, int flags, uint8_t* out, const uint64_t* out_offsets,
                      uint64_t* out_len, int32_t* status);

/* ---- fec.py decode (standalone ops: the reference never wires FEC into decode_from_buffer) ------------------
 * Blocks are independent; block i is in[in_offsets[i] .. in_offsets[i+1]), its slot out[out_offsets[i] ..
 * out_offsets[i+1]) (size it with fb_*_out_bound).  in_offsets / out_offsets are HOST arrays of n_blk+1 entries;
 * FB_SAMPLES_ON_DEVICE says `in` is a device pointer, FB_OUT_ON_DEVICE says out / out_len / crc_ok are.          */
uint64_t fb_rs_out_bound(uint64_t n);          /* ReedSolomonFEC.decode output length for n input bytes (fec.py:34-69) */
uint64_t fb_viterbi_out_bound(uint64_t n);     /* ViterbiDecoder.decode output length (fec.py:126-155)                */
/* Replaces ReedSolomonFEC.decode (fec.py:34-69).  crc_ok[i] = 1 when the trailing CRC32 matches the decoded block
 * (the reference only prints a warning when it does not), 0 otherwise, -1 when the out slot is too small.          */
int fb_rs_decode_batch(fb_handle* h, int n_blk, const uint8_t* in, const uint64_t* in_offsets, uint8_t* out,
                       const uint64_t* out_offsets, uint64_t* out_len, int32_t* crc_ok, int flags);
/* Same decode for blocks given as (start, length) spans of a DEVICE buffer (FB_SAMPLES_ON_DEVICE required; in_start /
 * in_len / out_offsets are HOST arrays): the payloads of frames found by fb_parse_frames_batch are decoded where the
 * demodulator left them, without a gather copy. */
int fb_rs_decode_spans(fb_handle* h, int n_blk, const uint8_t* in, const uint64_t* in_start, const uint64_t* in_len,
                       uint8_t* out, const uint64_t* out_offsets, uint64_t* out_len, int32_t* crc_ok, int flags);
/* Replaces ViterbiDecoder.decode (fec.py:126-155). */
int fb_viterbi_decode_batch(fb_handle* h, int n_blk, const uint8_t* in, const uint64_t* in_offsets, uint8_t* out,
                            const uint64_t* out_offsets, uint64_t* out_len, int flags);
/* zlib.crc32 / binascii.crc32 of n_blk byte ranges (decoder.py:100,194; fec.py:65). */
int fb_crc32_batch(fb_handle* h, int n_blk, const uint8_t* data, const uint64_t* offsets, uint32_t* crc, int flags);

/* ---- FBPC frame parser: replaces decoder.parse_fbp_stream_enhanced (decoder.py:142-208) -------------------------
 * Frame layout (encoder.py:94-114): "FBPC" | u8 name_len | name | <6 x u32 LE: part, total, file_size, file_crc,
 * data_len, crc32(data)> | data.  Offsets in fb_frame are relative to the recording's raw stream.                   */
typedef struct fb_frame {
  uint64_t offset, name_off, payload_off;
  uint32_t name_len, part, total, file_size, file_crc, data_len, payload_crc, reserved;
} fb_frame;
/* raw stream of recording r: raw[raw_offsets[r] ..) with raw_len[r] valid bytes (raw_offsets HOST, n_rec+1 entries;
 * raw_len lives where `raw` lives: device when FB_SAMPLES_ON_DEVICE -- e.g. fb_psk_demod_batch's out / out_len).
 * frames: n_rec * max_frames records, CRC-valid frames of recording r in stream order at frames[r*max_frames ..];
 * n_frames[r] = count of CRC-valid frames, every "FBPC" occurrence considered (no candidate limit, like the reference);
 * it may exceed max_frames: only the first max_frames are stored -- call again with a larger table;
 * payload_bytes[r] = sum of their data_len.                                                                        */
int fb_parse_frames_batch(fb_handle* h, int n_rec, const uint8_t* raw, const uint64_t* raw_offsets,
                          const uint64_t* raw_len, int max_frames, fb_frame* frames, int32_t* n_frames,
                          uint64_t* payload_bytes, int flags);

/* ---- WAV ingest on the device: replaces the host steps of decode_wav_file (decoder.py:381-387) ---------------------------
 * in: n_frames interleaved frames of n_channels samples (FB_S16 = WAV PCM16, scaled by 1/32768 as soundfile does; FB_F32;
 * FB_F64); channel 0 is kept (data[:, 0], decoder.py:382).  out: n_out float64 samples = scipy.signal.resample(channel0,
 * n_out) (FFT method, decoder.py:385-387: n_out = int(round(n_frames * 96000 / sr)); the transforms are the library's own
 * float64 FFT, csrc/fft.cu, any length); n_out == n_frames just converts.
 * FB_SAMPLES_ON_DEVICE / FB_OUT_ON_DEVICE say where in / out live; the result can be fed to the *_demod_batch calls as
 * FB_F64 without leaving the device.                                                                                  */
int fb_ingest_resample(fb_handle* h, const void* in, uint64_t n_frames, int n_channels, int dtype, uint64_t n_out,
                       double* out, int flags);

/* ---- batch modulators (TX side): replace bpsk_modulate (modem.py:28-65), qpsk_modulate (modem.py:138-186;
 * also psk8_modulate / ofdm_modulate_simple / apsk16_modulate, modem.py:344,371,379) and fsk_modulate (modem.py:270-295;
 * also fsk_high_speed_modulate, modem.py:351) for n_rec payloads in one call.  The host wrapper builds the sps-entry
 * tables with the reference's own numpy expressions: PSK base[j] = 2 pi carrier t_j and env[j] (10 % linear ramps,
 * modem.py:57-61,179-183); CPFSK base[j] = t_j, env = NULL.  Symbols per payload: DBPSK 80 + 8 n, DQPSK 40 + 4 n,
 * CPFSK 8 (4 + n) (preambles included).  data: payload bytes back to back (CSR data_offsets, n_rec + 1; host, or device
 * with FB_SAMPLES_ON_DEVICE); out: float32 samples, payload r at out_offsets[r] (slot >= fb_mod_out_samples; host, or
 * device with FB_OUT_ON_DEVICE).  Phases are accumulated sequentially in float64 exactly as the reference does.       */
enum { FB_MOD_DBPSK = 0, FB_MOD_DQPSK = 1, FB_MOD_CPFSK = 2 };
typedef struct fb_mod_params {
  int32_t kind, sps;                 /* FB_MOD_*; int(fs / baud) for PSK, int(round(fs * (1 / baud))) for CPFSK          */
  double  inc[4];                    /* PSK: phase change per code 2 b0 + b1 (DBPSK: [0, pi]); CPFSK: [space, mark] phase
                                        advance per bit, 2 pi f (spb / fs)                                            */
  double  wfreq[2];                  /* CPFSK: 2 pi f for [space, mark]                                               */
  float   gain, pad;                 /* CPFSK: 0.9f (float32 multiply, modem.py:295)                                  */
} fb_mod_params;
uint64_t fb_mod_out_samples(const fb_mod_params* p, uint64_t n_bytes);
int fb_modulate_batch(fb_handle* h, const fb_mod_params* p, const double* base, const double* env, int n_rec,
                      const uint8_t* data, const uint64_t* data_offsets, float* out, const uint64_t* out_offsets, int flags);

#ifdef __cplusplus
}
#endif
#endif /* FBDSP_H */
