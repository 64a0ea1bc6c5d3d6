// Declarations shared by the two DPSK interior kernels (psk_v2.cu: fp32 CUDA-core path, psk_mma.cu: tensor-core path).
#pragma once
#include "common.cuh"

// One interior tile, resolved once per call by psk_tiles_kernel so that a CTA starts with ONE coalesced 32-byte load
// instead of a dependent search (tile_first -> tile_first -> plans).
struct __align__(16) PskTile {
  uint64_t off, n, word_off;    // the recording: first sample (elements), samples, first word of its bit stream
  int32_t d0, d1;               // differential symbols [d0, d1) of this tile; symbols d0 .. d1
};

// psk_mma.cu
bool fb_psk_mma_usable(fb_handle* h, const fb_psk_design& d, const float* taps);
int fb_psk_mma_tile_syms();
int fb_psk_mma_launch(fb_handle* h, const fb_psk_design& d, const float* taps, const void* d_samples, uint64_t total_samples, int dtype,
                      const PskTile* d_tiles, uint32_t n_tiles, uint32_t* d_bits, uint32_t* d_redo, int edge_ctas);
void fb_psk_mma_release(fb_handle* h);
