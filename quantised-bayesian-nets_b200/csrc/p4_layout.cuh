// Layouts shared by the planar-C4 convolution kernel (umma_conv_p4.cu), the blocked weight sampler and the
// layout helpers (core.cu).
//
// Activations, "planar C4":  [C/4 chunk planes][rows][4 floats], rows = n_samples * B * Hp * Wp pixels of the
// zero-bordered maps + a zero tail.  Zeros are shared between neighbours: ph zero rows on TOP of every map, pw zero
// columns on the LEFT of every row (Hp = H + ph, Wp = W + pw), tail = ph * Wp + pw pixels after the last map.  One 16-byte K-chunk of a pixel is contiguous and consecutive pixels of one chunk plane
// are contiguous, which is exactly one column of UMMA's K-major no-swizzle operand layout: a tile's chunk
// plane is ONE bulk (TMA-engine) copy, and an epilogue thread (= one pixel) stores 16 bytes next to its
// neighbour's (fully coalesced).
//
// Sampled weights, "blocked":  [sample][channel block cb][tap t][chunk j][n_pad rows][4 floats] — every
// (cb, tap) block is already the smem image of the B operand, so it is one bulk copy too.
#pragma once

// channels per block: whole C up to 32, else the first of {32, 16, 24, 8} dividing C — small enough that two
// CTAs (two tcgen05 issuers) share an SM even when the whole sampled tensor of a 48-channel layer is resident.
// A stride-2 3x3 conv stages four phase strips per block, so its blocks are at most 24 channels.
static __host__ __device__ inline int qbn_p4_block_channels(int C, int stride = 1, int taps = 9) {
  if (stride == 2 && taps > 1) {
    const int c2[3] = {24, 16, 8};
    for (int i = 0; i < 3; ++i)
      if (C % c2[i] == 0) return c2[i];
    return 0;
  }
  if (C <= 32) return C;
  const int cand[4] = {32, 16, 24, 8};
  for (int i = 0; i < 4; ++i)
    if (C % cand[i] == 0) return cand[i];
  return 0;
}
static __host__ __device__ inline int qbn_p4_n_pad(int N) { return (N + 15) / 16 * 16; }

// ---- int8 twin, "planar C16": the same byte layout with 16 s8 channels per 16-byte chunk; maps hold (q - zero_point), so the
// shared zero border is the padding of a quint8 convolution.  Channel counts are zero-padded to a multiple of 32 (one kind::i8
// MMA consumes K = 32 = two chunks).  Blocked weights carry one extra output row N of ones over the real input channels: its
// accumulator column is sum_k (x - z_x), the z_w correction of sum (x - z_x)(w - z_w).
static __host__ __device__ inline int qbn_p16_block_channels(int C, int stride = 1, int taps = 9) {
  if (C % 32 != 0) return 0;
  const int cap = (stride == 2 && taps > 1) ? 96 : 192;      // a stride-2 3x3 conv stages four phase strips per block
  if (C <= cap) return C;
  const int cand[4] = {192, 96, 64, 32};
  for (int i = 0; i < 4; ++i)
    if (cand[i] <= cap && C % cand[i] == 0) return cand[i];
  return 0;
}
static __host__ __device__ inline int qbn_p16_n_pad(int N) { return (N + 1 + 15) / 16 * 16; }
