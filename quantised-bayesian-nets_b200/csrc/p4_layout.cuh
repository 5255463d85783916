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
#include <stdint.h>

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

// geometry of one blocked weight tensor (shared by the blocker, the samplers and the LRT weight preparation)
struct P4Block { int N, C, taps, CB, cbc, n_pad, K; int64_t total4; };
static inline bool p4_block_geom(int N, int C, int taps, int stride, P4Block& g, int cb_override = 0) {
  g.N = N; g.C = C; g.taps = taps; g.CB = cb_override > 0 ? cb_override : qbn_p4_block_channels(C, stride, taps);
  if (g.CB > 0 && (C % g.CB != 0 || g.CB % 8 != 0)) return false;
  if (C % 8 != 0 || g.CB == 0 || N > 256 || N <= 0 || taps <= 0) return false;
  g.cbc = g.CB / 4; g.n_pad = qbn_p4_n_pad(N); g.K = taps * C;
  g.total4 = (int64_t)(C / g.CB) * taps * g.cbc * g.n_pad;
  return true;
}
// blocked float4 index -> canonical OHWI element index (or -1 for the zero rows n >= N).  32-bit arithmetic: a layer's
// blocked tensor has < 2^31 chunks (64-bit divisions were most of the sampler's instructions)
static __host__ __device__ __forceinline__ int64_t p4_canonical(const P4Block& g, int64_t i64) {
  const uint32_t i = (uint32_t)i64;
  const uint32_t t1 = i / (uint32_t)g.n_pad;
  const uint32_t n = i - t1 * (uint32_t)g.n_pad;
  const uint32_t t2 = t1 / (uint32_t)g.cbc;
  const uint32_t j = t1 - t2 * (uint32_t)g.cbc;
  const uint32_t cb = t2 / (uint32_t)g.taps;
  const uint32_t t = t2 - cb * (uint32_t)g.taps;
  if ((int)n >= g.N) return -1;
  return (int64_t)n * g.K + (int64_t)t * g.C + cb * g.CB + 4 * j;
}


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
