// Shared device/host helpers for libqbn (sm_100a only).
//
// Everything in csrc/ is hand-written for B200; there is no CPU fallback: every
// extern "C" entry point launches a CUDA kernel or returns a negative status.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/qbn.h"

#define QBN_DEVINL __device__ __forceinline__

// ---------------------------------------------------------------------------------------------
// error plumbing (never throw across the C ABI)
// ---------------------------------------------------------------------------------------------
void qbn_set_error(const char* fmt, ...);

#define QBN_CHECK_ARG(cond, msg)                                                     \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      qbn_set_error("%s: invalid argument: %s", __func__, msg);                      \
      return QBN_ERR_INVALID_ARG;                                                    \
    }                                                                                \
  } while (0)

#define QBN_CHECK_LAUNCH()                                                           \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) {                                                        \
      qbn_set_error("%s: CUDA error: %s", __func__, cudaGetErrorString(e__));        \
      return QBN_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

#define QBN_CUDA(call)                                                               \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      qbn_set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e__));  \
      return QBN_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

int qbn_sm_count();  // cached cudaDevAttrMultiProcessorCount of the current device
// device scalar added to every Philox draw index (qbn_set_sample_base), or nullptr
const uint32_t* qbn_sample_base_ptr();

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid for a grid-stride elementwise kernel: a whole number of waves of 148-SM multiples
static inline int qbn_grid_for(int64_t work_items, int threads, int ctas_per_sm = 8) {
  int64_t want = ceil_div64(work_items, threads);
  int64_t cap = (int64_t)qbn_sm_count() * ctas_per_sm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG.  Stream = (seed, stream_a, stream_b); counter = element_index / 4.
//   stream_a : layer / purpose id     stream_b : GLOBAL Monte-Carlo sample index (or step index)
// so the draw for (sample s, layer l, element i) does not depend on how samples are sharded
// over GPUs.  oracle/philox.py restates this bit for bit.
// ---------------------------------------------------------------------------------------------
struct Philox4 {
  uint32_t x, y, z, w;
};

QBN_DEVINL Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                 uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

QBN_DEVINL Philox4 philox_at(uint64_t seed, uint32_t stream_a, uint32_t stream_b, uint64_t ctr) {
  return philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), stream_a, stream_b, (uint32_t)seed,
                       (uint32_t)(seed >> 32));
}

// 24-bit uniform strictly inside (0,1): (x>>8)*2^-24 + 2^-25
QBN_DEVINL float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f + 2.98023223876953125e-08f; }

// Box-Muller on a pair of words -> two N(0,1) draws
QBN_DEVINL void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  float u1 = u01(a), u2 = u01(b);
  float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

// four normals for counter `ctr` (elements 4*ctr .. 4*ctr+3 of the stream)
QBN_DEVINL void philox_normal4(uint64_t seed, uint32_t sa, uint32_t sb, uint64_t ctr, float z[4]) {
  Philox4 p = philox_at(seed, sa, sb, ctr);
  box_muller(p.x, p.y, z[0], z[1]);
  box_muller(p.z, p.w, z[2], z[3]);
}

// one normal for stream element `idx` (used where a thread owns scattered elements)
QBN_DEVINL float philox_normal1(uint64_t seed, uint32_t sa, uint32_t sb, uint64_t idx) {
  float z[4];
  philox_normal4(seed, sa, sb, idx >> 2, z);
  return z[idx & 3];
}

QBN_DEVINL float philox_uniform1(uint64_t seed, uint32_t sa, uint32_t sb, uint64_t idx) {
  Philox4 p = philox_at(seed, sa, sb, idx >> 2);
  uint32_t v = (idx & 3) == 0 ? p.x : (idx & 3) == 1 ? p.y : (idx & 3) == 2 ? p.z : p.w;
  return u01(v);
}

// torch.nn.functional.softplus (beta=1, threshold=20): reference linear.py:25,35,43
QBN_DEVINL float softplus_f(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
QBN_DEVINL float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// round-to-nearest (ties away) fp32 -> tf32, kept in an fp32 container (cvt.rna.tf32.f32).  The
// tensor core ignores the 13 low mantissa bits of a kind::tf32 operand, i.e. truncates; rounding
// first halves the error and removes its bias, which otherwise compounds through 21 layers.
QBN_DEVINL uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
QBN_DEVINL float tf32_round(float x) { return __uint_as_float(tf32_rna(x)); }

QBN_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
QBN_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
QBN_DEVINL float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
