// Zero-copy-im2col tcgen05 convolution on the planar-C4 zero-bordered layout (see p4_layout.cuh).
//
//   out[q][n] = act( scale[n] * sum_{tap, c} x[strip(tap)][q + shift(tap)][c] * w[z][n][tap][c] + shift[n] + residual[q][n] )
//
// q = flat pixel index of the zero-bordered map [n_samples*B][Hp][Wp]; the zero border IS the padding.
//  * stride 1, odd RxS "same" conv: one input strip, shift(r,s) = (r-ph)*Wp + (s-pw).
//  * stride 2 (3x3 pad 1, or 1x1 pad 0): the input arrives PHASE-SPLIT — four half-resolution zero-bordered
//    maps, phase (h&1, w&1) of the full-resolution one, each with the OUTPUT's geometry — so tap (r,s) reads
//    strip ((r-1)&1, (s-1)&1) at shift floor((r-1)/2)*Wp + floor((s-1)/2): again a pure row shift.
//    The producing layer writes that layout directly from its epilogue (QBN_FLAG_OUT_PHASE_SPLIT).
// Every tap is therefore a UMMA descriptor whose start address is shifted by a whole number of 16-byte rows
// inside ONE smem image of the rows [q0 - d_before, q0 + 128 + d_after) (K-major, no swizzle: a pixel's
// 16-byte K-chunk is one row of an 8-row core matrix).  Nothing is gathered and nothing is re-read.
//
// Persistent, warp-specialised, and no thread ever touches operand data:
//   warp 5, one lane : bulk copies (cp.async.bulk -> mbarrier complete_tx): one per chunk plane of a tile
//                      (contiguous in the planar layout), one per weight block (pre-blocked by the sampler)
//   warp 4           : tcgen05.mma kind::tf32 issue, tcgen05.commit releases the smem slots / signals the epilogue
//   warps 0-3        : epilogue, one TMEM lane = one pixel per thread: affine (BN/bias) + residual + ReLU + RNA
//                      rounding, 16-byte stores that are contiguous across the warp
// TMEM holds ACC accumulator tiles so the MMAs of tile i+1.. overlap the epilogue of tile i.
#include <stdlib.h>
#include <string.h>
#include "p4_layout.cuh"
#include "umma_common.cuh"

namespace {

// kernel-tuning knobs (QBN_P4_OCC / _TG / _SA / _ACC / _VERBOSE / _PROF) exist only in -DQBN_TUNING builds (QBN_TUNING=1 python -m
// ..._build; scripts/p4_sweep.sh): the product library reads nothing from the environment
#ifdef QBN_TUNING
static inline const char* tune_env(const char* name) { return getenv(name); }
#else
static inline const char* tune_env(const char*) { return nullptr; }
#endif

constexpr int TM = 128;
constexpr int P4_THREADS = 192;
constexpr int MAX_TAPS = 25;

struct P4Params {
  int Hp, Wp, bh, bw, B;
  int N, n_pad, Qs, tiles_per_sample, total_tiles;
  int cbc, n_cb, n_strips, taps, nk;
  int RA, RA_p, d_before;
  int SA, SB, TG, b_res, ACC, tmem_cols, flags, w_shared;
  uint32_t a_bytes, bt_bytes, b_slot_bytes;
  uint32_t idesc, mg_plane, mg_wp;
  long long x_plane, strip_rows, out_plane, res_plane, w_sample_floats;
  int out_split, Hp2, Wp2;
  int tile0;                      // first tile of the launch's window (qbn_p4_set_window; 0 = the whole tile space)
  int tile_rr;                    // tiles dealt round-robin over the CTAs (streamed weights) instead of in contiguous ranges
  int stacked, n_chunks, cps;     // stacked: the N columns are `n_chunks / cps` samples x cps channel chunks (shared input)
  long long q2_total;
  int tap_off[MAX_TAPS];          // 16-byte units inside an A slot: strip * cbc * RA_p + d_before + shift
  const float* x; const float* w; const float* scale; const float* shift; const float* residual; float* out;
  // fused 1x1 stride-2 shortcut (models_bbb.py:163-166): a second input (phase (0,0) of the block input) accumulated into the
  // same tile by n_cb2 extra channel blocks of one tap; its weight blocks follow the main ones in every sample's weight tensor
  const float* x2; long long x2_plane; int n_cb2, cbc2, nk2; uint32_t bt2_bytes;
  const float* out_mask; float out_mask_mult;     // A8: MC-Dropout of the OUTPUT, mask [n_img][N] (dropout.py:35-39)
  int x_shared;                                   // every sample reads the SAME input maps (first layer of the int8 network)
  // ---- int8 (KIND_I8): operands are (q - zero_point) as s8, 16 channels per 16-byte chunk; FBGEMM requantisation epilogue
  int z_w, z_out, q_lo, q_hi, n_out_chunks;       // q = clamp(rint((acc - z_w*rowsum + bias/atw) * mult) + z_out, q_lo, q_hi)
  float atw, mult;
  int has_add, z_res, z_add, add_lo, add_hi, z_fin;   // quantized::add[_relu] with the residual, then stored as q - z_fin
  float s_a, p_a, s_b, p_b, inv_s_add;
  int32_t* acc_dump;                              // optional [rows][N] int32 accumulators (tests)
  // ---- LRT (KIND_LRT): two contractions per tile, mean = x * mu and var = x^2 * sigma^2 (linear.py:32-35, conv.py:24-27), side by
  // side in TMEM.  The second operand pair enters as n_cb more channel blocks: x_sq planes and the sigma^2 blocks that follow the
  // mu blocks in the weight tensor.  lrt_mode 0: out = mean + sqrt(1e-8 + var) * eps + bias, out2 = sqrt(1e-8 + var).
  // lrt_mode 1 (input gradient, SURVEY 8a A3: the operands are g and dv, the weights flipped/transposed):
  // out = acc0 + 2 * xin .* acc1 (+ residual).
  int dual, acc_cols, lrt_mode;
  const float* x_sq; const float* eps; const float* xin; float* out2;
  long long aux_plane;                            // rows per chunk plane of eps / xin / out2 (the output's geometry)
  unsigned long long seed; uint32_t stream_a, stream_b; const uint32_t* sbase;      // sbase: device-side draw offset (qbn_set_sample_base)
  // module boundary (ops.LRTFunction): out / out2 / eps / xin are dense NHWC [B][H_out][W_out][N] tensors; the padded pixel
  // (b, hh, ww) of this launch's geometry is the NHWC pixel (b, (hh - bh) * oh_mul + oh_add, (ww - bw) * ow_mul + ow_add)
  // (mul 2 / add phase: the four phase launches of a stride-2 layer's input gradient)
  int nhwc, H_out, W_out, oh_mul, oh_add, ow_mul, ow_add;
  int phase_z;                                    // the launch's 'samples' z = 0..3 are the phases (z >> 1, z & 1) of a stride-2 input gradient
};
enum { KIND_TF32 = 0, KIND_I8 = 1, KIND_LRT = 2 };

// cycle accounting of CTA 0 (QBN_P4_PROF=1): [role*8 + category], summed over its tiles
__device__ unsigned long long g_p4_prof[32];
// (accumulated in registers, flushed once at the end: a global read-modify-write per sample would dominate)
#define PROF_BEGIN() long long t_prof = PROF ? clock64() : 0
#define PROF_ADD(slot)                                   \
  do {                                                   \
    if (PROF) {                                          \
      const long long t_now = clock64();                 \
      prof_acc[(slot) & 7] += (unsigned long long)(t_now - t_prof); \
      t_prof = t_now;                                    \
    }                                                    \
  } while (0)

QBN_DEVINL void warp_wait(uint64_t* bar, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait(smem_u32(bar), parity);
  __syncwarp();
}
QBN_DEVINL float4 ld_nc4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
// n / d for n*1 < 2^32 with m = floor(2^32 / d): estimate is exact or one too small
QBN_DEVINL void divmod(uint32_t n, uint32_t d, uint32_t m, uint32_t& q, uint32_t& r) {
  q = __umulhi(n, m);
  r = n - q * d;
  if (r >= d) { ++q; r -= d; }
}


// ---- int8 requantisation of one column group (FBGEMM ReQuantizeOutput + ATen quantized::add), CNT = 8 or 16 valid channels.
// Same integers as the reference sequence rint -> + zero point -> clamp (-> dequantise -> add -> requantise), with the float<->int
// conversions (half-rate instructions; the int8 epilogue is issue-bound, profiles/r02_i8_p4_kernels_ncu_full.csv) replaced by
// exact float arithmetic: clamping to INTEGER bounds commutes with rint, and for |y| < 2^22  y + 1.5*2^23  rounds y to the nearest
// even integer in one FADD (the same RN rounding as cvt.rni) and leaves it, two's complement, in the low mantissa bits.
// out[j]: a 32-bit pattern whose LOW BYTE is the stored value (q - zero point); the packer takes the low bytes.
constexpr float I8_MAGIC = 12582912.0f;      // 1.5 * 2^23
template <int CNT, bool ADD>
QBN_DEVINL void i8_requant_group(const uint32_t (&v)[16], int corr, const float* s_shift16, const uint4& rres, const P4Params& p, uint32_t (&out)[16]) {
  float sh[16];
#pragma unroll
  for (int k = 0; k < CNT / 4; ++k) {
    const float4 t = *reinterpret_cast<const float4*>(s_shift16 + 4 * k);
    sh[4 * k] = t.x; sh[4 * k + 1] = t.y; sh[4 * k + 2] = t.z; sh[4 * k + 3] = t.w;
  }
  if constexpr (!ADD) {
    const float lo = (float)(p.q_lo - p.z_out), hi = (float)(p.q_hi - p.z_out);
#pragma unroll
    for (int j = 0; j < CNT; ++j) {
      const float y = __fmul_rn(__fadd_rn((float)((int)v[j] - corr), sh[j]), p.mult);
      out[j] = __float_as_uint(__fadd_rn(fminf(fmaxf(y, lo), hi), I8_MAGIC));
    }
  } else {
    const float lo1 = (float)(p.q_lo - p.z_out), hi1 = (float)(p.q_hi - p.z_out);
    const float unbias1 = I8_MAGIC - (float)p.z_out;                 // (yc + M) - (M - z_out) = rint(yc) + z_out = the conv's quint8 output
    const float lo2 = (float)(p.add_lo - p.z_add), hi2 = (float)(p.add_hi - p.z_add);
    // residual bytes hold (q - z_res) as s8: (byte ^ 0x80) = q - z_res + 128 in [0, 255], placed in the low mantissa byte of 1.5 * 2^23
    const uint32_t rw[4] = {rres.x ^ 0x80808080u, rres.y ^ 0x80808080u, rres.z ^ 0x80808080u, rres.w ^ 0x80808080u};
    const float unbias_r = I8_MAGIC + 128.0f - (float)p.z_res;
#pragma unroll
    for (int j = 0; j < CNT; ++j) {
      const float y = __fmul_rn(__fadd_rn((float)((int)v[j] - corr), sh[j]), p.mult);
      const float qo = __fadd_rn(__fadd_rn(fminf(fmaxf(y, lo1), hi1), I8_MAGIC), -unbias1);
      const float rb = __fadd_rn(__uint_as_float(__byte_perm(rw[j >> 2], 0x4B400000u, 0x7650u + (uint32_t)(j & 3))), -unbias_r);
      // quantized::add[_relu] (vector body of ATen's kernel: dequantise with one FMA per operand)
      const float da = __fmaf_rn(p.s_a, qo, p.p_a);
      const float db = __fmaf_rn(p.s_b, rb, p.p_b);
      const float z = __fmul_rn(__fadd_rn(da, db), p.inv_s_add);
      out[j] = __float_as_uint(__fadd_rn(fminf(fmaxf(z, lo2), hi2), I8_MAGIC));
    }
  }
#pragma unroll
  for (int j = CNT; j < 16; ++j) out[j] = 0u;
}

template <int DBG_MODE, bool STACKED, bool MASKED, int KIND = KIND_TF32>
__global__ void __launch_bounds__(P4_THREADS, KIND == KIND_I8 ? 4 : (KIND == KIND_LRT ? 2 : 3)) umma_conv_p4_kernel(const __grid_constant__ P4Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr bool PROF = DBG_MODE == 1;                    // cycle accounting (QBN_P4_PROF, diagnostics only)
  constexpr bool I8 = KIND == KIND_I8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + (size_t)p.SA * p.a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + (size_t)p.SB * p.b_slot_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + p.SA;
  uint64_t* b_full = a_empty + p.SA;
  uint64_t* b_empty = b_full + p.SB;
  uint64_t* acc_full = b_empty + p.SB;
  uint64_t* acc_empty = acc_full + p.ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + p.ACC);
  float* s_scale = reinterpret_cast<float*>(tmem_slot + 4);
  float* s_shift = s_scale + 256;
  // operand offsets of every MMA of one channel block, in issue order (tap-major, then K step), 16-byte units: .x from the
  // A slot's start, .y from the weight slot's start.  The issuing thread reads one entry per MMA instead of re-deriving it.
  uint2* s_ops = reinterpret_cast<uint2*>(s_shift + 256);
  {
    const uint32_t a_kk = (uint32_t)(2 * p.RA_p), b_kk = (uint32_t)(2 * p.n_pad), bt16_ = p.bt_bytes >> 4;
    for (int e = tid; e < p.taps * p.nk; e += P4_THREADS) {
      const int t = e / p.nk, jj = e - t * p.nk;
      s_ops[e] = make_uint2((uint32_t)p.tap_off[t] + (uint32_t)jj * a_kk, (uint32_t)(p.b_res ? t : t % p.TG) * bt16_ + (uint32_t)jj * b_kk);
    }
  }
  if (tid == 0) {
    for (int i = 0; i < p.SA; ++i) { mbar_init(smem_u32(&a_full[i]), 1); mbar_init(smem_u32(&a_empty[i]), 1); }
    for (int i = 0; i < p.SB; ++i) { mbar_init(smem_u32(&b_full[i]), 1); mbar_init(smem_u32(&b_empty[i]), 1); }
    for (int i = 0; i < p.ACC; ++i) { mbar_init(smem_u32(&acc_full[i]), 1); mbar_init(smem_u32(&acc_empty[i]), 4); }
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  // Programmatic dependent launch (qbn_set_pdl): everything above — operand list, barrier init, TMEM allocation — overlaps the
  // tail of the previous kernel in the stream; nothing below touches global memory before that kernel has completed.  The next
  // kernel may be scheduled as soon as every CTA of this grid has passed this point (it then waits at its own
  // griddepcontrol.wait).  Without the launch attribute both instructions are no-ops.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  for (int i = tid; i < 256; i += P4_THREADS) {
    s_scale[i] = (p.scale && i < p.N) ? p.scale[i] : 1.f;
    // int8: the bias enters the accumulator domain exactly like FBGEMM's: fp32(bias) / fp32(s_x * s_w)
    s_shift[i] = (p.shift && i < p.N) ? (I8 ? __fdiv_rn(p.shift[i], p.atw) : p.shift[i]) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  unsigned long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // contiguous tile range per CTA: consecutive tiles share the sample (=> the resident weights) and their halos hit L2
  // Streamed weights (b_res == 0; the 96/192-channel layers): tiles are dealt round-robin instead, so that at any moment the whole
  // grid works on the tiles of ~2-3 samples — the weight blocks every CTA streams per tile then come from L2 (one DRAM read per
  // block instead of one per CTA group; ncu, profiles/r02_p4_dram_traffic.json: 0.6-1.0 GB of re-reads per layer-4 launch before).
  const int tile_step = p.tile_rr ? (int)gridDim.x : 1;
  const int tile_begin = p.tile0 + (p.tile_rr ? (int)blockIdx.x : (int)(((long long)p.total_tiles * blockIdx.x) / gridDim.x));
  const int tile_end = p.tile0 + (p.tile_rr ? p.total_tiles : (int)(((long long)p.total_tiles * (blockIdx.x + 1)) / gridDim.x));

  if (warp == 5) {
    // ======================================= PRODUCER (one lane) ================================
    // (issuing a tile's copies from several lanes was measured: no gain, the producer runs ahead of the MMAs anyway)
    if (lane == 0) {
      int sa = 0, sb = 0, cur_z = -1;
      uint32_t pa = 0, pb = 0;
      const uint32_t strip_bytes = (uint32_t)p.cbc * p.RA_p * 16;
      for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
        const int z = tile / p.tiles_per_sample;
        const int q0 = (tile - z * p.tiles_per_sample) * TM;
        const float* ws = p.w + (p.w_shared ? 0 : (size_t)z * p.w_sample_floats);
        PROF_BEGIN();
        if (p.b_res && z != cur_z) {
          mbar_wait(smem_u32(&b_empty[0]), pb ^ 1);          // MMAs of the previous sample have retired
          const uint32_t total = p.bt_bytes * (uint32_t)(p.n_cb * (p.dual ? 2 : 1) * p.taps) + p.bt2_bytes * (uint32_t)p.n_cb2;
          mbar_arrive_expect_tx(smem_u32(&b_full[0]), total);
          for (uint32_t off = 0; off < total; off += 32768u)
            bulk_load_g2s(smem_u32(b_ring) + off, reinterpret_cast<const uint8_t*>(ws) + off, min(32768u, total - off), smem_u32(&b_full[0]));
          pb ^= 1;
          cur_z = z;
        }
        // rows [g0, g0 + RA) of every strip, clamped to the tensor (rows outside feed border outputs only)
        const long long g0 = (long long)(p.x_shared ? 0 : z) * p.Qs + q0 - p.d_before;
        const long long lo = g0 < 0 ? 0 : g0;
        const long long lim = p.x_plane - (long long)(p.n_strips - 1) * p.strip_rows;      // rows readable from a strip's start (incl. the zero tail)
        const long long hi = (g0 + p.RA > lim) ? lim : g0 + p.RA;
        const uint32_t row_bytes = (uint32_t)(hi - lo) * 16;
        const uint32_t dst_off = (uint32_t)(lo - g0) * 16;
        const int n_cb_all = p.dual ? 2 * p.n_cb : p.n_cb;     // LRT: the x^2 blocks follow the x blocks
        for (int cb = 0; cb < n_cb_all; ++cb) {
          const float* xsrc = cb >= p.n_cb ? p.x_sq : p.x;
          const int cbl = cb >= p.n_cb ? cb - p.n_cb : cb;
          PROF_ADD(16);
          mbar_wait(smem_u32(&a_empty[sa]), pa ^ 1);
          PROF_ADD(17);
          const uint32_t bar = smem_u32(&a_full[sa]);
          mbar_arrive_expect_tx(bar, row_bytes * (uint32_t)(p.cbc * p.n_strips));
          const uint32_t slot = smem_u32(a_ring + (size_t)sa * p.a_bytes) + dst_off;
          for (int s2 = 0; s2 < p.n_strips; ++s2) {
            const float* src = xsrc + ((size_t)(cbl * p.cbc) * p.x_plane + (size_t)s2 * p.strip_rows + lo) * 4;
            for (int j = 0; j < p.cbc; ++j)
              bulk_load_g2s(slot + (uint32_t)s2 * strip_bytes + (uint32_t)(j * p.RA_p) * 16, src + (size_t)j * p.x_plane * 4, row_bytes, bar);
          }
          if (++sa == p.SA) { sa = 0; pa ^= 1; }
          PROF_ADD(18);
          if (!p.b_res) {
            const uint8_t* wb = reinterpret_cast<const uint8_t*>(ws) + (size_t)cb * p.taps * p.bt_bytes;
            for (int t0 = 0; t0 < p.taps; t0 += p.TG) {
              mbar_wait(smem_u32(&b_empty[sb]), pb ^ 1);
              PROF_ADD(19);
              const uint32_t bytes = p.bt_bytes * (uint32_t)min(p.TG, p.taps - t0);
              mbar_arrive_expect_tx(smem_u32(&b_full[sb]), bytes);
              bulk_load_g2s(smem_u32(b_ring + (size_t)sb * p.b_slot_bytes), wb + (size_t)t0 * p.bt_bytes, bytes, smem_u32(&b_full[sb]));
              if (++sb == p.SB) { sb = 0; pb ^= 1; }
              PROF_ADD(20);
            }
          }
        }
        // ---- fused shortcut: rows [q0, q0+128) of the second input, no halo, one tap
        if (p.n_cb2) {
          const long long r0 = (long long)z * p.Qs + q0;
          const long long r1 = (r0 + TM > p.x2_plane) ? p.x2_plane : r0 + TM;
          const uint32_t rb2 = (uint32_t)(r1 - r0) * 16;
          for (int cb = 0; cb < p.n_cb2; ++cb) {
            mbar_wait(smem_u32(&a_empty[sa]), pa ^ 1);
            const uint32_t bar = smem_u32(&a_full[sa]);
            mbar_arrive_expect_tx(bar, rb2 * (uint32_t)p.cbc2);
            const uint32_t slot = smem_u32(a_ring + (size_t)sa * p.a_bytes);
            const float* src = p.x2 + ((size_t)(cb * p.cbc2) * p.x2_plane + r0) * 4;
            for (int j = 0; j < p.cbc2; ++j)
              bulk_load_g2s(slot + (uint32_t)(j * TM) * 16, src + (size_t)j * p.x2_plane * 4, rb2, bar);
            if (++sa == p.SA) { sa = 0; pa ^= 1; }
            if (!p.b_res) {
              const uint8_t* wb = reinterpret_cast<const uint8_t*>(ws) + (size_t)p.n_cb * p.taps * p.bt_bytes + (size_t)cb * p.bt2_bytes;
              mbar_wait(smem_u32(&b_empty[sb]), pb ^ 1);
              mbar_arrive_expect_tx(smem_u32(&b_full[sb]), p.bt2_bytes);
              bulk_load_g2s(smem_u32(b_ring + (size_t)sb * p.b_slot_bytes), wb, p.bt2_bytes, smem_u32(&b_full[sb]));
              if (++sb == p.SB) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 4) {
    // ======================================= MMA ISSUER =========================================
    // ONE elected thread runs the whole issue loop (waits, tcgen05.mma, tcgen05.commit).  Inside an
    // elect.sync region ptxas keeps the descriptors in uniform registers; issuing from all lanes under a
    // leader predicate made it wrap every MMA in an ELECT / R2UR.BROADCAST waterfall loop (~100 cycles each).
    if (elect_one()) {
      int sa = 0, sb = 0, as = 0, cur_z = -1;
      uint32_t pa = 0, pb = 0, pacc = 0;
      const uint32_t lbo_a = (uint32_t)p.RA_p * 16, lbo_b = (uint32_t)p.n_pad * 16;
      const uint64_t adesc_hi = make_smem_desc(0, lbo_a, 128), bdesc_hi = make_smem_desc(0, lbo_b, 128);
      const uint32_t a_k = (2 * lbo_a) >> 4, b_k = (2 * lbo_b) >> 4;      // one K=8 step = two 16-byte chunks
      const uint32_t bt16 = p.bt_bytes >> 4;
      for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
        const int z = tile / p.tiles_per_sample;
        const bool last_of_z = (tile + tile_step >= tile_end) || ((tile + tile_step) / p.tiles_per_sample != z);
        PROF_BEGIN();
        if (p.b_res && z != cur_z) {
          mbar_wait(smem_u32(&b_full[0]), pb);
          pb ^= 1;
          cur_z = z;
        }
        PROF_ADD(8);
        mbar_wait(smem_u32(&acc_empty[as]), pacc ^ 1);      // epilogue has drained this accumulator
        PROF_ADD(9);
        uint32_t tacc = tmem_base + (uint32_t)(as * p.acc_cols);
        uint32_t accum = 0;
        const int n_cb_all = p.dual ? 2 * p.n_cb : p.n_cb;
        for (int cb = 0; cb < n_cb_all; ++cb) {
          if (p.dual && cb == p.n_cb) { tacc += (uint32_t)p.n_pad; accum = 0; }      // second accumulator: the variance contraction
          mbar_wait(smem_u32(&a_full[sa]), pa);
          PROF_ADD(10);
          tc_fence_after();
          const uint32_t a16 = smem_u32(a_ring + (size_t)sa * p.a_bytes) >> 4;
          uint32_t b16 = (smem_u32(b_ring) >> 4) + (uint32_t)(cb * p.taps) * bt16;
          // taps in groups of TG: one streamed weight slot per group (resident weights: a single group, no handshake)
#pragma unroll 1
          for (int t0 = 0; t0 < p.taps; t0 += p.TG) {
            if (!p.b_res) {
              mbar_wait(smem_u32(&b_full[sb]), pb);
              PROF_ADD(11);
              tc_fence_after();
              b16 = smem_u32(b_ring + (size_t)sb * p.b_slot_bytes) >> 4;
            }
            int e = t0 * p.nk;
            const int e1 = min(t0 + p.TG, p.taps) * p.nk;
            uint2 o = s_ops[e];
            if (!accum) {                      // the tile's first MMA overwrites the accumulator; every other one adds (constant predicate)
              const uint2 on = s_ops[min(e + 1, e1 - 1)];
              umma_mma_c<I8 ? MODE_I8 : MODE_EVAL, false>(tacc, adesc_hi | (uint64_t)((a16 + o.x) & 0x3FFF), bdesc_hi | (uint64_t)((b16 + o.y) & 0x3FFF), p.idesc);
              accum = 1;
              o = on;
              ++e;
            }
#pragma unroll 2
            for (; e < e1; ++e) {
              const uint2 on = s_ops[min(e + 1, e1 - 1)];
              umma_mma_c<I8 ? MODE_I8 : MODE_EVAL, true>(tacc, adesc_hi | (uint64_t)((a16 + o.x) & 0x3FFF), bdesc_hi | (uint64_t)((b16 + o.y) & 0x3FFF), p.idesc);
              o = on;
            }
            if (!p.b_res) {
              umma_commit(smem_u32(&b_empty[sb]));
              if (++sb == p.SB) { sb = 0; pb ^= 1; }
            }
            PROF_ADD(12);
          }
          umma_commit(smem_u32(&a_empty[sa]));
          if (++sa == p.SA) { sa = 0; pa ^= 1; }
        }
        for (int cb = 0; cb < p.n_cb2; ++cb) {              // fused shortcut blocks: one tap at row 0 of the slot
          mbar_wait(smem_u32(&a_full[sa]), pa);
          tc_fence_after();
          uint32_t ad = smem_u32(a_ring + (size_t)sa * p.a_bytes) >> 4;
          uint32_t bd;
          if (p.b_res) {
            bd = (smem_u32(b_ring) >> 4) + (uint32_t)(p.n_cb * p.taps) * bt16 + (uint32_t)cb * (p.bt2_bytes >> 4);
          } else {
            mbar_wait(smem_u32(&b_full[sb]), pb);
            tc_fence_after();
            bd = smem_u32(b_ring + (size_t)sb * p.b_slot_bytes) >> 4;
          }
          const uint64_t adesc2_hi = make_smem_desc(0, TM * 16, 128);      // chunk planes of 128 rows, no halo
#pragma unroll 1
          for (int jj = 0; jj < p.nk2; ++jj) {
            umma_mma<I8 ? MODE_I8 : MODE_EVAL>(tacc, adesc2_hi | (uint64_t)(ad & 0x3FFF), bdesc_hi | (uint64_t)(bd & 0x3FFF), p.idesc, accum);
            accum = 1;
            ad += (2 * TM * 16) >> 4; bd += b_k;
          }
          if (!p.b_res) {
            umma_commit(smem_u32(&b_empty[sb]));
            if (++sb == p.SB) { sb = 0; pb ^= 1; }
          }
          umma_commit(smem_u32(&a_empty[sa]));
          if (++sa == p.SA) { sa = 0; pa ^= 1; }
        }
        umma_commit(smem_u32(&acc_full[as]));
        if (p.b_res && last_of_z) umma_commit(smem_u32(&b_empty[0]));
        if (++as == p.ACC) { as = 0; pacc ^= 1; }
        PROF_ADD(13);
      }
      if (PROF && blockIdx.x == 0)
        for (int i = 0; i < 8; ++i) g_p4_prof[8 + i] = prof_acc[i];
    }
    __syncwarp();
  } else {
    // ======================================= EPILOGUE ===========================================
    int as = 0;
    uint32_t pacc = 0;
    const uint32_t plane = (uint32_t)(p.Hp * p.Wp);
    const int n_groups = p.n_pad >> 4;
    const int n_chunks = p.n_chunks;                         // 16-byte output chunks per row (all stacked samples)
    const int cps = p.cps;                                   // chunks per sample
    const bool relu = p.flags & QBN_FLAG_RELU, relu_pre = p.flags & QBN_FLAG_RELU_PRE, rnd = p.flags & QBN_FLAG_OUT_ROUND_TF32;
    for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
      const int z = tile / p.tiles_per_sample;
      const int q0 = (tile - z * p.tiles_per_sample) * TM;
      PROF_BEGIN();
      const int q = q0 + tid;
      const bool qv = q < p.Qs;
      uint32_t b, rem, hh, ww;
      divmod(qv ? (uint32_t)q : 0u, plane, p.mg_plane, b, rem);
      divmod(rem, (uint32_t)p.Wp, p.mg_wp, hh, ww);
      const bool interior = qv && (int)hh >= p.bh && (int)ww >= p.bw;      // zero rows on TOP of every map, zero columns on its LEFT
      const long long in_row = (long long)z * p.Qs + q;
      long long orow = in_row;
      bool store = qv;
      if (p.out_split) {
        // phase-split output for a stride-2 consumer: pixel (h, w) -> map (h&1, w&1), position (h>>1, w>>1)
        const int h = (int)hh - p.bh, w = (int)ww - p.bw;
        orow = (long long)((h & 1) * 2 + (w & 1)) * p.q2_total + ((long long)(z * p.B + (int)b) * p.Hp2 + (h >> 1) + 1) * p.Wp2 + (w >> 1) + 1;
        store = interior;                                   // its border is never written (pre-zeroed buffer)
      }
      if constexpr (KIND == KIND_LRT) {
        // ---- LRT training: acc0 = columns [0, n_pad), acc1 = [n_pad, 2 n_pad) of this accumulator slot.
        long long out_off, aux_off, aux_step, out_step, res_step;      // float offsets of chunk 0 / steps between chunks
        unsigned long long ctr0;                                       // Philox counter of chunk 0 (the existing NHWC path's: out offset / 4)
        if (p.nhwc) {
          const int oh_add = p.phase_z ? (z >> 1) : p.oh_add, ow_add = p.phase_z ? (z & 1) : p.ow_add;
          const long long pix = ((long long)b * p.H_out + ((int)hh - p.bh) * p.oh_mul + oh_add) * p.W_out + ((int)ww - p.bw) * p.ow_mul + ow_add;
          store = interior;
          out_off = aux_off = interior ? pix * p.N : 0;
          aux_step = out_step = res_step = 4;
          ctr0 = (unsigned long long)pix * (unsigned long long)n_chunks;
        } else {
          out_off = (store ? orow : 0) * 4; aux_off = in_row * 4;
          aux_step = p.aux_plane * 4; out_step = p.out_plane * 4; res_step = p.res_plane * 4;
          ctr0 = (unsigned long long)in_row * (unsigned long long)n_chunks;
        }
        float* optr = p.out + out_off;
        float* o2 = (p.out2 && store) ? p.out2 + aux_off : nullptr;
        const float* eptr = (p.eps && interior) ? p.eps + aux_off : nullptr;
        const float* xptr = (p.xin && interior) ? p.xin + aux_off : nullptr;
        const float* rptr = (p.residual && interior) ? p.residual + aux_off : nullptr;
        uint32_t v0[16], v1[16];
        // the per-element operand of the epilogue (forward: the noise tensor, input gradient: x) is fetched one column group ahead,
        // the first group BEFORE the accumulator is waited for: at a training step's launch sizes (a few tiles per CTA) an exposed
        // global-load round trip per 16-byte chunk made the epilogue, not the MMAs, the critical path
        const float* aptr = p.lrt_mode == 0 ? eptr : xptr;
        float4 acur[4], anext[4];
        auto prefetch = [&](float4* dst, int g) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int ch = g * 4 + i;
            dst[i] = (aptr && ch < n_chunks) ? ld_nc4(aptr + (long long)ch * aux_step) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        };
        prefetch(acur, 0);
        PROF_ADD(0);
        warp_wait(&acc_full[as], pacc, lane);
        PROF_ADD(1);
        tc_fence_after();
        const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(as * p.acc_cols);
        for (int g = 0; g < n_groups; ++g) {
          tmem_ld16(tlane + (uint32_t)(g * 16), v0);
          tmem_ld16(tlane + (uint32_t)(p.n_pad + g * 16), v1);
          if (g + 1 < n_groups) prefetch(anext, g + 1);
          tmem_ld_wait();
          if (g + 1 == n_groups) {                              // accumulators fully read: hand the slot back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&acc_empty[as]));
          }
          PROF_ADD(2);
          if (store) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int ch = g * 4 + i;
              if (ch < n_chunks) {
                float o[4] = {0.f, 0.f, 0.f, 0.f}, sd[4] = {0.f, 0.f, 0.f, 0.f};
                if (interior) {
                  const float av[4] = {acur[i].x, acur[i].y, acur[i].z, acur[i].w};
                  if (p.lrt_mode == 0) {
                    float e[4] = {av[0], av[1], av[2], av[3]};
                    if (!eptr)       // one Philox call = the four channels of this chunk; the backward regenerates the same draw
                      philox_normal4(p.seed, p.stream_a, p.stream_b + (p.sbase ? *p.sbase : 0u), ctr0 + (unsigned long long)ch, e);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                      sd[k] = sqrtf(__fadd_rn(1e-8f, __uint_as_float(v1[4 * i + k])));
                      o[k] = __fadd_rn(__fadd_rn(__uint_as_float(v0[4 * i + k]), __fmul_rn(sd[k], e[k])), s_shift[ch * 4 + k]);
                    }
                  } else {
                    const float4 rr = rptr ? ld_nc4(rptr + (long long)ch * res_step) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float rv[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k)      // dx = g*mu + 2x .* (dv*sigma^2) (+ the gradient that reached x through another branch)
                      o[k] = __fadd_rn(__fadd_rn(__uint_as_float(v0[4 * i + k]), __fmul_rn(__fmul_rn(2.0f, av[k]), __uint_as_float(v1[4 * i + k]))), rv[k]);
                  }
                }
                *reinterpret_cast<float4*>(optr + (long long)ch * out_step) = make_float4(o[0], o[1], o[2], o[3]);
                if (o2) *reinterpret_cast<float4*>(o2 + (long long)ch * aux_step) = make_float4(sd[0], sd[1], sd[2], sd[3]);
              }
            }
          }
          if (g + 1 < n_groups) {
#pragma unroll
            for (int i = 0; i < 4; ++i) acur[i] = anext[i];
          }
          PROF_ADD(3);
        }
      } else
      if constexpr (I8) {
        // ---- int8: one TMEM column group = 16 output channels = ONE 16-byte chunk of the planar-C16 s8 map.
        // FBGEMM's ReQuantizeOutput with a float bias (conv_q.py:120-125): every fp32 step rounded separately (no FMA).
        int8_t* optr8 = reinterpret_cast<int8_t*>(p.out) + (store ? orow : 0) * 16;
        const int8_t* rptr8 = (p.has_add && interior) ? reinterpret_cast<const int8_t*>(p.residual) + in_row * 16 : nullptr;
        const long long res_step8 = p.res_plane * 16, out_step8 = p.out_plane * 16;
        uint32_t v[16], vn[16];
        uint4 rres = make_uint4(0, 0, 0, 0), rnext = make_uint4(0, 0, 0, 0);
        auto ld_res = [&](int g) {
          uint4 r = make_uint4(0, 0, 0, 0);
          if (rptr8) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(rptr8 + (long long)g * res_step8));
          return r;
        };
        rres = ld_res(0);
        PROF_ADD(0);
        warp_wait(&acc_full[as], pacc, lane);
        PROF_ADD(1);
        tc_fence_after();
        const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(as * p.acc_cols);
        // column N of the accumulator = sum_k (x - z_x) (an all-ones weight row): the z_w correction of sum (x-z_x)(w-z_w)
        // TMEM reads are 64 bytes per cycle and SM, shared by the 4 resident CTAs (role accounting: the loads were 40 % of a tile's
        // epilogue): one column for the row sum, 8 columns for a half-filled last group
        int corr = 0;
        uint32_t rs_raw = 0;
        if (p.z_w != 0) rs_raw = tmem_ld1(tlane + (uint32_t)p.N);
        auto ld_group = [&](int g, uint32_t (&dst)[16]) {
          if (p.N - g * 16 > 8) {
            tmem_ld16(tlane + (uint32_t)(g * 16), dst);
          } else {
            tmem_ld8(tlane + (uint32_t)(g * 16), dst);      // (no register of dst may be touched before tcgen05.wait::ld)
          }
        };
        ld_group(0, v);
        const int n_out = p.n_out_chunks;
        for (int g = 0; g < n_out; ++g) {
          tmem_ld_wait();
          if (g == 0) corr = p.z_w * (int)rs_raw;
          if (g + 1 < n_out) {
            ld_group(g + 1, vn);
            rnext = ld_res(g + 1);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&acc_empty[as]));
          }
          PROF_ADD(2);
          if (store) {
            uint32_t packed[4] = {0u, 0u, 0u, 0u};
            if (interior) {
              if (p.acc_dump) {                 // tests only
                for (int j = 0; j < 16; ++j)
                  if (g * 16 + j < p.N) p.acc_dump[in_row * p.N + g * 16 + j] = (int)v[j] - corr;
              }
              // branch-free inner loops (independent chains per channel); the zero points ride the clamp bounds
              uint32_t sv[16];
              const int nvalid_g = p.N - g * 16;
              if (!p.has_add) {
                if (nvalid_g > 8) i8_requant_group<16, false>(v, corr, &s_shift[g * 16], rres, p, sv);
                else i8_requant_group<8, false>(v, corr, &s_shift[g * 16], rres, p, sv);
              } else {
                if (nvalid_g > 8) i8_requant_group<16, true>(v, corr, &s_shift[g * 16], rres, p, sv);
                else i8_requant_group<8, true>(v, corr, &s_shift[g * 16], rres, p, sv);
              }
              const int nvalid = p.N - g * 16;               // channels >= N of the last chunk stay zero (padding planes)
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t lo2 = __byte_perm(sv[4 * k], sv[4 * k + 1], 0x0040);
                const uint32_t hi2 = __byte_perm(sv[4 * k + 2], sv[4 * k + 3], 0x0040);
                uint32_t w4 = __byte_perm(lo2, hi2, 0x5410);
                const int left = nvalid - 4 * k;
                if (left < 4) w4 = left <= 0 ? 0u : (w4 & (0xFFFFFFFFu >> (8 * (4 - left))));
                packed[k] = w4;
              }
            }
            *reinterpret_cast<uint4*>(optr8 + (long long)g * out_step8) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          }
          if (g + 1 < n_out) {
            rres = rnext;
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = vn[i];
          }
          PROF_ADD(3);
        }
      } else {
      float* optr = p.out + (store ? orow : 0) * 4;
      // dropout mask row of this pixel's image (stacked: sample sidx of the shared input adds sidx * B images)
      const float* mrow = (MASKED && p.out_mask && interior) ? p.out_mask + (size_t)(z * p.B + (int)b) * p.N : nullptr;
      const float* rptr = (p.residual && interior) ? p.residual + in_row * 4 : nullptr;
      float4 rres[4], rnext[4];
      uint32_t v[16], vn[16];
      // running pointers: one 64-bit add per 16-byte chunk instead of a 64-bit multiply (60 chunks per row on the stacked first layer)
      const long long res_step = p.res_plane * 4, out_step = p.out_plane * 4;
      const long long out_wrap = STACKED ? ((long long)p.Qs - (long long)cps * p.out_plane) * 4 : 0;    // next stacked sample
      const float* rp = rptr;
      auto prefetch = [&](float4* dst, int g) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ch = g * 4 + i;
          dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rptr && ch < n_chunks) {
            dst[i] = ld_nc4(rp);
            rp += res_step;
          }
        }
      };
      prefetch(rres, 0);
      PROF_ADD(0);
      warp_wait(&acc_full[as], pacc, lane);
      PROF_ADD(1);
      tc_fence_after();
      const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(as * p.acc_cols);
      tmem_ld16(tlane, v);
      float* oc = optr;                                      // output pointer of the current chunk
      int jc = 0, sidx = 0;                                  // channel chunk inside the sample / stacked sample
      for (int g = 0; g < n_groups; ++g) {
        tmem_ld_wait();
        if (g + 1 < n_groups) {
          tmem_ld16(tlane + (uint32_t)((g + 1) * 16), vn);
          prefetch(rnext, g + 1);
        } else {                                            // accumulator fully read: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&acc_empty[as]));
        }
        PROF_ADD(2);
        if (store) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int ch = g * 4 + i;
            if (ch < n_chunks) {
              float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
              if (interior) {
                const float4 sc = *reinterpret_cast<const float4*>(&s_scale[jc * 4]);
                const float4 sh = *reinterpret_cast<const float4*>(&s_shift[jc * 4]);
                if (MASKED) {   // general order: affine -> ReLU(pre) -> mask*mult -> +residual
                  o.x = fmaf(__uint_as_float(v[4 * i + 0]), sc.x, sh.x);
                  o.y = fmaf(__uint_as_float(v[4 * i + 1]), sc.y, sh.y);
                  o.z = fmaf(__uint_as_float(v[4 * i + 2]), sc.z, sh.z);
                  o.w = fmaf(__uint_as_float(v[4 * i + 3]), sc.w, sh.w);
                  if (relu_pre) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                  if (mrow) {     // x = mul(x, mask); x = mul_scalar(x, multiplier): two roundings like the reference
                    const float4 mk = *reinterpret_cast<const float4*>(mrow + (STACKED ? (size_t)sidx * p.B * p.N : (size_t)0) + jc * 4);
                    o.x = __fmul_rn(__fmul_rn(o.x, mk.x), p.out_mask_mult); o.y = __fmul_rn(__fmul_rn(o.y, mk.y), p.out_mask_mult);
                    o.z = __fmul_rn(__fmul_rn(o.z, mk.z), p.out_mask_mult); o.w = __fmul_rn(__fmul_rn(o.w, mk.w), p.out_mask_mult);
                  }
                  o.x += rres[i].x; o.y += rres[i].y; o.z += rres[i].z; o.w += rres[i].w;
                } else {
                  o.x = fmaf(__uint_as_float(v[4 * i + 0]), sc.x, sh.x) + rres[i].x;
                  o.y = fmaf(__uint_as_float(v[4 * i + 1]), sc.y, sh.y) + rres[i].y;
                  o.z = fmaf(__uint_as_float(v[4 * i + 2]), sc.z, sh.z) + rres[i].z;
                  o.w = fmaf(__uint_as_float(v[4 * i + 3]), sc.w, sh.w) + rres[i].w;
                }
                if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                if (rnd) { o.x = tf32_round(o.x); o.y = tf32_round(o.y); o.z = tf32_round(o.z); o.w = tf32_round(o.w); }
              }
              *reinterpret_cast<float4*>(oc) = o;
              oc += out_step;
              ++jc;
              if (STACKED && jc == cps) { jc = 0; ++sidx; oc += out_wrap; }
            }
          }
        }
        if (g + 1 < n_groups) {
#pragma unroll
          for (int i = 0; i < 4; ++i) rres[i] = rnext[i];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = vn[i];
        }
        PROF_ADD(3);
      }
      }
      if (++as == p.ACC) { as = 0; pacc ^= 1; }
    }
  }
  if (PROF && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 5)) {   // (producer numbers: lane 0's view)
    const int role = warp == 0 ? 0 : 16;
    for (int i = 0; i < 8; ++i) g_p4_prof[role + i] = prof_acc[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

}  // namespace

extern "C" int qbn_p4_weight_floats(int C, int N, int R, int S, int stride, long long* out_floats) {
  QBN_CHECK_ARG(out_floats && C > 0 && N > 0 && R > 0 && S > 0, "sizes");
  const int CB = qbn_p4_block_channels(C, stride, R * S);
  if (C % 8 != 0 || CB == 0 || N > 256) {
    qbn_set_error("planar-C4 conv: needs C %% 8 == 0 and N <= 256 (C=%d N=%d)", C, N);
    return QBN_ERR_UNSUPPORTED;
  }
  *out_floats = (long long)(C / CB) * R * S * (CB / 4) * qbn_p4_n_pad(N) * 4;
  return QBN_OK;
}

// Programmatic dependent launch of the planar conv kernels (off by default; the MC engines switch it on for their graphs)
static int g_p4_pdl = 0;
extern "C" int qbn_set_pdl(int enabled) {
  g_p4_pdl = enabled ? 1 : 0;
  return QBN_OK;
}
// Unit window of the sample-sharded evaluation (dist.shard_units): of the `n_samples` samples of the following launches the FIRST
// only needs images [first_img, B), the LAST only [0, end_img) — a launch over that many samples then covers the tiles of exactly
// those rows (plus the top border of image end_img, which is the bottom padding of image end_img - 1).  Launches over another
// sample count (a fixed-weight layer on the shared input runs once for all samples), launches with the samples stacked along N
// (the shared-input first layer) and the LRT kinds ignore it.  n_samples 0 switches it off.  Sticky like qbn_set_pdl.
static int g_p4_win_first = 0, g_p4_win_end = 0, g_p4_win_n = 0;
extern "C" int qbn_p4_set_window(int first_img, int end_img, int n_samples) {
  QBN_CHECK_ARG(first_img >= 0 && end_img >= 0 && n_samples >= 0, "window");
  g_p4_win_first = first_img;
  g_p4_win_end = end_img;
  g_p4_win_n = n_samples;
  return QBN_OK;
}
template <typename K>
static cudaError_t p4_launch(K kernel, int grid, size_t smem, cudaStream_t st, const P4Params& p) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(P4_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_p4_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, p);
}

struct P4Planes { long long x, res, out, x2; };      // rows per chunk plane of each tensor (phases * maps + zero tail)
struct P4I8 {                                        // int8 extras of a launch (NULL: TF32)
  int x_shared, z_w, z_out, q_lo, q_hi;
  float atw, mult;
  int has_add, z_res, z_add, add_lo, add_hi, z_fin;
  float s_a, p_a, s_b, p_b, inv_s_add;
  int32_t* acc_dump;
};
struct P4LRT {                                       // LRT extras of a launch (NULL: eval)
  int mode;                                          // 0 forward, 1 input gradient
  const float* x_sq; const float* eps; const float* xin; float* out2; long long aux_plane;
  unsigned long long seed; uint32_t stream_a, stream_b;
  int nhwc, H_out, W_out, oh_mul, oh_add, ow_mul, ow_add;      // dense NHWC out / out2 / eps / xin (see P4Params)
  int phase_z;                                       // all four phases of a stride-2 input gradient in one launch (n_samples = 4)
  int n_taps;                                        // > 0: explicit tap list on maps with a (1, 1) border: tap t reads row q + shift[t] >= q
  int shift[MAX_TAPS];
};
static int conv_p4_launch(int n_samples, int B, int Hp, int Wp, int C, int N, int R, int S, int stride, const float* x, const float* w,
                          int w_shared, const float* scale, const float* shift, const float* residual, const float* out_mask,
                          float out_mask_mult, int flags, float* out, const float* x2, int C2, int CB2, P4Planes pl, void* stream,
                          const P4I8* i8 = nullptr, const P4LRT* lrt = nullptr);

extern "C" int qbn_conv_p4_fwd(int n_samples, int B, int Hp, int Wp, int C, int N, int R, int S, int stride, const float* x,
                               long long x_plane_rows, const float* w, int w_shared, const float* scale, const float* shift,
                               const float* residual, long long res_plane_rows, const float* out_mask, float out_mask_mult, int flags,
                               float* out, long long out_plane_rows, void* stream) {
  P4Planes pl = {x_plane_rows, res_plane_rows, out_plane_rows, 0};
  return conv_p4_launch(n_samples, B, Hp, Wp, C, N, R, S, stride, x, w, w_shared, scale, shift, residual, out_mask, out_mask_mult, flags, out,
                        nullptr, 0, 0, pl, stream);
}

// channels per block of the fused shortcut input: largest multiple of 8 dividing C2 that fits the main conv's activation slot
// (geometry-free rule, shared with the sampler: at most the main conv's block, so the slots — and the occupancy — stay the
// same; a bigger shortcut block was measured slower on the 48-channel layer: it costs the second CTA per SM)
static int p4_shortcut_block(int C, int C2) {
  const int cb_main = qbn_p4_block_channels(C, 1, 9);
  const int cap = cb_main;
  for (int cb = (C2 < cap ? C2 : cap) / 8 * 8; cb >= 8; cb -= 8)
    if (C2 % cb == 0) return cb;
  return 0;
}
extern "C" int qbn_p4_shortcut_block_channels(int C, int C2) { return p4_shortcut_block(C, C2); }

extern "C" int qbn_conv_p4_shortcut_fwd(int n_samples, int B, int Hp, int Wp, int C, int N, int R, int S, const float* x,
                                        long long x_plane_rows, const float* w, const float* x2, long long x2_plane_rows, int C2,
                                        const float* scale, const float* shift, int flags, float* out, long long out_plane_rows,
                                        void* stream) {
  QBN_CHECK_ARG(x2 && C2 > 0 && C2 % 8 == 0, "second input");
  const int cb2 = p4_shortcut_block(C, C2);
  if (cb2 == 0) {
    qbn_set_error("qbn_conv_p4_shortcut_fwd: no channel blocking for C2=%d", C2);
    return QBN_ERR_UNSUPPORTED;
  }
  P4Planes pl = {x_plane_rows, 0, out_plane_rows, x2_plane_rows};
  return conv_p4_launch(n_samples, B, Hp, Wp, C, N, R, S, 1, x, w, 0, scale, shift, nullptr, nullptr, 1.0f, flags, out, x2, C2, cb2, pl, stream);
}

static int conv_p4_launch(int n_samples, int B, int Hp, int Wp, int C, int N, int R, int S, int stride, const float* x, const float* w,
                          int w_shared, const float* scale, const float* shift, const float* residual, const float* out_mask,
                          float out_mask_mult, int flags, float* out, const float* x2, int C2, int CB2, P4Planes pl, void* stream,
                          const P4I8* i8, const P4LRT* lrt) {
  cudaStream_t st = (cudaStream_t)stream;
  QBN_CHECK_ARG(x && w && out, "null pointer");
  QBN_CHECK_ARG(n_samples > 0 && B > 0 && Hp > 2 && Wp > 2 && C > 0 && N > 0 && R > 0 && S > 0, "sizes");
  const int E = i8 ? 16 : 4;                       // channels per 16-byte K-chunk
  const int CB = i8 ? qbn_p16_block_channels(C, stride, R * S) : qbn_p4_block_channels(C, stride, R * S);
  const bool ct = lrt && lrt->n_taps > 0;           // explicit taps (phase launches of a stride-2 input gradient): R x S = 1 x n_taps
  const bool s1 = stride == 1 && (R & 1) && (S & 1);
  const bool s2 = stride == 2 && ((R == 3 && S == 3) || (R == 1 && S == 1));
  if (i8 && (C % 32 != 0 || CB == 0 || N + 1 > 256 || !(s1 || s2) || R * S > MAX_TAPS || x2 || out_mask || (flags & QBN_FLAG_X_SHARED_STACKED))) {
    qbn_set_error("qbn_i8_conv_p16_fwd: needs C %% 32 == 0 (zero-padded channels), N <= 255 and stride 1 (odd kernel) or stride 2 (3x3 / 1x1) "
                  "(C=%d N=%d R=%d S=%d stride=%d)", C, N, R, S, stride);
    return QBN_ERR_UNSUPPORTED;
  }
  if (!i8 && (C % 8 != 0 || CB == 0 || N % 4 != 0 || N > 256 || !(s1 || s2 || ct) || R * S > MAX_TAPS)) {
    qbn_set_error("qbn_conv_p4_fwd: needs C %% 8 == 0, N %% 4 == 0, N <= 256 and stride 1 (odd kernel) or stride 2 (3x3 / 1x1) "
                  "(C=%d N=%d R=%d S=%d stride=%d)", C, N, R, S, stride);
    return QBN_ERR_UNSUPPORTED;
  }
  const bool stacked = flags & QBN_FLAG_X_SHARED_STACKED;
  if (stacked && (n_samples * N > 256 || residual || (flags & QBN_FLAG_OUT_PHASE_SPLIT) || stride != 1)) {
    qbn_set_error("qbn_conv_p4_fwd: sample-stacked mode needs n_samples * N <= 256, stride 1, no residual, normal output (n_samples=%d N=%d)",
                  n_samples, N);
    return QBN_ERR_UNSUPPORTED;
  }
  P4Params p;
  memset(&p, 0, sizeof(p));
  p.Hp = Hp; p.Wp = Wp; p.B = B; p.N = N;
  p.stacked = stacked ? 1 : 0;
  p.cps = N / 4;
  p.n_chunks = (stacked ? n_samples : 1) * N / 4;
  if (i8) {
    p.x_shared = i8->x_shared; p.z_w = i8->z_w; p.z_out = i8->z_out; p.q_lo = i8->q_lo; p.q_hi = i8->q_hi; p.atw = i8->atw; p.mult = i8->mult;
    p.has_add = i8->has_add; p.z_res = i8->z_res; p.z_add = i8->z_add; p.add_lo = i8->add_lo; p.add_hi = i8->add_hi; p.z_fin = i8->z_fin;
    p.s_a = i8->s_a; p.p_a = i8->p_a; p.s_b = i8->s_b; p.p_b = i8->p_b; p.inv_s_add = i8->inv_s_add; p.acc_dump = i8->acc_dump;
    p.n_out_chunks = (N + 15) / 16;
  }
  if (lrt) {
    QBN_CHECK_ARG(!i8 && !stacked && !x2 && !out_mask && (n_samples == 1 || (lrt->phase_z && n_samples == 4)) && lrt->x_sq,
                  "LRT launch: one 'sample' (or the four phases), two operand tensors");
    QBN_CHECK_ARG(!(flags & QBN_FLAG_OUT_PHASE_SPLIT), "LRT launch: normal output layout");
    p.dual = 1; p.lrt_mode = lrt->mode; p.x_sq = lrt->x_sq; p.eps = lrt->eps; p.xin = lrt->xin; p.out2 = lrt->out2; p.aux_plane = lrt->aux_plane;
    p.seed = lrt->seed; p.stream_a = lrt->stream_a; p.stream_b = lrt->stream_b; p.sbase = qbn_sample_base_ptr();
    p.nhwc = lrt->nhwc; p.H_out = lrt->H_out; p.W_out = lrt->W_out;
    p.oh_mul = lrt->oh_mul; p.oh_add = lrt->oh_add; p.ow_mul = lrt->ow_mul; p.ow_add = lrt->ow_add;
    p.phase_z = lrt->phase_z;
    if (lrt->phase_z) p.x_shared = 1;               // every phase reads the same g / dv maps
    QBN_CHECK_ARG(!ct || (stride == 1 && R == 1 && S == lrt->n_taps && S <= MAX_TAPS), "explicit tap list: R = 1, S = n_taps, stride 1");
  }
  p.bh = (s1 && !ct) ? (R - 1) / 2 : 1;
  p.bw = (s1 && !ct) ? (S - 1) / 2 : 1;
  QBN_CHECK_ARG(Hp > p.bh && Wp > p.bw, "padded extent must exceed the border");
  p.Qs = B * Hp * Wp;
  p.tiles_per_sample = (p.Qs + TM - 1) / TM;
  p.total_tiles = p.tiles_per_sample * (stacked ? 1 : n_samples);
  p.tile0 = 0;
  if (g_p4_win_n > 0 && g_p4_win_n == n_samples && (g_p4_win_first > 0 || g_p4_win_end > 0) && !stacked && !lrt) {
    const long long map_rows = (long long)Hp * Wp;
    const int first = g_p4_win_first, end = g_p4_win_end > 0 ? g_p4_win_end : B;
    QBN_CHECK_ARG(first < B && end <= B && (n_samples > 1 || first < end), "unit window outside the batch");
    long long q_end = end * map_rows + (end < B ? (long long)p.bh * Wp + p.bw : 0);
    if (q_end > p.Qs) q_end = p.Qs;
    p.tile0 = (int)(first * map_rows / TM);
    p.total_tiles = (n_samples - 1) * p.tiles_per_sample + (int)((q_end + TM - 1) / TM) - p.tile0;
  }
  p.n_pad = i8 ? qbn_p16_n_pad(N) : qbn_p4_n_pad(stacked ? n_samples * N : N);
  p.n_cb = C / CB;
  p.cbc = CB / E;
  p.nk = p.cbc / 2;
  p.taps = R * S;
  p.strip_rows = (long long)((stacked || (i8 && i8->x_shared) || (lrt && lrt->phase_z)) ? 1 : n_samples) * p.Qs;
  int d_after = 0;
  if (ct) {
    p.n_strips = 1;
    p.d_before = 0;
    d_after = Wp + 1;
  } else if (s1) {
    p.n_strips = 1;
    p.d_before = p.bh * Wp + p.bw;
    d_after = p.d_before;
  } else {
    p.n_strips = (R == 3) ? 4 : 1;
    p.d_before = (R == 3) ? Wp + 1 : 0;
  }
  // plane strides come from the caller: a plane = phases * maps + the zero tail that the last map's bottom/right taps read
  const long long tail = (s1 || ct) ? (long long)p.bh * Wp + p.bw : 0;
  p.x_plane = pl.x;
  if (p.x_plane < p.strip_rows * (s2 ? 4 : 1) + tail) {
    qbn_set_error("qbn_conv_p4_fwd: x plane has %lld rows, needs %lld (maps) + %lld (zero tail)", p.x_plane, p.strip_rows * (s2 ? 4 : 1), tail);
    return QBN_ERR_INVALID_ARG;
  }
  p.RA = TM + p.d_before + d_after;
  p.RA_p = (p.RA + 7) / 8 * 8;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) {
      int strip = 0, sh;
      if (ct) {
        sh = lrt->shift[s];
        if (sh < 0 || sh > d_after) { qbn_set_error("explicit tap shift %d outside [0, Wp + 1]", sh); return QBN_ERR_INVALID_ARG; }
      } else if (s1) {
        sh = (r - p.bh) * Wp + (s - p.bw);
      } else if (R == 3) {
        const int dr = r - 1, ds = s - 1;
        strip = (dr & 1) * 2 + (ds & 1);
        sh = (dr < 0 ? -1 : 0) * Wp + (ds < 0 ? -1 : 0);
      } else {
        sh = 0;
      }
      p.tap_off[r * S + s] = strip * p.cbc * p.RA_p + p.d_before + sh;
    }
  p.a_bytes = (uint32_t)p.n_strips * p.cbc * p.RA_p * 16;
  p.bt_bytes = (uint32_t)p.cbc * p.n_pad * 16;
  if (x2) {
    QBN_CHECK_ARG(stride == 1 && !stacked && !(flags & QBN_FLAG_OUT_PHASE_SPLIT), "fused shortcut: stride-1 main conv, normal output");
    p.x2 = x2;
    p.x2_plane = pl.x2;                                  // phase (0,0) of the phase-split block input is read
    QBN_CHECK_ARG(p.x2_plane >= 4 * p.strip_rows, "x2 plane too small");
    p.cbc2 = CB2 / 4; p.n_cb2 = C2 / CB2; p.nk2 = p.cbc2 / 2;
    p.bt2_bytes = (uint32_t)p.cbc2 * p.n_pad * 16;
    if ((uint32_t)p.cbc2 * TM * 16 > p.a_bytes) p.a_bytes = (uint32_t)p.cbc2 * TM * 16;       // the slots must hold a shortcut block too
    QBN_CHECK_ARG(p.bt2_bytes <= p.bt_bytes * (uint32_t)p.taps, "shortcut weight block larger than the main conv's");
  }
  p.w_sample_floats = ((long long)p.n_cb * (p.dual ? 2 : 1) * p.taps * p.bt_bytes + (long long)p.n_cb2 * p.bt2_bytes) / 4;
  p.flags = flags; p.w_shared = stacked ? 1 : w_shared;
  p.x = x; p.w = w; p.scale = scale; p.shift = shift; p.residual = residual; p.out = out;
  p.out_mask = out_mask; p.out_mask_mult = out_mask_mult;
  p.idesc = i8 ? ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_pad >> 3) << 17) | ((uint32_t)(TM >> 4) << 24))      // S32 += S8 x S8
               : ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.n_pad >> 3) << 17) | ((uint32_t)(TM >> 4) << 24));    // F32 += TF32 x TF32
  p.mg_plane = (uint32_t)(0x100000000ull / (uint64_t)(Hp * Wp));
  p.mg_wp = (uint32_t)(0x100000000ull / (uint64_t)Wp);
  p.res_plane = pl.res;
  p.out_plane = pl.out;
  QBN_CHECK_ARG(!residual || p.res_plane >= (long long)n_samples * p.Qs, "residual plane too small");
  if (flags & QBN_FLAG_OUT_PHASE_SPLIT) {
    const int H = Hp - p.bh, W = Wp - p.bw;
    QBN_CHECK_ARG((H % 2 == 0) && (W % 2 == 0), "phase-split output needs even H, W");
    p.out_split = 1;
    p.Hp2 = H / 2 + 1; p.Wp2 = W / 2 + 1;
    p.q2_total = (long long)n_samples * B * p.Hp2 * p.Wp2;
    QBN_CHECK_ARG(p.out_plane >= 4 * p.q2_total, "phase-split out plane too small");
  }
  // ---- shared memory / occupancy policy ----
  const size_t b_all = (size_t)p.bt_bytes * p.n_cb * (p.dual ? 2 : 1) * p.taps + (size_t)p.bt2_bytes * p.n_cb2;
  p.acc_cols = p.dual ? 2 * p.n_pad : p.n_pad;
  QBN_CHECK_ARG(p.acc_cols <= 512, "two accumulators of this width do not fit TMEM (N <= 256 for LRT)");
  const size_t fixed = 2 * 256 * 4 + 16 + 8 * 64 + 8 * 256;          // affine tables, TMEM slot, barriers, MMA operand list
  QBN_CHECK_ARG(p.taps * p.nk <= 256, "too many MMAs per channel block");
  const size_t cap = 225 * 1024;
  int want_occ;
  size_t smem;
  const char* e_occ = tune_env("QBN_P4_OCC");
  if (b_all <= 100 * 1024 && b_all < (1u << 20)) {
    p.b_res = 1; p.SB = 1; p.TG = p.taps; p.b_slot_bytes = (uint32_t)b_all;
    want_occ = e_occ ? atoi(e_occ) : (lrt ? 2 : (i8 ? 4 : 3));      // (the LRT epilogue keeps two operand groups in flight: 2 CTAs per SM by registers)
    while (want_occ > 1 && 2 * (size_t)p.a_bytes + b_all + fixed > cap / want_occ - 1024) --want_occ;
    p.SA = 2;
    while (p.SA < 4 && (size_t)(p.SA + 1) * p.a_bytes + b_all + fixed <= cap / want_occ - 1024) ++p.SA;
    smem = (size_t)p.SA * p.a_bytes + b_all + fixed;
  } else {
    // streamed weights: one slot = TG consecutive taps of a channel block (contiguous in the blocked layout).
    // Wide outputs (n_pad > 128) get ONE CTA per SM — two accumulators in TMEM so the epilogue overlaps the MMAs —
    // and big slots (few handshakes); narrower ones two CTAs (two issuers) with ~16-28 KB slots.  (measured)
    p.b_res = 0;
    want_occ = e_occ ? atoi(e_occ) : (p.n_pad > 128 ? 1 : 2);
    const size_t slot_cap = want_occ == 1 ? 76 * 1024 : 28 * 1024;
    p.TG = 1;
    for (int tg = 2; tg <= p.taps; ++tg)
      if (p.taps % tg == 0 && (size_t)tg * p.bt_bytes <= slot_cap) p.TG = tg;
    if (tune_env("QBN_P4_TG")) p.TG = atoi(tune_env("QBN_P4_TG"));
    p.b_slot_bytes = p.bt_bytes * (uint32_t)p.TG;
    while (want_occ > 1 && 2 * (size_t)p.a_bytes + 3 * (size_t)p.b_slot_bytes + fixed > cap / want_occ - 1024) --want_occ;
    while (p.TG > 1 && 2 * (size_t)p.a_bytes + 2 * (size_t)p.b_slot_bytes + fixed > cap / want_occ - 1024) {   // shrink the slots until two fit
      do { --p.TG; } while (p.taps % p.TG != 0);
      p.b_slot_bytes = p.bt_bytes * (uint32_t)p.TG;
    }
    p.SA = 2; p.SB = 2;
    p.tile_rr = tune_env("QBN_P4_RR") ? atoi(tune_env("QBN_P4_RR")) : 1;
    while (p.SB < 8 && (size_t)p.SA * p.a_bytes + (size_t)(p.SB + 1) * p.b_slot_bytes + fixed <= cap / want_occ - 1024) ++p.SB;
    smem = (size_t)p.SA * p.a_bytes + (size_t)p.SB * p.b_slot_bytes + fixed;
  }
  if (tune_env("QBN_P4_SA")) {
    const int sa_new = atoi(tune_env("QBN_P4_SA"));
    smem += (size_t)(sa_new - p.SA) * p.a_bytes;
    p.SA = sa_new;
  }
  if (p.n_cb2 && !p.b_res && p.bt2_bytes > p.b_slot_bytes) {
    qbn_set_error("qbn_conv_p4_shortcut_fwd: shortcut weight block (%u B) exceeds the streamed weight slot (%u B)", p.bt2_bytes, p.b_slot_bytes);
    return QBN_ERR_UNSUPPORTED;
  }
  if (smem > cap || p.SA < 1) {
    qbn_set_error("qbn_conv_p4_fwd: tile does not fit shared memory (%zu bytes)", smem);
    return QBN_ERR_UNSUPPORTED;
  }
  {
    int share = 512 / want_occ, c2 = 32;
    while (c2 * 2 <= share) c2 <<= 1;
    p.ACC = c2 / p.acc_cols;
    if (p.ACC > 4) p.ACC = 4;
    if (p.ACC < 1) { p.ACC = 1; }
    if (tune_env("QBN_P4_ACC")) p.ACC = atoi(tune_env("QBN_P4_ACC"));
    p.tmem_cols = 32;
    while (p.tmem_cols < p.ACC * p.acc_cols) p.tmem_cols <<= 1;
    if (p.tmem_cols > 512) { p.ACC = 512 / p.acc_cols; p.tmem_cols = 512; }
    while (p.tmem_cols * want_occ > 512) --want_occ;
  }
  static bool attr_set = false;
  if (!attr_set) {
    QBN_CUDA(cudaFuncSetAttribute(umma_conv_p4_kernel<0, false, false, KIND_I8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    QBN_CUDA(cudaFuncSetAttribute(umma_conv_p4_kernel<0, false, false, KIND_LRT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
#ifdef QBN_TUNING
    QBN_CUDA(cudaFuncSetAttribute(umma_conv_p4_kernel<1, false, false, KIND_LRT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    QBN_CUDA(cudaFuncSetAttribute(umma_conv_p4_kernel<1, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    QBN_CUDA(cudaFuncSetAttribute(umma_conv_p4_kernel<1, false, false, KIND_I8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
#endif
    QBN_CUDA(cudaFuncSetAttribute(umma_conv_p4_kernel<0, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    QBN_CUDA(cudaFuncSetAttribute(umma_conv_p4_kernel<0, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    QBN_CUDA(cudaFuncSetAttribute(umma_conv_p4_kernel<0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    QBN_CUDA(cudaFuncSetAttribute(umma_conv_p4_kernel<0, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    attr_set = true;
  }
  int occ = (int)((227 * 1024) / (smem + 1024));
  if (occ > want_occ) occ = want_occ;
  if (occ < 1) occ = 1;
  int grid = qbn_sm_count() * occ;
  if (grid > p.total_tiles) grid = p.total_tiles;
  if (tune_env("QBN_P4_VERBOSE"))
    fprintf(stderr, "[p4] C=%d N=%d %dx%d k%d s%d: tiles=%d grid=%d occ=%d SA=%d SB=%d TG=%d ACC=%d b_res=%d smem=%zu a_bytes=%u bt=%u tmem=%d\n", C, N, Hp,
            Wp, R, stride, p.total_tiles, grid, occ, p.SA, p.SB, p.TG, p.ACC, p.b_res, smem, p.a_bytes, p.bt_bytes, p.tmem_cols);
  // the general epilogue order (pre-ReLU, output mask) is a separate instantiation: the plain one keeps its fused fma + residual
  bool masked = out_mask != nullptr;
  if (!masked && (p.flags & QBN_FLAG_RELU_PRE)) {
    if (residual) masked = true;                                  // ReLU before the residual add: general order
    else p.flags = (p.flags & ~QBN_FLAG_RELU_PRE) | QBN_FLAG_RELU;  // no mask, no residual: pre == post
  }
#ifdef QBN_TUNING
  auto prof_report = [&](const char* kind) {
    unsigned long long h[32];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(h, g_p4_prof, sizeof(h));
    const unsigned long long t0 = (unsigned long long)(p.total_tiles / grid > 0 ? p.total_tiles / grid : 1);
    fprintf(stderr, "[p4 prof %s] C=%d N=%d k%d s%d grid=%d tiles/CTA=%llu SA=%d SB=%d ACC=%d b_res=%d | epi: pre %llu wait_acc %llu tmem_ld %llu "
            "compute+store %llu | mma: bres %llu wait_acc_empty %llu wait_a %llu wait_b %llu issue %llu commit %llu | prod: bres %llu "
            "wait_a_empty %llu issueA %llu wait_b_empty %llu issueB %llu (cycles per tile)\n",
            kind, C, N, R, stride, grid, t0, p.SA, p.SB, p.ACC, p.b_res, h[0] / t0, h[1] / t0, h[2] / t0, h[3] / t0, h[8] / t0, h[9] / t0, h[10] / t0,
            h[11] / t0, h[12] / t0, h[13] / t0, h[16] / t0, h[17] / t0, h[18] / t0, h[19] / t0, h[20] / t0);
  };
#endif
  if (lrt) {
#ifdef QBN_TUNING
    if (tune_env("QBN_P4_PROF")) {
      unsigned long long h[32] = {0};
      cudaMemcpyToSymbol(g_p4_prof, h, sizeof(h));
      p4_launch(umma_conv_p4_kernel<1, false, false, KIND_LRT>, grid, smem, st, p);
      QBN_CHECK_LAUNCH();
      prof_report(lrt->mode ? "lrt-dgrad" : "lrt-fwd");
      return QBN_OK;
    }
#endif
    p4_launch(umma_conv_p4_kernel<0, false, false, KIND_LRT>, grid, smem, st, p);
    QBN_CHECK_LAUNCH();
    return QBN_OK;
  }
  if (i8) {
#ifdef QBN_TUNING
    if (tune_env("QBN_P4_PROF")) {       // diagnostics: cycle accounting of CTA 0 (synchronises the stream)
      unsigned long long h[32] = {0};
      cudaMemcpyToSymbol(g_p4_prof, h, sizeof(h));
      p4_launch(umma_conv_p4_kernel<1, false, false, KIND_I8>, grid, smem, st, p);
      QBN_CHECK_LAUNCH();
      prof_report("i8");
      return QBN_OK;
    }
#endif
    p4_launch(umma_conv_p4_kernel<0, false, false, KIND_I8>, grid, smem, st, p);
    QBN_CHECK_LAUNCH();
    return QBN_OK;
  }
  if (stacked || masked) {
    if (stacked && masked) p4_launch(umma_conv_p4_kernel<0, true, true>, grid, smem, st, p);
    else if (stacked) p4_launch(umma_conv_p4_kernel<0, true, false>, grid, smem, st, p);
    else p4_launch(umma_conv_p4_kernel<0, false, true>, grid, smem, st, p);
    QBN_CHECK_LAUNCH();
    return QBN_OK;
  }
#ifdef QBN_TUNING
  if (tune_env("QBN_P4_PROF")) {
    unsigned long long h[32] = {0};
    cudaMemcpyToSymbol(g_p4_prof, h, sizeof(h));
    p4_launch(umma_conv_p4_kernel<1, false, false>, grid, smem, st, p);
    QBN_CHECK_LAUNCH();
    prof_report("tf32");
    return QBN_OK;
  }
#endif
  p4_launch(umma_conv_p4_kernel<0, false, false>, grid, smem, st, p);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

// ---- int8 on the same zero-copy kernel (SURVEY 8a row A6 step 6 + A11 glue): u8 x s8 -> s32 with FBGEMM's requantisation,
// the BasicBlock's quantized::add_relu and the activation clamp in the epilogue.  Maps hold (q - zero_point) as s8 in the
// planar-C16 layout (16 channels per 16-byte chunk, channel count zero-padded to a multiple of 32 = one kind::i8 MMA).
extern "C" int qbn_i8_conv_p16_fwd(int n_samples, int B, int Hp, int Wp, int C, int N, int R, int S, int stride, const int8_t* x,
                                   long long x_plane_rows, int x_shared, const int8_t* w_blocked, int w_shared, const float* bias,
                                   const qbn_i8_requant* rq, const int8_t* residual, long long res_plane_rows, int flags, int8_t* out,
                                   long long out_plane_rows, int32_t* acc_dump, void* stream) {
  QBN_CHECK_ARG(rq, "requantisation parameters");
  QBN_CHECK_ARG(rq->s_x > 0 && rq->s_w > 0 && rq->s_out > 0, "scales must be > 0");
  QBN_CHECK_ARG(rq->act_max >= 1 && rq->act_max <= 127, "activations must fit 7 bits (quant_utils.py:120): the maps hold q - zero_point as s8");
  QBN_CHECK_ARG(rq->z_out >= 0 && rq->z_out <= 127 && rq->z_w >= -128 && rq->z_w <= 127, "zero points");
  P4I8 e;
  memset(&e, 0, sizeof(e));
  e.x_shared = x_shared; e.z_w = rq->z_w; e.z_out = rq->z_out;
  // ATen qconv (fbgemm): act_times_w = s_x * s_w ; multiplier = act_times_w / s_out, all fp32 (same as qbn_i8_conv_fwd)
  e.atw = rq->s_x * rq->s_w;
  e.mult = e.atw / rq->s_out;
  e.q_lo = rq->relu ? rq->z_out : 0;
  e.q_hi = rq->act_max;
  e.z_fin = rq->z_out;
  e.acc_dump = acc_dump;
  if (residual) {
    QBN_CHECK_ARG(rq->s_res > 0 && rq->s_add > 0 && rq->z_res >= 0 && rq->z_res <= 127 && rq->z_add >= 0 && rq->z_add <= 127, "residual add parameters");
    e.has_add = 1; e.z_res = rq->z_res; e.z_add = rq->z_add;
    e.s_a = rq->s_out; e.p_a = rq->s_out * (float)(-rq->z_out);
    e.s_b = rq->s_res; e.p_b = rq->s_res * (float)(-rq->z_res);
    e.inv_s_add = 1.0f / rq->s_add;
    e.add_lo = rq->add_relu ? rq->z_add : 0;
    e.add_hi = rq->act_max;
    e.z_fin = rq->z_add;
  }
  P4Planes pl = {x_plane_rows, res_plane_rows, out_plane_rows, 0};
  return conv_p4_launch(n_samples, B, Hp, Wp, C, N, R, S, stride, reinterpret_cast<const float*>(x), reinterpret_cast<const float*>(w_blocked),
                        w_shared, nullptr, bias, reinterpret_cast<const float*>(residual), nullptr, 1.0f, flags & QBN_FLAG_OUT_PHASE_SPLIT,
                        reinterpret_cast<float*>(out), nullptr, 0, 0, pl, stream, &e);
}

extern "C" int qbn_p16_weight_bytes(int C, int N, int R, int S, int stride, long long* out_bytes) {
  QBN_CHECK_ARG(out_bytes && C > 0 && N > 0 && R > 0 && S > 0, "sizes");
  const int CB = qbn_p16_block_channels(C, stride, R * S);
  if (C % 32 != 0 || CB == 0 || N + 1 > 256) {
    qbn_set_error("planar-C16 int8 conv: needs C %% 32 == 0 and N <= 255 (C=%d N=%d)", C, N);
    return QBN_ERR_UNSUPPORTED;
  }
  *out_bytes = (long long)(C / CB) * R * S * (CB / 16) * qbn_p16_n_pad(N) * 16;
  return QBN_OK;
}

// ---- LRT training on the same zero-copy kernel (linear.py:32-40, conv.py:24-32 and their autograd).
// Operands: planar C4 maps staged by qbn_p4_stage_input / qbn_p4_stage_grad (lrt_p4.cu); results: dense NHWC, the layout of the
// module boundary.  Weights: blocked like the eval path, the sigma^2 blocks right after the mu blocks (qbn_lrt_p4_weight_prep).
// Forward: out = conv(x, mu) + sqrt(1e-8 + conv(x_sq, sigma^2)) * eps + bias; std_out = the square root (kept for the backward).
// eps: NHWC like the output, or NULL -> Philox(seed, stream_a, stream_b, offset in out / 4) — the draw qbn_lrt_fwd makes.
// Hp, Wp: padded extent of the OUTPUT maps (Ho + border, Wo + border; border = (R-1)/2 x (S-1)/2 at stride 1, 1 x 1 at stride 2).
extern "C" int qbn_lrt_conv_p4_fwd(int B, int Hp, int Wp, int C, int N, int R, int S, int stride, const float* x, const float* x_sq,
                                   long long x_plane_rows, const float* w_blocked, const float* bias, const float* eps, uint64_t seed,
                                   uint32_t stream_a, uint32_t stream_b, float* out, float* std_out, void* stream) {
  QBN_CHECK_ARG(x_sq && std_out, "x_sq / std_out");
  P4LRT l;
  memset(&l, 0, sizeof(l));
  l.mode = 0; l.x_sq = x_sq; l.eps = eps; l.out2 = std_out; l.seed = seed; l.stream_a = stream_a; l.stream_b = stream_b;
  const int bh = stride == 1 ? (R - 1) / 2 : 1, bw = stride == 1 ? (S - 1) / 2 : 1;
  l.nhwc = 1; l.H_out = Hp - bh; l.W_out = Wp - bw; l.oh_mul = l.ow_mul = 1;
  P4Planes pl = {x_plane_rows, 0, 0, 0};
  return conv_p4_launch(1, B, Hp, Wp, C, N, R, S, stride, x, w_blocked, 1, nullptr, bias, nullptr, nullptr, 1.0f, 0, out, nullptr, 0, 0, pl, stream,
                        nullptr, &l);
}
// Input gradient of a stride-1 'same' layer: dx = convT(g, mu) + 2 x .* convT(dv, sigma^2), the SAME kernel on the flipped /
// transposed weights (tap (r,s) <- (R-1-r, S-1-s), in/out channels swapped: qbn_lrt_p4_weight_prep mode 1).  C = channels of g
// (the layer's outputs), N = channels of dx; g, dv planar with the layer's geometry; xin, dx dense NHWC [B][Hp-bh][Wp-bw][N].
extern "C" int qbn_lrt_conv_p4_dgrad(int B, int Hp, int Wp, int C, int N, int R, int S, const float* g, const float* dv, long long g_plane_rows,
                                     const float* w_flipped_blocked, const float* xin, float* dx, void* stream) {
  QBN_CHECK_ARG(dv && xin, "dv / xin");
  P4LRT l;
  memset(&l, 0, sizeof(l));
  l.mode = 1; l.x_sq = dv; l.xin = xin;
  l.nhwc = 1; l.H_out = Hp - (R - 1) / 2; l.W_out = Wp - (S - 1) / 2; l.oh_mul = l.ow_mul = 1;
  P4Planes pl = {g_plane_rows, 0, 0, 0};
  return conv_p4_launch(1, B, Hp, Wp, C, N, R, S, 1, g, w_flipped_blocked, 1, nullptr, nullptr, nullptr, nullptr, 1.0f, 0, dx, nullptr, 0, 0, pl,
                        stream, nullptr, &l);
}
// Input gradient of a stride-2 layer (3x3 pad 1, or 1x1 pad 0), one launch per phase (a, b) of dx: the pixels (2i+a, 2j+b) receive
// sum over the taps (r, s) with r = a+1 (mod 2), s = b+1 (mod 2) of g[i + (r == 0), j + (s == 0)] * w[r, s] — a convolution over
// g's zero-bordered maps (Hp = Ho + 1, Wp = Wo + 1) whose taps shift by 0 / +1 rows and columns; the shared zero border supplies
// the out-of-range reads.  shifts[t] = (r_t == 0) * Wp + (s_t == 0) in the order of the n_taps blocks of w_phase_blocked
// (qbn_lrt_p4_weight_prep mode 2).  xin, dx: dense NHWC [B][2(Hp-1)][2(Wp-1)][N].  A 1x1 layer has the single phase (0, 0).
// All four phases of a 3x3 stride-2 layer's input gradient in ONE launch: phase z = (a, b) is 'sample' z with its own blocked weights
// (qbn_lrt_p4_weight_prep mode 3: four taps (dr, ds) in {0,1}^2 per phase, shift dr * Wp + ds, the taps a phase does not have zeroed),
// all reading the same g / dv maps.  Same result as four qbn_lrt_conv_p4_dgrad_phase calls; a training step's stride-2 layers are
// launch-latency bound.
extern "C" int qbn_lrt_conv_p4_dgrad_s2(int B, int Hp, int Wp, int C, int N, const float* g, const float* dv, long long g_plane_rows,
                                        const float* w_phases_blocked, const float* xin, float* dx, void* stream) {
  QBN_CHECK_ARG(dv && xin, "dv / xin");
  P4LRT l;
  memset(&l, 0, sizeof(l));
  l.mode = 1; l.x_sq = dv; l.xin = xin;
  l.nhwc = 1; l.H_out = 2 * (Hp - 1); l.W_out = 2 * (Wp - 1); l.oh_mul = l.ow_mul = 2;
  l.phase_z = 1;
  l.n_taps = 4;
  l.shift[0] = 0; l.shift[1] = 1; l.shift[2] = Wp; l.shift[3] = Wp + 1;
  P4Planes pl = {g_plane_rows, 0, 0, 0};
  return conv_p4_launch(4, B, Hp, Wp, C, N, 1, 4, 1, g, w_phases_blocked, 0, nullptr, nullptr, nullptr, nullptr, 1.0f, 0, dx, nullptr, 0, 0, pl,
                        stream, nullptr, &l);
}
extern "C" int qbn_lrt_conv_p4_dgrad_phase(int B, int Hp, int Wp, int C, int N, int n_taps, const int* shifts, int phase_a, int phase_b,
                                           const float* g, const float* dv, long long g_plane_rows, const float* w_phase_blocked,
                                           const float* xin, float* dx, void* stream) {
  QBN_CHECK_ARG(dv && xin && shifts && n_taps > 0 && n_taps <= 4, "dv / xin / taps");
  QBN_CHECK_ARG((phase_a | phase_b) >= 0 && phase_a < 2 && phase_b < 2, "phase");
  P4LRT l;
  memset(&l, 0, sizeof(l));
  l.mode = 1; l.x_sq = dv; l.xin = xin;
  l.nhwc = 1; l.H_out = 2 * (Hp - 1); l.W_out = 2 * (Wp - 1); l.oh_mul = l.ow_mul = 2; l.oh_add = phase_a; l.ow_add = phase_b;
  l.n_taps = n_taps;
  for (int t = 0; t < n_taps; ++t) l.shift[t] = shifts[t];
  P4Planes pl = {g_plane_rows, 0, 0, 0};
  return conv_p4_launch(1, B, Hp, Wp, C, N, 1, n_taps, 1, g, w_phase_blocked, 1, nullptr, nullptr, nullptr, nullptr, 1.0f, 0, dx, nullptr, 0, 0, pl,
                        stream, nullptr, &l);
}
