// tcgen05 / TMEM implicit-GEMM kernels (QBN_MATH_TF32 and the int8 path) for sm_100a.
//
//   D[128 output pixels][N] (TMEM, fp32 or s32)  +=  A[128][K] (smem)  x  B[N][K]^T (smem)
//
// * One CTA owns a 128-pixel tile of ONE Monte-Carlo sample and the FULL output-channel extent,
//   so the activation tile is read exactly once.  grid = (M/128, 1, samples).
// * Operands are staged by four producer warps (LDG.128 -> registers -> STS.128) straight into
//   the UMMA canonical K-major no-swizzle ("interleaved") layout: 16-byte K-chunks, 8-row core
//   matrices of 128 contiguous bytes; LBO = distance between K-chunks, SBO = 128 B.  Going
//   through registers is what lets the operand load be *fused*: im2col gather + zero padding,
//   MC-Dropout mask (A8), x^2 for the LRT variance contraction (A1/A2) and the (x - z_x) shift
//   of the int8 path all happen here, so none of those tensors ever exists in HBM.
// * One elected thread of warp 4 issues tcgen05.mma (kind::tf32 / kind::i8); accumulators live in
//   TMEM (LRT: two accumulators side by side: mean in columns [0,N), variance in [N,2N)).
// * smem ring of `stages` slots, mbarrier full/empty pipeline; tcgen05.commit releases slots and
//   finally signals the epilogue.
// * Epilogue (the four producer warps again, one TMEM lane = one output pixel each):
//   tcgen05.ld -> LRT: mean + sqrt(1e-8+var)*eps(+Philox) + bias | eval: affine (BN/bias),
//   residual add, ReLU | int8: FBGEMM requantisation -> global.
// All mbarrier waits are bounded (trap instead of hanging the GPU).
#include <stdlib.h>
#include <string.h>
#include "umma_common.cuh"

// tuning knobs exist only in -DQBN_TUNING builds: the product library reads nothing from the environment
#ifdef QBN_TUNING
static inline const char* tune_env(const char* name) { return getenv(name); }
#else
static inline const char* tune_env(const char*) { return nullptr; }
#endif

namespace {

constexpr int UM = 128;          // rows per CTA tile = TMEM lanes
constexpr int KCH = 8;           // 16-byte K-chunks per stage (BLOCK_K = 128 bytes)
constexpr int NPROD = 128;       // producer threads (warps 0-3), also the epilogue warps
constexpr int NTHREADS = 160;    // + warp 4: TMEM allocator and MMA issuer

struct UParams {
  // geometry
  int B, H, W, C, N, R, S, sh, sw, ph, pw, dh, dw, Ho, Wo, K;
  int M;            // B*Ho*Wo (per sample)
  int n_pad;        // MMA N (multiple of 16)
  int acc_cols;     // TMEM columns used
  int tmem_cols;    // allocated (power of two >= 32)
  int stages;
  int a_pitch;      // rows+pad per K-chunk in A stage (chunks of 16 B)
  int b_pitch;
  int x_shared, w_shared, flags;
  int oph, opw;     // zero-bordered output layout (interior written only)
  int n_split;      // >0: the N columns are n_split channels of N/n_split stacked Monte-Carlo samples (shared input)
  long long sample_out_stride;   // elements between consecutive samples' output tensors
  long long w_ld;                // elements between consecutive samples' weight tensors (0: N*K); N-split launches of wide layers
  int out_ld;                    // channels per stored output row (0: N)
  long long out_plane;           // QBN_FLAG_OUT_P4: rows per chunk plane of the planar-C4 output (all samples)
  uint32_t idesc;
  // tensors
  const void* x; const void* w; const void* w2;
  const float* x2;               // LRT mode: second A operand from its own tensor (dgrad: dv) instead of x^2
  const float* scale; const float* shift; const float* residual; const float* in_mask; float in_mult;
  int tr_sh, tr_sw;              // transposed convolution (dgrad of a strided conv): input coordinate = (h0 + r) / tr_s when divisible
  const float* mul2x;            // LRT dgrad: acc *= 2 * mul2x[out index] before the residual add (dx = dx_mean + 2x .* dx_var)
  const float* bias; const float* eps; uint64_t seed; uint32_t sa, sb; const uint32_t* sbase;
  void* out; float* std_out;
  // int8
  int z_x, z_w, z_out, lo, hi; float atw, mult; int32_t* acc_dump;
};

template <int MODE>
__global__ void __launch_bounds__(NTHREADS) umma_conv_kernel(const UParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr bool LRT = MODE == MODE_LRT;
  constexpr bool I8 = MODE == MODE_I8;
  constexpr int ESZ = I8 ? 1 : 4;           // operand element size
  constexpr int EPC = 16 / ESZ;             // elements per 16-byte chunk
  constexpr int BLOCK_K = KCH * EPC;        // 32 (tf32) or 128 (i8) K-elements per stage
  constexpr int MMA_K = 32 / ESZ;           // 8 (tf32) or 32 (i8) per instruction = 2 chunks

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int z = blockIdx.z;
  const int m0 = blockIdx.x * UM;

  // ---- shared memory carve-up -----------------------------------------------------------------
  const uint32_t a_bytes = (uint32_t)KCH * p.a_pitch * 16;
  const uint32_t b_bytes = (uint32_t)KCH * p.b_pitch * 16;
  const uint32_t stage_bytes = (LRT ? 2 : 1) * (a_bytes + b_bytes);
  uint8_t* ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + p.stages;
  uint64_t* accum_bar = bars + 2 * p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 1);

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), NPROD);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(accum_bar), 1);
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;

  if (warp < 4) {
    // =========================== PRODUCERS ======================================================
    const int kc = tid & 7;          // this thread's 16-byte K-chunk inside every stage
    const int r0 = tid >> 3;         // rows r0 + 16*i
    const uint8_t* xs = reinterpret_cast<const uint8_t*>(p.x) + (p.x_shared ? 0 : (size_t)z * p.B * p.H * p.W * p.C * ESZ);
    const uint8_t* ws = reinterpret_cast<const uint8_t*>(p.w) + (p.w_shared ? 0 : (size_t)z * (p.w_ld ? (size_t)p.w_ld : (size_t)p.N * p.K) * ESZ);
    const uint8_t* ws2 = LRT ? reinterpret_cast<const uint8_t*>(p.w2) : nullptr;
    const float* msk = p.in_mask ? p.in_mask + (size_t)z * p.B * p.C : nullptr;

    int rbase[8], rh0[8], rw0[8], rb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int m = m0 + r0 + 16 * i;
      bool v = m < p.M;
      int mm = v ? m : 0;
      int wo = mm % p.Wo;
      int t = mm / p.Wo;
      int ho = t % p.Ho;
      int b = t / p.Ho;
      rh0[i] = v ? ho * p.sh - p.ph : -(1 << 28);   // invalid rows fail every bounds test -> zeros
      rw0[i] = wo * p.sw - p.pw;
      rbase[i] = b * p.H * p.W * p.C;
      rb[i] = b;
    }

    // Two operand-staging paths:
    //  (a) cp.async (LDGSTS, zero-fill for padding) straight into the UMMA layout — used whenever the
    //      operand needs no arithmetic on the way (eval A without dropout mask and TF32-ready, all B);
    //      the thread never waits for the data: cp.async.mbarrier.arrive.noinc signals the stage's
    //      full barrier when its copies land, so up to `stages` stages of loads are in flight per CTA.
    //  (b) registers (LDG.128 -> transform -> STS.128) for fused transforms: MC-Dropout mask, RNA
    //      rounding to TF32, x^2 for the LRT variance operand, (x - z_x) for int8.  Loads of stage
    //      kb+1 are issued before stage kb is published (one stage of register prefetch).
    const bool a_async = (MODE == MODE_EVAL) && msk == nullptr && (p.flags & QBN_FLAG_A_TF32_READY);
    constexpr int AV = I8 ? 2 : 1;               // 8-byte pieces (i8) or one 16-byte chunk (fp32)
    uint4 areg[8];
    uint4 areg2[LRT ? 8 : 1];            // second A operand when it is a tensor of its own (p.x2)

    auto tap_of = [&](int k, int& c, int& dr, int& ds) {
      const int kk = k < p.K ? k : 0;
      c = kk % p.C;
      const int rs = kk / p.C;
      ds = (rs % p.S) * p.dw;
      dr = (rs / p.S) * p.dh;
    };
    auto load_a_regs = [&](int kb) {
      const int k = kb * BLOCK_K + kc * EPC;
      if constexpr (!I8) {
        int c, dr, ds;
        tap_of(k, c, dr, ds);
        const bool kv = k < p.K;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          int hi = rh0[i] + dr, wi = rw0[i] + ds;
          bool ok = kv && hi >= 0 && wi >= 0;
          if (p.tr_sh > 1 || p.tr_sw > 1) {       // transposed conv: only coordinates on the stride grid carry a gradient
            ok = ok && (hi % p.tr_sh == 0) && (wi % p.tr_sw == 0);
            hi /= p.tr_sh; wi /= p.tr_sw;
          }
          ok = ok && hi < p.H && wi < p.W;
          areg[i] = make_uint4(0u, 0u, 0u, 0u);
          if (ok) areg[i] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(xs) + rbase[i] + (hi * p.W + wi) * p.C + c));
          if constexpr (LRT) {
            if (p.x2) {
              areg2[i] = make_uint4(0u, 0u, 0u, 0u);
              if (ok) areg2[i] = __ldg(reinterpret_cast<const uint4*>(p.x2 + rbase[i] + (hi * p.W + wi) * p.C + c));
            }
          }
        }
      } else {
#pragma unroll
        for (int h = 0; h < AV; ++h) {
          const int k8 = k + 8 * h;
          int c, dr, ds;
          tap_of(k8, c, dr, ds);
          const bool kv = k8 < p.K;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int hi = rh0[i] + dr, wi = rw0[i] + ds;
            const bool ok = kv && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
            uint2 q = make_uint2(0u, 0u);
            if (ok) {
              // int8: two 8-byte pieces per chunk (C % 8 == 0 keeps each piece inside one filter tap);
              // operand = (x - z_x) as s8 (activations are <= 7 bit, quant_utils.py:120), padding -> 0
              q = __ldg(reinterpret_cast<const uint2*>(xs + rbase[i] + (hi * p.W + wi) * p.C + c));
              const uint32_t zz = (uint32_t)p.z_x * 0x01010101u;
              q.x = __vsub4(q.x, zz);
              q.y = __vsub4(q.y, zz);
            }
            if (h == 0) { areg[i].x = q.x; areg[i].y = q.y; } else { areg[i].z = q.x; areg[i].w = q.y; }
          }
        }
      }
    };
    auto store_a_regs = [&](int kb, uint8_t* sa, uint8_t* sa2) {
      if constexpr (I8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(sa + ((size_t)kc * p.a_pitch + r0 + 16 * i) * 16) = areg[i];
      } else {
        const int k = kb * BLOCK_K + kc * EPC;
        int c, dr, ds;
        tap_of(k, c, dr, ds);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 v = make_float4(__uint_as_float(areg[i].x), __uint_as_float(areg[i].y), __uint_as_float(areg[i].z), __uint_as_float(areg[i].w));
          if (msk) {  // A8: x * mask[b,c] * 1/(1-p) in the operand load (dropout.py:38-39)
            const float4 mk = __ldg(reinterpret_cast<const float4*>(msk + (size_t)rb[i] * p.C + c));
            v.x = __fmul_rn(__fmul_rn(v.x, mk.x), p.in_mult);
            v.y = __fmul_rn(__fmul_rn(v.y, mk.y), p.in_mult);
            v.z = __fmul_rn(__fmul_rn(v.z, mk.z), p.in_mult);
            v.w = __fmul_rn(__fmul_rn(v.w, mk.w), p.in_mult);
          }
          const size_t off = ((size_t)kc * p.a_pitch + r0 + 16 * i) * 16;
          *reinterpret_cast<uint4*>(sa + off) = make_uint4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
          if constexpr (LRT) {
            if (p.x2)
              *reinterpret_cast<uint4*>(sa2 + off) = make_uint4(tf32_rna(__uint_as_float(areg2[i].x)), tf32_rna(__uint_as_float(areg2[i].y)),
                                                                tf32_rna(__uint_as_float(areg2[i].z)), tf32_rna(__uint_as_float(areg2[i].w)));
            else
              *reinterpret_cast<uint4*>(sa2 + off) = make_uint4(tf32_rna(v.x * v.x), tf32_rna(v.y * v.y), tf32_rna(v.z * v.z), tf32_rna(v.w * v.w));
          }
        }
      }
    };

    int stage = 0;
    uint32_t phase = 0;
    if (!a_async) load_a_regs(0);
    for (int kb = 0; kb < num_kb; ++kb) {
      if (lane == 0) mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);   // one lane polls for the warp
      __syncwarp();
      uint8_t* sa = ring + (size_t)stage * stage_bytes;
      uint8_t* sa2 = sa + a_bytes;                              // LRT only
      uint8_t* sb = sa + (LRT ? 2 : 1) * a_bytes;
      uint8_t* sb2 = sb + b_bytes;                              // LRT only
      const int k = kb * BLOCK_K + kc * EPC;                    // first K element of this chunk
      // ---- A ---------------------------------------------------------------------------------------
      if (a_async) {
        int c, dr, ds;
        tap_of(k, c, dr, ds);
        const bool kv = k < p.K;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          int hi = rh0[i] + dr, wi = rw0[i] + ds;
          bool ok = kv && hi >= 0 && wi >= 0;
          if (p.tr_sh > 1 || p.tr_sw > 1) {
            ok = ok && (hi % p.tr_sh == 0) && (wi % p.tr_sw == 0);
            hi /= p.tr_sh; wi /= p.tr_sw;
          }
          ok = ok && hi < p.H && wi < p.W;
          const float* src = reinterpret_cast<const float*>(xs) + (ok ? rbase[i] + (hi * p.W + wi) * p.C + c : 0);
          cp_async16(smem_u32(sa + ((size_t)kc * p.a_pitch + r0 + 16 * i) * 16), src, ok ? 16u : 0u);
        }
      } else {
        store_a_regs(kb, sa, sa2);
        if (kb + 1 < num_kb) load_a_regs(kb + 1);               // prefetch the next stage into registers
      }
      // ---- B: weights [N][K] (per-sample output of the sampling kernel, L2 resident) -----------------
      for (int n = r0; n < p.n_pad; n += 16) {
        const uint32_t dst = smem_u32(sb + ((size_t)kc * p.b_pitch + n) * 16);
        if (I8 && n == p.N) {
          // extra all-ones row: D[:, N] = sum_k (x - z_x), the row sums needed for the z_w correction
          const uint32_t one = k < p.K ? 0x01010101u : 0u, two = (k + 8 < p.K) ? 0x01010101u : 0u;
          *reinterpret_cast<uint4*>(sb + ((size_t)kc * p.b_pitch + n) * 16) = make_uint4(one, one, two, two);
          continue;
        }
        const bool ok = n < p.N && k < p.K;
        const uint8_t* src = ws + (ok ? ((size_t)n * p.K + k) * ESZ : 0);
        if constexpr (I8) {
          // rows are only 8-byte aligned when K % 16 == 8 (e.g. C = 24): two 8-byte LDGSTS
          cp_async8(dst, src, ok ? 8u : 0u);
          const bool ok2 = ok && k + 8 < p.K;
          cp_async8(dst + 8, ok2 ? src + 8 : ws, ok2 ? 8u : 0u);
        } else {
          cp_async16(dst, src, ok ? 16u : 0u);
          if constexpr (LRT) cp_async16(smem_u32(sb2 + ((size_t)kc * p.b_pitch + n) * 16), ws2 + (ok ? ((size_t)n * p.K + k) * ESZ : 0), ok ? 16u : 0u);
        }
      }
      fence_proxy_async();            // this thread's st.shared -> visible to the tensor core (async proxy)
      cp_async_arrive_noinc(smem_u32(&full_bar[stage]));   // arrives once this thread's cp.asyncs have landed
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else {
    // =========================== MMA ISSUER (warp 4) =============================================
    // One elected thread runs the whole loop: inside an elect.sync region ptxas keeps the per-MMA descriptors in uniform
    // registers (an `if (lane == 0)` region made it wrap every MMA in an ELECT / R2UR.BROADCAST waterfall loop).
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t lbo_a = (uint32_t)p.a_pitch * 16, lbo_b = (uint32_t)p.b_pitch * 16;
      const uint64_t adesc_hi = make_smem_desc(0, lbo_a, 128), bdesc_hi = make_smem_desc(0, lbo_b, 128);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        fence_proxy_async();            // cp.async / st.shared (generic proxy) data -> ordered before the MMAs' async-proxy reads
        tc_fence_after();
        const uint32_t sa = smem_u32(ring + (size_t)stage * stage_bytes);
        const uint32_t sa2 = sa + a_bytes;
        const uint32_t sb = sa + (LRT ? 2 : 1) * a_bytes;
        const uint32_t sb2 = sb + b_bytes;
        const int krem = p.K - kb * BLOCK_K;
        const int nmma = krem >= BLOCK_K ? KCH / 2 : (krem + MMA_K - 1) / MMA_K;
#pragma unroll 1
        for (int j = 0; j < nmma; ++j) {
          const uint32_t acc = (kb > 0 || j > 0) ? 1u : 0u;
          const uint64_t ad = adesc_hi | (uint64_t)(((sa + 2 * j * lbo_a) >> 4) & 0x3FFF);
          const uint64_t bd = bdesc_hi | (uint64_t)(((sb + 2 * j * lbo_b) >> 4) & 0x3FFF);
          umma_mma<MODE>(tmem_base, ad, bd, p.idesc, acc);
          if constexpr (LRT) {
            const uint64_t ad2 = adesc_hi | (uint64_t)(((sa2 + 2 * j * lbo_a) >> 4) & 0x3FFF);
            const uint64_t bd2 = bdesc_hi | (uint64_t)(((sb2 + 2 * j * lbo_b) >> 4) & 0x3FFF);
            umma_mma<MODE>(tmem_base + (uint32_t)p.n_pad, ad2, bd2, p.idesc, acc);
          }
        }
        umma_commit(smem_u32(&empty_bar[stage]));            // slot free once these MMAs retire
        if (kb == num_kb - 1) umma_commit(smem_u32(accum_bar));  // accumulators complete
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  }

  // =============================== EPILOGUE (warps 0-3) ==========================================
  if (warp < 4) {
    if (lane == 0) mbar_wait(smem_u32(accum_bar), 0);
    __syncwarp();
    tc_fence_after();
    const int m = m0 + warp * 32 + lane;        // TMEM lane == tile row
    const bool mv = m < p.M;
    const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int Nrow = p.n_split ? p.n_split : (p.out_ld ? p.out_ld : p.N);      // channels per stored row
    size_t prow = (size_t)z * p.M + (mv ? m : 0);          // pixel row of the output tensor
    if (p.oph | p.opw) {
      const int mm = mv ? m : 0;
      const int wo = mm % p.Wo, t2 = mm / p.Wo, ho = t2 % p.Ho, b = t2 / p.Ho;
      // planar-C4 maps share their zeros: borders on top / left only (p4_layout.cuh); NHWC bordered maps have them all around
      const bool p4 = p.flags & QBN_FLAG_OUT_P4;
      const int Hop = p.Ho + (p4 ? 1 : 2) * p.oph, Wop = p.Wo + (p4 ? 1 : 2) * p.opw;
      prow = (((size_t)z * p.B + b) * Hop + ho + p.oph) * Wop + wo + p.opw;
    }
    const size_t orow = prow * Nrow;
    const bool out_p4 = (MODE == MODE_EVAL) && (p.flags & QBN_FLAG_OUT_P4);
    int rowsum = 0;
    if constexpr (I8) {
      uint32_t v[8];
      tmem_ld8(tlane + (uint32_t)(p.N & ~7), v);   // column N holds sum_k (x - z_x)
      tmem_ld_wait();
      rowsum = (int)v[p.N & 7];
    }
    for (int c0 = 0; c0 < p.N; c0 += 8) {
      uint32_t v[8], v2[8];
      tmem_ld8(tlane + (uint32_t)c0, v);
      if constexpr (LRT) tmem_ld8(tlane + (uint32_t)(p.n_pad + c0), v2);
      tmem_ld_wait();
      if (!mv) continue;
      const int nvalid = min(8, p.N - c0);
      if constexpr (MODE == MODE_EVAL) {
        float* out = reinterpret_cast<float*>(p.out);
        // sample-stacked columns (shared input, first layer): column c = sample (c / n_split), channel (c % n_split)
        const int ch0 = p.n_split ? c0 % p.n_split : c0;
        const size_t obase = p.n_split ? orow + (size_t)(c0 / p.n_split) * (size_t)p.sample_out_stride : orow;
        // planar-C4 output (p4_layout.cuh): channel c of pixel row r lives at ((c/4) * out_plane + r) * 4 + c%4
        const size_t prow_s = p.n_split ? prow + (size_t)(c0 / p.n_split) * (size_t)(p.sample_out_stride / p.n_split) : prow;
        auto p4_index = [&](int c) { return ((size_t)(c >> 2) * (size_t)p.out_plane + prow_s) * 4 + (size_t)(c & 3); };
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a = __uint_as_float(v[j]);
          if (j < nvalid) {
            if (p.scale) a = __fmul_rn(a, __ldg(p.scale + ch0 + j));
            if (p.shift) a = __fadd_rn(a, __ldg(p.shift + ch0 + j));
            if (p.mul2x) a = __fmul_rn(a, 2.0f * __ldg(p.mul2x + obase + ch0 + j));
            if (p.residual) a = __fadd_rn(a, __ldg(p.residual + (out_p4 ? p4_index(ch0 + j) : obase + ch0 + j)));
            if (p.flags & QBN_FLAG_RELU) a = fmaxf(a, 0.f);
            if (p.flags & QBN_FLAG_OUT_ROUND_TF32) a = __uint_as_float(tf32_rna(a));
          }
          o[j] = a;
        }
        if (out_p4) {
          *reinterpret_cast<float4*>(out + p4_index(ch0)) = make_float4(o[0], o[1], o[2], o[3]);
          if (nvalid > 4) *reinterpret_cast<float4*>(out + p4_index(ch0 + 4)) = make_float4(o[4], o[5], o[6], o[7]);
        } else if (nvalid == 8 && ((obase + ch0) & 3) == 0) {
          *reinterpret_cast<float4*>(out + obase + ch0) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(out + obase + ch0 + 4) = make_float4(o[4], o[5], o[6], o[7]);
        } else {
          for (int j = 0; j < nvalid; ++j) out[obase + ch0 + j] = o[j];
        }
      } else if constexpr (LRT) {
        float* out = reinterpret_cast<float*>(p.out);
        if (p.mul2x) {     // dgrad: dx = conv(g, mu') + 2x .* conv(dv, sigma2'), both contractions of this one launch
          for (int j = 0; j < nvalid; ++j)
            out[orow + c0 + j] = __fadd_rn(__uint_as_float(v[j]), __fmul_rn(2.0f * __ldg(p.mul2x + orow + c0 + j), __uint_as_float(v2[j])));
          continue;
        }
        float e[8];
        if (p.eps) {
          for (int j = 0; j < 8; ++j) e[j] = j < nvalid ? __ldg(p.eps + orow + c0 + j) : 0.f;
        } else {
          for (int j = 0; j < 8; ++j) e[j] = j < nvalid ? philox_normal1(p.seed, p.sa, p.sb + (p.sbase ? *p.sbase : 0u), (uint64_t)(orow + c0 + j)) : 0.f;
        }
        for (int j = 0; j < nvalid; ++j) {
          float sd = sqrtf(1e-8f + __uint_as_float(v2[j]));
          float r = __fadd_rn(__uint_as_float(v[j]), __fmul_rn(sd, e[j]));
          if (p.bias) r = __fadd_rn(r, __ldg(p.bias + c0 + j));
          out[orow + c0 + j] = r;
          if (p.std_out) p.std_out[orow + c0 + j] = sd;
        }
      } else {
        uint8_t* out = reinterpret_cast<uint8_t*>(p.out);
        for (int j = 0; j < nvalid; ++j) {
          // sum (x-z_x)(w-z_w) = sum (x-z_x) w  -  z_w * sum (x-z_x)
          int acc = (int)v[j] - p.z_w * rowsum;
          float xf = (float)acc;
          if (p.bias) xf = __fadd_rn(xf, __fdiv_rn(__ldg(p.bias + c0 + j), p.atw));
          int q = (int)rintf(__fmul_rn(xf, p.mult)) + p.z_out;
          out[orow + c0 + j] = (uint8_t)max(p.lo, min(p.hi, q));
          if (p.acc_dump) p.acc_dump[orow + c0 + j] = acc;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

static int pow2_cols(int c) {
  int v = 32;
  while (v < c) v <<= 1;
  return v;
}

template <int MODE>
static int launch_umma(UParams& p, int n_samples, cudaStream_t st, const char* who) {
  constexpr bool LRT = MODE == MODE_LRT;
  constexpr bool I8 = MODE == MODE_I8;
  const int n_eff = I8 ? p.N + 1 : p.N;            // int8: one extra all-ones row (row sums)
  p.n_pad = (n_eff + 15) / 16 * 16;
  if (p.n_pad > 256 || (LRT && 2 * p.n_pad > 512)) {
    qbn_set_error("%s: N=%d too wide for one TMEM tile", who, p.N);
    return QBN_ERR_UNSUPPORTED;
  }
  p.acc_cols = (LRT ? 2 : 1) * p.n_pad;
  p.tmem_cols = pow2_cols(p.acc_cols);
  p.a_pitch = UM + 1;                              // +1 chunk: conflict-free STS for the (row, kc) thread map
  p.b_pitch = p.n_pad + 1;
  // instruction descriptor (cute::UMMA::InstrDescriptor): c_format[4,6) a_format[7,10) b_format[10,13)
  // a/b major = K (0), n_dim[17,23) = N>>3, m_dim[24,29) = M>>4
  if (I8)
    p.idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_pad >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);  // S32, S8 x S8
  else
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.n_pad >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);  // F32, TF32 x TF32
  const size_t stage_bytes = (size_t)(LRT ? 2 : 1) * KCH * 16 * (p.a_pitch + p.b_pitch);
  constexpr int BLOCK_K = KCH * (I8 ? 16 : 4);
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  // The kernel is not persistent and its epilogue does not overlap its main loop, so co-resident CTAs are what keeps an SM busy:
  // size the ring for TWO CTAs per SM when at least two stages fit in half the shared memory (else one CTA with a deeper ring).
  int stages = (int)((110 * 1024) / stage_bytes);
  if (tune_env("QBN_V1_DEEP") || stages < 2) stages = (int)((200 * 1024) / stage_bytes);
  if (stages > 6) stages = 6;
  if (stages > num_kb) stages = num_kb;
  if (stages < 1) {
    qbn_set_error("%s: stage does not fit shared memory", who);
    return QBN_ERR_UNSUPPORTED;
  }
  p.stages = stages;
  const size_t smem = stages * stage_bytes + (2 * stages + 1) * 8 + 16;
  static bool attr_set[3] = {false, false, false};
  if (!attr_set[MODE]) {
    QBN_CUDA(cudaFuncSetAttribute(umma_conv_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set[MODE] = true;
  }
  dim3 grid((unsigned)((p.M + UM - 1) / UM), 1, (unsigned)n_samples);
  umma_conv_kernel<MODE><<<grid, NTHREADS, smem, st>>>(p);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

static void fill_geom(UParams& p, const qbn_conv_desc* d) {
  memset(&p, 0, sizeof(p));
  p.B = d->B; p.H = d->H; p.W = d->W; p.C = d->C; p.N = d->N; p.R = d->R; p.S = d->S;
  p.sh = d->stride_h; p.sw = d->stride_w; p.ph = d->pad_h; p.pw = d->pad_w; p.dh = d->dil_h; p.dw = d->dil_w;
  p.Ho = d->Ho; p.Wo = d->Wo; p.K = d->R * d->S * d->C;
  p.M = d->B * d->Ho * d->Wo;
  p.oph = d->out_pad_h; p.opw = d->out_pad_w;
}

}  // namespace

int qbn_umma_lrt_fwd(const qbn_conv_desc* d, const float* x, const float* mu_p, const float* sig2_p, const float* bias,
                     const float* eps, uint64_t seed, uint32_t sa, uint32_t sb, float* out, float* std_out, cudaStream_t st) {
  if (d->C % 4 != 0) {
    qbn_set_error("qbn_lrt_fwd(TF32): C=%d must be a multiple of 4 (pad the input channels)", d->C);
    return QBN_ERR_UNSUPPORTED;
  }
  UParams p;
  fill_geom(p, d);
  p.x = x; p.w = mu_p; p.w2 = sig2_p; p.x_shared = 1; p.w_shared = 1;
  p.bias = bias; p.eps = eps; p.seed = seed; p.sa = sa; p.sb = sb; p.sbase = qbn_sample_base_ptr(); p.out = out; p.std_out = std_out;
  return launch_umma<MODE_LRT>(p, 1, st, "qbn_lrt_fwd(TF32)");
}

int qbn_umma_conv_fwd(const qbn_conv_desc* d, int n_samples, int x_shared, const float* x, const float* w, int w_shared,
                      const float* scale, const float* shift, const float* residual, int flags, const float* in_mask,
                      float in_mult, float* out, cudaStream_t st) {
  if (d->C % 4 != 0) {
    qbn_set_error("qbn_conv_fwd(TF32): C=%d must be a multiple of 4 (pad the input channels)", d->C);
    return QBN_ERR_UNSUPPORTED;
  }
  UParams p;
  fill_geom(p, d);
  p.x = x; p.w = w; p.x_shared = x_shared; p.w_shared = w_shared;
  p.scale = scale; p.shift = shift; p.residual = residual; p.flags = flags; p.in_mask = in_mask; p.in_mult = in_mult;
  p.out = out;
  if (flags & QBN_FLAG_OUT_P4) {
    if (d->N % 4 != 0) {
      qbn_set_error("qbn_conv_fwd: planar-C4 output needs N %% 4 == 0 (N=%d)", d->N);
      return QBN_ERR_UNSUPPORTED;
    }
    const long long hop = d->Ho + d->out_pad_h, wop = d->Wo + d->out_pad_w;
    p.out_plane = (long long)n_samples * d->B * hop * wop + (long long)d->out_pad_h * wop + d->out_pad_w;     // maps + zero tail
  }
  // Shared input (first layer): stack the samples' weights along N — [S][N][K] IS an [S*N][K] matrix — so
  // the input tile is staged once for all samples and one accumulator tile holds every sample's channels.
  if (x_shared && !w_shared && n_samples > 1 && d->N % 8 == 0 && n_samples * d->N <= 256 && !residual && !in_mask) {
    p.n_split = d->N;
    p.N = n_samples * d->N;
    const int bmul = (flags & QBN_FLAG_OUT_P4) ? 1 : 2;
    p.sample_out_stride = (long long)d->B * (d->Ho + bmul * d->out_pad_h) * (d->Wo + bmul * d->out_pad_w) * d->N;
    return launch_umma<MODE_EVAL>(p, 1, st, "qbn_conv_fwd(TF32, sample-stacked)");
  }
  // Wide linear layers (N > 256, e.g. LeNet's 2450 -> 500): one accumulator tile holds at most 256 columns, so the output
  // channels are split over several launches that write column ranges of the same rows.
  if (d->N > 256) {
    if (!(d->R == 1 && d->S == 1 && d->H == 1 && d->W == 1 && d->out_pad_h == 0 && d->out_pad_w == 0) || (flags & QBN_FLAG_OUT_P4)) {
      qbn_set_error("qbn_conv_fwd(TF32): N=%d > 256 is supported for linear (1x1 on a 1x1 map) geometry only", d->N);
      return QBN_ERR_UNSUPPORTED;
    }
    for (int n0 = 0; n0 < d->N; n0 += 256) {
      UParams q = p;
      q.N = d->N - n0 < 256 ? d->N - n0 : 256;
      q.w = w + (size_t)n0 * q.K;
      q.w_ld = (long long)d->N * q.K;
      q.out = out + n0;
      q.out_ld = d->N;
      if (scale) q.scale = scale + n0;
      if (shift) q.shift = shift + n0;
      if (residual) q.residual = residual + n0;
      int rc = launch_umma<MODE_EVAL>(q, n_samples, st, "qbn_conv_fwd(TF32, N-split)");
      if (rc != QBN_OK) return rc;
    }
    return QBN_OK;
  }
  return launch_umma<MODE_EVAL>(p, n_samples, st, "qbn_conv_fwd(TF32)");
}

// A3 dx on the tensor cores (undilated layers): the transposed convolution is the forward kernel run on the output gradient
// with flipped, transposed weights;  dx = conv(g, mu') + 2x .* conv(dv, sigma2')  as ONE launch of the dual-accumulator kernel.
int qbn_umma_lrt_dgrad(const qbn_conv_desc* d, const float* g, const float* dv, const float* mu_t, const float* sig2_t, const float* x,
                       float* dx, cudaStream_t st) {
  qbn_conv_desc t;
  memset(&t, 0, sizeof(t));
  t.B = d->B; t.H = d->Ho; t.W = d->Wo; t.C = d->N; t.N = d->C; t.R = d->R; t.S = d->S;
  t.stride_h = t.stride_w = 1; t.dil_h = t.dil_w = 1;
  t.pad_h = d->R - 1 - d->pad_h; t.pad_w = d->S - 1 - d->pad_w;
  t.Ho = d->H; t.Wo = d->W;                       // rows of the GEMM = input pixels of the forward conv
  UParams p;
  fill_geom(p, &t);
  p.tr_sh = d->stride_h; p.tr_sw = d->stride_w;   // strided forward conv: its gradient lives on the stride grid (transposed conv)
  // ONE launch of the dual-accumulator (LRT) kernel: A = g and dv, B = mu' and sigma2', epilogue dx = acc1 + 2x .* acc2
  p.x = g; p.x2 = dv; p.w = mu_t; p.w2 = sig2_t; p.x_shared = 1; p.w_shared = 1; p.out = dx; p.mul2x = x;
  return launch_umma<MODE_LRT>(p, 1, st, "qbn_lrt_bwd(TF32 dgrad)");
}

int qbn_umma_i8_conv_fwd(const qbn_conv_desc* d, int n_samples, int x_shared, const uint8_t* x, int z_x, const int8_t* w,
                         int w_shared, int z_w, const float* bias, float act_times_w, float mult, int z_out, int lo, int hi,
                         uint8_t* out, int32_t* acc_dump, cudaStream_t st) {
  if (d->C % 8 != 0 || z_x < 0 || z_x > 127) {
    qbn_set_error("qbn_i8_conv_fwd(tcgen05): needs C %% 8 == 0 and 0 <= z_x <= 127 (C=%d, z_x=%d)", d->C, z_x);
    return QBN_ERR_UNSUPPORTED;
  }
  UParams p;
  fill_geom(p, d);
  p.x = x; p.w = w; p.x_shared = x_shared; p.w_shared = w_shared;
  p.z_x = z_x; p.z_w = z_w; p.bias = bias; p.atw = act_times_w; p.mult = mult; p.z_out = z_out; p.lo = lo; p.hi = hi;
  p.out = out; p.acc_dump = acc_dump;
  return launch_umma<MODE_I8>(p, n_samples, st, "qbn_i8_conv_fwd(tcgen05)");
}
