// LRT weight gradients on tcgen05:
//
//   dmu   [n][tap][c] = sum_q g [q][n] * x   [strip(tap)][q + shift(tap)][c]
//   dsig2 [n][tap][c] = sum_q dv[q][n] * x^2 [strip(tap)][q + shift(tap)][c]         q = padded pixel of the OUTPUT maps
//
// i.e. per tap a GEMM whose reduction dimension is the PIXEL, so both operands are MN-major (channels contiguous, pixels = K).
// For kind::tf32 the tensor core accepts MN-major operands in exactly one shared-memory layout (measured with scripts/dbg/mn_probe.py,
// and what CUTLASS's sm100 builder states): SWIZZLE_128B_BASE32B — rows of 128 bytes = 32 consecutive channels of one pixel, consecutive
// pixels 128 bytes apart, the 32-byte chunk index XORed with bits [7,9) of the ABSOLUTE shared-memory address, 4-row groups SBO apart,
// 32-channel blocks LBO apart.  (The no-swizzle MN-major descriptor silently yields zeros.)  The operands therefore come in the
// "W32" layout, [C/32 blocks][rows][32 floats] with the chunks of row r pre-XORed by r & 3 in GLOBAL memory (qbn_w32_from_p4): a tile's
// rows are then ONE bulk copy per block, placed so that (shared row & 3) == (global row & 3), and a tap is again a row shift of the
// descriptor's start address inside one image of the x rows — the probe confirms that the XOR key follows the absolute address, so
// any row shift is valid with base_offset 0.  Nothing is transposed or gathered by threads.
//
// Work item (one CTA) = (mean | variance) x (block of <= 128 output channels) x (group of <= 3 input-channel blocks) x (tap group) x
// (pixel range); its partial D stays in TMEM over the whole pixel range and is added to the result with fp32 atomics once.
//   warp 5, one lane : bulk copies of the g/dv tile (128 pixels, its 32-channel blocks) and the x/x^2 rows (+ halo)
//   warp 4, one lane : tcgen05.mma kind::tf32, A and B MN-major, K = 8 pixels per instruction
//   warps 0-3        : final epilogue (TMEM lane = output channel n): red.global.add of the [n][tap][c] partial
#include <stdlib.h>
#include <string.h>
#include "p4_layout.cuh"
#include "umma_common.cuh"

namespace {

constexpr int WG_TM = 128;          // pixels per tile
constexpr int WG_THREADS = 192;
constexpr int WG_MAX_TAPS = 25;
constexpr int WG_ST = 2;            // pipeline stages (1 when two do not fit: the stride-2 layers stage four phase strips)

struct WGParams {
  int Qs, n_tiles;
  int N_out, C_real, taps;
  int n_blk_a_total, n_blk_b_total;  // 32-channel blocks of g / x
  int n_strips, d_before, RA, RA_pad;
  long long strip_rows;
  int tap_off[WG_MAX_TAPS];         // rows inside a B block image: strip region + alignment pad + d_before + shift
  int strip_pad[4];                 // (first needed global row of the strip) & 3 for tiles q0 = 0 mod 4
  int n_mb, n_cb, n_tg, n_ps;       // output-channel blocks (128), input-channel block groups (NB), tap groups, pixel splits
  int NB, TPG, n_cols;              // 32-channel blocks per MMA N, taps per group, MMA N = 32 * NB
  uint32_t a_bytes, b_bytes, b_blk_bytes, b_strip_bytes, idesc;
  int tmem_cols, ST;
  int stack, stack_bw, stack_rows;   // stack = S > 0: the S column shifts of a filter row stacked along N (stride 1, one 32-channel input block)
  uint32_t b_lbo;                    // bytes between the 32-channel blocks of the B operand
  const float* g; const float* dv; long long g_plane;       // W32 tensors: rows per block plane
  const float* x; const float* xsq; long long x_plane;
  float* dmu; float* dsig2;
};

// SWIZZLE_128B_BASE32B descriptor (layout type 1), MN-major: LBO = bytes between 32-channel blocks, SBO = bytes between 4-row groups
QBN_DEVINL uint64_t make_mn_desc(uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return make_smem_desc(0, lbo_bytes, sbo_bytes) | ((uint64_t)1 << 61);
}

__global__ void __launch_bounds__(WG_THREADS, 1) umma_wgrad_p4_kernel(const __grid_constant__ WGParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* a_ring = smem;                                          // [WG_ST][a_bytes], a_bytes a multiple of 16 KB
  uint8_t* b_ring = smem + (size_t)p.ST * p.a_bytes;               // [ST][b_bytes], b_bytes a multiple of 512
  __shared__ uint64_t bars[2 * WG_ST + 1];
  __shared__ uint32_t tmem_slot;
  uint64_t* full = bars;
  uint64_t* empty = bars + WG_ST;
  uint64_t* done = bars + 2 * WG_ST;
  // decode the work item
  int it = blockIdx.x;
  const int ps = it % p.n_ps; it /= p.n_ps;
  const int tg = it % p.n_tg; it /= p.n_tg;
  const int cb = it % p.n_cb; it /= p.n_cb;
  const int mb = it % p.n_mb; it /= p.n_mb;
  const int kind = it;                                           // 0: (g, x) -> dmu   1: (dv, x^2) -> dsig2
  const float* A = kind ? p.dv : p.g;
  const float* Bx = kind ? p.xsq : p.x;
  float* out = kind ? p.dsig2 : p.dmu;
  const int a_blk0 = mb * 4, a_load = min(4, p.n_blk_a_total - a_blk0);
  const int b_blk0 = cb * p.NB, b_load = min(p.NB, p.n_blk_b_total - b_blk0);
  const int n_acc = p.stack ? p.taps / p.stack : p.taps;          // accumulators: filter rows when the columns are stacked, else taps
  const int t_begin = tg * p.TPG, t_end = min(n_acc, t_begin + p.TPG);
  const int tile_begin = (int)(((long long)p.n_tiles * ps) / p.n_ps), tile_end = (int)(((long long)p.n_tiles * (ps + 1)) / p.n_ps);

  // stale shared memory must not hold NaNs where zeros are expected (rows in front of the first map, blocks beyond the tensor: they
  // only feed accumulator rows / columns nobody reads, but 0 * NaN inside a read row would poison it)
  const uint32_t dyn_bytes = (uint32_t)p.ST * (p.a_bytes + p.b_bytes);
  for (uint32_t i = tid; i < dyn_bytes / 16; i += WG_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < WG_ST; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
    mbar_init(smem_u32(done), 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  if (warp == 4) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 5) {
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const long long q0 = (long long)tile * WG_TM;
        mbar_wait(smem_u32(&empty[st]), ph ^ 1);
        const uint32_t bar = smem_u32(&full[st]);
        const uint32_t a_slot = smem_u32(a_ring + (size_t)st * p.a_bytes), b_slot = smem_u32(b_ring + (size_t)st * p.b_bytes);
        if (p.stack) {
          // block s = the rows shifted by (s - bw) pixels: rows [q0 - bh*Wp + s - bw, ... + stack_rows); blocks are b_lbo = RA_pad*128 + 128
          // bytes apart, so that block s starts one row further (mod 4) and the XOR key of every block follows its global rows
          uint32_t tx = (uint32_t)a_load * WG_TM * 128;
          long long los[8];
          for (int sft = 0; sft < p.stack; ++sft) {
            const long long gs = q0 - p.d_before + p.stack_bw + (sft - p.stack_bw);
            los[sft] = gs < 0 ? 0 : gs;
            tx += (uint32_t)(gs + p.stack_rows - los[sft]) * 128;
          }
          mbar_arrive_expect_tx(bar, tx);
          for (int j = 0; j < a_load; ++j)
            bulk_load_g2s(a_slot + (uint32_t)j * WG_TM * 128, A + ((size_t)(a_blk0 + j) * p.g_plane + q0) * 32, WG_TM * 128, bar);
          for (int sft = 0; sft < p.stack; ++sft) {
            const long long gs = q0 - p.d_before + sft;
            bulk_load_g2s(b_slot + (uint32_t)sft * p.b_lbo + (uint32_t)(p.strip_pad[0] + (int)(los[sft] - gs)) * 128,
                          Bx + ((size_t)b_blk0 * p.x_plane + los[sft]) * 32, (uint32_t)(gs + p.stack_rows - los[sft]) * 128, bar);
          }
          if (++st == p.ST) { st = 0; ph ^= 1; }
          continue;
        }
        // x rows [g0, g0 + RA) of every strip, clamped at the tensor's first row; shared row = strip_pad + (global row - g0)
        const long long g0 = q0 - p.d_before;
        const long long lo = g0 < 0 ? 0 : g0;
        const uint32_t b_rows = (uint32_t)(g0 + p.RA - lo);
        mbar_arrive_expect_tx(bar, (uint32_t)a_load * WG_TM * 128 + (uint32_t)(b_load * p.n_strips) * b_rows * 128);
        for (int j = 0; j < a_load; ++j)
          bulk_load_g2s(a_slot + (uint32_t)j * WG_TM * 128, A + ((size_t)(a_blk0 + j) * p.g_plane + q0) * 32, WG_TM * 128, bar);
        for (int s2 = 0; s2 < p.n_strips; ++s2)
          for (int j = 0; j < b_load; ++j)
            bulk_load_g2s(b_slot + (uint32_t)s2 * p.b_strip_bytes + (uint32_t)j * p.b_blk_bytes + (uint32_t)(p.strip_pad[s2] + (int)(lo - g0)) * 128,
                          Bx + ((size_t)(b_blk0 + j) * p.x_plane + (size_t)s2 * p.strip_rows + lo) * 32, b_rows * 128, bar);
        if (++st == p.ST) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 4) {
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0;
      const uint64_t adesc_hi = make_mn_desc(WG_TM * 128, 512), bdesc_hi = make_mn_desc(p.b_lbo, 512);
      bool first = true;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        mbar_wait(smem_u32(&full[st]), ph);
        tc_fence_after();
        const uint32_t a16 = smem_u32(a_ring + (size_t)st * p.a_bytes) >> 4, b16 = smem_u32(b_ring + (size_t)st * p.b_bytes) >> 4;
#pragma unroll 1
        for (int t = t_begin; t < t_end; ++t) {
          const uint32_t tcol = tmem_base + (uint32_t)((t - t_begin) * p.n_cols);
          const uint32_t bt = b16 + (uint32_t)p.tap_off[t] * 8;                  // rows of 128 bytes = 8 16-byte units
#pragma unroll 4
          for (int kb = 0; kb < WG_TM / 8; ++kb) {                              // 8 pixels = 8 rows = 1024 bytes per K step
            const uint64_t ad = adesc_hi | (uint64_t)((a16 + kb * 64) & 0x3FFF), bd = bdesc_hi | (uint64_t)((bt + kb * 64) & 0x3FFF);
            if (first && kb == 0) umma_mma_c<MODE_EVAL, false>(tcol, ad, bd, p.idesc);
            else umma_mma_c<MODE_EVAL, true>(tcol, ad, bd, p.idesc);
          }
        }
        first = false;
        umma_commit(smem_u32(&empty[st]));
        if (++st == p.ST) { st = 0; ph ^= 1; }
      }
      umma_commit(smem_u32(done));
    }
    __syncwarp();
  } else {
    // ---- final epilogue: lane = output channel of this block
    if (tile_end > tile_begin) {
      if (lane == 0) mbar_wait(smem_u32(done), 0);
      __syncwarp();
      tc_fence_after();
      const int n = mb * 128 + tid;                              // tid in [0, 128)
      const bool mine = n < p.N_out;
      const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16);
      const int K = p.taps * p.C_real;
      const int c_base = b_blk0 * 32;
      const bool vec4 = (p.C_real & 3) == 0;                     // rows of the result are 16-byte aligned
      const int n_cg = p.stack ? 32 * p.stack : 32 * b_load;
      for (int t = t_begin; t < t_end; ++t) {
        for (int cg = 0; cg < n_cg; cg += 16) {
          uint32_t v[16];
          tmem_ld16(tlane + (uint32_t)((t - t_begin) * p.n_cols + cg), v);
          tmem_ld_wait();
          // stacked: accumulator t = filter row, column block cg / 32 = filter column; else accumulator t = tap, columns = channels
          const int tap = p.stack ? t * p.stack + (cg >> 5) : t;
          const int c0 = p.stack ? (cg & 31) : c_base + cg;
          if (mine) {
            float* dst = out + (size_t)n * K + (size_t)tap * p.C_real + c0;
            if (vec4 && c0 + 16 <= p.C_real) {
              // 16-byte vector reductions (sm_90+): a quarter of the atomic instructions, 16 bytes per sector instead of 4
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])),
                             "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                             : "memory");
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < p.C_real) atomicAdd(dst + j, __uint_as_float(v[j]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// planar C4 [chunks][plane_rows][4] -> W32 [ceil(chunks / 8)][plane_rows][32], the 32-byte chunks of row r XORed with r & 3
__global__ void w32_from_p4_kernel(const float4* __restrict__ s0, const float4* __restrict__ s1, int chunks, long long plane_rows, float4* __restrict__ d0,
                                   float4* __restrict__ d1) {
  // one thread = one 16-byte piece of a W32 row: a warp writes four whole rows (512 contiguous bytes) and reads 64 contiguous bytes
  // (4 rows) from each of 8 planes
  const int n_blk = (chunks + 7) / 8;
  const long long total = (long long)n_blk * plane_rows * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int piece = (int)(i & 7);
    const long long br = i >> 3;                                  // blk * plane_rows + r
    const int blk = (int)(br / plane_rows);
    const long long r = br - (long long)blk * plane_rows;
    const int c8 = (piece >> 1) ^ (int)(r & 3);                   // the 8-channel chunk stored at this position of row r
    const int plane = blk * 8 + c8 * 2 + (piece & 1);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool has = plane < chunks;
    const long long src = (long long)plane * plane_rows + r;
    d0[i] = has ? s0[src] : z;
    if (s1) d1[i] = has ? s1[src] : z;
  }
}

}  // namespace

// Layout conversion for the weight-gradient operands: planar-C4 maps (as staged by qbn_p4_stage_input / qbn_p4_stage_grad, zero
// borders and tail included) -> W32.  src1 / dst1 nullable (a second tensor of the same shape in the same launch).
extern "C" int qbn_w32_from_p4(const float* src0, const float* src1, int C_pad, long long plane_rows, float* dst0, float* dst1, void* stream) {
  QBN_CHECK_ARG(src0 && dst0 && (!src1 || dst1), "null pointer");
  QBN_CHECK_ARG(C_pad > 0 && C_pad % 4 == 0 && plane_rows > 0, "sizes");
  const int chunks = C_pad / 4;
  const long long total = (long long)((chunks + 7) / 8) * plane_rows * 8;
  w32_from_p4_kernel<<<qbn_grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(src0), reinterpret_cast<const float4*>(src1),
                                                                                chunks, plane_rows, reinterpret_cast<float4*>(dst0),
                                                                                reinterpret_cast<float4*>(dst1));
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

// dmu_p / dsig2_p: [N][R*S][C_real] fp32 (the packed OHWI order of qbn_weight_prep / qbn_weight_grad_post), OVERWRITTEN.
// g, dv: W32 maps of the layer's OUTPUT geometry [ceil(N/32)][g_plane_rows][32]; x, x_sq: the layer's input in W32 (phase-split for
// a stride-2 layer, as the forward reads it), [ceil(C/32)][x_plane_rows][32]; channels >= C_real are zero.  Every plane must be
// readable (zeros) for 128 + 2 (Wp + 1) + 8 rows past the last map (ops.lrt_p4_plane_rows).
extern "C" int qbn_lrt_wgrad_p4(int B, int Hp, int Wp, int C_real, int N, int R, int S, int stride, const float* g, const float* dv,
                                long long g_plane_rows, const float* x, const float* x_sq, long long x_plane_rows, float* dmu_p,
                                float* dsig2_p, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  QBN_CHECK_ARG(g && dv && x && x_sq && dmu_p && dsig2_p, "null pointer");
  QBN_CHECK_ARG(B > 0 && Hp > 2 && Wp > 2 && C_real > 0 && N > 0 && R > 0 && S > 0, "sizes");
  const bool s1 = stride == 1 && (R & 1) && (S & 1);
  const bool s2 = stride == 2 && ((R == 3 && S == 3) || (R == 1 && S == 1));
  if (!(s1 || s2) || R * S > WG_MAX_TAPS) {
    qbn_set_error("qbn_lrt_wgrad_p4: needs stride 1 (odd kernel) or stride 2 (3x3 / 1x1) (R=%d S=%d stride=%d)", R, S, stride);
    return QBN_ERR_UNSUPPORTED;
  }
  WGParams p;
  memset(&p, 0, sizeof(p));
  p.Qs = B * Hp * Wp;
  p.n_tiles = (p.Qs + WG_TM - 1) / WG_TM;
  p.N_out = N; p.C_real = C_real; p.taps = R * S;
  p.n_blk_a_total = (N + 31) / 32;
  p.n_blk_b_total = (C_real + 31) / 32;
  const int bh = s1 ? (R - 1) / 2 : 1, bw = s1 ? (S - 1) / 2 : 1;
  int d_after;
  if (s1) { p.n_strips = 1; p.d_before = bh * Wp + bw; d_after = p.d_before; }
  else { p.n_strips = (R == 3) ? 4 : 1; p.d_before = (R == 3) ? Wp + 1 : 0; d_after = 0; }
  p.strip_rows = p.Qs;
  p.RA = WG_TM + p.d_before + d_after;
  p.RA_pad = (p.RA + 3 + 3) / 4 * 4;                              // up to 3 alignment rows in front
  p.b_blk_bytes = (uint32_t)p.RA_pad * 128;
  p.n_mb = (N + 127) / 128;
  const int a_blocks = p.n_blk_a_total < 4 ? p.n_blk_a_total : 4;
  p.a_bytes = (uint32_t)a_blocks * WG_TM * 128;
  // input-channel blocks per MMA: the most (<= 3, N <= 96 columns) whose two stages fit next to the g tiles
  const size_t cap = 224 * 1024;
  size_t smem = 0;
  // stride 1, one input block (C <= 32), all R accumulators of 32 * S columns in TMEM: the S column shifts of a filter row are stacked
  // along N (one MMA covers S taps; its cost is set by the A operand read), each from its own copy of the rows
  if (s1 && p.n_blk_b_total == 1 && S > 1 && S <= 8 && R * 32 * S <= 512) {
    p.stack = S; p.stack_bw = bw;
    p.stack_rows = WG_TM + 2 * bh * Wp;
    p.RA_pad = (p.stack_rows + 3 + 3) / 4 * 4;
    p.b_lbo = (uint32_t)p.RA_pad * 128 + 128;
    p.NB = 1;
    p.b_bytes = ((uint32_t)S * p.b_lbo + 511) / 512 * 512;
    p.b_strip_bytes = p.b_bytes;
    for (p.ST = WG_ST; p.ST >= 1; --p.ST) {
      size_t need = (size_t)p.ST * ((size_t)p.a_bytes + p.b_bytes);
      const size_t window = (size_t)(p.ST - 1) * p.a_bytes + 4 * (size_t)WG_TM * 128;
      if (need < window) need = window;
      if (need <= cap) { smem = need; break; }
    }
    if (!smem) { p.stack = 0; p.RA_pad = (p.RA + 3 + 3) / 4 * 4; }
  }
  if (!p.stack) {
  for (p.ST = WG_ST; p.ST >= 1 && !smem; --p.ST)
    for (p.NB = p.n_blk_b_total < 3 ? p.n_blk_b_total : 3; p.NB >= 1 && !smem; --p.NB) {
      size_t need = (size_t)p.ST * ((size_t)p.a_bytes + (size_t)p.n_strips * p.NB * p.b_blk_bytes);
      // an MMA reads M = 128 rows = four 32-channel blocks LBO apart whatever a_blocks is: the window must stay inside the allocation
      const size_t window = (size_t)(p.ST - 1) * p.a_bytes + 4 * (size_t)WG_TM * 128;
      if (need < window) need = window;
      if (need <= cap) { smem = need; goto fits; }
    }
  qbn_set_error("qbn_lrt_wgrad_p4: tile does not fit shared memory (Wp=%d, %d strips)", Wp, p.n_strips);
  return QBN_ERR_UNSUPPORTED;
fits:
  p.b_strip_bytes = (uint32_t)p.NB * p.b_blk_bytes;
  p.b_bytes = (uint32_t)p.n_strips * p.b_strip_bytes;
  p.b_lbo = p.b_blk_bytes;
  }
  p.n_cb = (p.n_blk_b_total + p.NB - 1) / p.NB;
  p.n_cols = p.stack ? 32 * p.stack : 32 * p.NB;
  const int n_acc = p.stack ? R : p.taps;                      // accumulators: filter rows (stacked) or taps
  p.TPG = 512 / p.n_cols;
  if (p.TPG > n_acc) p.TPG = n_acc;
  p.n_tg = (n_acc + p.TPG - 1) / p.TPG;
  p.TPG = (n_acc + p.n_tg - 1) / p.n_tg;                       // balanced groups
  p.tmem_cols = 32;
  while (p.tmem_cols < p.TPG * p.n_cols) p.tmem_cols <<= 1;
  if (p.stack) {
    // block s holds the rows from q0 - bh*Wp + (s - bw) on: its first row's parity is (s - (bh*Wp + bw)) & 3 and the block itself starts
    // s rows further (mod 4), so one pad serves every block; filter row r starts r * Wp rows into the block
    p.strip_pad[0] = (int)((((-(long long)p.d_before) % 4) + 4) % 4);
    for (int r = 0; r < R; ++r) p.tap_off[r] = p.strip_pad[0] + r * Wp;
  } else {
  for (int s = 0; s < p.n_strips; ++s) {
    const long long first = (long long)s * p.strip_rows - p.d_before;      // first needed global row of strip s for the tile q0 = 0
    p.strip_pad[s] = (int)(((first % 4) + 4) % 4);
  }
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) {
      int strip = 0, sh;
      if (s1) sh = (r - bh) * Wp + (s - bw);
      else if (R == 3) { const int dr = r - 1, ds = s - 1; strip = (dr & 1) * 2 + (ds & 1); sh = (dr < 0 ? -1 : 0) * Wp + (ds < 0 ? -1 : 0); }
      else sh = 0;
      p.tap_off[r * S + s] = strip * p.NB * p.RA_pad + p.strip_pad[strip] + p.d_before + sh;
    }
  }
  const long long need_g = (long long)p.n_tiles * WG_TM, need_x = (long long)p.n_tiles * WG_TM + d_after + 8 + (long long)(p.n_strips - 1) * p.strip_rows;
  if (g_plane_rows < need_g || x_plane_rows < need_x) {
    qbn_set_error("qbn_lrt_wgrad_p4: planes too short (g %lld < %lld or x %lld < %lld rows): allocate a zero tail of 128 + 2 (Wp + 1) + 8 rows", g_plane_rows,
                  need_g, x_plane_rows, need_x);
    return QBN_ERR_INVALID_ARG;
  }
  p.g = g; p.dv = dv; p.g_plane = g_plane_rows; p.x = x; p.xsq = x_sq; p.x_plane = x_plane_rows; p.dmu = dmu_p; p.dsig2 = dsig2_p;
  // F32 += TF32 x TF32, A and B MN-major (bits 15, 16), N at bit 17, M = 128 at bit 24
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.n_cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const int items = 2 * p.n_mb * p.n_cb * p.n_tg;
  // pixel splits: every CTA adds its whole [128][taps][columns] partial to the result at the end, so a split costs TPG * n_cols * 128
  // reductions while a tile costs TPG * 16 MMAs: one wave of CTAs, and at least ~4 tiles per CTA where the layer has them
  p.n_ps = (qbn_sm_count() + items - 1) / items;
  if (p.n_ps > (p.n_tiles + 3) / 4) p.n_ps = (p.n_tiles + 3) / 4;
  if (p.n_ps > p.n_tiles) p.n_ps = p.n_tiles;
  if (p.n_ps < 1) p.n_ps = 1;
  const size_t grad_bytes = sizeof(float) * (size_t)N * p.taps * C_real;
  if (reinterpret_cast<char*>(dmu_p) + grad_bytes == reinterpret_cast<char*>(dsig2_p)) {      // one allocation (ops.lrt_p4_backward): one memset
    QBN_CUDA(cudaMemsetAsync(dmu_p, 0, 2 * grad_bytes, st));
  } else {
    QBN_CUDA(cudaMemsetAsync(dmu_p, 0, grad_bytes, st));
    QBN_CUDA(cudaMemsetAsync(dsig2_p, 0, grad_bytes, st));
  }
  static bool attr_set = false;
  if (!attr_set) {
    QBN_CUDA(cudaFuncSetAttribute(umma_wgrad_p4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    attr_set = true;
  }
  umma_wgrad_p4_kernel<<<items * p.n_ps, WG_THREADS, smem, st>>>(p);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}
