// LRT weight gradients straight from the planar-C4 tensors:
//
//   dmu   [n][tap][c] = sum_q g [q][n] * x   [strip(tap)][q + shift(tap)][c]
//   dsig2 [n][tap][c] = sum_q dv[q][n] * x^2 [strip(tap)][q + shift(tap)][c]         q = padded pixel of the OUTPUT maps
//
// i.e. per tap a GEMM whose reduction dimension is the PIXEL.  A chunk plane of the planar layout — consecutive pixels 16 bytes
// apart, four channels per 16 bytes — is exactly UMMA's MN-major no-swizzle operand layout (8 consecutive K = pixels are one
// 128-byte core matrix, 4-channel groups are SBO = one plane apart), so both operands go to the tensor core as they lie in HBM:
// bulk copies, no transpose, no gather, and a tap is again a row shift of the start address inside one smem image of the x rows.
//
// Work item (one CTA) = (mean | variance) x (block of <= 128 output channels) x (block of input channels) x (tap group) x
// (pixel range); its partial D stays in TMEM over the whole pixel range and is added to the result with fp32 atomics once.
//   warp 5, one lane : bulk copies of the g/dv tile (128 pixels, all its channel planes) and the x/x^2 rows (+ halo)
//   warp 4, one lane : tcgen05.mma kind::tf32, A and B MN-major, K = 8 pixels per instruction
//   warps 0-3        : final epilogue (TMEM lane = output channel n): red.global.add of the [n][tap][c] partial
#include <string.h>
#include "p4_layout.cuh"
#include "umma_common.cuh"

#include <stdlib.h>
namespace {

#ifdef QBN_TUNING
static inline const char* tune_env(const char* name) { return getenv(name); }
#else
static inline const char* tune_env(const char*) { return nullptr; }
#endif

constexpr int WG_TM = 128;          // pixels per tile
constexpr int WG_THREADS = 192;
constexpr int WG_MAX_TAPS = 25;

struct WGParams {
  int Qs, n_tiles;
  int N_out, C, C_real, taps;      // C: channels staged in the planar input (a multiple of 4), C_real <= C: those of the parameter
  int n_strips, d_before, RA_p;
  long long strip_rows;
  int tap_off[WG_MAX_TAPS];         // rows inside the B slot: strip * b_planes * RA_p + d_before + shift
  int n_mb, n_cb, n_tg, n_ps;       // output-channel blocks, input-channel blocks, tap groups, pixel splits
  int CBW, TPG, n_pad;              // channels per input block, taps per group, MMA N (CBW rounded up to 16)
  int a_planes, b_planes;           // chunk planes staged per tile
  uint32_t a_bytes, b_bytes, idesc;
  int tmem_cols, variant;           // variant: descriptor experiments of -DQBN_TUNING builds (QBN_WG_V), 0 in the product
  const float* g; const float* dv; long long g_plane;
  const float* x; const float* xsq; long long x_plane;
  float* dmu; float* dsig2;
};

__global__ void __launch_bounds__(WG_THREADS, 1) umma_wgrad_p4_kernel(const __grid_constant__ WGParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int ST = 2;
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + (size_t)ST * p.a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + (size_t)ST * p.b_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + ST;
  uint64_t* done = bars + 2 * ST;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * ST + 1);
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
  // output-channel block: 128 rows = 32 chunk planes; the last block of a >128-channel layer is shifted back so that it stays inside
  // the tensor (rows it shares with the previous block are skipped in the epilogue)
  const int n_planes = p.N_out / 4;
  int plane0 = mb * 32;
  if (plane0 + 32 > n_planes && n_planes >= 32) plane0 = n_planes - 32;
  const int n_first = mb * 128;                                  // first output channel this block is responsible for
  const int c0 = cb * p.CBW;
  const int t_begin = tg * p.TPG, t_end = min(p.taps, t_begin + p.TPG);
  const int tile_begin = (int)(((long long)p.n_tiles * ps) / p.n_ps), tile_end = (int)(((long long)p.n_tiles * (ps + 1)) / p.n_ps);

  // stale shared memory must not hold NaNs where zeros are expected (rows in front of the first map; planes beyond the tensor are
  // only ever multiplied into accumulator rows / columns nobody reads, but 0 * NaN inside a read row would poison it)
  for (uint32_t i = tid; i < (ST * (p.a_bytes + p.b_bytes)) / 16; i += WG_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < ST; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
    mbar_init(smem_u32(done), 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 5) {
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      const int a_load = min(p.a_planes, n_planes - plane0);
      const int b_load = min(p.b_planes, p.C / 4 - c0 / 4);
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const long long q0 = (long long)tile * WG_TM;
        mbar_wait(smem_u32(&empty[st]), ph ^ 1);
        const long long g0 = q0 - p.d_before;
        const long long lo = g0 < 0 ? 0 : g0;
        const uint32_t b_rows = (uint32_t)(g0 + p.RA_p - lo);
        const uint32_t b_off = (uint32_t)(lo - g0) * 16;
        const uint32_t bar = smem_u32(&full[st]);
        mbar_arrive_expect_tx(bar, (uint32_t)a_load * WG_TM * 16 + (uint32_t)(b_load * p.n_strips) * b_rows * 16);
        const uint32_t a_slot = smem_u32(a_ring + (size_t)st * p.a_bytes), b_slot = smem_u32(b_ring + (size_t)st * p.b_bytes);
        for (int j = 0; j < a_load; ++j)
          bulk_load_g2s(a_slot + (uint32_t)j * WG_TM * 16, A + ((size_t)(plane0 + j) * p.g_plane + q0) * 4, WG_TM * 16, bar);
        for (int s2 = 0; s2 < p.n_strips; ++s2)
          for (int j = 0; j < b_load; ++j)
            bulk_load_g2s(b_slot + (uint32_t)((s2 * p.b_planes + j) * p.RA_p) * 16 + b_off,
                          Bx + ((size_t)(c0 / 4 + j) * p.x_plane + (size_t)s2 * p.strip_rows + lo) * 4, b_rows * 16, bar);
        if (++st == ST) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 4) {
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0;
      // MN-major, no swizzle: 16-byte units are 4 channels of one pixel; 8 consecutive pixels = one 128-byte core matrix along K;
      // the next 4-channel group along M / N is one chunk plane further (SBO); LBO (next 8 pixels) = 128 bytes
      const uint64_t adesc_hi = (p.variant & 1) ? make_smem_desc(0, WG_TM * 16, 128) : make_smem_desc(0, 128, WG_TM * 16);
      const uint64_t bdesc_hi = (p.variant & 2) ? make_smem_desc(0, (uint32_t)p.RA_p * 16, 128) : make_smem_desc(0, 128, (uint32_t)p.RA_p * 16);
      bool first = true;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        mbar_wait(smem_u32(&full[st]), ph);
        tc_fence_after();
        const uint32_t a16 = smem_u32(a_ring + (size_t)st * p.a_bytes) >> 4, b16 = smem_u32(b_ring + (size_t)st * p.b_bytes) >> 4;
#pragma unroll 1
        for (int t = t_begin; t < t_end; ++t) {
          const uint32_t tcol = tmem_base + (uint32_t)((t - t_begin) * p.n_pad);
          const uint32_t bt = b16 + (uint32_t)p.tap_off[t];
#pragma unroll 4
          for (int kb = 0; kb < WG_TM / 8; ++kb) {
            const uint64_t ad = adesc_hi | (uint64_t)((a16 + kb * 8) & 0x3FFF), bd = bdesc_hi | (uint64_t)((bt + kb * 8) & 0x3FFF);
            if (first && kb == 0) umma_mma_c<MODE_EVAL, false>(tcol, ad, bd, p.idesc);
            else umma_mma_c<MODE_EVAL, true>(tcol, ad, bd, p.idesc);
          }
        }
        first = false;
        umma_commit(smem_u32(&empty[st]));
        if (++st == ST) { st = 0; ph ^= 1; }
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
      const int n = plane0 * 4 + tid;                            // tid in [0, 128)
      const bool mine = n >= n_first && n < p.N_out;
      const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16);
      const int K = p.taps * p.C_real;
      for (int t = t_begin; t < t_end; ++t) {
        for (int cg = 0; cg < p.n_pad; cg += 16) {
          uint32_t v[16];
          tmem_ld16(tlane + (uint32_t)((t - t_begin) * p.n_pad + cg), v);
          tmem_ld_wait();
          if (mine) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int c = c0 + cg + j;
              if (cg + j < p.CBW && c < p.C_real) atomicAdd(out + (size_t)n * K + (size_t)t * p.C_real + c, __uint_as_float(v[j]));
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

}  // namespace

// dmu_p / dsig2_p: [N][R*S][C_real] fp32 (the packed OHWI order of qbn_weight_prep / qbn_weight_grad_post), OVERWRITTEN.
// C: channels of the planar input (zero planes beyond C_real, e.g. the 3 -> 8 padded first layer).
// g, dv: planar maps of the layer's OUTPUT geometry [N/4][g_plane_rows][4]; x, x_sq: the layer's planar input (phase-split for a
// stride-2 layer, as the forward reads it).  Every plane must be readable (zeros) for 128 + (Wp + 1) rows past the last map.
extern "C" int qbn_lrt_wgrad_p4(int B, int Hp, int Wp, int C, int C_real, int N, int R, int S, int stride, const float* g, const float* dv,
                                long long g_plane_rows, const float* x, const float* x_sq, long long x_plane_rows, float* dmu_p,
                                float* dsig2_p, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  QBN_CHECK_ARG(g && dv && x && x_sq && dmu_p && dsig2_p, "null pointer");
  QBN_CHECK_ARG(B > 0 && Hp > 2 && Wp > 2 && C > 0 && C_real > 0 && C_real <= C && N > 0 && R > 0 && S > 0, "sizes");
  const bool s1 = stride == 1 && (R & 1) && (S & 1);
  const bool s2 = stride == 2 && ((R == 3 && S == 3) || (R == 1 && S == 1));
  if (C % 4 != 0 || N % 4 != 0 || !(s1 || s2) || R * S > WG_MAX_TAPS) {
    qbn_set_error("qbn_lrt_wgrad_p4: needs C %% 4 == 0, N %% 4 == 0 and stride 1 (odd kernel) or stride 2 (3x3 / 1x1) (C=%d N=%d R=%d S=%d stride=%d)",
                  C, N, R, S, stride);
    return QBN_ERR_UNSUPPORTED;
  }
  WGParams p;
  memset(&p, 0, sizeof(p));
  p.Qs = B * Hp * Wp;
  p.n_tiles = (p.Qs + WG_TM - 1) / WG_TM;
  p.N_out = N; p.C = C; p.C_real = C_real; p.taps = R * S;
  const int bh = s1 ? (R - 1) / 2 : 1, bw = s1 ? (S - 1) / 2 : 1;
  int d_after;
  if (s1) { p.n_strips = 1; p.d_before = bh * Wp + bw; d_after = p.d_before; }
  else { p.n_strips = (R == 3) ? 4 : 1; p.d_before = (R == 3) ? Wp + 1 : 0; d_after = 0; }
  p.strip_rows = p.Qs;
  p.RA_p = (WG_TM + p.d_before + d_after + 7) / 8 * 8;
  // input-channel blocks: MMA N = channels of the block rounded up to 16; the largest block whose two stages fit next to the g tiles
  p.a_planes = 32;
  p.a_bytes = (uint32_t)p.a_planes * WG_TM * 16;
  {
    const int cand[7] = {C <= 96 ? C : 0, 96, 64, 48, 32, 16, 8};
    p.CBW = 0;
    for (int i = 0; i < 7 && !p.CBW; ++i) {
      const int cb = cand[i];
      if (cb <= 0 || cb > C || C % cb != 0) continue;
      const size_t bb = (size_t)p.n_strips * ((cb + 15) / 16 * 4) * p.RA_p * 16;
      if (2 * ((size_t)p.a_bytes + bb) + 128 <= 225 * 1024) p.CBW = cb;
    }
    if (!p.CBW) {
      qbn_set_error("qbn_lrt_wgrad_p4: no channel blocking fits shared memory (C=%d Wp=%d)", C, Wp);
      return QBN_ERR_UNSUPPORTED;
    }
  }
  p.n_cb = (C + p.CBW - 1) / p.CBW;
  p.n_pad = (p.CBW + 15) / 16 * 16;
  p.b_planes = p.n_pad / 4;
  p.n_mb = (N + 127) / 128;
  p.TPG = 512 / p.n_pad;
  if (p.TPG > p.taps) p.TPG = p.taps;
  p.n_tg = (p.taps + p.TPG - 1) / p.TPG;
  p.TPG = (p.taps + p.n_tg - 1) / p.n_tg;                      // balanced groups
  p.tmem_cols = 32;
  while (p.tmem_cols < p.TPG * p.n_pad) p.tmem_cols <<= 1;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) {
      int strip = 0, sh;
      if (s1) sh = (r - bh) * Wp + (s - bw);
      else if (R == 3) { const int dr = r - 1, ds = s - 1; strip = (dr & 1) * 2 + (ds & 1); sh = (dr < 0 ? -1 : 0) * Wp + (ds < 0 ? -1 : 0); }
      else sh = 0;
      p.tap_off[r * S + s] = strip * p.b_planes * p.RA_p + p.d_before + sh;
    }
  p.b_bytes = (uint32_t)p.n_strips * p.b_planes * p.RA_p * 16;
  const size_t smem = 2 * ((size_t)p.a_bytes + p.b_bytes) + 128;
  if (smem > 225 * 1024) {
    qbn_set_error("qbn_lrt_wgrad_p4: tile does not fit shared memory (%zu bytes)", smem);
    return QBN_ERR_UNSUPPORTED;
  }
  const long long need_g = (long long)p.n_tiles * WG_TM, need_x = (long long)p.n_tiles * WG_TM + d_after + 8 + (long long)(p.n_strips - 1) * p.strip_rows;
  if (g_plane_rows < need_g || x_plane_rows < need_x) {
    qbn_set_error("qbn_lrt_wgrad_p4: planes too short (g %lld < %lld or x %lld < %lld rows): allocate a zero tail of 128 + Wp + 1 rows", g_plane_rows,
                  need_g, x_plane_rows, need_x);
    return QBN_ERR_INVALID_ARG;
  }
  p.g = g; p.dv = dv; p.g_plane = g_plane_rows; p.x = x; p.xsq = x_sq; p.x_plane = x_plane_rows; p.dmu = dmu_p; p.dsig2 = dsig2_p;
  // F32 += TF32 x TF32, A and B MN-major (bits 15, 16), N at bit 17, M = 128 at bit 24
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.n_pad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (tune_env("QBN_WG_V")) {
    p.variant = atoi(tune_env("QBN_WG_V"));
    if (p.variant & 4) p.idesc &= ~((1u << 15) | (1u << 16));
  }
  const int items = 2 * p.n_mb * p.n_cb * p.n_tg;
  p.n_ps = (2 * qbn_sm_count() + items - 1) / items;           // ~2 waves of single-CTA SMs: pixel ranges long enough to amortise the epilogue
  if (p.n_ps > p.n_tiles) p.n_ps = p.n_tiles;
  if (p.n_ps < 1) p.n_ps = 1;
  QBN_CUDA(cudaMemsetAsync(dmu_p, 0, sizeof(float) * (size_t)N * p.taps * C_real, st));
  QBN_CUDA(cudaMemsetAsync(dsig2_p, 0, sizeof(float) * (size_t)N * p.taps * C_real, st));
  static bool attr_set = false;
  if (!attr_set) {
    QBN_CUDA(cudaFuncSetAttribute(umma_wgrad_p4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    attr_set = true;
  }
  umma_wgrad_p4_kernel<<<items * p.n_ps, WG_THREADS, smem, st>>>(p);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}
