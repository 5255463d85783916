// tcgen05 "zero-copy im2col" kernel for the stride-1 same-padding convolutions (16 of the 21
// stochastic layers of the ResNet, ~85 % of its FLOPs), persistent and warp-specialised.
//
// Activations live in HBM in a zero-bordered ("padded") NHWC layout [B][H+2ph][W+2pw][C]; the output
// uses the same layout with N channels.  With q = the flat pixel index of that padded grid, the
// input row needed by output row q for filter tap (r,s) is simply q + (r-ph)*Wp + (s-pw): a constant
// shift.  So ONE shared-memory tile holding rows [q0-D, q0+128+D) of a channel block serves all R*S
// taps — each tap is the same tile addressed through a UMMA descriptor whose start address is moved
// by the shift (the canonical K-major no-swizzle layout keeps 8 consecutive rows in one 128-byte
// core matrix, and consecutive rows are consecutive q).  Every input element is fetched from
// L2/HBM once per tile instead of R*S times, with no bounds tests (the zero border IS the padding)
// and no per-element address arithmetic.  Border rows are computed but stored as zeros, which keeps
// the border of the next layer's input intact.
//
//   warps 0-3  epilogue   : TMEM -> registers -> affine(BN)/residual/ReLU/TF32-round -> global
//   warp  4    MMA issuer : tcgen05.mma kind::tf32, double-buffered TMEM accumulators
//   warps 5-8  producers  : cp.async (LDGSTS) into two rings — A (activation tile, one slot per
//                           channel block) and B (sampled weights, one slot per (channel block, tap))
// CTAs are persistent (grid = SMs x occupancy) and walk the tile list round-robin, so TMEM
// allocation, barrier setup and the epilogue of tile i overlap the loads/MMAs of tile i+1.
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

QBN_DEVINL void epi_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the four epilogue warps only
QBN_DEVINL float4 ld_nc_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// tuning/diagnostic switches are compiled out of the product build (-DQBN_TUNING enables them)
#ifdef QBN_TUNING
#define S1DBG (p.dbg)
#else
#define S1DBG 0
#endif
// cycle accounting of CTA 0 (QBN_S1_DBG bit 8192): [role*8 + category] summed over its tiles
__device__ unsigned long long g_s1_prof[32];
#define PROF_BEGIN() long long _t0 = (S1DBG & 8192) ? clock64() : 0
#define PROF_ADD(slot)                                                            \
  do {                                                                            \
    if ((S1DBG & 8192) && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 4 || warp == 4 + NW_MMA)) { \
      long long _t1 = clock64();                                                  \
      g_s1_prof[slot] += (unsigned long long)(_t1 - _t0);                         \
      _t0 = _t1;                                                                  \
    }                                                                             \
  } while (0)

constexpr int TM = 128;
constexpr int N_EPI = 128, N_PROD = 128;
constexpr int NW_MMA = 1;                          // MMA-issuing warps (one tcgen05.mma costs ~83 cycles to ISSUE per thread on B200; two issuers x two CTAs/SM approach the SM-wide dispatch rate, scripts/ubench.py)
constexpr int NTHREADS_S1 = N_EPI + 32 * NW_MMA + N_PROD;   // 320

struct S1Params {
  int Hp, Wp, C, N, R, S, ph, pw;
  int Qs;                         // rows (padded pixels) per Monte-Carlo sample
  int tiles_per_sample, total_tiles;
  int n_pad, CB, cbc, n_cb, K;    // MMA N, channels per block, 16-byte chunks per block (even), #blocks, R*S*C
  int D, RA, RA_p, b_pitch;       // halo rows, A rows per slot, pitches (in 16-byte chunks)
  int SA, SB;                     // ring depths
  int b_res;                      // 1: the whole per-sample weight tensor stays resident in smem (reloaded on sample change)
  int TG;                         // streaming mode: filter taps per B slot
  int bulk_out;                   // narrow layers: stage the output tile in smem and write it with one bulk copy
  int ACC;                        // TMEM accumulator stages (MMA may run ACC tiles ahead of the epilogue)
  int dbg;                        // tuning knobs (QBN_S1_DBG): 1 round-robin tiles, 2 all-lane polling
  int tmem_cols, flags, w_shared;
  uint32_t idesc;
  const float* x; const float* w; const float* scale; const float* shift; const float* residual; float* out;
};

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(NTHREADS_S1, MIN_BLOCKS) umma_conv_s1_kernel(const S1Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_bytes = (uint32_t)p.cbc * p.RA_p * 16;
  const uint32_t bt_bytes = (uint32_t)p.cbc * p.b_pitch * 16;                 // one (channel block, tap) weight block
  const uint32_t b_bytes = p.b_res ? bt_bytes * p.n_cb * p.R * p.S : bt_bytes * p.TG;   // one B slot
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + (size_t)p.SA * a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + (size_t)p.SB * b_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + p.SA;
  uint64_t* b_full = a_empty + p.SA;
  uint64_t* b_empty = b_full + p.SB;
  uint64_t* acc_full = b_empty + p.SB;         // [ACC]
  uint64_t* acc_empty = acc_full + p.ACC;      // [ACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + p.ACC);
  float* s_scale = reinterpret_cast<float*>(tmem_slot + 4);        // [256+4] per-channel affine (eval BatchNorm / bias)
  float* s_shift = s_scale + 260;                                  // [256+4]
  float* out_stage = s_shift + 260;                                // [2][128][N] when bulk_out (16-byte aligned)
  for (int i = tid; i < 260; i += NTHREADS_S1) {
    s_scale[i] = (p.scale && i < p.N) ? p.scale[i] : 1.f;
    s_shift[i] = (p.shift && i < p.N) ? p.shift[i] : 0.f;
  }

  if (tid == 0) {
    for (int i = 0; i < p.SA; ++i) { mbar_init(smem_u32(&a_full[i]), N_PROD); mbar_init(smem_u32(&a_empty[i]), 1); }
    for (int i = 0; i < p.SB; ++i) { mbar_init(smem_u32(&b_full[i]), N_PROD); mbar_init(smem_u32(&b_empty[i]), p.b_res ? NW_MMA : 1); }
    for (int i = 0; i < p.ACC; ++i) { mbar_init(smem_u32(&acc_full[i]), 1); mbar_init(smem_u32(&acc_empty[i]), N_EPI); }
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int taps = p.R * p.S;
  // contiguous tile range per CTA: consecutive tiles share the sample (=> the weights) and their halos hit L2
  const bool rr = S1DBG & 1;
  const int tile_begin = rr ? (int)blockIdx.x : (int)(((long long)p.total_tiles * blockIdx.x) / gridDim.x);
  const int tile_end = rr ? p.total_tiles : (int)(((long long)p.total_tiles * (blockIdx.x + 1)) / gridDim.x);
  const int tile_step = rr ? (int)gridDim.x : 1;
  // one lane polls an mbarrier on behalf of its warp (32x fewer smem polls / issue slots)
  auto warp_wait = [&](uint64_t* bar, uint32_t parity) {
    if (lane == 0 || (S1DBG & 2)) {
      if (S1DBG & 256) mbar_spin(smem_u32(bar), parity); else mbar_wait(smem_u32(bar), parity);
    }
    __syncwarp();
  };

  if (warp >= 4 + NW_MMA) {
    // ======================================= PRODUCERS ==========================================
    const int pt = tid - (N_EPI + 32 * NW_MMA);   // 0..127
    const int CH = p.cbc;
    const int j = pt % CH;                        // this thread's 16-byte chunk inside every slot
    const int lane_row = pt / CH;
    const int RS = N_PROD / CH;                   // rows advanced per pass
    const bool active = lane_row < RS;            // (128 % CH) threads idle but still arrive
    const int ra_iters = ((p.RA + RS - 1) / RS) * RS;
    int sa = 0, sb = 0, cur_z = -1;
    uint32_t pa = 0, pb = 0;
    // Completion signalling: cp.async.mbarrier.arrive.noinc — each producer thread's arrival on the slot's
    // full barrier fires when ITS copies have landed, so the threads never wait for data and every ring
    // slot can be in flight at once.  (A wait_group-based one-arrival-per-warp scheme was measured
    // slower: it serialises the weight stream behind the L2 latency.)
    auto publish = [&](uint32_t bar) { cp_async_arrive_noinc(bar); };
    auto signal_older = [&](int) {};
    auto load_b_block = [&](uint32_t dst_block, const float* ws, int cb, int t) {      // one (cb, tap) weight block
      const int c = cb * p.CB + 4 * j;
      const bool cv = active && c < p.C && 4 * j < p.CB;
      const uint32_t dst0 = dst_block + (uint32_t)(j * p.b_pitch) * 16;
      const float* wt = ws + (size_t)t * p.C + c;
      for (int n = lane_row; active && n < p.n_pad; n += RS) {
        const bool ok = cv && n < p.N;
        cp_async16(dst0 + (uint32_t)n * 16, ok ? wt + (size_t)n * p.K : p.w, ok ? 16u : 0u);
      }
    };
    for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
      const int z = tile / p.tiles_per_sample;
      const int q0 = (tile - z * p.tiles_per_sample) * TM;
      const float* xs = p.x + (size_t)z * p.Qs * p.C;
      const float* ws = p.w + (p.w_shared ? 0 : (size_t)z * p.N * p.K);
      PROF_BEGIN();
      if (p.b_res && z != cur_z) {
        // resident weights: (re)load the whole sampled tensor of sample z once.  The MMA warp frees the
        // weight buffer only after the previous sample's last tiles, so their activation slots must be
        // signalled before blocking here.
        signal_older(0);
        warp_wait(&b_empty[0], pb ^ 1);
        const uint32_t base = smem_u32(b_ring);
        for (int cb = 0; cb < p.n_cb; ++cb)
          for (int t = 0; t < taps; ++t) load_b_block(base + (uint32_t)(cb * taps + t) * bt_bytes, ws, cb, t);
        publish(smem_u32(&b_full[0]));
        signal_older(0);                                  // one-slot 'ring': signal at once (rare: once per sample)
        pb ^= 1;
        cur_z = z;
      }
      PROF_ADD(16);
      for (int cb = 0; cb < p.n_cb; ++cb) {
        const int c = cb * p.CB + 4 * j;
        const bool cv = active && c < p.C && 4 * j < p.CB;
        // ---- A slot: rows [q0-D, q0+TM+D) of channel block cb, one contiguous strip of the padded map
        warp_wait(&a_empty[sa], pa ^ 1);
        PROF_ADD(17);
        {
          // same trip count for every lane (no divergence); address advanced incrementally
          uint32_t dst = smem_u32(a_ring + (size_t)sa * a_bytes) + (uint32_t)(j * p.RA_p + lane_row) * 16;
          int qq = q0 - p.D + lane_row;
          const float* src = xs + (ptrdiff_t)qq * p.C + c;
          const ptrdiff_t src_step = (ptrdiff_t)RS * p.C;
          for (int rho = lane_row; rho < ra_iters && !(S1DBG & 4096); rho += RS) {
            const bool ok = cv && rho < p.RA && qq >= 0 && qq < p.Qs;
            if (active && rho < p.RA && !(S1DBG & 64)) cp_async16(dst, ok ? src : p.x, ok ? 16u : 0u);
            dst += (uint32_t)RS * 16; qq += RS; src += src_step;
          }
        }
        PROF_ADD(18);
        publish(smem_u32(&a_full[sa]));
        PROF_ADD(19);
        if (++sa == p.SA) { sa = 0; pa ^= 1; }
        // ---- streamed weights: one slot per group of TG taps of this channel block
        if (!p.b_res) {
          for (int t0 = 0; t0 < taps; t0 += p.TG) {
            warp_wait(&b_empty[sb], pb ^ 1);
            PROF_ADD(20);
            const uint32_t base = smem_u32(b_ring + (size_t)sb * b_bytes);
            for (int t = t0; t < t0 + p.TG && t < taps; ++t) load_b_block(base + (uint32_t)(t - t0) * bt_bytes, ws, cb, t);
            PROF_ADD(21);
            publish(smem_u32(&b_full[sb]));
            PROF_ADD(22);
            if (++sb == p.SB) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
    signal_older(0);                                      // drain
  } else if (warp >= 4) {
    // ======================================= MMA ISSUERS ========================================
    // Tiles are dealt round-robin to the NW_MMA issuer warps; every warp walks the whole tile list to
    // keep its ring/phase counters in step, but only issues (and commits) for the tiles it owns.
    const int mw = warp - 4;
    int sa = 0, sb = 0, as = 0, cur_z = -1;
    uint32_t pa = 0, pb = 0, pacc = 0;
    const uint32_t lbo_a = (uint32_t)p.RA_p * 16, lbo_b = (uint32_t)p.b_pitch * 16;
    const int nk = p.cbc / 2;
    // descriptor = constant high part | (smem address >> 4): only the 14-bit start-address field moves
    const uint64_t adesc_hi = make_smem_desc(0, lbo_a, 128), bdesc_hi = make_smem_desc(0, lbo_b, 128);
    const uint32_t a_k = (2 * lbo_a) >> 4, b_k = (2 * lbo_b) >> 4;      // K-step (two 16-byte chunks) in 16-byte units
    const uint32_t leader = lane == 0 ? 1u : 0u;     // all lanes run the uniform issue code; only this one's tcgen05 ops fire
    auto issue_tap = [&](uint32_t tacc, uint32_t abase, uint32_t bblock, int t, uint32_t& first) {
      const int r = t / p.S, s = t - r * p.S;
      const int shift = p.D + (r - p.ph) * p.Wp + (s - p.pw);       // slot row that output row 0 reads for this tap
      uint32_t a16 = (abase >> 4) + (uint32_t)shift, b16 = bblock >> 4;
      for (int jj = 0; jj < ((S1DBG & 32) ? (t == 0 ? 1 : 0) : nk); ++jj) {
        umma_mma_tf32_pred(tacc, adesc_hi | (uint64_t)(a16 & 0x3FFF), bdesc_hi | (uint64_t)(b16 & 0x3FFF), p.idesc, first ? 0u : 1u, leader);
        first = 0;
        a16 += a_k; b16 += b_k;
      }
    };
    int local = 0;
    for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++local) {
      const int z = tile / p.tiles_per_sample;
      const bool last_of_z = (tile + tile_step >= tile_end) || ((tile + tile_step) / p.tiles_per_sample != z);
      const bool mine = (local % NW_MMA) == mw;
      PROF_BEGIN();
      if (p.b_res && z != cur_z) {
        warp_wait(&b_full[0], pb);
        pb ^= 1;
        cur_z = z;
      }
      if (mine) warp_wait(&acc_empty[as], pacc ^ 1);      // epilogue has drained this accumulator
      PROF_ADD(8);
      const uint32_t tacc = tmem_base + (uint32_t)(as * p.n_pad);
      uint32_t first = 1;
      for (int cb = 0; cb < p.n_cb; ++cb) {
        if (mine) warp_wait(&a_full[sa], pa);
        PROF_ADD(9);
        const uint32_t abase = smem_u32(a_ring + (size_t)sa * a_bytes);
        if (p.b_res) {
          if (mine) {
            if (!(S1DBG & 512)) fence_proxy_async();      // cp.async data (generic proxy) -> ordered before the async-proxy reads
            tc_fence_after();
            for (int t = 0; t < taps; ++t) issue_tap(tacc, abase, smem_u32(b_ring) + (uint32_t)(cb * taps + t) * bt_bytes, t, first);
            umma_commit_pred(smem_u32(&a_empty[sa]), leader);
            if (cb == p.n_cb - 1) umma_commit_pred(smem_u32(&acc_full[as]), leader);
          }
          // every issuer warp releases the resident weights of sample z (its commit covers its own MMAs)
          if (cb == p.n_cb - 1 && last_of_z) umma_commit_pred(smem_u32(&b_empty[0]), leader);
          PROF_ADD(10);
        } else {
          for (int t0 = 0; t0 < taps; t0 += p.TG) {
            if (mine) {
              warp_wait(&b_full[sb], pb);
              PROF_ADD(11);
              fence_proxy_async();
              tc_fence_after();
              const uint32_t bbase = smem_u32(b_ring + (size_t)sb * b_bytes);
              for (int t = t0; t < t0 + p.TG && t < taps; ++t) issue_tap(tacc, abase, bbase + (uint32_t)(t - t0) * bt_bytes, t, first);
              umma_commit_pred(smem_u32(&b_empty[sb]), leader);
              if (t0 + p.TG >= taps) {
                umma_commit_pred(smem_u32(&a_empty[sa]), leader);
                if (cb == p.n_cb - 1) umma_commit_pred(smem_u32(&acc_full[as]), leader);
              }
              PROF_ADD(10);
            }
            if (++sb == p.SB) { sb = 0; pb ^= 1; }
          }
        }
        if (++sa == p.SA) { sa = 0; pa ^= 1; }
      }
      if (++as == p.ACC) { as = 0; pacc ^= 1; }
    }
  } else {
    // ======================================= EPILOGUE ===========================================
    // Per 32-column chunk: TMEM -> registers (one lane = one output row) -> smem transpose -> a
    // coalesced pass in which 8 consecutive lanes own one 128-byte row segment: residual read,
    // affine/ReLU/round, store.  The residual loads of a chunk are issued BEFORE the accumulator is
    // waited for, so their DRAM latency hides behind the MMAs.
    int as = 0;
    uint32_t pacc = 0;
    // One TMEM lane = one output row per thread, processed in 16-column chunks with the next chunk's
    // TMEM load and residual fetch already in flight.  Narrow layers (128 x N x 4 B <= 24 KB: the
    // HBM-bound ones) stage the finished tile in smem and write it with ONE bulk (TMA-engine) copy —
    // the tile's rows are contiguous in the zero-bordered layout — instead of 32-way scattered STG.
    const int plane = p.Hp * p.Wp;
    const int n_chunks = p.n_pad / 16;
    const bool vec_ok = (p.N & 3) == 0;
    const bool relu = p.flags & QBN_FLAG_RELU, rnd = p.flags & QBN_FLAG_OUT_ROUND_TF32;
    const bool bulk_out = p.bulk_out;
    int buf = 0;
    for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
      PROF_BEGIN();
      const int z = tile / p.tiles_per_sample;
      const int q0 = (tile - z * p.tiles_per_sample) * TM;
      const int q = q0 + tid;
      const bool qv = q < p.Qs;
      const int rem = qv ? q % plane : 0;
      const int hh = rem / p.Wp, ww = rem - hh * p.Wp;
      const bool interior = qv && hh >= p.ph && hh < p.Hp - p.ph && ww >= p.pw && ww < p.Wp - p.pw;
      const size_t orow = ((size_t)z * p.Qs + (qv ? q : 0)) * p.N;
      float* optr = p.out + orow;
      float* sptr = out_stage + (size_t)buf * TM * p.N + (size_t)tid * p.N;      // this row inside the staging tile
      const float* rptr = (p.residual && interior && !(S1DBG & 128)) ? p.residual + orow : nullptr;
      float4 rres[4], rnext[4];
      uint32_t v[16], vn[16];
      auto prefetch = [&](float4* dst, int cc) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int col = cc * 16 + 4 * i;
          dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rptr && col < p.N) {
            if (vec_ok) {
              dst[i] = ld_nc_f4(rptr + col);
            } else {                                      // N % 4 != 0: rows are not 16-byte aligned
              dst[i].x = rptr[col];
              if (col + 1 < p.N) dst[i].y = rptr[col + 1];
              if (col + 2 < p.N) dst[i].z = rptr[col + 2];
              if (col + 3 < p.N) dst[i].w = rptr[col + 3];
            }
          }
        }
      };
      prefetch(rres, 0);
      PROF_ADD(0);
      warp_wait(&acc_full[as], pacc);
      tc_fence_after();
      PROF_ADD(1);
      const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(as * p.n_pad);
      if (!(S1DBG & 2048)) tmem_ld16(tlane, v);
      if (bulk_out) {                                     // staging buffer `buf` must have been drained by its bulk store
        if (tid == 0) bulk_wait_read<1>();
        epi_sync();
      }
      for (int cc = 0; cc < n_chunks; ++cc) {
        const int c0 = cc * 16;
        tmem_ld_wait();                                   // chunk cc has landed in v
        if (cc + 1 < n_chunks) {
          if (!(S1DBG & 2048)) tmem_ld16(tlane + (uint32_t)(c0 + 16), vn);
          prefetch(rnext, cc + 1);
        } else {                                          // accumulator fully read: release it to the MMA warp
          tc_fence_before();
          mbar_arrive(smem_u32(&acc_empty[as]));
        }
        PROF_ADD(2);
        if (qv && !(S1DBG & 16)) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int col = c0 + 4 * i;
            if (col < p.N) {
              float a[4] = {0.f, 0.f, 0.f, 0.f};
              if (interior) {
                const float4 sc = *reinterpret_cast<const float4*>(&s_scale[col]);
                const float4 sh = *reinterpret_cast<const float4*>(&s_shift[col]);
                a[0] = fmaf(__uint_as_float(v[4 * i + 0]), sc.x, sh.x) + rres[i].x;
                a[1] = fmaf(__uint_as_float(v[4 * i + 1]), sc.y, sh.y) + rres[i].y;
                a[2] = fmaf(__uint_as_float(v[4 * i + 2]), sc.z, sh.z) + rres[i].z;
                a[3] = fmaf(__uint_as_float(v[4 * i + 3]), sc.w, sh.w) + rres[i].w;
                if (relu) { a[0] = fmaxf(a[0], 0.f); a[1] = fmaxf(a[1], 0.f); a[2] = fmaxf(a[2], 0.f); a[3] = fmaxf(a[3], 0.f); }
                if (rnd) { a[0] = tf32_round(a[0]); a[1] = tf32_round(a[1]); a[2] = tf32_round(a[2]); a[3] = tf32_round(a[3]); }
              }
              if (bulk_out) {
                *reinterpret_cast<float4*>(sptr + col) = make_float4(a[0], a[1], a[2], a[3]);
              } else if (vec_ok) {
                *reinterpret_cast<float4*>(optr + col) = make_float4(a[0], a[1], a[2], a[3]);
              } else {
                for (int e = 0; e < 4 && col + e < p.N; ++e) optr[col + e] = a[e];
              }
            }
          }
        }
        if (cc + 1 < n_chunks) {
#pragma unroll
          for (int i = 0; i < 4; ++i) rres[i] = rnext[i];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = vn[i];
        }
        PROF_ADD(3);
      }
      if (bulk_out) {
        fence_proxy_async();                              // staged rows (generic proxy) -> visible to the bulk-copy engine
        epi_sync();
        if (tid == 0 && !(S1DBG & 16)) {
          const int rows = min(TM, p.Qs - q0);
          bulk_store_s2g(p.out + ((size_t)z * p.Qs + q0) * p.N, smem_u32(out_stage + (size_t)buf * TM * p.N), (uint32_t)(rows * p.N * 4));
          bulk_commit();
        }
        buf ^= 1;
      }
      if (++as == p.ACC) { as = 0; pacc ^= 1; }
    }
    if (bulk_out && tid == 0) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

}  // namespace

extern "C" int qbn_conv_s1_fwd(int n_samples, int B, int Hp, int Wp, int C, int N, int R, int S, const float* x, const float* w,
                               int w_shared, const float* scale, const float* shift, const float* residual, int flags, float* out,
                               void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  QBN_CHECK_ARG(x && w && out, "null pointer");
  QBN_CHECK_ARG(n_samples > 0 && B > 0 && Hp > 0 && Wp > 0 && C > 0 && N > 0 && R > 0 && S > 0, "sizes");
  QBN_CHECK_ARG(Hp > R - 1 && Wp > S - 1, "padded extent must exceed the halo");
  if (C % 4 != 0 || (R & 1) == 0 || (S & 1) == 0 || N > 256) {
    qbn_set_error("qbn_conv_s1_fwd: needs C %% 4 == 0, odd kernel, N <= 256 (C=%d R=%d S=%d N=%d)", C, R, S, N);
    return QBN_ERR_UNSUPPORTED;
  }
  S1Params p;
  memset(&p, 0, sizeof(p));
  p.Hp = Hp; p.Wp = Wp; p.C = C; p.N = N; p.R = R; p.S = S; p.ph = (R - 1) / 2; p.pw = (S - 1) / 2;
  p.Qs = B * Hp * Wp;
  p.tiles_per_sample = (p.Qs + TM - 1) / TM;
  p.total_tiles = p.tiles_per_sample * n_samples;
  p.n_pad = (N + 15) / 16 * 16;
  p.K = R * S * C;
  // channel blocking: whole C when small, else 32-channel blocks; chunks per block rounded to even
  p.CB = C <= 48 ? C : 32;
  p.n_cb = (C + p.CB - 1) / p.CB;
  p.cbc = ((p.CB + 7) / 8) * 2;
  p.D = p.ph * Wp + p.pw;
  p.RA = TM + 2 * p.D;
  p.RA_p = p.RA | 1;                // odd pitch: conflict-free LDGSTS for the (row, chunk) thread map
  p.b_pitch = p.n_pad + 1;
  p.flags = flags; p.w_shared = w_shared;
  p.x = x; p.w = w; p.scale = scale; p.shift = shift; p.residual = residual; p.out = out;
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.n_pad >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
  const size_t a_bytes = (size_t)p.cbc * p.RA_p * 16, bt_bytes = (size_t)p.cbc * p.b_pitch * 16;
  p.bulk_out = ((size_t)TM * N * 4 <= 24 * 1024 && N % 4 == 0 && tune_env("QBN_S1_BULK")) ? 1 : 0;   // measured slower than direct row stores (extra barriers); kept as a tuning knob
  const size_t epi_bytes = 16 + 2 * 260 * 4 + 16 + (p.bulk_out ? 2 * (size_t)TM * N * 4 : 0);
  const size_t cap = 220 * 1024;
  const int taps = R * S;
  const size_t b_all = bt_bytes * p.n_cb * taps;
  size_t b_bytes, smem;
  const char* dbg_env = tune_env("QBN_S1_DBG");
  p.dbg = dbg_env ? atoi(dbg_env) : 0;
  // Policy: several small CTAs per SM rather than one deep pipeline — one tcgen05.mma costs ~83 cycles to
  // issue from a thread (scripts/ubench.py), so issuers in different CTAs are what fills the tensor pipe,
  // and co-resident CTAs hide each other's barrier round trips.
  const size_t occ_budget[4] = {0, cap, 110 * 1024, 72 * 1024};
  int want_occ = 1;
  if (b_all <= 96 * 1024 && !(p.dbg & 4)) {
    // resident weights: one "slot" holding the whole sampled tensor of the current sample
    p.b_res = 1; p.SB = 1; p.TG = taps;
    b_bytes = b_all;
    auto total = [&](int sa) { return sa * a_bytes + b_bytes + (2 * sa + 2 + 2 * 8) * 8 + epi_bytes; };
    want_occ = total(2) <= occ_budget[2] ? 2 : 1;      // (3 CTAs/SM measured slower: the SM-wide issue rate, not latency, binds)
    p.SA = 2;
    while (p.SA < 4 && total(p.SA + 1) <= occ_budget[want_occ]) p.SA++;
    smem = total(p.SA);
  } else {
    p.b_res = 0;
    p.TG = 1;
    b_bytes = bt_bytes;
    auto total = [&](int sa, int sb) { return sa * a_bytes + sb * b_bytes + (2 * sa + 2 * sb + 2 * 8) * 8 + epi_bytes; };
    want_occ = (total(2, 3) <= occ_budget[2] && p.n_pad <= 128) ? 2 : 1;   // wide layers: one CTA with deep rings + 2 accumulators
    p.SA = 2; p.SB = 2;
    while (p.SB < 6 && total(p.SA, p.SB + 1) <= occ_budget[want_occ]) p.SB++;
    smem = total(p.SA, p.SB);
  }
  // TMEM: want_occ CTAs share 512 columns
  {
    int cols_per_cta = 512 / want_occ;
    int c2 = 32;
    while (c2 * 2 <= cols_per_cta) c2 <<= 1;           // power of two <= share
    p.ACC = c2 / p.n_pad;
    if (p.ACC > 4) p.ACC = 4;
    if (p.ACC < 1) { p.ACC = 1; want_occ = 1; }
    int cols = p.ACC * p.n_pad;
    p.tmem_cols = 32;
    while (p.tmem_cols < cols) p.tmem_cols <<= 1;
    if (p.tmem_cols * want_occ > 512) want_occ = 512 / p.tmem_cols;
  }
  if (tune_env("QBN_S1_SA")) {   // tuning: force the activation ring depth
    const int sa_new = atoi(tune_env("QBN_S1_SA"));
    smem += (size_t)(sa_new - p.SA) * (a_bytes + 16);
    p.SA = sa_new;
  }
  if (tune_env("QBN_S1_ACC")) { p.ACC = atoi(tune_env("QBN_S1_ACC")); }
  if (smem > cap) {
    qbn_set_error("qbn_conv_s1_fwd: tile does not fit shared memory (%zu bytes)", smem);
    return QBN_ERR_UNSUPPORTED;
  }
  static bool attr_set = false;
  if (!attr_set) {
    QBN_CUDA(cudaFuncSetAttribute(umma_conv_s1_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    attr_set = true;
  }
  int occ = (int)((227 * 1024) / (smem + 1024));
  if (occ > want_occ) occ = want_occ;
  if (occ > 2) occ = 2;
  if (occ < 1) occ = 1;
  if (p.dbg & 8) occ = 1;
  int grid = qbn_sm_count() * occ;
  if (grid > p.total_tiles) grid = p.total_tiles;
  if (p.dbg & 8192) {
    unsigned long long z32[32] = {0};
    cudaMemcpyToSymbol(g_s1_prof, z32, sizeof(z32));
  }
  umma_conv_s1_kernel<2><<<grid, NTHREADS_S1, smem, st>>>(p);  // a third resident CTA would need <= 75 registers (spills)
  QBN_CHECK_LAUNCH();
  if (p.dbg & 8192) {
    unsigned long long h[32];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(h, g_s1_prof, sizeof(h));
    const int tiles0 = p.total_tiles / grid;
    fprintf(stderr, "[s1 prof] C=%d N=%d grid=%d tiles/CTA=%d SA=%d SB=%d ACC=%d b_res=%d TG=%d | epi: pre %llu waitacc %llu tmemld %llu compute+store %llu | mma: wait_acc_empty %llu wait_a %llu issue %llu wait_b %llu | prod: bres %llu wait_a_empty %llu issueA %llu publishA %llu wait_b_empty %llu issueB %llu publishB %llu (cycles per tile)\n",
            C, N, grid, tiles0, p.SA, p.SB, p.ACC, p.b_res, p.TG, h[0] / tiles0, h[1] / tiles0, h[2] / tiles0, h[3] / tiles0, h[8] / tiles0, h[9] / tiles0,
            h[10] / tiles0, h[11] / tiles0, h[16] / tiles0, h[17] / tiles0, h[18] / tiles0, h[19] / tiles0, h[20] / tiles0, h[21] / tiles0, h[22] / tiles0);
  }
  return QBN_OK;
}
