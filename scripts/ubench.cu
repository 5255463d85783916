// Diagnostic micro-benchmark of tcgen05 instruction costs on this device (one CTA, clock64).
// Not on any product path; used to size the pipelines of umma_conv*.cu (numbers quoted in DESIGN.md).
#include "umma_common.cuh"

namespace {

__global__ void __launch_bounds__(128) ubench_kernel(unsigned long long* out, int n_cols, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1);
    fence_mbar_init();
    fence_proxy_async();
  }
  for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem) + 24 * 1024;
  const uint32_t lbo_a = 129 * 16, lbo_b = (uint32_t)(n_cols + 1) * 16;
  if (warp == 0) {
    const uint32_t leader = lane == 0;
    uint32_t phase = 0;
    long long t0, t1;
    // (0) clock overhead
    t0 = clock64(); t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
    // (1) tcgen05.fence::after_thread_sync
    t0 = clock64();
    for (int i = 0; i < reps; ++i) tc_fence_after();
    t1 = clock64();
    if (lane == 0) out[1] = (t1 - t0) / reps;
    // (2) fence.proxy.async
    t0 = clock64();
    for (int i = 0; i < reps; ++i) fence_proxy_async();
    t1 = clock64();
    if (lane == 0) out[2] = (t1 - t0) / reps;
    // (3) commit with nothing outstanding + wait for its arrival (round trip)
    t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      umma_commit_pred(smem_u32(&bar[0]), leader);
      if (lane == 0) mbar_spin(smem_u32(&bar[0]), phase);
      __syncwarp();
      phase ^= 1;
    }
    t1 = clock64();
    if (lane == 0) out[3] = (t1 - t0) / reps;
    // (4) issue cost of `reps` dependent MMAs (same accumulator), aligned operands
    uint64_t ad = make_smem_desc(a_addr, lbo_a, 128), bd = make_smem_desc(b_addr, lbo_b, 128);
    t0 = clock64();
    for (int i = 0; i < reps; ++i) umma_mma_tf32_pred(tm, ad, bd, idesc, 1u, leader);
    t1 = clock64();
    umma_commit_pred(smem_u32(&bar[1]), leader);
    if (lane == 0) mbar_spin(smem_u32(&bar[1]), 0);
    __syncwarp();
    long long t2 = clock64();
    if (lane == 0) { out[4] = (t1 - t0) / reps; out[5] = (t2 - t0) / reps; }
    // (5) same, A start address shifted by 3 rows (48 bytes): misaligned core matrices
    ad = make_smem_desc(a_addr + 48, lbo_a, 128);
    t0 = clock64();
    for (int i = 0; i < reps; ++i) umma_mma_tf32_pred(tm, ad, bd, idesc, 1u, leader);
    t1 = clock64();
    umma_commit_pred(smem_u32(&bar[2]), leader);
    if (lane == 0) mbar_spin(smem_u32(&bar[2]), 0);
    __syncwarp();
    t2 = clock64();
    if (lane == 0) { out[6] = (t1 - t0) / reps; out[7] = (t2 - t0) / reps; }
    // (6) independent accumulators (round-robin over 4 column ranges)
    ad = make_smem_desc(a_addr, lbo_a, 128);
    t0 = clock64();
    for (int i = 0; i < reps; ++i) umma_mma_tf32_pred(tm + (uint32_t)((i & 3) * n_cols), ad, bd, idesc, 1u, leader);
    t1 = clock64();
    umma_commit_pred(smem_u32(&bar[3]), leader);
    if (lane == 0) mbar_spin(smem_u32(&bar[3]), 0);
    __syncwarp();
    t2 = clock64();
    if (lane == 0) { out[8] = (t1 - t0) / reps; out[9] = (t2 - t0) / reps; }
    // (7) LDTM x16 + wait round trip
    uint32_t v[16];
    t0 = clock64();
    for (int i = 0; i < reps; ++i) { tmem_ld16(tm, v); tmem_ld_wait(); }
    t1 = clock64();
    if (lane == 0) out[10] = (t1 - t0) / reps + (v[0] & 0);
    // (8) commit issue cost only (arrivals pile up on a barrier nobody waits on)
    t0 = clock64();
    for (int i = 0; i < reps; ++i) umma_commit_pred(smem_u32(&bar[0]), leader);
    t1 = clock64();
    if (lane == 0) out[11] = (t1 - t0) / reps;
  }
  __syncthreads();
  {
    // (9) all four warps issue dependent MMA chains concurrently, each into its own accumulator
    __shared__ uint64_t bar2[4];
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar2[i]), 1); fence_mbar_init(); }
    __syncthreads();
    const uint32_t leader = lane == 0;
    uint64_t ad = make_smem_desc(a_addr, lbo_a, 128), bd = make_smem_desc(b_addr, lbo_b, 128);
    long long t0 = clock64();
    for (int i = 0; i < reps; ++i) umma_mma_tf32_pred(tm + (uint32_t)(warp * n_cols), ad, bd, idesc, 1u, leader);
    long long t1 = clock64();
    umma_commit_pred(smem_u32(&bar2[warp]), leader);
    if (lane == 0) mbar_spin(smem_u32(&bar2[warp]), 0);
    __syncwarp();
    long long t2 = clock64();
    if (lane == 0) { out[12 + warp] = ((unsigned long long)((t1 - t0) / reps) << 32) | (unsigned long long)((t2 - t0) / reps); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

// Several CTAs per SM, `n_warps` MMA issuer warps per CTA, every issuer running a dependent chain into its own
// accumulator: does the tensor pipe overlap MMAs of DIFFERENT CTAs as well as those of different warps of ONE CTA?
__global__ void __launch_bounds__(128) ubench_multi_kernel(unsigned long long* out, int n_cols, int reps, int n_warps, int tmem_cols) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1);
    fence_mbar_init();
    fence_proxy_async();
  }
  for (int i = tid; i < 40 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)tmem_cols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t lbo_a = 136 * 16, lbo_b = (uint32_t)n_cols * 16;
  if (warp < n_warps) {
    const uint32_t leader = lane == 0;
    const uint32_t a_addr = smem_u32(smem) + (uint32_t)warp * 2 * lbo_a, b_addr = smem_u32(smem) + 24 * 1024;
    uint64_t ad = make_smem_desc(a_addr, lbo_a, 128), bd = make_smem_desc(b_addr, lbo_b, 128);
    long long t0 = clock64();
    for (int i = 0; i < reps; ++i) umma_mma_tf32_pred(tm + (uint32_t)(warp * n_cols), ad, bd, idesc, 1u, leader);
    long long t1 = clock64();
    umma_commit_pred(smem_u32(&bar[warp]), leader);
    if (lane == 0) mbar_spin(smem_u32(&bar[warp]), 0);
    __syncwarp();
    long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[warp] = ((unsigned long long)((t1 - t0) / reps) << 32) | (unsigned long long)((t2 - t0) / reps);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, (uint32_t)tmem_cols); }
}

}  // namespace

extern "C" int qbn_ubench_tcgen05_multi(unsigned long long* out_dev /* >= 4 */, int n_cols, int reps, int ctas_per_sm, int n_warps, void* stream) {
  QBN_CHECK_ARG(out_dev && n_cols >= 16 && n_cols <= 128 && n_cols % 16 == 0 && reps > 0, "args");
  QBN_CHECK_ARG(ctas_per_sm >= 1 && ctas_per_sm <= 4 && n_warps >= 1 && n_warps <= 4, "ctas_per_sm, n_warps in 1..4");
  int tmem_cols = 32;
  while (tmem_cols < n_warps * n_cols) tmem_cols <<= 1;
  QBN_CHECK_ARG(tmem_cols * ctas_per_sm <= 512, "TMEM");
  const size_t smem = (size_t)(200 * 1024) / ctas_per_sm;          // forces exactly ctas_per_sm resident CTAs
  QBN_CUDA(cudaFuncSetAttribute(ubench_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  ubench_multi_kernel<<<qbn_sm_count() * ctas_per_sm, 128, smem, (cudaStream_t)stream>>>(out_dev, n_cols, reps, n_warps, tmem_cols);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

extern "C" int qbn_ubench_tcgen05(unsigned long long* out_dev /* >= 16 */, int n_cols, int reps, void* stream) {
  QBN_CHECK_ARG(out_dev && n_cols >= 16 && n_cols <= 128 && n_cols % 16 == 0 && reps > 0, "args");
  QBN_CUDA(cudaFuncSetAttribute(ubench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  ubench_kernel<<<1, 128, 48 * 1024 + 1024, (cudaStream_t)stream>>>(out_dev, n_cols, reps);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}
