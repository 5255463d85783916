// One tcgen05.mma kind::tf32 on host-built shared-memory images: a probe of the MN-major operand layouts (diagnostics only).
#include "umma_common.cuh"

__global__ void __launch_bounds__(128, 1) mn_probe_kernel(const uint4* a_img, const uint4* b_img, int a_bytes, int b_bytes, unsigned long long adesc_t,
                                                          unsigned long long bdesc_t, int a_off, int b_off, unsigned idesc, int n_cols, int n_kstep,
                                                          int a_kadv, int b_kadv, float* D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sa = smem;                       // A image at 0, B image at 64 KB (both 1024-byte aligned)
  uint8_t* sb = smem + 65536;
  for (int i = tid; i < a_bytes / 16; i += 128) reinterpret_cast<uint4*>(sa)[i] = a_img[i];
  for (int i = tid; i < b_bytes / 16; i += 128) reinterpret_cast<uint4*>(sb)[i] = b_img[i];
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    for (int ks = 0; ks < n_kstep; ++ks) {
      const uint64_t ad = adesc_t | (uint64_t)(((smem_u32(sa) + a_off + ks * a_kadv) >> 4) & 0x3FFF);
      const uint64_t bd = bdesc_t | (uint64_t)(((smem_u32(sb) + b_off + ks * b_kadv) >> 4) & 0x3FFF);
      umma_mma<MODE_EVAL>(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(smem_u32(&bar));
  }
  if ((tid & 31) == 0) mbar_wait(smem_u32(&bar), 0);
  __syncwarp();
  tc_fence_after();
  for (int c = 0; c < n_cols; c += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[tid * n_cols + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

extern "C" int mn_probe(const void* a_img, const void* b_img, int a_bytes, int b_bytes, unsigned long long adesc_t, unsigned long long bdesc_t, int a_off,
                        int b_off, unsigned idesc, int n_cols, int n_kstep, int a_kadv, int b_kadv, float* D) {
  static bool set = false;
  if (!set) { cudaFuncSetAttribute(mn_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072); set = true; }
  mn_probe_kernel<<<1, 128, 131072>>>((const uint4*)a_img, (const uint4*)b_img, a_bytes, b_bytes, adesc_t, bdesc_t, a_off, b_off, idesc, n_cols, n_kstep,
                                      a_kadv, b_kadv, D);
  return (int)cudaGetLastError();
}
