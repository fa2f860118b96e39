// K5 (bf16 production path): tcgen05 / TMEM tensor-core kernels.
#include "common.cuh"
#include "umma.cuh"

namespace mps {
using namespace umma;

// ------------------------------------------------------------------------------------------
// Diagnostic: one 128 x N x K tile through the exact building blocks of the fused kernels:
//   A written by threads into the SWIZZLE_128B canonical layout (generic proxy + proxy fence),
//   B brought in pre-swizzled by one bulk async copy (TMA engine) signalled on an mbarrier,
//   K/16 tcgen05.mma steps issued by one thread, tcgen05.commit -> mbarrier,
//   epilogue tcgen05.ld 32x32b (thread = TMEM lane = output row).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
selftest_umma_kernel(const uint16_t* __restrict__ a, const uint8_t* __restrict__ b_packed, float* __restrict__ d,
                     int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_b, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int chunks = K / 64;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)chunks * 16384;

  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 256);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(&bar_b, 1);
    mbar_init(&bar_mma, 1);
    mbar_fence_init();
  }
  for (int k0 = 0; k0 < K; k0 += 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(a + (size_t)tid * K + k0);
    store_a8(sA, 128, tid, k0, v);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;

  if (tid == 0) {
    const uint32_t bytes = (uint32_t)chunks * (uint32_t)N * 128u;
    mbar_arrive_expect_tx(&bar_b, bytes);
    bulk_g2s(sB, b_packed, bytes, &bar_b);
    mbar_wait(&bar_b, 0);
    tc_fence_after();
    const uint32_t idesc = instr_desc_bf16(N);
    for (int c = 0; c < chunks; ++c) {
      const uint32_t a0 = smem_u32(sA + (size_t)c * 16384);
      const uint32_t b0 = smem_u32(sB + (size_t)c * N * 128);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        mma_bf16_ss(tm, smem_desc_sw128(a0 + k4 * 32), smem_desc_sw128(b0 + k4 * 32), idesc, (c | k4) ? 1u : 0u);
      }
    }
    mma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c = 0; c < N; c += 16) {
    float v[16];
    tmem_ld_x16(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) d[(size_t)tid * N + c + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

}  // namespace mps

extern "C" int mpsnerf_selftest_umma(const uint16_t* a, const uint8_t* b_packed, float* d, int N, int K,
                                     void* stream) {
  MPS_REQUIRE(a && b_packed && d);
  MPS_REQUIRE(K >= 64 && K % 64 == 0 && N >= 16 && N <= 256 && N % 16 == 0);
  const size_t smem = (size_t)(K / 64) * (16384 + (size_t)N * 128);
  MPS_REQUIRE(smem <= 200 * 1024);
  MPS_CUDA(cudaFuncSetAttribute(mps::selftest_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mps::selftest_umma_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a, b_packed, d, N, K);
  MPS_LAUNCH_CHECK();
  return MPSNERF_OK;
}

extern "C" size_t mpsnerf_dense_bf16_workspace(int64_t count, int n_views) {
  (void)count; (void)n_views;
  return 256;
}

extern "C" int mpsnerf_dense_bf16(const float* tokens, int32_t ld, const float* xc, int64_t count,
                                  int n_views, const void* packed, size_t packed_bytes,
                                  const int32_t* act_pid, int64_t first, float* raw, void* workspace,
                                  void* stream) {
  (void)tokens; (void)ld; (void)xc; (void)count; (void)n_views; (void)packed; (void)packed_bytes;
  (void)act_pid; (void)first; (void)raw; (void)workspace; (void)stream;
  mps::set_error("mpsnerf_dense_bf16: not built in this revision");
  return MPSNERF_EINVAL;
}
