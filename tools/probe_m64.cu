// Probe for next round's T-kernel design: where does an M = 64 tcgen05.mma (cta_group::1)
// put its 64 accumulator rows in tensor memory, and does it accept a D address with lane offset 16 -- i.e. can two
// 64-row tiles share the same columns in the lower / upper 16 lanes of each 32-lane quarter?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I mps-nerf_b200/csrc -I include \
//        tools/probe_m64.cu -o build/probe_m64 && build/probe_m64
//
// A (64 x 16, smem, K-major SWIZZLE_128B) has A[r][0] = r + 1, B (64 x 16) has B[n][0] = n + 1, so D[r][n] =
// (r + 1)(n + 1): reading all 128 lanes x 64 columns back shows which lane holds which row.  Each case runs in its
// own launch so that a trap in one (an address the hardware rejects) does not hide the others.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "umma.cuh"

using namespace mps::umma;

__global__ void __launch_bounds__(128, 1) probe(float* out, int m, int lane_off, int zero_first) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* A = smem;             // 128 rows x 128 B
  uint8_t* B = smem + 16384;     // 64 rows x 128 B
  for (int i = tid; i < (16384 + 8192) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  if (tid < m) {                 // element k = 0 of row r lives in 16-byte unit (0 ^ (r & 7))
    __nv_bfloat16 v = __float2bfloat16((float)(tid + 1));
    *reinterpret_cast<__nv_bfloat16*>(A + tid * 128 + ((0 ^ (tid & 7)) << 4)) = v;
  }
  if (tid < 64) {
    __nv_bfloat16 v = __float2bfloat16((float)(tid + 1));
    *reinterpret_cast<__nv_bfloat16*>(B + tid * 128 + ((0 ^ (tid & 7)) << 4)) = v;
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base_s, 64); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16);
  if (zero_first) {
    uint32_t z[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) z[c] = __float_as_uint(-7.0f);      // sentinel: "not written by the MMA"
    tmem_st_u32(tl, z);
    tmem_st_u32(tl + 32, z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    mma_bf16_ss(tm + ((uint32_t)lane_off << 16), smem_desc_sw128(smem_u32(A)), smem_desc_sw128(smem_u32(B)),
                instr_desc_bf16(64, m), 0u);
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  float v[32];
  tmem_ld_x32(tl, v);
  tmem_ld_wait();
  for (int c = 0; c < 32; ++c) out[tid * 64 + c] = v[c];
  tmem_ld_x32(tl + 32, v);
  tmem_ld_wait();
  for (int c = 0; c < 32; ++c) out[tid * 64 + 32 + c] = v[c];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 64);
}

// TS form (A operand in tensor memory, as in the fused kernels): A row r is written to lane 32 (r / 16) + r % 16 +
// lane_off, columns [64, 72) (K = 16 bf16 = 8 packed columns), D goes to columns [0, 64) at the same lane offset.
__global__ void __launch_bounds__(128, 1) probe_ts(float* out, int lane_off) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* B = smem;
  for (int i = tid; i < 8192 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  if (tid < 64) {
    __nv_bfloat16 v = __float2bfloat16((float)(tid + 1));
    *reinterpret_cast<__nv_bfloat16*>(B + tid * 128 + ((0 ^ (tid & 7)) << 4)) = v;
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base_s, 128); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16);
  {
    uint32_t z[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) z[c] = __float_as_uint(-7.0f);
    tmem_st_u32(tl, z);
    tmem_st_u32(tl + 32, z);
    // A: every lane writes its 8 packed columns; lanes outside the tile's 16-lane window hold a poison row value
    const bool mine = (lane >= lane_off) && (lane < lane_off + 16);
    const float rowv = mine ? (float)(16 * warp + (lane - lane_off) + 1) : 999.0f;
    uint32_t a8[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) a8[c] = 0;
    a8[0] = pack_bf16x2(rowv, 0.f);
    tmem_st_x8(tl + 64, a8);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    mma_bf16_ts(tm + ((uint32_t)lane_off << 16), tm + 64 + ((uint32_t)lane_off << 16), smem_desc_sw128(smem_u32(B)),
                instr_desc_bf16(64, 64), 0u);
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  float v[32];
  tmem_ld_x32(tl, v);
  tmem_ld_wait();
  for (int c = 0; c < 32; ++c) out[tid * 64 + c] = v[c];
  tmem_ld_x32(tl + 32, v);
  tmem_ld_wait();
  for (int c = 0; c < 32; ++c) out[tid * 64 + 32 + c] = v[c];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 128);
}

// cycles per MMA (K = 16 step) for M in {64, 128} x N: 512 MMAs back to back from one thread, TS form.
template <int kM, int kN>
__global__ void __launch_bounds__(128, 1) probe_time(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32768 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  if (tid == 0) {
    const uint64_t bdesc = smem_desc_sw128(smem_u32(smem));
    const long long t0 = clock64();
    for (int i = 0; i < 512; ++i) mma_bf16_ts(tm, tm + 256 + (i & 3) * 8, bdesc + (uint64_t)((i & 3) * 2), instr_desc_bf16(kN, kM), 1u);
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    out[0] = clock64() - t0;
  } else {
    mbar_wait(&bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int kM, int kN>
static void run_time() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(probe_time<kM, kN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  probe_time<kM, kN><<<1, 128, 32768>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("M=%3d N=%3d K=16 (A in TMEM): %s, %.1f cycles per MMA\n", kM, kN, e == cudaSuccess ? "ok" : cudaGetErrorString(e), h / 512.0);
  cudaFree(d);
}

static void run(int m, int lane_off, bool ts = false) {
  float* d;
  cudaMalloc(&d, 128 * 64 * 4);
  cudaMemset(d, 0, 128 * 64 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  cudaFuncSetAttribute(probe_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  if (ts) probe_ts<<<1, 128, 32768>>>(d, lane_off);
  else probe<<<1, 128, 32768>>>(d, m, lane_off, 1);
  cudaError_t e = cudaDeviceSynchronize();
  printf("== M=%d, %s, D%s lane offset %d: %s\n", m, ts ? "A in TMEM" : "A in smem", ts ? " and A" : "", lane_off,
         e == cudaSuccess ? "ok" : cudaGetErrorString(e));
  if (e != cudaSuccess) { exit(0); }       // the context is gone after a trap: stop here
  std::vector<float> h(128 * 64);
  cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
  // row held by each lane (from column 0: value = row + 1; sentinel -7 = untouched), and a consistency check over columns
  int bad = 0;
  printf("lane->row: ");
  for (int l = 0; l < 128; ++l) {
    const float v0 = h[l * 64];
    const int row = v0 == -7.0f ? -1 : (int)v0 - 1;
    if (row >= 0)
      for (int c = 0; c < 64; ++c) bad += h[l * 64 + c] != (float)((row + 1) * (c + 1));
    if (l % 16 == 0) printf("| ");
    if (row < 0) printf(". "); else printf("%d ", row);
  }
  printf("\ncolumn-consistency violations: %d\n", bad);
  cudaFree(d);
}

int main() {
  run(128, 0);
  run(64, 0);
  run(64, 16);
  run(64, 0, true);
  run(64, 16, true);
  run_time<128, 192>(); run_time<64, 192>(); run_time<128, 128>(); run_time<64, 128>(); run_time<64, 64>(); run_time<128, 64>();
  return 0;
}
