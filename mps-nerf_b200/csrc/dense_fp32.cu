// K5 (fp32 option): cross-view transformer + canonical NeRF MLP on CUDA cores.
//
// Restates Transformer.forward (lib/transformer.py:13-86) and the MLP of
// SKinningBatch.forward (lib/skinnning_batch.py:438-473) layer by layer with fp32 FMA
// accumulation.  This is the high-precision mode (rgb within 1e-4 of the reference); the
// production path is the bf16 tcgen05 kernel in dense_tc.cu.
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace mps {

// ------------------------------------------------------------------ generic SIMT linear
// Y[m, n] = act(sum_k X[m,k] * W[n,k] + b[n]) (+ Res[m,n]);  W is torch nn.Linear layout.
constexpr int BM = 128, BN = 64, BK = 16, kLinThreads = 256;

template <int ACT /*0 none, 1 relu, 2 gelu(erf)*/>
__global__ void __launch_bounds__(kLinThreads)
linear_f32_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int K,
                  const float* __restrict__ b, const float* Res, int ldr, float* Y,
                  int ldy, int64_t M, int N) {
  __shared__ float Xs[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;          // 16 x 16 threads; each 8 rows x 4 cols
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // X tile: 128 x 16 = 2048 elements, 8 per thread; k fastest for coalescing
    for (int e = tid; e < BM * BK; e += kLinThreads) {
      const int kk = e % BK, mm = e / BK;
      const int64_t m = m0 + mm;
      const int k = k0 + kk;
      Xs[kk][mm] = (m < M && k < K) ? X[m * ldx + k] : 0.f;
    }
    for (int e = tid; e < BN * BK; e += kLinThreads) {
      const int kk = e % BK, nn = e / BK;
      const int n = n0 + nn, k = k0 + kk;
      Ws[kk][nn] = (n < N && k < K) ? W[(size_t)n * K + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], w[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = Xs[kk][ty * 8 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (b ? b[n] : 0.f);
      if (ACT == 1) v = fmaxf(v, 0.f);
      if (ACT == 2) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
      if (Res) v += Res[m * ldr + n];
      Y[m * ldy + n] = v;
    }
  }
}

// ------------------------------------------------------------------ the same linear on the tensor cores, fp32-accurate
// 3xTF32: every fp32 operand is split into a TF32 "hi" part and a TF32 "lo" remainder (x = hi + lo exactly to 2^-22
// relative), and a product is formed as  a_lo b_hi + a_hi b_lo + a_hi b_hi  with fp32 accumulation in the tensor core
// (mma.sync.m16n8k8, the warp-level path -- the operands live in registers, which is what makes the split cheap; the
// dropped a_lo b_lo term is 2^-22 of the product).  Error ~1e-6 relative: the <= 1e-4 rgb bound of the high-precision
// option holds with the same margin as on the CUDA cores (tests: fp32 goldens, dense train fwd / bwd), at 3 tensor-core
// MMAs per product instead of one FMA per MAC.  Block tile 128 x 64 x 16, 8 warps as 4 (M) x 2 (N), warp tile 32 x 32.
constexpr int kTcPad = 8;      // row stride 136 / 72 floats: the fragment loads (8 t + g) hit 32 distinct banks

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float rem = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(rem));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int ACT /*0 none, 1 relu, 2 gelu(erf)*/>
__global__ void __launch_bounds__(kLinThreads)
linear_tf32x3_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int K,
                     const float* __restrict__ b, const float* Res, int ldr, float* Y,
                     int ldy, int64_t M, int N) {
  __shared__ float Xs[BK][BM + kTcPad];
  __shared__ float Ws[BK][BN + kTcPad];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    for (int e = tid; e < BM * BK; e += kLinThreads) {
      const int kk = e % BK, mm = e / BK;
      const int64_t m = m0 + mm;
      const int k = k0 + kk;
      Xs[kk][mm] = (m < M && k < K) ? X[m * ldx + k] : 0.f;
    }
    for (int e = tid; e < BN * BK; e += kLinThreads) {
      const int kk = e % BK, nn = e / BK;
      const int n = n0 + nn, k = k0 + kk;
      Ws[kk][nn] = (n < N && k < K) ? W[(size_t)n * K + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < BK; ks += 8) {
      uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int row = wm * 32 + mt * 16 + g;
        split_tf32(Xs[ks + t][row], ah[mt][0], al[mt][0]);
        split_tf32(Xs[ks + t][row + 8], ah[mt][1], al[mt][1]);
        split_tf32(Xs[ks + t + 4][row], ah[mt][2], al[mt][2]);
        split_tf32(Xs[ks + t + 4][row + 8], ah[mt][3], al[mt][3]);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = wn * 32 + nt * 8 + g;
        split_tf32(Ws[ks + t][col], bh[nt][0], bl[nt][0]);
        split_tf32(Ws[ks + t + 4][col], bh[nt][1], bl[nt][1]);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          mma_tf32(acc[mt][nt], al[mt], bh[nt]);      // the small terms first
          mma_tf32(acc[mt][nt], ah[mt], bl[nt]);
          mma_tf32(acc[mt][nt], ah[mt], bh[nt]);
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int64_t m = m0 + wm * 32 + mt * 16 + g + ((e & 2) ? 8 : 0);
        const int n = n0 + wn * 32 + nt * 8 + 2 * t + (e & 1);
        if (m >= M || n >= N) continue;
        float v = acc[mt][nt][e] + (b ? b[n] : 0.f);
        if (ACT == 1) v = fmaxf(v, 0.f);
        if (ACT == 2) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
        if (Res) v += Res[m * ldr + n];
        Y[m * ldy + n] = v;
      }
}

// ------------------------------------------------------------------ the same 3xTF32 linear on the tcgen05 tensor cores
// Tile = 128 rows x kN outputs, K in chunks of 32 (one 128-byte SWIZZLE_128B row per operand row).  Every chunk is
// staged in shared memory four times over -- A_hi, A_lo, B_hi, B_lo, hi = the value truncated to TF32, lo = the exact
// remainder -- by a coalesced cooperative load (a warp reads one row's 128 bytes per instruction), and one thread issues
// 4 K-steps x 3 MMAs (lo.hi, hi.lo, hi.hi; kind::tf32, SS form, fp32 accumulator in tensor memory).  Two stages:
// the loads of chunk c + 1 run under the MMAs of chunk c; a stage is reused when the commit of the MMAs that read
// it has arrived.  Epilogue: thread = row (TMEM lane), tcgen05.ld 16 columns at a time -> bias / activation /
// residual -> global.  This is a plain tiled GEMM, not a fused kernel: the high-precision option is the
// validation / training arithmetic, launched layer by layer.
constexpr int kGRows = 128, kGChunk = 32, kGThreads = 512;      // 16 warps stage the operands, warps 0-3 drain TMEM
template <int kN> __host__ __device__ constexpr uint32_t g_stage_bytes() { return 2u * 128u * 128u + 2u * (uint32_t)kN * 128u; }

__device__ __forceinline__ uint32_t sw128_f32_off(int r, int k) {      // byte offset of element (row r, K index k < 32)
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 2) ^ (r & 7))) << 4) + (k & 3) * 4);
}

template <int ACT, int kN>
__global__ void __launch_bounds__(kGThreads, 1)
linear_tcgen05_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int K,
                      const float* __restrict__ b, const float* Res, int ldr, float* Y, int ldy, int64_t M, int N) {
  using namespace umma;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr uint32_t kA = 128u * 128u, kB = (uint32_t)kN * 128u, kStage = g_stage_bytes<kN>();
  constexpr uint32_t kCols = kN <= 64 ? 64u : (kN <= 128 ? 128u : 256u);      // TMEM allocations are powers of two
  if (warp == 0) { tmem_alloc(&tmem_base_s, kCols); tmem_relinquish(); }
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  const int64_t m0 = (int64_t)blockIdx.x * kGRows;
  const int n0 = blockIdx.y * kN;
  const int nchunks = (K + kGChunk - 1) / kGChunk;
  constexpr uint32_t idesc = instr_desc_tf32(kN);
  uint32_t ph[2] = {0u, 0u};
  // Thread (kk = tid & 31, r0 = tid >> 5) stages column kk of rows r0, r0 + 16, ...: a warp reads 128 contiguous
  // bytes of one row per load, and because 16 is a multiple of the 8-row swizzle period the shared-memory offset of a
  // thread's elements advances by a constant 2048 bytes -- no per-element address arithmetic.  The loads of chunk
  // c + 1 are issued into registers before the barrier that releases the MMAs of chunk c, so the global-memory
  // latency runs under those MMAs instead of in front of them.
  const int kk = tid & 31, r0 = tid >> 5;
  const uint32_t sw = sw128_f32_off(r0, kk);
  float ra[kGRows / 16], rb[kN / 16];
  auto fetch = [&](int c) {
    const int k = c * kGChunk + kk;
    const float* xp = X + (m0 + r0) * ldx + k;
#pragma unroll
    for (int i = 0; i < kGRows / 16; ++i)                            // A: rows of X
      ra[i] = (m0 + r0 + 16 * i < M && k < K) ? xp[(int64_t)(16 * i) * ldx] : 0.f;
    const float* wp = W + (size_t)(n0 + r0) * K + k;
#pragma unroll
    for (int i = 0; i < kN / 16; ++i)                                // B: rows of W (nn.Linear layout: (N, K), K-major)
      rb[i] = (n0 + r0 + 16 * i < N && k < K) ? wp[(size_t)(16 * i) * K] : 0.f;
  };
  fetch(0);
  for (int c = 0; c < nchunks; ++c) {
    const int s = c & 1;
    if (c >= 2) { mbar_wait(&bar[s], ph[s]); ph[s] ^= 1u; }      // the MMAs of chunk c - 2 have read stage s
    uint8_t* st = smem + (size_t)s * kStage;
#pragma unroll
    for (int i = 0; i < kGRows / 16; ++i) {
      const float hi = __uint_as_float(__float_as_uint(ra[i]) & 0xffffe000u);
      *reinterpret_cast<float*>(st + sw + 2048u * i) = hi;
      *reinterpret_cast<float*>(st + kA + sw + 2048u * i) = ra[i] - hi;
    }
#pragma unroll
    for (int i = 0; i < kN / 16; ++i) {
      const float hi = __uint_as_float(__float_as_uint(rb[i]) & 0xffffe000u);
      *reinterpret_cast<float*>(st + 2 * kA + sw + 2048u * i) = hi;
      *reinterpret_cast<float*>(st + 2 * kA + kB + sw + 2048u * i) = rb[i] - hi;
    }
    if (c + 1 < nchunks) fetch(c + 1);
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_hi = smem_u32(st), a_lo = a_hi + kA, b_hi = a_hi + 2 * kA, b_lo = b_hi + kB;
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const uint64_t dah = smem_desc_sw128(a_hi + k4 * 32), dal = smem_desc_sw128(a_lo + k4 * 32);
        const uint64_t dbh = smem_desc_sw128(b_hi + k4 * 32), dbl = smem_desc_sw128(b_lo + k4 * 32);
        mma_tf32_ss(tm, dal, dbh, idesc, (c | k4) ? 1u : 0u);       // the small terms first
        mma_tf32_ss(tm, dah, dbl, idesc, 1u);
        mma_tf32_ss(tm, dah, dbh, idesc, 1u);
      }
      mma_commit(&bar[s]);
    }
  }
  {
    const int s = (nchunks - 1) & 1;       // commits arrive in issue order: the last one covers every MMA
    mbar_wait(&bar[s], ph[s]);
  }
  tc_fence_after();
  // Epilogue.  A thread owns a row (its TMEM lane); writing rows straight to global would scatter every store
  // instruction over 32 rows, so the tile goes through shared memory (the operand stages are free: every MMA has
  // completed) and leaves with coalesced stores, bias / activation / residual applied on the way out.
  float* tile = reinterpret_cast<float*>(smem);              // [128][kN + 1]
  if (warp < 4) {                                            // warps 0-3 own the four TMEM lane quarters
    for (int cb = 0; cb < kN; cb += 16) {
      float v[16];
      tmem_ld_x16(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) tile[tid * (kN + 1) + cb + i] = v[i];
    }
  }
  __syncthreads();
  constexpr int kBatch = 8;      // residual loads of 8 elements in flight per thread before the first store
  for (int e0 = tid; e0 < kGRows * kN; e0 += kGThreads * kBatch) {
    float res[kBatch];
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      const int e = e0 + j * kGThreads, rr = e / kN, cc = e - rr * kN;
      const bool ok = e < kGRows * kN && m0 + rr < M && n0 + cc < N;
      res[j] = (ok && Res) ? Res[(m0 + rr) * ldr + n0 + cc] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      const int e = e0 + j * kGThreads, rr = e / kN, cc = e - rr * kN;
      const int n = n0 + cc;
      if (e >= kGRows * kN || m0 + rr >= M || n >= N) continue;
      float y = tile[rr * (kN + 1) + cc] + (b ? b[n] : 0.f);
      if (ACT == 1) y = fmaxf(y, 0.f);
      if (ACT == 2) y = 0.5f * y * (1.f + erff(y * 0.70710678118654752440f));
      Y[(m0 + rr) * ldy + n] = y + res[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, kCols);
}

// The 1- and 3-wide heads (alpha, rgb): a warp per row, lanes stride over K (128 contiguous bytes per load), exact fp32
// FMAs, shuffle reduction.  On a 16 x 8 MMA tile these layers waste 7/8 of the tensor work and took as long as a
// 256 x 256 layer (47 us per 32 768 rows); they are bound by reading X once (33 MB: a few microseconds).
template <int ACT>
__global__ void __launch_bounds__(256)
linear_narrow_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int K, const float* __restrict__ b,
                     const float* Res, int ldr, float* Y, int ldy, int64_t M, int N) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= M) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float* x = X + row * ldx;
  for (int k = lane; k < K; k += 32) {
    const float v = x[k];
#pragma unroll
    for (int n = 0; n < 4; ++n)
      if (n < N) acc[n] = fmaf(v, __ldg(W + (size_t)n * K + k), acc[n]);
  }
#pragma unroll
  for (int n = 0; n < 4; ++n)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
  if (lane == 0) {
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      if (n >= N) break;
      float v = acc[n] + (b ? b[n] : 0.f);
      if (ACT == 1) v = fmaxf(v, 0.f);
      if (ACT == 2) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
      if (Res) v += Res[row * ldr + n];
      Y[row * ldy + n] = v;
    }
  }
}

// MPSNERF_FP32_GEMM = tcgen05 (default: 3xTF32, kind::tf32 tcgen05 tiles) | mma (3xTF32, warp-level mma.sync) |
// simt (plain fp32 FMAs on the CUDA cores)
static int fp32_gemm_mode() {
  static int mode = -1;
  if (mode < 0) { const char* e = getenv("MPSNERF_FP32_GEMM"); mode = !e ? 2 : (e[0] == 's' ? 0 : (e[0] == 'm' ? 1 : 2)); }
  return mode;
}

template <int ACT>
static int launch_linear(const float* X, int ldx, const float* W, int K, const float* b, const float* Res, int ldr,
                         float* Y, int ldy, int64_t M, int N, cudaStream_t st) {
  const int mode = fp32_gemm_mode();
  if (mode == 2 && N >= 32) {      // (the 1- and 3-wide heads: linear_narrow_kernel below)
    if (N > 160) {                 // wide layers: 256 outputs per tile, the activation rows are staged half as often
      constexpr int kN = 256;
      const size_t smem = 2 * (size_t)g_stage_bytes<kN>();
      cudaFuncSetAttribute(linear_tcgen05_kernel<ACT, kN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      dim3 grid((unsigned)((M + kGRows - 1) / kGRows), (unsigned)((N + kN - 1) / kN));
      linear_tcgen05_kernel<ACT, kN><<<grid, kGThreads, smem, st>>>(X, ldx, W, K, b, Res, ldr, Y, ldy, M, N);
    } else if (N > 128) {          // the 155-wide token layers (to_out, net.3): one 160-column tile
      constexpr int kN = 160;
      const size_t smem = 2 * (size_t)g_stage_bytes<kN>();
      cudaFuncSetAttribute(linear_tcgen05_kernel<ACT, kN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      dim3 grid((unsigned)((M + kGRows - 1) / kGRows), (unsigned)((N + kN - 1) / kN));
      linear_tcgen05_kernel<ACT, kN><<<grid, kGThreads, smem, st>>>(X, ldx, W, K, b, Res, ldr, Y, ldy, M, N);
    } else if (N > 64) {
      constexpr int kN = 128;
      const size_t smem = 2 * (size_t)g_stage_bytes<kN>();
      cudaFuncSetAttribute(linear_tcgen05_kernel<ACT, kN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      dim3 grid((unsigned)((M + kGRows - 1) / kGRows), (unsigned)((N + kN - 1) / kN));
      linear_tcgen05_kernel<ACT, kN><<<grid, kGThreads, smem, st>>>(X, ldx, W, K, b, Res, ldr, Y, ldy, M, N);
    } else {
      constexpr int kN = 64;
      const size_t smem = 2 * (size_t)g_stage_bytes<kN>();
      cudaFuncSetAttribute(linear_tcgen05_kernel<ACT, kN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      dim3 grid((unsigned)((M + kGRows - 1) / kGRows), (unsigned)((N + kN - 1) / kN));
      linear_tcgen05_kernel<ACT, kN><<<grid, kGThreads, smem, st>>>(X, ldx, W, K, b, Res, ldr, Y, ldy, M, N);
    }
    return 0;
  }
  if (mode == 2 && N <= 4) {
    linear_narrow_kernel<ACT><<<(unsigned)((M * 32 + 255) / 256), 256, 0, st>>>(X, ldx, W, K, b, Res, ldr, Y, ldy, M, N);
    return 0;
  }
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
  if (mode >= 1)
    linear_tf32x3_kernel<ACT><<<grid, kLinThreads, 0, st>>>(X, ldx, W, K, b, Res, ldr, Y, ldy, M, N);
  else
    linear_f32_kernel<ACT><<<grid, kLinThreads, 0, st>>>(X, ldx, W, K, b, Res, ldr, Y, ldy, M, N);
  return 0;
}

// ------------------------------------------------------------------ LayerNorm (eps 1e-5), warp per row
__global__ void ln_rows_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ g,
                               const float* __restrict__ be, float* __restrict__ Y, int ldy, int64_t rows, int D) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = w0; r < rows; r += nw) {
    const float* x = X + r * ldx;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += x[c];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)D;
    float v = 0.f;
    for (int c = lane; c < D; c += 32) { const float d = x[c] - mean; v += d * d; }
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = rsqrtf(v / (float)D + 1e-5f);
    for (int c = lane; c < D; c += 32) Y[r * ldy + c] = (x[c] - mean) * rstd * g[c] + be[c];
  }
}

// ------------------------------------------------------------------ 4-head attention over the V view tokens
// QKV (count*V, 768): [q(256) | k(256) | v(256)], head h = columns h*64..h*64+63 of each part
// ('b n (h d) -> b h n d', lib/transformer.py:62).  One warp per point: lane l serves head l / 8 with the eight head
// dims 8 (l % 8) .. -- every load and store is a coalesced 128-bit access of a 1 KB row part, the q.k dots are reduced
// over the eight lanes of a head with three shuffle steps.  (Round 1 ran one thread per (point, head, query) over
// 768-float-strided rows: 17 ms of a 128 ms fp32 frame.)
__global__ void __launch_bounds__(256) attention_kernel(const float* __restrict__ QKV, float* __restrict__ O, int64_t count, int V) {
  const int lane = threadIdx.x & 31, h = lane >> 3, sub = lane & 7;
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int col = h * 64 + 8 * sub;
  for (int64_t p = w0; p < count; p += nw) {
    const float* base = QKV + p * V * 768 + col;
    for (int i = 0; i < V; ++i) {
      const float4 qa = *reinterpret_cast<const float4*>(base + i * 768), qb = *reinterpret_cast<const float4*>(base + i * 768 + 4);
      float dots[MPSNERF_MAX_VIEWS];
      float mx = -1e30f;
      for (int j = 0; j < V; ++j) {
        const float4 ka = *reinterpret_cast<const float4*>(base + j * 768 + 256), kb = *reinterpret_cast<const float4*>(base + j * 768 + 260);
        float d = qa.x * ka.x;
        d = fmaf(qa.y, ka.y, d); d = fmaf(qa.z, ka.z, d); d = fmaf(qa.w, ka.w, d);
        d = fmaf(qb.x, kb.x, d); d = fmaf(qb.y, kb.y, d); d = fmaf(qb.z, kb.z, d); d = fmaf(qb.w, kb.w, d);
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        dots[j] = d * 0.125f;                       // dim_head ** -0.5
        mx = fmaxf(mx, dots[j]);
      }
      float den = 0.f;
      for (int j = 0; j < V; ++j) { dots[j] = expf(dots[j] - mx); den += dots[j]; }
      float4 oa = make_float4(0.f, 0.f, 0.f, 0.f), ob = oa;
      for (int j = 0; j < V; ++j) {
        const float w = dots[j] / den;
        const float4 va = *reinterpret_cast<const float4*>(base + j * 768 + 512), vb = *reinterpret_cast<const float4*>(base + j * 768 + 516);
        oa.x = fmaf(w, va.x, oa.x); oa.y = fmaf(w, va.y, oa.y); oa.z = fmaf(w, va.z, oa.z); oa.w = fmaf(w, va.w, oa.w);
        ob.x = fmaf(w, vb.x, ob.x); ob.y = fmaf(w, vb.y, ob.y); ob.z = fmaf(w, vb.z, ob.z); ob.w = fmaf(w, vb.w, ob.w);
      }
      float* o = O + (p * V + i) * 256 + col;
      *reinterpret_cast<float4*>(o) = oa;
      *reinterpret_cast<float4*>(o + 4) = ob;
    }
  }
}

// ------------------------------------------------------------------ MLP input / skip / scatter helpers
// Hcat (count, 450): cols [0,39) = PE6(xc), [39,194) = token 0.   F (count, 411): cols [256,411) = token 1.
__global__ void mlp_inputs_kernel(const float* __restrict__ xc, const float* __restrict__ Xtok, int V,
                                  float* __restrict__ Hcat, float* __restrict__ F, int64_t count) {
  const int64_t total = count * 349;    // 39 + 155 + 155
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = t / 349;
    const int e = (int)(t % 349);
    if (e < 39) {
      const int ch = (e < 3) ? e : ((e - 3) % 3);
      const float x = xc[3 * p + ch];
      float val = x;
      if (e >= 3) {
        const int k = (e - 3) / 6;
        const bool is_cos = ((e - 3) % 6) >= 3;
        val = sinf(fmaf(x, 3.14159265358979323846f * (float)(1 << k), is_cos ? 1.57079632679489661923f : 0.0f));
      }
      Hcat[p * 450 + e] = val;
    } else if (e < 194) {
      Hcat[p * 450 + e] = Xtok[(p * V + 0) * 155 + (e - 39)];
    } else {
      F[p * 411 + 256 + (e - 194)] = Xtok[(p * V + 1) * 155 + (e - 194)];
    }
  }
}

__global__ void scatter_raw_kernel(const float* __restrict__ out4, const int32_t* __restrict__ act_pid,
                                   int64_t first, int64_t count, float* __restrict__ raw) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    reinterpret_cast<float4*>(raw)[act_pid[first + i]] = reinterpret_cast<const float4*>(out4)[i];
  }
}

__global__ void copy_rows_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd,
                                 int64_t rows, int cols) {
  const int64_t total = rows * cols;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    dst[(t / cols) * ldd + (t % cols)] = src[(t / cols) * lds + (t % cols)];
  }
}

static inline int blocks_for(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  return (int)(b > kNumSMs * 32 ? kNumSMs * 32 : (b < 1 ? 1 : b));
}

struct Fp32Ws {
  float *X, *Y, *QKV, *O, *Hff, *Hcat, *H1, *H2, *F, *G, *out4;
};

static size_t carve(Fp32Ws& w, char* base, int64_t count, int V) {
  size_t off = 0;
  auto take = [&](size_t nfloats) {
    float* p = reinterpret_cast<float*>(base + off);
    off += ((nfloats * sizeof(float) + 255) / 256) * 256;
    return p;
  };
  const size_t M3 = (size_t)count * V;
  w.X = take(M3 * 155); w.Y = take(M3 * 155); w.QKV = take(M3 * 768); w.O = take(M3 * 256); w.Hff = take(M3 * 128);
  w.Hcat = take((size_t)count * 450); w.H1 = take((size_t)count * 256); w.H2 = take((size_t)count * 256);
  w.F = take((size_t)count * 411); w.G = take((size_t)count * 128); w.out4 = take((size_t)count * 4);
  return off;
}

}  // namespace mps

extern "C" size_t mpsnerf_dense_fp32_workspace(int64_t count, int n_views) {
  mps::Fp32Ws w;
  return mps::carve(w, nullptr, count > 0 ? count : 0, n_views) + 256;
}

extern "C" int mpsnerf_dense_fp32(const float* tokens, int32_t ld, const float* xc, int64_t count,
                                  int n_views, const float* const* weights, const int32_t* act_pid,
                                  int64_t first, float* raw, void* workspace, void* stream) {
  using namespace mps;
  MPS_REQUIRE(count >= 0 && n_views >= 2 && n_views <= MPSNERF_MAX_VIEWS);
  if (count == 0) return MPSNERF_OK;
  MPS_REQUIRE(tokens && xc && weights && act_pid && raw && workspace && ld >= MPSNERF_TOKEN_DIM);
  MPS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  Fp32Ws w;
  carve(w, static_cast<char*>(workspace), count, n_views);
  const int V = n_views;
  const int64_t M3 = count * V;
  const float* const* L = weights;
  copy_rows_kernel<<<blocks_for(M3 * 155, 256), 256, 0, st>>>(tokens, ld, w.X, 155, M3, 155);
  for (int l = 0; l < 2; ++l) {
    const float* const* P = L + 11 * l;   // ln1_w ln1_b qkv_w out_w out_b ln2_w ln2_b ff1_w ff1_b ff2_w ff2_b
    ln_rows_kernel<<<blocks_for(M3 * 32, 256), 256, 0, st>>>(w.X, 155, P[0], P[1], w.Y, 155, M3, 155);
    launch_linear<0>(w.Y, 155, P[2], 155, nullptr, nullptr, 0, w.QKV, 768, M3, 768, st);
    attention_kernel<<<blocks_for(count * 32, 256), 256, 0, st>>>(w.QKV, w.O, count, V);
    launch_linear<0>(w.O, 256, P[3], 256, P[4], w.X, 155, w.X, 155, M3, 155, st);       // x += out(o)
    ln_rows_kernel<<<blocks_for(M3 * 32, 256), 256, 0, st>>>(w.X, 155, P[5], P[6], w.Y, 155, M3, 155);
    launch_linear<2>(w.Y, 155, P[7], 155, P[8], nullptr, 0, w.Hff, 128, M3, 128, st);
    launch_linear<0>(w.Hff, 128, P[9], 128, P[10], w.X, 155, w.X, 155, M3, 155, st);    // x += ff(y)
  }
  const float* const* Q = L + 22;         // pts_linears.{0..7}.{weight,bias}
  mlp_inputs_kernel<<<blocks_for(count * 349, 256), 256, 0, st>>>(xc, w.X, V, w.Hcat, w.F, count);
  launch_linear<1>(w.Hcat, 450, Q[0], 194, Q[1], nullptr, 0, w.H1, 256, count, 256, st);
  launch_linear<1>(w.H1, 256, Q[2], 256, Q[3], nullptr, 0, w.H2, 256, count, 256, st);
  launch_linear<1>(w.H2, 256, Q[4], 256, Q[5], nullptr, 0, w.H1, 256, count, 256, st);
  launch_linear<1>(w.H1, 256, Q[6], 256, Q[7], nullptr, 0, w.H2, 256, count, 256, st);
  launch_linear<1>(w.H2, 256, Q[8], 256, Q[9], nullptr, 0, w.Hcat + 194, 450, count, 256, st);   // skip: [x | h]
  launch_linear<1>(w.Hcat, 450, Q[10], 450, Q[11], nullptr, 0, w.H1, 256, count, 256, st);
  launch_linear<1>(w.H1, 256, Q[12], 256, Q[13], nullptr, 0, w.H2, 256, count, 256, st);
  launch_linear<1>(w.H2, 256, Q[14], 256, Q[15], nullptr, 0, w.H1, 256, count, 256, st);
  const float* const* T = L + 38;         // alpha_w alpha_b feature_w feature_b views_w views_b rgb_w rgb_b
  launch_linear<0>(w.H1, 256, T[0], 256, T[1], nullptr, 0, w.out4 + 3, 4, count, 1, st);
  launch_linear<0>(w.H1, 256, T[2], 256, T[3], nullptr, 0, w.F, 411, count, 256, st);
  launch_linear<1>(w.F, 411, T[4], 411, T[5], nullptr, 0, w.G, 128, count, 128, st);
  launch_linear<0>(w.G, 128, T[6], 128, T[7], nullptr, 0, w.out4, 4, count, 3, st);
  scatter_raw_kernel<<<blocks_for(count, 256), 256, 0, st>>>(w.out4, act_pid, first, count, raw);
  MPS_LAUNCH_CHECK();
  return MPSNERF_OK;
}

// =====================================================================================================
// Training: the same dense stage with every intermediate kept, and its backward (BASELINE config 4:
// fwd + bwd of the render path, run_nerf_batch.py:544-570).  fp32 on CUDA cores -- a training batch is
// 1024 rays per GPU, ~5 K active points, 50 GFLOP for forward + backward: launch-bound, not math-bound.
// Gradients w.r.t. every live parameter of the transformer and the MLP (accumulated into caller-owned
// buffers, like torch's .grad) and w.r.t. the tokens (-> K4 backward -> encoder trunk).
// =====================================================================================================
namespace mps {

// C[i, j] (+)= sum_r A(i, r) * B(j, r) with arbitrary strides: A(i, r) = A[i * sa_i + r * sa_r], B(j, r) likewise.
//   data gradient   dX[m, k] = sum_n dY[m, n] W[n, k]        i = m, j = k, r = n   (B: sb_j = 1, sb_r = K)
//   weight gradient dW[n, k] = sum_m dY[m, n] X[m, k]        i = n, j = k, r = m   (A: sa_i = 1, sa_r = ldy)
// The reduction range is split over blockIdx.z (weight gradients reduce over the points: few output tiles, long
// reduction) and the partial tiles are combined with atomicAdd when `atomic` is set; otherwise C is overwritten
// (accumulate = 0) or added to (accumulate = 1).
constexpr int GM = 64, GN = 64, GK = 16, kGemmThreads = 256;

__global__ void __launch_bounds__(kGemmThreads)
gemm_strided_kernel(const float* __restrict__ A, int64_t sa_i, int64_t sa_r, const float* __restrict__ B, int64_t sb_j,
                    int64_t sb_r, float* C, int64_t ldc, int64_t I, int J, int64_t R, int64_t r_per_split,
                    int accumulate, int atomic) {
  __shared__ float As[GK][GM + 4];
  __shared__ float Bs[GK][GN + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;      // 16 x 16 threads, 4 x 4 outputs each
  const int64_t i0 = (int64_t)blockIdx.x * GM;
  const int j0 = blockIdx.y * GN;
  const int64_t r_begin = (int64_t)blockIdx.z * r_per_split, r_end = min(R, r_begin + r_per_split);
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const bool a_r_fast = (sa_r == 1), b_r_fast = (sb_r == 1);
  for (int64_t r0 = r_begin; r0 < r_end; r0 += GK) {
    for (int e = tid; e < GM * GK; e += kGemmThreads) {
      const int rr = a_r_fast ? e % GK : e / GM, ii = a_r_fast ? e / GK : e % GM;
      const int64_t i = i0 + ii, r = r0 + rr;
      As[rr][ii] = (i < I && r < r_end) ? A[i * sa_i + r * sa_r] : 0.f;
    }
    for (int e = tid; e < GN * GK; e += kGemmThreads) {
      const int rr = b_r_fast ? e % GK : e / GN, jj = b_r_fast ? e / GK : e % GN;
      const int64_t r = r0 + rr;
      const int j = j0 + jj;
      Bs[rr][jj] = (j < J && r < r_end) ? B[(int64_t)j * sb_j + r * sb_r] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < GK; ++rr) {
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = As[rr][ty * 4 + u]; b[u] = Bs[rr][tx * 4 + u]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int64_t i = i0 + ty * 4 + u;
    if (i >= I) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int j = j0 + tx * 4 + v;
      if (j >= J) continue;
      float* c = C + i * ldc + j;
      if (atomic) atomicAdd(c, acc[u][v]);
      else *c = accumulate ? *c + acc[u][v] : acc[u][v];
    }
  }
}

// dX (M, K) (+)= dY (M, N) . W (N, K)
static void bwd_data(const float* dY, int ldy, const float* W, int K, float* dX, int ldx, int64_t M, int N, int accumulate,
                     cudaStream_t st) {
  dim3 grid((unsigned)((M + GM - 1) / GM), (unsigned)((K + GN - 1) / GN), 1);
  gemm_strided_kernel<<<grid, kGemmThreads, 0, st>>>(dY, ldy, 1, W, 1, K, dX, ldx, M, K, N, N, accumulate, 0);
}
// dW (N, K) += dY^T (N, M) . X (M, K), the reduction over the M points split into slices of 1024 rows
static void bwd_weight(const float* dY, int ldy, const float* X, int ldx, float* dW, int K, int64_t M, int N, cudaStream_t st) {
  const int64_t per = 1024;
  dim3 grid((unsigned)((N + GM - 1) / GM), (unsigned)((K + GN - 1) / GN), (unsigned)((M + per - 1) / per));
  gemm_strided_kernel<<<grid, kGemmThreads, 0, st>>>(dY, 1, ldy, X, 1, ldx, dW, K, N, K, M, per, 1, 1);
}

// db[n] += sum_m dY[m, n]
__global__ void colsum_kernel(const float* __restrict__ dY, int ldy, float* __restrict__ db, int64_t M, int N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int64_t per = (M + gridDim.y - 1) / gridDim.y, m0 = blockIdx.y * per, m1 = min(M, m0 + per);
  float s = 0.f;
  for (int64_t m = m0; m < m1; ++m) s += dY[m * ldy + n];
  atomicAdd(&db[n], s);
}
static void bwd_bias(const float* dY, int ldy, float* db, int64_t M, int N, cudaStream_t st) {
  dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 255) / 256 < 64 ? (M + 255) / 256 : 64));
  colsum_kernel<<<grid, 64, 0, st>>>(dY, ldy, db, M, N);
}

// in place: dY *= (Y > 0)   (Y = relu output), row-strided
__global__ void relu_bwd_kernel(float* __restrict__ dY, int ldd, const float* __restrict__ Y, int ldy, int64_t rows, int cols) {
  const int64_t total = rows * cols;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / cols;
    const int c = (int)(t % cols);
    if (!(Y[r * ldy + c] > 0.f)) dY[r * ldd + c] = 0.f;
  }
}
// H = gelu(Z) (erf form, nn.GELU default)
__global__ void gelu_fwd_kernel(const float* __restrict__ Z, float* __restrict__ H, int64_t n) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const float z = Z[t];
    H[t] = 0.5f * z * (1.f + erff(z * 0.70710678118654752440f));
  }
}
// in place: dH *= gelu'(Z) = Phi(z) + z phi(z)
__global__ void gelu_bwd_kernel(float* __restrict__ dH, const float* __restrict__ Z, int64_t n) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const float z = Z[t];
    const float cdf = 0.5f * (1.f + erff(z * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * z * z);
    dH[t] *= cdf + z * pdf;
  }
}

// LayerNorm backward, one warp per row.  dX[r, :] += rstd * (g dy - mean(g dy) - xhat mean(g dy xhat)); the parameter
// gradients dg += sum_r dy xhat, db += sum_r dy are first summed per block in shared memory.
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ X, const float* __restrict__ g, float* dX,
              float* __restrict__ dg, float* __restrict__ db, int64_t rows, int D) {
  __shared__ float s_dg[160], s_db[160];
  for (int c = threadIdx.x; c < D; c += blockDim.x) { s_dg[c] = 0.f; s_db[c] = 0.f; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = w0; r < rows; r += nw) {
    const float* x = X + r * D;
    const float* dy = dY + r * D;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += x[c];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)D;
    float v = 0.f;
    for (int c = lane; c < D; c += 32) { const float d = x[c] - mean; v += d * d; }
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = rsqrtf(v / (float)D + 1e-5f);
    float a = 0.f, b = 0.f;                       // sum g dy, sum g dy xhat
    for (int c = lane; c < D; c += 32) {
      const float xh = (x[c] - mean) * rstd, gd = g[c] * dy[c];
      a += gd; b += gd * xh;
      atomicAdd(&s_dg[c], dy[c] * xh);
      atomicAdd(&s_db[c], dy[c]);
    }
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    a /= (float)D; b /= (float)D;
    for (int c = lane; c < D; c += 32) {
      const float xh = (x[c] - mean) * rstd;
      dX[r * D + c] += rstd * (g[c] * dy[c] - a - xh * b);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) { atomicAdd(&dg[c], s_dg[c]); atomicAdd(&db[c], s_db[c]); }
}

// Attention backward (lib/transformer.py:59-71): one thread per (point, head, d-slice of 16 of the 64 head dims would
// need cross-thread sums; the point counts of a training batch are small) -- one thread per (point, head).
__global__ void attention_bwd_kernel(const float* __restrict__ QKV, const float* __restrict__ dO, float* __restrict__ dQKV,
                                     int64_t count, int V) {
  const int64_t total = count * 4;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int h = (int)(t % 4);
    const int64_t p = t / 4;
    float P[MPSNERF_MAX_VIEWS][MPSNERF_MAX_VIEWS], dS[MPSNERF_MAX_VIEWS][MPSNERF_MAX_VIEWS];
    for (int i = 0; i < V; ++i) {
      const float* q = QKV + (p * V + i) * 768 + h * 64;
      float mx = -1e30f;
      for (int j = 0; j < V; ++j) {
        const float* k = QKV + (p * V + j) * 768 + 256 + h * 64;
        float d = 0.f;
        for (int c = 0; c < 64; ++c) d = fmaf(q[c], k[c], d);
        P[i][j] = d * 0.125f;
        mx = fmaxf(mx, P[i][j]);
      }
      float den = 0.f;
      for (int j = 0; j < V; ++j) { P[i][j] = expf(P[i][j] - mx); den += P[i][j]; }
      for (int j = 0; j < V; ++j) P[i][j] /= den;
      // dP[i][j] = dO_i . v_j ; dS = P (dP - sum_j P dP)
      const float* go = dO + (p * V + i) * 256 + h * 64;
      float dP[MPSNERF_MAX_VIEWS], dot = 0.f;
      for (int j = 0; j < V; ++j) {
        const float* v = QKV + (p * V + j) * 768 + 512 + h * 64;
        float d = 0.f;
        for (int c = 0; c < 64; ++c) d = fmaf(go[c], v[c], d);
        dP[j] = d;
        dot = fmaf(P[i][j], d, dot);
      }
      for (int j = 0; j < V; ++j) dS[i][j] = P[i][j] * (dP[j] - dot) * 0.125f;
    }
    for (int i = 0; i < V; ++i) {
      float* dq = dQKV + (p * V + i) * 768 + h * 64;
      float* dk = dq + 256;
      float* dv = dq + 512;
      for (int c = 0; c < 64; ++c) {
        float aq = 0.f, ak = 0.f, av = 0.f;
        for (int j = 0; j < V; ++j) {
          aq = fmaf(dS[i][j], QKV[(p * V + j) * 768 + 256 + h * 64 + c], aq);      // dq_i = sum_j dS_ij k_j
          ak = fmaf(dS[j][i], QKV[(p * V + j) * 768 + h * 64 + c], ak);            // dk_i = sum_j dS_ji q_j
          av = fmaf(P[j][i], dO[(p * V + j) * 256 + h * 64 + c], av);              // dv_i = sum_j P_ji dO_j
        }
        dq[c] = aq; dk[c] = ak; dv[c] = av;
      }
    }
  }
}

// dX2 rows of token 0 / token 1 <- the MLP's input gradients; every other token row is zero
__global__ void scatter_token_grads_kernel(const float* __restrict__ dHcat, const float* __restrict__ dF, float* __restrict__ dX,
                                           int64_t count, int V) {
  const int64_t total = count * V * 155;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % 155);
    const int tok = (int)((t / 155) % V);
    const int64_t p = t / (155 * (int64_t)V);
    dX[t] = tok == 0 ? dHcat[p * 450 + 39 + c] : (tok == 1 ? dF[p * 411 + 256 + c] : 0.f);
  }
}
__global__ void add_rows_kernel(float* __restrict__ dst, int ldd, const float* __restrict__ src, int lds, int64_t rows, int cols) {
  const int64_t total = rows * cols;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
    dst[(t / cols) * ldd + (t % cols)] += src[(t / cols) * lds + (t % cols)];
}
// dH7 (count, 256) += d_alpha (count) x w_alpha (256)
__global__ void alpha_bwd_kernel(float* __restrict__ dH, const float* __restrict__ d_out4, const float* __restrict__ w_alpha, int64_t count) {
  const int64_t total = count * 256;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
    dH[t] = fmaf(d_out4[(t / 256) * 4 + 3], w_alpha[t % 256], dH[t]);
}

struct TrainWs {
  // forward, kept for the backward
  float *X[3], *Xmid[2], *Y1[2], *QKV[2], *O[2], *Y2[2], *Z[2], *Hff[2];   // transformer (M3 rows); X[l] = input of layer l
  float *Hcat, *H[8], *F, *G;                                         // MLP (count rows); H[4] lives in Hcat[:, 194:]
  // backward scratch
  float *dXa, *dXb, *dBig, *dO, *dHf, *dA, *dB, *dF, *dG;
};

static size_t carve_train(TrainWs& w, char* base, int64_t count, int V) {
  size_t off = 0;
  auto take = [&](size_t nfloats) {
    float* p = reinterpret_cast<float*>(base + off);
    off += ((nfloats * sizeof(float) + 255) / 256) * 256;
    return p;
  };
  const size_t M3 = (size_t)count * V, C = (size_t)count;
  for (int l = 0; l < 3; ++l) w.X[l] = take(M3 * 155);
  for (int l = 0; l < 2; ++l) {
    w.Xmid[l] = take(M3 * 155); w.Y1[l] = take(M3 * 155); w.QKV[l] = take(M3 * 768); w.O[l] = take(M3 * 256);
    w.Y2[l] = take(M3 * 155); w.Z[l] = take(M3 * 128); w.Hff[l] = take(M3 * 128);
  }
  w.Hcat = take(C * 450);
  for (int i = 0; i < 8; ++i) w.H[i] = (i == 4) ? nullptr : take(C * 256);
  w.F = take(C * 411); w.G = take(C * 128);
  w.dXa = take(M3 * 155); w.dXb = take(M3 * 155); w.dBig = take(M3 * 768); w.dO = take(M3 * 256); w.dHf = take(M3 * 128);
  w.dA = take(C * 450); w.dB = take(C * 450); w.dF = take(C * 411); w.dG = take(C * 128);
  return off;
}

}  // namespace mps

extern "C" size_t mpsnerf_dense_train_workspace(int64_t count, int n_views) {
  mps::TrainWs w;
  return mps::carve_train(w, nullptr, count > 0 ? count : 0, n_views) + 256;
}

// Forward of the dense stage with every intermediate kept in `workspace` for mpsnerf_dense_train_bwd.
// out4 (count, 4) = [rgb, alpha] per active point (the caller scatters / composites).
extern "C" int mpsnerf_dense_train_fwd(const float* tokens, int32_t ld, const float* xc, int64_t count, int n_views,
                                       const float* const* weights, float* out4, void* workspace, void* stream) {
  using namespace mps;
  MPS_REQUIRE(count >= 0 && n_views >= 2 && n_views <= MPSNERF_MAX_VIEWS);
  if (count == 0) return MPSNERF_OK;
  MPS_REQUIRE(tokens && xc && weights && out4 && workspace && ld >= MPSNERF_TOKEN_DIM);
  MPS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  TrainWs w;
  carve_train(w, static_cast<char*>(workspace), count, n_views);
  const int V = n_views;
  const int64_t M3 = count * V;
  const float* const* L = weights;
  copy_rows_kernel<<<blocks_for(M3 * 155, 256), 256, 0, st>>>(tokens, ld, w.X[0], 155, M3, 155);
  for (int l = 0; l < 2; ++l) {
    const float* const* P = L + 11 * l;   // ln1_w ln1_b qkv_w out_w out_b ln2_w ln2_b ff1_w ff1_b ff2_w ff2_b
    ln_rows_kernel<<<blocks_for(M3 * 32, 256), 256, 0, st>>>(w.X[l], 155, P[0], P[1], w.Y1[l], 155, M3, 155);
    launch_linear<0>(w.Y1[l], 155, P[2], 155, nullptr, nullptr, 0, w.QKV[l], 768, M3, 768, st);
    attention_kernel<<<blocks_for(count * 32, 256), 256, 0, st>>>(w.QKV[l], w.O[l], count, V);
    launch_linear<0>(w.O[l], 256, P[3], 256, P[4], w.X[l], 155, w.Xmid[l], 155, M3, 155, st);
    ln_rows_kernel<<<blocks_for(M3 * 32, 256), 256, 0, st>>>(w.Xmid[l], 155, P[5], P[6], w.Y2[l], 155, M3, 155);
    launch_linear<0>(w.Y2[l], 155, P[7], 155, P[8], nullptr, 0, w.Z[l], 128, M3, 128, st);
    gelu_fwd_kernel<<<blocks_for(M3 * 128, 256), 256, 0, st>>>(w.Z[l], w.Hff[l], M3 * 128);
    launch_linear<0>(w.Hff[l], 128, P[9], 128, P[10], w.Xmid[l], 155, w.X[l + 1], 155, M3, 155, st);
  }
  const float* const* Q = L + 22;         // pts_linears.{0..7}.{weight,bias}
  mlp_inputs_kernel<<<blocks_for(count * 349, 256), 256, 0, st>>>(xc, w.X[2], V, w.Hcat, w.F, count);
  float* H4 = w.Hcat + 194;               // ld 450
  launch_linear<1>(w.Hcat, 450, Q[0], 194, Q[1], nullptr, 0, w.H[0], 256, count, 256, st);
  launch_linear<1>(w.H[0], 256, Q[2], 256, Q[3], nullptr, 0, w.H[1], 256, count, 256, st);
  launch_linear<1>(w.H[1], 256, Q[4], 256, Q[5], nullptr, 0, w.H[2], 256, count, 256, st);
  launch_linear<1>(w.H[2], 256, Q[6], 256, Q[7], nullptr, 0, w.H[3], 256, count, 256, st);
  launch_linear<1>(w.H[3], 256, Q[8], 256, Q[9], nullptr, 0, H4, 450, count, 256, st);            // skip: [x | h]
  launch_linear<1>(w.Hcat, 450, Q[10], 450, Q[11], nullptr, 0, w.H[5], 256, count, 256, st);
  launch_linear<1>(w.H[5], 256, Q[12], 256, Q[13], nullptr, 0, w.H[6], 256, count, 256, st);
  launch_linear<1>(w.H[6], 256, Q[14], 256, Q[15], nullptr, 0, w.H[7], 256, count, 256, st);
  const float* const* T = L + 38;         // alpha_w alpha_b feature_w feature_b views_w views_b rgb_w rgb_b
  launch_linear<0>(w.H[7], 256, T[0], 256, T[1], nullptr, 0, out4 + 3, 4, count, 1, st);
  launch_linear<0>(w.H[7], 256, T[2], 256, T[3], nullptr, 0, w.F, 411, count, 256, st);
  launch_linear<1>(w.F, 411, T[4], 411, T[5], nullptr, 0, w.G, 128, count, 128, st);
  launch_linear<0>(w.G, 128, T[6], 128, T[7], nullptr, 0, out4, 4, count, 3, st);
  MPS_LAUNCH_CHECK();
  return MPSNERF_OK;
}

// Backward of mpsnerf_dense_train_fwd.  d_out4 (count, 4): gradient w.r.t. [rgb, alpha].  grads: 46 device pointers in
// the order of `weights` (DENSE_FP32_ORDER), each ACCUMULATED into (zero them for a fresh gradient).  d_tokens
// (count, V, 155): gradient w.r.t. the tokens (overwritten).  Positional code and canonical points carry no gradient
// (no parameter upstream of them under the shipped configs: skinning_field = correction_field = 0).
extern "C" int mpsnerf_dense_train_bwd(const float* d_out4, int64_t count, int n_views, const float* const* weights,
                                       float* const* grads, float* d_tokens, void* workspace, void* stream) {
  using namespace mps;
  MPS_REQUIRE(count >= 0 && n_views >= 2 && n_views <= MPSNERF_MAX_VIEWS);
  if (count == 0) return MPSNERF_OK;
  MPS_REQUIRE(d_out4 && weights && grads && d_tokens && workspace);
  cudaStream_t st = (cudaStream_t)stream;
  TrainWs w;
  carve_train(w, static_cast<char*>(workspace), count, n_views);
  const int V = n_views;
  const int64_t M3 = count * V, C = count;
  const float* const* T = weights + 38;
  float* const* gT = grads + 38;
  const float* const* Q = weights + 22;
  float* const* gQ = grads + 22;
  auto relu_bwd = [&](float* d, int ldd, const float* y, int ldy, int64_t rows, int cols) {
    relu_bwd_kernel<<<blocks_for(rows * cols, 256), 256, 0, st>>>(d, ldd, y, ldy, rows, cols);
  };
  // ---- colour head: rgb_linear <- G = relu(views_linear(F)), F = [feature_linear(H7) | tok1]
  bwd_weight(d_out4, 4, w.G, 128, gT[6], 128, C, 3, st);
  bwd_bias(d_out4, 4, gT[7], C, 3, st);
  bwd_data(d_out4, 4, T[6], 128, w.dG, 128, C, 3, 0, st);
  relu_bwd(w.dG, 128, w.G, 128, C, 128);
  bwd_weight(w.dG, 128, w.F, 411, gT[4], 411, C, 128, st);
  bwd_bias(w.dG, 128, gT[5], C, 128, st);
  bwd_data(w.dG, 128, T[4], 411, w.dF, 411, C, 128, 0, st);           // dF[:, :256] = d feature, dF[:, 256:] = d tok1
  bwd_weight(w.dF, 411, w.H[7], 256, gT[2], 256, C, 256, st);
  bwd_bias(w.dF, 411, gT[3], C, 256, st);
  float* dH = w.dA;                                                    // (count, 256) views of the (count, 450) scratch
  bwd_data(w.dF, 411, T[2], 256, dH, 256, C, 256, 0, st);
  // ---- density head: alpha_linear on H7
  bwd_weight(d_out4 + 3, 4, w.H[7], 256, gT[0], 256, C, 1, st);
  bwd_bias(d_out4 + 3, 4, gT[1], C, 1, st);
  alpha_bwd_kernel<<<blocks_for(C * 256, 256), 256, 0, st>>>(dH, d_out4, T[0], C);
  // ---- pts_linears 7, 6, 5
  float* dcur = dH;
  float* dnext = w.dB;
  for (int i = 7; i >= 6; --i) {
    relu_bwd(dcur, 256, w.H[i], 256, C, 256);
    bwd_weight(dcur, 256, w.H[i - 1], 256, gQ[2 * i], 256, C, 256, st);
    bwd_bias(dcur, 256, gQ[2 * i + 1], C, 256, st);
    bwd_data(dcur, 256, Q[2 * i], 256, dnext, 256, C, 256, 0, st);
    float* t = dcur; dcur = dnext; dnext = t;
  }
  relu_bwd(dcur, 256, w.H[5], 256, C, 256);                            // layer 5: input Hcat = [x (194) | H4 (256)]
  bwd_weight(dcur, 256, w.Hcat, 450, gQ[10], 450, C, 256, st);
  bwd_bias(dcur, 256, gQ[11], C, 256, st);
  float* dHcat = dnext;                                                // dA / dB are (count, 450)
  bwd_data(dcur, 256, Q[10], 450, dHcat, 450, C, 256, 0, st);          // dHcat[:, :194] = dx (skip), [:, 194:] = dH4
  // ---- pts_linears 4 .. 1 (H4 lives in Hcat[:, 194:], its gradient in dHcat[:, 194:])
  float* dx_skip = dHcat;                                              // keep: columns [0, 194) are added to dx below
  float* d4 = dHcat + 194;                                             // ld 450
  relu_bwd(d4, 450, w.Hcat + 194, 450, C, 256);
  bwd_weight(d4, 450, w.H[3], 256, gQ[8], 256, C, 256, st);
  bwd_bias(d4, 450, gQ[9], C, 256, st);
  float* da = dcur;                                                    // free again (its contents were consumed above)
  bwd_data(d4, 450, Q[8], 256, da, 256, C, 256, 0, st);
  float* dbuf = w.dBig;                                                // idle until the transformer backward: (count, 256) here
  for (int i = 3; i >= 1; --i) {
    relu_bwd(da, 256, w.H[i], 256, C, 256);
    bwd_weight(da, 256, w.H[i - 1], 256, gQ[2 * i], 256, C, 256, st);
    bwd_bias(da, 256, gQ[2 * i + 1], C, 256, st);
    bwd_data(da, 256, Q[2 * i], 256, dbuf, 256, C, 256, 0, st);
    float* t = da; da = dbuf; dbuf = t;
  }
  relu_bwd(da, 256, w.H[0], 256, C, 256);                              // layer 0: input x = Hcat[:, :194]
  bwd_weight(da, 256, w.Hcat, 450, gQ[0], 194, C, 256, st);
  bwd_bias(da, 256, gQ[1], C, 256, st);
  bwd_data(da, 256, Q[0], 194, dx_skip, 450, C, 256, 1, st);           // dx = skip part + layer-0 part (accumulate)
  // ---- gradient w.r.t. the transformer output: token 0 <- dx[:, 39:194], token 1 <- dF[:, 256:411]
  float* dX = w.dXa;
  scatter_token_grads_kernel<<<blocks_for(M3 * 155, 256), 256, 0, st>>>(dx_skip, w.dF, dX, C, V);
  // ---- transformer layers 1, 0
  for (int l = 1; l >= 0; --l) {
    const float* const* P = weights + 11 * l;
    float* const* gP = grads + 11 * l;
    // feed-forward: x_out = x_mid + W2 gelu(W1 LN2(x_mid) + b1) + b2
    bwd_weight(dX, 155, w.Hff[l], 128, gP[9], 128, M3, 155, st);
    bwd_bias(dX, 155, gP[10], M3, 155, st);
    bwd_data(dX, 155, P[9], 128, w.dHf, 128, M3, 155, 0, st);
    gelu_bwd_kernel<<<blocks_for(M3 * 128, 256), 256, 0, st>>>(w.dHf, w.Z[l], M3 * 128);
    bwd_weight(w.dHf, 128, w.Y2[l], 155, gP[7], 155, M3, 128, st);
    bwd_bias(w.dHf, 128, gP[8], M3, 128, st);
    bwd_data(w.dHf, 128, P[7], 155, w.dXb, 155, M3, 128, 0, st);       // dY2
    ln_bwd_kernel<<<blocks_for(M3 * 32, 256), 256, 0, st>>>(w.dXb, w.Xmid[l], P[5], dX, gP[5], gP[6], M3, 155);   // dX = dX_mid
    // attention: x_mid = x_in + Wo attn(Wqkv LN1(x_in)) + bo
    bwd_weight(dX, 155, w.O[l], 256, gP[3], 256, M3, 155, st);
    bwd_bias(dX, 155, gP[4], M3, 155, st);
    bwd_data(dX, 155, P[3], 256, w.dO, 256, M3, 155, 0, st);
    attention_bwd_kernel<<<blocks_for(C * 4, 64), 64, 0, st>>>(w.QKV[l], w.dO, w.dBig, C, V);
    bwd_weight(w.dBig, 768, w.Y1[l], 155, gP[2], 155, M3, 768, st);
    bwd_data(w.dBig, 768, P[2], 155, w.dXb, 155, M3, 768, 0, st);      // dY1
    ln_bwd_kernel<<<blocks_for(M3 * 32, 256), 256, 0, st>>>(w.dXb, w.X[l], P[0], dX, gP[0], gP[1], M3, 155);      // dX = dX_in
  }
  copy_rows_kernel<<<blocks_for(M3 * 155, 256), 256, 0, st>>>(dX, 155, d_tokens, 155, M3, 155);
  MPS_LAUNCH_CHECK();
  return MPSNERF_OK;
}
