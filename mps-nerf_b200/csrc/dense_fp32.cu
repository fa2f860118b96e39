// K5 (fp32 option): cross-view transformer + canonical NeRF MLP on CUDA cores.
//
// Restates Transformer.forward (lib/transformer.py:13-86) and the MLP of
// SKinningBatch.forward (lib/skinnning_batch.py:438-473) layer by layer with fp32 FMA
// accumulation.  This is the high-precision mode (rgb within 1e-4 of the reference); the
// production path is the bf16 tcgen05 kernel in dense_tc.cu.
#include "common.cuh"

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

template <int ACT>
static int launch_linear(const float* X, int ldx, const float* W, int K, const float* b, const float* Res, int ldr,
                         float* Y, int ldy, int64_t M, int N, cudaStream_t st) {
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
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
// ('b n (h d) -> b h n d', lib/transformer.py:62).  One thread per (point, head, query token).
__global__ void attention_kernel(const float* __restrict__ QKV, float* __restrict__ O, int64_t count, int V) {
  const int64_t total = count * 4 * V;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(t % V);
    const int h = (int)((t / V) % 4);
    const int64_t p = t / (4 * V);
    const float* q = QKV + (p * V + i) * 768 + h * 64;
    float dots[MPSNERF_MAX_VIEWS];
    float mx = -1e30f;
    for (int j = 0; j < V; ++j) {
      const float* k = QKV + (p * V + j) * 768 + 256 + h * 64;
      float d = 0.f;
      for (int c = 0; c < 64; ++c) d = fmaf(q[c], k[c], d);
      dots[j] = d * 0.125f;                       // dim_head ** -0.5
      mx = fmaxf(mx, dots[j]);
    }
    float den = 0.f;
    for (int j = 0; j < V; ++j) { dots[j] = expf(dots[j] - mx); den += dots[j]; }
    float* o = O + (p * V + i) * 256 + h * 64;
    for (int c = 0; c < 64; ++c) {
      float a = 0.f;
      for (int j = 0; j < V; ++j) a = fmaf(dots[j] / den, QKV[(p * V + j) * 768 + 512 + h * 64 + c], a);
      o[c] = a;
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
    attention_kernel<<<blocks_for(count * 4 * V, 128), 128, 0, st>>>(w.QKV, w.O, count, V);
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
