// K5 (bf16 production path): fused tcgen05 / TMEM kernels for the cross-view transformer ("T")
// and the canonical NeRF MLP ("M").  Restates lib/transformer.py:13-86 and
// lib/skinnning_batch.py:438-473 with bf16 operands and fp32 accumulation.
//
// Common structure (one persistent CTA per SM, 192 threads):
//   warps 0-3  epilogue: thread r owns TMEM lane r = row r of the 128-row tile.  They turn
//              accumulators into the next A operand (bias / ReLU / GELU / LayerNorm / attention)
//              written straight into the SWIZZLE_128B K-major smem layout the MMA reads;
//   warp 4     weight producer: one thread streams the pre-swizzled weight chunks with bulk
//              async copies (TMA engine) into an mbarrier ring, in consumption order;
//   warp 5     MMA issuer: one thread issues tcgen05.mma (M = 128) and commits to mbarriers.
// Activations never leave the SM between layers; per tile only the inputs are read and the
// outputs written.  MMA <-> epilogue hand-over is a pair of mbarriers (a_bar: "A operand
// ready", 128 arrivals; d_bar: "accumulator ready", tcgen05.commit).
#include "common.cuh"
#include "umma.cuh"

namespace mps {
using namespace umma;

constexpr int kTcThreads = 192;
constexpr int kEpiThreads = 128;

// ---- blob layout (must match mps-nerf_b200/pack.py)
constexpr uint32_t kQkvChunk = 192 * 128, kWoChunk = 160 * 128, kW1Chunk = 128 * 128, kW2Chunk = 160 * 128;
constexpr uint32_t kTLayerBytes = 12 * kQkvChunk + 4 * kWoChunk + 3 * kW1Chunk + 2 * kW2Chunk;
constexpr uint32_t kTBytes = 2 * kTLayerBytes;
constexpr uint32_t kMBytes = 40 * 32768 + 7 * 16384;
constexpr uint32_t kFloatOff = kTBytes + kMBytes;
constexpr int kTLayerFloats = 1088, kTFloats = 2 * kTLayerFloats + 160, kMFloats = 8 * 256 + 256 + 256 + 128 + 384 + 4;
constexpr size_t kBlobBytes = (size_t)kFloatOff + 4 * (size_t)(kTFloats + kMFloats);
static_assert(kTLayerBytes == 466944 && kMBytes == 1425408, "blob layout drifted from pack.py");

constexpr int kTokLd = 160;   // bf16 row stride of tok0 / tok1 handed from T to M

// ------------------------------------------------------------------------------------------
// small shared pieces
// ------------------------------------------------------------------------------------------
struct Pipe {            // barriers of one CTA (in dynamic smem)
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t a_bar;
  uint64_t d_bar;
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ float gelu_erf(float x) {
  // 0.5 x (1 + erf(x / sqrt 2)), erf by Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7)
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.0f - p * t * __expf(-z * z);
  return 0.5f * x * (1.0f + copysignf(e, x));
}

__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 t = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// ------------------------------------------------------------------------------------------
// T: cross-view transformer.  Tile = ppt = 128 / V points, row r = (point r / V, token r % V).
// TMEM columns: X [0,160) fp32 residual stream (the out-proj and FF2 GEMMs accumulate into it,
// which is the residual add; their biases are deferred, see pack.py), R [160,352) scratch
// accumulator (q|k|v of one head, or the FF hidden layer).
// ------------------------------------------------------------------------------------------
constexpr uint32_t kT_YA = 0;                       // LN output, A operand, 3 chunks
constexpr uint32_t kT_OA = kT_YA + 3 * 16384;       // attention output (4 chunks = 4 heads); FF hidden aliases chunks 0-1
constexpr uint32_t kT_KX = kT_OA + 4 * 16384;       // k of the current head, fp16 [128][64], unit-swizzled
constexpr uint32_t kT_VX = kT_KX + 16384;
constexpr uint32_t kT_RING = kT_VX + 16384;
constexpr int kT_Slots = 3;
constexpr uint32_t kT_SlotBytes = kQkvChunk;
constexpr uint32_t kT_FP = kT_RING + kT_Slots * kT_SlotBytes;
constexpr uint32_t kT_PIPE = kT_FP + kTFloats * 4;
constexpr uint32_t kT_Smem = kT_PIPE + sizeof(Pipe);
constexpr uint32_t kT_ColX = 0, kT_ColR = 160;
static_assert(kT_Smem <= 232448 - 1024, "T kernel shared memory over budget");

struct TArgs {
  const float* tokens;   // (count, V, ld) fp32
  int ld;
  int64_t count;
  int V;
  const uint8_t* blob;
  __nv_bfloat16* tok0;   // (count, 160)
  __nv_bfloat16* tok1;
};

// LayerNorm over the 155 real columns of x (+ optional pending bias), result -> bf16 A operand.
// gamma/beta are zero in the 5 pad columns, so the pad of the operand is exactly zero.
__device__ __forceinline__ void ln_to_operand(float (&x)[160], const float* __restrict__ pend,
                                              const float* __restrict__ g, const float* __restrict__ b,
                                              uint8_t* YA, int r) {
  float s = 0.f;
  if (pend) {
    const float4* p4 = reinterpret_cast<const float4*>(pend);
#pragma unroll
    for (int c = 0; c < 40; ++c) {
      const float4 t = p4[c];
      x[4 * c] += t.x; x[4 * c + 1] += t.y; x[4 * c + 2] += t.z; x[4 * c + 3] += t.w;
    }
  }
#pragma unroll
  for (int c = 0; c < 155; ++c) s += x[c];
  const float mean = s * (1.0f / 155.0f);
  float v = 0.f;
#pragma unroll
  for (int c = 0; c < 155; ++c) { const float d = x[c] - mean; v = fmaf(d, d, v); }
  const float rstd = rsqrtf(v * (1.0f / 155.0f) + 1e-5f);
#pragma unroll
  for (int c0 = 0; c0 < 160; c0 += 8) {
    float y[8];
    const float4 g0 = *reinterpret_cast<const float4*>(g + c0), g1 = *reinterpret_cast<const float4*>(g + c0 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(b + c0), b1 = *reinterpret_cast<const float4*>(b + c0 + 4);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = fmaf((x[c0 + i] - mean) * rstd, gg[i], bb[i]);
    store_a8(YA, 128, r, c0, pack8_bf16(y));
  }
}

__device__ __forceinline__ void load_x160(uint32_t taddr, float (&x)[160]) {
#pragma unroll
  for (int c = 0; c < 160; c += 32) {
    float t[32];
    tmem_ld_x32(taddr + c, t);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) x[c + i] = t[i];
  }
}

template <int kV>
__global__ void __launch_bounds__(kTcThreads, 1) xformer_tc_kernel(const TArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Pipe* pipe = reinterpret_cast<Pipe*>(smem + kT_PIPE);
  float* FP = reinterpret_cast<float*>(smem + kT_FP);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int V = kV;
  constexpr int ppt = 128 / V;
  const int64_t ntiles = (a.count + ppt - 1) / ppt;

  if (tid == 0) {
    for (int i = 0; i < kT_Slots; ++i) { mbar_init(&pipe->full[i], 1); mbar_init(&pipe->empty[i], 1); }
    mbar_init(&pipe->a_bar, kEpiThreads);
    mbar_init(&pipe->d_bar, 1);
    mbar_fence_init();
  }
  if (warp == 5) { tmem_alloc(&pipe->tmem_base, 512); tmem_relinquish(); }
  {
    const float* src = reinterpret_cast<const float*>(a.blob + kFloatOff);
    for (int i = tid; i < kTFloats; i += kTcThreads) FP[i] = src[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = pipe->tmem_base;

  if (warp == 4) {
    // ================= weight producer =================
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int l = 0; l < 2; ++l) {
          const uint8_t* src = a.blob + (size_t)l * kTLayerBytes;
          auto push = [&](uint32_t bytes) {
            const uint32_t slot = it % kT_Slots;
            mbar_wait(&pipe->empty[slot], ((it / kT_Slots) & 1) ^ 1);
            mbar_arrive_expect_tx(&pipe->full[slot], bytes);
            bulk_g2s(smem + kT_RING + slot * kT_SlotBytes, src, bytes, &pipe->full[slot]);
            src += bytes;
            ++it;
          };
          for (int c = 0; c < 3; ++c) push(kQkvChunk);
          for (int h = 1; h < 4; ++h) { push(kWoChunk); for (int c = 0; c < 3; ++c) push(kQkvChunk); }
          push(kWoChunk);
          for (int c = 0; c < 3; ++c) push(kW1Chunk);
          for (int c = 0; c < 2; ++c) push(kW2Chunk);
        }
      }
    }
  } else if (warp == 5) {
    // ================= MMA issuer =================
    if (lane == 0) {
      uint32_t it = 0, g = 0;
      const uint32_t sYA = smem_u32(smem + kT_YA), sOA = smem_u32(smem + kT_OA), sRing = smem_u32(smem + kT_RING);
      auto wait_a = [&]() { mbar_wait(&pipe->a_bar, g & 1); tc_fence_after(); };
      auto done = [&]() { mma_commit(&pipe->d_bar); ++g; };
      auto slot_wait = [&]() -> uint32_t {
        const uint32_t slot = it % kT_Slots;
        mbar_wait(&pipe->full[slot], (it / kT_Slots) & 1);
        tc_fence_after();
        return sRing + slot * kT_SlotBytes;
      };
      auto slot_free = [&]() { mma_commit(&pipe->empty[it % kT_Slots]); ++it; };
      // Y (K = 160: 10 k-steps over 3 chunks) x weight chunks with N rows -> D
      auto gemm_y = [&](uint32_t dcol, int N, int nchunks) {
        const uint32_t idesc = instr_desc_bf16(N);
        for (int c = 0; c < nchunks; ++c) {
          const uint32_t b0 = slot_wait();
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int ks = c * 4 + k4;
            if (ks < 10) mma_bf16_ss(tm + dcol, smem_desc_sw128(sYA + c * 16384 + k4 * 32), smem_desc_sw128(b0 + k4 * 32), idesc, ks > 0);
          }
          slot_free();
        }
      };
      // A = chunk(s) of OA (64 K each) x weight chunk(s) with 160 rows, accumulated into X
      auto gemm_into_x = [&](uint32_t a0, int nchunks) {
        const uint32_t idesc = instr_desc_bf16(160);
        for (int c = 0; c < nchunks; ++c) {
          const uint32_t b0 = slot_wait();
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            mma_bf16_ss(tm + kT_ColX, smem_desc_sw128(a0 + c * 16384 + k4 * 32), smem_desc_sw128(b0 + k4 * 32), idesc, 1u);
          slot_free();
        }
      };
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int l = 0; l < 2; ++l) {
          wait_a(); gemm_y(kT_ColR, 192, 3); done();                        // q|k|v of head 0
          for (int h = 1; h < 4; ++h) {
            wait_a();
            gemm_into_x(sOA + (h - 1) * 16384, 1);                          // x += o_{h-1} Wo_{h-1}^T
            gemm_y(kT_ColR, 192, 3);                                        // q|k|v of head h
            done();
          }
          wait_a(); gemm_into_x(sOA + 3 * 16384, 1); done();
          wait_a(); gemm_y(kT_ColR, 128, 3); done();                        // FF hidden
          wait_a(); gemm_into_x(sOA, 2); done();                            // x += gelu(.) W2^T
        }
      }
    }
  } else {
    // ================= epilogue warps =================
    const int r = tid;
    uint8_t* YA = smem + kT_YA;
    uint8_t* OA = smem + kT_OA;
    uint8_t* KX = smem + kT_KX;
    uint8_t* VX = smem + kT_VX;
    const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16);
    uint32_t g = 0;
    auto hand_over = [&]() { tc_fence_before(); fence_proxy_async(); mbar_arrive(&pipe->a_bar); };
    auto wait_d = [&]() { mbar_wait(&pipe->d_bar, g & 1); ++g; tc_fence_after(); };
    constexpr int rows = ppt * V;
    const int p0 = (r < rows) ? (r / V) * V : 0;      // first row of this row's point (attention partners)

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t pnt = tile * ppt + r / V;
      const int tok = r % V;
      const bool valid = (r < rows) && (pnt < a.count);
      {
        // ---- tile load: tokens -> X (TMEM), LN1 of layer 0 -> YA
        float x[160];
        if (valid) {
          const float4* src = reinterpret_cast<const float4*>(a.tokens + (pnt * V + tok) * (int64_t)a.ld);
#pragma unroll
          for (int c = 0; c < 40; ++c) {
            const float4 t = __ldg(src + c);
            x[4 * c] = t.x; x[4 * c + 1] = t.y; x[4 * c + 2] = t.z; x[4 * c + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int c = 0; c < 160; ++c) x[c] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < 160; c += 16) {
          float t[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) t[i] = x[c + i];
          tmem_st_x16(tl + kT_ColX + c, t);
        }
        tmem_st_wait();
        ln_to_operand(x, nullptr, FP, FP + 160, YA, r);
        hand_over();
      }
      for (int l = 0; l < 2; ++l) {
        const float* fp = FP + l * kTLayerFloats;   // ln1_g ln1_b pend_in ln2_g ln2_b pend_mid b1
        for (int h = 0; h < 4; ++h) {
          wait_d();
          // ---- attention of head h (lib/transformer.py:59-71): R = [q | k | v], 64 columns each
          {
            float t[32];
#pragma unroll
            for (int half = 0; half < 2; ++half) {      // k -> KX, v -> VX as fp16
#pragma unroll
              for (int part = 0; part < 2; ++part) {
                tmem_ld_x32(tl + kT_ColR + 64 + 64 * part + 32 * half, t);
                tmem_ld_wait();
                uint8_t* dst = (part == 0 ? KX : VX) + r * 128;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const uint4 pk = make_uint4(pack_h2(t[8 * u], t[8 * u + 1]), pack_h2(t[8 * u + 2], t[8 * u + 3]),
                                              pack_h2(t[8 * u + 4], t[8 * u + 5]), pack_h2(t[8 * u + 6], t[8 * u + 7]));
                  *reinterpret_cast<uint4*>(dst + (((half * 4 + u) ^ (r & 7)) << 4)) = pk;
                }
              }
            }
          }
          named_bar_sync(1, kEpiThreads);
          float q[64];
          {
            float t[32];
            tmem_ld_x32(tl + kT_ColR, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) q[i] = t[i];
            tmem_ld_x32(tl + kT_ColR + 32, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) q[32 + i] = t[i];
          }
          float dots[V];
          float mx = -1e30f;
#pragma unroll
          for (int j = 0; j < V; ++j) {
            const int rj = p0 + j;
            const uint8_t* src = KX + rj * 128;
            float d = 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const uint4 pk = *reinterpret_cast<const uint4*>(src + ((u ^ (rj & 7)) << 4));
              const __half2* h2 = reinterpret_cast<const __half2*>(&pk);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(h2[i]);
                d = fmaf(q[8 * u + 2 * i], f.x, d);
                d = fmaf(q[8 * u + 2 * i + 1], f.y, d);
              }
            }
            dots[j] = d * 0.125f;
            mx = fmaxf(mx, dots[j]);
          }
          float den = 0.f;
#pragma unroll
          for (int j = 0; j < V; ++j) { dots[j] = __expf(dots[j] - mx); den += dots[j]; }
          const float inv = 1.0f / den;
          float o[64];
#pragma unroll
          for (int i = 0; i < 64; ++i) o[i] = 0.f;
#pragma unroll
          for (int j = 0; j < V; ++j) {
            const int rj = p0 + j;
            const float w = dots[j] * inv;
            const uint8_t* src = VX + rj * 128;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const uint4 pk = *reinterpret_cast<const uint4*>(src + ((u ^ (rj & 7)) << 4));
              const __half2* h2 = reinterpret_cast<const __half2*>(&pk);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(h2[i]);
                o[8 * u + 2 * i] = fmaf(w, f.x, o[8 * u + 2 * i]);
                o[8 * u + 2 * i + 1] = fmaf(w, f.y, o[8 * u + 2 * i + 1]);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) store_a8(OA + h * 16384, 128, r, 8 * u, pack8_bf16(o + 8 * u));
          hand_over();
        }
        {
          // ---- x (+ deferred biases) -> LN2 -> YA
          wait_d();
          float x[160];
          load_x160(tl + kT_ColX, x);
          ln_to_operand(x, fp + 800, fp + 480, fp + 640, YA, r);
          hand_over();
        }
        {
          // ---- FF hidden: GELU(acc + b1) -> bf16 operand (K = 128, aliases OA chunks 0-1)
          wait_d();
          const float* b1 = fp + 960;
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) {
            float t[32];
            tmem_ld_x32(tl + kT_ColR + 32 * cb, t);
            tmem_ld_wait();
            const float4* b4 = reinterpret_cast<const float4*>(b1 + 32 * cb);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 bb = b4[i];
              t[4 * i] = gelu_erf(t[4 * i] + bb.x); t[4 * i + 1] = gelu_erf(t[4 * i + 1] + bb.y);
              t[4 * i + 2] = gelu_erf(t[4 * i + 2] + bb.z); t[4 * i + 3] = gelu_erf(t[4 * i + 3] + bb.w);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) store_a8(OA, 128, r, 32 * cb + 8 * u, pack8_bf16(t + 8 * u));
          }
          hand_over();
        }
        {
          wait_d();
          float x[160];
          load_x160(tl + kT_ColX, x);
          if (l == 0) {
            const float* f1 = FP + kTLayerFloats;        // layer 1: LN1 on x + pend_in
            ln_to_operand(x, f1 + 320, f1, f1 + 160, YA, r);
            hand_over();
          } else if (valid && tok < 2) {
            // ---- output tokens 0 (density branch) and 1 (colour branch), lib/skinnning_batch.py:441-442
            const float* pend = FP + 2 * kTLayerFloats;
            __nv_bfloat16* dst = (tok == 0 ? a.tok0 : a.tok1) + pnt * kTokLd;
#pragma unroll
            for (int c0 = 0; c0 < 160; c0 += 8) {
              float y[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) y[i] = x[c0 + i] + pend[c0 + i];
              *reinterpret_cast<uint4*>(dst + c0) = pack8_bf16(y);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tm, 512);
}

// ------------------------------------------------------------------------------------------
// M: canonical NeRF MLP.  Tile = 128 points.  TMEM: one 256-column accumulator.
// A operands: XA = [tok0 155 | 0 x5 | PE6(xc) 39 | 0] (K = 208, reused for tok1 after layer 5),
// HA = hidden activations (K = 256).  Weight ring slots are half chunks (128 rows x 128 B).
// ------------------------------------------------------------------------------------------
constexpr uint32_t kM_XA = 0;
constexpr uint32_t kM_HA = kM_XA + 4 * 16384;
constexpr uint32_t kM_RING = kM_HA + 4 * 16384;
constexpr int kM_Slots = 5;
constexpr uint32_t kM_SlotBytes = 16384;
constexpr uint32_t kM_FP = kM_RING + kM_Slots * kM_SlotBytes;
constexpr uint32_t kM_PIPE = kM_FP + kMFloats * 4;
constexpr uint32_t kM_Smem = kM_PIPE + sizeof(Pipe);
static_assert(kM_Smem <= 232448 - 1024, "M kernel shared memory over budget");

struct MArgs {
  const __nv_bfloat16* tok0;
  const __nv_bfloat16* tok1;
  const float* xc;          // (count, 3) canonical points
  int64_t count;
  const uint8_t* blob;
  const int32_t* act_pid;   // already offset by `first`
  float* raw;               // (P, 4)
};

// Hidden-layer epilogue: 256 accumulator columns of this thread's row -> (+bias, ReLU) -> bf16 A
// operand.  64 columns per TMEM round trip, biases as 128-bit broadcast loads, everything
// compile-time so the inner loops are branch-free.  Returns sum_j act_j * w_alpha_j when kAlpha.
template <bool kRelu, bool kAlpha>
__device__ __forceinline__ float epi_hidden(uint32_t tl, const float* __restrict__ b,
                                            const float* __restrict__ w_alpha, uint8_t* HA, int r) {
  float acc4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int cb = 0; cb < 4; ++cb) {
    float t[64];
    tmem_ld_x32(tl + 64 * cb, *reinterpret_cast<float(*)[32]>(&t[0]));
    tmem_ld_x32(tl + 64 * cb + 32, *reinterpret_cast<float(*)[32]>(&t[32]));
    tmem_ld_wait();
    const float4* b4 = reinterpret_cast<const float4*>(b + 64 * cb);
    const float4* w4 = reinterpret_cast<const float4*>(w_alpha + 64 * cb);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 bb = b4[j];
      float v0 = t[4 * j] + bb.x, v1 = t[4 * j + 1] + bb.y, v2 = t[4 * j + 2] + bb.z, v3 = t[4 * j + 3] + bb.w;
      if (kRelu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
      if (kAlpha) {
        const float4 ww = w4[j];
        acc4[0] = fmaf(v0, ww.x, acc4[0]); acc4[1] = fmaf(v1, ww.y, acc4[1]);
        acc4[2] = fmaf(v2, ww.z, acc4[2]); acc4[3] = fmaf(v3, ww.w, acc4[3]);
      }
      t[4 * j] = v0; t[4 * j + 1] = v1; t[4 * j + 2] = v2; t[4 * j + 3] = v3;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) store_a8(HA, 128, r, 64 * cb + 8 * u, pack8_bf16(t + 8 * u));
  }
  return (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]);
}

__global__ void __launch_bounds__(kTcThreads, 1) mlp_tc_kernel(const MArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Pipe* pipe = reinterpret_cast<Pipe*>(smem + kM_PIPE);
  float* FP = reinterpret_cast<float*>(smem + kM_FP);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (a.count + 127) / 128;

  if (tid == 0) {
    for (int i = 0; i < kM_Slots; ++i) { mbar_init(&pipe->full[i], 1); mbar_init(&pipe->empty[i], 1); }
    mbar_init(&pipe->a_bar, kEpiThreads);
    mbar_init(&pipe->d_bar, 1);
    mbar_fence_init();
  }
  if (warp == 5) { tmem_alloc(&pipe->tmem_base, 256); tmem_relinquish(); }
  {
    const float* src = reinterpret_cast<const float*>(a.blob + kFloatOff) + kTFloats;
    for (int i = tid; i < kMFloats; i += kTcThreads) FP[i] = src[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = pipe->tmem_base;

  if (warp == 4) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint8_t* src = a.blob + kTBytes;
        constexpr int kItems = (kMBytes / kM_SlotBytes);      // the whole MLP section, in order
        for (int i = 0; i < kItems; ++i) {
          const uint32_t slot = it % kM_Slots;
          mbar_wait(&pipe->empty[slot], ((it / kM_Slots) & 1) ^ 1);
          mbar_arrive_expect_tx(&pipe->full[slot], kM_SlotBytes);
          bulk_g2s(smem + kM_RING + slot * kM_SlotBytes, src, kM_SlotBytes, &pipe->full[slot]);
          src += kM_SlotBytes;
          ++it;
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      uint32_t it = 0, g = 0;
      const uint32_t sXA = smem_u32(smem + kM_XA), sHA = smem_u32(smem + kM_HA), sRing = smem_u32(smem + kM_RING);
      const uint32_t idesc = instr_desc_bf16(128);
      auto wait_a = [&]() { mbar_wait(&pipe->a_bar, g & 1); tc_fence_after(); };
      auto done = [&]() { mma_commit(&pipe->d_bar); ++g; };
      // one K-chunk of A (at a0, `ksteps` valid K=16 steps) against `nhalves` 128-row weight halves
      auto chunk = [&](uint32_t a0, int ksteps, int nhalves, bool first) {
        for (int n = 0; n < nhalves; ++n) {
          const uint32_t slot = it % kM_Slots;
          mbar_wait(&pipe->full[slot], (it / kM_Slots) & 1);
          tc_fence_after();
          const uint32_t b0 = sRing + slot * kM_SlotBytes;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            if (k4 < ksteps)
              mma_bf16_ss(tm + n * 128, smem_desc_sw128(a0 + k4 * 32), smem_desc_sw128(b0 + k4 * 32), idesc, (!first || k4 > 0) ? 1u : 0u);
          mma_commit(&pipe->empty[slot]);
          ++it;
        }
      };
      auto x_part = [&](int nhalves, int ksteps_total, bool first) {       // XA: 13 steps (x) or 10 (tok1)
        const int nch = (ksteps_total + 3) / 4;
        for (int c = 0; c < 4; ++c) {
          const int ks = min(4, ksteps_total - 4 * c);
          if (c < nch) chunk(sXA + c * 16384, ks, nhalves, first && c == 0);
          else if (nhalves == 2) chunk(sXA + c * 16384, 0, nhalves, false);   // keep the weight stream in step
        }
      };
      auto h_part = [&](int nhalves, bool first) {
        for (int c = 0; c < 4; ++c) chunk(sHA + c * 16384, 4, nhalves, first && c == 0);
      };
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        wait_a(); x_part(2, 13, true); done();                              // L0
        for (int L = 1; L < 5; ++L) { wait_a(); h_part(2, true); done(); }  // L1..L4
        wait_a(); x_part(2, 13, true); h_part(2, false); done();            // L5: [x | h]
        for (int L = 6; L < 8; ++L) { wait_a(); h_part(2, true); done(); }  // L6, L7
        wait_a(); h_part(2, true); done();                                  // feature
        wait_a();                                                           // views: [tok1 | feature], N = 128
        for (int c = 0; c < 3; ++c) chunk(sXA + c * 16384, c < 2 ? 4 : 2, 1, c == 0);
        for (int c = 0; c < 4; ++c) chunk(sHA + c * 16384, 4, 1, false);
        done();
      }
    }
  } else {
    const int r = tid;
    uint8_t* XA = smem + kM_XA;
    uint8_t* HA = smem + kM_HA;
    const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16);
    uint32_t g = 0;
    auto hand_over = [&]() { tc_fence_before(); fence_proxy_async(); mbar_arrive(&pipe->a_bar); };
    auto wait_d = [&]() { mbar_wait(&pipe->d_bar, g & 1); ++g; tc_fence_after(); };
    const float* bias = FP;                       // 8 x 256
    const float* w_alpha = FP + 2048;
    const float* b_feat = FP + 2304;
    const float* b_views = FP + 2560;
    const float* w_rgb = FP + 2688;               // 3 x 128
    const float* b_tail = FP + 3072;              // b_alpha, b_rgb[3]

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t i = tile * 128 + r;
      const bool valid = i < a.count;
      {
        // ---- tile load: XA = [tok0 | 0 | PE6(xc) | 0]
        const uint4* t0 = reinterpret_cast<const uint4*>(a.tok0 + (valid ? i : 0) * kTokLd);
#pragma unroll
        for (int u = 0; u < 20; ++u) store_a8(XA, 128, r, 8 * u, valid ? __ldg(t0 + u) : make_uint4(0, 0, 0, 0));
        float pe[48];
        float xc[3] = {0.f, 0.f, 0.f};
        if (valid) { xc[0] = a.xc[3 * i]; xc[1] = a.xc[3 * i + 1]; xc[2] = a.xc[3 * i + 2]; }
#pragma unroll
        for (int e = 0; e < 48; ++e) {
          float v = 0.f;
          if (e < 3) v = xc[e];
          else if (e < 39) {
            const int k = (e - 3) / 6, ch = (e - 3) % 3;
            const bool is_cos = ((e - 3) % 6) >= 3;
            v = sinf(fmaf(xc[ch], 3.14159265358979323846f * (float)(1 << k), is_cos ? 1.57079632679489661923f : 0.0f));
          }
          pe[e] = v;
        }
#pragma unroll
        for (int u = 0; u < 6; ++u) store_a8(XA, 128, r, 160 + 8 * u, pack8_bf16(pe + 8 * u));
        hand_over();
      }
      float alpha = 0.f;
      for (int L = 0; L < 9; ++L) {          // L0..L7 (ReLU) and 8 = feature (no activation)
        wait_d();
        if (L == 5) {                        // x is dead after layer 5: XA <- tok1 for the views layer
          const uint4* t1 = reinterpret_cast<const uint4*>(a.tok1 + (valid ? i : 0) * kTokLd);
#pragma unroll
          for (int u = 0; u < 20; ++u) store_a8(XA, 128, r, 8 * u, valid ? __ldg(t1 + u) : make_uint4(0, 0, 0, 0));
        }
        // the layer kind is warp-uniform: pick a fully specialised epilogue (no per-element branches)
        if (L == 7) alpha = epi_hidden<true, true>(tl, bias + 256 * 7, w_alpha, HA, r);   // + alpha_linear on fp32 acts
        else if (L == 8) epi_hidden<false, false>(tl, b_feat, nullptr, HA, r);
        else epi_hidden<true, false>(tl, bias + 256 * L, nullptr, HA, r);
        hand_over();
      }
      {
        // ---- views layer epilogue: relu -> rgb_linear on CUDA cores -> raw[pid] = (rgb, alpha)
        wait_d();
        float c0[2] = {0.f, 0.f}, c1[2] = {0.f, 0.f}, c2[2] = {0.f, 0.f};
#pragma unroll 1
        for (int cb = 0; cb < 2; ++cb) {
          float t[64];
          tmem_ld_x32(tl + 64 * cb, *reinterpret_cast<float(*)[32]>(&t[0]));
          tmem_ld_x32(tl + 64 * cb + 32, *reinterpret_cast<float(*)[32]>(&t[32]));
          tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(b_views + 64 * cb);
          const float4* r0 = reinterpret_cast<const float4*>(w_rgb + 64 * cb);
          const float4* r1 = reinterpret_cast<const float4*>(w_rgb + 128 + 64 * cb);
          const float4* r2 = reinterpret_cast<const float4*>(w_rgb + 256 + 64 * cb);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 bb = b4[j], w0 = r0[j], w1 = r1[j], w2 = r2[j];
            const float v0 = fmaxf(t[4 * j] + bb.x, 0.f), v1 = fmaxf(t[4 * j + 1] + bb.y, 0.f);
            const float v2 = fmaxf(t[4 * j + 2] + bb.z, 0.f), v3 = fmaxf(t[4 * j + 3] + bb.w, 0.f);
            c0[0] = fmaf(v0, w0.x, c0[0]); c0[1] = fmaf(v1, w0.y, c0[1]); c0[0] = fmaf(v2, w0.z, c0[0]); c0[1] = fmaf(v3, w0.w, c0[1]);
            c1[0] = fmaf(v0, w1.x, c1[0]); c1[1] = fmaf(v1, w1.y, c1[1]); c1[0] = fmaf(v2, w1.z, c1[0]); c1[1] = fmaf(v3, w1.w, c1[1]);
            c2[0] = fmaf(v0, w2.x, c2[0]); c2[1] = fmaf(v1, w2.y, c2[1]); c2[0] = fmaf(v2, w2.z, c2[0]); c2[1] = fmaf(v3, w2.w, c2[1]);
          }
        }
        if (valid)
          reinterpret_cast<float4*>(a.raw)[a.act_pid[i]] =
              make_float4(c0[0] + c0[1] + b_tail[1], c1[0] + c1[1] + b_tail[2], c2[0] + c2[1] + b_tail[3], alpha + b_tail[0]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tm, 256);
}

// ------------------------------------------------------------------------------------------
// Diagnostic: one 128 x N x K tile through the exact building blocks of the fused kernels.
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
  (void)n_views;
  const size_t c = (size_t)(count > 0 ? count : 0);
  return 2 * (c * mps::kTokLd * sizeof(__nv_bfloat16) + 256) + 256;
}

extern "C" int mpsnerf_dense_bf16(const float* tokens, int32_t ld, const float* xc, int64_t count,
                                  int n_views, const void* packed, size_t packed_bytes,
                                  const int32_t* act_pid, int64_t first, float* raw, void* workspace,
                                  void* stream) {
  using namespace mps;
  MPS_REQUIRE(count >= 0 && n_views >= 2 && n_views <= 4);   // tensor-core path: 2..4 input views (fp32 path: up to 8)
  if (count == 0) return MPSNERF_OK;
  MPS_REQUIRE(tokens && xc && packed && act_pid && raw && workspace);
  MPS_REQUIRE(ld == MPSNERF_TOKEN_LD);
  MPS_REQUIRE(packed_bytes == kBlobBytes);
  MPS_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  MPS_REQUIRE((reinterpret_cast<uintptr_t>(tokens) & 15) == 0 && (reinterpret_cast<uintptr_t>(raw) & 15) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t tok_bytes = ((size_t)count * kTokLd * sizeof(__nv_bfloat16) + 255) / 256 * 256;
  __nv_bfloat16* tok0 = reinterpret_cast<__nv_bfloat16*>(workspace);
  __nv_bfloat16* tok1 = reinterpret_cast<__nv_bfloat16*>(static_cast<char*>(workspace) + tok_bytes);

  static bool attr_done = false;     // idempotent attribute set; benign if raced
  if (!attr_done) {
    MPS_CUDA(cudaFuncSetAttribute(xformer_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kT_Smem));
    MPS_CUDA(cudaFuncSetAttribute(xformer_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kT_Smem));
    MPS_CUDA(cudaFuncSetAttribute(xformer_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kT_Smem));
    MPS_CUDA(cudaFuncSetAttribute(mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kM_Smem));
    attr_done = true;
  }
  {
    TArgs ta{tokens, ld, count, n_views, static_cast<const uint8_t*>(packed), tok0, tok1};
    const int ppt = 128 / n_views;
    int64_t tiles = (count + ppt - 1) / ppt;
    const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
    if (n_views == 2) xformer_tc_kernel<2><<<grid, kTcThreads, kT_Smem, st>>>(ta);
    else if (n_views == 3) xformer_tc_kernel<3><<<grid, kTcThreads, kT_Smem, st>>>(ta);
    else xformer_tc_kernel<4><<<grid, kTcThreads, kT_Smem, st>>>(ta);
    MPS_LAUNCH_CHECK();
  }
  {
    MArgs ma{tok0, tok1, xc, count, static_cast<const uint8_t*>(packed), act_pid + first, raw};
    int64_t tiles = (count + 127) / 128;
    const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
    mlp_tc_kernel<<<grid, kTcThreads, kM_Smem, st>>>(ma);
    MPS_LAUNCH_CHECK();
  }
  return MPSNERF_OK;
}
