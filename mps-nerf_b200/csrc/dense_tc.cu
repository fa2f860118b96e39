// K5 (bf16 production path): fused tcgen05 / TMEM kernels for the cross-view transformer ("T")
// and the canonical NeRF MLP ("M").  Restates lib/transformer.py:13-86 and
// lib/skinnning_batch.py:438-473 with bf16 operands and fp32 accumulation.
//
// Structure (one persistent CTA per SM, 320 threads):
//   warps 0-7  epilogue, 256 threads: thread (r = tid & 127, half = tid >> 7) owns half of the
//              columns of row r (TMEM lane r; warps w and w+4 share lane quarter w & 3).  They turn
//              accumulators into the next A operand (bias / ReLU / GELU / LayerNorm / attention)
//              and store it, bf16-packed, back into TENSOR MEMORY (tcgen05.st);
//   warp 8     weight producer: one thread streams the pre-swizzled weight chunks with bulk
//              async copies (TMA engine) into an mbarrier ring, in consumption order
//              (multicast across the kC CTAs of a cluster);
//   warp 9     MMA issuer: one thread issues TS-form tcgen05.mma (A from TMEM, B from smem, M = 128).
// Activations never leave the SM between layers and never touch shared memory: shared memory
// holds only the weight ring (6-7 chunks in flight), which is what hides the L2 latency.
// MMA <-> epilogue hand-over is a pair of mbarriers (a_bar: "A operand ready", 256 arrivals;
// d_bar: "accumulator ready", tcgen05.commit).
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace mps {
using namespace umma;

constexpr int kTcThreads = 320;
constexpr int kEpiThreads = 256;
constexpr int kProdWarp = 8, kMmaWarp = 9;

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

// Optional in-kernel cycle accounting (MPSNERF_TC_PROF=1): per kernel 8 counters summed over CTAs:
// 0 epilogue wait-for-MMA, 1 epilogue work, 2 tile load, 3 MMA wait-for-A, 4 MMA wait-for-weights,
// 5 MMA thread total, 6 producer wait-for-free-slot, 7 tiles.  Read with mpsnerf_debug_read_prof().
__device__ unsigned long long g_prof[2][16];   // [8..15]: T epilogue sections, see tools/step.py
struct Prof {
  bool on;
  long long t;
  __device__ __forceinline__ void start() { if (on) t = clock64(); }
  __device__ __forceinline__ void stop(long long& acc) { if (on) { const long long n = clock64(); acc += n - t; t = n; } }
};

struct Pipe {            // barriers of one CTA (in dynamic smem)
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t a_bar;
  uint64_t d_bar;
  uint64_t r_bar;       // T kernel: "scratch accumulator R has been copied out" (early release)
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ float gelu_erf(float x) {
  // 0.5 x (1 + erf(x / sqrt 2)) with erf(z) = z * P(z^2) on |z| <= 3 (degree-9 Chebyshev fit, max
  // |err| 1.6e-5 in fp32; saturates to +-0.99998 beyond).  Deliberately free of MUFU ops (ex2 / rcp):
  // the special-function unit was the bottleneck of this epilogue (2 MUFU per element).
  const float z = fminf(fmaxf(x * 0.70710678118654752440f, -3.0f), 3.0f);
  const float w = z * z;
  float p = -4.6617889859e-09f;
  p = fmaf(p, w, 2.3821795289e-07f);
  p = fmaf(p, w, -5.4625094310e-06f);
  p = fmaf(p, w, 7.5274732228e-05f);
  p = fmaf(p, w, -7.0841461755e-04f);
  p = fmaf(p, w, 4.9218977801e-03f);
  p = fmaf(p, w, -2.6500707362e-02f);
  p = fmaf(p, w, 1.1261424783e-01f);
  p = fmaf(p, w, -3.7607604539e-01f);
  p = fmaf(p, w, 1.1283780006e+00f);
  const float hx = 0.5f * x;
  return fmaf(hx, z * p, hx);
}

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 t = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// epilogue-side helpers shared by both kernels
__device__ __forceinline__ void epi_bar() { named_bar_sync(1, kEpiThreads); }

template <int kSlots, uint32_t kSlotBytes, int kC>
struct Producer {            // used by the single producer thread
  Pipe* pipe;
  uint8_t* ring;
  uint32_t crank;
  uint32_t it = 0;
  Prof pf;
  long long acc_e = 0;
  __device__ __forceinline__ void push(const uint8_t*& src, uint32_t bytes) {
    const uint32_t slot = it % kSlots;
    pf.start();
    mbar_wait(&pipe->empty[slot], ((it / kSlots) & 1) ^ 1);
    pf.stop(acc_e);
    mbar_arrive_expect_tx(&pipe->full[slot], bytes);          // the whole chunk lands here (kC slices)
    if (kC == 1) {
      bulk_g2s(ring + slot * kSlotBytes, src, bytes, &pipe->full[slot]);
    } else {
      const uint32_t slice = bytes / kC;
      bulk_g2s_mc(ring + slot * kSlotBytes + crank * slice, src + crank * slice, slice, &pipe->full[slot],
                  (uint16_t)((1u << kC) - 1u));
    }
    src += bytes;
    ++it;
  }
};

template <int kSlots, uint32_t kSlotBytes, int kC>
struct Consumer {            // used by the single MMA thread
  Pipe* pipe;
  uint32_t ring_addr;
  uint32_t it = 0, g = 0;
  Prof pf;
  long long acc_a = 0, acc_w = 0;
  __device__ __forceinline__ void wait_a() { pf.start(); mbar_wait(&pipe->a_bar, g & 1); pf.stop(acc_a); tc_fence_after(); }
  __device__ __forceinline__ void done() { mma_commit(&pipe->d_bar); ++g; }
  __device__ __forceinline__ uint32_t slot_wait() {
    const uint32_t slot = it % kSlots;
    pf.start();
    mbar_wait(&pipe->full[slot], (it / kSlots) & 1);
    pf.stop(acc_w);
    tc_fence_after();
    return ring_addr + slot * kSlotBytes;
  }
  __device__ __forceinline__ void slot_free() {
    if (kC == 1) mma_commit(&pipe->empty[it % kSlots]);
    else mma_commit_mc(&pipe->empty[it % kSlots], (uint16_t)((1u << kC) - 1u));
    ++it;
  }
};

// ------------------------------------------------------------------------------------------
// T: cross-view transformer.  Tile = ppt = 128 / V points, row r = (point r / V, token r % V).
// TMEM columns:
//   X  [0,160)    fp32 residual stream; the out-proj and FF2 GEMMs accumulate into it (that IS the
//                 residual add; their biases are deferred, see pack.py)
//   R  [160,352)  scratch accumulator: q|k|v of one head (192) or the FF hidden layer (128)
//   YT [352,432)  LayerNorm output, bf16 x2 per column (K = 160)           -> A of qkv / FF1
//   OT [432,496)  attention output of one head (K = 64, 32 cols) or GELU(FF hidden) (K = 128, 64 cols)
// ------------------------------------------------------------------------------------------
constexpr uint32_t kT_ColX = 0, kT_ColR = 160, kT_ColY = 352, kT_ColO = 432;
constexpr uint32_t kT_KX = 0;                      // k of the current head, fp32 [128][64], rows of 256 B, unit-swizzled
constexpr uint32_t kT_VX = kT_KX + 32768;          // v of the current head, fp16 [128][64], rows of 128 B, unit-swizzled
constexpr uint32_t kT_PD = kT_VX + 16384;          // partial q.k dots  float[2][128][4]
constexpr uint32_t kT_LS = kT_PD + 4096;           // LayerNorm partial sums float[2][2][128]
constexpr uint32_t kT_RING = kT_LS + 2048;         // 1024-aligned: 32768 + 16384 + 4096 + 2048 = 55296 = 54 * 1024
constexpr int kT_Slots = 6;
constexpr uint32_t kT_SlotBytes = kQkvChunk;
constexpr uint32_t kT_FP = kT_RING + kT_Slots * kT_SlotBytes;
constexpr uint32_t kT_PIPE = kT_FP + ((kTFloats * 4 + 15) / 16) * 16;
constexpr uint32_t kT_Smem = kT_PIPE + sizeof(Pipe);
static_assert(kT_RING % 1024 == 0 && kT_SlotBytes % 1024 == 0, "ring slots must be 1024-byte aligned");
static_assert(kT_Smem <= 232448 - 1024, "T kernel shared memory over budget");

struct TArgs {
  const float* tokens;   // (count, V, ld) fp32
  int ld;
  int64_t count;
  int V;
  const uint8_t* blob;
  __nv_bfloat16* tok0;   // (count, 160)
  __nv_bfloat16* tok1;
  int prof;
};

// LayerNorm over the 155 real columns of a row whose 160 columns are split between two threads
// (this thread: x[0..80) = columns 80*half ..; real columns: 80 or 75), result bf16-packed -> YT.
// gamma/beta are zero in the 5 pad columns, so the pad of the operand is exactly zero.
__device__ __forceinline__ void ln_to_tmem(float (&x)[80], const float* __restrict__ pend,
                                           const float* __restrict__ g, const float* __restrict__ b, float* LS,
                                           int r, int half, uint32_t tl) {
  const int nreal = half ? 75 : 80;
  if (pend) {
    const float4* p4 = reinterpret_cast<const float4*>(pend + 80 * half);
#pragma unroll
    for (int c = 0; c < 20; ++c) {
      const float4 t = p4[c];
      x[4 * c] += t.x; x[4 * c + 1] += t.y; x[4 * c + 2] += t.z; x[4 * c + 3] += t.w;
    }
  }
  float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 80; ++c) if (c < nreal) s[c & 3] += x[c];
  LS[half * 128 + r] = (s[0] + s[1]) + (s[2] + s[3]);
  epi_bar();
  const float mean = (LS[r] + LS[128 + r]) * (1.0f / 155.0f);
  float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 80; ++c) if (c < nreal) { const float d = x[c] - mean; v[c & 3] = fmaf(d, d, v[c & 3]); }
  LS[256 + half * 128 + r] = (v[0] + v[1]) + (v[2] + v[3]);
  epi_bar();
  const float rstd = rsqrtf((LS[256 + r] + LS[384 + r]) * (1.0f / 155.0f) + 1e-5f);
  const float4* g4 = reinterpret_cast<const float4*>(g + 80 * half);
  const float4* b4 = reinterpret_cast<const float4*>(b + 80 * half);
  uint32_t pk[40];
#pragma unroll
  for (int c = 0; c < 20; ++c) {
    const float4 gg = g4[c], bb = b4[c];
    const float y0 = fmaf((x[4 * c] - mean) * rstd, gg.x, bb.x), y1 = fmaf((x[4 * c + 1] - mean) * rstd, gg.y, bb.y);
    const float y2 = fmaf((x[4 * c + 2] - mean) * rstd, gg.z, bb.z), y3 = fmaf((x[4 * c + 3] - mean) * rstd, gg.w, bb.w);
    pk[2 * c] = pack_bf16x2(y0, y1);
    pk[2 * c + 1] = pack_bf16x2(y2, y3);
  }
#pragma unroll
  for (int i = 0; i < 5; ++i)      // every TMEM access is aligned to its own width (here 8 columns)
    tmem_st_x8(tl + kT_ColY + 40 * half + 8 * i, *reinterpret_cast<const uint32_t(*)[8]>(&pk[8 * i]));
  tmem_st_wait();
}

__device__ __forceinline__ void load_x80(uint32_t taddr, float (&x)[80]) {
#pragma unroll
  for (int i = 0; i < 5; ++i) tmem_ld_x16(taddr + 16 * i, *reinterpret_cast<float(*)[16]>(&x[16 * i]));
  tmem_ld_wait();
}

// kC = CTAs per cluster.  The kC CTAs of a cluster run the same program on neighbouring tiles and
// share the weight stream: CTA j loads slice j of every chunk and multicasts it to all of them, so
// each chunk crosses L2 -> SM once per cluster.  A ring slot is reused only after the MMA warps of
// *all* kC CTAs have released it (empty barriers count kC arrivals, delivered by multicast commits).
template <int kV, int kC>
__global__ void __launch_bounds__(kTcThreads, 1) xformer_tc_kernel(const TArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Pipe* pipe = reinterpret_cast<Pipe*>(smem + kT_PIPE);
  float* FP = reinterpret_cast<float*>(smem + kT_FP);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int V = kV;
  constexpr int ppt = 128 / V;
  const int64_t ntiles = (a.count + ppt - 1) / ppt;
  const uint32_t crank = (kC > 1) ? cluster_ctarank() : 0u;
  const int64_t ncl = gridDim.x / kC, cid = blockIdx.x / kC;

  if (tid == 0) {
    for (int i = 0; i < kT_Slots; ++i) { mbar_init(&pipe->full[i], 1); mbar_init(&pipe->empty[i], kC); }
    mbar_init(&pipe->a_bar, kEpiThreads);
    mbar_init(&pipe->d_bar, 1);
    mbar_init(&pipe->r_bar, kEpiThreads);
    mbar_fence_init();
  }
  if (warp == kMmaWarp) { tmem_alloc(&pipe->tmem_base, 512); tmem_relinquish(); }
  {
    const float* src = reinterpret_cast<const float*>(a.blob + kFloatOff);
    for (int i = tid; i < kTFloats; i += kTcThreads) FP[i] = src[i];
  }
  tc_fence_before();
  __syncthreads();
  if (kC > 1) cluster_sync_all();      // peers' barriers are initialised before any multicast touches them
  tc_fence_after();
  const uint32_t tm = pipe->tmem_base;

  if (warp == kProdWarp) {
    // ================= weight producer =================
    if (lane == 0) {
      Producer<kT_Slots, kT_SlotBytes, kC> P{pipe, smem + kT_RING, crank};
      P.pf = Prof{a.prof != 0, 0};
      for (int64_t tbase = cid * kC; tbase < ntiles; tbase += ncl * kC) {   // cluster-uniform trip count
        for (int l = 0; l < 2; ++l) {
          const uint8_t* src = a.blob + (size_t)l * kTLayerBytes;
          for (int c = 0; c < 3; ++c) P.push(src, kQkvChunk);                                   // qkv_0
          for (int h = 0; h < 3; ++h) { for (int c = 0; c < 3; ++c) P.push(src, kQkvChunk); P.push(src, kWoChunk); }   // qkv_{h+1}, Wo_h
          P.push(src, kWoChunk);                                                                 // Wo_3
          for (int c = 0; c < 3; ++c) P.push(src, kW1Chunk);
          for (int c = 0; c < 2; ++c) P.push(src, kW2Chunk);
        }
      }
      if (a.prof) atomicAdd(&g_prof[0][6], (unsigned long long)P.acc_e);
    }
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer =================
    if (lane == 0) {
      Consumer<kT_Slots, kT_SlotBytes, kC> Cn{pipe, smem_u32(smem + kT_RING)};
      Cn.pf = Prof{a.prof != 0, 0};
      uint32_t rr = 0;
      const long long t_begin = clock64();
      // A (TMEM, `ksteps` K=16 steps starting at column acol) x weight chunks with N rows -> D column dcol
      auto gemm = [&](uint32_t dcol, uint32_t acol, int ksteps, int N, bool accumulate) {
        const uint32_t idesc = instr_desc_bf16(N);
        for (int k0 = 0; k0 < ksteps; k0 += 4) {
          const uint32_t b0 = Cn.slot_wait();
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int ks = k0 + k4;
            if (ks < ksteps)
              mma_bf16_ts(tm + dcol, tm + acol + ks * 8, smem_desc_sw128(b0 + k4 * 32), idesc, (accumulate || ks > 0) ? 1u : 0u);
          }
          Cn.slot_free();
        }
      };
      for (int64_t tbase = cid * kC; tbase < ntiles; tbase += ncl * kC) {
        for (int l = 0; l < 2; ++l) {
          Cn.wait_a(); gemm(kT_ColR, kT_ColY, 10, 192, false); Cn.done();           // q|k|v of head 0
          for (int h = 0; h < 4; ++h) {
            if (h < 3) {                                                             // R copied out by the epilogue:
              mbar_wait(&pipe->r_bar, rr & 1); ++rr; tc_fence_after();               // q|k|v of head h+1 overlaps the
              gemm(kT_ColR, kT_ColY, 10, 192, false);                                // attention math of head h
            }
            Cn.wait_a();
            gemm(kT_ColX, kT_ColO, 4, 160, true);                                    // x += o_h Wo_h^T
            Cn.done();
          }
          Cn.wait_a(); gemm(kT_ColR, kT_ColY, 10, 128, false); Cn.done();            // FF hidden
          Cn.wait_a(); gemm(kT_ColX, kT_ColO, 8, 160, true); Cn.done();              // x += gelu(.) W2^T
        }
      }
      if (a.prof) {
        atomicAdd(&g_prof[0][3], (unsigned long long)Cn.acc_a);
        atomicAdd(&g_prof[0][4], (unsigned long long)Cn.acc_w);
        atomicAdd(&g_prof[0][5], (unsigned long long)(clock64() - t_begin));
      }
    }
  } else {
    // ================= epilogue warps =================
    const int r = tid & 127, half = tid >> 7;
    uint8_t* KX = smem + kT_KX;
    uint8_t* VX = smem + kT_VX;
    float* PD = reinterpret_cast<float*>(smem + kT_PD);
    float* LS = reinterpret_cast<float*>(smem + kT_LS);
    const uint32_t tl = tm + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t g = 0;
    Prof pf{a.prof != 0 && tid == 0, 0};
    long long acc_d = 0, acc_tl = 0, n_tiles = 0;
    long long sec[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // publish, bar1, dots, bar2, softmax+o, LN2, GELU, LN1/final
    const long long t_begin = clock64();
    auto hand_over = [&]() { tc_fence_before(); mbar_arrive(&pipe->a_bar); };
    auto wait_d = [&]() { pf.start(); mbar_wait(&pipe->d_bar, g & 1); pf.stop(acc_d); ++g; tc_fence_after(); };
    constexpr int rows = ppt * V;
    const int p0 = (r < rows) ? (r / V) * V : 0;      // first row of this row's point (attention partners)

    for (int64_t tbase = cid * kC; tbase < ntiles; tbase += ncl * kC) {
      const int64_t tile = tbase + crank;              // tile >= ntiles: all rows invalid
      const int64_t pnt = tile * ppt + r / V;
      const int tok = r % V;
      const bool valid = (r < rows) && (pnt < a.count);
      {
        // ---- tile load: tokens -> X (TMEM), LN1 of layer 0 -> YT
        pf.start();
        float x[80];
        if (valid) {
          const float4* src = reinterpret_cast<const float4*>(a.tokens + (pnt * V + tok) * (int64_t)a.ld + 80 * half);
#pragma unroll
          for (int c = 0; c < 20; ++c) {
            const float4 t = __ldg(src + c);
            x[4 * c] = t.x; x[4 * c + 1] = t.y; x[4 * c + 2] = t.z; x[4 * c + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int c = 0; c < 80; ++c) x[c] = 0.f;
        }
        {
          uint32_t xb[80];
#pragma unroll
          for (int c = 0; c < 80; ++c) xb[c] = __float_as_uint(x[c]);
#pragma unroll
          for (int i = 0; i < 5; ++i) tmem_st_u16(tl + kT_ColX + 80 * half + 16 * i, xb + 16 * i);   // 16-aligned
        }
        ln_to_tmem(x, nullptr, FP, FP + 160, LS, r, half, tl);
        pf.stop(acc_tl);
        hand_over();
      }
      for (int l = 0; l < 2; ++l) {
        const float* fp = FP + l * kTLayerFloats;   // ln1_g ln1_b pend_in ln2_g ln2_b pend_mid b1
        for (int h = 0; h < 4; ++h) {
          wait_d();
          pf.start();
          // ---- attention of head h (lib/transformer.py:59-71): R = [q | k | v], 64 columns each.
          // half 0 publishes k as fp32, half 1 publishes v as fp16 (unit-swizzled rows)
          {
            float t[64];
            tmem_ld_x32(tl + kT_ColR + 64 + 64 * half, *reinterpret_cast<float(*)[32]>(&t[0]));
            tmem_ld_x32(tl + kT_ColR + 64 + 64 * half + 32, *reinterpret_cast<float(*)[32]>(&t[32]));
            tmem_ld_wait();
            if (half == 0) {
              uint8_t* dst = KX + r * 256;
#pragma unroll
              for (int u = 0; u < 16; ++u)
                *reinterpret_cast<float4*>(dst + ((u ^ (r & 15)) << 4)) = make_float4(t[4 * u], t[4 * u + 1], t[4 * u + 2], t[4 * u + 3]);
            } else {
              uint8_t* dst = VX + r * 128;
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const uint4 pk = make_uint4(pack_h2(t[8 * u], t[8 * u + 1]), pack_h2(t[8 * u + 2], t[8 * u + 3]),
                                            pack_h2(t[8 * u + 4], t[8 * u + 5]), pack_h2(t[8 * u + 6], t[8 * u + 7]));
                *reinterpret_cast<uint4*>(dst + ((u ^ (r & 7)) << 4)) = pk;
              }
            }
          }
          float q[32];
          tmem_ld_x32(tl + kT_ColR + 32 * half, q);
          tmem_ld_wait();
          if (h < 3) { tc_fence_before(); mbar_arrive(&pipe->r_bar); }   // R is free: the next head's q|k|v may land
          pf.stop(sec[0]);
          epi_bar();
          pf.stop(sec[1]);
          // partial dots over this thread's 32 of the 64 head dims (fp32, no conversions)
#pragma unroll
          for (int j = 0; j < V; ++j) {
            const int rj = p0 + j;
            const uint8_t* src = KX + rj * 256;
            float4 kk[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)      // all loads of the row first: one exposed smem latency per row
              kk[u] = *reinterpret_cast<const float4*>(src + (((8 * half + u) ^ (rj & 15)) << 4));
            float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              d[0] = fmaf(q[4 * u], kk[u].x, d[0]); d[1] = fmaf(q[4 * u + 1], kk[u].y, d[1]);
              d[2] = fmaf(q[4 * u + 2], kk[u].z, d[2]); d[3] = fmaf(q[4 * u + 3], kk[u].w, d[3]);
            }
            PD[(half * 128 + r) * 4 + j] = (d[0] + d[1]) + (d[2] + d[3]);
          }
          pf.stop(sec[2]);
          epi_bar();
          pf.stop(sec[3]);
          float w[V];
          float mx = -1e30f;
#pragma unroll
          for (int j = 0; j < V; ++j) {
            w[j] = (PD[r * 4 + j] + PD[(128 + r) * 4 + j]) * 0.125f;      // dim_head ** -0.5
            mx = fmaxf(mx, w[j]);
          }
          float den = 0.f;
#pragma unroll
          for (int j = 0; j < V; ++j) { w[j] = __expf(w[j] - mx); den += w[j]; }
          const float inv = 1.0f / den;
          // o = sum_j softmax_j * v_j over this thread's 32 dims, accumulated as half2 (3 terms;
          // the result is rounded to bf16 for the out-projection anyway)
          __half2 oh[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) oh[i] = __float2half2_rn(0.f);
#pragma unroll
          for (int j = 0; j < V; ++j) {
            const int rj = p0 + j;
            const __half2 wj = __float2half2_rn(w[j] * inv);
            const uint8_t* src = VX + rj * 128;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint4 pkv = *reinterpret_cast<const uint4*>(src + (((4 * half + u) ^ (rj & 7)) << 4));
              const __half2* h2 = reinterpret_cast<const __half2*>(&pkv);
#pragma unroll
              for (int i = 0; i < 4; ++i) oh[4 * u + i] = __hfma2(wj, h2[i], oh[4 * u + i]);
            }
          }
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) { const float2 f = __half22float2(oh[i]); pk[i] = pack_bf16x2(f.x, f.y); }
          tmem_st_u16(tl + kT_ColO + 16 * half, pk);
          tmem_st_wait();
          pf.stop(sec[4]);
          hand_over();
        }
        {
          // ---- x (+ deferred biases) -> LN2 -> YT
          wait_d();
          pf.start();
          float x[80];
          load_x80(tl + kT_ColX + 80 * half, x);
          ln_to_tmem(x, fp + 800, fp + 480, fp + 640, LS, r, half, tl);
          pf.stop(sec[5]);
          hand_over();
        }
        {
          // ---- FF hidden: GELU(acc + b1) -> bf16 operand (K = 128 -> 64 packed columns)
          wait_d();
          pf.start();
          float t[64];
          tmem_ld_x32(tl + kT_ColR + 64 * half, *reinterpret_cast<float(*)[32]>(&t[0]));
          tmem_ld_x32(tl + kT_ColR + 64 * half + 32, *reinterpret_cast<float(*)[32]>(&t[32]));
          tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(fp + 960 + 64 * half);
          uint32_t pk[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 bb = b4[i];
            pk[2 * i] = pack_bf16x2(gelu_erf(t[4 * i] + bb.x), gelu_erf(t[4 * i + 1] + bb.y));
            pk[2 * i + 1] = pack_bf16x2(gelu_erf(t[4 * i + 2] + bb.z), gelu_erf(t[4 * i + 3] + bb.w));
          }
          tmem_st_u16(tl + kT_ColO + 32 * half, pk);
          tmem_st_u16(tl + kT_ColO + 32 * half + 16, pk + 16);
          tmem_st_wait();
          pf.stop(sec[6]);
          hand_over();
        }
        {
          wait_d();
          pf.start();
          float x[80];
          load_x80(tl + kT_ColX + 80 * half, x);
          if (l == 0) {
            const float* f1 = FP + kTLayerFloats;        // layer 1: LN1 on x + pend_in
            ln_to_tmem(x, f1 + 320, f1, f1 + 160, LS, r, half, tl);
            pf.stop(sec[7]);
            hand_over();
          } else if (valid && tok < 2) {
            // ---- output tokens 0 (density branch) and 1 (colour branch), lib/skinnning_batch.py:441-442
            const float4* p4 = reinterpret_cast<const float4*>(FP + 2 * kTLayerFloats + 80 * half);
            uint4* dst = reinterpret_cast<uint4*>((tok == 0 ? a.tok0 : a.tok1) + pnt * kTokLd + 80 * half);
#pragma unroll
            for (int c = 0; c < 10; ++c) {
              const float4 pa = p4[2 * c], pb = p4[2 * c + 1];
              dst[c] = make_uint4(pack_bf16x2(x[8 * c] + pa.x, x[8 * c + 1] + pa.y), pack_bf16x2(x[8 * c + 2] + pa.z, x[8 * c + 3] + pa.w),
                                  pack_bf16x2(x[8 * c + 4] + pb.x, x[8 * c + 5] + pb.y), pack_bf16x2(x[8 * c + 6] + pb.z, x[8 * c + 7] + pb.w));
            }
          }
        }
      }
      ++n_tiles;
    }
    if (pf.on) {
      atomicAdd(&g_prof[0][0], (unsigned long long)acc_d);
      atomicAdd(&g_prof[0][1], (unsigned long long)(clock64() - t_begin - acc_d - acc_tl));
      atomicAdd(&g_prof[0][2], (unsigned long long)acc_tl);
      atomicAdd(&g_prof[0][7], (unsigned long long)n_tiles);
      for (int k = 0; k < 8; ++k) atomicAdd(&g_prof[0][8 + k], (unsigned long long)sec[k]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kC > 1) cluster_sync_all();      // no CTA may exit while a peer can still multicast into it
  if (warp == kMmaWarp) tmem_dealloc(tm, 512);
}

// ------------------------------------------------------------------------------------------
// M: canonical NeRF MLP.  Tile = 128 points.  TMEM columns:
//   ACC [0,256)    accumulator
//   HT  [256,384)  hidden activations, bf16 x2 per column (K = 256)
//   XT  [384,496)  x = [tok0 155 | 0 x5 | PE6(xc) 39 | 0 x25] (K = 224, 14 K-steps);
//                  reused for tok1 (K = 160) after layer 5
// Shared memory holds only the weight ring (6 x 32 KB) and the fp32 parameter vectors.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kM_ColAcc = 0, kM_ColH = 256, kM_ColX = 384;
constexpr uint32_t kM_RING = 0;
constexpr int kM_Slots = 6;
constexpr uint32_t kM_SlotBytes = 32768;
constexpr uint32_t kM_FP = kM_RING + kM_Slots * kM_SlotBytes;
constexpr uint32_t kM_PART = kM_FP + ((kMFloats * 4 + 15) / 16) * 16;   // float[128][4]: alpha / rgb partials of half 1
constexpr uint32_t kM_PIPE = kM_PART + 128 * 16;
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
  int prof;
};

// Hidden-layer epilogue of one thread: 128 accumulator columns (its half of the row) ->
// (+bias, ReLU) -> bf16 x2 -> HT.  Everything compile-time so the inner loops are branch-free.
// Returns this half's sum_j act_j * w_alpha_j when kAlpha.
template <bool kRelu, bool kAlpha>
__device__ __forceinline__ float epi_hidden(uint32_t tl, int half, const float* __restrict__ b,
                                            const float* __restrict__ w_alpha) {
  float acc4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int cb = 0; cb < 2; ++cb) {
    const int c0 = 128 * half + 64 * cb;
    float t[64];
    tmem_ld_x32(tl + kM_ColAcc + c0, *reinterpret_cast<float(*)[32]>(&t[0]));
    tmem_ld_x32(tl + kM_ColAcc + c0 + 32, *reinterpret_cast<float(*)[32]>(&t[32]));
    tmem_ld_wait();
    const float4* b4 = reinterpret_cast<const float4*>(b + c0);
    const float4* w4 = reinterpret_cast<const float4*>(w_alpha + c0);
    uint32_t pk[32];
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
      pk[2 * j] = pack_bf16x2(v0, v1);
      pk[2 * j + 1] = pack_bf16x2(v2, v3);
    }
    tmem_st_u32(tl + kM_ColH + c0 / 2, pk);
  }
  tmem_st_wait();
  return (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]);
}

// 80 packed columns of a bf16 token row (160 values), split between the two halves of the row
__device__ __forceinline__ void token_to_tmem(const __nv_bfloat16* row, bool valid, int half, uint32_t taddr) {
  uint32_t w[40];
  const uint4* src = reinterpret_cast<const uint4*>(row) + 10 * half;     // 80 bf16 = 10 x 16 bytes per half
#pragma unroll
  for (int u = 0; u < 10; ++u) {
    const uint4 t = valid ? __ldg(src + u) : make_uint4(0, 0, 0, 0);
    w[4 * u] = t.x; w[4 * u + 1] = t.y; w[4 * u + 2] = t.z; w[4 * u + 3] = t.w;
  }
#pragma unroll
  for (int i = 0; i < 5; ++i) tmem_st_x8(taddr + 40 * half + 8 * i, *reinterpret_cast<const uint32_t(*)[8]>(&w[8 * i]));
}

// 32 elements (16 packed words) of the 39-wide positional code [x, sin(f0 x), cos(f0 x), ...]
// (run_nerf_helpers.py:337-353; cos as sin(. + fl(pi/2))); elements >= 39 are zero padding.
template <int kHalf>
__device__ __forceinline__ void pe_words(const float (&xc)[3], uint32_t (&pk)[16]) {
#pragma unroll
  for (int w2 = 0; w2 < 16; ++w2) {
    float v[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      constexpr int kBase = 32 * kHalf;
      const int e = kBase + 2 * w2 + q;            // compile-time after unrolling
      float val = 0.f;
      if (e < 3) val = xc[e];
      else if (e < 39) {
        const int k = (e - 3) / 6, ch = (e - 3) % 3;
        const bool is_cos = ((e - 3) % 6) >= 3;
        val = sinf(fmaf(xc[ch], 3.14159265358979323846f * (float)(1 << k), is_cos ? 1.57079632679489661923f : 0.0f));
      }
      v[q] = val;
    }
    pk[w2] = pack_bf16x2(v[0], v[1]);
  }
}

template <int kC>
__global__ void __launch_bounds__(kTcThreads, 1) mlp_tc_kernel(const MArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Pipe* pipe = reinterpret_cast<Pipe*>(smem + kM_PIPE);
  float* FP = reinterpret_cast<float*>(smem + kM_FP);
  float* PART = reinterpret_cast<float*>(smem + kM_PART);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (a.count + 127) / 128;
  const uint32_t crank = (kC > 1) ? cluster_ctarank() : 0u;
  const int64_t ncl = gridDim.x / kC, cid = blockIdx.x / kC;

  if (tid == 0) {
    for (int i = 0; i < kM_Slots; ++i) { mbar_init(&pipe->full[i], 1); mbar_init(&pipe->empty[i], kC); }
    mbar_init(&pipe->a_bar, kEpiThreads);
    mbar_init(&pipe->d_bar, 1);
    mbar_fence_init();
  }
  if (warp == kMmaWarp) { tmem_alloc(&pipe->tmem_base, 512); tmem_relinquish(); }
  {
    const float* src = reinterpret_cast<const float*>(a.blob + kFloatOff) + kTFloats;
    for (int i = tid; i < kMFloats; i += kTcThreads) FP[i] = src[i];
  }
  tc_fence_before();
  __syncthreads();
  if (kC > 1) cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = pipe->tmem_base;

  if (warp == kProdWarp) {
    if (lane == 0) {
      Producer<kM_Slots, kM_SlotBytes, kC> P{pipe, smem + kM_RING, crank};
      P.pf = Prof{a.prof != 0, 0};
      for (int64_t tbase = cid * kC; tbase < ntiles; tbase += ncl * kC) {
        const uint8_t* src = a.blob + kTBytes;
        for (int i = 0; i < 40; ++i) P.push(src, 32768);      // L0..L7, feature: 256-row chunks
        for (int i = 0; i < 7; ++i) P.push(src, 16384);       // views: 128-row chunks
      }
      if (a.prof) atomicAdd(&g_prof[1][6], (unsigned long long)P.acc_e);
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0) {
      Consumer<kM_Slots, kM_SlotBytes, kC> Cn{pipe, smem_u32(smem + kM_RING)};
      Cn.pf = Prof{a.prof != 0, 0};
      const long long t_begin = clock64();
      // A (TMEM at column acol, `ksteps` valid K=16 steps out of `chunks` weight chunks) -> ACC
      auto gemm = [&](uint32_t acol, int ksteps, int chunks, int N, bool accumulate) {
        const uint32_t idesc = instr_desc_bf16(N);
        for (int c = 0; c < chunks; ++c) {
          const uint32_t b0 = Cn.slot_wait();
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const int ks = 4 * c + k4;
            if (ks < ksteps)
              mma_bf16_ts(tm + kM_ColAcc, tm + acol + ks * 8, smem_desc_sw128(b0 + k4 * 32), idesc, (accumulate || ks > 0) ? 1u : 0u);
          }
          Cn.slot_free();
        }
      };
      for (int64_t tbase = cid * kC; tbase < ntiles; tbase += ncl * kC) {
        Cn.wait_a(); gemm(kM_ColX, 14, 4, 256, false); Cn.done();                                   // L0
        for (int L = 1; L < 5; ++L) { Cn.wait_a(); gemm(kM_ColH, 16, 4, 256, false); Cn.done(); }   // L1..L4
        Cn.wait_a(); gemm(kM_ColX, 14, 4, 256, false); gemm(kM_ColH, 16, 4, 256, true); Cn.done();  // L5: [x | h]
        for (int L = 6; L < 8; ++L) { Cn.wait_a(); gemm(kM_ColH, 16, 4, 256, false); Cn.done(); }   // L6, L7
        Cn.wait_a(); gemm(kM_ColH, 16, 4, 256, false); Cn.done();                                   // feature
        Cn.wait_a(); gemm(kM_ColX, 10, 3, 128, false); gemm(kM_ColH, 16, 4, 128, true); Cn.done();  // views: [tok1 | feature]
      }
      if (a.prof) {
        atomicAdd(&g_prof[1][3], (unsigned long long)Cn.acc_a);
        atomicAdd(&g_prof[1][4], (unsigned long long)Cn.acc_w);
        atomicAdd(&g_prof[1][5], (unsigned long long)(clock64() - t_begin));
      }
    }
  } else {
    const int r = tid & 127, half = tid >> 7;
    const uint32_t tl = tm + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t g = 0;
    Prof pf{a.prof != 0 && tid == 0, 0};
    long long acc_d = 0, acc_tl = 0, n_tiles = 0;
    const long long t_begin = clock64();
    auto hand_over = [&]() { tc_fence_before(); mbar_arrive(&pipe->a_bar); };
    auto wait_d = [&]() { pf.start(); mbar_wait(&pipe->d_bar, g & 1); pf.stop(acc_d); ++g; tc_fence_after(); };
    const float* bias = FP;                       // 8 x 256
    const float* w_alpha = FP + 2048;
    const float* b_feat = FP + 2304;
    const float* b_views = FP + 2560;
    const float* w_rgb = FP + 2688;               // 3 x 128
    const float* b_tail = FP + 3072;              // b_alpha, b_rgb[3]

    for (int64_t tbase = cid * kC; tbase < ntiles; tbase += ncl * kC) {
      const int64_t tile = tbase + crank;
      const int64_t i = tile * 128 + r;
      const bool valid = i < a.count;
      {
        // ---- tile load: XT = [tok0 | 0 | PE6(xc) | 0]; this thread: 40 token columns + 16 PE columns
        pf.start();
        token_to_tmem(a.tok0 + (valid ? i : 0) * kTokLd, valid, half, tl + kM_ColX);
        float xc[3] = {0.f, 0.f, 0.f};
        if (valid) { xc[0] = a.xc[3 * i]; xc[1] = a.xc[3 * i + 1]; xc[2] = a.xc[3 * i + 2]; }
        uint32_t pk[16];
        if (half == 0) pe_words<0>(xc, pk); else pe_words<1>(xc, pk);
        tmem_st_u16(tl + kM_ColX + 80 + 16 * half, pk);
        tmem_st_wait();
        pf.stop(acc_tl);
        hand_over();
      }
      float alpha = 0.f;
      for (int L = 0; L < 9; ++L) {          // L0..L7 (ReLU) and 8 = feature (no activation)
        wait_d();
        if (L == 5)                          // x is dead after layer 5: XT <- tok1 for the views layer
          token_to_tmem(a.tok1 + (valid ? i : 0) * kTokLd, valid, half, tl + kM_ColX);
        // the layer kind is warp-uniform: pick a fully specialised epilogue (no per-element branches)
        if (L == 7) alpha = epi_hidden<true, true>(tl, half, bias + 256 * 7, w_alpha);   // + alpha_linear on fp32 acts
        else if (L == 8) epi_hidden<false, false>(tl, half, b_feat, nullptr);
        else epi_hidden<true, false>(tl, half, bias + 256 * L, nullptr);
        hand_over();
      }
      {
        // ---- views layer epilogue: relu -> rgb_linear on CUDA cores -> raw[pid] = (rgb, alpha)
        wait_d();
        float c0[2] = {0.f, 0.f}, c1[2] = {0.f, 0.f}, c2[2] = {0.f, 0.f};
        float t[64];
        tmem_ld_x32(tl + kM_ColAcc + 64 * half, *reinterpret_cast<float(*)[32]>(&t[0]));
        tmem_ld_x32(tl + kM_ColAcc + 64 * half + 32, *reinterpret_cast<float(*)[32]>(&t[32]));
        tmem_ld_wait();
        const float4* b4 = reinterpret_cast<const float4*>(b_views + 64 * half);
        const float4* r0 = reinterpret_cast<const float4*>(w_rgb + 64 * half);
        const float4* r1 = reinterpret_cast<const float4*>(w_rgb + 128 + 64 * half);
        const float4* r2 = reinterpret_cast<const float4*>(w_rgb + 256 + 64 * half);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 bb = b4[j], w0 = r0[j], w1 = r1[j], w2 = r2[j];
          const float v0 = fmaxf(t[4 * j] + bb.x, 0.f), v1 = fmaxf(t[4 * j + 1] + bb.y, 0.f);
          const float v2 = fmaxf(t[4 * j + 2] + bb.z, 0.f), v3 = fmaxf(t[4 * j + 3] + bb.w, 0.f);
          c0[0] = fmaf(v0, w0.x, c0[0]); c0[1] = fmaf(v1, w0.y, c0[1]); c0[0] = fmaf(v2, w0.z, c0[0]); c0[1] = fmaf(v3, w0.w, c0[1]);
          c1[0] = fmaf(v0, w1.x, c1[0]); c1[1] = fmaf(v1, w1.y, c1[1]); c1[0] = fmaf(v2, w1.z, c1[0]); c1[1] = fmaf(v3, w1.w, c1[1]);
          c2[0] = fmaf(v0, w2.x, c2[0]); c2[1] = fmaf(v1, w2.y, c2[1]); c2[0] = fmaf(v2, w2.z, c2[0]); c2[1] = fmaf(v3, w2.w, c2[1]);
        }
        if (half == 1)
          *reinterpret_cast<float4*>(PART + 4 * r) = make_float4(c0[0] + c0[1], c1[0] + c1[1], c2[0] + c2[1], alpha);
        epi_bar();
        if (half == 0 && valid) {
          const float4 o = *reinterpret_cast<const float4*>(PART + 4 * r);
          reinterpret_cast<float4*>(a.raw)[a.act_pid[i]] =
              make_float4(c0[0] + c0[1] + o.x + b_tail[1], c1[0] + c1[1] + o.y + b_tail[2],
                          c2[0] + c2[1] + o.z + b_tail[3], alpha + o.w + b_tail[0]);
        }
        epi_bar();      // PART is rewritten by the next tile
      }
      ++n_tiles;
    }
    if (pf.on) {
      atomicAdd(&g_prof[1][0], (unsigned long long)acc_d);
      atomicAdd(&g_prof[1][1], (unsigned long long)(clock64() - t_begin - acc_d - acc_tl));
      atomicAdd(&g_prof[1][2], (unsigned long long)acc_tl);
      atomicAdd(&g_prof[1][7], (unsigned long long)n_tiles);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kC > 1) cluster_sync_all();
  if (warp == kMmaWarp) tmem_dealloc(tm, 512);
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

// Same tile with the A operand written to TENSOR MEMORY by the epilogue threads (tcgen05.st,
// two bf16 per 32-bit column) and consumed by the TS form of tcgen05.mma.
__global__ void __launch_bounds__(128, 1)
selftest_umma_ts_kernel(const uint16_t* __restrict__ a, const uint8_t* __restrict__ b_packed, float* __restrict__ d,
                        int N, int K, int acol) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_b, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int chunks = K / 64;
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  if (tid == 0) { mbar_init(&bar_b, 1); mbar_init(&bar_mma, 1); mbar_fence_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16);
  for (int k0 = 0; k0 < K; k0 += 16) {          // 16 bf16 = 8 packed columns per store
    const uint4 lo = *reinterpret_cast<const uint4*>(a + (size_t)tid * K + k0);
    const uint4 hi = *reinterpret_cast<const uint4*>(a + (size_t)tid * K + k0 + 8);
    const uint32_t v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    tmem_st_x8(tl + acol + k0 / 2, v);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)chunks * (uint32_t)N * 128u;
    mbar_arrive_expect_tx(&bar_b, bytes);
    bulk_g2s(smem, b_packed, bytes, &bar_b);
    mbar_wait(&bar_b, 0);
    tc_fence_after();
    const uint32_t idesc = instr_desc_bf16(N);
    for (int c = 0; c < chunks; ++c) {
      const uint32_t b0 = smem_u32(smem + (size_t)c * N * 128);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4)
        mma_bf16_ts(tm, tm + acol + (c * 4 + k4) * 8, smem_desc_sw128(b0 + k4 * 32), idesc, (c | k4) ? 1u : 0u);
    }
    mma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c = 0; c < N; c += 16) {
    float v[16];
    tmem_ld_x16(tl + (uint32_t)c, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) d[(size_t)tid * N + c + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

}  // namespace mps

extern "C" int mpsnerf_selftest_umma_ts(const uint16_t* a, const uint8_t* b_packed, float* d, int N, int K,
                                        int acol, void* stream) {
  MPS_REQUIRE(a && b_packed && d);
  MPS_REQUIRE(acol >= N && acol % 8 == 0 && acol + K / 2 <= 512);
  MPS_REQUIRE(K >= 64 && K % 64 == 0 && K <= 512 && N >= 16 && N <= 256 && N % 16 == 0);
  const size_t smem = (size_t)(K / 64) * (size_t)N * 128;
  MPS_REQUIRE(smem <= 200 * 1024);
  MPS_CUDA(cudaFuncSetAttribute(mps::selftest_umma_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mps::selftest_umma_ts_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a, b_packed, d, N, K, acol);
  MPS_LAUNCH_CHECK();
  return MPSNERF_OK;
}

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

  // cluster size of the weight multicast: MPSNERF_CLUSTER = 1 | 2 | 4 (default 2)
  static int cluster = 0;
  if (cluster == 0) {
    const char* e = getenv("MPSNERF_CLUSTER");
    cluster = e ? atoi(e) : 2;
    if (cluster != 1 && cluster != 2 && cluster != 4) cluster = 2;
  }
  static int prof = -1;
  if (prof < 0) { const char* e = getenv("MPSNERF_TC_PROF"); prof = (e && atoi(e)) ? 1 : 0; }
  TArgs ta{tokens, ld, count, n_views, static_cast<const uint8_t*>(packed), tok0, tok1, prof};
  MArgs ma{tok0, tok1, xc, count, static_cast<const uint8_t*>(packed), act_pid + first, raw, prof};
  const int ppt = 128 / n_views;
  const int64_t t_tiles = (count + ppt - 1) / ppt, m_tiles = (count + 127) / 128;
  auto launch = [&](auto kernel, const auto& args, size_t smem_bytes, int64_t tiles, int kc) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    int64_t ctas = (tiles + kc - 1) / kc * kc;                  // whole clusters
    const int64_t cap = (kNumSMs / kc) * kc;
    if (ctas > cap) ctas = cap;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)kc;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args);
  };
#define MPS_T_CASE(V_, C_) if (n_views == V_ && cluster == C_) MPS_CUDA(launch(xformer_tc_kernel<V_, C_>, ta, kT_Smem, t_tiles, C_));
  MPS_T_CASE(2, 1) MPS_T_CASE(2, 2) MPS_T_CASE(2, 4)
  MPS_T_CASE(3, 1) MPS_T_CASE(3, 2) MPS_T_CASE(3, 4)
  MPS_T_CASE(4, 1) MPS_T_CASE(4, 2) MPS_T_CASE(4, 4)
#undef MPS_T_CASE
  if (cluster == 1) MPS_CUDA(launch(mlp_tc_kernel<1>, ma, kM_Smem, m_tiles, 1));
  if (cluster == 2) MPS_CUDA(launch(mlp_tc_kernel<2>, ma, kM_Smem, m_tiles, 2));
  if (cluster == 4) MPS_CUDA(launch(mlp_tc_kernel<4>, ma, kM_Smem, m_tiles, 4));
  return MPSNERF_OK;
}

// Debug: copy (and clear) the in-kernel cycle counters; out = 16 unsigned 64-bit values (T then M).
extern "C" int mpsnerf_debug_read_prof(unsigned long long* host_out) {
  MPS_REQUIRE(host_out != nullptr);
  MPS_CUDA(cudaMemcpyFromSymbol(host_out, mps::g_prof, sizeof(unsigned long long) * 32));
  unsigned long long zero[32] = {0};
  MPS_CUDA(cudaMemcpyToSymbol(mps::g_prof, zero, sizeof(zero)));
  return MPSNERF_OK;
}
