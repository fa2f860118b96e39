// K5 (bf16 production path): fused tcgen05 / TMEM kernels for the cross-view transformer ("T")
// and the canonical NeRF MLP ("M").  Restates lib/transformer.py:13-86 and
// lib/skinnning_batch.py:438-473 with bf16 operands and fp32 accumulation.
//
// Structure (one persistent CTA per SM, 640 threads = 5 warpgroups, setmaxnreg 104 / 40):
//   warps 0-15 epilogue, 512 threads: thread (r = tid & 127, q = tid >> 7) owns a quarter of the
//              columns of row r (TMEM lane r; warps w, w+4, w+8, w+12 share lane quarter w & 3).  They
//              turn accumulators into the next A operand (bias / ReLU / GELU / LayerNorm / attention)
//              and store it, bf16-packed, back into TENSOR MEMORY (tcgen05.st);
//   warp 16    weight producer: streams the pre-swizzled weight chunks with bulk async copies (TMA
//              engine) into an mbarrier ring, in consumption order (multicast across the kC CTAs of
//              a cluster);
//   warp 17    MMA issuer: TS-form tcgen05.mma (A from TMEM, B from smem, M = 128).  Both run fully
//              converged with one elected lane issuing, so that operands stay in uniform registers.
// Activations never leave the SM between layers and never touch shared memory: shared memory
// holds only the weight ring, which is what hides the L2 latency.  MMA <-> epilogue hand-over is
// by mbarriers (a_bar: "A operand ready", 512 arrivals; d_bar: "accumulator ready", tcgen05.commit).
// Measured building blocks and the design rules derived from them: DESIGN.md section 5.1,
// tools/ubench_tc.cu.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "umma.cuh"

namespace mps {
using namespace umma;


// ---- blob layout (must match mps-nerf_b200/pack.py)
constexpr uint32_t kQkvChunk = 192 * 128, kWoChunk = 160 * 128, kW1Chunk = 128 * 128, kW2Chunk = 160 * 128;
constexpr uint32_t kTLayerBytes = 12 * kQkvChunk + 4 * kWoChunk + 3 * kW1Chunk + 2 * kW2Chunk;
constexpr uint32_t kTBytes = 2 * kTLayerBytes;
constexpr uint32_t kMBytes = 44 * 32768;
constexpr uint32_t kFloatOff = kTBytes + kMBytes;
constexpr int kTLayerFloats = 1088, kTFloats = 2 * kTLayerFloats + 160, kMFloats = 8 * 256 + 256 + 256 + 128 + 384 + 4;
constexpr size_t kBlobBytes = (size_t)kFloatOff + 4 * (size_t)(kTFloats + kMFloats);
static_assert(kTLayerBytes == 466944 && kMBytes == 1441792, "blob layout drifted from pack.py");

constexpr int kTokLd = 160;   // bf16 row stride of tok0 / tok1 handed from T to M

// Optional in-kernel cycle accounting (MPSNERF_TC_PROF=1): per kernel 8 counters summed over CTAs:
// 0 epilogue wait-for-MMA, 1 epilogue work, 2 tile load, 3 MMA wait-for-A, 4 MMA wait-for-weights,
// 5 MMA thread total, 6 producer wait-for-free-slot, 7 tiles.  Read with mpsnerf_debug_read_prof().
__device__ unsigned long long g_prof[2][16];   // [8..15]: T epilogue sections, see tools/step.py
// Optional event trace of CTA 0's second tile (MPSNERF_TC_PROF=2): [0,256) MMA thread, [256,512) epilogue thread 0;
// each entry = (tag << 48) | (clock64 & 0xffffffffffff).  Read with mpsnerf_debug_read_trace().
__device__ unsigned long long g_trace[512];
struct Trace {
  bool on; int n; int base;
  __device__ __forceinline__ void ev(int tag) { if (on && n < 256) { g_trace[base + n] = ((unsigned long long)tag << 48) | ((unsigned long long)clock64() & 0xffffffffffffull); ++n; } }
};

struct Prof {
  bool on;
  long long t;
  __device__ __forceinline__ void start() { if (on) t = clock64(); }
  __device__ __forceinline__ void stop(long long& acc) { if (on) { const long long n = clock64(); acc += n - t; t = n; } }
};

struct Pipe {            // barriers of one CTA (in dynamic smem)
  uint64_t full[16];
  uint64_t empty[16];
  uint64_t a_bar[2];    // "A operand ready" (M kernel: one per K-half)
  uint64_t d_bar[2];    // "accumulator ready" (M kernel: one per N-half)
  uint64_t k_bar;       // M kernel: "[h1: K0] done, HT K-half 0 may be overwritten"
  uint64_t r_bar;       // T kernel: "scratch accumulator R has been copied out" (early release)
  uint64_t q_bar[2];    // T kernel: "q|k|v of this team's next head is in R" (tcgen05.commit)
  uint64_t o_bar[2];    // T kernel: "this team's attention output is in its half of OT" (256 arrivals)
  uint64_t w_bar[2];    // T kernel: "the out-projection of this team's first head has read its half of OT" (commit)
  uint64_t h_bar[2];    // T kernel: "half j (64 columns) of the FF hidden layer is in R" (commit)
  uint64_t g_bar[2];    // T kernel: "GELU of hidden half j is in OT" (256 arrivals: the threads with q >> 1 == j)
  uint32_t tmem_base;
  uint32_t pad;
};

// GELU on two values at once in half precision: 0.5 x (1 + tanh(x (c0 + c1 x^2 + c2 x^4))) with the hardware
// tanh.approx.f16x2 (one MUFU op per pair).  The three coefficients are a minimax fit of the erf form
// (max |error| 2.5e-5 on the real line, 20x tighter than the usual two-term tanh form); the f16 arithmetic
// of the tanh argument adds < 4e-4 absolute, below the bf16 rounding (4e-3) of the operand this feeds.  x^2 is clamped at 36:
// beyond |x| = 6 the polynomial would turn around, tanh is +-1 there anyway.  Returns bf16 x2.
__device__ __forceinline__ uint32_t gelu_pair_bf16(float x0, float x1) {
  const __half2 x = __floats2half2_rn(x0, x1);
  const __half2 x2 = __hmin2(__hmul2(x, x), __float2half2_rn(36.0f));
  __half2 p = __hfma2(x2, __float2half2_rn(-3.51516790e-04f), __float2half2_rn(3.70056460e-02f));
  p = __hfma2(x2, p, __float2half2_rn(7.97507884e-01f));
  const __half2 u = __hmul2(x, p);
  uint32_t ti;
  const uint32_t ui = *reinterpret_cast<const uint32_t*>(&u);
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(ti) : "r"(ui));
  const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&ti));
  const float h0 = 0.5f * x0, h1 = 0.5f * x1;          // the linear part stays in fp32
  return pack_bf16x2(fmaf(h0, t.x, h0), fmaf(h1, t.y, h1));
}

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 t = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// Weight ring: the whole warp runs the control flow, one elected lane issues.
template <int kSlots, uint32_t kSlotBytes, int kC>
struct RingProducer {
  Pipe* pipe;
  uint8_t* ring;
  uint32_t crank;
  uint32_t slot = 0, phase = 0;
  Prof pf{false, 0};
  long long acc_e = 0;
  __device__ __forceinline__ void push(const uint8_t*& src, uint32_t bytes) {
    pf.start();
    mbar_wait(&pipe->empty[slot], phase ^ 1);
    pf.stop(acc_e);
    if (elect_one()) {
      mbar_arrive_expect_tx(&pipe->full[slot], bytes);          // the whole chunk lands here (kC slices)
      if (kC == 1) {
        bulk_g2s(ring + slot * kSlotBytes, src, bytes, &pipe->full[slot]);
      } else {
        const uint32_t slice = bytes / kC;
        bulk_g2s_mc(ring + slot * kSlotBytes + crank * slice, src + crank * slice, slice, &pipe->full[slot],
                    (uint16_t)((1u << kC) - 1u));
      }
    }
    __syncwarp();
    src += bytes;
    if (++slot == kSlots) { slot = 0; phase ^= 1; }
  }
};

template <int kSlots, uint32_t kSlotBytes, int kC>
struct RingConsumer {
  Pipe* pipe;
  uint32_t ring_addr;
  uint32_t slot = 0, phase = 0;
  Prof pf{false, 0};
  long long acc_w = 0;
  __device__ __forceinline__ uint32_t acquire() {       // all lanes
    pf.start();
    mbar_wait(&pipe->full[slot], phase);
    pf.stop(acc_w);
    tc_fence_after();
    return ring_addr + slot * kSlotBytes;
  }
  __device__ __forceinline__ void release_elected() {   // elected lane, after its MMAs
    if (kC == 1) mma_commit(&pipe->empty[slot]);
    else mma_commit_mc(&pipe->empty[slot], (uint16_t)((1u << kC) - 1u));
  }
  __device__ __forceinline__ void advance() { if (++slot == kSlots) { slot = 0; phase ^= 1; } }
};

// Both fused kernels run 640 threads = 5 warpgroups: four epilogue warpgroups (16 warps, four
// threads per row: thread (r = tid & 127, q = tid >> 7) owns a quarter of the row's columns; raised to
// 104 registers) and one holding the producer warp (16), the MMA warp (17) and two idle warps (lowered
// to 40), via setmaxnreg.  16 epilogue warps instead of 8: the epilogues are latency bound (issue
// slots 30 % busy with 2 warps per scheduler), and per-thread state halves.
constexpr int kFThreads = 640, kFEpiThreads = 512, kFProdWarp = 16, kFMmaWarp = 17;
template <int kRegs> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }
__device__ __forceinline__ void f_epi_bar() { named_bar_sync(1, kFEpiThreads); }
// the four threads of a row live in warps w, w + 4, w + 8, w + 12: barrier among those 128 threads
__device__ __forceinline__ void f_row_bar(int warp) { named_bar_sync(2 + (warp & 3), 128); }

__device__ __forceinline__ void tmem_st_x4(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3])
               : "memory");
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

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
// attention exchange area, one set per team (the two teams work on different heads at the same time):
constexpr uint32_t kT_KX = 0;                      // k of the team's head, fp16 [128][64], rows of 128 B, unit-swizzled
constexpr uint32_t kT_VX = 16384;                  // v of the team's head, fp16, same layout
constexpr uint32_t kT_PD = 32768;                  // partial q.k dots  float[2][128][4]
constexpr uint32_t kT_TeamBytes = 36864;
constexpr uint32_t kT_LS = 2 * kT_TeamBytes;       // LayerNorm partial sums float[2][4][128]
constexpr uint32_t kT_RING = kT_LS + 4096;         // 1024-aligned: 2 * 36864 + 4096 = 77824 = 76 * 1024
constexpr int kT_Slots = 6;          // (5 until the LayerNorm vectors left shared memory: the tensor pipe waited ~9 % of a tile for weights)
constexpr int kT_PendFloats = 5 * 160;   // the deferred-bias vectors kept in shared memory: pend_in / pend_mid of both layers, pend_out
constexpr uint32_t kT_SlotBytes = kQkvChunk;
constexpr uint32_t kT_FP = kT_RING + kT_Slots * kT_SlotBytes;
constexpr uint32_t kT_PIPE = kT_FP + ((kT_PendFloats * 4 + 15) / 16) * 16;
constexpr uint32_t kT_Smem = kT_PIPE + sizeof(Pipe);
static_assert(kT_RING % 1024 == 0 && kT_SlotBytes % 1024 == 0, "ring slots must be 1024-byte aligned");
static_assert(kT_Smem <= 232448 - 1024, "T kernel shared memory over budget");

struct TArgs {
  const __half* tokens;  // (count, V, 160) fp16
  int ld;
  int64_t count;
  int V;
  const uint8_t* blob;
  __nv_bfloat16* tok0;   // (count, 160)
  __nv_bfloat16* tok1;
  int prof;
  const int32_t* count_dev;   // optional device-side active count (count is then the slab capacity)
  int64_t first;
  int opt;                    // experiment switches (MPSNERF_T_OPT): bit 0 = partner dots in fp32 instead of packed half2 FMAs
};

// LayerNorm over the 155 real columns of a row whose 160 columns are split between four threads
// (this thread: x[0..40) = columns 40 q ..; real columns: 40 or 35), result bf16-packed -> YT.
// The affine part (gamma, beta) lives in the weights of the consuming GEMM (pack.py); the pad of the operand is exactly
// zero, except the constant 1.0 of column 155.  Two passes (mean, then centred variance) like the reference's
// nn.LayerNorm; the partial sums of a row meet in LS.
__device__ __forceinline__ void ln_to_tmem(float (&x)[40], const float* __restrict__ pend,
                                           const float* __restrict__ g, const float* __restrict__ b, float* LS,
                                           int r, int q, int warp, uint32_t tl) {
  const int nreal = (q == 3) ? 35 : 40;
  if (pend) {
    const float4* p4 = reinterpret_cast<const float4*>(pend + 40 * q);
#pragma unroll
    for (int c = 0; c < 10; ++c) {
      const float4 t = p4[c];
      x[4 * c] += t.x; x[4 * c + 1] += t.y; x[4 * c + 2] += t.z; x[4 * c + 3] += t.w;
    }
  }
  float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 40; ++c) if (c < nreal) s[c & 3] += x[c];
  LS[q * 128 + r] = (s[0] + s[1]) + (s[2] + s[3]);
  f_row_bar(warp);
  const float mean = ((LS[r] + LS[128 + r]) + (LS[256 + r] + LS[384 + r])) * (1.0f / 155.0f);
  float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 40; ++c) if (c < nreal) { const float d = x[c] - mean; v[c & 3] = fmaf(d, d, v[c & 3]); }
  LS[512 + q * 128 + r] = (v[0] + v[1]) + (v[2] + v[3]);
  f_row_bar(warp);
  const float rstd = rsqrtf(((LS[512 + r] + LS[640 + r]) + (LS[768 + r] + LS[896 + r])) * (1.0f / 155.0f) + 1e-5f);
  // gamma / beta are folded into the consuming GEMM's weights (pack.py): the operand is xhat = (x - mean) rstd, exactly
  // zero in the pad columns except a constant 1.0 in column 155 (K row that carries W beta for the q|k|v projection).
  (void)g; (void)b;
  const float nm = -mean * rstd;
  uint32_t pk[20];
#pragma unroll
  for (int c = 0; c < 10; ++c) {
    float y[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int col = 4 * c + e;
      y[e] = (col < nreal) ? fmaf(x[col], rstd, nm) : ((q == 3 && col == 35) ? 1.0f : 0.0f);
    }
    pk[2 * c] = pack_bf16x2(y[0], y[1]);
    pk[2 * c + 1] = pack_bf16x2(y[2], y[3]);
  }
#pragma unroll
  for (int i = 0; i < 5; ++i) tmem_st_x4(tl + kT_ColY + 20 * q + 4 * i, &pk[4 * i]);   // every TMEM access aligned to its own width
  tmem_st_wait();
}

__device__ __forceinline__ void load_x40(uint32_t taddr, float (&x)[40]) {
#pragma unroll
  for (int i = 0; i < 5; ++i) tmem_ld_x8(taddr + 8 * i, &x[8 * i]);
  tmem_ld_wait();
}

// kC = CTAs per cluster.  The kC CTAs of a cluster run the same program on neighbouring tiles and
// share the weight stream: CTA j loads slice j of every chunk and multicasts it to all of them, so
// each chunk crosses L2 -> SM once per cluster.  A ring slot is reused only after the MMA warps of
// *all* kC CTAs have released it (empty barriers count kC arrivals, delivered by multicast commits).
template <int kV, int kC, bool kProf>
__global__ void __launch_bounds__(kFThreads, 1) xformer_tc_kernel(const TArgs a_in) {
  TArgs a = a_in;
  if (a.count_dev != nullptr) {
    const int64_t n = (int64_t)*a.count_dev - a.first;
    a.count = n < a.count ? (n > 0 ? n : 0) : a.count;
  }
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
    mbar_init(&pipe->a_bar[0], kFEpiThreads);
    mbar_init(&pipe->d_bar[0], 1);
    mbar_init(&pipe->r_bar, kFEpiThreads / 2);
    for (int t = 0; t < 2; ++t) { mbar_init(&pipe->q_bar[t], 1); mbar_init(&pipe->o_bar[t], kFEpiThreads / 2); mbar_init(&pipe->w_bar[t], 1);
                                  mbar_init(&pipe->h_bar[t], 1); mbar_init(&pipe->g_bar[t], kFEpiThreads / 2); }
    mbar_fence_init();
  }
  if (warp == kFMmaWarp) { tmem_alloc(&pipe->tmem_base, 512); tmem_relinquish(); }
  {
    const float* src = reinterpret_cast<const float*>(a.blob + kFloatOff);
    // Only the deferred-bias vectors are needed on chip: the LayerNorm affine and the feed-forward bias live in the
    // packed weights (pack.py).  FP[160 j ..]: j = 0 pend_in(l0), 1 pend_mid(l0), 2 pend_in(l1), 3 pend_mid(l1), 4 pend_out.
    for (int i = tid; i < kT_PendFloats; i += kFThreads) {
      const int j = i / 160, c = i - 160 * j;
      FP[i] = src[j == 4 ? 2 * kTLayerFloats + c : (j >> 1) * kTLayerFloats + ((j & 1) ? 800 : 320) + c];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kC > 1) cluster_sync_all();      // peers' barriers are initialised before any multicast touches them
  tc_fence_after();
  const uint32_t tm = pipe->tmem_base;

  if (warp >= 16) {
    reg_dec<40>();           // fifth warpgroup: producer, MMA issuer, two idle warps
  if (warp == kFProdWarp) {
    // ================= weight producer (whole warp converged, one elected lane issues) =================
    RingProducer<kT_Slots, kT_SlotBytes, kC> P{pipe, smem + kT_RING, crank};
    P.pf = Prof{kProf && a.prof == 1 && lane == 0, 0};
    for (int64_t tbase = cid * kC; tbase < ntiles; tbase += ncl * kC) {   // cluster-uniform trip count
#pragma unroll 1
      for (int l = 0; l < 2; ++l) {
        const uint8_t* src = a.blob + (size_t)l * kTLayerBytes;
        // consumption order of the attention block: qkv_0 qkv_1 qkv_2 Wo_0 qkv_3 Wo_1 Wo_2 Wo_3
        for (int c = 0; c < 9; ++c) P.push(src, kQkvChunk);
        P.push(src, kWoChunk);
        for (int c = 0; c < 3; ++c) P.push(src, kQkvChunk);
        for (int c = 0; c < 3; ++c) P.push(src, kWoChunk);
        for (int c = 0; c < 2; ++c) P.push(src, 3 * (kW1Chunk / 2));    // W1 rows 0..63 (3 K-chunks = one slot), then rows 64..127
        for (int c = 0; c < 2; ++c) P.push(src, kW2Chunk);
      }
    }
    if (P.pf.on) atomicAdd(&g_prof[0][6], (unsigned long long)P.acc_e);
  } else if (warp == kFMmaWarp) {
    // ================= MMA issuer (whole warp converged, one elected lane issues) =================
    RingConsumer<kT_Slots, kT_SlotBytes, kC> Cn{pipe, smem_u32(smem + kT_RING)};
    Prof pf{kProf && a.prof == 1 && lane == 0, 0};
    Cn.pf = pf;
    long long acc_a = 0;
    uint32_t g = 0, rr = 0, po[2] = {0, 0}, pg = 0;
    const long long t_begin = clock64();
    auto wait_a = [&]() { pf.start(); mbar_wait(&pipe->a_bar[0], g & 1); pf.stop(acc_a); tc_fence_after(); };
    auto done = [&]() { if (elect_one()) mma_commit(&pipe->d_bar[0]); __syncwarp(); ++g; };
    // A (TMEM, kSteps K=16 steps starting at column acol) x weight chunks with kN rows -> D column dcol.
    // One ring slot per 64 K columns: the barrier probe of a chunk hides behind the queued MMAs of the previous one.
    auto gemm = [&](auto ksteps_c, auto n_c, uint32_t dcol, uint32_t acol, bool accumulate, auto f16_c) {
      constexpr int kSteps = decltype(ksteps_c)::value, kN = decltype(n_c)::value;
      // operands bf16 x bf16, or fp16 x fp16 for the out-projection (its A operand, the attention output, is produced
      // in fp16 and handed over without a repack; its weights are packed as fp16, pack.py)
      constexpr uint32_t idesc = decltype(f16_c)::value ? instr_desc_f16(kN) : instr_desc_bf16(kN);
#pragma unroll
      for (int k0 = 0; k0 < kSteps; k0 += 4) {
        const uint64_t bdesc = smem_desc_sw128(Cn.acquire());
        if (elect_one()) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            if (k0 + k4 < kSteps)
              mma_bf16_ts(tm + dcol, tm + acol + (k0 + k4) * 8, bdesc + (uint64_t)(k4 * 2), idesc, (accumulate || k0 + k4 > 0) ? 1u : 0u);
          Cn.release_elected();
        }
        __syncwarp();
        Cn.advance();
      }
    };
    // Same, with all K-chunks of the operand in ONE ring slot (chunk stride kN * 128 bytes): the two 64-row halves of
    // W1 are 3 x 8 KB each -- as three slots apiece they filled the 5-slot ring with a quarter of its bytes and the
    // feed-forward burst (W1 + W2 = 8 slots) made the tensor pipe wait for weights (3 K cycles per tile).
    auto gemm_one_slot = [&](auto ksteps_c, auto n_c, uint32_t dcol, uint32_t acol) {
      constexpr int kSteps = decltype(ksteps_c)::value, kN = decltype(n_c)::value;
      constexpr uint32_t idesc = instr_desc_bf16(kN);
      static_assert(((kSteps + 3) / 4) * kN * 128 <= (int)kT_SlotBytes, "operand must fit one slot");
      const uint32_t base = Cn.acquire();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kSteps; ++k) {
          const uint64_t bdesc = smem_desc_sw128(base + (uint32_t)(k >> 2) * (uint32_t)(kN * 128));
          mma_bf16_ts(tm + dcol, tm + acol + k * 8, bdesc + (uint64_t)((k & 3) * 2), idesc, k > 0 ? 1u : 0u);
        }
        Cn.release_elected();
      }
      __syncwarp();
      Cn.advance();
    };
    using BF = std::false_type; using F16 = std::true_type;
    using K10 = std::integral_constant<int, 10>; using K4 = std::integral_constant<int, 4>; using K8 = std::integral_constant<int, 8>;
    using N192 = std::integral_constant<int, 192>; using N160 = std::integral_constant<int, 160>; using N128 = std::integral_constant<int, 128>; using N64 = std::integral_constant<int, 64>;
    for (int64_t tbase = cid * kC; tbase < ntiles; tbase += ncl * kC) {
#pragma unroll 1
      for (int l = 0; l < 2; ++l) {
        // Attention block.  The epilogue warps form two teams (A: heads 0, 2; B: heads 1, 3) that run their
        // heads concurrently, so the issue slots one team leaves empty in its barrier / latency stalls are
        // filled by the other.  R is single: the next head's q|k|v is issued as soon as the team that owns
        // the current one has copied it out (r_bar); each team's output has its own half of OT.
        auto wait_r = [&]() { mbar_wait(&pipe->r_bar, rr & 1); ++rr; tc_fence_after(); };
        auto wait_o = [&](int t) { pf.start(); mbar_wait(&pipe->o_bar[t], po[t] & 1); ++po[t]; pf.stop(acc_a); tc_fence_after(); };
        auto commit_q = [&](int t) { if (elect_one()) mma_commit(&pipe->q_bar[t]); __syncwarp(); };
        auto commit_w = [&](int t) { if (elect_one()) mma_commit(&pipe->w_bar[t]); __syncwarp(); };
        wait_a(); gemm(K10{}, N192{}, kT_ColR, kT_ColY, false, BF{}); commit_q(0);           // q|k|v of head 0 -> team A
        wait_r(); gemm(K10{}, N192{}, kT_ColR, kT_ColY, false, BF{}); commit_q(1);     // head 1 -> team B
        wait_r(); gemm(K10{}, N192{}, kT_ColR, kT_ColY, false, BF{}); commit_q(0);     // head 2 -> team A (waits in R while A finishes head 0)
        wait_o(0); gemm(K4{}, N160{}, kT_ColX, kT_ColO, true, F16{}); commit_w(0);     // x += o_0 Wo_0^T
        wait_r(); gemm(K10{}, N192{}, kT_ColR, kT_ColY, false, BF{}); commit_q(1);     // head 3 -> team B
        wait_o(1); gemm(K4{}, N160{}, kT_ColX, kT_ColO + 32, true, F16{}); commit_w(1); // x += o_1 Wo_1^T
        wait_o(0); gemm(K4{}, N160{}, kT_ColX, kT_ColO, true, F16{});                  // x += o_2 Wo_2^T
        wait_o(1); gemm(K4{}, N160{}, kT_ColX, kT_ColO + 32, true, F16{});             // x += o_3 Wo_3^T
        done();
        // Feed-forward, pipelined in two halves of the hidden layer: hidden columns [0,64) are handed to the
        // epilogue threads that own them (q < 2) while [64,128) is still being computed, and x += gelu(.) W2^T runs
        // over the first half's K = 64 while the other threads are still in their GELU.
        wait_a();
        gemm_one_slot(K10{}, N64{}, kT_ColR, kT_ColY);
        if (elect_one()) mma_commit(&pipe->h_bar[0]);
        __syncwarp();
        gemm_one_slot(K10{}, N64{}, kT_ColR + 64, kT_ColY);
        if (elect_one()) mma_commit(&pipe->h_bar[1]);
        __syncwarp();
        pf.start(); mbar_wait(&pipe->g_bar[0], pg & 1); pf.stop(acc_a); tc_fence_after();
        gemm(K4{}, N160{}, kT_ColX, kT_ColO, true, BF{});
        pf.start(); mbar_wait(&pipe->g_bar[1], pg & 1); pf.stop(acc_a); tc_fence_after();
        ++pg;
        gemm(K4{}, N160{}, kT_ColX, kT_ColO + 32, true, BF{});
        done();
      }
    }
    if (pf.on) {
      atomicAdd(&g_prof[0][3], (unsigned long long)acc_a);
      atomicAdd(&g_prof[0][4], (unsigned long long)Cn.acc_w);
      atomicAdd(&g_prof[0][5], (unsigned long long)(clock64() - t_begin));
    }
  }
  } else {
    // ================= epilogue warps =================
    reg_inc<104>();
    const int r = tid & 127, q = tid >> 7;
    float* LS = reinterpret_cast<float*>(smem + kT_LS);
    const uint32_t tl = tm + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t g = 0, pq = 0, pw = 0, ph = 0;
    Prof pf{kProf && a.prof == 1 && tid == 0, 0};
    long long acc_d = 0, acc_tl = 0, n_tiles = 0;
    long long sec[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // publish, bar1, dots, bar2, softmax+o, LN2, GELU, LN1/final
    const long long t_begin = clock64();
    auto hand_over = [&]() { tc_fence_before(); mbar_arrive(&pipe->a_bar[0]); };
    auto wait_d = [&]() { pf.start(); mbar_wait(&pipe->d_bar[0], g & 1); pf.stop(acc_d); ++g; tc_fence_after(); };
    constexpr int rows = ppt * V;
    const int p0 = (r < rows) ? (r / V) * V : 0;      // first row of this row's point (attention partners)
    // The next tile's tokens (this thread's 40 fp16 columns, 5 x 16 B) are staged asynchronously (LDGSTS) in the
    // attention exchange buffers, which are idle between the last attention block of a tile and the first one of
    // the next ([unit][thread] layout, conflict-free), and picked up into registers when they are needed.
    uint8_t* xstage = smem + tid * 16;
    static_assert(5 * kFEpiThreads * 16 <= 2 * kT_TeamBytes, "token staging lives in the exchange buffers");
    auto stage_tokens = [&](int64_t tile) {
      const int64_t row = min((tile * ppt + min(r, rows - 1) / V) * V + r % V, a.count * V - 1);
      const uint4* src = reinterpret_cast<const uint4*>(a.tokens + row * (int64_t)kTokLd + 40 * q);
#pragma unroll
      for (int c = 0; c < 5; ++c) cp_async16(xstage + c * (kFEpiThreads * 16), src + c);
      cp_async_commit();
    };
    uint4 xn[5];
    auto fetch_staged = [&]() {
      cp_async_wait_all();
#pragma unroll
      for (int c = 0; c < 5; ++c) xn[c] = *reinterpret_cast<const uint4*>(xstage + c * (kFEpiThreads * 16));
    };
    auto unpack_tokens = [&](float (&x)[40]) {
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&xn[c]);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h2[i]); x[8 * c + 2 * i] = f.x; x[8 * c + 2 * i + 1] = f.y; }
      }
    };
    auto store_x = [&](const float (&x)[40]) {
      uint32_t xb[40];
#pragma unroll
      for (int c = 0; c < 40; ++c) xb[c] = __float_as_uint(x[c]);
#pragma unroll
      for (int i = 0; i < 5; ++i) tmem_st_x8(tl + kT_ColX + 40 * q + 8 * i, *reinterpret_cast<const uint32_t(*)[8]>(&xb[8 * i]));
    };
    // LN1 of layer 0 for the staged tile -> YT, and the hand-over that lets its first q|k|v GEMM start
    auto open_tile = [&]() {
      fetch_staged();
      float x[40];
      unpack_tokens(x);
      ln_to_tmem(x, nullptr, nullptr, nullptr, LS, r, q, warp, tl);
      hand_over();
    };

    if (cid * kC < ntiles) {
      // first tile of this CTA: tokens -> LN1 -> YT, tokens -> X.  Every later tile is opened inside the previous
      // one (below), under its last GEMM.
      pf.start();
      stage_tokens(cid * kC + crank);
      open_tile();
      float x[40];
      unpack_tokens(x);
      store_x(x);
      tmem_st_wait();
      pf.stop(acc_tl);
    }

    for (int64_t tbase = cid * kC; tbase < ntiles; tbase += ncl * kC) {
      const int64_t tile = tbase + crank;              // tile >= ntiles: all rows invalid
      const int64_t pnt = tile * ppt + r / V;
      const int tok = r % V;
      const bool valid = (r < rows) && (pnt < a.count);
      const bool has_next = tbase + ncl * kC < ntiles;
#pragma unroll 1
      for (int l = 0; l < 2; ++l) {
        const float* fp = FP + l * 320;             // pend_in, pend_mid of this layer
        {
          // ---- attention (lib/transformer.py:59-71).  Two teams of 8 warps (team = q >> 1: A = heads 0, 2;
          // B = heads 1, 3) work on different heads at the same time.  R = [q | k | v] of one head, 64 columns
          // each; thread (r, hf = q & 1) owns head dims [32 hf, 32 hf + 32) of its row.  It keeps its own q, k, v
          // slices in registers and publishes k and v as fp16 (unit-swizzled 128-byte rows) for the V - 1
          // partner rows of its point; softmax and the weighted sum are symmetric in the key order, so keys
          // are visited as (self, partner 1, ..).
          const int team = q >> 1, hf = q & 1;
          uint8_t* KXt = smem + team * kT_TeamBytes + kT_KX;
          uint8_t* VXt = smem + team * kT_TeamBytes + kT_VX;
          float* PDt = reinterpret_cast<float*>(smem + team * kT_TeamBytes + kT_PD);
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {
            pf.start();
            if (hh == 1) {
              // the team's second head may already be waiting in R: its exchange buffers are free only when every
              // thread of the team is done with the first head (and its half of OT only when that head's
              // out-projection has read it: w_bar, checked right before o is stored)
              named_bar_sync(6 + team, kFEpiThreads / 2);
            }
            mbar_wait(&pipe->q_bar[team], pq & 1);
            ++pq;
            tc_fence_after();
            pf.stop(acc_d);
            float qv[32];
            __half2 vh[16];
            {
              float kk[32];
              tmem_ld_x32(tl + kT_ColR + 64 + 32 * hf, kk);
              tmem_ld_x32(tl + kT_ColR + 32 * hf, qv);
              tmem_ld_wait();
              uint8_t* kd = KXt + r * 128;
#pragma unroll
              for (int u = 0; u < 4; ++u)
                *reinterpret_cast<uint4*>(kd + (((4 * hf + u) ^ (r & 7)) << 4)) =
                    make_uint4(pack_h2(kk[8 * u], kk[8 * u + 1]), pack_h2(kk[8 * u + 2], kk[8 * u + 3]),
                               pack_h2(kk[8 * u + 4], kk[8 * u + 5]), pack_h2(kk[8 * u + 6], kk[8 * u + 7]));
              float d[4] = {0.f, 0.f, 0.f, 0.f};      // own key, fp32
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                d[0] = fmaf(qv[4 * u], kk[4 * u], d[0]); d[1] = fmaf(qv[4 * u + 1], kk[4 * u + 1], d[1]);
                d[2] = fmaf(qv[4 * u + 2], kk[4 * u + 2], d[2]); d[3] = fmaf(qv[4 * u + 3], kk[4 * u + 3], d[3]);
              }
              PDt[(hf * 128 + r) * 4] = (d[0] + d[1]) + (d[2] + d[3]);
            }
            {
              float vv[32];
              tmem_ld_x32(tl + kT_ColR + 128 + 32 * hf, vv);
              tmem_ld_wait();
              if (team + 2 * hh < 3) { tc_fence_before(); mbar_arrive(&pipe->r_bar); }   // R is free: the next head's q|k|v may land
              uint8_t* vd = VXt + r * 128;
#pragma unroll
              for (int i = 0; i < 16; ++i) vh[i] = __floats2half2_rn(vv[2 * i], vv[2 * i + 1]);
#pragma unroll
              for (int u = 0; u < 4; ++u)
                *reinterpret_cast<uint4*>(vd + (((4 * hf + u) ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(&vh[4 * u]);
            }
            pf.stop(sec[0]);
            named_bar_sync(6 + team, kFEpiThreads / 2);      // a point's rows may sit in different lane quarters: team-wide
            pf.stop(sec[1]);
            // partial dots with the partner keys over this thread's 32 of the 64 head dims.  The partner keys arrive as
            // fp16 anyway; q is rounded to fp16 as well and the products run as packed half2 FMAs into eight short
            // partial sums (four products each, so the fp16 accumulation error stays ~2^-11 of a term), which are
            // added up in fp32 -- a third of the instructions of unpacking every key to fp32 first.
            if (a.opt & 1) {
#pragma unroll
              for (int j = 1; j < V; ++j) {
                const int rj = p0 + ((tok + j >= V) ? tok + j - V : tok + j);
                const uint8_t* src = KXt + rj * 128;
                uint4 kp[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) kp[u] = *reinterpret_cast<const uint4*>(src + (((4 * hf + u) ^ (rj & 7)) << 4));
                float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const __half2* h2 = reinterpret_cast<const __half2*>(&kp[u]);
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(h2[i]);
                    d[(2 * i) & 3] = fmaf(qv[8 * u + 2 * i], f.x, d[(2 * i) & 3]);
                    d[(2 * i + 1) & 3] = fmaf(qv[8 * u + 2 * i + 1], f.y, d[(2 * i + 1) & 3]);
                  }
                }
                PDt[(hf * 128 + r) * 4 + j] = (d[0] + d[1]) + (d[2] + d[3]);
              }
            } else {
              __half2 qh[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) qh[i] = __floats2half2_rn(qv[2 * i], qv[2 * i + 1]);
#pragma unroll
              for (int j = 1; j < V; ++j) {
                const int rj = p0 + ((tok + j >= V) ? tok + j - V : tok + j);
                const uint8_t* src = KXt + rj * 128;
                uint4 kp[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) kp[u] = *reinterpret_cast<const uint4*>(src + (((4 * hf + u) ^ (rj & 7)) << 4));
                __half2 ac[4] = {__float2half2_rn(0.f), __float2half2_rn(0.f), __float2half2_rn(0.f), __float2half2_rn(0.f)};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const __half2* h2 = reinterpret_cast<const __half2*>(&kp[u]);
#pragma unroll
                  for (int i = 0; i < 4; ++i) ac[i] = __hfma2(qh[4 * u + i], h2[i], ac[i]);
                }
                const float2 f0 = __half22float2(ac[0]), f1 = __half22float2(ac[1]), f2 = __half22float2(ac[2]), f3 = __half22float2(ac[3]);
                PDt[(hf * 128 + r) * 4 + j] = ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
              }
            }
            pf.stop(sec[2]);
            named_bar_sync(8 + 4 * team + (warp & 3), 64);    // the two threads of a row (warps w, w + 4 of the team)
            pf.stop(sec[3]);
            float w[V];
            float mx = -1e30f;
#pragma unroll
            for (int j = 0; j < V; ++j) {
              w[j] = (PDt[r * 4 + j] + PDt[(128 + r) * 4 + j]) * 0.125f;      // dim_head ** -0.5
              mx = fmaxf(mx, w[j]);
            }
            float den = 0.f;
#pragma unroll
            for (int j = 0; j < V; ++j) { w[j] = __expf(w[j] - mx); den += w[j]; }
            const float inv = 1.0f / den;
            // o = sum_j softmax_j * v_j over this thread's 32 dims, accumulated as half2 (V terms; the result is
            // rounded to bf16 for the out-projection anyway)
            {
              const __half2 w0 = __float2half2_rn(w[0] * inv);
#pragma unroll
              for (int i = 0; i < 16; ++i) vh[i] = __hmul2(w0, vh[i]);
            }
#pragma unroll
            for (int j = 1; j < V; ++j) {
              const int rj = p0 + ((tok + j >= V) ? tok + j - V : tok + j);
              const __half2 wj = __float2half2_rn(w[j] * inv);
              const uint8_t* src = VXt + rj * 128;
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const uint4 pkv = *reinterpret_cast<const uint4*>(src + (((4 * hf + u) ^ (rj & 7)) << 4));
                const __half2* h2 = reinterpret_cast<const __half2*>(&pkv);
#pragma unroll
                for (int i = 0; i < 4; ++i) vh[4 * u + i] = __hfma2(wj, h2[i], vh[4 * u + i]);
              }
            }
            uint32_t pk[16];         // o leaves as the fp16 pairs it was accumulated in: the out-projection is an fp16 GEMM
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = *reinterpret_cast<const uint32_t*>(&vh[i]);
            if (hh == 1) { mbar_wait(&pipe->w_bar[team], pw & 1); ++pw; }     // (long complete by now)
            tmem_st_u16(tl + kT_ColO + 32 * team + 16 * hf, pk);
            tmem_st_wait();
            pf.stop(sec[4]);
            tc_fence_before();
            mbar_arrive(&pipe->o_bar[team]);
          }
        }
        {
          // ---- x (+ deferred biases) -> LN2 -> YT
          wait_d();
          // every thread of both teams is past its attention: the exchange buffers are free for the staging
          if (l == 1 && has_next) stage_tokens(tile + ncl * kC);
          pf.start();
          float x[40];
          load_x40(tl + kT_ColX + 40 * q, x);
          ln_to_tmem(x, fp + 160, nullptr, nullptr, LS, r, q, warp, tl);
          pf.stop(sec[5]);
          hand_over();
        }
        {
          // ---- FF hidden: GELU(acc + b1) -> bf16 operand (K = 128 -> 64 packed columns); per hidden half
          pf.start();
          mbar_wait(&pipe->h_bar[q >> 1], ph & 1);
          ++ph;
          tc_fence_after();
          pf.stop(acc_d);
          pf.start();
          float t[32];
          tmem_ld_x32(tl + kT_ColR + 32 * q, t);
          tmem_ld_wait();
          uint32_t pk[16];         // (the bias b1 rides in the GEMM: K column 155 of W1 against the operand's constant 1.0, pack.py)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            pk[2 * i] = gelu_pair_bf16(t[4 * i], t[4 * i + 1]);
            pk[2 * i + 1] = gelu_pair_bf16(t[4 * i + 2], t[4 * i + 3]);
          }
          tmem_st_u16(tl + kT_ColO + 16 * q, pk);
          tmem_st_wait();
          pf.stop(sec[6]);
          tc_fence_before();
          mbar_arrive(&pipe->g_bar[q >> 1]);
        }
        {
          // Last layer: the next tile is opened first -- its LN1 needs only the staged tokens and YT, which the FF
          // hidden GEMM has finished reading -- so that the tensor pipe runs its first q|k|v GEMM right behind this
          // tile's last GEMM, while the epilogue writes this tile's output.  X is overwritten with the next tile's
          // tokens only after it has been read; the first GEMM that accumulates into X (Wo of head 0) is issued
          // after every thread of both teams has arrived at r_bar / o_bar, i.e. after these stores.
          if (l == 1 && has_next) { pf.start(); open_tile(); pf.stop(acc_tl); }
          wait_d();
          pf.start();
          float x[40];
          load_x40(tl + kT_ColX + 40 * q, x);
          if (l == 0) {
            ln_to_tmem(x, FP + 320, nullptr, nullptr, LS, r, q, warp, tl);      // layer 1: LN1 on x + pend_in
            pf.stop(sec[7]);
            hand_over();
          } else {
            if (valid && tok < 2) {
              // ---- output tokens 0 (density branch) and 1 (colour branch), lib/skinnning_batch.py:441-442
              const float4* p4 = reinterpret_cast<const float4*>(FP + 640 + 40 * q);
              uint4* dst = reinterpret_cast<uint4*>((tok == 0 ? a.tok0 : a.tok1) + pnt * kTokLd + 40 * q);
#pragma unroll
              for (int c = 0; c < 5; ++c) {
                const float4 pa = p4[2 * c], pb = p4[2 * c + 1];
                dst[c] = make_uint4(pack_bf16x2(x[8 * c] + pa.x, x[8 * c + 1] + pa.y), pack_bf16x2(x[8 * c + 2] + pa.z, x[8 * c + 3] + pa.w),
                                    pack_bf16x2(x[8 * c + 4] + pb.x, x[8 * c + 5] + pb.y), pack_bf16x2(x[8 * c + 6] + pb.z, x[8 * c + 7] + pb.w));
              }
            }
            if (has_next) {
              unpack_tokens(x);
              store_x(x);
              tmem_st_wait();
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
  if (warp == kFMmaWarp) tmem_dealloc(tm, 512);
}

// ------------------------------------------------------------------------------------------
// M: canonical NeRF MLP.  Tile = 128 points.  TMEM columns:
//   ACC0 [0,128), ACC1 [128,256)   the two N-halves of the layer output (fp32)
//   HT   [256,384)  hidden activations, bf16 x2 per column (K = 256); K-half j = columns 256 + 64 j ..
//   XT   [384,496)  x = [tok0 155 | 0 x5 | PE6(xc) 39 | 0 x25] (K = 224, 14 K-steps);
//                   reused for tok1 (K = 160) after layer 5
// Software pipeline inside one tile (all MMAs are M = 128, N = 128, K = 16):
//   a layer is issued as  [h0: K0 K1] commit d_bar0  [h1: K0] commit k_bar  [h1: K1] commit d_bar1
//   the epilogue turns half 0 (e0) into K-half 0 of the next layer's operand while the tensor pipe
//   computes half 1, and half 1 (e1) while it already runs the next layer's [h0: K0]:
//     [h0': K0] needs e0 (a_bar0), [h0': K1] needs e1 (a_bar1); e0 may overwrite HT K-half 0 only
//     after [h1: K0] has read it (k_bar).
// The next tile's x is prefetched into registers and stored to XT inside the views epilogue,
// before the colour math, so the tensor pipe restarts while the tile's output is still being written.
// Shared memory holds only the weight ring (12 x 16 KB) and the fp32 parameter vectors.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kM_ColAcc = 0, kM_ColH = 256, kM_ColX = 384;
constexpr uint32_t kM_RING = 0;
constexpr int kM_Slots = 3;
static_assert(kM_Slots <= 16 && kT_Slots <= 16, "Pipe holds 16 ring barriers");
constexpr uint32_t kM_SlotBytes = 65536;         // one half-layer: 128 weight rows x K = 256 (4 chunks, 16 MMAs)
constexpr int kM_SlotsPerTile = 22;
constexpr uint32_t kM_FP = kM_RING + kM_Slots * kM_SlotBytes;
constexpr uint32_t kM_PART = kM_FP + ((kMFloats * 4 + 15) / 16) * 16;   // float4[3][128]: alpha / rgb partials of quarters 1..3
constexpr uint32_t kM_PIPE = kM_PART + 3 * 128 * 16;
constexpr uint32_t kM_Smem = kM_PIPE + sizeof(Pipe);
static_assert(kM_Smem <= 232448 - 1024, "M kernel shared memory over budget");
static_assert(kM_SlotsPerTile * kM_SlotBytes == kMBytes, "M weight stream drifted from pack.py");

struct MArgs {
  const __nv_bfloat16* tok0;
  const __nv_bfloat16* tok1;
  const float* xc;          // (count, 3) canonical points
  int64_t count;
  const uint8_t* blob;
  const int32_t* act_pid;   // already offset by `first`
  float* raw;               // (P, 4)
  int prof;
  const int32_t* count_dev;   // optional device-side active count (count is then the slab capacity)
  int64_t first;
};

__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// The MMA warp's program for one tile, as data (keeps its code one small loop that stays in the
// instruction cache).  A step = 8 MMAs = one half (K = 128) of a ring slot; a slot = the weights
// of one half-layer (128 rows x K = 256, 64 KB), numbered 0..21 within a tile:
//   0,1 L0 h0/h1 | 2L+2, 2L+3 for L1..L4 | 10 L5.h0 x part, 11 L5.h0 h part, 12 L5.h1 x, 13 L5.h1 h |
//   14..19 L6, L7, feature | 20 views tok1 part, 21 views feature part.
// Issue-side budget (tools/ubench_tc.cu, in-kernel traces): the tensor pipe queues only ~3 MMAs
// (~190 cycles) and every mbarrier probe costs the issuing warp ~100 cycles, not pipelined.  So the
// MMA warp waits on at most one barrier per step: the "weights landed" barriers are waited on by the
// EPILOGUE warps (which have slack) before they arrive on a_bar -- a_bar0/1 then stand for
// "operand ready, accumulator drained, weights of the steps up to the next a_bar wait landed".
// wait bits: 1 a_bar0, 2 a_bar1, 4 full barrier of the own slot (the second and later slots of L5 and
// views, which cannot be resident early enough in a 3-slot ring); commit: 1 d_bar0, 2 k_bar, 3 d_bar1; sub = slot half;
// K is padded to whole slots: x (K = 224) and tok1 (K = 160) read on into TMEM columns that hold
// finite values (stale PE, zeroed [496,512)) against zero weight columns.
// One packed word per step (a single constant load, prefetched one step ahead): bits 0-1 (acol - 256) / 64,
// 2-4 wait, 5 dhalf, 6 fresh, 7-8 commit (after the step), 9-10 mode (0 = first half of the slot, 8 MMAs;
// 1 = second half, 8 MMAs, releases the slot; 2 = whole slot, 16 MMAs, releases it), 11 midk (commit k_bar
// after the first 8 MMAs of a whole-slot step).  Steps are as long as the dependencies allow: every step
// boundary costs the MMA warp ~250 cycles of descriptor set-up and bookkeeping that the shallow tensor
// queue (~3 MMAs) cannot hide.
#define MPS_OP(acol, wait, dhalf, fresh, commit, mode, midk) \
  ((uint32_t)(((acol) - 256) / 64) | ((wait) << 2) | ((dhalf) << 5) | ((fresh) << 6) | ((commit) << 7) | ((mode) << 9) | ((midk) << 11))
#define MPS_M_HIDDEN MPS_OP(256, 1, 0, 1, 0, 0, 0), MPS_OP(320, 2, 0, 0, 1, 1, 0), MPS_OP(256, 0, 1, 1, 3, 2, 1)
__constant__ uint32_t kMSchedule[] = {
    MPS_OP(384, 1 | 2, 0, 1, 1, 2, 0), MPS_OP(384, 0, 1, 1, 3, 2, 1),                                        // L0: A = x
    MPS_M_HIDDEN, MPS_M_HIDDEN, MPS_M_HIDDEN, MPS_M_HIDDEN,                                                  // L1..L4
    MPS_OP(384, 1, 0, 1, 0, 2, 0), MPS_OP(256, 4, 0, 0, 0, 0, 0), MPS_OP(320, 2, 0, 0, 1, 1, 0),             // L5: [x | h], x part first
    MPS_OP(384, 4, 1, 1, 0, 2, 0), MPS_OP(256, 4, 1, 0, 3, 2, 1),
    MPS_M_HIDDEN, MPS_M_HIDDEN, MPS_M_HIDDEN,                                                                // L6, L7, feature
    MPS_OP(384, 1, 0, 1, 0, 2, 0), MPS_OP(256, 4, 0, 0, 0, 0, 0), MPS_OP(320, 2, 0, 0, 1, 1, 0),             // views: [tok1 | feature]
    MPS_OP(384, 1 | 2, 0, 1, 1, 2, 0)};                                                                      // (prefetch target past the end = step 0)
#undef MPS_M_HIDDEN
struct MOp {
  uint32_t w;
  __device__ __forceinline__ uint32_t acol() const { return 256u + 64u * (w & 3u); }
  __device__ __forceinline__ uint32_t wait() const { return (w >> 2) & 7u; }
  __device__ __forceinline__ uint32_t dcol() const { return ((w >> 5) & 1u) * 128u; }
  __device__ __forceinline__ bool fresh() const { return (w >> 6) & 1u; }
  __device__ __forceinline__ uint32_t commit() const { return (w >> 7) & 3u; }
  __device__ __forceinline__ uint32_t mode() const { return (w >> 9) & 3u; }
  __device__ __forceinline__ bool midk() const { return (w >> 11) & 1u; }
};
constexpr int kMScheduleLen = 2 + 7 * 3 + 5 + 3;
// slots whose arrival the epilogue of (layer L, half j) vouches for when it arrives on a_bar[j]: {first, count}
__constant__ uint8_t kMCarry[9][2][2] = {
    {{2, 1}, {3, 1}}, {{4, 1}, {5, 1}}, {{6, 1}, {7, 1}}, {{8, 1}, {9, 1}}, {{10, 1}, {0, 0}},
    {{14, 1}, {15, 1}}, {{16, 1}, {17, 1}}, {{18, 1}, {19, 1}}, {{20, 1}, {0, 0}}};
static_assert(sizeof(kMSchedule) / sizeof(uint32_t) == kMScheduleLen + 1, "M schedule length");

// 640 threads = 5 warpgroups: four epilogue warpgroups (16 warps, four threads per row: thread
// (r = tid & 127, q = tid >> 7) owns 32 of the 128 columns of an accumulator half; raised to 104
// registers) and one holding the producer warp, the MMA warp and two idle warps (lowered to 40).
constexpr int kMThreads = kFThreads, kMEpiThreads = kFEpiThreads, kMProdWarp = kFProdWarp, kMMmaWarp = kFMmaWarp;
__device__ __forceinline__ void m_epi_bar() { named_bar_sync(1, kMEpiThreads); }


// 16 elements (8 packed words) [16 q, 16 q + 16) of the 39-wide positional code [x, sin(f0 x), cos(f0 x), ...],
// f_k = pi 2^k (run_nerf_helpers.py:337-353); elements >= 39 are zero padding.  One sincospi per
// channel, then the double-angle recurrence for the 5 higher octaves (abs error < 3e-6 after 5
// doublings, three orders of magnitude below the bf16 rounding of the operand).
__device__ __forceinline__ void pe_words(const float (&xc)[3], int q, uint32_t (&pk)[8]) {
  float pe[64];
#pragma unroll
  for (int e = 0; e < 64; ++e) pe[e] = 0.f;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    pe[ch] = xc[ch];
    float s, c;
    sincospif(xc[ch], &s, &c);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      pe[3 + 6 * k + ch] = s;
      pe[3 + 6 * k + 3 + ch] = c;
      const float s2 = 2.0f * s * c, c2 = fmaf(-2.0f * s, s, 1.0f);
      s = s2; c = c2;
    }
  }
#pragma unroll
  for (int w2 = 0; w2 < 8; ++w2) {      // q is warp-uniform: three selects per word instead of four code copies
    const float lo = q == 0 ? pe[2 * w2] : q == 1 ? pe[16 + 2 * w2] : q == 2 ? pe[32 + 2 * w2] : pe[48 + 2 * w2];
    const float hi = q == 0 ? pe[2 * w2 + 1] : q == 1 ? pe[16 + 2 * w2 + 1] : q == 2 ? pe[32 + 2 * w2 + 1] : pe[48 + 2 * w2 + 1];
    pk[w2] = pack_bf16x2(lo, hi);
  }
}

// this thread's 20 packed columns of a bf16 token row (160 values split between the four quarters).
// Unconditional loads (the caller clamps the row index): nothing consumes the registers until the
// token is stored, so the global latency hides behind the work in between.
__device__ __forceinline__ void token_load(const __nv_bfloat16* row, int q, uint32_t (&w)[20]) {
  const uint4* src = reinterpret_cast<const uint4*>(row) + 5 * q;     // 40 bf16 = 5 x 16 bytes per quarter
#pragma unroll
  for (int u = 0; u < 5; ++u) {
    const uint4 t = __ldg(src + u);
    w[4 * u] = t.x; w[4 * u + 1] = t.y; w[4 * u + 2] = t.z; w[4 * u + 3] = t.w;
  }
}
__device__ __forceinline__ void token_store(uint32_t taddr, int q, const uint32_t (&w)[20]) {
#pragma unroll
  for (int i = 0; i < 5; ++i) tmem_st_x4(taddr + 20 * q + 4 * i, &w[4 * i]);
}

template <int kC>
__global__ void __launch_bounds__(kMThreads, 1) mlp_tc_kernel(const MArgs a_in) {
  MArgs a = a_in;
  if (a.count_dev != nullptr) {
    const int64_t n = (int64_t)*a.count_dev - a.first;
    a.count = n < a.count ? (n > 0 ? n : 0) : a.count;
  }
  extern __shared__ __align__(1024) uint8_t smem[];
  Pipe* pipe = reinterpret_cast<Pipe*>(smem + kM_PIPE);
  float* FP = reinterpret_cast<float*>(smem + kM_FP);
  float* PART = reinterpret_cast<float*>(smem + kM_PART);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (a.count + 127) / 128;
  const uint32_t crank = (kC > 1) ? cluster_ctarank() : 0u;
  const int64_t ncl = gridDim.x / kC, cid = blockIdx.x / kC;
  const int64_t tstride = ncl * kC;

  if (tid == 0) {
    for (int i = 0; i < kM_Slots; ++i) { mbar_init(&pipe->full[i], 1); mbar_init(&pipe->empty[i], kC); }
    mbar_init(&pipe->a_bar[0], kMEpiThreads);
    mbar_init(&pipe->a_bar[1], kMEpiThreads);
    mbar_init(&pipe->d_bar[0], 1);
    mbar_init(&pipe->d_bar[1], 1);
    mbar_init(&pipe->k_bar, 1);
    mbar_fence_init();
  }
  if (warp == kMMmaWarp) { tmem_alloc(&pipe->tmem_base, 512); tmem_relinquish(); }
  {
    const float* src = reinterpret_cast<const float*>(a.blob + kFloatOff) + kTFloats;
    for (int i = tid; i < kMFloats; i += kMThreads) FP[i] = src[i];
  }
  tc_fence_before();
  __syncthreads();
  if (kC > 1) cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = pipe->tmem_base;

  if (warp >= 16) {
    reg_dec<40>();           // fifth warpgroup: producer, MMA issuer, two idle warps
  if (warp == kMProdWarp) {
    RingProducer<kM_Slots, kM_SlotBytes, kC> P{pipe, smem + kM_RING, crank};
    P.pf = Prof{a.prof == 1 && lane == 0, 0};
    for (int64_t tbase = cid * kC; tbase < ntiles; tbase += tstride) {
      const uint8_t* src = a.blob + kTBytes;
      for (int i = 0; i < kM_SlotsPerTile; ++i) P.push(src, kM_SlotBytes);
    }
    if (P.pf.on) atomicAdd(&g_prof[1][6], (unsigned long long)P.acc_e);
  } else if (warp == kMMmaWarp) {
    // The whole warp runs the schedule; one elected lane issues the tcgen05 instructions.
    const uint32_t ring_addr = smem_u32(smem + kM_RING);
    constexpr uint32_t idesc = instr_desc_bf16(128);
    uint32_t pa0 = 0, pa1 = 0, slot = 0, phase = 0, opw = kMSchedule[0];
    Trace tr{false, 0, 0};
    Prof pf{a.prof == 1 && lane == 0, 0};
    long long acc_a = 0;
    const long long t_begin = clock64();
    for (int64_t tbase = cid * kC; tbase < ntiles; tbase += tstride) {
      tr.on = (a.prof == 2) && blockIdx.x == 0 && lane == 0 && tbase == cid * kC + tstride;
      if (a.prof == 5 && blockIdx.x == 0 && lane == 0) { Trace t5{true, tr.n, 0}; t5.ev(30); tr.n = t5.n; }
#pragma unroll 1
      for (int o = 0; o < kMScheduleLen; ++o) {
        const MOp op{opw};
        opw = kMSchedule[o + 1];               // prefetch the next step's word (entry [len] = step 0)
        const uint32_t wait = op.wait(), mode = op.mode();
        // operands first (uniform-datapath work), then the barrier probes, then the MMAs back to back
        const uint64_t bdesc = smem_desc_sw128(ring_addr + slot * kM_SlotBytes + (mode == 1 ? 32768u : 0u));
        const uint32_t d_addr = tm + op.dcol(), a_addr = tm + op.acol(), commit = op.commit();
        const uint32_t acc0 = op.fresh() ? 0u : 1u;
        if (wait) {
          tr.ev(20);
          pf.start();
          if (wait & 1) { mbar_wait(&pipe->a_bar[0], pa0 & 1); ++pa0; }
          if (wait & 2) { mbar_wait(&pipe->a_bar[1], pa1 & 1); ++pa1; }
          if (wait & 4) mbar_wait(&pipe->full[slot], phase);
          pf.stop(acc_a);
          tc_fence_after();
          tr.ev(21);
        }
        if (elect_one()) {       // K-steps of A (TMEM columns acol ..) x 128 weight rows -> accumulator half
          // descriptor of K-step k = base + (chunk k / 4) * 16 KB + (k % 4) * 32 B, in 16-byte units
#pragma unroll
          for (int k = 0; k < 8; ++k)
            mma_bf16_ts(d_addr, a_addr + k * 8, bdesc + (uint64_t)((k >> 2) * 1024 + (k & 3) * 2), idesc, k ? 1u : acc0);
          if (mode == 2) {
            if (op.midk()) mma_commit(&pipe->k_bar);
#pragma unroll
            for (int k = 8; k < 16; ++k)
              mma_bf16_ts(d_addr, a_addr + k * 8, bdesc + (uint64_t)((k >> 2) * 1024 + (k & 3) * 2), idesc, 1u);
          }
          if (mode != 0) {          // slot consumed
            if (kC == 1) mma_commit(&pipe->empty[slot]);
            else mma_commit_mc(&pipe->empty[slot], (uint16_t)((1u << kC) - 1u));
          }
          if (commit) mma_commit(commit == 1 ? &pipe->d_bar[0] : &pipe->d_bar[1]);
        }
        __syncwarp();
        if (commit) tr.ev(4 + commit);
        if (mode != 0) { if (++slot == kM_Slots) { slot = 0; phase ^= 1; } }
      }
    }
    if (pf.on) {
      atomicAdd(&g_prof[1][3], (unsigned long long)acc_a);
      atomicAdd(&g_prof[1][5], (unsigned long long)(clock64() - t_begin));
    }
  }
  } else {
    reg_inc<104>();
    const int r = tid & 127, q = tid >> 7;
    const uint32_t tl = tm + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t pd[2] = {0, 0}, pk_ = 0;
    Prof pf{a.prof == 1 && tid == 0, 0};
    long long acc_d = 0, acc_tl = 0, n_tiles = 0;
    const long long t_begin = clock64();
    Trace tr{false, 0, 256};
    auto wait_d = [&](int j) { tr.ev(10 + j); pf.start(); mbar_wait(&pipe->d_bar[j], pd[j] & 1); pf.stop(acc_d); ++pd[j]; tc_fence_after(); tr.ev(12 + j); };
    auto wait_k = [&]() { pf.start(); mbar_wait(&pipe->k_bar, pk_ & 1); pf.stop(acc_d); ++pk_; tc_fence_after(); };
    auto arrive = [&](int j) { tc_fence_before(); mbar_arrive(&pipe->a_bar[j]); tr.ev(16 + j); };
    const float* bias = FP;                       // 8 x 256
    const float* w_alpha = FP + 2048;
    const float* b_feat = FP + 2304;
    const float* b_views = FP + 2560;
    const float* w_rgb = FP + 2688;               // 3 x 128
    const float* b_tail = FP + 3072;              // b_alpha, b_rgb[3]

    // one warp per a_bar waits for the weight slots [first, first + n) (tile-relative numbering) to land
    int64_t seq0 = 0;            // ring sequence number of this tile's slot 0
    auto vouch = [&](int j, int first_slot, int n) {
      if (warp == 4 * j) {
        for (int u = 0; u < n; ++u) {
          const int64_t sq = seq0 + first_slot + u;
          mbar_wait(&pipe->full[sq % kM_Slots], (uint32_t)((sq / kM_Slots) & 1));
        }
      }
    };
    // The next tile's x is fetched in steps spread over the current tile: global loads before the
    // feature layer, positional code after it, TMEM store inside the views epilogue.  Rows past the
    // end are clamped to the last point: rows are independent and their results are never written.
    uint32_t xw[20], xpe[8];
    float xcn[3];
    auto load_x = [&](int64_t tile) {
      const int64_t i = min(tile * 128 + r, a.count - 1);
      token_load(a.tok0 + i * kTokLd, q, xw);
      xcn[0] = __ldg(a.xc + 3 * i); xcn[1] = __ldg(a.xc + 3 * i + 1); xcn[2] = __ldg(a.xc + 3 * i + 2);
    };
    auto store_x = [&](int64_t seq_next) {       // XT = [tok0 | 0 | PE6(xc) | 0]; this thread: 20 token columns + 8 PE columns
      token_store(tl + kM_ColX, q, xw);
      tmem_st_x8(tl + kM_ColX + 80 + 8 * q, xpe);
      if (warp == 0) mbar_wait(&pipe->full[seq_next % kM_Slots], (uint32_t)((seq_next / kM_Slots) & 1));
      if (warp == 4) mbar_wait(&pipe->full[(seq_next + 1) % kM_Slots], (uint32_t)(((seq_next + 1) / kM_Slots) & 1));
      tmem_st_wait();
      arrive(0); arrive(1);
    };
    // half j of a hidden layer: 32 accumulator columns of this thread -> (+bias, ReLU) -> bf16 x2 -> HT
    auto epi_half = [&](auto relu_c, auto alpha_c, int L, int j, const float* __restrict__ b, float& alpha) {
      constexpr bool relu = decltype(relu_c)::value, with_alpha = decltype(alpha_c)::value;
      const int c0 = 128 * j + 32 * q;
      wait_d(j);
      float t[32];
      tmem_ld_x32(tl + kM_ColAcc + c0, t);
      vouch(j, kMCarry[L][j][0], kMCarry[L][j][1]);      // overlaps the TMEM load
      tmem_ld_wait();
      tr.ev(18);
      const float4* b4 = reinterpret_cast<const float4*>(b + c0);
      const float4* w4 = reinterpret_cast<const float4*>(w_alpha + c0);
      uint32_t pk[16];
      float a4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float4 bb = b4[u];
        const float v0 = t[4 * u] + bb.x, v1 = t[4 * u + 1] + bb.y, v2 = t[4 * u + 2] + bb.z, v3 = t[4 * u + 3] + bb.w;
        if (with_alpha) {
          const float4 ww = w4[u];
          a4[0] = fmaf(fmaxf(v0, 0.f), ww.x, a4[0]); a4[1] = fmaf(fmaxf(v1, 0.f), ww.y, a4[1]);
          a4[2] = fmaf(fmaxf(v2, 0.f), ww.z, a4[2]); a4[3] = fmaf(fmaxf(v3, 0.f), ww.w, a4[3]);
        }
        if (relu) { pk[2 * u] = pack_relu_bf16x2(v0, v1); pk[2 * u + 1] = pack_relu_bf16x2(v2, v3); }
        else { pk[2 * u] = pack_bf16x2(v0, v1); pk[2 * u + 1] = pack_bf16x2(v2, v3); }
      }
      if (with_alpha) alpha += (a4[0] + a4[1]) + (a4[2] + a4[3]);
      if (j == 0) wait_k();                      // [h1: K0] has read the old K-half 0
      tmem_st_u16(tl + kM_ColH + c0 / 2, pk);
    };
    using T_ = std::true_type; using F_ = std::false_type;

    if (q == 0) {      // K padding of the x / tok1 operand: TMEM columns [496, 512) must hold finite values
      const uint32_t z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      tmem_st_u16(tl + 496, z);
      tmem_st_wait();
    }
    bool first = true;
    for (int64_t tbase = cid * kC; tbase < ntiles; tbase += tstride) {
      const int64_t tile = tbase + crank;
      const int64_t i = tile * 128 + r;
      const bool valid = i < a.count;
      const bool has_next = tbase + tstride < ntiles;
      tr.on = (a.prof == 2) && blockIdx.x == 0 && tid == 0 && tbase == cid * kC + tstride;
      if (first) {
        pf.start();
        load_x(tile);
        pe_words(xcn, q, xpe);
        store_x(0);
        pf.stop(acc_tl);
        first = false;
      }
      float alpha = 0.f;
#pragma unroll 1
      for (int L = 0; L < 5; ++L) {          // L0..L4
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
          epi_half(T_{}, F_{}, L, j, bias + 256 * L, alpha);
          tmem_st_wait();
          arrive(j);
        }
      }
      {                                      // L5: afterwards x is dead and XT takes tok1 for the views layer
        uint32_t t1[20];
        token_load(a.tok1 + min(i, a.count - 1) * kTokLd, q, t1);
        epi_half(T_{}, F_{}, 5, 0, bias + 256 * 5, alpha);
        tmem_st_wait();
        arrive(0);
        epi_half(T_{}, F_{}, 5, 1, bias + 256 * 5, alpha);
        token_store(tl + kM_ColX, q, t1);    // layer 5 has been accumulated (d_bar1): nobody reads x any more
        tmem_st_wait();
        arrive(1);
      }
#pragma unroll 1
      for (int L = 6; L < 8; ++L) {          // L6, L7 (+ alpha_linear on the fp32 activations of L7)
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
          if (L == 7) epi_half(T_{}, T_{}, L, j, bias + 256 * L, alpha); else epi_half(T_{}, F_{}, L, j, bias + 256 * L, alpha);
          tmem_st_wait();
          arrive(j);
        }
      }
      if (has_next) load_x(tile + tstride);  // in flight during the feature layer
#pragma unroll 1
      for (int j = 0; j < 2; ++j) {          // feature (no activation)
        epi_half(F_{}, F_{}, 8, j, b_feat, alpha);
        tmem_st_wait();
        arrive(j);
      }
      {
        // ---- views layer epilogue: relu -> rgb_linear on CUDA cores -> raw[pid] = (rgb, alpha)
        if (has_next) pe_words(xcn, q, xpe);
        wait_d(0);
        float t[32];
        tmem_ld_x32(tl + kM_ColAcc + 32 * q, t);
        tmem_ld_wait();
        if (has_next) store_x(seq0 + kM_SlotsPerTile);   // ACC0 and XT are free: the next tile's layer 0 starts now
        float c0[2] = {0.f, 0.f}, c1[2] = {0.f, 0.f}, c2[2] = {0.f, 0.f};
        const float4* b4 = reinterpret_cast<const float4*>(b_views + 32 * q);
        const float4* r0 = reinterpret_cast<const float4*>(w_rgb + 32 * q);
        const float4* r1 = reinterpret_cast<const float4*>(w_rgb + 128 + 32 * q);
        const float4* r2 = reinterpret_cast<const float4*>(w_rgb + 256 + 32 * q);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float4 bb = b4[u], w0 = r0[u], w1 = r1[u], w2 = r2[u];
          const float v0 = fmaxf(t[4 * u] + bb.x, 0.f), v1 = fmaxf(t[4 * u + 1] + bb.y, 0.f);
          const float v2 = fmaxf(t[4 * u + 2] + bb.z, 0.f), v3 = fmaxf(t[4 * u + 3] + bb.w, 0.f);
          c0[0] = fmaf(v0, w0.x, c0[0]); c0[1] = fmaf(v1, w0.y, c0[1]); c0[0] = fmaf(v2, w0.z, c0[0]); c0[1] = fmaf(v3, w0.w, c0[1]);
          c1[0] = fmaf(v0, w1.x, c1[0]); c1[1] = fmaf(v1, w1.y, c1[1]); c1[0] = fmaf(v2, w1.z, c1[0]); c1[1] = fmaf(v3, w1.w, c1[1]);
          c2[0] = fmaf(v0, w2.x, c2[0]); c2[1] = fmaf(v1, w2.y, c2[1]); c2[0] = fmaf(v2, w2.z, c2[0]); c2[1] = fmaf(v3, w2.w, c2[1]);
        }
        if (q != 0)
          *reinterpret_cast<float4*>(PART + 4 * (128 * (q - 1) + r)) = make_float4(c0[0] + c0[1], c1[0] + c1[1], c2[0] + c2[1], alpha);
        m_epi_bar();
        if (q == 0 && valid) {
          const float4 o1 = *reinterpret_cast<const float4*>(PART + 4 * r);
          const float4 o2 = *reinterpret_cast<const float4*>(PART + 4 * (128 + r));
          const float4 o3 = *reinterpret_cast<const float4*>(PART + 4 * (256 + r));
          reinterpret_cast<float4*>(a.raw)[a.act_pid[i]] =
              make_float4(c0[0] + c0[1] + (o1.x + o2.x + o3.x) + b_tail[1], c1[0] + c1[1] + (o1.y + o2.y + o3.y) + b_tail[2],
                          c2[0] + c2[1] + (o1.z + o2.z + o3.z) + b_tail[3], alpha + (o1.w + o2.w + o3.w) + b_tail[0]);
        }
        m_epi_bar();      // PART is rewritten by the next tile
      }
      ++n_tiles;
      seq0 += kM_SlotsPerTile;
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
  if (warp == kMMmaWarp) tmem_dealloc(tm, 512);
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

// which: bit 0 = transformer (tokens -> tok0 / tok1 in the workspace), bit 1 = MLP (workspace -> raw)
static int dense_bf16_impl(const void* tokens, int32_t ld, const float* xc, int64_t count,
                           int n_views, const void* packed, size_t packed_bytes,
                           const int32_t* act_pid, int64_t first, float* raw, void* workspace,
                           void* stream, int which, const int32_t* count_dev = nullptr) {
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
  if (prof < 0) { const char* e = getenv("MPSNERF_TC_PROF"); prof = e ? atoi(e) : 0; }
  static int topt = -1;
  if (topt < 0) { const char* e = getenv("MPSNERF_T_OPT"); topt = e ? atoi(e) : 0; }
  TArgs ta{static_cast<const __half*>(tokens), ld, count, n_views, static_cast<const uint8_t*>(packed), tok0, tok1, prof, count_dev, first, topt};
  MArgs ma{tok0, tok1, xc, count, static_cast<const uint8_t*>(packed), act_pid + first, raw, prof, count_dev, first};
  const int ppt = 128 / n_views;
  const int64_t t_tiles = (count + ppt - 1) / ppt, m_tiles = (count + 127) / 128;
  auto launch = [&](auto kernel, const auto& args, size_t smem_bytes, int64_t tiles, int kc, int threads) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    int64_t ctas = (tiles + kc - 1) / kc * kc;                  // whole clusters
    const int64_t cap = (kNumSMs / kc) * kc;
    if (ctas > cap) ctas = cap;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(threads);
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
#define MPS_T_CASE(V_, C_) if ((which & 1) && n_views == V_ && cluster == C_) { \
    if (prof == 1) MPS_CUDA(launch(xformer_tc_kernel<V_, C_, true>, ta, kT_Smem, t_tiles, C_, kFThreads)); \
    else MPS_CUDA(launch(xformer_tc_kernel<V_, C_, false>, ta, kT_Smem, t_tiles, C_, kFThreads)); }
  MPS_T_CASE(2, 1) MPS_T_CASE(2, 2) MPS_T_CASE(2, 4)
  MPS_T_CASE(3, 1) MPS_T_CASE(3, 2) MPS_T_CASE(3, 4)
  MPS_T_CASE(4, 1) MPS_T_CASE(4, 2) MPS_T_CASE(4, 4)
#undef MPS_T_CASE
  if ((which & 2) && cluster == 1) MPS_CUDA(launch(mlp_tc_kernel<1>, ma, kM_Smem, m_tiles, 1, kMThreads));
  if ((which & 2) && cluster == 2) MPS_CUDA(launch(mlp_tc_kernel<2>, ma, kM_Smem, m_tiles, 2, kMThreads));
  if ((which & 2) && cluster == 4) MPS_CUDA(launch(mlp_tc_kernel<4>, ma, kM_Smem, m_tiles, 4, kMThreads));
  return MPSNERF_OK;
}

extern "C" int mpsnerf_dense_bf16(const void* tokens, int32_t ld, const float* xc, int64_t count,
                                  int n_views, const void* packed, size_t packed_bytes,
                                  const int32_t* act_pid, int64_t first, float* raw, void* workspace,
                                  void* stream) {
  return dense_bf16_impl(tokens, ld, xc, count, n_views, packed, packed_bytes, act_pid, first, raw, workspace, stream, 3);
}
extern "C" int mpsnerf_xformer_bf16(const void* tokens, int32_t ld, const float* xc, int64_t count,
                                    int n_views, const void* packed, size_t packed_bytes,
                                    const int32_t* act_pid, int64_t first, float* raw, void* workspace,
                                    void* stream) {
  return dense_bf16_impl(tokens, ld, xc, count, n_views, packed, packed_bytes, act_pid, first, raw, workspace, stream, 1);
}
// Device-side count variants: `capacity` points of buffer space starting at active-list position `first`; the number
// actually processed is clamp(*count_dev - first, 0, capacity), read on the device -- no host round trip.
extern "C" int mpsnerf_xformer_bf16_dc(const void* tokens, const float* xc, int64_t first, int64_t capacity,
                                       const int32_t* count_dev, int n_views, const void* packed, size_t packed_bytes,
                                       const int32_t* act_pid, float* raw, void* workspace, void* stream) {
  MPS_REQUIRE(count_dev != nullptr);
  return dense_bf16_impl(tokens, MPSNERF_TOKEN_LD, xc, capacity, n_views, packed, packed_bytes, act_pid, first, raw, workspace,
                         stream, 1, count_dev);
}
extern "C" int mpsnerf_mlp_bf16_dc(const void* tokens, const float* xc, int64_t first, int64_t capacity,
                                   const int32_t* count_dev, int n_views, const void* packed, size_t packed_bytes,
                                   const int32_t* act_pid, float* raw, void* workspace, void* stream) {
  MPS_REQUIRE(count_dev != nullptr);
  return dense_bf16_impl(tokens, MPSNERF_TOKEN_LD, xc, capacity, n_views, packed, packed_bytes, act_pid, first, raw, workspace,
                         stream, 2, count_dev);
}
extern "C" int mpsnerf_mlp_bf16(const void* tokens, int32_t ld, const float* xc, int64_t count,
                                int n_views, const void* packed, size_t packed_bytes,
                                const int32_t* act_pid, int64_t first, float* raw, void* workspace,
                                void* stream) {
  return dense_bf16_impl(tokens, ld, xc, count, n_views, packed, packed_bytes, act_pid, first, raw, workspace, stream, 2);
}

// Debug: copy the event trace (512 entries, see g_trace).
extern "C" int mpsnerf_debug_read_trace(unsigned long long* host_out) {
  MPS_REQUIRE(host_out != nullptr);
  MPS_CUDA(cudaMemcpyFromSymbol(host_out, mps::g_trace, sizeof(unsigned long long) * 512));
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
