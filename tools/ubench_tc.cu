// Micro-benchmarks of the sm_100a building blocks the fused kernels are designed around
// (numbers quoted in DESIGN.md section 5.1).  Stand-alone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I mps-nerf_b200/csrc -I include \
//        tools/ubench_tc.cu -o gpurun_out/ubench_tc && gpurun_out/ubench_tc
// One CTA per SM (148), 320 threads like the fused kernels; every figure is cycles (clock64) of CTA 0.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "umma.cuh"

using namespace mps::umma;

__device__ __forceinline__ void bar_epi() { named_bar_sync(1, 256); }

struct Res {
  long long ld8, ld4, st8, ldst8;
  long long mma[4];        // N = 128, 160, 192, 256: 64 back-to-back TS MMAs
  long long mma_ss256;     // 64 SS MMAs N = 256
  long long mma_ld;        // 64 TS MMAs N=256 while 8 warps stream tcgen05.ld
  long long ld_mma;        // ... and the ld side of the same experiment
  long long pingpong;      // 64 round trips commit -> epilogue wake -> arrive -> mma wake
  long long mma128_commit_each;   // 64 x (MMA N=128 + commit + wait)
};

__global__ void __launch_bounds__(320, 1) ubench(Res* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_a, bar_d, bar_x;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 9) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  if (tid == 0) { mbar_init(&bar_a, 256); mbar_init(&bar_d, 1); mbar_init(&bar_x, 1); mbar_fence_init(); }
  for (int i = tid; i < 65536 / 4; i += 320) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  const uint32_t tl = tm + ((uint32_t)((warp & 3) * 32) << 16);
  Res r{};
  const int REP = 64;
  float sink = 0.f;
  uint32_t ph = 0;            // phase of bar_x (MMA thread only)

  // ---- 1. tcgen05.ld: 8 warps, each 128 columns per rep (x32 x 4) => whole CTA reads 128 lanes x 256 cols x 4 B = 128 KB per rep
  if (warp < 8) {
    bar_epi();
    long long t0 = clock64();
    for (int it = 0; it < REP; ++it) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        tmem_ld_x32(tl + 128 * (warp >> 2) + 32 * j, v);
        tmem_ld_wait();
        sink += v[0] + v[31];
      }
    }
    bar_epi();
    r.ld8 = clock64() - t0;
  }
  __syncthreads();
  // ---- 2. tcgen05.ld: 4 warps, each 256 columns per rep (same 128 KB per rep)
  if (warp < 4) {
    named_bar_sync(2, 128);
    long long t0 = clock64();
    for (int it = 0; it < REP; ++it) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        tmem_ld_x32(tl + 32 * j, v);
        tmem_ld_wait();
        sink += v[0] + v[31];
      }
    }
    named_bar_sync(2, 128);
    r.ld4 = clock64() - t0;
  }
  __syncthreads();
  // ---- 3. tcgen05.st: 8 warps, 128 columns each per rep
  if (warp < 8) {
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0x3c003c00u + i;
    bar_epi();
    long long t0 = clock64();
    for (int it = 0; it < REP; ++it) {
#pragma unroll
      for (int j = 0; j < 4; ++j) tmem_st_u32(tl + 128 * (warp >> 2) + 32 * j, v);
      tmem_st_wait();
    }
    bar_epi();
    r.st8 = clock64() - t0;
    // ---- 4. the M-kernel half-layer epilogue shape: ld 64 cols -> relu/pack -> st 32 cols, per rep
    bar_epi();
    t0 = clock64();
    for (int it = 0; it < REP; ++it) {
      float a[32], b[32];
      tmem_ld_x32(tl + 64 * (warp >> 2), a);
      tmem_ld_x32(tl + 64 * (warp >> 2) + 32, b);
      tmem_ld_wait();
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(pk[i]) : "f"(a[2 * i + 1]), "f"(a[2 * i]));
        asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(pk[16 + i]) : "f"(b[2 * i + 1]), "f"(b[2 * i]));
      }
      tmem_st_u32(tl + 256 + 32 * (warp >> 2), pk);
      tmem_st_wait();
    }
    bar_epi();
    r.ldst8 = clock64() - t0;
  }
  __syncthreads();
  // ---- 5. MMA issue rates (TS form, A = TMEM cols 384.., B = smem), 64 back to back + one commit
  if (warp == 9 && lane == 0) {
    const int Ns[4] = {128, 160, 192, 256};
    for (int c = 0; c < 4; ++c) {
      const uint32_t idesc = instr_desc_bf16(Ns[c]);
      const uint32_t b0 = smem_u32(smem);
      long long t0 = clock64();
      for (int it = 0; it < REP; ++it)
        mma_bf16_ts(tm, tm + 384 + (it & 7) * 8, smem_desc_sw128(b0 + (it & 3) * 32), idesc, it ? 1u : 0u);
      mma_commit(&bar_x);
      mbar_wait(&bar_x, ph & 1); ++ph;
      r.mma[c] = clock64() - t0;
    }
    {
      const uint32_t idesc = instr_desc_bf16(256);
      const uint32_t b0 = smem_u32(smem);
      long long t0 = clock64();
      for (int it = 0; it < REP; ++it)
        mma_bf16_ss(tm, smem_desc_sw128(b0 + 32768 + (it & 3) * 32), smem_desc_sw128(b0 + (it & 3) * 32), idesc, it ? 1u : 0u);
      mma_commit(&bar_x);
      mbar_wait(&bar_x, ph & 1); ++ph;
      r.mma_ss256 = clock64() - t0;
    }
    {
      // one N=128 MMA + commit + wait, 64 times: the cost of a fine-grained hand-over
      const uint32_t idesc = instr_desc_bf16(128);
      const uint32_t b0 = smem_u32(smem);
      long long t0 = clock64();
      for (int it = 0; it < REP; ++it) {
        mma_bf16_ts(tm, tm + 384, smem_desc_sw128(b0), idesc, 0u);
        mma_commit(&bar_x);
        mbar_wait(&bar_x, ph & 1); ++ph;
      }
      r.mma128_commit_each = clock64() - t0;
    }
  }
  __syncthreads();
  // ---- 6. MMA (N=256 into cols 0..255) concurrent with 8 warps of tcgen05.ld on cols 256..383
  if (warp == 9 && lane == 0) {
    const uint32_t idesc = instr_desc_bf16(256);
    const uint32_t b0 = smem_u32(smem);
    long long t0 = clock64();
    for (int it = 0; it < 4 * REP; ++it)
      mma_bf16_ts(tm, tm + 384 + (it & 7) * 8, smem_desc_sw128(b0 + (it & 3) * 32), idesc, it ? 1u : 0u);
    mma_commit(&bar_x);
    mbar_wait(&bar_x, ph & 1); ++ph;
    r.mma_ld = clock64() - t0;
  } else if (warp < 8) {
    bar_epi();
    long long t0 = clock64();
    for (int it = 0; it < REP; ++it) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        tmem_ld_x32(tl + 256 + 32 * (j & 1) + 64 * (warp >> 2), v);
        tmem_ld_wait();
        sink += v[0] + v[31];
      }
    }
    bar_epi();
    r.ld_mma = clock64() - t0;
  }
  __syncthreads();
  // ---- 7. ping-pong latency: MMA thread commits (empty), 256 epilogue threads wake and arrive, MMA thread wakes
  if (warp == 9 && lane == 0) {
    long long t0 = clock64();
    for (int it = 0; it < REP; ++it) {
      mma_commit(&bar_d);
      mbar_wait(&bar_a, it & 1);
      tc_fence_after();
    }
    r.pingpong = clock64() - t0;
  } else if (warp < 8) {
    for (int it = 0; it < REP; ++it) {
      mbar_wait(&bar_d, it & 1);
      tc_fence_after();
      tc_fence_before();
      mbar_arrive(&bar_a);
    }
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    if (tid == 0) { out->ld8 = r.ld8; out->st8 = r.st8; out->ldst8 = r.ldst8; out->ld4 = r.ld4; out->ld_mma = r.ld_mma; }
    if (warp == 9 && lane == 0) {
      for (int c = 0; c < 4; ++c) out->mma[c] = r.mma[c];
      out->mma_ss256 = r.mma_ss256; out->mma_ld = r.mma_ld; out->pingpong = r.pingpong;
      out->mma128_commit_each = r.mma128_commit_each;
    }
  }
  if (sink == 123.456f) out->ld8 = 0;
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tm, 512);
}

// ---- fragment-layout probe: which (lane, column) lands in which register for the 16x256b / 16x128b shapes
__global__ void __launch_bounds__(128, 1) probe(float* out) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) { tmem_alloc(&tmem_base_s, 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16);
  {   // pattern via 32x32b: value = 1000 * lane_in_tile + column
    uint32_t v[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(1000.f * tid + c);
    tmem_st_u32(tl, v);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t r[8];
  // 16x256b.x2 at lane offset 0: 16 lanes x 16 columns
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(tl) : "memory");
  tmem_ld_wait();
  for (int i = 0; i < 8; ++i) out[(tid * 3 + 0) * 8 + i] = __uint_as_float(r[i]);
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(tl + (16u << 16)) : "memory");
  tmem_ld_wait();
  for (int i = 0; i < 8; ++i) out[(tid * 3 + 1) * 8 + i] = __uint_as_float(r[i]);
  tc_fence_before();
  __syncthreads();
  // 16x128b.x2 store at lane offset 0 (cols 32..39): register i of thread t = 100000 + 100 * t + i, read back with 32x32b
  {
    uint32_t w[4];
    for (int i = 0; i < 4; ++i) w[i] = __float_as_uint(100000.f + 100.f * lane + i);
    asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1,%2,%3,%4};" :: "r"(tl + 32), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
    for (int i = 0; i < 4; ++i) w[i] = __float_as_uint(200000.f + 100.f * lane + i);
    asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1,%2,%3,%4};" :: "r"(tl + 32 + (16u << 16)), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  {
    uint32_t q[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]) : "r"(tl + 32) : "memory");
    tmem_ld_wait();
    for (int i = 0; i < 8; ++i) out[(tid * 3 + 2) * 8 + i] = __uint_as_float(q[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 64);
}

// ---- issue-side costs of the MMA warp (whole warp converged, elected lane issues)
__global__ void __launch_bounds__(128, 1) issue_cost(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_x, bar_done[8];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  if (tid == 0) { mbar_init(&bar_x, 1); for (int i = 0; i < 8; ++i) mbar_init(&bar_done[i], 1); mbar_fence_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  if (warp == 1) {
    const uint32_t b0 = smem_u32(smem);
    uint32_t ph = 0;
    const int Ns[4] = {16, 32, 64, 128};
    for (int c = 0; c < 4; ++c) {
      const uint32_t idesc = instr_desc_bf16(Ns[c]);
      __syncwarp();
      long long t0 = clock64();
      if (elect_one()) {
#pragma unroll 4
        for (int it = 0; it < 64; ++it)
          mma_bf16_ts(tm, tm + 384 + (it & 7) * 8, smem_desc_sw128(b0 + (it & 3) * 32), idesc, it ? 1u : 0u);
      }
      __syncwarp();
      long long t1 = clock64();
      if (elect_one()) mma_commit(&bar_x);
      __syncwarp();
      mbar_wait(&bar_x, ph & 1); ++ph;
      long long t2 = clock64();
      if (tid == 32) { out[2 * c] = t1 - t0; out[2 * c + 1] = t2 - t0; }
    }
    // 64 x (try_wait on an already-completed barrier)
    {
      if (elect_one()) mma_commit(&bar_done[0]);
      __syncwarp();
      mbar_wait(&bar_done[0], 0);
      long long t0 = clock64();
      for (int it = 0; it < 64; ++it) mbar_wait(&bar_done[0], 0);
      long long t1 = clock64();
      if (tid == 32) out[8] = t1 - t0;
    }
    // 64 x (4 MMAs N=128 + commit to a ring-like barrier), no waiting: pure issue loop of the M kernel
    {
      const uint32_t idesc = instr_desc_bf16(128);
      long long t0 = clock64();
      for (int it = 0; it < 64; ++it) {
        if (elect_one()) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            mma_bf16_ts(tm, tm + 384 + k4 * 8, smem_desc_sw128(b0 + k4 * 32), idesc, 1u);
          mma_commit(&bar_done[1 + (it & 3)]);
        }
        __syncwarp();
      }
      long long t1 = clock64();
      if (elect_one()) mma_commit(&bar_x);
      __syncwarp();
      mbar_wait(&bar_x, ph & 1); ++ph;
      long long t2 = clock64();
      if (tid == 32) { out[9] = t1 - t0; out[10] = t2 - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// ---- does the weight stream (bulk copies into smem) slow the MMAs down?  148 CTAs; warp 1 issues 256 MMAs
// (N = 128, B from a 64 KB smem window) while warp 2 streams `stream_kb_per_mma` worth of bulk copies into the window.
__global__ void __launch_bounds__(128, 1) smem_contention(const uint8_t* __restrict__ gsrc, long long* out, int stream) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_x, bar_t;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  if (tid == 0) { mbar_init(&bar_x, 1); mbar_init(&bar_t, 1); mbar_fence_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  if (warp == 1) {
    const uint32_t b0 = smem_u32(smem);
    const uint32_t idesc = instr_desc_bf16(128);
    long long t0 = clock64();
    if (elect_one()) {
#pragma unroll 4
      for (int it = 0; it < 1024; ++it)
        mma_bf16_ts(tm, tm + 384 + (it & 7) * 8, smem_desc_sw128(b0 + ((it >> 2) & 3) * 16384 + (it & 3) * 32), idesc, it ? 1u : 0u);
      mma_commit(&bar_x);
    }
    __syncwarp();
    mbar_wait(&bar_x, 0);
    long long t1 = clock64();
    if (tid == 32 && blockIdx.x == 0) out[0] = t1 - t0;
  } else if (warp == 2 && stream) {
    // 256 chunks of 16 KB = the M kernel's bytes per MMA (4 KB per N=128 MMA), 4 in flight
    uint32_t ph = 0;
    long long t0 = clock64();
    for (int c = 0; c < 256; c += 4) {
      if (elect_one()) {
        mbar_arrive_expect_tx(&bar_t, 4 * 16384);
        for (int j = 0; j < 4; ++j)
          bulk_g2s(smem + 65536 + j * 16384, gsrc + (size_t)((c + j) % 87) * 16384, 16384, &bar_t);
      }
      __syncwarp();
      mbar_wait(&bar_t, ph & 1); ++ph;
    }
    long long t1 = clock64();
    if (tid == 64 && blockIdx.x == 0) out[1] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// ---- data dependence of the MMA rate: 148 CTAs x 4096 MMAs (N = 128) on zeros vs pseudo-random bf16 operands
__global__ void __launch_bounds__(128, 1) data_dep(long long* out, int random_data) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_x;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  if (tid == 0) { mbar_init(&bar_x, 1); mbar_fence_init(); }
  uint32_t rng = 0x9e3779b9u * (tid + 1) + blockIdx.x;
  auto next = [&]() { rng = rng * 1664525u + 1013904223u; const uint32_t m = random_data ? rng : 0u;
                      return (m & 0x807f807fu) | 0x3f003f00u * (random_data ? 1u : 0u); };   // two bf16 in +-[0.5, 1)
  for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = next();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16);
  {
    uint32_t v[32];
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = next();
      tmem_st_u32(tl + 384 + 32 * c, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    const uint32_t b0 = smem_u32(smem);
    const uint32_t idesc = instr_desc_bf16(128);
    long long t0 = clock64();
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    if (elect_one()) {
#pragma unroll 4
      for (int it = 0; it < 4096; ++it)
        mma_bf16_ts(tm + ((it >> 4) & 1) * 128, tm + 384 + (it & 15) * 8, smem_desc_sw128(b0 + ((it >> 2) & 3) * 16384 + (it & 3) * 32), idesc, (it & 15) ? 1u : 0u);
      mma_commit(&bar_x);
    }
    __syncwarp();
    mbar_wait(&bar_x, 0);
    long long t1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (tid == 32 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)(g1 - g0); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// ---- do warps parked in mbarrier try_wait loops slow the MMA issue down?  `spin` extra warps wait on a
// barrier that the MMA warp completes at the end.
__global__ void __launch_bounds__(640, 1) spin_effect(long long* out, int spin, int suspend_hint) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_x, bar_never;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  if (tid == 0) { mbar_init(&bar_x, 1); mbar_init(&bar_never, 1); mbar_fence_init(); }
  for (int i = tid; i < 65536 / 4; i += 640) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  if (warp == 17) {
    const uint32_t b0 = smem_u32(smem);
    const uint32_t idesc = instr_desc_bf16(128);
    long long t0 = clock64();
    for (int step = 0; step < 128; ++step) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          mma_bf16_ts(tm + (step & 1) * 128, tm + 384 + k * 8, smem_desc_sw128(b0 + (k >> 2) * 16384 + (k & 3) * 32), idesc, k ? 1u : 0u);
        mma_commit(&bar_x);
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (tid == 17 * 32 && blockIdx.x == 0) out[0] = t1 - t0;
    if (elect_one()) mbar_arrive(&bar_never);
  } else if (warp < spin) {
    if (suspend_hint) mbar_wait(&bar_never, 0);
    else { while (!mbar_test(&bar_never, 0)) { __nanosleep(64); } }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// ---- bursts: B MMAs (N = 128) + commit + wait for completion, repeated; and "B MMAs + commit" bursts separated by an idle gap
__global__ void __launch_bounds__(640, 1) burst_cost(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_x;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  if (tid == 0) { mbar_init(&bar_x, 1); mbar_fence_init(); }
  for (int i = tid; i < 65536 / 4; i += 640) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  if (warp == 17) {
    const uint32_t b0 = smem_u32(smem);
    const uint32_t idesc = instr_desc_bf16(128);
    uint32_t ph = 0;
    int slot = 0;
    for (int B = 8; B <= 32; B *= 2) {
      long long t_issue = 0, t_total = 0;
      for (int rep = 0; rep < 32; ++rep) {
        long long t0 = clock64();
        if (elect_one()) {
          for (int k = 0; k < B; ++k)
            mma_bf16_ts(tm, tm + 384 + (k & 7) * 8, smem_desc_sw128(b0 + ((k >> 2) & 3) * 16384 + (k & 3) * 32), idesc, k ? 1u : 0u);
          mma_commit(&bar_x);
        }
        __syncwarp();
        long long t1 = clock64();
        mbar_wait(&bar_x, ph & 1); ++ph;
        long long t2 = clock64();
        t_issue += t1 - t0; t_total += t2 - t0;
      }
      if (tid == 17 * 32 && blockIdx.x == 0) { out[slot] = t_issue / 32; out[slot + 1] = t_total / 32; }
      slot += 2;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  {
    long long* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
    cudaFuncSetAttribute(burst_cost, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    burst_cost<<<148, 640, 65536>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("burst_cost CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[8]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int i = 0, B = 8; i < 3; ++i, B *= 2)
      printf("burst of %2d MMAs N=128 on an idle pipe: issue %lld cyc, until complete %lld cyc (exec floor %d)\n", B, h[2 * i], h[2 * i + 1], 64 * B);
  }
  {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(spin_effect, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int hint = 1; hint >= 0; --hint) for (int spin : {0, 4, 16}) {
      cudaMemset(d, 0, 16);
      spin_effect<<<148, 640, 65536>>>(d, spin, hint);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("spin_effect CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
      long long h[2]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      printf("1024 MMAs N=128 in steps of 8, %2d warps parked in %s: %.1f cyc/MMA\n", spin, hint ? "try_wait loops   " : "test_wait+nanosleep", h[0] / 1024.0);
    }
  }
  {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(data_dep, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int rd = 0; rd < 2; ++rd) for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(d, 0, 16);
      data_dep<<<148, 128, 65536>>>(d, rd);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("data_dep CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
      long long h[2]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      printf("4096 MMAs N=128 on %s data: %.1f cyc/MMA, %.1f ns/MMA (=> %.0f MHz)\n", rd ? "RANDOM" : "zero  ", h[0] / 4096.0, h[1] / 4096.0,
             1e3 * h[0] / (double)h[1]);
    }
  }
  {
    uint8_t* g; cudaMalloc(&g, 87 * 16384); cudaMemset(g, 0, 87 * 16384);
    long long* d; cudaMalloc(&d, 16); 
    cudaFuncSetAttribute(smem_contention, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    for (int stream = 0; stream < 2; ++stream) {
      cudaMemset(d, 0, 16);
      smem_contention<<<148, 128, 131072>>>(g, d, stream);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("smem_contention CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
      long long h[2]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      printf("1024 MMAs N=128 %s weight stream: %.1f cyc/MMA; stream of 4 MB took %lld cyc (%.1f B/cyc/SM)\n", stream ? "WITH" : "without",
             h[0] / 1024.0, h[1], h[1] ? 4194304.0 / h[1] : 0.0);
    }
  }
  {
    long long* d; cudaMalloc(&d, 16 * 8); cudaMemset(d, 0, 128);
    cudaFuncSetAttribute(issue_cost, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    issue_cost<<<1, 128, 65536>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("issue_cost CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const int Ns[4] = {16, 32, 64, 128};
    for (int c = 0; c < 4; ++c) printf("64 MMAs N=%3d: issue loop %.1f cyc/MMA, until complete %.1f cyc/MMA (exec floor %d)\n", Ns[c], h[2 * c] / 64.0, h[2 * c + 1] / 64.0, Ns[c] / 2);
    printf("mbar try_wait on a completed barrier: %.1f cyc\n", h[8] / 64.0);
    printf("chunk loop (4 MMA N=128 + commit): issue %.1f cyc/chunk, complete %.1f cyc/chunk (exec floor 256)\n", h[9] / 64.0, h[10] / 64.0);
  }
  {
    float* d; cudaMalloc(&d, 128 * 24 * 4);
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    static float h[128 * 24];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int t : {0, 1, 2, 3, 4, 5, 31, 32, 37}) {
      printf("thread %3d  ld16x256b.x2 lane0 :", t); for (int i = 0; i < 8; ++i) printf(" %7.0f", h[(t * 3) * 8 + i]); printf("\n");
      printf("            ld16x256b.x2 lane16:"); for (int i = 0; i < 8; ++i) printf(" %7.0f", h[(t * 3 + 1) * 8 + i]); printf("\n");
      printf("            32x32b readback of st16x128b.x2 (cols 32..39):"); for (int i = 0; i < 8; ++i) printf(" %7.0f", h[(t * 3 + 2) * 8 + i]); printf("\n");
    }
    for (int t : {8, 16, 24}) { printf("thread %3d  readback:", t); for (int i = 0; i < 8; ++i) printf(" %7.0f", h[(t * 3 + 2) * 8 + i]); printf("\n"); }
  }
  Res* d;
  cudaMalloc(&d, sizeof(Res));
  cudaFuncSetAttribute(ubench, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int rep = 0; rep < 2; ++rep) {
    ubench<<<148, 320, 65536>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  }
  Res h;
  cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost);
  const double R = 64.0;
  printf("tcgen05.ld  8 warps: %.0f cyc per 128 KB (128x256 fp32)  => %.1f B/cyc/SM\n", h.ld8 / R, 131072.0 / (h.ld8 / R));
  printf("tcgen05.ld  4 warps: %.0f cyc per 128 KB                 => %.1f B/cyc/SM\n", h.ld4 / R, 131072.0 / (h.ld4 / R));
  printf("tcgen05.st  8 warps: %.0f cyc per 128 KB                 => %.1f B/cyc/SM\n", h.st8 / R, 131072.0 / (h.st8 / R));
  printf("ld64+cvt.relu+st32 (8 warps, 128 rows x 128 cols): %.0f cyc per half-layer\n", h.ldst8 / R);
  const int Ns[4] = {128, 160, 192, 256};
  for (int c = 0; c < 4; ++c) printf("TS MMA M=128 N=%d K=16: %.1f cyc each (floor %d)\n", Ns[c], h.mma[c] / R, Ns[c] / 2);
  printf("SS MMA M=128 N=256 K=16: %.1f cyc each\n", h.mma_ss256 / R);
  printf("TS MMA N=128 + commit + wait each: %.1f cyc\n", h.mma128_commit_each / R);
  printf("TS MMA N=256 under 8-warp tcgen05.ld load: %.1f cyc each; ld side: %.0f cyc per 128 KB\n", h.mma_ld / (4 * R), h.ld_mma / R);
  printf("commit -> 256-thread wake -> arrive -> MMA-thread wake: %.0f cyc per round trip\n", h.pingpong / R);
  return 0;
}
