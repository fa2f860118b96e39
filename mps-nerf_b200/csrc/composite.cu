// K6: alpha compositing (raw2outputs, run_nerf_batch.py:369-398).  One warp per ray.
//
// Each lane owns a contiguous run of S/32 samples, forms its local transmittance product,
// the warp does an exclusive multiplicative scan over lanes with shuffles, and the sums
// (rgb, depth, acc) are warp-reduced.  z is regenerated from (near, far, t, u) with the same
// pinned formula K1 used, so no z buffer is ever stored.
#include "common.cuh"

namespace mps {

constexpr int kK6Threads = 256;
constexpr int kMaxPerLane = 8;   // S <= 256

__device__ __forceinline__ float softplus_torch(float x) {   // F.softplus, beta=1, threshold=20
  return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float wide_sigmoid(float x) {     // run_nerf_helpers.py:19
  return 1.0002f * (1.0f / (1.0f + expf(-x))) - 0.0001f;
}

// kPer = samples per lane, a compile-time bound so that the per-lane loops carry no dead iterations
// (S = 64 -> 2); the kernel handles every S <= 32 kPer.
template <int kPer>
__global__ void __launch_bounds__(kK6Threads)
composite_kernel(const float* __restrict__ raw, const float* __restrict__ rays, int64_t n_rays, int S,
                 const float* __restrict__ t_vals, const float* __restrict__ u, const float* __restrict__ z_vals,
                 int occupancy, float* __restrict__ rgb_out, float* __restrict__ disp_out, float* __restrict__ acc_out,
                 float* __restrict__ depth_out, float* __restrict__ w_out, float* __restrict__ ts_out) {
  const int lane = threadIdx.x & 31;
  const int per = (S + 31) / 32;
  const int64_t warp0 = ((int64_t)blockIdx.x * kK6Threads + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * kK6Threads) >> 5;
  for (int64_t r = warp0; r < n_rays; r += nwarps) {
    const float* ray = rays + 8 * r;
    const float near = ray[6], far = ray[7];
    const float dn = sqrtf(ray[3] * ray[3] + ray[4] * ray[4] + ray[5] * ray[5]);
    const float* u_row = u ? u + r * S : nullptr;
    const float* z_row = z_vals ? z_vals + r * S : nullptr;
    const float4* raw4 = reinterpret_cast<const float4*>(raw) + r * S;
    float alpha[kPer], zs[kPer];
    float4 c[kPer];
    float prod = 1.f;
    const int s0 = lane * per;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int s = s0 + k;
      alpha[k] = 0.f; zs[k] = 0.f; c[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < per && s < S) {
        const float4 v = __ldg(raw4 + s);
        float a, z = 0.f;
        if (!occupancy && v.w == -80.f) {
          // the fill value of masked-out samples (lib/skinnning_batch.py:493), ~94 % of a frame:
          // softplus(-81) = 6.6e-36, so alpha = 1 - exp(-6.6e-36 * dist) is exactly 0 for every dist <= 1e10 * |d|
          // -- a strength reduction, not an approximation; the weight, colour, depth and transmittance terms
          // vanish with it, so not even z is needed
          a = 0.f;
        } else {
          z = z_row ? z_row[s] : sample_z(near, far, t_vals, s, S, u_row);
          if (!occupancy) {
            const float zn = (s == S - 1) ? 0.f : (z_row ? z_row[s + 1] : sample_z(near, far, t_vals, s + 1, S, u_row));
            const float dist = (s == S - 1) ? 1e10f : (zn - z);
            a = 1.f - expf(-softplus_torch(v.w - 1.f) * (dist * dn));
          } else {
            a = wide_sigmoid(v.w);
          }
        }
        alpha[k] = a; zs[k] = z; c[k] = v;
        prod *= (1.f - a + 1e-10f);
      }
    }
    {
      // rays on which no sample absorbs (every alpha exactly 0; most rays of a frame miss the body):
      // weights 0, transmittance 1, sums 0, disp = 1 / max(1e-10, 0/0) = NaN -- what the general path produces
      bool none = true;
#pragma unroll
      for (int k = 0; k < kPer; ++k) none = none && (alpha[k] == 0.f);
      if (__all_sync(0xffffffffu, none)) {
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
          const int s = s0 + k;
          if (k < per && s < S) {
            if (w_out) w_out[r * S + s] = 0.f;
            if (ts_out) ts_out[r * S + s] = 1.f;
          }
        }
        if (lane == 0) {
          rgb_out[3 * r] = 0.f; rgb_out[3 * r + 1] = 0.f; rgb_out[3 * r + 2] = 0.f;
          acc_out[r] = 0.f;
          if (depth_out) depth_out[r] = 0.f;
          disp_out[r] = __int_as_float(0x7fffffff);
        }
        continue;
      }
    }
    // exclusive product scan across lanes
    float incl = prod;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl *= t;
    }
    float T = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) T = 1.f;
    float sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f, sa = 0.f;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int s = s0 + k;
      if (k < per && s < S) {
        const float w = alpha[k] * T;
        if (alpha[k] != 0.f) {          // w == 0 adds nothing and 1 - 0 + 1e-10 == 1 in fp32
          sr += w * wide_sigmoid(c[k].x);
          sg += w * wide_sigmoid(c[k].y);
          sb += w * wide_sigmoid(c[k].z);
          sd += w * zs[k];
          sa += w;
        }
        if (w_out) w_out[r * S + s] = w;
        if (ts_out) ts_out[r * S + s] = T;
        if (alpha[k] != 0.f) T *= (1.f - alpha[k] + 1e-10f);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sr += __shfl_xor_sync(0xffffffffu, sr, o);
      sg += __shfl_xor_sync(0xffffffffu, sg, o);
      sb += __shfl_xor_sync(0xffffffffu, sb, o);
      sd += __shfl_xor_sync(0xffffffffu, sd, o);
      sa += __shfl_xor_sync(0xffffffffu, sa, o);
    }
    if (lane == 0) {
      rgb_out[3 * r] = sr; rgb_out[3 * r + 1] = sg; rgb_out[3 * r + 2] = sb;
      acc_out[r] = sa;
      if (depth_out) depth_out[r] = sd;
      // 1/max(1e-10, depth/acc); torch.max propagates the NaN of 0/0 (empty rays) -> disp = NaN
      const float ratio = sd / sa;
      disp_out[r] = (ratio != ratio) ? ratio : 1.f / fmaxf(1e-10f, ratio);
    }
  }
}

}  // namespace mps

extern "C" int mpsnerf_composite(const float* raw, const float* rays, int64_t n_rays, int32_t S,
                                 const float* t_vals, const float* u, const float* z_vals, int occupancy,
                                 float* rgb, float* disp, float* acc, float* depth, float* weights,
                                 float* trans, void* stream) {
  MPS_REQUIRE(n_rays >= 0 && S >= 1 && S <= 32 * mps::kMaxPerLane);
  if (n_rays == 0) return MPSNERF_OK;
  MPS_REQUIRE(raw && rays && (t_vals || z_vals) && rgb && disp && acc);
  MPS_REQUIRE((reinterpret_cast<uintptr_t>(raw) & 15) == 0);
  int64_t blocks = (n_rays * 32 + mps::kK6Threads - 1) / mps::kK6Threads;
  if (blocks > mps::kNumSMs * 16) blocks = mps::kNumSMs * 16;
  const int per = (S + 31) / 32;
#define MPS_K6(P) mps::composite_kernel<P><<<(int)blocks, mps::kK6Threads, 0, (cudaStream_t)stream>>>( \
      raw, rays, n_rays, S, t_vals, u, z_vals, occupancy, rgb, disp, acc, depth, weights, trans)
  if (per <= 1) MPS_K6(1); else if (per <= 2) MPS_K6(2); else if (per <= 4) MPS_K6(4); else MPS_K6(8);
#undef MPS_K6
  MPS_LAUNCH_CHECK();
  return MPSNERF_OK;
}
