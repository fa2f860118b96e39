// Backward of the two kernels either side of the dense stage, for the training step (BASELINE config 4;
// run_nerf_batch.py:544-570: render -> img2mse(rgb) + img2mse(acc) -> backward):
//   K6 backward: d(rgb_map, acc_map) -> d raw          (raw2outputs, run_nerf_batch.py:369-398)
//   K4 backward: d tokens -> d latent (NHWC), a scatter-add of the bilinear taps
//                                                      (SpatialEncoder.index / grid_sample, lib/encoder.py:12-62, 225-253)
// Nothing upstream of the canonical points carries a parameter under the shipped configs (skinning_field =
// correction_field = 0), so K1 / K3 have no backward; the RGB half of a token is a function of the input images only.
#include "common.cuh"

namespace mps {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// One thread per ray (a training batch is ~1 K rays): two passes over the S samples.
//   w_s = alpha_s T_s,  T_s = prod_{j<s} (1 - alpha_j + 1e-10),  rgb_map = sum w_s c_s,  acc = sum w_s
//   G_s = c_s . d_rgb + d_acc   (gradient w.r.t. w_s)
//   dL/dalpha_s = T_s G_s - (sum_{j>s} w_j G_j) / (1 - alpha_s + 1e-10)
// Samples whose raw is the -80 fill of masked-out points (lib/skinnning_batch.py:493) are constants: gradient 0.
__global__ void __launch_bounds__(128)
composite_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ rays, int64_t n_rays, int S,
                     const float* __restrict__ t_vals, const float* __restrict__ u, const float* __restrict__ z_vals,
                     int occupancy, const float* __restrict__ d_rgb, const float* __restrict__ d_acc,
                     float* __restrict__ d_raw) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  const float* ray = rays + 8 * r;
  const float near = ray[6], far = ray[7];
  const float dn = sqrtf(ray[3] * ray[3] + ray[4] * ray[4] + ray[5] * ray[5]);
  const float* u_row = u ? u + r * S : nullptr;
  const float* z_row = z_vals ? z_vals + r * S : nullptr;
  const float4* raw4 = reinterpret_cast<const float4*>(raw) + r * S;
  float4* out4 = reinterpret_cast<float4*>(d_raw) + r * S;
  const float gr = d_rgb[3 * r], gg = d_rgb[3 * r + 1], gb = d_rgb[3 * r + 2], ga = d_acc ? d_acc[r] : 0.f;
  auto alpha_of = [&](int s, const float4& v, float& dalpha_draw) -> float {
    if (!occupancy) {
      const float z = z_row ? z_row[s] : sample_z(near, far, t_vals, s, S, u_row);
      const float zn = (s == S - 1) ? 0.f : (z_row ? z_row[s + 1] : sample_z(near, far, t_vals, s + 1, S, u_row));
      const float dist = ((s == S - 1) ? 1e10f : (zn - z)) * dn;
      const float x = v.w - 1.f;
      const float sp = x > 20.f ? x : log1pf(expf(x));           // F.softplus (beta 1, threshold 20)
      const float e = expf(-sp * dist);                          // 1 - alpha
      dalpha_draw = e * dist * (x > 20.f ? 1.f : sigmoidf_(x));
      return 1.f - e;
    }
    const float sg = sigmoidf_(v.w);
    dalpha_draw = 1.0002f * sg * (1.f - sg);
    return 1.0002f * sg - 0.0001f;
  };
  // pass 1: total = sum_j w_j G_j
  float T = 1.f, total = 0.f;
  for (int s = 0; s < S; ++s) {
    const float4 v = __ldg(raw4 + s);
    if (!occupancy && v.w == -80.f) continue;                    // alpha exactly 0: w = 0, T unchanged
    float da;
    const float a = alpha_of(s, v, da);
    const float w = a * T;
    const float G = (1.0002f * sigmoidf_(v.x) - 0.0001f) * gr + (1.0002f * sigmoidf_(v.y) - 0.0001f) * gg +
                    (1.0002f * sigmoidf_(v.z) - 0.0001f) * gb + ga;
    total = fmaf(w, G, total);
    T *= (1.f - a + 1e-10f);
  }
  // pass 2
  T = 1.f;
  float prefix = 0.f;
  for (int s = 0; s < S; ++s) {
    const float4 v = __ldg(raw4 + s);
    if (!occupancy && v.w == -80.f) { out4[s] = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
    float da;
    const float a = alpha_of(s, v, da);
    const float w = a * T;
    const float sx = sigmoidf_(v.x), sy = sigmoidf_(v.y), sz = sigmoidf_(v.z);
    const float G = (1.0002f * sx - 0.0001f) * gr + (1.0002f * sy - 0.0001f) * gg + (1.0002f * sz - 0.0001f) * gb + ga;
    prefix = fmaf(w, G, prefix);
    const float dLda = T * G - (total - prefix) / (1.f - a + 1e-10f);
    out4[s] = make_float4(w * gr * 1.0002f * sx * (1.f - sx), w * gg * 1.0002f * sy * (1.f - sy),
                          w * gb * 1.0002f * sz * (1.f - sz), dLda * da);
    T *= (1.f - a + 1e-10f);
  }
}

// One warp per (point, view) row: the 128 latent channels of d_tokens are scattered to the four taps with the
// forward's bilinear weights (weights from the unclamped corners, indices clamped: border padding).
__global__ void __launch_bounds__(256)
gather_tokens_bwd_kernel(const float* __restrict__ uv, int64_t n_rows, int V, const mpsnerf_frame* __restrict__ frame,
                         const float* __restrict__ d_tokens, int ld, float* __restrict__ d_latent) {
  const int img_w = frame->img_w, img_h = frame->img_h, FW = frame->feat_w, FH = frame->feat_h;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp0; row < n_rows; row += nwarps) {
    const int v = (int)(row % V);
    const float u_ = uv[2 * row], v_ = uv[2 * row + 1];
    const float gx = 2.0f * u_ / (float)img_w - 1.0f, gy = 2.0f * v_ / (float)img_h - 1.0f;
    const float ix = ((gx + 1.0f) / 2.0f) * (float)(FW - 1), iy = ((gy + 1.0f) / 2.0f) * (float)(FH - 1);
    const float x0 = floorf(ix), y0 = floorf(iy), x1 = x0 + 1.0f, y1 = y0 + 1.0f;
    const float w00 = (x1 - ix) * (y1 - iy), w01 = (ix - x0) * (y1 - iy), w10 = (x1 - ix) * (iy - y0), w11 = (ix - x0) * (iy - y0);
    const float fw = (float)(FW - 1), fh = (float)(FH - 1);
    const int cx0 = (int)fminf(fmaxf(x0, 0.f), fw), cx1 = (int)fminf(fmaxf(x1, 0.f), fw);
    const int cy0 = (int)fminf(fmaxf(y0, 0.f), fh), cy1 = (int)fminf(fmaxf(y1, 0.f), fh);
    float* base = d_latent + (size_t)v * FH * FW * 128;
    float* p00 = base + (size_t)(cy0 * FW + cx0) * 128;
    float* p01 = base + (size_t)(cy0 * FW + cx1) * 128;
    float* p10 = base + (size_t)(cy1 * FW + cx0) * 128;
    float* p11 = base + (size_t)(cy1 * FW + cx1) * 128;
    const float* g = d_tokens + row * ld;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = lane + 32 * k;
      const float gv = g[c];
      atomicAdd(p00 + c, w00 * gv); atomicAdd(p01 + c, w01 * gv);
      atomicAdd(p10 + c, w10 * gv); atomicAdd(p11 + c, w11 * gv);
    }
  }
}

// out4 (count, 4) <-> raw[act_pid[i]] helpers for the training path
__global__ void gather_rows4_kernel(const float* __restrict__ src, const int32_t* __restrict__ act_pid, int64_t count,
                                    float* __restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[act_pid[i]];
}
__global__ void scatter_rows4_kernel(const float* __restrict__ src, const int32_t* __restrict__ act_pid, int64_t count,
                                     float* __restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    reinterpret_cast<float4*>(dst)[act_pid[i]] = reinterpret_cast<const float4*>(src)[i];
}

}  // namespace mps

extern "C" int mpsnerf_composite_bwd(const float* raw, const float* rays, int64_t n_rays, int32_t S, const float* t_vals,
                                     const float* u, const float* z_vals, int occupancy, const float* d_rgb,
                                     const float* d_acc, float* d_raw, void* stream) {
  MPS_REQUIRE(n_rays >= 0 && S >= 1);
  if (n_rays == 0) return MPSNERF_OK;
  MPS_REQUIRE(raw && rays && (t_vals || z_vals) && d_rgb && d_raw);
  MPS_REQUIRE((reinterpret_cast<uintptr_t>(raw) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_raw) & 15) == 0);
  mps::composite_bwd_kernel<<<(unsigned)((n_rays + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      raw, rays, n_rays, S, t_vals, u, z_vals, occupancy, d_rgb, d_acc, d_raw);
  MPS_LAUNCH_CHECK();
  return MPSNERF_OK;
}

extern "C" int mpsnerf_gather_tokens_bwd(const float* uv, int64_t count, int n_views, const mpsnerf_frame* frame,
                                         const float* d_tokens, int32_t ld, float* d_latent, void* stream) {
  MPS_REQUIRE(count >= 0 && n_views >= 1 && n_views <= MPSNERF_MAX_VIEWS && ld >= 128);
  if (count == 0) return MPSNERF_OK;
  MPS_REQUIRE(uv && frame && d_tokens && d_latent);
  const int64_t rows = count * n_views;
  int64_t blocks = (rows * 32 + 255) / 256;
  if (blocks > mps::kNumSMs * 16) blocks = mps::kNumSMs * 16;
  mps::gather_tokens_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(uv, rows, n_views, frame, d_tokens, ld,
                                                                                 d_latent);
  MPS_LAUNCH_CHECK();
  return MPSNERF_OK;
}

extern "C" int mpsnerf_rows4_gather(const float* src, const int32_t* act_pid, int64_t count, float* dst, void* stream) {
  MPS_REQUIRE(count >= 0);
  if (count == 0) return MPSNERF_OK;
  MPS_REQUIRE(src && act_pid && dst);
  mps::gather_rows4_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, act_pid, count, dst);
  MPS_LAUNCH_CHECK();
  return MPSNERF_OK;
}

extern "C" int mpsnerf_rows4_scatter(const float* src, const int32_t* act_pid, int64_t count, float* dst, void* stream) {
  MPS_REQUIRE(count >= 0);
  if (count == 0) return MPSNERF_OK;
  MPS_REQUIRE(src && act_pid && dst);
  mps::scatter_rows4_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, act_pid, count, dst);
  MPS_LAUNCH_CHECK();
  return MPSNERF_OK;
}

// ------------------------------------------------------------------ the gradient all-reduce of data-parallel training
// The one collective of the path (run_nerf_batch.py:344-348: DistributedDataParallel around the network).  The
// communicator belongs to the host (a C / C++ trainer creates it with ncclCommInitRank; the Python mirror uses
// torch.distributed instead and never calls this).  The library does not link NCCL: ncclAllReduce is bound at first
// use from the libnccl the process has already loaded, else from libnccl.so.2 on the loader path.
#include <dlfcn.h>
#include <nccl.h>

namespace mps {
using AllReduceFn = ncclResult_t (*)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
using ErrStrFn = const char* (*)(ncclResult_t);
static AllReduceFn g_allreduce = nullptr;
static ErrStrFn g_errstr = nullptr;

static bool bind_nccl() {
  if (g_allreduce) return true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return false;
  g_errstr = reinterpret_cast<ErrStrFn>(dlsym(h, "ncclGetErrorString"));
  g_allreduce = reinterpret_cast<AllReduceFn>(dlsym(h, "ncclAllReduce"));
  return g_allreduce != nullptr;
}
}  // namespace mps

extern "C" int mpsnerf_allreduce_mean(void* nccl_comm, float* bucket, int64_t count, void* stream) {
  MPS_REQUIRE(count >= 0);
  if (count == 0) return MPSNERF_OK;
  MPS_REQUIRE(nccl_comm && bucket);
  if (!mps::bind_nccl()) {
    mps::set_error("%s", "mpsnerf_allreduce_mean: libnccl.so.2 not found (dlopen)");
    return MPSNERF_ENCCL;
  }
  // in place, averaged by the collective itself (ncclAvg: with NVLS the sum and the scale happen in the switch)
  const ncclResult_t r = mps::g_allreduce(bucket, bucket, (size_t)count, ncclFloat32, ncclAvg,
                                          static_cast<ncclComm_t>(nccl_comm), (cudaStream_t)stream);
  if (r != ncclSuccess) {
    mps::set_error("mpsnerf_allreduce_mean: ncclAllReduce: %s", mps::g_errstr ? mps::g_errstr(r) : "failed");
    return MPSNERF_ENCCL;
  }
  return MPSNERF_OK;
}
