// K4: multiview bilinear feature + RGB lookup -> view tokens.
//
// Replaces SpatialEncoder.index / grid_sample (lib/encoder.py:12-62, 225-253) and the RGB
// append with its 4-frequency code (lib/skinnning_batch.py:428-435).  Semantics: bilinear,
// align_corners=True, weights from the unclamped corners, indices clamped (border padding).
// One warp per (point, view): the latent is NHWC so the 128 channels of a tap are one
// coalesced 512-byte row (a float4 per lane); the 4 taps are independent 128-bit loads.
#include <cuda_fp16.h>

#include "common.cuh"

namespace mps {

constexpr int kK4Threads = 256;

struct Taps {
  int o00, o01, o10, o11;      // pixel offsets (y*W + x) of nw, ne, sw, se
  float w00, w01, w10, w11;
};

__device__ __forceinline__ Taps make_taps(float u, float v, int img_w, int img_h, int IW, int IH) {
  // uv -> [-1,1] -> source pixel, exactly the reference's sequence (encoder.py:239, :19-20)
  const float gx = 2.0f * u / (float)img_w - 1.0f;
  const float gy = 2.0f * v / (float)img_h - 1.0f;
  const float ix = ((gx + 1.0f) / 2.0f) * (float)(IW - 1);
  const float iy = ((gy + 1.0f) / 2.0f) * (float)(IH - 1);
  const float x0 = floorf(ix), y0 = floorf(iy);
  const float x1 = x0 + 1.0f, y1 = y0 + 1.0f;
  Taps t;
  t.w00 = (x1 - ix) * (y1 - iy);
  t.w01 = (ix - x0) * (y1 - iy);
  t.w10 = (x1 - ix) * (iy - y0);
  t.w11 = (ix - x0) * (iy - y0);
  const float fw = (float)(IW - 1), fh = (float)(IH - 1);
  const int cx0 = (int)fminf(fmaxf(x0, 0.f), fw), cx1 = (int)fminf(fmaxf(x1, 0.f), fw);
  const int cy0 = (int)fminf(fmaxf(y0, 0.f), fh), cy1 = (int)fminf(fmaxf(y1, 0.f), fh);
  t.o00 = cy0 * IW + cx0; t.o01 = cy0 * IW + cx1;
  t.o10 = cy1 * IW + cx0; t.o11 = cy1 * IW + cx1;
  return t;
}

__device__ __forceinline__ float4 lerp4(const float4& a, const float4& b, const float4& c, const float4& d, const Taps& t) {
  float4 r;
  r.x = a.x * t.w00 + b.x * t.w01 + c.x * t.w10 + d.x * t.w11;
  r.y = a.y * t.w00 + b.y * t.w01 + c.y * t.w10 + d.y * t.w11;
  r.z = a.z * t.w00 + b.z * t.w01 + c.z * t.w10 + d.z * t.w11;
  r.w = a.w * t.w00 + b.w * t.w01 + c.w * t.w10 + d.w * t.w11;
  return r;
}

// two floats -> packed fp16 pair, values beyond the finite fp16 range saturate to +-65504
__device__ __forceinline__ uint32_t pack_h2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// kHalf: tokens are written as fp16 (row stride ld halfs, ld % 4 == 0) for the tensor-core path, which
// halves the only large HBM stream of this kernel (the latent taps are L2 hits); values are clamped to the
// finite fp16 range.  Otherwise fp32 with row stride ld floats.
//
// The kernel is issue bound, not bandwidth bound: the tap set-up (two make_taps with their divisions,
// floors and clamps) is scalar work per row.  A warp therefore takes 32 consecutive rows at a time: lane L
// sets up row L (taps of the latent, and the complete bilinear RGB lookup), then the warp walks the 32
// rows, broadcasting one row's taps per step with shuffles and spending its 32 lanes on the 128 channels.
template <bool kHalf>
__global__ void __launch_bounds__(kK4Threads)
gather_tokens_kernel(const float* __restrict__ uv, int64_t n_rows /* count*V */, int V, const mpsnerf_frame* __restrict__ frame,
                     const float* __restrict__ latent, const float* __restrict__ img4, void* __restrict__ tokens_v, int ld,
                     const int32_t* __restrict__ count_dev, int64_t first) {
  if (count_dev != nullptr) {          // device-side active count: n_rows is only the capacity of the slab
    const int64_t n = ((int64_t)*count_dev - first) * V;
    n_rows = n < n_rows ? (n > 0 ? n : 0) : n_rows;
  }
  float* tokens = static_cast<float*>(tokens_v);
  __half* tokens_h = static_cast<__half*>(tokens_v);
  const int img_w = frame->img_w, img_h = frame->img_h, FW = frame->feat_w, FH = frame->feat_h;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * kK4Threads + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * kK4Threads) >> 5;
  // element e of the 27-wide RGB code: e<3 identity; else k=(e-3)/6, ch=(e-3)%3, cos=((e-3)%6)>=3
  const int e = lane;
  const int ch = (e < 3) ? e : ((e - 3) % 3);
  const float freq = (e < 3) ? 0.f : 3.14159265358979323846f * (float)(1 << ((e - 3) / 6));
  const float phase = (e >= 3 && ((e - 3) % 6) >= 3) ? 1.57079632679489661923f : 0.0f;
  for (int64_t base = warp0 * 32; base < n_rows; base += nwarps * 32) {
    // ---- set-up of row base + lane
    const int64_t my = min(base + lane, n_rows - 1);
    const int mv = (int)(my % V);
    const float u_ = __ldg(&uv[2 * my]), v_ = __ldg(&uv[2 * my + 1]);
    const Taps t = make_taps(u_, v_, img_w, img_h, FW, FH);
    float4 rgb;
    {
      const Taps ti = make_taps(u_, v_, img_w, img_h, img_w, img_h);
      const float4* ib = reinterpret_cast<const float4*>(img4 + (size_t)mv * img_h * img_w * 4);
      rgb = lerp4(__ldg(ib + ti.o00), __ldg(ib + ti.o01), __ldg(ib + ti.o10), __ldg(ib + ti.o11), ti);
    }
    const int nrow = (int)min((int64_t)32, n_rows - base);
    // the four tap offsets travel as one word: o00 plus the (clamped) x / y steps to the other corners
    const int tpack = t.o00 | ((t.o01 - t.o00) << 28) | ((t.o10 != t.o00 ? 1 : 0) << 29);
    int v = (int)(base % V);
    // ---- the 32 rows, one per step: latent 128 channels, 4 per lane
    for (int i = 0; i < nrow; ++i) {
      const int64_t row = base + i;
      Taps s;
      const int tp = __shfl_sync(0xffffffffu, tpack, i);
      s.o00 = tp & 0x0fffffff;
      s.o01 = s.o00 + ((tp >> 28) & 1);
      s.o10 = s.o00 + ((tp >> 29) & 1) * FW;
      s.o11 = s.o10 + ((tp >> 28) & 1);
      s.w00 = __shfl_sync(0xffffffffu, t.w00, i); s.w01 = __shfl_sync(0xffffffffu, t.w01, i);
      s.w10 = __shfl_sync(0xffffffffu, t.w10, i); s.w11 = __shfl_sync(0xffffffffu, t.w11, i);
      const float cx = __shfl_sync(0xffffffffu, rgb.x, i), cy = __shfl_sync(0xffffffffu, rgb.y, i), cz = __shfl_sync(0xffffffffu, rgb.z, i);
      const float4* lb = reinterpret_cast<const float4*>(latent + (size_t)v * FH * FW * 128) + lane;
      v = (v + 1 == V) ? 0 : v + 1;
      const float4 a = __ldg(lb + (size_t)s.o00 * 32), b = __ldg(lb + (size_t)s.o01 * 32);
      const float4 c = __ldg(lb + (size_t)s.o10 * 32), d = __ldg(lb + (size_t)s.o11 * 32);
      const float4 r = lerp4(a, b, c, d, s);
      // rgb: 27-wide code [x, sin(f0 x), cos(f0 x), ...] with cos as sin(. + fl(pi/2)) (run_nerf_helpers.py:337-353);
      // the fp16 path uses the MUFU sine: its error on |arg| < 27 (~1e-5) is far below the fp16 rounding of the token
      const float x = ch == 0 ? cx : (ch == 1 ? cy : cz);
      const float arg = fmaf(x, freq, phase);
      const float code = (e < 3) ? x : (kHalf ? __sinf(arg) : sinf(arg));
      if (kHalf) {
        __half* out_h = tokens_h + row * ld;
        reinterpret_cast<uint2*>(out_h)[lane] = make_uint2(pack_h2_sat(r.x, r.y), pack_h2_sat(r.z, r.w));
        if (128 + lane < ld) out_h[128 + lane] = __float2half_rn(lane < 27 ? code : 0.f);      // pad columns 155.. are zero
      } else {
        float* out = tokens + row * ld;
        if ((ld & 3) == 0) {
          reinterpret_cast<float4*>(out)[lane] = r;
        } else {
          out[4 * lane] = r.x; out[4 * lane + 1] = r.y; out[4 * lane + 2] = r.z; out[4 * lane + 3] = r.w;
        }
        if (128 + lane < ld) out[128 + lane] = lane < 27 ? code : 0.f;
      }
    }
  }
}

}  // namespace mps

extern "C" int mpsnerf_gather_tokens(const float* uv, int64_t count, int n_views, const mpsnerf_frame* frame,
                                     const float* latent, const float* img4, float* tokens, int32_t ld,
                                     void* stream) {
  MPS_REQUIRE(count >= 0 && n_views >= 1 && n_views <= MPSNERF_MAX_VIEWS);
  if (count == 0) return MPSNERF_OK;
  MPS_REQUIRE(uv && frame && latent && img4 && tokens);
  MPS_REQUIRE(ld >= MPSNERF_TOKEN_DIM && ld <= MPSNERF_TOKEN_LD);
  MPS_REQUIRE((reinterpret_cast<uintptr_t>(latent) & 15) == 0 && (reinterpret_cast<uintptr_t>(img4) & 15) == 0);
  MPS_REQUIRE((reinterpret_cast<uintptr_t>(tokens) & 15) == 0);
  const int64_t rows = count * n_views;
  int64_t blocks = (rows + mps::kK4Threads - 1) / mps::kK4Threads;      // a warp takes 32 rows per visit
  if (blocks > mps::kNumSMs * 16) blocks = mps::kNumSMs * 16;
  mps::gather_tokens_kernel<false><<<(int)blocks, mps::kK4Threads, 0, (cudaStream_t)stream>>>(
      uv, rows, n_views, frame, latent, img4, tokens, ld, nullptr, 0);
  MPS_LAUNCH_CHECK();
  return MPSNERF_OK;
}

static int gather_tokens_f16_impl(const float* uv, int64_t count, int n_views, const mpsnerf_frame* frame,
                                  const float* latent, const float* img4, void* tokens, const int32_t* count_dev,
                                  int64_t first, void* stream) {
  MPS_REQUIRE(count >= 0 && n_views >= 1 && n_views <= MPSNERF_MAX_VIEWS);
  if (count == 0) return MPSNERF_OK;
  MPS_REQUIRE(uv && frame && latent && img4 && tokens);
  MPS_REQUIRE((reinterpret_cast<uintptr_t>(latent) & 15) == 0 && (reinterpret_cast<uintptr_t>(img4) & 15) == 0);
  MPS_REQUIRE((reinterpret_cast<uintptr_t>(tokens) & 15) == 0);
  const int64_t rows = count * n_views;
  int64_t blocks = (rows + mps::kK4Threads - 1) / mps::kK4Threads;
  if (blocks > mps::kNumSMs * 16) blocks = mps::kNumSMs * 16;
  mps::gather_tokens_kernel<true><<<(int)blocks, mps::kK4Threads, 0, (cudaStream_t)stream>>>(
      uv, rows, n_views, frame, latent, img4, tokens, MPSNERF_TOKEN_LD, count_dev, first);
  MPS_LAUNCH_CHECK();
  return MPSNERF_OK;
}

extern "C" int mpsnerf_gather_tokens_f16(const float* uv, int64_t count, int n_views, const mpsnerf_frame* frame,
                                         const float* latent, const float* img4, void* tokens, void* stream) {
  return gather_tokens_f16_impl(uv, count, n_views, frame, latent, img4, tokens, nullptr, 0, stream);
}
extern "C" int mpsnerf_gather_tokens_f16_dc(const float* uv, int64_t first, int64_t capacity, const int32_t* count_dev,
                                            int n_views, const mpsnerf_frame* frame, const float* latent,
                                            const float* img4, void* tokens, void* stream) {
  MPS_REQUIRE(count_dev != nullptr && first >= 0);
  return gather_tokens_f16_impl(uv, capacity, n_views, frame, latent, img4, tokens, count_dev, first, stream);
}
