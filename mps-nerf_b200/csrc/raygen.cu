// K-1: ray generation + axis-aligned-box near/far, in front of K1.
//
// Replaces get_rays and get_near_far of the reference (lib/if_nerf_data_utils.py:11-25, 55-92), which run
// in numpy (float64) on the CPU once per target view.  One thread per pixel; the arithmetic is done in
// double and follows the reference's sequence (pixel -> camera through K^-1, -> world through (. - T) R,
// six plane parameters, the "exactly two of the six intersections lie on the box" rule with its 1e-6
// slack, distances as |p - o| / |d|), then rounded to fp32 -- the dataset casts to float32 at the same point.
// Output is the (N, 8) ray layout K1 consumes: o(3), d(3), near, far; rays that miss the box get
// near = 0, far = 1 and mask 0 (the full-frame convention of lib/THuman_dataset.py:719-724).
#include "common.cuh"

namespace mps {

struct RayCam {
  double kinv[9];   // inverse intrinsics, row-major
  double r[9];      // world -> camera rotation, row-major
  double t[3];
  double o[3];      // camera centre -R^T T
  double bmin[3], bmax[3];   // bounds already widened by 0.01 (ref :57)
};

__global__ void __launch_bounds__(256)
raygen_kernel(const RayCam cam, int H, int W, const int32_t* __restrict__ rows, int n_rows, float* __restrict__ rays8,
              uint8_t* __restrict__ mask_at_box) {
  // rows == nullptr: the whole H x W view; else only the listed image rows, packed in list order (ray p = pixel
  // (rows[p / W], p % W)): how one frame is dealt out to several GPUs
  const int64_t n = (int64_t)(rows ? n_rows : H) * W;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const double i = (double)(p % W), j = (double)(rows ? rows[p / W] : (int)(p / W));
    // pixel_camera = [i, j, 1] @ inv(K).T ; pixel_world = (pixel_camera - T) @ R   (ref :19-20)
    double pc[3], d[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) pc[a] = i * cam.kinv[3 * a] + j * cam.kinv[3 * a + 1] + cam.kinv[3 * a + 2] - cam.t[a];
#pragma unroll
    for (int a = 0; a < 3; ++a) d[a] = (pc[0] * cam.r[a] + pc[1] * cam.r[3 + a] + pc[2] * cam.r[6 + a]) - cam.o[a];
    // the reference hands float32 rays to get_near_far (the datasets cast first): same here
    float of[3], df[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { of[a] = (float)cam.o[a]; df[a] = (float)d[a]; }
    // get_near_far: float32 rays against float64 bounds -- numpy promotes everything after ref :57 to
    // float64, so the six plane parameters, the intersections and the distances are doubles here too
    double dd[3], od[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { dd[a] = (double)((df[a] == 0.0f) ? 1e-8f : df[a]); od[a] = (double)of[a]; }     // ref :58
    int n_in = 0;
    double dmin = 1e300, dmax = -1e300;
    const double norm_d = sqrt(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2]);
#pragma unroll
    for (int side = 0; side < 2; ++side) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const double plane = side == 0 ? cam.bmin[a] : cam.bmax[a];
        const double tpar = (plane - od[a]) / dd[a];                     // ref :59-61
        double q[3];
        bool in = true;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          q[c] = __dadd_rn(__dmul_rn(tpar, dd[c]), od[c]);                                   // ref :63 (no fma: numpy rounds the product)
          in = in && (q[c] >= cam.bmin[c] - 1e-6) && (q[c] <= cam.bmax[c] + 1e-6);
        }
        if (in) {
          ++n_in;
          const double e0 = q[0] - od[0], e1 = q[1] - od[1], e2 = q[2] - od[2];
          const double dist = sqrt(e0 * e0 + e1 * e1 + e2 * e2) / norm_d;   // ref :85-87
          dmin = fmin(dmin, dist);
          dmax = fmax(dmax, dist);
        }
      }
    }
    const bool hit = (n_in == 2);
    float* out = rays8 + 8 * p;
    reinterpret_cast<float4*>(out)[0] = make_float4(of[0], of[1], of[2], df[0]);
    reinterpret_cast<float4*>(out)[1] = make_float4(df[1], df[2], hit ? (float)dmin : 0.0f, hit ? (float)dmax : 1.0f);
    if (mask_at_box) mask_at_box[p] = hit ? 1 : 0;
  }
}

}  // namespace mps

// K, R: 3x3 row-major, T: 3, bounds: (2,3) min/max -- HOST pointers (a handful of doubles per view).
static int gen_rays_impl(const double* K, const double* R, const double* T, const double* bounds, int32_t H, int32_t W,
                         const int32_t* rows, int32_t n_rows, float* rays8, uint8_t* mask_at_box, void* stream) {
  MPS_REQUIRE(K && R && T && bounds && rays8);
  MPS_REQUIRE(H >= 1 && W >= 1 && n_rows >= 0);
  MPS_REQUIRE((reinterpret_cast<uintptr_t>(rays8) & 15) == 0);
  mps::RayCam cam;
  // inverse of K (general 3x3, adjugate / determinant)
  const double a = K[0], b = K[1], c = K[2], d = K[3], e = K[4], f = K[5], g = K[6], h = K[7], i = K[8];
  const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
  MPS_REQUIRE(det != 0.0);
  const double inv[9] = {(e * i - f * h) / det, (c * h - b * i) / det, (b * f - c * e) / det,
                         (f * g - d * i) / det, (a * i - c * g) / det, (c * d - a * f) / det,
                         (d * h - e * g) / det, (b * g - a * h) / det, (a * e - b * d) / det};
  for (int k = 0; k < 9; ++k) { cam.kinv[k] = inv[k]; cam.r[k] = R[k]; }
  for (int k = 0; k < 3; ++k) {
    cam.t[k] = T[k];
    cam.o[k] = -(R[k] * T[0] + R[3 + k] * T[1] + R[6 + k] * T[2]);     // -R^T T  (ref :13)
    cam.bmin[k] = bounds[k] - 0.01;                                      // ref :57
    cam.bmax[k] = bounds[3 + k] + 0.01;
  }
  const int64_t n = (int64_t)(rows ? n_rows : H) * W;
  if (n == 0) return MPSNERF_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > mps::kNumSMs * 16) blocks = mps::kNumSMs * 16;
  mps::raygen_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(cam, H, W, rows, n_rows, rays8, mask_at_box);
  MPS_LAUNCH_CHECK();
  return MPSNERF_OK;
}

extern "C" int mpsnerf_gen_rays(const double* K, const double* R, const double* T, const double* bounds, int32_t H,
                                int32_t W, float* rays8, uint8_t* mask_at_box, void* stream) {
  return gen_rays_impl(K, R, T, bounds, H, W, nullptr, 0, rays8, mask_at_box, stream);
}

extern "C" int mpsnerf_gen_rays_rows(const double* K, const double* R, const double* T, const double* bounds, int32_t H,
                                     int32_t W, const int32_t* rows, int32_t n_rows, float* rays8, uint8_t* mask_at_box,
                                     void* stream) {
  MPS_REQUIRE(rows != nullptr || n_rows == 0);
  if (n_rows == 0) return MPSNERF_OK;
  return gen_rays_impl(K, R, T, bounds, H, W, rows, n_rows, rays8, mask_at_box, stream);
}
