// Mesh-extraction post-step on the density grid (extract_thuman_mesh.py:125-153): for every grid point
//   occupancy = shifted_softplus(raw[..., 3])                                  (:125, run_nerf_helpers.py:18)
//   pts_mask  = (d2 to the nearest vertex) < 0.05^2                            (:132-137, knn_points K = 1)
//   outside   = dot(normalise(p - mean of the 5 nearest vertices),
//                   mean of their vertex normals) > 0                          (:148-154, knn_points K = 5)
//   occupancy = pts_mask ? occupancy : (outside ? 0 : 100)                     (:156-158)
// The reference runs two brute-force knn_points passes over 256^3 x 6890 pairs; here one pass keeps the five
// best (d2, index) pairs per point in registers while the vertices stream through shared memory in tiles.
// Distances follow the library's kNN contract: d2 = (dx*dx + dy*dy) + dz*dz in pinned
// fp32, ties -> lowest index, so K = 1 is the first of the five and the mask is bit-exact.  The grid points of
// a mesh extraction are mostly far from the body, where a cell search has no radius guarantee: brute force
// over the 6890 vertices (11 instructions per pair) is the exact and simplest form.
#include "common.cuh"

namespace mps {

constexpr int kOccThreads = 256;
constexpr int kOccTile = 2048;      // vertices per shared-memory tile (32 KB)

__device__ __forceinline__ float softplus_shifted(float x) {      // F.softplus(x - 1), beta 1, threshold 20
  const float y = x - 1.0f;
  return y > 20.f ? y : log1pf(expf(y));
}

__global__ void __launch_bounds__(kOccThreads)
occupancy_fix_kernel(const float* __restrict__ pts, int64_t n, const float* __restrict__ verts,
                     const float* __restrict__ normals, int nv, const float* __restrict__ raw, int raw_stride,
                     float* __restrict__ occ_out, int32_t* __restrict__ mask_out, uint8_t* __restrict__ outside_out,
                     int32_t* __restrict__ idx5_out, float* __restrict__ d2_out) {
  __shared__ float4 s_v[kOccTile];
  const float INF = __int_as_float(0x7f800000);
  for (int64_t base = (int64_t)blockIdx.x * kOccThreads; base < n; base += (int64_t)gridDim.x * kOccThreads) {
    const int64_t i = base + threadIdx.x;
    const bool valid = i < n;
    const int64_t ii = valid ? i : n - 1;
    const float qx = pts[3 * ii], qy = pts[3 * ii + 1], qz = pts[3 * ii + 2];
    float bd[5] = {INF, INF, INF, INF, INF};
    int bi[5] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
    for (int t0 = 0; t0 < nv; t0 += kOccTile) {
      const int tn = min(kOccTile, nv - t0);
      __syncthreads();
      for (int k = threadIdx.x; k < tn; k += kOccThreads)
        s_v[k] = make_float4(verts[3 * (t0 + k)], verts[3 * (t0 + k) + 1], verts[3 * (t0 + k) + 2], 0.f);
      __syncthreads();
#pragma unroll 4
      for (int k = 0; k < tn; ++k) {
        const float4 v = s_v[k];
        const float dx = psub(qx, v.x), dy = psub(qy, v.y), dz = psub(qz, v.z);
        const float d2 = padd(padd(pmul(dx, dx), pmul(dy, dy)), pmul(dz, dz));
        if (d2 < bd[4]) {
          // sorted insertion; vertices arrive in index order, so a strict < keeps the lowest index on ties
          const int id = t0 + k;
          bd[4] = d2; bi[4] = id;
#pragma unroll
          for (int s = 4; s > 0; --s) {
            if (bd[s] < bd[s - 1]) {
              const float td = bd[s]; bd[s] = bd[s - 1]; bd[s - 1] = td;
              const int ti = bi[s]; bi[s] = bi[s - 1]; bi[s - 1] = ti;
            }
          }
        }
      }
    }
    if (!valid) continue;
    // mean of the five neighbours / of their normals (sum in neighbour order, then / 5), :150-153
    float mx = 0.f, my = 0.f, mz = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
    const int kn = nv < 5 ? nv : 5;
    for (int s = 0; s < kn; ++s) {
      const int id = bi[s];
      mx += verts[3 * id]; my += verts[3 * id + 1]; mz += verts[3 * id + 2];
      nx += normals[3 * id]; ny += normals[3 * id + 1]; nz += normals[3 * id + 2];
    }
    const float inv = 1.0f / (float)kn;
    float px = qx - mx * inv, py = qy - my * inv, pz = qz - mz * inv;
    const float len = sqrtf(px * px + py * py + pz * pz);
    px /= len; py /= len; pz /= len;                       // 0/0 -> NaN -> "not outside", as in the reference
    const float dot = (px * (nx * inv) + py * (ny * inv)) + pz * (nz * inv);
    const bool outside = dot > 0.f;
    const bool inmask = bd[0] < kMaskThresh;
    float occ = softplus_shifted(raw[i * raw_stride + 3]);
    if (!inmask) occ = outside ? 0.f : 100.f;
    occ_out[i] = occ;
    if (mask_out) mask_out[i] = inmask ? 1 : 0;
    if (outside_out) outside_out[i] = outside ? 1 : 0;
    if (d2_out) d2_out[i] = bd[0];
    if (idx5_out) {
#pragma unroll
      for (int s = 0; s < 5; ++s) idx5_out[5 * i + s] = bi[s];
    }
  }
}

}  // namespace mps

extern "C" int mpsnerf_occupancy_fix(const float* pts, int64_t n, const float* verts, const float* normals,
                                     int32_t n_verts, const float* raw, int32_t raw_stride, float* occupancy,
                                     int32_t* pts_mask, uint8_t* outside, int32_t* idx5, float* d2_nearest,
                                     void* stream) {
  MPS_REQUIRE(n >= 0 && n_verts >= 1 && raw_stride >= 4);
  if (n == 0) return MPSNERF_OK;
  MPS_REQUIRE(pts && verts && normals && raw && occupancy);
  int64_t blocks = (n + mps::kOccThreads - 1) / mps::kOccThreads;
  if (blocks > mps::kNumSMs * 8) blocks = mps::kNumSMs * 8;
  mps::occupancy_fix_kernel<<<(int)blocks, mps::kOccThreads, 0, (cudaStream_t)stream>>>(
      pts, n, verts, normals, n_verts, raw, raw_stride, occupancy, pts_mask, outside, idx5, d2_nearest);
  MPS_LAUNCH_CHECK();
  return MPSNERF_OK;
}
