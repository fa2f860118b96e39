// Exact nearest-vertex search on a uniform grid.
//
// Replaces the brute-force pytorch3d knn_points(K=1) calls of the reference
// (lib/skinnning_batch.py:214,256,357).  Exactness:
//  * cell >= 1.01 * r, so every vertex with d2 < r^2 lies in the 27-cell neighbourhood of
//    the query's cell; the human-region mask and the argmin of masked-in points are
//    therefore exact;
//  * without a radius guarantee the 27-cell result is accepted only when its d2 is below
//    safe_r2 = (0.98*cell)^2 (anything outside the neighbourhood is at least one full cell
//    away); otherwise a warp-cooperative brute-force scan over all vertices decides.
// Distances use the pinned formula d2 = (dx*dx + dy*dy) + dz*dz, ties -> lowest index.
#pragma once
#include "common.cuh"

namespace mps {

constexpr int kGridMaxCells = MPSNERF_GRID_MAX_DIM * MPSNERF_GRID_MAX_DIM * MPSNERF_GRID_MAX_DIM;

struct GridHdr {
  float ox, oy, oz;   // origin = bbox min of the vertices
  float inv_cell;
  float cell;
  float safe_r2;
  int nx, ny, nz;
  int ncells;
  int nv;
  int pad[5];
};
static_assert(sizeof(GridHdr) == 64, "GridHdr must be 64 bytes");

// byte offsets inside the grid buffer
constexpr size_t kGridOffStart = 64;
constexpr size_t kGridOffOcc = kGridOffStart + sizeof(int) * (size_t)(kGridMaxCells + 4);
constexpr size_t kGridOffCursor = kGridOffOcc + sizeof(uint32_t) * (size_t)(kGridMaxCells / 32);
constexpr size_t kGridOffSorted = kGridOffCursor + sizeof(int) * (size_t)kGridMaxCells;

struct GridView {
  const GridHdr* hdr;
  const int* cell_start;
  const uint32_t* occ;
  const float4* sorted;
};

__host__ __device__ inline GridView grid_view(const void* buf) {
  const char* p = static_cast<const char*>(buf);
  GridView g;
  g.hdr = reinterpret_cast<const GridHdr*>(p);
  g.cell_start = reinterpret_cast<const int*>(p + kGridOffStart);
  g.occ = reinterpret_cast<const uint32_t*>(p + kGridOffOcc);
  g.sorted = reinterpret_cast<const float4*>(p + kGridOffSorted);
  return g;
}

__device__ __forceinline__ int cell_coord(float x, float o, float inv_cell) {
  return (int)floorf((x - o) * inv_cell);   // monotone in x; same formula for vertices and queries
}

__device__ __forceinline__ float dist2_pinned(float qx, float qy, float qz, float vx, float vy, float vz) {
  float dx = psub(qx, vx), dy = psub(qy, vy), dz = psub(qz, vz);
  return padd(padd(pmul(dx, dx), pmul(dy, dy)), pmul(dz, dz));
}

__device__ __forceinline__ void nn_update(float d2, int idx, float& bd2, int& bidx) {
  if (d2 < bd2 || (d2 == bd2 && idx < bidx)) {
    bd2 = d2;
    bidx = idx;
  }
}

// Is the (dilated) neighbourhood of the query's cell non-empty?  One bit test rejects most
// sample points of a frame.
__device__ __forceinline__ bool grid_maybe_near(const GridHdr& h, const uint32_t* __restrict__ occ,
                                                int cx, int cy, int cz) {
  if ((unsigned)cx >= (unsigned)h.nx || (unsigned)cy >= (unsigned)h.ny || (unsigned)cz >= (unsigned)h.nz) {
    // outside the vertex bbox: near only if within one cell of it; those border cells are
    // not in the bitmap, so fall through to the search (cheap: mostly empty rows)
    return cx >= -1 && cx <= h.nx && cy >= -1 && cy <= h.ny && cz >= -1 && cz <= h.nz;
  }
  int c = (cz * h.ny + cy) * h.nx + cx;
  return (__ldg(&occ[c >> 5]) >> (c & 31)) & 1u;
}

// Evaluate every vertex of the 27-cell neighbourhood of (cx,cy,cz): nine independent x-runs.  K1 uses this
// form: its candidates see ~10 vertices each and run 2-4 warps per block, so the search is bound by the
// latency of the dependent loads, which the independent rows overlap; the pruned, ordered variant below
// (fewer vertices, but every row waits for the previous one's result) measured 4 % slower there and 20 % faster
// in K3, where all lanes are busy and the neighbourhoods are denser.
__device__ __forceinline__ void nn_search27_all(const GridHdr& h, const int* __restrict__ cell_start,
                                                const float4* __restrict__ sorted, int cx, int cy, int cz,
                                                float qx, float qy, float qz, float& bd2, int& bidx) {
  int x0 = max(cx - 1, 0), x1 = min(cx + 1, h.nx - 1);
  if (x0 > x1) return;
  for (int z = max(cz - 1, 0); z <= min(cz + 1, h.nz - 1); ++z) {
    for (int y = max(cy - 1, 0); y <= min(cy + 1, h.ny - 1); ++y) {
      int row = (z * h.ny + y) * h.nx;
      int b = __ldg(&cell_start[row + x0]);
      int e = __ldg(&cell_start[row + x1 + 1]);
      for (int i = b; i < e; ++i) {
        float4 v = __ldg(&sorted[i]);
        nn_update(dist2_pinned(qx, qy, qz, v.x, v.y, v.z), __float_as_int(v.w), bd2, bidx);
      }
    }
  }
}

// Nearest vertex of the 27-cell neighbourhood of (cx,cy,cz), exact, with pruning: the nine x-runs (rows) are
// visited nearest first -- the query's own row, the four rows sharing a face with it, the four diagonal ones --
// and a row, or the outer cell of a row, is skipped when a lower bound of its distance to the query already
// exceeds min(best d2 so far, cap).  The lower bound is the gap to the cell's face shrunk by 1e-3 cell, which
// covers the rounding of the binning formula (~1e-6 cell) and of the pinned d2, so no vertex that could win --
// or tie, lb^2 > bound is strict -- is ever skipped: the result equals the plain scan's (the (d2, index)
// minimum does not depend on the visiting order).  cap: callers that only care about vertices with d2 < cap
// (the human-region mask of K1) pass it to prune rows beyond it; d2 >= cap results are then unspecified.
__device__ __forceinline__ void nn_search27(const GridHdr& h, const int* __restrict__ cell_start,
                                            const float4* __restrict__ sorted, int cx, int cy, int cz,
                                            float qx, float qy, float qz, float& bd2, int& bidx,
                                            float cap = __builtin_huge_valf()) {
  const float m = 1e-3f * h.cell;
  const float fx = qx - (h.ox + (float)cx * h.cell), fy = qy - (h.oy + (float)cy * h.cell),
              fz = qz - (h.oz + (float)cz * h.cell);
  float gl, gr;
  gl = fmaxf(fx - m, 0.f); gr = fmaxf(h.cell - fx - m, 0.f);
  const float gxl = gl * gl, gxr = gr * gr;
  gl = fmaxf(fy - m, 0.f); gr = fmaxf(h.cell - fy - m, 0.f);
  const float gyl = gl * gl, gyr = gr * gr;
  gl = fmaxf(fz - m, 0.f); gr = fmaxf(h.cell - fz - m, 0.f);
  const float gzl = gl * gl, gzr = gr * gr;
  // (dy, dz) + 1 of the k-th row visited, two bits each
  constexpr uint32_t kDy = 1u | (0u << 2) | (2u << 4) | (1u << 6) | (1u << 8) | (0u << 10) | (2u << 12) | (0u << 14) | (2u << 16);
  constexpr uint32_t kDz = 1u | (1u << 2) | (1u << 4) | (0u << 6) | (2u << 8) | (0u << 10) | (0u << 12) | (2u << 14) | (2u << 16);
#pragma unroll 1
  for (int k = 0; k < 9; ++k) {
    const int dy = (int)((kDy >> (2 * k)) & 3u) - 1, dz = (int)((kDz >> (2 * k)) & 3u) - 1;
    const int y = cy + dy, z = cz + dz;
    if ((unsigned)y >= (unsigned)h.ny || (unsigned)z >= (unsigned)h.nz) continue;
    const float rowd2 = (dy < 0 ? gyl : (dy > 0 ? gyr : 0.f)) + (dz < 0 ? gzl : (dz > 0 ? gzr : 0.f));
    const float bound = fminf(bd2, cap);
    if (rowd2 > bound) continue;
    const int x0 = max(cx - ((gxl + rowd2 > bound) ? 0 : 1), 0);
    const int x1 = min(cx + ((gxr + rowd2 > bound) ? 0 : 1), h.nx - 1);
    if (x0 > x1) continue;
    const int row = (z * h.ny + y) * h.nx;
    const int b = __ldg(&cell_start[row + x0]);
    const int e = __ldg(&cell_start[row + x1 + 1]);
    for (int i = b; i < e; ++i) {
      const float4 v = __ldg(&sorted[i]);
      nn_update(dist2_pinned(qx, qy, qz, v.x, v.y, v.z), __float_as_int(v.w), bd2, bidx);
    }
  }
}

// Warp-cooperative exact scan of all vertices for the lanes whose `need` flag is set.
// Must be called by all 32 lanes of a warp (converged).
__device__ __forceinline__ void nn_brute_warp(const GridHdr& h, const float4* __restrict__ sorted, bool need,
                                              float qx, float qy, float qz, float& bd2, int& bidx) {
  unsigned todo = __ballot_sync(0xffffffffu, need);
  const int lane = threadIdx.x & 31;
  while (todo) {
    int src = __ffs(todo) - 1;
    todo &= todo - 1;
    float sx = __shfl_sync(0xffffffffu, qx, src);
    float sy = __shfl_sync(0xffffffffu, qy, src);
    float sz = __shfl_sync(0xffffffffu, qz, src);
    float d = __int_as_float(0x7f800000);
    int id = 0x7fffffff;
    for (int i = lane; i < h.nv; i += 32) {
      float4 v = __ldg(&sorted[i]);
      nn_update(dist2_pinned(sx, sy, sz, v.x, v.y, v.z), __float_as_int(v.w), d, id);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float od = __shfl_xor_sync(0xffffffffu, d, o);
      int oi = __shfl_xor_sync(0xffffffffu, id, o);
      nn_update(od, oi, d, id);
    }
    if (lane == src) {
      bd2 = d;
      bidx = id;
    }
  }
}

}  // namespace mps
