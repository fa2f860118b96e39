// Build the uniform vertex grid (one CTA; 6890 vertices) and the diagnostic exact kNN entry.
#include "grid.cuh"

namespace mps {

__device__ __forceinline__ float3 load_vertex(const float* __restrict__ verts, int i, const float* Th,
                                              const float* R) {
  float x = verts[3 * i], y = verts[3 * i + 1], z = verts[3 * i + 2];
  if (Th != nullptr) {
    // v' = (v - Th) @ R, pinned: ((d0*R0k + d1*R1k) + d2*R2k)   (lib/skinnning_batch.py:355-356)
    float d0 = psub(x, Th[0]), d1 = psub(y, Th[1]), d2 = psub(z, Th[2]);
    x = padd(padd(pmul(d0, R[0]), pmul(d1, R[3])), pmul(d2, R[6]));
    y = padd(padd(pmul(d0, R[1]), pmul(d1, R[4])), pmul(d2, R[7]));
    z = padd(padd(pmul(d0, R[2]), pmul(d1, R[5])), pmul(d2, R[8]));
  }
  return make_float3(x, y, z);
}

constexpr int kBuildThreads = 1024;
constexpr int kBuildSmemInts = 51200;      // 200 KB of dynamic shared memory

__global__ void __launch_bounds__(kBuildThreads, 1)
grid_build_kernel(const float* __restrict__ verts, int nv, const float* __restrict__ Th,
                  const float* __restrict__ R, float cell_req, char* buf) {
  GridHdr* hdr = reinterpret_cast<GridHdr*>(buf);
  int* cell_start = reinterpret_cast<int*>(buf + kGridOffStart);
  uint32_t* occ = reinterpret_cast<uint32_t*>(buf + kGridOffOcc);
  int* cursor = reinterpret_cast<int*>(buf + kGridOffCursor);
  float4* sorted = reinterpret_cast<float4*>(buf + kGridOffSorted);

  __shared__ float s_red[6][32];
  __shared__ GridHdr s_hdr;
  __shared__ int s_scan[kBuildThreads / 32];
  extern __shared__ int s_cells[];        // kBuildSmemInts ints: [cursor | cell_start] when the grid is small enough
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

  // 1. bounding box
  float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
  for (int i = tid; i < nv; i += kBuildThreads) {
    float3 v = load_vertex(verts, i, Th, R);
    lo[0] = fminf(lo[0], v.x); hi[0] = fmaxf(hi[0], v.x);
    lo[1] = fminf(lo[1], v.y); hi[1] = fmaxf(hi[1], v.y);
    lo[2] = fminf(lo[2], v.z); hi[2] = fmaxf(hi[2], v.z);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
    }
    if (lane == 0) { s_red[k][wid] = lo[k]; s_red[3 + k][wid] = hi[k]; }
  }
  __syncthreads();
  if (tid == 0) {
    float l[3], h[3];
    for (int k = 0; k < 3; ++k) {
      l[k] = s_red[k][0]; h[k] = s_red[3 + k][0];
      for (int w = 1; w < kBuildThreads / 32; ++w) { l[k] = fminf(l[k], s_red[k][w]); h[k] = fmaxf(h[k], s_red[3 + k][w]); }
    }
    float ext = fmaxf(fmaxf(h[0] - l[0], h[1] - l[1]), h[2] - l[2]);
    float cell = fmaxf(cell_req, ext / (MPSNERF_GRID_MAX_DIM - 0.5f));
    GridHdr g;
    g.ox = l[0]; g.oy = l[1]; g.oz = l[2];
    g.cell = cell;
    g.inv_cell = 1.0f / cell;
    g.safe_r2 = (0.98f * cell) * (0.98f * cell);
    g.nx = min(MPSNERF_GRID_MAX_DIM, cell_coord(h[0], l[0], g.inv_cell) + 1);
    g.ny = min(MPSNERF_GRID_MAX_DIM, cell_coord(h[1], l[1], g.inv_cell) + 1);
    g.nz = min(MPSNERF_GRID_MAX_DIM, cell_coord(h[2], l[2], g.inv_cell) + 1);
    g.ncells = g.nx * g.ny * g.nz;
    g.nv = nv;
    for (int k = 0; k < 5; ++k) g.pad[k] = 0;
    s_hdr = g;
    *hdr = g;
  }
  __syncthreads();
  const GridHdr g = s_hdr;
  // The build is one CTA walking several dependent phases over per-cell arrays; with those arrays in shared memory
  // (cells <= kBuildSmemInts / 2 - 1: 25 K cells; a posed body at 5 cm cells has ~15-20 K) every phase runs at
  // shared-memory latency instead of L2 latency -- the target-pose build sits on the frame's critical path in front
  // of K1 (160 us -> see DESIGN.md).  Larger grids work in the global arrays directly, as before.
  const bool in_smem = 2 * (g.ncells + 1) <= kBuildSmemInts;
  int* const cell_start_g = cell_start;
  if (in_smem) { cursor = s_cells; cell_start = s_cells + g.ncells + 1; }

  // 2. histogram
  for (int c = tid; c < g.ncells; c += kBuildThreads) cursor[c] = 0;
  __syncthreads();
  for (int i = tid; i < nv; i += kBuildThreads) {
    float3 v = load_vertex(verts, i, Th, R);
    int cx = min(cell_coord(v.x, g.ox, g.inv_cell), g.nx - 1);
    int cy = min(cell_coord(v.y, g.oy, g.inv_cell), g.ny - 1);
    int cz = min(cell_coord(v.z, g.oz, g.inv_cell), g.nz - 1);
    atomicAdd(&cursor[(cz * g.ny + cy) * g.nx + cx], 1);
  }
  __syncthreads();

  // 3. exclusive scan -> cell_start (each thread owns a contiguous span of cells)
  const int span = (g.ncells + kBuildThreads - 1) / kBuildThreads;
  const int c0 = min(tid * span, g.ncells), c1 = min(c0 + span, g.ncells);
  int local = 0;
  for (int c = c0; c < c1; ++c) local += cursor[c];
  int incl = local;
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_scan[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int v = s_scan[lane];
    int iv = v;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, iv, o);
      if (lane >= o) iv += t;
    }
    s_scan[lane] = iv - v;
  }
  __syncthreads();
  int run = s_scan[wid] + incl - local;
  for (int c = c0; c < c1; ++c) {
    int n = cursor[c];
    cell_start[c] = run;
    run += n;
  }
  if (tid == kBuildThreads - 1) cell_start[g.ncells] = nv;
  __syncthreads();

  // 4. scatter vertices into cell order (w carries the original index)
  for (int c = tid; c < g.ncells; c += kBuildThreads) cursor[c] = cell_start[c];
  __syncthreads();
  for (int i = tid; i < nv; i += kBuildThreads) {
    float3 v = load_vertex(verts, i, Th, R);
    int cx = min(cell_coord(v.x, g.ox, g.inv_cell), g.nx - 1);
    int cy = min(cell_coord(v.y, g.oy, g.inv_cell), g.ny - 1);
    int cz = min(cell_coord(v.z, g.oz, g.inv_cell), g.nz - 1);
    int pos = atomicAdd(&cursor[(cz * g.ny + cy) * g.nx + cx], 1);
    sorted[pos] = make_float4(v.x, v.y, v.z, __int_as_float(i));
  }

  // 5. dilated occupancy bitmap: bit c set iff some cell of c's 27-neighbourhood holds a vertex.  One thread per
  // cell, a warp's 32 consecutive cells form one word (ballot).
  const int nwords = (g.ncells + 31) / 32;
  for (int w = wid; w < nwords; w += kBuildThreads / 32) {
    const int c = w * 32 + lane;
    bool any = false;
    if (c < g.ncells) {
      const int cx = c % g.nx, cy = (c / g.nx) % g.ny, cz = c / (g.nx * g.ny);
      for (int z = max(cz - 1, 0); z <= min(cz + 1, g.nz - 1) && !any; ++z)
        for (int y = max(cy - 1, 0); y <= min(cy + 1, g.ny - 1) && !any; ++y) {
          const int row = (z * g.ny + y) * g.nx;
          any = cell_start[row + min(cx + 1, g.nx - 1) + 1] > cell_start[row + max(cx - 1, 0)];
        }
    }
    const uint32_t bits = __ballot_sync(0xffffffffu, any);
    if (lane == 0) occ[w] = bits;
  }
  if (in_smem)
    for (int c = tid; c <= g.ncells; c += kBuildThreads) cell_start_g[c] = cell_start[c];
}

// Vertices clamped into the last cell (coordinate == bbox max) keep the neighbourhood
// argument valid: the clamp only ever moves a vertex to the cell next to its computed one
// when the extent is an exact multiple of the cell, and that cell is still in the 27-block.

__global__ void knn1_kernel(const float* __restrict__ q, int64_t n, const char* __restrict__ buf,
                            float* __restrict__ d2_out, int32_t* __restrict__ idx_out) {
  GridView g = grid_view(buf);
  const GridHdr h = *g.hdr;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < n; base += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = base + threadIdx.x;
    bool valid = i < n;
    float qx = 0, qy = 0, qz = 0;
    if (valid) { qx = q[3 * i]; qy = q[3 * i + 1]; qz = q[3 * i + 2]; }
    float bd2 = __int_as_float(0x7f800000);
    int bidx = 0x7fffffff;
    if (valid) {
      nn_search27(h, g.cell_start, g.sorted, cell_coord(qx, h.ox, h.inv_cell), cell_coord(qy, h.oy, h.inv_cell),
                  cell_coord(qz, h.oz, h.inv_cell), qx, qy, qz, bd2, bidx);
    }
    nn_brute_warp(h, g.sorted, valid && !(bd2 < h.safe_r2), qx, qy, qz, bd2, bidx);
    if (valid) { d2_out[i] = bd2; idx_out[i] = bidx; }
  }
}

}  // namespace mps

extern "C" size_t mpsnerf_grid_bytes(int n_verts) {
  return mps::kGridOffSorted + sizeof(float4) * (size_t)(n_verts > 0 ? n_verts : 0) + 256;
}

extern "C" int mpsnerf_grid_build(const float* verts, int n_verts, const float* Th, const float* R,
                                  float cell, void* grid, size_t grid_bytes, void* stream) {
  MPS_REQUIRE(verts != nullptr && grid != nullptr);
  MPS_REQUIRE(n_verts > 0 && n_verts < (1 << 24));
  MPS_REQUIRE((Th == nullptr) == (R == nullptr));
  MPS_REQUIRE(cell > 0.f);
  MPS_REQUIRE(grid_bytes >= mpsnerf_grid_bytes(n_verts));
  MPS_REQUIRE((reinterpret_cast<uintptr_t>(grid) & 15) == 0);
  const size_t smem = sizeof(int) * (size_t)mps::kBuildSmemInts;
  MPS_CUDA(cudaFuncSetAttribute(mps::grid_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mps::grid_build_kernel<<<1, mps::kBuildThreads, smem, (cudaStream_t)stream>>>(verts, n_verts, Th, R, cell,
                                                                               static_cast<char*>(grid));
  MPS_LAUNCH_CHECK();
  return MPSNERF_OK;
}

extern "C" int mpsnerf_knn1(const float* query, int64_t n, const void* grid, float* d2_out,
                            int32_t* idx_out, void* stream) {
  MPS_REQUIRE(grid != nullptr && n >= 0);
  if (n == 0) return MPSNERF_OK;
  MPS_REQUIRE(query != nullptr && d2_out != nullptr && idx_out != nullptr);
  int blocks = (int)((n + 127) / 128);
  if (blocks > mps::kNumSMs * 16) blocks = mps::kNumSMs * 16;
  mps::knn1_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(query, n, static_cast<const char*>(grid), d2_out, idx_out);
  MPS_LAUNCH_CHECK();
  return MPSNERF_OK;
}
