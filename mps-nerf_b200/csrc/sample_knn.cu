// K1: stratified sampling + world->SMPL + human-region mask + nearest posed vertex + compaction.
//
// Two kernels with identical results:
//  * sample_knn_warp_kernel (default): warp-autonomous.  A warp owns 32 consecutive sample points per iteration
//    and never meets a block barrier: every lane generates its point, moves it to SMPL space (pinned fp32) and
//    tests one bit of the dilated occupancy bitmap; the survivors ("candidates", ~15 % of a frame) are searched
//    four at a time, eight lanes per candidate -- the group walks the nine x-runs of the 27-cell neighbourhood
//    together, runs beyond the mask radius pruned, vertices dealt out eight at a time; the (d2, index) minimum is
//    reduced with three shuffle steps -- and the active points are
//    compacted with one ballot and one atomicAdd per warp.  Per-point outputs are written straight from
//    registers (raw / mask) or through a 384-byte per-warp staging row (smpl_query, so that the stores are
//    128-bit and coalesced).
//  * sample_knn_kernel (MPSNERF_K1=block): the round-1 block-synchronous form -- 256 points per block
//    iteration, candidates compacted across the block, six barriers per iteration.  Kept for A/B timing.
#include <stdlib.h>

#include "grid.cuh"

namespace mps {

constexpr int kK1Threads = 256;
constexpr int kStrip = 8;             // chunks of 32 points per warp strip (warp-autonomous kernel)
constexpr int kK1QueueBytes = (kK1Threads / 32) * 5 * kStrip * 32 * 4;

__global__ void __launch_bounds__(kK1Threads)
sample_knn_kernel(const float* __restrict__ rays, int64_t n_points, int S, const float* __restrict__ t_vals,
                  const float* __restrict__ u, const float* __restrict__ points,
                  const mpsnerf_frame* __restrict__ frame, const char* __restrict__ grid_buf,
                  float* __restrict__ raw, float* __restrict__ pts_mask, float* __restrict__ smpl_query,
                  float* __restrict__ smpl_src, int32_t* __restrict__ act_pid, int32_t* __restrict__ act_idx2,
                  float* __restrict__ act_q, int32_t* __restrict__ act_count) {
  const GridView g = grid_view(grid_buf);
  __shared__ GridHdr s_hdr;
  __shared__ float s_fr[12];                 // Th(3) R(9)
  __shared__ __align__(16) float s_q[kK1Threads * 3];   // read back as float4 in phase 3
  __shared__ float s_d2[kK1Threads];
  __shared__ int s_idx[kK1Threads];
  __shared__ int s_cand[kK1Threads];
  __shared__ int s_warp_cnt[kK1Threads / 32];
  __shared__ int s_ncand, s_nact, s_base;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_hdr = *g.hdr;
  if (tid < 3) s_fr[tid] = frame->Th_tp[tid];
  if (tid >= 3 && tid < 12) s_fr[tid] = frame->R_tp[tid - 3];
  __syncthreads();
  const GridHdr h = s_hdr;
  const float INF = __int_as_float(0x7f800000);

  for (int64_t base = (int64_t)blockIdx.x * kK1Threads; base < n_points; base += (int64_t)gridDim.x * kK1Threads) {
    const int64_t pid = base + tid;
    const bool valid = pid < n_points;
    // ---- phase 1
    float qx = 0.f, qy = 0.f, qz = 0.f;
    bool cand = false;
    if (valid) {
      float px, py, pz;
      if (points != nullptr) {
        px = points[3 * pid]; py = points[3 * pid + 1]; pz = points[3 * pid + 2];
      } else {
        // P < 2^31 (checked by the host wrapper): 32-bit division instead of the ~100-instruction 64-bit one
        const uint32_t r = (uint32_t)pid / (uint32_t)S;
        const int s = (int)((uint32_t)pid - r * (uint32_t)S);
        const float4 ra = __ldg(reinterpret_cast<const float4*>(rays) + 2 * (size_t)r);       // o.xyz, d.x
        const float4 rb = __ldg(reinterpret_cast<const float4*>(rays) + 2 * (size_t)r + 1);   // d.yz, near, far
        const float z = sample_z(rb.z, rb.w, t_vals, s, S, u ? u + (size_t)r * S : nullptr);
        px = padd(ra.x, pmul(ra.w, z));     // run_nerf_batch.py:424
        py = padd(ra.y, pmul(rb.x, z));
        pz = padd(ra.z, pmul(rb.y, z));
      }
      const float d0 = psub(px, s_fr[0]), d1 = psub(py, s_fr[1]), d2 = psub(pz, s_fr[2]);
      qx = padd(padd(pmul(d0, s_fr[3]), pmul(d1, s_fr[6])), pmul(d2, s_fr[9]));   // (p-Th)@R, :347
      qy = padd(padd(pmul(d0, s_fr[4]), pmul(d1, s_fr[7])), pmul(d2, s_fr[10]));
      qz = padd(padd(pmul(d0, s_fr[5]), pmul(d1, s_fr[8])), pmul(d2, s_fr[11]));
      cand = grid_maybe_near(h, g.occ, cell_coord(qx, h.ox, h.inv_cell), cell_coord(qy, h.oy, h.inv_cell),
                             cell_coord(qz, h.oz, h.inv_cell));
    }
    s_q[3 * tid] = qx; s_q[3 * tid + 1] = qy; s_q[3 * tid + 2] = qz;
    s_d2[tid] = INF;
    s_idx[tid] = -1;
    unsigned bal = __ballot_sync(0xffffffffu, cand);
    if (lane == 0) s_warp_cnt[wid] = __popc(bal);
    __syncthreads();
    {
      int off = 0;
      for (int w = 0; w < wid; ++w) off += s_warp_cnt[w];
      if (cand) s_cand[off + __popc(bal & ((1u << lane) - 1))] = tid;
      if (tid == kK1Threads - 1) s_ncand = off + __popc(bal);
    }
    __syncthreads();
    // ---- phase 2 (dense over candidates)
    const int ncand = s_ncand;
    if (tid < ncand) {
      const int t = s_cand[tid];
      const float cxq = s_q[3 * t], cyq = s_q[3 * t + 1], czq = s_q[3 * t + 2];
      float bd2 = INF;
      int bidx = 0x7fffffff;
      nn_search27_all(h, g.cell_start, g.sorted, cell_coord(cxq, h.ox, h.inv_cell), cell_coord(cyq, h.oy, h.inv_cell),
                  cell_coord(czq, h.oz, h.inv_cell), cxq, cyq, czq, bd2, bidx);
      s_d2[t] = bd2;
      s_idx[t] = bidx;
    }
    __syncthreads();
    // ---- phase 3
    const bool active = valid && (s_d2[tid] < kMaskThresh);     // lib/skinnning_batch.py:360-361
    bal = __ballot_sync(0xffffffffu, active);
    if (lane == 0) s_warp_cnt[wid] = __popc(bal);
    __syncthreads();
    int off = 0;
    for (int w = 0; w < wid; ++w) off += s_warp_cnt[w];
    if (tid == kK1Threads - 1) {
      const int n = off + __popc(bal);
      s_nact = n;
      s_base = n ? atomicAdd(act_count, n) : 0;
    }
    __syncthreads();
    if (active) {
      const int64_t slot = (int64_t)s_base + off + __popc(bal & ((1u << lane) - 1));
      act_pid[slot] = (int32_t)pid;
      act_idx2[slot] = s_idx[tid];
      act_q[3 * slot] = qx; act_q[3 * slot + 1] = qy; act_q[3 * slot + 2] = qz;
    }
    if (valid) {
      pts_mask[pid] = active ? 1.0f : 0.0f;
      if (!active) reinterpret_cast<float4*>(raw)[pid] = make_float4(-80.f, -80.f, -80.f, -80.f);   // :493
    }
    // smpl_query / smpl_src rows (ref :483-484): staged so that the stores are 128-bit and coalesced
    if (!active) { s_q[3 * tid] = 0.f; s_q[3 * tid + 1] = 0.f; s_q[3 * tid + 2] = 0.f; }
    __syncthreads();
    if (base + kK1Threads <= n_points) {
      if (tid < kK1Threads * 3 / 4) {
        reinterpret_cast<float4*>(smpl_query + 3 * base)[tid] = reinterpret_cast<const float4*>(s_q)[tid];
        reinterpret_cast<float4*>(smpl_src + 3 * base)[tid] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else if (valid) {
      for (int k = 0; k < 3; ++k) { smpl_query[3 * pid + k] = s_q[3 * tid + k]; smpl_src[3 * pid + k] = 0.f; }
    }
    __syncthreads();
  }
}

// ---- warp-autonomous form ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kK1Threads)
sample_knn_warp_kernel(const float* __restrict__ rays, int64_t n_points, int S, const float* __restrict__ t_vals,
                       const float* __restrict__ u, const float* __restrict__ points,
                       const mpsnerf_frame* __restrict__ frame, const char* __restrict__ grid_buf,
                       float* __restrict__ raw, float* __restrict__ pts_mask, float* __restrict__ smpl_query,
                       float* __restrict__ smpl_src, int32_t* __restrict__ act_pid, int32_t* __restrict__ act_idx2,
                       float* __restrict__ act_q, int32_t* __restrict__ act_count) {
  const GridView g = grid_view(grid_buf);
  __shared__ GridHdr s_hdr;
  __shared__ float s_fr[12];                                     // Th(3) R(9)
  __shared__ __align__(16) float s_stage[kK1Threads / 32][96];   // one smpl_query row block per warp
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_hdr = *g.hdr;
  if (tid < 3) s_fr[tid] = frame->Th_tp[tid];
  if (tid >= 3 && tid < 12) s_fr[tid] = frame->R_tp[tid - 3];
  __syncthreads();
  const GridHdr h = s_hdr;
  const float INF = __int_as_float(0x7f800000);
  const int sub = lane & 7, grp = lane >> 3;
  float* stage = s_stage[wid];
  const int64_t nchunks = (n_points + 31) >> 5;
  constexpr int kWarps = kK1Threads / 32;
  // A warp works through strips of kStrip consecutive chunks (256 points = 4 rays at S = 64) and queues the strip's
  // active points in shared memory; the queue is flushed with ONE atomicAdd and coalesced stores per strip.  The
  // active list then keeps runs of up to 256 points' worth of spatial neighbours -- K3 and K4 read it 32 entries per
  // warp and their grid / texture reads coalesce only if those are neighbours (per-chunk flushes interleave the
  // chunks of 700 concurrent warps: K3 +23 %) -- and the atomic traffic on the one counter drops 8x.
  extern __shared__ __align__(16) int32_t s_queue_all[];                 // [kWarps][5][kStrip * 32]
  int32_t* q_pid = s_queue_all + wid * (5 * kStrip * 32);
  int32_t* q_idx = q_pid + kStrip * 32;
  float* q_q = reinterpret_cast<float*>(q_idx + kStrip * 32);          // [3][kStrip * 32], struct of arrays
  const int64_t nstrips = (nchunks + kStrip - 1) / kStrip;

  for (int64_t strip = (int64_t)blockIdx.x * kWarps + wid; strip < nstrips; strip += (int64_t)gridDim.x * kWarps) {
   int qn = 0;
   for (int64_t c = strip * kStrip; c < min((strip + 1) * kStrip, nchunks); ++c) {
    const int64_t base = c << 5, pid = base + lane;
    const bool valid = pid < n_points;
    // ---- every lane: its point in SMPL space, one bitmap test
    float qx = 0.f, qy = 0.f, qz = 0.f;
    int cx = 0, cy = 0, cz = 0;
    bool cand = false;
    if (valid) {
      float px, py, pz;
      if (points != nullptr) {
        px = points[3 * pid]; py = points[3 * pid + 1]; pz = points[3 * pid + 2];
      } else {
        // P < 2^31 (checked by the host wrapper): 32-bit division instead of the ~100-instruction 64-bit one
        const uint32_t r = (uint32_t)pid / (uint32_t)S;
        const int s = (int)((uint32_t)pid - r * (uint32_t)S);
        const float4 ra = __ldg(reinterpret_cast<const float4*>(rays) + 2 * (size_t)r);       // o.xyz, d.x
        const float4 rb = __ldg(reinterpret_cast<const float4*>(rays) + 2 * (size_t)r + 1);   // d.yz, near, far
        const float z = sample_z(rb.z, rb.w, t_vals, s, S, u ? u + (size_t)r * S : nullptr);
        px = padd(ra.x, pmul(ra.w, z));     // run_nerf_batch.py:424
        py = padd(ra.y, pmul(rb.x, z));
        pz = padd(ra.z, pmul(rb.y, z));
      }
      const float d0 = psub(px, s_fr[0]), d1 = psub(py, s_fr[1]), d2 = psub(pz, s_fr[2]);
      qx = padd(padd(pmul(d0, s_fr[3]), pmul(d1, s_fr[6])), pmul(d2, s_fr[9]));   // (p-Th)@R, :347
      qy = padd(padd(pmul(d0, s_fr[4]), pmul(d1, s_fr[7])), pmul(d2, s_fr[10]));
      qz = padd(padd(pmul(d0, s_fr[5]), pmul(d1, s_fr[8])), pmul(d2, s_fr[11]));
      cx = cell_coord(qx, h.ox, h.inv_cell); cy = cell_coord(qy, h.oy, h.inv_cell); cz = cell_coord(qz, h.oz, h.inv_cell);
      cand = grid_maybe_near(h, g.occ, cx, cy, cz);
    }
    // ---- candidates, four per pass: lanes [8 k, 8 k + 8) search the k-th remaining candidate
    float bd2 = INF;
    int bidx = 0x7fffffff;
    unsigned cm = __ballot_sync(0xffffffffu, cand);
    while (cm) {
      const int b0 = __ffs(cm) - 1;
      const unsigned m1 = cm & (cm - 1);
      const int b1 = m1 ? __ffs(m1) - 1 : -1;
      const unsigned m2 = m1 & (m1 - 1);
      const int b2 = m2 ? __ffs(m2) - 1 : -1;
      const unsigned m3 = m2 & (m2 - 1);
      const int b3 = m3 ? __ffs(m3) - 1 : -1;
      cm = m3 & (m3 - 1);
      const int src = grp == 0 ? b0 : grp == 1 ? b1 : grp == 2 ? b2 : b3;
      const int sl = src < 0 ? 0 : src;
      const float sx = __shfl_sync(0xffffffffu, qx, sl), sy = __shfl_sync(0xffffffffu, qy, sl), sz = __shfl_sync(0xffffffffu, qz, sl);
      const int ccx = __shfl_sync(0xffffffffu, cx, sl), ccy = __shfl_sync(0xffffffffu, cy, sl), ccz = __shfl_sync(0xffffffffu, cz, sl);
      float d = INF;
      int id = 0x7fffffff;
      if (src >= 0) {
        // The eight lanes of a group walk the nine x-runs of the candidate's 27-cell neighbourhood together, a run's
        // vertices dealt out eight at a time (the runs are uneven -- a cell on a hand holds 100+ vertices, most
        // hold one or two -- so a lane per run would leave the warp waiting for its longest run).  A run, or the
        // outer cells of a run, is skipped when a lower bound of its distance to the query already exceeds the
        // mask radius: only vertices with d2 < r^2 can make the point active or be its nearest vertex.  Bound =
        // gap to the cell face shrunk by 1e-3 cell (covers the rounding of the binning and of the pinned d2), the
        // same rule as nn_search27 with cap = r^2.
        const float m = 1e-3f * h.cell;
        const float fx = sx - (h.ox + (float)ccx * h.cell), fy = sy - (h.oy + (float)ccy * h.cell),
                    fz = sz - (h.oz + (float)ccz * h.cell);
        float gl, gr;
        gl = fmaxf(fx - m, 0.f); gr = fmaxf(h.cell - fx - m, 0.f);
        const float gxl = gl * gl, gxr = gr * gr;
        gl = fmaxf(fy - m, 0.f); gr = fmaxf(h.cell - fy - m, 0.f);
        const float gyl = gl * gl, gyr = gr * gr;
        gl = fmaxf(fz - m, 0.f); gr = fmaxf(h.cell - fz - m, 0.f);
        const float gzl = gl * gl, gzr = gr * gr;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const int dy = k % 3 - 1, dz = k / 3 - 1;
          const int y = ccy + dy, z = ccz + dz;
          const float rowd2 = (dy < 0 ? gyl : (dy > 0 ? gyr : 0.f)) + (dz < 0 ? gzl : (dz > 0 ? gzr : 0.f));
          if ((unsigned)y >= (unsigned)h.ny || (unsigned)z >= (unsigned)h.nz || rowd2 > kMaskThresh) continue;
          const int x0 = max(ccx - ((gxl + rowd2 > kMaskThresh) ? 0 : 1), 0);
          const int x1 = min(ccx + ((gxr + rowd2 > kMaskThresh) ? 0 : 1), h.nx - 1);
          if (x0 > x1) continue;
          const int row = (z * h.ny + y) * h.nx;
          const int b = __ldg(&g.cell_start[row + x0]);
          const int e = __ldg(&g.cell_start[row + x1 + 1]);
          for (int i = b + sub; i < e; i += 8) {
            const float4 v = __ldg(&g.sorted[i]);
            nn_update(dist2_pinned(sx, sy, sz, v.x, v.y, v.z), __float_as_int(v.w), d, id);
          }
        }
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {                // (d2, index) minimum over the group's eight lanes
        const float od = __shfl_xor_sync(0xffffffffu, d, o);
        const int oi = __shfl_xor_sync(0xffffffffu, id, o);
        nn_update(od, oi, d, id);
      }
      // hand each group's result to the lane that owns the candidate
      const float r0 = __shfl_sync(0xffffffffu, d, 0), r1 = __shfl_sync(0xffffffffu, d, 8), r2 = __shfl_sync(0xffffffffu, d, 16),
                  r3 = __shfl_sync(0xffffffffu, d, 24);
      const int i0 = __shfl_sync(0xffffffffu, id, 0), i1 = __shfl_sync(0xffffffffu, id, 8), i2 = __shfl_sync(0xffffffffu, id, 16),
                i3 = __shfl_sync(0xffffffffu, id, 24);
      if (lane == b0) { bd2 = r0; bidx = i0; }
      if (lane == b1) { bd2 = r1; bidx = i1; }
      if (lane == b2) { bd2 = r2; bidx = i2; }
      if (lane == b3) { bd2 = r3; bidx = i3; }
    }
    // ---- outputs and compaction
    const bool active = valid && (bd2 < kMaskThresh);     // lib/skinnning_batch.py:360-361
    const unsigned am = __ballot_sync(0xffffffffu, active);
    if (active) {
      const int slot = qn + __popc(am & ((1u << lane) - 1));
      q_pid[slot] = (int32_t)pid;
      q_idx[slot] = bidx;
      q_q[slot] = qx; q_q[kStrip * 32 + slot] = qy; q_q[2 * kStrip * 32 + slot] = qz;
    }
    qn += __popc(am);
    if (valid) {
      pts_mask[pid] = active ? 1.0f : 0.0f;
      if (!active) reinterpret_cast<float4*>(raw)[pid] = make_float4(-80.f, -80.f, -80.f, -80.f);   // :493
    }
    // smpl_query / smpl_src rows (ref :483-484): q where active, zeros elsewhere
    if (base + 32 <= n_points) {
      stage[3 * lane] = active ? qx : 0.f; stage[3 * lane + 1] = active ? qy : 0.f; stage[3 * lane + 2] = active ? qz : 0.f;
      __syncwarp();
      if (lane < 24) {
        reinterpret_cast<float4*>(smpl_query + 3 * base)[lane] = reinterpret_cast<const float4*>(stage)[lane];
        reinterpret_cast<float4*>(smpl_src + 3 * base)[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncwarp();
    } else if (valid) {
      smpl_query[3 * pid] = active ? qx : 0.f; smpl_query[3 * pid + 1] = active ? qy : 0.f; smpl_query[3 * pid + 2] = active ? qz : 0.f;
      smpl_src[3 * pid] = 0.f; smpl_src[3 * pid + 1] = 0.f; smpl_src[3 * pid + 2] = 0.f;
    }
   }
   // ---- flush the strip's queue: one atomicAdd, coalesced stores
   __syncwarp();
   if (qn) {
     int slot0 = 0;
     if (lane == 0) slot0 = atomicAdd(act_count, qn);
     slot0 = __shfl_sync(0xffffffffu, slot0, 0);
     for (int i = lane; i < qn; i += 32) {
       act_pid[slot0 + i] = q_pid[i];
       act_idx2[slot0 + i] = q_idx[i];
     }
     for (int i = lane; i < 3 * qn; i += 32) {      // act_q rows are (x, y, z): element i = component i % 3 of entry i / 3
       const int e = i / 3, k = i - 3 * e;
       act_q[3 * (int64_t)slot0 + i] = q_q[k * (kStrip * 32) + e];
     }
   }
   __syncwarp();
  }
}

}  // namespace mps

extern "C" int mpsnerf_sample_knn(const float* rays, int64_t n_rays, int32_t S, const float* t_vals,
                                  const float* u, const float* points, const mpsnerf_frame* frame,
                                  const void* grid_tp, float* raw, float* pts_mask, float* smpl_query,
                                  float* smpl_src, int32_t* act_pid, int32_t* act_idx2, float* act_q,
                                  int32_t* act_count, void* stream) {
  MPS_REQUIRE(n_rays >= 0 && S >= 1);
  const int64_t P = n_rays * S;
  MPS_REQUIRE(P < ((int64_t)1 << 31));
  if (P == 0) return MPSNERF_OK;
  MPS_REQUIRE(points != nullptr || (rays != nullptr && t_vals != nullptr));
  MPS_REQUIRE(frame && grid_tp && raw && pts_mask && smpl_query && smpl_src);
  MPS_REQUIRE(act_pid && act_idx2 && act_q && act_count);
  MPS_REQUIRE((reinterpret_cast<uintptr_t>(raw) & 15) == 0 && (reinterpret_cast<uintptr_t>(rays) & 15) == 0);
  MPS_REQUIRE((reinterpret_cast<uintptr_t>(smpl_query) & 15) == 0 && (reinterpret_cast<uintptr_t>(smpl_src) & 15) == 0);
  int64_t blocks = (P + mps::kK1Threads - 1) / mps::kK1Threads;
  if (blocks > mps::kNumSMs * 8) blocks = mps::kNumSMs * 8;
  static int form = -1;       // MPSNERF_K1 = warp (default) | block
  if (form < 0) { const char* e = getenv("MPSNERF_K1"); form = (e && e[0] == 'b') ? 1 : 0; }
  if (form == 1)
    mps::sample_knn_kernel<<<(int)blocks, mps::kK1Threads, 0, (cudaStream_t)stream>>>(
        rays, P, S, t_vals, u, points, frame, static_cast<const char*>(grid_tp), raw, pts_mask, smpl_query,
        smpl_src, act_pid, act_idx2, act_q, act_count);
  else {
    // persistent warps: exactly as many blocks as are resident at once (no second, partial wave)
    static int resident = 0;
    if (resident == 0) {
      int per_sm = 0;
      MPS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mps::sample_knn_warp_kernel, mps::kK1Threads,
                                                             mps::kK1QueueBytes));
      resident = mps::kNumSMs * (per_sm > 0 ? per_sm : 1);
    }
    const int64_t want = (P + mps::kK1Threads - 1) / mps::kK1Threads;
    blocks = want < resident ? want : resident;
    mps::sample_knn_warp_kernel<<<(int)blocks, mps::kK1Threads, mps::kK1QueueBytes, (cudaStream_t)stream>>>(
        rays, P, S, t_vals, u, points, frame, static_cast<const char*>(grid_tp), raw, pts_mask, smpl_query,
        smpl_src, act_pid, act_idx2, act_q, act_count);
  }
  MPS_LAUNCH_CHECK();
  return MPSNERF_OK;
}
