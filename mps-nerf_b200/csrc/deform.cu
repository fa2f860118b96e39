// K3: inverse LBS target -> canonical -> source, then projection into the input views.
//
// Restates coarse_deform_target2c (lib/skinnning_batch.py:203-251), coarse_deform_c2source
// (:253-300) and projection (:177-184) for mean_shape = 0, weights_correction = 0.  The blend,
// the 3x3 adjugate inverse and the point transforms are "pinned" (explicit rounding, fixed
// order, mirrored by oracle/oracle.py) so that the nearest canonical-template vertex (idx3)
// is bit-exact against the oracle.
#include "grid.cuh"

namespace mps {

constexpr int kK3Threads = 128;

struct Xf { float m[12]; };   // 3x4 row-major

// Two blends with the same weights (sum_j w_j A0_j and sum_j w_j A1_j), sequential over the 24 joints in
// joint order; zero weights are skipped (adding +-0 never changes a value), the
// j = 0 term always initialises.  kNorm: the weights are first divided by their (sequential) sum
// (lib/skinnning_batch.py:261-262).  The joint loop is rolled in groups of four (one 16-byte weight load per
// group, transforms addressed in shared memory at run time): fully unrolled, the four blends of the kernel were
// 2.7 K instructions of mostly skipped code and the kernel stalled on instruction fetch.
template <bool kNorm>
__device__ __forceinline__ void blend2(const float* __restrict__ skin_w, int v, const float* __restrict__ A0,
                                       const float* __restrict__ A1, Xf& M0, Xf& M1) {
  const float4* p = reinterpret_cast<const float4*>(skin_w + (size_t)v * 24);
  float4 t = __ldg(p);
  float s = 1.f;
  if (kNorm) {
    s = padd(padd(padd(t.x, t.y), t.z), t.w);
#pragma unroll 1
    for (int k = 1; k < 6; ++k) {
      const float4 n = __ldg(p + k);
      s = padd(padd(padd(padd(s, n.x), n.y), n.z), n.w);
    }
  }
#pragma unroll 1
  for (int k = 0; k < 6; ++k) {
    float wv[4] = {t.x, t.y, t.z, t.w};
    if (k < 5) t = __ldg(p + k + 1);
    const float* a0 = A0 + 48 * k;
    const float* a1 = A1 + 48 * k;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float w = kNorm ? pdiv(wv[i], s) : wv[i];
      if (i == 0 && k == 0) {
#pragma unroll
        for (int e = 0; e < 12; ++e) { M0.m[e] = pmul(w, a0[e]); M1.m[e] = pmul(w, a1[e]); }
      } else if (w != 0.f) {
#pragma unroll
        for (int e = 0; e < 12; ++e) {
          M0.m[e] = padd(M0.m[e], pmul(w, a0[12 * i + e]));
          M1.m[e] = padd(M1.m[e], pmul(w, a1[12 * i + e]));
        }
      }
    }
  }
}

__device__ __forceinline__ void apply_inv(const Xf& M, float x, float y, float z, float& ox, float& oy, float& oz) {
#define A_(r, c) M.m[4 * (r) + (c)]
  const float c00 = psub(pmul(A_(1, 1), A_(2, 2)), pmul(A_(1, 2), A_(2, 1)));
  const float c01 = psub(pmul(A_(0, 2), A_(2, 1)), pmul(A_(0, 1), A_(2, 2)));
  const float c02 = psub(pmul(A_(0, 1), A_(1, 2)), pmul(A_(0, 2), A_(1, 1)));
  const float c10 = psub(pmul(A_(1, 2), A_(2, 0)), pmul(A_(1, 0), A_(2, 2)));
  const float c11 = psub(pmul(A_(0, 0), A_(2, 2)), pmul(A_(0, 2), A_(2, 0)));
  const float c12 = psub(pmul(A_(0, 2), A_(1, 0)), pmul(A_(0, 0), A_(1, 2)));
  const float c20 = psub(pmul(A_(1, 0), A_(2, 1)), pmul(A_(1, 1), A_(2, 0)));
  const float c21 = psub(pmul(A_(0, 1), A_(2, 0)), pmul(A_(0, 0), A_(2, 1)));
  const float c22 = psub(pmul(A_(0, 0), A_(1, 1)), pmul(A_(0, 1), A_(1, 0)));
  const float det = padd(padd(pmul(A_(0, 0), c00), pmul(A_(0, 1), c10)), pmul(A_(0, 2), c20));
  const float r = pdiv(1.0f, det);
  const float v0 = psub(x, A_(0, 3)), v1 = psub(y, A_(1, 3)), v2 = psub(z, A_(2, 3));
  ox = padd(padd(pmul(pmul(c00, r), v0), pmul(pmul(c01, r), v1)), pmul(pmul(c02, r), v2));
  oy = padd(padd(pmul(pmul(c10, r), v0), pmul(pmul(c11, r), v1)), pmul(pmul(c12, r), v2));
  oz = padd(padd(pmul(pmul(c20, r), v0), pmul(pmul(c21, r), v1)), pmul(pmul(c22, r), v2));
#undef A_
}

__device__ __forceinline__ void apply_fwd(const Xf& M, float x, float y, float z, float& ox, float& oy, float& oz) {
  ox = padd(padd(padd(pmul(M.m[0], x), pmul(M.m[1], y)), pmul(M.m[2], z)), M.m[3]);
  oy = padd(padd(padd(pmul(M.m[4], x), pmul(M.m[5], y)), pmul(M.m[6], z)), M.m[7]);
  oz = padd(padd(padd(pmul(M.m[8], x), pmul(M.m[9], y)), pmul(M.m[10], z)), M.m[11]);
}

__global__ void __launch_bounds__(kK3Threads)
deform_project_kernel(const int32_t* __restrict__ act_pid, const int32_t* __restrict__ act_idx2,
                      const float* __restrict__ act_q, int64_t first, int64_t count,
                      const float* __restrict__ skin_w, const mpsnerf_frame* __restrict__ frame,
                      const char* __restrict__ grid_buf, float* __restrict__ xc_out, float* __restrict__ uv_out,
                      float* __restrict__ smpl_src, int32_t* __restrict__ idx3_out, float* __restrict__ xw_out,
                      int identity_canonical, const int32_t* __restrict__ count_dev) {
  if (count_dev != nullptr) {          // device-side active count: `count` is only the capacity of the slab
    const int64_t n = (int64_t)*count_dev - first;
    count = n < count ? (n > 0 ? n : 0) : count;
  }
  __shared__ mpsnerf_frame s_fr;
  __shared__ GridHdr s_hdr;
  const GridView g = grid_view(grid_buf);
  {
    const int* src = reinterpret_cast<const int*>(frame);
    int* dst = reinterpret_cast<int*>(&s_fr);
    for (int i = threadIdx.x; i < (int)(sizeof(mpsnerf_frame) / 4); i += kK3Threads) dst[i] = src[i];
    if (threadIdx.x == 0) s_hdr = *g.hdr;
  }
  __syncthreads();
  const GridHdr h = s_hdr;
  const int V = s_fr.n_views;
  const float INF = __int_as_float(0x7f800000);

  for (int64_t base = (int64_t)blockIdx.x * kK3Threads; base < count; base += (int64_t)gridDim.x * kK3Threads) {
    const int64_t i = base + threadIdx.x;
    const bool valid = i < count;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    Xf M, M2;
    if (valid) {
      const int64_t a = first + i;
      const float qx = act_q[3 * a], qy = act_q[3 * a + 1], qz = act_q[3 * a + 2];
      if (identity_canonical) {                 // extract_mesh: canonical_pts = world_query_pts (:394-396)
        cx = qx; cy = qy; cz = qz;
      } else {
        blend2<false>(skin_w, act_idx2[a], s_fr.A_tp, s_fr.A_big_tp, M, M2);
        float tx, ty, tz;
        apply_inv(M, qx, qy, qz, tx, ty, tz);
        apply_fwd(M2, tx, ty, tz, cx, cy, cz);
      }
    }
    // nearest template vertex (no radius guarantee: accept below safe_r2, else exact scan)
    float bd2 = INF;
    int bidx = 0x7fffffff;
    if (valid) {
      nn_search27(h, g.cell_start, g.sorted, cell_coord(cx, h.ox, h.inv_cell), cell_coord(cy, h.oy, h.inv_cell),
                  cell_coord(cz, h.oz, h.inv_cell), cx, cy, cz, bd2, bidx);
    }
    nn_brute_warp(h, g.sorted, valid && !(bd2 < h.safe_r2), cx, cy, cz, bd2, bidx);
    if (valid) {
    blend2<true>(skin_w, bidx, s_fr.A_big_sp, s_fr.A_sp, M, M2);   // weights normalised, :261-262
    float tx, ty, tz, sx, sy, sz;
    apply_inv(M, cx, cy, cz, tx, ty, tz);
    apply_fwd(M2, tx, ty, tz, sx, sy, sz);                      // smpl_src_pts
    const float* Ri = s_fr.Rinv_sp;
    const float wx = padd(padd(padd(pmul(sx, Ri[0]), pmul(sy, Ri[3])), pmul(sz, Ri[6])), s_fr.Th_sp[0]);   // :297-298
    const float wy = padd(padd(padd(pmul(sx, Ri[1]), pmul(sy, Ri[4])), pmul(sz, Ri[7])), s_fr.Th_sp[1]);
    const float wz = padd(padd(padd(pmul(sx, Ri[2]), pmul(sy, Ri[5])), pmul(sz, Ri[8])), s_fr.Th_sp[2]);

    xc_out[3 * i] = cx; xc_out[3 * i + 1] = cy; xc_out[3 * i + 2] = cz;
    if (smpl_src != nullptr) {
      const int64_t pid = act_pid[first + i];
      smpl_src[3 * pid] = sx; smpl_src[3 * pid + 1] = sy; smpl_src[3 * pid + 2] = sz;
    }
    if (idx3_out != nullptr) idx3_out[i] = bidx;
    if (xw_out != nullptr) { xw_out[3 * i] = wx; xw_out[3 * i + 1] = wy; xw_out[3 * i + 2] = wz; }
    for (int v = 0; v < V; ++v) {                               // projection, :177-184
      const float* R = s_fr.cam_R + 9 * v;
      const float* T = s_fr.cam_T + 3 * v;
      const float* K = s_fr.cam_K + 9 * v;
      const float c0 = R[0] * wx + R[1] * wy + R[2] * wz + T[0];
      const float c1 = R[3] * wx + R[4] * wy + R[5] * wz + T[1];
      const float c2 = R[6] * wx + R[7] * wy + R[8] * wz + T[2];
      const float i0 = K[0] * c0 + K[1] * c1 + K[2] * c2;
      const float i1 = K[3] * c0 + K[4] * c1 + K[5] * c2;
      const float i2 = K[6] * c0 + K[7] * c1 + K[8] * c2;
      const float den = i2 + 1e-5f;
      uv_out[(i * V + v) * 2] = i0 / den;
      uv_out[(i * V + v) * 2 + 1] = i1 / den;
    }
    }  // valid
  }
}

}  // namespace mps

static int deform_project_impl(const int32_t* act_pid, const int32_t* act_idx2, const float* act_q,
                               int64_t first, int64_t count, const float* skin_w,
                               const mpsnerf_frame* frame, const void* grid_tv, float* xc, float* uv,
                               float* smpl_src, int32_t* idx3, float* xw, int identity_canonical,
                               const int32_t* count_dev, void* stream) {
  MPS_REQUIRE(first >= 0 && count >= 0);
  if (count == 0) return MPSNERF_OK;
  MPS_REQUIRE(act_pid && act_q && skin_w && frame && grid_tv && xc && uv);
  MPS_REQUIRE(identity_canonical || act_idx2 != nullptr);
  MPS_REQUIRE((reinterpret_cast<uintptr_t>(skin_w) & 15) == 0);
  int64_t blocks = (count + mps::kK3Threads - 1) / mps::kK3Threads;
  if (blocks > mps::kNumSMs * 16) blocks = mps::kNumSMs * 16;
  mps::deform_project_kernel<<<(int)blocks, mps::kK3Threads, 0, (cudaStream_t)stream>>>(
      act_pid, act_idx2, act_q, first, count, skin_w, frame, static_cast<const char*>(grid_tv), xc, uv,
      smpl_src, idx3, xw, identity_canonical, count_dev);
  MPS_LAUNCH_CHECK();
  return MPSNERF_OK;
}

extern "C" int mpsnerf_deform_project(const int32_t* act_pid, const int32_t* act_idx2, const float* act_q,
                                      int64_t first, int64_t count, const float* skin_w,
                                      const mpsnerf_frame* frame, const void* grid_tv, float* xc, float* uv,
                                      float* smpl_src, int32_t* idx3, float* xw, int identity_canonical,
                                      void* stream) {
  return deform_project_impl(act_pid, act_idx2, act_q, first, count, skin_w, frame, grid_tv, xc, uv, smpl_src, idx3, xw,
                             identity_canonical, nullptr, stream);
}
extern "C" int mpsnerf_deform_project_dc(const int32_t* act_pid, const int32_t* act_idx2, const float* act_q,
                                         int64_t first, int64_t capacity, const int32_t* count_dev, const float* skin_w,
                                         const mpsnerf_frame* frame, const void* grid_tv, float* xc, float* uv,
                                         float* smpl_src, void* stream) {
  MPS_REQUIRE(count_dev != nullptr);
  return deform_project_impl(act_pid, act_idx2, act_q, first, capacity, skin_w, frame, grid_tv, xc, uv, smpl_src, nullptr,
                             nullptr, 0, count_dev, stream);
}
