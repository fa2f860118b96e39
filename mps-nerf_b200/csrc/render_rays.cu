// Fused render_rays entry point: one C call enqueues the whole per-ray path of a frame -- K1 (sample + mask + argmin +
// compaction), K3 (deform + project), K4 (gather, fp16 tokens), T + M (tcgen05 transformer + MLP), K6 (composite) --
// on one stream, with the active count read on the device by every stage.  Replaces render_rays of the reference
// (run_nerf_batch.py:401-444: sampling -> network_query_fn -> raw2outputs) for the tensor-core path.
#include "common.cuh"

namespace mps {

struct RenderWs {
  int32_t *act_pid, *act_idx2;
  float *act_q, *xc, *uv;
  void* tokens;
  void* dense;
};

static size_t carve_render(RenderWs& w, char* base, int64_t P, int64_t cap, int V) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base + off;
    off += (bytes + 255) / 256 * 256;
    return p;
  };
  w.act_pid = reinterpret_cast<int32_t*>(take(4 * (size_t)P));
  w.act_idx2 = reinterpret_cast<int32_t*>(take(4 * (size_t)P));
  w.act_q = reinterpret_cast<float*>(take(12 * (size_t)P));
  w.xc = reinterpret_cast<float*>(take(12 * (size_t)cap));
  w.uv = reinterpret_cast<float*>(take(8 * (size_t)V * cap));
  w.tokens = take(2 * (size_t)MPSNERF_TOKEN_LD * V * cap);
  w.dense = take(mpsnerf_dense_bf16_workspace(cap, V));
  return off;
}

}  // namespace mps

extern "C" size_t mpsnerf_render_rays_workspace(int64_t n_rays, int32_t S, int n_views, int64_t capacity) {
  mps::RenderWs w;
  const int64_t P = (n_rays > 0 ? n_rays : 0) * (int64_t)(S > 0 ? S : 0);
  return mps::carve_render(w, nullptr, P, capacity > 0 ? capacity : 0, n_views) + 256;
}

extern "C" int mpsnerf_render_rays_bf16(const float* rays, int64_t n_rays, int32_t S, const float* t_vals, const float* u,
                                        const mpsnerf_frame* frame, const void* grid_tp, const void* grid_tv,
                                        const float* skin_w, const float* latent, const float* img4, const void* packed,
                                        size_t packed_bytes, int n_views, int occupancy, float* raw, float* pts_mask,
                                        float* smpl_query, float* smpl_src, float* rgb, float* disp, float* acc, float* depth,
                                        int32_t* act_count, int64_t capacity, void* workspace, size_t workspace_bytes,
                                        void* event_lbs, void* event_trunk, int32_t* host_count, void* event_count,
                                        void* stream) {
  MPS_REQUIRE(n_rays >= 0 && S >= 1 && capacity >= 1 && n_views >= 2 && n_views <= 4);
  const int64_t P = n_rays * S;
  if (P == 0) return MPSNERF_OK;
  MPS_REQUIRE(rays && t_vals && frame && grid_tp && grid_tv && skin_w && latent && img4 && packed);
  MPS_REQUIRE(raw && pts_mask && smpl_query && smpl_src && rgb && disp && acc && act_count && workspace);
  MPS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  MPS_REQUIRE(workspace_bytes >= mpsnerf_render_rays_workspace(n_rays, S, n_views, capacity));
  cudaStream_t st = (cudaStream_t)stream;
  mps::RenderWs w;
  mps::carve_render(w, static_cast<char*>(workspace), P, capacity, n_views);
  MPS_CUDA(cudaMemsetAsync(act_count, 0, sizeof(int32_t), st));
  int rc = mpsnerf_sample_knn(rays, n_rays, S, t_vals, u, nullptr, frame, grid_tp, raw, pts_mask, smpl_query, smpl_src, w.act_pid,
                              w.act_idx2, w.act_q, act_count, stream);
  if (rc != MPSNERF_OK) return rc;
  // the count goes to the host right behind K1 (not behind the frame): the caller can look at it while T / M still run
  if (host_count) MPS_CUDA(cudaMemcpyAsync(host_count, act_count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (event_count) MPS_CUDA(cudaEventRecord((cudaEvent_t)event_count, st));
  // the LBS transform sets / template grid and the encoder latent may be produced on other streams (engine.py)
  if (event_lbs) MPS_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)event_lbs, 0));
  rc = mpsnerf_deform_project_dc(w.act_pid, w.act_idx2, w.act_q, 0, capacity, act_count, skin_w, frame, grid_tv, w.xc, w.uv,
                                 smpl_src, stream);
  if (rc != MPSNERF_OK) return rc;
  if (event_trunk) MPS_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)event_trunk, 0));
  rc = mpsnerf_gather_tokens_f16_dc(w.uv, 0, capacity, act_count, n_views, frame, latent, img4, w.tokens, stream);
  if (rc != MPSNERF_OK) return rc;
  rc = mpsnerf_xformer_bf16_dc(w.tokens, w.xc, 0, capacity, act_count, n_views, packed, packed_bytes, w.act_pid, raw, w.dense,
                               stream);
  if (rc != MPSNERF_OK) return rc;
  rc = mpsnerf_mlp_bf16_dc(w.tokens, w.xc, 0, capacity, act_count, n_views, packed, packed_bytes, w.act_pid, raw, w.dense,
                           stream);
  if (rc != MPSNERF_OK) return rc;
  return mpsnerf_composite(raw, rays, n_rays, S, t_vals, u, nullptr, occupancy, rgb, disp, acc, depth, nullptr, nullptr, stream);
}

// Pointers into a mpsnerf_render_rays_bf16 workspace: the compacted active list the frame produced (needed by a caller
// that must run the points beyond `capacity` through the staged entry points).
extern "C" int mpsnerf_render_rays_active_list(void* workspace, int64_t n_rays, int32_t S, int n_views, int64_t capacity,
                                               int32_t** act_pid, int32_t** act_idx2, float** act_q) {
  MPS_REQUIRE(workspace && act_pid && act_idx2 && act_q);
  mps::RenderWs w;
  mps::carve_render(w, static_cast<char*>(workspace), n_rays * (int64_t)S, capacity, n_views);
  *act_pid = w.act_pid; *act_idx2 = w.act_idx2; *act_q = w.act_q;
  return MPSNERF_OK;
}
