/*
 * mpsnerf.h -- C ABI of libmpsnerf_b200.so: the B200 (sm_100a) implementation of
 * MPS-NeRF's per-ray render hot path.
 *
 * The reference (gaoxiangjun/MPS-NeRF) is pure Python/PyTorch and has no FFI layer,
 * so every entry point below replaces a *Python* function of the reference; the
 * file:line it replaces is cited per function (paths relative to the reference root).
 * The host-side mirror of the reference API (mpsnerf_b200.run_nerf_batch, .lib.*)
 * binds these symbols with ctypes; INTEGRATION.md shows the stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless named host_*; all floats are fp32, row major;
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *  - no entry point allocates, synchronises the device or throws; each returns 0 on success
 *    or a negative MPSNERF_E* code, with a message available from mpsnerf_last_error()
 *    (thread-local);
 *  - every entry point is re-entrant per stream: all state lives in caller-owned buffers.
 */
#ifndef MPSNERF_H_
#define MPSNERF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPSNERF_ABI_VERSION 1

#define MPSNERF_OK 0
#define MPSNERF_EINVAL (-1)   /* bad argument (null pointer, size out of range) */
#define MPSNERF_ECUDA (-2)    /* a CUDA runtime call or launch failed */
#define MPSNERF_EARCH (-3)    /* device is not sm_100 */
#define MPSNERF_ENCCL (-4)    /* libnccl could not be bound, or ncclAllReduce returned an error */

#define MPSNERF_MAX_VIEWS 8
#define MPSNERF_NUM_JOINTS 24
#define MPSNERF_GRID_MAX_DIM 64
#define MPSNERF_TOKEN_DIM 155       /* 128 latent + 27 rgb code (lib/skinnning_batch.py:161) */
#define MPSNERF_TOKEN_LD 160        /* padded row stride used by the tensor-core path */

/* Per-(source,target) frame constants, filled by the host once per render() call.
 * A_* are the 24 LBS transforms as 3x4 row-major blocks (lib/run_nerf_helpers.py:227-254):
 *   A_tp      target pose, target shape      (lib/skinnning_batch.py:206)
 *   A_big_tp  "big pose", target shape       (lib/skinnning_batch.py:243)
 *   A_big_sp  "big pose", source shape       (lib/skinnning_batch.py:266)
 *   A_sp      source pose, source shape      (lib/skinnning_batch.py:289) */
typedef struct mpsnerf_frame {
  float Th_tp[3];
  float R_tp[9];
  float Rinv_sp[9];
  float Th_sp[3];
  float A_tp[MPSNERF_NUM_JOINTS * 12];
  float A_big_tp[MPSNERF_NUM_JOINTS * 12];
  float A_big_sp[MPSNERF_NUM_JOINTS * 12];
  float A_sp[MPSNERF_NUM_JOINTS * 12];
  float cam_R[MPSNERF_MAX_VIEWS * 9];   /* R_all  (lib/skinnning_batch.py:177-184) */
  float cam_T[MPSNERF_MAX_VIEWS * 3];   /* T_all */
  float cam_K[MPSNERF_MAX_VIEWS * 9];   /* K_all */
  int32_t n_views;
  int32_t img_w, img_h;                 /* size of img_all: image_shape = [W, H] (:186-189) */
  int32_t feat_w, feat_h;               /* size of the encoder latent */
  int32_t reserved[3];
} mpsnerf_frame;

const char* mpsnerf_last_error(void);
int mpsnerf_abi_version(void);
/* 0 if device `dev` can run this library (compute capability 10.x), else MPSNERF_EARCH. */
int mpsnerf_check_device(int dev);

/* ---- K0: per-frame constants on the device ----------------------------------------------
 * Replaces get_transform_params_torch / batch_rodrigues_torch / get_rigid_transformation_torch
 * (lib/run_nerf_helpers.py:174-254) and big_pose_params (lib/skinnning_batch.py:193-201), which
 * the reference runs on the host 4x per chunk.  Inputs are the device tensors of the dataset
 * dicts (tp = target, sp = source: poses 72, shapes 10, R 9, Th 3; R_all/T_all/K_all of the
 * input views) and the SMPL tables (v_template (nv,3), shapedirs (nv,3,10), dense J_regressor
 * (24,nv), parents int32 (24)); fills *out (device). */
int mpsnerf_frame_prepare(const float* poses_tp, const float* shapes_tp, const float* R_tp, const float* Th_tp,
                          const float* poses_sp, const float* shapes_sp, const float* R_sp, const float* Th_sp,
                          const float* cam_R, const float* cam_T, const float* cam_K, int n_views,
                          int img_w, int img_h, int feat_w, int feat_h, const float* v_template,
                          const float* shapedirs, const float* J_regressor, const int32_t* parents,
                          int n_verts, mpsnerf_frame* out, void* stream);

/* The two halves of mpsnerf_frame_prepare as separate launches (calling both, in any order or on different streams,
 * equals the call above; they write disjoint fields of *out):
 *  - header: the fields that are copies of the inputs -- Th_tp, R_tp (lib/skinnning_batch.py:345-347), Th_sp, the 3x3
 *    inverse of R_sp (:297), cameras, sizes.  A few microseconds, and all that K1 needs of the frame.
 *  - transforms: the four LBS transform sets (lib/run_nerf_helpers.py:174-254, lib/skinnning_batch.py:193-201),
 *    evaluated in float64 and rounded to fp32 once.  Needed from K3 on. */
int mpsnerf_frame_header(const float* R_tp, const float* Th_tp, const float* R_sp, const float* Th_sp,
                         const float* cam_R, const float* cam_T, const float* cam_K, int n_views,
                         int img_w, int img_h, int feat_w, int feat_h, mpsnerf_frame* out, void* stream);
int mpsnerf_frame_transforms(const float* poses_tp, const float* shapes_tp, const float* poses_sp,
                             const float* shapes_sp, const float* v_template, const float* shapedirs,
                             const float* J_regressor, const int32_t* parents, int n_verts,
                             mpsnerf_frame* out, void* stream);

/* ---- nearest-vertex acceleration grid -------------------------------------------------
 * Replaces the brute-force pytorch3d knn_points calls (lib/skinnning_batch.py:214,256,357)
 * with an exact uniform-grid search.  If Th/R are non-null the vertices are first taken to
 * SMPL space, v' = (v - Th) @ R (lib/skinnning_batch.py:355-356), with pinned fp32 ops. */
size_t mpsnerf_grid_bytes(int n_verts);
int mpsnerf_grid_build(const float* verts, int n_verts, const float* Th, const float* R,
                       float cell, void* grid, size_t grid_bytes, void* stream);

/* Exact K=1 nearest vertex for arbitrary queries (diagnostic / mesh-extraction use;
 * extract_thuman_mesh.py:132).  d2 = (dx*dx+dy*dy)+dz*dz, ties -> lowest index. */
int mpsnerf_knn1(const float* query, int64_t n, const void* grid, float* d2_out,
                 int32_t* idx_out, void* stream);

/* Mesh-extraction post-step on the density grid (replaces extract_thuman_mesh.py:125-158: shifted_softplus,
 * knn_points K=1 mask, knn_points K=5 mean-normal inside/outside test, occupancy override).  One brute-force
 * pass over the n_verts vertices per point keeps the five nearest (pinned d2, ties -> lowest index).
 * raw: (n, raw_stride) floats, channel 3 = density logit; occupancy (n); optional outputs (may be NULL):
 * pts_mask (n) int32, outside (n) uint8, idx5 (n,5) int32 sorted by (d2, index), d2_nearest (n). */
int mpsnerf_occupancy_fix(const float* pts, int64_t n, const float* verts, const float* normals,
                          int32_t n_verts, const float* raw, int32_t raw_stride, float* occupancy,
                          int32_t* pts_mask, uint8_t* outside, int32_t* idx5, float* d2_nearest, void* stream);

/* ---- K1: stratified sampling + world->SMPL + human-region mask + argmin + compaction ---
 * Replaces render_rays sampling (run_nerf_batch.py:406-424), run_network flattening
 * (:42-52) and SKinningBatch.forward steps 2,4 (lib/skinnning_batch.py:345-365).
 *   rays     (n_rays, 8): o(3) d(3) near far
 *   t_vals   (S) = torch.linspace(0,1,S);  u (n_rays,S) uniforms or NULL (perturb == 0)
 *   points   optional (n_rays*S,3): if non-null the sample points are READ from here
 *            instead of being generated (network_fn called directly on points)
 * Outputs for all P = n_rays*S points:
 *   raw (P,4) = -80 where inactive (left untouched where active), pts_mask (P) 0/1,
 *   smpl_query (P,3), smpl_src (P,3) zero-filled for inactive points;
 * and the compacted active list (capacity P): act_pid, act_idx2, act_q (cap,3), *act_count.
 * act_count must be zeroed by the caller before the first launch of a frame. */
int mpsnerf_sample_knn(const float* rays, int64_t n_rays, int32_t S, const float* t_vals,
                       const float* u, const float* points, const mpsnerf_frame* frame,
                       const void* grid_tp, float* raw, float* pts_mask, float* smpl_query,
                       float* smpl_src, int32_t* act_pid, int32_t* act_idx2, float* act_q,
                       int32_t* act_count, void* stream);

/* ---- K3: inverse LBS target->canonical->source + projection ---------------------------
 * Replaces coarse_deform_target2c (lib/skinnning_batch.py:203-251), coarse_deform_c2source
 * (:253-300) and projection (:177-184) for mean_shape = 0.
 *   skin_w (n_verts,24) SMPL blend weights.  Works on active points [first, first+count).
 * Outputs: xc (count,3) canonical points, uv (count,V,2) pixels, smpl_src[pid] scattered;
 * optional idx3 / xw (count) for diagnostics (may be NULL).
 * identity_canonical != 0 reproduces extract_mesh mode (:394-396): canonical = query point. */
int mpsnerf_deform_project(const int32_t* act_pid, const int32_t* act_idx2, const float* act_q,
                           int64_t first, int64_t count, const float* skin_w,
                           const mpsnerf_frame* frame, const void* grid_tv, float* xc, float* uv,
                           float* smpl_src, int32_t* idx3, float* xw, int identity_canonical,
                           void* stream);

/* ---- K-1: ray generation + box near/far (in front of K1) -------------------------------
 * Replaces get_rays and get_near_far (lib/if_nerf_data_utils.py:11-25, 55-92), numpy on the CPU in the
 * reference.  K, R (3x3 row-major), T (3), bounds (2,3 = min,max; widened by 0.01 inside like the
 * reference) are HOST doubles.  rays8 (H*W, 8) = [o, d, near, far] fp32, pixel (row j, column i) at
 * j*W + i; rays that do not hit the box exactly twice get near = 0, far = 1 and mask_at_box = 0
 * (mask_at_box may be NULL). */
int mpsnerf_gen_rays(const double* K, const double* R, const double* T, const double* bounds, int32_t H,
                     int32_t W, float* rays8, uint8_t* mask_at_box, void* stream);

/* Same for a subset of the image rows: rows (n_rows) is a DEVICE list of row indices in [0, H); output ray
 * p = pixel (rows[p / W], p % W), so rays8 holds n_rows * W rays.  This is how one target view is dealt out to
 * several GPUs (each rank generates only its own rows; no ray upload, no scatter). */
int mpsnerf_gen_rays_rows(const double* K, const double* R, const double* T, const double* bounds, int32_t H,
                          int32_t W, const int32_t* rows, int32_t n_rows, float* rays8, uint8_t* mask_at_box,
                          void* stream);

/* ---- K4: multiview bilinear feature + RGB lookup, RGB positional code -> tokens --------
 * Replaces SpatialEncoder.index / grid_sample (lib/encoder.py:12-62, 225-253) and the RGB
 * append (lib/skinnning_batch.py:428-435).  latent is NHWC (V,Hf,Wf,128); img is NHWC with
 * 4 floats per pixel (V,H,W,4: r,g,b,0).  tokens (count, V, ld) with ld >= 155; pad = 0. */
int mpsnerf_gather_tokens(const float* uv, int64_t count, int n_views, const mpsnerf_frame* frame,
                          const float* latent, const float* img4, float* tokens, int32_t ld,
                          void* stream);
/* Same lookup, tokens written as IEEE fp16 (count, V, 160), pad = 0, values clamped to +-65504:
 * the input format of the tensor-core path (mpsnerf_dense_bf16). */
int mpsnerf_gather_tokens_f16(const float* uv, int64_t count, int n_views, const mpsnerf_frame* frame,
                              const float* latent, const float* img4, void* tokens, void* stream);

/* ---- K5: cross-view transformer + canonical NeRF MLP -----------------------------------
 * Replaces Transformer.forward (lib/transformer.py:74-86) and the MLP of
 * SKinningBatch.forward (lib/skinnning_batch.py:438-473); scatters [rgb, alpha] to
 * raw[act_pid[first + i]].
 * fp32 variant: SIMT, plain torch-layout weights passed as an array of device pointers in
 * the order documented in mpsnerf_b200/engine.py::DENSE_FP32_ORDER.  workspace >=
 * mpsnerf_dense_fp32_workspace(count) bytes. */
size_t mpsnerf_dense_fp32_workspace(int64_t count, int n_views);
int mpsnerf_dense_fp32(const float* tokens, int32_t ld, const float* xc, int64_t count,
                       int n_views, const float* const* weights, const int32_t* act_pid,
                       int64_t first, float* raw, void* workspace, void* stream);

/* ---- training step: the dense stage with saved intermediates, and the backward of K6 / dense / K4 -----------
 * Replaces what autograd does for the reference's training step (run_nerf_batch.py:544-570:
 * render -> img2mse(rgb) + img2mse(acc) -> loss.backward()) on the render path.  fp32 on CUDA cores: a training
 * batch is ~1 K rays / ~5 K active points per GPU.  Under the shipped configs (skinning_field = correction_field = 0)
 * no parameter sits upstream of the canonical points, so K1 / K3 have no backward; gradients reach the live
 * transformer + MLP parameters (mpsnerf_dense_train_bwd) and, through the latent, the encoder trunk
 * (mpsnerf_gather_tokens_bwd -> torch / cuDNN).
 *  - dense_train_fwd: as mpsnerf_dense_fp32, but writes out4 (count, 4) = [rgb, alpha] per active point instead
 *    of scattering, and keeps every intermediate in `workspace` (>= mpsnerf_dense_train_workspace bytes) for
 *  - dense_train_bwd: d_out4 (count, 4) -> `grads` (46 device pointers in the order of `weights`, each ACCUMULATED
 *    into) and d_tokens (count, V, 155) (overwritten);
 *  - composite_bwd: d_rgb (n_rays, 3), d_acc (n_rays, may be NULL) -> d_raw (n_rays, S, 4) (overwritten; zero for the
 *    -80-filled samples of masked-out points); same sampling arguments as mpsnerf_composite;
 *  - gather_tokens_bwd: scatter-add of d_tokens[:, :, 0:128] (row stride ld) into d_latent (V, Hf, Wf, 128), which the
 *    caller zeroes; the RGB half of a token depends on the input images only;
 *  - rows4_gather / rows4_scatter: dst[i] = src[act_pid[i]] / dst[act_pid[i]] = src[i] on float4 rows (raw <-> out4). */
size_t mpsnerf_dense_train_workspace(int64_t count, int n_views);
int mpsnerf_dense_train_fwd(const float* tokens, int32_t ld, const float* xc, int64_t count, int n_views,
                            const float* const* weights, float* out4, void* workspace, void* stream);
int mpsnerf_dense_train_bwd(const float* d_out4, int64_t count, int n_views, const float* const* weights,
                            float* const* grads, float* d_tokens, void* workspace, void* stream);
int mpsnerf_composite_bwd(const float* raw, const float* rays, int64_t n_rays, int32_t S, const float* t_vals,
                          const float* u, const float* z_vals, int occupancy, const float* d_rgb,
                          const float* d_acc, float* d_raw, void* stream);
int mpsnerf_gather_tokens_bwd(const float* uv, int64_t count, int n_views, const mpsnerf_frame* frame,
                              const float* d_tokens, int32_t ld, float* d_latent, void* stream);
int mpsnerf_rows4_gather(const float* src, const int32_t* act_pid, int64_t count, float* dst, void* stream);
int mpsnerf_rows4_scatter(const float* src, const int32_t* act_pid, int64_t count, float* dst, void* stream);

/* The gradient all-reduce of data-parallel training -- the one collective of the path.  Replaces what
 * DistributedDataParallel does behind loss.backward() (run_nerf_batch.py:344-348, :560): `bucket` (count floats, e.g. the
 * flat buffer the dense backward accumulated the 46 gradients into) is averaged IN PLACE over the ranks of `nccl_comm`
 * (an ncclComm_t owned by the host, passed as void*), enqueued on `stream` behind the backward kernels: one
 * ncclAllReduce with ncclAvg.  The library does not link NCCL: the symbol is bound at first use from the libnccl already
 * loaded in the process (else libnccl.so.2 from the loader path); MPSNERF_ENCCL if that fails or NCCL reports an error.
 * The Python mirror (mps-nerf_b200/train.py) uses torch.distributed's communicator instead and does not call this. */
int mpsnerf_allreduce_mean(void* nccl_comm, float* bucket, int64_t count, void* stream);

/* bf16 tensor-core variant (tcgen05 / TMEM / bulk-async weight streaming).  `packed` is the
 * blob produced by mpsnerf_b200.engine.pack_weights_bf16 (layout: DESIGN.md section 5). */
size_t mpsnerf_dense_bf16_workspace(int64_t count, int n_views);
/* tokens: fp16 (count, V, 160) as written by mpsnerf_gather_tokens_f16; ld must be 160. */
int mpsnerf_dense_bf16(const void* tokens, int32_t ld, const float* xc, int64_t count,
                       int n_views, const void* packed, size_t packed_bytes,
                       const int32_t* act_pid, int64_t first, float* raw, void* workspace,
                       void* stream);
/* The two halves of mpsnerf_dense_bf16, same arguments: the cross-view transformer alone (tokens -> the two
 * output tokens per point, bf16, in the workspace; lib/transformer.py:74-86) and the NeRF MLP alone
 * (workspace -> raw; lib/skinnning_batch.py:449-473).  Calling one after the other equals the fused call. */
int mpsnerf_xformer_bf16(const void* tokens, int32_t ld, const float* xc, int64_t count,
                       int n_views, const void* packed, size_t packed_bytes,
                       const int32_t* act_pid, int64_t first, float* raw, void* workspace,
                       void* stream);
int mpsnerf_mlp_bf16(const void* tokens, int32_t ld, const float* xc, int64_t count,
                       int n_views, const void* packed, size_t packed_bytes,
                       const int32_t* act_pid, int64_t first, float* raw, void* workspace,
                       void* stream);

/* ---- device-side active count ("_dc") variants of K3, K4 and K5 ------------------------
 * Same work as the functions above for the active-list positions [first, first + n), where
 * n = clamp(*count_dev - first, 0, capacity) is read ON THE DEVICE (count_dev = the counter K1 increments):
 * the host can enqueue a whole frame without waiting for K1.  Buffers must hold `capacity` points; the
 * debug outputs (idx3, xw) are not available.  The caller checks *count_dev <= first + capacity afterwards
 * and runs the remainder, if any, through the host-count entry points. */
int mpsnerf_deform_project_dc(const int32_t* act_pid, const int32_t* act_idx2, const float* act_q,
                              int64_t first, int64_t capacity, const int32_t* count_dev, const float* skin_w,
                              const mpsnerf_frame* frame, const void* grid_tv, float* xc, float* uv,
                              float* smpl_src, void* stream);
int mpsnerf_gather_tokens_f16_dc(const float* uv, int64_t first, int64_t capacity, const int32_t* count_dev,
                                 int n_views, const mpsnerf_frame* frame, const float* latent,
                                 const float* img4, void* tokens, void* stream);
int mpsnerf_xformer_bf16_dc(const void* tokens, const float* xc, int64_t first, int64_t capacity,
                            const int32_t* count_dev, int n_views, const void* packed, size_t packed_bytes,
                            const int32_t* act_pid, float* raw, void* workspace, void* stream);
int mpsnerf_mlp_bf16_dc(const void* tokens, const float* xc, int64_t first, int64_t capacity,
                        const int32_t* count_dev, int n_views, const void* packed, size_t packed_bytes,
                        const int32_t* act_pid, float* raw, void* workspace, void* stream);

/* ---- fused render_rays: the whole per-ray path of a frame in ONE call ------------------------------------------
 * Replaces render_rays (run_nerf_batch.py:401-444: sampling -> network_query_fn -> raw2outputs) for the tensor-core
 * path: enqueues K1, K3, K4, T, M and K6 on `stream`, every stage reading the active count on the device.
 * Inputs: rays (n_rays, 8), t_vals (S), u (n_rays, S) or NULL; the prepared frame state (frame, the two grids, skin_w,
 * NHWC latent, img4) and the packed weights.  Outputs: the per-sample extras (raw, pts_mask, smpl_query, smpl_src; P =
 * n_rays * S rows), the per-ray maps (rgb, disp, acc, depth; depth may be NULL) and *act_count (device) = the number
 * of active points.  `capacity` = active points the workspace holds; points beyond it are NOT evaluated -- the caller
 * compares *act_count with capacity afterwards and, if it was exceeded, runs the remainder through the staged entry
 * points (active list: mpsnerf_render_rays_active_list) and mpsnerf_composite again.  event_lbs / event_trunk:
 * optional cudaEvent_t the stream waits on before K3 / K4 (the frame preparation may run on other streams).
 * host_count (pinned host memory) / event_count (cudaEvent_t), both optional: the count is copied to the host and the
 * event recorded right behind K1, so the caller can read it long before the frame has finished.
 * No allocation, no synchronisation; workspace >= mpsnerf_render_rays_workspace(...) bytes, 256-byte aligned. */
size_t mpsnerf_render_rays_workspace(int64_t n_rays, int32_t S, int n_views, int64_t capacity);
int mpsnerf_render_rays_bf16(const float* rays, int64_t n_rays, int32_t S, const float* t_vals, const float* u,
                             const mpsnerf_frame* frame, const void* grid_tp, const void* grid_tv,
                             const float* skin_w, const float* latent, const float* img4, const void* packed,
                             size_t packed_bytes, int n_views, int occupancy, float* raw, float* pts_mask,
                             float* smpl_query, float* smpl_src, float* rgb, float* disp, float* acc, float* depth,
                             int32_t* act_count, int64_t capacity, void* workspace, size_t workspace_bytes,
                             void* event_lbs, void* event_trunk, int32_t* host_count, void* event_count,
                             void* stream);
int mpsnerf_render_rays_active_list(void* workspace, int64_t n_rays, int32_t S, int n_views, int64_t capacity,
                                    int32_t** act_pid, int32_t** act_idx2, float** act_q);

/* ---- K6: alpha compositing --------------------------------------------------------------
 * Replaces raw2outputs (run_nerf_batch.py:369-398).  One warp per ray.
 *   raw (n_rays,S,4); z is regenerated from rays/t_vals/u exactly as in K1, or read from
 *   z_vals (n_rays,S) when that is non-null (rays then only supplies the directions).
 * Outputs rgb (n_rays,3), disp, acc, depth (n_rays; depth may be NULL); optional (NULL ok)
 * weights and transmittance T_s, both (n_rays,S). */
int mpsnerf_composite(const float* raw, const float* rays, int64_t n_rays, int32_t S,
                      const float* t_vals, const float* u, const float* z_vals, int occupancy,
                      float* rgb, float* disp, float* acc, float* depth, float* weights,
                      float* trans, void* stream);

/* Diagnostic: one 128 x N x K bf16 tcgen05 GEMM tile through the same shared-memory layout,
 * descriptors and TMEM epilogue the fused kernels use.  a (128,K) row-major bf16 (as uint16);
 * b_packed = the (N,K) weight in the SWIZZLE_128B chunk layout produced by
 * mpsnerf_b200.engine.pack_kmajor_sw128; d (128,N) fp32.  K % 64 == 0, N % 16 == 0, N <= 256. */
int mpsnerf_selftest_umma(const uint16_t* a, const uint8_t* b_packed, float* d, int N, int K,
                          void* stream);

/* Same diagnostic with the A operand staged in tensor memory at column `acol` (tcgen05.st +
 * TS-form MMA); acol >= N, acol % 8 == 0, acol + K/2 <= 512. */
int mpsnerf_selftest_umma_ts(const uint16_t* a, const uint8_t* b_packed, float* d, int N, int K,
                             int acol, void* stream);

/* Debug: with MPSNERF_TC_PROF=1 the tensor-core kernels accumulate per-role cycle counters
 * (see csrc/dense_tc.cu); this copies the 2 x 16 counters to the host and clears them. */
int mpsnerf_debug_read_prof(unsigned long long* host_out);
/* Debug: copy the 512-entry event trace recorded with MPSNERF_TC_PROF=2 (tag << 48 | clock). */
int mpsnerf_debug_read_trace(unsigned long long* host_out);

#ifdef __cplusplus
}
#endif
#endif /* MPSNERF_H_ */
