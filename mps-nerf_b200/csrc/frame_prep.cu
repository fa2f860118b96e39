// K0: per-frame constants on the device (no host round trip).
//
// Restates get_transform_params_torch and its helpers (lib/run_nerf_helpers.py:174-254:
// shape blend -> joint regression -> Rodrigues -> 24-joint kinematic chain -> rest-joint
// removal) and big_pose_params (lib/skinnning_batch.py:193-201) for the four transform sets a
// (source, target) pair needs, and fills mpsnerf_frame in place.  The reference recomputes
// these on the host 4x per chunk; here it is one single-CTA kernel per render() call.
#include "common.cuh"

namespace mps {

constexpr int kPrepThreads = 1024;

struct PrepArgs {
  const float* poses[2];    // [0] target, [1] source: (72)
  const float* shapes[2];   // (10)
  const float* R[2];        // (9)
  const float* Th[2];       // (3)
  const float* cam_R;       // (V,9)
  const float* cam_T;       // (V,3)
  const float* cam_K;       // (V,9)
  const float* v_template;  // (nv,3)
  const float* shapedirs;   // (nv,3,10)
  const float* J_regressor; // (24,nv)
  const int32_t* parents;   // (24)
  int nv, n_views, img_w, img_h, feat_w, feat_h;
  mpsnerf_frame* out;
};

__device__ void rodrigues(const float* r, float* R) {     // run_nerf_helpers.py:174-192
  const float x = r[0] + 1e-8f, y = r[1] + 1e-8f, z = r[2] + 1e-8f;
  const float angle = sqrtf(x * x + y * y + z * z);
  const float kx = r[0] / angle, ky = r[1] / angle, kz = r[2] / angle;
  const float s = sinf(angle), c = 1.0f - cosf(angle);
  const float K[9] = {0.f, -kz, ky, kz, 0.f, -kx, -ky, kx, 0.f};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float kk = 0.f;
      for (int k = 0; k < 3; ++k) kk += K[3 * i + k] * K[3 * k + j];
      R[3 * i + j] = (i == j ? 1.0f : 0.0f) + s * K[3 * i + j] + c * kk;
    }
}

// 24-joint chain for one pose vector; joints (24,3) in smem; writes 24 x (3x4) to A.
__device__ void lbs_chain(const float* poses, bool big_pose, const float* joints, const int32_t* parents, float* G /*24*12 scratch*/,
                          float* A) {
  for (int j = 0; j < MPSNERF_NUM_JOINTS; ++j) {
    float r[3] = {0.f, 0.f, 0.f};
    if (!big_pose) { r[0] = poses[3 * j]; r[1] = poses[3 * j + 1]; r[2] = poses[3 * j + 2]; }
    else {   // lib/skinnning_batch.py:193-201: flat indices 5, 8 = +-45 deg, 23, 26 = -+30 deg
      const float pi = 3.14159265358979323846f;
      for (int k = 0; k < 3; ++k) {
        const int f = 3 * j + k;
        r[k] = f == 5 ? 45.f / 180.f * pi : f == 8 ? -45.f / 180.f * pi : f == 23 ? -30.f / 180.f * pi : f == 26 ? 30.f / 180.f * pi : 0.f;
      }
    }
    float T[12];
    float Rm[9];
    rodrigues(r, Rm);
    const int p = j == 0 ? -1 : parents[j];
    for (int i = 0; i < 3; ++i) {
      for (int k = 0; k < 3; ++k) T[4 * i + k] = Rm[3 * i + k];
      T[4 * i + 3] = joints[3 * j + i] - (j == 0 ? 0.f : joints[3 * p + i]);
    }
    float* Gj = G + 12 * j;
    if (j == 0) {
      for (int e = 0; e < 12; ++e) Gj[e] = T[e];
    } else {
      const float* Gp = G + 12 * p;
      for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 4; ++k) {
          float v = Gp[4 * i] * T[k] + Gp[4 * i + 1] * T[4 + k] + Gp[4 * i + 2] * T[8 + k];
          if (k == 3) v += Gp[4 * i + 3];
          Gj[4 * i + k] = v;
        }
    }
  }
  for (int j = 0; j < MPSNERF_NUM_JOINTS; ++j) {
    const float* Gj = G + 12 * j;
    for (int i = 0; i < 3; ++i) {
      for (int k = 0; k < 3; ++k) A[12 * j + 4 * i + k] = Gj[4 * i + k];
      A[12 * j + 4 * i + 3] = Gj[4 * i + 3] - (Gj[4 * i] * joints[3 * j] + Gj[4 * i + 1] * joints[3 * j + 1] + Gj[4 * i + 2] * joints[3 * j + 2]);
    }
  }
}

__global__ void __launch_bounds__(kPrepThreads, 1) frame_prep_kernel(const PrepArgs a) {
  extern __shared__ float s_vshaped[];               // (nv,3)
  __shared__ float s_joints[2][MPSNERF_NUM_JOINTS * 3];
  __shared__ float s_G[4][MPSNERF_NUM_JOINTS * 12];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int s = 0; s < 2; ++s) {
    float beta[10];
    for (int k = 0; k < 10; ++k) beta[k] = a.shapes[s][k];
    for (int e = tid; e < a.nv * 3; e += kPrepThreads) {
      float v = 0.f;
      for (int k = 0; k < 10; ++k) v += a.shapedirs[(size_t)e * 10 + k] * beta[k];
      s_vshaped[e] = a.v_template[e] + v;
    }
    __syncthreads();
    for (int o = warp; o < MPSNERF_NUM_JOINTS * 3; o += kPrepThreads / 32) {
      const int j = o / 3, c = o % 3;
      float acc = 0.f;
      for (int v = lane; v < a.nv; v += 32) acc = fmaf(a.J_regressor[(size_t)j * a.nv + v], s_vshaped[3 * v + c], acc);
      for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
      if (lane == 0) s_joints[s][o] = acc;
    }
    __syncthreads();
  }
  mpsnerf_frame* f = a.out;
  if (tid == 0) lbs_chain(a.poses[0], false, s_joints[0], a.parents, s_G[0], f->A_tp);
  if (tid == 32) lbs_chain(nullptr, true, s_joints[0], a.parents, s_G[1], f->A_big_tp);
  if (tid == 64) lbs_chain(nullptr, true, s_joints[1], a.parents, s_G[2], f->A_big_sp);
  if (tid == 96) lbs_chain(a.poses[1], false, s_joints[1], a.parents, s_G[3], f->A_sp);
  if (tid == 128) {
    for (int k = 0; k < 3; ++k) { f->Th_tp[k] = a.Th[0][k]; f->Th_sp[k] = a.Th[1][k]; }
    for (int k = 0; k < 9; ++k) f->R_tp[k] = a.R[0][k];
    const float* m = a.R[1];                       // inverse of R_sp (torch.inverse at lib/skinnning_batch.py:297)
    const float c00 = m[4] * m[8] - m[5] * m[7], c01 = m[2] * m[7] - m[1] * m[8], c02 = m[1] * m[5] - m[2] * m[4];
    const float c10 = m[5] * m[6] - m[3] * m[8], c11 = m[0] * m[8] - m[2] * m[6], c12 = m[2] * m[3] - m[0] * m[5];
    const float c20 = m[3] * m[7] - m[4] * m[6], c21 = m[1] * m[6] - m[0] * m[7], c22 = m[0] * m[4] - m[1] * m[3];
    const float rd = 1.0f / (m[0] * c00 + m[1] * c10 + m[2] * c20);
    const float inv[9] = {c00 * rd, c01 * rd, c02 * rd, c10 * rd, c11 * rd, c12 * rd, c20 * rd, c21 * rd, c22 * rd};
    for (int k = 0; k < 9; ++k) f->Rinv_sp[k] = inv[k];
    f->n_views = a.n_views; f->img_w = a.img_w; f->img_h = a.img_h; f->feat_w = a.feat_w; f->feat_h = a.feat_h;
    f->reserved[0] = f->reserved[1] = f->reserved[2] = 0;
  }
  for (int e = tid; e < a.n_views * 9; e += kPrepThreads) { f->cam_R[e] = a.cam_R[e]; f->cam_K[e] = a.cam_K[e]; }
  for (int e = tid; e < a.n_views * 3; e += kPrepThreads) f->cam_T[e] = a.cam_T[e];
}

}  // namespace mps

extern "C" int mpsnerf_frame_prepare(const float* poses_tp, const float* shapes_tp, const float* R_tp, const float* Th_tp,
                                     const float* poses_sp, const float* shapes_sp, const float* R_sp, const float* Th_sp,
                                     const float* cam_R, const float* cam_T, const float* cam_K, int n_views,
                                     int img_w, int img_h, int feat_w, int feat_h, const float* v_template,
                                     const float* shapedirs, const float* J_regressor, const int32_t* parents,
                                     int n_verts, mpsnerf_frame* out, void* stream) {
  MPS_REQUIRE(poses_tp && shapes_tp && R_tp && Th_tp && poses_sp && shapes_sp && R_sp && Th_sp);
  MPS_REQUIRE(cam_R && cam_T && cam_K && v_template && shapedirs && J_regressor && parents && out);
  MPS_REQUIRE(n_views >= 1 && n_views <= MPSNERF_MAX_VIEWS && n_verts > 0 && n_verts * 12 <= 200 * 1024);
  mps::PrepArgs a;
  a.poses[0] = poses_tp; a.poses[1] = poses_sp;
  a.shapes[0] = shapes_tp; a.shapes[1] = shapes_sp;
  a.R[0] = R_tp; a.R[1] = R_sp;
  a.Th[0] = Th_tp; a.Th[1] = Th_sp;
  a.cam_R = cam_R; a.cam_T = cam_T; a.cam_K = cam_K;
  a.v_template = v_template; a.shapedirs = shapedirs; a.J_regressor = J_regressor; a.parents = parents;
  a.nv = n_verts; a.n_views = n_views; a.img_w = img_w; a.img_h = img_h; a.feat_w = feat_w; a.feat_h = feat_h;
  a.out = out;
  const size_t smem = (size_t)n_verts * 3 * sizeof(float);
  MPS_CUDA(cudaFuncSetAttribute(mps::frame_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mps::frame_prep_kernel<<<1, mps::kPrepThreads, smem, (cudaStream_t)stream>>>(a);
  MPS_LAUNCH_CHECK();
  return MPSNERF_OK;
}
