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

// Arithmetic contract of K0: the reference's formulas evaluated in float64 on the fp32
// inputs and rounded to fp32 ONCE, when the result is stored into mpsnerf_frame.  The oracle
// (oracle/oracle.py::frame_constants) does the same in numpy float64; the two agree bit for bit (a float64
// re-association or a last-bit sin/cos difference moves a result by ~1e-16 relative, 2^-29 of an fp32 ulp), and
// both sit within ~1 fp32 ulp of the reference's own fp32 evaluation.  That makes every downstream index
// (kNN #3 on the canonical points) exact against the oracle end to end, not only on substituted constants.
__device__ void rodrigues(const double* r, double* R) {     // run_nerf_helpers.py:174-192
  const double x = r[0] + 1e-8, y = r[1] + 1e-8, z = r[2] + 1e-8;
  const double angle = sqrt(x * x + y * y + z * z);
  const double kx = r[0] / angle, ky = r[1] / angle, kz = r[2] / angle;
  const double s = sin(angle), c = 1.0 - cos(angle);
  const double K[9] = {0., -kz, ky, kz, 0., -kx, -ky, kx, 0.};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double kk = 0.;
      for (int k = 0; k < 3; ++k) kk += K[3 * i + k] * K[3 * k + j];
      R[3 * i + j] = (i == j ? 1.0 : 0.0) + s * K[3 * i + j] + c * kk;
    }
}

// axis-angle vector of joint j: the pose vector, or the fixed "big pose" of lib/skinnning_batch.py:193-201
// (flat indices 5, 8 = +-45 deg, 23, 26 = -+30 deg; the reference forms them as fp32 values)
__device__ void joint_axis_angle(const float* poses, bool big_pose, int j, double* r) {
  for (int k = 0; k < 3; ++k) {
    const int f = 3 * j + k;
    if (!big_pose) r[k] = (double)poses[f];
    else {
      const float pi = 3.14159265358979323846f;
      r[k] = (double)(f == 5 ? 45.f / 180.f * pi : f == 8 ? -45.f / 180.f * pi : f == 23 ? -30.f / 180.f * pi : f == 26 ? 30.f / 180.f * pi : 0.f);
    }
  }
}

// 24-joint chain; rot (24,9) joint rotations, joints (24,3); G = 24 x (3x4) scratch; writes 24 x (3x4) fp32 to A.
__device__ void lbs_chain(const double* rot, const double* joints, const int32_t* parents, double* G, float* A) {
  for (int j = 0; j < MPSNERF_NUM_JOINTS; ++j) {
    double T[12];
    const double* Rm = rot + 9 * j;
    const int p = j == 0 ? -1 : parents[j];
    for (int i = 0; i < 3; ++i) {
      for (int k = 0; k < 3; ++k) T[4 * i + k] = Rm[3 * i + k];
      T[4 * i + 3] = joints[3 * j + i] - (j == 0 ? 0. : joints[3 * p + i]);
    }
    double* Gj = G + 12 * j;
    if (j == 0) {
      for (int e = 0; e < 12; ++e) Gj[e] = T[e];
    } else {
      const double* Gp = G + 12 * p;
      for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 4; ++k) {
          double v = Gp[4 * i] * T[k] + Gp[4 * i + 1] * T[4 + k] + Gp[4 * i + 2] * T[8 + k];
          if (k == 3) v += Gp[4 * i + 3];
          Gj[4 * i + k] = v;
        }
    }
  }
  for (int j = 0; j < MPSNERF_NUM_JOINTS; ++j) {
    const double* Gj = G + 12 * j;
    for (int i = 0; i < 3; ++i) {
      for (int k = 0; k < 3; ++k) A[12 * j + 4 * i + k] = (float)Gj[4 * i + k];
      A[12 * j + 4 * i + 3] = (float)(Gj[4 * i + 3] - (Gj[4 * i] * joints[3 * j] + Gj[4 * i + 1] * joints[3 * j + 1] + Gj[4 * i + 2] * joints[3 * j + 2]));
    }
  }
}

__global__ void __launch_bounds__(kPrepThreads, 1) frame_prep_kernel(const PrepArgs a) {
  extern __shared__ double s_vshaped[];               // (nv,3)
  __shared__ double s_joints[2][MPSNERF_NUM_JOINTS * 3];
  __shared__ double s_rot[4][MPSNERF_NUM_JOINTS * 9];
  __shared__ double s_G[4][MPSNERF_NUM_JOINTS * 12];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // joint rotations of the four transform sets: one thread per (set, joint)
  if (tid < 4 * MPSNERF_NUM_JOINTS) {
    const int set = tid / MPSNERF_NUM_JOINTS, j = tid % MPSNERF_NUM_JOINTS;     // 0 A_tp, 1 A_big_tp, 2 A_big_sp, 3 A_sp
    double r[3];
    joint_axis_angle(set == 0 ? a.poses[0] : a.poses[1], set == 1 || set == 2, j, r);
    rodrigues(r, &s_rot[set][9 * j]);
  }
  for (int s = 0; s < 2; ++s) {
    double beta[10];
    for (int k = 0; k < 10; ++k) beta[k] = (double)a.shapes[s][k];
    for (int e = tid; e < a.nv * 3; e += kPrepThreads) {
      double v = 0.;
      for (int k = 0; k < 10; ++k) v += (double)a.shapedirs[(size_t)e * 10 + k] * beta[k];
      s_vshaped[e] = (double)a.v_template[e] + v;
    }
    __syncthreads();
    for (int o = warp; o < MPSNERF_NUM_JOINTS * 3; o += kPrepThreads / 32) {
      const int j = o / 3, c = o % 3;
      double acc = 0.;
      for (int v = lane; v < a.nv; v += 32) acc = fma((double)a.J_regressor[(size_t)j * a.nv + v], s_vshaped[3 * v + c], acc);
      for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
      if (lane == 0) s_joints[s][o] = acc;
    }
    __syncthreads();
  }
  mpsnerf_frame* f = a.out;
  if (tid == 0) lbs_chain(s_rot[0], s_joints[0], a.parents, s_G[0], f->A_tp);
  if (tid == 32) lbs_chain(s_rot[1], s_joints[0], a.parents, s_G[1], f->A_big_tp);
  if (tid == 64) lbs_chain(s_rot[2], s_joints[1], a.parents, s_G[2], f->A_big_sp);
  if (tid == 96) lbs_chain(s_rot[3], s_joints[1], a.parents, s_G[3], f->A_sp);
}

// The fields that are copies of the inputs (rigid transforms of target and source, cameras, sizes) plus the 3x3
// inverse of R_sp: a few microseconds, and all that K1 (sample + mask) needs of the frame -- so the engine
// launches it alone in front of K1 while the LBS chains above run beside it.
__global__ void __launch_bounds__(128, 1) frame_header_kernel(const PrepArgs a) {
  const int tid = threadIdx.x;
  mpsnerf_frame* f = a.out;
  if (tid == 0) {
    for (int k = 0; k < 3; ++k) { f->Th_tp[k] = a.Th[0][k]; f->Th_sp[k] = a.Th[1][k]; }
    for (int k = 0; k < 9; ++k) f->R_tp[k] = a.R[0][k];
    double m[9];                                   // inverse of R_sp (torch.inverse at lib/skinnning_batch.py:297)
    for (int k = 0; k < 9; ++k) m[k] = (double)a.R[1][k];
    const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[2] * m[7] - m[1] * m[8], c02 = m[1] * m[5] - m[2] * m[4];
    const double c10 = m[5] * m[6] - m[3] * m[8], c11 = m[0] * m[8] - m[2] * m[6], c12 = m[2] * m[3] - m[0] * m[5];
    const double c20 = m[3] * m[7] - m[4] * m[6], c21 = m[1] * m[6] - m[0] * m[7], c22 = m[0] * m[4] - m[1] * m[3];
    const double rd = 1.0 / (m[0] * c00 + m[1] * c10 + m[2] * c20);
    const double inv[9] = {c00 * rd, c01 * rd, c02 * rd, c10 * rd, c11 * rd, c12 * rd, c20 * rd, c21 * rd, c22 * rd};
    for (int k = 0; k < 9; ++k) f->Rinv_sp[k] = (float)inv[k];
    f->n_views = a.n_views; f->img_w = a.img_w; f->img_h = a.img_h; f->feat_w = a.feat_w; f->feat_h = a.feat_h;
    f->reserved[0] = f->reserved[1] = f->reserved[2] = 0;
  }
  for (int e = tid; e < a.n_views * 9; e += 128) { f->cam_R[e] = a.cam_R[e]; f->cam_K[e] = a.cam_K[e]; }
  for (int e = tid; e < a.n_views * 3; e += 128) f->cam_T[e] = a.cam_T[e];
}

}  // namespace mps

extern "C" int mpsnerf_frame_header(const float* R_tp, const float* Th_tp, const float* R_sp, const float* Th_sp,
                                    const float* cam_R, const float* cam_T, const float* cam_K, int n_views,
                                    int img_w, int img_h, int feat_w, int feat_h, mpsnerf_frame* out, void* stream) {
  MPS_REQUIRE(R_tp && Th_tp && R_sp && Th_sp && cam_R && cam_T && cam_K && out);
  MPS_REQUIRE(n_views >= 1 && n_views <= MPSNERF_MAX_VIEWS);
  mps::PrepArgs a{};
  a.R[0] = R_tp; a.R[1] = R_sp;
  a.Th[0] = Th_tp; a.Th[1] = Th_sp;
  a.cam_R = cam_R; a.cam_T = cam_T; a.cam_K = cam_K;
  a.n_views = n_views; a.img_w = img_w; a.img_h = img_h; a.feat_w = feat_w; a.feat_h = feat_h;
  a.out = out;
  mps::frame_header_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(a);
  MPS_LAUNCH_CHECK();
  return MPSNERF_OK;
}

extern "C" int mpsnerf_frame_transforms(const float* poses_tp, const float* shapes_tp, const float* poses_sp,
                                        const float* shapes_sp, const float* v_template, const float* shapedirs,
                                        const float* J_regressor, const int32_t* parents, int n_verts,
                                        mpsnerf_frame* out, void* stream) {
  MPS_REQUIRE(poses_tp && shapes_tp && poses_sp && shapes_sp);
  MPS_REQUIRE(v_template && shapedirs && J_regressor && parents && out);
  MPS_REQUIRE(n_verts > 0 && n_verts * 24 <= 200 * 1024);
  mps::PrepArgs a{};
  a.poses[0] = poses_tp; a.poses[1] = poses_sp;
  a.shapes[0] = shapes_tp; a.shapes[1] = shapes_sp;
  a.v_template = v_template; a.shapedirs = shapedirs; a.J_regressor = J_regressor; a.parents = parents;
  a.nv = n_verts;
  a.out = out;
  const size_t smem = (size_t)n_verts * 3 * sizeof(double);
  MPS_CUDA(cudaFuncSetAttribute(mps::frame_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mps::frame_prep_kernel<<<1, mps::kPrepThreads, smem, (cudaStream_t)stream>>>(a);
  MPS_LAUNCH_CHECK();
  return MPSNERF_OK;
}

extern "C" int mpsnerf_frame_prepare(const float* poses_tp, const float* shapes_tp, const float* R_tp, const float* Th_tp,
                                     const float* poses_sp, const float* shapes_sp, const float* R_sp, const float* Th_sp,
                                     const float* cam_R, const float* cam_T, const float* cam_K, int n_views,
                                     int img_w, int img_h, int feat_w, int feat_h, const float* v_template,
                                     const float* shapedirs, const float* J_regressor, const int32_t* parents,
                                     int n_verts, mpsnerf_frame* out, void* stream) {
  const int rc = mpsnerf_frame_header(R_tp, Th_tp, R_sp, Th_sp, cam_R, cam_T, cam_K, n_views, img_w, img_h, feat_w, feat_h,
                                      out, stream);
  if (rc != MPSNERF_OK) return rc;
  return mpsnerf_frame_transforms(poses_tp, shapes_tp, poses_sp, shapes_sp, v_template, shapedirs, J_regressor, parents,
                                  n_verts, out, stream);
}
