// Shared helpers for libmpsnerf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "mpsnerf.h"

namespace mps {

void set_error(const char* fmt, ...);

#define MPS_REQUIRE(cond)                                                        \
  do {                                                                           \
    if (!(cond)) {                                                               \
      mps::set_error("%s: requirement failed: %s", __func__, #cond);             \
      return MPSNERF_EINVAL;                                                     \
    }                                                                            \
  } while (0)

#define MPS_LAUNCH_CHECK()                                                       \
  do {                                                                           \
    cudaError_t e_ = cudaGetLastError();                                         \
    if (e_ != cudaSuccess) {                                                     \
      mps::set_error("%s: launch failed: %s", __func__, cudaGetErrorString(e_)); \
      return MPSNERF_ECUDA;                                                      \
    }                                                                            \
  } while (0)

#define MPS_CUDA(call)                                                           \
  do {                                                                           \
    cudaError_t e_ = (call);                                                     \
    if (e_ != cudaSuccess) {                                                     \
      mps::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e_)); \
      return MPSNERF_ECUDA;                                                      \
    }                                                                            \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// ---- pinned fp32 arithmetic: every op individually rounded, never contracted to FMA.
// The oracle (oracle/oracle.py) executes the same sequences in numpy, which makes mask and
// vertex indices bit-exact by construction.
__device__ __forceinline__ float pmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float padd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float psub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float pdiv(float a, float b) { return __fdiv_rn(a, b); }

// fl32(0.05**2): the human-region threshold of lib/skinnning_batch.py:360-361
__device__ constexpr float kMaskThresh = 0.0025f;

// z of sample s on a ray (run_nerf_batch.py:411-422), pinned.
__device__ __forceinline__ float sample_z_base(float near, float far, float t) {
  return padd(pmul(near, psub(1.0f, t)), pmul(far, t));
}

__device__ __forceinline__ float sample_z(float near, float far, const float* __restrict__ t_vals,
                                          int s, int S, const float* __restrict__ u_row) {
  float z = sample_z_base(near, far, t_vals[s]);
  if (u_row != nullptr) {
    float z0 = sample_z_base(near, far, t_vals[0]);
    float zl = sample_z_base(near, far, t_vals[S - 1]);
    float lower = (s == 0) ? z0 : pmul(0.5f, padd(z, sample_z_base(near, far, t_vals[s - 1])));
    float upper = (s == S - 1) ? zl : pmul(0.5f, padd(sample_z_base(near, far, t_vals[s + 1]), z));
    z = padd(lower, pmul(psub(upper, lower), u_row[s]));
  }
  return z;
}

}  // namespace mps
