// Shared device/host helpers for libskp_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/skp_b200.h"

namespace skp {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define SKP_REQUIRE(cond, ...)                  \
  do {                                          \
    if (!(cond)) {                              \
      skp::set_error(__VA_ARGS__);              \
      return SKP_ERR_INVALID;                   \
    }                                           \
  } while (0)

#define SKP_CHECK_LAUNCH(name)                                              \
  do {                                                                      \
    cudaError_t e__ = cudaGetLastError();                                   \
    if (e__ != cudaSuccess) {                                               \
      skp::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return SKP_ERR_LAUNCH;                                                \
    }                                                                       \
    skp::count_launch();                                                    \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum; `red` is >= 32 floats of shared memory.  All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
    r = warp_sum(r);
    if (lane == 0) red[0] = r;
  }
  __syncthreads();
  r = red[0];
  return r;
}

// Programmatic dependent launch (PDL): a kernel launched with skp::launch_pdl may be scheduled while its predecessor on the
// stream is still draining; it must call pdl_wait() before its first global-memory access (the wait returns once the
// predecessor has completed and flushed), and calls pdl_launch_dependents() at its top so that ITS successor may be
// scheduled early too.  Everything before pdl_wait() (barrier init, TMEM allocation, index set-up) overlaps the
// predecessor's tail.  With several thousand 5-15 us launches per step this launch latency is a visible share.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();   // skp_api.cu: SKP_PDL=1 (opt-in; no measurable gain inside the step graph)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// PyTorch upsample_bicubic2d coefficients (A = -0.75), align_corners=False.
__device__ __forceinline__ void cubic_coeffs(float t, float w[4]) {
  const float A = -0.75f;
  float x;
  x = t + 1.f;  w[0] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
  x = t;        w[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 1.f - t;  w[2] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 2.f - t;  w[3] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
}

// Source coordinate of destination index `dst` for a resize in_size -> out_size, align_corners=False.
// Bicubic keeps negative coordinates (taps are clamped); bilinear clamps the coordinate at 0.
__device__ __forceinline__ float src_coord(int dst, float scale /* in/out */, bool cubic) {
  float s = scale * (dst + 0.5f) - 0.5f;
  return (!cubic && s < 0.f) ? 0.f : s;
}

}  // namespace skp
