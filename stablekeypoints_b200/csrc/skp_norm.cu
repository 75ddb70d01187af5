// GroupNorm(+SiLU) of the frozen trunk on channels-last activations x[rows = H*W, C], fused with what follows it:
//   * forward never materialises the normalised tensor on its own: gn_apply emits it as fp32 and/or directly as the
//     split-bf16 K-major operand that the tcgen05 GEMM / implicit-GEMM convolution reads through TMA;
//     gn_im2col3x3_split emits the split-bf16 3x3 im2col operand (strided / odd-shaped convolutions);
//   * statistics are (sum, sum of squares) per group: fp32 partials per thread, fp64 across CTAs;
//   * backward (d wrt the input only; gamma/beta are frozen): dz = g * silu'(z), then the GroupNorm input-gradient
//     with two group reductions.
// Thread mapping (HBM-bound, no integer division in the row loops): a thread owns a FIXED set of 4-channel chunks
// (c4 = cx + k*TPR) and walks rows, so its per-channel scale/shift live in registers and every access is a coalesced
// 128-bit load/store.
// Replaces diffusers ResnetBlock2D / Transformer2DModel GroupNorm + SiLU inside the UNet / VAE the
// path runs through (ptp_utils.py:227-229, 299-302).
#include "skp_common.cuh"
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <stdlib.h>

namespace skp {

constexpr int GN_THREADS = 256;
constexpr int GN_MAX_GROUPS = 64;
constexpr int GN_NCH = 3;   // 4-channel chunks per thread -> C <= 4 * 256 * 3 = 3072
constexpr int GN_REPL = 8;  // replicas of the fp64 (sum, sumsq) accumulators: CTA b adds into replica b % 8 (8x less
                            // same-address atomic contention), consumers add the replicas up

__device__ __forceinline__ float silu_f(float z) { return z / (1.f + __expf(-z)); }
__device__ __forceinline__ float silu_grad(float z) {
  float s = 1.f / (1.f + __expf(-z));
  return s * (1.f + z * (1.f - s));
}

struct GnMap {   // how the 256 threads of a CTA tile [rows, C/4]
  int tpr;       // threads along the channel axis (each owns chunks cx, cx+tpr, ...)
  int rl;        // row lanes
};
__host__ __device__ inline GnMap gn_map(int C) {
  GnMap m;
  int c4 = C / 4;
  m.tpr = c4 < GN_THREADS ? c4 : GN_THREADS;
  m.rl = GN_THREADS / m.tpr;
  return m;
}

// per-group (mean, rstd) from the fp64 sums into shared memory
__device__ __forceinline__ void load_group_stats(const double* __restrict__ sums, int G, double count, float eps, float* mean, float* rstd) {
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int r = 0; r < GN_REPL; ++r) {
      s1 += sums[r * 2 * G + 2 * g];
      s2 += sums[r * 2 * G + 2 * g + 1];
    }
    double m = s1 / count;
    double var = s2 / count - m * m;
    if (var < 0.0) var = 0.0;
    mean[g] = (float)m;
    rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
}

__device__ __forceinline__ void store_split4(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t idx, const float v[4]) {
  __nv_bfloat16 h[4], l[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { h[k] = __float2bfloat16_rn(v[k]); l[k] = __float2bfloat16_rn(v[k] - __bfloat162float(h[k])); }
  __nv_bfloat162 ha = __halves2bfloat162(h[0], h[1]), hb = __halves2bfloat162(h[2], h[3]);
  __nv_bfloat162 la = __halves2bfloat162(l[0], l[1]), lb = __halves2bfloat162(l[2], l[3]);
  *reinterpret_cast<uint2*>(hi + idx) = make_uint2(*reinterpret_cast<uint32_t*>(&ha), *reinterpret_cast<uint32_t*>(&hb));
  *reinterpret_cast<uint2*>(lo + idx) = make_uint2(*reinterpret_cast<uint32_t*>(&la), *reinterpret_cast<uint32_t*>(&lb));
}
__device__ __forceinline__ void store_split(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t idx, float v) {
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[idx] = h;
  lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// ------------------------------------------------------------------------------------------------ statistics
// sums[g] += (sum x, sum x^2) over this CTA's rows.  MODE 0: plain statistics of x.  MODE 1: backward reductions
// (sum gamma dz, sum gamma dz xhat) with dz = g * silu'(z).
template <int MODE>
__global__ void __launch_bounds__(GN_THREADS) gn_reduce_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ gr,
                                                               int64_t ldg, int rows, int C, int cg, const double* __restrict__ sums,
                                                               float eps, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, int silu, int rows_per_cta,
                                                               double* __restrict__ out) {
  __shared__ float mean[GN_MAX_GROUPS], rstd[GN_MAX_GROUPS];
  __shared__ float sg[GN_MAX_GROUPS * 2];
  const int G = C / cg;
  for (int i = threadIdx.x; i < 2 * G; i += GN_THREADS) sg[i] = 0.f;
  if (MODE == 1) load_group_stats(sums, G, (double)rows * cg, eps, mean, rstd);
  else __syncthreads();
  const GnMap mp = gn_map(C);
  const int cx = threadIdx.x % mp.tpr, ry = threadIdx.x / mp.tpr;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  const bool same_group4 = (cg % 4) == 0;   // the 4 channels of a chunk share a group
#pragma unroll
  for (int k = 0; k < GN_NCH; ++k) {
    if ((k * mp.tpr) * 4 >= C) break;       // uniform over the CTA
    const int c = (cx + k * mp.tpr) * 4;
    const bool valid = ry < mp.rl && c < C;
    float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
    if (valid) {
      float sc[4], sh[4], gm[4];
      if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int g = (c + j) / cg;
          gm[j] = __ldg(gamma + c + j);
          sc[j] = rstd[g];
          sh[j] = -mean[g] * rstd[g];
        }
      }
      for (int r = r0 + ry; r < r1; r += mp.rl) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * ldx + c));
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
        if (MODE == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) { a1[j] += xs[j]; a2[j] = fmaf(xs[j], xs[j], a2[j]); }
        } else {
          const float4 gv = __ldg(reinterpret_cast<const float4*>(gr + (size_t)r * ldg + c));
          const float gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float xh = fmaf(xs[j], sc[j], sh[j]);
            float dz = gs[j];
            if (silu) dz *= silu_grad(fmaf(gm[j], xh, __ldg(beta + c + j)));
            const float d = dz * gm[j];
            a1[j] += d; a2[j] = fmaf(d, xh, a2[j]);
          }
        }
      }
    }
    // Warp-segmented reduction by group before touching shared memory: lanes of a warp cover a handful of groups, so
    // one leader per (warp, group) issues the shared atomic instead of every thread (256 threads hammering <= 64
    // addresses serialised for ~10 us on the small UNet activations).
    const int lane = threadIdx.x & 31;
    const int npass = same_group4 ? 1 : 4;
    for (int j = 0; j < npass; ++j) {
      float v1 = same_group4 ? (a1[0] + a1[1]) + (a1[2] + a1[3]) : a1[j];
      float v2 = same_group4 ? (a2[0] + a2[1]) + (a2[2] + a2[3]) : a2[j];
      const int key = valid ? (c + j) / cg : -1;
      unsigned remaining = __ballot_sync(0xffffffffu, key >= 0);
      while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const int gk = __shfl_sync(0xffffffffu, key, leader);
        const bool mine = key == gk;
        const float s1 = warp_sum(mine ? v1 : 0.f), s2 = warp_sum(mine ? v2 : 0.f);
        if (lane == leader) {
          atomicAdd(&sg[2 * gk], s1);
          atomicAdd(&sg[2 * gk + 1], s2);
        }
        remaining &= ~__ballot_sync(0xffffffffu, mine);
      }
    }
  }
  __syncthreads();
  double* dst = out + (size_t)(blockIdx.x % GN_REPL) * 2 * G;
  for (int i = threadIdx.x; i < 2 * G; i += GN_THREADS) atomicAdd(dst + i, (double)sg[i]);
}

// ------------------------------------------------------------------------------------------------ apply
// MODE 0: y = act(gamma (x - mean) rstd + beta) -> fp32 y (nullable) and/or split-bf16 hi/lo (nullable, row pitch Kpad).
// MODE 1: dx = rstd (gamma dz - m1 - xhat m2)   (backward; `gr` = d loss / d y, `bs` = the two backward reductions)
template <int MODE>
__global__ void __launch_bounds__(GN_THREADS) gn_rowwise_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ gr,
                                                                int64_t ldg, int rows, int C, int cg, const double* __restrict__ sums,
                                                                float eps, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, int silu, const double* __restrict__ bs,
                                                                int rows_per_cta, float* __restrict__ y, int64_t ldy,
                                                                __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int Kpad) {
  __shared__ float mean[GN_MAX_GROUPS], rstd[GN_MAX_GROUPS], m1[GN_MAX_GROUPS], m2[GN_MAX_GROUPS];
  const int G = C / cg;
  const double count = (double)rows * cg;
  if (MODE == 1)
    for (int g = threadIdx.x; g < G; g += GN_THREADS) {
      double b1 = 0.0, b2 = 0.0;
#pragma unroll
      for (int r = 0; r < GN_REPL; ++r) {
        b1 += bs[r * 2 * G + 2 * g];
        b2 += bs[r * 2 * G + 2 * g + 1];
      }
      m1[g] = (float)(b1 / count);
      m2[g] = (float)(b2 / count);
    }
  load_group_stats(sums, G, count, eps, mean, rstd);
  const GnMap mp = gn_map(C);
  const int cx = threadIdx.x % mp.tpr, ry = threadIdx.x / mp.tpr;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  if (ry >= mp.rl) return;
#pragma unroll
  for (int k = 0; k < GN_NCH; ++k) {
    const int c = (cx + k * mp.tpr) * 4;
    if (c >= C) break;
    float sc[4], sh[4], gm[4], bt[4], a1[4], a2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int g = (c + j) / cg;
      gm[j] = __ldg(gamma + c + j);
      bt[j] = __ldg(beta + c + j);
      if (MODE == 0) {             // y = x * sc + sh  with gamma/beta folded in
        sc[j] = rstd[g] * gm[j];
        sh[j] = fmaf(-mean[g], sc[j], bt[j]);
      } else {                     // xhat = x * sc + sh
        sc[j] = rstd[g];
        sh[j] = -mean[g] * rstd[g];
        a1[j] = m1[g]; a2[j] = m2[g];
      }
    }
    for (int r = r0 + ry; r < r1; r += mp.rl) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * ldx + c));
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
      float o[4];
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float z = fmaf(xs[j], sc[j], sh[j]);
          o[j] = silu ? silu_f(z) : z;
        }
        if (y) *reinterpret_cast<float4*>(y + (size_t)r * ldy + c) = make_float4(o[0], o[1], o[2], o[3]);
        if (hi) store_split4(hi, lo, (size_t)r * Kpad + c, o);
      } else {
        const float4 gv = __ldg(reinterpret_cast<const float4*>(gr + (size_t)r * ldg + c));
        const float gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float xh = fmaf(xs[j], sc[j], sh[j]);
          float dz = gs[j];
          if (silu) dz *= silu_grad(fmaf(gm[j], xh, bt[j]));
          o[j] = sc[j] * (dz * gm[j] - a1[j] - xh * a2[j]);
        }
        *reinterpret_cast<float4*>(y + (size_t)r * ldy + c) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  if (MODE == 0 && hi && Kpad > C) {   // zero the K padding of the operand
    for (int r = r0 + threadIdx.x / 32; r < r1; r += GN_THREADS / 32)
      for (int c = C + (threadIdx.x & 31); c < Kpad; c += 32) store_split(hi, lo, (size_t)r * Kpad + c, 0.f);
  }
}

// ------------------------------------------------------------------------------------------------ scalar fallbacks (C % 4 != 0)
__global__ void __launch_bounds__(GN_THREADS) gn_stats_scalar_kernel(const float* __restrict__ x, int64_t ldx, int rows, int C, int cg,
                                                                     int rows_per_cta, double* __restrict__ sums) {
  __shared__ float sg[GN_MAX_GROUPS * 2];
  const int G = C / cg;
  for (int i = threadIdx.x; i < 2 * G; i += GN_THREADS) sg[i] = 0.f;
  __syncthreads();
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  for (int c = threadIdx.x; c < C; c += GN_THREADS) {
    float s = 0.f, ss = 0.f;
    for (int r = r0; r < r1; ++r) {
      float v = __ldg(x + (size_t)r * ldx + c);
      s += v; ss = fmaf(v, v, ss);
    }
    atomicAdd(&sg[2 * (c / cg)], s);
    atomicAdd(&sg[2 * (c / cg) + 1], ss);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * G; i += GN_THREADS) atomicAdd(sums + i, (double)sg[i]);
}

__global__ void __launch_bounds__(GN_THREADS) gn_apply_scalar_kernel(const float* __restrict__ x, int64_t ldx, int rows, int C, int cg,
                                                                     const double* __restrict__ sums, float eps,
                                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                     int silu, float* __restrict__ y, int64_t ldy,
                                                                     __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int Kpad) {
  __shared__ float mean[GN_MAX_GROUPS], rstd[GN_MAX_GROUPS];
  load_group_stats(sums, C / cg, (double)rows * cg, eps, mean, rstd);
  const int width = hi ? Kpad : C;
  const size_t total = (size_t)rows * width;
  for (size_t i = blockIdx.x * (size_t)GN_THREADS + threadIdx.x; i < total; i += (size_t)gridDim.x * GN_THREADS) {
    const int r = (int)(i / width), c = (int)(i - (size_t)r * width);
    float v = 0.f;
    if (c < C) {
      const int g = c / cg;
      v = (__ldg(x + (size_t)r * ldx + c) - mean[g]) * rstd[g] * __ldg(gamma + c) + __ldg(beta + c);
      if (silu) v = silu_f(v);
      if (y) y[(size_t)r * ldy + c] = v;
    }
    if (hi) store_split(hi, lo, i, v);
  }
}

__global__ void __launch_bounds__(GN_THREADS) gn_bwd_reduce_scalar_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ gr,
                                                                          int64_t ldg, int rows, int C, int cg,
                                                                          const double* __restrict__ sums, float eps,
                                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                          int silu, int rows_per_cta, double* __restrict__ bs) {
  __shared__ float mean[GN_MAX_GROUPS], rstd[GN_MAX_GROUPS];
  __shared__ float sg[GN_MAX_GROUPS * 2];
  const int G = C / cg;
  for (int i = threadIdx.x; i < 2 * G; i += GN_THREADS) sg[i] = 0.f;
  load_group_stats(sums, G, (double)rows * cg, eps, mean, rstd);
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  for (int c = threadIdx.x; c < C; c += GN_THREADS) {
    const int g = c / cg;
    const float gm = __ldg(gamma + c), bt = __ldg(beta + c), mu = mean[g], rs = rstd[g];
    float s1 = 0.f, s2 = 0.f;
    for (int r = r0; r < r1; ++r) {
      float xh = (__ldg(x + (size_t)r * ldx + c) - mu) * rs;
      float dz = __ldg(gr + (size_t)r * ldg + c);
      if (silu) dz *= silu_grad(gm * xh + bt);
      float d = dz * gm;
      s1 += d; s2 = fmaf(d, xh, s2);
    }
    atomicAdd(&sg[2 * g], s1);
    atomicAdd(&sg[2 * g + 1], s2);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * G; i += GN_THREADS) atomicAdd(bs + i, (double)sg[i]);
}

__global__ void __launch_bounds__(GN_THREADS) gn_bwd_apply_scalar_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ gr,
                                                                         int64_t ldg, int rows, int C, int cg,
                                                                         const double* __restrict__ sums, float eps,
                                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                         int silu, const double* __restrict__ bs, float* __restrict__ dx,
                                                                         int64_t ldd) {
  __shared__ float mean[GN_MAX_GROUPS], rstd[GN_MAX_GROUPS], m1[GN_MAX_GROUPS], m2[GN_MAX_GROUPS];
  const int G = C / cg;
  const double count = (double)rows * cg;
  for (int g = threadIdx.x; g < G; g += GN_THREADS) {
    double b1 = 0.0, b2 = 0.0;
    for (int r = 0; r < GN_REPL; ++r) {
      b1 += bs[r * 2 * G + 2 * g];
      b2 += bs[r * 2 * G + 2 * g + 1];
    }
    m1[g] = (float)(b1 / count);
    m2[g] = (float)(b2 / count);
  }
  load_group_stats(sums, G, count, eps, mean, rstd);
  const size_t total = (size_t)rows * C;
  for (size_t i = blockIdx.x * (size_t)GN_THREADS + threadIdx.x; i < total; i += (size_t)gridDim.x * GN_THREADS) {
    const int r = (int)(i / C), c = (int)(i - (size_t)r * C);
    const int g = c / cg;
    const float gm = __ldg(gamma + c), rs = rstd[g];
    float xh = (__ldg(x + (size_t)r * ldx + c) - mean[g]) * rs;
    float dz = __ldg(gr + (size_t)r * ldg + c);
    if (silu) dz *= silu_grad(gm * xh + __ldg(beta + c));
    dx[(size_t)r * ldd + c] = rs * (dz * gm - m1[g] - xh * m2[g]);
  }
}

// fused GroupNorm(+SiLU) + 3x3 im2col to split-bf16 (column = tap*C + c); one warp per (output pixel, tap)
__global__ void __launch_bounds__(GN_THREADS) gn_im2col3x3_split_kernel(const float* __restrict__ x, int64_t ldx, int H, int W, int C,
                                                                        int cg, const double* __restrict__ sums, float eps,
                                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                        int silu, int Ho, int Wo, int stride, int pad, int Kpad,
                                                                        __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  __shared__ float mean[GN_MAX_GROUPS], rstd[GN_MAX_GROUPS];
  load_group_stats(sums, C / cg, (double)H * W * cg, eps, mean, rstd);
  const int lane = threadIdx.x & 31;
  const long warp = (blockIdx.x * (long)GN_THREADS + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * GN_THREADS) >> 5;
  const long items = (long)Ho * Wo * 9;
  const bool vec = (C & 3) == 0 && (ldx & 3) == 0 && ((((uintptr_t)x) | ((uintptr_t)gamma) | ((uintptr_t)beta)) & 15) == 0;
  for (long it = warp; it < items; it += nwarps) {
    const int tap = (int)(it % 9);
    const long row = it / 9;
    const int oy = (int)(row / Wo), ox = (int)(row - (long)oy * Wo);
    const int iy = oy * stride + tap / 3 - pad, ix = ox * stride + tap % 3 - pad;
    const bool inside = iy >= 0 && iy < H && ix >= 0 && ix < W;
    const float* src = x + ((size_t)(inside ? iy : 0) * W + (inside ? ix : 0)) * ldx;
    const size_t dst = (size_t)row * Kpad + (size_t)tap * C;
    // zero padding applies AFTER the normalisation/activation (the convolution pads its input)
    if (vec) {
      for (int c = lane * 4; c < C; c += 128) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (inside) {
          const float4 xv = *reinterpret_cast<const float4*>(src + c);
          const float4 gm = *reinterpret_cast<const float4*>(gamma + c), bt = *reinterpret_cast<const float4*>(beta + c);
          const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gm.x, gm.y, gm.z, gm.w}, bs[4] = {bt.x, bt.y, bt.z, bt.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int g = (c + k) / cg;
            float z = (xs[k] - mean[g]) * rstd[g] * gs[k] + bs[k];
            v[k] = silu ? silu_f(z) : z;
          }
        }
        store_split4(hi, lo, dst + c, v);
      }
    } else {
      for (int c = lane; c < C; c += 32) {
        float v = 0.f;
        if (inside) {
          const int g = c / cg;
          v = (__ldg(src + c) - mean[g]) * rstd[g] * __ldg(gamma + c) + __ldg(beta + c);
          if (silu) v = silu_f(v);
        }
        store_split(hi, lo, dst + c, v);
      }
    }
    if (tap == 8)
      for (int c = 9 * C + lane; c < Kpad; c += 32) store_split(hi, lo, (size_t)row * Kpad + c, 0.f);
  }
}

// ------------------------------------------------------------------------------------------------ one cluster per group
// Small and medium activations (the UNet at one image per rank: 64..4096 rows): statistics AND normalisation in ONE launch,
// one thread-block CLUSTER of 1..8 CTAs per group -- no atomics, no memset, no second kernel, bit-reproducible.  Each CTA
// sweeps its share of the rows of the [rows x cg] slab twice (the second sweep hits L1/L2); the per-CTA partial sums meet
// through distributed shared memory (every CTA reads the partials of all ranks in rank order, so all of them hold the same
// totals).  Thread = (channel pair, row lane).  The (sum, sumsq) pair is still published (replica 0; the others zeroed) for
// the backward.
namespace cgx = cooperative_groups;

// sum over the cluster of one (a, b) pair of block totals, in rank order; the two cluster syncs bracket the remote reads
__device__ __forceinline__ void cluster_sum2(double& a, double& b, double* part) {
  cgx::cluster_group cl = cgx::this_cluster();
  const unsigned n = cl.num_blocks();
  if (n == 1) return;
  if (threadIdx.x == 0) { part[0] = a; part[1] = b; }
  cl.sync();
  double sa = 0.0, sb = 0.0;
  for (unsigned r = 0; r < n; ++r) {
    const double* rp = cl.map_shared_rank(part, r);
    sa += rp[0];
    sb += rp[1];
  }
  cl.sync();          // nobody leaves (or overwrites part) while a peer still reads it
  a = sa;
  b = sb;
}
__device__ __forceinline__ double block_sum_d(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double r = 0.0;
  for (int i = 0; i < nw; ++i) r += red[i];
  return r;
}

__global__ void __launch_bounds__(GN_THREADS) gn_group_fwd_kernel(const float* __restrict__ x, int64_t ldx, int rows, int C, int cg,
                                                                  float eps, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, int silu, float* __restrict__ y,
                                                                  int64_t ldy, __nv_bfloat16* __restrict__ hi,
                                                                  __nv_bfloat16* __restrict__ lo, int Kpad, double* __restrict__ sums) {
  __shared__ double red[GN_THREADS / 32];
  __shared__ double part[2];
  const int g = blockIdx.x, G = gridDim.x, c0 = g * cg;
  const int rank = blockIdx.y, CL = gridDim.y;                       // cluster = the CL CTAs of this group
  const int rbeg = (int)((long)rows * rank / CL), rend = (int)((long)rows * (rank + 1) / CL);
  const int PP = cg >> 1, RL = GN_THREADS / PP;
  const int p = threadIdx.x % PP, ry = threadIdx.x / PP;
  const bool active = ry < RL;
  const float* xc = x + c0 + 2 * p;
  float s = 0.f, ss = 0.f;
  if (active)
    for (int r = rbeg + ry; r < rend; r += 4 * RL) {   // 4 independent loads in flight per thread
      float2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = (r + u * RL < rend) ? __ldg(reinterpret_cast<const float2*>(xc + (size_t)(r + u * RL) * ldx)) : make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s += v[u].x + v[u].y;
        ss = fmaf(v[u].x, v[u].x, fmaf(v[u].y, v[u].y, ss));
      }
    }
  double S = block_sum_d((double)s, red), SS = block_sum_d((double)ss, red);
  cluster_sum2(S, SS, part);
  const double count = (double)rows * cg;
  const double md = S / count;
  double var = SS / count - md * md;
  if (var < 0.0) var = 0.0;
  const float mean = (float)md, rstd = (float)(1.0 / sqrt(var + (double)eps));
  if (rank == 0 && threadIdx.x < 2 * GN_REPL) {   // replica 0 carries the sums, the rest must read as zero
    const int r = threadIdx.x >> 1, w = threadIdx.x & 1;
    sums[(size_t)r * 2 * G + 2 * g + w] = r == 0 ? (w == 0 ? S : SS) : 0.0;
  }
  if (active) {
    const float g0 = __ldg(gamma + c0 + 2 * p), g1 = __ldg(gamma + c0 + 2 * p + 1);
    const float sc0 = rstd * g0, sc1 = rstd * g1;
    const float sh0 = fmaf(-mean, sc0, __ldg(beta + c0 + 2 * p)), sh1 = fmaf(-mean, sc1, __ldg(beta + c0 + 2 * p + 1));
    for (int r0 = rbeg + ry; r0 < rend; r0 += 4 * RL) {
      float2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = (r0 + u * RL < rend) ? __ldg(reinterpret_cast<const float2*>(xc + (size_t)(r0 + u * RL) * ldx)) : make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u * RL;
        if (r >= rend) break;
        float z0 = fmaf(v[u].x, sc0, sh0), z1 = fmaf(v[u].y, sc1, sh1);
        if (silu) { z0 = silu_f(z0); z1 = silu_f(z1); }
        if (y) *reinterpret_cast<float2*>(y + (size_t)r * ldy + c0 + 2 * p) = make_float2(z0, z1);
        if (hi) {
          const __nv_bfloat162 h = __floats2bfloat162_rn(z0, z1);
          const float2 f = __bfloat1622float2(h);
          const size_t o = (size_t)r * Kpad + c0 + 2 * p;
          *reinterpret_cast<__nv_bfloat162*>(hi + o) = h;
          *reinterpret_cast<__nv_bfloat162*>(lo + o) = __floats2bfloat162_rn(z0 - f.x, z1 - f.y);
        }
      }
    }
  }
  if (hi && Kpad > C && g == G - 1) {   // zero the K padding of the operand (this CTA's rows)
    const int padp = (Kpad - C) >> 1;
    for (int i = threadIdx.x; i < (rend - rbeg) * padp; i += GN_THREADS) {
      const int r = rbeg + i / padp, c = C + 2 * (i % padp);
      *reinterpret_cast<uint32_t*>(hi + (size_t)r * Kpad + c) = 0u;
      *reinterpret_cast<uint32_t*>(lo + (size_t)r * Kpad + c) = 0u;
    }
  }
}

__global__ void __launch_bounds__(GN_THREADS) gn_group_bwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ gr,
                                                                  int64_t ldg, int rows, int C, int cg,
                                                                  const double* __restrict__ sums, float eps,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  int silu, float* __restrict__ dx, int64_t ldd) {
  __shared__ double red[GN_THREADS / 32];
  __shared__ double part[2];
  __shared__ float st[2];
  const int g = blockIdx.x, G = gridDim.x, c0 = g * cg;
  const int rank = blockIdx.y, CL = gridDim.y;
  const int rbeg = (int)((long)rows * rank / CL), rend = (int)((long)rows * (rank + 1) / CL);
  const double count = (double)rows * cg;
  if (threadIdx.x == 0) {
    double s1 = 0.0, s2 = 0.0;
    for (int r = 0; r < GN_REPL; ++r) {
      s1 += sums[(size_t)r * 2 * G + 2 * g];
      s2 += sums[(size_t)r * 2 * G + 2 * g + 1];
    }
    const double m = s1 / count;
    double var = s2 / count - m * m;
    if (var < 0.0) var = 0.0;
    st[0] = (float)m;
    st[1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const float mean = st[0], rstd = st[1];
  const int PP = cg >> 1, RL = GN_THREADS / PP;
  const int p = threadIdx.x % PP, ry = threadIdx.x / PP;
  const bool active = ry < RL;
  const int c = c0 + 2 * p;
  float g0 = 0.f, g1 = 0.f, b0 = 0.f, b1 = 0.f;
  if (active) { g0 = __ldg(gamma + c); g1 = __ldg(gamma + c + 1); b0 = __ldg(beta + c); b1 = __ldg(beta + c + 1); }
  const float sc = rstd, sh = -mean * rstd;
  float a1 = 0.f, a2 = 0.f;
  if (active)
    for (int r0 = rbeg + ry; r0 < rend; r0 += 4 * RL) {
      float2 v[4], gv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool ok = r0 + u * RL < rend;
        v[u] = ok ? __ldg(reinterpret_cast<const float2*>(x + (size_t)(r0 + u * RL) * ldx + c)) : make_float2(0.f, 0.f);
        gv[u] = ok ? __ldg(reinterpret_cast<const float2*>(gr + (size_t)(r0 + u * RL) * ldg + c)) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float xh0 = fmaf(v[u].x, sc, sh), xh1 = fmaf(v[u].y, sc, sh);
        float dz0 = gv[u].x, dz1 = gv[u].y;     // rows past the end carry g = 0 and contribute nothing
        if (silu) { dz0 *= silu_grad(fmaf(g0, xh0, b0)); dz1 *= silu_grad(fmaf(g1, xh1, b1)); }
        const float d0 = dz0 * g0, d1 = dz1 * g1;
        a1 += d0 + d1;
        a2 = fmaf(d0, xh0, fmaf(d1, xh1, a2));
      }
    }
  double A1 = block_sum_d((double)a1, red), A2 = block_sum_d((double)a2, red);
  cluster_sum2(A1, A2, part);
  const float m1 = (float)(A1 / count), m2 = (float)(A2 / count);
  if (active)
    for (int r0 = rbeg + ry; r0 < rend; r0 += 4 * RL) {
      float2 v[4], gv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool ok = r0 + u * RL < rend;
        v[u] = ok ? __ldg(reinterpret_cast<const float2*>(x + (size_t)(r0 + u * RL) * ldx + c)) : make_float2(0.f, 0.f);
        gv[u] = ok ? __ldg(reinterpret_cast<const float2*>(gr + (size_t)(r0 + u * RL) * ldg + c)) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u * RL;
        if (r >= rend) break;
        const float xh0 = fmaf(v[u].x, sc, sh), xh1 = fmaf(v[u].y, sc, sh);
        float dz0 = gv[u].x, dz1 = gv[u].y;
        if (silu) { dz0 *= silu_grad(fmaf(g0, xh0, b0)); dz1 *= silu_grad(fmaf(g1, xh1, b1)); }
        *reinterpret_cast<float2*>(dx + (size_t)r * ldd + c) =
            make_float2(sc * (dz0 * g0 - m1 - xh0 * m2), sc * (dz1 * g1 - m1 - xh1 * m2));
      }
    }
}

// the one-cluster-per-group kernels take even group widths up to 512 channels and slabs of at most 20K elements per CTA
// of the cluster (1, 2, 4 or 8 CTAs: 32 .. 256 CTAs in flight); beyond that the two-kernel path streams better
constexpr int GN_SLAB_PER_CTA = 20480;
static int g_gn_cluster_max = getenv("SKP_GN_CLUSTER") ? atoi(getenv("SKP_GN_CLUSTER")) : 8;
static int g_gn_cta_elems = getenv("SKP_GN_CTA_ELEMS") ? atoi(getenv("SKP_GN_CTA_ELEMS")) : 4096;   // elements per CTA before the cluster grows
static inline int gn_cluster_size(int rows, int cg) {
  const size_t slab = (size_t)rows * cg;
  int cl = 1;
  while (cl < g_gn_cluster_max && slab > (size_t)g_gn_cta_elems * cl) cl *= 2;     // the sweep is latency-bound
  while (cl > 1 && rows / cl < 1) cl /= 2;
  return cl;
}
static inline bool gn_group_ok(int rows, int C, int groups, const float* x, int64_t ldx) {
  const int cg = C / groups;
  return (cg % 2 == 0) && cg <= 2 * GN_THREADS && (size_t)rows * cg <= (size_t)GN_SLAB_PER_CTA * g_gn_cluster_max && (ldx % 2 == 0) &&
         ((((uintptr_t)x) & 7) == 0);
}
template <typename... KArgs, typename... Args>
static cudaError_t gn_launch_cluster(void (*kernel)(KArgs...), int groups, int cl, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups, cl);
  cfg.blockDim = dim3(GN_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1;
  at[0].val.clusterDim.y = cl;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Rows per CTA.  Large activations (VAE): ~6 CTAs per SM so enough loads are in flight.  Small ones (the UNet at one image
// per rank is 0.1-5 MB per tensor): few rows per thread -- the kernels are latency-bound, and for the reductions the grid
// is capped so that the fp64 global atomics of the CTAs do not pile up on the same 64 addresses.
static inline int gn_rows_per_cta(int rows, int C, bool reduce = false) {
  GnMap m = gn_map(C >= 4 ? C : 4);
  const bool small = (size_t)rows * C < ((size_t)4 << 20);
  int per = (rows + 148 * 6 - 1) / (148 * 6);
  if (small && reduce) per = (rows + 295) / 296;
  int min_rows = m.rl * (small ? (reduce ? 2 : 1) : 4);
  return per < min_rows ? min_rows : per;
}
static inline int gn_grid(size_t n) {
  size_t b = (n + GN_THREADS - 1) / GN_THREADS;
  if (b > 148 * 16) b = 148 * 16;
  return b < 1 ? 1 : (int)b;
}
static inline bool gn_vec_ok(const float* x, int64_t ldx, int C, const float* g = nullptr, int64_t ldg = 0) {
  bool ok = (C % 4 == 0) && (C <= 4 * GN_THREADS * GN_NCH) && (ldx % 4 == 0) && ((((uintptr_t)x) & 15) == 0);
  if (g) ok = ok && (ldg % 4 == 0) && ((((uintptr_t)g) & 15) == 0);
  return ok;
}

}  // namespace skp

using namespace skp;

#define SKP_GN_CHECK(name)                                                                                   \
  SKP_REQUIRE(x && rows > 0 && C > 0 && groups > 0 && C % groups == 0, name ": bad arguments");              \
  SKP_REQUIRE(groups <= GN_MAX_GROUPS, name ": at most %d groups supported", GN_MAX_GROUPS);

extern "C" int skp_gn_stats(const float* x, int64_t ldx, int rows, int C, int groups, double* sums, void* stream) {
  SKP_GN_CHECK("gn_stats");
  SKP_REQUIRE(sums, "gn_stats: null sums");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(sums, 0, sizeof(double) * 2 * groups * GN_REPL, st);
  int per = gn_rows_per_cta(rows, C, true);
  int grid = (rows + per - 1) / per;
  if (gn_vec_ok(x, ldx, C))
    gn_reduce_kernel<0><<<grid, GN_THREADS, 0, st>>>(x, ldx, nullptr, 0, rows, C, C / groups, nullptr, 0.f, nullptr, nullptr, 0, per, sums);
  else
    gn_stats_scalar_kernel<<<grid, GN_THREADS, 0, st>>>(x, ldx, rows, C, C / groups, per, sums);
  SKP_CHECK_LAUNCH("gn_stats");
  return SKP_OK;
}

extern "C" int skp_gn_apply(const float* x, int64_t ldx, int rows, int C, int groups, const double* sums, float eps,
                            const float* gamma, const float* beta, int silu, float* y, int64_t ldy, void* hi, void* lo, int Kpad,
                            void* stream) {
  SKP_GN_CHECK("gn_apply");
  SKP_REQUIRE(sums && gamma && beta && (y || (hi && lo)), "gn_apply: null pointer");
  SKP_REQUIRE(!hi || (Kpad >= C && Kpad % 64 == 0), "gn_apply: Kpad=%d must be a multiple of 64 >= C", Kpad);
  cudaStream_t st = (cudaStream_t)stream;
  bool vec = gn_vec_ok(x, ldx, C) && (!y || (ldy % 4 == 0 && ((((uintptr_t)y) & 15) == 0))) &&
             (!hi || (((((uintptr_t)hi) | ((uintptr_t)lo)) & 7) == 0));
  if (vec) {
    int per = gn_rows_per_cta(rows, C);
    gn_rowwise_kernel<0><<<(rows + per - 1) / per, GN_THREADS, 0, st>>>(x, ldx, nullptr, 0, rows, C, C / groups, sums, eps, gamma, beta,
                                                                        silu, nullptr, per, y, ldy, (__nv_bfloat16*)hi,
                                                                        (__nv_bfloat16*)lo, Kpad);
  } else {
    size_t total = (size_t)rows * (hi ? Kpad : C);
    gn_apply_scalar_kernel<<<gn_grid(total), GN_THREADS, 0, st>>>(x, ldx, rows, C, C / groups, sums, eps, gamma, beta, silu, y, ldy,
                                                                  (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, Kpad);
  }
  SKP_CHECK_LAUNCH("gn_apply");
  return SKP_OK;
}

extern "C" int skp_gn_fwd(const float* x, int64_t ldx, int rows, int C, int groups, float eps, const float* gamma,
                          const float* beta, int silu, float* y, int64_t ldy, void* hi, void* lo, int Kpad, double* sums,
                          void* stream) {
  SKP_GN_CHECK("gn_fwd");
  SKP_REQUIRE(sums && gamma && beta && (y || (hi && lo)), "gn_fwd: null pointer");
  SKP_REQUIRE(!hi || (Kpad >= C && Kpad % 64 == 0), "gn_fwd: Kpad=%d must be a multiple of 64 >= C", Kpad);
  static const bool group_off = getenv("SKP_GN_GROUP") != nullptr && atoi(getenv("SKP_GN_GROUP")) == 0;
  if (!group_off && gn_group_ok(rows, C, groups, x, ldx) && (!y || (ldy % 2 == 0 && ((((uintptr_t)y) & 7) == 0))) &&
      (!hi || (((((uintptr_t)hi) | ((uintptr_t)lo)) & 3) == 0))) {
    cudaError_t le = gn_launch_cluster(gn_group_fwd_kernel, groups, gn_cluster_size(rows, C / groups), (cudaStream_t)stream, x, ldx, rows, C,
                                       C / groups, eps, gamma, beta, silu, y, ldy, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, Kpad, sums);
    if (le != cudaSuccess) { set_error("gn_group_fwd: cluster launch: %s", cudaGetErrorString(le)); return SKP_ERR_LAUNCH; }
    SKP_CHECK_LAUNCH("gn_group_fwd");
    return SKP_OK;
  }
  int rc = skp_gn_stats(x, ldx, rows, C, groups, sums, stream);
  if (rc != SKP_OK) return rc;
  return skp_gn_apply(x, ldx, rows, C, groups, sums, eps, gamma, beta, silu, y, ldy, hi, lo, Kpad, stream);
}

extern "C" int skp_gn_im2col3x3_split(const float* x, int64_t ldx, int H, int W, int C, int groups, const double* sums, float eps,
                                      const float* gamma, const float* beta, int silu, int Ho, int Wo, int stride, int pad,
                                      int Kpad, void* hi, void* lo, void* stream) {
  const int rows = H * W;
  SKP_GN_CHECK("gn_im2col3x3_split");
  SKP_REQUIRE(sums && gamma && beta && hi && lo && Ho > 0 && Wo > 0 && stride > 0, "gn_im2col3x3_split: bad arguments");
  SKP_REQUIRE(Kpad >= 9 * C && Kpad % 64 == 0, "gn_im2col3x3_split: Kpad=%d must be a multiple of 64 >= 9*C", Kpad);
  long items = (long)Ho * Wo * 9;
  long blocks = (items * 32 + GN_THREADS - 1) / GN_THREADS;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gn_im2col3x3_split_kernel<<<(int)blocks, GN_THREADS, 0, (cudaStream_t)stream>>>(x, ldx, H, W, C, C / groups, sums, eps, gamma, beta,
                                                                                 silu, Ho, Wo, stride, pad, Kpad, (__nv_bfloat16*)hi,
                                                                                 (__nv_bfloat16*)lo);
  SKP_CHECK_LAUNCH("gn_im2col3x3_split");
  return SKP_OK;
}

extern "C" int skp_gn_bwd(const float* x, int64_t ldx, const float* g, int64_t ldg, int rows, int C, int groups, const double* sums,
                          float eps, const float* gamma, const float* beta, int silu, double* bsums, float* dx, int64_t ldd,
                          void* stream) {
  SKP_GN_CHECK("gn_bwd");
  SKP_REQUIRE(g && sums && gamma && beta && bsums && dx, "gn_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  static const bool group_off = getenv("SKP_GN_GROUP") != nullptr && atoi(getenv("SKP_GN_GROUP")) == 0;
  if (!group_off && gn_group_ok(rows, C, groups, x, ldx) && (ldg % 2 == 0) && (ldd % 2 == 0) &&
      (((((uintptr_t)g) | ((uintptr_t)dx)) & 7) == 0)) {
    cudaError_t le = gn_launch_cluster(gn_group_bwd_kernel, groups, gn_cluster_size(rows, C / groups), st, x, ldx, g, ldg, rows, C, C / groups,
                                       sums, eps, gamma, beta, silu, dx, ldd);
    if (le != cudaSuccess) { set_error("gn_group_bwd: cluster launch: %s", cudaGetErrorString(le)); return SKP_ERR_LAUNCH; }
    SKP_CHECK_LAUNCH("gn_group_bwd");
    return SKP_OK;
  }
  cudaMemsetAsync(bsums, 0, sizeof(double) * 2 * groups * GN_REPL, st);
  int per = gn_rows_per_cta(rows, C, true);
  int grid = (rows + per - 1) / per;
  const int per_rw = gn_rows_per_cta(rows, C, false), grid_rw = (rows + per_rw - 1) / per_rw;
  const bool vec = gn_vec_ok(x, ldx, C, g, ldg) && (ldd % 4 == 0) && ((((uintptr_t)dx) & 15) == 0);
  if (vec)
    gn_reduce_kernel<1><<<grid, GN_THREADS, 0, st>>>(x, ldx, g, ldg, rows, C, C / groups, sums, eps, gamma, beta, silu, per, bsums);
  else
    gn_bwd_reduce_scalar_kernel<<<grid, GN_THREADS, 0, st>>>(x, ldx, g, ldg, rows, C, C / groups, sums, eps, gamma, beta, silu, per, bsums);
  SKP_CHECK_LAUNCH("gn_bwd_reduce");
  if (vec)
    gn_rowwise_kernel<1><<<grid_rw, GN_THREADS, 0, st>>>(x, ldx, g, ldg, rows, C, C / groups, sums, eps, gamma, beta, silu, bsums, per_rw, dx,
                                                         ldd, nullptr, nullptr, 0);
  else
    gn_bwd_apply_scalar_kernel<<<gn_grid((size_t)rows * C), GN_THREADS, 0, st>>>(x, ldx, g, ldg, rows, C, C / groups, sums, eps, gamma, beta,
                                                                                silu, bsums, dx, ldd);
  SKP_CHECK_LAUNCH("gn_bwd_apply");
  return SKP_OK;
}
