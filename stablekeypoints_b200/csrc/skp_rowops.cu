// Row-wise pieces of the frozen BasicTransformerBlock (diffusers) fused with the operand split of
// the projection that follows them, so the normalised / gated activation is never stored as fp32:
//   LayerNorm  -> split-bf16 K-major operand of to_q / qkv / ff.net.0.proj      (norm1, norm2, norm3)
//   GEGLU      a * gelu(gate) (exact erf GELU) -> split-bf16 operand of ff.net.2
// and their input gradients (gamma / beta are frozen: only dx exists, optimize_token.py:71-76).
// One warp owns one row; every access is a coalesced 128-bit load.  HBM-bound: 4 B read + 4 B written per element.
#include "skp_common.cuh"
#include <cuda_bf16.h>

namespace skp {

constexpr int ROW_WARPS = 4;   // rows per CTA (one warp each)

__device__ __forceinline__ void split_store4(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t idx, float a, float b, float c, float d) {
  __nv_bfloat162 h0 = __floats2bfloat162_rn(a, b), h1 = __floats2bfloat162_rn(c, d);
  float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
  __nv_bfloat162 l0 = __floats2bfloat162_rn(a - f0.x, b - f0.y), l1 = __floats2bfloat162_rn(c - f1.x, d - f1.y);
  *reinterpret_cast<uint2*>(hi + idx) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
  *reinterpret_cast<uint2*>(lo + idx) = make_uint2(*reinterpret_cast<uint32_t*>(&l0), *reinterpret_cast<uint32_t*>(&l1));
}

constexpr int LN_MAXCH = 10;   // float4 chunks a lane keeps in registers: rows of up to 1280 channels are read from L2 ONCE

// x[rows, C] (C % 4 == 0) -> hi/lo[rows, Kpad] = split(LayerNorm(x) * gamma + beta); stats[row] = (mean, rstd)
// CACHED: C <= 1280, the row lives in registers between the three sweeps (mean, variance, normalise).
template <bool CACHED>
__global__ void __launch_bounds__(ROW_WARPS * 32) ln_split_kernel(const float* __restrict__ x, int64_t ldx, int rows, int C,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  float eps, __nv_bfloat16* __restrict__ hi,
                                                                  __nv_bfloat16* __restrict__ lo, int Kpad, float2* __restrict__ stats) {
  const int lane = threadIdx.x & 31, row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * ldx);
  const int C4 = C >> 2;
  float4 v[LN_MAXCH];
  float s = 0.f;
  if (CACHED) {
#pragma unroll
    for (int k = 0; k < LN_MAXCH; ++k) {
      const int i = lane + 32 * k;
      v[k] = i < C4 ? __ldg(xr + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
  } else {
    for (int i = lane; i < C4; i += 32) {
      float4 t = __ldg(xr + i);
      s += (t.x + t.y) + (t.z + t.w);
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
  if (CACHED) {
#pragma unroll
    for (int k = 0; k < LN_MAXCH; ++k)
      if (lane + 32 * k < C4) {
        float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
  } else {
    for (int i = lane; i < C4; i += 32) {
      float4 t = __ldg(xr + i);
      float a = t.x - mean, b = t.y - mean, c = t.z - mean, d = t.w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  if (lane == 0 && stats != nullptr) stats[row] = make_float2(mean, rstd);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  const size_t base = (size_t)row * Kpad;
  if (CACHED) {
#pragma unroll
    for (int k = 0; k < LN_MAXCH + 2; ++k) {     // + the zero padding up to Kpad (< 64 columns past C)
      const int i = lane + 32 * k;
      if (i >= (Kpad >> 2)) break;
      float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < LN_MAXCH && i < C4) {
        const float4 t = v[k < LN_MAXCH ? k : 0], g = __ldg(g4 + i), b = __ldg(b4 + i);
        y.x = fmaf((t.x - mean) * rstd, g.x, b.x);
        y.y = fmaf((t.y - mean) * rstd, g.y, b.y);
        y.z = fmaf((t.z - mean) * rstd, g.z, b.z);
        y.w = fmaf((t.w - mean) * rstd, g.w, b.w);
      }
      split_store4(hi, lo, base + 4 * (size_t)i, y.x, y.y, y.z, y.w);
    }
  } else {
    for (int i = lane; i < (Kpad >> 2); i += 32) {
      float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < C4) {
        float4 t = __ldg(xr + i), g = __ldg(g4 + i), b = __ldg(b4 + i);
        y.x = fmaf((t.x - mean) * rstd, g.x, b.x);
        y.y = fmaf((t.y - mean) * rstd, g.y, b.y);
        y.z = fmaf((t.z - mean) * rstd, g.z, b.z);
        y.w = fmaf((t.w - mean) * rstd, g.w, b.w);
      }
      split_store4(hi, lo, base + 4 * (size_t)i, y.x, y.y, y.z, y.w);
    }
  }
}

// dx = rstd * (a - mean(a) - xhat * mean(a * xhat)),  a = g * gamma,  xhat = (x - mean) * rstd
template <bool CACHED>
__global__ void __launch_bounds__(ROW_WARPS * 32) ln_bwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ g,
                                                                int64_t ldg, int rows, int C, const float* __restrict__ gamma,
                                                                const float2* __restrict__ stats, float* __restrict__ dx,
                                                                int64_t lddx) {
  const int lane = threadIdx.x & 31, row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * ldx);
  const float4* gr = reinterpret_cast<const float4*>(g + (size_t)row * ldg);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float2 st = stats[row];
  const int C4 = C >> 2;
  float4 xh[LN_MAXCH], av[LN_MAXCH];   // CACHED: xhat and a = g * gamma of this lane's chunks
  float s1 = 0.f, s2 = 0.f;
  if (CACHED) {
#pragma unroll
    for (int k = 0; k < LN_MAXCH; ++k) {
      const int i = lane + 32 * k;
      xh[k] = av[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < C4) {
        const float4 v = __ldg(xr + i), gg = __ldg(gr + i), w = __ldg(g4 + i);
        xh[k] = make_float4((v.x - st.x) * st.y, (v.y - st.x) * st.y, (v.z - st.x) * st.y, (v.w - st.x) * st.y);
        av[k] = make_float4(gg.x * w.x, gg.y * w.y, gg.z * w.z, gg.w * w.w);
        s1 += (av[k].x + av[k].y) + (av[k].z + av[k].w);
        s2 += av[k].x * xh[k].x + av[k].y * xh[k].y + av[k].z * xh[k].z + av[k].w * xh[k].w;
      }
    }
  } else {
    for (int i = lane; i < C4; i += 32) {
      float4 v = __ldg(xr + i), gg = __ldg(gr + i), w = __ldg(g4 + i);
      float a0 = gg.x * w.x, a1 = gg.y * w.y, a2 = gg.z * w.z, a3 = gg.w * w.w;
      s1 += (a0 + a1) + (a2 + a3);
      s2 += a0 * ((v.x - st.x) * st.y) + a1 * ((v.y - st.x) * st.y) + a2 * ((v.z - st.x) * st.y) + a3 * ((v.w - st.x) * st.y);
    }
  }
  const float m1 = warp_sum(s1) / (float)C, m2 = warp_sum(s2) / (float)C;
  float4* dr = reinterpret_cast<float4*>(dx + (size_t)row * lddx);
  if (CACHED) {
#pragma unroll
    for (int k = 0; k < LN_MAXCH; ++k) {
      const int i = lane + 32 * k;
      if (i < C4)
        dr[i] = make_float4(st.y * (av[k].x - m1 - xh[k].x * m2), st.y * (av[k].y - m1 - xh[k].y * m2),
                            st.y * (av[k].z - m1 - xh[k].z * m2), st.y * (av[k].w - m1 - xh[k].w * m2));
    }
  } else {
    for (int i = lane; i < C4; i += 32) {
      float4 v = __ldg(xr + i), gg = __ldg(gr + i), w = __ldg(g4 + i), o;
      o.x = st.y * (gg.x * w.x - m1 - (v.x - st.x) * st.y * m2);
      o.y = st.y * (gg.y * w.y - m1 - (v.y - st.x) * st.y * m2);
      o.z = st.y * (gg.z * w.z - m1 - (v.z - st.x) * st.y * m2);
      o.w = st.y * (gg.w * w.w - m1 - (v.w - st.x) * st.y * m2);
      dr[i] = o;
    }
  }
}

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float v) {
  return 0.5f * (1.f + erff(v * 0.70710678118654752f)) + v * 0.3989422804014327f * __expf(-0.5f * v * v);
}

// proj[rows, 2*H] = (a | gate) -> hi/lo[rows, Kpad] = split(a * gelu(gate))
__global__ void __launch_bounds__(256) geglu_split_kernel(const float* __restrict__ proj, int64_t ld, int rows, int H,
                                                          __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int Kpad) {
  const int K4 = Kpad >> 2, H4 = H >> 2;
  const int64_t total = (int64_t)rows * K4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int r = (int)(i / K4), c = (int)(i - (int64_t)r * K4);
    float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < H4) {
      const float4* pr = reinterpret_cast<const float4*>(proj + (size_t)r * ld);
      float4 a = __ldg(pr + c), gt = __ldg(pr + H4 + c);
      y.x = a.x * gelu_erf(gt.x);
      y.y = a.y * gelu_erf(gt.y);
      y.z = a.z * gelu_erf(gt.z);
      y.w = a.w * gelu_erf(gt.w);
    }
    split_store4(hi, lo, (size_t)r * Kpad + 4 * (size_t)c, y.x, y.y, y.z, y.w);
  }
}

// d_proj[rows, 2*H] = (g * gelu(gate) | g * a * gelu'(gate))
__global__ void __launch_bounds__(256) geglu_bwd_kernel(const float* __restrict__ proj, int64_t ld, const float* __restrict__ g,
                                                        int64_t ldg, int rows, int H, float* __restrict__ dproj, int64_t ldd) {
  const int H4 = H >> 2;
  const int64_t total = (int64_t)rows * H4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int r = (int)(i / H4), c = (int)(i - (int64_t)r * H4);
    const float4* pr = reinterpret_cast<const float4*>(proj + (size_t)r * ld);
    float4 a = __ldg(pr + c), gt = __ldg(pr + H4 + c), gg = __ldg(reinterpret_cast<const float4*>(g + (size_t)r * ldg) + c);
    float4 da, dg;
    da.x = gg.x * gelu_erf(gt.x); dg.x = gg.x * a.x * gelu_erf_grad(gt.x);
    da.y = gg.y * gelu_erf(gt.y); dg.y = gg.y * a.y * gelu_erf_grad(gt.y);
    da.z = gg.z * gelu_erf(gt.z); dg.z = gg.z * a.z * gelu_erf_grad(gt.z);
    da.w = gg.w * gelu_erf(gt.w); dg.w = gg.w * a.w * gelu_erf_grad(gt.w);
    float4* dr = reinterpret_cast<float4*>(dproj + (size_t)r * ldd);
    dr[c] = da;
    dr[H4 + c] = dg;
  }
}

// x[rows, cols] -> hi/lo[rows, Kpad] = split(softmax(x, -1)).  One CTA per row, three sweeps (max, sum, write) over a row
// that stays in L1/L2.  Used by the dense (GEMM-formulated) attention of the VAE mid block: one head of 512 channels.
__global__ void __launch_bounds__(256) softmax_split_kernel(const float* __restrict__ x, int64_t ldx, int cols,
                                                            __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int Kpad) {
  __shared__ float red[32];
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)blockIdx.x * ldx);
  const int C4 = cols >> 2, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float m = -3.0e38f;
  for (int i = threadIdx.x; i < C4; i += 256) {
    float4 v = __ldg(xr + i);
    m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  m = warp_max(m);
  if (lane == 0) red[w] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < C4; i += 256) {
    float4 v = __ldg(xr + i);
    sum += (__expf(v.x - m) + __expf(v.y - m)) + (__expf(v.z - m) + __expf(v.w - m));
  }
  const float inv = 1.f / block_sum(sum, red);
  const size_t base = (size_t)blockIdx.x * Kpad;
  for (int i = threadIdx.x; i < (Kpad >> 2); i += 256) {
    float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < C4) {
      float4 v = __ldg(xr + i);
      y = make_float4(__expf(v.x - m) * inv, __expf(v.y - m) * inv, __expf(v.z - m) * inv, __expf(v.w - m) * inv);
    }
    split_store4(hi, lo, base + 4 * (size_t)i, y.x, y.y, y.z, y.w);
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace skp

using namespace skp;

extern "C" int skp_ln_split_fwd(const float* x, int64_t ldx, int rows, int C, const float* gamma, const float* beta, float eps,
                                void* hi, void* lo, int Kpad, float* stats, void* stream) {
  SKP_REQUIRE(x && gamma && beta && hi && lo, "skp_ln_split_fwd: null pointer");
  SKP_REQUIRE(rows > 0 && C > 0 && C % 4 == 0 && Kpad >= C && Kpad % 4 == 0 && ldx % 4 == 0, "skp_ln_split_fwd: bad sizes rows=%d C=%d Kpad=%d", rows, C, Kpad);
  SKP_REQUIRE(aligned16(x) && aligned16(gamma) && aligned16(beta) && (reinterpret_cast<uintptr_t>(hi) & 7) == 0 &&
                  (reinterpret_cast<uintptr_t>(lo) & 7) == 0 && (stats == nullptr || (reinterpret_cast<uintptr_t>(stats) & 7) == 0),
              "skp_ln_split_fwd: misaligned pointer");
  const int grid = (rows + ROW_WARPS - 1) / ROW_WARPS;
  if (C <= 128 * LN_MAXCH)
    ln_split_kernel<true><<<grid, ROW_WARPS * 32, 0, (cudaStream_t)stream>>>(x, ldx, rows, C, gamma, beta, eps, (__nv_bfloat16*)hi,
                                                                             (__nv_bfloat16*)lo, Kpad, reinterpret_cast<float2*>(stats));
  else
    ln_split_kernel<false><<<grid, ROW_WARPS * 32, 0, (cudaStream_t)stream>>>(x, ldx, rows, C, gamma, beta, eps, (__nv_bfloat16*)hi,
                                                                              (__nv_bfloat16*)lo, Kpad, reinterpret_cast<float2*>(stats));
  SKP_CHECK_LAUNCH("ln_split_kernel");
  return SKP_OK;
}

extern "C" int skp_ln_bwd(const float* x, int64_t ldx, const float* g, int64_t ldg, int rows, int C, const float* gamma,
                          const float* stats, float* dx, int64_t lddx, void* stream) {
  SKP_REQUIRE(x && g && gamma && stats && dx, "skp_ln_bwd: null pointer");
  SKP_REQUIRE(rows > 0 && C > 0 && C % 4 == 0 && ldx % 4 == 0 && ldg % 4 == 0 && lddx % 4 == 0, "skp_ln_bwd: bad sizes rows=%d C=%d", rows, C);
  SKP_REQUIRE(aligned16(x) && aligned16(g) && aligned16(gamma) && aligned16(dx) && (reinterpret_cast<uintptr_t>(stats) & 7) == 0,
              "skp_ln_bwd: misaligned pointer");
  const int grid = (rows + ROW_WARPS - 1) / ROW_WARPS;
  if (C <= 128 * LN_MAXCH)
    ln_bwd_kernel<true><<<grid, ROW_WARPS * 32, 0, (cudaStream_t)stream>>>(x, ldx, g, ldg, rows, C, gamma,
                                                                           reinterpret_cast<const float2*>(stats), dx, lddx);
  else
    ln_bwd_kernel<false><<<grid, ROW_WARPS * 32, 0, (cudaStream_t)stream>>>(x, ldx, g, ldg, rows, C, gamma,
                                                                            reinterpret_cast<const float2*>(stats), dx, lddx);
  SKP_CHECK_LAUNCH("ln_bwd_kernel");
  return SKP_OK;
}

extern "C" int skp_geglu_split_fwd(const float* proj, int64_t ld, int rows, int H, void* hi, void* lo, int Kpad, void* stream) {
  SKP_REQUIRE(proj && hi && lo, "skp_geglu_split_fwd: null pointer");
  SKP_REQUIRE(rows > 0 && H > 0 && H % 4 == 0 && Kpad >= H && Kpad % 4 == 0 && ld % 4 == 0 && aligned16(proj), "skp_geglu_split_fwd: bad sizes rows=%d H=%d Kpad=%d", rows, H, Kpad);
  int64_t total = (int64_t)rows * (Kpad / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  geglu_split_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(proj, ld, rows, H, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, Kpad);
  SKP_CHECK_LAUNCH("geglu_split_kernel");
  return SKP_OK;
}

extern "C" int skp_geglu_bwd(const float* proj, int64_t ld, const float* g, int64_t ldg, int rows, int H, float* dproj,
                             int64_t ldd, void* stream) {
  SKP_REQUIRE(proj && g && dproj, "skp_geglu_bwd: null pointer");
  SKP_REQUIRE(rows > 0 && H > 0 && H % 4 == 0 && ld % 4 == 0 && ldg % 4 == 0 && ldd % 4 == 0 && aligned16(proj) && aligned16(g) && aligned16(dproj),
              "skp_geglu_bwd: bad sizes rows=%d H=%d", rows, H);
  int64_t total = (int64_t)rows * (H / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  geglu_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(proj, ld, g, ldg, rows, H, dproj, ldd);
  SKP_CHECK_LAUNCH("geglu_bwd_kernel");
  return SKP_OK;
}

extern "C" int skp_softmax_split_fwd(const float* x, int64_t ldx, int rows, int cols, void* hi, void* lo, int Kpad, void* stream) {
  SKP_REQUIRE(x && hi && lo, "skp_softmax_split_fwd: null pointer");
  SKP_REQUIRE(rows > 0 && cols > 0 && cols % 4 == 0 && Kpad >= cols && Kpad % 4 == 0 && ldx % 4 == 0 && aligned16(x),
              "skp_softmax_split_fwd: bad sizes rows=%d cols=%d Kpad=%d", rows, cols, Kpad);
  softmax_split_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(x, ldx, cols, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, Kpad);
  SKP_CHECK_LAUNCH("softmax_split_kernel");
  return SKP_OK;
}
