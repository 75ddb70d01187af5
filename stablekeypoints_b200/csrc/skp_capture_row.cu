// Attention-store kernel, row formulation (forward, STORE): ptp_utils.py:508-538 by linearity (see skp_capture.cu):
//   probs[h, Y*R + X, :] = softmax_tokens( bicubic(logits[h])(Y, X, :) )
// The 40 MB/layer store is the only HBM traffic that matters (logits <= 2.5 MB, L2-resident), so the kernel is built
// around getting a full output row out of shared memory with as few instructions per element as possible:
//
//   CTA = (2 consecutive output rows, head h, x-block), two threads per output pixel X (each owns half of the token
//   axis).  An x-block is the whole row when it fits ~100 KB of staging (N = 77, 100), else 32 pixels (N = 500).
//   1. vertical pass   V[xs][n] = sum_j wy[j] * L[h, row_j, xs, n]   (thread = column x 4 tokens, one 128-bit shared
//      store), with 2 replicated halo columns on each side so the horizontal taps never clamp; M = max |V| of the row.
//   2. horizontal pass, thread = (pixel, token slice): x_n = sum_i wx[i] * V[ix-1+i][n] with 128-bit shared loads (4 tokens
//      per load; lanes of a warp share <= 6 distinct columns -> broadcast, conflict-free because the column stride is
//      an odd number of float4), e_n = exp2(x_n - U) with U = M * sum_i |wx[i]| an upper bound of max_n x_n (no max
//      sweep; softmax is shift-invariant and fp32 keeps its relative precision), e_n staged at [X][n] -- exactly the
//      global layout of the row -- while the thread accumulates the sum of its token slice (the two slices of a pixel
//      meet through one shared float each).  The bicubic weights of a pixel are computed once for both rows.
//   3. the thread rescales its slice by 1/sum in place; then ONE bulk asynchronous copy (cp.async.bulk, the TMA
//      engine) moves the contiguous x-block of the row from shared memory to HBM while the CTA starts its second row.
// 9.3 M warp instructions for the 10.1 M elements of a layer at N = 77 (56 M for the tile kernel it replaced).
// If U - max_n x_n is so loose that the sum underflows (pathological logits), the pixel is redone with the exact max.
#include "skp_common.cuh"
#include <math_constants.h>
#include <stdlib.h>

namespace skp {

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int ROW_MAX_THREADS = 512;

// token slices per pixel: 2 for the whole-row blocks of N = 77 / 100 (128 pixel lanes); wide token axes (N = 500: x-blocks
// of 32 pixels) get as many slices as fill the CTA -- 64 threads per SM cannot hide anything
static inline int row_token_slices(int N, int P) {
  if (N < 16) return 1;
  int ts = 2;
  if (P <= 64) {
    ts = ROW_MAX_THREADS / P;
    if (ts > 16) ts = 16;
    while (ts > 2 && (N >> 2) / ts < 4) ts >>= 1;       // keep >= 4 float4 token groups per slice
  }
  return ts;
}
constexpr int ROWS_PER_CTA = 2;   // consecutive output rows per CTA: the per-pixel horizontal setup is shared

template <bool ONE_BLOCK>
__global__ void __launch_bounds__(ROW_MAX_THREADS, 2) capture_store_row_kernel(const float* __restrict__ logits,
                                                                            float* __restrict__ probs, int s, int N, int R,
                                                                            int NV, int P, int TS, int XB_) {
  const int XB = ONE_BLOCK ? R : XB_;
  extern __shared__ __align__(16) unsigned char row_smem[];
  float* stage = reinterpret_cast<float*>(row_smem);                 // [XB][N]  (an x-block of the output row, global layout)
  float* Vs = stage + (((size_t)XB * N + 3) & ~(size_t)3);           // [s+4][NV]
  float* red = Vs + (size_t)(s + 4) * NV;                            // [32] per-warp max |V|
  float* psum = red + 32;                                            // [TS][P] partial softmax sums
  const int h = blockIdx.y;
  const int tid = threadIdx.x, NT = blockDim.x;
  const float scale = (float)s / (float)R;
  const float LOG2E = 1.4426950408889634f;
  // P pixel lanes x TS token slices: thread = (pixel X, slice of the token axis); slices meet through psum[].
  const int N4 = N >> 2;
  const int X_lane = tid % P, part = tid / P;
  const int g0 = (part * N4) / TS, g1 = ((part + 1) * N4) / TS;      // float4 token groups of this slice
  const int n_lo = 4 * g0, n_hi = (part == TS - 1) ? N : 4 * g1;     // the last slice also takes the N % 4 tail
  const bool vec4 = (N & 3) == 0;
  // wide token axes (N = 500) do not fit a whole row: the row leaves in x-blocks of XB pixels, one after the other, out of
  // the SAME vertical pass.  With a single x-block the taps of this thread's pixel are the same for every row of the CTA.
  constexpr bool one_block = ONE_BLOCK;
  float pwx[4] = {0.f, 0.f, 0.f, 0.f}, psabs = 0.f;
  int pc0 = 1;
  if (one_block && X_lane < R) {
    float rx = scale * (X_lane + 0.5f) - 0.5f, fx = floorf(rx);
    cubic_coeffs(rx - fx, pwx);
    pc0 = (int)fx + 1;   // column of tap 0 in the halo'd array (ix - 1 + 2)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      pwx[i] *= LOG2E;
      psabs += fabsf(pwx[i]);
    }
  }
  bool store_pending = false;                                        // tid 0: a bulk copy may still be reading the staging tile

  for (int rr = 0; rr < ROWS_PER_CTA; ++rr) {
  const int Y = blockIdx.x * ROWS_PER_CTA + rr;
  if (Y >= R) break;
  // ---- 1. vertical pass: thread = (low-res column xs, group of 4 tokens); one 128-bit shared store per item
  float amax = 0.f;
  {
    float ry = scale * (Y + 0.5f) - 0.5f, fy = floorf(ry);
    float wy[4];
    cubic_coeffs(ry - fy, wy);
    int iy = (int)fy;
    const float* rows[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int r = iy - 1 + j;
      r = r < 0 ? 0 : (r > s - 1 ? s - 1 : r);
      rows[j] = logits + ((size_t)h * s + r) * s * N;
    }
    const int NV4 = NV >> 2, items = s * NV4;
    const float inv_nv4 = 1.f / (float)NV4;
    for (int i = tid; i < items; i += NT) {
      int xs = (int)((i + 0.5f) * inv_nv4);          // i / NV4 (exact: the quotient is never within 0.5/NV4 of an integer)
      int n0 = (i - xs * NV4) << 2;
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int n = n0 + k;
        float a = 0.f;
        if (n < N) {
          int o = xs * N + n;
          a = wy[0] * __ldg(rows[0] + o);
          a = fmaf(wy[1], __ldg(rows[1] + o), a);
          a = fmaf(wy[2], __ldg(rows[2] + o), a);
          a = fmaf(wy[3], __ldg(rows[3] + o), a);
          amax = fmaxf(amax, fabsf(a));
        }
        v[k] = a;   // token padding = 0
      }
      float4 q = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(Vs + (xs + 2) * NV + n0) = q;
      if (xs == 0) {            // replicated halo columns: the horizontal taps never clamp
        *reinterpret_cast<float4*>(Vs + n0) = q;
        *reinterpret_cast<float4*>(Vs + NV + n0) = q;
      }
      if (xs == s - 1) {
        *reinterpret_cast<float4*>(Vs + (s + 2) * NV + n0) = q;
        *reinterpret_cast<float4*>(Vs + (s + 3) * NV + n0) = q;
      }
    }
    amax = warp_max(amax);
    if ((tid & 31) == 0) red[tid >> 5] = amax;
  }
  // single x-block: the wait for the previous row's copy rides on the barrier that publishes the vertical pass
  if (one_block && tid == 0 && store_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncthreads();
  // M = max |V| over the row's footprint: x_n = sum_i wx[i] V_i[n] <= (sum_i |wx[i]|) * M for every token
  float M = 0.f;
  for (int w = 0; w < (NT >> 5); ++w) M = fmaxf(M, red[w]);

  for (int xb0 = 0; xb0 < R; xb0 += XB) {
  const int xb1 = min(R, xb0 + XB);
  if (!one_block) {
    if (tid == 0 && store_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging is free again
    __syncthreads();
  }
  // ---- 2. + 3. horizontal pass, softmax over tokens, in-place normalisation.
  for (int X0 = xb0; X0 < xb1; X0 += P) {
    const int X = X0 + X_lane;
    const bool live = X < xb1;
    float wx[4] = {0.f, 0.f, 0.f, 0.f};
    int c0 = 1;
    float U = 0.f, sum = 0.f;
    float* orow = stage + (size_t)(live ? X - xb0 : 0) * N;
    if (live) {
      if (one_block && X0 == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) wx[i] = pwx[i];
        c0 = pc0;
        U = psabs * M;
      } else {
        float rx = scale * (X + 0.5f) - 0.5f, fx = floorf(rx);
        cubic_coeffs(rx - fx, wx);
        c0 = (int)fx + 1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          wx[i] *= LOG2E;
          U += fabsf(wx[i]);
        }
        U *= M;
      }
      const float4* v0 = reinterpret_cast<const float4*>(Vs + (size_t)c0 * NV);
      const float4* v1 = reinterpret_cast<const float4*>(Vs + (size_t)(c0 + 1) * NV);
      const float4* v2 = reinterpret_cast<const float4*>(Vs + (size_t)(c0 + 2) * NV);
      const float4* v3 = reinterpret_cast<const float4*>(Vs + (size_t)(c0 + 3) * NV);
#pragma unroll 4
      for (int g = g0; g < g1; ++g) {
        float4 a = v0[g], b = v1[g], c = v2[g], d = v3[g];
        float e0 = ex2_approx(fmaf(wx[3], d.x, fmaf(wx[2], c.x, fmaf(wx[1], b.x, fmaf(wx[0], a.x, -U)))));
        float e1 = ex2_approx(fmaf(wx[3], d.y, fmaf(wx[2], c.y, fmaf(wx[1], b.y, fmaf(wx[0], a.y, -U)))));
        float e2 = ex2_approx(fmaf(wx[3], d.z, fmaf(wx[2], c.z, fmaf(wx[1], b.z, fmaf(wx[0], a.z, -U)))));
        float e3 = ex2_approx(fmaf(wx[3], d.w, fmaf(wx[2], c.w, fmaf(wx[1], b.w, fmaf(wx[0], a.w, -U)))));
        if (vec4) {                       // N % 4 == 0: the pixel rows of the staging tile are 16-byte aligned (conflict-free
          *reinterpret_cast<float4*>(orow + 4 * g) = make_float4(e0, e1, e2, e3);   // 128-bit stores; scalar ones collide 4-way
        } else {                          // when the row pitch N is a multiple of 4 floats)
          orow[4 * g] = e0;
          orow[4 * g + 1] = e1;
          orow[4 * g + 2] = e2;
          orow[4 * g + 3] = e3;
        }
        sum += (e0 + e1) + (e2 + e3);
      }
      if (part == TS - 1)
        for (int n = N4 * 4; n < N; ++n) {
          float x = fmaf(wx[3], Vs[(c0 + 3) * NV + n], fmaf(wx[2], Vs[(c0 + 2) * NV + n],
                    fmaf(wx[1], Vs[(c0 + 1) * NV + n], fmaf(wx[0], Vs[c0 * NV + n], -U))));
          float e = ex2_approx(x);
          orow[n] = e;
          sum += e;
        }
      psum[part * P + X_lane] = sum;
    }
    __syncthreads();
    if (live) {
      float total = 0.f;
      for (int q = 0; q < TS; ++q) total += psum[q * P + X_lane];
      if (!(total > 1e-30f) || !(total < 1e30f)) {
        // the bound was too loose (or not finite): slice 0 redoes the whole pixel with the exact max, the others stand by
        if (part == 0) {
          float m = -CUDART_INF_F;
          for (int n = 0; n < N; ++n) {
            float x = fmaf(wx[3], Vs[(c0 + 3) * NV + n], fmaf(wx[2], Vs[(c0 + 2) * NV + n],
                      fmaf(wx[1], Vs[(c0 + 1) * NV + n], wx[0] * Vs[c0 * NV + n])));
            orow[n] = x;
            m = fmaxf(m, x);
          }
          float t = 0.f;
          for (int n = 0; n < N; ++n) {
            float e = exp2f(orow[n] - m);
            orow[n] = e;
            t += e;
          }
          const float inv = 1.f / t;
          for (int n = 0; n < N; ++n) orow[n] *= inv;
        }
      } else {
        const float inv = 1.f / total;
        if (vec4) {
          float4* o4 = reinterpret_cast<float4*>(orow);
#pragma unroll 4
          for (int g = g0; g < g1; ++g) {
            float4 v = o4[g];
            v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
            o4[g] = v;
          }
        } else {
#pragma unroll 8
          for (int n = n_lo; n < n_hi; ++n) orow[n] *= inv;
        }
      }
    }
    __syncthreads();   // psum is reused by the next pixel block
  }
  // ---- bulk store of the x-block: generic-proxy writes -> async proxy, then one thread drives the copy engine
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)((size_t)(xb1 - xb0) * N * sizeof(float));
    char* dst = reinterpret_cast<char*>(probs + (((size_t)h * R + Y) * R + xb0) * N);
    uint32_t src = (uint32_t)__cvta_generic_to_shared(stage);
    for (uint32_t off = 0; off < bytes; off += 16384u) {
      uint32_t n = bytes - off < 16384u ? bytes - off : 16384u;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(src + off), "r"(n)
                   : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    store_pending = true;
  }
  }   // x-blocks of the row
  }   // rows of this CTA
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem must outlive the engine's reads
}

// ------------------------------------------------------------------------------------------------ backward (fused mean)
// d_logits[l][h, ys, xs, n] += bicubic^T( p o (g - <p, g>) ),  g[n] = w * d_maps[n, Y, X],  p = softmax_tokens(bicubic(logits)):
// the input gradient of  maps = mean_{l,h} softmax_tokens(bicubic(logits[l][h]))  (ptp_utils.py:508-538 + optimize.py:50-75).
// Two kernels, no atomics, bit-reproducible:
//  (1) capture_mean_row_bwd_kernel, CTA = (output row Y, head h): the same row formulation as the forward.  The vertical
//      pass runs once; then, x-block by x-block (the whole row when it fits, 32 pixels at N = 500), the probabilities are
//      recomputed, the gradient of the scores replaces them in the staging tile, and the TRANSPOSED horizontal stencil is
//      applied as a gather (thread = (low-res column, token) sums the <= 4*F pixels whose taps touch that column) into a
//      shared accumulator that lives across the x-blocks.  Result: dV[h, Y, xs, n], the gradient with respect to the
//      vertically interpolated logits of row Y, written once, coalesced.
//  (2) capture_mean_vgather_kernel: the transposed VERTICAL stencil as a gather too -- thread = (h, ys, xs, 4 tokens) sums
//      wy_j(Y) * dV[h, Y, xs, n] over the <= 4*F+ rows Y that have ys among their four (clamped) taps.
// (The one-kernel version scattered 4*s*N global atomics per (row, x-block, head): 131 M per layer at N = 500, 3.8 ms.)
__global__ void __launch_bounds__(ROW_MAX_THREADS, 2) capture_mean_row_bwd_kernel(const float* __restrict__ logits,
                                                                               const float* __restrict__ d_maps,
                                                                               float* __restrict__ dvrow, int s, int N, int R,
                                                                               int NV, int P, int TS, int XB, int XBW, float w) {
  extern __shared__ __align__(16) unsigned char row_smem[];
  const int NS = (N + 3) & ~3;                                       // staging row pitch: whole float4 token groups; the padding
                                                                     // tokens carry exp = 0 (vertical tile padded with a large
                                                                     // negative number), so every N runs the 128-bit path
  float* stage = reinterpret_cast<float*>(row_smem);                 // [XB][NS]  e_n, then dS_n
  float* Vs = stage + (size_t)XB * NS;                               // [s+4][NV] vertically interpolated logits
  float* dVs = Vs + (size_t)(s + 4) * NV;                            // [s+4][NV] gradient wrt Vs (accumulated over the x-blocks)
  float* red = dVs + (size_t)(s + 4) * NV;                           // [32]
  float* psum = red + 32;                                            // [2][TS][P]  partial (sum e, sum e*g)
  float* wtab = psum + 2 * TS * P;                                   // [XB][4] raw horizontal weights
  int* c0tab = reinterpret_cast<int*>(wtab + 4 * XB);                // [XB]    first halo'd column of the pixel's taps
  int* xrange = c0tab + XB;                                          // [s+4][2] pixel range (local) touching a column
  float* wcol = reinterpret_cast<float*>(xrange + 2 * (s + 4));      // [s+4][XBW] transposed-stencil weights per (column, pixel of its range)
  const int Y = blockIdx.x, h = blockIdx.y;
  const int tid = threadIdx.x, NT = blockDim.x;
  const float scale = (float)s / (float)R;
  const float LOG2E = 1.4426950408889634f;
  const int N4 = NS >> 2;
  const int X_lane = tid % P, part = tid / P;
  const int g0 = (part * N4) / TS, g1 = ((part + 1) * N4) / TS;
  const int n_lo = 4 * g0, n_hi = 4 * g1;                            // whole groups; tokens >= N are padding

  // vertical weights / rows of this output row
  float wy[4];
  int rowj[4];
  {
    float ry = scale * (Y + 0.5f) - 0.5f, fy = floorf(ry);
    cubic_coeffs(ry - fy, wy);
    const int iy = (int)fy;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int r = iy - 1 + j;
      rowj[j] = r < 0 ? 0 : (r > s - 1 ? s - 1 : r);
    }
  }
  // ---- 1. vertical pass (as the forward), once per row; the dVs accumulator starts at zero
  float amax = 0.f;
  {
    const float* rows[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) rows[j] = logits + ((size_t)h * s + rowj[j]) * s * N;
    const int NV4 = NV >> 2, items = s * NV4;
    const float inv_nv4 = 1.f / (float)NV4;
    for (int i = tid; i < items; i += NT) {
      int xs = (int)((i + 0.5f) * inv_nv4);
      int n0 = (i - xs * NV4) << 2;
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int n = n0 + k;
        float a = -1.0e4f;                                           // padding token: exp2(w . pad - U) == 0 exactly
        if (n < N) {
          int o = xs * N + n;
          a = wy[0] * __ldg(rows[0] + o);
          a = fmaf(wy[1], __ldg(rows[1] + o), a);
          a = fmaf(wy[2], __ldg(rows[2] + o), a);
          a = fmaf(wy[3], __ldg(rows[3] + o), a);
          amax = fmaxf(amax, fabsf(a));
        }
        v[k] = a;
      }
      float4 q = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(Vs + (xs + 2) * NV + n0) = q;
      if (xs == 0) {
        *reinterpret_cast<float4*>(Vs + n0) = q;
        *reinterpret_cast<float4*>(Vs + NV + n0) = q;
      }
      if (xs == s - 1) {
        *reinterpret_cast<float4*>(Vs + (s + 2) * NV + n0) = q;
        *reinterpret_cast<float4*>(Vs + (s + 3) * NV + n0) = q;
      }
    }
    for (int i = tid; i < (s + 4) * NV; i += NT) dVs[i] = 0.f;
    amax = warp_max(amax);
    if ((tid & 31) == 0) red[tid >> 5] = amax;
  }
  __syncthreads();
  float M = 0.f;
  for (int q = 0; q < (NT >> 5); ++q) M = fmaxf(M, red[q]);

  for (int xb0 = 0; xb0 < R; xb0 += XB) {
    const int xb1 = min(R, xb0 + XB), nx = xb1 - xb0;
    // per-pixel tap tables of this x-block (raw weights: the transposed stencil needs them un-scaled)
    for (int i = tid; i < nx; i += NT) {
      float rx = scale * (xb0 + i + 0.5f) - 0.5f, fx = floorf(rx);
      float cw[4];
      cubic_coeffs(rx - fx, cw);
      wtab[4 * i] = cw[0]; wtab[4 * i + 1] = cw[1]; wtab[4 * i + 2] = cw[2]; wtab[4 * i + 3] = cw[3];
      c0tab[i] = (int)fx + 1;
    }
    __syncthreads();
    // pixel range of every halo'd column: pixels with c0 <= c <= c0 + 3.  One shared-memory atomic min / max per (pixel, tap)
    // (the first version let s + 4 threads scan all nx pixels serially: ~2 us of one warp per x-block with everybody waiting)
    for (int c = tid; c < s + 4; c += NT) {
      xrange[2 * c] = nx;
      xrange[2 * c + 1] = -1;
    }
    __syncthreads();
    for (int i = tid; i < nx; i += NT) {
      const int c0 = c0tab[i];
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        atomicMin(&xrange[2 * (c0 + d)], i);
        atomicMax(&xrange[2 * (c0 + d) + 1], i);
      }
    }
    // ---- 2. horizontal pass: e_n staged, partial (sum e, sum e*g) per token slice; g read coalesced over X from d_maps
    for (int X0 = 0; X0 < nx; X0 += P) {
      const int xi = X0 + X_lane;           // local pixel
      const bool live = xi < nx;
      const int X = xb0 + xi;
      float wx[4] = {0.f, 0.f, 0.f, 0.f};
      int c0 = 1;
      float U = 0.f, s1 = 0.f, s2 = 0.f;
      float* orow = stage + (size_t)(live ? xi : 0) * NS;
      const float* grow = d_maps + (size_t)Y * R + (live ? X : 0);     // + n*R*R per token
      if (live) {
        c0 = c0tab[xi];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          wx[i] = wtab[4 * xi + i] * LOG2E;
          U += fabsf(wx[i]);
        }
        U *= M;
        {                                                            // four tokens per step: 128-bit tile loads and staging stores
          const float4* v0 = reinterpret_cast<const float4*>(Vs + c0 * NV);
          const float4* v1 = reinterpret_cast<const float4*>(Vs + (c0 + 1) * NV);
          const float4* v2 = reinterpret_cast<const float4*>(Vs + (c0 + 2) * NV);
          const float4* v3 = reinterpret_cast<const float4*>(Vs + (c0 + 3) * NV);
          const size_t RR = (size_t)R * R;
          for (int n = n_lo; n < n_hi; n += 4) {
            const float4 a = v0[n >> 2], b = v1[n >> 2], c = v2[n >> 2], dd = v3[n >> 2];
            const float* gp = grow + (size_t)n * RR;
            const float g0v = __ldg(gp), g1v = n + 1 < N ? __ldg(gp + RR) : 0.f, g2v = n + 2 < N ? __ldg(gp + 2 * RR) : 0.f,
                        g3v = n + 3 < N ? __ldg(gp + 3 * RR) : 0.f;
            float4 e;
            e.x = ex2_approx(fmaf(wx[3], dd.x, fmaf(wx[2], c.x, fmaf(wx[1], b.x, fmaf(wx[0], a.x, -U)))));
            e.y = ex2_approx(fmaf(wx[3], dd.y, fmaf(wx[2], c.y, fmaf(wx[1], b.y, fmaf(wx[0], a.y, -U)))));
            e.z = ex2_approx(fmaf(wx[3], dd.z, fmaf(wx[2], c.z, fmaf(wx[1], b.z, fmaf(wx[0], a.z, -U)))));
            e.w = ex2_approx(fmaf(wx[3], dd.w, fmaf(wx[2], c.w, fmaf(wx[1], b.w, fmaf(wx[0], a.w, -U)))));
            *reinterpret_cast<float4*>(orow + n) = e;
            s1 += (e.x + e.y) + (e.z + e.w);
            s2 = fmaf(e.x, g0v, fmaf(e.y, g1v, fmaf(e.z, g2v, fmaf(e.w, g3v, s2))));
          }
        }
        psum[part * P + X_lane] = s1;
        psum[(TS + part) * P + X_lane] = s2;
      }
      __syncthreads();
      if (live) {
        float t1 = 0.f, t2 = 0.f;
        for (int q = 0; q < TS; ++q) {
          t1 += psum[q * P + X_lane];
          t2 += psum[(TS + q) * P + X_lane];
        }
        if (!(t1 > 1e-30f) || !(t1 < 1e30f)) {
          // loose bound: slice 0 redoes the whole pixel with the exact max
          if (part == 0) {
            float m = -CUDART_INF_F;
            for (int n = 0; n < N; ++n) {
              float x = fmaf(wx[3], Vs[(c0 + 3) * NV + n], fmaf(wx[2], Vs[(c0 + 2) * NV + n],
                        fmaf(wx[1], Vs[(c0 + 1) * NV + n], wx[0] * Vs[c0 * NV + n])));
              orow[n] = x;
              m = fmaxf(m, x);
            }
            float a1 = 0.f, a2 = 0.f;
            for (int n = 0; n < N; ++n) {
              float e = exp2f(orow[n] - m);
              orow[n] = e;
              a1 += e;
              a2 = fmaf(e, __ldg(grow + (size_t)n * R * R), a2);
            }
            const float inv = 1.f / a1, dot = a2 * inv;
            for (int n = 0; n < N; ++n) orow[n] = orow[n] * inv * (__ldg(grow + (size_t)n * R * R) - dot) * w;
            for (int n = N; n < NS; ++n) orow[n] = 0.f;
          }
        } else {
          const float inv = 1.f / t1, dot = t2 * inv;
          const float iw = inv * w;
          const size_t RR = (size_t)R * R;
          for (int n = n_lo; n < n_hi; n += 4) {
            const float* gp = grow + (size_t)n * RR;
            float4 e = *reinterpret_cast<const float4*>(orow + n);
            e.x *= (__ldg(gp) - dot) * iw;
            e.y *= ((n + 1 < N ? __ldg(gp + RR) : 0.f) - dot) * iw;   // (padding tokens: e == 0)
            e.z *= ((n + 2 < N ? __ldg(gp + 2 * RR) : 0.f) - dot) * iw;
            e.w *= ((n + 3 < N ? __ldg(gp + 3 * RR) : 0.f) - dot) * iw;
            *reinterpret_cast<float4*>(orow + n) = e;
          }
        }
      }
      __syncthreads();
    }
    // ---- 3. transposed horizontal stencil as a gather, accumulated over the x-blocks:
    //         dVs[c][n] += sum_X wx[X][c - c0[X]] * dS[X][n]
    // The weights of a column do not depend on the token: they are tabulated once per x-block (wcol), then a warp takes one
    // (column, 128-token chunk) item at a time -- four tokens per lane, 128-bit staging loads -- so the inner loop is one
    // broadcast weight load, one staging load and four FMAs.  (The first version looked the weight up per (column, token,
    // pixel) through two dependent table loads and divided a flat index by N per item: half of the kernel's instructions.)
    for (int i = tid; i < (s + 4) * XBW; i += NT) {
      const int c = i / XBW, k = i - c * XBW;
      const int xi = xrange[2 * c] + k;
      wcol[i] = xi <= xrange[2 * c + 1] ? wtab[4 * xi + (c - c0tab[xi])] : 0.f;
    }
    __syncthreads();
    {
      const int lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
      const int nchunks = (NS + 127) / 128;                          // 128 tokens per item
      for (int item = warp; item < (s + 4) * nchunks; item += nwarps) {
        const int c = item / nchunks, ch = item - c * nchunks;
        const int lo = xrange[2 * c], cnt = xrange[2 * c + 1] - lo + 1;
        if (cnt <= 0) continue;
        const float* wc = wcol + c * XBW;
        const int n = ch * 128 + 4 * lane;
        if (n < NS) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          const float* sp = stage + (size_t)lo * NS + n;
          for (int k = 0; k < cnt; ++k) {
            const float wk = wc[k];
            const float4 v = *reinterpret_cast<const float4*>(sp + (size_t)k * NS);
            a.x = fmaf(wk, v.x, a.x); a.y = fmaf(wk, v.y, a.y); a.z = fmaf(wk, v.z, a.z); a.w = fmaf(wk, v.w, a.w);
          }
          float4* d4 = reinterpret_cast<float4*>(dVs + c * NV + n);
          float4 o = *d4;
          o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
          *d4 = o;
        }
      }
    }
    __syncthreads();
  }
  // ---- 4. fold the replicated halo columns; the row's dV leaves once, coalesced
  float* dv = dvrow + (((size_t)h * R + Y) * s) * N;
  for (int i = tid; i < s * N; i += NT) {
    const int xs = i / N, n = i - xs * N;
    float v = dVs[(xs + 2) * NV + n];
    if (xs == 0) v += dVs[n] + dVs[NV + n];
    if (xs == s - 1) v += dVs[(s + 2) * NV + n] + dVs[(s + 3) * NV + n];
    dv[i] = v;
  }
}

// d_logits[h, ys, xs, n] += sum over the output rows Y whose (clamped) vertical taps include ys of wy_j(Y) * dV[h, Y, xs, n].
// CTA = (head, low-res row ys): the (<= VG_MAXW) contributing rows and their summed tap weights are tabulated once in shared
// memory, then every thread streams its (xs, n) elements: one coalesced load + FMA per contributing row.
constexpr int VG_MAXW = 96;
__global__ void __launch_bounds__(256) capture_mean_vgather_kernel(const float* __restrict__ dvrow, float* __restrict__ d_logits,
                                                                   int s, int N, int R) {
  __shared__ float wtab[VG_MAXW];
  __shared__ int ytab[VG_MAXW];
  __shared__ int cnt;
  const int ys = blockIdx.x, h = blockIdx.y;
  const float scale = (float)s / (float)R;
  if (threadIdx.x == 0) {
    // rows Y with iy = floor(scale*(Y+0.5)-0.5) in [ys-2, ys+1] touch ys un-clamped; at the borders clamping adds more:
    // scan a conservative window and test the taps exactly as the forward computes them
    int ylo = (int)floorf(((float)(ys - 2) + 0.5f) / scale - 0.5f) - 1, yhi = (int)ceilf(((float)(ys + 2) + 0.5f) / scale - 0.5f) + 1;
    if (ys == 0) ylo = 0;
    if (ys == s - 1) yhi = R - 1;
    ylo = max(ylo, 0);
    yhi = min(yhi, R - 1);
    int c = 0;
    for (int Y = ylo; Y <= yhi; ++Y) {
      const float ry = scale * (Y + 0.5f) - 0.5f, fy = floorf(ry);
      const int iy = (int)fy;
      float wyv[4];
      cubic_coeffs(ry - fy, wyv);
      float wsum = 0.f;
      bool hit = false;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int r = iy - 1 + j;
        r = r < 0 ? 0 : (r > s - 1 ? s - 1 : r);
        if (r == ys) { wsum += wyv[j]; hit = true; }
      }
      if (hit && c < VG_MAXW) { wtab[c] = wsum; ytab[c] = Y; ++c; }
    }
    cnt = c;
  }
  __syncthreads();
  const int nrows = cnt;
  const size_t row_stride = (size_t)s * N;                      // floats between consecutive output rows of dV
  const float* src = dvrow + (size_t)h * R * row_stride;
  float* dst = d_logits + ((size_t)h * s + ys) * row_stride;
  const int total = s * N;
  // blockIdx.z splits the (xs, token) range so that long token axes fill the GPU (N = 500: 128 -> 512 CTAs)
  const int chunk = (((total + (int)gridDim.z - 1) / (int)gridDim.z) + 3) & ~3;
  const int i0 = (int)blockIdx.z * chunk, i1 = min(total, i0 + chunk);
  if ((row_stride & 3) == 0 && ((reinterpret_cast<uintptr_t>(dvrow) | reinterpret_cast<uintptr_t>(d_logits)) & 15) == 0) {
    for (int i = i0 + 4 * (int)threadIdx.x; i < i1; i += 4 * 256) {      // total % 4 == 0: whole float4 groups
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < nrows; ++k) {
        const float wk = wtab[k];
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + (size_t)ytab[k] * row_stride + i));
        acc.x = fmaf(wk, v.x, acc.x); acc.y = fmaf(wk, v.y, acc.y); acc.z = fmaf(wk, v.z, acc.z); acc.w = fmaf(wk, v.w, acc.w);
      }
      float4* d4 = reinterpret_cast<float4*>(dst + i);
      float4 o = *d4;
      o.x += acc.x; o.y += acc.y; o.z += acc.z; o.w += acc.w;
      *d4 = o;
    }
  } else {
    for (int i = i0 + (int)threadIdx.x; i < i1; i += 256) {
      float acc = 0.f;
      for (int k = 0; k < nrows; ++k) acc = fmaf(wtab[k], __ldg(src + (size_t)ytab[k] * row_stride + i), acc);
      dst[i] += acc;
    }
  }
}

// pixels of one x-block that can touch a given low-res column: 4 taps x ceil(R / s) pixels per source cell (+ slack), at most XB
static int mean_row_bwd_xbw(int s, int R, int XB) {
  const int wmax = 4 * ((R + s - 1) / s) + 8;
  return wmax < XB ? wmax : XB;
}

// shared-memory plan of the backward row kernel: pixels per x-block (0 = does not fit) and bytes
static int mean_row_bwd_plan(int s, int N, int R, int* NV_out, int* P_out, int* TS_out, size_t* bytes_out) {
  int Np4 = (N + 3) & ~3;
  int NV = ((Np4 >> 2) & 1) ? Np4 : Np4 + 4;
  int XB = R;
  if ((size_t)R * N * sizeof(float) > 100 * 1024) {
    XB = (int)((96 * 1024) / ((size_t)N * sizeof(float)));
    XB = XB >= 32 ? (XB / 32) * 32 : (XB / 4) * 4;
  }
  for (; XB >= 4; XB = (XB > 32 ? 32 : XB / 2)) {
    int P = ((XB + 31) / 32) * 32;
    if (P > 256) P = 256;
    const int TS = row_token_slices(N, P);
    size_t floats = (size_t)XB * ((N + 3) & ~3) + 2 * (size_t)(s + 4) * NV + 32 + 2 * (size_t)TS * P + 5 * (size_t)XB +
                    2 * (size_t)(s + 4) + (size_t)(s + 4) * mean_row_bwd_xbw(s, R, XB);
    if (floats * sizeof(float) <= 200 * 1024) {
      *NV_out = NV; *P_out = P; *TS_out = TS; *bytes_out = floats * sizeof(float);
      return XB;
    }
  }
  return 0;
}

bool capture_mean_row_bwd_fits(int s, int N, int R) {
  int NV, P, TS;
  size_t bytes;
  return mean_row_bwd_plan(s, N, R, &NV, &P, &TS, &bytes) > 0 && 4 * ((R + s - 1) / s) + 8 <= 96;
}

// floats of the dV workspace of one layer: [heads, R, s, N]
size_t capture_mean_row_bwd_workspace(int heads, int s, int N, int R) { return (size_t)heads * R * s * N; }

// d_logits is accumulated into (single writer per element).  workspace: capture_mean_row_bwd_workspace floats.
// *handled = false: shape not taken.
int capture_mean_row_bwd(const float* logits, const float* d_maps, float* d_logits, float* workspace, int heads, int s, int N, int R,
                         float w, cudaStream_t st, bool* handled) {
  *handled = false;
  int NV, P, TS;
  size_t bytes;
  const int XB = mean_row_bwd_plan(s, N, R, &NV, &P, &TS, &bytes);
  // every low-res row is touched by <= 4 * ceil(R / s) + a few output rows (more only when down-sampling, R < s)
  if (XB == 0 || workspace == nullptr || 4 * ((R + s - 1) / s) + 8 > VG_MAXW) return SKP_OK;
  static size_t configured = 0;
  if (bytes > configured) {
    cudaError_t e = cudaFuncSetAttribute(capture_mean_row_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      set_error("capture_mean_row_bwd: smem attr: %s", cudaGetErrorString(e));
      return SKP_ERR_LAUNCH;
    }
    cudaFuncSetAttribute(capture_mean_row_bwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    configured = bytes;
  }
  dim3 grid(R, heads);
  capture_mean_row_bwd_kernel<<<grid, P * TS, bytes, st>>>(logits, d_maps, workspace, s, N, R, NV, P, TS, XB, mean_row_bwd_xbw(s, R, XB), w);
  SKP_CHECK_LAUNCH("capture_mean_row_bwd");
  const int vz = (s * N + 2047) / 2048;                                  // ~2K elements per CTA
  capture_mean_vgather_kernel<<<dim3(s, heads, vz < 1 ? 1 : vz), 256, 0, st>>>(workspace, d_logits, s, N, R);
  SKP_CHECK_LAUNCH("capture_mean_vgather");
  *handled = true;
  return SKP_OK;
}

// Returns SKP_OK with *handled = true when the row kernel ran; *handled = false when the shape does not fit it.
int capture_store_row(const float* logits, float* probs, int heads, int s, int N, int R, cudaStream_t st, bool* handled) {
  *handled = false;
  if (((size_t)R * N) % 4 != 0 || (reinterpret_cast<uintptr_t>(probs) & 15) != 0) return SKP_OK;   // 16-byte rows for the bulk copy
  int Np4 = (N + 3) & ~3;
  int NV = ((Np4 >> 2) & 1) ? Np4 : Np4 + 4;   // NV/4 odd: distinct columns land in distinct bank groups
  // pixels per CTA: the whole row when it fits ~100 KB of staging (4 CTAs per SM at N = 77), else the largest multiple of
  // 32 (>= 4) pixels that does -- each x-block is still one contiguous, 16-byte aligned chunk of the row
  int XB = R;
  if ((size_t)R * N * sizeof(float) > 100 * 1024) {
    XB = (int)((96 * 1024) / ((size_t)N * sizeof(float)));
    XB = XB >= 32 ? (XB / 32) * 32 : (XB / 4) * 4;
    if (XB < 4) return SKP_OK;
  }
  if (((size_t)XB * N) % 4 != 0) return SKP_OK;
  int P = ((XB + 31) / 32) * 32;               // pixel lanes (whole warps)
  if (P > 256) P = 256;
  int TS = row_token_slices(N, P);             // token slices per pixel: more warps to hide the LDS->FMA->EX2->STS chain
  static const int ts_env = getenv("SKP_ROW_TS") ? atoi(getenv("SKP_ROW_TS")) : 0;   // tuning knob (1, 2 or 4)
  if (ts_env > 0 && N >= 8 * ts_env && P * ts_env <= ROW_MAX_THREADS) TS = ts_env;
  size_t floats = (((size_t)XB * N + 3) & ~(size_t)3) + (size_t)(s + 4) * NV + 32 + (size_t)TS * P;
  size_t bytes = floats * sizeof(float);
  if (bytes > 200 * 1024) return SKP_OK;
  static size_t configured[2] = {0, 0};
  const bool one = XB >= R;
  auto kern = one ? capture_store_row_kernel<true> : capture_store_row_kernel<false>;
  if (bytes > configured[one]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      set_error("capture_store_row: smem attr: %s", cudaGetErrorString(e));
      return SKP_ERR_LAUNCH;
    }
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    configured[one] = bytes;
  }
  dim3 grid((R + ROWS_PER_CTA - 1) / ROWS_PER_CTA, heads);
  kern<<<grid, P * TS, bytes, st>>>(logits, probs, s, N, R, NV, P, TS, XB);
  SKP_CHECK_LAUNCH("capture_store_row");
  *handled = true;
  return SKP_OK;
}

}  // namespace skp
