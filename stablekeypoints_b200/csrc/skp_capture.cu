// Attention-store ("capture") kernels: ptp_utils.py:508-538 + optimize.py:50-75.
//
// The reference upsamples the layer input bicubically to R x R, re-projects it with to_q and takes
// softmax_tokens(q' k^T scale).  to_q has no bias and bicubic resampling is linear, so
// q' k^T == bicubic_pixels(q k^T): we resample the layer's own low-res scaled logits [h, s, s, N]
// (L2-resident, <= 2.5 MB) and never form the [R*R, C] activations or the second projection.
//
// One CTA owns a tile of TY x 16 output pixels.  Per (layer, head) it stages the low-res footprint of the
// tile in shared memory (<= 40 slots x N floats), then each thread = (pixel, token-slice) evaluates the
// 4x4 bicubic stencil with 128-bit shared loads (4 tokens per load, conflict-free because the slot
// stride is an odd number of float4), does an online softmax over its token slice, and combines the
// slices through shared memory.  A second sweep re-evaluates the stencil and either
//   STORE: writes probs[h, pix, :] through a [pix][N] staging tile so the 40 MB/layer store is
//          coalesced 128 B lines (HBM-bound: the "attn-store" kernel), or
//   MEAN : accumulates the 1/(layers*heads) mean in shared memory and writes maps[N, R, R] once.
// Backward recomputes the probabilities, forms dS' = p (g - <p, g>) per pixel, and applies the
// transposed stencil separably (x then y) inside the tile before one atomicAdd per footprint slot.
#include "skp_common.cuh"
#include <math_constants.h>

namespace skp {

constexpr int CAP_TX = 16;
constexpr int CAP_THREADS = 256;
constexpr int CAP_MAX_W = 48;  // max footprint extent per axis we stage (covers 2x down-sampling of a 16-px tile)

struct CapParams {
  const float* logits[SKP_MAX_LAYERS];
  float* dlogits[SKP_MAX_LAYERS];
  int s[SKP_MAX_LAYERS];
  int n_layers;
  int heads, N, R;
  int Nf;       // footprint row stride (floats), multiple of 4, Nf/4 odd
  int Nb;       // [pix][Nb] tile stride, odd
  int max_slots;
  int mwy, mwx;  // max footprint extent per axis over all tiles/layers (table strides)
  const float* g;   // MEAN bwd: d_maps [N,R,R]; STORE bwd: d_probs [h,R*R,N]
  float* out;       // MEAN fwd: maps [N,R,R]; STORE fwd: probs [h,R*R,N]
  float w;          // 1/(layers*heads) for MEAN
};

struct Taps {
  float wy[4], wx[4];
  int off[4][4];  // float4 index of (slot, token 0) for tap (t,u)
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Window of low-res rows touched by output rows [d0, d0+n) (clamped taps), align_corners=False bicubic.
__device__ __host__ inline void window(int d0, int n, int R, int s, int* w0, int* w1) {
  float scale = (float)s / (float)R;
  int last = d0 + n - 1;
  if (last > R - 1) last = R - 1;
  float a = scale * (d0 + 0.5f) - 0.5f, b = scale * (last + 0.5f) - 0.5f;
  int lo = (int)floorf(a) - 1, hi = (int)floorf(b) + 2;
  lo = lo < 0 ? 0 : (lo > s - 1 ? s - 1 : lo);
  hi = hi < 0 ? 0 : (hi > s - 1 ? s - 1 : hi);
  *w0 = lo; *w1 = hi;
}

__device__ __forceinline__ void make_taps(Taps& tp, int Y, int X, int R, int s, int wy0, int wx0, int wsx, int Nf4) {
  float scale = (float)s / (float)R;
  float ry = scale * (Y + 0.5f) - 0.5f, rx = scale * (X + 0.5f) - 0.5f;
  float fy = floorf(ry), fx = floorf(rx);
  cubic_coeffs(ry - fy, tp.wy);
  cubic_coeffs(rx - fx, tp.wx);
  int iy = (int)fy, ix = (int)fx;
  int cy[4], cx[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    cy[t] = clampi(iy - 1 + t, 0, s - 1) - wy0;
    cx[t] = clampi(ix - 1 + t, 0, s - 1) - wx0;
  }
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int u = 0; u < 4; ++u) tp.off[t][u] = (cy[t] * wsx + cx[u]) * Nf4;
}

__device__ __forceinline__ float4 stencil4(const float4* __restrict__ fp4, const Taps& tp, int n4) {
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float4 v = fp4[tp.off[t][u] + n4];
      float w = tp.wx[u];
      r.x = fmaf(w, v.x, r.x); r.y = fmaf(w, v.y, r.y); r.z = fmaf(w, v.z, r.z); r.w = fmaf(w, v.w, r.w);
    }
    float w = tp.wy[t];
    a.x = fmaf(w, r.x, a.x); a.y = fmaf(w, r.y, a.y); a.z = fmaf(w, r.z, a.z); a.w = fmaf(w, r.w, a.w);
  }
  return a;
}

// Stage the footprint of (layer l, head h) for this tile: fp[slot][Nf], pad tokens = 0.
__device__ __forceinline__ void load_footprint(float* fp, const float* __restrict__ lg, int h, int s, int N, int Nf,
                                               int wy0, int wx0, int wsy, int wsx) {
  int rowlen = wsx * Nf;
  for (int sy = 0; sy < wsy; ++sy) {
    const float* src = lg + ((size_t)(h * s + wy0 + sy) * s + wx0) * N;
    float* dst = fp + sy * rowlen;
    for (int i = threadIdx.x; i < rowlen; i += CAP_THREADS) {
      int sx = i / Nf, n = i - sx * Nf;
      dst[i] = (n < N) ? __ldg(src + (size_t)sx * N + n) : 0.f;
    }
  }
}

template <int TY, bool STORE>
__global__ void __launch_bounds__(CAP_THREADS) capture_fwd_kernel(CapParams p) {
  constexpr int TP = TY * CAP_TX;
  constexpr int NQ = CAP_THREADS / TP;
  extern __shared__ __align__(16) float smem[];
  const int N = p.N, R = p.R, Nf = p.Nf, Nb = p.Nb, Nf4 = Nf >> 2;
  float* fp = smem;                              // [max_slots][Nf]
  float* tile = fp + (size_t)p.max_slots * Nf;   // [TP][Nb]  (MEAN: accumulator; STORE: probs staging)
  float* st_m = tile + (size_t)TP * Nb;          // [NQ][TP]
  float* st_s = st_m + NQ * TP;                  // [NQ][TP]

  const int pix = threadIdx.x % TP, q = threadIdx.x / TP;
  const int py = pix / CAP_TX, px = pix % CAP_TX;
  const int Y0 = blockIdx.y * TY, X0 = blockIdx.x * CAP_TX;
  const int Y = Y0 + py, X = X0 + px;
  const bool valid = (Y < R) && (X < R);
  const int Yc = valid ? Y : (Y < R ? Y : R - 1), Xc = X < R ? X : R - 1;
  int ch = (N + NQ - 1) / NQ;
  ch = (ch + 3) & ~3;
  const int n0 = q * ch, n1 = min(N, n0 + ch);

  if (!STORE)
    for (int i = threadIdx.x; i < TP * Nb; i += CAP_THREADS) tile[i] = 0.f;

  const int h_begin = STORE ? blockIdx.z : 0, h_end = STORE ? blockIdx.z + 1 : p.heads;
  for (int l = 0; l < p.n_layers; ++l) {
    const int s = p.s[l];
    int wy0, wy1, wx0, wx1;
    window(Y0, TY, R, s, &wy0, &wy1);
    window(X0, CAP_TX, R, s, &wx0, &wx1);
    const int wsy = wy1 - wy0 + 1, wsx = wx1 - wx0 + 1;
    Taps tp;
    make_taps(tp, Yc, Xc, R, s, wy0, wx0, wsx, Nf4);
    for (int h = h_begin; h < h_end; ++h) {
      __syncthreads();  // previous footprint / stats fully consumed
      load_footprint(fp, p.logits[l], h, s, N, Nf, wy0, wx0, wsy, wsx);
      __syncthreads();
      const float4* fp4 = reinterpret_cast<const float4*>(fp);
      // sweep 1: online softmax statistics over this thread's token slice
      float m = -CUDART_INF_F, ssum = 0.f;
      for (int n = n0; n < n1; n += 4) {
        float4 v4 = stencil4(fp4, tp, n >> 2);
        float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (n + c < n1) {
            if (v[c] > m) { ssum *= __expf(m - v[c]); m = v[c]; }
            ssum += __expf(v[c] - m);
          }
      }
      st_m[q * TP + pix] = m;
      st_s[q * TP + pix] = ssum;
      __syncthreads();
      float M = -CUDART_INF_F;
#pragma unroll
      for (int k = 0; k < NQ; ++k) M = fmaxf(M, st_m[k * TP + pix]);
      float Ssum = 0.f;
#pragma unroll
      for (int k = 0; k < NQ; ++k) Ssum += st_s[k * TP + pix] * __expf(st_m[k * TP + pix] - M);
      const float inv = 1.f / Ssum;
      // sweep 2: probabilities
      float* trow = tile + (size_t)pix * Nb;
      for (int n = n0; n < n1; n += 4) {
        float4 v4 = stencil4(fp4, tp, n >> 2);
        float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (n + c < n1) {
            float pr = __expf(v[c] - M) * inv;
            if (STORE) trow[n + c] = pr;
            else trow[n + c] += p.w * pr;
          }
      }
      if (STORE) {
        __syncthreads();
        // coalesced copy-out: each tile row is a contiguous run of (valid pixels)*N floats
        for (int ry = 0; ry < TY; ++ry) {
          int Yr = Y0 + ry;
          if (Yr >= R) break;
          int npx = min(CAP_TX, R - X0);
          float* dst = p.out + ((size_t)h * R * R + (size_t)Yr * R + X0) * N;
          int total = npx * N;
          for (int i = threadIdx.x; i < total; i += CAP_THREADS) {
            int pxx = i / N, n = i - pxx * N;
            dst[i] = tile[(size_t)(ry * CAP_TX + pxx) * Nb + n];
          }
        }
      }
    }
  }
  if (!STORE) {
    __syncthreads();
    for (int i = threadIdx.x; i < N * TP; i += CAP_THREADS) {
      int n = i / TP, pp = i - n * TP;
      int yy = Y0 + pp / CAP_TX, xx = X0 + pp % CAP_TX;
      if (yy < R && xx < R) p.out[((size_t)n * R + yy) * R + xx] = tile[(size_t)pp * Nb + n];
    }
  }
}

template <int TY, bool STORE>
__global__ void __launch_bounds__(CAP_THREADS) capture_bwd_kernel(CapParams p) {
  constexpr int TP = TY * CAP_TX;
  constexpr int NQ = CAP_THREADS / TP;
  extern __shared__ __align__(16) float smem[];
  const int N = p.N, R = p.R, Nf = p.Nf, Nb = p.Nb, Nf4 = Nf >> 2;
  float* fp = smem;                                // [max_slots][Nf]
  float* gt = fp + (size_t)p.max_slots * Nf;       // [TP][Nb] upstream gradient g
  float* ds = gt + (size_t)TP * Nb;                // [TP][Nb] probs, then dS'
  const int MWY = p.mwy, MWX = p.mwx;
  float* tmpx = ds + (size_t)TP * Nb;              // [TY][MWX][Nb]
  float* st_m = tmpx + (size_t)TY * MWX * Nb;      // [NQ][TP]
  float* st_s = st_m + NQ * TP;                    // [NQ][TP]
  float* Wy = st_s + NQ * TP;                      // [TY][MWY]
  float* Wx = Wy + TY * MWY;                       // [CAP_TX][MWX]

  const int pix = threadIdx.x % TP, q = threadIdx.x / TP;
  const int py = pix / CAP_TX, px = pix % CAP_TX;
  const int Y0 = blockIdx.y * TY, X0 = blockIdx.x * CAP_TX;
  const int Y = Y0 + py, X = X0 + px;
  const bool valid = (Y < R) && (X < R);
  const int Yc = Y < R ? Y : R - 1, Xc = X < R ? X : R - 1;
  int ch = (N + NQ - 1) / NQ;
  ch = (ch + 3) & ~3;
  const int n0 = q * ch, n1 = min(N, n0 + ch);

  if (!STORE) {
    // g[pix][n] = w * d_maps[n, Y, X]  (shared by every layer/head)
    for (int i = threadIdx.x; i < N * TP; i += CAP_THREADS) {
      int n = i / TP, pp = i - n * TP;
      int yy = Y0 + pp / CAP_TX, xx = X0 + pp % CAP_TX;
      gt[(size_t)pp * Nb + n] = (yy < R && xx < R) ? p.w * __ldg(p.g + ((size_t)n * R + yy) * R + xx) : 0.f;
    }
  }
  const int h_begin = STORE ? blockIdx.z : 0, h_end = STORE ? blockIdx.z + 1 : p.heads;
  for (int l = 0; l < p.n_layers; ++l) {
    const int s = p.s[l];
    int wy0, wy1, wx0, wx1;
    window(Y0, TY, R, s, &wy0, &wy1);
    window(X0, CAP_TX, R, s, &wx0, &wx1);
    const int wsy = wy1 - wy0 + 1, wsx = wx1 - wx0 + 1;
    Taps tp;
    make_taps(tp, Yc, Xc, R, s, wy0, wx0, wsx, Nf4);
    __syncthreads();
    // dense separable transposed-stencil tables for this tile/layer (clamped duplicate taps add up)
    for (int i = threadIdx.x; i < TY * MWY + CAP_TX * MWX; i += CAP_THREADS) Wy[i] = 0.f;
    __syncthreads();
    if (threadIdx.x < TY + CAP_TX) {
      bool isy = threadIdx.x < TY;
      int r = isy ? threadIdx.x : threadIdx.x - TY;
      int d = (isy ? Y0 : X0) + r;
      if (d < R) {
        float scale = (float)s / (float)R;
        float rc = scale * (d + 0.5f) - 0.5f, f = floorf(rc);
        float cw[4];
        cubic_coeffs(rc - f, cw);
        int w0 = isy ? wy0 : wx0;
        float* row = isy ? (Wy + r * MWY) : (Wx + r * MWX);
        for (int t = 0; t < 4; ++t) row[clampi((int)f - 1 + t, 0, s - 1) - w0] += cw[t];
      }
    }
    for (int h = h_begin; h < h_end; ++h) {
      __syncthreads();
      load_footprint(fp, p.logits[l], h, s, N, Nf, wy0, wx0, wsy, wsx);
      if (STORE) {
        for (int ry = 0; ry < TY; ++ry) {
          int Yr = Y0 + ry;
          int npx = (Yr < R) ? min(CAP_TX, R - X0) : 0;
          const float* src = p.g + ((size_t)h * R * R + (size_t)(Yr < R ? Yr : 0) * R + X0) * N;
          for (int i = threadIdx.x; i < CAP_TX * N; i += CAP_THREADS) {
            int pxx = i / N, n = i - pxx * N;
            gt[(size_t)(ry * CAP_TX + pxx) * Nb + n] = (pxx < npx) ? __ldg(src + i) : 0.f;
          }
        }
      }
      __syncthreads();
      const float4* fp4 = reinterpret_cast<const float4*>(fp);
      float m = -CUDART_INF_F, ssum = 0.f;
      for (int n = n0; n < n1; n += 4) {
        float4 v4 = stencil4(fp4, tp, n >> 2);
        float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (n + c < n1) {
            if (v[c] > m) { ssum *= __expf(m - v[c]); m = v[c]; }
            ssum += __expf(v[c] - m);
          }
      }
      st_m[q * TP + pix] = m;
      st_s[q * TP + pix] = ssum;
      __syncthreads();
      float M = -CUDART_INF_F;
#pragma unroll
      for (int k = 0; k < NQ; ++k) M = fmaxf(M, st_m[k * TP + pix]);
      float Ssum = 0.f;
#pragma unroll
      for (int k = 0; k < NQ; ++k) Ssum += st_s[k * TP + pix] * __expf(st_m[k * TP + pix] - M);
      const float inv = 1.f / Ssum;
      __syncthreads();  // stats consumed before they are reused for the dot products
      float* drow = ds + (size_t)pix * Nb;
      const float* grow = gt + (size_t)pix * Nb;
      float dot = 0.f;
      for (int n = n0; n < n1; n += 4) {
        float4 v4 = stencil4(fp4, tp, n >> 2);
        float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (n + c < n1) {
            float pr = __expf(v[c] - M) * inv;
            drow[n + c] = pr;
            dot = fmaf(pr, grow[n + c], dot);
          }
      }
      st_s[q * TP + pix] = dot;
      __syncthreads();
      dot = 0.f;
#pragma unroll
      for (int k = 0; k < NQ; ++k) dot += st_s[k * TP + pix];
      for (int n = n0; n < n1; ++n) drow[n] = valid ? drow[n] * (grow[n] - dot) : 0.f;
      __syncthreads();
      // transposed stencil, x first: tmpx[py][sx][n] = sum_px Wx[px][sx] * ds[py,px][n]
      for (int i = threadIdx.x; i < TY * wsx * N; i += CAP_THREADS) {
        int n = i % N, r = i / N;
        int sx = r % wsx, yy = r / wsx;
        float a = 0.f;
#pragma unroll
        for (int xx = 0; xx < CAP_TX; ++xx) a = fmaf(Wx[xx * MWX + sx], ds[(size_t)(yy * CAP_TX + xx) * Nb + n], a);
        tmpx[(size_t)(yy * MWX + sx) * Nb + n] = a;
      }
      __syncthreads();
      float* dl = p.dlogits[l];
      for (int i = threadIdx.x; i < wsy * wsx * N; i += CAP_THREADS) {
        int n = i % N, r = i / N;
        int sx = r % wsx, sy = r / wsx;
        float a = 0.f;
#pragma unroll
        for (int yy = 0; yy < TY; ++yy) a = fmaf(Wy[yy * MWY + sy], tmpx[(size_t)(yy * MWX + sx) * Nb + n], a);
        if (a != 0.f) atomicAdd(dl + ((size_t)(h * s + wy0 + sy) * s + wx0 + sx) * N + n, a);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ host
static int footprint_slots(int R, int s, int TY, int* ok, int* mwy, int* mwx) {
  int best = 0;
  *ok = 1;
  for (int Y0 = 0; Y0 < R; Y0 += TY) {
    int a, b;
    window(Y0, TY, R, s, &a, &b);
    int wy = b - a + 1;
    if (wy > CAP_MAX_W) *ok = 0;
    if (wy > *mwy) *mwy = wy;
    for (int X0 = 0; X0 < R; X0 += CAP_TX) {
      int c, d;
      window(X0, CAP_TX, R, s, &c, &d);
      int wx = d - c + 1;
      if (wx > CAP_MAX_W) *ok = 0;
      if (wx > *mwx) *mwx = wx;
      if (wy * wx > best) best = wy * wx;
    }
  }
  return best;
}

template <int TY>
static size_t smem_bytes(const CapParams& p, bool bwd) {
  constexpr int TP = TY * CAP_TX, NQ = CAP_THREADS / TP;
  size_t f = (size_t)p.max_slots * p.Nf + (size_t)TP * p.Nb + 2 * NQ * TP;
  if (bwd) f += (size_t)TP * p.Nb + (size_t)TY * p.mwx * p.Nb + (size_t)TY * p.mwy + (size_t)CAP_TX * p.mwx;
  return f * sizeof(float);
}

template <int TY, bool STORE, bool BWD>
static int launch_ty(CapParams& p, cudaStream_t st, bool* fits) {
  int slots = 0;
  p.mwy = p.mwx = 1;
  for (int l = 0; l < p.n_layers; ++l) {
    int ok;
    int v = footprint_slots(p.R, p.s[l], TY, &ok, &p.mwy, &p.mwx);
    if (!ok) { *fits = false; return SKP_OK; }
    if (v > slots) slots = v;
  }
  p.max_slots = slots;
  size_t bytes = smem_bytes<TY>(p, BWD);
  if (bytes > 200 * 1024) { *fits = false; return SKP_OK; }
  *fits = true;
  dim3 grid((p.R + CAP_TX - 1) / CAP_TX, (p.R + TY - 1) / TY, STORE ? p.heads : 1);
  cudaError_t e;
  if (BWD) {
    e = cudaFuncSetAttribute(capture_bwd_kernel<TY, STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error("capture: smem attr: %s", cudaGetErrorString(e)); return SKP_ERR_LAUNCH; }
    capture_bwd_kernel<TY, STORE><<<grid, CAP_THREADS, bytes, st>>>(p);
  } else {
    e = cudaFuncSetAttribute(capture_fwd_kernel<TY, STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error("capture: smem attr: %s", cudaGetErrorString(e)); return SKP_ERR_LAUNCH; }
    capture_fwd_kernel<TY, STORE><<<grid, CAP_THREADS, bytes, st>>>(p);
  }
  SKP_CHECK_LAUNCH("capture");
  return SKP_OK;
}

// skp_capture_row.cu: the row-per-CTA attn-store kernel (any R, s; needs the row + footprint to fit shared memory)
int capture_store_row(const float* logits, float* probs, int heads, int s, int N, int R, cudaStream_t st, bool* handled);
// skp_capture_store.cu: the register formulation of the same kernel (exponentials never re-read from shared memory)
int capture_store_reg(const float* logits, float* probs, int heads, int s, int N, int R, cudaStream_t st, bool* handled);
int capture_mean_row_bwd(const float* logits, const float* d_maps, float* d_logits, float* workspace, int heads, int s, int N, int R,
                         float w, cudaStream_t st, bool* handled);
bool capture_mean_row_bwd_fits(int s, int N, int R);
size_t capture_mean_row_bwd_workspace(int heads, int s, int N, int R);
// skp_capture_tc.cu: the tcgen05 formulation (horizontal bicubic pass as a GEMM, softmax thread-local on the TMEM lanes)
int capture_tc(const float* const* logits, const int* s, int n_layers, float* out, int heads, int N, int R, bool store,
               float* workspace, cudaStream_t st, bool* handled);
bool capture_tc_eligible(const int* s, int n_layers, int N, int R, bool store);
size_t capture_tc_workspace(const int* s, int n_layers, int heads);
void capture_tc_enable(int on);
void capture_tc_debug(long long* buf);

// kernel selection switches (tests / A-B measurements): initial value from the environment, skp_capture_select() at run time
// forward store: 2 = register kernel (skp_capture_store.cu), 1 = row kernel (skp_capture_row.cu), 0 = tile kernel
static int g_row_fwd = getenv("SKP_CAPTURE_ROW") != nullptr ? atoi(getenv("SKP_CAPTURE_ROW")) : 2;
static int g_row_bwd = (getenv("SKP_CAPTURE_BWD_ROW") != nullptr && atoi(getenv("SKP_CAPTURE_BWD_ROW")) == 0) ? 0 : 1;

template <bool STORE, bool BWD>
static int launch(CapParams& p, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (STORE && !BWD) {
    if (g_row_fwd >= 2) {
      bool handled = false;
      int rr = capture_store_reg(p.logits[0], p.out, p.heads, p.s[0], p.N, p.R, st, &handled);
      if (rr != SKP_OK || handled) return rr;
    }
    if (g_row_fwd) {
      bool handled = false;
      int rr = capture_store_row(p.logits[0], p.out, p.heads, p.s[0], p.N, p.R, st, &handled);
      if (rr != SKP_OK || handled) return rr;
    }
  }
  int Np4 = (p.N + 3) & ~3;
  p.Nf = ((Np4 >> 2) & 1) ? Np4 : Np4 + 4;
  p.Nb = p.N | 1;
  bool fits = false;
  int rc = launch_ty<4, STORE, BWD>(p, st, &fits);
  if (rc != SKP_OK || fits) return rc;
  rc = launch_ty<2, STORE, BWD>(p, st, &fits);
  if (rc != SKP_OK || fits) return rc;
  rc = launch_ty<1, STORE, BWD>(p, st, &fits);
  if (rc != SKP_OK || fits) return rc;
  set_error("capture: tile does not fit shared memory (N=%d R=%d)", p.N, p.R);
  return SKP_ERR_UNSUPPORTED;
}

static int fill(CapParams& p, const float* const* logits, float* const* dlogits, const int* s, int n_layers, int heads,
                int N, int R) {
  SKP_REQUIRE(n_layers >= 1 && n_layers <= SKP_MAX_LAYERS, "capture: n_layers=%d out of range", n_layers);
  SKP_REQUIRE(heads > 0 && N > 0 && R > 0, "capture: bad sizes heads=%d N=%d R=%d", heads, N, R);
  p.n_layers = n_layers; p.heads = heads; p.N = N; p.R = R;
  for (int l = 0; l < n_layers; ++l) {
    SKP_REQUIRE(logits[l] != nullptr && s[l] > 0, "capture: layer %d null/empty", l);
    p.logits[l] = logits[l];
    p.dlogits[l] = dlogits ? dlogits[l] : nullptr;
    p.s[l] = s[l];
  }
  return SKP_OK;
}

}  // namespace skp

using namespace skp;

extern "C" int skp_capture_store_fwd(const float* logits, float* probs, int heads, int s, int N, int R, void* stream);
extern "C" int skp_capture_mean_fwd(const float* const* logits, const int* s, int n_layers, float* maps, int heads, int N,
                                    int R, void* stream);

extern "C" void skp_capture_tc(int on) { capture_tc_enable(on); }
extern "C" void skp_capture_tc_trace(void* buf) { capture_tc_debug(reinterpret_cast<long long*>(buf)); }

/* 1 when skp_capture_store_fwd (store != 0) / skp_capture_mean_fwd (store == 0) would run the tcgen05 kernel for this shape */
extern "C" int skp_capture_tc_ok(const int* s, int n_layers, int N, int R, int store) {
  return s != nullptr && capture_tc_eligible(s, n_layers, N, R, store != 0) ? 1 : 0;
}

extern "C" int64_t skp_capture_tc_workspace(const int* s, int n_layers, int heads) {
  if (s == nullptr || n_layers < 1 || n_layers > SKP_MAX_LAYERS || heads < 1) return 0;
  return (int64_t)capture_tc_workspace(s, n_layers, heads);
}

extern "C" int skp_capture_store_tc_fwd(const float* logits, float* probs, int heads, int s, int N, int R, float* workspace,
                                        void* stream) {
  SKP_REQUIRE(logits != nullptr && probs != nullptr, "capture_store_tc_fwd: null pointer");
  bool handled = false;
  int rc = capture_tc(&logits, &s, 1, probs, heads, N, R, true, workspace, (cudaStream_t)stream, &handled);
  if (rc != SKP_OK || handled) return rc;
  return skp_capture_store_fwd(logits, probs, heads, s, N, R, stream);   // shapes the tensor-core kernel does not take
}

extern "C" int skp_capture_mean_tc_fwd(const float* const* logits, const int* s, int n_layers, float* maps, int heads, int N,
                                       int R, float* workspace, void* stream) {
  SKP_REQUIRE(logits != nullptr && s != nullptr && maps != nullptr, "capture_mean_tc_fwd: null pointer");
  bool handled = false;
  int rc = capture_tc(logits, s, n_layers, maps, heads, N, R, false, workspace, (cudaStream_t)stream, &handled);
  if (rc != SKP_OK || handled) return rc;
  return skp_capture_mean_fwd(logits, s, n_layers, maps, heads, N, R, stream);
}

extern "C" void skp_capture_select(int row_fwd, int row_bwd) {
  if (row_fwd >= 0) g_row_fwd = row_fwd;
  if (row_bwd >= 0) g_row_bwd = row_bwd != 0;
}

extern "C" int skp_capture_store_fwd(const float* logits, float* probs, int heads, int s, int N, int R, void* stream) {
  SKP_REQUIRE(probs != nullptr, "capture_store_fwd: null output");
  CapParams p{};
  int rc = fill(p, &logits, nullptr, &s, 1, heads, N, R);
  if (rc) return rc;
  p.out = probs; p.w = 1.f;
  return launch<true, false>(p, stream);
}

extern "C" int skp_capture_store_bwd(const float* logits, const float* d_probs, float* d_logits, int heads, int s,
                                     int N, int R, void* stream) {
  SKP_REQUIRE(d_probs != nullptr && d_logits != nullptr, "capture_store_bwd: null pointer");
  CapParams p{};
  float* dl = d_logits;
  int rc = fill(p, &logits, &dl, &s, 1, heads, N, R);
  if (rc) return rc;
  p.g = d_probs; p.w = 1.f;
  return launch<true, true>(p, stream);
}

extern "C" int skp_capture_mean_fwd(const float* const* logits, const int* s, int n_layers, float* maps, int heads,
                                    int N, int R, void* stream) {
  SKP_REQUIRE(logits != nullptr && s != nullptr && maps != nullptr, "capture_mean_fwd: null pointer");
  CapParams p{};
  int rc = fill(p, logits, nullptr, s, n_layers, heads, N, R);
  if (rc) return rc;
  p.out = maps; p.w = 1.f / (float)(n_layers * heads);
  return launch<false, false>(p, stream);
}

extern "C" int64_t skp_capture_mean_bwd_workspace(const int* s, int n_layers, int heads, int N, int R) {
  if (s == nullptr || n_layers < 1 || n_layers > SKP_MAX_LAYERS || heads < 1 || N < 1 || R < 1) return 0;
  size_t fl = 0;                       // the layers run one after the other on the stream: they share the buffer
  for (int l = 0; l < n_layers; ++l) {
    const size_t f = capture_mean_row_bwd_workspace(heads, s[l], N, R);
    if (f > fl) fl = f;
  }
  return (int64_t)(fl * sizeof(float));
}

extern "C" int skp_capture_mean_bwd(const float* const* logits, const int* s, int n_layers, const float* d_maps,
                                    float* const* d_logits, int heads, int N, int R, float* workspace, void* stream) {
  SKP_REQUIRE(logits != nullptr && s != nullptr && d_maps != nullptr && d_logits != nullptr,
              "capture_mean_bwd: null pointer");
  {   // row formulation (skp_capture_row.cu): two launches per layer, no atomics; needs the workspace.  Without one (or with
      // skp_capture_select(., 0)) the tile kernel runs
    if (g_row_bwd && workspace != nullptr && n_layers >= 1 && n_layers <= SKP_MAX_LAYERS && heads > 0 && N > 0 && R > 0) {
      bool all = true;
      for (int l = 0; l < n_layers; ++l)   // every layer must fit before any is accumulated
        if (logits[l] == nullptr || d_logits[l] == nullptr || s[l] <= 0 || !capture_mean_row_bwd_fits(s[l], N, R)) all = false;
      for (int l = 0; l < n_layers && all; ++l) {
        bool handled = false;
        int rr = capture_mean_row_bwd(logits[l], d_maps, d_logits[l], workspace, heads, s[l], N, R, 1.f / (float)(n_layers * heads),
                                      (cudaStream_t)stream, &handled);
        if (rr != SKP_OK) return rr;
        if (!handled) {
          SKP_REQUIRE(l == 0, "capture_mean_bwd: layer %d does not fit the row kernel after earlier layers were accumulated", l);
          all = false;
        }
      }
      if (all) return SKP_OK;
    }
  }
  CapParams p{};
  int rc = fill(p, logits, d_logits, s, n_layers, heads, N, R);
  if (rc) return rc;
  for (int l = 0; l < n_layers; ++l) SKP_REQUIRE(d_logits[l] != nullptr, "capture_mean_bwd: d_logits[%d] null", l);
  p.g = d_maps; p.w = 1.f / (float)(n_layers * heads);
  return launch<false, true>(p, stream);
}
