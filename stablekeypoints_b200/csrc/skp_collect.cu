// collect_maps (optimize.py:27-79): mean over (layer, batch*head) of the stored [BH, R*R, N] maps, optional
// token gather, permute to token-major, optional bilinear resize (align_corners=False).
//
// HBM-bound: the forward reads n_layers*BH*R*R*N floats once (161.5 MB at N=77, R=128) and writes T*R2*R2.
// A CTA owns 32 consecutive pixels; for each (layer, bh) slice the 32*N floats it needs are ONE contiguous
// run, read with coalesced 128-bit loads and accumulated in registers; the token-major transpose goes through
// shared memory so the [T,R,R] store is 128 B lines.  Resize-after-mean equals the reference's
// mean-after-resize (both linear) and touches 32x less data.
#include "skp_common.cuh"

namespace skp {

constexpr int COL_PIX = 32;
constexpr int COL_THREADS = 256;
constexpr int COL_MAX_ACC = 16;  // float4 accumulators per thread -> 32*N <= 256*16*4 -> N <= 512

struct ColPtrs {
  const float* in[SKP_MAX_LAYERS];
  float* out[SKP_MAX_LAYERS];
};

template <int NACC>
__global__ void __launch_bounds__(COL_THREADS) collect_mean_kernel(ColPtrs ptrs, int n_layers, int BH, int RR, int N,
                                                                    const int64_t* __restrict__ idx, int T,
                                                                    float* __restrict__ out /* [T, RR] */) {
  extern __shared__ float sm[];  // [32][N|1]
  const int Ns = N | 1;
  const int pix0 = blockIdx.x * COL_PIX;
  const int npix = min(COL_PIX, RR - pix0);
  const int total = npix * N;  // contiguous floats per slice
  const float inv = 1.f / (float)(n_layers * BH);
  const bool vec = ((N & 3) == 0) || (((size_t)pix0 * N) % 4 == 0 && (((size_t)RR * N) % 4 == 0));
  if (vec) {
    float4 acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int total4 = total >> 2;
    for (int l = 0; l < n_layers; ++l)
      for (int b = 0; b < BH; ++b) {
        const float4* src = reinterpret_cast<const float4*>(ptrs.in[l] + ((size_t)b * RR + pix0) * N);
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
          int e = threadIdx.x + i * COL_THREADS;
          if (e < total4) {
            float4 v = __ldcs(src + e);
            acc[i].x += v.x; acc[i].y += v.y; acc[i].z += v.z; acc[i].w += v.w;
          }
        }
      }
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      int e = threadIdx.x + i * COL_THREADS;
      if (e < total4) {
        float v[4] = {acc[i].x, acc[i].y, acc[i].z, acc[i].w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          int f = e * 4 + c;
          int pp = f / N, n = f - pp * N;
          sm[pp * Ns + n] = v[c] * inv;
        }
      }
    }
    // scalar tail (total not a multiple of 4)
    for (int f = (total4 << 2) + threadIdx.x; f < total; f += COL_THREADS) {
      float a = 0.f;
      for (int l = 0; l < n_layers; ++l)
        for (int b = 0; b < BH; ++b) a += __ldcs(ptrs.in[l] + ((size_t)b * RR + pix0) * N + f);
      int pp = f / N, n = f - pp * N;
      sm[pp * Ns + n] = a * inv;
    }
  } else {
    for (int f = threadIdx.x; f < total; f += COL_THREADS) {
      float a = 0.f;
      for (int l = 0; l < n_layers; ++l)
        for (int b = 0; b < BH; ++b) a += __ldcs(ptrs.in[l] + ((size_t)b * RR + pix0) * N + f);
      int pp = f / N, n = f - pp * N;
      sm[pp * Ns + n] = a * inv;
    }
  }
  __syncthreads();
  // token-major write: lanes over pixels
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int t = w; t < T; t += COL_THREADS / 32) {
    int n = idx ? (int)idx[t] : t;
    if (lane < npix) out[(size_t)t * RR + pix0 + lane] = sm[lane * Ns + n];
  }
}

// out[t, y2, x2] = bilinear(in[t], R -> R2), align_corners=False (PyTorch upsample_bilinear2d).
__global__ void bilinear_resize_kernel(const float* __restrict__ in, float* __restrict__ out, int T, int R, int R2) {
  const float scale = (float)R / (float)R2;
  size_t total = (size_t)T * R2 * R2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int x2 = (int)(i % R2);
    int y2 = (int)((i / R2) % R2);
    int t = (int)(i / ((size_t)R2 * R2));
    float sy = src_coord(y2, scale, false), sx = src_coord(x2, scale, false);
    int y0 = min((int)sy, R - 1), x0 = min((int)sx, R - 1);
    int y1 = y0 + (y0 < R - 1), x1 = x0 + (x0 < R - 1);
    float ly = sy - y0, lx = sx - x0;
    const float* m = in + (size_t)t * R * R;
    float v = (1.f - ly) * ((1.f - lx) * m[y0 * R + x0] + lx * m[y0 * R + x1]) +
              ly * ((1.f - lx) * m[y1 * R + x0] + lx * m[y1 * R + x1]);
    out[i] = v;
  }
}

// Transposed bilinear: din[t, y, x] = sum over output pixels whose stencil touches (y, x).  Gather form:
// each input pixel scans the (small) range of output pixels that can reference it.
__global__ void bilinear_resize_bwd_kernel(const float* __restrict__ dout, float* __restrict__ din, int T, int R, int R2) {
  const float scale = (float)R / (float)R2;
  const float inv_scale = (float)R2 / (float)R;
  size_t total = (size_t)T * R * R;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int x = (int)(i % R);
    int y = (int)((i / R) % R);
    int t = (int)(i / ((size_t)R * R));
    // output rows y2 with floor(src) in {y-1, y}: src in [y-1, y+1)  (plus clamping at the borders)
    int ylo = max(0, (int)floorf((y - 1 + 0.5f) * inv_scale - 0.5f) - 1);
    int yhi = min(R2 - 1, (int)ceilf((y + 1 + 0.5f) * inv_scale - 0.5f) + 1);
    int xlo = max(0, (int)floorf((x - 1 + 0.5f) * inv_scale - 0.5f) - 1);
    int xhi = min(R2 - 1, (int)ceilf((x + 1 + 0.5f) * inv_scale - 0.5f) + 1);
    if (y == 0) ylo = 0;
    if (y == R - 1) yhi = R2 - 1;
    if (x == 0) xlo = 0;
    if (x == R - 1) xhi = R2 - 1;
    const float* g = dout + (size_t)t * R2 * R2;
    float a = 0.f;
    for (int y2 = ylo; y2 <= yhi; ++y2) {
      float sy = src_coord(y2, scale, false);
      int y0 = min((int)sy, R - 1), y1 = y0 + (y0 < R - 1);
      float ly = sy - y0;
      float wy = (y0 == y ? 1.f - ly : 0.f) + (y1 == y ? ly : 0.f);
      if (wy == 0.f) continue;
      for (int x2 = xlo; x2 <= xhi; ++x2) {
        float sx = src_coord(x2, scale, false);
        int x0 = min((int)sx, R - 1), x1 = x0 + (x0 < R - 1);
        float lx = sx - x0;
        float wx = (x0 == x ? 1.f - lx : 0.f) + (x1 == x ? lx : 0.f);
        if (wx != 0.f) a = fmaf(wy * wx, g[(size_t)y2 * R2 + x2], a);
      }
    }
    din[i] = a;
  }
}

// d_stored[l][b, pix, n] = (n == idx[t] ? dmean[t, pix] : 0) / (n_layers*BH).  Writes every element.
__global__ void __launch_bounds__(COL_THREADS) collect_bwd_kernel(const float* __restrict__ dmean /* [T, RR] */, ColPtrs ptrs,
                                                                  int n_layers, int BH, int RR, int N,
                                                                  const int64_t* __restrict__ idx, int T) {
  extern __shared__ float sm[];  // [32][N|1]
  const int Ns = N | 1;
  const int pix0 = blockIdx.x * COL_PIX;
  const int npix = min(COL_PIX, RR - pix0);
  const float inv = 1.f / (float)(n_layers * BH);
  for (int i = threadIdx.x; i < COL_PIX * Ns; i += COL_THREADS) sm[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int t = w; t < T; t += COL_THREADS / 32) {
    int n = idx ? (int)idx[t] : t;
    // duplicate indices accumulate (matches autograd of advanced indexing)
    if (lane < npix) atomicAdd(&sm[lane * Ns + n], dmean[(size_t)t * RR + pix0 + lane] * inv);
  }
  __syncthreads();
  const int total = npix * N;
  for (int l = 0; l < n_layers; ++l)
    for (int b = 0; b < BH; ++b) {
      float* dst = ptrs.out[l] + ((size_t)b * RR + pix0) * N;
      for (int f = threadIdx.x; f < total; f += COL_THREADS) {
        int pp = f / N, n = f - pp * N;
        __stcs(dst + f, sm[pp * Ns + n]);
      }
    }
}

}  // namespace skp

using namespace skp;

extern "C" int skp_collect_maps_fwd(const float* const* stored, int n_layers, int BH, int R, int N, const int64_t* idx,
                                    int n_idx, int R2, float* tmp, float* out, void* stream) {
  SKP_REQUIRE(stored && out, "collect_maps_fwd: null pointer");
  SKP_REQUIRE(n_layers >= 1 && n_layers <= SKP_MAX_LAYERS, "collect_maps_fwd: n_layers=%d", n_layers);
  SKP_REQUIRE(BH > 0 && R > 0 && N > 0 && R2 > 0, "collect_maps_fwd: bad sizes");
  SKP_REQUIRE(N <= 512, "collect_maps_fwd: N=%d > 512 unsupported", N);
  SKP_REQUIRE(idx == nullptr || n_idx > 0, "collect_maps_fwd: empty index list");
  cudaStream_t st = (cudaStream_t)stream;
  ColPtrs p{};
  for (int l = 0; l < n_layers; ++l) {
    SKP_REQUIRE(stored[l], "collect_maps_fwd: stored[%d] null", l);
    p.in[l] = stored[l];
  }
  const int T = idx ? n_idx : N, RR = R * R;
  float* mean_dst = (R2 == R) ? out : tmp;
  SKP_REQUIRE(mean_dst, "collect_maps_fwd: tmp workspace required when resizing");
  size_t smem = (size_t)COL_PIX * (N | 1) * sizeof(float);
  int blocks = (RR + COL_PIX - 1) / COL_PIX;
  int nacc = (COL_PIX * N / 4 + COL_THREADS - 1) / COL_THREADS;
#define SKP_COL_LAUNCH(NA)                                                                                     \
  {                                                                                                            \
    cudaFuncSetAttribute(collect_mean_kernel<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
    collect_mean_kernel<NA><<<blocks, COL_THREADS, smem, st>>>(p, n_layers, BH, RR, N, idx, T, mean_dst);      \
  }
  if (nacc <= 1) SKP_COL_LAUNCH(1)
  else if (nacc <= 2) SKP_COL_LAUNCH(2)
  else if (nacc <= 4) SKP_COL_LAUNCH(4)
  else if (nacc <= 8) SKP_COL_LAUNCH(8)
  else SKP_COL_LAUNCH(16)
#undef SKP_COL_LAUNCH
  SKP_CHECK_LAUNCH("collect_mean");
  if (R2 != R) {
    size_t total = (size_t)T * R2 * R2;
    int b2 = (int)((total + 255) / 256);
    if (b2 > 148 * 16) b2 = 148 * 16;
    bilinear_resize_kernel<<<b2, 256, 0, st>>>(tmp, out, T, R, R2);
    SKP_CHECK_LAUNCH("bilinear_resize");
  }
  return SKP_OK;
}

extern "C" int skp_collect_maps_bwd(const float* d_out, int n_layers, int BH, int R, int N, const int64_t* idx,
                                    int n_idx, int R2, float* tmp, float* const* d_stored, void* stream) {
  SKP_REQUIRE(d_out && d_stored, "collect_maps_bwd: null pointer");
  SKP_REQUIRE(n_layers >= 1 && n_layers <= SKP_MAX_LAYERS, "collect_maps_bwd: n_layers=%d", n_layers);
  SKP_REQUIRE(BH > 0 && R > 0 && N > 0 && R2 > 0, "collect_maps_bwd: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  ColPtrs p{};
  for (int l = 0; l < n_layers; ++l) {
    SKP_REQUIRE(d_stored[l], "collect_maps_bwd: d_stored[%d] null", l);
    p.out[l] = d_stored[l];
  }
  const int T = idx ? n_idx : N, RR = R * R;
  const float* dmean = d_out;
  if (R2 != R) {
    SKP_REQUIRE(tmp, "collect_maps_bwd: tmp workspace required when resizing");
    size_t total = (size_t)T * RR;
    int b2 = (int)((total + 255) / 256);
    if (b2 > 148 * 16) b2 = 148 * 16;
    bilinear_resize_bwd_kernel<<<b2, 256, 0, st>>>(d_out, tmp, T, R, R2);
    SKP_CHECK_LAUNCH("bilinear_resize_bwd");
    dmean = tmp;
  }
  size_t smem = (size_t)COL_PIX * (N | 1) * sizeof(float);
  int blocks = (RR + COL_PIX - 1) / COL_PIX;
  cudaFuncSetAttribute(collect_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  collect_bwd_kernel<<<blocks, COL_THREADS, smem, st>>>(dmean, p, n_layers, BH, RR, N, idx, T);
  SKP_CHECK_LAUNCH("collect_bwd");
  return SKP_OK;
}
