// Loss-side kernels: sharpening (optimize.py:166-206, optimize_token.py:203-241), equivariance
// (optimize.py:157-163, invertable_transform.py:38-92), soft-arg-max (eval.py:113-155), Adam (optimize.py:320,424).
// All are small ([K,R,R] with K=10, R=128 -> 0.66 MB): launch-latency bound, so each loss is ONE multi-CTA kernel: a
// deterministic in-block reduction, then one atomicAdd of the CTA's share into the zeroed scalar (the logged loss value
// can differ in its last bits from run to run; no gradient depends on it).  The backward kernels fuse the loss weight.
#include "skp_common.cuh"

namespace skp {

__device__ __forceinline__ float gaussian_target(int y, int x, const int64_t* __restrict__ peaks, int num, int K, int k,
                                                 int H, int W, float denom) {
  float g = 0.f;
  for (int j = 0; j < num; ++j) {
    int64_t pk = peaks[(size_t)j * K + k];
    // pos = (idx + 0.5) / W (optimize.py:168) ; centre = pos * size with size = H (optimize_token.py:211)
    float cy = ((float)(pk / W) + 0.5f) / (float)W * (float)H, cx = ((float)(pk % W) + 0.5f) / (float)W * (float)H;
    float dx = ((float)x + 0.5f) - cx, dy = ((float)y + 0.5f) - cy;
    g += expf(-1.f * (dx * dx + dy * dy) / denom);
  }
  return g / (float)num;
}

// Multi-CTA reduction: each CTA adds its share of the mean into the zero-initialised device scalar (the loss sits on
// the serial part of the step between the two forwards' join and the backward, so one 1024-thread CTA cost 170 us).
__global__ void __launch_bounds__(256) sharpen_fwd_kernel(const float* __restrict__ maps, int H, int W,
                                                          const int64_t* __restrict__ sel, int K,
                                                          const int64_t* __restrict__ peaks, int num, float denom,
                                                          float* __restrict__ loss) {
  __shared__ float red[32];
  const int P = H * W;
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K * P; i += gridDim.x * blockDim.x) {
    int k = i / P, pix = i - k * P;
    int y = pix / W, x = pix - y * W;
    float d = maps[(size_t)sel[k] * P + pix] - gaussian_target(y, x, peaks, num, K, k, H, W, denom);
    acc = fmaf(d, d, acc);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(loss, acc / ((float)K * (float)P));
}

__global__ void sharpen_bwd_kernel(const float* __restrict__ maps, int H, int W, const int64_t* __restrict__ sel, int K,
                                   const int64_t* __restrict__ peaks, int num, float denom,
                                   const float* __restrict__ d_loss, float weight, float* __restrict__ d_maps) {
  const int P = H * W;
  const float c = 2.f * weight * (d_loss ? *d_loss : 1.f) / ((float)K * (float)P);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K * P; i += gridDim.x * blockDim.x) {
    int k = i / P, pix = i - k * P;
    int y = pix / W, x = pix - y * W;
    size_t o = (size_t)sel[k] * P + pix;
    float d = maps[o] - gaussian_target(y, x, peaks, num, K, k, H, W, denom);
    atomicAdd(d_maps + o, c * d);
  }
}

// affine_grid (align_corners=False) + bilinear grid_sample (zeros padding, align_corners=False) at output (y, x).
struct BilinearTap {
  int x0, y0;
  float w00, w01, w10, w11;  // (y0,x0) (y0,x1) (y1,x0) (y1,x1)
};

__device__ __forceinline__ BilinearTap affine_tap(const float* __restrict__ th, int y, int x, int H, int W) {
  float xn = (2.f * x + 1.f) / (float)W - 1.f, yn = (2.f * y + 1.f) / (float)H - 1.f;
  float gx = th[0] * xn + th[1] * yn + th[2];
  float gy = th[3] * xn + th[4] * yn + th[5];
  float ix = ((gx + 1.f) * (float)W - 1.f) * 0.5f, iy = ((gy + 1.f) * (float)H - 1.f) * 0.5f;
  float fx = floorf(ix), fy = floorf(iy);
  BilinearTap t;
  // clamp far-out-of-range coordinates so the int conversion is defined; such taps are out of bounds anyway
  fx = fminf(fmaxf(fx, -2.f), (float)W + 1.f);
  fy = fminf(fmaxf(fy, -2.f), (float)H + 1.f);
  t.x0 = (int)fx; t.y0 = (int)fy;
  float lx = ix - fx, ly = iy - fy;
  lx = fminf(fmaxf(lx, 0.f), 1.f); ly = fminf(fmaxf(ly, 0.f), 1.f);
  t.w00 = (1.f - ly) * (1.f - lx); t.w01 = (1.f - ly) * lx;
  t.w10 = ly * (1.f - lx);         t.w11 = ly * lx;
  return t;
}

__device__ __forceinline__ float sample_tap(const float* __restrict__ m, const BilinearTap& t, int H, int W) {
  float v = 0.f;
  bool xa = t.x0 >= 0 && t.x0 < W, xb = t.x0 + 1 >= 0 && t.x0 + 1 < W;
  bool ya = t.y0 >= 0 && t.y0 < H, yb = t.y0 + 1 >= 0 && t.y0 + 1 < H;
  if (ya && xa) v = fmaf(t.w00, m[t.y0 * W + t.x0], v);
  if (ya && xb) v = fmaf(t.w01, m[t.y0 * W + t.x0 + 1], v);
  if (yb && xa) v = fmaf(t.w10, m[(t.y0 + 1) * W + t.x0], v);
  if (yb && xb) v = fmaf(t.w11, m[(t.y0 + 1) * W + t.x0 + 1], v);
  return v;
}

__global__ void __launch_bounds__(256) equiv_fwd_kernel(const float* __restrict__ maps, const float* __restrict__ maps_t,
                                                        int H, int W, const int64_t* __restrict__ sel, int K,
                                                        const float* __restrict__ theta_inv, float* __restrict__ loss) {
  __shared__ float red[32];
  __shared__ float th[6];
  if (threadIdx.x < 6) th[threadIdx.x] = theta_inv[threadIdx.x];
  __syncthreads();
  const int P = H * W;
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    int y = i / W, x = i - y * W;
    BilinearTap t = affine_tap(th, y, x, H, W);
    for (int k = 0; k < K; ++k) {
      size_t base = (size_t)sel[k] * P;
      float d = maps[base + i] - sample_tap(maps_t + base, t, H, W);
      acc = fmaf(d, d, acc);
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(loss, acc / ((float)K * (float)P));
}

// Backward of the equivariance loss.  d_maps gets one contribution per pixel.  d_maps_t is the transposed bilinear
// sampling: instead of scattering four float atomics per source pixel (whose arrival order -- hence the rounding of the
// sum -- changes from run to run), every destination pixel q GATHERS, in a fixed order, from the source pixels whose
// bilinear footprint covers it.  The sampling position is affine in the pixel coordinates, (ix, iy) = A (x, y) + b, and
// a footprint covers q iff (ix, iy) lies in q + [-1, 1)^2, so the candidates sit in the bounding box of A^-1 applied to
// that square around A^-1 (q - b); membership is then decided by the very affine_tap() the forward uses, which makes
// every term bit-identical to the scattered one.  A near-singular theta (box wider than GATHER_MAX) keeps the scatter.
constexpr float EQUIV_GATHER_MAX = 6.f;

__global__ void equiv_bwd_kernel(const float* __restrict__ maps, const float* __restrict__ maps_t, int H, int W,
                                 const int64_t* __restrict__ sel, int K, const float* __restrict__ theta_inv,
                                 const float* __restrict__ d_loss, float weight, float* __restrict__ d_maps,
                                 float* __restrict__ d_maps_t) {
  __shared__ float th[6];
  if (threadIdx.x < 6) th[threadIdx.x] = theta_inv[threadIdx.x];
  __syncthreads();
  const int P = H * W;
  const float c = 2.f * weight * (d_loss ? *d_loss : 1.f) / ((float)K * (float)P);
  // pixel-space form of affine_tap: ix = a00 x + a01 y + bx, iy = a10 x + a11 y + by
  const float a00 = th[0], a01 = th[1] * (float)W / (float)H, a10 = th[3] * (float)H / (float)W, a11 = th[4];
  const float bx = ((th[0] * (1.f / (float)W - 1.f) + th[1] * (1.f / (float)H - 1.f) + th[2] + 1.f) * (float)W - 1.f) * 0.5f;
  const float by = ((th[3] * (1.f / (float)W - 1.f) + th[4] * (1.f / (float)H - 1.f) + th[5] + 1.f) * (float)H - 1.f) * 0.5f;
  const float det = a00 * a11 - a01 * a10;
  const float rdet = det != 0.f ? 1.f / det : 0.f;
  const float i00 = a11 * rdet, i01 = -a01 * rdet, i10 = -a10 * rdet, i11 = a00 * rdet;
  const float rx = fabsf(i00) + fabsf(i01) + 0.02f, ry = fabsf(i10) + fabsf(i11) + 0.02f;   // half extents of A^-1 [-1,1)^2
  const bool gather = det != 0.f && rx <= EQUIV_GATHER_MAX && ry <= EQUIV_GATHER_MAX;            // uniform over the grid
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K * P; i += gridDim.x * blockDim.x) {
    int k = i / P, pix = i - k * P;
    int y = pix / W, x = pix - y * W;
    size_t base = (size_t)sel[k] * P;
    const float* mt = maps_t + base;
    if (d_maps || !gather) {
      BilinearTap t = affine_tap(th, y, x, H, W);
      float g = c * (maps[base + pix] - sample_tap(mt, t, H, W));
      if (d_maps) atomicAdd(d_maps + base + pix, g);
      if (d_maps_t && !gather) {
        float* dt = d_maps_t + base;
        bool xa = t.x0 >= 0 && t.x0 < W, xb = t.x0 + 1 >= 0 && t.x0 + 1 < W;
        bool ya = t.y0 >= 0 && t.y0 < H, yb = t.y0 + 1 >= 0 && t.y0 + 1 < H;
        if (ya && xa) atomicAdd(dt + t.y0 * W + t.x0, -g * t.w00);
        if (ya && xb) atomicAdd(dt + t.y0 * W + t.x0 + 1, -g * t.w01);
        if (yb && xa) atomicAdd(dt + (t.y0 + 1) * W + t.x0, -g * t.w10);
        if (yb && xb) atomicAdd(dt + (t.y0 + 1) * W + t.x0 + 1, -g * t.w11);
      }
    }
    if (d_maps_t && gather) {
      // (x, y) now plays the destination pixel q of d_maps_t
      const float qx = (float)x - bx, qy = (float)y - by;
      const float cx = i00 * qx + i01 * qy, cy = i10 * qx + i11 * qy;
      const int px0 = max(0, (int)ceilf(cx - rx)), px1 = min(W - 1, (int)floorf(cx + rx));
      const int py0 = max(0, (int)ceilf(cy - ry)), py1 = min(H - 1, (int)floorf(cy + ry));
      float acc = 0.f;
      for (int py = py0; py <= py1; ++py)
        for (int px = px0; px <= px1; ++px) {
          BilinearTap t = affine_tap(th, py, px, H, W);
          const int ddx = x - t.x0, ddy = y - t.y0;
          if ((unsigned)ddx > 1u || (unsigned)ddy > 1u) continue;
          const float wq = ddy ? (ddx ? t.w11 : t.w10) : (ddx ? t.w01 : t.w00);
          const float g = c * (maps[base + py * W + px] - sample_tap(mt, t, H, W));
          acc += -g * wq;
        }
      atomicAdd(d_maps_t + base + pix, acc);   // one contribution per address (atomic only against duplicate tokens in sel)
    }
  }
}

__global__ void affine_warp_kernel(const float* __restrict__ img, int B, int C, int H, int W,
                                   const float* __restrict__ theta, float* __restrict__ out) {
  const int P = H * W;
  size_t total = (size_t)B * P;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int b = (int)(i / P), pix = (int)(i - (size_t)b * P);
    int y = pix / W, x = pix - y * W;
    BilinearTap t = affine_tap(theta + b * 6, y, x, H, W);
    for (int c = 0; c < C; ++c) {
      size_t base = ((size_t)b * C + c) * P;
      out[base + pix] = sample_tap(img + base, t, H, W);
    }
  }
}

__global__ void affine_warp_bwd_kernel(const float* __restrict__ dout, int B, int C, int H, int W,
                                       const float* __restrict__ theta, float* __restrict__ dimg) {
  const int P = H * W;
  size_t total = (size_t)B * P;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int b = (int)(i / P), pix = (int)(i - (size_t)b * P);
    int y = pix / W, x = pix - y * W;
    BilinearTap t = affine_tap(theta + b * 6, y, x, H, W);
    bool xa = t.x0 >= 0 && t.x0 < W, xb = t.x0 + 1 >= 0 && t.x0 + 1 < W;
    bool ya = t.y0 >= 0 && t.y0 < H, yb = t.y0 + 1 >= 0 && t.y0 + 1 < H;
    for (int c = 0; c < C; ++c) {
      size_t base = ((size_t)b * C + c) * P;
      float g = dout[base + pix];
      float* d = dimg + base;
      if (ya && xa) atomicAdd(d + t.y0 * W + t.x0, g * t.w00);
      if (ya && xb) atomicAdd(d + t.y0 * W + t.x0 + 1, g * t.w01);
      if (yb && xa) atomicAdd(d + (t.y0 + 1) * W + t.x0, g * t.w10);
      if (yb && xb) atomicAdd(d + (t.y0 + 1) * W + t.x0 + 1, g * t.w11);
    }
  }
}

// One CTA per heat-map: zero in place beyond `distance` of the peak, then expectation over what is left.
__global__ void __launch_bounds__(1024) soft_argmax_kernel(float* __restrict__ hm, int H, int W,
                                                           const int64_t* __restrict__ peaks, float distance,
                                                           float* __restrict__ out) {
  __shared__ float red[32];
  const int t = blockIdx.x, P = H * W;
  float* m = hm + (size_t)t * P;
  const int64_t pk = peaks[t];
  const float py = (float)(pk / W), px = (float)(pk % W);
  float s = 0.f, sy = 0.f, sx = 0.f;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    int y = i / W, x = i - y * W;
    float v;
    if (distance >= 0.f) {
      float dy = (float)y - py, dx = (float)x - px;
      float d = sqrtf(dy * dy + dx * dx);
      if (d > distance) { m[i] = 0.f; v = 0.f; }
      else v = m[i];
    } else {
      v = m[i];
    }
    s += v; sy = fmaf((float)y, v, sy); sx = fmaf((float)x, v, sx);
  }
  s = block_sum(s, red);
  sy = block_sum(sy, red);
  sx = block_sum(sx, red);
  if (threadIdx.x == 0) {
    float den = s + 1e-6f;
    out[2 * t + 0] = sy / den + 0.5f;
    out[2 * t + 1] = sx / den + 0.5f;
  }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int64_t n, float lr_over_bc1, float inv_sqrt_bc2, float b1, float b2, float eps, float gscale) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    float mi = m[i] * b1 + (1.f - b1) * gi;      // exp_avg.lerp_(grad, 1 - beta1)
    float vi = v[i] * b2 + (1.f - b2) * gi * gi;  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] = p[i] - lr_over_bc1 * (mi / denom);
  }
}

// CUDA-graph friendly variant: the step count lives on the device so a captured optimizer step stays correct on replay.
__global__ void adam_bump_step_kernel(int* step) { *step += 1; }

__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                int64_t n, const int* __restrict__ step, float lr, float b1, float b2, float eps, float gscale) {
  const double t = (double)*step;
  const float lr_over_bc1 = (float)((double)lr / (1.0 - pow((double)b1, t)));
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)b2, t)));
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    float mi = m[i] * b1 + (1.f - b1) * gi;
    float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] = p[i] - lr_over_bc1 * (mi / denom);
  }
}

static inline int grid_for(size_t n, int threads) {
  size_t b = (n + threads - 1) / threads;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace skp

using namespace skp;

extern "C" int skp_sharpen_loss_fwd(const float* maps, int H, int W, const int64_t* sel, int K, const int64_t* peaks,
                                    int num, float sigma, float* loss, void* stream) {
  SKP_REQUIRE(maps && sel && peaks && loss && H > 0 && W > 0 && K > 0 && num > 0 && sigma > 0.f, "sharpen_loss_fwd: bad arguments");
  float denom = (float)(2.0 * (double)sigma * (double)sigma);
  cudaMemsetAsync(loss, 0, sizeof(float), (cudaStream_t)stream);
  sharpen_fwd_kernel<<<grid_for((size_t)K * H * W, 256), 256, 0, (cudaStream_t)stream>>>(maps, H, W, sel, K, peaks, num, denom, loss);
  SKP_CHECK_LAUNCH("sharpen_fwd");
  return SKP_OK;
}

extern "C" int skp_sharpen_loss_bwd(const float* maps, int H, int W, const int64_t* sel, int K, const int64_t* peaks,
                                    int num, float sigma, const float* d_loss, float weight, float* d_maps, void* stream) {
  SKP_REQUIRE(maps && sel && peaks && d_maps && H > 0 && W > 0 && K > 0 && num > 0 && sigma > 0.f, "sharpen_loss_bwd: bad arguments");
  float denom = (float)(2.0 * (double)sigma * (double)sigma);
  sharpen_bwd_kernel<<<grid_for((size_t)K * H * W, 256), 256, 0, (cudaStream_t)stream>>>(maps, H, W, sel, K, peaks, num,
                                                                                        denom, d_loss, weight, d_maps);
  SKP_CHECK_LAUNCH("sharpen_bwd");
  return SKP_OK;
}

extern "C" int skp_equivariance_loss_fwd(const float* maps, const float* maps_t, int H, int W, const int64_t* sel, int K,
                                         const float* theta_inv, float* loss, void* stream) {
  SKP_REQUIRE(maps && maps_t && sel && theta_inv && loss && H > 0 && W > 0 && K > 0, "equivariance_loss_fwd: bad arguments");
  cudaMemsetAsync(loss, 0, sizeof(float), (cudaStream_t)stream);
  equiv_fwd_kernel<<<grid_for((size_t)H * W, 256), 256, 0, (cudaStream_t)stream>>>(maps, maps_t, H, W, sel, K, theta_inv, loss);
  SKP_CHECK_LAUNCH("equiv_fwd");
  return SKP_OK;
}

extern "C" int skp_equivariance_loss_bwd(const float* maps, const float* maps_t, int H, int W, const int64_t* sel, int K,
                                         const float* theta_inv, const float* d_loss, float weight, float* d_maps,
                                         float* d_maps_t, void* stream) {
  SKP_REQUIRE(maps && maps_t && sel && theta_inv && H > 0 && W > 0 && K > 0, "equivariance_loss_bwd: bad arguments");
  equiv_bwd_kernel<<<grid_for((size_t)K * H * W, 256), 256, 0, (cudaStream_t)stream>>>(maps, maps_t, H, W, sel, K, theta_inv,
                                                                                      d_loss, weight, d_maps, d_maps_t);
  SKP_CHECK_LAUNCH("equiv_bwd");
  return SKP_OK;
}

extern "C" int skp_affine_warp(const float* img, int B, int C, int H, int W, const float* theta, float* out, void* stream) {
  SKP_REQUIRE(img && theta && out && B > 0 && C > 0 && H > 0 && W > 0, "affine_warp: bad arguments");
  affine_warp_kernel<<<grid_for((size_t)B * H * W, 256), 256, 0, (cudaStream_t)stream>>>(img, B, C, H, W, theta, out);
  SKP_CHECK_LAUNCH("affine_warp");
  return SKP_OK;
}

extern "C" int skp_affine_warp_bwd(const float* d_out, int B, int C, int H, int W, const float* theta, float* d_img,
                                   void* stream) {
  SKP_REQUIRE(d_out && theta && d_img && B > 0 && C > 0 && H > 0 && W > 0, "affine_warp_bwd: bad arguments");
  affine_warp_bwd_kernel<<<grid_for((size_t)B * H * W, 256), 256, 0, (cudaStream_t)stream>>>(d_out, B, C, H, W, theta, d_img);
  SKP_CHECK_LAUNCH("affine_warp_bwd");
  return SKP_OK;
}

extern "C" int skp_soft_argmax(float* heatmaps, int T, int H, int W, const int64_t* peaks, float distance, float* out,
                               void* stream) {
  SKP_REQUIRE(heatmaps && peaks && out && T > 0 && H > 0 && W > 0, "soft_argmax: bad arguments");
  soft_argmax_kernel<<<T, 1024, 0, (cudaStream_t)stream>>>(heatmaps, H, W, peaks, distance, out);
  SKP_CHECK_LAUNCH("soft_argmax");
  return SKP_OK;
}

extern "C" int skp_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int step,
                             float lr, float beta1, float beta2, float eps, float grad_scale, void* stream) {
  SKP_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "adam_step: bad arguments");
  double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_kernel<<<grid_for((size_t)n, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n,
                                                                         (float)((double)lr / bc1), (float)(1.0 / sqrt(bc2)),
                                                                         beta1, beta2, eps, grad_scale);
  SKP_CHECK_LAUNCH("adam");
  return SKP_OK;
}

extern "C" int skp_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int* step_dev,
                                 float lr, float beta1, float beta2, float eps, float grad_scale, void* stream) {
  SKP_REQUIRE(param && grad && exp_avg && exp_avg_sq && step_dev && n > 0, "adam_step_dev: bad arguments");
  adam_bump_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
  SKP_CHECK_LAUNCH("adam_bump_step");
  adam_dev_kernel<<<grid_for((size_t)n, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, step_dev, lr,
                                                                             beta1, beta2, eps, grad_scale);
  SKP_CHECK_LAUNCH("adam_dev");
  return SKP_OK;
}

// ------------------------------------------------------------------ eval-time augmentation ensemble (eval.py:250-262,333-336)
namespace skp {
// sum[k,p] += unwarp(maps[k])[p]; num[k,p] += unwarp(ones)[p]  (one pass instead of two grid_samples + two adds)
__global__ void unwarp_accumulate_kernel(const float* __restrict__ maps, int K, int H, int W, const float* __restrict__ theta_inv,
                                         float* __restrict__ sum, float* __restrict__ num) {
  __shared__ float th[6];
  if (threadIdx.x < 6) th[threadIdx.x] = theta_inv[threadIdx.x];
  __syncthreads();
  const int P = H * W;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < P; pix += gridDim.x * blockDim.x) {
    const int y = pix / W, x = pix - y * W;
    const BilinearTap t = affine_tap(th, y, x, H, W);
    const bool xa = t.x0 >= 0 && t.x0 < W, xb = t.x0 + 1 >= 0 && t.x0 + 1 < W;
    const bool ya = t.y0 >= 0 && t.y0 < H, yb = t.y0 + 1 >= 0 && t.y0 + 1 < H;
    float ones = 0.f;
    if (ya && xa) ones += t.w00;
    if (ya && xb) ones += t.w01;
    if (yb && xa) ones += t.w10;
    if (yb && xb) ones += t.w11;
    for (int k = 0; k < K; ++k) {
      const size_t o = (size_t)k * P + pix;
      sum[o] += sample_tap(maps + (size_t)k * P, t, H, W);
      num[o] += ones;
    }
  }
}
// out = sum / num with 0/0 -> 0 (eval.py:333-336: NaNs replaced by 0)
__global__ void ensemble_finalize_kernel(const float* __restrict__ sum, const float* __restrict__ num, float* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v = sum[i] / num[i];
    out[i] = (v != v) ? 0.f : v;
  }
}
}  // namespace skp

extern "C" int skp_unwarp_accumulate(const float* maps, int K, int H, int W, const float* theta_inv, float* sum_samples,
                                     float* num_samples, void* stream) {
  SKP_REQUIRE(maps && theta_inv && sum_samples && num_samples && K > 0 && H > 0 && W > 0, "unwarp_accumulate: bad arguments");
  unwarp_accumulate_kernel<<<grid_for((size_t)H * W, 256), 256, 0, (cudaStream_t)stream>>>(maps, K, H, W, theta_inv, sum_samples, num_samples);
  SKP_CHECK_LAUNCH("unwarp_accumulate");
  return SKP_OK;
}

extern "C" int skp_ensemble_finalize(const float* sum_samples, const float* num_samples, float* out, int64_t n, void* stream) {
  SKP_REQUIRE(sum_samples && num_samples && out && n > 0, "ensemble_finalize: bad arguments");
  ensemble_finalize_kernel<<<grid_for((size_t)n, 256), 256, 0, (cudaStream_t)stream>>>(sum_samples, num_samples, out, (size_t)n);
  SKP_CHECK_LAUNCH("ensemble_finalize");
  return SKP_OK;
}
