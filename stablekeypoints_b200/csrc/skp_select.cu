// Arg-max / token-selection kernels: eval.py:39-111, ptp_utils.py:86-159.
// Integer outputs must be bit-exact with the reference: first-occurrence arg-max, stable ascending sort,
// strict-'>' furthest-point sampling with the same fp32 operation order (no FMA contraction).
#include "skp_common.cuh"
#include <math_constants.h>

namespace skp {

// (value, index) max with lowest index winning ties -- matches torch.argmax first-occurrence behaviour.
// NaN handling: torch treats NaN as the maximum; maps here are softmax means (finite), not replicated.
__device__ __forceinline__ void argmax_combine(float& v, int& i, float ov, int oi) {
  if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

__device__ __forceinline__ void block_argmax(float& v, int& i, float* sv, int* si) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, i, o);
    argmax_combine(v, i, ov, oi);
  }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) { sv[w] = v; si[w] = i; }
  __syncthreads();
  if (w == 0) {
    v = lane < nw ? sv[lane] : -CUDART_INF_F;
    i = lane < nw ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, v, o);
      int oi = __shfl_xor_sync(0xffffffffu, i, o);
      argmax_combine(v, i, ov, oi);
    }
    if (lane == 0) { sv[0] = v; si[0] = i; }
  }
  __syncthreads();
  v = sv[0]; i = si[0];
}

// One CTA per map.  mask_* : optional previous peaks (flat) within `radius` of which values are multiplied by 0
// (eval.py:83-111: integer pixel coords against the +0.5 peak centre, strict '>' keeps the value).
__global__ void __launch_bounds__(512) argmax_rows_kernel(const float* __restrict__ maps, int P, int W,
                                                          const int64_t* __restrict__ prev, int n_prev, int T,
                                                          float radius2, int64_t* __restrict__ out) {
  __shared__ float sv[32];
  __shared__ int si[32];
  const int t = blockIdx.x;
  const float* m = maps + (size_t)t * P;
  float best = -CUDART_INF_F;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    float v = m[i];
    if (n_prev > 0) {
      int y = i / W, x = i - y * W;
      for (int k = 0; k < n_prev; ++k) {
        int64_t pk = prev[(size_t)k * T + t];
        float cy = (float)(pk / W) + 0.5f, cx = (float)(pk % W) + 0.5f;
        float dx = (float)x - cx, dy = (float)y - cy;
        float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        v = v * ((d2 > radius2) ? 1.f : 0.f);
      }
    }
    if (v > best) { best = v; bi = i; }  // strided scan keeps the lowest index per thread
  }
  block_argmax(best, bi, sv, si);
  if (threadIdx.x == 0) out[t] = (bi == 0x7fffffff) ? 0 : bi;
}

// KL( tgt || softmax(map + eps) ), tgt = normalised (mean_k Gaussian_k + eps)   (ptp_utils.py:97-108)
// One CTA per token; three passes over the [H*W] map held in L1/L2 (64 KB at 128^2).
__global__ void __launch_bounds__(512) gaussian_kl_kernel(const float* __restrict__ maps, int T, int H, int W,
                                                          const int64_t* __restrict__ peaks, int num, float sigma,
                                                          float eps, float* __restrict__ kl) {
  __shared__ float red[32];
  const int t = blockIdx.x, P = H * W;
  const float* m = maps + (size_t)t * P;
  const float denom = sigma;  // host passes float(2.0 * sigma**2.0), optimize_token.py:220
  // pass 1: max of (map + eps), sum of target
  float mx = -CUDART_INF_F, tsum = 0.f;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    mx = fmaxf(mx, m[i] + eps);
    int y = i / W, x = i - y * W;
    float g = 0.f;
    for (int k = 0; k < num; ++k) {
      int64_t pk = peaks[(size_t)k * T + t];
      // peak position in pixels: ((idx + 0.5) / H) * size with size == H (optimize_token.py:211)
      float cy = ((float)(pk / W) + 0.5f) / (float)H * (float)H, cx = ((float)(pk % W) + 0.5f) / (float)H * (float)H;
      float dx = ((float)x + 0.5f) - cx, dy = ((float)y + 0.5f) - cy;
      g += expf(-1.f * (dx * dx + dy * dy) / denom);
    }
    tsum += g / (float)num + eps;
  }
  mx = warp_max(mx);
  {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = mx;
    __syncthreads();
    float r = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : -CUDART_INF_F;
    r = warp_max(r);
    __syncthreads();
    if (threadIdx.x == 0) red[0] = r;
    __syncthreads();
    mx = red[0];
  }
  tsum = block_sum(tsum, red);
  // pass 2: log-sum-exp
  float se = 0.f;
  for (int i = threadIdx.x; i < P; i += blockDim.x) se += expf(m[i] + eps - mx);
  se = block_sum(se, red);
  const float lse = mx + logf(se);
  // pass 3: KL
  float acc = 0.f;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    int y = i / W, x = i - y * W;
    float g = 0.f;
    for (int k = 0; k < num; ++k) {
      int64_t pk = peaks[(size_t)k * T + t];
      float cy = ((float)(pk / W) + 0.5f) / (float)H * (float)H, cx = ((float)(pk % W) + 0.5f) / (float)H * (float)H;
      float dx = ((float)x + 0.5f) - cx, dy = ((float)y + 0.5f) - cy;
      g += expf(-1.f * (dx * dx + dy * dy) / denom);
    }
    float tg = (g / (float)num + eps) / tsum;
    acc += tg * (logf(tg) - (m[i] + eps - lse));
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) kl[t] = acc;
}

// Stable ascending rank sort of T <= 1024 scalars; writes the first top_k indices.
__global__ void __launch_bounds__(1024) argsort_topk_kernel(const float* __restrict__ scores, int T, int top_k,
                                                            int64_t* __restrict__ out) {
  extern __shared__ float sc[];
  for (int i = threadIdx.x; i < T; i += blockDim.x) sc[i] = scores[i];
  __syncthreads();
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    float v = sc[i];
    const bool vn = v != v;
    int rank = 0;
    for (int j = 0; j < T; ++j) {
      float u = sc[j];
      const bool un = u != u;
      // total order (NaN last, like torch.argsort; ties by index): the ranks are a permutation, so every out[] slot is
      // written even for degenerate scores and no stale index can reach the loss kernels
      rank += (u < v) || (vn && !un) || ((u == v || (un && vn)) && j < i);
    }
    if (rank < top_k) out[rank] = i;
  }
}

// Single-thread furthest point sampling on <= a few dozen candidates (ptp_utils.py:115-159).  The reference
// is a Python double loop of a few hundred tiny kernels with a blocking .item() each; here it is one launch
// with no host round trip.  fp32 op order: loc = (idx + 0.5) / H ; d = sqrt((ay-by)^2 + (ax-bx)^2).
__device__ __forceinline__ float fps_dist(const float* ly, const float* lx, int a, int b) {
  float dy = __fsub_rn(ly[a], ly[b]), dx = __fsub_rn(lx[a], lx[b]);
  return __fsqrt_rn(__fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx)));
}

__global__ void fps_kernel(const int64_t* __restrict__ peaks, int H, int W, const int64_t* __restrict__ cand, int n_cand,
                           int top_k, int64_t* __restrict__ out, int32_t* __restrict__ n_out) {
  extern __shared__ float sh[];
  float* ly = sh;            // per candidate slot
  float* lx = sh + n_cand;
  int* chosen = reinterpret_cast<int*>(sh + 2 * n_cand);  // candidate slots picked so far
  for (int c = threadIdx.x; c < n_cand; c += blockDim.x) {
    int64_t pk = peaks[cand[c]];
    ly[c] = __fdiv_rn((float)(pk / W) + 0.5f, (float)H);
    lx[c] = __fdiv_rn((float)(pk % W) + 0.5f, (float)H);
  }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  // One warp; every reduction reproduces the serial loop's "first strict maximum": larger value wins, ties go to the
  // smaller (i, j) pair / candidate slot -- exactly the element the reference's Python loops keep.
  const int lane = threadIdx.x;
  float best = -1.f;
  int bp = 0x7fffffff;
  for (int p = lane; p < n_cand * n_cand; p += 32) {
    const int i = p / n_cand, j = p - i * n_cand;
    if (j > i) {
      const float dd = fps_dist(ly, lx, i, j);
      if (dd > best) { best = dd; bp = p; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int op = __shfl_xor_sync(0xffffffffu, bp, o);
    if (ob > best || (ob == best && op < bp)) { best = ob; bp = op; }
  }
  int n = 2;
  if (lane == 0) {
    chosen[0] = bp == 0x7fffffff ? 0 : bp / n_cand;
    chosen[1] = bp == 0x7fffffff ? 1 : bp % n_cand;
  }
  __syncwarp();
  for (int it = 0; it < top_k - 2; ++it) {
    float bmin = -1.f;
    int pick = 0x7fffffff;
    for (int c = lane; c < n_cand; c += 32) {
      // the reference skips by token VALUE (`i.item() in selected_indices`), so duplicate tokens are skipped too
      bool used = false;
      for (int k = 0; k < n; ++k) used |= (cand[chosen[k]] == cand[c]);
      if (used) continue;
      float mn = CUDART_INF_F;
      for (int k = 0; k < n; ++k) mn = fminf(mn, fps_dist(ly, lx, c, chosen[k]));
      if (mn > bmin) { bmin = mn; pick = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, bmin, o);
      const int op = __shfl_xor_sync(0xffffffffu, pick, o);
      if (ob > bmin || (ob == bmin && op < pick)) { bmin = ob; pick = op; }
    }
    if (pick != 0x7fffffff) {
      if (lane == 0) chosen[n] = pick;
      ++n;
    }
    __syncwarp();
  }
  for (int k = lane; k < n; k += 32) out[k] = cand[chosen[k]];
  if (lane == 0) *n_out = n;
}

}  // namespace skp

using namespace skp;

extern "C" int skp_argmax_rows(const float* maps, int T, int P, int64_t* flat_idx, void* stream) {
  SKP_REQUIRE(maps && flat_idx && T > 0 && P > 0, "argmax_rows: bad arguments");
  argmax_rows_kernel<<<T, 512, 0, (cudaStream_t)stream>>>(maps, P, P, nullptr, 0, T, 0.f, flat_idx);
  SKP_CHECK_LAUNCH("argmax_rows");
  return SKP_OK;
}

extern "C" int skp_k_argmax(const float* maps, int T, int H, int W, int num, float* work, int64_t* flat_idx,
                            void* stream) {
  SKP_REQUIRE(maps && flat_idx && T > 0 && H > 0 && W > 0 && num > 0, "k_argmax: bad arguments");
  (void)work;  // masking is evaluated on the fly from the previous peaks; no scratch copy of the maps is needed
  double radius = 0.05 * (double)H;  // eval.py:79: Python float, squared in double, compared as fp32
  float radius2 = (float)(radius * radius);
  for (int k = 0; k < num; ++k) {
    argmax_rows_kernel<<<T, 512, 0, (cudaStream_t)stream>>>(maps, H * W, W, flat_idx, k, T, radius2,
                                                            flat_idx + (size_t)k * T);
    SKP_CHECK_LAUNCH("k_argmax");
  }
  return SKP_OK;
}

// ptp_utils.py:165-187 entropy_sort: entropy of softmax-over-pixels of every token map,
//   H = lse - sum_i softmax_i * m_i   (== -sum p log p), one CTA per token, three block reductions.
__global__ void __launch_bounds__(512) entropy_kernel(const float* __restrict__ maps, int P, float* __restrict__ ent) {
  __shared__ float red[32];
  const float* m = maps + (size_t)blockIdx.x * P;
  float mx = -CUDART_INF_F;
  for (int i = threadIdx.x; i < P; i += blockDim.x) mx = fmaxf(mx, m[i]);
  mx = warp_max(mx);
  {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) red[w] = mx;
    __syncthreads();
    float r = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : -CUDART_INF_F;
    r = warp_max(r);
    __syncthreads();
    if (threadIdx.x == 0) red[0] = r;
    __syncthreads();
    mx = red[0];
  }
  float se = 0.f, sm = 0.f;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const float e = expf(m[i] - mx);
    se += e;
    sm = fmaf(e, m[i], sm);
  }
  se = block_sum(se, red);
  sm = block_sum(sm, red);
  if (threadIdx.x == 0) ent[blockIdx.x] = (mx + logf(se)) - sm / se;
}

extern "C" int skp_entropy_scores(const float* maps, int T, int P, float* ent, void* stream) {
  SKP_REQUIRE(maps && ent && T > 0 && P > 0, "entropy_scores: bad arguments");
  entropy_kernel<<<T, 512, 0, (cudaStream_t)stream>>>(maps, P, ent);
  SKP_CHECK_LAUNCH("entropy");
  return SKP_OK;
}

extern "C" int skp_gaussian_kl_scores(const float* maps, int T, int H, int W, const int64_t* peaks, int num,
                                      float sigma, float eps, float* kl, void* stream) {
  SKP_REQUIRE(maps && peaks && kl && T > 0 && H > 0 && W > 0 && num > 0 && sigma > 0.f, "gaussian_kl_scores: bad arguments");
  float denom = (float)(2.0 * (double)sigma * (double)sigma);
  gaussian_kl_kernel<<<T, 512, 0, (cudaStream_t)stream>>>(maps, T, H, W, peaks, num, denom, eps, kl);
  SKP_CHECK_LAUNCH("gaussian_kl");
  return SKP_OK;
}

extern "C" int skp_argsort_topk(const float* scores, int T, int top_k, int64_t* out_idx, void* stream) {
  SKP_REQUIRE(scores && out_idx && T > 0 && top_k > 0 && top_k <= T, "argsort_topk: bad arguments");
  SKP_REQUIRE(T <= 8192, "argsort_topk: T=%d too large", T);
  argsort_topk_kernel<<<1, 1024, T * sizeof(float), (cudaStream_t)stream>>>(scores, T, top_k, out_idx);
  SKP_CHECK_LAUNCH("argsort_topk");
  return SKP_OK;
}

extern "C" int skp_furthest_point_sampling(const int64_t* peaks_flat, int H, int W, const int64_t* candidates, int n_cand,
                                           int top_k, int64_t* out_idx, int32_t* n_out, void* stream) {
  SKP_REQUIRE(peaks_flat && candidates && out_idx && n_out, "fps: null pointer");
  SKP_REQUIRE(n_cand >= 2 && top_k >= 2 && H > 0 && W > 0, "fps: need >= 2 candidates and top_k >= 2");
  SKP_REQUIRE(n_cand <= 4096, "fps: too many candidates");
  size_t smem = (size_t)n_cand * 2 * sizeof(float) + (size_t)(top_k + 2) * sizeof(int);
  fps_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(peaks_flat, H, W, candidates, n_cand, top_k, out_idx, n_out);
  SKP_CHECK_LAUNCH("fps");
  return SKP_OK;
}
