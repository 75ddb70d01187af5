// Cross-attention core, exact-fp32 SIMT formulation (ptp_utils.py:493-506): per head
//   sim = q k^T * scale ; attn = softmax(sim, -1) ; out = attn v
// and its backward.  The token axis is short (N = 77..500) and there is one image per rank, so these layers
// are a few hundred MFLOP each: the point of this version is fp32-exact logits (they feed the captured maps
// whose parity budget is 1e-3 after a softmax) with K/V tiles staged once per CTA in shared memory.
// The scaled logits [h, S, N] are written out: they are both the saved tensor for backward and the input of
// the capture kernels (skp_capture.cu).
#include "skp_common.cuh"
#include <math_constants.h>

namespace skp {

constexpr int AT_ROWS = 32;      // query rows per CTA
constexpr int AT_THREADS = 256;  // 8 warps x 4 rows
constexpr int AT_TN = 64;        // K/V tokens staged per chunk
constexpr int AT_ND = 5;         // head-dim accumulators per lane -> d <= 160

// stage rows [r0, r0+AT_TN) x d of a [*, ld] matrix (head column offset applied by caller) into smem [AT_TN][d+1]
__device__ __forceinline__ void stage_tokens(float* dst, const float* __restrict__ src, int64_t ld, int r0, int nrows,
                                             int d) {
  for (int i = threadIdx.x; i < AT_TN * d; i += AT_THREADS) {
    int j = i / d, dd = i - j * d;
    dst[j * (d + 1) + dd] = (r0 + j < nrows) ? __ldg(src + (size_t)(r0 + j) * ld + dd) : 0.f;
  }
}

// acc[r][t] = sum_dd a[r][dd] * ks[t][dd] for the warp's 4 rows and tokens lane, lane+32 of the staged chunk
__device__ __forceinline__ void rows_dot_tokens(const float* __restrict__ as, const float* __restrict__ ks, int d, int w,
                                                int lane, float acc[4][2]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = 0.f;
  const float* k0 = ks + lane * (d + 1);
  const float* k1 = ks + (lane + 32) * (d + 1);
  const float* a0 = as + (w * 4) * d;
  for (int dd = 0; dd < d; ++dd) {
    float kv0 = k0[dd], kv1 = k1[dd];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float av = a0[r * d + dd];
      acc[r][0] = fmaf(av, kv0, acc[r][0]);
      acc[r][1] = fmaf(av, kv1, acc[r][1]);
    }
  }
}

// acc[r][i] += sum_{t in chunk} ps[r][c0+t] * vs[t][lane + 32 i]
__device__ __forceinline__ void probs_times_tokens(const float* __restrict__ ps, int Ns, const float* __restrict__ vs,
                                                   int d, int c0, int cn, int w, int lane, float acc[4][AT_ND]) {
  for (int t = 0; t < cn; ++t) {
    float vv[AT_ND];
#pragma unroll
    for (int i = 0; i < AT_ND; ++i) vv[i] = (lane + 32 * i < d) ? vs[t * (d + 1) + lane + 32 * i] : 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float pv = ps[(w * 4 + r) * Ns + c0 + t];
#pragma unroll
      for (int i = 0; i < AT_ND; ++i) acc[r][i] = fmaf(pv, vv[i], acc[r][i]);
    }
  }
}

__global__ void __launch_bounds__(AT_THREADS) cross_attn_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                                    int64_t ldk, const float* __restrict__ v, int64_t ldv,
                                                                    float* __restrict__ o, float* __restrict__ logits, int S,
                                                                    int N, int heads, int d, float scale) {
  extern __shared__ float sm[];
  const int Ns = N | 1, C = heads * d;
  float* qs = sm;                      // [32][d]
  float* ks = qs + AT_ROWS * d;        // [64][d+1]
  float* ps = ks + AT_TN * (d + 1);    // [32][Ns]
  const int h = blockIdx.y, s0 = blockIdx.x * AT_ROWS;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < AT_ROWS * d; i += AT_THREADS) {
    int r = i / d, dd = i - r * d;
    qs[i] = (s0 + r < S) ? __ldg(q + (size_t)(s0 + r) * C + h * d + dd) : 0.f;
  }
  // phase A: scaled logits
  for (int c0 = 0; c0 < N; c0 += AT_TN) {
    __syncthreads();
    stage_tokens(ks, k + h * d, ldk, c0, N, d);
    __syncthreads();
    float acc[4][2];
    rows_dot_tokens(qs, ks, d, w, lane, acc);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        int n = c0 + lane + 32 * t, row = s0 + w * 4 + r;
        if (n < N) {
          float l = acc[r][t] * scale;
          ps[(w * 4 + r) * Ns + n] = l;
          if (row < S) logits[((size_t)h * S + row) * N + n] = l;
        }
      }
  }
  __syncwarp();
  // phase B: softmax over tokens (each warp owns its 4 rows of ps: no block barrier needed)
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float* row = ps + (w * 4 + r) * Ns;
    float m = -CUDART_INF_F;
    for (int n = lane; n < N; n += 32) m = fmaxf(m, row[n]);
    m = warp_max(m);
    float ssum = 0.f;
    for (int n = lane; n < N; n += 32) {
      float e = __expf(row[n] - m);
      row[n] = e;
      ssum += e;
    }
    ssum = warp_sum(ssum);
    float inv = 1.f / ssum;
    for (int n = lane; n < N; n += 32) row[n] *= inv;
  }
  // phase C: out = P V
  float oacc[4][AT_ND];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int i = 0; i < AT_ND; ++i) oacc[r][i] = 0.f;
  for (int c0 = 0; c0 < N; c0 += AT_TN) {
    __syncthreads();
    stage_tokens(ks, v + h * d, ldv, c0, N, d);
    __syncthreads();
    probs_times_tokens(ps, Ns, ks, d, c0, min(AT_TN, N - c0), w, lane, oacc);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int row = s0 + w * 4 + r;
    if (row < S)
#pragma unroll
      for (int i = 0; i < AT_ND; ++i)
        if (lane + 32 * i < d) o[(size_t)row * C + h * d + lane + 32 * i] = oacc[r][i];
  }
}

// Backward part 1 (per head, 32 query rows): P from logits, dP = dO V^T, dS = P (dP - <P,dP>) + extra, dQ = scale dS K.
// Writes dS and the row statistics (max, 1/sum) for part 2.
__global__ void __launch_bounds__(AT_THREADS) cross_attn_bwd_dq_kernel(
    const float* __restrict__ d_o, const float* __restrict__ k, int64_t ldk, const float* __restrict__ v, int64_t ldv,
    const float* __restrict__ logits, const float* __restrict__ extra, float* __restrict__ ds_ws, float* __restrict__ stats,
    float* __restrict__ dq, int S, int N, int heads, int d, float scale) {
  extern __shared__ float sm[];
  const int Ns = N | 1, C = heads * d;
  float* dos = sm;                      // [32][d]
  float* ks = dos + AT_ROWS * d;        // [64][d+1]
  float* ps = ks + AT_TN * (d + 1);     // [32][Ns]  P
  float* dps = ps + AT_ROWS * Ns;       // [32][Ns]  dP then dS
  const int h = blockIdx.y, s0 = blockIdx.x * AT_ROWS;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < AT_ROWS * d; i += AT_THREADS) {
    int r = i / d, dd = i - r * d;
    dos[i] = (s0 + r < S) ? __ldg(d_o + (size_t)(s0 + r) * C + h * d + dd) : 0.f;
  }
  // P rows (warp-private)
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int row = s0 + w * 4 + r;
    float* pr = ps + (w * 4 + r) * Ns;
    if (row < S) {
      const float* lg = logits + ((size_t)h * S + row) * N;
      float m = -CUDART_INF_F;
      for (int n = lane; n < N; n += 32) { float l = __ldg(lg + n); pr[n] = l; m = fmaxf(m, l); }
      m = warp_max(m);
      float ssum = 0.f;
      for (int n = lane; n < N; n += 32) { float e = __expf(pr[n] - m); pr[n] = e; ssum += e; }
      ssum = warp_sum(ssum);
      float inv = 1.f / ssum;
      for (int n = lane; n < N; n += 32) pr[n] *= inv;
      if (lane == 0) { stats[((size_t)h * S + row) * 2] = m; stats[((size_t)h * S + row) * 2 + 1] = inv; }
    } else {
      for (int n = lane; n < N; n += 32) pr[n] = 0.f;
    }
  }
  // dP = dO V^T
  for (int c0 = 0; c0 < N; c0 += AT_TN) {
    __syncthreads();
    stage_tokens(ks, v + h * d, ldv, c0, N, d);
    __syncthreads();
    float acc[4][2];
    rows_dot_tokens(dos, ks, d, w, lane, acc);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        int n = c0 + lane + 32 * t;
        if (n < N) dps[(w * 4 + r) * Ns + n] = acc[r][t];
      }
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int row = s0 + w * 4 + r;
    float* pr = ps + (w * 4 + r) * Ns;
    float* dr = dps + (w * 4 + r) * Ns;
    float dot = 0.f;
    for (int n = lane; n < N; n += 32) dot = fmaf(pr[n], dr[n], dot);
    dot = warp_sum(dot);
    for (int n = lane; n < N; n += 32) {
      float g = pr[n] * (dr[n] - dot);
      if (row < S) {
        if (extra) g += __ldg(extra + ((size_t)h * S + row) * N + n);
        ds_ws[((size_t)h * S + row) * N + n] = g;
      } else {
        g = 0.f;
      }
      dr[n] = g;
    }
  }
  // dQ = scale * dS K
  float qacc[4][AT_ND];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int i = 0; i < AT_ND; ++i) qacc[r][i] = 0.f;
  for (int c0 = 0; c0 < N; c0 += AT_TN) {
    __syncthreads();
    stage_tokens(ks, k + h * d, ldk, c0, N, d);
    __syncthreads();
    probs_times_tokens(dps, Ns, ks, d, c0, min(AT_TN, N - c0), w, lane, qacc);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int row = s0 + w * 4 + r;
    if (row < S)
#pragma unroll
      for (int i = 0; i < AT_ND; ++i)
        if (lane + 32 * i < d) dq[(size_t)row * C + h * d + lane + 32 * i] = qacc[r][i] * scale;
  }
}

// Backward part 2 (per head, 32 tokens, chunk of query rows): dK = scale dS^T Q, dV = P^T dO; reduction over the
// S axis is split across CTAs and finished with fp32 atomics into zero-initialised dk/dv.
constexpr int KV_TOK = 32;
constexpr int KV_SCHUNK = 128;
constexpr int KV_SROWS = 32;

__global__ void __launch_bounds__(AT_THREADS) cross_attn_bwd_dkv_kernel(
    const float* __restrict__ d_o, const float* __restrict__ q, const float* __restrict__ logits,
    const float* __restrict__ ds_ws, const float* __restrict__ stats, float* __restrict__ dk, float* __restrict__ dv, int S,
    int N, int heads, int d, float scale) {
  extern __shared__ float sm[];
  const int C = heads * d;
  float* qs = sm;                        // [32 rows][d]
  float* dos = qs + KV_SROWS * d;        // [32 rows][d]
  float* dss = dos + KV_SROWS * d;       // [32 rows][33]
  float* pss = dss + KV_SROWS * 33;      // [32 rows][33]
  const int h = blockIdx.y, n0 = blockIdx.x * KV_TOK, sbeg = blockIdx.z * KV_SCHUNK;
  const int send = min(S, sbeg + KV_SCHUNK);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float ak[4][AT_ND], av[4][AT_ND];
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int i = 0; i < AT_ND; ++i) ak[t][i] = av[t][i] = 0.f;
  for (int r0 = sbeg; r0 < send; r0 += KV_SROWS) {
    __syncthreads();
    for (int i = threadIdx.x; i < KV_SROWS * d; i += AT_THREADS) {
      int r = i / d, dd = i - r * d;
      bool ok = r0 + r < send;
      qs[i] = ok ? __ldg(q + (size_t)(r0 + r) * C + h * d + dd) : 0.f;
      dos[i] = ok ? __ldg(d_o + (size_t)(r0 + r) * C + h * d + dd) : 0.f;
    }
    for (int i = threadIdx.x; i < KV_SROWS * KV_TOK; i += AT_THREADS) {
      int r = i >> 5, t = i & 31;
      bool ok = (r0 + r < send) && (n0 + t < N);
      size_t off = ((size_t)h * S + r0 + r) * N + n0 + t;
      float dsv = 0.f, pv = 0.f;
      if (ok) {
        dsv = __ldg(ds_ws + off);
        const float* st = stats + ((size_t)h * S + r0 + r) * 2;
        pv = __expf(__ldg(logits + off) - __ldg(st)) * __ldg(st + 1);
      }
      dss[r * 33 + t] = dsv;
      pss[r * 33 + t] = pv;
    }
    __syncthreads();
    for (int r = 0; r < KV_SROWS; ++r) {
      float qv[AT_ND], ov[AT_ND];
#pragma unroll
      for (int i = 0; i < AT_ND; ++i) {
        bool ok = lane + 32 * i < d;
        qv[i] = ok ? qs[r * d + lane + 32 * i] : 0.f;
        ov[i] = ok ? dos[r * d + lane + 32 * i] : 0.f;
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float dsv = dss[r * 33 + w * 4 + t], pv = pss[r * 33 + w * 4 + t];
#pragma unroll
        for (int i = 0; i < AT_ND; ++i) {
          ak[t][i] = fmaf(dsv, qv[i], ak[t][i]);
          av[t][i] = fmaf(pv, ov[i], av[t][i]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    int n = n0 + w * 4 + t;
    if (n < N)
#pragma unroll
      for (int i = 0; i < AT_ND; ++i)
        if (lane + 32 * i < d) {
          atomicAdd(dk + (size_t)n * C + h * d + lane + 32 * i, ak[t][i] * scale);
          atomicAdd(dv + (size_t)n * C + h * d + lane + 32 * i, av[t][i]);
        }
  }
}

// ---------------------------------------------------------------------------- fp32 SIMT GEMM (NT)
constexpr int GM = 64, GN = 64, GK = 16;

__global__ void __launch_bounds__(256) gemm_nt_simt_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                                                           int64_t ldb, float* __restrict__ Cm, int64_t ldc, int M, int N, int K,
                                                           float alpha, const float* __restrict__ bias,
                                                           const float* __restrict__ residual, int64_t ldr) {
  __shared__ float As[GK][GM + 4];
  __shared__ float Bs[GK][GN + 4];
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += GK) {
    // 64x16 tiles: thread loads 4 elements of A and 4 of B (k fastest for coalescing)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = threadIdx.x + e * 256;
      int r = idx >> 4, kk = idx & 15;
      As[kk][r] = (m0 + r < M && k0 + kk < K) ? __ldg(A + (size_t)(m0 + r) * lda + k0 + kk) : 0.f;
      Bs[kk][r] = (n0 + r < N && k0 + kk < K) ? __ldg(B + (size_t)(n0 + r) * ldb + k0 + kk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] * alpha;
      if (bias) v += bias[n];
      if (residual) v += residual[(size_t)m * ldr + n];
      Cm[(size_t)m * ldc + n] = v;
    }
  }
}

}  // namespace skp

using namespace skp;

extern "C" int skp_cross_attn_fwd(const float* q, const float* k, int64_t ldk, const float* v, int64_t ldv, float* o,
                                  float* logits, int S, int N, int heads, int d, float scale, void* stream) {
  SKP_REQUIRE(q && k && v && o && logits, "cross_attn_fwd: null pointer");
  SKP_REQUIRE(S > 0 && N > 0 && heads > 0 && d > 0, "cross_attn_fwd: bad sizes");
  SKP_REQUIRE(d <= 32 * AT_ND, "cross_attn_fwd: head dim %d > %d unsupported", d, 32 * AT_ND);
  size_t smem = ((size_t)AT_ROWS * d + (size_t)AT_TN * (d + 1) + (size_t)AT_ROWS * (N | 1)) * sizeof(float);
  SKP_REQUIRE(smem <= 220 * 1024, "cross_attn_fwd: N=%d too large for the shared-memory row buffer", N);
  cudaFuncSetAttribute(cross_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((S + AT_ROWS - 1) / AT_ROWS, heads);
  cross_attn_fwd_kernel<<<grid, AT_THREADS, smem, (cudaStream_t)stream>>>(q, k, ldk, v, ldv, o, logits, S, N, heads, d, scale);
  SKP_CHECK_LAUNCH("cross_attn_fwd");
  return SKP_OK;
}

extern "C" int skp_cross_attn_bwd(const float* d_o, const float* q, const float* k, int64_t ldk, const float* v, int64_t ldv,
                                  const float* logits, const float* d_logits_extra, float* ds_ws, float* dq, float* dk,
                                  float* dv, int S, int N, int heads, int d, float scale, void* stream) {
  SKP_REQUIRE(d_o && q && k && v && logits && ds_ws && dq && dk && dv, "cross_attn_bwd: null pointer");
  SKP_REQUIRE(S > 0 && N > 0 && heads > 0 && d > 0, "cross_attn_bwd: bad sizes");
  SKP_REQUIRE(d <= 32 * AT_ND, "cross_attn_bwd: head dim %d > %d unsupported", d, 32 * AT_ND);
  cudaStream_t st = (cudaStream_t)stream;
  float* stats = ds_ws + (size_t)heads * S * N;  // ds_ws holds heads*S*(N+2) floats
  size_t smem1 = ((size_t)AT_ROWS * d + (size_t)AT_TN * (d + 1) + 2 * (size_t)AT_ROWS * (N | 1)) * sizeof(float);
  SKP_REQUIRE(smem1 <= 220 * 1024, "cross_attn_bwd: N=%d too large for the shared-memory row buffers", N);
  cudaFuncSetAttribute(cross_attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
  dim3 g1((S + AT_ROWS - 1) / AT_ROWS, heads);
  cross_attn_bwd_dq_kernel<<<g1, AT_THREADS, smem1, st>>>(d_o, k, ldk, v, ldv, logits, d_logits_extra, ds_ws, stats, dq, S, N,
                                                         heads, d, scale);
  SKP_CHECK_LAUNCH("cross_attn_bwd_dq");
  size_t smem2 = (2 * (size_t)KV_SROWS * d + 2 * (size_t)KV_SROWS * 33) * sizeof(float);
  cudaFuncSetAttribute(cross_attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
  dim3 g2((N + KV_TOK - 1) / KV_TOK, heads, (S + KV_SCHUNK - 1) / KV_SCHUNK);
  cross_attn_bwd_dkv_kernel<<<g2, AT_THREADS, smem2, st>>>(d_o, q, logits, ds_ws, stats, dk, dv, S, N, heads, d, scale);
  SKP_CHECK_LAUNCH("cross_attn_bwd_dkv");
  return SKP_OK;
}

extern "C" int skp_gemm_nt_simt(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M, int N,
                                int K, float alpha, const float* bias, const float* residual, int64_t ldr, void* stream) {
  SKP_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "gemm_nt_simt: bad arguments");
  dim3 grid((N + GN - 1) / GN, (M + GM - 1) / GM);
  gemm_nt_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, B, ldb, C, ldc, M, N, K, alpha, bias, residual, ldr);
  SKP_CHECK_LAUNCH("gemm_nt_simt");
  return SKP_OK;
}
