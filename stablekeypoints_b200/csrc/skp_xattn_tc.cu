// Cross-attention (and short-sequence self-attention) forward on the 5th-generation tensor cores: ptp_utils.py:480-506
//   out = softmax(q k^T * scale) v   per head,   q [Sq, h*d],  k, v [Skv, h*d]  (Skv = N learned tokens for attn2)
// with, for the captured layers (ptp_utils.py:508-538), the scaled logits [h, Sq, Skv] written from the same kernel -- they
// are what the attn-store kernel up-samples -- so the capture costs no second pass over q and k.
//
// Generalises skp_attn_tc.cu (S % 128 == 0, d <= 64) to any Sq / Skv and head dims up to 160:
//   * the head dim is cut into KCH chunks of 64 columns (one 128B-swizzled TMA box each): QK^T = KCH x 4 K-steps x 3 split terms
//   * keys come in tiles of 64 (Skv padded with zero rows; the pad columns of the last tile are masked to -inf)
//   * the output accumulator O [128 x DV] stays in tensor memory across the key tiles: when the running maximum moves, the
//     softmax warps rescale it in place (tcgen05.ld / mul / tcgen05.st) before the next PV MMA accumulates onto it, so head
//     dims of 160 do not need 160 accumulator registers per thread
//   * query rows beyond Sq inside a 128-row tile (Sq = 64 in the mid block) are computed and dropped.
// Same numerics contract as every contraction of the trunk: split-bf16 operands (hi.hi + hi.lo + lo.hi), fp32 accumulation.
//
// One CTA = 128 query rows of one head:
//   warp 0      TMA producer (Q once; per key tile K [64 keys][KCH*64] and V^T [DV][64 keys])
//   warp 1      MMA issuer   (S -> TMEM columns 0..63, O -> columns 64..64+DV)
//   warps 2..5  softmax      (thread = query row = TMEM lane): scores -> (logits out) -> online softmax in base 2 -> P as
//               split-bf16 straight into the K-major swizzled A-operand layout -> O rescale -> final normalise + store
#include "skp_tc.cuh"
#include <math_constants.h>

namespace skp {

constexpr int XA_BM = 128;
constexpr int XA_BN = 64;
constexpr int XA_THREADS = 192;

__device__ __forceinline__ void xa_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float xa_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void xa_named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void xa_tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void xa_tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// q fp32 [Sq, heads*d] -> Qp [2][heads*Sq][KCH*64]  (hi, lo; scaled by qscale; zero padded columns)
// k fp32 [Skv, heads*d] -> Kp [2][heads*NKP][KCH*64] (zero rows beyond Skv)
// v fp32 [Skv, heads*d] -> VTp [2][heads*DV][NKP]    (V transposed per head; zero beyond Skv / d)
__global__ void xa_tc_split_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk,
                                   const float* __restrict__ v, int64_t ldv, __nv_bfloat16* __restrict__ Qp,
                                   __nv_bfloat16* __restrict__ Kp, __nv_bfloat16* __restrict__ VTp, int Sq, int Skv, int NKP,
                                   int heads, int d, int DK, int DV, float qscale) {
  const int dk2 = DK >> 1;
  const long nq = (long)heads * Sq * dk2, nk = (long)heads * NKP * dk2;   // bf16 pairs
  const long nvt = (long)heads * DV * (NKP >> 1);
  const long total = nq + nk + nvt;
  const size_t q_plane = (size_t)heads * Sq * DK, k_plane = (size_t)heads * NKP * DK, vt_plane = (size_t)heads * DV * NKP;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    float x = 0.f, y = 0.f;
    __nv_bfloat16 *hi, *lo;
    if (i < nq + nk) {
      const bool isk = i >= nq;
      const long t = isk ? i - nq : i;
      const int c = (int)(t % dk2) << 1;
      const long hr = t / dk2;                        // h * rows + row
      const int rows = isk ? NKP : Sq;
      const int h = (int)(hr / rows), row = (int)(hr - (long)h * rows);
      if (!isk || row < Skv) {
        const float* src = (isk ? k + (size_t)row * ldk : q + (size_t)row * ldq) + h * d;
        x = c < d ? __ldg(src + c) : 0.f;
        y = c + 1 < d ? __ldg(src + c + 1) : 0.f;
        if (!isk) { x *= qscale; y *= qscale; }
      }
      hi = (isk ? Kp : Qp) + (size_t)hr * DK + c;
      lo = hi + (isk ? k_plane : q_plane);
    } else {
      // channel fastest: the reads of V rows are coalesced, the transposed 4-byte writes scatter (absorbed by L2)
      const long t = i - nq - nk;
      const int half = NKP >> 1;
      const int c = (int)(t % DV);
      const long hs = t / DV;                        // h * half + key pair
      const int h = (int)(hs / half), s2 = (int)(hs - (long)h * half) << 1;
      if (c < d) {
        if (s2 < Skv) x = __ldg(v + (size_t)s2 * ldv + h * d + c);
        if (s2 + 1 < Skv) y = __ldg(v + (size_t)(s2 + 1) * ldv + h * d + c);
      }
      hi = VTp + ((size_t)h * DV + c) * NKP + s2;
      lo = hi + vt_plane;
    }
    __nv_bfloat162 hh = __floats2bfloat162_rn(x, y);
    float2 f = __bfloat1622float2(hh);
    *reinterpret_cast<__nv_bfloat162*>(hi) = hh;
    *reinterpret_cast<__nv_bfloat162*>(lo) = __floats2bfloat162_rn(x - f.x, y - f.y);
  }
}

template <int KCH, int DV>
struct XaCfg {
  static constexpr int Q_PLANE = XA_BM * 128;            // one chunk, one plane
  static constexpr int K_PLANE = XA_BN * 128;
  static constexpr int V_PLANE = DV * 128;               // DV channel rows x 64 keys
  static constexpr int P_PLANE = XA_BM * 128;
  static constexpr int Q_BYTES = KCH * 2 * Q_PLANE, K_BYTES = KCH * 2 * K_PLANE, V_BYTES = 2 * V_PLANE, P_BYTES = 2 * P_PLANE;
  static constexpr int SMEM = Q_BYTES + K_BYTES + V_BYTES + P_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = (64 + DV) <= 128 ? 128 : 256;
  static_assert(DV % 16 == 0 && DV <= 160, "DV: multiple of 16 up to 160");
  static_assert(SMEM <= 227 * 1024, "tile set does not fit shared memory");
};

template <int KCH, int DV>
__global__ void __launch_bounds__(XA_THREADS, 1)
xa_tc_fwd_kernel(const __grid_constant__ CUtensorMap tm_q_hi, const __grid_constant__ CUtensorMap tm_q_lo,
                 const __grid_constant__ CUtensorMap tm_k_hi, const __grid_constant__ CUtensorMap tm_k_lo,
                 const __grid_constant__ CUtensorMap tm_v_hi, const __grid_constant__ CUtensorMap tm_v_lo,
                 float* __restrict__ out, int64_t ldo, float* __restrict__ lse, float* __restrict__ logits, int Sq, int Skv,
                 int NKP, int d) {
  using Cfg = XaCfg<KCH, DV>;
  extern __shared__ uint8_t xa_smem_raw[];
  const uint32_t raw = smem_u32(xa_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = xa_smem_raw + (base - raw);
  const uint32_t sQ = base, sK = sQ + Cfg::Q_BYTES, sV = sK + Cfg::K_BYTES, sP = sV + Cfg::V_BYTES;
  const uint32_t bars = sP + Cfg::P_BYTES;
  enum { Q_FULL = 0, K_FULL, K_EMPTY, V_FULL, V_EMPTY, S_FULL, P_FULL, O_FULL, O_READY, NBARS };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + (bars - base) + 8 * NBARS);
  uint8_t* gPh = gen + (sP - base);
  uint8_t* gPl = gPh + Cfg::P_PLANE;
  float* scratch = reinterpret_cast<float*>(gPh);          // [128][64] fp32 (the P region, free while the scores are read)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, q0 = blockIdx.x * XA_BM;
  const int ntiles = NKP / XA_BN;

  if (warp == 0 && lane == 0) {
    for (int b = 0; b < NBARS; ++b) mbar_init(bars + 8 * b, (b == P_FULL || b == O_READY) ? 128u : 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_expect_tx(bars + 8 * Q_FULL, Cfg::Q_BYTES);
      for (int c = 0; c < KCH; ++c) {
        tma_load_2d(sQ + c * 2 * Cfg::Q_PLANE, &tm_q_hi, bars + 8 * Q_FULL, c * 64, h * Sq + q0);
        tma_load_2d(sQ + c * 2 * Cfg::Q_PLANE + Cfg::Q_PLANE, &tm_q_lo, bars + 8 * Q_FULL, c * 64, h * Sq + q0);
      }
      for (int j = 0; j < ntiles; ++j) {
        const uint32_t ph = (uint32_t)j & 1u;
        mbar_wait(bars + 8 * K_EMPTY, ph ^ 1u);
        mbar_expect_tx(bars + 8 * K_FULL, Cfg::K_BYTES);
        for (int c = 0; c < KCH; ++c) {
          tma_load_2d(sK + c * 2 * Cfg::K_PLANE, &tm_k_hi, bars + 8 * K_FULL, c * 64, h * NKP + j * XA_BN);
          tma_load_2d(sK + c * 2 * Cfg::K_PLANE + Cfg::K_PLANE, &tm_k_lo, bars + 8 * K_FULL, c * 64, h * NKP + j * XA_BN);
        }
        mbar_wait(bars + 8 * V_EMPTY, ph ^ 1u);
        mbar_expect_tx(bars + 8 * V_FULL, Cfg::V_BYTES);
        tma_load_2d(sV, &tm_v_hi, bars + 8 * V_FULL, j * XA_BN, h * DV);
        tma_load_2d(sV + Cfg::V_PLANE, &tm_v_lo, bars + 8 * V_FULL, j * XA_BN, h * DV);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(XA_BN >> 3) << 17) | ((uint32_t)(XA_BM >> 4) << 24);
      constexpr uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(DV >> 3) << 17) | ((uint32_t)(XA_BM >> 4) << 24);
      const uint64_t dVh = make_smem_desc(sV), dVl = make_smem_desc(sV + Cfg::V_PLANE);
      const uint64_t dPh = make_smem_desc(sP), dPl = make_smem_desc(sP + Cfg::P_PLANE);
      mbar_wait(bars + 8 * Q_FULL, 0);
      for (int j = 0; j < ntiles; ++j) {
        const uint32_t ph = (uint32_t)j & 1u;
        mbar_wait(bars + 8 * K_FULL, ph);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < KCH; ++c) {
          const uint64_t dQh = make_smem_desc(sQ + c * 2 * Cfg::Q_PLANE), dQl = make_smem_desc(sQ + c * 2 * Cfg::Q_PLANE + Cfg::Q_PLANE);
          const uint64_t dKh = make_smem_desc(sK + c * 2 * Cfg::K_PLANE), dKl = make_smem_desc(sK + c * 2 * Cfg::K_PLANE + Cfg::K_PLANE);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);
            umma_bf16(tmem, dQl + adv, dKh + adv, idesc_s, (c | k) != 0);
            umma_bf16(tmem, dQh + adv, dKl + adv, idesc_s, 1u);
            umma_bf16(tmem, dQh + adv, dKh + adv, idesc_s, 1u);
          }
        }
        umma_commit(bars + 8 * K_EMPTY);   // K tile free once these MMAs retire
        umma_commit(bars + 8 * S_FULL);    // ... and the scores are complete
        mbar_wait(bars + 8 * P_FULL, ph);  // P of this tile is in shared memory (and S has been consumed)
        mbar_wait(bars + 8 * V_FULL, ph);
        mbar_wait(bars + 8 * O_READY, ph); // O has been rescaled for this tile's running maximum
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < XA_BN / 16; ++k) {
          const uint64_t adv = (uint64_t)((k * 32) >> 4);
          umma_bf16(tmem + 64, dPl + adv, dVh + adv, idesc_o, (j | k) != 0);
          umma_bf16(tmem + 64, dPh + adv, dVl + adv, idesc_o, 1u);
          umma_bf16(tmem + 64, dPh + adv, dVh + adv, idesc_o, 1u);
        }
        umma_commit(bars + 8 * V_EMPTY);
        umma_commit(bars + 8 * O_FULL);
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;            // query row inside the tile
    const int row = q0 + r;
    const bool live = row < Sq;
    const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16);
    const uint32_t prow = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;   // K-major 128B-swizzled row
    const uint32_t sw = (uint32_t)(r & 7);
    float m = -CUDART_INF_F, l = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      const uint32_t ph = (uint32_t)j & 1u;
      mbar_wait(bars + 8 * S_FULL, ph);
      tc_fence_after();
      float s0[32], s1[32];
      tmem_ld32(trow, s0);
      tmem_ld32(trow + 32, s1);
      const int nvalid = min(XA_BN, Skv - j * XA_BN);   // keys of this tile that exist (>= 1)
      if (logits != nullptr) {
        // captured layer: scaled logits (natural units) out, row-major [h, Sq, Skv]; transposed through the (free) P region
        // so that the global stores are coalesced along the keys
        const float LN2 = 0.6931471805599453f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          scratch[r * 64 + (i ^ (r & 31))] = s0[i] * LN2;
          scratch[r * 64 + 32 + (i ^ (r & 31))] = s1[i] * LN2;
        }
        xa_named_bar(1, 128);
        const int rw = quarter * 32;
        for (int rr = 0; rr < 32; ++rr) {
          const int r2 = rw + rr, row2 = q0 + r2;
          if (row2 < Sq) {
            float* dst = logits + ((size_t)h * Sq + row2) * Skv + j * XA_BN;
            if (lane < nvalid) dst[lane] = scratch[r2 * 64 + (lane ^ (r2 & 31))];
            if (lane + 32 < nvalid) dst[lane + 32] = scratch[r2 * 64 + 32 + (lane ^ (r2 & 31))];
          }
        }
        xa_named_bar(1, 128);
      }
      float t = -CUDART_INF_F;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (i >= nvalid) s0[i] = -CUDART_INF_F;
        if (i + 32 >= nvalid) s1[i] = -CUDART_INF_F;
        t = fmaxf(t, fmaxf(s0[i], s1[i]));
      }
      const float mn = fmaxf(m, t);
      const float alpha = xa_ex2(m - mn);
      m = mn;
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        s0[i] = xa_ex2(s0[i] - mn);
        s1[i] = xa_ex2(s1[i] - mn);
        sum += s0[i] + s1[i];
      }
      l = l * alpha + sum;
      // P row -> split bf16, 16-byte chunks of 8 keys at chunk position (c ^ (row & 7)) of the 128-byte row
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float* src = c < 4 ? s0 + 8 * c : s1 + 8 * (c - 4);
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          __nv_bfloat162 hh = __floats2bfloat162_rn(src[2 * e], src[2 * e + 1]);
          float2 f = __bfloat1622float2(hh);
          __nv_bfloat162 ll = __floats2bfloat162_rn(src[2 * e] - f.x, src[2 * e + 1] - f.y);
          hw[e] = *reinterpret_cast<uint32_t*>(&hh);
          lw[e] = *reinterpret_cast<uint32_t*>(&ll);
        }
        const uint32_t off = prow + (((uint32_t)c ^ sw) << 4);
        *reinterpret_cast<uint4*>(gPh + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(gPl + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
      tc_fence_before();
      xa_mbar_arrive(bars + 8 * P_FULL);
      // rescale the accumulator (TMEM resident) for the new running maximum before PV of this tile accumulates onto it
      if (j > 0) {
        mbar_wait(bars + 8 * O_FULL, ph ^ 1u);     // PV of the previous tile has retired
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < DV / 16; ++c) {
          uint32_t rr[16];
          xa_tmem_ld16(trow + 64 + 16 * c, rr);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 16; ++i) rr[i] = __float_as_uint(__uint_as_float(rr[i]) * alpha);
          xa_tmem_st16(trow + 64 + 16 * c, rr);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
      }
      xa_mbar_arrive(bars + 8 * O_READY);
    }
    // final: O / l -> out, log-sum-exp (base 2) -> lse
    mbar_wait(bars + 8 * O_FULL, (uint32_t)(ntiles - 1) & 1u);
    tc_fence_after();
    const float inv = 1.f / l;
    float* orow = out + (size_t)row * ldo + h * d;
#pragma unroll
    for (int c = 0; c < DV / 16; ++c) {
      uint32_t rr[16];
      xa_tmem_ld16(trow + 64 + 16 * c, rr);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (live) {
#pragma unroll
        for (int i = 0; i < 16; i += 2)
          if (16 * c + i < d)
            *reinterpret_cast<float2*>(orow + 16 * c + i) = make_float2(__uint_as_float(rr[i]) * inv, __uint_as_float(rr[i + 1]) * inv);
      }
    }
    if (live) lse[(size_t)h * Sq + row] = m + log2f(l);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS));
  }
}

static int xa_dv(int d) { return (d + 15) & ~15; }
static int xa_kch(int d) { return (d + 63) / 64; }

template <int KCH, int DV>
static int xa_launch(const __nv_bfloat16* Qp, const __nv_bfloat16* Kp, const __nv_bfloat16* VTp, float* o, int64_t ldo, float* lse,
                     float* logits, int Sq, int Skv, int NKP, int heads, int d, cudaStream_t st) {
  using Cfg = XaCfg<KCH, DV>;
  const int DK = KCH * 64;
  CUtensorMap tq_hi, tq_lo, tk_hi, tk_lo, tv_hi, tv_lo;
  const size_t q_plane = (size_t)heads * Sq * DK, k_plane = (size_t)heads * NKP * DK, vt_plane = (size_t)heads * DV * NKP;
  int rc;
  if ((rc = tc_make_map(&tq_hi, Qp, heads * Sq, DK, XA_BM))) return rc;
  if ((rc = tc_make_map(&tq_lo, Qp + q_plane, heads * Sq, DK, XA_BM))) return rc;
  if ((rc = tc_make_map(&tk_hi, Kp, heads * NKP, DK, XA_BN))) return rc;
  if ((rc = tc_make_map(&tk_lo, Kp + k_plane, heads * NKP, DK, XA_BN))) return rc;
  if ((rc = tc_make_map(&tv_hi, VTp, heads * DV, NKP, DV))) return rc;
  if ((rc = tc_make_map(&tv_lo, VTp + vt_plane, heads * DV, NKP, DV))) return rc;
  cudaError_t e = cudaFuncSetAttribute(xa_tc_fwd_kernel<KCH, DV>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  if (e != cudaSuccess) { set_error("xattn_tc_fwd: smem attr: %s", cudaGetErrorString(e)); return SKP_ERR_LAUNCH; }
  dim3 grid((Sq + XA_BM - 1) / XA_BM, heads);
  xa_tc_fwd_kernel<KCH, DV><<<grid, XA_THREADS, Cfg::SMEM, st>>>(tq_hi, tq_lo, tk_hi, tk_lo, tv_hi, tv_lo, o, ldo, lse, logits, Sq, Skv,
                                                                   NKP, d);
  SKP_CHECK_LAUNCH("xa_tc_fwd_kernel");
  return SKP_OK;
}

}  // namespace skp

using namespace skp;

// workspace bytes for skp_xattn_tc_fwd: split-bf16 operand planes of q, k and v^T
extern "C" int64_t skp_xattn_tc_workspace(int Sq, int Skv, int heads, int d) {
  if (Sq <= 0 || Skv <= 0 || heads <= 0 || d <= 0 || d > 160 || (d & 1)) return 0;
  const int64_t DK = xa_kch(d) * 64, DV = xa_dv(d), NKP = (Skv + 63) / 64 * 64;
  return ((int64_t)2 * heads * Sq * DK + (int64_t)2 * heads * NKP * DK + (int64_t)2 * heads * DV * NKP) * 2;
}

extern "C" int skp_xattn_tc_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv, float* o,
                                int64_t ldo, float* lse, float* logits, void* workspace, int Sq, int Skv, int heads, int d,
                                float scale, void* stream) {
  SKP_REQUIRE(q && k && v && o && lse && workspace, "skp_xattn_tc_fwd: null pointer");
  SKP_REQUIRE(skp_xattn_tc_workspace(Sq, Skv, heads, d) > 0, "skp_xattn_tc_fwd: needs an even head dim <= 160 (d=%d)", d);
  SKP_REQUIRE(ldo % 2 == 0 && (reinterpret_cast<uintptr_t>(o) & 7) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 127) == 0,
              "skp_xattn_tc_fwd: o must be 8-byte aligned with even ld, the workspace 128-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int KCH = xa_kch(d), DK = KCH * 64, DV = xa_dv(d), NKP = (Skv + 63) / 64 * 64;
  __nv_bfloat16* Qp = (__nv_bfloat16*)workspace;
  __nv_bfloat16* Kp = Qp + (size_t)2 * heads * Sq * DK;
  __nv_bfloat16* VTp = Kp + (size_t)2 * heads * NKP * DK;
  const long total = (long)heads * Sq * (DK / 2) + (long)heads * NKP * (DK / 2) + (long)heads * DV * (NKP / 2);
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  xa_tc_split_kernel<<<(int)blocks, 256, 0, st>>>(q, ldq, k, ldk, v, ldv, Qp, Kp, VTp, Sq, Skv, NKP, heads, d, DK, DV,
                                                  scale * 1.4426950408889634f);
  SKP_CHECK_LAUNCH("xa_tc_split_kernel");
#define XA_CASE(K_, D_) \
  if (KCH == K_ && DV == D_) return xa_launch<K_, D_>(Qp, Kp, VTp, o, ldo, lse, logits, Sq, Skv, NKP, heads, d, st);
  XA_CASE(1, 16) XA_CASE(1, 32) XA_CASE(1, 48) XA_CASE(1, 64)
  XA_CASE(2, 80) XA_CASE(2, 96) XA_CASE(2, 112) XA_CASE(2, 128)
  XA_CASE(3, 144) XA_CASE(3, 160)
#undef XA_CASE
  set_error("skp_xattn_tc_fwd: head dim %d has no instantiation", d);
  return SKP_ERR_UNSUPPORTED;
}
