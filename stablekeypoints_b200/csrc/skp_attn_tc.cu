// Self-attention forward on the 5th-generation tensor cores (tcgen05 + TMEM + TMA) for the long-sequence layers of the
// frozen UNet (attn1 of the 64x64 blocks: S = 4096 tokens, 8 heads x d = 40; ptp_utils.py:480-506 with context=None).
// Same numerics contract as skp_selfattn.cu -- every contraction is a split-bf16 product (hi.hi + hi.lo + lo.hi, fp32
// accumulate) -- but QK^T and PV run as tcgen05.mma with the accumulators in tensor memory, three times the issue rate
// of the warp-level mma.sync path (profiles/r01_flash_attn.md).
//
// One CTA = 128 query rows of one head, looping over tiles of 64 keys (two CTAs per SM: one's softmax overlaps the
// other's MMAs):
//   warp 0      TMA producer: Q tile once; per tile the K tile [64 keys][64] and the V^T tile [DV channels][64 keys]
//               (cp.async.bulk.tensor, 128B swizzle, mbarrier expect_tx).  K is re-fetched as soon as QK^T of the tile
//               retired and V^T as soon as PV retired, so both loads hide behind the softmax of the tile in between.
//   warp 1      MMA issuer: S[128x64] = Q' K^T (M128 N64 K16 x ceil(d/16) steps x 3 split terms) into TMEM columns 0..63;
//               after the softmax warps published P: PV[128xDV] = P V (M128 N=DV K16 x 4 steps x 3 terms) into columns 64..
//   warps 2..5  softmax: thread = query row (its TMEM lane); tcgen05.ld the 64 scores, online softmax in base 2, write
//               P as split-bf16 straight into the K-major 128B-swizzled layout the MMA reads as its A operand,
//               the PV accumulator stays in TMEM across the key tiles (lazy rescale, see the softmax loop).
// Operands come pre-split from sa_tc_split_kernel: Q' (scaled by scale*log2 e) and K as [heads*S][64] planes, V
// transposed as [heads*DV][S] planes (so that P V is an ordinary K-major x K-major product).
// Eligibility (host): S % 128 == 0, d even and <= 64.  Everything else stays on the mma.sync kernels.
#include "skp_tc.cuh"
#include <math_constants.h>
#include <stdlib.h>

namespace skp {

constexpr int FT_BM = 128;       // query rows per CTA
constexpr int FT_BN = 64;        // keys per tile
constexpr int FT_THREADS = 224;  // TMA warp (Q, K; + V in the first form), MMA warp, 4 softmax warps, V TMA warp (pipelined form)
constexpr int FT_Q_BYTES = FT_BM * 128;   // one bf16 plane of the Q tile (64 columns = 128 B per row)
constexpr int FT_K_BYTES = FT_BN * 128;
constexpr int FT_V_BYTES = 64 * 128;      // up to 64 channel rows x 64 keys
constexpr int FT_P_BYTES = FT_BM * 128;
constexpr int FT_SMEM = 2 * FT_Q_BYTES + 2 * FT_K_BYTES + 2 * FT_V_BYTES + 2 * FT_P_BYTES + 1024 + 256;
constexpr int FT_TMEM_COLS = 128;         // first form: S columns 0..63, PV columns 64..64+DV
constexpr int FT_TMEM_COLS_PIPE = 256;    // pipelined form: two score buffers 0..63 | 64..127, O at 128..128+DV

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float ft_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// q,k,v fp32 [S, heads*d] (leading dims) -> Qp,Kp: [2][heads*S][64] bf16 (hi, lo; zero padded, Q scaled);
// VTp: [2][heads*DV][S] bf16 (V transposed per head).
__global__ void sa_tc_split_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk,
                                   const float* __restrict__ v, int64_t ldv, __nv_bfloat16* __restrict__ Qp,
                                   __nv_bfloat16* __restrict__ Kp, __nv_bfloat16* __restrict__ VTp, int S, int heads, int d,
                                   int DV, float qscale) {
  const long nqk = (long)heads * S * 32;            // bf16 pairs of one [heads*S][64] plane
  const long nvt = (long)heads * DV * (S / 2);      // bf16 pairs of one [heads*DV][S] plane
  const long total = 2 * nqk + nvt;
  const size_t qk_plane = (size_t)heads * S * 64, vt_plane = (size_t)heads * DV * S;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    float x, y;
    __nv_bfloat16 *hi, *lo;
    if (i < 2 * nqk) {
      const bool isk = i >= nqk;
      const long t = isk ? i - nqk : i;
      const int c = (int)(t & 31) << 1;
      const long hr = t >> 5;                        // h * S + row
      const int h = (int)(hr / S), row = (int)(hr - (long)h * S);
      const float* src = (isk ? k + (size_t)row * ldk : q + (size_t)row * ldq) + h * d;
      x = c < d ? __ldg(src + c) : 0.f;
      y = c + 1 < d ? __ldg(src + c + 1) : 0.f;
      if (!isk) { x *= qscale; y *= qscale; }
      hi = (isk ? Kp : Qp) + (size_t)hr * 64 + c;
      lo = hi + qk_plane;
    } else {
      // channel fastest: the reads of V rows are coalesced, the transposed 4-byte writes scatter (absorbed by L2)
      const long t = i - 2 * nqk;
      const int half = S >> 1;
      const int c = (int)(t % DV);
      const long hs = t / DV;                        // h * half + key pair
      const int h = (int)(hs / half), s2 = (int)(hs - (long)h * half) << 1;
      x = c < d ? __ldg(v + (size_t)s2 * ldv + h * d + c) : 0.f;
      y = c < d ? __ldg(v + (size_t)(s2 + 1) * ldv + h * d + c) : 0.f;
      hi = VTp + ((size_t)h * DV + c) * S + s2;
      lo = hi + vt_plane;
    }
    __nv_bfloat162 hh = __floats2bfloat162_rn(x, y);
    float2 f = __bfloat1622float2(hh);
    *reinterpret_cast<__nv_bfloat162*>(hi) = hh;
    *reinterpret_cast<__nv_bfloat162*>(lo) = __floats2bfloat162_rn(x - f.x, y - f.y);
  }
}

template <int DV, bool OTMEM>
__global__ void __launch_bounds__(FT_THREADS, 2)
sa_tc_fwd_kernel(const __grid_constant__ CUtensorMap tm_q_hi, const __grid_constant__ CUtensorMap tm_q_lo,
                 const __grid_constant__ CUtensorMap tm_k_hi, const __grid_constant__ CUtensorMap tm_k_lo,
                 const __grid_constant__ CUtensorMap tm_v_hi, const __grid_constant__ CUtensorMap tm_v_lo,
                 float* __restrict__ out, int64_t ldo, float* __restrict__ lse, int S, int d, int ksteps) {
  extern __shared__ uint8_t ft_smem_raw[];
  const uint32_t raw = smem_u32(ft_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = ft_smem_raw + (base - raw);
  const uint32_t sQh = base, sQl = sQh + FT_Q_BYTES, sKh = sQl + FT_Q_BYTES, sKl = sKh + FT_K_BYTES;
  const uint32_t sVh = sKl + FT_K_BYTES, sVl = sVh + FT_V_BYTES, sPh = sVl + FT_V_BYTES, sPl = sPh + FT_P_BYTES;
  const uint32_t bars = sPl + FT_P_BYTES;
  // barriers: 0 q_full, 1 k_full, 2 k_empty, 3 v_full, 4 v_empty, 5 s_full, 6 p_full (128), 7 o_full, 8 o_empty (128)
  // pipelined form: S_FULL / P_FULL exist per score buffer (index + (j & 1), phase (j >> 1) & 1)
  enum { Q_FULL = 0, K_FULL, K_EMPTY, V_FULL, V_EMPTY, S_FULL, S_FULL1, P_FULL, P_FULL1, O_FULL, O_EMPTY, NBARS };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + (bars - base) + 8 * NBARS);
  uint8_t* gPh = gen + (sPh - base);
  uint8_t* gPl = gen + (sPl - base);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, q0 = blockIdx.x * FT_BM;
  const int ntiles = S / FT_BN;

  if (warp == 0 && lane == 0) {
    for (int b = 0; b < NBARS; ++b) mbar_init(bars + 8 * b, (b == P_FULL || b == P_FULL1 || b == O_EMPTY) ? 128u : 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(OTMEM ? FT_TMEM_COLS_PIPE : FT_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_expect_tx(bars + 8 * Q_FULL, 2 * FT_Q_BYTES);
      tma_load_2d(sQh, &tm_q_hi, bars + 8 * Q_FULL, 0, h * S + q0);
      tma_load_2d(sQl, &tm_q_lo, bars + 8 * Q_FULL, 0, h * S + q0);
      for (int j = 0; j < ntiles; ++j) {
        const uint32_t ph = (uint32_t)j & 1u;
        mbar_wait(bars + 8 * K_EMPTY, ph ^ 1u);
        mbar_expect_tx(bars + 8 * K_FULL, 2 * FT_K_BYTES);
        tma_load_2d(sKh, &tm_k_hi, bars + 8 * K_FULL, 0, h * S + j * FT_BN);
        tma_load_2d(sKl, &tm_k_lo, bars + 8 * K_FULL, 0, h * S + j * FT_BN);
        if (OTMEM) continue;                      // pipelined form: V^T tiles come from their own warp, so that the K tile of
                                                  // tile j + 1 does not queue behind PV of tile j - 1
        mbar_wait(bars + 8 * V_EMPTY, ph ^ 1u);
        mbar_expect_tx(bars + 8 * V_FULL, 2 * DV * 128);
        tma_load_2d(sVh, &tm_v_hi, bars + 8 * V_FULL, j * FT_BN, h * DV);
        tma_load_2d(sVl, &tm_v_lo, bars + 8 * V_FULL, j * FT_BN, h * DV);
      }
    }
  } else if (warp == 6) {
    if (OTMEM && elect_one_sync()) {
      for (int j = 0; j < ntiles; ++j) {
        mbar_wait(bars + 8 * V_EMPTY, ((uint32_t)j & 1u) ^ 1u);
        mbar_expect_tx(bars + 8 * V_FULL, 2 * DV * 128);
        tma_load_2d(sVh, &tm_v_hi, bars + 8 * V_FULL, j * FT_BN, h * DV);
        tma_load_2d(sVl, &tm_v_lo, bars + 8 * V_FULL, j * FT_BN, h * DV);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      // instruction descriptors: D = f32, A = B = bf16, both K-major, M = 128, N = 64 (scores) / DV (output)
      constexpr uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FT_BN >> 3) << 17) | ((uint32_t)(FT_BM >> 4) << 24);
      constexpr uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(DV >> 3) << 17) | ((uint32_t)(FT_BM >> 4) << 24);
      const uint64_t dQh = make_smem_desc(sQh), dQl = make_smem_desc(sQl), dKh = make_smem_desc(sKh), dKl = make_smem_desc(sKl);
      const uint64_t dVh = make_smem_desc(sVh), dVl = make_smem_desc(sVl), dPh = make_smem_desc(sPh), dPl = make_smem_desc(sPl);
      mbar_wait(bars + 8 * Q_FULL, 0);
      if constexpr (OTMEM) {
        // software-pipelined issue: the scores of tile j + 2 go out right behind PV of tile j, so they are ready when the
        // softmax warps finish tile j + 1 (two score buffers in tensor memory)
        auto issue_scores = [&](int j) {
          const uint32_t ts = tmem + 64u * ((uint32_t)j & 1u);
          mbar_wait(bars + 8 * K_FULL, (uint32_t)j & 1u);
          tc_fence_after();
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);
            umma_bf16(ts, dQl + adv, dKh + adv, idesc_s, k != 0);
            umma_bf16(ts, dQh + adv, dKl + adv, idesc_s, 1u);
            umma_bf16(ts, dQh + adv, dKh + adv, idesc_s, 1u);
          }
          umma_commit(bars + 8 * K_EMPTY);
          umma_commit(bars + 8 * (S_FULL + (j & 1)));
        };
        issue_scores(0);
        if (ntiles > 1) issue_scores(1);
        for (int j = 0; j < ntiles; ++j) {
          mbar_wait(bars + 8 * (P_FULL + (j & 1)), ((uint32_t)j >> 1) & 1u);
          mbar_wait(bars + 8 * V_FULL, (uint32_t)j & 1u);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < FT_BN / 16; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);
            umma_bf16(tmem + 128, dPl + adv, dVh + adv, idesc_o, (j | k) != 0);   // O stays in TMEM across the key tiles
            umma_bf16(tmem + 128, dPh + adv, dVl + adv, idesc_o, 1u);
            umma_bf16(tmem + 128, dPh + adv, dVh + adv, idesc_o, 1u);
          }
          umma_commit(bars + 8 * V_EMPTY);
          umma_commit(bars + 8 * O_FULL);
          if (j + 2 < ntiles) issue_scores(j + 2);   // its buffer was read by the softmax warps before P_FULL of tile j
        }
      } else
      for (int j = 0; j < ntiles; ++j) {
        const uint32_t ph = (uint32_t)j & 1u;
        mbar_wait(bars + 8 * K_FULL, ph);
        tc_fence_after();
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t adv = (uint64_t)((k * 32) >> 4);
          umma_bf16(tmem, dQl + adv, dKh + adv, idesc_s, k != 0);
          umma_bf16(tmem, dQh + adv, dKl + adv, idesc_s, 1u);
          umma_bf16(tmem, dQh + adv, dKh + adv, idesc_s, 1u);
        }
        umma_commit(bars + 8 * K_EMPTY);   // K tile free once these MMAs retire
        umma_commit(bars + 8 * S_FULL);    // ... and the scores are complete
        mbar_wait(bars + 8 * P_FULL, ph);  // P of this tile is in shared memory (and S has been consumed)
        mbar_wait(bars + 8 * V_FULL, ph);
        mbar_wait(bars + 8 * O_EMPTY, ph ^ 1u);   // the previous tile's PV has been read out of TMEM
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < FT_BN / 16; ++k) {
          const uint64_t adv = (uint64_t)((k * 32) >> 4);
          umma_bf16(tmem + 64, dPl + adv, dVh + adv, idesc_o, k != 0);
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
    const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16);
    const uint32_t prow = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;   // K-major 128B-swizzled row
    const uint32_t sw = (uint32_t)(r & 7);
    float m = -CUDART_INF_F, l = 0.f;
    const int row = q0 + r;
    float* orow = out + (size_t)row * ldo + h * d;
    if constexpr (OTMEM) {
    // Pipelined form.  The output accumulator stays in tensor memory for the whole key loop: the exponent offset m lags the
    // running row maximum by up to 8 (probabilities up to 2^8: harmless in fp32 / split-bf16), and only when a row's maximum
    // exceeds m + 8 does its warp pull the accumulator out, rescale it and put it back (in practice in the first tiles only).
    for (int j = 0; j < ntiles; ++j) {
      const uint32_t sb = (uint32_t)j & 1u;
      const uint32_t ts = trow + 64u * sb;
      mbar_wait(bars + 8 * (S_FULL + sb), ((uint32_t)j >> 1) & 1u);
      tc_fence_after();
      float s0[32], s1[32];
      tmem_ld32(ts, s0);
      tmem_ld32(ts + 32, s1);
      float t = -CUDART_INF_F;
#pragma unroll
      for (int i = 0; i < 32; ++i) t = fmaxf(t, fmaxf(s0[i], s1[i]));
      if (j == 0) {
        m = t;
      } else if (__any_sync(0xffffffffu, t > m + 8.f)) {
        const bool up = t > m + 8.f;
        const float alpha = up ? ft_ex2(m - t) : 1.f;
        if (up) m = t;
        l *= alpha;
        mbar_wait(bars + 8 * O_FULL, ((uint32_t)j & 1u) ^ 1u);          // PV of tile j - 1 has landed
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < DV / 16; ++c) {
          uint32_t rr[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]), "=r"(rr[8]),
                "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15])
              : "r"(trow + 128 + 16 * c));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 16; ++i) rr[i] = __float_as_uint(__uint_as_float(rr[i]) * alpha);
          asm volatile(
              "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
                  trow + 128 + 16 * c),
              "r"(rr[0]), "r"(rr[1]), "r"(rr[2]), "r"(rr[3]), "r"(rr[4]), "r"(rr[5]), "r"(rr[6]), "r"(rr[7]), "r"(rr[8]), "r"(rr[9]),
              "r"(rr[10]), "r"(rr[11]), "r"(rr[12]), "r"(rr[13]), "r"(rr[14]), "r"(rr[15])
              : "memory");
        }
      }
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        s0[i] = ft_ex2(s0[i] - m);
        s1[i] = ft_ex2(s1[i] - m);
        sum += s0[i] + s1[i];
      }
      l += sum;
      // P row -> split bf16, 16-byte chunks of 8 keys at chunk position (c ^ (row & 7)) of the 128-byte staging row.
      // (Handing P back through tensor memory as the A operand of PV -- the form the backward kernel uses -- produced wrong rows
      // here whenever two CTAs shared an SM, and was not pursued.)
      uint32_t hw[32], lw[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const float x0 = e < 16 ? s0[2 * e] : s1[2 * e - 32], x1 = e < 16 ? s0[2 * e + 1] : s1[2 * e - 31];
        __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
        float2 f = __bfloat1622float2(hh);
        __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - f.x, x1 - f.y);
        hw[e] = *reinterpret_cast<uint32_t*>(&hh);
        lw[e] = *reinterpret_cast<uint32_t*>(&ll);
      }
      if (j > 0) mbar_wait(bars + 8 * O_FULL, ((uint32_t)j & 1u) ^ 1u);   // PV of tile j - 1 has read the staging tile
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t off = prow + (((uint32_t)c ^ sw) << 4);
        *reinterpret_cast<uint4*>(gPh + off) = make_uint4(hw[4 * c], hw[4 * c + 1], hw[4 * c + 2], hw[4 * c + 3]);
        *reinterpret_cast<uint4*>(gPl + off) = make_uint4(lw[4 * c], lw[4 * c + 1], lw[4 * c + 2], lw[4 * c + 3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");   // (the rescale path's stores)
      tc_fence_before();
      mbar_arrive(bars + 8 * (P_FULL + sb));
    }
    // ---- the accumulator leaves tensor memory once
    mbar_wait(bars + 8 * O_FULL, (uint32_t)(ntiles - 1) & 1u);
    tc_fence_after();
    const float inv = 1.f / l;
#pragma unroll
    for (int c = 0; c < DV / 16; ++c) {
      uint32_t rr[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]), "=r"(rr[8]),
            "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15])
          : "r"(trow + 128 + 16 * c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; i += 2)
        if (16 * c + i < d)
          *reinterpret_cast<float2*>(orow + 16 * c + i) = make_float2(__uint_as_float(rr[i]) * inv, __uint_as_float(rr[i + 1]) * inv);
    }
    } else {
    float acc[DV];
#pragma unroll
    for (int i = 0; i < DV; ++i) acc[i] = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      const uint32_t ph = (uint32_t)j & 1u;
      mbar_wait(bars + 8 * S_FULL, ph);
      tc_fence_after();
      float alpha;
      {
        float s0[32], s1[32];
        tmem_ld32(trow, s0);
        tmem_ld32(trow + 32, s1);
        float t = -CUDART_INF_F;
#pragma unroll
        for (int i = 0; i < 32; ++i) t = fmaxf(t, fmaxf(s0[i], s1[i]));
        const float mn = fmaxf(m, t);
        alpha = ft_ex2(m - mn);
        m = mn;
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          s0[i] = ft_ex2(s0[i] - mn);
          s1[i] = ft_ex2(s1[i] - mn);
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
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
      tc_fence_before();
      mbar_arrive(bars + 8 * P_FULL);
#pragma unroll
      for (int i = 0; i < DV; ++i) acc[i] *= alpha;                  // overlaps the PV MMAs
      mbar_wait(bars + 8 * O_FULL, ph);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < DV / 16; ++c) {
        uint32_t rr[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]), "=r"(rr[8]),
              "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15])
            : "r"(trow + 64 + 16 * c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[16 * c + i] += __uint_as_float(rr[i]);
      }
      tc_fence_before();
      mbar_arrive(bars + 8 * O_EMPTY);
    }
    const float inv = 1.f / l;
#pragma unroll
    for (int c = 0; c < DV; c += 2)
      if (c < d) *reinterpret_cast<float2*>(orow + c) = make_float2(acc[c] * inv, acc[c + 1] * inv);
    }
    lse[(size_t)h * S + row] = m + log2f(l);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(OTMEM ? FT_TMEM_COLS_PIPE : FT_TMEM_COLS));
  }
}

template <int DV, bool OTMEM>
static int sa_tc_launch_v(const __nv_bfloat16* Qp, const __nv_bfloat16* Kp, const __nv_bfloat16* VTp, float* o, int64_t ldo,
                        float* lse, int S, int heads, int d, cudaStream_t st) {
  CUtensorMap tq_hi, tq_lo, tk_hi, tk_lo, tv_hi, tv_lo;
  const size_t qk_plane = (size_t)heads * S * 64, vt_plane = (size_t)heads * DV * S;
  int rc;
  if ((rc = tc_make_map(&tq_hi, Qp, heads * S, 64, FT_BM))) return rc;
  if ((rc = tc_make_map(&tq_lo, Qp + qk_plane, heads * S, 64, FT_BM))) return rc;
  if ((rc = tc_make_map(&tk_hi, Kp, heads * S, 64, FT_BN))) return rc;
  if ((rc = tc_make_map(&tk_lo, Kp + qk_plane, heads * S, 64, FT_BN))) return rc;
  if ((rc = tc_make_map(&tv_hi, VTp, heads * DV, S, DV))) return rc;
  if ((rc = tc_make_map(&tv_lo, VTp + vt_plane, heads * DV, S, DV))) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(sa_tc_fwd_kernel<DV, OTMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM);
    if (e != cudaSuccess) { set_error("self_attn_tc_fwd: smem attr: %s", cudaGetErrorString(e)); return SKP_ERR_LAUNCH; }
    configured = true;
  }
  dim3 grid(S / FT_BM, heads);
  sa_tc_fwd_kernel<DV, OTMEM><<<grid, FT_THREADS, FT_SMEM, st>>>(tq_hi, tq_lo, tk_hi, tk_lo, tv_hi, tv_lo, o, ldo, lse, S, d, (d + 15) / 16);
  SKP_CHECK_LAUNCH("sa_tc_fwd_kernel");
  return SKP_OK;
}

// SKP_ATTN_FWD_PIPE=0: the first form of the kernel (one score buffer, the accumulator read back and rescaled in registers
// every key tile)
template <int DV>
static int sa_tc_launch(const __nv_bfloat16* Qp, const __nv_bfloat16* Kp, const __nv_bfloat16* VTp, float* o, int64_t ldo,
                        float* lse, int S, int heads, int d, cudaStream_t st) {
  static const bool o_tmem = !(getenv("SKP_ATTN_FWD_PIPE") && getenv("SKP_ATTN_FWD_PIPE")[0] == '0');
  return o_tmem ? sa_tc_launch_v<DV, true>(Qp, Kp, VTp, o, ldo, lse, S, heads, d, st)
                : sa_tc_launch_v<DV, false>(Qp, Kp, VTp, o, ldo, lse, S, heads, d, st);
}

static int sa_tc_dv(int d) { return d <= 16 ? 16 : d <= 32 ? 32 : d <= 48 ? 48 : 64; }

}  // namespace skp

using namespace skp;

// workspace bytes: (2*heads*S*64 * 2 + 2*heads*DV*S) bf16
extern "C" int64_t skp_self_attn_tc_workspace(int S, int heads, int d) {
  if (S <= 0 || heads <= 0 || d <= 0 || d > 64 || (d & 1) || S % FT_BM != 0) return 0;
  return ((int64_t)4 * heads * S * 64 + (int64_t)2 * heads * sa_tc_dv(d) * S) * 2;
}

extern "C" int skp_self_attn_tc_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                                    float* o, int64_t ldo, float* lse, void* workspace, int S, int heads, int d, float scale,
                                    void* stream) {
  SKP_REQUIRE(q && k && v && o && lse && workspace, "skp_self_attn_tc_fwd: null pointer");
  SKP_REQUIRE(skp_self_attn_tc_workspace(S, heads, d) > 0, "skp_self_attn_tc_fwd: needs S %% 128 == 0 and even d <= 64 (S=%d d=%d)", S, d);
  SKP_REQUIRE(ldo % 2 == 0 && (reinterpret_cast<uintptr_t>(o) & 7) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 127) == 0,
              "skp_self_attn_tc_fwd: o must be 8-byte aligned with even ld, the workspace 128-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int DV = sa_tc_dv(d);
  __nv_bfloat16* Qp = (__nv_bfloat16*)workspace;
  __nv_bfloat16* Kp = Qp + (size_t)2 * heads * S * 64;
  __nv_bfloat16* VTp = Kp + (size_t)2 * heads * S * 64;
  const long total = (long)2 * heads * S * 32 + (long)heads * DV * (S / 2);
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  sa_tc_split_kernel<<<(int)blocks, 256, 0, st>>>(q, ldq, k, ldk, v, ldv, Qp, Kp, VTp, S, heads, d, DV, scale * 1.4426950408889634f);
  SKP_CHECK_LAUNCH("sa_tc_split_kernel");
  switch (DV) {
    case 16: return sa_tc_launch<16>(Qp, Kp, VTp, o, ldo, lse, S, heads, d, st);
    case 32: return sa_tc_launch<32>(Qp, Kp, VTp, o, ldo, lse, S, heads, d, st);
    case 48: return sa_tc_launch<48>(Qp, Kp, VTp, o, ldo, lse, S, heads, d, st);
    default: return sa_tc_launch<64>(Qp, Kp, VTp, o, ldo, lse, S, heads, d, st);
  }
}
