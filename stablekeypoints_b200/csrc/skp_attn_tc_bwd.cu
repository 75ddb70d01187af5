// Self-attention BACKWARD on the 5th-generation tensor cores (tcgen05 + TMEM + TMA) for the long-sequence layers of the
// frozen UNet (attn1 of the 64x64 blocks: S = 4096 tokens, 8 heads x d = 40; the autograd pass through ptp_utils.py:493-506
// that carries d loss / d context back to the earlier cross-attention layers).  Numerics contract of skp_selfattn.cu and
// skp_attn_tc.cu: every contraction is a split-bf16 product (hi.hi + hi.lo + lo.hi, fp32 accumulate), the probabilities
// are recomputed from the forward's base-2 log-sum-exp, nothing [S, S]-sized touches HBM, no atomics (bit-reproducible).
//
// One launch, two kinds of CTA (blockIdx.z), both built from the same loop "resident 128-row block x streamed 64-row tiles":
//   z = 0  dK / dV of 128 keys:  S^T = K Q'^T and dP^T = V dO^T (M128 x N64 into a TMEM score buffer);
//          P^T = exp2(S^T - lse[q]), dS^T = P^T (dP^T - delta[q]) by the element-wise warps (thread = key = TMEM lane),
//          written back split-bf16 IN PLACE (tcgen05.st) as the A operands of dV += P^T dO and dK += dS^T Q', whose two
//          [128 x DV] accumulators stay resident in TMEM for the whole loop.
//   z = 1  dQ of 128 queries:    S = Q' K^T, dP = dO V^T, dS = P (dP - delta[row]); dQ += dS K in TMEM.
// Every MMA takes its A operand from tensor memory (the resident block is copied there once per CTA) and only B through
// shared memory; score buffers, streamed row tiles and streamed transposed tiles are double-buffered.
// Warp roles: 0 = TMA of the streamed row tiles, 1 = MMA issuer (software-pipelined: the score MMAs of tile j+2 are issued
// right behind the accumulation MMAs of tile j, so they run while the element-wise warps work on tile j+1), 2 = TMA of the
// streamed transposed tiles, 3..10 = element-wise (two warps per TMEM lane quarter, 32 score columns each).
// Operands come pre-split from sa_tc_bwd_split_kernel: Q' (scaled by scale*log2 e), K, V, dO as [heads*S][64] planes and
// Q'^T, K^T, dO^T as [heads*DV][S] planes (the B operands of the accumulation products), delta = rowsum(dO * O).
// Eligibility (host): S % 128 == 0, d even and <= 96.  Head dims above 64 use two 64-column chunks of the row planes and ONE
// score buffer (tensor memory holds 128 + 4 * DV columns: scores, two accumulators, the four planes of the resident block).
#include "skp_tc.cuh"
#include <math_constants.h>

namespace skp {

namespace {

constexpr int BT_BM = 128;                       // resident rows per CTA
constexpr int BT_BN = 64;                        // streamed rows per step
#ifndef SKP_BT_EW_WARPS
#define SKP_BT_EW_WARPS 8
#endif
constexpr int BT_EW_WARPS = SKP_BT_EW_WARPS;     // element-wise warps: 8 (32 score columns each) or 16 (16 columns each;
                                                 // measured slower on B200: 283 vs 267 us at S = 4096)
static_assert(BT_EW_WARPS == 8 || BT_EW_WARPS == 16, "two or four element-wise warps per TMEM lane quarter");
constexpr int BT_NPART = BT_EW_WARPS / 4;        // warps per lane quarter
constexpr int BT_CW = 64 / BT_NPART;             // score columns per element-wise warp
constexpr int BT_THREADS = 96 + 32 * BT_EW_WARPS;
constexpr int BT_B_BYTES = BT_BN * 128;
constexpr int BT_TMEM_COLS = 512;                // the whole tensor memory of the SM: map in front of bt_body
constexpr int bt_kch(int DV) { return DV > 64 ? 2 : 1; }     // 64-column chunks of a row plane
constexpr int bt_nsb(int DV) { return DV > 64 ? 1 : 2; }     // score buffers in tensor memory
// double-buffered streamed row tiles [chunks][64][64] and transposed tiles [DV][64], four bf16 planes each (the resident
// block and the P / dS staging live in tensor memory)
constexpr int bt_smem(int DV) { return 8 * bt_kch(DV) * BT_B_BYTES + 8 * DV * 128 + 256 + 1024; }

struct BtMaps {
  CUtensorMap b1h, b1l, b2h, b2l;                // streamed tiles (64 rows):  Q', dO (z = 0) / K, V (z = 1)
  CUtensorMap t1h, t1l, t2h, t2l;                // streamed transposed tiles [DV][64]: dO^T, Q'^T (z = 0) / -, K^T (z = 1)
};

__device__ __forceinline__ void bt_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float bt_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float v[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// two 32-column loads of this warp's TMEM lanes, one wait
__device__ __forceinline__ void tmem_ld32x2(uint32_t ta, uint32_t tb, float a[32], float b[32]) {
  uint32_t r[32], q[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(ta));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
        "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]),
        "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]),
        "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
      : "r"(tb));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    a[i] = __uint_as_float(r[i]);
    b[i] = __uint_as_float(q[i]);
  }
}
__device__ __forceinline__ void tmem_ld16x2(uint32_t ta, uint32_t tb, float a[16], float b[16]) {
  uint32_t r[16], q[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(ta));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]), "=r"(q[9]),
        "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
      : "r"(tb));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    a[i] = __uint_as_float(r[i]);
    b[i] = __uint_as_float(q[i]);
  }
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float v[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
// 2*W consecutive fp32 -> W bf16 pairs of the hi plane and W of the lo plane (element 2c in the low half of word c)
template <int W>
__device__ __forceinline__ void bt_split(const float* x, uint32_t* hi, uint32_t* lo) {
#pragma unroll
  for (int e = 0; e < W; ++e) {
    __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * e], x[2 * e + 1]);
    float2 f = __bfloat1622float2(hh);
    __nv_bfloat162 ll = __floats2bfloat162_rn(x[2 * e] - f.x, x[2 * e + 1] - f.y);
    hi[e] = *reinterpret_cast<uint32_t*>(&hh);
    lo[e] = *reinterpret_cast<uint32_t*>(&ll);
  }
}
template <int W>
__device__ __forceinline__ void tmem_st_w(uint32_t taddr, const uint32_t* r);
// tcgen05.st: this thread's TMEM lane, 8 / 16 consecutive columns (32: skp_tc.cuh; completion: tcgen05.wait::st by the caller)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t r[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t r[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
template <>
__device__ __forceinline__ void tmem_st_w<16>(uint32_t taddr, const uint32_t* r) { tmem_st16(taddr, r); }
template <>
__device__ __forceinline__ void tmem_st_w<8>(uint32_t taddr, const uint32_t* r) { tmem_st8(taddr, r); }
// q,k,v [S, heads*d] and d_o, o [S, heads*d] fp32 (leading dims) ->
//   row planes   RP[t][2][heads*S][KP]  t = Q' (scaled), K, V, dO          (hi, lo; zero padded to KP = 64 / 128 columns)
//   transposed   TP[t][2][heads*DV][S]  t = Q'^T, K^T, dO^T
//   delta[heads][S] = sum_c dO * O
// blockIdx.y selects the section (0..3 row planes -- the dO section also makes delta --, 4..6 transposed planes); all index arithmetic is 32-bit (the first
// version decomposed one flat 64-bit index with four 64-bit divisions per element and spent its time in them).
__global__ void sa_tc_bwd_split_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk,
                                       const float* __restrict__ v, int64_t ldv, const float* __restrict__ d_o, int64_t lddo,
                                       const float* __restrict__ o, int64_t ldo, __nv_bfloat16* __restrict__ RP,
                                       __nv_bfloat16* __restrict__ TP, float* __restrict__ delta, int S, int heads, int d, int DV,
                                       int kp_shift, float qscale) {
  const int KP = 1 << kp_shift;
  const size_t rp = (size_t)heads * S * KP, tp = (size_t)heads * DV * S;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int sec = blockIdx.y;
  if (sec < 4) {
    // one thread = eight adjacent columns of one row of one plane pair: 2 x 16-byte loads, 2 x 16-byte stores
    const int cp_shift = kp_shift - 3;             // 8-column groups per row
    if (idx >= (unsigned)heads * S << cp_shift) return;
    const int t = sec;
    const int c = (int)(idx & ((1u << cp_shift) - 1)) << 3;
    const unsigned hr = idx >> cp_shift;           // h * S + row
    const int h = (int)(hr / (unsigned)S), row = (int)(hr - (unsigned)h * S);
    const float* src = (t == 0 ? q + (size_t)row * ldq : t == 1 ? k + (size_t)row * ldk : t == 2 ? v + (size_t)row * ldv
                                                                                                : d_o + (size_t)row * lddo) + h * d + c;
    float x[8];
    if (c + 8 <= d && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b4 = __ldg(reinterpret_cast<const float4*>(src) + 1);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b4.x; x[5] = b4.y; x[6] = b4.z; x[7] = b4.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = c + i < d ? __ldg(src + i) : 0.f;
    }
    if (t == 3) {
      // delta = rowsum(dO * O): partial dot of this thread's eight columns, reduced over the 8 / 16 adjacent lanes of the row
      // (whole warps take this branch: heads * S * KP / 8 is a multiple of 32)
      const float* op = o + (size_t)row * ldo + h * d + c;
      float acc = 0.f;
      if (c + 8 <= d && (reinterpret_cast<uintptr_t>(op) & 15) == 0) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(op)), b4 = __ldg(reinterpret_cast<const float4*>(op) + 1);
        acc = x[0] * a.x + x[1] * a.y + x[2] * a.z + x[3] * a.w + x[4] * b4.x + x[5] * b4.y + x[6] * b4.z + x[7] * b4.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = fmaf(x[i], c + i < d ? __ldg(op + i) : 0.f, acc);
      }
      for (int off = 1; off < (1 << cp_shift); off <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (c == 0) delta[(size_t)h * S + row] = acc;
    }
    const float sc = t == 0 ? qscale : 1.f;
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float x0 = x[2 * e] * sc, x1 = x[2 * e + 1] * sc;
      __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
      float2 f = __bfloat1622float2(hh);
      __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - f.x, x1 - f.y);
      hw[e] = *reinterpret_cast<uint32_t*>(&hh);
      lw[e] = *reinterpret_cast<uint32_t*>(&ll);
    }
    __nv_bfloat16* hi = RP + (size_t)t * 2 * rp + ((size_t)hr << kp_shift) + c;
    *reinterpret_cast<uint4*>(hi) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(hi + rp) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  } else if (sec < 7) {
    // one thread = 16 consecutive rows of one channel: channel fastest across the warp, so each of the 16 row reads is a
    // coalesced segment and every store is one full 32-byte sector of the transposed plane
    const int seg = S >> 4;
    if (idx >= (unsigned)heads * DV * seg) return;
    const int t = sec - 4;
    const int c = (int)(idx % (unsigned)DV);
    const unsigned hs = idx / (unsigned)DV;        // h * seg + 16-row segment
    const int h = (int)(hs / (unsigned)seg), s0 = (int)(hs - (unsigned)h * seg) << 4;
    const float* src = (t == 0 ? q : t == 1 ? k : d_o) + h * d + c;
    const int64_t ld = t == 0 ? ldq : t == 1 ? ldk : lddo;
    const float sc = t == 0 ? qscale : 1.f;
    uint32_t hw[8], lw[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float x0 = c < d ? __ldg(src + (size_t)(s0 + 2 * e) * ld) * sc : 0.f;
      const float x1 = c < d ? __ldg(src + (size_t)(s0 + 2 * e + 1) * ld) * sc : 0.f;
      __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
      float2 f = __bfloat1622float2(hh);
      __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - f.x, x1 - f.y);
      hw[e] = *reinterpret_cast<uint32_t*>(&hh);
      lw[e] = *reinterpret_cast<uint32_t*>(&ll);
    }
    uint4* ph = reinterpret_cast<uint4*>(TP + (size_t)t * 2 * tp + ((size_t)h * DV + c) * S + s0);
    uint4* pl = reinterpret_cast<uint4*>(TP + (size_t)t * 2 * tp + tp + ((size_t)h * DV + c) * S + s0);
    ph[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    ph[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
    pl[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    pl[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
  }
}

// DKV = true: resident rows are keys, out1 = dV, out2 = dK; false: resident rows are queries, out2 = dQ.
// TMEM map (512 columns x 128 lanes, lane = resident row):
//   [0,128) and [128,256)  score buffers: S | dP (fp32) from the tensor core, overwritten IN PLACE by the element-wise warps
//                          with P hi | P lo | dS hi | dS lo (bf16 pairs, 32 columns each) -- the A operands of the accumulation MMAs
//   [256,320) [320,384)    accumulators (dV, dK  /  -, dQ), resident for the whole loop
//   [384,512)              the resident block itself, split-bf16: A1 hi | A1 lo | A2 hi | A2 lo (A operands of the score MMAs)
// Every MMA takes A from tensor memory and only B through shared memory: with three MMAs per product the shared-memory
// re-reads of A were what bound the first version of this kernel (ncu: 33 % tensor pipe, ~2 us per 64-row step).
template <int DV, bool DKV>
__device__ __forceinline__ void bt_body(const BtMaps& mp, const __nv_bfloat16* __restrict__ a1_planes,
                                        const __nv_bfloat16* __restrict__ a2_planes, size_t plane_stride,
                                        const float* __restrict__ lse, const float* __restrict__ delta,
                                        float* __restrict__ out1, int64_t ld1, float* __restrict__ out2, int64_t ld2, int S, int d,
                                        int ksteps, float scale2, uint8_t* smem_raw) {
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  constexpr int BT_T_BYTES = DV * 128;
  constexpr int KCH = bt_kch(DV), NSB = bt_nsb(DV), KP = 64 * KCH;
  constexpr int BP_BYTES = KCH * BT_B_BYTES;       // one plane of a streamed row tile: [KCH chunks][64 rows][64 columns]
  constexpr int ACOLS = DV / 2;                    // TMEM columns of one plane of the resident block (DV bf16 per row)
  const uint32_t sB = base;                        // [2 buffers][B1h, B1l, B2h, B2l]
  const uint32_t sT = sB + 8 * BP_BYTES;           // [2 buffers][T1h, T1l, T2h, T2l]
  const uint32_t bars = sT + 8 * BT_T_BYTES;
  // B_*, T_* exist per shared-memory buffer (index + (j & 1), phase (j >> 1) & 1), S_FULL / E_FULL per score buffer
  // (index + j % NSB, phase (j / NSB) & 1)
  enum { A_FULL = 0, B_FULL, B_FULL1, B_EMPTY, B_EMPTY1, T_FULL, T_FULL1, T_EMPTY, T_EMPTY1, S_FULL, S_FULL1, E_FULL, E_FULL1, DONE, NBARS };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + (bars - base) + 8 * NBARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, r0 = blockIdx.x * BT_BM;
  const int ntiles = S / BT_BN;
  constexpr uint32_t EW = 32u * BT_EW_WARPS;
  constexpr uint32_t TM_ACC1 = 128 * NSB, TM_ACC2 = TM_ACC1 + DV, TM_A = TM_ACC2 + DV;
  static_assert(TM_A + 4 * ACOLS <= 512, "tensor memory: scores + two accumulators + the resident block");

  if (warp == 0 && lane == 0) {
    for (int b = 0; b < NBARS; ++b) mbar_init(bars + 8 * b, (b == A_FULL || b == E_FULL || b == E_FULL1) ? EW : 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(BT_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (elect_one_sync()) {
      for (int j = 0; j < ntiles; ++j) {
        const int b = j & 1;
        const uint32_t full = bars + 8 * (B_FULL + b), dst = sB + (uint32_t)b * 4u * BP_BYTES;
        mbar_wait(bars + 8 * (B_EMPTY + b), (((uint32_t)j >> 1) & 1u) ^ 1u);
        mbar_expect_tx(full, 4 * BP_BYTES);
#pragma unroll
        for (int c = 0; c < KCH; ++c) {
          tma_load_2d(dst + c * BT_B_BYTES, &mp.b1h, full, c * 64, h * S + j * BT_BN);
          tma_load_2d(dst + BP_BYTES + c * BT_B_BYTES, &mp.b1l, full, c * 64, h * S + j * BT_BN);
          tma_load_2d(dst + 2 * BP_BYTES + c * BT_B_BYTES, &mp.b2h, full, c * 64, h * S + j * BT_BN);
          tma_load_2d(dst + 3 * BP_BYTES + c * BT_B_BYTES, &mp.b2l, full, c * 64, h * S + j * BT_BN);
        }
      }
    }
  } else if (warp == 2) {
    if (elect_one_sync()) {
      for (int j = 0; j < ntiles; ++j) {
        const int b = j & 1;
        const uint32_t full = bars + 8 * (T_FULL + b), dst = sT + (uint32_t)b * 4u * BT_T_BYTES;
        mbar_wait(bars + 8 * (T_EMPTY + b), (((uint32_t)j >> 1) & 1u) ^ 1u);
        mbar_expect_tx(full, (DKV ? 4 : 2) * BT_T_BYTES);
        if (DKV) {
          tma_load_2d(dst, &mp.t1h, full, j * BT_BN, h * DV);
          tma_load_2d(dst + BT_T_BYTES, &mp.t1l, full, j * BT_BN, h * DV);
        }
        tma_load_2d(dst + 2 * BT_T_BYTES, &mp.t2h, full, j * BT_BN, h * DV);
        tma_load_2d(dst + 3 * BT_T_BYTES, &mp.t2l, full, j * BT_BN, h * DV);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      // instruction descriptors: D = f32, A = B = bf16, both K-major, M = 128, N = 64 (scores) / DV (accumulators)
      constexpr uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BT_BN >> 3) << 17) | ((uint32_t)(BT_BM >> 4) << 24);
      constexpr uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(DV >> 3) << 17) | ((uint32_t)(BT_BM >> 4) << 24);
      const uint32_t tA1h = tmem + TM_A, tA1l = tA1h + ACOLS, tA2h = tA1h + 2 * ACOLS, tA2l = tA1h + 3 * ACOLS;
      auto issue_scores = [&](int j) {
        const int b = j & 1, sbuf = j % NSB;
        const uint32_t ts = tmem + (uint32_t)sbuf * 128u;
        const uint32_t sb = sB + (uint32_t)b * 4u * BP_BYTES;
        mbar_wait(bars + 8 * (B_FULL + b), ((uint32_t)j >> 1) & 1u);
        tc_fence_after();
        // (the score buffer's previous tenant, tile j - NSB, was consumed by accumulation MMAs issued before these: in-order pipe)
        for (int k = 0; k < ksteps; ++k) {
          const uint32_t koff = (uint32_t)(k >> 2) * BT_B_BYTES + (uint32_t)(k & 3) * 32u;   // chunk, then 16 bf16 = 32 B inside the swizzle row
          const uint32_t ka = 8u * (uint32_t)k;             // 16 bf16 of K = 8 TMEM columns
          const uint64_t dB1h = make_smem_desc(sb + koff), dB1l = make_smem_desc(sb + BP_BYTES + koff);
          umma_bf16_ts(ts, tA1l + ka, dB1h, idesc_s, k != 0);
          umma_bf16_ts(ts, tA1h + ka, dB1l, idesc_s, 1u);
          umma_bf16_ts(ts, tA1h + ka, dB1h, idesc_s, 1u);
        }
        for (int k = 0; k < ksteps; ++k) {
          const uint32_t koff = (uint32_t)(k >> 2) * BT_B_BYTES + (uint32_t)(k & 3) * 32u;
          const uint32_t ka = 8u * (uint32_t)k;
          const uint64_t dB2h = make_smem_desc(sb + 2 * BP_BYTES + koff), dB2l = make_smem_desc(sb + 3 * BP_BYTES + koff);
          umma_bf16_ts(ts + 64, tA2l + ka, dB2h, idesc_s, k != 0);
          umma_bf16_ts(ts + 64, tA2h + ka, dB2l, idesc_s, 1u);
          umma_bf16_ts(ts + 64, tA2h + ka, dB2h, idesc_s, 1u);
        }
        umma_commit(bars + 8 * (B_EMPTY + b));      // this buffer of streamed row tiles is free once these MMAs retire
        umma_commit(bars + 8 * (S_FULL + sbuf));    // ... and both score tiles are complete
      };
      mbar_wait(bars + 8 * A_FULL, 0);            // the resident block is in tensor memory
      tc_fence_after();
      for (int j = 0; j < NSB && j < ntiles; ++j) issue_scores(j);
      for (int j = 0; j < ntiles; ++j) {
        const int b = j & 1, sbuf = j % NSB;
        const uint32_t par = ((uint32_t)j >> 1) & 1u, ts = tmem + (uint32_t)sbuf * 128u;
        const uint32_t st = sT + (uint32_t)b * 4u * BT_T_BYTES;
        const uint64_t dT1h = make_smem_desc(st), dT1l = make_smem_desc(st + BT_T_BYTES);
        const uint64_t dT2h = make_smem_desc(st + 2 * BT_T_BYTES), dT2l = make_smem_desc(st + 3 * BT_T_BYTES);
        mbar_wait(bars + 8 * (E_FULL + sbuf), (uint32_t)(j / NSB) & 1u);    // P^T / dS of tile j sit in the score buffer
        mbar_wait(bars + 8 * (T_FULL + b), par);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < BT_BN / 16; ++k) {
          const uint64_t adv = (uint64_t)((k * 32) >> 4);
          const uint32_t ka = 8u * (uint32_t)k;
          const uint32_t acc = (j != 0 || k != 0) ? 1u : 0u;
          if (DKV) {
            umma_bf16_ts(tmem + TM_ACC1, ts + 32 + ka, dT1h + adv, idesc_o, acc);
            umma_bf16_ts(tmem + TM_ACC1, ts + ka, dT1l + adv, idesc_o, 1u);
            umma_bf16_ts(tmem + TM_ACC1, ts + ka, dT1h + adv, idesc_o, 1u);
          }
          umma_bf16_ts(tmem + TM_ACC2, ts + 96 + ka, dT2h + adv, idesc_o, acc);
          umma_bf16_ts(tmem + TM_ACC2, ts + 64 + ka, dT2l + adv, idesc_o, 1u);
          umma_bf16_ts(tmem + TM_ACC2, ts + 64 + ka, dT2h + adv, idesc_o, 1u);
        }
        umma_commit(bars + 8 * (T_EMPTY + b));
        if (j + NSB < ntiles) issue_scores(j + NSB);   // NSB = 2: runs while the element-wise warps work on tile j + 1
      }
      umma_commit(bars + 8 * DONE);
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    constexpr int CW = BT_CW, NPART = BT_NPART;
    const int part = (warp - 3) >> 2;             // which CW of the 64 score columns
    const int r = quarter * 32 + lane;            // resident row inside the block
    const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16);
    // ---- the resident block goes to tensor memory: this thread's row of planes part, part + NPART, .. (A1 hi, A1 lo, A2 hi, A2 lo)
    {
#pragma unroll
      for (int pl = 0; pl < 4; ++pl) {
        if (pl % NPART != part) continue;
        const __nv_bfloat16* src = (pl < 2 ? a1_planes : a2_planes) + (size_t)(pl & 1) * plane_stride + ((size_t)h * S + r0 + r) * KP;
        const uint4* p4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
        for (int i = 0; i < ACOLS / 8; ++i) {     // 8 TMEM columns = 16 bf16 = two 16-byte loads
          const uint4 v0 = __ldg(p4 + 2 * i), v1 = __ldg(p4 + 2 * i + 1);
          const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
          tmem_st8(trow + TM_A + pl * ACOLS + 8 * i, w);
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      bt_arrive(bars + 8 * A_FULL);
    }
    float lse_r = 0.f, delta_r = 0.f;
    if (!DKV) {
      lse_r = __ldg(lse + (size_t)h * S + r0 + r);
      delta_r = __ldg(delta + (size_t)h * S + r0 + r);
    }
    float lv[CW], dl[CW];
    auto load_stats = [&](int j) {                // per-column statistics of the streamed queries (warp-uniform addresses)
      const float4* lp = reinterpret_cast<const float4*>(lse + (size_t)h * S + j * BT_BN + CW * part);
      const float4* dp = reinterpret_cast<const float4*>(delta + (size_t)h * S + j * BT_BN + CW * part);
#pragma unroll
      for (int i = 0; i < CW / 4; ++i) {
        const float4 a = __ldg(lp + i), b = __ldg(dp + i);
        lv[4 * i] = a.x; lv[4 * i + 1] = a.y; lv[4 * i + 2] = a.z; lv[4 * i + 3] = a.w;
        dl[4 * i] = b.x; dl[4 * i + 1] = b.y; dl[4 * i + 2] = b.z; dl[4 * i + 3] = b.w;
      }
    };
    if (DKV) load_stats(0);
    for (int j = 0; j < ntiles; ++j) {
      const int sbuf = j % NSB;
      const uint32_t ts = trow + 128u * (uint32_t)sbuf;
      mbar_wait(bars + 8 * (S_FULL + sbuf), (uint32_t)(j / NSB) & 1u);
      tc_fence_after();
      float s[CW], g[CW];
      if constexpr (CW == 32) tmem_ld32x2(ts + CW * part, ts + 64 + CW * part, s, g);
      else tmem_ld16x2(ts + CW * part, ts + 64 + CW * part, s, g);
      // the bf16 results overwrite score columns the other warps of this lane quarter read: all must have loaded
      asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "n"(32 * NPART) : "memory");
#pragma unroll
      for (int i = 0; i < CW; ++i) {
        const float p = bt_ex2(s[i] - (DKV ? lv[i] : lse_r));
        s[i] = p;
        g[i] = p * (g[i] - (DKV ? dl[i] : delta_r));
      }
      if (DKV && j + 1 < ntiles) load_stats(j + 1);   // lands behind the conversions and the next tile's wait
      uint32_t hi[CW / 2], lo[CW / 2];
      if (DKV) {
        bt_split<CW / 2>(s, hi, lo);
        tmem_st_w<CW / 2>(ts + (CW / 2) * part, hi);            // P hi: columns [0,32), two queries per column
        tmem_st_w<CW / 2>(ts + 32 + (CW / 2) * part, lo);       // P lo: columns [32,64)
      }
      bt_split<CW / 2>(g, hi, lo);
      tmem_st_w<CW / 2>(ts + 64 + (CW / 2) * part, hi);         // dS hi: columns [64,96)
      tmem_st_w<CW / 2>(ts + 96 + (CW / 2) * part, lo);         // dS lo: columns [96,128)
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      bt_arrive(bars + 8 * (E_FULL + sbuf));
    }
    // ---- epilogue: the accumulators leave TMEM; this thread owns row r, columns [part*DV/NPART, (part+1)*DV/NPART)
    mbar_wait(bars + 8 * DONE, 0);
    tc_fence_after();
    const int row = r0 + r;
    constexpr int HC = DV / NPART;
#pragma unroll
    for (int which = DKV ? 0 : 1; which < 2; ++which) {
      float* orow = (which == 0 ? out1 + (size_t)row * ld1 : out2 + (size_t)row * ld2) + h * d;
      const float sc = which == 0 ? 1.f : scale2;
#pragma unroll
      for (int c4 = 0; c4 < HC / 4; ++c4) {
        float a[4];
        const int c0 = part * HC + 4 * c4;
        tmem_ld4(trow + (which == 0 ? TM_ACC1 : TM_ACC2) + c0, a);
#pragma unroll
        for (int i = 0; i < 4; i += 2)
          if (c0 + i < d) *reinterpret_cast<float2*>(orow + c0 + i) = make_float2(a[i] * sc, a[i + 1] * sc);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(BT_TMEM_COLS));
  }
}

template <int DV>
__global__ void __launch_bounds__(BT_THREADS, 1)
sa_tc_bwd_kernel(const __grid_constant__ BtMaps kv, const __grid_constant__ BtMaps qm, const __nv_bfloat16* __restrict__ RP,
                 size_t rp, const float* __restrict__ lse, const float* __restrict__ delta, float* __restrict__ dq, int64_t lddq,
                 float* __restrict__ dk, int64_t lddk, float* __restrict__ dv, int64_t lddv, int S, int d, int ksteps, float scale) {
  extern __shared__ uint8_t bt_smem_raw[];
  if (blockIdx.z == 0)        // resident K, V;  dK = ln 2 * dS^T Q'  (Q' carries scale * log2 e),  dV = P^T dO
    bt_body<DV, true>(kv, RP + 2 * rp, RP + 4 * rp, rp, lse, delta, dv, lddv, dk, lddk, S, d, ksteps, 0.6931471805599453f, bt_smem_raw);
  else                        // resident Q', dO;  dQ = scale * dS K
    bt_body<DV, false>(qm, RP, RP + 6 * rp, rp, lse, delta, nullptr, 0, dq, lddq, S, d, ksteps, scale, bt_smem_raw);
}

int bt_dv(int d) { return (d + 15) / 16 * 16; }   // 16 .. 96

template <int DV>
int bt_launch(const __nv_bfloat16* RP, const __nv_bfloat16* TP, const float* lse, const float* delta, float* dq, int64_t lddq,
              float* dk, int64_t lddk, float* dv, int64_t lddv, int S, int heads, int d, float scale, cudaStream_t st) {
  constexpr int KP = 64 * bt_kch(DV);
  const size_t rp = (size_t)heads * S * KP, tp = (size_t)heads * DV * S;
  const __nv_bfloat16 *Qh = RP, *Ql = RP + rp, *Kh = RP + 2 * rp, *Kl = RP + 3 * rp, *Vh = RP + 4 * rp, *Vl = RP + 5 * rp;
  const __nv_bfloat16 *Dh = RP + 6 * rp, *Dl = RP + 7 * rp;
  const __nv_bfloat16 *QTh = TP, *QTl = TP + tp, *KTh = TP + 2 * tp, *KTl = TP + 3 * tp, *DTh = TP + 4 * tp, *DTl = TP + 5 * tp;
  BtMaps kv, qm;
  int rc;
  const int rows = heads * S, trows = heads * DV;
#define BT_MAP(dst, ptr, R, KP, BOX) if ((rc = tc_make_map(&(dst), (ptr), (R), (KP), (BOX)))) return rc
  BT_MAP(kv.b1h, Qh, rows, KP, BT_BN); BT_MAP(kv.b1l, Ql, rows, KP, BT_BN); BT_MAP(kv.b2h, Dh, rows, KP, BT_BN); BT_MAP(kv.b2l, Dl, rows, KP, BT_BN);
  BT_MAP(kv.t1h, DTh, trows, S, DV); BT_MAP(kv.t1l, DTl, trows, S, DV); BT_MAP(kv.t2h, QTh, trows, S, DV); BT_MAP(kv.t2l, QTl, trows, S, DV);
  BT_MAP(qm.b1h, Kh, rows, KP, BT_BN); BT_MAP(qm.b1l, Kl, rows, KP, BT_BN); BT_MAP(qm.b2h, Vh, rows, KP, BT_BN); BT_MAP(qm.b2l, Vl, rows, KP, BT_BN);
  BT_MAP(qm.t2h, KTh, trows, S, DV); BT_MAP(qm.t2l, KTl, trows, S, DV);
#undef BT_MAP
  qm.t1h = qm.t2h;
  qm.t1l = qm.t2l;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(sa_tc_bwd_kernel<DV>, cudaFuncAttributeMaxDynamicSharedMemorySize, bt_smem(DV));
    if (e != cudaSuccess) { set_error("self_attn_tc_bwd: smem attr: %s", cudaGetErrorString(e)); return SKP_ERR_LAUNCH; }
    configured = true;
  }
  dim3 grid(S / BT_BM, heads, 2);
  sa_tc_bwd_kernel<DV><<<grid, BT_THREADS, bt_smem(DV), st>>>(kv, qm, RP, rp, lse, delta, dq, lddq, dk, lddk, dv, lddv, S, d, (d + 15) / 16, scale);
  SKP_CHECK_LAUNCH("sa_tc_bwd_kernel");
  return SKP_OK;
}

}  // namespace

}  // namespace skp

using namespace skp;

// workspace bytes: 4 row-plane pairs [heads*S][64] + 3 transposed pairs [heads*DV][S] (bf16) + delta [heads][S] (fp32)
extern "C" int64_t skp_self_attn_tc_bwd_workspace(int S, int heads, int d) {
  if (S <= 0 || heads <= 0 || d <= 0 || d > 96 || (d & 1) || S % BT_BM != 0) return 0;
  const int DV = bt_dv(d);
  return ((int64_t)8 * heads * S * 64 * bt_kch(DV) + (int64_t)6 * heads * DV * S) * 2 + (int64_t)heads * S * 4;
}

extern "C" int skp_self_attn_tc_bwd(const float* d_o, int64_t lddo, const float* o, int64_t ldo, const float* lse, const float* q,
                                    int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv, void* workspace,
                                    float* dq, int64_t lddq, float* dk, int64_t lddk, float* dv, int64_t lddv, int S, int heads,
                                    int d, float scale, void* stream) {
  SKP_REQUIRE(d_o && o && lse && q && k && v && workspace && dq && dk && dv, "skp_self_attn_tc_bwd: null pointer");
  SKP_REQUIRE(skp_self_attn_tc_bwd_workspace(S, heads, d) > 0, "skp_self_attn_tc_bwd: needs S %% 128 == 0 and even d <= 96 (S=%d d=%d)", S, d);
  SKP_REQUIRE((lddq | lddk | lddv) % 2 == 0 && ((reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv)) & 7) == 0 &&
                  (reinterpret_cast<uintptr_t>(workspace) & 127) == 0 && (reinterpret_cast<uintptr_t>(lse) & 15) == 0,
              "skp_self_attn_tc_bwd: gradients must be 8-byte aligned with even ld, lse 16-byte and the workspace 128-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int DV = bt_dv(d);
  __nv_bfloat16* RP = (__nv_bfloat16*)workspace;
  const int KP = 64 * bt_kch(DV);
  __nv_bfloat16* TP = RP + (size_t)8 * heads * S * KP;
  float* delta = reinterpret_cast<float*>(TP + (size_t)6 * heads * DV * S);
  const long n_row = (long)heads * S * (KP / 8), n_tr = (long)heads * DV * (S / 16);
  const long widest = n_row > n_tr ? n_row : n_tr;   // threads of the largest section
  SKP_REQUIRE((long)heads * S * KP < (1L << 31), "skp_self_attn_tc_bwd: problem too large for 32-bit indexing");
  dim3 sgrid((unsigned)((widest + 255) / 256), 7);
  sa_tc_bwd_split_kernel<<<sgrid, 256, 0, st>>>(q, ldq, k, ldk, v, ldv, d_o, lddo, o, ldo, RP, TP, delta, S, heads, d, DV,
                                                KP == 128 ? 7 : 6, scale * 1.4426950408889634f);
  SKP_CHECK_LAUNCH("sa_tc_bwd_split_kernel");
  switch (DV) {
    case 16: return bt_launch<16>(RP, TP, lse, delta, dq, lddq, dk, lddk, dv, lddv, S, heads, d, scale, st);
    case 32: return bt_launch<32>(RP, TP, lse, delta, dq, lddq, dk, lddk, dv, lddv, S, heads, d, scale, st);
    case 48: return bt_launch<48>(RP, TP, lse, delta, dq, lddq, dk, lddk, dv, lddv, S, heads, d, scale, st);
    case 64: return bt_launch<64>(RP, TP, lse, delta, dq, lddq, dk, lddk, dv, lddv, S, heads, d, scale, st);
    case 80: return bt_launch<80>(RP, TP, lse, delta, dq, lddq, dk, lddk, dv, lddv, S, heads, d, scale, st);
    default: return bt_launch<96>(RP, TP, lse, delta, dq, lddq, dk, lddk, dv, lddv, S, heads, d, scale, st);
  }
}
