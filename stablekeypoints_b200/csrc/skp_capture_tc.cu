// Attention-store kernel on the 5th-generation tensor cores (tcgen05 + TMEM) with the softmax and the store fused behind
// the MMA: ptp_utils.py:508-538 (STORE: probs[h, R*R, N]) and, fused with optimize.py:50-75, MEAN: maps[N, R, R].
//
// By linearity (skp_capture.cu) the captured logits of output row Y are  x[X, n] = sum_xs Wx[X, xs] * V[xs, n]  with
// V[xs, n] = sum_j wy_j(Y) * L[h, row_j(Y), xs, n]  (vertical bicubic taps) and Wx the [R x s] horizontal bicubic matrix
// (border taps accumulated onto the edge columns).  The horizontal pass is a GEMM whose M tile is exactly one row of
// 128 output pixels:
//
//   D[128 pixels, N tokens] = A[128, s] . B[N, s]^T ,  A = Wx * log2(e)  (split bf16, built once per CTA),
//                                                      B = (V - colmax)^T (split bf16, built per output row)
//
// so ONE tcgen05.mma group (M128 x N=pad16(N) x K16 x ceil(s/16) steps x 3 split terms) produces all logits of an output
// row in tensor memory, and the softmax over tokens becomes thread-local: thread = pixel = TMEM lane reads its N logits
// with tcgen05.ld, takes the exact max / exp2 / sum in registers (no shuffles, no shared-memory reads) and
//   STORE: stages the normalised row [128][N] in its global layout and hands it to the copy engine (cp.async.bulk), or
//   MEAN : accumulates  w * p  over the (layer, head) items of the row in registers and writes maps[n, Y, X] once.
// Before the bf16 split every column of V is shifted by the vertically interpolated row maximum of the source logits
// (sum_j wy_j * max_n L[h, row_j, xs, n], from a tiny pre-pass): bicubic weights sum to one and softmax is shift invariant,
// so this costs nothing in exactness, and the split error becomes relative to the GAP to the maximum, not to the raw logit.
//
// Warp roles (8 + PW + 1 warps, one CTA per SM; PW = 7 producers with 128 registers per thread, or 3 with 168):
//   warps 0..7     epilogue  : thread = (pixel, half of the token axis); TMEM lane quarter = warp & 3.  The two halves of a
//                              pixel meet through one (max, sum) pair in shared memory.
//   warps 8..8+PW  producers : vertical pass out of shared memory (the four source rows of a group of output rows arrive
//                              as cp.async.bulk copies, double buffered); unit = (pair of low-res columns, 32 tokens), lanes
//                              run over tokens: conflict-free shared loads, split, one swizzled 32-bit store per plane
//   DMA warp (the last of the PW): one lane polls two cursors -- the next group of source rows to fetch (cp.async.bulk into a
//                              free row slot) and, in STORE mode, the next staged row to hand to the copy engine -- so that
//                              neither the producers nor the epilogue ever wait for an issue slot of the engine
//   last warp      MMA issuer: one lane; tcgen05.commit releases the B stage and publishes the accumulator
// Pipelines: B stages in shared memory (full / empty mbarriers) and TMEM accumulators (acc_full / acc_empty), so the
// vertical pass of row i+2, the MMA of row i+1 and the softmax of row i overlap.
//
// Work split.  STORE: CTA = (chunk of the (head, Y) rows, x-tile); consecutive Y share their source rows in L1.
// MEAN: CTA = (output row Y, x-tile), looping over all (layer, head) items.
#include "skp_tc.cuh"
#include <math_constants.h>
#include <stdlib.h>

namespace skp {

constexpr int CT_MAX_STAGES = 3;
constexpr int CT_A_PLANE = 128 * 128;   // one bf16 plane of the A tile: 128 pixels x 64 columns (128 B rows)
constexpr int CT_MAX_SLOTS = 2;         // distinct low-res sides among the layers of one launch (SD: 16 and 32); their K
                                        // ranges sit side by side in the ONE 64-column A tile
constexpr float CT_PAD_LOGIT = -8192.f; // token padding rows of B: exp2(-8192 * log2 e) == 0, so no token predicate in the softmax

struct CapTcParams {
  const float* logits[SKP_MAX_LAYERS];
  const float* rowmax[SKP_MAX_LAYERS];   // [h, s, s] max over tokens of the layer's logits (workspace, pre-pass)
  int s[SKP_MAX_LAYERS];
  int kofs[SKP_MAX_LAYERS];     // first A column of the layer's side (multiple of 16)
  int slot_s[CT_MAX_SLOTS], slot_kofs[CT_MAX_SLOTS];
  int n_slots, n_layers;
  int heads, N, R, XT;          // XT = x tiles of 128 pixels per row
  int rows_per_cta;             // STORE: rows of the flattened (head, Y) sequence per CTA (x-tile = blockIdx.y)
  int nst;                      // B stages in shared memory (2 or 3)
  int nstg;                     // STORE: staging buffers (1 or 2)
  int slot_floats;              // floats of one source-row group slot: 4 rows x max_s x N, then 4 x 32 row maxima
  long long* dbg;               // optional per-role time stamps of CTA 0 (scripts/capture_tc_trace.py); nullptr normally
  int store_path;               // 0: copy engine (cp.async.bulk); 1: plain coalesced 128-bit stores; 2: no store (measurement only)
  float* out;                   // STORE: probs [h, R*R, N];  MEAN: maps [N, R, R]
  float w;                      // MEAN: 1 / (layers * heads)
};

__device__ __forceinline__ float ct_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float2 ct_add2(float2 a, float2 b) {
  float2 c;
  asm("add.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t&>(c)) : "l"(reinterpret_cast<uint64_t const&>(a)), "l"(reinterpret_cast<uint64_t const&>(b)));
  return c;
}
__device__ __forceinline__ float2 ct_mul2(float2 a, float2 b) {
  float2 c;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t&>(c)) : "l"(reinterpret_cast<uint64_t const&>(a)), "l"(reinterpret_cast<uint64_t const&>(b)));
  return c;
}
__device__ __forceinline__ float2 ct_fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<uint64_t const&>(a)), "l"(reinterpret_cast<uint64_t const&>(b)), "l"(reinterpret_cast<uint64_t const&>(c)));
  return d;
}
__device__ __forceinline__ void ct_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// wait with a back-off: the waiting warps share their scheduler with the warps they are waiting for
__device__ __forceinline__ void ct_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    asm volatile("nanosleep.u32 32;" ::: "memory");
  }
}
__device__ __forceinline__ bool ct_mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void ct_named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void ct_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

#define CT_STAMP(role, it, k)                                                                    \
  do {                                                                                           \
    if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && (it) < 64)                     \
      p.dbg[((role) * 64 + (it)) * 8 + (k)] = clock64();                                         \
  } while (0)

// byte offset of bf16 element (row, col) inside a K-major tile of 128-byte rows with the 128B swizzle (8-row groups 1024 B apart)
__device__ __forceinline__ uint32_t sw128_off(int row, int col) {
  return (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + ((((uint32_t)col >> 3) ^ ((uint32_t)row & 7u)) << 4) +
         ((uint32_t)col & 7u) * 2u;
}

// max over tokens of every low-res logit row of every layer of the launch (one warp per row, one launch); non-finite
// maxima are reported as 0 (no shift)
struct CapRowmaxParams {
  const float* logits[SKP_MAX_LAYERS];
  float* out[SKP_MAX_LAYERS];
  int row_end[SKP_MAX_LAYERS];   // prefix sums of heads * s * s
  int n_layers, N;
};
__global__ void capture_rowmax_kernel(const CapRowmaxParams rp) {
  const int lane = threadIdx.x & 31;
  int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (row >= rp.row_end[rp.n_layers - 1]) return;
  int l = 0;
  while (row >= rp.row_end[l]) ++l;
  if (l > 0) row -= rp.row_end[l - 1];
  const float* src = rp.logits[l] + (size_t)row * rp.N;
  float m = -CUDART_INF_F;
  for (int n = lane; n < rp.N; n += 32) m = fmaxf(m, __ldg(src + n));
  m = warp_max(m);
  if (lane == 0) rp.out[l][row] = fabsf(m) < 3.0e38f ? m : 0.f;
}

template <int NPT>
struct CtCfg {
  static constexpr int B_PLANE = NPT * 128;                 // NPT token rows x 64 columns
  static constexpr int B_STAGE = 2 * B_PLANE;
  static constexpr int ACC_STRIDE = (NPT + 31) / 32 * 32;   // TMEM columns per accumulator
  static constexpr int NACC = (512 / ACC_STRIDE) < 4 ? (512 / ACC_STRIDE) : 4;
  static constexpr int NBARS = 2 * CT_MAX_STAGES + 2 * NACC + 8;   // B full/empty, accumulator full/empty, row-slot full/empty, staging full/free
  static constexpr int HALF = NPT / 2;                      // tokens per epilogue half (multiple of 16)
};

// MODE 0 = STORE, 1 = MEAN;  PW = producer warps
template <int NPT, int MODE, int PW>
__global__ void __launch_bounds__((9 + PW) * 32, 1) capture_tc_kernel(const CapTcParams p, const int stage_floats) {
  using Cfg = CtCfg<NPT>;
  constexpr int HALF = Cfg::HALF;
  constexpr int NTHREADS = (9 + PW) * 32;
  constexpr int MMA_WARP = 8 + PW;
  extern __shared__ uint8_t ct_smem_raw[];
  const uint32_t raw = smem_u32(ct_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = ct_smem_raw + (base - raw);
  // layout: A tile | B stages | source-row group slots (2) | barriers, tmem slot | (max, sum) exchange | staging (STORE)
  const int nst = p.nst;
  const uint32_t sA = base;
  const uint32_t sB = sA + 2 * CT_A_PLANE;
  const uint32_t sRows = sB + nst * Cfg::B_STAGE;
  const uint32_t row_slot_bytes = (uint32_t)p.slot_floats * 4u;
  const uint32_t bars = sRows + 2 * row_slot_bytes;
  const uint32_t FULL = bars, EMPTY = bars + 8 * CT_MAX_STAGES, AFULL = bars + 16 * CT_MAX_STAGES, AEMPTY = AFULL + 8 * Cfg::NACC;
  const uint32_t RFULL = AEMPTY + 8 * Cfg::NACC, REMPTY = RFULL + 16, SFULL = REMPTY + 16, SFREE = SFULL + 16;
  constexpr int NPROD = PW - 1;                            // the last of the PW warps is the DMA warp
  uint8_t* after_bars = gen + (bars - base) + 8 * Cfg::NBARS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(after_bars);
  float2* xch = reinterpret_cast<float2*>(after_bars + 16);                       // [2 parities][2 halves][128 pixels]
  float* staging = reinterpret_cast<float*>(after_bars + 16 + 2 * 2 * 128 * 8);   // 16-byte aligned (STORE)
  const float* rows_smem = reinterpret_cast<const float*>(gen + (sRows - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.N, R = p.R;
  const int xt = blockIdx.y, x0 = xt * 128, cnt = min(128, R - x0);

  // ---- work of this CTA.  STORE: rows [row0, row0 + n_items) of the flattened (head, Y) order of the one layer.
  //                         MEAN : items (layer, head) for the fixed output row Y = blockIdx.x.
  int n_items, row0 = 0;
  if (MODE == 0) {
    const int total = p.heads * R;
    row0 = blockIdx.x * p.rows_per_cta;
    n_items = max(0, min(total, row0 + p.rows_per_cta) - row0);
  } else {
    n_items = p.n_layers * p.heads;
  }
  // item -> (layer, head, Y) and the first of the 4 (clamped) source rows; items with equal (layer, head, iy) share them
  auto decode = [&](int it, int& l, int& h, int& Y, int& iy) {
    if (MODE == 0) {
      const int r = row0 + it;
      l = 0;
      h = r / R;
      Y = r - h * R;
    } else {
      l = it / p.heads;
      h = it - l * p.heads;
      Y = blockIdx.x;
    }
    iy = (int)floorf((float)p.s[l] / (float)R * (Y + 0.5f) - 0.5f);
  };

  if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.dbg[(2 * 64 + 62) * 8 + 0] = clock64();
    p.dbg[(2 * 64 + 62) * 8 + 1] = (long long)gt;
  }
  // ---- one-time set-up
  if (warp == MMA_WARP && lane == 0) {
    for (int i = 0; i < CT_MAX_STAGES; ++i) {
      mbar_init(FULL + 8 * i, NPROD);          // one arrive per producer warp
      mbar_init(EMPTY + 8 * i, 1);             // tcgen05.commit
    }
    for (int i = 0; i < Cfg::NACC; ++i) {
      mbar_init(AFULL + 8 * i, 1);             // tcgen05.commit
      mbar_init(AEMPTY + 8 * i, 8);            // one arrive per epilogue warp
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(RFULL + 8 * i, 1);             // expect_tx of the loader + the copy engine's complete_tx
      mbar_init(REMPTY + 8 * i, NPROD);
      mbar_init(SFULL + 8 * i, 8);             // one arrive per epilogue warp: the row is staged
      mbar_init(SFREE + 8 * i, 1);             // the store warp: the copy engine has read the buffer
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  // B stages: token padding rows [N, NPT) carry a large negative logit in every K column (their probabilities are exactly
  // 0); the K padding columns [s, pad16(s)) of the real rows must read as zero -- one pass fills both
  {
    const uint32_t padw = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(CT_PAD_LOGIT)) * 0x00010001u;
    uint4* z = reinterpret_cast<uint4*>(gen + (sB - base));
    const int per_stage = Cfg::B_STAGE >> 4;
    for (int i = threadIdx.x; i < nst * per_stage; i += NTHREADS) {
      const int r = i % per_stage;                         // 16-byte chunk inside the stage: hi plane first, rows of 8 chunks
      const bool hi_pad = r < (Cfg::B_PLANE >> 4) && (r >> 3) >= N;
      z[i] = hi_pad ? make_uint4(padw, padw, padw, padw) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  // A tile: thread = pixel of the x tile writes its whole 128-byte row (both planes): zeros, then its 4 horizontal taps
  // (clamped at the border, duplicates accumulated) * log2(e); the sides of the launch occupy disjoint column ranges
  if (threadIdx.x < 128) {
    const int pix = threadIdx.x, X = x0 + pix;
    uint8_t* Ah = gen;
    uint8_t* Al = Ah + CT_A_PLANE;
    const uint32_t rowoff = (uint32_t)(pix >> 3) * 1024u + (uint32_t)(pix & 7) * 128u;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      *reinterpret_cast<uint4*>(Ah + rowoff + c * 16) = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(Al + rowoff + c * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    if (X < R) {
      for (int sl = 0; sl < p.n_slots; ++sl) {
        const int s = p.slot_s[sl], k0 = p.slot_kofs[sl];
        const float scale = (float)s / (float)R;
        const float rx = scale * (X + 0.5f) - 0.5f, fx = floorf(rx);
        float wx[4];
        cubic_coeffs(rx - fx, wx);
        int col[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = (int)fx - 1 + i;
          col[i] = c < 0 ? 0 : (c > s - 1 ? s - 1 : c);
          wx[i] *= 1.4426950408889634f;
        }
#pragma unroll
        for (int i = 1; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < i; ++j)
            if (col[i] >= 0 && col[j] == col[i]) {
              wx[j] += wx[i];
              col[i] = -1;
            }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (col[i] >= 0) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(wx[i]);
            const __nv_bfloat16 lo = __float2bfloat16_rn(wx[i] - __bfloat162float(hi));
            const uint32_t off = sw128_off(pix, k0 + col[i]);
            *reinterpret_cast<__nv_bfloat16*>(Ah + off) = hi;
            *reinterpret_cast<__nv_bfloat16*>(Al + off) = lo;
          }
      }
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // first item of the group after the one `it` belongs to (n_items: none)
  auto next_group_start = [&](int it) {
    if (MODE == 1) return it + 1;
    int l0, h0, Y0, iy0;
    decode(it, l0, h0, Y0, iy0);
    int j = it + 1;
    for (; j < n_items; ++j) {
      int l, h, Y, iy;
      decode(j, l, h, Y, iy);
      if (h != h0 || iy != iy0) break;
    }
    return j;
  };

  if (warp == 8 + PW - 1) {
    // =========================================================== DMA warp: source rows in, staged rows out
    if (elect_one_sync()) {
      const int max_rowfl = (p.slot_floats - 128) >> 2;      // floats of one source row inside a slot
      int lg = 0, lit = 0;                                   // next group to fetch and its first item
      int sit = 0;                                           // next item to store (STORE)
      const int n_store = MODE == 0 ? n_items : 0;
      while (lit < n_items || sit < n_store) {
        bool progressed = false;
        if (lit < n_items && ct_mbar_test(REMPTY + 8 * (lg & 1), (((uint32_t)lg >> 1) & 1u) ^ 1u)) {
          // the 4 source rows of the group and their row maxima -> row slot lg & 1
          int l, h, Y, iy;
          decode(lit, l, h, Y, iy);
          const int s = p.s[l];
          const uint32_t slot = (uint32_t)lg & 1u, bytes = (uint32_t)(s * N) * 4u, mbytes = (uint32_t)s * 4u;
          mbar_expect_tx(RFULL + 8 * slot, 4u * (bytes + mbytes));
          const uint32_t dst = sRows + slot * row_slot_bytes;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            int r = iy - 1 + j;
            r = r < 0 ? 0 : (r > s - 1 ? s - 1 : r);
            ct_bulk_load(dst + (uint32_t)(j * max_rowfl) * 4u, p.logits[l] + ((size_t)h * s + r) * s * N, bytes, RFULL + 8 * slot);
            ct_bulk_load(dst + (uint32_t)(4 * max_rowfl + 32 * j) * 4u, p.rowmax[l] + ((size_t)h * s + r) * s, mbytes, RFULL + 8 * slot);
          }
          lit = next_group_start(lit);
          ++lg;
          progressed = true;
        }
        if (sit < n_store && ct_mbar_test(SFULL + 8 * (sit & 1), ((uint32_t)sit >> 1) & 1u)) {
          int l, h, Y, iy;
          decode(sit, l, h, Y, iy);
          const uint32_t b = (uint32_t)sit & 1u;
          const float* stg = staging + (size_t)(p.nstg == 2 ? b : 0) * stage_floats;
          float* dst = p.out + (((size_t)h * R + Y) * R + x0) * N;
          const size_t bytes = (size_t)cnt * N * sizeof(float);
          const bool al16 = ((bytes & 15) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
          if (al16 && p.store_path == 0) {                 // otherwise the epilogue warps stored the row themselves
            const uint32_t src = smem_u32(stg);
            const char* d8 = reinterpret_cast<const char*>(dst);
            for (size_t off = 0; off < bytes; off += 32768) {
              const uint32_t nb = (uint32_t)(bytes - off < 32768 ? bytes - off : 32768);
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d8 + off), "r"(src + (uint32_t)off), "r"(nb)
                           : "memory");
            }
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          if (p.nstg == 2) {
            if (sit > 0) {
              asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              ct_mbar_arrive(SFREE + 8 * (b ^ 1u));       // the row of the previous item has left its buffer
            }
          } else {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            ct_mbar_arrive(SFREE + 8 * b);
          }
          ++sit;
          progressed = true;
        }
        if (!progressed) asm volatile("nanosleep.u32 20;" ::: "memory");
      }
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must outlive the engine's reads
    }
    __syncwarp();
  } else if (warp >= 8 && warp < 8 + NPROD) {
    // =========================================================== producers: vertical pass (shared memory -> B stage)
    const int pw = warp - 8;
    const int max_rowfl = (p.slot_floats - 128) >> 2;      // floats of one source row inside a slot
    constexpr int NC = NPT / 32;
    // unit = (pair of low-res columns, chunk of 32 tokens); the units of a side are dealt round-robin to the producer warps
    // and a warp's share is at most MAXU (s <= 32).  Their source / destination offsets depend only on the side s: they are
    // tabulated in registers when s changes (once per launch for STORE, once per layer-side for MEAN), so that the loop over
    // the items is loads, FMAs, converts and stores with immediate offsets -- no index arithmetic.
    constexpr int MAXU = (16 * NC + NPROD - 1) / NPROD;
    int u_src[MAXU];                       // float offset of (xs, n) inside a source row ((xs + 1, n) is N floats further)
    uint32_t u_dst[MAXU];                  // swizzled byte offset of (n, xs) inside a B plane
    int u_xs[MAXU];                        // xs, or -1: no unit / token beyond N (nothing is stored)
    int tab_s = -1, nunits_w = 0;
    auto tabulate = [&](int s) {
      tab_s = s;
      const int npairs = (s + 1) >> 1, nunits = npairs * NC;
      nunits_w = 0;
#pragma unroll
      for (int q = 0; q < MAXU; ++q) {
        const int u = pw + q * NPROD;
        const int uu = min(u, nunits - 1);
        const int pi = uu / NC, c = uu - pi * NC;
        const int xs = 2 * pi;
        const int n = c * 32 + lane, nn = min(n, N - 1);
        u_src[q] = xs * N + nn;
        u_dst[q] = sw128_off(nn, xs);
        u_xs[q] = (u < nunits && n < N) ? xs : -1;
        if (u < nunits) nunits_w = q + 1;
      }
    };
    // item iterator (no divisions in the loop)
    int l = 0, h = 0, Y = 0;
    if (MODE == 0) {
      h = row0 / R;
      Y = row0 - h * R;
    } else {
      Y = blockIdx.x;
    }
    int g = 0, prev_h = h, prev_iy = 0x7fffffff, prev_l = l;
    for (int it = 0; it < n_items; ++it) {
      const int s = p.s[l];
      const float scale = (float)s / (float)R;
      const float ry = scale * (Y + 0.5f) - 0.5f, fy = floorf(ry);
      const int iy = (int)fy;
      bool fresh = it == 0;
      if (it > 0 && (MODE == 1 || h != prev_h || iy != prev_iy || l != prev_l)) {   // next group: release the old row slot
        __syncwarp();
        if (lane == 0) ct_mbar_arrive(REMPTY + 8 * (g & 1));
        ++g;
        fresh = true;
      }
      prev_h = h; prev_iy = iy; prev_l = l;
      if (s != tab_s) tabulate(s);
      const int st = it % nst;
      const uint32_t ph = (uint32_t)(it / nst) & 1u;
      float wy[4];
      cubic_coeffs(ry - fy, wy);
      if (pw == 0 && lane == 0) CT_STAMP(0, it, 0);
      if (fresh) ct_mbar_wait(RFULL + 8 * (g & 1), ((uint32_t)g >> 1) & 1u);
      const float* src = rows_smem + (size_t)(g & 1) * p.slot_floats;
      // shift of column xs = lane: the vertically interpolated row maximum (lanes >= s hold garbage that is never used)
      const float* mxs = src + 4 * max_rowfl + lane;
      const float shift = fmaf(wy[3], mxs[96], fmaf(wy[2], mxs[64], fmaf(wy[1], mxs[32], wy[0] * mxs[0])));
      ct_mbar_wait(EMPTY + 8 * st, ph ^ 1u);
      if (pw == 0 && lane == 0) CT_STAMP(0, it, 1);
      uint8_t* Bh = gen + (sB - base) + st * Cfg::B_STAGE;
      uint8_t* Bl = Bh + Cfg::B_PLANE;
      // trips of TRIP units: all their shared loads are issued before the first use
      constexpr int TRIP = 4;
      const bool odd_s = (s & 1) != 0;
#pragma unroll
      for (int q0 = 0; q0 < MAXU; q0 += TRIP) {
        if (q0 < nunits_w) {
          float raw[TRIP][2][4];
#pragma unroll
          for (int q = 0; q < TRIP; ++q)
            if (q0 + q < MAXU) {
              const bool two = !(odd_s && u_xs[q0 + q] == s - 1);
              const float* a0 = src + u_src[q0 + q];
              const float* b0 = a0 + (two ? N : 0);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                raw[q][0][j] = a0[j * max_rowfl];
                raw[q][1][j] = b0[j * max_rowfl];
              }
            }
          if (pw == 0 && lane == 0 && q0 == 0) CT_STAMP(0, it, 4);
#pragma unroll
          for (int q = 0; q < TRIP; ++q)
            if (q0 + q < MAXU) {
              const int xq = u_xs[q0 + q];
              const int xs = xq < 0 ? 0 : xq;
              const bool two = !(odd_s && xq == s - 1);
              const float sh0 = __shfl_sync(0xffffffffu, shift, xs), sh1 = __shfl_sync(0xffffffffu, shift, min(xs + 1, 31));
              float a = fmaf(wy[0], raw[q][0][0], -sh0), b = fmaf(wy[0], raw[q][1][0], -sh1);
#pragma unroll
              for (int j = 1; j < 4; ++j) {
                a = fmaf(wy[j], raw[q][0][j], a);
                b = fmaf(wy[j], raw[q][1][j], b);
              }
              if (!two) b = 0.f;
              if (xq >= 0) {
                const __nv_bfloat162 hh = __floats2bfloat162_rn(a, b);
                const float2 f = __bfloat1622float2(hh);
                const __nv_bfloat162 ll = __floats2bfloat162_rn(a - f.x, b - f.y);
                *reinterpret_cast<__nv_bfloat162*>(Bh + u_dst[q0 + q]) = hh;
                *reinterpret_cast<__nv_bfloat162*>(Bl + u_dst[q0 + q]) = ll;
              }
            }
        }
      }
      if (pw == 0 && lane == 0) CT_STAMP(0, it, 2);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) ct_mbar_arrive(FULL + 8 * st);
      if (pw == 0 && lane == 0) CT_STAMP(0, it, 3);
      // next item
      if (MODE == 0) {
        if (++Y == R) { Y = 0; ++h; }
      } else {
        if (++h == p.heads) { h = 0; ++l; }
      }
    }
  } else if (warp == MMA_WARP) {
    // =========================================================== MMA issuer
    if (elect_one_sync()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NPT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      for (int it = 0; it < n_items; ++it) {
        int l, h, Y, iy;
        decode(it, l, h, Y, iy);
        const int st = it % nst, ac = it % Cfg::NACC;
        const uint32_t ph = (uint32_t)(it / nst) & 1u, aph = (uint32_t)(it / Cfg::NACC) & 1u;
        const int ksteps = (p.s[l] + 15) >> 4;
        CT_STAMP(1, it, 0);
        mbar_wait(AEMPTY + 8 * ac, aph ^ 1u);
        CT_STAMP(1, it, 1);
        mbar_wait(FULL + 8 * st, ph);
        CT_STAMP(1, it, 2);
        tc_fence_after();
        const uint32_t b0 = sB + st * Cfg::B_STAGE;
        const uint64_t aofs = (uint64_t)((p.kofs[l] * 2) >> 4);     // the layer's K range inside the A tile
        const uint64_t dAh = make_smem_desc(sA) + aofs, dAl = make_smem_desc(sA + CT_A_PLANE) + aofs;
        const uint64_t dBh = make_smem_desc(b0), dBl = make_smem_desc(b0 + Cfg::B_PLANE);
        const uint32_t d = tmem + (uint32_t)(ac * Cfg::ACC_STRIDE);
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t adv = (uint64_t)((k * 32) >> 4);
          umma_bf16(d, dAl + adv, dBh + adv, idesc, k != 0);
          umma_bf16(d, dAh + adv, dBl + adv, idesc, 1u);
          umma_bf16(d, dAh + adv, dBh + adv, idesc, 1u);
        }
        umma_commit(EMPTY + 8 * st);     // B stage free once these MMAs retire
        umma_commit(AFULL + 8 * ac);     // ... and the logits of the row are complete
        CT_STAMP(1, it, 3);
      }
    }
    __syncwarp();
  } else {
    // =========================================================== epilogue: thread = (pixel = TMEM lane, token half)
    const int half = warp >> 2;                                // tokens [half * HALF, half * HALF + HALF)
    const int pix = (warp & 3) * 32 + lane;
    const int n0 = half * HALF;
    const int nv = min(HALF, max(0, N - n0));                  // real tokens of this half
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)n0;
    float2 acc[MODE == 1 ? HALF / 2 : 1];
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < HALF / 2; ++i) acc[i] = make_float2(0.f, 0.f);
    }
    for (int it = 0; it < n_items; ++it) {
      int l, h, Y, iy;
      decode(it, l, h, Y, iy);
      const int ac = it % Cfg::NACC;
      const uint32_t aph = (uint32_t)(it / Cfg::NACC) & 1u;
      if (threadIdx.x == 0) CT_STAMP(2, it, 0);
      mbar_wait(AFULL + 8 * ac, aph);
      if (threadIdx.x == 0) CT_STAMP(2, it, 1);
      tc_fence_after();
      float x[HALF];
#pragma unroll
      for (int c = 0; c < HALF / 16; ++c) tmem_ld16_nowait(lane_base + (uint32_t)(ac * Cfg::ACC_STRIDE + 16 * c), x + 16 * c);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) ct_mbar_arrive(AEMPTY + 8 * ac);   // the accumulator may be overwritten: its logits live in registers now
      if (threadIdx.x == 0) CT_STAMP(2, it, 2);
      // exact softmax statistics of this half, four independent chains (padding tokens sit at -8192 * log2 e: exp2 -> 0)
      float m4[4] = {x[0], x[1], x[2], x[3]};
#pragma unroll
      for (int i = 4; i < HALF; ++i) m4[i & 3] = fmaxf(m4[i & 3], x[i]);
      const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      float2 e[HALF / 2];
      float2 s2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      const float2 negm = make_float2(-m, -m);
#pragma unroll
      for (int i = 0; i < HALF / 2; ++i) {
        if (2 * i < nv) {          // uniform: tokens beyond N (padding, exp2 == 0) take no SFU slot
          const float2 t = ct_add2(make_float2(x[2 * i], x[2 * i + 1]), negm);
          e[i] = make_float2(ct_ex2(t.x), ct_ex2(t.y));
          s2[i & 1] = ct_add2(s2[i & 1], e[i]);
        } else {
          e[i] = make_float2(0.f, 0.f);
        }
      }
      const float sum = (s2[0].x + s2[0].y) + (s2[1].x + s2[1].y);
      // the two halves of the pixel exchange (max, sum); the barrier also orders the staging buffer hand-over below
      float2* xc = xch + (size_t)(it & 1) * 256;
      xc[half * 128 + pix] = make_float2(m, sum);
      if (threadIdx.x == 0) CT_STAMP(2, it, 3);
      if (threadIdx.x == 0) CT_STAMP(2, it, 4);
      ct_named_bar(1, 256);
      if (threadIdx.x == 0) CT_STAMP(2, it, 5);
      const float2 o = xc[(half ^ 1) * 128 + pix];
      const float M = fmaxf(m, o.x);
      const float mine = ct_ex2(m - M);
      const float f = __fdividef(mine, fmaf(sum, mine, o.y * ct_ex2(o.x - M)));
      if (MODE == 1) {
        const float wi = f * p.w;
        const float2 w2 = make_float2(wi, wi);
#pragma unroll
        for (int i = 0; i < HALF / 2; ++i) acc[i] = ct_fma2(e[i], w2, acc[i]);
      } else {
        float* stg = staging + (size_t)(p.nstg == 2 ? (it & 1) : 0) * stage_floats;
        const float2 f2 = make_float2(f, f);
        // the buffer is free once the store warp saw the copy engine finish reading its previous row
        if (p.nstg == 2) mbar_wait(SFREE + 8 * (it & 1), (((uint32_t)it >> 1) & 1u) ^ 1u);
        else if (it > 0) mbar_wait(SFREE + 8 * ((it - 1) & 1), ((uint32_t)(it - 1) >> 1) & 1u);
        if (pix < cnt) {
          float* orow = stg + (size_t)pix * N + n0;
          if ((N & 3) == 0) {
#pragma unroll
            for (int i = 0; i < HALF / 4; ++i)
              if (4 * i < nv) {
                const float2 a = ct_mul2(e[2 * i], f2), b = ct_mul2(e[2 * i + 1], f2);
                *reinterpret_cast<float4*>(orow + 4 * i) = make_float4(a.x, a.y, b.x, b.y);
              }
          } else {
#pragma unroll
            for (int i = 0; i < HALF / 2; ++i) {
              const float2 a = ct_mul2(e[i], f2);
              if (2 * i < nv) orow[2 * i] = a.x;
              if (2 * i + 1 < nv) orow[2 * i + 1] = a.y;
            }
          }
        }
        if (threadIdx.x == 0) CT_STAMP(2, it, 6);
        float* dst = p.out + (((size_t)h * R + Y) * R + x0) * N;
        const size_t bytes = (size_t)cnt * N * sizeof(float);
        const bool al16 = ((bytes & 15) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
        if (al16 && p.store_path != 1) {
          // hand the staged row to the store warp (which drives the copy engine) and move on
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) ct_mbar_arrive(SFULL + 8 * (it & 1));
        } else {
          // rows that are not 16-byte multiples (or store_path 1): plain coalesced stores from the staging tile
          ct_named_bar(2, 256);
          if (al16) {
            const int tot4 = (cnt * N) >> 2;
            const float4* s4v = reinterpret_cast<const float4*>(stg);
            float4* d4 = reinterpret_cast<float4*>(dst);
            for (int i = threadIdx.x; i < tot4; i += 256) d4[i] = s4v[i];
          } else {
            const int tot = cnt * N;
            for (int i = threadIdx.x; i < tot; i += 256) dst[i] = stg[i];
          }
          __syncwarp();
          if (lane == 0) ct_mbar_arrive(SFULL + 8 * (it & 1));     // the store warp only recycles the buffer (nothing to copy)
        }
      }
    }
    if (threadIdx.x == 0) CT_STAMP(2, 63, 7);
    if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
      unsigned long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      p.dbg[(2 * 64 + 62) * 8 + 2] = clock64();
      p.dbg[(2 * 64 + 62) * 8 + 3] = (long long)gt;
    }
    if (MODE == 1) {
      const int X = x0 + pix;
      if (X < R) {
        float* o = p.out + (size_t)n0 * R * R + (size_t)blockIdx.x * R + X;
#pragma unroll
        for (int i = 0; i < HALF / 2; ++i) {
          if (2 * i < nv) o[(size_t)(2 * i) * R * R] = acc[i].x;
          if (2 * i + 1 < nv) o[(size_t)(2 * i + 1) * R * R] = acc[i].y;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------ host
// 0: never; 1 (default): where it is the faster kernel (measured on B200, profiles/r02_attn_store_tc.md: always for the fused
// capture+collect, for the attn-store when N > 96); 2: wherever the shape is eligible (tests, A/B measurements)
static int g_capture_tc = getenv("SKP_CAPTURE_TC") != nullptr ? atoi(getenv("SKP_CAPTURE_TC")) : 1;
static long long* g_dbg = nullptr;
static int g_store_path = getenv("SKP_CAPTURE_TC_STORE") ? atoi(getenv("SKP_CAPTURE_TC_STORE")) : 0;

static int npt_of(int N) { return N <= 32 ? 32 : N <= 64 ? 64 : N <= 96 ? 96 : N <= 128 ? 128 : 0; }

template <int NPT, int MODE>
static size_t capture_tc_smem(const CapTcParams& p, int nst, int nstg, int* stage_floats) {
  using Cfg = CtCfg<NPT>;
  *stage_floats = MODE == 0 ? ((128 * p.N + 3) & ~3) : 0;
  return 1024 + (size_t)2 * CT_A_PLANE + (size_t)nst * Cfg::B_STAGE + (size_t)2 * p.slot_floats * 4 + 8 * Cfg::NBARS + 16 + 2 * 2 * 128 * 8 +
         (size_t)nstg * *stage_floats * sizeof(float) + 16;
}

template <int NPT, int MODE, int PW>
static int capture_tc_launch(CapTcParams& p, cudaStream_t st) {
  int stage_floats = 0;
  // shared-memory plan: prefer 3 B stages and 2 staging buffers, shrink until the CTA fits
  static const int plans[3][2] = {{3, 2}, {2, 2}, {2, 1}};
  size_t bytes = 0;
  bool ok = false;
  for (int i = 0; i < 3 && !ok; ++i) {
    p.nst = plans[i][0];
    p.nstg = MODE == 0 ? plans[i][1] : 1;
    bytes = capture_tc_smem<NPT, MODE>(p, p.nst, MODE == 0 ? p.nstg : 0, &stage_floats);
    ok = bytes <= 227 * 1024;
  }
  if (!ok) return SKP_ERR_UNSUPPORTED;     // caller falls back to the SIMT kernels
  cudaError_t e = cudaFuncSetAttribute(capture_tc_kernel<NPT, MODE, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) { set_error("capture_tc: smem attr: %s", cudaGetErrorString(e)); return SKP_ERR_LAUNCH; }
  dim3 grid;
  if (MODE == 0) {
    const int total = p.heads * p.R;
    int ctas = 148 / p.XT;
    if (ctas < 1) ctas = 1;
    if (ctas > total) ctas = total;
    p.rows_per_cta = (total + ctas - 1) / ctas;
    grid = dim3((total + p.rows_per_cta - 1) / p.rows_per_cta, p.XT);
  } else {
    grid = dim3(p.R, p.XT);
  }
  capture_tc_kernel<NPT, MODE, PW><<<grid, (9 + PW) * 32, bytes, st>>>(p, stage_floats);
  SKP_CHECK_LAUNCH("capture_tc");
  return SKP_OK;
}

bool capture_tc_eligible(const int* s, int n_layers, int N, int R, bool store) {
  if (!g_capture_tc || npt_of(N) == 0 || n_layers < 1 || n_layers > SKP_MAX_LAYERS || R < 1) return false;
  if (store && n_layers != 1) return false;
  if (store && g_capture_tc == 1) return false;   // store mode: the register kernel (skp_capture_store.cu) is faster at every N (cfg5: 116 vs 168 us)
  int slots[CT_MAX_SLOTS], ns = 0, kcols = 0;
  for (int l = 0; l < n_layers; ++l) {
    if (s[l] < 4 || s[l] > 32 || (s[l] & 3)) return false;   // source rows and their maxima travel as 16-byte multiples
    bool seen = false;
    for (int k = 0; k < ns; ++k) seen |= slots[k] == s[l];
    if (!seen) {
      if (ns == CT_MAX_SLOTS) return false;
      slots[ns++] = s[l];
      kcols += (s[l] + 15) & ~15;
    }
  }
  return kcols <= 64;                                      // the sides share the one 64-column A tile
}

size_t capture_tc_workspace(const int* s, int n_layers, int heads) {
  size_t fl = 0;
  for (int l = 0; l < n_layers; ++l) fl += (((size_t)heads * s[l] * s[l]) + 3) & ~(size_t)3;
  return fl * sizeof(float);
}

// *handled = false: shape not taken (caller falls back to the SIMT kernels)
int capture_tc(const float* const* logits, const int* s, int n_layers, float* out, int heads, int N, int R, bool store,
               float* workspace, cudaStream_t st, bool* handled) {
  *handled = false;
  if (workspace == nullptr || !capture_tc_eligible(s, n_layers, N, R, store)) return SKP_OK;
  if ((reinterpret_cast<uintptr_t>(workspace) & 15) != 0) return SKP_OK;
  for (int l = 0; l < n_layers; ++l)
    if ((reinterpret_cast<uintptr_t>(logits[l]) & 15) != 0) return SKP_OK;
  CapTcParams p{};
  p.n_layers = n_layers; p.heads = heads; p.N = N; p.R = R; p.XT = (R + 127) / 128;
  p.out = out; p.w = 1.f / (float)(n_layers * heads);
  p.store_path = g_store_path;
  p.dbg = g_dbg;
  p.n_slots = 0;
  int kcols = 0, max_s = 0;
  float* ws = workspace;
  CapRowmaxParams rp{};
  rp.n_layers = n_layers; rp.N = N;
  int rows_total = 0;
  for (int l = 0; l < n_layers; ++l) {
    p.logits[l] = logits[l];
    p.s[l] = s[l];
    p.rowmax[l] = ws;
    const int rows = heads * s[l] * s[l];
    rp.logits[l] = logits[l];
    rp.out[l] = ws;
    rows_total += rows;
    rp.row_end[l] = rows_total;
    ws += (rows + 3) & ~3;
    if (s[l] > max_s) max_s = s[l];
    int k = 0;
    for (; k < p.n_slots; ++k)
      if (p.slot_s[k] == s[l]) break;
    if (k == p.n_slots) {
      p.slot_s[k] = s[l];
      p.slot_kofs[k] = kcols;
      kcols += (s[l] + 15) & ~15;
      ++p.n_slots;
    }
    p.kofs[l] = p.slot_kofs[k];
  }
  p.slot_floats = 4 * max_s * N + 128;
  capture_rowmax_kernel<<<(int)(((size_t)rows_total * 32 + 255) / 256), 256, 0, st>>>(rp);
  SKP_CHECK_LAUNCH("capture_rowmax");
  const int npt = npt_of(N);
  int rc;
  // 16 warps: 8 epilogue, 6 producers, the DMA warp, the MMA warp (128 registers per thread)
  if (store) {
    switch (npt) {
      case 32: rc = capture_tc_launch<32, 0, 7>(p, st); break;
      case 64: rc = capture_tc_launch<64, 0, 7>(p, st); break;
      case 96: rc = capture_tc_launch<96, 0, 7>(p, st); break;
      default: rc = capture_tc_launch<128, 0, 7>(p, st); break;
    }
  } else {
    switch (npt) {
      case 32: rc = capture_tc_launch<32, 1, 7>(p, st); break;
      case 64: rc = capture_tc_launch<64, 1, 7>(p, st); break;
      case 96: rc = capture_tc_launch<96, 1, 7>(p, st); break;
      default: rc = capture_tc_launch<128, 1, 7>(p, st); break;
    }
  }
  if (rc == SKP_ERR_UNSUPPORTED) return SKP_OK;
  if (rc == SKP_OK) *handled = true;
  return rc;
}

void capture_tc_enable(int on) { g_capture_tc = on < 0 ? 0 : (on > 2 ? 2 : on); }
void capture_tc_debug(long long* buf) { g_dbg = buf; }

}  // namespace skp
