// Spatial self-attention of STDiT / PixArt on tcgen05 (reference: t2v/opensora/models/layers/blocks.py:151-195 on the
// "(B T) S C" view, stdit.py:104-109; flash-attn / xformers there).  fp16 Q/K/V, fp32 scores and softmax statistics,
// fp16 probabilities, fp32 output accumulation: the arithmetic of the flash kernels the reference calls.
//
// One sequence = S tokens of one frame, 16 heads of 72 dims.  q|k|v are read IN PLACE from the fused q|k|v GEMM output
// [n_seq * S, 3 * H * 72] through 4-D TMA tensor maps (dim, head, {q,k,v}, token); the output is written token-major
// [n_seq * S, H * 72], i.e. already in the layout the projection's quantiser reads (the reference's
// transpose(1, 2).reshape copy, blocks.py:189-191, does not exist).
//
// Work item = (sequence, head, 256 queries): two 128-query tiles share one stream of 64-key K / V tiles, which halves the
// L2 -> shared-memory operand traffic (at 128 queries per K/V stream the kernel would sit on the L2 bandwidth roof).
// Persistent CTAs (one per SM), 12 warps:
//   warp 0      TMA producer: Q tiles of the item, then a 4-stage ring of K tiles and one of V tiles.  head_dim 72 of Q / K is
//               stored as a 64-dim SWIZZLE_128B tile plus a 16-dim SWIZZLE_32B tile whose dims 72..79 lie outside the
//               tensor map's innermost extent and are zero-filled by the TMA unit (K = 80 for the MMA, no padded copy);
//               V as five 16-dim SWIZZLE_32B boxes = one MN-major operand of N = 80.
//   warp 1      MMA issuer (warp-uniform control flow, one elected lane issues; descriptors live in uniform registers).
//               S = Q K^T: 5 x tcgen05.mma.kind::f16 (M128 N64 K16) with **Q read from TMEM** as the A operand (copied
//               there once per item by the softmax warps: from shared memory every N = 64 instruction re-read 4 KB of Q
//               and was bound by the shared-memory port) into one of TWO score buffers per query tile, so the scores run
//               two key tiles ahead of the softmax.  O += P V: P is read from TMEM as the A operand too (it overwrites S in
//               place, two fp16 per column), V is the MN-major B operand straight from the TMA tile: one N = 80
//               instruction per 16 keys.
//   warp 2      TMEM allocation (512 columns: S/P 2 tiles x 2 buffers x 64, O 2 x 80, Q 2 x 40).
//   warps 4-7   softmax of query tile 0, one thread per query row (tcgen05.ld 32x32b: lane = row, no shuffles);
//   warps 8-11  softmax of query tile 1.
// Online softmax with a lazy rescale: the running maximum is only raised (and O / the row sum rescaled in TMEM) when a
// row's new maximum exceeds the one in use by more than 2^8 — after the first K tiles that is rare, so the O round trip
// through registers disappears from the steady state.  exp2 on MUFU with the log2(e) / sqrt(d) factor folded into one
// FFMA; a quarter of the exponentials take an FMA-pipe polynomial instead.
// Measured (B200, 32 sequences x 16 heads x 1024^2): 215 us = 718 TFLOP/s at d = 72.  ncu: MUFU pipe
// 45 % busy, tensor-core pipe 53 % — the two add up to the whole kernel: softmax and MMA phases of the two tiles do not
// overlap (both tiles run in lock-step, and a softmax warp spends ~850 cycles per 64 keys in its TMEM / barrier dependency
// chain).  Bring-up variants that tried to break this are listed in DESIGN.md; all land within 5 % of each other.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "vq_internal.h"
#include "vq_ptx.cuh"

namespace vq {

constexpr int SA_D = 72;
constexpr int SA_BM = 128;                 // queries per tile
constexpr int SA_BN = 64;                  // keys per tile
constexpr int SA_QT = 2;                   // query tiles per item
constexpr int SA_STAGES = 4;               // K ring and V ring depth
constexpr int SA_TILE_A = SA_BM * 128;     // Q: dims 0..63: 128-byte rows, SWIZZLE_128B
constexpr int SA_TILE_B = SA_BM * 32;      // Q: dims 64..79: 32-byte rows, SWIZZLE_32B (72..79 zero)
constexpr int SA_TILE = SA_TILE_A + SA_TILE_B;
constexpr int SA_KV_A = SA_BN * 128;       // K tile: same two pieces with 64 rows; V tile: five 32-byte-row boxes, same bytes
constexpr int SA_KV_B = SA_BN * 32;
constexpr int SA_KV = SA_KV_A + SA_KV_B;
constexpr int SA_OSTAGE = 128 * SA_D * 2;  // dense [128][72] fp16 staging tile of the output
constexpr int SA_SMEM_K = SA_QT * SA_TILE;
constexpr int SA_SMEM_V = SA_SMEM_K + SA_STAGES * SA_KV;
constexpr int SA_SMEM_O = SA_SMEM_V + SA_STAGES * SA_KV;
constexpr int SA_SMEM_BAR = SA_SMEM_O + SA_QT * SA_OSTAGE;
constexpr int SA_SMEM_BYTES = SA_SMEM_BAR + 512 + 1024;   // barriers + alignment slack
constexpr int SA_THREADS = 384;
constexpr uint32_t SA_TMEM_COLS = 512;
constexpr uint32_t SA_O_COL = 256;         // O accumulators start here; S/P buffer b of tile t at t * 128 + b * 64
constexpr uint32_t SA_Q_COL = SA_O_COL + 80;   // Q of tile t (fp16 pairs, 40 columns) behind its O accumulator (80 columns)
constexpr float SA_RESCALE_LOG2 = 8.0f;    // lazy-rescale threshold (log2 units): P stays <= 2^8 in fp16

struct SpatialArgs {
  int n_seq, S, H;
  float scale_log2e;
  int debug;        // 0 = attention; 1 = P := 1; 2 = P := identity on the first key tile; 3 = also dump S of key tile 0
  float* dbg;       // debug 3: [items][2][128][128] raw scores of the first key tile
  uint32_t k_lbo;   // leading byte offset written into the K-major K descriptors (ignored by the hardware for swizzled K-major tiles)
  // cross attention (vq_attn_cross_tc): sequence = sample, S = image tokens per sample (queries), keys / values = that
  // sample's prompt rows [kv_start[b], kv_start[b] + kv_len[b]) of a separate packed k|v tensor, at most 128 of them
  int cross;
  const int* kv_start;
  const int* kv_len;
  int k_which, v_which;   // index of k / v along the {q,k,v} (self) or {k,v} (cross) dimension of the K/V tensor map
};

__device__ __forceinline__ float sa_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for a pair on the FMA pipe (no MUFU): round-to-nearest split x = n + f, |f| <= 0.5 with the 1.5 * 2^23 magic add,
// degree-3 minimax polynomial for 2^f (relative error 7.5e-5, a third of an fp16 half-ulp — P is rounded to fp16 next),
// then n is added into the exponent field.  One warp cannot issue MUFU.EX2 faster than every ~16 cycles, so moving a share of
// the exponentials here shortens every softmax iteration, not just the MUFU-bound ones.
__device__ __forceinline__ float2 sa_exp2_poly(float2 x) {
  const float magic = 12582912.0f;   // 1.5 * 2^23
  x.x = fmaxf(x.x, -125.0f);
  x.y = fmaxf(x.y, -125.0f);
  const float2 xr = __fadd2_rn(x, make_float2(magic, magic));
  const float2 xi = __fadd2_rn(xr, make_float2(-magic, -magic));
  const float2 f = __fadd2_rn(x, make_float2(-xi.x, -xi.y));
  float2 p = __ffma2_rn(make_float2(0.05517164617776871f, 0.05517164617776871f), f,
                        make_float2(0.2426111251115799f, 0.2426111251115799f));
  p = __ffma2_rn(p, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
  p = __ffma2_rn(p, f, make_float2(0.9999280571937561f, 0.9999280571937561f));
  float2 r;
  r.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(xr.x) << 23));
  r.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(xr.y) << 23));
  return r;
}
template <bool DBG, int EMU>
__global__ void __launch_bounds__(SA_THREADS, 1)
vq_attn_spatial_kernel(const __grid_constant__ CUtensorMap tmap_qa, const __grid_constant__ CUtensorMap tmap_qb,
                       const __grid_constant__ CUtensorMap tmap_ka, const __grid_constant__ CUtensorMap tmap_kb,
                       const __grid_constant__ CUtensorMap tmap_o, const SpatialArgs a) {
  extern __shared__ uint8_t sa_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sa_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + SA_SMEM_K;
  uint8_t* smem_v = smem + SA_SMEM_V;
  uint8_t* smem_o = smem + SA_SMEM_O;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SA_SMEM_BAR);
  uint64_t* q_full = bars;                     // Q tiles of the item landed
  uint64_t* q_empty = bars + 1;                // last S MMA of the item has read them
  uint64_t* k_full = bars + 2;                 // [SA_STAGES]
  uint64_t* k_empty = k_full + SA_STAGES;
  uint64_t* v_full = k_empty + SA_STAGES;
  uint64_t* v_empty = v_full + SA_STAGES;
  uint64_t* s_full = v_empty + SA_STAGES;      // [2 tiles][2 buffers] S is in TMEM
  // The softmax of a tile may run up to two key tiles ahead of what the MMA warp has observed (its scores are produced two
  // tiles ahead), so every softmax <-> MMA barrier exists once per score buffer: a barrier then never advances two phases
  // between two waits of its consumer (with one barrier per tile the parity wait aliases: deadlock / early pass).
  uint64_t* p_full = s_full + 2 * SA_QT;       // [2 tiles][2 buffers] P is in TMEM (and O rescaled if needed)
  uint64_t* o_full = p_full + 2 * SA_QT;       // [2 tiles][2 buffers] the P V reading that buffer has retired
  uint64_t* qt_full = o_full + 2 * SA_QT;      // [2 tiles] Q of tile t has been copied into TMEM (A operand of S = Q K^T)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qt_full + SA_QT);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nqp = a.S / (SA_QT * SA_BM);       // 256-query groups per sequence
  const int nkv = a.cross ? 2 : a.S / SA_BN;   // key tiles per sequence (even)
  const int num_items = a.n_seq * a.H * nqp;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qa);
    tma_prefetch_desc(&tmap_qb);
    tma_prefetch_desc(&tmap_ka);
    tma_prefetch_desc(&tmap_kb);
    tma_prefetch_desc(&tmap_o);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 8);   // all eight softmax warps have copied their Q rows out of shared memory
    mbar_init(&qt_full[0], 4);
    mbar_init(&qt_full[1], 4);
    for (int s = 0; s < SA_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < SA_QT; ++t) {
      for (int b = 0; b < 2; ++b) {
        mbar_init(&s_full[2 * t + b], 1);
        mbar_init(&p_full[2 * t + b], 4);   // one arrival per softmax warp of the tile
        mbar_init(&o_full[2 * t + b], 1);
      }
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, SA_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_sync();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint32_t kc = 0;
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int qp = item % nqp;
        const int h = (item / nqp) % a.H;
        const int seq = item / (nqp * a.H);
        const int row_q0 = seq * a.S + qp * (SA_QT * SA_BM);
        const int row_kv0 = a.cross ? a.kv_start[seq] : seq * a.S;
        mbar_wait(q_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, SA_QT * SA_TILE);
        for (int t = 0; t < SA_QT; ++t) {
          tma_load_4d_hint(smem_q + t * SA_TILE, &tmap_qa, q_full, 0, h, 0, row_q0 + t * SA_BM, kEvictFirst);
          tma_load_4d_hint(smem_q + t * SA_TILE + SA_TILE_A, &tmap_qb, q_full, 64, h, 0, row_q0 + t * SA_BM, kEvictFirst);
        }
        for (int j = 0; j < nkv; ++j, ++kc) {
          const int s = kc % SA_STAGES;
          const uint32_t ph = (kc / SA_STAGES) & 1;
          // K / V tiles of a (sequence, head) are re-read by the other query groups of that head: keep them in L2
          mbar_wait(&k_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[s], SA_KV);
          tma_load_4d_hint(smem_k + s * SA_KV, &tmap_ka, &k_full[s], 0, h, a.k_which, row_kv0 + j * SA_BN, kEvictLast);
          tma_load_4d_hint(smem_k + s * SA_KV + SA_KV_A, &tmap_kb, &k_full[s], 64, h, a.k_which, row_kv0 + j * SA_BN, kEvictLast);
          mbar_wait(&v_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&v_full[s], SA_KV);
          // V: five 16-dim SWIZZLE_32B boxes (2 KB each) = one MN-major operand of N = 80 for a single P V instruction
          // per 16 keys (dims 72..79 of the last box are zero fill)
#pragma unroll
          for (int db = 0; db < 5; ++db)
            tma_load_4d_hint(smem_v + s * SA_KV + db * SA_KV_B, &tmap_kb, &v_full[s], 16 * db, h, a.v_which,
                             row_kv0 + j * SA_BN, kEvictLast);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the (warp-uniform) control flow and polls the barriers; one elected lane issues.  Keeping loop
    // counters and shared-memory addresses warp-uniform lets ptxas build the matrix descriptors in uniform registers — with
    // the loop inside a single-lane branch every tcgen05.mma cost four R2UR round trips and the issuing thread, not the
    // tensor pipe or the MUFU, bounded the kernel (26 small MMAs per key-tile pair).
    constexpr uint32_t idesc_s = make_idesc_f16(SA_BM, SA_BN, 0, 0);
    constexpr uint32_t idesc_pv80 = make_idesc_f16(SA_BM, 80, 0, 1);   // B = V, MN-major
    constexpr uint64_t v_lbo_word = static_cast<uint64_t>((SA_KV_B >> 4) & 0x3FFFu) << 16;
    // descriptor = lo | hi << 32: lo = (addr >> 4) | (LBO >> 4) << 16, hi = (SBO >> 4) | version 1 << 14 | layout << 29
    constexpr uint64_t hi_sw128 = static_cast<uint64_t>((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
    constexpr uint64_t hi_sw32 = static_cast<uint64_t>((256u >> 4) | (1u << 14) | (6u << 29)) << 32;
    const uint32_t lbo = (a.k_lbo >> 4) << 16;
    const uint32_t k_lo = ((smem_u32(smem_k) & 0x3FFFFu) >> 4) | lbo;
    const uint32_t v_lo = (smem_u32(smem_v) & 0x3FFFFu) >> 4;
    // S[t][b] = Q[t] K^T: K dimension 80 = 4 steps of 16 inside the 128-byte swizzle rows + 1 step in the 32-byte tile
    auto issue_s = [&](int t, int b, int ks) {
      const uint32_t ka = k_lo + ks * (SA_KV >> 4);
      const uint32_t d = tmem_base + t * 128 + b * SA_BN;
      const uint32_t qt = tmem_base + SA_Q_COL + t * 128;   // Q[t] in TMEM: 40 columns, 8 per 16-dim step
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_f16_ts(d, qt + 8 * k, hi_sw128 | (ka + 2 * k), idesc_s, k > 0 ? 1u : 0u);
      tc_mma_f16_ts(d, qt + 32, hi_sw32 | (ka + (SA_KV_A >> 4)), idesc_s, 1u);
      tc_commit(&s_full[2 * t + b]);
    };
    // O[t] (+)= P[t][b] V: 4 steps of 16 keys
    auto issue_pv = [&](int t, int b, int vs, uint32_t acc) {
      const uint32_t va = v_lo + vs * (SA_KV >> 4);
      const uint32_t p = tmem_base + t * 128 + b * SA_BN;
      const uint32_t o = tmem_base + SA_O_COL + t * 128;
      // one N = 80 instruction per 16 keys: V is MN-major in five 16-dim SWIZZLE_32B atoms, 2 KB apart (leading byte
      // offset), 8-key groups 256 B apart (stride byte offset); a key step advances by 16 keys x 32 B
#pragma unroll
      for (int k = 0; k < SA_BN / 16; ++k)
        tc_mma_f16_ts(o, p + 8 * k, hi_sw32 | v_lbo_word | (va + k * (512 >> 4)), idesc_pv80,
                      (acc | static_cast<uint32_t>(k > 0)) ? 1u : 0u);
      tc_commit(&o_full[2 * t + b]);
    };
    uint32_t kc = 0, vc = 0;
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      mbar_wait(&qt_full[0], it & 1);
      mbar_wait(&qt_full[1], it & 1);
      // the score buffers run two key tiles ahead of the softmax: S_0 and S_1 first
      for (int i = 0; i < 2; ++i) {
        const int ks = kc % SA_STAGES;
        mbar_wait(&k_full[ks], (kc / SA_STAGES) & 1);
        tc_fence_after();
        if (elect_one()) {
          issue_s(0, i, ks);
          issue_s(1, i, ks);
          tc_commit(&k_empty[ks]);
        }
        __syncwarp();
        ++kc;
      }
      for (int j = 0; j < nkv; ++j) {
        const uint32_t bph = (static_cast<uint32_t>(it) * (nkv >> 1) + (j >> 1)) & 1;   // phase of the per-buffer barriers
        const int b = j & 1;
        const bool more = j + 2 < nkv;
        const int vs = vc % SA_STAGES;
        const int ks = kc % SA_STAGES;
        mbar_wait(&v_full[vs], (vc / SA_STAGES) & 1);
        if (more) mbar_wait(&k_full[ks], (kc / SA_STAGES) & 1);
#pragma unroll
        for (int t = 0; t < SA_QT; ++t) {
          mbar_wait(&p_full[2 * t + b], bph);
          tc_fence_after();
          if (elect_one()) {
            issue_pv(t, b, vs, j > 0 ? 1u : 0u);
            if (more) issue_s(t, b, ks);   // S_{j+2} into the buffer whose P this P V has just consumed (issue order)
          }
          __syncwarp();
        }
        if (elect_one()) {
          tc_commit(&v_empty[vs]);
          if (more) tc_commit(&k_empty[ks]);
        }
        __syncwarp();
        ++vc;
        if (more) ++kc;
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax / correction / epilogue =====================
    const int t = (warp - 4) >> 2;               // query tile of this warpgroup
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;               // query row inside the tile
    const int wg_thread = threadIdx.x - (4 + 4 * t) * 32;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_off + t * 128;
    const uint32_t o_addr = tmem_base + lane_off + SA_O_COL + t * 128;
    uint8_t* ostage = smem_o + t * SA_OSTAGE;
    const float c = a.scale_log2e;
    const float2 c2 = make_float2(c, c);
    float m_used = 0.f, l = 0.f;
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int klen = a.cross ? a.kv_len[item / (nqp * a.H)] : a.S;   // keys of this item's sequence
      {
        // Q[t] -> TMEM (the A operand of every S = Q K^T of this item): with Q read from shared memory each N = 64 score
        // MMA re-read 4 KB of it and sat on the shared-memory port; from TMEM the instruction is compute-bound.  The rows
        // are un-swizzled by hand (SWIZZLE_128B tile: 16-byte chunk c of row r sits at c ^ (r & 7); SWIZZLE_32B tile: at
        // c ^ ((r >> 2) & 1)).  The previous item's score MMAs have retired (this thread consumed their last result).
        mbar_wait(q_full, it & 1);
        const uint32_t qa = smem_u32(smem_q + t * SA_TILE), qb = qa + SA_TILE_A;
        uint32_t w[32], w8[8];
#pragma unroll
        for (int cch = 0; cch < 8; ++cch) {
          const int4 v = lds_v4_addr(qa + row * 128 + ((cch ^ (row & 7)) << 4));
          w[4 * cch] = v.x; w[4 * cch + 1] = v.y; w[4 * cch + 2] = v.z; w[4 * cch + 3] = v.w;
        }
#pragma unroll
        for (int cch = 0; cch < 2; ++cch) {
          const int4 v = lds_v4_addr(qb + row * 32 + ((cch ^ ((row >> 2) & 1)) << 4));
          w8[4 * cch] = v.x; w8[4 * cch + 1] = v.y; w8[4 * cch + 2] = v.z; w8[4 * cch + 3] = v.w;
        }
        const uint32_t qt = tmem_base + lane_off + SA_Q_COL + t * 128;
        tmem_st_32x32b_x32(qt, w);
        tmem_st_32x32b_x8(qt + 32, w8);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&qt_full[t]);
          mbar_arrive(q_empty);      // shared-memory Q may be refilled for the next item
        }
      }
      for (int j = 0; j < nkv; ++j) {
        const int b = j & 1;
        const uint32_t sa = s_addr + b * SA_BN;
        const uint32_t pairs = static_cast<uint32_t>(it) * (nkv >> 1);   // completions of every per-buffer barrier before this item
        mbar_wait(&s_full[2 * t + b], (pairs + (j >> 1)) & 1);
        tc_fence_after();
        uint32_t v0[32], v1[32];
        tmem_ld_32x32b_x32(sa, v0);
        tmem_ld_32x32b_x32(sa + 32, v1);
        tmem_ld_wait();
        const int n_valid = klen - j * SA_BN;
        if (n_valid < SA_BN) {
          // ragged prompt (cross attention): keys past the sample's length are other samples' rows or TMA zero fill —
          // score -inf, so that they neither raise the maximum nor get a probability
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i >= n_valid) v0[i] = 0xff800000u;
            if (32 + i >= n_valid) v1[i] = 0xff800000u;
          }
        }
        // ---- row maximum of the 64 scores (four independent chains)
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v0[i + 1])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(v0[i + 2]), __uint_as_float(v0[i + 3])));
          mx2 = fmaxf(mx2, fmaxf(__uint_as_float(v1[i]), __uint_as_float(v1[i + 1])));
          mx3 = fmaxf(mx3, fmaxf(__uint_as_float(v1[i + 2]), __uint_as_float(v1[i + 3])));
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        if (j == 0) {
          m_used = mx;
          l = 0.f;
        } else {
          const bool need = (mx - m_used) * c > SA_RESCALE_LOG2;
          if (__any_sync(0xffffffffu, need)) {
            // raise the maximum in use and rescale O and the row sum; rows that did not need it take the exact new maximum too
            const float m_new = fmaxf(m_used, mx);
            const float f = sa_exp2((m_used - m_new) * c);
            const float2 f2 = make_float2(f, f);
            m_used = m_new;
            l *= f;
            // the previous P V of this tile (key tile j-1, other buffer) has retired; the one before it certainly has
            // (S_j was issued after it and has completed), so this parity wait cannot alias
            mbar_wait(&o_full[2 * t + (b ^ 1)], (pairs + ((j - 1) >> 1)) & 1);
            tc_fence_after();
            // 80 accumulator columns in three sequential pieces (rare path: keep the live register set small)
#pragma unroll 1
            for (int piece = 0; piece < 3; ++piece) {
              uint32_t w[32];
              if (piece < 2) tmem_ld_32x32b_x32(o_addr + 32 * piece, w);
              else tmem_ld_32x32b_x16(o_addr + 64, *reinterpret_cast<uint32_t(*)[16]>(w));
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float2 x = __fmul2_rn(make_float2(__uint_as_float(w[i]), __uint_as_float(w[i + 1])), f2);
                w[i] = __float_as_uint(x.x);
                w[i + 1] = __float_as_uint(x.y);
              }
              if (piece < 2) tmem_st_32x32b_x32(o_addr + 32 * piece, w);
              else tmem_st_32x32b_x16(o_addr + 64, *reinterpret_cast<uint32_t(*)[16]>(w));
            }
            tmem_st_wait();
          }
        }
        // ---- P = exp2(S c - m c) as fp16 pairs, written over S (32 columns); row sum in fp32
        const float neg = -m_used * c;
        const float2 neg2 = make_float2(neg, neg);
        float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const uint32_t* src = i < 16 ? v0 : v1;
          const int k2 = 2 * (i & 15);
          const float2 sv = make_float2(__uint_as_float(src[k2]), __uint_as_float(src[k2 + 1]));
          const float2 x = __ffma2_rn(sv, c2, neg2);
          float e0, e1;
          if (((i * EMU) & 15) < EMU) {   // EMU of every 16 pairs take the FMA-pipe polynomial instead of MUFU
            const float2 e = sa_exp2_poly(x);
            e0 = e.x;
            e1 = e.y;
          } else {
            e0 = sa_exp2(x.x);
            e1 = sa_exp2(x.y);
          }
          if (DBG) {
            const int key = 2 * i;
            if (a.debug == 2) {
              e0 = (j < 2 && j * SA_BN + key == row) ? 1.0f : 0.0f;
              e1 = (j < 2 && j * SA_BN + key + 1 == row) ? 1.0f : 0.0f;
            }
            if (a.debug == 3 && j < 2) {
              float* d = a.dbg + ((static_cast<size_t>(item) * SA_QT + t) * SA_BM + row) * 128 + j * SA_BN + key;
              d[0] = sv.x;
              d[1] = sv.y;
            }
          }
          if (i & 1) lb = __fadd2_rn(lb, make_float2(e0, e1));
          else la = __fadd2_rn(la, make_float2(e0, e1));
          const __half2 hp = __floats2half2_rn(e0, e1);
          pk[i] = *reinterpret_cast<const uint32_t*>(&hp);
        }
        tmem_st_32x32b_x32(sa, pk);
        l += (la.x + la.y) + (lb.x + lb.y);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[2 * t + b]);
      }
      // ---- epilogue: O / l -> fp16 -> dense staging tile -> one TMA store of [128 queries x 72 dims] at (head, row)
      {
        // last key tile nkv-1 is odd: buffer 1, its (nkv/2)-th use in this item
        mbar_wait(&o_full[2 * t + 1], (static_cast<uint32_t>(it) * (nkv >> 1) + (nkv >> 1) - 1) & 1);
        tc_fence_after();
        uint32_t v0[32], v1[32], w[8];
        tmem_ld_32x32b_x32(o_addr, v0);
        tmem_ld_32x32b_x32(o_addr + 32, v1);
        tmem_ld_32x32b_x8(o_addr + 64, w);
        tmem_ld_wait();
        const float inv = __fdividef(1.0f, l);
        uint32_t pk[36];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          __half2 h0 = __floats2half2_rn(__uint_as_float(v0[2 * i]) * inv, __uint_as_float(v0[2 * i + 1]) * inv);
          __half2 h1 = __floats2half2_rn(__uint_as_float(v1[2 * i]) * inv, __uint_as_float(v1[2 * i + 1]) * inv);
          pk[i] = *reinterpret_cast<const uint32_t*>(&h0);
          pk[16 + i] = *reinterpret_cast<const uint32_t*>(&h1);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          __half2 h0 = __floats2half2_rn(__uint_as_float(w[2 * i]) * inv, __uint_as_float(w[2 * i + 1]) * inv);
          pk[32 + i] = *reinterpret_cast<const uint32_t*>(&h0);
        }
        // the previous item's store has finished reading the staging tile
        if (wg_thread == 0) tma_store_wait_read<0>();
        named_bar_sync(1 + t, 128);
        const uint32_t dst = smem_u32(ostage) + row * (SA_D * 2);   // 144-byte rows: 16-byte pieces of 8 consecutive rows hit 8 bank groups
#pragma unroll
        for (int i = 0; i < 9; ++i) sts_v4_addr(dst + 16 * i, pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        fence_proxy_async_smem();
        named_bar_sync(1 + t, 128);
        if (wg_thread == 0) {
          const int qp = item % nqp;
          const int h = (item / nqp) % a.H;
          const int seq = item / (nqp * a.H);
          tma_store_3d(&tmap_o, ostage, 0, h, seq * a.S + qp * (SA_QT * SA_BM) + t * SA_BM);
          tma_store_commit();
        }
      }
    }
    if (wg_thread == 0) tma_store_wait<0>();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, SA_TMEM_COLS);
  }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiledSA)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiledSA sa_encode_fn() {
  static PFN_encodeTiledSA fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) return nullptr;
  fn = reinterpret_cast<PFN_encodeTiledSA>(ptr);
  return fn;
}

// q|k|v as a 4-D tensor (dim 72, head H, {q,k,v} 3, token rows): box = box_d dims x 128 tokens of one (head, which).
// Dims past 72 are outside the innermost extent: the TMA unit fills them with zeros.
static int make_qkv_tmap(CUtensorMap* out, const void* base, uint64_t rows, int H, int n_which, uint32_t box_d,
                         uint32_t box_rows, CUtensorMapSwizzle sw) {
  PFN_encodeTiledSA enc = sa_encode_fn();
  if (!enc) return VQ_ERR_DRIVER;
  const uint64_t C = static_cast<uint64_t>(H) * SA_D;
  cuuint64_t gdim[4] = {SA_D, static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(n_which), rows};
  cuuint64_t gstride[3] = {SA_D * 2, C * 2, static_cast<cuuint64_t>(n_which) * C * 2};
  cuuint32_t box[4] = {box_d, 1u, 1u, box_rows};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? VQ_OK : VQ_ERR_TMAP;
}

// output [rows, H * 72] as (dim 72, head H, token rows): box = one head's 72 dims x 128 tokens, dense in shared memory
static int make_attn_out_tmap(CUtensorMap* out, const void* base, uint64_t rows, int H) {
  PFN_encodeTiledSA enc = sa_encode_fn();
  if (!enc) return VQ_ERR_DRIVER;
  const uint64_t C = static_cast<uint64_t>(H) * SA_D;
  cuuint64_t gdim[3] = {SA_D, static_cast<cuuint64_t>(H), rows};
  cuuint64_t gstride[2] = {SA_D * 2, C * 2};
  cuuint32_t box[3] = {SA_D, 1u, static_cast<cuuint32_t>(SA_BM)};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? VQ_OK : VQ_ERR_TMAP;
}

template <bool DBG, int EMU>
static int launch_spatial(const CUtensorMap& qa, const CUtensorMap& qb, const CUtensorMap& ka, const CUtensorMap& kb,
                          const CUtensorMap& to, const SpatialArgs& a, int grid, cudaStream_t st) {
  static bool attr_dev[kMaxDevices] = {};   // per-device function attribute (one process may drive several GPUs)
  bool& attr = attr_dev[current_device()];
  if (!attr) {
    if (cudaFuncSetAttribute(vq_attn_spatial_kernel<DBG, EMU>, cudaFuncAttributeMaxDynamicSharedMemorySize, SA_SMEM_BYTES) !=
        cudaSuccess)
      return VQ_ERR_LAUNCH;
    attr = true;
  }
  launch_pdl(vq_attn_spatial_kernel<DBG, EMU>, dim3(grid), dim3(SA_THREADS), SA_SMEM_BYTES, st, qa, qb, ka, kb, to, a);
  return cudaGetLastError() == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
}

// self attention: q_base = kv_base = fused q|k|v [n_seq * S, 3C]; cross attention: q [n_seq * S, C], kv [kv_rows, 2C]
static int attn_tc_impl(const void* q_base, const void* kv_base, void* out, int n_seq, int S, int H, int head_dim,
                        float scale, int cross, uint64_t kv_rows, const int* kv_start, const int* kv_len, int debug,
                        float* dbg, uint32_t v_lbo, void* stream) {
  if (!q_base || !kv_base || !out || n_seq <= 0 || S <= 0 || H <= 0) return VQ_ERR_ARG;
  if (head_dim != SA_D || (S % (SA_QT * SA_BM)) != 0 || S < 4 * SA_BN) return VQ_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(q_base) & 15) || (reinterpret_cast<uintptr_t>(kv_base) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 15))
    return VQ_ERR_ARG;
  const uint64_t rows = static_cast<uint64_t>(n_seq) * S;
  if (rows * 3 * H * SA_D >= (1ull << 40)) return VQ_ERR_UNSUPPORTED;
  const int nq = cross ? 1 : 3, nk = cross ? 2 : 3;
  if (!cross) kv_rows = rows;
  CUtensorMap qa, qb, ka, kb, to;
  int rc = make_qkv_tmap(&qa, q_base, rows, H, nq, 64, SA_BM, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != VQ_OK) return rc;
  rc = make_qkv_tmap(&qb, q_base, rows, H, nq, 16, SA_BM, CU_TENSOR_MAP_SWIZZLE_32B);
  if (rc != VQ_OK) return rc;
  rc = make_qkv_tmap(&ka, kv_base, kv_rows, H, nk, 64, SA_BN, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != VQ_OK) return rc;
  rc = make_qkv_tmap(&kb, kv_base, kv_rows, H, nk, 16, SA_BN, CU_TENSOR_MAP_SWIZZLE_32B);
  if (rc != VQ_OK) return rc;
  rc = make_attn_out_tmap(&to, out, rows, H);
  if (rc != VQ_OK) return rc;
  // share of the exponentials evaluated on the FMA pipe, in sixteenths (tuning knob)
  static const int emu = [] {
    const char* e = getenv("VQ_SA_EMU");
    return e ? atoi(e) : 4;
  }();
  SpatialArgs a{n_seq, S, H, scale * 1.4426950408889634f, debug, dbg, v_lbo, cross, kv_start, kv_len, cross ? 0 : 1,
                cross ? 1 : 2};
  const long long items = static_cast<long long>(n_seq) * H * (S / (SA_QT * SA_BM));
  const int grid = static_cast<int>(items < num_sms() ? items : num_sms());
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (debug) return launch_spatial<true, 4>(qa, qb, ka, kb, to, a, grid, st);
  switch (emu) {
    case 0: return launch_spatial<false, 0>(qa, qb, ka, kb, to, a, grid, st);
    case 8: return launch_spatial<false, 8>(qa, qb, ka, kb, to, a, grid, st);
    default: return launch_spatial<false, 4>(qa, qb, ka, kb, to, a, grid, st);
  }
}

// cross attention on the same kernel (declared in vq_internal.h, called by vq_attn_cross in vq_attention.cu)
int attn_cross_tc(const void* q, const void* kv, void* out, const int* kv_start, const int* kv_len, int B, int N, int H,
                  int head_dim, long long kv_rows, float scale, void* stream) {
  if (!kv_start || !kv_len || kv_rows <= 0) return VQ_ERR_ARG;
  return attn_tc_impl(q, kv, out, B, N, H, head_dim, scale, 1, static_cast<uint64_t>(kv_rows), kv_start, kv_len, 0,
                      nullptr, 16, stream);
}

}  // namespace vq

extern "C" int vq_attn_spatial(const void* qkv, void* out, int n_seq, int S, int H, int head_dim, float scale,
                               void* stream) {
  return vq::attn_tc_impl(qkv, qkv, out, n_seq, S, H, head_dim, scale, 0, 0, nullptr, nullptr, 0, nullptr, 16, stream);
}

// bring-up / bisection entry used by tools/attn_selftest.cu only (not part of the C ABI in include/viditq_b200.h)
extern "C" int vq_attn_spatial_debug(const void* qkv, void* out, int n_seq, int S, int H, int head_dim, float scale,
                                     int debug, float* dbg, unsigned v_lbo, void* stream) {
  return vq::attn_tc_impl(qkv, qkv, out, n_seq, S, H, head_dim, scale, 0, 0, nullptr, nullptr, debug, dbg, v_lbo, stream);
}
