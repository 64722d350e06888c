// ONE-LAUNCH fused QuantLinear for sm_100a (SURVEY.md section 8b entry 3, north-star "activation-quantize -> INT8 tensor-core
// GEMM -> per-token x per-channel dequant -> bias / GELU / residual in one launch").
//
// Replaces one whole QuantLayer-family forward — reference qdiff/models/quant_layer.py:185-211 (and
// stdit_quant_layer.py:68-96, dit_quant_layer.py:18-29): DynamicActQuantizer.forward (dynamic_quantizer.py:16-45) on the
// live fp16 activations, the (pre-quantised) weight, F.linear, bias — optionally with LayerNorm + t2i_modulate
// (blocks.py:51, stdit.py:104,125) in front and GELU / the gated residual (stdit.py:109-127) behind — by ONE kernel:
//
//   warps 4..11 : (A) PRODUCERS.  A CTA owns a 128-row activation panel (K = 1152: 147 KB of u8 codes).  One warp per
//                 token row: coalesced 8-byte loads of the fp16 row (4 rows in flight per warp), [LayerNorm + modulate],
//                 [/ smooth], row min/max, the exact packed quantiser of vq_quant_common.cuh, and the codes are written
//                 straight into shared memory in the 128-byte-swizzled K-major layout tcgen05.mma reads (what a TMA box
//                 {128 B, 128 rows} with SWIZZLE_128B would have written) — activation codes never exist in HBM.  Per-row
//                 {delta, zero point, code sum} stay in shared memory for the epilogue.
//                 (B) then the same warps are the EPILOGUE: tcgen05.ld, integer zero-point correction, dequant,
//                 bias / GELU / res + gate * y, fp16, 32x32 staging tile -> TMA store.
//   warp 0      : TMA producer of the WEIGHT tiles only (4-stage ring; CTA pairs stage half a tile each, cta_group::2)
//   warp 1      : MMA issuer: waits for the panel (a_ready), then per n-tile 9 K-blocks x 4 tcgen05.mma.kind::i8 with the
//                 A descriptor pointing into the resident panel
//   warp 2      : TMEM owner (2 accumulator stages);  warp 3: per-tile column records, one tile ahead
// Work split: grid = (panel pairs) x n_split; each CTA (pair) quantises its panel once and walks a contiguous range of
// 192-column n-tiles, so small-M problems (PixArt 512: M = 2048; cross-attention kv_linear: M ~ 120..480) still fill the
// machine — the regime where the two-launch sequence is launch- and latency-bound.  At video sizes (M >= 16384) the
// stand-alone quantise pass + the persistent GEMM stay the better schedule: a 227 KB SM cannot double-buffer a 147 KB
// panel, so quantising (~7 us per panel) and the MMAs of the previous panel cannot overlap.
// Pooled batches (G = 2, 4: PixArt's CFG pair, quirk Q1) interleave: a panel holds 128 / G tokens x all G batch entries.
#include <stdlib.h>

#include "vq_gemm_common.cuh"
#include "vq_quant_common.cuh"

namespace vq {

constexpr int FL_KB = 9;                                   // K = 1152 = 9 K-blocks of 128 codes
constexpr int FL_K = FL_KB * BK;
constexpr int FL_PANEL_BYTES = FL_KB * A_STAGE_BYTES;      // 9 x [128 rows x 128 B], SWIZZLE_128B K-major
constexpr int FL_STAGES_PAIR = 4;                          // x 12 KB (half a 192-row weight tile per CTA)
constexpr int FL_STAGES_SINGLE = 2;                        // x 24 KB
constexpr int FL_B_BYTES = FL_STAGES_PAIR * B_PAIR_STAGE_BYTES;
static_assert(FL_B_BYTES == FL_STAGES_SINGLE * B_STAGE_BYTES, "weight ring size");
// W4 (packed INT4 weights, two codes per byte: low nibble = even k): the TMA stages PACKED half-width tiles, converter warps 2-3
// expand them into the u8 SWIZZLE_128B stage tcgen05.mma reads (kind::i8 is the only integer MMA kind; there is no INT4 kind)
constexpr int FL_W4_U8_STAGES_PAIR = 3, FL_W4_PK_STAGES_PAIR = 3;        // 3 x 12 KB u8 + 3 x 6 KB packed
constexpr int FL_W4_U8_STAGES_SINGLE = 1, FL_W4_PK_STAGES_SINGLE = 2;    // 1 x 24 KB u8 + 2 x 12 KB packed
constexpr int FL_B_BYTES_W4 = FL_W4_U8_STAGES_PAIR * B_PAIR_STAGE_BYTES + FL_W4_PK_STAGES_PAIR * (B_PAIR_STAGE_BYTES / 2);
static_assert(FL_W4_U8_STAGES_SINGLE * B_STAGE_BYTES + FL_W4_PK_STAGES_SINGLE * (B_STAGE_BYTES / 2) <= FL_B_BYTES_W4, "W4 ring");
constexpr int FL_EPI_BYTES = NUM_EPI_WARPS * EPI_BUF_BYTES;   // one 32 x 32 fp16 staging sub-tile per epilogue warp
constexpr int FL_ROWP_BYTES = BM * 16;                        // per panel row {delta, zero point, code sum, -}
constexpr int FL_SMEM_BYTES = FL_PANEL_BYTES + FL_B_BYTES + FL_EPI_BYTES + COLBUF_BYTES + FL_ROWP_BYTES + 512 + 1024;
constexpr int FL_SMEM_BYTES_W4 = FL_SMEM_BYTES - FL_B_BYTES + FL_B_BYTES_W4;
static_assert(FL_SMEM_BYTES_W4 <= 232448, "exceeds the 227 KB dynamic shared memory limit");
constexpr int FL_PROD_WARPS = NUM_EPI_WARPS;               // 8
constexpr int FL_ROWS_IN_FLIGHT = 4;

struct FusedArgs {
  const __half* x;          // [G * rows, K] fp16
  int G, rows, N;
  int tpp;                  // tokens per panel = 128 / G
  const __half* smooth;     // [K] or null
  const __half* shift;      // LN mode: [G * rows / rows_per_mod, K]
  const __half* scale;
  int rows_per_mod;
  float qmax;
  const VqColParam* col;
  const __half* res;        // gated residual: [G * rows, ldr]
  int ldr;
  const __half* gate;       // [G * rows / rows_per_gate, N]
  int rows_per_gate;
  __half* out_delta;        // optional [rows]
  __half* out_zp;
  uint32_t* status;
  int n_split;              // CTAs (pairs) sharing one panel (pair): each walks a contiguous range of n-tiles
  uint64_t store_policy;
};

struct RowParam {
  float dx;
  int32_t zx, rs, pad;
};

// codes of unit `u` (4 codes = K columns 128 u + 4 lane .. + 3) of panel row `i`: K-block u, byte 4 * lane of the 128-byte
// row, 16-byte chunk index XOR (row & 7) — CU_TENSOR_MAP_SWIZZLE_128B, 8-row groups 1024 B apart
__device__ __forceinline__ uint32_t panel_addr(uint32_t panel, int i, int u, int lane) {
  return panel + u * A_STAGE_BYTES + (i >> 3) * 1024 + (i & 7) * 128 +
         ((((lane >> 2) ^ (i & 7)) << 4) | ((lane & 3) << 2));
}

__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

template <bool LN>
__device__ __forceinline__ void fl_transform(UnitRegs<FL_KB>& regs, const FusedArgs& p, int grow, int lane) {
  if (LN) {
    const size_t mo = static_cast<size_t>(grow / p.rows_per_mod) * FL_K;
    uapply_ln_modulate<FL_KB>(regs, p.shift + mo, p.scale + mo, FL_K, lane);
  }
  if (p.smooth) uapply_smooth<FL_KB>(regs, p.smooth, lane);
}

// quantise one (transformed) row into panel row i; returns the warp-wide code sum
__device__ __forceinline__ int fl_quant_row(const UnitRegs<FL_KB>& r, uint32_t panel, int i, int lane, const QuantConsts& qc) {
  uint32_t sum = 0;
#pragma unroll
  for (int u = 0; u < FL_KB; ++u) {
    const __half2* x = reinterpret_cast<const __half2*>(&r.u[u]);
    const uint32_t w = __byte_perm(quant_pair(x[0], qc), quant_pair(x[1], qc), 0x6420);
    sum = __dp4a(w, 0x01010101u, sum);
    sts_u32(panel_addr(panel, i, u, lane), w);
  }
  return warp_sum_i(static_cast<int>(sum));
}

__device__ __forceinline__ void fl_zero_row(uint32_t panel, int i, int lane) {
#pragma unroll
  for (int u = 0; u < FL_KB; ++u) sts_u32(panel_addr(panel, i, u, lane), 0u);
}

template <int EPI, bool PAIR, bool LN, bool W4>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
vq_linear_fused_kernel(const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_out,
                       const FusedArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NS = W4 ? (PAIR ? FL_W4_U8_STAGES_PAIR : FL_W4_U8_STAGES_SINGLE) : (PAIR ? FL_STAGES_PAIR : FL_STAGES_SINGLE);
  constexpr int NSP = PAIR ? FL_W4_PK_STAGES_PAIR : FL_W4_PK_STAGES_SINGLE;   // packed ring (W4 only)
  constexpr int BSB = PAIR ? B_PAIR_STAGE_BYTES : B_STAGE_BYTES;
  constexpr int PKB = BSB / 2;                                                // bytes of one packed stage
  uint8_t* smem_panel = smem;
  uint8_t* smem_b = smem + FL_PANEL_BYTES;
  uint8_t* smem_pk = smem_b + NS * BSB;
  uint8_t* smem_epi = smem_b + (W4 ? FL_B_BYTES_W4 : FL_B_BYTES);
  int4* colbuf = reinterpret_cast<int4*>(smem_epi + FL_EPI_BYTES);
  RowParam* rowp = reinterpret_cast<RowParam*>(smem_epi + FL_EPI_BYTES + COLBUF_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + FL_EPI_BYTES + COLBUF_BYTES + FL_ROWP_BYTES);
  uint64_t* full_bar = bars;                       // [4]
  uint64_t* empty_bar = bars + 4;                  // [4]
  uint64_t* tfull_bar = bars + 8;                  // [2]
  uint64_t* tempty_bar = bars + 10;                // [2]
  uint64_t* colfull_bar = bars + 12;               // [2]
  uint64_t* colempty_bar = bars + 14;              // [2]
  uint64_t* a_ready_bar = bars + 16;               // the activation panel(s) of this CTA (pair) are quantised
  uint64_t* pfull_bar = bars + 18;                 // [4] W4: a packed tile has landed (TMA)
  uint64_t* pempty_bar = bars + 22;                // [4] W4: both converter warps are done with a packed tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int TILE_M = PAIR ? 2 * BM : BM;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const int cluster_id = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int pp = cluster_id / p.n_split;           // panel (pair) index
  const int split = cluster_id - pp * p.n_split;
  const int panel = PAIR ? 2 * pp + static_cast<int>(cta_rank) : pp;
  const int num_n_tiles = (p.N + BN - 1) / BN;
  const int nt0 = static_cast<int>((static_cast<long long>(split) * num_n_tiles) / p.n_split);
  const int nt1 = static_cast<int>((static_cast<long long>(split + 1) * num_n_tiles) / p.n_split);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_out);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 4; ++s) {
      // W4: the u8 stage is filled by the two converter warps of each CTA of the pair, not by the TMA
      mbar_init(&full_bar[s], W4 ? (PAIR ? 4 : 2) : 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&pfull_bar[s], 1);
      mbar_init(&pempty_bar[s], 2);
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], PAIR ? 2 * NUM_EPI_WARPS : NUM_EPI_WARPS);
      mbar_init(&colfull_bar[a], 1);
      mbar_init(&colempty_bar[a], NUM_EPI_WARPS);
    }
    mbar_init(a_ready_bar, PAIR ? 2 : 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    if (PAIR) {
      tmem_alloc_pair(tmem_slot, TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_sync();

  if (warp == 0) {
    // ===================== TMA producer: weight tiles =====================
    if (elect_one()) {
      int s = 0;
      uint32_t phase = 0;
      constexpr uint32_t kStageBytes = PAIR ? 2 * B_PAIR_STAGE_BYTES : B_STAGE_BYTES;
      for (int tile = nt0; tile < nt1; ++tile) {
        const int n_idx = tile * BN + (PAIR ? static_cast<int>(cta_rank) * (BN / 2) : 0);
        for (int kb = 0; kb < FL_KB; ++kb) {
          if (W4) {   // packed tile (64 bytes per row) into this CTA's packed ring, on its own barrier
            mbar_wait(&pempty_bar[s], phase ^ 1);
            mbar_arrive_expect_tx(&pfull_bar[s], PKB);
            tma_load_2d_hint(smem_pk + s * PKB, &tmap_b, &pfull_bar[s], kb * (BK / 2), n_idx, kEvictLast);
            if (++s == NSP) { s = 0; phase ^= 1; }
            continue;
          }
          mbar_wait(&empty_bar[s], phase ^ 1);
          if (PAIR) {
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
            tma_load_2d_pair(smem_b + s * BSB, &tmap_b, &full_bar[s], kb * BK, n_idx, kEvictLast);
          } else {
            mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
            tma_load_2d_hint(smem_b + s * BSB, &tmap_b, &full_bar[s], kb * BK, n_idx, kEvictLast);
          }
          if (++s == NS) { s = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if ((!PAIR || cta_rank == 0) && elect_one()) {
      constexpr uint32_t idesc = make_idesc_i8(TILE_M, BN, 0, 0);
      mbar_wait(a_ready_bar, 0);      // both CTAs' panels: generic-proxy writes + fence.proxy.async + (remote) arrive
      tc_fence_after();
      int s = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = nt0; tile < nt1; ++tile, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
        for (int kb = 0; kb < FL_KB; ++kb) {
          mbar_wait(&full_bar[s], phase);
          tc_fence_after();
          const uint64_t a_desc = make_kmajor_sw128_desc(smem_u32(smem_panel + kb * A_STAGE_BYTES));
          const uint64_t b_desc = make_kmajor_sw128_desc(smem_u32(smem_b + s * BSB));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            if (PAIR) tc_mma_i8_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            else tc_mma_i8(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if (PAIR) tc_commit_pair(&empty_bar[s], 0b11);
          else tc_commit(&empty_bar[s]);
          if (++s == NS) { s = 0; phase ^= 1; }
        }
        if (PAIR) tc_commit_pair(&tfull_bar[acc], 0b11);
        else tc_commit(&tfull_bar[acc]);
      }
    }
    __syncwarp();
  } else if (warp == 3 || (W4 && warp == 2)) {
    // ===================== column-record producer (warp 3) [+ W4: INT4 -> u8 converters (warps 2 and 3)] =====================
    const int4* colg = reinterpret_cast<const int4*>(p.col);
    const int nmax = p.N - 1;
    int local = 0;
    int us = 0, ps = 0;
    uint32_t uphase = 0, pphase = 0;
    for (int tile = nt0; tile < nt1; ++tile, ++local) {
      if (warp == 3) {
        const int b = local & 1;
        mbar_wait(&colempty_bar[b], ((local >> 1) & 1) ^ 1);
        const int n0 = tile * BN;
#pragma unroll
        for (int i = 0; i < BN / 64; ++i) {
          const int pr = lane + 32 * i;
          const int n = n0 + 2 * pr;
          const int4 r0 = __ldg(colg + (n < nmax ? n : nmax));
          const int4 r1 = __ldg(colg + (n + 1 < nmax ? n + 1 : nmax));
          colbuf[b * BN + 2 * pr] = make_int4(r0.x, r1.x, r0.y, r1.y);
          colbuf[b * BN + 2 * pr + 1] = make_int4(r0.z, r1.z, r0.w, r1.w);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&colfull_bar[b]);
      }
      if (W4) {
        // each K block: packed [rows][64 B] -> u8 [rows][128 B] in the SWIZZLE_128B pattern (16-byte chunk ^ (row & 7)).
        // One item = 8 packed bytes -> one 16-byte chunk of 16 codes; the two warps split the rows of the stage.
        constexpr int ROWS = BSB / BK;                  // 96 (pair) | 192
        constexpr int ITEMS_PER_WARP = ROWS * 8 / 2;
        const int cw = warp - 2;
        for (int kb = 0; kb < FL_KB; ++kb) {
          mbar_wait(&pfull_bar[ps], pphase);
          mbar_wait(&empty_bar[us], uphase ^ 1);        // the MMAs that read this u8 stage have retired
          const uint32_t src = smem_u32(smem_pk + ps * PKB);
          const uint32_t dst = smem_u32(smem_b + us * BSB);
#pragma unroll 4
          for (int t = 0; t < ITEMS_PER_WARP / 32; ++t) {
            const int idx = cw * ITEMS_PER_WARP + t * 32 + lane;
            const int row = idx >> 3, j = idx & 7;
            uint32_t px, py;
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(px), "=r"(py) : "r"(src + row * (BK / 2) + j * 8));
            const uint32_t lo0 = px & 0x0F0F0F0Fu, hi0 = (px >> 4) & 0x0F0F0F0Fu;
            const uint32_t lo1 = py & 0x0F0F0F0Fu, hi1 = (py >> 4) & 0x0F0F0F0Fu;
            sts_v4_addr(dst + row * BK + ((j ^ (row & 7)) << 4), __byte_perm(lo0, hi0, 0x5140), __byte_perm(lo0, hi0, 0x7362),
                        __byte_perm(lo1, hi1, 0x5140), __byte_perm(lo1, hi1, 0x7362));
          }
          fence_proxy_async_smem();                     // generic-proxy writes -> the tensor core's operand reads
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_leader(&full_bar[us]);
            else mbar_arrive(&full_bar[us]);
            mbar_arrive(&pempty_bar[ps]);
          }
          if (++us == NS) { us = 0; uphase ^= 1; }
          if (++ps == NSP) { ps = 0; pphase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== (A) producers: quantise the activation panel into shared memory =====================
    const uint32_t panel_s = smem_u32(smem_panel);
    const int w = warp - 4;
    const int tok_per_warp = p.tpp / FL_PROD_WARPS;          // 16 (G = 1), 8 (G = 2), 4 (G = 4)
    const int tok0 = panel * p.tpp;
    // Each warp owns tok_per_warp tokens x G batch entries = 16 rows, handled in batches of FL_ROWS_IN_FLIGHT = 4 rows that are
    // all loaded before the first one is touched and stay in registers (one pass, also for pooled statistics): row b of a
    // batch is (token j / G, batch entry j % G), j = j0 + b — G divides 4, so a token's entries never straddle two batches.
    for (int j0 = 0; j0 < tok_per_warp * p.G; j0 += FL_ROWS_IN_FLIGHT) {
      UnitRegs<FL_KB> regs[FL_ROWS_IN_FLIGHT];
      int grow[FL_ROWS_IN_FLIGHT], irow[FL_ROWS_IN_FLIGHT], tokn[FL_ROWS_IN_FLIGHT];
#pragma unroll
      for (int b = 0; b < FL_ROWS_IN_FLIGHT; ++b) {
        const int j = j0 + b;
        const int tl = w * tok_per_warp + j / p.G, g = j % p.G;
        tokn[b] = tok0 + tl;
        grow[b] = g * p.rows + tokn[b];
        irow[b] = g * p.tpp + tl;
        if (tokn[b] < p.rows) uload_row<FL_KB>(regs[b], p.x + static_cast<size_t>(grow[b]) * FL_K, lane);
      }
      float mn[FL_ROWS_IN_FLIGHT], mx[FL_ROWS_IN_FLIGHT];
#pragma unroll
      for (int b = 0; b < FL_ROWS_IN_FLIGHT; ++b) {
        mn[b] = mx[b] = 0.f;
        if (tokn[b] < p.rows) {
          fl_transform<LN>(regs[b], p, grow[b], lane);
          __half2 mn2 = __float2half2_rn(0.f), mx2 = mn2;   // the range always contains zero
          urow_minmax<FL_KB>(regs[b], mn2, mx2);
          warp_minmax(mn2, mx2, mn[b], mx[b]);
        }
      }
      if (p.G == 2) {         // statistics pooled over the batch entries of a token (quirk Q1)
        mn[0] = mn[1] = fminf(mn[0], mn[1]);  mx[0] = mx[1] = fmaxf(mx[0], mx[1]);
        mn[2] = mn[3] = fminf(mn[2], mn[3]);  mx[2] = mx[3] = fmaxf(mx[2], mx[3]);
      } else if (p.G == 4) {
        mn[0] = mn[1] = mn[2] = mn[3] = fminf(fminf(mn[0], mn[1]), fminf(mn[2], mn[3]));
        mx[0] = mx[1] = mx[2] = mx[3] = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      }
#pragma unroll
      for (int b = 0; b < FL_ROWS_IN_FLIGHT; ++b) {
        if (tokn[b] < p.rows) {
          const RowStats st = make_stats(mn[b], mx[b], p.qmax);
          const QuantConsts qc = make_consts(st.delta, st.zp, p.qmax);
          const int rs = fl_quant_row(regs[b], panel_s, irow[b], lane, qc);
          if (lane == 0) {
            rowp[irow[b]] = RowParam{st.delta, __float2int_rn(st.zp), rs, 0};
            if (st.degenerate && p.status) atomicOr(p.status, static_cast<uint32_t>(VQ_STATUS_EPS_DEGENERATE));
            if (split == 0 && p.out_delta && grow[b] == tokn[b]) {      // once per token (batch entry 0)
              p.out_delta[tokn[b]] = __float2half_rn(st.delta);
              p.out_zp[tokn[b]] = __float2half_rn(st.zp);
            }
          }
        } else {
          fl_zero_row(panel_s, irow[b], lane);
          if (lane == 0) rowp[irow[b]] = RowParam{0.f, 0, 0, 0};
        }
      }
    }
    // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads, then hand the panel over
    fence_proxy_async_smem();
    named_bar_sync(1, FL_PROD_WARPS * 32);
    if (warp == 4 && lane == 0) {
      if (PAIR) mbar_arrive_leader(a_ready_bar);
      else mbar_arrive(a_ready_bar);
    }

    // ===================== (B) epilogue =====================
    const int q = warp & 3;          // TMEM lane quarter
    const int h = (warp - 4) >> 2;   // column half
    uint8_t* stage = smem_epi + (warp - 4) * EPI_BUF_BYTES;
    const int i_own = q * 32 + lane;                         // this thread's panel row
    const RowParam rp = rowp[i_own];
    const int g_own = i_own / p.tpp;
    const int token = tok0 + (i_own - g_own * p.tpp);
    const bool row_ok = token < p.rows;
    const int grow = g_own * p.rows + (row_ok ? token : 0);  // global row of this thread
    const int i0 = q * 32;                                   // first panel row of the warp's strip (same batch entry)
    const int g0 = i0 / p.tpp;
    const int token0 = tok0 + (i0 - g0 * p.tpp);
    const bool strip_ok = token0 < p.rows;
    const int row0 = g0 * p.rows + token0;
    const __half* gate_row = (EPI == VQ_EPI_GATE_RESIDUAL)
                                 ? p.gate + static_cast<size_t>(grow / p.rows_per_gate) * p.N : nullptr;
    const __half* res_row = (EPI == VQ_EPI_GATE_RESIDUAL) ? p.res + static_cast<size_t>(grow) * p.ldr : nullptr;
    int local = 0;
    for (int tile = nt0; tile < nt1; ++tile, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int cbase = tile * BN + h * EPI_COLS;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      mbar_wait(&colfull_bar[acc], acc_phase);
      const uint32_t t_base = tmem_base + acc * ACC_COLS + (static_cast<uint32_t>(q * 32) << 16) + h * EPI_COLS;
      const int4* ctile = colbuf + acc * BN + h * EPI_COLS;
#pragma unroll 1
      for (int c = 0; c < EPI_NCHUNK; ++c) {
        const int col0 = cbase + c * EPI_CHUNK;
        const bool chunk_ok = strip_ok && col0 < p.N;
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_base + c * EPI_CHUNK, v);
        uint4 rv[4];
        if (EPI == VQ_EPI_GATE_RESIDUAL) {
          // residual chunk of this thread's row: 64 contiguous bytes, in flight under the TMEM load and the dequant math
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            rv[g4] = make_uint4(0, 0, 0, 0);
            if (chunk_ok && row_ok && col0 + g4 * 8 < p.N)
              rv[g4] = *reinterpret_cast<const uint4*>(res_row + col0 + g4 * 8);
          }
        }
        tmem_ld_wait();
        if (c == EPI_NCHUNK - 1) {
          // last TMEM read of this accumulator stage: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_leader(&tempty_bar[acc]);
            else mbar_arrive(&tempty_bar[acc]);
          }
        }
        uint32_t packed[16];
        dequant_chunk<EPI>(v, rp.zx, rp.rs, rp.dx, ctile + c * EPI_CHUNK, packed);
        if (EPI == VQ_EPI_GATE_RESIDUAL) {
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            const int n = col0 + g4 * 8;
            uint4 gv = make_uint4(0, 0, 0, 0);
            if (n < p.N) gv = __ldg(reinterpret_cast<const uint4*>(gate_row + n));
            const __half2* g2 = reinterpret_cast<const __half2*>(&gv);
            const __half2* r2 = reinterpret_cast<const __half2*>(&rv[g4]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __half2 y = *reinterpret_cast<__half2*>(&packed[g4 * 4 + e]);
              const __half2 o = __hadd2_rn(r2[e], __hmul2_rn(g2[e], y));   // two roundings, never one fp16 FMA
              packed[g4 * 4 + e] = *reinterpret_cast<const uint32_t*>(&o);
            }
          }
        }
        if (chunk_ok) {   // warp-uniform
          if (lane == 0) tma_store_wait_read<0>();   // the previous chunk's store has finished reading the staging tile
          __syncwarp();
          stage_chunk<VQ_EPI_BIAS>(GemmArgs{}, packed, 0, true, 0, stage, lane);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d_hint(&tmap_out, stage, col0, row0, p.store_policy);
            tma_store_commit();
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&colempty_bar[acc]);
    }
    if (lane == 0) tma_store_wait<0>();
    __syncwarp();
  }

  tc_fence_before();
  if (PAIR) cluster_sync();
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int EPI, bool PAIR, bool LN, bool W4>
static int launch_fused_impl(const CUtensorMap& tb, const CUtensorMap& to, const FusedArgs& args, int grid, cudaStream_t stream) {
  static bool attr_set[kMaxDevices] = {};
  const int dev = current_device();
  constexpr int SMEM = W4 ? FL_SMEM_BYTES_W4 : FL_SMEM_BYTES;
  if (!attr_set[dev]) {
    if (cudaFuncSetAttribute(vq_linear_fused_kernel<EPI, PAIR, LN, W4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             SMEM) != cudaSuccess)
      return VQ_ERR_LAUNCH;
    attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, vq_linear_fused_kernel<EPI, PAIR, LN, W4>, tb, to, args);
  return e == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
}

template <int EPI, bool W4>
static int launch_fused_w(const CUtensorMap& tb, const CUtensorMap& to, const FusedArgs& a, int grid, bool pair, bool ln,
                          cudaStream_t st) {
  if (pair) return ln ? launch_fused_impl<EPI, true, true, W4>(tb, to, a, grid, st) : launch_fused_impl<EPI, true, false, W4>(tb, to, a, grid, st);
  return ln ? launch_fused_impl<EPI, false, true, W4>(tb, to, a, grid, st) : launch_fused_impl<EPI, false, false, W4>(tb, to, a, grid, st);
}

template <int EPI>
static int launch_fused(const CUtensorMap& tb, const CUtensorMap& to, const FusedArgs& a, int grid, bool pair, bool ln, bool w4,
                        cudaStream_t st) {
  return w4 ? launch_fused_w<EPI, true>(tb, to, a, grid, pair, ln, st) : launch_fused_w<EPI, false>(tb, to, a, grid, pair, ln, st);
}

// two INT4 codes per byte, low nibble = even k (what the converter warps of vq_linear_fused_kernel<.., W4> expand)
__global__ void __launch_bounds__(256) vq_pack_u4_kernel(const uint8_t* __restrict__ codes, uint8_t* __restrict__ packed,
                                                         long long n_groups) {
  grid_dep_sync();
  const long long gidx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;   // 16 codes -> 8 bytes
  if (gidx >= n_groups) return;
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(codes) + gidx);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t o[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    // word = codes k0..k3 -> byte (k0 | k1 << 4), (k2 | k3 << 4)
    const uint32_t a = w[2 * h], b = w[2 * h + 1];
    const uint32_t pa = (a & 0x000F000Fu) | ((a >> 4) & 0x00F000F0u);    // bytes 0 and 2 hold the packed pairs
    const uint32_t pb = (b & 0x000F000Fu) | ((b >> 4) & 0x00F000F0u);
    o[h] = __byte_perm(pa, pb, 0x6420);
  }
  reinterpret_cast<uint2*>(packed)[gidx] = make_uint2(o[0], o[1]);
}

static inline int64_t align16(int64_t v) { return (v + 15) & ~int64_t(15); }

}  // namespace vq

extern "C" int64_t vq_linear_workspace_bytes(int G, int rows, int K) {
  if (G <= 0 || rows <= 0 || K <= 0) return 0;
  const int64_t M = static_cast<int64_t>(G) * rows;
  // codes | delta | zp | rowsum | sync words of the overlapped schedule (work-queue head + one flag per 256-row panel)
  return vq::align16(M * K) + 2 * vq::align16(2LL * rows) + vq::align16(4 * M) + vq::align16(4 * (2 + (M + 255) / 256));
}

static bool fused_shape_supported(int G, int rows, int K) {
  if (K != vq::FL_K) return false;
  if (G == 1) return true;
  if (G == 2 || G == 4) return (rows % (vq::BM / G)) == 0;
  return false;
}

// Which shapes vq_linear_w8a8 runs as the single fused kernel.  Measured on B200 (profiles/r02_s3_linear_bench.md): inside
// a replayed CUDA graph the two-launch sequence wins at every size (9.0 vs 21.9 us at M = 109, 14.9 vs 35.7 us at M = 2048,
// 68 vs 74 us at M = 16384) — the stand-alone quantise pass spreads its rows over 148 SMs x 32 warps with no redundancy,
// the fused kernel's 8 producer warps per SM re-quantise the panel in every CTA that shares it and cannot overlap the
// MMAs (one 147 KB panel per SM) — so the default is OFF; eager callers with single-panel problems may opt in.
// mode: 0 never (default), 1 whenever the shape is supported, -1 up to max_m rows; 2: the OVERLAPPED schedule — the
// persistent GEMM with quantiser warpgroups running ahead of its MMAs (vq_gemm_w8a8.cu, QPRO) for K = 1152, G = 1, M > 128,
// bias / gated-residual epilogue; other shapes as mode 0.  Env VQ_LINEAR_FUSED / VQ_LINEAR_FUSED_MAX_M give the initial values.
static int g_fused_mode = [] {
  const char* e = getenv("VQ_LINEAR_FUSED");
  return e ? atoi(e) : 0;
}();
static long long g_fused_max_m = [] {
  const char* e = getenv("VQ_LINEAR_FUSED_MAX_M");
  return e ? atoll(e) : 0LL;
}();

extern "C" int vq_linear_set_fused_policy(int mode, int64_t max_m) {
  if (mode < -1 || mode > 2 || max_m < 0) return VQ_ERR_ARG;
  g_fused_mode = mode;
  g_fused_max_m = max_m;
  return VQ_OK;
}

extern "C" int vq_linear_launch_count(int G, int rows, int K) {
  if (G <= 0 || rows <= 0 || K <= 0) return VQ_ERR_ARG;
  const long long M = static_cast<long long>(G) * rows;
  if (g_fused_mode == 2) return (G == 1 && K == vq::FL_K && M >= 1024) ? 1 : 2;   // overlapped schedule (epilogue checked at call)
  const bool fused = fused_shape_supported(G, rows, K) && g_fused_mode != 0 && (g_fused_mode == 1 || M <= g_fused_max_m);
  return fused ? 1 : 2;
}

static int linear_impl(const void* x, int G, int rows, int K, const void* smooth, const void* ln_shift,
                       const void* ln_scale, int rows_per_mod, int n_bits, const uint8_t* w_codes, bool w4,
                       const VqColParam* col, int N, int epi, const void* res, int ldr, const void* gate,
                       int rows_per_gate, void* out, int ldo, void* out_delta, void* out_zp, void* workspace,
                       int64_t workspace_bytes, uint32_t* status, void* stream) {
  using namespace vq;
  if (!x || !w_codes || !col || !out || G <= 0 || rows <= 0 || K <= 0 || N <= 0) return VQ_ERR_ARG;
  if ((K % 16) != 0 || (N % 8) != 0 || (ldo % 8) != 0 || n_bits < 2 || n_bits > 8) return VQ_ERR_ARG;
  if (epi < VQ_EPI_BIAS || epi > VQ_EPI_GATE_RESIDUAL) return VQ_ERR_ARG;
  if (epi == VQ_EPI_GATE_RESIDUAL && (!res || !gate || rows_per_gate <= 0 || (ldr % 8) != 0)) return VQ_ERR_ARG;
  const bool ln = ln_shift != nullptr || ln_scale != nullptr;
  if (ln && (!ln_shift || !ln_scale || rows_per_mod <= 0)) return VQ_ERR_ARG;
  if ((out_delta == nullptr) != (out_zp == nullptr)) return VQ_ERR_ARG;
  const long long M = static_cast<long long>(G) * rows;
  if (M > 0x7fffffffLL) return VQ_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool overlapped = g_fused_mode == 2 && !w4 && G == 1 && K == FL_K && M >= 1024 &&
                          (epi == VQ_EPI_BIAS || epi == VQ_EPI_GATE_RESIDUAL);
  const bool fused = g_fused_mode != 2 && vq_linear_launch_count(G, rows, K) == 1;
  if (!fused && w4) return VQ_ERR_UNSUPPORTED;   // packed INT4 weights exist for the fused kernel only: pass the u8 codes
  if (overlapped) {
    if (!workspace || workspace_bytes < vq_linear_workspace_bytes(G, rows, K)) return VQ_ERR_ARG;
    if (ln && (rows_per_mod > rows || (rows % rows_per_mod) != 0)) return VQ_ERR_ARG;
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    uint8_t* codes = ws;
    void* delta = ws + align16(M * K);
    void* zp = ws + align16(M * K) + align16(2LL * rows);
    int32_t* rowsum = reinterpret_cast<int32_t*>(ws + align16(M * K) + 2 * align16(2LL * rows));
    uint32_t* sync = reinterpret_cast<uint32_t*>(ws + align16(M * K) + 2 * align16(2LL * rows) + align16(4 * M));
    int rc = gemm_w8a8_qpro(x, ln_shift, ln_scale, rows_per_mod, smooth, n_bits, codes, delta, zp, rowsum, sync, w_codes, col,
                            static_cast<int>(M), N, K, epi, res, ldr, gate, rows_per_gate, out, ldo, status, st);
    if (rc != VQ_OK) return rc;
    if (out_delta) {
      if (cudaMemcpyAsync(out_delta, delta, 2 * static_cast<size_t>(rows), cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
          cudaMemcpyAsync(out_zp, zp, 2 * static_cast<size_t>(rows), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return VQ_ERR_LAUNCH;
    }
    return VQ_OK;
  }
  if (!fused) {
    // two launches through the caller's workspace: (LayerNorm + modulate +) quantise pass, then the persistent GEMM
    if (!workspace || workspace_bytes < vq_linear_workspace_bytes(G, rows, K)) return VQ_ERR_ARG;
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    uint8_t* codes = ws;
    void* delta = ws + align16(M * K);
    void* zp = ws + align16(M * K) + align16(2LL * rows);
    int32_t* rowsum = reinterpret_cast<int32_t*>(ws + align16(M * K) + 2 * align16(2LL * rows));
    int rc;
    if (ln) {
      if (rows_per_mod > rows || (rows % rows_per_mod) != 0) return VQ_ERR_ARG;
      rc = vq_ln_modulate_act_quant(x, ln_shift, ln_scale, smooth, G, rows, K, rows_per_mod, n_bits, nullptr, codes, delta,
                                    zp, rowsum, status, stream);
    } else {
      rc = vq_act_quant(x, G, rows, K, static_cast<int64_t>(rows) * K, K, smooth, n_bits, codes, delta, zp, rowsum, status,
                        stream);
    }
    if (rc != VQ_OK) return rc;
    if (out_delta) {
      if (cudaMemcpyAsync(out_delta, delta, 2 * static_cast<size_t>(rows), cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
          cudaMemcpyAsync(out_zp, zp, 2 * static_cast<size_t>(rows), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return VQ_ERR_LAUNCH;
    }
    return vq_gemm_w8a8(codes, delta, zp, rowsum, rows, w_codes, col, static_cast<int>(M), N, K, epi, res, ldr, gate,
                        rows_per_gate, out, ldo, stream);
  }
  const int tpp = BM / G;
  const int panels = (rows + tpp - 1) / tpp;
  const bool pair = panels > 1;
  CUtensorMap tb, to;
  int rc = w4 ? make_u8_tmap_ex(&tb, w_codes, static_cast<uint64_t>(N), static_cast<uint64_t>(K / 2),
                                static_cast<uint64_t>(K / 2), BK / 2, pair ? BN / 2 : BN, /*swizzle128=*/false)
              : make_u8_kmajor_tmap(&tb, w_codes, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(K),
                                    pair ? BN / 2 : BN);
  if (rc != VQ_OK) return rc;
  rc = make_f16_out_tmap(&to, out, static_cast<uint64_t>(M), static_cast<uint64_t>(N), static_cast<uint64_t>(ldo));
  if (rc != VQ_OK) return rc;
  FusedArgs a{};
  a.x = static_cast<const __half*>(x);
  a.G = G; a.rows = rows; a.N = N; a.tpp = tpp;
  a.smooth = static_cast<const __half*>(smooth);
  a.shift = static_cast<const __half*>(ln_shift);
  a.scale = static_cast<const __half*>(ln_scale);
  a.rows_per_mod = ln ? rows_per_mod : 1;
  a.qmax = static_cast<float>((1 << n_bits) - 1);
  a.col = col;
  a.res = static_cast<const __half*>(res);
  a.ldr = ldr;
  a.gate = static_cast<const __half*>(gate);
  a.rows_per_gate = rows_per_gate > 0 ? rows_per_gate : 1;
  a.out_delta = static_cast<__half*>(out_delta);
  a.out_zp = static_cast<__half*>(out_zp);
  a.status = status;
  a.store_policy = kEvictFirst;
  const int clusters = pair ? (panels + 1) / 2 : panels;
  const int slots = pair ? num_sms() / 2 : num_sms();
  const int num_n_tiles = (N + BN - 1) / BN;
  int n_split = slots / clusters;
  if (n_split < 1) n_split = 1;
  if (n_split > num_n_tiles) n_split = num_n_tiles;
  a.n_split = n_split;
  const int grid = clusters * n_split * (pair ? 2 : 1);
  switch (epi) {
    case VQ_EPI_BIAS: return launch_fused<VQ_EPI_BIAS>(tb, to, a, grid, pair, ln, w4, st);
    case VQ_EPI_GELU_TANH: return launch_fused<VQ_EPI_GELU_TANH>(tb, to, a, grid, pair, ln, w4, st);
    default: return launch_fused<VQ_EPI_GATE_RESIDUAL>(tb, to, a, grid, pair, ln, w4, st);
  }
}

extern "C" int vq_linear_w8a8(const void* x, int G, int rows, int K, const void* smooth, const void* ln_shift,
                              const void* ln_scale, int rows_per_mod, int n_bits, const uint8_t* w_codes,
                              const VqColParam* col, int N, int epi, const void* res, int ldr, const void* gate,
                              int rows_per_gate, void* out, int ldo, void* out_delta, void* out_zp, void* workspace,
                              int64_t workspace_bytes, uint32_t* status, void* stream) {
  return linear_impl(x, G, rows, K, smooth, ln_shift, ln_scale, rows_per_mod, n_bits, w_codes, false, col, N, epi, res, ldr,
                     gate, rows_per_gate, out, ldo, out_delta, out_zp, workspace, workspace_bytes, status, stream);
}

extern "C" int vq_linear_w4a8(const void* x, int G, int rows, int K, const void* smooth, const void* ln_shift,
                              const void* ln_scale, int rows_per_mod, int n_bits, const uint8_t* w_packed,
                              const VqColParam* col, int N, int epi, const void* res, int ldr, const void* gate,
                              int rows_per_gate, void* out, int ldo, void* out_delta, void* out_zp, uint32_t* status,
                              void* stream) {
  return linear_impl(x, G, rows, K, smooth, ln_shift, ln_scale, rows_per_mod, n_bits, w_packed, true, col, N, epi, res, ldr,
                     gate, rows_per_gate, out, ldo, out_delta, out_zp, nullptr, 0, status, stream);
}

extern "C" int vq_pack_u4(const uint8_t* codes, int N, int K, uint8_t* packed, void* stream) {
  using namespace vq;
  if (!codes || !packed || N <= 0 || K <= 0 || (K % 16) != 0) return VQ_ERR_ARG;
  const long long groups = static_cast<long long>(N) * K / 16;
  launch_pdl(vq_pack_u4_kernel, dim3(static_cast<unsigned>((groups + 255) / 256)), dim3(256), 0,
             static_cast<cudaStream_t>(stream), codes, packed, groups);
  return cudaGetLastError() == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
}
