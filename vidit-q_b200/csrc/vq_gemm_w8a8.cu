// W8A8 QuantLinear GEMM for sm_100a: u8 activation codes x u8 weight codes -> s32 (tcgen05.mma.kind::i8,
// accumulators in TMEM), fused zero-point correction + per-token x per-channel dequant + bias
// (+ GELU-tanh | + gate*y + residual) epilogue, fp16 out.
//
// Replaces, for one QuantLayer-family forward (reference qdiff/models/quant_layer.py:185-211,
// stdit_quant_layer.py:76-96, dit_quant_layer.py:18-29), the chain
//     x_hat = (x_q - zx) * dx ; w_hat = (w_q - zw) * dw ; F.linear(x_hat, w_hat, bias)
// by the algebraically identical integer form
//     out[m,n] = (sum_k xq[m,k] wq[n,k] - zx[m] * c1[n] - zw[n] * rowsum[m]) * dx[m] * dw[n] + bias[n]
// with c1[n] = sum_k wq[n,k] - K * zw[n]  (prepared once per weight by vq_prep_weight).
//
// Structure (one CTA per SM, persistent over 128x192 output tiles, m-fastest so co-resident CTAs share a B tile in L2):
//   warp 0      : TMA producer   (one elected lane; A box 128 rows x 128 B, B box 192 rows x 128 B, SWIZZLE_128B)
//   warp 1      : MMA issuer     (one elected lane; 4 x tcgen05.mma 128x192x32 per 128-byte K block)
//   warp 2      : TMEM allocator (512 columns = 2 accumulator stages x 256)
//   warp 3      : stages each tile's per-column dequant records in smem one tile ahead (mbarrier hand-off)
//   warps 4..11 : epilogue       (tcgen05.ld 32x32b.x32; warp%4 selects the TMEM lane quarter, (warp-4)/4 the column half)
// Pipelines: smem full/empty ring (4 stages x 40 KB; CTA pairs 6 x 28 KB) and TMEM full/empty (2 stages), all mbarrier based.
// Epilogue output path: registers -> per-warp 32x32 fp16 staging tile in smem (64B swizzle, bank-conflict free)
// -> TMA store (cp.async.bulk.tensor, coalesced, clipped at the M/N edges by the tensor map). For the gated-residual
// epilogue the residual tile is TMA-loaded into the same staging tile one chunk ahead (per-warp mbarriers), so the
// SM never issues row-scattered global loads; out may alias res (in-place residual stream).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "vq_gemm_common.cuh"
#include "vq_quant_common.cuh"

namespace vq {

// PAIR = true: launched as 2-CTA clusters; one output tile is 256 rows (128 per CTA) x 192 columns, the leader CTA
// (cluster rank 0) issues tcgen05.mma.cta_group::2 for both, every CTA TMA-loads its own 128 A rows and HALF of the
// B tile (96 rows) — 30 % less operand traffic per SM than two independent CTAs — and drains its own TMEM half.
// QPRO = 1 | 2 (CTA pairs, K = 1152): the activation codes are produced INSIDE the kernel.  Two extra warpgroups (warps
// 12-19) take fp16 rows from a global work queue in ascending order — the order the rasterisation consumes m-panels —,
// run the exact quantiser ([LayerNorm + modulate for QPRO = 2], [/ smooth], min / max, codes, row sum) and write codes and
// per-row parameters to an L2-resident scratch; every finished row is counted on its m-panel's flag.  The TMA producer
// acquires a panel's flag before loading its A tiles, the epilogue before reading its row parameters.  The quantise
// arithmetic thus overlaps the tensor pipe instead of running as a pass in front of the GEMM.  Register budgets per
// warpgroup via setmaxnreg (control 40, epilogue 144, quantisers 72).
template <int EPI, bool PAIR, int QPRO = 0>
__global__ void __launch_bounds__(QPRO ? QP_THREADS : GEMM_THREADS, 1)
vq_gemm_w8a8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                    const GemmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NS = PAIR ? PAIR_STAGES : STAGES;
  constexpr int BSB = PAIR ? B_PAIR_STAGE_BYTES : B_STAGE_BYTES;   // per-CTA bytes of one B stage
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + NS * A_STAGE_BYTES;
  uint8_t* smem_epi = smem + OPERAND_BYTES;
  int4* colbuf = reinterpret_cast<int4*>(smem_epi + EPI_STAGING_BYTES);   // [2][BN] records
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + EPI_STAGING_BYTES + COLBUF_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + MAX_STAGES;
  uint64_t* tfull_bar = bars + 2 * MAX_STAGES;
  uint64_t* tempty_bar = bars + 2 * MAX_STAGES + ACC_STAGES;
  uint64_t* res_bar = bars + 2 * MAX_STAGES + 2 * ACC_STAGES;   // [NUM_EPI_WARPS] residual strip landed
  uint64_t* colfull_bar = res_bar + NUM_EPI_WARPS;           // [2] column records of a tile are in colbuf[b]
  uint64_t* colempty_bar = colfull_bar + 2;                  // [2] all epilogue warps are done with colbuf[b]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(colempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile = (m-tile, n-tile); with PAIR an m-tile is 256 rows shared by the two CTAs of a cluster
  constexpr int TILE_M = PAIR ? 2 * BM : BM;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const int worker = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int num_workers = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int num_m_tiles = (p.M + TILE_M - 1) / TILE_M;
  const int num_n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = num_m_tiles * num_n_tiles;
  const int num_kb = (p.K + BK - 1) / BK;
  const int m_cta = static_cast<int>(cta_rank) * BM;   // this CTA's row offset inside a tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_out);
    if (EPI == VQ_EPI_GATE_RESIDUAL) tma_prefetch_desc(&tmap_res);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], PAIR ? 2 * NUM_EPI_WARPS : NUM_EPI_WARPS);   // leader collects both CTAs' epilogues
    }
    for (int i = 0; i < NUM_EPI_WARPS; ++i) mbar_init(&res_bar[i], 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&colfull_bar[b], 1);
      mbar_init(&colempty_bar[b], NUM_EPI_WARPS);
    }
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
  if (PAIR) cluster_sync();   // peer barriers must be initialised before remote arrives / multicast commits
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_sync();   // barriers, TMEM and descriptor prefetch above overlap the previous kernel's tail
  static_assert(QPRO == 0 || PAIR, "quantise-producer variant: CTA pairs only");

  // Role dispatch by warpGROUP first: with QPRO every group sets its register budget at the top of its own branch
  // (setmaxnreg dominates the code it governs; one instruction executed by all four warps of the group).
  if (warp < 4) {
  if (QPRO) setmaxnreg_dec<QP_REGS_CONTROL>();
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int s = 0;
      uint32_t phase = 0;
      constexpr uint32_t kStageBytes = PAIR ? 2 * (A_STAGE_BYTES + B_PAIR_STAGE_BYTES) : (A_STAGE_BYTES + B_STAGE_BYTES);
      for (int tile = worker; tile < num_tiles; tile += num_workers) {
        int tm, tn;
        tile_to_mn(tile, num_m_tiles, num_n_tiles, p.group_m, tm, tn);
        const int m_idx = tm * TILE_M + m_cta;
        const int n_idx = tn * BN + (PAIR ? static_cast<int>(cta_rank) * (BN / 2) : 0);
        if (QPRO) {
          // the quantiser warpgroups (of any CTA) have finished every row of this m-panel: acquire, then order the
          // generic-proxy writes of the codes before this thread's async-proxy (TMA) reads
          const int need = min(TILE_M, p.M - tm * TILE_M);
          wait_counter(p.q_sync + 1 + tm, static_cast<uint32_t>(need));
          fence_proxy_async_all();
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], phase ^ 1);
          // operand tiles are re-read by other CTAs (A by every n-tile, B by every m-tile): keep them in L2
          if (PAIR) {
            // both CTAs' bytes are accounted on the leader's full barrier, which the leader arms for the pair
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
            tma_load_2d_pair(smem_a + s * A_STAGE_BYTES, &tmap_a, &full_bar[s], kb * BK, m_idx, kEvictLast);
            tma_load_2d_pair(smem_b + s * BSB, &tmap_b, &full_bar[s], kb * BK, n_idx, kEvictLast);
          } else {
            mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
            tma_load_2d_hint(smem_a + s * A_STAGE_BYTES, &tmap_a, &full_bar[s], kb * BK, m_idx, kEvictLast);
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
      constexpr uint32_t idesc = make_idesc_i8(TILE_M, BN, /*a_signed=*/0, /*b_signed=*/0);
      int s = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], phase);
          tc_fence_after();
          const uint64_t a_desc = make_kmajor_sw128_desc(smem_u32(smem_a + s * A_STAGE_BYTES));
          const uint64_t b_desc = make_kmajor_sw128_desc(smem_u32(smem_b + s * BSB));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advancing K by 32 bytes inside the 128B swizzle row: +2 in the (addr >> 4) start-address field
            if (PAIR) tc_mma_i8_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            else tc_mma_i8(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // frees this smem stage (in both CTAs of a pair) once the MMAs above have read it
          if (PAIR) tc_commit_pair(&empty_bar[s], 0b11);
          else tc_commit(&empty_bar[s]);
          if (++s == NS) { s = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (PAIR) tc_commit_pair(&tfull_bar[acc], 0b11);
        else tc_commit(&tfull_bar[acc]);
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ===================== column-record producer =====================
    // stages each tile's 192 {c1, zw, dw, bias} records in shared memory one tile ahead of the epilogue warps
    const int4* colg = reinterpret_cast<const int4*>(p.col);
    const int nmax = p.N - 1;
    int local = 0;
    for (int tile = worker; tile < num_tiles; tile += num_workers, ++local) {
      const int b = local & 1;
      mbar_wait(&colempty_bar[b], ((local >> 1) & 1) ^ 1);
      int tm, tn;
      tile_to_mn(tile, num_m_tiles, num_n_tiles, p.group_m, tm, tn);
      const int n0 = tn * BN;
#pragma unroll
      for (int i = 0; i < BN / 64; ++i) {   // 96 column pairs per tile, 3 per lane
        const int pr = lane + 32 * i;
        const int n = n0 + 2 * pr;
        const int4 r0 = __ldg(colg + (n < nmax ? n : nmax));
        const int4 r1 = __ldg(colg + (n + 1 < nmax ? n + 1 : nmax));
        colbuf[b * BN + 2 * pr] = make_int4(r0.x, r1.x, r0.y, r1.y);       // {c1, c1', zw, zw'}
        colbuf[b * BN + 2 * pr + 1] = make_int4(r0.z, r1.z, r0.w, r1.w);   // {dw, dw', bias, bias'}
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&colfull_bar[b]);
    }
  }
  } else if (QPRO && warp >= 12) {
    // ===================== quantiser warpgroups (QPRO) =====================
    setmaxnreg_dec<QP_REGS_QUANT>();
    // work queue of row batches in ascending order; one warp per row, QP_ROWS rows in flight
    const int KQ = 9 * 128;
    for (;;) {
      int g0 = 0;
      if (lane == 0) g0 = static_cast<int>(atomicAdd(p.q_sync, static_cast<uint32_t>(QP_GRAB)));
      g0 = __shfl_sync(0xffffffffu, g0, 0);
      if (g0 >= p.M) break;
      int nvalid = 0;
#pragma unroll 1
      for (int r0 = g0; r0 < g0 + QP_GRAB && r0 < p.M; r0 += QP_ROWS) {
        UnitRegs<9> regs[QP_ROWS];
#pragma unroll
        for (int b = 0; b < QP_ROWS; ++b)
          if (r0 + b < p.M) uload_row<9>(regs[b], p.qx + static_cast<size_t>(r0 + b) * KQ, lane);
#pragma unroll
        for (int b = 0; b < QP_ROWS; ++b) {
          const int row = r0 + b;
          if (row < p.M) {
            if (QPRO == 2) {
              const size_t mo = static_cast<size_t>(row / p.q_rows_per_mod) * KQ;
              uapply_ln_modulate<9>(regs[b], p.q_shift + mo, p.q_scale + mo, KQ, lane);
            }
            if (p.q_smooth) uapply_smooth<9>(regs[b], p.q_smooth, lane);
            __half2 mn2 = __float2half2_rn(0.f), mx2 = mn2;   // the range always contains zero
            urow_minmax<9>(regs[b], mn2, mx2);
            float mn, mx;
            warp_minmax(mn2, mx2, mn, mx);
            const RowStats st = make_stats(mn, mx, p.q_qmax);
            const QuantConsts qc = make_consts(st.delta, st.zp, p.q_qmax);
            int sum = uquant_store_row<9>(regs[b], p.q_codes + static_cast<size_t>(row) * KQ, lane, qc);
            sum = warp_sum_i(sum);
            if (lane == 0) {
              const_cast<__half*>(p.a_delta)[row] = __float2half_rn(st.delta);
              const_cast<__half*>(p.a_zp)[row] = __float2half_rn(st.zp);
              const_cast<int32_t*>(p.a_rowsum)[row] = sum;
              if (st.degenerate && p.q_status) atomicOr(p.q_status, static_cast<uint32_t>(VQ_STATUS_EPS_DEGENERATE));
            }
            ++nvalid;
          }
        }
      }
      // all lanes' stores are ordered before lane 0's release by the warp barrier (causality order; the release is
      // cumulative); QP_GRAB divides TILE_M, so a grab lies in one m-panel
      __syncwarp();
      if (lane == 0) red_release_gpu_add(p.q_sync + 1 + g0 / TILE_M, static_cast<uint32_t>(nvalid));
    }
  } else {
    // ===================== epilogue (warps 4-11) =====================
    if (QPRO) setmaxnreg_inc<QP_REGS_EPILOGUE>();
    // Per tile and warp: 32 rows x 96 columns. The strip is dequantised into registers chunk by chunk (TMEM loads
    // software-pipelined), then staged and handed to the TMA with ONE proxy fence and three bulk stores per tile.
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int h = (warp - 4) >> 2;   // column half
    uint8_t* stage0 = smem_epi + (warp - 4) * EPI_NCHUNK * EPI_BUF_BYTES;
    uint64_t* my_res_bar = res_bar + (warp - 4);
    uint32_t res_uses = 0;
    int local = 0;
    // per-row dequant parameters {delta, zero point, row sum}, fetched one tile ahead
    struct RowP { float dx; int32_t zx, rs; };
    auto load_rowp = [&](int tile_) {
      RowP r;
      int tm_, tn_;
      tile_to_mn(tile_, num_m_tiles, num_n_tiles, p.group_m, tm_, tn_);
      const int row_ = tm_ * TILE_M + m_cta + q * 32 + lane;
      const int rc = row_ < p.M ? row_ : p.M - 1;
      const int sr = p.a_period >= p.M ? rc : rc % p.a_period;
      r.dx = __half2float(p.a_delta[sr]);
      r.zx = __float2int_rn(__half2float(p.a_zp[sr]));
      r.rs = p.a_rowsum[rc];
      return r;
    };
    // QPRO: the parameters are produced inside this kernel — acquire the m-panel's flag, then read them past the L1
    auto load_rowp_acquired = [&](int tile_) {
      RowP r;
      int tm_, tn_;
      tile_to_mn(tile_, num_m_tiles, num_n_tiles, p.group_m, tm_, tn_);
      wait_counter(p.q_sync + 1 + tm_, static_cast<uint32_t>(min(TILE_M, p.M - tm_ * TILE_M)));
      const int row_ = tm_ * TILE_M + m_cta + q * 32 + lane;
      const int rc = row_ < p.M ? row_ : p.M - 1;
      r.dx = __half2float(__ldcg(p.a_delta + rc));
      r.zx = __float2int_rn(__half2float(__ldcg(p.a_zp + rc)));
      r.rs = __ldcg(p.a_rowsum + rc);
      return r;
    };
    RowP rp_next = RowP{0.f, 0, 0};
    if (!QPRO) rp_next = load_rowp(worker < num_tiles ? worker : 0);
    for (int tile = worker; tile < num_tiles; tile += num_workers, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      int tm, tn;
      tile_to_mn(tile, num_m_tiles, num_n_tiles, p.group_m, tm, tn);
      const int m_idx = tm * TILE_M + m_cta;
      const int n_idx = tn * BN;
      const int row0 = m_idx + q * 32;
      const int row = row0 + lane;
      const bool row_ok = row < p.M;
      RowP rp = rp_next;
      if (QPRO) rp = load_rowp_acquired(tile);   // in flight long before the accumulator is complete
      else if (tile + num_workers < num_tiles) rp_next = load_rowp(tile + num_workers);
      const int cbase = n_idx + h * EPI_COLS;
      // active sub-tiles of this warp: rows in range and first column in range (N is a multiple of 8)
      int nact = 0;
      if (EPI != VQ_EPI_DEBUG_MAINLOOP && EPI != VQ_EPI_DEBUG_LOADS && EPI != VQ_EPI_DEBUG_MATH && row0 < p.M) {
#pragma unroll
        for (int c = 0; c < EPI_NCHUNK; ++c) nact += (cbase + c * EPI_CHUNK < p.N) ? 1 : 0;
      }

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + acc * ACC_COLS + (static_cast<uint32_t>(q * 32) << 16) + h * EPI_COLS;
      uint32_t v[2][32];
      tmem_ld_32x32b_x32(t_base, v[0]);
      // the previous tile's TMA stores have finished reading the staging strip (they were issued a whole tile ago)
      if (lane == 0 && nact > 0) {
        tma_store_wait_read<0>();
        if (EPI == VQ_EPI_GATE_RESIDUAL) {   // residual strip -> staging (same swizzle), landed on my_res_bar
          mbar_arrive_expect_tx(my_res_bar, nact * EPI_BUF_BYTES);
          for (int c = 0; c < nact; ++c)
            tma_load_2d_hint(stage0 + c * EPI_BUF_BYTES, &tmap_res, my_res_bar, cbase + c * EPI_CHUNK, row0, kEvictFirst);
        }
      }
      mbar_wait(&colfull_bar[acc], acc_phase);
      const int4* ctile = colbuf + acc * BN + h * EPI_COLS;
      uint32_t packed[EPI_NCHUNK][16];
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < EPI_NCHUNK; ++c) {
        // software pipeline: the TMEM load of chunk c+1 is in flight while chunk c is dequantised
        if (c + 1 < EPI_NCHUNK) {
          tmem_ld_32x32b_x32(t_base + (c + 1) * EPI_CHUNK, v[(c + 1) & 1]);
        } else {
          // the last TMEM load of this accumulator stage has completed: hand the stage back to the MMA warp now
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_leader(&tempty_bar[acc]);   // the leader's MMA thread waits for both CTAs
            else mbar_arrive(&tempty_bar[acc]);
          }
        }
        if (EPI == VQ_EPI_DEBUG_LOADS || EPI == VQ_EPI_DEBUG_STORES) {
#pragma unroll
          for (int j = 0; j < 16; ++j) packed[c][j] = v[c & 1][j] ^ v[c & 1][j + 16];
        } else if (EPI != VQ_EPI_DEBUG_MAINLOOP) {
          dequant_chunk<EPI>(v[c & 1], rp.zx, rp.rs, rp.dx, ctile + c * EPI_CHUNK, packed[c]);
        }
        if (c + 1 < EPI_NCHUNK) tmem_ld_wait();
      }
      // the column records of this tile are consumed: release them before touching shared staging
      __syncwarp();
      if (lane == 0) mbar_arrive(&colempty_bar[acc]);
      if (EPI == VQ_EPI_DEBUG_LOADS || EPI == VQ_EPI_DEBUG_MATH) {
        uint32_t acc_x = 0;
#pragma unroll
        for (int c = 0; c < EPI_NCHUNK; ++c)
#pragma unroll
          for (int j = 0; j < 16; ++j) acc_x ^= packed[c][j];
        if (acc_x == 0x9e3779b9u && p.M < 0) p.out[0] = __float2half(1.0f);   // never true: defeats dead-code elimination
      }
      if (nact > 0) {
        if (EPI == VQ_EPI_GATE_RESIDUAL) {
          mbar_wait(my_res_bar, res_uses & 1);
          ++res_uses;
        } else {
          __syncwarp();   // lane 0's wait_read above precedes every lane's staging writes
        }
#pragma unroll
        for (int c = 0; c < EPI_NCHUNK; ++c)
          if (c < nact) stage_chunk<EPI>(p, packed[c], row, row_ok, cbase + c * EPI_CHUNK, stage0 + c * EPI_BUF_BYTES, lane);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          // streaming output: evict-first so it does not push the operand tiles out of L2
          for (int c = 0; c < nact; ++c)
            tma_store_2d_hint(&tmap_out, stage0 + c * EPI_BUF_BYTES, cbase + c * EPI_CHUNK, row0, p.store_policy);
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait<0>();
    __syncwarp();
  }

  tc_fence_before();
  if (PAIR) cluster_sync();   // the peer may still address this CTA's barriers / shared memory
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  return fn;
}

// rows x cols_bytes u8 matrix, row pitch = pitch bytes; box = box_rows x 128 B, 128B swizzle, zero OOB fill.
int make_u8_kmajor_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch,
                        uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return VQ_ERR_DRIVER;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch};
  cuuint32_t box[2] = {128u, box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? VQ_OK : VQ_ERR_TMAP;
}

// general u8 matrix map: box = box_rows x box_cols bytes, 128-byte swizzle or none (dense rows: the packed INT4 weight tiles
// the converter warps of vq_linear_fused_kernel expand themselves)
int make_u8_tmap_ex(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t box_cols,
                    uint32_t box_rows, bool swizzle128) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return VQ_ERR_DRIVER;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? VQ_OK : VQ_ERR_TMAP;
}

// rows x cols fp16 matrix (row pitch ld elements); box = 32 rows x EPI_CHUNK cols, matching swizzle: the epilogue staging tile.
int make_f16_out_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return VQ_ERR_DRIVER;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(EPI_CHUNK), 32u};   // 32 columns = 64 B
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? VQ_OK : VQ_ERR_TMAP;
}

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev < 0 ? 0 : (dev >= kMaxDevices ? kMaxDevices - 1 : dev);
}

int num_sms() {   // per device ordinal: one process may drive several GPUs
  static int n[kMaxDevices] = {};
  const int dev = current_device();
  if (n[dev] == 0) cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
  return n[dev];
}

template <int EPI, bool PAIR>
static int launch_gemm_impl(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr,
                            const GemmArgs& args, int grid, cudaStream_t stream) {
  static bool attr_set[kMaxDevices] = {};   // the > 48 KB dynamic shared memory opt-in is a per-device function attribute
  const int dev = current_device();
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(vq_gemm_w8a8_kernel<EPI, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         SMEM_BYTES);
    if (e != cudaSuccess) return VQ_ERR_LAUNCH;
    attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, vq_gemm_w8a8_kernel<EPI, PAIR>, ta, tb, to, tr, args);
  return e == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
}

template <int EPI, int QPRO>
static int launch_gemm_qpro(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr,
                            const GemmArgs& args, int grid, cudaStream_t stream) {
  static bool attr_set[kMaxDevices] = {};
  const int dev = current_device();
  if (!attr_set[dev]) {
    if (cudaFuncSetAttribute(vq_gemm_w8a8_kernel<EPI, true, QPRO>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) !=
        cudaSuccess)
      return VQ_ERR_LAUNCH;
    attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(QP_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, vq_gemm_w8a8_kernel<EPI, true, QPRO>, ta, tb, to, tr, args);
  return e == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
}

// One QuantLinear with the quantise pass running INSIDE the GEMM kernel (quantiser warpgroups ahead of the MMAs).
// K = 1152, un-pooled statistics (one scale pair per row), M > 128, bias or gated-residual epilogue.  codes / delta / zp /
// rowsum / sync are caller scratch; sync (2 + ceil(M / 256) words) is zeroed here (a memset node in a captured graph).
int gemm_w8a8_qpro(const void* x, const void* shift, const void* scale, int rows_per_mod, const void* smooth, int n_bits,
                   uint8_t* codes, void* delta, void* zp, int32_t* rowsum, uint32_t* sync, const uint8_t* w_codes,
                   const VqColParam* col, int M, int N, int K, int epi, const void* res, int ldr, const void* gate,
                   int rows_per_gate, void* out, int ldo, uint32_t* status, cudaStream_t st) {
  if (K != 9 * 128 || M <= BM || (epi != VQ_EPI_BIAS && epi != VQ_EPI_GATE_RESIDUAL)) return VQ_ERR_UNSUPPORTED;
  const bool ln = shift != nullptr;
  CUtensorMap ta, tb, to;
  int rc = make_u8_kmajor_tmap(&ta, codes, (uint64_t)M, (uint64_t)K, (uint64_t)K, BM);
  if (rc != VQ_OK) return rc;
  rc = make_u8_kmajor_tmap(&tb, w_codes, (uint64_t)N, (uint64_t)K, (uint64_t)K, BN / 2);
  if (rc != VQ_OK) return rc;
  rc = make_f16_out_tmap(&to, out, (uint64_t)M, (uint64_t)N, (uint64_t)ldo);
  if (rc != VQ_OK) return rc;
  CUtensorMap tr = to;
  if (epi == VQ_EPI_GATE_RESIDUAL) {
    rc = make_f16_out_tmap(&tr, res, (uint64_t)M, (uint64_t)N, (uint64_t)ldr);
    if (rc != VQ_OK) return rc;
  }
  const int tile_m = 2 * BM;
  const int m_tiles = (M + tile_m - 1) / tile_m;
  if (cudaMemsetAsync(sync, 0, sizeof(uint32_t) * (2 + m_tiles), st) != cudaSuccess) return VQ_ERR_LAUNCH;
  GemmArgs args{};
  args.M = M; args.N = N; args.K = K;
  args.a_delta = static_cast<const __half*>(delta);
  args.a_zp = static_cast<const __half*>(zp);
  args.a_rowsum = rowsum;
  args.a_period = M;
  args.col = col;
  args.out = static_cast<__half*>(out);
  args.ldo = ldo;
  args.epi = epi;
  args.res = static_cast<const __half*>(res);
  args.ldr = ldr;
  args.gate = static_cast<const __half*>(gate);
  args.rows_per_gate = rows_per_gate > 0 ? rows_per_gate : 1;
  args.store_policy = kEvictFirst;
  static const int group_m_env = [] {
    const char* e = getenv("VQ_GEMM_GROUP_M");
    return e ? atoi(e) : 8;
  }();
  args.group_m = group_m_env <= 0 ? m_tiles : (group_m_env < m_tiles ? group_m_env : m_tiles);
  args.qx = static_cast<const __half*>(x);
  args.q_shift = static_cast<const __half*>(shift);
  args.q_scale = static_cast<const __half*>(scale);
  args.q_smooth = static_cast<const __half*>(smooth);
  args.q_rows_per_mod = rows_per_mod > 0 ? rows_per_mod : 1;
  args.q_qmax = static_cast<float>((1 << n_bits) - 1);
  args.q_codes = codes;
  args.q_sync = sync;
  args.q_status = status;
  const int tiles = m_tiles * ((N + BN - 1) / BN);
  const int workers = num_sms() / 2;
  const int grid = (tiles < workers ? tiles : workers) * 2;
  if (epi == VQ_EPI_BIAS)
    return ln ? launch_gemm_qpro<VQ_EPI_BIAS, 2>(ta, tb, to, tr, args, grid, st)
              : launch_gemm_qpro<VQ_EPI_BIAS, 1>(ta, tb, to, tr, args, grid, st);
  return ln ? launch_gemm_qpro<VQ_EPI_GATE_RESIDUAL, 2>(ta, tb, to, tr, args, grid, st)
            : launch_gemm_qpro<VQ_EPI_GATE_RESIDUAL, 1>(ta, tb, to, tr, args, grid, st);
}

template <int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr,
                       const GemmArgs& args, int grid, bool pair, cudaStream_t stream) {
  return pair ? launch_gemm_impl<EPI, true>(ta, tb, to, tr, args, grid, stream)
              : launch_gemm_impl<EPI, false>(ta, tb, to, tr, args, grid, stream);
}

}  // namespace vq

extern "C" int vq_gemm_w8a8(const uint8_t* a_codes, const void* a_delta, const void* a_zp, const int32_t* a_rowsum,
                            int a_rows_period, const uint8_t* w_codes, const VqColParam* col, int M, int N, int K, int epi,
                            const void* res, int ldr, const void* gate, int rows_per_gate, void* out, int ldo,
                            void* stream) {
  using namespace vq;
  if (M <= 0 || N <= 0 || K <= 0 || a_rows_period <= 0) return VQ_ERR_ARG;
  if ((K % 16) != 0 || (N % 8) != 0 || (ldo % 8) != 0 || !out) return VQ_ERR_ARG;
  if (epi == VQ_EPI_GATE_RESIDUAL && (!res || !gate || rows_per_gate <= 0 || (ldr % 8) != 0)) return VQ_ERR_ARG;
#ifdef VQ_DEBUG_EPI
  if (epi < 0 || epi > VQ_EPI_DEBUG_STORES || epi == 4) return VQ_ERR_ARG;
#else
  if (epi < 0 || epi > VQ_EPI_GATE_RESIDUAL) return VQ_ERR_ARG;   // the bisection epilogues exist in the -DVQ_DEBUG_EPI build only
#endif
  // CTA pairs (cta_group::2) whenever there is more than one 128-row tile; VQ_GEMM_PAIR=0 forces single-CTA tiles
  static const bool allow_pair = [] {
    const char* e = getenv("VQ_GEMM_PAIR");
    return !(e && e[0] == '0');
  }();
  const bool pair = allow_pair && M > BM;
  CUtensorMap ta, tb, to;
  int rc = make_u8_kmajor_tmap(&ta, a_codes, (uint64_t)M, (uint64_t)K, (uint64_t)K, BM);
  if (rc != VQ_OK) return rc;
  rc = make_u8_kmajor_tmap(&tb, w_codes, (uint64_t)N, (uint64_t)K, (uint64_t)K, pair ? BN / 2 : BN);
  if (rc != VQ_OK) return rc;
  rc = make_f16_out_tmap(&to, out, (uint64_t)M, (uint64_t)N, (uint64_t)ldo);
  if (rc != VQ_OK) return rc;
  CUtensorMap tr = to;
  if (epi == VQ_EPI_GATE_RESIDUAL) {
    rc = make_f16_out_tmap(&tr, res, (uint64_t)M, (uint64_t)N, (uint64_t)ldr);
    if (rc != VQ_OK) return rc;
  }
  GemmArgs args;
  args.M = M; args.N = N; args.K = K;
  args.a_delta = static_cast<const __half*>(a_delta);
  args.a_zp = static_cast<const __half*>(a_zp);
  args.a_rowsum = a_rowsum;
  args.a_period = a_rows_period;
  args.col = col;
  args.out = static_cast<__half*>(out);
  args.ldo = ldo;
  args.epi = epi;
  args.res = static_cast<const __half*>(res);
  args.ldr = ldr;
  args.gate = static_cast<const __half*>(gate);
  args.rows_per_gate = rows_per_gate;
  static const uint64_t store_policy = [] {
    const char* e = getenv("VQ_STORE_POLICY");   // tuning knob: "normal" keeps outputs L2-resident for the consumer
    return (e && e[0] == 'n') ? kEvictNormal : kEvictFirst;
  }();
  args.store_policy = store_policy;
  // Rasterisation: groups of `group_m` m-panels, m fastest inside a group, then n (tile_to_mn).  A wave of CTAs then covers
  // few m-panels x several n-tiles: an A panel is fetched from HBM once and its other n-tiles hit L2 (with m fastest over ALL
  // panels, K = 4608 re-streamed the whole A operand once per n-tile column as soon as it exceeded the L2, and the first wave
  // of a K = 1152 GEMM pulled every A panel at once).  VQ_GEMM_GROUP_M=0 restores m-fastest order (A/B knob).
  static const int group_m_env = [] {
    const char* e = getenv("VQ_GEMM_GROUP_M");
    return e ? atoi(e) : 8;
  }();
  const int tile_m = pair ? 2 * BM : BM;
  const int m_tiles = (M + tile_m - 1) / tile_m;
  args.group_m = group_m_env <= 0 ? m_tiles : (group_m_env < m_tiles ? group_m_env : m_tiles);
  const int tiles = ((M + tile_m - 1) / tile_m) * ((N + BN - 1) / BN);
  const int workers = pair ? num_sms() / 2 : num_sms();
  const int grid = (tiles < workers ? tiles : workers) * (pair ? 2 : 1);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (epi) {
    case VQ_EPI_BIAS: return launch_gemm<VQ_EPI_BIAS>(ta, tb, to, tr, args, grid, pair, st);
    case VQ_EPI_GELU_TANH: return launch_gemm<VQ_EPI_GELU_TANH>(ta, tb, to, tr, args, grid, pair, st);
    case VQ_EPI_GATE_RESIDUAL: return launch_gemm<VQ_EPI_GATE_RESIDUAL>(ta, tb, to, tr, args, grid, pair, st);
#ifdef VQ_DEBUG_EPI
    case VQ_EPI_DEBUG_LOADS: return launch_gemm<VQ_EPI_DEBUG_LOADS>(ta, tb, to, tr, args, grid, pair, st);
    case VQ_EPI_DEBUG_MATH: return launch_gemm<VQ_EPI_DEBUG_MATH>(ta, tb, to, tr, args, grid, pair, st);
    case VQ_EPI_DEBUG_STORES: return launch_gemm<VQ_EPI_DEBUG_STORES>(ta, tb, to, tr, args, grid, pair, st);
    case VQ_EPI_DEBUG_MAINLOOP: return launch_gemm<VQ_EPI_DEBUG_MAINLOOP>(ta, tb, to, tr, args, grid, pair, st);
#endif
    default: return VQ_ERR_ARG;
  }
}
