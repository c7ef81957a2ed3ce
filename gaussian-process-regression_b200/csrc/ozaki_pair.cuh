// ozaki_pair.cuh -- CTA-pair variant of the wide INT8 substitution update.
//
// STATUS: compiles for sm_100a; NOT YET RUN ON HARDWARE (written after the round's GPU budget was spent).  It is not
// part of libgprc: only tools/oz_test includes it (`tools/oz_test check|time S n mc j 3 0`, j = pair index), which
// compares it with the same host digit emulation that pins the other two kernels.
//
// One cluster of two CTAs owns 256 L-rows (block rows 2j and 2j+1, against k < 256 j) x 128 test points:
//   * tcgen05.mma.cta_group::2, M = 256, N = 128, K = 32: CTA r keeps rows 128 r .. 128 r + 127 of A in its own shared
//     memory and of D in its own tensor memory (4 accumulators x 128 columns), and HALF of the B tile (64 of the 128 test
//     points); the leader CTA (rank 0) issues every MMA.  Per SM a 256 x 128 x 32 MMA has the 64 clk math floor of the
//     wide kernel's 128 x 128 x 32, but each V digit tile is fetched from HBM once per 256 rows instead of once per 128,
//     and each CTA reads only 2 KB of B per MMA from shared memory.
//   * orders in two passes per drain interval exactly as update128_kernel (ozaki.cuh): orders [0, S-4) then [S-4, S).
//   * synchronisation: every CTA's producer fills its own ring (local `full` barriers); the peer's idle MMA warp relays
//     "stage landed" to the leader (`peer_ready`, remote mbarrier.arrive through mapa); the leader's tcgen05.commit is
//     multicast to both CTAs for `empty`, `tmem_full` and `pass_done`; both CTAs' drain warps arrive on the leader's
//     `tmem_empty` (count 8).
// The substitution then needs one extra FP64 step per pair: block row 2j+1 still has to subtract L[2j+1, 2j] V[2j]
// (128 columns) after the diagonal solve of block row 2j -- TrsmLeftUpdatePolicy with a k range.
#pragma once
#include "ozaki.cuh"

namespace gprc {
namespace oz {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 remAddr32;\n"
      "mapa.shared::cluster.u32 remAddr32, %0, %1;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [remAddr32];\n"
      "}" ::"r"(local_bar),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void mma_i8_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the barrier at this offset in BOTH CTAs of the pair arrives once the MMAs issued so far have completed
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
  asm volatile(
      "{\n"
      ".reg .b16 mask;\n"
      "mov.b16 mask, 3;\n"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], mask;\n"
      "}" ::"r"(bar)
      : "memory");
}

template <int S>
struct CfgPair {
  static_assert(S >= 5 && S <= 8, "5..8 digits");
  static constexpr int NLO = S - 4;
  static constexpr int HALF_B = B2_TILE / 2;                  // 64 test points x 32 k = 2048 B per digit and CTA
  static constexpr int STAGE0 = NLO * (A_TILE + HALF_B);
  static constexpr int STAGE1 = S * (A_TILE + HALF_B);
  static constexpr int RING = (215 * 1024) / STAGE1 * STAGE1;
  static constexpr int STAGES1 = RING / STAGE1;
  static constexpr int STAGES0 = (RING / STAGE0) > 8 ? 8 : (RING / STAGE0);
  static constexpr int SMEM_BYTES = RING + DRAIN_STAGING_BYTES + 1024;
  static constexpr int KC = (S <= 7) ? 16384 : 8192;
};

// p.i = PAIR index j (block rows 2j and 2j + 1, k < 256 j); grid = 2 * (mpad / 128) CTAs, clusters of 2
template <int S>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) update_pair_kernel(const UpdateParams p) {
  using C = CfgPair<S>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stg_all = reinterpret_cast<double*>(smem_raw + C::RING);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + C::RING + DRAIN_STAGING_BYTES);  // [16] local: this CTA's stage landed
  uint64_t* peer_ready = full + 16;                                                        // [16] leader: the peer's stage landed
  uint64_t* empty = peer_ready + 16;                                                       // [16] local: stage may be refilled
  uint64_t* tmem_full = empty + 16;
  uint64_t* tmem_empty = tmem_full + 1;  // leader only: 8 drain warps
  uint64_t* pass_done = tmem_empty + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pass_done + 1);
  static_assert((16 * 3 + 3) * 8 + 8 <= 1024, "barrier block");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int tc = blockIdx.x >> 1;               // 128-point tile
  const int brow = 2 * p.i + (int)rank;         // this CTA's block row
  const int KT = 8 * p.i;
  constexpr int KT_CHUNK = C::KC / BK;
  const int nchunks = (KT + KT_CHUNK - 1) / KT_CHUNK;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < 16; ++s) {
      mbar_init(smem_u32(full + s), 1);
      mbar_init(smem_u32(peer_ready + s), 1);
      mbar_init(smem_u32(empty + s), 1);
    }
    mbar_init(smem_u32(tmem_full), 1);
    mbar_init(smem_u32(tmem_empty), 8);
    mbar_init(smem_u32(pass_done), 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_pair(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // barrier inits of both CTAs are visible before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== producer: own A rows (block row `brow`) and own half of the B tile =====
    const int8_t* a_src = p.Ls + (long)brow * p.KB * (long)(S * A_TILE);
    const int8_t* b_src = p.Vs + (long)tc * p.KB * (long)(S * B2_TILE) + (long)rank * C::HALF_B;
    int cnt[2] = {0, 0};
    for (int r = 0; r < 2 * nchunks; ++r) {
      const int pass = r & 1, c = r >> 1;
      const int kt0 = c * KT_CHUNK, kt1 = min(KT, kt0 + KT_CHUNK);
      const int nst = pass ? C::STAGES1 : C::STAGES0, sbytes = pass ? C::STAGE1 : C::STAGE0;
      const int planes = pass ? S : C::NLO;
      if (r > 0) mbar_wait_guarded(smem_u32(pass_done), (r - 1) & 1, p.error, 8);
      for (int kt = kt0; kt < kt1; ++kt) {
        const int n = cnt[pass]++, s = n % nst;
        if (kt - kt0 >= nst) mbar_wait_guarded(smem_u32(empty + pass * 8 + s), ((n / nst) - 1) & 1, p.error, 16);
        if (elect_one()) {
          const uint32_t bar = smem_u32(full + pass * 8 + s);
          const uint32_t dst = smem_u32(smem_raw + s * sbytes);
          mbar_arrive_expect_tx(bar, (uint32_t)sbytes);
          bulk_g2s(dst, a_src + (long)kt * (S * A_TILE), planes * A_TILE, bar);
          for (int q = 0; q < planes; ++q)  // this CTA's 64 points of every digit plane of the 128-point tile
            bulk_g2s(dst + planes * A_TILE + q * C::HALF_B, b_src + (long)kt * (S * B2_TILE) + (long)q * B2_TILE,
                     C::HALF_B, bar);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int cnt[2] = {0, 0};
      if (!leader) {
        // ===== peer: relay "my stage has landed" to the leader =====
        for (int r = 0; r < 2 * nchunks; ++r) {
          const int pass = r & 1, c = r >> 1;
          const int kt0 = c * KT_CHUNK, kt1 = min(KT, kt0 + KT_CHUNK);
          const int nst = pass ? C::STAGES1 : C::STAGES0;
          for (int kt = kt0; kt < kt1; ++kt) {
            const int n = cnt[pass]++, s = n % nst;
            mbar_wait_guarded(smem_u32(full + pass * 8 + s), (n / nst) & 1, p.error, 64);
            mbar_arrive_remote(smem_u32(peer_ready + pass * 8 + s), 0);
          }
        }
      } else {
        // ===== leader: MMA issuer for the pair =====
        constexpr uint32_t idesc = instr_desc_i8(2 * BM, BN2);
        for (int r = 0; r < 2 * nchunks; ++r) {
          const int pass = r & 1, c = r >> 1;
          const int kt0 = c * KT_CHUNK, kt1 = min(KT, kt0 + KT_CHUNK);
          const int nst = pass ? C::STAGES1 : C::STAGES0, sbytes = pass ? C::STAGE1 : C::STAGE0;
          if (r > 0) {
            mbar_wait_guarded(smem_u32(tmem_empty), (r - 1) & 1, p.error, 32);
            tc_fence_after();
          }
          for (int kt = kt0; kt < kt1; ++kt) {
            const int n = cnt[pass]++, s = n % nst;
            mbar_wait_guarded(smem_u32(full + pass * 8 + s), (n / nst) & 1, p.error, 64);
            mbar_wait_guarded(smem_u32(peer_ready + pass * 8 + s), (n / nst) & 1, p.error, 256);
            tc_fence_after();
            const uint32_t a0 = smem_u32(smem_raw + s * sbytes);
            const bool first = (kt == kt0);
            if (pass == 0) {
              const uint64_t ad0 = smem_desc(a0, 128, 256), bd0 = smem_desc(a0 + C::NLO * A_TILE, 128, 256);
#pragma unroll
              for (int a = 0; a < C::NLO; ++a)
#pragma unroll
                for (int b = 0; b < C::NLO - a; ++b)
                  mma_i8_pair(tmem_base + (uint32_t)((a + b) * BN2), ad0 + (uint64_t)(a * (A_TILE >> 4)),
                              bd0 + (uint64_t)(b * (C::HALF_B >> 4)), idesc, (!first || a > 0) ? 1u : 0u);
            } else {
              const uint64_t ad0 = smem_desc(a0, 128, 256), bd0 = smem_desc(a0 + S * A_TILE, 128, 256);
#pragma unroll
              for (int a = 0; a < S; ++a)
#pragma unroll
                for (int b = 0; b < S - a; ++b)
                  if (a + b >= C::NLO)
                    mma_i8_pair(tmem_base + (uint32_t)((a + b - C::NLO) * BN2), ad0 + (uint64_t)(a * (A_TILE >> 4)),
                                bd0 + (uint64_t)(b * (C::HALF_B >> 4)), idesc, (!first || a > 0) ? 1u : 0u);
            }
            tc_commit_pair(smem_u32(empty + pass * 8 + s));
          }
          tc_commit_pair(smem_u32(tmem_full));
          tc_commit_pair(smem_u32(pass_done));
        }
      }
    }
    __syncwarp();
  } else {
    // ===== drain: this CTA's 128 rows x 128 test points =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const double sr = p.scale_row[(long)brow * BM + row];
    const double* sc = p.scale_col + (long)tc * BN2;
    double* stg = stg_all + q * (32 * 9);
    double* cbase = p.T + (long)tc * BN2 + ((long)brow * BM + q * 32 + (lane >> 3)) * p.ldt + (lane & 7);
    {
      const double* prow = p.T + (long)tc * BN2 + ((long)brow * BM + row) * p.ldt;
#pragma unroll
      for (int j = 0; j < 8; ++j) prefetch_l2(prow + j * 16);
    }
    constexpr int NHI = 4;
    for (int r = 0; r < 2 * nchunks; ++r) {
      const int pass = r & 1;
      mbar_wait_guarded(smem_u32(tmem_full), r & 1, p.error, 128);
      tc_fence_after();
      const double srp = pass ? ldexp(sr, -8 * C::NLO) : sr;
#pragma unroll 1
      for (int g = 0; g < BN2 / 8; ++g) {
        int32_t acc[NHI][8];
        if (pass == 0) {
#pragma unroll
          for (int o = 0; o < C::NLO; ++o)
            tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(o * BN2 + g * 8), acc[o]);
        } else {
#pragma unroll
          for (int o = 0; o < NHI; ++o)
            tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(o * BN2 + g * 8), acc[o]);
        }
        double cur[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) cur[it] = cbase[(long)(it * 4) * p.ldt + g * 8];
        double scv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) scv[j] = __ldg(sc + g * 8 + j);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          double h;
          if (pass == 0) {
            h = (double)acc[C::NLO - 1][j];
#pragma unroll
            for (int o = C::NLO - 2; o >= 0; --o) h = fma(h, 0.00390625, (double)acc[o][j]);
          } else {
            h = (double)acc[NHI - 1][j];
#pragma unroll
            for (int o = NHI - 2; o >= 0; --o) h = fma(h, 0.00390625, (double)acc[o][j]);
          }
          stg[lane * 9 + j] = (srp * scv[j]) * h;
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) cur[it] -= stg[(it * 4 + (lane >> 3)) * 9 + (lane & 7)];
#pragma unroll
        for (int it = 0; it < 8; ++it) cbase[(long)(it * 4) * p.ldt + g * 8] = cur[it];
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(smem_u32(tmem_empty));
        else mbar_arrive_remote(smem_u32(tmem_empty), 0);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves (or frees tensor memory) while the other may still read its shared memory
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

}  // namespace oz
}  // namespace gprc
