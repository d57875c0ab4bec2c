// EXPERIMENTAL (B200_DYNAMIC_TILES=1; not yet run on hardware): umma_kernel with a dynamic tile
// scheduler instead of static striding.
//
// Why: umma_kernel gives every CTA num_tiles / gridDim tiles.  During backward, the comm stream's
// all-reduce and optimizer CTAs hold register space on some SMs (an all-reduce CTA 32 K, a dgrad CTA
// 51 K of the SM's 64 K registers), so the persistent CTA for such an SM starts late and still has
// its full share to do: the kernel ends when the unluckiest CTA ends.  Here the producer warp
// claims tiles from a global counter (one atomicAdd per tile) and publishes them to the MMA and
// epilogue warps through a 4-deep shared-memory queue with full/empty mbarriers; a CTA that starts
// late simply claims fewer tiles.  Tiles are still handed out in increasing order, so neighbouring
// tiles (which share operand tiles in L2) still run at the same time on different SMs.
//
// The counter is never reset: the host knows how many claims every launch makes (num_tiles + one
// over-claim per CTA) and passes the counter's value at launch as `base` (TileTicket).
#pragma once
#include "umma_core.cuh"

namespace b200 {

struct TileTicket {
  unsigned int* counter;
  unsigned int base;
};

constexpr int UMMA_TQ = 4;
constexpr int UMMA_DYN_BAR_BYTES = 384;       // the static kernel's 256 + queue barriers and entries

template <class P>
__global__ void __launch_bounds__(UMMA_THREADS, 1)
umma_kernel_dyn(const __grid_constant__ typename P::Params prm, const TileTicket ticket) {
  constexpr int BN = P::BN;
  constexpr int STAGES = P::STAGES;
  constexpr int B_BYTES = umma_b_bytes<BN>();
  constexpr uint32_t STAGE_TX = UMMA_A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = umma_tmem_cols<BN>();
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "tile N must be a multiple of 32 in [32,256]");
  static_assert(!P::B_MN || BN % 64 == 0, "MN-major B is staged in 64-wide slabs");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * UMMA_A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;
  uint64_t* acc_empty = acc_full + UMMA_ACC_BUFS;
  uint64_t* tq_full = acc_empty + UMMA_ACC_BUFS;        // tile queue: scheduler (warp 0) -> MMA + epilogue warps
  uint64_t* tq_empty = tq_full + UMMA_TQ;
  int* tile_q = reinterpret_cast<int*>(tq_empty + UMMA_TQ);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tile_q + UMMA_TQ);
  float* epi_smem = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + UMMA_DYN_BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    P::prefetch(prm);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < UMMA_ACC_BUFS; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], UMMA_EPI_WARPS);   // one arrival per epilogue warp
    }
    for (int s = 0; s < UMMA_TQ; ++s) {
      mbar_init(&tq_full[s], 1);
      mbar_init(&tq_empty[s], 1 + UMMA_EPI_WARPS);  // MMA warp + every epilogue warp has read the entry
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int num_tiles = prm.num_tiles;

  // Producer and MMA warps run CONVERGED (all 32 lanes execute the loops and poll the barriers);
  // only the TMA / tcgen05 instructions themselves are predicated on elect.sync.  Addresses and
  // descriptors are then warp-uniform values the compiler keeps in uniform registers, instead of
  // per-lane values that need a waterfall loop around every UTCHMMA / UTMALDG.
  if (warp == 0) {
    uint32_t s = 0, ph = 0;                 // ring position / phase kept as counters (no div/mod)
    uint32_t qs = 0, qph = 0;
    for (;;) {
      // claim the next tile of the launch and publish it to the other warps of this CTA
      mbar_wait(&tq_empty[qs], qph ^ 1, 5);
      unsigned int claimed = 0;
      if (lane == 0) claimed = atomicAdd(ticket.counter, 1u) - ticket.base;
      claimed = __shfl_sync(0xffffffffu, claimed, 0);
      const int tile = claimed < static_cast<unsigned int>(num_tiles) ? static_cast<int>(claimed) : num_tiles;
      if (lane == 0) {
        tile_q[qs] = tile;
        mbar_arrive(&tq_full[qs]);          // release: the entry is visible to whoever acquires the phase
      }
      __syncwarp();
      if (++qs == UMMA_TQ) { qs = 0; qph ^= 1; }
      if (tile >= num_tiles) break;         // sentinel published: every role leaves its loop
      const typename P::Ctx ctx = P::make_ctx(prm, tile);
      const int nk = P::num_k_iters(prm, ctx);
      for (int i = 0; i < nk; ++i) {
        mbar_wait(&empty_bar[s], ph ^ 1, 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[s], STAGE_TX);
          P::load(prm, ctx, i, sA + s * UMMA_A_BYTES, sB + s * B_BYTES, &full_bar[s]);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(UMMA_BM, BN, P::A_MN, P::B_MN);
    // byte step per UMMA_K=16: K-major -> 32 B inside the swizzle row; MN-major -> 16 k-rows.
    constexpr uint32_t A_KSTEP = P::A_MN ? 16 * 128 : 32;
    constexpr uint32_t B_KSTEP = P::B_MN ? 16 * 128 : 32;
    constexpr uint32_t A_LBO = P::A_MN ? UMMA_SLAB_BYTES : 16;
    constexpr uint32_t B_LBO = P::B_MN ? UMMA_SLAB_BYTES : 16;
    uint32_t s = 0, ph = 0, t = 0;
    // Stage descriptors are base + s * stage size: the address field is addr >> 4 and shared
    // memory is < 256 KB, so the add never carries out of the 14-bit field.
    const uint64_t ad_base = umma_smem_desc_sw128(smem_u32(sA), A_LBO, 1024);
    const uint64_t bd_base = umma_smem_desc_sw128(smem_u32(sB), B_LBO, 1024);
    uint32_t qs = 0, qph = 0;
    for (;; ++t) {
      mbar_wait(&tq_full[qs], qph, 6);
      const int tile = tile_q[qs];
      __syncwarp();
      if (lane == 0) mbar_arrive(&tq_empty[qs]);
      if (++qs == UMMA_TQ) { qs = 0; qph ^= 1; }
      if (tile >= num_tiles) break;
      const typename P::Ctx ctx = P::make_ctx(prm, tile);
      const int nk = P::num_k_iters(prm, ctx);
      const uint32_t buf = t & 1;
      mbar_wait(&acc_empty[buf], ((t >> 1) & 1) ^ 1, 4);     // epilogue drained this buffer
      tc_fence_after_sync();
      const uint32_t tmem_acc = tmem_base + buf * BN;
      for (int i = 0; i < nk; ++i) {
        mbar_wait(&full_bar[s], ph, 2);
        tc_fence_after_sync();
        const uint64_t ad0 = ad_base + s * (UMMA_A_BYTES >> 4);
        const uint64_t bd0 = bd_base + s * (B_BYTES >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < UMMA_BK / 16; ++k)     // start-address field is in 16-byte units
            umma_f16(tmem_acc, ad0 + k * (A_KSTEP >> 4), bd0 + k * (B_KSTEP >> 4), idesc, (i | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);       // frees the smem stage when these MMAs retire
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(&acc_full[buf]);        // accumulator of this tile complete
      __syncwarp();
    }
  } else {
    // Eight epilogue warps: warp w reads TMEM lane quadrant (w & 3); the two warps of a quadrant
    // split the tile's 32-column chunks between them (even / odd), so every SM sub-partition has
    // two epilogue warps to overlap TMEM loads, global loads and stores.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    if constexpr (P::EPI_SMEM > 0) {
      P::epi_begin(prm, epi_smem, threadIdx.x - 64);
      asm volatile("bar.sync 1, 256;" ::: "memory");      // the epilogue warps only
    }
    uint32_t t = 0, qs = 0, qph = 0;
    for (;; ++t) {
      mbar_wait(&tq_full[qs], qph, 7);
      const int tile = tile_q[qs];
      __syncwarp();
      if (lane == 0) mbar_arrive(&tq_empty[qs]);
      if (++qs == UMMA_TQ) { qs = 0; qph ^= 1; }
      if (tile >= num_tiles) break;
      const typename P::Ctx ctx = P::make_ctx(prm, tile);
      const int nk = P::num_k_iters(prm, ctx);
      const typename P::RowCtx rc = P::row_ctx(prm, ctx, row);
      const uint32_t buf = t & 1;
      mbar_wait(&acc_full[buf], (t >> 1) & 1, 3);
      tc_fence_after_sync();
      if (nk > 0) {
        uint32_t acc[32];
#pragma unroll 1
        for (int c = half * 32; c < BN; c += 64) {
          __syncwarp();
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + c, acc);
          tmem_ld_wait();
          P::epilogue(prm, ctx, rc, row, c, acc, epi_smem);
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    if constexpr (P::EPI_SMEM > 0) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      P::epi_end(prm, epi_smem, threadIdx.x - 64);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace b200
