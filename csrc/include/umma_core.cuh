// One persistent, warp-specialised tcgen05 pipeline shared by every GEMM-shaped kernel in this
// tree (FC fwd/dgrad/wgrad, conv fprop/dgrad/wgrad, the layer-0 im2col GEMM).
//
// One CTA per SM loops over output tiles (tile = blockIdx.x, += gridDim.x):
//   warp 0      TMA producer: runs ahead across tile boundaries; waits "empty[s]", arms "full[s]"
//               with the stage byte count, issues the policy's cp.async.bulk.tensor loads.
//   warp 1      owns TMEM (2 accumulator buffers of BN fp32 columns).  One lane waits
//               "acc_empty[buf]", then per k-block waits "full[s]" and issues four tcgen05.mma
//               (M=128, N=BN, K=16), tcgen05.commit -> "empty[s]"; after the last k-block of the tile
//               tcgen05.commit -> "acc_full[buf]".
//   warps 2..9  epilogue: wait "acc_full[buf]", tcgen05.ld their 32-lane TMEM quadrant 32 columns
//               at a time, hand the fp32 fragment to the policy (bias/ReLU/mask/bf16 pack/atomics),
//               then arrive on "acc_empty[buf]".  The epilogue of tile t overlaps the MMAs of t+1.
//
// Tile: 128 (M) x BN (N), BK = 64 bf16 = one 128-byte swizzle row.  Operands are staged in the
// canonical SWIZZLE_128B layouts, K-major (rows = M/N index) or MN-major (rows = K index,
// 64-element slabs) -- see ptx.cuh.
//
// A Policy provides:
//   static constexpr int  BN, STAGES;  static constexpr bool A_MN, B_MN;
//   struct Params (POD, holds CUtensorMaps and `int num_tiles`; passed __grid_constant__);
//   struct Ctx;
//   __device__ static void prefetch(const Params&);
//   __device__ static Ctx  make_ctx(const Params&, int tile);
//   __device__ static int  num_k_iters(const Params&, const Ctx&);
//   __device__ static void load(const Params&, const Ctx&, int kiter, uint8_t* sA, uint8_t* sB,
//                               uint64_t* full_bar);                                  // one thread
//   struct RowCtx;  __device__ static RowCtx row_ctx(const Params&, const Ctx&, int row);
//                                                   // per-thread, per-tile address arithmetic
//   __device__ static void epilogue(const Params&, const Ctx&, const RowCtx&, int row, int col0,
//                                   const uint32_t (&acc)[32], float* epi_smem);
//                                                                 // row in [0,128), 32 columns
//   static constexpr int EPI_SMEM;                                // bytes of epilogue scratch
//   __device__ static void epi_begin(const Params&, float* epi_smem, int tid);  // tid in [0,256)
//   __device__ static void epi_end(const Params&, float* epi_smem, int tid);    // after all tiles
#pragma once
#include "ptx.cuh"

namespace b200 {

constexpr int UMMA_BM = 128;
constexpr int UMMA_BK = 64;                           // elements per k-block (128 bytes of bf16)
constexpr int UMMA_A_BYTES = UMMA_BM * UMMA_BK * 2;   // 16 KB per stage
constexpr int UMMA_SLAB_BYTES = 64 * 128;             // one MN-major slab: 64 k-rows x 128 B
constexpr int UMMA_EPI_WARPS = 8;                     // two per TMEM lane quadrant
constexpr int UMMA_THREADS = 64 + 32 * UMMA_EPI_WARPS;
constexpr int UMMA_ACC_BUFS = 2;

template <int BN>
__host__ __device__ constexpr int umma_b_bytes() { return BN * UMMA_BK * 2; }
template <int BN, int STAGES>
__host__ __device__ constexpr int umma_smem_bytes() {
  return STAGES * (UMMA_A_BYTES + umma_b_bytes<BN>()) + 1024 /*alignment slack*/ + 256 /*barriers*/;
}
template <int BN>
__host__ __device__ constexpr uint32_t umma_tmem_cols() {   // power of two >= 2*BN, >= 32
  return 2 * BN <= 32 ? 32u : 2 * BN <= 64 ? 64u : 2 * BN <= 128 ? 128u : 2 * BN <= 256 ? 256u : 512u;
}

template <class P>
__global__ void __launch_bounds__(UMMA_THREADS, 1)
umma_kernel(const __grid_constant__ typename P::Params prm) {
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
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + UMMA_ACC_BUFS);
  float* epi_smem = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);

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
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
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
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
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
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
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
