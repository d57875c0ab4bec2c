// One warp-specialised tcgen05 pipeline shared by every GEMM-shaped kernel in this tree
// (FC fwd/dgrad/wgrad, conv fprop/dgrad/wgrad, the layer-0 im2col GEMM):
//
//   warp 0      TMA producer: waits "empty[s]", arms "full[s]" with the stage byte count and
//               issues the policy's cp.async.bulk.tensor loads into stage s.
//   warp 1      allocates TMEM, then one elected lane issues tcgen05.mma (M=128, N=BN, K=16) four
//               times per 64-deep k-block and tcgen05.commit's the stage's "empty" barrier; after
//               the last k-block it commits "accum_full".
//   warps 2..5  epilogue: wait "accum_full", tcgen05.ld their 32-lane TMEM quadrant 32 columns at a
//               time and hand the fp32 fragment to the policy (bias/ReLU/mask/bf16 pack/atomics).
//
// Tile: 128 (M) x BN (N) per CTA, fp32 accumulator in BN TMEM columns, BK = 64 bf16 = one
// 128-byte swizzle row.  Operands are staged in the canonical SWIZZLE_128B layouts, either
// K-major (rows = M/N index) or MN-major (rows = K index, 64-element slabs) -- see ptx.cuh.
//
// A Policy provides:
//   static constexpr int  BN, STAGES;  static constexpr bool A_MN, B_MN;
//   struct Params (POD, holds CUtensorMaps; passed __grid_constant__);  struct Ctx;
//   __device__ static Ctx  make_ctx(const Params&);            // per-CTA tile coordinates
//   __device__ static int  num_k_iters(const Params&, const Ctx&);
//   __device__ static void load(const Params&, const Ctx&, int kiter, uint8_t* sA, uint8_t* sB,
//                               uint64_t* full_bar);           // one thread
//   __device__ static void epilogue(const Params&, const Ctx&, int row, int col0,
//                                   const uint32_t (&acc)[32]); // row in [0,128), 32 columns
#pragma once
#include "ptx.cuh"

namespace b200 {

constexpr int UMMA_BM = 128;
constexpr int UMMA_BK = 64;                       // elements per k-block (128 bytes of bf16)
constexpr int UMMA_A_BYTES = UMMA_BM * UMMA_BK * 2;   // 16 KB per stage
constexpr int UMMA_SLAB_BYTES = 64 * 128;             // one MN-major slab: 64 k-rows x 128 B
constexpr int UMMA_THREADS = 192;

template <int BN>
__host__ __device__ constexpr int umma_b_bytes() { return BN * UMMA_BK * 2; }
template <int BN, int STAGES>
__host__ __device__ constexpr int umma_smem_bytes() {
  return STAGES * (UMMA_A_BYTES + umma_b_bytes<BN>()) + 1024 /*alignment slack*/ + 256 /*barriers*/;
}
template <int BN>
__host__ __device__ constexpr uint32_t umma_tmem_cols() { return BN <= 32 ? 32u : BN <= 64 ? 64u : BN <= 128 ? 128u : 256u; }

template <class P>
__global__ void __launch_bounds__(UMMA_THREADS, 1)
umma_kernel(const __grid_constant__ typename P::Params prm) {
  constexpr int BN = P::BN;
  constexpr int STAGES = P::STAGES;
  constexpr int B_BYTES = umma_b_bytes<BN>();
  constexpr uint32_t STAGE_TX = UMMA_A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = umma_tmem_cols<BN>();
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N must be a multiple of 16 in [16,256]");
  static_assert(!P::B_MN || BN % 64 == 0, "MN-major B is staged in 64-wide slabs");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * UMMA_A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    P::prefetch(prm);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
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

  const typename P::Ctx ctx = P::make_ctx(prm);
  const int nk = P::num_k_iters(prm, ctx);

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nk; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1, 1);
        mbar_arrive_expect_tx(&full_bar[s], STAGE_TX);
        P::load(prm, ctx, i, sA + s * UMMA_A_BYTES, sB + s * B_BYTES, &full_bar[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(UMMA_BM, BN, P::A_MN, P::B_MN);
      // byte step per UMMA_K=16: K-major -> 32 B inside the swizzle row; MN-major -> 16 k-rows.
      constexpr uint32_t A_KSTEP = P::A_MN ? 16 * 128 : 32;
      constexpr uint32_t B_KSTEP = P::B_MN ? 16 * 128 : 32;
      constexpr uint32_t A_LBO = P::A_MN ? UMMA_SLAB_BYTES : 16;
      constexpr uint32_t B_LBO = P::B_MN ? UMMA_SLAB_BYTES : 16;
      for (int i = 0; i < nk; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], ph, 2);
        tc_fence_after_sync();
        const uint32_t a0 = smem_u32(sA + s * UMMA_A_BYTES);
        const uint32_t b0 = smem_u32(sB + s * B_BYTES);
#pragma unroll
        for (int k = 0; k < UMMA_BK / 16; ++k) {
          const uint64_t ad = umma_smem_desc_sw128(a0 + k * A_KSTEP, A_LBO, 1024);
          const uint64_t bd = umma_smem_desc_sw128(b0 + k * B_KSTEP, B_LBO, 1024);
          umma_f16(tmem_base, ad, bd, idesc, (i | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);       // frees the smem stage when these MMAs retire
      }
      umma_commit(accum_bar);             // accumulator complete
    }
  } else {
    if (nk > 0) {
      mbar_wait(accum_bar, 0, 3);
      tc_fence_after_sync();
      const int q = warp & 3;             // TMEM lane quadrant this warp may read
      const int row = q * 32 + lane;
      uint32_t acc[32];
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        __syncwarp();
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, acc);
        tmem_ld_wait();
        P::epilogue(prm, ctx, row, c, acc);
      }
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
