// Halo-tile 3x3 convolution (fprop / dgrad) for the large-spatial layers.
//
// The generic ConvPolicy issues nine TMA loads of the same activation pixels (one per filter tap)
// per 64-channel chunk, which makes the early VGG layers (64/128 channels, 224^2 / 112^2 pixels)
// L2-bandwidth bound: 16 KB of A per 128x64x64 MMA block.  Here one TMA box brings the (8+2) x
// (16+2) pixel halo of an 8 x 16 output tile into shared memory ONCE per channel chunk (23 KB),
// and the nine taps are nine *views* of it: the K-major SWIZZLE_128B descriptor of tap (kh, kw)
// starts at row kh*10 + kw and steps 10 rows (SBO = 1280 B) between 8-pixel groups.  Measured on
// B200 (bench/probe_shift.py): tcgen05.mma swizzles on absolute shared-memory address bits, so such
// unaligned, strided views of a TMA-written tile are read exactly (descriptor base_offset = 0).
// A traffic drops 6.25x.
//
// Two rings instead of one: A halos (NA slots) and B weight tiles (NB slots, one per (tap, chunk)).
// When all 9 * chunks weight tiles fit in the B ring and there is a single channel tile, they are
// loaded once per CTA and stay resident for every tile the persistent CTA processes.
//
// Warp roles, TMEM double buffering and the epilogue (bias / ReLU / ReLU-mask / fused column sum,
// shared with ConvPolicy) are those of umma_core.cuh.
#pragma once
#include "umma_policies.cuh"

namespace b200 {

constexpr int HALO_WT = 8;                    // output tile: 8 (w) x 16 (h) pixels = 128 rows
constexpr int HALO_HT = 16;
constexpr int HALO_PITCH = HALO_WT + 2;       // halo rows per image row
constexpr int HALO_ROWS = HALO_PITCH * (HALO_HT + 2);       // 180
constexpr int HALO_BYTES = HALO_ROWS * 128;                 // 23040 (TMA transaction size)
constexpr int HALO_SLOT = 23 * 1024;                        // 1 KB aligned slot

template <int BN>
struct HaloCfg {
  static constexpr int NA = BN >= 256 ? 2 : 3;
  static constexpr int NB = BN >= 256 ? 5 : (BN >= 128 ? 9 : 18);
  static constexpr int B_SLOT = BN * 128;
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM = NA * HALO_SLOT + NB * B_SLOT + 1024 + BAR_BYTES;
};

template <int BN, bool DGRAD, bool POOL = false>
__global__ void __launch_bounds__(UMMA_THREADS, 1)
conv_halo_kernel(const __grid_constant__ ConvParams prm) {
  using Cfg = HaloCfg<BN>;
  using Epi = ConvPolicy<BN, 1, DGRAD, POOL>;
  constexpr int NA = Cfg::NA, NB = Cfg::NB, B_SLOT = Cfg::B_SLOT;
  constexpr uint32_t TMEM_COLS = umma_tmem_cols<BN>();

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + NA * HALO_SLOT;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + NB * B_SLOT);
  uint64_t* a_empty = a_full + NA;
  uint64_t* b_full = a_empty + NA;
  uint64_t* b_empty = b_full + NB;
  uint64_t* acc_full = b_empty + NB;
  uint64_t* acc_empty = acc_full + UMMA_ACC_BUFS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + UMMA_ACC_BUFS);
  float* epi_smem = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a_full) + Cfg::BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&prm.mapA);
    tma_prefetch_desc(&prm.mapB);
    for (int s = 0; s < NA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < NB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int b = 0; b < UMMA_ACC_BUFS; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], UMMA_EPI_WARPS); }
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
  const int cchunks = prm.Ca / UMMA_BK;
  const bool resident = prm.resident != 0;

  // Converged producer / MMA warps with elect.sync-predicated issue (see umma_core.cuh).
  if (warp == 0) {
    uint32_t a = 0, aph = 0, bs = 0, bph = 0;      // ring positions / phases as counters
    bool first = true;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const typename Epi::Ctx ctx = Epi::make_ctx(prm, tile);
      for (int c = 0; c < cchunks; ++c) {
        mbar_wait(&a_empty[a], aph ^ 1, 21);
        if (elect_one()) {
          mbar_arrive_expect_tx(&a_full[a], HALO_BYTES);
          tma_load_4d(sA + a * HALO_SLOT, &prm.mapA, &a_full[a], c * UMMA_BK, ctx.w0 - 1, ctx.h0 - 1, ctx.n0);
        }
        __syncwarp();
        if (++a == NA) { a = 0; aph ^= 1; }
        if (resident && !first) continue;
        for (int tap = 0; tap < 9; ++tap) {
          uint32_t b;
          if (resident) {
            b = c * 9 + tap;
          } else {
            b = bs;
            mbar_wait(&b_empty[b], bph ^ 1, 22);
            if (++bs == NB) { bs = 0; bph ^= 1; }
          }
          if (elect_one()) {
            mbar_arrive_expect_tx(&b_full[b], B_SLOT);
            uint8_t* dst = sB + b * B_SLOT;
            if constexpr (!DGRAD) {
              tma_load_2d(dst, &prm.mapB, &b_full[b], tap * prm.wcols_per_tap + c * UMMA_BK, ctx.c0);
            } else {
              const int wt = 8 - tap;
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_2d(dst + j * UMMA_SLAB_BYTES, &prm.mapB, &b_full[b],
                            wt * prm.wcols_per_tap + ctx.c0 + 64 * j, c * UMMA_BK);
            }
          }
          __syncwarp();
        }
      }
      first = false;
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(UMMA_BM, BN, false, DGRAD);
    constexpr uint32_t B_KSTEP = DGRAD ? 16 * 128 : 32;
    constexpr uint32_t B_LBO = DGRAD ? UMMA_SLAB_BYTES : 16;
    // The MMA warp has to keep up with N=64 MMAs of 32 tensor cycles each, so its loops carry no
    // address arithmetic: every descriptor is a base plus a compile-time constant (the address
    // field is addr >> 4 and shared memory is < 256 KB: the add cannot carry out of the field),
    // the nine taps are unrolled, ring positions are counters.  (The first version rebuilt both
    // descriptors per tap in a rolled loop: ~300 issue cycles per tap against 128 of tensor time;
    // the 64-channel layers ran at 25-32 % tensor-pipe active.)
    const uint64_t ad_base = umma_smem_desc_sw128(smem_u32(sA), 16, HALO_PITCH * 128);
    const uint64_t bd_base = umma_smem_desc_sw128(smem_u32(sB), B_LBO, 1024);
    uint32_t a = 0, aph = 0, bs = 0, bph = 0, t = 0;
    bool first = true;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const uint32_t buf = t & 1;
      mbar_wait(&acc_empty[buf], ((t >> 1) & 1) ^ 1, 24);
      tc_fence_after_sync();
      const uint32_t tmem_acc = tmem_base + buf * BN;
      for (int c = 0; c < cchunks; ++c) {
        mbar_wait(&a_full[a], aph, 25);
        tc_fence_after_sync();
        const uint64_t ad = ad_base + a * (HALO_SLOT >> 4);
        const uint32_t acc_c = c != 0 ? 1u : 0u;
        if (resident && !first) {
          // steady state, resident weights: nothing to wait for between taps -- 36 MMAs back to back
          const uint64_t bd = bd_base + c * (9 * B_SLOT >> 4);
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
              for (int k = 0; k < UMMA_BK / 16; ++k)
                umma_f16(tmem_acc, ad + (((tap / 3) * HALO_PITCH + tap % 3) * 128 >> 4) + k * 2,
                         bd + (tap * B_SLOT >> 4) + k * (B_KSTEP >> 4), idesc,
                         (tap | k) != 0 ? 1u : acc_c);
            }
            umma_commit(&a_empty[a]);
          }
          __syncwarp();
        } else {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            uint32_t b;
            if (resident) {          // first tile of this CTA: the weights are still arriving
              b = c * 9 + tap;
              mbar_wait(&b_full[b], 0, 26);
            } else {
              b = bs;
              mbar_wait(&b_full[b], bph, 27);
              if (++bs == NB) { bs = 0; bph ^= 1; }
            }
            tc_fence_after_sync();
            // tap view of the halo: starts kh rows of 10 pixels + kw pixels in, 10-pixel group pitch
            const uint64_t bd = bd_base + b * (B_SLOT >> 4);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < UMMA_BK / 16; ++k)
                umma_f16(tmem_acc, ad + (((tap / 3) * HALO_PITCH + tap % 3) * 128 >> 4) + k * 2,
                         bd + k * (B_KSTEP >> 4), idesc, (tap | k) != 0 ? 1u : acc_c);
              if (!resident) umma_commit(&b_empty[b]);
            }
            __syncwarp();
          }
          if (elect_one()) umma_commit(&a_empty[a]);
          __syncwarp();
        }
        if (++a == NA) { a = 0; aph ^= 1; }
      }
      first = false;
      if (elect_one()) umma_commit(&acc_full[buf]);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    if constexpr (Epi::EPI_SMEM > 0) {
      Epi::epi_begin(prm, epi_smem, threadIdx.x - 64);
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    uint32_t t = 0;
    // Fused bias gradient with a single channel tile (the 64-channel layers): every tile of this
    // CTA lands on the same columns, so a thread keeps running sums of what it stored and the
    // warp transpose-reduce runs once per kernel instead of once per tile.
    float keep[32];
    const bool keep_on = DGRAD && BN == 64 && (prm.flags & CONV_COLSUM) && prm.Cn <= BN;
#pragma unroll
    for (int i = 0; i < 32; ++i) keep[i] = 0.f;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const typename Epi::Ctx ctx = Epi::make_ctx(prm, tile);
      const typename Epi::RowCtx rc = Epi::row_ctx(prm, ctx, row);
      const uint32_t buf = t & 1;
      uint32_t acc[32];
      if constexpr (DGRAD) {
        // dgrad: fetch the ReLU masks of this warp's chunks before waiting for the accumulator
        constexpr int NCH = BN / 64;
        uint32_t mk[NCH][16];
        const bool masked = (prm.flags & CONV_MASK) != 0;
        if (masked) {
#pragma unroll
          for (int ci = 0; ci < NCH; ++ci) Epi::load_mask(prm, ctx, rc, half * 32 + ci * 64, mk[ci]);
        }
        mbar_wait(&acc_full[buf], (t >> 1) & 1, 28);
        tc_fence_after_sync();
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const int c = half * 32 + ci * 64;
          __syncwarp();
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + c, acc);
          tmem_ld_wait();
          Epi::epilogue(prm, ctx, rc, row, c, acc, epi_smem, masked ? mk[ci] : nullptr, keep, keep_on);
        }
      } else {
        mbar_wait(&acc_full[buf], (t >> 1) & 1, 28);
        tc_fence_after_sync();
#pragma unroll 1
        for (int c = half * 32; c < BN; c += 64) {
          __syncwarp();
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + c, acc);
          tmem_ld_wait();
          Epi::epilogue(prm, ctx, rc, row, c, acc, epi_smem);
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    if constexpr (DGRAD) {
      if (keep_on) Epi::colsum_flush(prm, half * 32, lane, keep, epi_smem);
    }
    if constexpr (Epi::EPI_SMEM > 0) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      Epi::epi_end(prm, epi_smem, threadIdx.x - 64);
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
