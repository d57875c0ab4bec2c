// EXPERIMENTAL (B200_HALO_TMA_EPI=1; written after the round-1 GPU budget was spent, not yet run on
// hardware): the halo-tile convolution of conv_halo.cuh with the epilogue's global traffic moved to
// TMA.
//
// Why (profiles/r1_ncu_halo_v6.md): on the 64/128-channel layers the L1 data path is the saturated
// unit -- tcgen05.mma operand reads (1 728 wavefronts per 128x64 tile) plus the epilogue's
// row-per-thread global accesses (a warp-wide 32-byte access touches 32 lines: 512 store wavefronts
// per tile, and as many again for the dgrad ReLU mask).  Here the epilogue writes its bf16 rows
// into a SWIZZLE_128B staging tile in shared memory (conflict-free 16-byte stores: 128 wavefronts)
// and one elected thread issues a single cp.async.bulk.tensor store per 64-channel slab; the dgrad
// mask tile arrives the same way (TMA load, issued one tile ahead by the same thread), so the
// epilogue performs no global loads or stores of its own.
//
// Differences from conv_halo_kernel: ring sizes are run-time (the launcher fits A ring + B ring +
// staging into 227 KB per layer), the tile must be full (H % 16 == 0, W % 8 == 0: guaranteed by
// halo_applicable), the optional fused bias gradient works as before.
#pragma once
#include "conv_halo.cuh"

namespace b200 {

struct ConvTmaExtra {
  CUtensorMap mapOut;     // output, dims {Cn, W, H, N}, box {64, 8, 16, 1}, SWIZZLE_128B (store)
  CUtensorMap mapMask;    // dgrad ReLU mask source, same geometry (load)
  int na, nb;             // ring sizes chosen by the launcher (na <= 4, nb <= 18)
};

constexpr int HALO_STAGE_BYTES = UMMA_BM * 128;         // one 64-channel slab of a 128-pixel tile
constexpr int HALO_TMA_MAX_NA = 4;
constexpr int HALO_TMA_MAX_NB = 18;

template <int BN, bool DGRAD>
__host__ __device__ constexpr int halo_tma_smem(int na, int nb, bool masked) {
  return na * HALO_SLOT + nb * BN * 128 + (BN / 64) * HALO_STAGE_BYTES * (masked ? 2 : 1) + 1024 + 512 +
         ConvPolicy<BN, 1, DGRAD>::EPI_SMEM;
}

// byte offset of 16-byte chunk j of row r inside a SWIZZLE_128B slab
__device__ __forceinline__ uint32_t sw128_off(int row, int j) { return row * 128 + ((j ^ (row & 7)) << 4); }

template <int BN, bool DGRAD>
__global__ void __launch_bounds__(UMMA_THREADS, 1)
conv_halo_tma_kernel(const __grid_constant__ ConvParams prm, const __grid_constant__ ConvTmaExtra ext) {
  using Epi = ConvPolicy<BN, 1, DGRAD>;
  constexpr int B_SLOT = BN * 128;
  constexpr int SLABS = BN / 64;
  constexpr int NCH = BN / 64;                   // 32-column chunks per epilogue warp
  constexpr uint32_t TMEM_COLS = umma_tmem_cols<BN>();
  const int NA = ext.na, NB = ext.nb;
  const bool masked = DGRAD && (prm.flags & CONV_MASK) != 0;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + NA * HALO_SLOT;
  uint8_t* sOut = sB + NB * B_SLOT;                                  // 1 KB aligned: all sizes are
  uint8_t* sMask = sOut + SLABS * HALO_STAGE_BYTES;
  uint8_t* after = sMask + (masked ? SLABS * HALO_STAGE_BYTES : 0);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(after);
  uint64_t* a_empty = a_full + HALO_TMA_MAX_NA;
  uint64_t* b_full = a_empty + HALO_TMA_MAX_NA;
  uint64_t* b_empty = b_full + HALO_TMA_MAX_NB;
  uint64_t* acc_full = b_empty + HALO_TMA_MAX_NB;
  uint64_t* acc_empty = acc_full + UMMA_ACC_BUFS;
  uint64_t* mask_full = acc_empty + UMMA_ACC_BUFS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mask_full + 1);
  float* epi_smem = reinterpret_cast<float*>(after + 512);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&prm.mapA);
    tma_prefetch_desc(&prm.mapB);
    tma_prefetch_desc(&ext.mapOut);
    if (masked) tma_prefetch_desc(&ext.mapMask);
    for (int s = 0; s < NA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < NB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int b = 0; b < UMMA_ACC_BUFS; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], UMMA_EPI_WARPS); }
    mbar_init(mask_full, 1);
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

  if (warp == 0) {
    // ---- TMA producer: identical protocol to conv_halo_kernel, ring sizes at run time ----------
    uint32_t a = 0, aph = 0, bs = 0, bph = 0;
    bool first = true;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const typename Epi::Ctx ctx = Epi::make_ctx(prm, tile);
      for (int c = 0; c < cchunks; ++c) {
        mbar_wait(&a_empty[a], aph ^ 1, 61);
        if (elect_one()) {
          mbar_arrive_expect_tx(&a_full[a], HALO_BYTES);
          tma_load_4d(sA + a * HALO_SLOT, &prm.mapA, &a_full[a], c * UMMA_BK, ctx.w0 - 1, ctx.h0 - 1, ctx.n0);
        }
        __syncwarp();
        if (++a == static_cast<uint32_t>(NA)) { a = 0; aph ^= 1; }
        if (resident && !first) continue;
        for (int tap = 0; tap < 9; ++tap) {
          uint32_t b;
          if (resident) {
            b = c * 9 + tap;
          } else {
            b = bs;
            mbar_wait(&b_empty[b], bph ^ 1, 62);
            if (++bs == static_cast<uint32_t>(NB)) { bs = 0; bph ^= 1; }
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
    // ---- MMA issuer: as conv_halo_kernel --------------------------------------------------------
    constexpr uint32_t idesc = umma_idesc_bf16(UMMA_BM, BN, false, DGRAD);
    constexpr uint32_t B_KSTEP = DGRAD ? 16 * 128 : 32;
    constexpr uint32_t B_LBO = DGRAD ? UMMA_SLAB_BYTES : 16;
    const uint64_t ad_base = umma_smem_desc_sw128(smem_u32(sA), 16, HALO_PITCH * 128);
    const uint64_t bd_base = umma_smem_desc_sw128(smem_u32(sB), B_LBO, 1024);
    uint32_t a = 0, aph = 0, bs = 0, bph = 0, t = 0;
    bool first = true;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const uint32_t buf = t & 1;
      mbar_wait(&acc_empty[buf], ((t >> 1) & 1) ^ 1, 64);
      tc_fence_after_sync();
      const uint32_t tmem_acc = tmem_base + buf * BN;
      for (int c = 0; c < cchunks; ++c) {
        mbar_wait(&a_full[a], aph, 65);
        tc_fence_after_sync();
        const uint64_t ad = ad_base + a * (HALO_SLOT >> 4);
        const uint32_t acc_c = c != 0 ? 1u : 0u;
        if (resident && !first) {
          const uint64_t bd = bd_base + c * (9 * B_SLOT >> 4);
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
              for (int k = 0; k < UMMA_BK / 16; ++k)
                umma_f16(tmem_acc, ad + (((tap / 3) * HALO_PITCH + tap % 3) * 128 >> 4) + k * 2,
                         bd + (tap * B_SLOT >> 4) + k * (B_KSTEP >> 4), idesc, (tap | k) != 0 ? 1u : acc_c);
            }
            umma_commit(&a_empty[a]);
          }
          __syncwarp();
        } else {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            uint32_t b;
            if (resident) {
              b = c * 9 + tap;
              mbar_wait(&b_full[b], 0, 66);
            } else {
              b = bs;
              mbar_wait(&b_full[b], bph, 67);
              if (++bs == static_cast<uint32_t>(NB)) { bs = 0; bph ^= 1; }
            }
            tc_fence_after_sync();
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
        if (++a == static_cast<uint32_t>(NA)) { a = 0; aph ^= 1; }
      }
      first = false;
      if (elect_one()) umma_commit(&acc_full[buf]);
      __syncwarp();
    }
  } else {
    // ---- epilogue: TMEM -> registers -> swizzled staging tile -> one TMA store per slab ---------
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const bool issuer = threadIdx.x == 64;            // warp 2, lane 0
    const bool colsum = DGRAD && (prm.flags & CONV_COLSUM);
    if constexpr (Epi::EPI_SMEM > 0) {
      Epi::epi_begin(prm, epi_smem, threadIdx.x - 64);
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    float keep[32];
    const bool keep_on = colsum && BN == 64 && prm.Cn <= BN;
#pragma unroll
    for (int i = 0; i < 32; ++i) keep[i] = 0.f;
    if (masked && issuer && static_cast<int>(blockIdx.x) < num_tiles) {      // mask tile of the first tile
      const typename Epi::Ctx c0 = Epi::make_ctx(prm, blockIdx.x);
      mbar_arrive_expect_tx(mask_full, SLABS * HALO_STAGE_BYTES);
#pragma unroll
      for (int s = 0; s < SLABS; ++s)
        tma_load_4d(sMask + s * HALO_STAGE_BYTES, &ext.mapMask, mask_full, c0.c0 + 64 * s, c0.w0, c0.h0, c0.n0);
    }
    uint32_t t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const typename Epi::Ctx ctx = Epi::make_ctx(prm, tile);
      const uint32_t buf = t & 1;
      mbar_wait(&acc_full[buf], (t >> 1) & 1, 68);
      tc_fence_after_sync();
      if (masked) mbar_wait(mask_full, t & 1, 69);
      if (issuer) tma_store_wait_read<0>();            // the previous tile's store has left the staging tile
      asm volatile("bar.sync 1, 256;" ::: "memory");
      uint32_t acc[32];
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) {
        const int c = half * 32 + ci * 64;               // column inside the BN-wide tile
        const int ch = ctx.c0 + c;
        __syncwarp();
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + c, acc);
        tmem_ld_wait();
        const bool chan_ok = ch < prm.Cn;                // warp-uniform
        uint8_t* slab = sOut + (c >> 6) * HALO_STAGE_BYTES;
        const uint8_t* mslab = sMask + (c >> 6) * HALO_STAGE_BYTES;
        const int j0 = (c & 63) >> 3;                    // first 16-byte chunk of this thread's 64 bytes
        float cs[32];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = __uint_as_float(acc[jj * 8 + u]);
          if constexpr (!DGRAD) {
            if (chan_ok && (prm.flags & CONV_BIAS)) {
              const float4 b0 = *reinterpret_cast<const float4*>(epi_smem + ch + jj * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(epi_smem + ch + jj * 8 + 4);
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
          }
          if (prm.flags & CONV_RELU) {
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = fmaxf(v[u], 0.f);
          }
          if (masked) {
            const uint4 m = *reinterpret_cast<const uint4*>(mslab + sw128_off(row, j0 + jj));
            const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float2 f = unpack_bf16x2(mw[u]);
              if (!(f.x > 0.f)) v[2 * u] = 0.f;
              if (!(f.y > 0.f)) v[2 * u + 1] = 0.f;
            }
          }
          const uint4 pk = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                      pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
          if (chan_ok) *reinterpret_cast<uint4*>(slab + sw128_off(row, j0 + jj)) = pk;
          if (colsum) {
            const uint32_t pw[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float2 f = unpack_bf16x2(pw[u]);
              cs[jj * 8 + 2 * u] = chan_ok ? f.x : 0.f;
              cs[jj * 8 + 2 * u + 1] = chan_ok ? f.y : 0.f;
            }
          }
        }
        if (colsum) {
          if (keep_on) {
#pragma unroll
            for (int i = 0; i < 32; ++i) keep[i] += cs[i];
          } else {
            Epi::colsum_flush(prm, ch, lane, cs, epi_smem);
          }
        }
      }
      fence_proxy_async_smem();                          // staging writes -> visible to the TMA store
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);       // TMEM buffer drained
      asm volatile("bar.sync 1, 256;" ::: "memory");    // staging complete, mask tile consumed
      if (issuer) {
#pragma unroll
        for (int s = 0; s < SLABS; ++s)
          if (ctx.c0 + 64 * s < prm.Cn)
            tma_store_4d(&ext.mapOut, sOut + s * HALO_STAGE_BYTES, ctx.c0 + 64 * s, ctx.w0, ctx.h0, ctx.n0);
        tma_store_commit();
        const int next = tile + gridDim.x;
        if (masked && next < num_tiles) {                // every warp is past its mask reads (barrier above)
          const typename Epi::Ctx cn = Epi::make_ctx(prm, next);
          mbar_arrive_expect_tx(mask_full, SLABS * HALO_STAGE_BYTES);
#pragma unroll
          for (int s = 0; s < SLABS; ++s)
            tma_load_4d(sMask + s * HALO_STAGE_BYTES, &ext.mapMask, mask_full, cn.c0 + 64 * s, cn.w0, cn.h0, cn.n0);
        }
      }
    }
    if (issuer) tma_store_wait<0>();                     // all stores complete before the CTA exits
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
