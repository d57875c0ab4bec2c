// Weight gradient of a 3x3 convolution with 64 input channels (VGG features.2 / features.5),
// as nine views of ONE activation halo per pixel tile.
//
//   dW[co][tap][ci] = sum over pixels p of dZ[p][co] * X[p + tap][ci]
//
// The generic WgradPolicy gives each filter tap its own CTA and TMA-loads the shifted X tile nine
// times; with Cin = Cout = 64 that is 24 KB of operand traffic per 128 cycles of tensor work and
// the kernel sits at 22 % tensor-pipe utilisation, L2/DRAM bound (profiles/r1_ncu_conv_summary_v2).
// Here, per 8x8-pixel tile (one k-block, K = 64 pixels):
//   * X comes in once as a (8+2)x(8+2) halo box (100 rows x 128 B: one row = the 64 channels of a
//     pixel), dZ as the dense 8x8 box of one 64-channel slab;
//   * tensor-core operands are MN-major views (rows = K = pixels): tap (kh,kw) of X starts at halo
//     row kh*10+kw with 8-pixel groups 10 rows apart (SBO = 1280 B) -- exact on B200 with
//     base_offset 0, see bench/probe_shift.py -- and TWO taps are stacked along M through the
//     descriptor's leading-byte-offset (LBO = byte distance between the two tap views), so the
//     MMA is a full 128 (2 taps x 64 ci) x 64 (co) x 16;
//   * five such tap-pair accumulators (320 TMEM columns) cover the nine taps; they stay in TMEM
//     for the CTA's whole pixel range (split-K over pixel tiles, one work item per CTA);
//   * epilogue: lane = ci, column = co, so a warp's red.global.add hits 32 consecutive floats of
//     dW[co][tap][:] -- one 128-byte atomic transaction.
// Operand traffic per k-block drops from 9 x 24 KB to 21 KB.
#pragma once
#include "umma_core.cuh"

namespace b200 {

constexpr int WH_T = 8;                                   // 8 x 8 pixel tile = 64 K rows
constexpr int WH_PITCH = WH_T + 2;
constexpr int WH_X_BYTES = WH_PITCH * WH_PITCH * 128;     // 12800 (TMA transaction)
constexpr int WH_X_SLOT = 13 * 1024;                      // slot (holds the +1 row the dummy tap touches)
constexpr int WH_Z_BYTES = WH_T * WH_T * 128;             // 8192
constexpr int WH_STAGE = WH_X_SLOT + WH_Z_BYTES;
constexpr int WH_STAGES = 8;
constexpr int WH_GROUPS = 5;                              // tap pairs (0,1) (2,3) (4,5) (6,7) (8,-)
constexpr int WH_SMEM = WH_STAGES * WH_STAGE + 1024 + 512;

struct WgradHaloParams {
  CUtensorMap mapX;        // X  {64, W, H, N}, box {64, 10, 10, 1}
  CUtensorMap mapZ;        // dZ {Cout, W, H, N}, box {64, 8, 8, 1}
  int num_items;           // tiles_n * ksplit
  int tiles_n;             // Cout / 64
  int tiles_w, tiles_h;    // W / 8, H / 8
  int total_tiles;         // N * tiles_h * tiles_w
  int tiles_per_split;
  int Cout;
  float* dW;               // [Cout][9][64] fp32
  float scale;
};

__device__ __forceinline__ int wh_tap_row(int tap) { return (tap / 3) * WH_PITCH + tap % 3; }

__global__ void __launch_bounds__(UMMA_THREADS, 1)
wgrad_halo64_kernel(const __grid_constant__ WgradHaloParams prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WH_STAGES * WH_STAGE);
  uint64_t* empty_bar = full_bar + WH_STAGES;
  uint64_t* acc_full = empty_bar + WH_STAGES;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&prm.mapX);
    tma_prefetch_desc(&prm.mapZ);
    for (int s = 0; s < WH_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, UMMA_EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  auto item_range = [&](int item, int& n_tile, int& t0, int& nk) {
    n_tile = item % prm.tiles_n;
    const int split = item / prm.tiles_n;
    t0 = split * prm.tiles_per_split;
    const int rem = prm.total_tiles - t0;
    nk = rem < 0 ? 0 : (rem < prm.tiles_per_split ? rem : prm.tiles_per_split);
  };

  if (warp == 0) {
    uint32_t it = 0;
    for (int item = blockIdx.x; item < prm.num_items; item += gridDim.x) {
      int n_tile, t0, nk;
      item_range(item, n_tile, t0, nk);
      for (int i = 0; i < nk; ++i, ++it) {
        const uint32_t s = it % WH_STAGES;
        mbar_wait(&empty_bar[s], ((it / WH_STAGES) & 1) ^ 1, 31);
        const int tile = t0 + i;
        const int tw = tile % prm.tiles_w;
        const int r = tile / prm.tiles_w;
        const int th = r % prm.tiles_h;
        const int n = r / prm.tiles_h;
        if (elect_one()) {
          uint8_t* st = smem + s * WH_STAGE;
          mbar_arrive_expect_tx(&full_bar[s], WH_X_BYTES + WH_Z_BYTES);
          tma_load_4d(st, &prm.mapX, &full_bar[s], 0, tw * WH_T - 1, th * WH_T - 1, n);
          tma_load_4d(st + WH_X_SLOT, &prm.mapZ, &full_bar[s], n_tile * 64, tw * WH_T, th * WH_T, n);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, true, true);
    uint32_t it = 0, t = 0;
    for (int item = blockIdx.x; item < prm.num_items; item += gridDim.x, ++t) {
      int n_tile, t0, nk;
      item_range(item, n_tile, t0, nk);
      mbar_wait(acc_empty, (t & 1) ^ 1, 32);
      tc_fence_after_sync();
      for (int i = 0; i < nk; ++i, ++it) {
        const uint32_t s = it % WH_STAGES;
        mbar_wait(&full_bar[s], (it / WH_STAGES) & 1, 33);
        tc_fence_after_sync();
        const uint32_t x0 = smem_u32(smem + s * WH_STAGE);
        const uint64_t zd0 = umma_smem_desc_sw128(x0 + WH_X_SLOT, UMMA_SLAB_BYTES, 1024);
        if (elect_one()) {
#pragma unroll
          for (int g = 0; g < WH_GROUPS; ++g) {
            const int r0 = wh_tap_row(2 * g);
            const int r1 = g < 4 ? wh_tap_row(2 * g + 1) : r0 + 1;      // dummy second slab for tap 8
            // A: two tap views of the X halo stacked along M (LBO = distance between the views)
            const uint64_t xd0 = umma_smem_desc_sw128(x0 + r0 * 128, (r1 - r0) * 128, WH_PITCH * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k)        // 16 pixels = two 8-pixel rows: +2 halo rows / +16 dZ rows
              umma_f16(tmem_base + g * 64, xd0 + k * ((2 * WH_PITCH * 128) >> 4), zd0 + k * ((16 * 128) >> 4),
                       idesc, (i | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(acc_full);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;            // = j * 64 + ci
    const int j = row >> 6, ci = row & 63;
    uint32_t t = 0;
    for (int item = blockIdx.x; item < prm.num_items; item += gridDim.x, ++t) {
      int n_tile, t0, nk;
      item_range(item, n_tile, t0, nk);
      mbar_wait(acc_full, t & 1, 34);
      tc_fence_after_sync();
      if (nk > 0) {
        uint32_t acc[32];
#pragma unroll 1
        for (int pc = half; pc < 2 * WH_GROUPS; pc += 2) {       // (group, 32-column chunk) pairs
          const int g = pc >> 1, c = (pc & 1) * 32;
          __syncwarp();
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * 64 + c, acc);
          tmem_ld_wait();
          const int tap = 2 * g + j;
          if (tap < 9) {
            float* o = prm.dW + (static_cast<long long>(n_tile * 64 + c) * 9 + tap) * 64 + ci;
#pragma unroll
            for (int u = 0; u < 32; ++u)
              if (n_tile * 64 + c + u < prm.Cout)
                atomicAdd(o + static_cast<long long>(u) * 9 * 64, __uint_as_float(acc[u]) * prm.scale);
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace b200
