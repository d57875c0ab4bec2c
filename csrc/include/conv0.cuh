// First convolution (3 -> 64 channels, K = 27) without an im2col matrix in HBM.
//
// The first version expanded the 3-channel image into im2col rows [pixels][64] in global memory
// (411 MB written by the input kernel, read by the fprop GEMM, read again by the wgrad GEMM).  Here
// eight producer warps build the 128B-swizzled operand tile DIRECTLY IN SHARED MEMORY from the
// 8-byte NHWC4 pixels (row = pixel, 64 bf16 = 27 taps*channels + zero tail; 16-byte chunk j of row r
// lives at r*128 + ((j ^ (r & 7)) << 4), exactly what a SWIZZLE_128B TMA load would have written),
// make it visible to the async proxy (fence.proxy.async) and hand it to tcgen05.mma through the
// same full/empty mbarrier ring a TMA producer would use:
//   fprop : the tile is the K-major A operand (rows = M = pixels), B = the 64x64 weight (resident),
//           epilogue bias + ReLU -> bf16 NHWC;
//   wgrad : the same bytes are the MN-major B operand (rows = K = pixels); A = dZ via TMA (MN-major),
//           accumulators stay in TMEM over the CTA's pixel range, fp32 red.add epilogue.
#pragma once
#include "umma_core.cuh"

namespace b200 {

constexpr int C0_STAGES = 4;
constexpr int C0_PROD_WARPS = 8;                                  // 256 producer threads, in row-groups
constexpr int C0_THREADS = UMMA_THREADS + 32 * C0_PROD_WARPS;     // 10 + 8 warps
constexpr int C0_TILE_BYTES = 128 * 128;                          // 128 rows x 64 bf16

struct Conv0Params {
  CUtensorMap mapW;            // fprop: weight [64][64] bf16, box {64, 64}
  CUtensorMap mapZ;            // wgrad: dZ [pixels][64] bf16 as {64, pixels}, box {64, 64}
  const __nv_bfloat16* x4;     // [N][H][W][4] bf16 (3 channels + pad)
  __nv_bfloat16* out;          // fprop: [pixels][64] bf16
  const float* bias;
  float* dW;                   // wgrad: [64][64] fp32
  int N, H, W;
  long long pixels;            // N*H*W
  int num_tiles;               // fprop: ceil(pixels/128); wgrad: k-blocks of 64 pixels
  int blocks_per_cta;          // wgrad: k-blocks per CTA
};

// One im2col row (pixel `pix`) into a swizzled 128-byte smem row.  32-bit index arithmetic: the
// pixel count of a batch (3.2 M at 64 x 224 x 224) is far below 2^31.
__device__ __forceinline__ void c0_build_row(const Conv0Params& p, int pix, uint8_t* tile, int r) {
  uint32_t h[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) h[i] = 0u;
  if (pix < static_cast<int>(p.pixels)) {
    const int w = pix % p.W;
    const int t = pix / p.W;
    const int hh = t % p.H;
    const uint2* base = reinterpret_cast<const uint2*>(p.x4);       // one uint2 = one NHWC4 pixel
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int dy = tap / 3 - 1, dx = tap % 3 - 1;
      const int yy = hh + dy, xx = w + dx;
      uint2 px = make_uint2(0u, 0u);
      if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) px = __ldg(base + (pix + dy * p.W + dx));
      h[tap * 3 + 0] = px.x & 0xffffu;
      h[tap * 3 + 1] = px.x >> 16;
      h[tap * 3 + 2] = px.y & 0xffffu;
    }
  }
  uint8_t* row = tile + r * 128;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (j < 4) v = make_uint4(h[8 * j] | (h[8 * j + 1] << 16), h[8 * j + 2] | (h[8 * j + 3] << 16),
                              h[8 * j + 4] | (h[8 * j + 5] << 16), h[8 * j + 6] | (h[8 * j + 7] << 16));
    *reinterpret_cast<uint4*>(row + ((j ^ (r & 7)) << 4)) = v;
  }
}

template <bool WGRAD>
__global__ void __launch_bounds__(C0_THREADS, 1) conv0_kernel(const __grid_constant__ Conv0Params prm) {
  constexpr int ROWS = WGRAD ? 64 : 128;                       // pixels per stage
  constexpr int A_BYTES = WGRAD ? 2 * UMMA_SLAB_BYTES : 0;     // wgrad: dZ slabs (second one zero-filled)
  constexpr int STAGE = C0_TILE_BYTES + A_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sW = smem + C0_STAGES * STAGE;                      // fprop: resident weight (8 KB)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sW + 8192);
  uint64_t* empty_bar = full_bar + C0_STAGES;
  uint64_t* acc_full = empty_bar + C0_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* w_bar = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C0_STAGES; ++s) {
      // producers arrive once per thread that writes rows; wgrad adds the TMA transaction arrival
      mbar_init(&full_bar[s], (WGRAD ? 64 : 128) + (WGRAD ? 1 : 0));
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], UMMA_EPI_WARPS); }
    mbar_init(w_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // wgrad: this CTA's contiguous range of 64-pixel k-blocks
  const int kb0 = WGRAD ? blockIdx.x * prm.blocks_per_cta : 0;
  const int kb1 = WGRAD ? min(prm.num_tiles, kb0 + prm.blocks_per_cta) : 0;

  if (warp == 0) {
    if constexpr (!WGRAD) {
      if (elect_one()) {
        tma_prefetch_desc(&prm.mapW);
        mbar_arrive_expect_tx(w_bar, 8192);
        tma_load_2d(sW, &prm.mapW, w_bar, 0, 0);
      }
    } else {
      tma_prefetch_desc(&prm.mapZ);
      uint32_t it = 0;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const uint32_t s = it % C0_STAGES;
        mbar_wait(&empty_bar[s], ((it / C0_STAGES) & 1) ^ 1, 41);
        if (elect_one()) {
          uint8_t* sa = smem + s * STAGE + C0_TILE_BYTES;
          mbar_arrive_expect_tx(&full_bar[s], 2 * UMMA_SLAB_BYTES);
          tma_load_2d(sa, &prm.mapZ, &full_bar[s], 0, kb * 64);
          tma_load_2d(sa + UMMA_SLAB_BYTES, &prm.mapZ, &full_bar[s], 64, kb * 64);   // out of bounds -> zeros
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if constexpr (!WGRAD) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, false, false);
      mbar_wait(w_bar, 0, 42);
      tc_fence_after_sync();
      const uint64_t bd0 = umma_smem_desc_sw128(smem_u32(sW), 16, 1024);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x, ++it) {
        const uint32_t s = it % C0_STAGES, buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1, 43);
        mbar_wait(&full_bar[s], (it / C0_STAGES) & 1, 44);
        tc_fence_after_sync();
        const uint64_t ad0 = umma_smem_desc_sw128(smem_u32(smem + s * STAGE), 16, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_base + buf * 64, ad0 + k * 2, bd0 + k * 2, idesc, k != 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);
          umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    } else {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, true, true);
      uint32_t it = 0;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const uint32_t s = it % C0_STAGES;
        mbar_wait(&full_bar[s], (it / C0_STAGES) & 1, 45);
        tc_fence_after_sync();
        const uint32_t st = smem_u32(smem + s * STAGE);
        const uint64_t bd0 = umma_smem_desc_sw128(st, UMMA_SLAB_BYTES, 1024);                    // col tile, MN-major
        const uint64_t ad0 = umma_smem_desc_sw128(st + C0_TILE_BYTES, UMMA_SLAB_BYTES, 1024);    // dZ slabs
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base, ad0 + k * 128, bd0 + k * 128, idesc, (static_cast<int>(it) | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&acc_full[0]);
      __syncwarp();
    }
  } else if (warp < 2 + UMMA_EPI_WARPS) {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    uint32_t acc[32];
    if constexpr (!WGRAD) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it & 1;
        mbar_wait(&acc_full[buf], (it >> 1) & 1, 46);
        tc_fence_after_sync();
        const int pix = tile * 128 + row;
        const int c = half * 32;
        __syncwarp();
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * 64 + c, acc);
        tmem_ld_wait();
        if (pix < prm.pixels) {
          __nv_bfloat16* o = prm.out + static_cast<long long>(pix) * 64 + c;
#pragma unroll
          for (int j = 0; j < 32; j += 16) {     // 64 contiguous bytes per thread: two 32-byte stores
            uint32_t pk[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float lo = fmaxf(__uint_as_float(acc[j + 2 * u]) + __ldg(prm.bias + c + j + 2 * u), 0.f);
              const float hi = fmaxf(__uint_as_float(acc[j + 2 * u + 1]) + __ldg(prm.bias + c + j + 2 * u + 1), 0.f);
              pk[u] = pack_bf16x2(lo, hi);
            }
            st_global_v8(o + j, pk);
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
      }
    } else {
      if (kb1 > kb0) {
        mbar_wait(&acc_full[0], 0, 47);
        tc_fence_after_sync();
        const int c = half * 32;
        __syncwarp();
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, acc);
        tmem_ld_wait();
        if (row < 64) {                       // rows 64..127 belong to the zero-filled dZ slab
          float* o = prm.dW + row * 64 + c;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + j), "f"(__uint_as_float(acc[j])),
                         "f"(__uint_as_float(acc[j + 1])), "f"(__uint_as_float(acc[j + 2])),
                         "f"(__uint_as_float(acc[j + 3]))
                         : "memory");
        }
      }
    }
  } else {
    // ---- operand producers: build swizzled im2col rows in shared memory -----------------------
    // 256 threads in groups of ROWS (one thread per row); group g builds the stages of iterations
    // it = g, g + NGROUPS, ... so several tiles are under construction at once.
    constexpr int NGROUPS = (32 * C0_PROD_WARPS) / ROWS;
    const int pt = threadIdx.x - UMMA_THREADS;          // 0..255
    const int grp = pt / ROWS, r = pt % ROWS;
    if constexpr (!WGRAD) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x, ++it) {
        if (static_cast<int>(it % NGROUPS) != grp) continue;
        const uint32_t s = it % C0_STAGES;
        mbar_wait(&empty_bar[s], ((it / C0_STAGES) & 1) ^ 1, 48);
        c0_build_row(prm, tile * 128 + r, smem + s * STAGE, r);
        fence_proxy_async_smem();                        // generic-proxy stores -> visible to tcgen05
        mbar_arrive(&full_bar[s]);
      }
    } else {
      uint32_t it = 0;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        if (static_cast<int>(it % NGROUPS) != grp) continue;
        const uint32_t s = it % C0_STAGES;
        mbar_wait(&empty_bar[s], ((it / C0_STAGES) & 1) ^ 1, 49);
        c0_build_row(prm, kb * 64 + r, smem + s * STAGE, r);
        fence_proxy_async_smem();
        mbar_arrive(&full_bar[s]);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace b200
