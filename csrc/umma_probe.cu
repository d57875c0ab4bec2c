// Hardware probe (not on the training path): does tcgen05.mma accept a K-major SWIZZLE_128B
// operand whose start address is NOT 1024-byte aligned (descriptor base_offset = (addr >> 7) & 7)
// and whose 8-row groups are NOT 1024 bytes apart (SBO != 1024)?  If yes, the nine taps of a 3x3
// convolution can be read as nine views of ONE halo tile in shared memory instead of nine TMA
// loads.  The kernel loads `rows` rows of A (64 bf16 each) with a single TMA box, then multiplies
// the 128-row view  row(r) = A[(r / 8) * group_pitch + r % 8 + shift]  by B^T (N = 64, K = 64).
#include <cudaTypedefs.h>

#include <stdexcept>
#include <string>

#include "api.h"
#include "ptx.cuh"

namespace b200 {

struct ProbeParams {
  CUtensorMap mapA;   // [rows][64] bf16, box {64, rows}
  CUtensorMap mapB;   // [64][64] bf16, box {64, 64}
  float* out;         // [128][64]
  int rows;
  int shift;          // first row of the view
  int group_pitch;    // rows between consecutive 8-row groups (8 = dense)
  int use_base_offset;
  int mode;           // 0: shifted K-major A view; 1: A dense, B read MN-major from the `rows` buffer
                      //    (mapA holds the [rows][64] K-by-N matrix, mapB the dense [128..][64] A)
};

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const __grid_constant__ ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                    // up to 256 rows x 128 B = 32 KB
  uint8_t* sB = smem + 32768;            // up to 128 x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 16384);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 64);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, p.rows * 128 + 64 * 128);
    tma_load_2d(sA, &p.mapA, bar, 0, 0);
    tma_load_2d(sB, &p.mapB, bar, 0, 0);
    mbar_wait(bar, 0, 10);
    tc_fence_after_sync();
    if (p.mode == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, false, false);
      const uint32_t a0 = smem_u32(sA) + p.shift * 128;
      const uint32_t b0 = smem_u32(sB);
      for (int k = 0; k < 4; ++k) {
        uint64_t ad = umma_smem_desc_sw128(a0 + k * 32, 16, p.group_pitch * 128);
        if (!p.use_base_offset) ad &= ~(7ull << 49);
        const uint64_t bd = umma_smem_desc_sw128(b0 + k * 32, 16, 1024);
        umma_f16(tmem, ad, bd, idesc, k != 0);
      }
    } else {
      // D[m][n] = sum_k Adense[m][k] * KN[k + shift][n]; Adense (64 rows valid) sits in sB, KN in sA.
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, false, true);
      const uint32_t a0 = smem_u32(sB);
      const uint32_t b0 = smem_u32(sA) + p.shift * 128;
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = umma_smem_desc_sw128(a0 + k * 32, 16, 1024);
        uint64_t bd = umma_smem_desc_sw128(b0 + k * 16 * 128, 8192, 1024);
        if (!p.use_base_offset) bd &= ~(7ull << 49);
        umma_f16(tmem, ad, bd, idesc, k != 0);
      }
    }
    umma_commit(done);
  }
  __syncwarp();
  mbar_wait(done, 0, 11);
  tc_fence_after_sync();
  uint32_t acc[32];
  const int row = warp * 32 + lane;
  for (int c = 0; c < 64; c += 32) {
    __syncwarp();
    tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, acc);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) p.out[row * 64 + c + j] = __uint_as_float(acc[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem, 64);
  }
}

void umma_shift_probe(const bf16* A, int rows, const bf16* B, float* out, int shift, int group_pitch,
                      int use_base_offset, int mode, cudaStream_t stream) {
  if (rows > 256 || (mode == 0 && (127 / 8) * group_pitch + 7 + shift >= rows) || (mode == 1 && 64 + shift > rows))
    throw std::runtime_error("[b200] umma_shift_probe: view exceeds the loaded rows");
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  auto fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fnp);
  ProbeParams p;
  cuuint32_t estr[2] = {1, 1};
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    cuuint64_t str[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)rows};
    fn(&p.mapA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(A), dims, str, box, estr,
       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  {
    cuuint64_t dims[2] = {64, 64};
    cuuint64_t str[1] = {128};
    cuuint32_t box[2] = {64, 64};
    fn(&p.mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(B), dims, str, box, estr,
       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  p.out = out; p.rows = rows; p.shift = shift; p.group_pitch = group_pitch; p.use_base_offset = use_base_offset; p.mode = mode;
  const int smem = 32768 + 16384 + 64 + 1024;
  cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  umma_probe_kernel<<<1, 128, smem, stream>>>(p);
  check_last("umma_probe_kernel");
}

}  // namespace b200
