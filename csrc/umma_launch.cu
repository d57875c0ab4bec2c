// Host side of the tcgen05 GEMM / convolution family: builds the TMA tensor maps
// (cuTensorMapEncodeTiled through the runtime's driver entry point, so nothing links libcuda),
// picks the tile width and launches umma_kernel<Policy>.
#include <cudaTypedefs.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "api.h"
#include "conv0.cuh"
#include "conv_halo.cuh"
#include "conv_halo_tma.cuh"
#include "wgrad_halo.cuh"
#include "umma_policies.cuh"
#include "umma_core_dyn.cuh"

namespace b200 {

static std::atomic<long long> g_launches{0};
long long launch_count() { return g_launches.load(); }
void count_launch(int n) { g_launches.fetch_add(n); }
void check_last(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    throw std::runtime_error(std::string("[b200] ") + what + ": " + cudaGetErrorString(e));
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
      throw std::runtime_error("[b200] cuTensorMapEncodeTiled entry point not available");
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

static void encode(CUtensorMap* map, const void* ptr, int rank, const cuuint64_t* dims,
                   const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0)
    throw std::runtime_error("[b200] TMA base pointer must be 16-byte aligned");
  for (int i = 0; i < rank - 1; ++i)
    if (strides_bytes[i] % 16 != 0)
      throw std::runtime_error("[b200] TMA strides must be multiples of 16 bytes");
  CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr),
                           dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf,
             "[b200] cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu box %u %u", (int)r,
             rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    throw std::runtime_error(buf);
  }
}

// 2-D row-major matrix [rows][cols] (row stride ld elements); box = {box_cols, box_rows}.
static void map_2d(CUtensorMap* m, const bf16* p, long long rows, long long cols, long long ld,
                   int box_cols, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t str[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  encode(m, p, 2, dims, str, box);
}
// NHWC activation as {C, W, H, N}; box = {64, Wb, Hb, Nb}.
static void map_nhwc(CUtensorMap* m, const bf16* p, int N, int H, int W, int C, int Wb, int Hb,
                     int Nb) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)Wb, (cuuint32_t)Hb, (cuuint32_t)Nb};
  encode(m, p, 4, dims, str, box);
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) throw std::runtime_error("[b200] cannot query the SM count of the current device");
  }
  return n;
}
static int num_sms() { return sm_count(); }

bool comm_carveout_enabled() {
  static const bool on = [] {
    const char* e = getenv("B200_COMM_CARVEOUT");
    return !(e && e[0] == '0');
  }();
  return on;
}
void prefer_max_shared_carveout(const void* kernel) {
  if (!comm_carveout_enabled()) return;
  const cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                               cudaSharedmemCarveoutMaxShared);
  if (err != cudaSuccess)
    throw std::runtime_error(std::string("[b200] cudaFuncSetAttribute(carveout): ") + cudaGetErrorString(err));
}

// EXPERIMENTAL (B200_DYNAMIC_TILES=1, not yet run on hardware): dynamic tile scheduler, see
// umma_core_dyn.cuh.  64 device counters used round-robin; a counter is never reset -- the host
// tracks how far every launch advances it (num_tiles claims + one over-claim per CTA).
static bool dynamic_tiles_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200_DYNAMIC_TILES");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v != 0;
}

static TileTicket next_ticket(int num_tiles, int grid) {
  constexpr int SLOTS = 64;
  static unsigned int* counters = nullptr;
  static unsigned int bases[SLOTS] = {};
  static unsigned int seq = 0;
  static int device = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!counters) {
    if (cudaMalloc(&counters, SLOTS * sizeof(unsigned int)) != cudaSuccess ||
        cudaMemset(counters, 0, SLOTS * sizeof(unsigned int)) != cudaSuccess)
      throw std::runtime_error("[b200] tile counters: allocation failed");
    device = dev;
  }
  if (dev != device) throw std::runtime_error("[b200] dynamic tile scheduling supports one device per process");
  const unsigned int j = seq++ % SLOTS;
  TileTicket t{counters + j, bases[j]};
  bases[j] += static_cast<unsigned int>(num_tiles) + static_cast<unsigned int>(grid);
  return t;
}

template <class P>
static void launch_dyn(const typename P::Params& prm, cudaStream_t stream) {
  dim3 grid(prm.num_tiles < num_sms() ? prm.num_tiles : num_sms());
  constexpr int smem = umma_smem_bytes<P::BN, P::STAGES>() + (UMMA_DYN_BAR_BYTES - 256) + P::EPI_SMEM;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(umma_kernel_dyn<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess)
      throw std::runtime_error(std::string("[b200] cudaFuncSetAttribute(dyn): ") + cudaGetErrorString(e));
    configured = true;
  }
  umma_kernel_dyn<P><<<grid, UMMA_THREADS, smem, stream>>>(prm, next_ticket(prm.num_tiles, grid.x));
  count_launch();
  check_last("umma_kernel_dyn launch");
}

// Persistent launch: one CTA per SM (or fewer when there are fewer tiles).
template <class P>
static void launch(const typename P::Params& prm, cudaStream_t stream) {
  if (prm.num_tiles <= 0) return;
  if (dynamic_tiles_enabled()) {
    launch_dyn<P>(prm, stream);
    return;
  }
  dim3 grid(prm.num_tiles < num_sms() ? prm.num_tiles : num_sms());
  constexpr int smem = umma_smem_bytes<P::BN, P::STAGES>() + P::EPI_SMEM;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(umma_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess)
      throw std::runtime_error(std::string("[b200] cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    configured = true;
  }
  umma_kernel<P><<<grid, UMMA_THREADS, smem, stream>>>(prm);
  count_launch();
  check_last("umma_kernel launch");
}

static int largest_pow2_divisor(int x, int cap) {
  int p = 1;
  while (p * 2 <= cap && x % (p * 2) == 0) p *= 2;
  return p;
}
// Pixel box {Wb, Hb, Nb} with Wb*Hb*Nb == pixels, following the power-of-two factors of W and H
// (224 = 7*32 -> 32x4x1, 112 -> 16x8x1, 56 -> 8x8x2, 28 -> 4x4x8, 14 -> 2x2x32).
static ConvTile make_tile(int N, int H, int W, int pixels) {
  ConvTile t;
  t.N = N; t.H = H; t.W = W;
  t.Wb = largest_pow2_divisor(W, pixels < 32 ? pixels : 32);
  t.Hb = largest_pow2_divisor(H, pixels / t.Wb);
  t.Nb = pixels / (t.Wb * t.Hb);
  t.tiles_w = (W + t.Wb - 1) / t.Wb;
  t.tiles_h = (H + t.Hb - 1) / t.Hb;
  t.wb_shift = 0; while ((1 << t.wb_shift) < t.Wb) ++t.wb_shift;
  t.hb_shift = 0; while ((1 << t.hb_shift) < t.Hb) ++t.hb_shift;
  t.div_tw = make_fastdiv(t.tiles_w);
  t.div_th = make_fastdiv(t.tiles_h);
  return t;
}

// NHWC activation with an explicit pixel box {64, bw, bh, bn} (the halo kernel's (8+2) x (16+2)).
static void map_nhwc_box(CUtensorMap* m, const bf16* p, int N, int H, int W, int C, int bw, int bh, int bn) {
  map_nhwc(m, p, N, H, W, C, bw, bh, bn);
}

static bool halo_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200_CONV_HALO");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}
// The halo kernel tiles the image in 8 x 16 pixel blocks: use it where that wastes nothing.
static bool halo_applicable(int H, int W) { return halo_enabled() && W % HALO_WT == 0 && H % HALO_HT == 0 && W >= 32; }

template <int BN, bool DGRAD, bool POOL = false>
static void conv_halo_launch(ConvParams& prm, const bf16* act, cudaStream_t stream) {
  using Cfg = HaloCfg<BN>;
  ConvTile& t = prm.t;
  t.Wb = HALO_WT; t.Hb = HALO_HT; t.Nb = 1; t.wb_shift = 3; t.hb_shift = 4;
  t.tiles_w = t.W / HALO_WT; t.tiles_h = t.H / HALO_HT;
  t.div_tw = make_fastdiv(t.tiles_w); t.div_th = make_fastdiv(t.tiles_h);
  prm.tiles_m = t.tiles_w * t.tiles_h * t.N;
  prm.div_tm = make_fastdiv(prm.tiles_m);
  const int tiles_n = (prm.Cn + BN - 1) / BN;
  prm.num_tiles = prm.tiles_m * tiles_n;
  prm.resident = (tiles_n == 1 && 9 * (prm.Ca / UMMA_BK) <= Cfg::NB) ? 1 : 0;
  map_nhwc_box(&prm.mapA, act, t.N, t.H, t.W, prm.Ca, HALO_PITCH, HALO_HT + 2, 1);
  constexpr int smem = Cfg::SMEM + ConvPolicy<BN, 1, DGRAD, POOL>::EPI_SMEM;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<BN, DGRAD, POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess)
      throw std::runtime_error(std::string("[b200] cudaFuncSetAttribute(halo): ") + cudaGetErrorString(e));
    configured = true;
  }
  dim3 grid(prm.num_tiles < num_sms() ? prm.num_tiles : num_sms());
  conv_halo_kernel<BN, DGRAD, POOL><<<grid, UMMA_THREADS, smem, stream>>>(prm);
  count_launch();
  check_last("conv_halo_kernel launch");
}

// EXPERIMENTAL (B200_HALO_TMA_EPI=1, not yet run on hardware): halo kernel whose epilogue goes
// through a shared-memory staging tile and TMA (conv_halo_tma.cuh).  Returns false when the layer's
// rings + staging do not fit in shared memory: the caller then uses the regular halo kernel.
static bool halo_tma_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200_HALO_TMA_EPI");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v != 0;
}

template <int BN, bool DGRAD>
static bool conv_halo_tma_launch(ConvParams& prm, const bf16* act, cudaStream_t stream) {
  constexpr int B_SLOT = BN * 128;
  constexpr int MAX_SMEM = 227 * 1024;
  ConvTile& t = prm.t;
  t.Wb = HALO_WT; t.Hb = HALO_HT; t.Nb = 1; t.wb_shift = 3; t.hb_shift = 4;
  t.tiles_w = t.W / HALO_WT; t.tiles_h = t.H / HALO_HT;
  t.div_tw = make_fastdiv(t.tiles_w); t.div_th = make_fastdiv(t.tiles_h);
  prm.tiles_m = t.tiles_w * t.tiles_h * t.N;
  prm.div_tm = make_fastdiv(prm.tiles_m);
  const int tiles_n = (prm.Cn + BN - 1) / BN;
  prm.num_tiles = prm.tiles_m * tiles_n;
  const int cch = prm.Ca / UMMA_BK;
  const bool masked = DGRAD && prm.mask_src != nullptr;
  const int budget = MAX_SMEM - halo_tma_smem<BN, DGRAD>(0, 0, masked);
  ConvTmaExtra ext;
  if (tiles_n == 1 && 9 * cch <= HALO_TMA_MAX_NB && 9 * cch * B_SLOT + 2 * HALO_SLOT <= budget) {
    prm.resident = 1;
    ext.nb = 9 * cch;
    ext.na = (budget - ext.nb * B_SLOT) / HALO_SLOT;
  } else {
    prm.resident = 0;
    ext.na = 3;
    ext.nb = (budget - ext.na * HALO_SLOT) / B_SLOT;
  }
  if (ext.na > HALO_TMA_MAX_NA) ext.na = HALO_TMA_MAX_NA;
  if (ext.nb > HALO_TMA_MAX_NB) ext.nb = HALO_TMA_MAX_NB;
  if (ext.na < 2 || ext.nb < 2) return false;
  map_nhwc_box(&prm.mapA, act, t.N, t.H, t.W, prm.Ca, HALO_PITCH, HALO_HT + 2, 1);
  map_nhwc(&ext.mapOut, prm.out, t.N, t.H, t.W, prm.Cn, HALO_WT, HALO_HT, 1);
  if (masked) map_nhwc(&ext.mapMask, prm.mask_src, t.N, t.H, t.W, prm.Cn, HALO_WT, HALO_HT, 1);
  else ext.mapMask = ext.mapOut;
  const int smem = halo_tma_smem<BN, DGRAD>(ext.na, ext.nb, masked);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_tma_kernel<BN, DGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         MAX_SMEM);
    if (e != cudaSuccess)
      throw std::runtime_error(std::string("[b200] cudaFuncSetAttribute(halo tma): ") + cudaGetErrorString(e));
    configured = true;
  }
  dim3 grid(prm.num_tiles < num_sms() ? prm.num_tiles : num_sms());
  conv_halo_tma_kernel<BN, DGRAD><<<grid, UMMA_THREADS, smem, stream>>>(prm, ext);
  count_launch();
  check_last("conv_halo_tma_kernel launch");
  return true;
}

// ------------------------------------------------------------------------- first conv (no im2col)
template <bool WGRAD>
static void conv0_launch(Conv0Params& prm, int grid, cudaStream_t stream) {
  constexpr int stage = C0_TILE_BYTES + (WGRAD ? 2 * UMMA_SLAB_BYTES : 0);
  constexpr int smem = C0_STAGES * stage + 8192 + 512 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv0_kernel<WGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess)
      throw std::runtime_error(std::string("[b200] cudaFuncSetAttribute(conv0): ") + cudaGetErrorString(e));
    configured = true;
  }
  conv0_kernel<WGRAD><<<grid, C0_THREADS, smem, stream>>>(prm);
  count_launch();
  check_last("conv0_kernel launch");
}

void conv0_fprop(const bf16* x4, const bf16* w0, const float* bias, bf16* y, int N, int H, int W,
                 cudaStream_t stream) {
  Conv0Params prm;
  prm.x4 = x4; prm.out = y; prm.bias = bias; prm.dW = nullptr;
  prm.N = N; prm.H = H; prm.W = W; prm.pixels = static_cast<long long>(N) * H * W;
  prm.num_tiles = static_cast<int>((prm.pixels + 127) / 128);
  prm.blocks_per_cta = 0;
  map_2d(&prm.mapW, w0, 64, 64, 64, 64, 64);
  prm.mapZ = prm.mapW;
  conv0_launch<false>(prm, prm.num_tiles < num_sms() ? prm.num_tiles : num_sms(), stream);
}

void conv0_wgrad(const bf16* dz, const bf16* x4, float* dw0, int N, int H, int W, cudaStream_t stream) {
  Conv0Params prm;
  prm.x4 = x4; prm.out = nullptr; prm.bias = nullptr; prm.dW = dw0;
  prm.N = N; prm.H = H; prm.W = W; prm.pixels = static_cast<long long>(N) * H * W;
  prm.num_tiles = static_cast<int>((prm.pixels + 63) / 64);
  int grid = prm.num_tiles < num_sms() ? prm.num_tiles : num_sms();
  prm.blocks_per_cta = (prm.num_tiles + grid - 1) / grid;
  grid = (prm.num_tiles + prm.blocks_per_cta - 1) / prm.blocks_per_cta;
  map_2d(&prm.mapZ, dz, prm.pixels, 64, 64, 64, 64);
  prm.mapW = prm.mapZ;
  conv0_launch<true>(prm, grid, stream);
}

// ---------------------------------------------------------------------------------------- GEMM
template <int BN, bool A_MN, bool B_MN, int EPI>
static void gemm_launch(GemmParams& prm, const bf16* A, long long lda, const bf16* B, long long ldb,
                        int M, int N, int K, int ksplit, cudaStream_t stream) {
  constexpr int STAGES = BN >= 256 ? 4 : (BN >= 128 ? 6 : 8);
  using P = GemmPolicy<BN, STAGES, A_MN, B_MN, EPI>;
  if (A_MN) map_2d(&prm.mapA, A, K, M, lda, 64, 64); else map_2d(&prm.mapA, A, M, K, lda, 64, UMMA_BM);
  if (B_MN) map_2d(&prm.mapB, B, K, N, ldb, 64, 64); else map_2d(&prm.mapB, B, N, K, ldb, 64, BN);
  prm.tiles_m = (M + UMMA_BM - 1) / UMMA_BM;
  prm.tiles_n = (N + BN - 1) / BN;
  prm.num_tiles = prm.tiles_m * prm.tiles_n * ksplit;
  launch<P>(prm, stream);
}

#define GEMM_BN_SWITCH(AMN, BMN, EPI)                                                       \
  switch (bn) {                                                                             \
    case 64:  gemm_launch<64,  AMN, BMN, EPI>(prm, A, lda, B, ldb, M, N, K, ksplit, stream); return; \
    case 128: gemm_launch<128, AMN, BMN, EPI>(prm, A, lda, B, ldb, M, N, K, ksplit, stream); return; \
    case 256: gemm_launch<256, AMN, BMN, EPI>(prm, A, lda, B, ldb, M, N, K, ksplit, stream); return; \
    default: break;                                                                         \
  }
#define GEMM_BN_SWITCH_32(AMN, BMN, EPI)                                                    \
  if (bn == 32) { gemm_launch<32, AMN, BMN, EPI>(prm, A, lda, B, ldb, M, N, K, ksplit, stream); return; } \
  GEMM_BN_SWITCH(AMN, BMN, EPI)

void gemm_bf16(const bf16* A, long long lda, bool a_mn, const bf16* B, long long ldb, bool b_mn,
               int M, int N, int K, void* out, long long ldo, int epi, const float* bias,
               float alpha, int ksplit, int bn, cudaStream_t stream) {
  if (bn == 0) bn = N <= 32 && !b_mn ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
  if (ksplit < 1) ksplit = 1;
  GemmParams prm;
  prm.M = M; prm.N = N;
  prm.k_iters_total = (K + UMMA_BK - 1) / UMMA_BK;
  prm.k_iters_per_split = (prm.k_iters_total + ksplit - 1) / ksplit;
  ksplit = (prm.k_iters_total + prm.k_iters_per_split - 1) / prm.k_iters_per_split;
  prm.out = out; prm.ldo = ldo; prm.bias = bias; prm.alpha = alpha;
  if (ksplit > 1 && !(epi == EPI_F32_ATOMIC || epi == EPI_F32_ATOMIC_T))
    throw std::runtime_error("[b200] gemm_bf16: split-K needs an atomic epilogue");
  if (!a_mn && !b_mn) {
    if (epi == EPI_F32_ATOMIC_T) { GEMM_BN_SWITCH_32(false, false, EPI_F32_ATOMIC_T) }
    if (epi == EPI_F32_STORE_T) { GEMM_BN_SWITCH_32(false, false, EPI_F32_STORE_T) }
    if (epi == EPI_BF16_BIAS_RELU) { GEMM_BN_SWITCH(false, false, EPI_BF16_BIAS_RELU) }
    if (epi == EPI_F32_STORE) { GEMM_BN_SWITCH(false, false, EPI_F32_STORE) }
  } else if (a_mn && !b_mn) {
    if (epi == EPI_F32_ATOMIC_T) { GEMM_BN_SWITCH_32(true, false, EPI_F32_ATOMIC_T) }
    if (epi == EPI_F32_STORE_T) { GEMM_BN_SWITCH_32(true, false, EPI_F32_STORE_T) }
  } else if (a_mn && b_mn) {
    if (epi == EPI_F32_STORE) { GEMM_BN_SWITCH(true, true, EPI_F32_STORE) }
    if (epi == EPI_F32_ATOMIC) { GEMM_BN_SWITCH(true, true, EPI_F32_ATOMIC) }
    if (epi == EPI_BF16_STORE) { GEMM_BN_SWITCH(true, true, EPI_BF16_STORE) }
  }
  throw std::runtime_error("[b200] gemm_bf16: unsupported (layout, epilogue, bn) combination");
}

// ---------------------------------------------------------------------------------------- conv
static void check_channels(int c, const char* what) {
  if (c % 64 != 0) throw std::runtime_error(std::string("[b200] ") + what + " must be a multiple of 64");
}

template <int BN, bool DGRAD, bool POOL = false>
static void conv_launch(ConvParams& prm, cudaStream_t stream) {
  constexpr int STAGES = BN >= 256 ? 4 : (BN >= 128 ? 6 : 8);
  using P = ConvPolicy<BN, STAGES, DGRAD, POOL>;
  const ConvTile& t = prm.t;
  const int tiles_n = (t.N + t.Nb - 1) / t.Nb;
  prm.tiles_m = t.tiles_w * t.tiles_h * tiles_n;
  prm.div_tm = make_fastdiv(prm.tiles_m);
  prm.num_tiles = prm.tiles_m * ((prm.Cn + BN - 1) / BN);
  launch<P>(prm, stream);
}

static int auto_bn(int cn, int bn) {
  if (bn) return bn;
  return cn <= 64 ? 64 : (cn <= 128 ? 128 : 256);
}

static void fprop_impl(const bf16* x, const bf16* w, const float* bias, bf16* y, bf16* pool_out,
                       uint32_t* pool_mask, int N, int H, int W, int Cin, int Cout, bool relu, int bn,
                       cudaStream_t stream) {
  check_channels(Cin, "conv3x3_fprop Cin");
  check_channels(Cout, "conv3x3_fprop Cout");
  ConvParams prm;
  prm.t = make_tile(N, H, W, UMMA_BM);
  prm.Ca = Cin; prm.Cn = Cout; prm.wcols_per_tap = Cin;
  prm.out = y; prm.bias = bias; prm.mask_src = nullptr; prm.colsum = nullptr;
  prm.pool_out = pool_out; prm.pool_mask = pool_mask;
  prm.flags = (bias ? CONV_BIAS : 0) | (relu ? CONV_RELU : 0) | (pool_out ? CONV_POOL : 0);
  bn = auto_bn(Cout, bn);
  prm.resident = 0;
  map_2d(&prm.mapB, w, Cout, 9LL * Cin, 9LL * Cin, 64, bn);
  const bool pool = pool_out != nullptr;
  if (halo_applicable(H, W) && halo_tma_enabled() && !pool && bn <= 128) {
    if (bn == 64 ? conv_halo_tma_launch<64, false>(prm, x, stream) : conv_halo_tma_launch<128, false>(prm, x, stream))
      return;
  }
  if (halo_applicable(H, W)) {
    switch (bn) {
      case 64: pool ? conv_halo_launch<64, false, true>(prm, x, stream) : conv_halo_launch<64, false>(prm, x, stream); return;
      case 128: pool ? conv_halo_launch<128, false, true>(prm, x, stream) : conv_halo_launch<128, false>(prm, x, stream); return;
      case 256: pool ? conv_halo_launch<256, false, true>(prm, x, stream) : conv_halo_launch<256, false>(prm, x, stream); return;
      default: throw std::runtime_error("[b200] conv3x3_fprop: bn must be 64/128/256");
    }
  }
  map_nhwc(&prm.mapA, x, N, H, W, Cin, prm.t.Wb, prm.t.Hb, prm.t.Nb);
  switch (bn) {
    case 64: pool ? conv_launch<64, false, true>(prm, stream) : conv_launch<64, false>(prm, stream); break;
    case 128: pool ? conv_launch<128, false, true>(prm, stream) : conv_launch<128, false>(prm, stream); break;
    case 256: pool ? conv_launch<256, false, true>(prm, stream) : conv_launch<256, false>(prm, stream); break;
    default: throw std::runtime_error("[b200] conv3x3_fprop: bn must be 64/128/256");
  }
}

void conv3x3_fprop(const bf16* x, const bf16* w, const float* bias, bf16* y, int N, int H, int W,
                   int Cin, int Cout, bool relu, int bn, cudaStream_t stream) {
  fprop_impl(x, w, bias, y, nullptr, nullptr, N, H, W, Cin, Cout, relu, bn, stream);
}

// EXPERIMENTAL: the pool windows must lie inside one epilogue warp (see ConvPolicy::epilogue_pool).
bool conv3x3_pool_fusable(int N, int H, int W) {
  if (H % 2 || W % 2) return false;
  if (halo_applicable(H, W)) return true;               // 8 x 16 tiles
  const ConvTile t = make_tile(N, H, W, UMMA_BM);
  return t.Wb <= 16 && t.Hb >= 2 && W % t.Wb == 0 && H % t.Hb == 0;
}

void conv3x3_fprop_pool(const bf16* x, const bf16* w, const float* bias, bf16* pool_out, uint32_t* pool_mask,
                        int N, int H, int W, int Cin, int Cout, int bn, cudaStream_t stream) {
  if (!conv3x3_pool_fusable(N, H, W))
    throw std::runtime_error("[b200] conv3x3_fprop_pool: this image size cannot fuse the 2x2 pool");
  if (!pool_out || !pool_mask) throw std::runtime_error("[b200] conv3x3_fprop_pool: missing outputs");
  fprop_impl(x, w, bias, nullptr, pool_out, pool_mask, N, H, W, Cin, Cout, true, bn, stream);
}

void conv3x3_dgrad(const bf16* dz, const bf16* w, const bf16* mask_src, bf16* dx, float* colsum, int N,
                   int H, int W, int Cin, int Cout, int bn, cudaStream_t stream) {
  check_channels(Cin, "conv3x3_dgrad Cin");
  check_channels(Cout, "conv3x3_dgrad Cout");
  ConvParams prm;
  prm.t = make_tile(N, H, W, UMMA_BM);
  prm.Ca = Cout; prm.Cn = Cin; prm.wcols_per_tap = Cin;
  prm.out = dx; prm.bias = nullptr; prm.mask_src = mask_src; prm.colsum = colsum;
  prm.pool_out = nullptr; prm.pool_mask = nullptr;
  prm.flags = (mask_src ? CONV_MASK : 0) | (colsum ? CONV_COLSUM : 0);
  if (colsum && Cin > 512) throw std::runtime_error("[b200] conv3x3_dgrad: fused column sum supports Cin <= 512");
  bn = auto_bn(Cin, bn);
  prm.resident = 0;
  map_2d(&prm.mapB, w, Cout, 9LL * Cin, 9LL * Cin, 64, 64);   // MN-major: 64 ci x 64 co boxes
  if (halo_applicable(H, W) && halo_tma_enabled() && bn <= 128) {
    if (bn == 64 ? conv_halo_tma_launch<64, true>(prm, dz, stream) : conv_halo_tma_launch<128, true>(prm, dz, stream))
      return;
  }
  if (halo_applicable(H, W)) {
    switch (bn) {
      case 64: conv_halo_launch<64, true>(prm, dz, stream); return;
      case 128: conv_halo_launch<128, true>(prm, dz, stream); return;
      case 256: conv_halo_launch<256, true>(prm, dz, stream); return;
      default: throw std::runtime_error("[b200] conv3x3_dgrad: bn must be 64/128/256");
    }
  }
  map_nhwc(&prm.mapA, dz, N, H, W, Cout, prm.t.Wb, prm.t.Hb, prm.t.Nb);
  switch (bn) {
    case 64: conv_launch<64, true>(prm, stream); break;
    case 128: conv_launch<128, true>(prm, stream); break;
    case 256: conv_launch<256, true>(prm, stream); break;
    default: throw std::runtime_error("[b200] conv3x3_dgrad: bn must be 64/128/256");
  }
}

template <int BN>
static void wgrad_launch(WgradParams& prm, cudaStream_t stream) {
  constexpr int STAGES = BN >= 256 ? 4 : (BN >= 128 ? 6 : 8);
  using P = WgradPolicy<BN, STAGES>;
  prm.tiles_m = (prm.Cout + UMMA_BM - 1) / UMMA_BM;
  prm.tiles_n = (prm.Cin + BN - 1) / BN;
  prm.num_tiles = prm.tiles_m * prm.tiles_n * 9 * prm.ksplit;
  launch<P>(prm, stream);
}

// 64-input-channel layers on 8-divisible maps: nine tap views of one X halo per pixel tile.
static void wgrad_halo64_launch(const bf16* dz, const bf16* x, float* dw, int N, int H, int W, int Cout,
                                float scale, cudaStream_t stream) {
  WgradHaloParams prm;
  prm.tiles_n = (Cout + 63) / 64;
  prm.tiles_w = W / WH_T; prm.tiles_h = H / WH_T;
  prm.total_tiles = N * prm.tiles_w * prm.tiles_h;
  int ksplit = num_sms() / prm.tiles_n;
  if (ksplit < 1) ksplit = 1;
  if (ksplit > prm.total_tiles) ksplit = prm.total_tiles;
  prm.tiles_per_split = (prm.total_tiles + ksplit - 1) / ksplit;
  ksplit = (prm.total_tiles + prm.tiles_per_split - 1) / prm.tiles_per_split;
  prm.num_items = prm.tiles_n * ksplit;
  prm.Cout = Cout; prm.dW = dw; prm.scale = scale;
  map_nhwc(&prm.mapX, x, N, H, W, 64, WH_PITCH, WH_PITCH, 1);
  map_nhwc(&prm.mapZ, dz, N, H, W, Cout, WH_T, WH_T, 1);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_halo64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WH_SMEM);
    if (e != cudaSuccess)
      throw std::runtime_error(std::string("[b200] cudaFuncSetAttribute(wgrad halo): ") + cudaGetErrorString(e));
    configured = true;
  }
  dim3 grid(prm.num_items < num_sms() ? prm.num_items : num_sms());
  wgrad_halo64_kernel<<<grid, UMMA_THREADS, WH_SMEM, stream>>>(prm);
  count_launch();
  check_last("wgrad_halo64_kernel launch");
}

void conv3x3_wgrad(const bf16* dz, const bf16* x, float* dw, int N, int H, int W, int Cin,
                   int Cout, float scale, int ksplit, int bn, cudaStream_t stream) {
  check_channels(Cin, "conv3x3_wgrad Cin");
  check_channels(Cout, "conv3x3_wgrad Cout");
  if (Cin == 64 && H % WH_T == 0 && W % WH_T == 0 && W >= 16 && halo_enabled() && bn == 0 && ksplit <= 0) {
    wgrad_halo64_launch(dz, x, dw, N, H, W, Cout, scale, stream);
    return;
  }
  WgradParams prm;
  prm.t = make_tile(N, H, W, 64);
  const int tiles_n = (N + prm.t.Nb - 1) / prm.t.Nb;
  prm.Cout = Cout; prm.Cin = Cin;
  prm.total_tiles = prm.t.tiles_w * prm.t.tiles_h * tiles_n;
  bn = auto_bn(Cin, bn);
  if (ksplit <= 0) {     // work items of <= ~64 k-blocks (4096 pixels), and at least ~4 per SM
    const int base = ((Cout + UMMA_BM - 1) / UMMA_BM) * ((Cin + bn - 1) / bn) * 9;
    ksplit = (prm.total_tiles + 63) / 64;
    const int min_split = (4 * num_sms() + base - 1) / base;
    if (ksplit < min_split) ksplit = min_split;
  }
  if (ksplit > prm.total_tiles) ksplit = prm.total_tiles;
  prm.tiles_per_split = (prm.total_tiles + ksplit - 1) / ksplit;
  prm.ksplit = (prm.total_tiles + prm.tiles_per_split - 1) / prm.tiles_per_split;
  prm.dW = dw; prm.scale = scale;
  map_nhwc(&prm.mapA, dz, N, H, W, Cout, prm.t.Wb, prm.t.Hb, prm.t.Nb);
  map_nhwc(&prm.mapB, x, N, H, W, Cin, prm.t.Wb, prm.t.Hb, prm.t.Nb);
  switch (bn) {
    case 64: wgrad_launch<64>(prm, stream); break;
    case 128: wgrad_launch<128>(prm, stream); break;
    case 256: wgrad_launch<256>(prm, stream); break;
    default: throw std::runtime_error("[b200] conv3x3_wgrad: bn must be 64/128/256");
  }
}

}  // namespace b200
