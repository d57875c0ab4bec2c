// Host-callable entry points of the native library (raw pointers + stream; no torch types so the
// .cu files compile in seconds).  bindings.cpp adapts torch tensors onto these.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace b200 {

typedef __nv_bfloat16 bf16;

// Number of kernels launched by this library since load (bench.py reports it as gpu_launches).
long long launch_count();
void count_launch(int n = 1);
void check_last(const char* what);
// multiProcessorCount of the current device (cached; grids are sized from it, never from a literal)
int sm_count();
// Kernels that run on the comm stream UNDER the persistent conv kernels must ask for the same L1 /
// shared-memory split as those (maximum shared memory): an SM only changes its carve-out when it is
// empty, so a resident CTA of a kernel with the default (L1-heavy) preference keeps every conv CTA --
// which needs > 200 KB of shared memory -- off that SM until it exits (bench/interference.py: a
// background kernel that only sleeps slowed the convolutions by 1.3-1.6x).  Call once per kernel.
void prefer_max_shared_carveout(const void* kernel);
bool comm_carveout_enabled();   // B200_COMM_CARVEOUT=0 switches the above off (A/B measurements)

// ---- tcgen05 GEMM family (umma_launch.cu) -----------------------------------------------------
// D[M,N] = alpha * sum_k A[m,k] * B[n,k].  a_mn: A is stored [K][M] (row stride lda) instead of
// [M][K]; same for b_mn.  epi: GemmEpi.  ksplit > 1 needs an ATOMIC epilogue and a zeroed output.
void gemm_bf16(const bf16* A, long long lda, bool a_mn, const bf16* B, long long ldb, bool b_mn,
               int M, int N, int K, void* out, long long ldo, int epi, const float* bias,
               float alpha, int ksplit, int bn, cudaStream_t stream);

// First convolution (3 -> 64 channels) straight from NHWC4 pixels: the im2col operand tile is built
// in shared memory by producer warps (conv0.cuh).  w0 is [64][64] bf16 (k = (kh*3+kw)*3+c, zero tail).
void conv0_fprop(const bf16* x4, const bf16* w0, const float* bias, bf16* y, int N, int H, int W,
                 cudaStream_t stream);
void conv0_wgrad(const bf16* dz, const bf16* x4, float* dw0, int N, int H, int W, cudaStream_t stream);

// 3x3 / stride 1 / pad 1 convolution, NHWC bf16, weights [Cout][3][3][Cin] bf16.
void conv3x3_fprop(const bf16* x, const bf16* w, const float* bias, bf16* y, int N, int H, int W,
                   int Cin, int Cout, bool relu, int bn, cudaStream_t stream);
// EXPERIMENTAL (B200_FUSE_POOL=1): conv + bias + ReLU + 2x2/2 max-pool in one kernel.  Writes the
// pooled activation [N,H/2,W/2,Cout] and, per pooled pixel and 32 channels, four words {argmax bit 0,
// argmax bit 1, max > 0, 0}; the un-pooled activation is never materialised.  unpool2x2 is the
// matching backward (elementwise.cu).
bool conv3x3_pool_fusable(int N, int H, int W);
void conv3x3_fprop_pool(const bf16* x, const bf16* w, const float* bias, bf16* pool_out, uint32_t* pool_mask,
                        int N, int H, int W, int Cin, int Cout, int bn, cudaStream_t stream);
// dx[N,H,W,Cin] = conv_transpose(dz[N,H,W,Cout], w); optional mask: zero where mask_src <= 0;
// optional colsum[Cin] += sum over pixels of dx (the previous layer's bias gradient, fused).
void conv3x3_dgrad(const bf16* dz, const bf16* w, const bf16* mask_src, bf16* dx, float* colsum, int N,
                   int H, int W, int Cin, int Cout, int bn, cudaStream_t stream);
// dw[Cout][3][3][Cin] (fp32) += scale * sum_pixels dz * shifted x.
void conv3x3_wgrad(const bf16* dz, const bf16* x, float* dw, int N, int H, int W, int Cin,
                   int Cout, float scale, int ksplit, int bn, cudaStream_t stream);

// Hardware probe for unaligned / strided SWIZZLE_128B operand views (umma_probe.cu).
void umma_shift_probe(const bf16* A, int rows, const bf16* B, float* out, int shift, int group_pitch,
                      int use_base_offset, int mode, cudaStream_t stream);

// ---- element-wise / reduction kernels (elementwise.cu) ----------------------------------------
void maxpool2x2_fwd(const bf16* x, bf16* y, int N, int H, int W, int C, cudaStream_t s);
// dz[N,H,W,C] = (y == pooled(y) first match && y > 0) ? dp[N,H/2,W/2,C] : 0
// optional colsum[C] += sum over pixels of dz (the layer's bias gradient, fused).
void maxpool2x2_relu_bwd(const bf16* y, const bf16* dp, bf16* dz, float* colsum, int N, int H, int W,
                         int C, cudaStream_t s);
// EXPERIMENTAL: backward of conv3x3_fprop_pool's pooling: dz[N,H,W,C] gets dp[N,H/2,W/2,C] at the
// recorded argmax where the maximum was positive, zero elsewhere; optional colsum as above.
void unpool2x2(const bf16* dp, const uint32_t* mask, bf16* dz, float* colsum, int N, int H, int W, int C,
               cudaStream_t s);
void adaptive_avgpool_fwd(const bf16* x, bf16* y, int N, int H, int W, int C, int OH, int OW,
                          cudaStream_t s);
void adaptive_avgpool_bwd(const bf16* dy, bf16* dx, int N, int H, int W, int C, int OH, int OW,
                          cudaStream_t s);
// db[c] += scale * sum_rows dz[row][c]      (rows = N*H*W for conv, batch for FC)
void bias_grad(const bf16* dz, float* db, long long rows, int C, float scale, cudaStream_t s);
// same for C % 8 != 0 (row stride ld), e.g. the 3-class logits gradient
void bias_grad_ld(const bf16* dz, float* db, long long rows, int C, int ld, float scale,
                  cudaStream_t s);
// FC epilogue: y = dropout(relu(acc + bias)) -> bf16; acc (fp32 [B][N]) is cleared for the next use.
void fc_bias_act(float* acc, const float* bias, bf16* y, float* y_f32, int B, int N, bool relu,
                 float drop_p, unsigned long long seed, unsigned long long offset, bool clear,
                 cudaStream_t s);
// FC backward epilogue: dz = acc * (act > 0) * dropmask/(1-p) -> bf16; acc cleared.
void fc_grad_act(float* acc, const bf16* act, bf16* dz, int B, int N, bool relu, float drop_p,
                 unsigned long long seed, unsigned long long offset, bool clear, cudaStream_t s);
// co-residency probe (bench/interference.py): thin background kernel exercising one SM resource
void probe_background(float* buf, long long n, int mode, int ctas, int reps, float* sink, cudaStream_t s);
void cast_f32_to_bf16(const float* x, bf16* y, long long n, cudaStream_t s);
void cast_bf16_to_f32(const bf16* x, float* y, long long n, cudaStream_t s);

// Fused cross-entropy forward+backward+metrics over fp32 logits [B][C].
// dlogits = (softmax - onehot) * grad_scale (bf16, ld = C rounded up to `ldd`);
// meter += {sum loss, #correct, B}; loss_out (optional) = mean loss of this batch.
void cross_entropy_fused(const float* logits, const long long* target, bf16* dlogits, int ldd,
                         float* meter, float* loss_out, int B, int C, float grad_scale,
                         const float* class_weights, cudaStream_t s);

// K-FUN2+CE: last Linear(K -> C) + cross-entropy + metrics + the layer's whole backward in one launch
// (C <= 8, K <= 1024, B <= 256).  h [B][K] bf16 (the previous FC layer's output), W [C][K] bf16.
// Writes logits [B][C] fp32, meter / loss_out as cross_entropy_fused, optional dlogits (bf16, ld ldd);
// with dW != nullptr also dW [C][K] (fp32, stored), db [C] (fp32, accumulated) and
// dh [B][K] = (h > 0 or !relu) * drop_scale * dlogits W  (bf16: the previous layer's dz).
bool head_ce_supported(int B, int C, int K);
void head_ce_fused(const bf16* h, const bf16* W, const float* bias, const long long* target, float* logits,
                   bf16* dlogits, int ldd, float* dW, float* db, bf16* dh, float drop_scale, bool relu, float* meter,
                   float* loss_out, int B, int C, int K, const float* class_weights, cudaStream_t s);

// Fused optimizers over flat arenas.  grad is fp32 (local) or bf16 (the reduced wire buffer).
void adam_fused(float* p, float* m, float* v, const float* g32, const bf16* g16, bf16* shadow,
                long long n, float lr, float beta1, float beta2, float eps, float weight_decay,
                int step, float grad_scale, bool zero_grad, float* g32_to_zero, cudaStream_t s);
void sgd_fused(float* p, float* mom, const float* g32, const bf16* g16, bf16* shadow, long long n,
               float lr, float momentum, float weight_decay, bool first_step, float grad_scale,
               float* g32_to_zero, cudaStream_t s);

// Fused input transform: uint8 HWC source -> normalised bf16.  mode 0: NHWC with C padded to
// `cpad`; mode 1: layer-0 im2col matrix [N*OH*OW][kpad] (k = (kh*3+kw)*3 + c, zero padded).
void augment_fused(const uint8_t* src, const float* params, bf16* out, int N, int SH, int SW,
                   int RH, int RW, int OH, int OW, int mode, int pad, const float* mean,
                   const float* stdv, cudaStream_t s);
// im2col of an NHWC(3, padded to cpad) bf16 image for layer 0 (used when the input is given as a
// float tensor instead of uint8 + params).
void im2col3x3_c3(const bf16* x, bf16* out, int N, int H, int W, int cpad, int kpad,
                  cudaStream_t s);
void nchw_f32_to_nhwc_bf16(const float* x, bf16* y, int N, int C, int H, int W, int cpad,
                           cudaStream_t s);

// ---- cross-GPU (allreduce.cu) ------------------------------------------------------------------
struct CommCtx {
  int rank, world;
  void* const* wire_ptrs;        // device array [world]: peers' wire buffers (bf16 or fp32)
  uint32_t* const* signal_ptrs;  // device array [world]: peers' signal pads
  void* wire_mc;                 // multicast alias of the wire buffer (NVLS) or nullptr
  uint32_t* signal_mc;           // multicast alias of the signal pad or nullptr
  unsigned long long timeout_ns; // wall-clock bound of every cross-GPU wait (B200_BARRIER_TIMEOUT_S)
};
enum AllreduceAlgo : int { AR_ONESHOT = 0, AR_TWOSHOT = 1, AR_NVLS = 2 };
// Fused gradient all-reduce of arena range [start, start+n):
//   pack:   wire[i] = bf16(grad_f32[i] * inv_world)            (skipped when grad_f32 == nullptr)
//   reduce: every rank ends with wire[i] = sum over ranks (fp32 accumulation, bf16 on the wire)
//   unpack: optional fp32 write-back of the averaged gradient into grad_out_f32.
// `slot` selects one of the signal-pad barrier groups; `epoch` is incremented by the caller.
void allreduce_fused(const CommCtx& ctx, const float* grad_f32, float* grad_out_f32,
                     long long start, long long n, float inv_world, int algo, bool wire_fp32,
                     int slot, uint32_t epoch, int max_ctas, cudaStream_t s);
// CTAs allreduce_fused / zero1_step use for a range of n elements (the cell decomposition follows).
int allreduce_grid(long long n, int world, int max_ctas, bool wire_fp32);
// EXPERIMENTAL: ZeRO-1 step of arena range [start, start+n) in one kernel -- reduce-scatter, Adam on
// the cells this rank owns, all-gather of the new bf16 weights into every rank's shadow (allreduce.cu).
void zero1_step(const CommCtx& ctx, float* grad_f32, float* p, float* m, float* v, bf16* shadow,
                long long start, long long n, float inv_world, int algo, int slot, uint32_t epoch, int max_ctas,
                float lr, float beta1, float beta2, float eps, float weight_decay, int step, cudaStream_t s);
// Parameter broadcast rank `root` -> all through the wire buffer (K-BCAST).
void broadcast_fused(const CommCtx& ctx, float* data_f32, long long n, int root, int slot,
                     uint32_t epoch, cudaStream_t s);
void device_barrier(const CommCtx& ctx, int slot, uint32_t epoch, cudaStream_t s);
// uint32 words of signal pad needed for `slots` barrier groups
int allreduce_signal_words(int slots);

}  // namespace b200
