// Fused cross-entropy (forward + backward + metrics) and the fused flat-arena optimizers.
//
// Reference semantics: F.cross_entropy(outputs, targets) with mean reduction
// (distributedVggf.py:168), Accuracy2 = argmax == target count (distributedUtil.py:92-99),
// Average = sample-weighted loss mean (distributedUtil.py:55-58), torch.optim.Adam with default
// betas/eps (distributedVggf.py:230).  The reference pays two host syncs per step for the metrics;
// here they are three atomics into a device meter.
#include <cstdlib>
#include <stdexcept>

#include "api.h"
#include "ptx.cuh"

namespace b200 {

// One block, one warp per row (round-robin).  C is small (3) or moderate (1000).
__global__ void cross_entropy_kernel(const float* __restrict__ logits, const long long* __restrict__ target,
                                     bf16* __restrict__ dlogits, int ldd, float* __restrict__ meter,
                                     float* __restrict__ loss_out, int B, int C, float grad_scale,
                                     const float* __restrict__ cw) {
  __shared__ float s_loss[32], s_correct[32], s_wsum[32];
  __shared__ float s_total_w;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;

  // pass 0 (only with class weights): normaliser = sum_i w[target_i]
  float wsum_local = 0.f;
  if (cw) {
    for (int r = threadIdx.x; r < B; r += blockDim.x) wsum_local += cw[target[r]];
    for (int o = 16; o; o >>= 1) wsum_local += __shfl_xor_sync(0xffffffffu, wsum_local, o);
    if (lane == 0) s_wsum[warp] = wsum_local;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < nwarps; ++i) t += s_wsum[i];
      s_total_w = t;
    }
    __syncthreads();
  }
  const float norm = cw ? 1.f / s_total_w : grad_scale;   // grad_scale == 1/B for the mean

  float loss_acc = 0.f, correct_acc = 0.f;
  for (int r = warp; r < B; r += nwarps) {
    const float* row = logits + static_cast<long long>(r) * C;
    const int t = static_cast<int>(target[r]);
    float mx = -INFINITY;
    int amax = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
      const float v = row[c];
      if (v > mx) { mx = v; amax = c; }
    }
    for (int o = 16; o; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oa = __shfl_xor_sync(0xffffffffu, amax, o);
      if (om > mx || (om == mx && oa < amax)) { mx = om; amax = oa; }
    }
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += __expf(row[c] - mx);
    for (int o = 16; o; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    const float lse = mx + __logf(se);
    const float w = cw ? cw[t] : 1.f;
    if (dlogits) {
      for (int c = lane; c < ldd; c += 32) {
        float g = 0.f;
        if (c < C) g = (__expf(row[c] - lse) - (c == t ? 1.f : 0.f)) * w * norm;
        dlogits[static_cast<long long>(r) * ldd + c] = __float2bfloat16(g);
      }
    }
    if (lane == 0) {
      loss_acc += w * (lse - row[t]);
      correct_acc += (amax == t) ? 1.f : 0.f;
    }
  }
  if (lane == 0) { s_loss[warp] = loss_acc; s_correct[warp] = correct_acc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tl = 0.f, tc = 0.f;
    for (int i = 0; i < nwarps; ++i) { tl += s_loss[i]; tc += s_correct[i]; }
    // batch loss as torch reports it: mean over the batch, or sum_i w_i nll_i / sum_i w_i with
    // class weights; the meter accumulates loss * B like Average.update(loss.item(), B)
    const float mean = cw ? tl / s_total_w : tl / static_cast<float>(B);
    if (meter) {
      atomicAdd(meter + 0, mean * static_cast<float>(B));
      atomicAdd(meter + 1, tc);
      atomicAdd(meter + 2, static_cast<float>(B));
    }
    if (loss_out) *loss_out = mean;
  }
}

void cross_entropy_fused(const float* logits, const long long* target, bf16* dlogits, int ldd,
                         float* meter, float* loss_out, int B, int C, float grad_scale,
                         const float* class_weights, cudaStream_t s) {
  cross_entropy_kernel<<<1, 1024, 0, s>>>(logits, target, dlogits, ldd, meter, loss_out, B, C,
                                          grad_scale, class_weights);
  count_launch();
  check_last("cross_entropy_fused");
}

// ------------------------------------------------------------- fused head: last Linear + CE + bwd
// K-FUN2+CE: the funnel's Linear(512, C) (distributedVggf.py:56), cross_entropy (:168),
// Accuracy2 / Average (distributedUtil.py:95-96, :55-58) and the whole backward of that layer in ONE
// launch of one CTA -- seven launches before (GEMM, bias epilogue, CE, bias grad, wgrad GEMM, dgrad
// GEMM, ReLU/dropout-mask epilogue).  C <= 8 classes, K <= 1024 inputs, B <= 256 rows; everything lives
// in shared memory / registers:
//   logits[r][c] = bias[c] + sum_k h[r][k] W[c][k]            (warp per row, fp32 accumulation)
//   loss, argmax == target count, B  -> device meter;  dlogits = (softmax - onehot) * w / norm -> bf16
//   db[c] += sum_r dl[r][c];  dW[c][k] = sum_r dl[r][c] h[r][k];                (thread per k)
//   dh[r][k] = (h[r][k] > 0) * drop_scale * sum_c dl[r][c] W[c][k]  -> bf16     (previous layer's dz)
// Rounding points are those of the unfused path: dlogits are rounded to bf16 before they are used.
constexpr int HEAD_MAXC = 8, HEAD_MAXK = 1024, HEAD_MAXB = 256, HEAD_THREADS = 512;

__global__ void __launch_bounds__(HEAD_THREADS, 1)
head_ce_kernel(const bf16* __restrict__ h, const bf16* __restrict__ W, const float* __restrict__ bias,
               const long long* __restrict__ target, float* __restrict__ logits, bf16* __restrict__ dlogits, int ldd,
               float* __restrict__ dW, float* __restrict__ db, bf16* __restrict__ dh, float drop_scale, int relu,
               float* __restrict__ meter, float* __restrict__ loss_out, int B, int C, int K,
               const float* __restrict__ cw) {
  __shared__ float sW[HEAD_MAXC * HEAD_MAXK];
  __shared__ float sdl[HEAD_MAXB * HEAD_MAXC];
  __shared__ float s_loss[32], s_correct[32], s_wsum[32];
  __shared__ float s_total_w;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < C * K; i += blockDim.x) sW[i] = __bfloat162float(W[i]);

  float wsum_local = 0.f;
  if (cw) {      // class weights: normaliser = sum_i w[target_i]   (same as cross_entropy_kernel)
    for (int r = threadIdx.x; r < B; r += blockDim.x) wsum_local += cw[target[r]];
    for (int o = 16; o; o >>= 1) wsum_local += __shfl_xor_sync(0xffffffffu, wsum_local, o);
    if (lane == 0) s_wsum[warp] = wsum_local;
  }
  __syncthreads();
  if (cw && threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < nwarps; ++i) t += s_wsum[i];
    s_total_w = t;
  }
  __syncthreads();
  const float norm = cw ? 1.f / s_total_w : 1.f / static_cast<float>(B);

  // ---- forward + loss + dlogits ---------------------------------------------------------------
  float loss_acc = 0.f, correct_acc = 0.f;
  for (int r = warp; r < B; r += nwarps) {
    float acc[HEAD_MAXC];
#pragma unroll
    for (int c = 0; c < HEAD_MAXC; ++c) acc[c] = 0.f;
    const bf16* hr = h + static_cast<long long>(r) * K;
    for (int k = lane; k < K; k += 32) {
      const float hv = __bfloat162float(hr[k]);
#pragma unroll
      for (int c = 0; c < HEAD_MAXC; ++c)
        if (c < C) acc[c] = fmaf(hv, sW[c * K + k], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < HEAD_MAXC; ++c)
      for (int o = 16; o; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    const int t = static_cast<int>(target[r]);
    float mx = -INFINITY;
    int amax = 0;
#pragma unroll
    for (int c = 0; c < HEAD_MAXC; ++c)
      if (c < C) {
        acc[c] += bias[c];
        if (acc[c] > mx) { mx = acc[c]; amax = c; }       // first maximum, like torch.argmax
      }
    float se = 0.f, zt = 0.f;
#pragma unroll
    for (int c = 0; c < HEAD_MAXC; ++c)
      if (c < C) {
        se += __expf(acc[c] - mx);
        if (c == t) zt = acc[c];
      }
    const float lse = mx + __logf(se);
    const float w = cw ? cw[t] : 1.f;
    if (lane < ldd || lane < C) {
      float g = 0.f, z = 0.f;
#pragma unroll
      for (int c = 0; c < HEAD_MAXC; ++c)
        if (c == lane && c < C) {
          z = acc[c];
          g = (__expf(acc[c] - lse) - (c == t ? 1.f : 0.f)) * w * norm;
        }
      if (lane < C) {
        logits[static_cast<long long>(r) * C + lane] = z;
        sdl[r * HEAD_MAXC + lane] = __bfloat162float(__float2bfloat16(g));
      }
      if (dlogits && lane < ldd) dlogits[static_cast<long long>(r) * ldd + lane] = __float2bfloat16(g);
    }
    if (lane == 0) {
      loss_acc += w * (lse - zt);
      correct_acc += (amax == t) ? 1.f : 0.f;
    }
  }
  if (lane == 0) { s_loss[warp] = loss_acc; s_correct[warp] = correct_acc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tl = 0.f, tc = 0.f;
    for (int i = 0; i < nwarps; ++i) { tl += s_loss[i]; tc += s_correct[i]; }
    const float mean = cw ? tl / s_total_w : tl / static_cast<float>(B);
    if (meter) {
      atomicAdd(meter + 0, mean * static_cast<float>(B));
      atomicAdd(meter + 1, tc);
      atomicAdd(meter + 2, static_cast<float>(B));
    }
    if (loss_out) *loss_out = mean;
  }
  if (!dW) return;                       // evaluation: forward + metrics only

  // ---- backward of the layer --------------------------------------------------------------------
  if (threadIdx.x < C) {
    float sum = 0.f;
    for (int r = 0; r < B; ++r) sum += sdl[r * HEAD_MAXC + threadIdx.x];
    db[threadIdx.x] += sum;
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float wk[HEAD_MAXC], gw[HEAD_MAXC];
#pragma unroll
    for (int c = 0; c < HEAD_MAXC; ++c) { wk[c] = c < C ? sW[c * K + k] : 0.f; gw[c] = 0.f; }
    for (int r = 0; r < B; ++r) {
      const float hv = __bfloat162float(h[static_cast<long long>(r) * K + k]);
      float dx = 0.f;
#pragma unroll
      for (int c = 0; c < HEAD_MAXC; ++c)
        if (c < C) {
          const float g = sdl[r * HEAD_MAXC + c];
          gw[c] = fmaf(g, hv, gw[c]);
          dx = fmaf(g, wk[c], dx);
        }
      if (dh) dh[static_cast<long long>(r) * K + k] = __float2bfloat16((!relu || hv > 0.f) ? dx * drop_scale : 0.f);
    }
#pragma unroll
    for (int c = 0; c < HEAD_MAXC; ++c)
      if (c < C) dW[static_cast<long long>(c) * K + k] = gw[c];
  }
}

bool head_ce_supported(int B, int C, int K) { return C <= HEAD_MAXC && K <= HEAD_MAXK && B <= HEAD_MAXB && B > 0; }

void head_ce_fused(const bf16* h, const bf16* W, const float* bias, const long long* target, float* logits,
                   bf16* dlogits, int ldd, float* dW, float* db, bf16* dh, float drop_scale, bool relu, float* meter,
                   float* loss_out, int B, int C, int K, const float* class_weights, cudaStream_t s) {
  if (!head_ce_supported(B, C, K)) throw std::runtime_error("[b200] head_ce_fused: needs C <= 8, K <= 1024, B <= 256");
  if (ldd > 32) throw std::runtime_error("[b200] head_ce_fused: dlogits row stride must be <= 32");
  head_ce_kernel<<<1, HEAD_THREADS, 0, s>>>(h, W, bias, target, logits, dlogits, ldd, dW, db, dh, drop_scale,
                                            relu ? 1 : 0, meter, loss_out, B, C, K, class_weights);
  count_launch();
  check_last("head_ce_fused");
}

// ---------------------------------------------------------------------------------------- Adam
// One pass over the flat arena: read grad (fp32 local or bf16 reduced wire), update fp32 master
// weight and both moments, emit the bf16 shadow used by the GEMMs, and zero the fp32 gradient so
// the red.add wgrad epilogues of the next step start from zero.
// Footprint (same reasoning as allreduce.cu): the update runs on the comm stream UNDER the persistent
// conv kernels of backward.  256 threads x <= 40 registers and no shared memory co-reside with any of
// them, and the grid is one CTA per SM -- a background stream of HBM traffic.  (Round 1 launched 8
// CTAs per SM: at every conv-kernel boundary the high-priority stream filled the SMs with optimizer
// CTAs and the next conv kernel's persistent CTAs started late; bench/step_timeline.py.)
__global__ void __launch_bounds__(256, 6) adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                            const float* __restrict__ g32, const bf16* __restrict__ g16,
                            bf16* __restrict__ shadow, long long n4, float lr, float b1, float b2,
                            float eps, float wd, float bc1_inv, float bc2_inv_sqrt, float gscale,
                            float* __restrict__ gzero) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    // Every stream here is touched exactly once per step: cache-streaming (evict-first) loads and
    // stores, so that ~4 GB of optimizer traffic does not flush the conv kernels' L2-resident operand
    // tiles while it runs under backward (bench/step_timeline.py: +30 % on the 512-channel layers).
    float4 pv = __ldcs(reinterpret_cast<const float4*>(p) + i);
    float4 mv = __ldcs(reinterpret_cast<const float4*>(m) + i);
    float4 vv = __ldcs(reinterpret_cast<const float4*>(v) + i);
    float g[4];
    if (g16) {
      const uint2 raw = __ldcs(reinterpret_cast<const uint2*>(g16 + i * 4));
      const float2 a = unpack_bf16x2(raw.x), b = unpack_bf16x2(raw.y);
      g[0] = a.x; g[1] = a.y; g[2] = b.x; g[3] = b.y;
    } else {
      const float4 gv = __ldcs(reinterpret_cast<const float4*>(g32) + i);
      g[0] = gv.x; g[1] = gv.y; g[2] = gv.z; g[3] = gv.w;
    }
    float pp[4] = {pv.x, pv.y, pv.z, pv.w};
    float mm[4] = {mv.x, mv.y, mv.z, mv.w};
    float vs[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gk = g[k] * gscale + wd * pp[k];
      mm[k] = b1 * mm[k] + (1.f - b1) * gk;
      vs[k] = b2 * vs[k] + (1.f - b2) * gk * gk;
      const float denom = sqrtf(vs[k]) * bc2_inv_sqrt + eps;
      pp[k] -= lr * bc1_inv * mm[k] / denom;
    }
    __stcs(reinterpret_cast<float4*>(p) + i, make_float4(pp[0], pp[1], pp[2], pp[3]));
    __stcs(reinterpret_cast<float4*>(m) + i, make_float4(mm[0], mm[1], mm[2], mm[3]));
    __stcs(reinterpret_cast<float4*>(v) + i, make_float4(vs[0], vs[1], vs[2], vs[3]));
    if (shadow)
      __stcs(reinterpret_cast<uint2*>(shadow + i * 4), make_uint2(pack_bf16x2(pp[0], pp[1]), pack_bf16x2(pp[2], pp[3])));
    if (gzero) __stcs(reinterpret_cast<float4*>(gzero) + i, make_float4(0.f, 0.f, 0.f, 0.f));
  }
}

void adam_fused(float* p, float* m, float* v, const float* g32, const bf16* g16, bf16* shadow,
                long long n, float lr, float beta1, float beta2, float eps, float weight_decay,
                int step, float grad_scale, bool, float* g32_to_zero, cudaStream_t s) {
  if (n % 4) throw std::runtime_error("[b200] adam_fused: n must be a multiple of 4");
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  const long long n4 = n / 4;
  // Tunable without a rebuild: B200_ADAM_CTAS_PER_SM (default 4).
  static const int per_sm = [] {
    const char* e = getenv("B200_ADAM_CTAS_PER_SM");
    const int v = e ? atoi(e) : 4;
    return v >= 1 && v <= 16 ? v : 4;
  }();
  static const bool carve = (prefer_max_shared_carveout(reinterpret_cast<const void*>(adam_kernel)), true);
  (void)carve;
  const long long cap = static_cast<long long>(sm_count()) * per_sm;
  const int blocks = static_cast<int>(n4 / 256 + 1 < cap ? n4 / 256 + 1 : cap);
  adam_kernel<<<blocks, 256, 0, s>>>(p, m, v, g32, g16, shadow, n4, lr, beta1, beta2, eps, weight_decay,
                                     static_cast<float>(1.0 / bc1), static_cast<float>(1.0 / sqrt(bc2)),
                                     grad_scale, g32_to_zero);
  count_launch();
  check_last("adam_fused");
}

// --------------------------------------------------------------------------------------- SGD
// torch.optim.SGD(momentum) semantics: buf = g (first step) | momentum*buf + g ; p -= lr*buf.
__global__ void __launch_bounds__(256, 6) sgd_kernel(float* __restrict__ p, float* __restrict__ mom, const float* __restrict__ g32,
                           const bf16* __restrict__ g16, bf16* __restrict__ shadow, long long n4, float lr,
                           float momentum, float wd, int first, float gscale, float* __restrict__ gzero) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 pv = __ldcs(reinterpret_cast<const float4*>(p) + i);       // streaming, see adam_kernel
    float4 bv = __ldcs(reinterpret_cast<const float4*>(mom) + i);
    float g[4];
    if (g16) {
      const uint2 raw = __ldcs(reinterpret_cast<const uint2*>(g16 + i * 4));
      const float2 a = unpack_bf16x2(raw.x), b = unpack_bf16x2(raw.y);
      g[0] = a.x; g[1] = a.y; g[2] = b.x; g[3] = b.y;
    } else {
      const float4 gv = __ldcs(reinterpret_cast<const float4*>(g32) + i);
      g[0] = gv.x; g[1] = gv.y; g[2] = gv.z; g[3] = gv.w;
    }
    float pp[4] = {pv.x, pv.y, pv.z, pv.w};
    float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = g[k] * gscale + wd * pp[k];
      bb[k] = first ? gk : momentum * bb[k] + gk;
      pp[k] -= lr * bb[k];
    }
    __stcs(reinterpret_cast<float4*>(p) + i, make_float4(pp[0], pp[1], pp[2], pp[3]));
    __stcs(reinterpret_cast<float4*>(mom) + i, make_float4(bb[0], bb[1], bb[2], bb[3]));
    if (shadow)
      __stcs(reinterpret_cast<uint2*>(shadow + i * 4), make_uint2(pack_bf16x2(pp[0], pp[1]), pack_bf16x2(pp[2], pp[3])));
    if (gzero) __stcs(reinterpret_cast<float4*>(gzero) + i, make_float4(0.f, 0.f, 0.f, 0.f));
  }
}

void sgd_fused(float* p, float* mom, const float* g32, const bf16* g16, bf16* shadow, long long n,
               float lr, float momentum, float weight_decay, bool first_step, float grad_scale,
               float* g32_to_zero, cudaStream_t s) {
  if (n % 4) throw std::runtime_error("[b200] sgd_fused: n must be a multiple of 4");
  const long long n4 = n / 4;
  static const bool carve = (prefer_max_shared_carveout(reinterpret_cast<const void*>(sgd_kernel)), true);
  (void)carve;
  const long long cap = static_cast<long long>(sm_count()) * 4;
  const int blocks = static_cast<int>(n4 / 256 + 1 < cap ? n4 / 256 + 1 : cap);
  sgd_kernel<<<blocks, 256, 0, s>>>(p, mom, g32, g16, shadow, n4, lr, momentum, weight_decay,
                                    first_step ? 1 : 0, grad_scale, g32_to_zero);
  count_launch();
  check_last("sgd_fused");
}

}  // namespace b200
