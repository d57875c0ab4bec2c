// Fused input transform (K-AUG): the reference's six PIL/torchvision stages
// (distributedVggf.py:88-95 / :103-108) evaluated as one coordinate chain per output pixel, reading
// the decoded uint8 HWC source and writing normalised bf16 either as NHWC (C padded) or directly
// as the layer-0 im2col matrix [N*OH*OW][kpad] (k = (kh*3+kw)*3 + c) that feeds the tcgen05 GEMM
// -- the 3-channel tensor never exists in HBM in that mode.
// The arithmetic mirrors data/transforms.py::augment_reference (the oracle in the tests).
#include <stdexcept>

#include "api.h"
#include "ptx.cuh"

namespace b200 {

struct AugGeom {
  int N, SH, SW, RH, RW, OH, OW;
  float mean[3], inv_std[3];
};

// Normalised value of output pixel (oy, ox) of sample n, all 3 channels.
__device__ __forceinline__ void aug_sample(const uint8_t* __restrict__ src, const float* __restrict__ prm,
                                           const AugGeom& g, int oy, int ox, float (&out)[3]) {
  const float top = prm[0], left = prm[1], ch = prm[2], cw = prm[3], cs = prm[4], sn = prm[5];
  const bool flip = prm[6] > 0.5f;
  const int off_y = __float2int_rn((g.RH - g.OH) * 0.5f);
  const int off_x = __float2int_rn((g.RW - g.OW) * 0.5f);
  const float ay = static_cast<float>(oy + off_y);
  float ax = static_cast<float>(ox + off_x);
  if (flip) ax = static_cast<float>(g.RW - 1) - ax;
  const float cx = g.RW * 0.5f, cy = g.RH * 0.5f;
  const float dx = ax + 0.5f - cx, dy = ay + 0.5f - cy;
  const float rx = floorf(cs * dx - sn * dy + cx);
  const float ry = floorf(sn * dx + cs * dy + cy);
  if (rx < 0.f || rx >= g.RW || ry < 0.f || ry >= g.RH) {   // rotation fill = 0 (pre-normalise)
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c] = (0.f - g.mean[c]) * g.inv_std[c];
    return;
  }
  float sy = (ry + 0.5f) * (ch / g.RH) - 0.5f;
  float sx = (rx + 0.5f) * (cw / g.RW) - 0.5f;
  sy = fminf(fmaxf(sy, 0.f), ch - 1.f);
  sx = fminf(fmaxf(sx, 0.f), cw - 1.f);
  const float y0 = floorf(sy), x0 = floorf(sx);
  const float wy = sy - y0, wx = sx - x0;
  const float y1 = fminf(y0 + 1.f, ch - 1.f), x1 = fminf(x0 + 1.f, cw - 1.f);
  const int iy0 = min(max(static_cast<int>(y0 + top), 0), g.SH - 1);
  const int iy1 = min(max(static_cast<int>(y1 + top), 0), g.SH - 1);
  const int ix0 = min(max(static_cast<int>(x0 + left), 0), g.SW - 1);
  const int ix1 = min(max(static_cast<int>(x1 + left), 0), g.SW - 1);
  const uint8_t* p00 = src + (static_cast<long long>(iy0) * g.SW + ix0) * 3;
  const uint8_t* p01 = src + (static_cast<long long>(iy0) * g.SW + ix1) * 3;
  const uint8_t* p10 = src + (static_cast<long long>(iy1) * g.SW + ix0) * 3;
  const uint8_t* p11 = src + (static_cast<long long>(iy1) * g.SW + ix1) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = p00[c] * (1.f - wy) * (1.f - wx) + p01[c] * (1.f - wy) * wx +
                    p10[c] * wy * (1.f - wx) + p11[c] * wy * wx;
    out[c] = (v * (1.f / 255.f) - g.mean[c]) * g.inv_std[c];
  }
}

// mode 0: NHWC bf16, C padded to `pad` (pad in {4, 8}); one thread per output pixel.
__global__ void augment_nhwc_kernel(const uint8_t* __restrict__ src, const float* __restrict__ params,
                                    bf16* __restrict__ out, AugGeom g, int pad) {
  const long long total = static_cast<long long>(g.N) * g.OH * g.OW;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ox = i % g.OW;
    const int oy = (i / g.OW) % g.OH;
    const int n = static_cast<int>(i / (static_cast<long long>(g.OW) * g.OH));
    float v[3];
    aug_sample(src + static_cast<long long>(n) * g.SH * g.SW * 3, params + n * 8, g, oy, ox, v);
    bf16* o = out + i * pad;
    o[0] = __float2bfloat16(v[0]);
    o[1] = __float2bfloat16(v[1]);
    o[2] = __float2bfloat16(v[2]);
    for (int c = 3; c < pad; ++c) o[c] = __float2bfloat16(0.f);
  }
}

// mode 1: im2col rows for the 3x3/pad-1 first convolution.  One thread per (pixel, tap): it
// evaluates the transform at the tap's neighbour (or 0 outside the image: conv zero padding acts
// on the *normalised* tensor) and writes 3 bf16; tap 9 writes the zero tail [27, kpad).
__global__ void augment_im2col_kernel(const uint8_t* __restrict__ src, const float* __restrict__ params,
                                      bf16* __restrict__ out, AugGeom g, int kpad) {
  const long long total = static_cast<long long>(g.N) * g.OH * g.OW * 10;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int tap = i % 10;
    const long long pix = i / 10;
    bf16* row = out + pix * kpad;
    if (tap == 9) {
      for (int k = 27; k < kpad; ++k) row[k] = __float2bfloat16(0.f);
      continue;
    }
    const int ox = pix % g.OW;
    const int oy = (pix / g.OW) % g.OH;
    const int n = static_cast<int>(pix / (static_cast<long long>(g.OW) * g.OH));
    const int yy = oy + tap / 3 - 1, xx = ox + tap % 3 - 1;
    float v[3] = {0.f, 0.f, 0.f};
    if (yy >= 0 && yy < g.OH && xx >= 0 && xx < g.OW)
      aug_sample(src + static_cast<long long>(n) * g.SH * g.SW * 3, params + n * 8, g, yy, xx, v);
    row[tap * 3 + 0] = __float2bfloat16(v[0]);
    row[tap * 3 + 1] = __float2bfloat16(v[1]);
    row[tap * 3 + 2] = __float2bfloat16(v[2]);
  }
}

void augment_fused(const uint8_t* src, const float* params, bf16* out, int N, int SH, int SW, int RH,
                   int RW, int OH, int OW, int mode, int pad, const float* mean, const float* stdv,
                   cudaStream_t s) {
  AugGeom g;
  g.N = N; g.SH = SH; g.SW = SW; g.RH = RH; g.RW = RW; g.OH = OH; g.OW = OW;
  for (int c = 0; c < 3; ++c) { g.mean[c] = mean[c]; g.inv_std[c] = 1.f / stdv[c]; }
  const long long pixels = static_cast<long long>(N) * OH * OW;
  if (mode == 0) {
    const int blocks = static_cast<int>((pixels + 255) / 256 < 148 * 16 ? (pixels + 255) / 256 : 148 * 16);
    augment_nhwc_kernel<<<blocks, 256, 0, s>>>(src, params, out, g, pad);
  } else {
    if (pad < 27) throw std::runtime_error("[b200] augment_fused: kpad must be >= 27");
    const long long total = pixels * 10;
    const int blocks = static_cast<int>((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    augment_im2col_kernel<<<blocks, 256, 0, s>>>(src, params, out, g, pad);
  }
  count_launch();
  check_last("augment_fused");
}

// im2col of an already-normalised NHWC (C padded to cpad) bf16 image -- the path used when the
// caller supplies float tensors (reference pipeline, tests) instead of uint8 + parameters.
__global__ void im2col3x3_c3_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int N, int H, int W,
                                    int cpad, int kpad) {
  const long long total = static_cast<long long>(N) * H * W * 10;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int tap = i % 10;
    const long long pix = i / 10;
    bf16* row = out + pix * kpad;
    if (tap == 9) {
      for (int k = 27; k < kpad; ++k) row[k] = __float2bfloat16(0.f);
      continue;
    }
    const int w = pix % W;
    const int h = (pix / W) % H;
    const int n = static_cast<int>(pix / (static_cast<long long>(W) * H));
    const int yy = h + tap / 3 - 1, xx = w + tap % 3 - 1;
    const bf16 z = __float2bfloat16(0.f);
    bf16 v0 = z, v1 = z, v2 = z;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      const bf16* p = x + ((static_cast<long long>(n) * H + yy) * W + xx) * cpad;
      v0 = p[0]; v1 = p[1]; v2 = p[2];
    }
    row[tap * 3 + 0] = v0; row[tap * 3 + 1] = v1; row[tap * 3 + 2] = v2;
  }
}

void im2col3x3_c3(const bf16* x, bf16* out, int N, int H, int W, int cpad, int kpad, cudaStream_t s) {
  const long long total = static_cast<long long>(N) * H * W * 10;
  const int blocks = static_cast<int>((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
  im2col3x3_c3_kernel<<<blocks, 256, 0, s>>>(x, out, N, H, W, cpad, kpad);
  count_launch();
  check_last("im2col3x3_c3");
}

__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, int N, int C,
                                             int H, int W, int cpad) {
  const long long total = static_cast<long long>(N) * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long hw = i % (static_cast<long long>(H) * W);
    const long long n = i / (static_cast<long long>(H) * W);
    for (int c = 0; c < cpad; ++c) {
      const float v = c < C ? x[(n * C + c) * H * W + hw] : 0.f;
      y[i * cpad + c] = __float2bfloat16(v);
    }
  }
}

void nchw_f32_to_nhwc_bf16(const float* x, bf16* y, int N, int C, int H, int W, int cpad, cudaStream_t s) {
  const long long total = static_cast<long long>(N) * H * W;
  const int blocks = static_cast<int>((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  nchw_f32_to_nhwc_bf16_kernel<<<blocks, 256, 0, s>>>(x, y, N, C, H, W, cpad);
  count_launch();
  check_last("nchw_f32_to_nhwc_bf16");
}

}  // namespace b200
