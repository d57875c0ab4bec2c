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
    if (pad == 4) {          // one 8-byte store per pixel
      *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], 0.f));
    } else {
      o[0] = __float2bfloat16(v[0]);
      o[1] = __float2bfloat16(v[1]);
      o[2] = __float2bfloat16(v[2]);
      for (int c = 3; c < pad; ++c) o[c] = __float2bfloat16(0.f);
    }
  }
}

// mode 1: im2col rows for the 3x3/pad-1 first convolution.  One thread per (pixel, 8-element
// chunk of the row): it evaluates the transform at the (up to 4) taps its chunk covers (0 outside
// the image: conv zero padding acts on the *normalised* tensor) and issues ONE 16-byte store, so a
// warp writes 512 contiguous bytes.  Chunks beyond k = 27 are the zero tail.
__global__ void augment_im2col_kernel(const uint8_t* __restrict__ src, const float* __restrict__ params,
                                      bf16* __restrict__ out, AugGeom g, int kpad) {
  const int chunks = kpad / 8;
  const long long total = static_cast<long long>(g.N) * g.OH * g.OW * chunks;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int chunk = static_cast<int>(i % chunks);
    const long long pix = i / chunks;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int k0 = chunk * 8;
    if (k0 < 27) {
      const int ox = pix % g.OW;
      const int oy = (pix / g.OW) % g.OH;
      const int n = static_cast<int>(pix / (static_cast<long long>(g.OW) * g.OH));
      const uint8_t* img = src + static_cast<long long>(n) * g.SH * g.SW * 3;
      const float* prm = params + n * 8;
      const int tap_lo = k0 / 3;
      const int tap_hi = min((k0 + 7) / 3, 8);
      for (int tap = tap_lo; tap <= tap_hi; ++tap) {
        const int yy = oy + tap / 3 - 1, xx = ox + tap % 3 - 1;
        float s3[3] = {0.f, 0.f, 0.f};
        if (yy >= 0 && yy < g.OH && xx >= 0 && xx < g.OW) aug_sample(img, prm, g, yy, xx, s3);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int k = tap * 3 + c - k0;
          if (k >= 0 && k < 8) v[k] = s3[c];
        }
      }
    }
    const uint4 pk = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    *reinterpret_cast<uint4*>(out + pix * kpad + k0) = pk;
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
    const int blocks = static_cast<int>((pixels + 255) / 256 < sm_count() * 16 ? (pixels + 255) / 256 : sm_count() * 16);
    augment_nhwc_kernel<<<blocks, 256, 0, s>>>(src, params, out, g, pad);
  } else {
    if (pad < 32 || pad % 8) throw std::runtime_error("[b200] augment_fused: kpad must be a multiple of 8, >= 32");
    const long long total = pixels * (pad / 8);
    const int blocks = static_cast<int>((total + 255) / 256 < sm_count() * 32 ? (total + 255) / 256 : sm_count() * 32);
    augment_im2col_kernel<<<blocks, 256, 0, s>>>(src, params, out, g, pad);
  }
  count_launch();
  check_last("augment_fused");
}

// im2col of an already-normalised NHWC (3 channels padded to cpad = 4) bf16 image: the second half
// of the two-pass input path (augment -> 8-byte pixels, 25 MB, stays in L2 -> im2col rows).  One
// thread per (pixel, 8-element chunk): up to four 8-byte pixel loads, one 16-byte store.
__global__ void im2col3x3_c3_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int N, int H, int W,
                                    int cpad, int kpad) {
  const int chunks = kpad / 8;
  const long long total = static_cast<long long>(N) * H * W * chunks;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int chunk = static_cast<int>(i % chunks);
    const long long pix = i / chunks;
    const int k0 = chunk * 8;
    uint32_t h[8];                                   // bf16 bit patterns, compile-time indexed only
    if (k0 < 27) {
      const int w = pix % W;
      const int hh = (pix / W) % H;
      const long long img = (pix / (static_cast<long long>(W) * H)) * H * W;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = k0 + e;
        const int tap = k / 3, c = k - 3 * tap;
        const int yy = hh + tap / 3 - 1, xx = w + tap % 3 - 1;
        uint2 px = make_uint2(0u, 0u);
        if (k < 27 && yy >= 0 && yy < H && xx >= 0 && xx < W)
          px = __ldg(reinterpret_cast<const uint2*>(x + (img + static_cast<long long>(yy) * W + xx) * 4));
        h[e] = c == 0 ? (px.x & 0xffffu) : (c == 1 ? (px.x >> 16) : (px.y & 0xffffu));
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) h[e] = 0u;
    }
    const uint4 pk = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
    *reinterpret_cast<uint4*>(out + pix * kpad + k0) = pk;
  }
}

void im2col3x3_c3(const bf16* x, bf16* out, int N, int H, int W, int cpad, int kpad, cudaStream_t s) {
  if (cpad != 4 || kpad % 8 || kpad < 32) throw std::runtime_error("[b200] im2col3x3_c3: cpad must be 4, kpad a multiple of 8 >= 32");
  const long long total = static_cast<long long>(N) * H * W * (kpad / 8);
  const int blocks = static_cast<int>((total + 255) / 256 < sm_count() * 32 ? (total + 255) / 256 : sm_count() * 32);
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
  const int blocks = static_cast<int>((total + 255) / 256 < sm_count() * 16 ? (total + 255) / 256 : sm_count() * 16);
  nchw_f32_to_nhwc_bf16_kernel<<<blocks, 256, 0, s>>>(x, y, N, C, H, W, cpad);
  count_launch();
  check_last("nchw_f32_to_nhwc_bf16");
}

}  // namespace b200
