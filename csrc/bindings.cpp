// pybind11 surface of the native library: adapts torch tensors to the raw-pointer API of api.h.
// Every op runs on torch's current CUDA stream of the tensor's device.
#include <cstdlib>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "api.h"

namespace py = pybind11;
using b200::bf16;

static cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }

static const bf16* bfp(const at::Tensor& t) {
  TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kBFloat16, "expected a CUDA bf16 tensor");
  return reinterpret_cast<const bf16*>(t.data_ptr());
}
static bf16* bfp_mut(at::Tensor& t) { return const_cast<bf16*>(bfp(t)); }
static float* f32p(const at::Tensor& t) {
  TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kFloat, "expected a CUDA fp32 tensor");
  return reinterpret_cast<float*>(t.data_ptr());
}
static const bf16* bfp_opt(const c10::optional<at::Tensor>& t) { return t.has_value() ? bfp(*t) : nullptr; }
static float* f32p_opt(const c10::optional<at::Tensor>& t) { return t.has_value() ? f32p(*t) : nullptr; }

// ------------------------------------------------------------------------------------- GEMM / conv
static void gemm(const at::Tensor& A, bool a_mn, const at::Tensor& B, bool b_mn, int64_t M, int64_t N,
                 int64_t K, at::Tensor out, int64_t ldo, int64_t epi, c10::optional<at::Tensor> bias,
                 double alpha, int64_t ksplit, int64_t bn) {
  c10::cuda::CUDAGuard g(A.device());
  TORCH_CHECK(A.dim() == 2 && B.dim() == 2 && A.stride(1) == 1 && B.stride(1) == 1, "gemm: 2-D row-major operands");
  b200::gemm_bf16(bfp(A), A.stride(0), a_mn, bfp(B), B.stride(0), b_mn, (int)M, (int)N, (int)K, out.data_ptr(),
                  ldo, (int)epi, f32p_opt(bias), (float)alpha, (int)ksplit, (int)bn, cur_stream());
}

static void conv_fprop(const at::Tensor& x, const at::Tensor& w, c10::optional<at::Tensor> bias, at::Tensor y,
                       bool relu, int64_t bn) {
  c10::cuda::CUDAGuard g(x.device());
  TORCH_CHECK(x.dim() == 4 && x.is_contiguous() && w.is_contiguous() && y.is_contiguous(), "conv_fprop: contiguous NHWC");
  const int N = x.size(0), H = x.size(1), W = x.size(2), Cin = x.size(3), Cout = w.size(0);
  TORCH_CHECK(w.numel() == (int64_t)Cout * 9 * Cin && y.numel() == (int64_t)N * H * W * Cout, "conv_fprop: shape mismatch");
  b200::conv3x3_fprop(bfp(x), bfp(w), f32p_opt(bias), bfp_mut(y), N, H, W, Cin, Cout, relu, (int)bn, cur_stream());
}
// EXPERIMENTAL (B200_FUSE_POOL=1): pooled activation + int32 mask words [N,H/2,W/2,Cout/32,4]
static void conv_fprop_pool(const at::Tensor& x, const at::Tensor& w, c10::optional<at::Tensor> bias,
                            at::Tensor pooled, at::Tensor mask, int64_t bn) {
  c10::cuda::CUDAGuard g(x.device());
  TORCH_CHECK(x.dim() == 4 && x.is_contiguous() && w.is_contiguous() && pooled.is_contiguous() && mask.is_contiguous(),
              "conv_fprop_pool: contiguous NHWC");
  const int N = x.size(0), H = x.size(1), W = x.size(2), Cin = x.size(3), Cout = w.size(0);
  TORCH_CHECK(w.numel() == (int64_t)Cout * 9 * Cin && pooled.numel() == (int64_t)N * (H / 2) * (W / 2) * Cout &&
                  pooled.scalar_type() == at::kBFloat16,
              "conv_fprop_pool: shape mismatch");
  TORCH_CHECK(mask.scalar_type() == at::kInt && mask.numel() == (int64_t)N * (H / 2) * (W / 2) * (Cout / 32) * 4,
              "conv_fprop_pool: mask must be int32 [N,H/2,W/2,Cout/32,4]");
  b200::conv3x3_fprop_pool(bfp(x), bfp(w), f32p_opt(bias), bfp_mut(pooled),
                           reinterpret_cast<uint32_t*>(mask.data_ptr<int>()), N, H, W, Cin, Cout, (int)bn, cur_stream());
}
static bool conv_pool_fusable(int64_t N, int64_t H, int64_t W) { return b200::conv3x3_pool_fusable((int)N, (int)H, (int)W); }
static void unpool(const at::Tensor& dp, const at::Tensor& mask, at::Tensor dz, c10::optional<at::Tensor> colsum) {
  c10::cuda::CUDAGuard g(dp.device());
  TORCH_CHECK(dp.dim() == 4 && dz.dim() == 4 && dp.is_contiguous() && dz.is_contiguous() && mask.is_contiguous() &&
                  mask.scalar_type() == at::kInt,
              "unpool2x2: contiguous NHWC bf16 + int32 mask");
  const int N = dz.size(0), H = dz.size(1), W = dz.size(2), C = dz.size(3);
  TORCH_CHECK(dp.size(0) == N && dp.size(1) == H / 2 && dp.size(2) == W / 2 && dp.size(3) == C &&
                  mask.numel() == (int64_t)N * (H / 2) * (W / 2) * (C / 32) * 4,
              "unpool2x2: shape mismatch");
  b200::unpool2x2(bfp(dp), reinterpret_cast<const uint32_t*>(mask.data_ptr<int>()), bfp_mut(dz), f32p_opt(colsum), N, H,
                  W, C, cur_stream());
}

static void conv_dgrad(const at::Tensor& dz, const at::Tensor& w, c10::optional<at::Tensor> mask, at::Tensor dx,
                       c10::optional<at::Tensor> colsum, int64_t bn) {
  c10::cuda::CUDAGuard g(dz.device());
  TORCH_CHECK(dz.dim() == 4 && dz.is_contiguous() && w.is_contiguous() && dx.is_contiguous(), "conv_dgrad: contiguous NHWC");
  const int N = dz.size(0), H = dz.size(1), W = dz.size(2), Cout = dz.size(3), Cin = dx.size(3);
  TORCH_CHECK(w.numel() == (int64_t)Cout * 9 * Cin, "conv_dgrad: weight shape mismatch");
  b200::conv3x3_dgrad(bfp(dz), bfp(w), bfp_opt(mask), bfp_mut(dx), f32p_opt(colsum), N, H, W, Cin, Cout, (int)bn, cur_stream());
}
static void conv_wgrad(const at::Tensor& dz, const at::Tensor& x, at::Tensor dw, double scale, int64_t ksplit,
                       int64_t bn) {
  c10::cuda::CUDAGuard g(dz.device());
  TORCH_CHECK(dz.dim() == 4 && dz.is_contiguous() && x.is_contiguous(), "conv_wgrad: contiguous NHWC");
  const int N = dz.size(0), H = dz.size(1), W = dz.size(2), Cout = dz.size(3), Cin = x.size(3);
  TORCH_CHECK(dw.numel() == (int64_t)Cout * 9 * Cin, "conv_wgrad: dw shape mismatch");
  b200::conv3x3_wgrad(bfp(dz), bfp(x), f32p(dw), N, H, W, Cin, Cout, (float)scale, (int)ksplit, (int)bn, cur_stream());
}

static void conv0_fprop(const at::Tensor& x4, const at::Tensor& w0, const at::Tensor& bias, at::Tensor y) {
  c10::cuda::CUDAGuard g(x4.device());
  TORCH_CHECK(x4.dim() == 4 && x4.size(3) == 4 && x4.is_contiguous() && w0.numel() == 64 * 64 && y.is_contiguous(),
              "conv0_fprop: x4 [N,H,W,4], w0 [64,64]");
  b200::conv0_fprop(bfp(x4), bfp(w0), f32p(bias), bfp_mut(y), x4.size(0), x4.size(1), x4.size(2), cur_stream());
}
static void conv0_wgrad(const at::Tensor& dz, const at::Tensor& x4, at::Tensor dw0) {
  c10::cuda::CUDAGuard g(x4.device());
  TORCH_CHECK(x4.dim() == 4 && x4.size(3) == 4 && x4.is_contiguous() && dz.is_contiguous() && dw0.numel() == 64 * 64,
              "conv0_wgrad: x4 [N,H,W,4], dw0 [64,64]");
  b200::conv0_wgrad(bfp(dz), bfp(x4), f32p(dw0), x4.size(0), x4.size(1), x4.size(2), cur_stream());
}

static void shift_probe(const at::Tensor& A, const at::Tensor& B, at::Tensor out, int64_t shift, int64_t pitch,
                        bool use_base_offset, int64_t mode) {
  c10::cuda::CUDAGuard g(A.device());
  b200::umma_shift_probe(bfp(A), (int)A.size(0), bfp(B), f32p(out), (int)shift, (int)pitch, use_base_offset ? 1 : 0,
                         (int)mode, cur_stream());
}

// ------------------------------------------------------------------------------------ element-wise
static void maxpool_fwd(const at::Tensor& x, at::Tensor y) {
  c10::cuda::CUDAGuard g(x.device());
  b200::maxpool2x2_fwd(bfp(x), bfp_mut(y), x.size(0), x.size(1), x.size(2), x.size(3), cur_stream());
}
static void maxpool_relu_bwd(const at::Tensor& y, const at::Tensor& dp, at::Tensor dz,
                             c10::optional<at::Tensor> colsum) {
  c10::cuda::CUDAGuard g(y.device());
  b200::maxpool2x2_relu_bwd(bfp(y), bfp(dp), bfp_mut(dz), f32p_opt(colsum), y.size(0), y.size(1), y.size(2), y.size(3),
                            cur_stream());
}
static void avgpool_fwd(const at::Tensor& x, at::Tensor y) {
  c10::cuda::CUDAGuard g(x.device());
  b200::adaptive_avgpool_fwd(bfp(x), bfp_mut(y), x.size(0), x.size(1), x.size(2), x.size(3), y.size(1), y.size(2), cur_stream());
}
static void avgpool_bwd(const at::Tensor& dy, at::Tensor dx) {
  c10::cuda::CUDAGuard g(dy.device());
  b200::adaptive_avgpool_bwd(bfp(dy), bfp_mut(dx), dx.size(0), dx.size(1), dx.size(2), dx.size(3), dy.size(1), dy.size(2), cur_stream());
}
static void bias_grad(const at::Tensor& dz, at::Tensor db, int64_t rows, int64_t C, int64_t ld, double scale) {
  c10::cuda::CUDAGuard g(dz.device());
  if (C % 8 == 0 && ld == C) b200::bias_grad(bfp(dz), f32p(db), rows, (int)C, (float)scale, cur_stream());
  else b200::bias_grad_ld(bfp(dz), f32p(db), rows, (int)C, (int)ld, (float)scale, cur_stream());
}
static void fc_bias_act(at::Tensor acc, c10::optional<at::Tensor> bias, c10::optional<at::Tensor> y,
                        c10::optional<at::Tensor> y_f32, int64_t B, int64_t N, bool relu, double drop_p,
                        int64_t seed, int64_t offset, bool clear) {
  c10::cuda::CUDAGuard g(acc.device());
  b200::fc_bias_act(f32p(acc), f32p_opt(bias), y.has_value() ? bfp_mut(*y) : nullptr, f32p_opt(y_f32), (int)B,
                    (int)N, relu, (float)drop_p, (unsigned long long)seed, (unsigned long long)offset, clear,
                    cur_stream());
}
static void fc_grad_act(at::Tensor acc, c10::optional<at::Tensor> act, at::Tensor dz, int64_t B, int64_t N,
                        bool relu, double drop_p, bool clear) {
  c10::cuda::CUDAGuard g(acc.device());
  b200::fc_grad_act(f32p(acc), bfp_opt(act), bfp_mut(dz), (int)B, (int)N, relu, (float)drop_p, 0, 0, clear, cur_stream());
}
static void cast_to_bf16(const at::Tensor& x, at::Tensor y) {
  c10::cuda::CUDAGuard g(x.device());
  b200::cast_f32_to_bf16(f32p(x), bfp_mut(y), x.numel(), cur_stream());
}
static void cast_to_f32(const at::Tensor& x, at::Tensor y) {
  c10::cuda::CUDAGuard g(x.device());
  b200::cast_bf16_to_f32(bfp(x), f32p(y), x.numel(), cur_stream());
}

static void cross_entropy(const at::Tensor& logits, const at::Tensor& target, c10::optional<at::Tensor> dlogits,
                          int64_t ldd, c10::optional<at::Tensor> meter, c10::optional<at::Tensor> loss_out,
                          double grad_scale, c10::optional<at::Tensor> class_weights) {
  c10::cuda::CUDAGuard g(logits.device());
  TORCH_CHECK(target.scalar_type() == at::kLong && logits.is_contiguous(), "cross_entropy: int64 targets, contiguous logits");
  b200::cross_entropy_fused(f32p(logits), reinterpret_cast<const long long*>(target.data_ptr()),
                            dlogits.has_value() ? bfp_mut(*dlogits) : nullptr, (int)ldd, f32p_opt(meter),
                            f32p_opt(loss_out), (int)logits.size(0), (int)logits.size(1), (float)grad_scale,
                            f32p_opt(class_weights), cur_stream());
}

static void adam(at::Tensor p, at::Tensor m, at::Tensor v, c10::optional<at::Tensor> g32,
                 c10::optional<at::Tensor> g16, c10::optional<at::Tensor> shadow, double lr, double b1, double b2,
                 double eps, double wd, int64_t step, double grad_scale, c10::optional<at::Tensor> gzero) {
  c10::cuda::CUDAGuard g(p.device());
  b200::adam_fused(f32p(p), f32p(m), f32p(v), f32p_opt(g32), bfp_opt(g16),
                   shadow.has_value() ? bfp_mut(*shadow) : nullptr, p.numel(), (float)lr, (float)b1, (float)b2,
                   (float)eps, (float)wd, (int)step, (float)grad_scale, false, f32p_opt(gzero), cur_stream());
}
static void sgd(at::Tensor p, at::Tensor mom, c10::optional<at::Tensor> g32, c10::optional<at::Tensor> g16,
                c10::optional<at::Tensor> shadow, double lr, double momentum, double wd, bool first,
                double grad_scale, c10::optional<at::Tensor> gzero) {
  c10::cuda::CUDAGuard g(p.device());
  b200::sgd_fused(f32p(p), f32p(mom), f32p_opt(g32), bfp_opt(g16), shadow.has_value() ? bfp_mut(*shadow) : nullptr,
                  p.numel(), (float)lr, (float)momentum, (float)wd, first, (float)grad_scale, f32p_opt(gzero),
                  cur_stream());
}

static void augment(const at::Tensor& src, const at::Tensor& params, at::Tensor out, int64_t RH, int64_t RW,
                    int64_t OH, int64_t OW, int64_t mode, int64_t pad, std::vector<double> mean,
                    std::vector<double> stdv) {
  c10::cuda::CUDAGuard g(src.device());
  TORCH_CHECK(src.is_cuda() && src.scalar_type() == at::kByte && src.is_contiguous() && src.dim() == 4 && src.size(3) == 3,
              "augment: uint8 [N,H,W,3] CUDA tensor");
  float m[3] = {(float)mean[0], (float)mean[1], (float)mean[2]}, s[3] = {(float)stdv[0], (float)stdv[1], (float)stdv[2]};
  b200::augment_fused(src.data_ptr<uint8_t>(), f32p(params), bfp_mut(out), src.size(0), src.size(1), src.size(2),
                      (int)RH, (int)RW, (int)OH, (int)OW, (int)mode, (int)pad, m, s, cur_stream());
}
static void im2col_c3(const at::Tensor& x, at::Tensor out, int64_t kpad) {
  c10::cuda::CUDAGuard g(x.device());
  b200::im2col3x3_c3(bfp(x), bfp_mut(out), x.size(0), x.size(1), x.size(2), x.size(3), (int)kpad, cur_stream());
}
static void nchw_to_nhwc(const at::Tensor& x, at::Tensor y) {
  c10::cuda::CUDAGuard g(x.device());
  TORCH_CHECK(x.is_contiguous() && x.dim() == 4, "nchw_to_nhwc: contiguous NCHW fp32");
  b200::nchw_f32_to_nhwc_bf16(f32p(x), bfp_mut(y), x.size(0), x.size(1), x.size(2), x.size(3), y.size(3), cur_stream());
}

// ------------------------------------------------------------------------------------ PNG decode
namespace b200 {
struct PngImage {
  int w = 0, h = 0;
  std::vector<uint8_t> rgb;
  bool ok = false;
};
void decode_png_files(const std::vector<std::string>& paths, int threads, std::vector<PngImage>& out);
}  // namespace b200

// All files decoded natively into one uint8 [N,H,W,3] tensor; throws if a file is not a plain 8-bit
// PNG or the sizes differ (the caller then falls back to PIL).
static at::Tensor decode_pngs(const std::vector<std::string>& paths, int64_t threads) {
  std::vector<b200::PngImage> imgs;
  {
    py::gil_scoped_release nogil;
    b200::decode_png_files(paths, (int)threads, imgs);
  }
  TORCH_CHECK(!imgs.empty(), "decode_pngs: no files");
  const int w = imgs[0].w, h = imgs[0].h;
  for (size_t i = 0; i < imgs.size(); ++i)
    TORCH_CHECK(imgs[i].ok && imgs[i].w == w && imgs[i].h == h, "decode_pngs: unsupported or non-uniform file ", paths[i]);
  at::Tensor out = at::empty({(int64_t)imgs.size(), h, w, 3}, at::kByte);
  uint8_t* dst = out.data_ptr<uint8_t>();
  const size_t per = (size_t)w * h * 3;
  for (size_t i = 0; i < imgs.size(); ++i) memcpy(dst + i * per, imgs[i].rgb.data(), per);
  return out;
}

// ------------------------------------------------------------------------------- batch prefetcher
namespace b200 {
struct PrefetchConfig {
  const uint8_t* cache;
  const int64_t* labels;
  int H, W, mb, slots;
  uint8_t* ring_img;
  float* ring_par;
  int64_t* ring_lab;
  bool train;
  float scale_lo, scale_hi, ratio_lo, ratio_hi, degrees;
};
class BatchPrefetcher;
BatchPrefetcher* prefetcher_create(const PrefetchConfig& c);
void prefetcher_destroy(BatchPrefetcher* p);
void prefetcher_start(BatchPrefetcher* p, std::vector<int64_t> idx, uint64_t seed);
bool prefetcher_next(BatchPrefetcher* p, int& slot, int& count);
void prefetcher_release(BatchPrefetcher* p, int slot, uint64_t cuda_event);
void prefetcher_stop(BatchPrefetcher* p);
}  // namespace b200

struct PyPrefetcher {
  b200::BatchPrefetcher* p = nullptr;
  at::Tensor cache, labels, ring_img, ring_par, ring_lab;     // keep the storage alive
  PyPrefetcher(at::Tensor cache_, at::Tensor labels_, at::Tensor ring_img_, at::Tensor ring_par_, at::Tensor ring_lab_,
               bool train, double scale_lo, double scale_hi, double ratio_lo, double ratio_hi, double degrees)
      : cache(cache_), labels(labels_), ring_img(ring_img_), ring_par(ring_par_), ring_lab(ring_lab_) {
    TORCH_CHECK(cache.dim() == 4 && cache.size(3) == 3 && cache.scalar_type() == at::kByte && cache.is_contiguous() &&
                !cache.is_cuda(), "Prefetcher: cache must be a CPU uint8 [N,H,W,3] tensor");
    TORCH_CHECK(ring_img.dim() == 5 && ring_img.is_contiguous() && ring_par.is_contiguous() && ring_lab.is_contiguous() &&
                labels.scalar_type() == at::kLong && ring_lab.scalar_type() == at::kLong && ring_par.scalar_type() == at::kFloat,
                "Prefetcher: bad ring tensors");
    b200::PrefetchConfig c;
    c.cache = cache.data_ptr<uint8_t>(); c.labels = labels.data_ptr<int64_t>();
    c.H = (int)cache.size(1); c.W = (int)cache.size(2);
    c.slots = (int)ring_img.size(0); c.mb = (int)ring_img.size(1);
    c.ring_img = ring_img.data_ptr<uint8_t>(); c.ring_par = ring_par.data_ptr<float>(); c.ring_lab = ring_lab.data_ptr<int64_t>();
    c.train = train; c.scale_lo = (float)scale_lo; c.scale_hi = (float)scale_hi;
    c.ratio_lo = (float)ratio_lo; c.ratio_hi = (float)ratio_hi; c.degrees = (float)degrees;
    p = b200::prefetcher_create(c);
  }
  ~PyPrefetcher() { if (p) b200::prefetcher_destroy(p); }
  void start(std::vector<int64_t> idx, uint64_t seed) {
    py::gil_scoped_release nogil;
    b200::prefetcher_start(p, std::move(idx), seed);
  }
  py::object next() {
    int slot = -1, count = 0;
    bool ok;
    {
      py::gil_scoped_release nogil;
      ok = b200::prefetcher_next(p, slot, count);
    }
    if (!ok) return py::none();
    return py::make_tuple(slot, count);
  }
  void release(int64_t slot, uint64_t ev) { b200::prefetcher_release(p, (int)slot, ev); }
  void stop() {
    py::gil_scoped_release nogil;
    b200::prefetcher_stop(p);
  }
};

// ------------------------------------------------------------------------------------------- comm
struct PyComm {
  b200::CommCtx c;
};
static PyComm make_comm(int64_t rank, int64_t world, int64_t wire_ptrs_dev, int64_t signal_ptrs_dev,
                        int64_t wire_mc, int64_t signal_mc) {
  PyComm p;
  p.c.rank = (int)rank; p.c.world = (int)world;
  p.c.wire_ptrs = reinterpret_cast<void* const*>(wire_ptrs_dev);
  p.c.signal_ptrs = reinterpret_cast<uint32_t* const*>(signal_ptrs_dev);
  p.c.wire_mc = reinterpret_cast<void*>(wire_mc);
  p.c.signal_mc = reinterpret_cast<uint32_t*>(signal_mc);
  // cross-GPU waits are bounded by wall clock; minutes by default so that host-side skew (rank 0
  // writing a checkpoint, a slow loader) never trips it -- only a dead peer does
  double secs = 600.0;
  if (const char* e = std::getenv("B200_BARRIER_TIMEOUT_S")) secs = std::atof(e) > 0 ? std::atof(e) : secs;
  p.c.timeout_ns = static_cast<unsigned long long>(secs * 1e9);
  return p;
}
static void allreduce(const PyComm& comm, c10::optional<at::Tensor> grad, c10::optional<at::Tensor> grad_out,
                      int64_t start, int64_t n, double inv_world, int64_t algo, bool wire_fp32, int64_t slot,
                      int64_t epoch, int64_t max_ctas) {
  b200::allreduce_fused(comm.c, f32p_opt(grad), f32p_opt(grad_out), start, n, (float)inv_world, (int)algo,
                        wire_fp32, (int)slot, (uint32_t)epoch, (int)max_ctas, cur_stream());
}
// EXPERIMENTAL (--zero1): fused reduce-scatter + Adam on the owned cells + all-gather of bf16 weights
static void zero1_step(const PyComm& comm, c10::optional<at::Tensor> grad, at::Tensor p, at::Tensor m, at::Tensor v,
                       at::Tensor shadow, int64_t start, int64_t n, double inv_world, int64_t algo, int64_t slot,
                       int64_t epoch, int64_t max_ctas, double lr, double beta1, double beta2, double eps,
                       double weight_decay, int64_t step) {
  TORCH_CHECK(p.scalar_type() == at::kFloat && m.scalar_type() == at::kFloat && v.scalar_type() == at::kFloat &&
                  shadow.scalar_type() == at::kBFloat16 && p.numel() == m.numel() && p.numel() == v.numel() &&
                  p.numel() == shadow.numel() && start + n <= p.numel(),
              "zero1_step: full fp32 arenas (p, m, v) and the bf16 shadow of the same length");
  b200::zero1_step(comm.c, grad.has_value() ? grad->data_ptr<float>() : nullptr, p.data_ptr<float>(),
                   m.data_ptr<float>(), v.data_ptr<float>(), bfp_mut(shadow), start, n, (float)inv_world, (int)algo,
                   (int)slot, (uint32_t)epoch, (int)max_ctas, (float)lr, (float)beta1, (float)beta2, (float)eps,
                   (float)weight_decay, (int)step, cur_stream());
}
static int64_t allreduce_grid(int64_t n, int64_t world, int64_t max_ctas, bool wire_fp32) {
  return b200::allreduce_grid(n, (int)world, (int)max_ctas, wire_fp32);
}
static void broadcast(const PyComm& comm, at::Tensor data, int64_t root, int64_t slot, int64_t epoch) {
  b200::broadcast_fused(comm.c, f32p(data), data.numel(), (int)root, (int)slot, (uint32_t)epoch, cur_stream());
}
static void barrier(const PyComm& comm, int64_t slot, int64_t epoch) {
  b200::device_barrier(comm.c, (int)slot, (uint32_t)epoch, cur_stream());
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "distributed-vgg-f_b200 native sm_100a kernels";
  m.def("launch_count", &b200::launch_count);
  m.def("decode_pngs", &decode_pngs);
  m.def("gemm", &gemm);
  m.def("shift_probe", &shift_probe);
  m.def("conv0_fprop", &conv0_fprop);
  m.def("conv0_wgrad", &conv0_wgrad);
  m.def("conv_fprop", &conv_fprop);
  m.def("conv_fprop_pool", &conv_fprop_pool);
  m.def("conv_pool_fusable", &conv_pool_fusable);
  m.def("unpool2x2", &unpool);
  m.def("conv_dgrad", &conv_dgrad);
  m.def("conv_wgrad", &conv_wgrad);
  m.def("maxpool_fwd", &maxpool_fwd);
  m.def("maxpool_relu_bwd", &maxpool_relu_bwd);
  m.def("avgpool_fwd", &avgpool_fwd);
  m.def("avgpool_bwd", &avgpool_bwd);
  m.def("bias_grad", &bias_grad);
  m.def("fc_bias_act", &fc_bias_act);
  m.def("fc_grad_act", &fc_grad_act);
  m.def("cast_to_bf16", &cast_to_bf16);
  m.def("cast_to_f32", &cast_to_f32);
  m.def("cross_entropy", &cross_entropy);
  m.def("head_ce_supported", [](int64_t B, int64_t C, int64_t K) { return b200::head_ce_supported((int)B, (int)C, (int)K); });
  m.def("head_ce", [](const at::Tensor& h, const at::Tensor& W, const at::Tensor& bias, const at::Tensor& target,
                      at::Tensor logits, c10::optional<at::Tensor> dlogits, int64_t ldd, c10::optional<at::Tensor> dW,
                      c10::optional<at::Tensor> db, c10::optional<at::Tensor> dh, double drop_scale, bool relu,
                      c10::optional<at::Tensor> meter, c10::optional<at::Tensor> loss_out,
                      c10::optional<at::Tensor> class_weights) {
    c10::cuda::CUDAGuard g(h.device());
    TORCH_CHECK(h.dim() == 2 && W.dim() == 2 && h.size(1) == W.size(1) && h.is_contiguous() && W.is_contiguous(),
                "head_ce: h [B][K], W [C][K], contiguous");
    TORCH_CHECK(target.scalar_type() == at::kLong && logits.is_contiguous() && logits.size(0) == h.size(0) &&
                logits.size(1) == W.size(0), "head_ce: int64 targets, logits [B][C]");
    TORCH_CHECK(dW.has_value() == db.has_value(), "head_ce: dW and db come together");
    b200::head_ce_fused(bfp(h), bfp(W), f32p(bias), reinterpret_cast<const long long*>(target.data_ptr()), f32p(logits),
                        dlogits.has_value() ? bfp_mut(*dlogits) : nullptr, (int)ldd, f32p_opt(dW), f32p_opt(db),
                        dh.has_value() ? bfp_mut(*dh) : nullptr, (float)drop_scale, relu, f32p_opt(meter),
                        f32p_opt(loss_out), (int)h.size(0), (int)W.size(0), (int)h.size(1), f32p_opt(class_weights),
                        cur_stream());
  });
  m.def("adam", &adam);
  m.def("sgd", &sgd);
  m.def("augment", &augment);
  m.def("im2col_c3", &im2col_c3);
  m.def("nchw_to_nhwc", &nchw_to_nhwc);
  py::class_<PyPrefetcher>(m, "Prefetcher")
      .def(py::init<at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor, bool, double, double, double, double, double>())
      .def("start", &PyPrefetcher::start)
      .def("next", &PyPrefetcher::next)
      .def("release", &PyPrefetcher::release)
      .def("stop", &PyPrefetcher::stop);
  py::class_<PyComm>(m, "Comm");
  m.def("probe_background", [](at::Tensor buf, int64_t mode, int64_t ctas, int64_t reps, at::Tensor sink) {
    b200::probe_background(f32p(buf), buf.numel(), (int)mode, (int)ctas, (int)reps, f32p(sink), cur_stream());
  });
  m.def("make_comm", &make_comm);
  m.def("allreduce", &allreduce);
  m.def("zero1_step", &zero1_step);
  m.def("allreduce_grid", &allreduce_grid);
  m.def("broadcast", &broadcast);
  m.def("barrier", &barrier);
  m.def("allreduce_signal_words", &b200::allreduce_signal_words);
  m.attr("EPI_F32_STORE") = 0;
  m.attr("EPI_F32_ATOMIC") = 1;
  m.attr("EPI_F32_ATOMIC_T") = 2;
  m.attr("EPI_BF16_BIAS_RELU") = 3;
  m.attr("EPI_F32_STORE_T") = 4;
  m.attr("EPI_BF16_STORE") = 5;
  m.attr("AR_ONESHOT") = 0;
  m.attr("AR_TWOSHOT") = 1;
  m.attr("AR_NVLS") = 2;
}
