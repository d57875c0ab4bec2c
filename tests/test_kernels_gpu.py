"""Numerics of every sm_100a kernel against a plain PyTorch fp32 reference of the same op.

bf16 operands: inputs are rounded to bf16 first, the reference is evaluated in fp32 on those
rounded values, so the only differences are accumulation order and the final bf16 rounding.
"""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from distributed_vgg_f_b200 import ops
    ops.require()
    return ops


def rel_err(got: torch.Tensor, ref: torch.Tensor) -> float:
    got, ref = got.float(), ref.float()
    return float((got - ref).abs().max() / (ref.abs().max() + 1e-12))


def bf(t):
    return t.to(torch.bfloat16)


# ------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 64, 512), (4096, 64, 4096), (200, 48, 200),
                                   (4096, 96, 25088 // 8)])
@pytest.mark.parametrize("ksplit", [1, 4])
def test_gemm_kk_transposed_out(M, N, K, ksplit):
    """FC forward form (swap-AB): out[b, n] += sum_k W[n, k] X[b, k]."""
    ops = _ops()
    torch.manual_seed(0)
    Kp = (K + 7) // 8 * 8
    A = bf(torch.randn(M, Kp, device=DEV))[:, :K]
    B = bf(torch.randn(N, Kp, device=DEV))[:, :K]
    out = torch.zeros(N, M, device=DEV)
    ops.gemm(A, B, out, M=M, N=N, K=K, epi="f32_atomic_t", ksplit=ksplit)
    ref = B.float() @ A.float().t()
    assert rel_err(out, ref) < 2e-3


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (25088 // 4, 64, 4096), (300, 40, 130)])
def test_gemm_mnk_store_t(M, N, K):
    """FC dgrad form: A stored [K][M] (W[n_out][k_in] read as A[m=k_in, k=n_out])."""
    ops = _ops()
    torch.manual_seed(1)
    Mp = (M + 7) // 8 * 8
    Kp = (K + 7) // 8 * 8
    A_km = bf(torch.randn(K, Mp, device=DEV))[:, :M]          # [K][M]
    B = bf(torch.randn(N, Kp, device=DEV))[:, :K]             # [N][K]
    out = torch.zeros(N, M, device=DEV)
    ops.gemm(A_km, B, out, M=M, N=N, K=K, a_mn=True, epi="f32_store_t")
    ref = B.float() @ A_km.float()
    assert rel_err(out, ref) < 2e-3


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (4096, 1024, 64), (512, 256, 96), (3 * 8, 512, 64)])
def test_gemm_mnmn_store(M, N, K):
    """FC wgrad form: dW[n_out, k_in] = sum_b dY[b, n_out] X[b, k_in]; both operands [K][*]."""
    ops = _ops()
    torch.manual_seed(2)
    A_km = bf(torch.randn(K, M, device=DEV))
    B_kn = bf(torch.randn(K, N, device=DEV))
    out = torch.empty(M, N, device=DEV)
    ops.gemm(A_km, B_kn, out, M=M, N=N, K=K, a_mn=True, b_mn=True, epi="f32_store", alpha=0.5)
    ref = 0.5 * (A_km.float().t() @ B_kn.float())
    assert rel_err(out, ref) < 2e-3


@pytest.mark.parametrize("M,N,K,off", [(128, 64, 64, 0), (4096, 1024, 64, 0), (512, 256, 96, 8), (256, 72, 64, 0),
                                       (24, 512, 64, 4)])
def test_gemm_mnmn_bf16_store(M, N, K, off):
    """FC wgrad written straight to the bf16 wire: out = bf16(alpha * dY^T X).  ``off`` shifts the
    output start by a few elements (32-byte vs 16-byte vs scalar store paths), N=72 is ragged."""
    ops = _ops()
    torch.manual_seed(4)
    A_km = bf(torch.randn(K, M, device=DEV))
    B_kn = bf(torch.randn(K, N, device=DEV))
    buf = torch.full((M * N + 64,), 7.0, dtype=torch.bfloat16, device=DEV)
    out = buf[off:off + M * N].view(M, N)
    ops.gemm(A_km, B_kn, out, M=M, N=N, K=K, a_mn=True, b_mn=True, epi="bf16_store", alpha=0.125)
    ref = 0.125 * (A_km.float().t() @ B_kn.float())
    assert rel_err(out, ref) < 1e-2
    assert bool((buf[:off] == 7.0).all()) and bool((buf[off + M * N:] == 7.0).all())


def test_gemm_bf16_bias_relu():
    """Layer-0 im2col GEMM: bf16(relu(A W^T + b))."""
    ops = _ops()
    torch.manual_seed(3)
    M, N, K = 1000, 64, 64
    A = bf(torch.randn(M, K, device=DEV))
    W = bf(torch.randn(N, K, device=DEV) * 0.2)
    bias = torch.randn(N, device=DEV)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(A, W, out, M=M, N=N, K=K, epi="bf16_bias_relu", bias=bias)
    ref = torch.relu(A.float() @ W.float().t() + bias)
    assert rel_err(out, ref) < 1e-2


# ------------------------------------------------------------------------------------------ conv
def _conv_case(N, H, W, Cin, Cout, seed=0):
    torch.manual_seed(seed)
    x = bf(torch.randn(N, H, W, Cin, device=DEV))
    w = bf(torch.randn(Cout, 3, 3, Cin, device=DEV) * (1.0 / math.sqrt(9 * Cin)))
    b = torch.randn(Cout, device=DEV) * 0.1
    return x, w, b


def _conv_ref(x, w, b):
    xr = x.float().permute(0, 3, 1, 2)
    wr = w.float().permute(0, 3, 1, 2)
    return F.conv2d(xr, wr, b, padding=1)


CONV_SHAPES = [(2, 16, 16, 64, 64), (1, 28, 28, 128, 256), (4, 14, 14, 512, 512), (2, 56, 56, 64, 128),
               (3, 12, 20, 64, 64), (1, 224, 224, 64, 64),
               # halo-tile kernel (H % 16 == 0, W % 8 == 0, W >= 32): resident / streaming weights
               (2, 32, 32, 64, 64), (1, 32, 32, 64, 128), (1, 48, 64, 128, 128), (1, 32, 32, 256, 256),
               (3, 16, 40, 128, 64)]


@pytest.mark.parametrize("N,H,W,Cin,Cout", CONV_SHAPES)
def test_conv_fprop(N, H, W, Cin, Cout):
    ops = _ops()
    x, w, b = _conv_case(N, H, W, Cin, Cout)
    y = ops.conv3x3_fprop(x, w, b, relu=True)
    ref = torch.relu(_conv_ref(x, w, b)).permute(0, 2, 3, 1)
    assert rel_err(y, ref) < 1e-2


@pytest.mark.parametrize("N,H,W,Cin,Cout", CONV_SHAPES)
def test_conv_dgrad(N, H, W, Cin, Cout):
    ops = _ops()
    x, w, _ = _conv_case(N, H, W, Cin, Cout, seed=1)
    dz = bf(torch.randn(N, H, W, Cout, device=DEV))
    dx = ops.conv3x3_dgrad(dz, w, Cin)
    ref = F.conv_transpose2d(dz.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2),
                             padding=1).permute(0, 2, 3, 1)
    assert rel_err(dx, ref) < 1e-2
    # fused ReLU mask
    mask_src = bf(torch.randn(N, H, W, Cin, device=DEV))
    dx2 = ops.conv3x3_dgrad(dz, w, Cin, mask_src=mask_src)
    assert rel_err(dx2, ref * (mask_src.float() > 0)) < 1e-2
    # fused column sum (= bias gradient of the previous layer): sums exactly what was stored
    cs = torch.zeros(Cin, device=DEV)
    dx3 = ops.conv3x3_dgrad(dz, w, Cin, mask_src=mask_src, colsum=cs)
    assert torch.equal(dx3, dx2)
    assert rel_err(cs, dx2.float().sum((0, 1, 2))) < 1e-3


@pytest.mark.parametrize("N,H,W,Cin,Cout", CONV_SHAPES)
def test_conv_wgrad(N, H, W, Cin, Cout):
    ops = _ops()
    x, w, _ = _conv_case(N, H, W, Cin, Cout, seed=2)
    dz = bf(torch.randn(N, H, W, Cout, device=DEV))
    dw = torch.zeros(Cout, 3, 3, Cin, device=DEV)
    ops.conv3x3_wgrad(dz, x, dw)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(False)
    wr = w.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    F.conv2d(xr, wr, None, padding=1).backward(dz.float().permute(0, 3, 1, 2))
    ref = wr.grad.permute(0, 2, 3, 1)
    assert rel_err(dw, ref) < 5e-3


@pytest.mark.parametrize("N,H,W", [(2, 16, 16), (1, 30, 22), (3, 64, 64)])
def test_conv0_fused(N, H, W):
    """First conv straight from NHWC4 pixels (im2col tile built in shared memory): fprop + wgrad."""
    ops = _ops()
    C = ops.require()
    torch.manual_seed(0)
    x = bf(torch.randn(N, H, W, 3, device=DEV))
    x4 = torch.zeros(N, H, W, 4, dtype=torch.bfloat16, device=DEV)
    x4[..., :3] = x
    w = bf(torch.randn(64, 3, 3, 3, device=DEV) * 0.2)           # [co][kh][kw][c]
    w0 = torch.zeros(64, 64, dtype=torch.bfloat16, device=DEV)
    w0[:, :27] = w.reshape(64, 27)
    bias = torch.randn(64, device=DEV)
    y = torch.empty(N, H, W, 64, dtype=torch.bfloat16, device=DEV)
    C.conv0_fprop(x4, w0, bias, y)
    xr = x.float().permute(0, 3, 1, 2)
    wr = w.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    ref = torch.relu(F.conv2d(xr, wr, bias, padding=1))
    assert rel_err(y, ref.permute(0, 2, 3, 1)) < 1e-2
    dz = bf(torch.randn(N, H, W, 64, device=DEV))
    dw0 = torch.zeros(64, 64, device=DEV)
    C.conv0_wgrad(dz, x4, dw0)
    F.conv2d(xr, wr, None, padding=1).backward(dz.float().permute(0, 3, 1, 2))
    assert rel_err(dw0[:, :27], wr.grad.permute(0, 2, 3, 1).reshape(64, 27)) < 5e-3
    assert torch.count_nonzero(dw0[:, 27:]) == 0


# ------------------------------------------------------------------------------------- pointwise
EXPERIMENTAL = pytest.mark.skipif(os.environ.get("B200_EXPERIMENTAL", "0") != "1",
                                  reason="experimental kernels: set B200_EXPERIMENTAL=1")


@EXPERIMENTAL
@pytest.mark.parametrize("N,H,W,Cin,Cout", [(2, 32, 32, 64, 64), (1, 48, 64, 128, 128), (2, 56, 56, 64, 128),
                                            (3, 28, 28, 128, 256), (4, 14, 14, 512, 512), (1, 224, 224, 64, 64)])
def test_conv_fprop_pool_fused(N, H, W, Cin, Cout):
    """conv + bias + ReLU + 2x2 max-pool in the conv epilogue, and its unpool backward, against the
    separate kernels (bit-identical by construction: same rounding point, same first-maximum rule)."""
    ops = _ops()
    x, w, b = _conv_case(N, H, W, Cin, Cout, seed=5)
    y = ops.conv3x3_fprop(x, w, b, relu=True)
    p_ref = ops.maxpool2x2(y)
    p, mask = ops.conv3x3_fprop_pool(x, w, b)
    assert torch.equal(p, p_ref)
    dp = bf(torch.randn(N, H // 2, W // 2, Cout, device=DEV))
    cs_ref = torch.zeros(Cout, device=DEV)
    dz_ref = ops.maxpool2x2_relu_bwd(y, dp, colsum=cs_ref)
    cs = torch.zeros(Cout, device=DEV)
    dz = ops.unpool2x2(dp, mask, colsum=cs)
    assert torch.equal(dz, dz_ref)
    assert rel_err(cs, cs_ref) < 1e-3


def test_maxpool_fwd_bwd():
    ops = _ops()
    torch.manual_seed(0)
    x = bf(torch.relu(torch.randn(3, 8, 12, 64, device=DEV)))
    y = ops.maxpool2x2(x)
    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    yr = F.max_pool2d(xr, 2, 2)
    assert torch.equal(y.float(), yr.permute(0, 2, 3, 1))
    dp = bf(torch.randn_like(y.float()))
    dz = ops.maxpool2x2_relu_bwd(x, dp)
    yr.backward(dp.float().permute(0, 3, 1, 2))
    ref = (xr.grad * (xr > 0)).permute(0, 2, 3, 1)
    assert rel_err(dz, ref) < 1e-6
    # fused bias gradient (column sum of dz)
    cs = torch.zeros(64, device=DEV)
    dz2 = ops.maxpool2x2_relu_bwd(x, dp, colsum=cs)
    assert torch.equal(dz2, dz)
    assert rel_err(cs, dz.float().sum((0, 1, 2))) < 1e-4


@pytest.mark.parametrize("H", [4, 7, 14])
def test_adaptive_avgpool(H):
    ops = _ops()
    torch.manual_seed(0)
    x = bf(torch.randn(2, H, H, 64, device=DEV))
    y = ops.adaptive_avgpool(x, 7, 7)
    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    yr = F.adaptive_avg_pool2d(xr, (7, 7))
    assert rel_err(y, yr.permute(0, 2, 3, 1)) < 1e-2
    dy = bf(torch.randn_like(y.float()))
    dx = ops.adaptive_avgpool_bwd(dy, H, H)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    assert rel_err(dx, xr.grad.permute(0, 2, 3, 1)) < 1e-2


@pytest.mark.parametrize("rows,C", [(64, 4096), (5000, 64), (3000, 512), (64, 3)])
def test_bias_grad(rows, C):
    ops = _ops()
    torch.manual_seed(0)
    ld = (C + 7) // 8 * 8
    dz = bf(torch.randn(rows, ld, device=DEV))
    db = torch.zeros(C, device=DEV)
    ops.bias_grad(dz, db, rows, C, ld=ld, scale=0.5)
    ref = 0.5 * dz.float()[:, :C].sum(0)
    assert rel_err(db, ref) < 1e-3


def test_fc_epilogues_and_dropout():
    ops = _ops()
    torch.manual_seed(0)
    B, N = 64, 4096
    acc0 = torch.randn(B, N, device=DEV)
    bias = torch.randn(N, device=DEV)
    acc = acc0.clone()
    y = torch.empty(B, N, dtype=torch.bfloat16, device=DEV)
    ops.fc_bias_act(acc, bias, y, B=B, N=N, relu=True, drop_p=0.0)
    assert torch.count_nonzero(acc) == 0
    assert rel_err(y, torch.relu(acc0 + bias)) < 1e-2
    acc = acc0.clone()
    ops.fc_bias_act(acc, bias, y, B=B, N=N, relu=True, drop_p=0.5, seed=7, offset=3)
    base = torch.relu(acc0 + bias)
    kept = (y.float() != 0) | (base == 0)
    frac = float(((y.float() == 0) & (base > 0)).sum() / (base > 0).sum())
    assert 0.47 < frac < 0.53                      # Bernoulli(0.5) drop rate
    sel = (y.float() != 0)
    assert rel_err(y.float()[sel], 2.0 * base[sel]) < 1e-2
    y2 = torch.empty_like(y)
    acc = acc0.clone()
    ops.fc_bias_act(acc, bias, y2, B=B, N=N, relu=True, drop_p=0.5, seed=7, offset=3)
    assert torch.equal(y, y2)                      # same (seed, offset) -> same mask
    # backward epilogue
    g = torch.randn(B, N, device=DEV)
    dz = torch.empty(B, N, dtype=torch.bfloat16, device=DEV)
    ops.fc_grad_act(g.clone(), y, dz, B=B, N=N, relu=True, drop_p=0.5)
    assert rel_err(dz, g * (y.float() > 0) * 2.0) < 1e-2
    del kept


@pytest.mark.parametrize("B,C", [(64, 3), (96, 1000), (7, 10)])
def test_cross_entropy(B, C):
    ops = _ops()
    torch.manual_seed(0)
    logits = torch.randn(B, C, device=DEV) * 3
    target = torch.randint(0, C, (B,), device=DEV)
    ldd = (C + 7) // 8 * 8
    dl = torch.empty(B, ldd, dtype=torch.bfloat16, device=DEV)
    meter = torch.zeros(4, device=DEV)
    loss = torch.zeros(1, device=DEV)
    ops.cross_entropy(logits, target, dl, ldd, meter, loss)
    lr = logits.clone().requires_grad_(True)
    ref = F.cross_entropy(lr, target)
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-4
    assert rel_err(dl[:, :C], lr.grad) < 1e-2
    assert torch.count_nonzero(dl[:, C:]) == 0
    assert abs(float(meter[0]) - float(ref) * B) < 1e-2
    assert int(meter[1]) == int((logits.argmax(1) == target).sum())
    assert int(meter[2]) == B


@pytest.mark.parametrize("B,C", [(64, 3), (33, 10)])
def test_cross_entropy_class_weights(B, C):
    """The reference's optional weighted loss (distributedVggf.py:164-166): torch semantics,
    loss = sum_i w[y_i] nll_i / sum_i w[y_i], and the gradient that goes with it."""
    ops = _ops()
    torch.manual_seed(1)
    logits = torch.randn(B, C, device=DEV) * 2
    target = torch.randint(0, C, (B,), device=DEV)
    w = torch.rand(C, device=DEV) + 0.1
    ldd = (C + 7) // 8 * 8
    dl = torch.empty(B, ldd, dtype=torch.bfloat16, device=DEV)
    meter = torch.zeros(4, device=DEV)
    loss = torch.zeros(1, device=DEV)
    ops.cross_entropy(logits, target, dl, ldd, meter, loss, class_weights=w)
    lr = logits.clone().requires_grad_(True)
    ref = F.cross_entropy(lr, target, weight=w)
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-4
    assert rel_err(dl[:, :C], lr.grad) < 1e-2
    assert abs(float(meter[0]) - float(ref) * B) < 1e-2 and int(meter[2]) == B


def test_adam_matches_torch():
    ops = _ops()
    torch.manual_seed(0)
    n = 4096 * 5
    p0 = torch.randn(n, device=DEV)
    p = p0.clone()
    m = torch.zeros(n, device=DEV)
    v = torch.zeros(n, device=DEV)
    shadow = torch.empty(n, dtype=torch.bfloat16, device=DEV)
    ref_p = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=1e-3)
    for step in range(1, 4):
        g = torch.randn(n, device=DEV)
        gbuf = g.clone()
        ops.adam_step(p, m, v, g32=gbuf, shadow=shadow, lr=1e-3, step=step, zero=gbuf)
        assert torch.count_nonzero(gbuf) == 0
        ref_p.grad = g.clone()
        opt.step()
    assert rel_err(p, ref_p.detach()) < 1e-5
    assert rel_err(shadow, ref_p.detach()) < 1e-2
    # bf16 gradient source
    g16 = bf(torch.randn(n, device=DEV))
    p2, m2, v2 = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    ops.adam_step(p2, m2, v2, g16=g16, lr=1e-3, step=1)
    r2 = p0.clone().requires_grad_(True)
    o2 = torch.optim.Adam([r2], lr=1e-3)
    r2.grad = g16.float()
    o2.step()
    assert rel_err(p2, r2.detach()) < 1e-5


def test_sgd_matches_torch():
    ops = _ops()
    torch.manual_seed(0)
    n = 4096
    p0 = torch.randn(n, device=DEV)
    p, mom = p0.clone(), torch.zeros(n, device=DEV)
    r = p0.clone().requires_grad_(True)
    opt = torch.optim.SGD([r], lr=0.01, momentum=0.9)
    for step in range(3):
        g = torch.randn(n, device=DEV)
        ops.sgd_step(p, mom, g32=g, lr=0.01, momentum=0.9, first=(step == 0))
        r.grad = g.clone()
        opt.step()
    assert rel_err(p, r.detach()) < 1e-5


def test_augment_matches_reference():
    ops = _ops()
    from distributed_vgg_f_b200.data import transforms as T
    from distributed_vgg_f_b200.data.synthetic import synthetic_uint8_batch

    imgs, _ = synthetic_uint8_batch(4, 128, seed=3)
    src = torch.from_numpy(imgs).to(DEV)
    g = torch.Generator().manual_seed(5)
    params = T.sample_train_params(4, 128, 128, g)
    ref = T.augment_reference(src, params, (256, 256))            # [4,3,224,224] fp32
    out = torch.empty(4, 224, 224, 8, dtype=torch.bfloat16, device=DEV)
    ops.augment(src, params.to(DEV), out, (256, 256), mode="nhwc", pad=8)
    got = out[..., :3].float().permute(0, 3, 1, 2)
    # bf16 rounding of values up to ~2.6 plus rare one-pixel floor() disagreements at exact ties
    diff = (got - ref).abs()
    assert float(diff.median()) < 5e-3
    assert float((diff > 0.05).float().mean()) < 2e-3
    # im2col mode equals unfolding the NHWC result
    col = torch.empty(4 * 224 * 224, 64, dtype=torch.bfloat16, device=DEV)
    ops.augment(src, params.to(DEV), col, (256, 256), mode="im2col", pad=64)
    unf = F.unfold(got, 3, padding=1).view(4, 3, 9, 224 * 224).permute(0, 3, 2, 1).reshape(-1, 27)
    assert torch.equal(col[:, :27].float(), unf)
    assert torch.count_nonzero(col[:, 27:]) == 0


def test_augment_kernel_matches_torchvision_chain():
    """The sm_100a input kernel against torchvision's OWN transform ops on PIL images with the same draws
    (distributedVggf.py:88-95) -- not against this repo's torch oracle."""
    ops = _ops()
    import numpy as np

    from distributed_vgg_f_b200.data import transforms as T
    from distributed_vgg_f_b200.data.synthetic import synthetic_uint8_batch

    imgs, _ = synthetic_uint8_batch(4, 128, seed=13)
    rng = np.random.default_rng(1)
    imgs = np.clip(imgs.astype(np.int16) + rng.integers(-20, 20, imgs.shape), 0, 255).astype(np.uint8)
    params = T.sample_train_params(4, 128, 128, torch.Generator().manual_seed(9))
    params[0, 6], params[1, 6] = 0.0, 1.0
    out = torch.empty(4, 224, 224, 4, dtype=torch.bfloat16, device=DEV)
    ops.augment(torch.from_numpy(imgs).to(DEV), params.to(DEV), out, (256, 256), mode="nhwc", pad=4)
    got = out[..., :3].float().permute(0, 3, 1, 2).cpu()
    for i in range(4):
        theirs = T.torchvision_chain_fixed(imgs[i], params[i])
        d = (got[i] - theirs).abs()
        # PIL's uint8 rounding of the resized image (0.009 after Normalize) + bf16 output rounding (2^-8 relative)
        assert float((d < 0.04).float().mean()) > 0.995, (i, float((d < 0.04).float().mean()), float(d.max()))
        assert float(d.mean()) < 0.012, (i, float(d.mean()))


# ------------------------------------------------------------------ fused head (K-FUN2+CE)
@pytest.mark.parametrize("B,C,K,weighted", [(64, 3, 512, False), (37, 5, 96, False), (64, 3, 512, True), (1, 8, 1024, False)])
def test_head_ce_fused_matches_torch(B, C, K, weighted):
    """Last Linear + cross-entropy + metrics + the layer's backward in one launch (loss_optim.cu::head_ce_kernel)
    against plain torch fp32 with the same rounding points (bf16 operands, dlogits rounded to bf16)."""
    from distributed_vgg_f_b200 import ops

    Cx = ops.require()
    g = torch.Generator(device="cuda").manual_seed(B * 131 + C)
    h = torch.relu(torch.randn(B, K, device="cuda", generator=g)).to(torch.bfloat16)       # post-ReLU: has zeros
    W = (torch.randn(C, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(C, device="cuda", generator=g) * 0.1
    y = torch.randint(0, C, (B,), device="cuda", generator=g)
    cw = (torch.rand(C, device="cuda", generator=g) + 0.5) if weighted else None
    logits = torch.empty(B, C, device="cuda")
    dl = torch.full((B, 8), 7.0, device="cuda", dtype=torch.bfloat16)
    dW = torch.empty(C, K, device="cuda")
    db = torch.ones(C, device="cuda")                   # accumulated into
    dh = torch.empty(B, K, device="cuda", dtype=torch.bfloat16)
    meter = torch.zeros(4, device="cuda")
    loss = torch.zeros(1, device="cuda")
    Cx.head_ce(h, W, bias, y, logits, dl, 8, dW, db, dh, 2.0, True, meter, loss, cw)

    z = h.float() @ W.float().t() + bias
    ref_loss = torch.nn.functional.cross_entropy(z, y, weight=cw)
    wt = cw[y] if weighted else torch.ones(B, device="cuda")
    norm = wt.sum() if weighted else torch.tensor(float(B), device="cuda")
    g_ref = ((torch.softmax(z, 1) - torch.nn.functional.one_hot(y, C).float()) * wt[:, None] / norm)
    g_bf = g_ref.to(torch.bfloat16).float()
    torch.testing.assert_close(logits, z, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(loss[0], ref_loss, rtol=1e-4, atol=1e-5)
    assert float(meter[1]) == float((z.argmax(1) == y).sum()) and float(meter[2]) == B
    torch.testing.assert_close(meter[0], ref_loss * B, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dl[:, :C].float(), g_bf, rtol=0, atol=1e-6 + 8e-3 * float(g_ref.abs().max()))
    assert float(dl[:, C:].float().abs().max()) == 0.0 if C < 8 else True
    used = dl[:, :C].float()                            # what the kernel itself rounded
    torch.testing.assert_close(dW, used.t() @ h.float(), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(db, 1.0 + used.sum(0), rtol=1e-4, atol=1e-6)
    dh_ref = ((used @ W.float()) * (h.float() > 0) * 2.0).to(torch.bfloat16)
    torch.testing.assert_close(dh.float(), dh_ref.float(), rtol=1e-2, atol=1e-6)
    # evaluation form: forward + metrics only
    logits2 = torch.empty_like(logits)
    Cx.head_ce(h, W, bias, y, logits2, None, 0, None, None, None, 1.0, False, None, loss, None)
    torch.testing.assert_close(logits2, z, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(loss[0], torch.nn.functional.cross_entropy(z, y), rtol=1e-4, atol=1e-5)
