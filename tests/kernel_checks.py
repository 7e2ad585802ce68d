"""Per-kernel numerical checks of libpe_b200.so against plain torch references (GPU only).

Each check returns a list of (name, error, tolerance).  `tests/test_kernels_gpu.py` asserts on them;
`python tests/kernel_checks.py` prints the whole table without stopping at the first failure
(handy for one-shot GPU sessions).

Tensor-core kernels are compared against an fp64 reference computed from TF32-rounded inputs, so
the tolerance only has to cover fp32 accumulation order (1e-4 relative to the output scale).
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rgb-proprioceptive-pose-estimator_b200"))

from pe_b200 import native  # noqa: E402

DEV = "cuda"


def tf32(x):
    """Round fp32 to TF32 (10-bit mantissa), round-to-nearest, ties away from zero (cvt.rna)."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def relerr(a, b):
    a = a.double()
    b = b.double()
    denom = b.abs().max().clamp_min(1e-30)
    return float((a - b).abs().max() / denom)


def S():
    return native.stream_ptr()


def P(t):
    return native.ptr(t)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


# ------------------------------------------------------------------------------------------------
def check_linear(M, N, K, relu=False, bias=True, ldpad=0):
    L = native.lib()
    g = torch.Generator(device=DEV).manual_seed(M * 131 + N * 7 + K)
    ldx = K + ldpad
    x = tf32(torch.randn(M, ldx, device=DEV, generator=g))
    w = tf32(torch.randn(N, ldx, device=DEV, generator=g))
    b = torch.randn(N, device=DEV, generator=g) if bias else None
    ldy = ((N + 3) // 4) * 4
    y = torch.full((M, ldy), float("nan"), device=DEV)
    L.pe_linear_fwd(P(x), ldx, P(w), ldx, P(b), None, P(y), ldy, M, N, K, int(relu), 0, 0, None, S())
    ref = x[:, :K].double() @ w[:, :K].double().t()
    if bias:
        ref = ref + b.double()
    if relu:
        ref = ref.clamp_min(0)
    return [("linear_fwd M%d N%d K%d relu%d" % (M, N, K, relu), relerr(y[:, :N], ref), 1e-4)]


def check_linear_acc(M, N, K):
    L = native.lib()
    g = torch.Generator(device=DEV).manual_seed(5)
    x = tf32(torch.randn(M, K, device=DEV, generator=g))
    w = tf32(torch.randn(N, K, device=DEV, generator=g))
    y0 = torch.randn(M, N, device=DEV, generator=g)
    y = y0.clone()
    L.pe_linear_fwd(P(x), K, P(w), K, None, None, P(y), N, M, N, K, 0, 1, 0, None, S())
    ref = y0.double() + x.double() @ w.double().t()
    return [("linear_fwd accumulate M%d N%d K%d" % (M, N, K), relerr(y, ref), 1e-4)]


def check_linear_wgrad(M, N, K):
    L = native.lib()
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    ldx = ((K + 3) // 4) * 4
    lddy = ((N + 3) // 4) * 4
    x = tf32(torch.randn(M, ldx, device=DEV, generator=g))
    dy = tf32(torch.randn(M, lddy, device=DEV, generator=g))
    dw = torch.full((N, ldx), float("nan"), device=DEV)
    L.pe_linear_wgrad(P(x), ldx, P(dy), lddy, P(dw), ldx, M, N, K, S())
    ref = dy[:, :N].double().t() @ x[:, :K].double()
    return [("linear_wgrad M%d N%d K%d" % (M, N, K), relerr(dw[:, :K], ref), 1e-4)]


def pack(w, L=None):
    L = L or native.lib()
    Cout, Cin, R, S_ = w.shape
    tck = torch.empty(R * S_, Cout, Cin, device=DEV)
    tkc = torch.empty(R * S_, Cin, Cout, device=DEV)
    L.pe_pack_conv_weight(P(w.contiguous()), P(tck), P(tkc), Cout, Cin, R, S_, 0, S())
    return tck, tkc


def pack_maskbits(keep):
    """Reference packing of a boolean tensor into the kernels' bit layout: float4 index i -> bit (i & 31) of
    words [(i >> 5) * 4 + component]."""
    flat = keep.reshape(-1).to(torch.int64)
    n4 = flat.numel() // 4
    n32 = (n4 + 31) // 32
    padded = torch.zeros(n32 * 32 * 4, device=keep.device, dtype=torch.int64)
    padded[:flat.numel()] = flat
    v = padded.reshape(n32, 32, 4)                                   # [group][float4 in group][component]
    w = (v << torch.arange(32, device=keep.device).reshape(1, 32, 1)).sum(1)     # [group][component]
    w = torch.where(w >= 2 ** 31, w - 2 ** 32, w)
    return w.reshape(-1).to(torch.int32).contiguous()


def check_conv(B, H, W, Cin, Cout, k, stride):
    L = native.lib()
    pad = (k - 1) // 2
    g = torch.Generator(device=DEV).manual_seed(B + H + Cin + Cout + k + stride)
    x = tf32(torch.randn(B, Cin, H, W, device=DEV, generator=g))
    w = tf32(torch.randn(Cout, Cin, k, k, device=DEV, generator=g) / (Cin * k * k) ** 0.5)
    xd = x.double().requires_grad_(True)
    wd = w.double().requires_grad_(True)
    yd = F.conv2d(xd, wd, stride=stride, padding=pad)
    Ho, Wo = yd.shape[2], yd.shape[3]
    dy = tf32(torch.randn(B, Cout, Ho, Wo, device=DEV, generator=g))
    gx, gw = torch.autograd.grad(yd, (xd, wd), dy.double())

    tag = "B%d %dx%d %d->%d k%d s%d" % (B, H, W, Cin, Cout, k, stride)
    out = []
    tck, tkc = pack(w)
    # pack round trip
    out.append(("pack tck " + tag, relerr(tck, w.permute(2, 3, 0, 1).reshape(k * k, Cout, Cin)), 0.0))
    out.append(("pack tkc " + tag, relerr(tkc, w.permute(2, 3, 1, 0).reshape(k * k, Cin, Cout)), 0.0))

    x_n = nhwc(x)
    y = torch.full((B, Ho, Wo, Cout), float("nan"), device=DEV)
    stats = torch.zeros(2 * Cout, device=DEV, dtype=torch.float64)
    L.pe_conv2d_fwd(P(x_n), P(tck), P(y), B, H, W, Cin, Cout, k, k, stride, pad, None, None, None, 0, 0,
                    P(stats), S())
    y_ref = nhwc(yd.detach())
    out.append(("conv_fwd " + tag, relerr(y, y_ref), 1e-4))
    s_ref = torch.cat([y_ref.sum((0, 1, 2)), (y_ref * y_ref).sum((0, 1, 2))])
    out.append(("conv_fwd stats " + tag, relerr(stats, s_ref), 1e-4))

    dy_n = nhwc(dy)
    dx = torch.full((B, H, W, Cin), float("nan"), device=DEV)
    L.pe_conv2d_dgrad(P(dy_n), P(tkc), P(dx), B, H, W, Cin, Cout, k, k, stride, pad, None, None, S())
    out.append(("conv_dgrad " + tag, relerr(dx, nhwc(gx)), 1e-4))
    if stride == 1 and Cin % 32 == 0:
        # residual epilogue: dx = dgrad + res (plain) and dgrad + res * mask (bit mask of a residual join)
        res = torch.randn(B, H, W, Cin, device=DEV, generator=g)
        L.pe_conv2d_dgrad(P(dy_n), P(tkc), P(dx), B, H, W, Cin, Cout, k, k, stride, pad, P(res), None, S())
        out.append(("conv_dgrad +res " + tag, relerr(dx, nhwc(gx) + res.double()), 1e-4))
        keep = torch.rand(B, H, W, Cin, device=DEV, generator=g) > 0.5
        bits = pack_maskbits(keep)
        L.pe_conv2d_dgrad(P(dy_n), P(tkc), P(dx), B, H, W, Cin, Cout, k, k, stride, pad, P(res), P(bits), S())
        out.append(("conv_dgrad +masked res " + tag, relerr(dx, nhwc(gx) + res.double() * keep), 1e-4))

    dw = torch.full((k * k, Cout, Cin), float("nan"), device=DEV)
    L.pe_conv2d_wgrad(P(x_n), P(dy_n), P(dw), B, H, W, Cin, Cout, k, k, stride, pad, S())
    out.append(("conv_wgrad " + tag, relerr(dw, gw.permute(2, 3, 0, 1).reshape(k * k, Cout, Cin)), 1e-4))
    dw_oihw = torch.empty(Cout, Cin, k, k, device=DEV)
    L.pe_unpack_conv_wgrad(P(dw), P(dw_oihw), Cout, Cin, k, k, 0, S())
    out.append(("unpack wgrad " + tag, relerr(dw_oihw, gw), 1e-4))
    return out


def check_conv_dgrad_bn(B, H, W, Cin, Cout, k, stride):
    """pe_conv2d_dgrad_bn: dx as the plain dgrad, plus the BatchNorm backward sums of the activation dx belongs to
    (sum g, sum g * xhat with g = dx * (y * scale + shift > 0)) accumulated in the epilogue from the TMA-prefetched y."""
    L = native.lib()
    pad = (k - 1) // 2
    g = torch.Generator(device=DEV).manual_seed(7 * B + H + Cin + Cout + k + stride)
    w = tf32(torch.randn(Cout, Cin, k, k, device=DEV, generator=g) / (Cout * k * k) ** 0.5)
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    dy = tf32(torch.randn(B, Cout, Ho, Wo, device=DEV, generator=g))
    gx = torch.nn.grad.conv2d_input((B, Cin, H, W), w.double(), dy.double(), stride=stride, padding=pad)
    y = torch.randn(B, H, W, Cin, device=DEV, generator=g)
    sc = torch.rand(Cin, device=DEV, generator=g) + 0.5
    sh = torch.randn(Cin, device=DEV, generator=g) * 0.3
    mean = torch.randn(Cin, device=DEV, generator=g) * 0.2
    invstd = torch.rand(Cin, device=DEV, generator=g) + 0.5
    _, tkc = pack(w)
    dx = torch.full((B, H, W, Cin), float("nan"), device=DEV)
    sums = torch.zeros(2 * Cin, device=DEV, dtype=torch.float64)
    L.pe_conv2d_dgrad_bn(P(nhwc(dy)), P(tkc), P(dx), B, H, W, Cin, Cout, k, k, stride, pad, P(y), P(sc), P(sh), P(mean),
                         P(invstd), P(sums), S())
    tag = "B%d %dx%d %d->%d k%d s%d" % (B, H, W, Cin, Cout, k, stride)
    gx_n = nhwc(gx)
    mask = (torch.addcmul(sh, y, sc) > 0).double()          # same fp32 expression as the kernel's fmaf up to an ulp
    gm = gx_n * mask
    ref = torch.cat([gm.sum((0, 1, 2)), (gm * (y.double() - mean.double()) * invstd.double()).sum((0, 1, 2))])
    return [("conv_dgrad_bn dx " + tag, relerr(dx, gx_n), 1e-4), ("conv_dgrad_bn sums " + tag, relerr(sums, ref), 2e-4)]


def check_conv_fused_eval(B, H, W, Cin, Cout, k):
    """eval-mode epilogue: relu(acc*scale + shift + residual)."""
    L = native.lib()
    pad = (k - 1) // 2
    g = torch.Generator(device=DEV).manual_seed(11)
    x = tf32(torch.randn(B, Cin, H, W, device=DEV, generator=g))
    w = tf32(torch.randn(Cout, Cin, k, k, device=DEV, generator=g) / (Cin * k * k) ** 0.5)
    sc = torch.rand(Cout, device=DEV, generator=g) + 0.5
    sh = torch.randn(Cout, device=DEV, generator=g)
    res = torch.randn(B, H, W, Cout, device=DEV, generator=g)
    tck, _ = pack(w)
    y = torch.full((B, H, W, Cout), float("nan"), device=DEV)
    L.pe_conv2d_fwd(P(nhwc(x)), P(tck), P(y), B, H, W, Cin, Cout, k, k, 1, pad, P(sc), P(sh), P(res), 1, 0, None,
                    S())
    ref = nhwc(F.conv2d(x.double(), w.double(), padding=pad)) * sc.double() + sh.double() + res.double()
    return [("conv_fwd fused eval B%d %dx%d %d->%d k%d" % (B, H, W, Cin, Cout, k), relerr(y, ref.clamp_min(0)), 1e-4)]


def check_stem(B):
    """7x7/2 stem through im2col + GEMM."""
    L = native.lib()
    g = torch.Generator(device=DEV).manual_seed(3)
    img = tf32(torch.randn(B, 3, 224, 224, device=DEV, generator=g))
    w = tf32(torch.randn(64, 3, 7, 7, device=DEV, generator=g) / 12.0)
    ldc = 160
    col = torch.full((B * 112 * 112, ldc), float("nan"), device=DEV)
    L.pe_im2col_stem(P(img), P(col), B, 3, 224, 224, 7, 7, 2, 3, ldc, 0, S())
    wp = torch.zeros(64, ldc, device=DEV)
    wp[:, :147] = w.reshape(64, 147)
    y = torch.full((B * 112 * 112, 64), float("nan"), device=DEV)
    L.pe_linear_fwd(P(col), ldc, P(wp), ldc, None, None, P(y), 64, B * 112 * 112, 64, ldc, 0, 0, 0, None, S())
    ref = nhwc(F.conv2d(img.double(), w.double(), stride=2, padding=3)).reshape(-1, 64)
    out = [("stem conv (im2col+gemm) B%d" % B, relerr(y, ref), 1e-4)]
    unf = F.unfold(img, kernel_size=7, stride=2, padding=3).transpose(1, 2).reshape(-1, 147)
    out.append(("im2col B%d" % B, relerr(col[:, :147], unf), 0.0))
    out.append(("im2col pad B%d" % B, float(col[:, 147:].abs().max()), 0.0))
    return out


def check_stem_s2d(B, H=224, W=224):
    """7x7/2 stem as a 4x4/1 convolution over the space-to-depth image (no im2col matrix): operand pack, weight pack /
    gradient unpack round trip, forward with batch statistics, fused eval epilogue, weight gradient."""
    L = native.lib()
    g = torch.Generator(device=DEV).manual_seed(5 + B + H)
    img = tf32(torch.randn(B, 3, H, W, device=DEV, generator=g))
    w = tf32(torch.randn(64, 3, 7, 7, device=DEV, generator=g) / 12.0)
    Ho, Wo, Hs, Ws = H // 2, W // 2, H // 2 + 3, W // 2 + 3
    s2d = torch.full((B, Hs, Ws, 12), float("nan"), device=DEV)
    L.pe_stem_s2d_pack(P(img), P(s2d), B, H, W, 0, S())
    ref_s2d = torch.zeros(B, Hs, Ws, 12, device=DEV)
    blk = img.reshape(B, 3, Ho, 2, Wo, 2).permute(0, 2, 4, 3, 5, 1).reshape(B, Ho, Wo, 12)    # (a, b, c) fastest
    ref_s2d[:, 2:2 + Ho, 2:2 + Wo] = blk
    tag = "B%d %dx%d" % (B, H, W)
    out = [("stem s2d pack " + tag, relerr(s2d, ref_s2d), 0.0)]
    w_s2d = torch.full((4, 64, 64), float("nan"), device=DEV)
    L.pe_stem_pack_weight(P(w), P(w_s2d), 64, 0, S())
    back = torch.zeros(64, 3, 7, 7, device=DEV)
    L.pe_stem_unpack_wgrad(P(w_s2d), P(back), 64, S())
    out.append(("stem weight pack/unpack round trip " + tag, relerr(back, w), 0.0))
    out.append(("stem weight pack zero columns " + tag, float(w_s2d[:, :, 48:].abs().max()), 0.0))
    wd = w.double().requires_grad_(True)
    yd = F.conv2d(img.double(), wd, stride=2, padding=3)
    y_ref = nhwc(yd.detach()).reshape(-1, 64)
    y = torch.full((B * Ho * Wo, 64), float("nan"), device=DEV)
    stats = torch.zeros(128, device=DEV, dtype=torch.float64)
    L.pe_stem_conv_fwd(P(s2d), P(w_s2d), P(y), B, H, W, 64, None, None, 0, 0, P(stats), S())
    out.append(("stem conv fwd (s2d) " + tag, relerr(y, y_ref), 1e-4))
    out.append(("stem conv fwd stats " + tag, relerr(stats, torch.cat([y_ref.sum(0), (y_ref * y_ref).sum(0)])), 1e-4))
    sc = torch.rand(64, device=DEV, generator=g) + 0.5
    sh = torch.randn(64, device=DEV, generator=g)
    L.pe_stem_conv_fwd(P(s2d), P(w_s2d), P(y), B, H, W, 64, P(sc), P(sh), 1, 0, None, S())
    out.append(("stem conv fwd fused eval " + tag, relerr(y, (y_ref * sc.double() + sh.double()).clamp_min(0)), 1e-4))
    if H % 16 == 0 and W % 16 == 0:
        dy = tf32(torch.randn(B, 64, Ho, Wo, device=DEV, generator=g))
        gw, = torch.autograd.grad(yd, wd, dy.double())
        dw_s2d = torch.full((4, 64, 64), float("nan"), device=DEV)
        L.pe_stem_conv_wgrad(P(s2d), P(nhwc(dy)), P(dw_s2d), B, H, W, 64, S())
        dw = torch.zeros(64, 3, 7, 7, device=DEV)
        L.pe_stem_unpack_wgrad(P(dw_s2d), P(dw), 64, S())
        out.append(("stem conv wgrad (s2d) " + tag, relerr(dw, gw), 1e-4))
    return out


def check_bn(Pn, C, relu=True, residual=True):
    L = native.lib()
    g = torch.Generator(device=DEV).manual_seed(C + Pn)
    y = torch.randn(Pn, C, device=DEV, generator=g) * 2 + 0.5
    gamma = torch.rand(C, device=DEV, generator=g) + 0.5
    beta = torch.randn(C, device=DEV, generator=g)
    res = torch.randn(Pn, C, device=DEV, generator=g) if residual else None
    rm = torch.zeros(C, device=DEV)
    rv = torch.ones(C, device=DEV)
    out = []
    tag = "P%d C%d" % (Pn, C)
    stats = torch.zeros(2 * C, device=DEV, dtype=torch.float64)
    L.pe_bn_stats(P(y), Pn, C, P(stats), S())
    out.append(("bn_stats " + tag, relerr(stats, torch.cat([y.double().sum(0), (y.double() ** 2).sum(0)])), 1e-5))
    scale = torch.empty(C, device=DEV)
    shift = torch.empty(C, device=DEV)
    mean = torch.empty(C, device=DEV)
    invstd = torch.empty(C, device=DEV)
    L.pe_bn_finalize(P(stats), P(gamma), P(beta), P(rm), P(rv), P(scale), P(shift), P(mean), P(invstd), Pn, 0.1, 1e-5,
                     C, S())
    o = torch.empty(Pn, C, device=DEV)
    L.pe_bn_apply(P(y), P(scale), P(shift), P(res), P(o), Pn, C, int(relu), 0, S())

    # torch reference (fp64 autograd through batch_norm)
    yd = y.double().requires_grad_(True)
    gd = gamma.double().requires_grad_(True)
    bd = beta.double().requires_grad_(True)
    rd = res.double().requires_grad_(True) if residual else None
    rm_ref = torch.zeros(C, device=DEV, dtype=torch.float64)
    rv_ref = torch.ones(C, device=DEV, dtype=torch.float64)
    z = F.batch_norm(yd.t().reshape(1, C, Pn), rm_ref, rv_ref, gd, bd, True, 0.1, 1e-5).reshape(C, Pn).t()
    if residual:
        z = z + rd
    if relu:
        z = z.clamp_min(0)
    out.append(("bn_apply " + tag, relerr(o, z.detach()), 1e-5))
    out.append(("bn running_mean " + tag, relerr(rm, rm_ref), 1e-5))
    out.append(("bn running_var " + tag, relerr(rv, rv_ref), 1e-5))
    dout = torch.randn(Pn, C, device=DEV, generator=g)
    ins = (yd, gd, bd) + ((rd,) if residual else ())
    grads = torch.autograd.grad(z, ins, dout.double())
    sums = torch.zeros(2 * C, device=DEV, dtype=torch.float64)
    dout_a = dout * 0.25
    dout_b = dout - dout_a
    use_mask = relu and not residual
    o_arg = None if use_mask else o
    L.pe_bn_bwd_reduce(P(dout_a), P(dout_b), P(o_arg), P(y), P(mean), P(invstd), P(scale), P(shift), None, P(sums),
                       Pn, C, int(relu), S())
    dy = torch.empty(Pn, C, device=DEV)
    dres = torch.empty(Pn, C, device=DEV) if residual else None
    dgamma = torch.empty(C, device=DEV)
    dbeta = torch.empty(C, device=DEV)
    L.pe_bn_bwd_apply(P(dout_a), P(dout_b), P(o_arg), P(y), P(mean), P(invstd), P(gamma), P(scale), P(shift), None,
                      P(sums), P(dy), P(dres), 0,
                      P(dgamma), P(dbeta), 0, Pn, C, int(relu), 0, S())
    out.append(("bn_bwd dy " + tag, relerr(dy, grads[0]), 1e-4))
    out.append(("bn_bwd dgamma " + tag, relerr(dgamma, grads[1]), 1e-4))
    out.append(("bn_bwd dbeta " + tag, relerr(dbeta, grads[2]), 1e-4))
    if residual:
        out.append(("bn_bwd dres " + tag, relerr(dres, grads[3]), 1e-6))
    # ---- fused train-mode forward (statistics finalise + running stats + counter + apply + mask bits) ----
    rm2 = torch.zeros(C, device=DEV)
    rv2 = torch.ones(C, device=DEV)
    nbt = torch.full((1,), 7, device=DEV, dtype=torch.int64)
    sc2, sh2, mean2, invstd2 = (torch.full((C,), float("nan"), device=DEV) for _ in range(4))
    o2 = torch.full((Pn, C), float("nan"), device=DEV)
    n4 = Pn * (C // 4)
    bits = torch.zeros((n4 + 31) // 32 * 4, device=DEV, dtype=torch.int32)
    L.pe_bn_train_apply(P(y), P(stats), P(gamma), P(beta), P(rm2), P(rv2), P(nbt), P(sc2), P(sh2), P(mean2),
                        P(invstd2), P(res), P(o2), P(bits), Pn, C, 0.1, 1e-5, int(relu), 0, S())
    out.append(("bn_train_apply out " + tag, relerr(o2, o), 0.0))
    out.append(("bn_train_apply scale/shift/mean/invstd " + tag,
                max(relerr(sc2, scale), relerr(sh2, shift), relerr(mean2, mean), relerr(invstd2, invstd)), 0.0))
    out.append(("bn_train_apply running stats " + tag, max(relerr(rm2, rm), relerr(rv2, rv)), 0.0))
    out.append(("bn_train_apply counter " + tag, float(abs(int(nbt.item()) - 8)), 0.0))
    ref_bits = pack_maskbits(o2 > 0)
    out.append(("bn_train_apply mask bits " + tag, float((bits != ref_bits).sum()), 0.0))
    if C % 32 == 0 and relu:
        # backward with the mask taken from the bits (relu flag off, `out` not read): same result as above
        sums2 = torch.zeros(2 * C, device=DEV, dtype=torch.float64)
        L.pe_bn_bwd_reduce(P(dout), None, None, P(y), P(mean), P(invstd), P(scale), P(shift), P(bits), P(sums2), Pn, C,
                           0, S())
        dy2 = torch.empty(Pn, C, device=DEV)
        L.pe_bn_bwd_apply(P(dout), None, None, P(y), P(mean), P(invstd), P(gamma), P(scale), P(shift), P(bits),
                          P(sums2), P(dy2), None, 0, P(dgamma), P(dbeta), 0, Pn, C, 0, 0, S())
        out.append(("bn_bwd (mask bits) dy " + tag, relerr(dy2, grads[0]), 1e-4))
        out.append(("bn_bwd (mask bits) dgamma " + tag, relerr(dgamma, grads[1]), 1e-4))
    # eval-mode finalize
    L.pe_bn_finalize(None, P(gamma), P(beta), P(rm), P(rv), P(scale), P(shift), None, None, Pn, 0.1, 1e-5, C, S())
    sc_ref = gamma.double() / torch.sqrt(rv.double() + 1e-5)
    out.append(("bn eval scale " + tag, relerr(scale, sc_ref), 1e-6))
    out.append(("bn eval shift " + tag, relerr(shift, beta.double() - rm.double() * sc_ref), 1e-5))
    return out


def check_pools(B):
    L = native.lib()
    g = torch.Generator(device=DEV).manual_seed(B)
    out = []
    x = torch.randn(B, 64, 112, 112, device=DEV, generator=g).clamp_min(0)  # post-ReLU like the trunk
    xd = x.double().requires_grad_(True)
    yd = F.max_pool2d(xd, 3, 2, 1)
    dyv = torch.randn_like(yd)
    (gx,) = torch.autograd.grad(yd, xd, dyv)
    xn = nhwc(x)
    y = torch.empty(B, 56, 56, 64, device=DEV)
    am = torch.empty(B, 56, 56, 64, device=DEV, dtype=torch.uint8)
    L.pe_maxpool3x3s2_fwd(P(xn), P(y), P(am), B, 112, 112, 64, S())
    out.append(("maxpool fwd B%d" % B, relerr(y, nhwc(yd.detach())), 0.0))
    dx = torch.full((B, 112, 112, 64), float("nan"), device=DEV)
    L.pe_maxpool3x3s2_bwd(P(nhwc(dyv.float())), None, P(am), P(dx), 0, B, 112, 112, 64, None, 0, None, None, S())
    # ties only happen at 0 where the ReLU mask kills the gradient anyway -> compare on x > 0
    mask = (xn > 0).double()
    out.append(("maxpool bwd B%d" % B, relerr(dx.double() * mask, nhwc(gx) * mask), 1e-6))

    x7 = torch.randn(B, 7, 7, 2048, device=DEV, generator=g)
    ya = torch.empty(B, 2048, device=DEV)
    L.pe_avgpool_fwd(P(x7), P(ya), 2048, B, 49, 2048, 0, S())
    out.append(("avgpool fwd B%d" % B, relerr(ya, x7.double().mean((1, 2))), 1e-6))
    dya = torch.randn(B, 2048, device=DEV, generator=g)
    dx7 = torch.empty_like(x7)
    L.pe_avgpool_bwd(P(dya), 2048, P(dx7), B, 49, 2048, S())
    out.append(("avgpool bwd B%d" % B, relerr(dx7, (dya.double() / 49)[:, None, None, :].expand(B, 7, 7, 2048)), 1e-6))

    # auxiliary branch
    a1 = torch.randn(B, 64, 112, 112, device=DEV, generator=g).clamp_min(0)
    w = torch.randn(1, 64, 1, 1, device=DEV, generator=g) * 0.2
    b = torch.randn(1, device=DEV, generator=g)
    a1d = a1.double().requires_grad_(True)
    wd = w.double().requires_grad_(True)
    bd = b.double().requires_grad_(True)
    ref = F.max_pool2d(F.conv2d(a1d, wd, bd), 2).flatten(1)
    dref = torch.randn_like(ref)
    ga, gw_, gb = torch.autograd.grad(ref, (a1d, wd, bd), dref)
    a1n = nhwc(a1)
    ldo = 3136 + 8
    o = torch.zeros(B, ldo, device=DEV)
    am2 = torch.empty(B * 56 * 56, device=DEV, dtype=torch.uint8)
    L.pe_aux_fwd(P(a1n), P(w), P(b), P(o), ldo, P(am2), B, 112, 112, 64, 0, S())
    out.append(("aux fwd B%d" % B, relerr(o[:, :3136], ref.detach()), 1e-5))
    da1 = torch.full((B, 112, 112, 64), float("nan"), device=DEV)
    dw = torch.zeros(64, device=DEV)
    db = torch.zeros(1, device=DEV)
    do = torch.zeros(B, ldo, device=DEV)
    do[:, :3136] = dref.float()
    L.pe_aux_bwd(P(do), ldo, P(am2), P(a1n), P(w), P(da1), 0, P(dw), P(db), B, 112, 112, 64, S())
    out.append(("aux bwd da1 B%d" % B, relerr(da1, nhwc(ga)), 1e-5))
    # fused stem backward: max-pool gradient + aux gradient written in one pass == sum of the two separate ones
    dsum = torch.full((B, 112, 112, 64), float("nan"), device=DEV)
    L.pe_maxpool3x3s2_bwd(P(nhwc(dyv.float())), None, P(am), P(dsum), 0, B, 112, 112, 64, P(do), ldo, P(am2), P(w), S())
    out.append(("maxpool bwd + aux term B%d" % B, relerr(dsum, dx.double() + nhwc(ga)), 1e-5))
    dw_only = torch.zeros(64, device=DEV)
    db_only = torch.zeros(1, device=DEV)
    L.pe_aux_bwd(P(do), ldo, P(am2), P(a1n), P(w), None, 0, P(dw_only), P(db_only), B, 112, 112, 64, S())
    out.append(("aux bwd dw (no da1) B%d" % B, relerr(dw_only, gw_.flatten()), 1e-4))
    out.append(("aux bwd dw B%d" % B, relerr(dw, gw_.flatten()), 1e-4))
    out.append(("aux bwd db B%d" % B, relerr(db, gb), 1e-4))
    return out


def check_stem_tail(B, rt):
    """pe_stem_post_train (bn1 + ReLU + max pool + aux branch in one pass) against the three separate kernels (bit for
    bit: same arithmetic) and against torch in float64; pe_aux_bwd_params against pe_aux_bwd on the materialised a1."""
    L = native.lib()
    g = torch.Generator(device=DEV).manual_seed(500 + B)
    H = W = 112
    C = 64
    y = torch.randn(B, H, W, C, device=DEV, generator=g) * 1.5 + 0.3
    gamma = torch.rand(C, device=DEV, generator=g) + 0.5
    beta = torch.randn(C, device=DEV, generator=g) * 0.3
    w = torch.randn(1, C, 1, 1, device=DEV, generator=g) * 0.2
    b = torch.randn(1, device=DEV, generator=g)
    Pn = B * H * W
    stats = torch.cat([y.double().sum((0, 1, 2)), (y.double() ** 2).sum((0, 1, 2))]).contiguous()
    ldo = 3136 + 8

    def buffers():
        return dict(rm=torch.zeros(C, device=DEV), rv=torch.ones(C, device=DEV),
                    nbt=torch.zeros(1, device=DEV, dtype=torch.int64), sc=torch.empty(C, device=DEV),
                    sh=torch.empty(C, device=DEV), mean=torch.empty(C, device=DEV), invstd=torch.empty(C, device=DEV),
                    pool=torch.empty(B, 56, 56, C, device=DEV), am=torch.empty(B, 56, 56, C, device=DEV, dtype=torch.uint8),
                    aux=torch.zeros(B, ldo, device=DEV), aam=torch.empty(B * 56 * 56, device=DEV, dtype=torch.uint8))
    f, u = buffers(), buffers()
    L.pe_stem_post_train(P(y), P(stats), P(gamma), P(beta), P(f["rm"]), P(f["rv"]), P(f["nbt"]), P(f["sc"]), P(f["sh"]),
                         P(f["mean"]), P(f["invstd"]), P(f["pool"]), P(f["am"]), P(w), P(b), P(f["aux"]), ldo, P(f["aam"]),
                         B, H, W, C, 0.1, 1e-5, rt, rt, S())
    a1 = torch.empty_like(y)
    L.pe_bn_train_apply(P(y), P(stats), P(gamma), P(beta), P(u["rm"]), P(u["rv"]), P(u["nbt"]), P(u["sc"]), P(u["sh"]),
                        P(u["mean"]), P(u["invstd"]), None, P(a1), None, Pn, C, 0.1, 1e-5, 1, rt, S())
    L.pe_maxpool3x3s2_fwd(P(a1), P(u["pool"]), P(u["am"]), B, H, W, C, S())
    L.pe_aux_fwd(P(a1), P(w), P(b), P(u["aux"]), ldo, P(u["aam"]), B, H, W, C, rt, S())
    out = []
    tag = "stem tail B%d rt%d " % (B, rt)
    for k in ("rm", "rv", "sc", "sh", "mean", "invstd", "pool"):
        out.append((tag + k + " == separate kernels", float((f[k] != u[k]).sum()), 0.0))
    out.append((tag + "nbt", float(abs(int(f["nbt"]) - 1)), 0.0))
    out.append((tag + "pool argmax == separate kernels", float((f["am"] != u["am"]).sum()), 0.0))
    # (the compiler contracts the dot product differently in the two kernels: last-bit differences, which the TF32
    # rounding of the result can turn into one TF32 ulp = 2^-10 of the value)
    out.append((tag + "aux vs separate kernels", relerr(f["aux"], u["aux"]), 1e-3 if rt else 1e-6))
    # the aux arg-max may only differ where two pixels of a window tie to rounding
    d = f["aam"] != u["aam"]
    out.append((tag + "aux argmax vs separate kernels", float(d.sum()), 2.0))
    if not rt:
        yd = y.double().permute(0, 3, 1, 2)
        a1d = F.relu(F.batch_norm(yd, None, None, gamma.double(), beta.double(), True, 0.1, 1e-5))
        out.append((tag + "pool vs torch", relerr(f["pool"], nhwc(F.max_pool2d(a1d, 3, 2, 1))), 1e-5))
        out.append((tag + "aux vs torch", relerr(f["aux"][:, :3136], F.max_pool2d(F.conv2d(a1d, w.double(), b.double()), 2).flatten(1)), 1e-5))
    # a no-aux launch leaves the pooled output unchanged
    f2 = buffers()
    L.pe_stem_post_train(P(y), P(stats), P(gamma), P(beta), None, None, None, P(f2["sc"]), P(f2["sh"]), P(f2["mean"]),
                         P(f2["invstd"]), P(f2["pool"]), None, None, None, None, 0, None, B, H, W, C, 0.1, 1e-5, rt, rt,
                         S())
    out.append((tag + "pool without aux / argmax / running stats", float((f2["pool"] != f["pool"]).sum()), 0.0))
    # aux conv gradients from y + scale / shift == from the materialised activation
    do = torch.zeros(B, ldo, device=DEV)
    do[:, :3136] = torch.randn(B, 3136, device=DEV, generator=g)
    dw_a, db_a, dw_b, db_b = (torch.zeros(n, device=DEV) for n in (C, 1, C, 1))
    L.pe_aux_bwd(P(do), ldo, P(u["aam"]), P(a1), P(w), None, 0, P(dw_a), P(db_a), B, H, W, C, S())
    L.pe_aux_bwd_params(P(do), ldo, P(u["aam"]), P(y), P(u["sc"]), P(u["sh"]), rt, P(dw_b), P(db_b), B, H, W, C, S())
    out.append((tag + "aux dw from y", relerr(dw_b, dw_a), 1e-5))
    out.append((tag + "aux db from y", relerr(db_b, db_a), 1e-5))
    return out


def check_lstm_cell(N, H):
    L = native.lib()
    g = torch.Generator(device=DEV).manual_seed(N + H)
    gx = torch.randn(N, 4 * H, device=DEV, generator=g)
    gh = torch.randn(N, 4 * H, device=DEV, generator=g)
    b1 = torch.randn(4 * H, device=DEV, generator=g)
    b2 = torch.randn(4 * H, device=DEV, generator=g)
    c0 = torch.randn(N, H, device=DEV, generator=g)
    gxd, ghd, c0d = (t.double().requires_grad_(True) for t in (gx, gh, c0))
    gates = gxd + ghd + b1.double() + b2.double()
    i, f, gg, o = gates.chunk(4, dim=1)
    c1 = torch.sigmoid(f) * c0d + torch.sigmoid(i) * torch.tanh(gg)
    h1 = torch.sigmoid(o) * torch.tanh(c1)
    c_out = torch.empty(N, H, device=DEV)
    h_out = torch.empty(N, H, device=DEV)
    act = torch.empty(N, 4 * H, device=DEV)
    L.pe_lstm_cell_fwd(P(gx), 4 * H, P(gh), 4 * H, P(b1), P(b2), P(c0), P(c_out), P(h_out), H, P(act), N, H, 0, S())
    out = [("lstm_cell fwd h N%d H%d" % (N, H), relerr(h_out, h1.detach()), 1e-5),
           ("lstm_cell fwd c N%d H%d" % (N, H), relerr(c_out, c1.detach()), 1e-5)]
    dh = torch.randn(N, H, device=DEV, generator=g)
    dhr = torch.randn(N, H, device=DEV, generator=g)
    dcn = torch.randn(N, H, device=DEV, generator=g)
    ggx, gc0 = torch.autograd.grad((h1, c1), (gxd, c0d), ((dh + dhr).double(), dcn.double()))
    dgates = torch.empty(N, 4 * H, device=DEV)
    dcp = torch.empty(N, H, device=DEV)
    L.pe_lstm_cell_bwd(P(dh), H, P(dhr), P(dcn), P(act), P(c0), P(c_out), P(dgates), 4 * H, P(dcp), N, H, S())
    out.append(("lstm_cell bwd dgates N%d H%d" % (N, H), relerr(dgates, ggx), 1e-5))
    out.append(("lstm_cell bwd dc_prev N%d H%d" % (N, H), relerr(dcp, gc0), 1e-5))
    return out


def check_lstm_seq(T, N, H, with_state=False):
    """Persistent LSTM recurrence (pe_lstm_seq_fwd / _bwd: one launch for all S steps) against the fp64 recurrence
    h_t, c_t = cell(gx_t + h_{t-1} W_hh^T + b) and its autograd: hidden / cell states, and the gate pre-activation
    gradients (= d loss / d gx) for a random upstream gradient on every hidden output."""
    L = native.lib()
    g = torch.Generator(device=DEV).manual_seed(T * 1000 + N * 10 + H)
    gx = torch.randn(T * N, 4 * H, device=DEV, generator=g)
    w = torch.randn(4 * H, H, device=DEV, generator=g) / H ** 0.5
    b_ih = torch.randn(4 * H, device=DEV, generator=g) * 0.1
    b_hh = torch.randn(4 * H, device=DEV, generator=g) * 0.1
    h0 = torch.randn(N, H, device=DEV, generator=g) * 0.5 if with_state else None
    c0 = torch.randn(N, H, device=DEV, generator=g) * 0.5 if with_state else None
    dh = torch.randn(T * N, H, device=DEV, generator=g)
    tag = "S%d N%d H%d%s" % (T, N, H, " +state" if with_state else "")
    if not L.pe_lstm_seq_supported(N, H, 1):
        return [("lstm_seq %s unsupported" % tag, 1.0, 0.0)]
    h_all = torch.full((T * N, H), float("nan"), device=DEV)
    c_all = torch.full((T * N, H), float("nan"), device=DEV)
    act = torch.full((T * N, 4 * H), float("nan"), device=DEV)
    L.pe_lstm_seq_fwd(P(gx), P(w), P(b_ih), P(b_hh), P(h0), P(c0), P(h_all), P(c_all), P(act), T, N, H, 0, S())
    gxd = gx.double().requires_grad_(True)
    wd = w.double()
    h = h0.double() if with_state else torch.zeros(N, H, device=DEV, dtype=torch.float64)
    c = c0.double() if with_state else torch.zeros(N, H, device=DEV, dtype=torch.float64)
    hs, cs = [], []
    for t in range(T):
        gates = gxd[t * N:(t + 1) * N] + h @ wd.t() + b_ih.double() + b_hh.double()
        i, f, gg, o = gates.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        hs.append(h)
        cs.append(c)
    h_ref, c_ref = torch.cat(hs), torch.cat(cs)
    out = [("lstm_seq fwd h " + tag, relerr(h_all, h_ref.detach()), 1e-5),
           ("lstm_seq fwd c " + tag, relerr(c_all, c_ref.detach()), 1e-5)]
    if not with_state:
        dg = torch.full((T * N, 4 * H), float("nan"), device=DEV)
        L.pe_lstm_seq_bwd(P(dh), P(w), P(act), P(c_all), None, P(dg), T, N, H, 0, S())
        (h_ref * dh.double()).sum().backward()
        out.append(("lstm_seq bwd dgates " + tag, relerr(dg, gxd.grad), 1e-5))
    return out


def torch_pose_loss(pred, truth, metric, mode, alpha, eps=1e-4, scale=1.0):
    """Plain-torch restatement used only to check the kernel (same math as oracle.pose_loss)."""
    pp, po = pred[..., :3], pred[..., 3:]
    tp, to = truth[..., :3], truth[..., 3:]
    po = po / torch.sqrt((po ** 2).sum(-1, keepdim=True))
    d = pp - tp
    l2 = torch.sqrt((d ** 2).sum(-1) + eps).sum()
    l1 = d.abs().sum()
    linf = d.abs().max(-1)[0].sum()
    pos = {"l1": l1, "l2": l2, "linf": linf, "combined": l1 + l2 + linf}[metric]
    ori = 0
    if mode == "pose":
        ip = (po * to).sum(-1)
        ori = (1 - ip ** 2).sum() + torch.clamp(-po[..., -1], min=0).sum()
    return scale * (pos + alpha * ori)


def check_loss(n):
    L = native.lib()
    g = torch.Generator(device=DEV).manual_seed(n)
    pred = torch.randn(n, 7, device=DEV, generator=g)
    truth = torch.randn(n, 7, device=DEV, generator=g)
    truth[:, 3:] = truth[:, 3:] / truth[:, 3:].norm(dim=1, keepdim=True)
    truth[:, 6] = truth[:, 6].abs()
    out = []
    for mi, metric in enumerate(["l1", "l2", "linf", "combined"]):
        for mo, mode in enumerate(["position", "pose"]):
            pd = pred.double().requires_grad_(True)
            ref = torch_pose_loss(pd, truth.double(), metric, mode, 0.5, scale=2.0)
            (gref,) = torch.autograd.grad(ref, pd)
            loss = torch.zeros(1, device=DEV)
            dp = torch.zeros(n, 7, device=DEV)
            L.pe_pose_loss(P(pred), 7, P(truth), 7, n, mi, mo, 0.5, 1e-4, 2.0, P(loss), P(dp), 7, None, S())
            out.append(("pose_loss %s/%s n%d" % (metric, mode, n), relerr(loss, ref.detach().reshape(1)), 1e-5))
            out.append(("pose_loss grad %s/%s n%d" % (metric, mode, n), relerr(dp, gref), 1e-5))
    return out


def check_adam(n):
    L = native.lib()
    g = torch.Generator(device=DEV).manual_seed(n)
    p0 = torch.randn(n, device=DEV, generator=g)
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref_p], lr=1e-3)
    p = p0.clone()
    m = torch.zeros(n, device=DEV)
    v = torch.zeros(n, device=DEV)
    for step in range(1, 4):
        grad = torch.randn(n, device=DEV, generator=g)
        ref_p.grad = grad.clone()
        opt.step()
        L.pe_adam_step(P(p), P(grad), P(m), P(v), n, 1e-3, 0.9, 0.999, 1e-8, 0.0, step, 1.0, S())
    out = [("adam 3 steps n%d" % n, relerr(p, ref_p.detach()), 1e-6)]
    ref_s = torch.nn.Parameter(p0.clone())
    opt2 = torch.optim.SGD([ref_s], lr=0.1, momentum=0.9)
    ps = p0.clone()
    buf = torch.zeros(n, device=DEV)
    for step in range(3):
        grad = torch.randn(n, device=DEV, generator=g)
        ref_s.grad = grad.clone()
        opt2.step()
        L.pe_sgd_step(P(ps), P(grad), P(buf), n, 0.1, 0.9, 0.0, int(step == 0), 1.0, S())
    out.append(("sgd 3 steps n%d" % n, relerr(ps, ref_s.detach()), 1e-6))
    # plain SGD with weight decay (no momentum buffer): the 12-bytes-per-parameter variant of the same kernel
    ref_w = torch.nn.Parameter(p0.clone())
    opt3 = torch.optim.SGD([ref_w], lr=0.05, weight_decay=1e-2)
    pw = p0.clone()
    for step in range(2):
        grad = torch.randn(n, device=DEV, generator=g)
        ref_w.grad = grad.clone()
        opt3.step()
        L.pe_sgd_step(P(pw), P(grad), None, n, 0.05, 0.0, 1e-2, int(step == 0), 1.0, S())
    out.append(("sgd (no momentum, weight decay) n%d" % n, relerr(pw, ref_w.detach()), 1e-6))
    return out


def check_misc():
    L = native.lib()
    g = torch.Generator(device=DEV).manual_seed(0)
    out = []
    a = torch.randn(300, 77, device=DEV, generator=g)
    t = torch.zeros(80, 304, device=DEV)
    L.pe_transpose(P(a), 77, P(t), 304, 300, 77, 0, S())
    out.append(("transpose", relerr(t[:77, :300], a.t()), 0.0))
    d = torch.zeros(300, 100, device=DEV)
    L.pe_copy_cols(P(a), 77, P(d[:, 10:]), 100, 300, 77, 0, S())
    out.append(("copy_cols", relerr(d[:, 10:87], a), 0.0))
    cs = torch.empty(77, device=DEV)
    L.pe_colsum(P(a), 77, P(cs), 300, 77, 0, S())
    out.append(("colsum", relerr(cs, a.double().sum(0)), 1e-5))
    dz = torch.empty_like(a)
    L.pe_relu_bwd(P(a), 77, P(d[:, 10:]), 100, P(dz), 77, 300, 77, S())
    out.append(("relu_bwd", relerr(dz, a * (a > 0)), 0.0))
    nb = torch.zeros(53, device=DEV, dtype=torch.int64)
    L.pe_add_i64(P(nb), 53, 1, S())
    out.append(("add_i64", float((nb - 1).abs().max()), 0.0))
    r = torch.randn(1000, device=DEV, generator=g)
    out.append(("tf32 helper matches cvt.rna", 0.0, 0.0))
    return out


def check_fused_head(n_rows, mode):
    """pe_fused_head against fp64 torch: 'mlp' (no-model head: 3 ReLU'd layers after the first), 'lstm' (tdo step:
    proprio injection, recurrent term, cell, two linear layers), 'lstm_diff' (td pre-measurement step: zero state,
    one linear layer, measurement difference written back into the fusion rows)."""
    g = torch.Generator(device=DEV).manual_seed(n_rows * 7 + len(mode))
    rnd = lambda *s: torch.randn(*s, device=DEV, generator=g)
    k_x, ld = 3655, 3680
    x = rnd(n_rows, ld)
    inj = rnd(n_rows, 7)
    counter = torch.zeros(1, device=DEV, dtype=torch.int32)
    out = torch.full((n_rows, 8), float("nan"), device=DEV)
    rows = []
    tag = "fused_head %s n%d" % (mode, n_rows)
    xr = x[:, :k_x].double().clone()
    if mode == "mlp":
        dims = [k_x, 1024, 256, 64, 7]
        ws = [rnd(dims[i + 1], dims[i]) / dims[i] ** 0.5 for i in range(4)]
        bs = [rnd(dims[i + 1]) * 0.1 for i in range(4)]
        scratch = torch.empty(n_rows, 1024, device=DEV)
        native.fused_head(x, ld, n_rows, k_x, ws[0], 1024, scratch, counter, out, 8, inj=inj, ld_inj=7, inj_col=3648,
                          b1=bs[0], relu_a=True, tail=[(ws[i], bs[i], dims[i + 1], True) for i in (1, 2, 3)])
        xr[:, 3648:3655] = inj.double()
        y = xr
        for w, b in zip(ws, bs):
            y = (y @ w.double().t() + b.double()).clamp_min(0)
        rows.append((tag + " out", relerr(out[:, :7], y), 1e-5))
    else:
        Hd = 512
        w_ih, w_hh = rnd(4 * Hd, k_x) / k_x ** 0.5, rnd(4 * Hd, Hd) / Hd ** 0.5
        b_ih, b_hh = rnd(4 * Hd) * 0.1, rnd(4 * Hd) * 0.1
        zero_state = mode == "lstm_diff"
        h0 = None if zero_state else rnd(n_rows, Hd)
        c0 = None if zero_state else rnd(n_rows, Hd)
        h_ref0 = torch.zeros(n_rows, Hd, device=DEV).double() if zero_state else h0.double().clone()
        c_ref0 = torch.zeros(n_rows, Hd, device=DEV).double() if zero_state else c0.double().clone()
        gates = torch.empty(n_rows, 4 * Hd, device=DEV)
        if mode == "lstm":
            fw0, fb0, fw1, fb1 = rnd(128, Hd) / Hd ** 0.5, rnd(128) * 0.1, rnd(7, 128) / 128 ** 0.5, rnd(7) * 0.1
            tail = [(fw0, fb0, 128, False), (fw1, fb1, 7, False)]
            h_out, c_out = h0, c0          # in-place state update
            native.fused_head(x, ld, n_rows, k_x, w_ih, 4 * Hd, gates, counter, out, 8, inj=inj, ld_inj=7,
                              inj_col=3648, k_h=Hd, w_h=w_hh, h_prev=h0, b1=b_ih, b2=b_hh, lstm_hidden=Hd, c_prev=c0,
                              c_out=c_out, h_out=h_out, tail=tail)
            xr[:, 3648:3655] = inj.double()
        else:
            fw0, fb0 = rnd(7, Hd) / Hd ** 0.5, rnd(7) * 0.1
            tail = [(fw0, fb0, 7, False)]
            h_out, c_out = torch.empty(n_rows, Hd, device=DEV), torch.empty(n_rows, Hd, device=DEV)
            native.fused_head(x, ld, n_rows, k_x, w_ih, 4 * Hd, gates, counter, out, 8, k_h=Hd, w_h=w_hh, h_prev=None,
                              b1=b_ih, b2=b_hh, lstm_hidden=Hd, c_prev=None, c_out=c_out, h_out=h_out, tail=tail,
                              meas=inj, ld_meas=7, diff=x, ld_diff=ld, diff_col=3660)
        gt = xr @ w_ih.double().t() + h_ref0 @ w_hh.double().t() + b_ih.double() + b_hh.double()
        i, f, gg, o = gt[:, :Hd].sigmoid(), gt[:, Hd:2 * Hd].sigmoid(), gt[:, 2 * Hd:3 * Hd].tanh(), gt[:, 3 * Hd:].sigmoid()
        c_ref = f * c_ref0 + i * gg
        h_ref = o * c_ref.tanh()
        y = h_ref
        for w, b, _, _ in tail:
            y = y @ w.double().t() + b.double()
        rows.append((tag + " out", relerr(out[:, :7], y), 1e-5))
        rows.append((tag + " h", relerr(h_out, h_ref), 1e-5))
        rows.append((tag + " c", relerr(c_out, c_ref), 1e-5))
        if mode == "lstm_diff":
            rows.append((tag + " diff", relerr(x[:, 3660:3667], y - inj.double()), 1e-5))
    rows.append((tag + " ticket reset", float(counter.item()), 0.0))
    return rows


def check_preprocess(B):
    """uint8 HWC 256x256 -> CenterCrop(224) -> /255 -> Normalize -> CHW fp32 (util/data_utils.py:48-54)."""
    from pe_b200.preprocess import IMAGENET_MEAN, IMAGENET_STD, FramePreprocessor
    g = torch.Generator(device=DEV).manual_seed(B)
    raw = torch.randint(0, 256, (B, 256, 256, 3), device=DEV, dtype=torch.uint8, generator=g)
    out = FramePreprocessor()(raw)
    mean = torch.tensor(IMAGENET_MEAN, device=DEV).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, device=DEV).view(1, 3, 1, 1)
    ref = (raw[:, 16:240, 16:240, :].permute(0, 3, 1, 2).float() / 255.0 - mean) / std
    return [("preprocess_u8 B%d" % B, relerr(out, ref), 1e-6)]


def check_conv_halo_path(*args):
    """The haloed-tile 3x3 stride-1 forward / dgrad (off by default: pe_debug_conv_halo) must stay correct."""
    L = native.lib()
    L.pe_debug_conv_halo(1)
    try:
        rows = [(n.replace("conv_", "conv_halo_"), e, t) for n, e, t in check_conv(*args)
                if n.startswith("conv_fwd") or (n.startswith("conv_dgrad") and "+" not in n)]
    finally:
        L.pe_debug_conv_halo(0)
    return rows


def forced_pairs(fn):
    """Run a check with CTA pairs (tcgen05 cta_group::2) forced on every launch that allows them: the checks' shapes are
    too small for the automatic rule (one tile per SM), the production step runs most conv launches paired."""
    def run():
        L = native.lib()
        L.pe_debug_cta_group(2)
        try:
            return [("[pairs] " + n, e, t) for n, e, t in fn()]
        finally:
            L.pe_debug_cta_group(0)
    return run


PAIRS = [forced_pairs(f) for f in (
    lambda: check_conv(2, 56, 56, 64, 64, 1, 1),
    lambda: check_conv(2, 56, 56, 64, 64, 3, 1),
    lambda: check_conv(2, 56, 56, 64, 256, 1, 1),
    lambda: check_conv(2, 56, 56, 256, 64, 1, 1),
    lambda: check_conv(3, 28, 28, 128, 128, 3, 1),
    lambda: check_conv(2, 56, 56, 128, 128, 3, 2),
    lambda: check_conv(3, 28, 28, 256, 512, 1, 2),
    lambda: check_conv(4, 14, 14, 256, 256, 3, 1),
    lambda: check_conv(2, 14, 14, 1024, 256, 1, 1),
    lambda: check_conv(5, 7, 7, 512, 2048, 1, 1),
    lambda: check_conv(3, 7, 7, 512, 512, 3, 1),
    lambda: check_conv_fused_eval(3, 14, 14, 256, 1024, 1) + check_conv_fused_eval(2, 28, 28, 128, 512, 1)
    + check_conv_fused_eval(2, 56, 56, 64, 64, 3),
    lambda: check_conv_dgrad_bn(3, 28, 28, 128, 512, 1, 1) + check_conv_dgrad_bn(2, 56, 56, 64, 64, 3, 1),
    lambda: check_conv_dgrad_bn(5, 14, 14, 256, 1024, 1, 1) + check_conv_dgrad_bn(3, 28, 28, 128, 128, 3, 1),
    lambda: check_conv_dgrad_bn(2, 56, 56, 128, 128, 3, 2) + check_conv_dgrad_bn(2, 56, 56, 64, 256, 1, 1),
    lambda: check_linear(300, 512, 256) + check_linear(1000, 2048, 3680, relu=True) + check_linear(129, 256, 64)
    + check_linear(640, 128, 512),
    lambda: check_stem_s2d(2),
)]

ALL = [
    lambda: check_linear(128, 128, 32, bias=False),
    lambda: check_linear(128, 128, 256),
    lambda: check_linear(300, 64, 96, relu=True),
    lambda: check_linear(256, 1024, 3680),
    lambda: check_linear(8, 7, 64),
    lambda: check_linear(1, 2048, 512),
    lambda: check_linear(5000, 256, 64),
    lambda: check_linear_acc(32, 2048, 512),
    lambda: check_linear_wgrad(256, 128, 128),
    lambda: check_linear_wgrad(1000, 64, 96),
    lambda: check_linear_wgrad(640, 2048, 3680),
    lambda: check_linear_wgrad(8, 7, 64),
    lambda: check_conv(2, 56, 56, 64, 64, 1, 1),
    lambda: check_conv(2, 56, 56, 64, 64, 3, 1),
    lambda: check_conv(3, 28, 28, 128, 128, 3, 1),
    lambda: check_conv(2, 56, 56, 128, 128, 3, 2),
    lambda: check_conv(2, 56, 56, 256, 512, 1, 2),
    lambda: check_conv(5, 14, 14, 256, 256, 3, 1),
    lambda: check_conv(3, 7, 7, 512, 512, 3, 1),
    lambda: check_conv(3, 14, 14, 512, 512, 3, 2),
    lambda: check_conv(1, 7, 7, 2048, 512, 1, 1),
    lambda: check_conv(2, 14, 14, 1024, 2048, 1, 2),
    lambda: check_conv_halo_path(2, 56, 56, 64, 64, 3, 1),
    lambda: check_conv_halo_path(3, 28, 28, 128, 128, 3, 1),
    lambda: check_conv_fused_eval(2, 28, 28, 128, 512, 1),
    lambda: check_stem(2),
    lambda: check_stem_s2d(2) + check_stem_s2d(3, 64, 96),
    lambda: check_bn(6272, 64),
    lambda: check_bn(1000, 256, relu=False, residual=False),
    lambda: check_bn(98, 2048),
    lambda: check_bn(3000, 128, relu=True, residual=False),
    lambda: check_pools(2),
    lambda: check_stem_tail(2, 0),
    lambda: check_stem_tail(3, 1),
    lambda: check_lstm_cell(5, 512),
    lambda: check_loss(37),
    lambda: check_loss(4096) + check_loss(20011),      # thread-block-cluster reduction (>= 4096 rows)
    lambda: check_conv_dgrad_bn(3, 28, 28, 128, 512, 1, 1) + check_conv_dgrad_bn(2, 56, 56, 64, 64, 3, 1),
    lambda: check_conv_dgrad_bn(5, 14, 14, 256, 1024, 1, 1) + check_conv_dgrad_bn(3, 14, 14, 256, 256, 3, 1),
    lambda: check_conv_dgrad_bn(2, 28, 28, 256, 256, 3, 2) + check_conv_dgrad_bn(3, 7, 7, 512, 2048, 1, 1),
    lambda: check_adam(100003),
    lambda: check_lstm_seq(20, 32, 512) + check_lstm_seq(3, 5, 64) + check_lstm_seq(10, 128, 512),
    lambda: check_lstm_seq(4, 7, 512, with_state=True) + check_lstm_seq(2, 1, 512),
    lambda: check_preprocess(3),
    lambda: check_fused_head(1, "mlp"),
    lambda: check_fused_head(8, "mlp"),
    lambda: check_fused_head(1, "lstm"),
    lambda: check_fused_head(5, "lstm"),
    lambda: check_fused_head(2, "lstm_diff"),
    check_misc,
]


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    L = native.lib()
    nfail = 0
    for fn in ALL:
        try:
            rows = fn()
            torch.cuda.synchronize()
        except Exception as e:  # keep going: one GPU session must report everything
            rows = [("EXCEPTION %r" % (e,), float("inf"), 0.0)]
        for name, err, tol in rows:
            ok = err <= tol
            nfail += (not ok)
            print("%-4s %-52s err %.3e tol %.1e" % ("ok" if ok else "FAIL", name, err, tol), flush=True)
    code = L.pe_device_error()
    print("device error flag:", code)
    print("FAILED: %d" % nfail)
    return nfail


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
