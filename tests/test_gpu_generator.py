"""Parity of the CUDA generators (through the C ABI) with the torch-CPU oracle on identical weights / inputs / eps.
Tolerances (BASELINE.json north_star): fp32 mode <= 1e-4, bf16 tensor-core mode <= 1e-2, max abs error on the
normalised output (outputs are O(1))."""
import os

import numpy as np
import pytest

from moonsuperresolution_b200 import weights as W
from oracle import generator as OG

pytestmark = pytest.mark.gpu

TOL_FP32 = 1e-4
TOL_BF16 = 1e-2


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def msr(torch):
    import moonsuperresolution_b200 as m
    return m


def bf16_round(a, torch):
    return torch.from_numpy(a).to(torch.bfloat16).to(torch.float32).numpy()


@pytest.mark.parametrize("n,r,cin,cout", [(1, 8, 64, 128), (2, 16, 128, 256), (3, 4, 128, 128), (1, 32, 256, 384),
                                           (5, 2, 64, 128), (1, 128, 64, 128), (2, 1, 128, 256),
                                           (2, 128, 64, 256), (8, 64, 64, 128), (3, 128, 64, 512),    # CTA-pair schedule
                                           (2, 256, 256, 128), (3, 128, 128, 128)])   # two output rows per work unit
def test_conv3x3_tensor_core_operator(torch, n, r, cin, cout):
    """tcgen05 implicit GEMM vs torch fp32 convolution on bf16-rounded operands."""
    from moonsuperresolution_b200 import _lib
    rng = np.random.default_rng(n * 1000 + r)
    x = bf16_round(rng.standard_normal((n, r, r, cin)).astype(np.float32), torch)
    w = bf16_round((rng.standard_normal((3, 3, cin, cout)) / np.sqrt(9 * cin)).astype(np.float32), torch)
    b = rng.standard_normal(cout).astype(np.float32)
    want = OG.conv2d_same(torch.from_numpy(x).permute(0, 3, 1, 2), w, b).permute(0, 2, 3, 1).numpy()
    d_x = torch.from_numpy(x).cuda().to(torch.bfloat16).contiguous()
    wt = np.ascontiguousarray(w.reshape(9 * cin, cout).T)                     # [cout][9*cin], k = tap*cin + ci
    d_w = torch.from_numpy(wt).cuda().to(torch.bfloat16).contiguous()
    d_b = torch.from_numpy(b).cuda()
    d_y = torch.zeros((n, r, r, cout), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().msr_op_conv3x3_bf16(d_x.data_ptr(), d_w.data_ptr(), d_b.data_ptr(), d_y.data_ptr(), n, r, cin,
                                              cout, _lib.stream_ptr()), "msr_op_conv3x3_bf16")
    torch.cuda.synchronize()
    got = d_y.cpu().numpy()
    assert np.abs(got - want).max() < 2e-3, np.abs(got - want).max()


@pytest.mark.parametrize("n,r_out,cin,cout,taps,stride,pad,act", [
    (2, 16, 64, 128, 9, 2, 0, 0),      # encoder block: 3x3 stride 2, SAME pad (0, 1)
    (1, 64, 128, 256, 9, 2, 0, 0),
    (3, 8, 64, 64, 1, 1, 0, 2),        # 1x1 (im2col GEMM), leaky-relu, bf16 out, 64 columns
    (2, 32, 64, 128, 1, 1, 0, 1),      # 1x1, relu, bf16 out
    (1, 16, 128, 32, 9, 1, 1, 0),      # 32 columns
    (2, 16, 64, 128, 9, 1, 1, 0),      # fp32 out + fused statistics
    (2, 256, 64, 128, 9, 1, 1, 0),     # the same on the two-rows-per-unit strip schedule (statistics rows per tile)
    (5, 128, 64, 128, 9, 1, 1, 0),
])
def test_conv_tc_general_operator(torch, n, r_out, cin, cout, taps, stride, pad, act):
    from moonsuperresolution_b200 import _lib
    import torch.nn.functional as F
    rng = np.random.default_rng(r_out * 7 + cout)
    rin = r_out * stride
    ks = 3 if taps == 9 else 1
    x = bf16_round(rng.standard_normal((n, rin, rin, cin)).astype(np.float32), torch)
    w = bf16_round((rng.standard_normal((ks, ks, cin, cout)) / np.sqrt(taps * cin)).astype(np.float32), torch)
    b = rng.standard_normal(cout).astype(np.float32)
    xt = torch.from_numpy(x).permute(0, 3, 1, 2)
    if taps == 9:
        after = (r_out - 1) * stride + 3 - pad - rin       # zero rows / cols needed after the last input element
        xt = F.pad(xt, (pad, max(after, 0), pad, max(after, 0)))
    wt_t = torch.from_numpy(w).permute(3, 2, 0, 1)
    want = F.conv2d(xt, wt_t, torch.from_numpy(b), stride=stride).permute(0, 2, 3, 1).numpy()
    assert want.shape == (n, r_out, r_out, cout)
    d_x = torch.from_numpy(x).cuda().to(torch.bfloat16).contiguous()
    d_w = torch.from_numpy(np.ascontiguousarray(w.reshape(taps * cin, cout).T)).cuda().to(torch.bfloat16).contiguous()
    d_b = torch.from_numpy(b).cuda()
    bf16_out = taps == 1
    use_stats = (not bf16_out) and stride == 1 and r_out * r_out >= 128 and cout == 128
    d_y = torch.zeros((n, r_out, r_out, cout), dtype=torch.bfloat16 if bf16_out else torch.float32, device="cuda")
    pairs = torch.zeros((n * r_out * r_out // 128 * 4, cout, 2), dtype=torch.float32, device="cuda") if use_stats else None
    _lib.check(_lib.lib().msr_op_conv_tc(d_x.data_ptr(), d_w.data_ptr(), d_b.data_ptr(),
                                         None if bf16_out else d_y.data_ptr(), d_y.data_ptr() if bf16_out else None,
                                         n, r_out, cin, cout, taps, stride, pad, act, 0.2, _lib.ptr(pairs),
                                         _lib.stream_ptr()), "msr_op_conv_tc")
    torch.cuda.synchronize()
    got = d_y.float().cpu().numpy()
    if act == 1:
        want = np.maximum(want, 0)
    elif act == 2:
        want = np.where(want > 0, want, 0.2 * want)
    tol = 2e-2 if bf16_out else 2e-3
    assert np.abs(got - want).max() < tol, np.abs(got - want).max()
    if use_stats:
        p = pairs.cpu().numpy().astype(np.float64)
        flat = got.reshape(-1, cout).astype(np.float64)
        np.testing.assert_allclose(p[:, :, 0].sum(0), flat.sum(0), rtol=1e-5, atol=1e-3)
        np.testing.assert_allclose(p[:, :, 1].sum(0), (flat ** 2).sum(0), rtol=1e-5, atol=1e-3)
        # each (tile, warp) row covers 32 consecutive pixels
        np.testing.assert_allclose(p[5, :, 0], flat[5 * 32:6 * 32].sum(0), rtol=1e-5, atol=1e-4)
        for row in (0, 7, p.shape[0] // 2 + 3, p.shape[0] - 1):
            np.testing.assert_allclose(p[row, :, 0], flat[row * 32:(row + 1) * 32].sum(0), rtol=1e-5, atol=1e-4)
            np.testing.assert_allclose(p[row, :, 1], (flat[row * 32:(row + 1) * 32] ** 2).sum(0), rtol=1e-5, atol=1e-4)


def _phase_filters(kernel, transposed):
    """(4, 4, cin) kernel of the one-channel last layer -> (4, 3, 3, cin) sub-pixel phase filters on the LOW-resolution
    tensor.  transposed=False: UpSampling2D(2) -> Conv2D(1, 4, 'same') (networks.py:54-56; SAME pads 1 before, 2 after);
    transposed=True: Conv2DTranspose(1, 4, strides=2, 'same') (pix2pix.py:91-95)."""
    cin = kernel.shape[-1]
    w4 = np.zeros((4, 3, 3, cin), np.float32)
    for py in range(2):
        for px in range(2):
            if not transposed:
                for ky in range(4):
                    for kx in range(4):
                        ty, tx = (py - 1 + ky) // 2 + 1, (px - 1 + kx) // 2 + 1
                        w4[py * 2 + px, ty, tx] += kernel[ky, kx]
            else:
                for ty in range(3):
                    for tx in range(3):
                        ky, kx = py + 1 - 2 * (ty - 1), px + 1 - 2 * (tx - 1)
                        if 0 <= ky <= 3 and 0 <= kx <= 3:
                            w4[py * 2 + px, ty, tx] = kernel[ky, kx]
    return w4


@pytest.mark.parametrize("n,r,cin,transposed", [
    (2, 128, 64, True),       # few patches: 2 output rows per work unit
    (3, 256, 128, False),     # ragged batch, r = 256: one image row per 256-pixel group
    (5, 128, 128, True),      # pix2pix's last layer (two image rows per group), tanh
    (40, 256, 128, False),    # the bench call shape's schedule: 16 output rows per work unit, several units per CTA
    (80, 128, 128, False),
])
def test_phase_layer_tensor_core_operator(torch, n, r, cin, transposed):
    """The last layer in its "contract once per pixel, then 3x3 stencil of scalars" form (csrc/phase_tc.cu) against
    torch's float64 convolution at FULL resolution on bf16-rounded operands (the phase filters are summed in float32 and
    rounded to bf16 before both sides use them, as generator.cu does)."""
    from moonsuperresolution_b200 import _lib
    import torch.nn.functional as F
    rng = np.random.default_rng(n * 1000 + r + cin)
    x = bf16_round(rng.standard_normal((n, r, r, cin)).astype(np.float32), torch)
    x[0, 0, :, :] = 3.0                                  # borders carry weight: SAME padding must be zeros there
    x[-1, :, -1, :] = -2.0
    kernel = (rng.standard_normal((4, 4, cin)) / np.sqrt(16 * cin)).astype(np.float32)
    w4 = bf16_round(_phase_filters(kernel, transposed), torch)              # (4, 3, 3, cin)
    bias = np.array([0.3], np.float32)
    # reference: the 4 phase filters as a 3x3 convolution to 4 channels + pixel shuffle, float64 on the device
    xt = torch.from_numpy(x).cuda().double().permute(0, 3, 1, 2)
    wt = torch.from_numpy(w4).cuda().double().permute(0, 3, 1, 2)           # (4, cin, 3, 3)
    ph = F.conv2d(xt, wt, padding=1) + float(bias[0])                       # (n, 4, r, r)
    want = F.pixel_shuffle(ph, 2)[:, 0]                                     # channel py*2+px -> (2h+py, 2w+px)
    if transposed:
        want = torch.tanh(want)
    # and the phase decomposition itself against the full-resolution layer it stands for (unrounded kernel, small n)
    if n <= 3:
        kt = torch.from_numpy(kernel).cuda().double()
        xs = xt[:1]
        if transposed:
            full = F.conv_transpose2d(xs, kt.permute(2, 0, 1)[:, None], stride=2, padding=1)[:, 0]
        else:
            up = F.interpolate(xs, scale_factor=2, mode="nearest")
            full = F.conv2d(F.pad(up, (1, 2, 1, 2)), kt.permute(2, 0, 1)[None])[:, 0]
        w4f = torch.from_numpy(_phase_filters(kernel, transposed)).cuda().double().permute(0, 3, 1, 2)
        dec = F.pixel_shuffle(F.conv2d(xs, w4f, padding=1), 2)[:, 0]
        assert (full - dec).abs().max().item() < 1e-5       # the phase filters are summed in float32
    d_x = torch.from_numpy(x).cuda().to(torch.bfloat16).contiguous()
    h_w4 = torch.from_numpy(w4.reshape(4, 9 * cin)).to(torch.bfloat16).contiguous().view(torch.int16).numpy()
    d_b = torch.from_numpy(bias).cuda()
    d_y = torch.full((n, 2 * r, 2 * r), float("nan"), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().msr_op_phase_tc(d_x.data_ptr(), h_w4.ctypes.data, d_b.data_ptr(), d_y.data_ptr(), n, r, cin,
                                          3 if transposed else 0, _lib.stream_ptr()), "msr_op_phase_tc")
    torch.cuda.synchronize()
    err = (d_y.double() - want).abs().max().item()
    assert err < 1e-4, err


@pytest.mark.parametrize("n,i,r", [(3, 64, 8), (2, 128, 128), (5, 256, 64), (16, 512, 256), (1, 64, 4), (7, 64, 1)])
def test_mask_convolution_tensor_core_operator(torch, n, i, r):
    """SPADE's mask convolution with the operand tile built inside the kernel (csrc/mask_tc.cu) against torch float64:
    nearest resize with half-pixel centres (source pixel h * I/r + I/(2r)), 3x3 SAME conv 2 -> 128, bias, ReLU.  The
    split-bf16 operands give ~float32 products, so the only visible rounding is the bf16 output (2^-9 relative)."""
    from moonsuperresolution_b200 import _lib
    import torch.nn.functional as F
    rng = np.random.default_rng(n * 100 + r)
    src = rng.uniform(-0.5, 0.5, (n, i, i, 2)).astype(np.float32)
    src[0, :, 0] = 0.5                                   # values on the border: SAME padding must contribute zeros
    w = (rng.standard_normal((3, 3, 2, 128)) / np.sqrt(18)).astype(np.float32)
    b = (0.1 * rng.standard_normal(128)).astype(np.float32)
    f = i // r
    idx = np.arange(r) * f + f // 2
    mask = torch.from_numpy(src[:, idx][:, :, idx]).cuda().double().permute(0, 3, 1, 2)
    want = torch.relu(F.conv2d(mask, torch.from_numpy(w).cuda().double().permute(3, 2, 0, 1), torch.from_numpy(b).cuda().double(),
                               padding=1)).permute(0, 2, 3, 1)
    d_src = torch.from_numpy(src).cuda()
    d_y = torch.full((n, r, r, 128), float("nan"), dtype=torch.bfloat16, device="cuda")
    guard = torch.full((4096,), 7.0, dtype=torch.bfloat16, device="cuda")     # nothing may be written past the output
    _lib.check(_lib.lib().msr_op_mask_tc(d_src.data_ptr(), i, w.ctypes.data, b.ctypes.data, d_y.data_ptr(), n, r,
                                         _lib.stream_ptr()), "msr_op_mask_tc")
    torch.cuda.synchronize()
    got = d_y.double()
    assert torch.isfinite(got).all()
    err = ((got - want).abs() / (1.0 + want.abs())).max().item()
    assert err < 2.0 ** -8, err
    # against the bf16 rounding of the exact result: at most one bf16 step away (products are ~float32 exact)
    exact_bf16 = want.float().to(torch.bfloat16).double()
    assert ((got - exact_bf16).abs() <= 2.0 ** -7 * (want.abs() + 1e-3)).all()
    assert (guard == 7.0).all()


@pytest.mark.parametrize("n,i", [(3, 64), (2, 256), (16, 512), (1, 8)])
def test_encoder_block1_tensor_core_operator(torch, n, i):
    """Encoder block 1 (Conv2D(64, 3, strides=2, 'same', no bias) -> LeakyReLU(0.2)) with the operand tile built inside the
    kernel (csrc/mask_tc.cu, MODE 1) against torch float64; SAME on an even input with stride 2 pads (0, 1).  The output is
    the split-bf16 pair hi | lo: hi must be the bf16 rounding of the value (one step at most), hi + lo the value to ~2^-16."""
    from moonsuperresolution_b200 import _lib
    import torch.nn.functional as F
    rng = np.random.default_rng(n * 10 + i)
    src = rng.uniform(-0.5, 0.5, (n, i, i, 2)).astype(np.float32)
    src[0, -1, :] = 0.5                                  # the padded side (bottom / right) carries weight
    src[0, :, -1] = -0.5
    w = (rng.standard_normal((3, 3, 2, 64)) / np.sqrt(18)).astype(np.float32)
    xt = F.pad(torch.from_numpy(src).cuda().double().permute(0, 3, 1, 2), (0, 1, 0, 1))
    want = F.leaky_relu(F.conv2d(xt, torch.from_numpy(w).cuda().double().permute(3, 2, 0, 1), stride=2), 0.2).permute(0, 2, 3, 1)
    r = i // 2
    assert want.shape == (n, r, r, 64)
    d_src = torch.from_numpy(src).cuda()
    d_y = torch.full((n, r, r, 128), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.check(_lib.lib().msr_op_enc1_tc(d_src.data_ptr(), i, w.ctypes.data, d_y.data_ptr(), n, 0.2, _lib.stream_ptr()),
               "msr_op_enc1_tc")
    torch.cuda.synchronize()
    hi, lo = d_y[..., :64].double(), d_y[..., 64:].double()
    assert torch.isfinite(hi).all() and torch.isfinite(lo).all()
    assert ((hi - want).abs() <= 2.0 ** -7 * (want.abs() + 1e-3)).all()
    err = ((hi + lo - want).abs() / (1.0 + want.abs())).max().item()
    assert err < 3e-5, err


@pytest.mark.parametrize("n,r,C,x_shift,spg", [(4, 16, 128, 1, 2), (2, 64, 256, 0, 2), (2, 128, 128, 1, 1), (3, 4, 64, 0, 3)])
def test_fused_spade_operator(torch, n, r, C, x_shift, spg):
    """gamma|beta conv + normalise + modulate + LeakyReLU(0.2) (spade.py:19-24, blocks.py:30) with the nearest x2
    upsampling of x fused (x stored at r >> x_shift)."""
    from moonsuperresolution_b200 import _lib
    rng = np.random.default_rng(r + C)
    a = bf16_round(np.maximum(rng.standard_normal((n, r, r, 128)), 0).astype(np.float32), torch)
    wg = bf16_round((rng.standard_normal((3, 3, 128, C)) / np.sqrt(1152)).astype(np.float32), torch)
    wb = bf16_round((rng.standard_normal((3, 3, 128, C)) / np.sqrt(1152)).astype(np.float32), torch)
    bg, bb = rng.standard_normal(C).astype(np.float32), rng.standard_normal(C).astype(np.float32)
    rs = r >> x_shift
    x = rng.standard_normal((n, rs, rs, C)).astype(np.float32) * 2 + 1
    groups = n // spg
    mean = rng.standard_normal((groups, C)).astype(np.float32)
    rstd = rng.uniform(0.5, 2.0, (groups, C)).astype(np.float32)
    at = torch.from_numpy(a).permute(0, 3, 1, 2)
    gamma = OG.conv2d_same(at, wg, bg).permute(0, 2, 3, 1).numpy()
    beta = OG.conv2d_same(at, wb, bb).permute(0, 2, 3, 1).numpy()
    xu = x.repeat(1 << x_shift, axis=1).repeat(1 << x_shift, axis=2)
    gi = np.arange(n) // spg
    t = gamma * ((xu - mean[gi][:, None, None, :]) * rstd[gi][:, None, None, :]) + beta
    want = np.where(t > 0, t, 0.2 * t)
    # interleave rows per 64 channels: [64 gamma | 64 beta] ...
    K = 1152
    wt = np.zeros((2 * C, K), np.float32)
    bias = np.zeros(2 * C, np.float32)
    for c in range(C):
        j, q = divmod(c, 64)
        wt[128 * j + q] = wg.reshape(K, C)[:, c]
        wt[128 * j + 64 + q] = wb.reshape(K, C)[:, c]
        bias[128 * j + q], bias[128 * j + 64 + q] = bg[c], bb[c]
    d_a = torch.from_numpy(a).cuda().to(torch.bfloat16).contiguous()
    d_w = torch.from_numpy(wt).cuda().to(torch.bfloat16).contiguous()
    d_b, d_x, d_m, d_r = (torch.from_numpy(v).cuda() for v in (bias, x, mean, rstd))
    d_o = torch.zeros((n, r, r, C), dtype=torch.bfloat16, device="cuda")
    _lib.check(_lib.lib().msr_op_spade_tc(d_a.data_ptr(), d_w.data_ptr(), d_b.data_ptr(), d_x.data_ptr(), x_shift,
                                          d_m.data_ptr(), d_r.data_ptr(), spg, d_o.data_ptr(), n, r, C,
                                          _lib.stream_ptr()), "msr_op_spade_tc")
    torch.cuda.synchronize()
    got = d_o.float().cpu().numpy()
    assert np.abs(got - want).max() <= 0.02 * max(1.0, np.abs(want).max())      # bf16 output rounding


@pytest.mark.parametrize("n,r,cin,cout", [(2, 8, 5, 7), (1, 16, 64, 33), (3, 4, 128, 128)])
def test_conv3x3_fp32_operator(torch, n, r, cin, cout):
    from moonsuperresolution_b200 import _lib
    rng = np.random.default_rng(7)
    x = rng.standard_normal((n, r, r, cin)).astype(np.float32)
    w = (rng.standard_normal((3, 3, cin, cout)) / np.sqrt(9 * cin)).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    want = OG.conv2d_same(torch.from_numpy(x).permute(0, 3, 1, 2), w, b).permute(0, 2, 3, 1).numpy()
    d_x, d_w, d_b = (torch.from_numpy(a).cuda() for a in (x, w, b))
    d_y = torch.zeros((n, r, r, cout), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().msr_op_conv3x3_f32(d_x.data_ptr(), d_w.data_ptr(), d_b.data_ptr(), d_y.data_ptr(), n, r, cin,
                                             cout, _lib.stream_ptr()), "msr_op_conv3x3_f32")
    torch.cuda.synchronize()
    assert np.abs(d_y.cpu().numpy() - want).max() < 1e-5


def inputs(i, b, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-0.5, 0.5, (b, i, i, 2)).astype(np.float32)
    x[-1] = 0.0                                    # a zero padding slot takes part in the batch statistics (:468-474)
    eps = rng.standard_normal((b, 256)).astype(np.float32)
    return x, eps


@pytest.mark.parametrize("arch,i,b", [("cnn", 64, 3), ("spade", 64, 2), ("spade", 128, 2)])
def test_spade_generator_fp32_mode(msr, arch, i, b):
    w = W.random_init(arch, i, seed=11, perturb_affine=True)
    x, eps = inputs(i, b)
    want, want_latent = OG.gaugan_call(x, w, eps, arch, return_latent=True)
    cls = msr.GauGAN if arch == "spade" else msr.CNNSpade
    model = cls(i, b, precision="fp32", weights=w)
    got = model(x, training=False, eps=eps)
    assert got.shape == want.shape == (b, i, i, 1)
    lat = model.read_activation("latent").reshape(b, 256)
    assert np.abs(lat - want_latent).max() < 1e-4 * max(1.0, np.abs(want_latent).max())
    err = np.abs(got - want).max()
    assert err <= TOL_FP32 * max(1.0, np.abs(want).max()), err


@pytest.mark.parametrize("arch,i,b", [("cnn", 64, 3), ("spade", 128, 2), ("cnn", 256, 1), ("spade", 512, 2)])
def test_spade_generator_bf16_tensor_core_mode(msr, arch, i, b):
    """(spade, 512, 2) is BASELINE.json's model size: its r = 128 / 256 layers run the CTA-pair + strip-mode schedule."""
    w = W.random_init(arch, i, seed=12, perturb_affine=True)
    x, eps = inputs(i, b, seed=1)
    want, want_latent = OG.gaugan_call(x, w, eps, arch, return_latent=True)
    cls = msr.GauGAN if arch == "spade" else msr.CNNSpade
    model = cls(i, b, precision="bf16", weights=w)
    got = model(x, training=False, eps=eps)
    # the encoder runs with split-bf16 operands (~fp32 products): the latent must be far inside the bf16 tolerance
    lat = model.read_activation("latent").reshape(b, 256)
    assert np.abs(lat - want_latent).max() < 1e-3 * max(1.0, np.abs(want_latent).max())
    err = np.abs(got - want).max()
    assert err <= TOL_BF16 * max(1.0, np.abs(want).max()), err


def test_gaugan512_at_the_bench_call_shape_bf16(msr, torch):
    """BASELINE.json configs[2] at bench.py's own call shape: GauGAN-512, bf16 tensor-core mode, batch 16, EIGHT groups in
    one generator call (128 patches, the n = 128 plans with CTA pairs / strips on every layer from rb2 on).  Group 5
    carries 7 zero padding slots, the way processTile pads the last batch of a tile (process_full_tiles.py:468-474): they
    take part in that group's batch statistics.  Groups 0, 5 and 7 are held to the oracle at north_star's bf16 bar
    (<= 1e-2 max abs on the O(1) output); every group must equal the same model called on that group alone (batch
    statistics never cross groups)."""
    i, b, groups = 512, 16, 8
    w = W.random_init("spade", i, seed=0, perturb_affine=True)
    rng = np.random.default_rng(42)
    x = rng.uniform(-0.5, 0.5, (groups * b, i, i, 2)).astype(np.float32)
    x[5 * b + 9:6 * b] = 0.0
    eps = rng.standard_normal((groups * b, 256)).astype(np.float32)
    model = msr.GauGAN(i, b, precision="bf16", weights=w, max_groups=groups)
    src = torch.from_numpy(x).cuda()
    d_eps = torch.from_numpy(eps).cuda()
    out = torch.empty((groups * b, i, i), dtype=torch.float32, device="cuda")
    model.forward_device(src, out, d_eps, groups)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert np.isfinite(got).all()
    worst = 0.0
    for g in (0, 5, 7):
        sl = slice(g * b, (g + 1) * b)
        want = OG.gaugan_call(x[sl], w, eps[sl], "spade")[..., 0]
        err = np.abs(got[sl] - want).max() / max(1.0, np.abs(want).max())
        worst = max(worst, err)
        assert err <= TOL_BF16, (g, err)
    one = torch.empty((b, i, i), dtype=torch.float32, device="cuda")
    for g in range(groups):
        sl = slice(g * b, (g + 1) * b)
        model.forward_device(src[sl], one, d_eps[sl], 1)
        torch.cuda.synchronize()
        # same kernels, same per-group statistics; only the dense layers' split-K shape depends on the row count, which
        # moves the latent by float32 rounding and, through six blocks of bf16 roundings, the output by a few 1e-3
        # (measured 6.3e-3): both runs sit inside the bf16 bar around the oracle, so they may differ by at most that
        assert np.abs(one.cpu().numpy() - got[sl]).max() <= TOL_BF16, g
    print("GauGAN-512 bf16 B=16 x 8 groups: worst normalised max-abs error vs oracle", worst)


def test_gaugan512_fp32_mode(msr):
    """fp32 parity mode at BASELINE.json's model size (I = 512), batch 2 with one zero padding slot: <= 1e-4."""
    i, b = 512, 2
    w = W.random_init("spade", i, seed=5, perturb_affine=True)
    x, eps = inputs(i, b, seed=9)
    want, want_latent = OG.gaugan_call(x, w, eps, "spade", return_latent=True)
    model = msr.GauGAN(i, b, precision="fp32", weights=w)
    got = model(x, training=False, eps=eps)
    lat = model.read_activation("latent").reshape(b, 256)
    assert np.abs(lat - want_latent).max() < 1e-4 * max(1.0, np.abs(want_latent).max())
    err = np.abs(got - want).max() / max(1.0, np.abs(want).max())
    assert err <= TOL_FP32, err


@pytest.mark.parametrize("b,precision,atol", [(2, "fp32", 1e-6), (9, "fp32", 1e-5), (9, "bf16", 5e-3)])
def test_groups_have_independent_batch_statistics(msr, torch, b, precision, atol):
    """Two batches pushed through one forward call (max_groups = 2) equal two separate calls (b = 9: 18 rows go through
    the many-row dense kernel, whose different summation order flips a few bf16 roundings downstream)."""
    i = 64
    w = W.random_init("cnn", i, seed=3)
    x, _ = inputs(i, 2 * b, seed=2)
    one = msr.CNNSpade(i, b, precision=precision, weights=w)
    two = msr.CNNSpade(i, b, precision=precision, weights=w, max_groups=2)
    sep = np.concatenate([one(x[:b]), one(x[b:])])
    src = torch.from_numpy(x).cuda()
    out = torch.empty((2 * b, i, i), dtype=torch.float32, device="cuda")
    two.forward_device(src, out, None, 2)
    np.testing.assert_allclose(out.cpu().numpy()[..., None], sep, atol=atol)


@pytest.mark.parametrize("precision,tol", [("fp32", TOL_FP32), ("bf16", TOL_BF16)])
def test_pix2pix_generator(msr, precision, tol):
    """pix2pix U-Net (pix2pix.py:64-108); bf16 = every (transposed) convolution on the tcgen05 kernel (4x4 stride-2
    implicit GEMMs, transposed convs as sub-pixel phases, BatchNorm folded into the epilogue)."""
    w = W.random_init("pix2pix", 256, seed=4, perturb_affine=True)
    x, _ = inputs(256, 2, seed=3)
    want = OG.pix2pix_call(x, w)
    got = msr.Pix2Pix(batch_size=2, weights=w, precision=precision)(x, training=False)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= tol, np.abs(got - want).max()


def test_engine_with_device_model_matches_oracle_pipeline(msr):
    """Full path with the CUDA generator plugged in vs the oracle pipeline with the oracle generator."""
    from oracle import tiling as OT
    i, s, b, t = 64, 32, 4, 128
    rng = np.random.default_rng(0)
    h, w_ = 150, 170
    dem = np.cumsum(np.cumsum(rng.standard_normal((h, w_)), 0), 1).astype(np.float32)
    img = rng.uniform(1, 255, (h, w_)).astype(np.float32)
    dem[60:63, 80:90] = -32768.0
    weights = W.random_init("cnn", i, seed=1, perturb_affine=True)
    cfg = msr.DSRConfig(image_size=i, stride=s, batch_size=b, tile_size=t)
    ref = OT.process_map(dem, img, i, s, b, t, cfg.no_value, OG.OracleModel("cnn", weights))
    scale = float(dem[dem > -32768].max() - dem[dem > -32768].min())
    for precision, tol in (("fp32", TOL_FP32), ("bf16", TOL_BF16)):
        eng = msr.DEMSuperResolution(cfg, model=msr.CNNSpade(i, b, precision=precision, weights=weights, max_groups=3))
        mean, std, good = eng.run(dem, img)
        np.testing.assert_array_equal(good, ref[2])
        g = good.astype(bool)
        # north_star: max abs error <= 1e-2 (bf16) / 1e-4 (fp32) in normalised units, i.e. relative to the DEM's range
        assert np.abs(mean[g] - ref[0][g]).max() / scale <= tol, (precision, np.abs(mean[g] - ref[0][g]).max() / scale)
        assert np.abs(std[g] - ref[1][g]).max() / scale <= tol, (precision, np.abs(std[g] - ref[1][g]).max() / scale)
        assert (mean[~g] == cfg.no_value).all()


def test_pix2pix_tiled_pipeline_matches_oracle(msr):
    """BASELINE.json configs[1] in miniature: pix2pix-256, stride 32, batch 16, tile 1024 over a 448 x 480 raster (49
    valid patches, last batch padded) against the oracle pipeline with the oracle generator."""
    from oracle import tiling as OT
    i, s, b, t = 256, 32, 16, 1024
    rng = np.random.default_rng(4)
    h, w_ = 448, 480
    dem = np.cumsum(np.cumsum(rng.standard_normal((h, w_)), 0), 1).astype(np.float32)
    img = rng.uniform(1, 255, (h, w_)).astype(np.float32)
    weights = W.random_init("pix2pix", i, seed=5, perturb_affine=True)
    cfg = msr.DSRConfig(image_size=i, stride=s, batch_size=b, tile_size=t)
    ref = OT.process_map(dem, img, i, s, b, t, cfg.no_value, OG.OracleModel("pix2pix", weights))
    eng = msr.DEMSuperResolution(cfg, model=msr.Pix2Pix(batch_size=b, weights=weights))
    mean, std, good = eng.run(dem, img)
    np.testing.assert_array_equal(good, ref[2])
    g = good.astype(bool)
    assert g.any()
    scale = float(dem.max() - dem.min())
    assert np.abs(mean[g] - ref[0][g]).max() / scale <= TOL_FP32, np.abs(mean[g] - ref[0][g]).max() / scale
    assert np.abs(std[g] - ref[1][g]).max() / scale <= TOL_FP32, np.abs(std[g] - ref[1][g]).max() / scale


def test_repeated_forward_calls_honour_new_output_buffers(msr, torch):
    """The tensor-core plans are cached after the first call; a later call with different output / eps buffers must
    write there (regression: the cached plan of the final layer kept the first call's output pointer)."""
    i, b = 64, 4
    w = W.random_init("spade", i, seed=8)
    model = msr.GauGAN(i, b, precision="bf16", weights=w)
    x, eps = inputs(i, b, seed=5)
    src = torch.from_numpy(x).cuda()
    e = torch.from_numpy(eps).cuda()
    out1 = torch.full((b, i, i), 7.0, device="cuda")
    out2 = torch.full((b, i, i), 7.0, device="cuda")
    model.forward_device(src, out1, e, 1)
    model.forward_device(src, out2, e, 1)
    torch.cuda.synchronize()
    assert not (out2 == 7.0).any()
    assert torch.equal(out1, out2)


def test_load_gan_model_reads_saved_model_directories(msr, tmp_path):
    """load_GAN_model(path, I, B) (process_full_tiles.py:13-31) on the directory layout GauGAN.save writes
    (spade/models/model.py:569-605): <path>/generator, <path>/discriminator, <path>/encoder as Keras SavedModel
    directories, read without TensorFlow (savedmodel.py); falls back to <path>/weights.npz."""
    import tf_bundle_writer as TW
    i, b = 64, 2
    weights = W.random_init("spade", i, seed=11, perturb_affine=True)
    root = str(tmp_path / "model") + os.sep
    TW.write_gaugan_saved_models(root, weights, compress=True)
    x, eps = inputs(i, b, seed=2)
    want = msr.GauGAN(i, b, precision="fp32", weights=weights, eps_fn=lambda call, n: eps)(x, training=False)
    loaded = msr.load_GAN_model(root, i, b, precision="fp32", eps_fn=lambda call, n: eps)
    np.testing.assert_array_equal(loaded(x, training=False), want)
    cnn = msr.load_CNN_model(root, i, b, precision="fp32")
    assert cnn(x, training=False).shape == (b, i, i, 1)
    # npz fallback
    root2 = str(tmp_path / "model2") + os.sep
    os.makedirs(root2)
    W.save_npz(os.path.join(root2, "weights.npz"), weights)
    again = msr.load_GAN_model(root2, i, b, precision="fp32", eps_fn=lambda call, n: eps)
    np.testing.assert_array_equal(again(x, training=False), want)
    os.makedirs(str(tmp_path / "empty"))
    with pytest.raises(ValueError):
        msr.load_GAN_model(str(tmp_path / "empty") + os.sep, i, b)


def test_command_line_run_end_to_end(msr, tmp_path, capsys):
    """The reference's ``python3 process_full_tiles.py --source_folder_path ... --model_path ...`` (:589-594) through this
    package's CLI: GeoTIFF inputs, GauGAN weights from SavedModel directories, preprocess, tiles, three GeoTIFF outputs
    (mean f32, std f32, good UInt16) with the DEM's extent; the same rasters as driving the engine by hand."""
    import tf_bundle_writer as TW
    from moonsuperresolution_b200 import engine, geotiff
    i, b = 64, 4
    weights = W.random_init("spade", i, seed=12, perturb_affine=True)
    root = str(tmp_path / "weights") + os.sep
    TW.write_gaugan_saved_models(root, weights)
    rng = np.random.default_rng(6)
    h = w_ = 256                                             # square, like every raster the reference can preprocess
    dem = (np.cumsum(np.cumsum(rng.standard_normal((h, w_)), 0), 1) * 0.5 + 1500.0).astype(np.float32)
    img = rng.uniform(1, 255, (h, w_)).astype(np.float32)
    src = tmp_path / "in"
    src.mkdir()
    geotiff.write(str(src / "run-DEM.tif"), dem)
    geotiff.write(str(src / "run-DRG.tif"), img)
    out = tmp_path / "out"
    out.mkdir()
    argv = ["--source_folder_path", str(src), "--map_name", "crater", "--save_path", str(out), "--model_path", root,
            "--image_size", str(i), "--stride", "32", "--batch_size", str(b), "--tile_size", "128", "--seed", "5"]
    engine.main(argv)
    assert "Cutting the image in" in capsys.readouterr().out
    mean, _ = geotiff.read(str(out / "crater_mean.tiff"))
    std, _ = geotiff.read(str(out / "crater_std.tiff"))
    good, _ = geotiff.read(str(out / "crater_good.tiff"))
    assert mean.shape == std.shape == good.shape == (h, w_)
    assert mean.dtype == np.float32 and std.dtype == np.float32 and good.dtype == np.uint16
    assert good.any() and np.isfinite(mean[good > 0]).all() and (mean[good == 0] == -32768.0).all()
    # by hand: same config, same seed
    cfg = engine.parse_args(argv)
    eng = msr.DEMSuperResolution(cfg, model=msr.load_GAN_model(root, i, b, max_groups=8))
    eng.setRasters(dem, img)
    eng.preprocess()
    eng.padInputs()
    eng.processTiles()
    m2, s2, g2, _ = eng.results()
    np.testing.assert_array_equal(mean, m2)
    np.testing.assert_array_equal(std, s2)
    np.testing.assert_array_equal(good, g2.astype(np.uint16))


@pytest.mark.parametrize("kind", ["bias_f32", "act_bf16_t", "act_bf16", "spade", "stats"])
def test_tensor_core_epilogues_write_only_their_output(torch, kind):
    """Bounds check of our own (compute-sanitizer is closed on the GPU pool, profiles/r02_compute_sanitizer_closed.txt):
    every output of the tcgen05 kernel's epilogues sits between two guard regions filled with a canary; ragged shapes
    (3 images of 8 x 8: the last 128-row tile is half empty; 5 images of 16 x 16) exercise the row masks.  The guards
    must come back untouched and the payload fully written."""
    from moonsuperresolution_b200 import _lib
    L, st = _lib.lib(), _lib.stream_ptr()
    guard = 1 << 16
    rng = np.random.default_rng(7)

    def guarded(count, dtype):
        buf = torch.full((count + 2 * guard,), -7777.0, dtype=dtype, device="cuda")
        return buf, buf[guard:guard + count]

    def check(buf, count, what):
        assert bool((buf[:guard] == -7777.0).all()) and bool((buf[guard + count:] == -7777.0).all()), what + ": guard overwritten"
        assert not bool((buf[guard:guard + count] == -7777.0).any()), what + ": output not fully written"

    for (n, r) in ((3, 8), (5, 16), (1, 128)):
        if kind in ("bias_f32", "stats"):
            cin, cout = 64, 128
            x = torch.from_numpy(rng.standard_normal((n, r, r, cin)).astype(np.float32)).cuda().to(torch.bfloat16)
            w = torch.from_numpy((rng.standard_normal((cout, 9 * cin)) * 0.05).astype(np.float32)).cuda().to(torch.bfloat16)
            b = torch.zeros(cout, device="cuda") + 0.5
            ybuf, y = guarded(n * r * r * cout, torch.float32)
            pairs_n = n * r * r // 128 * 4 * cout * 2
            use_stats = kind == "stats" and r * r >= 128
            pbuf, pairs = guarded(max(pairs_n, 1), torch.float32)
            _lib.check(L.msr_op_conv_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), None, n, r, cin, cout, 9, 1,
                                        1, 0, 0.2, pairs.data_ptr() if use_stats else None, st), kind)
            torch.cuda.synchronize()
            check(ybuf, n * r * r * cout, kind)
            if use_stats:
                check(pbuf, pairs_n, kind + " pairs")
        elif kind in ("act_bf16_t", "act_bf16"):
            cout = 128 if kind == "act_bf16_t" else 64          # 128 columns: whole-row stores through shared memory
            x = torch.from_numpy(rng.standard_normal((n, r, r, 64)).astype(np.float32)).cuda().to(torch.bfloat16)
            w = torch.from_numpy((rng.standard_normal((cout, 64)) * 0.1).astype(np.float32)).cuda().to(torch.bfloat16)
            b = torch.ones(cout, device="cuda")
            ybuf, y = guarded(n * r * r * cout, torch.bfloat16)
            _lib.check(L.msr_op_conv_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), None, y.data_ptr(), n, r, 64, cout, 1, 1,
                                        0, 2, 0.2, None, st), kind)
            torch.cuda.synchronize()
            check(ybuf, n * r * r * cout, kind)
            want = torch.nn.functional.leaky_relu(x.float().reshape(-1, 64) @ w.float().T + 1.0, 0.2)
            assert (y.float().reshape(-1, cout) - want).abs().max().item() < 0.05
        else:
            C = 64
            a = torch.from_numpy(np.maximum(rng.standard_normal((n, r, r, 128)), 0).astype(np.float32)).cuda().to(torch.bfloat16)
            w = torch.from_numpy((rng.standard_normal((2 * C, 1152)) * 0.03).astype(np.float32)).cuda().to(torch.bfloat16)
            b = torch.zeros(2 * C, device="cuda") + 0.25
            xs = torch.randn((n, r, r, C), device="cuda")
            mean, rstd = torch.zeros((n, C), device="cuda"), torch.ones((n, C), device="cuda")
            ybuf, y = guarded(n * r * r * C, torch.bfloat16)
            _lib.check(L.msr_op_spade_tc(a.data_ptr(), w.data_ptr(), b.data_ptr(), xs.data_ptr(), 0, mean.data_ptr(),
                                         rstd.data_ptr(), 1, y.data_ptr(), n, r, C, st), kind)
            torch.cuda.synchronize()
            check(ybuf, n * r * r * C, kind)


def test_pix2pix_at_the_cfg2_call_shape_bf16(msr, torch):
    """BASELINE.json configs[1] at its call shape: pix2pix-256, bf16 tensor-core mode, batch 16, eight groups per call
    (128 patches); pix2pix has no batch coupling at inference, so every group must equal the oracle within the bf16 bar
    and a zero padding slot must not disturb its batch-mates."""
    b, groups = 16, 8
    w = W.random_init("pix2pix", 256, seed=6, perturb_affine=True)
    rng = np.random.default_rng(12)
    x = rng.uniform(-0.5, 0.5, (groups * b, 256, 256, 2)).astype(np.float32)
    x[3 * b + 5:4 * b] = 0.0
    model = msr.Pix2Pix(batch_size=b, weights=w, precision="bf16", max_groups=groups)
    src = torch.from_numpy(x).cuda()
    out = torch.empty((groups * b, 256, 256), dtype=torch.float32, device="cuda")
    model.forward_device(src, out, None, groups)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    for g in (0, 3, 7):
        sl = slice(g * b, (g + 1) * b)
        want = OG.pix2pix_call(x[sl], w)[..., 0]
        err = np.abs(got[sl] - want).max()
        assert err <= TOL_BF16, (g, err)
