"""Host-side halves of the round-2b tensor-core layers (weight packing), checked against numpy.  No GPU: the C-ABI library
loads and these entry points run on the CPU (csrc/mask_tc.cu, csrc/phase_tc.cu)."""
import ctypes as C

import numpy as np
import pytest

from moonsuperresolution_b200 import _lib


def bf16_bits(a):
    """float32 -> bf16 bit pattern, round to nearest even (finite inputs)."""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)


def bf16_value(bits):
    return (bits.astype(np.uint32) << 16).view(np.float32)


@pytest.mark.parametrize("cout,with_bias", [(128, True), (64, False)])
def test_mask_weight_layout(cout, with_bias):
    """spade.py:18 / blocks.py:53-60 kernels -> the split-bf16 K layout the in-kernel operand builder pairs with
    (x_hi0, x_hi1, x_lo0, x_lo1 per tap, then x_hi0, x_hi1 per tap, then two ones for the bias)."""
    rng = np.random.default_rng(cout)
    w = rng.standard_normal((3, 3, 2, cout)).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32) if with_bias else None
    out = np.full((cout, 64), 0xFFFF, np.uint16)
    _lib.check(_lib.lib().msr_host_pack_mask_weights(w.ctypes.data, None if b is None else b.ctypes.data, cout,
                                                     out.ctypes.data), "msr_host_pack_mask_weights")
    hi = bf16_bits(w)
    lo = bf16_bits(w - bf16_value(hi))
    want = np.zeros((cout, 64), np.uint16)
    for t in range(9):
        for c in range(2):
            want[:, 4 * t + c] = hi[t // 3, t % 3, c]
            want[:, 4 * t + 2 + c] = hi[t // 3, t % 3, c]
            want[:, 36 + 2 * t + c] = lo[t // 3, t % 3, c]
    if with_bias:
        want[:, 54] = bf16_bits(b)
        want[:, 55] = bf16_bits(b - bf16_value(bf16_bits(b)))
    assert np.array_equal(out, want)
    # what the GEMM then evaluates for one pixel: x_hi*w_hi + x_lo*w_hi + x_hi*w_lo (+ bias) ~ the float32 product
    x = rng.uniform(-0.5, 0.5, (9, 2)).astype(np.float32)
    xh = bf16_value(bf16_bits(x))
    xl = bf16_value(bf16_bits(x - xh))
    row = np.zeros(64, np.float64)
    for t in range(9):
        row[4 * t:4 * t + 4] = [xh[t, 0], xh[t, 1], xl[t, 0], xl[t, 1]]
        row[36 + 2 * t:38 + 2 * t] = xh[t]
    row[54:56] = 1.0
    got = bf16_value(out).astype(np.float64) @ row
    ref = np.einsum("tc,tco->o", x.astype(np.float64), w.reshape(9, 2, cout).astype(np.float64))
    if with_bias:
        ref = ref + b
    assert np.abs(got - ref).max() < 2e-4 * max(1.0, np.abs(ref).max())


def phase_filters(kernel, transposed):
    cin = kernel.shape[-1]
    w4 = np.zeros((4, 3, 3, cin), np.float32)
    for py in range(2):
        for px in range(2):
            if not transposed:      # networks.py:54-56: UpSampling2D(2) -> Conv2D(1, 4, 'same'), SAME pads (1, 2)
                for ky in range(4):
                    for kx in range(4):
                        w4[py * 2 + px, (py - 1 + ky) // 2 + 1, (px - 1 + kx) // 2 + 1] += kernel[ky, kx]
            else:                   # pix2pix.py:91-95: Conv2DTranspose(1, 4, strides=2, 'same')
                for ty in range(3):
                    for tx in range(3):
                        ky, kx = py + 1 - 2 * (ty - 1), px + 1 - 2 * (tx - 1)
                        if 0 <= ky <= 3 and 0 <= kx <= 3:
                            w4[py * 2 + px, ty, tx] = kernel[ky, kx]
    return w4


@pytest.mark.parametrize("transposed,cin,ncols", [(False, 128, 25), (True, 128, 16), (False, 64, 25)])
def test_phase_weight_columns(transposed, cin, ncols):
    """The (phase, tap) pairs with a non-zero filter become the columns of the per-pixel GEMM, phase-major then tap order;
    summing column j at pixel offset (ty - 1, tx - 1) into phase q reproduces the 3x3 phase convolution."""
    rng = np.random.default_rng(cin + transposed)
    w4 = phase_filters(rng.standard_normal((4, 4, cin)).astype(np.float32), transposed)
    bits = bf16_bits(w4.reshape(4, 9 * cin))
    wg = np.full((32, cin), 0xFFFF, np.uint16)
    kind, n = C.c_int(-7), C.c_int(-7)
    _lib.check(_lib.lib().msr_host_pack_phase_weights(bits.ctypes.data, cin, wg.ctypes.data, C.byref(kind), C.byref(n)),
               "msr_host_pack_phase_weights")
    assert kind.value == (1 if transposed else 0) and n.value == ncols
    pairs = [(q, t) for q in range(4) for t in range(9) if np.any(w4[q, t // 3, t % 3] != 0)]
    assert len(pairs) == ncols
    for j, (q, t) in enumerate(pairs):
        assert np.array_equal(wg[j], bits[q, t * cin:(t + 1) * cin])
    assert not wg[ncols:].any()
    # end to end on a small tensor: G = x . wg^T per pixel, then the stencil of scalars == the 3x3 phase convolution
    x = bf16_value(bf16_bits(rng.standard_normal((6, 7, cin)).astype(np.float32))).astype(np.float64)
    wf = bf16_value(bits).reshape(4, 3, 3, cin).astype(np.float64)
    G = x @ bf16_value(wg[:ncols]).astype(np.float64).T                       # (6, 7, ncols)
    xp = np.pad(x, ((1, 1), (1, 1), (0, 0)))
    Gp = np.pad(G, ((1, 1), (1, 1), (0, 0)))
    for q in range(4):
        want = sum(np.einsum("hwc,c->hw", xp[ty:ty + 6, tx:tx + 7], wf[q, ty, tx]) for ty in range(3) for tx in range(3))
        got = sum(Gp[t // 3:t // 3 + 6, t % 3:t % 3 + 7, j] for j, (qq, t) in enumerate(pairs) if qq == q)
        assert np.abs(got - want).max() < 1e-9


def test_phase_pattern_outside_the_two_kinds_is_declined():
    """A filter bank whose non-zero (phase, tap) pattern is neither layer kind keeps the 9-tap implicit GEMM."""
    cin = 64
    w4 = np.ones((4, 9 * cin), np.float32)          # all 36 pairs non-zero
    bits = bf16_bits(w4)
    wg = np.zeros((32, cin), np.uint16)
    kind, n = C.c_int(0), C.c_int(0)
    _lib.check(_lib.lib().msr_host_pack_phase_weights(bits.ctypes.data, cin, wg.ctypes.data, C.byref(kind), C.byref(n)),
               "msr_host_pack_phase_weights")
    assert kind.value == -1 and n.value == 0 and not wg.any()
